// CSR contraction kernels: the drop-in for the reference extension THB_eval
//   forward : out[i] = sum_j PHI[j] * ctrl_pts[Jm[j]]          (compute_pts_cuda_kernels.cu:3-26)
//   backward: grad_ctrl_pts[Jm[j]] += PHI[j] * grad_out[i]     (compute_pts_cuda_kernels.cu:28-47)
//
// Reference mapping: one thread per point walking its segment serially (lane stride = S elements,
// uncoalesced; 3*nnz global float atomics).  Here: a warp owns a point's segment, so PHI / Jm are read
// coalesced; the backward keeps per-lane partial sums while consecutive points share one support list
// (points of the same cell do) and only issues atomics when the list changes -- "warp-aggregated"
// scatter-add: 3*S atomics per run of points instead of per point.
#include <cstdlib>
#include <cstdint>

#include "thb_common.cuh"
#include "thb_csr_impl.cuh"

namespace {

using namespace thb_csr;

constexpr int kWarpsPerCta = 8;

template <typename JT, typename CT, int CPD>
__global__ void __launch_bounds__(kWarpsPerCta * 32) k_csr_forward(const float* __restrict__ cp, const JT* __restrict__ jm,
                                                                     const float* __restrict__ phi,
                                                                     const CT* __restrict__ cumsum, float* __restrict__ out,
                                                                     int64_t npoints) {
    constexpr int NPT = 4;
    static_assert(kFwdBatch % NPT == 0 && kFwdPoints % kFwdBatch == 0 && kFwdPoints <= 32, "point grouping");
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    const int64_t p0 = warp * kFwdPoints;
    if (p0 >= npoints) return;
    const int np = (int)min((int64_t)kFwdPoints, npoints - p0);
    const int64_t base = p0 ? (int64_t)cumsum[p0 - 1] : 0;
    // segment ends of the warp's points as offsets from `base`: lane l holds the END of point p0 + l.  A warp whose
    // points hold more than 2^31 - 1 pairs together cannot occur with lists that fit the register path; longer lists take
    // the streamed path below, which works on 64-bit positions.
    const int64_t end64 = lane < np ? (int64_t)cumsum[p0 + lane] : 0;
    const int64_t span = __shfl_sync(0xffffffffu, end64, np - 1) - base;
    const bool small = span < (int64_t)0x7fffffff;
    const int my_end = small ? (int)(end64 - base) : 0;

    FwdCache<JT, CPD> c;
    c.runS = -1;
#pragma unroll
    for (int s = 0; s < kFwdStrips; ++s) c.key[s] = (JT)-2;

    int beg = 0;
    for (int i0 = 0; i0 < np; i0 += kFwdBatch) {
        float acc[CPD][kFwdBatch];
#pragma unroll
        for (int j0 = 0; j0 < kFwdBatch; j0 += NPT) {
            float a[NPT][CPD];
#pragma unroll
            for (int h = 0; h < NPT; ++h)
#pragma unroll
                for (int k = 0; k < CPD; ++k) a[h][k] = 0.0f;
            if (i0 + j0 < np) {   // warp-uniform
#ifndef THB_FWD_NO_PREFETCH
                {   // the lines of the group two ahead -> L2 (one prefetch per lane and array): the loads of that group then
                    // see L2 latency instead of HBM latency; bytes in flight per warp no longer limited by its registers
                    const int gi = i0 + j0 + THB_FWD_PF_DIST * NPT;
                    if (small && gi < np) {
                        const int e0 = __shfl_sync(0xffffffffu, my_end, (gi - 1) & 31);
                        const int e1 = __shfl_sync(0xffffffffu, my_end, (min(gi + NPT, np) - 1) & 31);
                        const char* pa = reinterpret_cast<const char*>(phi + base + e0);
                        const char* ja = reinterpret_cast<const char*>(jm + base + e0);
                        const int nb_p = (e1 - e0) * 4, nb_j = (e1 - e0) * (int)sizeof(JT);
                        if (lane * 128 < nb_p) prefetch_l2(pa + lane * 128);
                        if (lane * 128 < nb_j) prefetch_l2(ja + lane * 128);
                    }
                }
#endif
                int ob[NPT], S[NPT], smax = 0;
#pragma unroll
                for (int h = 0; h < NPT; ++h) {
                    const int i = i0 + j0 + h;
                    const int e = (small && i < np) ? __shfl_sync(0xffffffffu, my_end, i & 31) : beg;
                    ob[h] = beg;
                    S[h] = e - beg;
                    beg = e;
                    smax = max(smax, S[h]);
                }
                if (small && smax <= 64) fwd_points<JT, CPD, 2, NPT>(cp, jm, phi, base, ob, S, c, a, lane);
                else if (small && smax <= 96) fwd_points<JT, CPD, 3, NPT>(cp, jm, phi, base, ob, S, c, a, lane);
                else if (small && smax <= 32 * kFwdStrips) fwd_points<JT, CPD, kFwdStrips, NPT>(cp, jm, phi, base, ob, S, c, a, lane);
                else {  // very long lists (or a warp spanning > 2^31 pairs): streamed on 64-bit positions, nothing cached
#pragma unroll
                    for (int h = 0; h < NPT; ++h) {
                        const int i = i0 + j0 + h;
                        if (i >= np) break;
                        const int64_t pb = i ? __shfl_sync(0xffffffffu, end64, (i - 1) & 31) : base;
                        const int64_t pe = __shfl_sync(0xffffffffu, end64, i & 31);
                        for (int64_t j = pb + lane; j < pe; j += 32) {
                            const float w = __ldg(phi + j);
                            const int64_t r = (int64_t)__ldg(jm + j);
#pragma unroll
                            for (int k = 0; k < CPD; ++k) a[h][k] = fmaf(w, __ldg(cp + r * CPD + k), a[h][k]);
                        }
                    }
                    c.runS = -1;
                }
            }
#pragma unroll
            for (int h = 0; h < NPT; ++h)
#pragma unroll
                for (int k = 0; k < CPD; ++k) acc[k][j0 + h] = a[h][k];
        }
        // ---- eight points reduced together; lane l with (l & 3) == 0 holds point ((l>>4)&1)*4 + ((l>>3)&1)*2 + ((l>>2)&1)
        const int mine = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
#pragma unroll
        for (int k = 0; k < CPD; ++k) {
            const float r = reduce8(acc[k], lane);
            if ((lane & 3) == 0 && i0 + mine < np) out[(p0 + i0 + mine) * CPD + k] = r;
        }
    }
}

template <typename JT, typename CT, int CPD>
__global__ void __launch_bounds__(kWarpsPerCta * 32) k_csr_backward(const float* __restrict__ gout, const JT* __restrict__ jm,
                                                                      const float* __restrict__ phi,
                                                                      const CT* __restrict__ cumsum, float* __restrict__ gcp,
                                                                      int64_t npoints, int pts_per_warp, int64_t perm_stride) {
    // points in flight per warp: four with 32-bit indices; two with 64-bit ones, whose bodies are twice as large -- the
    // kernel is sensitive to its code size (instruction-cache misses): 5.2 vs 4.0 ms int64, 3.0 vs 3.7 ms int32 on config 3
#ifdef THB_BWD_NPT
    constexpr int NPT = THB_BWD_NPT;
#else
    constexpr int NPT = sizeof(JT) == 8 ? 2 : 4;
#endif
    const int lane = threadIdx.x & 31;
    int64_t warp = (int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    {   // Warps that run at the same time take point blocks far apart (block index times a stride coprime to the block
        // count): neighbouring points add into the same control points, and atomics on one address serialise in L2 --
        // with the blocks dealt in order the config-3 backward took 5.5 ms, spread out 3.0 ms (int32 indices).
        const int64_t nw = (npoints + pts_per_warp - 1) / pts_per_warp;
        if (warp >= nw) return;
        warp = (warp * perm_stride) % nw;
    }
    const int64_t p0 = warp * pts_per_warp;
    if (p0 >= npoints) return;
    const int64_t p1 = (p0 + pts_per_warp < npoints) ? p0 + pts_per_warp : npoints;

    BwdCache<JT, CPD> c;
#pragma unroll
    for (int s = 0; s < kStrips; ++s) {
        c.key[s] = (JT)-1;
#pragma unroll
        for (int k = 0; k < CPD; ++k) c.acc[s][k] = 0.0f;
    }
    c.runS = -1;

    int64_t base = p0 ? (int64_t)cumsum[p0 - 1] : 0;
    for (int64_t c0 = p0; c0 < p1; c0 += 32) {
        // ends of up to 32 points at once (one coalesced load) as 32-bit offsets from the first entry of the group
        const int nc = (int)min((int64_t)32, p1 - c0);
        const int64_t end64 = lane < nc ? (int64_t)cumsum[c0 + lane] : 0;
        const int64_t last = __shfl_sync(0xffffffffu, end64, nc - 1);
        const bool small = last - base < (int64_t)0x7fffffff;
        const int my_end = small ? (int)(end64 - base) : 0;
        int beg = 0;
        for (int i = 0; i < nc; i += NPT) {
#ifndef THB_BWD_PREFETCH
#define THB_BWD_PREFETCH 2   // as in the forward kernel (3.96 -> 3.75 ms int64, flat with int32 indices: that one waits on L2 atomics)
#endif
            {
                const int gi = i + THB_BWD_PREFETCH * NPT;
                if (small && gi < nc) {
                    const int e0 = __shfl_sync(0xffffffffu, my_end, (gi - 1) & 31);
                    const int e1 = __shfl_sync(0xffffffffu, my_end, (min(gi + NPT, nc) - 1) & 31);
                    const char* pa = reinterpret_cast<const char*>(phi + base + e0);
                    const char* ja = reinterpret_cast<const char*>(jm + base + e0);
                    const int nb_p = (e1 - e0) * 4, nb_j = (e1 - e0) * (int)sizeof(JT);
                    if (lane * 128 < nb_p) prefetch_l2(pa + lane * 128);
                    if (lane * 128 < nb_j) prefetch_l2(ja + lane * 128);
                }
            }
            int ob[NPT], S[NPT], smax = 0;
#pragma unroll
            for (int h = 0; h < NPT; ++h) {
                const int e = (small && i + h < nc) ? __shfl_sync(0xffffffffu, my_end, (i + h) & 31) : beg;
                ob[h] = beg;
                S[h] = e - beg;
                beg = e;
                smax = max(smax, S[h]);
            }
            const int nvalid = min(NPT, nc - i);
#ifndef THB_BWD_SPEC2
            if (small && smax <= 64) bwd_points<JT, CPD, 2, NPT>(gout, jm, phi, base, c0 + i, nvalid, ob, S, c, gcp, lane);
            else
#endif
            if (small && smax <= 96) bwd_points<JT, CPD, 3, NPT>(gout, jm, phi, base, c0 + i, nvalid, ob, S, c, gcp, lane);
            else if (small && smax <= 32 * kStrips) bwd_points<JT, CPD, kStrips, NPT>(gout, jm, phi, base, c0 + i, nvalid, ob, S, c, gcp, lane);
            else {  // very long lists: no aggregation, 64-bit positions
                bwd_flush<JT, CPD>(c, gcp);
                for (int h = 0; h < nvalid; ++h) {
                    const int64_t pb = (i + h) ? __shfl_sync(0xffffffffu, end64, (i + h - 1) & 31) : base;
                    const int64_t pe = __shfl_sync(0xffffffffu, end64, (i + h) & 31);
                    float g[CPD];
#pragma unroll
                    for (int k = 0; k < CPD; ++k) g[k] = __ldg(gout + (c0 + i + h) * CPD + k);
                    for (int64_t j = pb + lane; j < pe; j += 32) {
                        const float w = __ldg(phi + j);
                        const int64_t r = (int64_t)__ldg(jm + j);
#pragma unroll
                        for (int k = 0; k < CPD; ++k) atomicAdd(gcp + r * CPD + k, w * g[k]);
                    }
                }
            }
        }
        base = last;
    }
    bwd_flush<JT, CPD>(c, gcp);
}

template <typename JT, typename CT>
int launch_fwd(const float* cp, const void* jm, const float* phi, const void* cs, float* out, int64_t n, int cpd,
               cudaStream_t st) {
    const int64_t warps = (n + kFwdPoints - 1) / kFwdPoints;
    const int grid = (int)((warps + kWarpsPerCta - 1) / kWarpsPerCta);
#define THB_FWD(C)                                                                                             \
    k_csr_forward<JT, CT, C><<<grid, kWarpsPerCta * 32, 0, st>>>(cp, (const JT*)jm, phi, (const CT*)cs, out, n)
    switch (cpd) {
        case 1: THB_FWD(1); break;
        case 2: THB_FWD(2); break;
        case 3: THB_FWD(3); break;
        case 4: THB_FWD(4); break;
        default: return thb_set_error(THB_E_INVALID, "cp_dim must be 1..4");
    }
#undef THB_FWD
    THB_LAUNCHED();
    return THB_OK;
}

template <typename JT, typename CT>
int launch_bwd(const float* go, const void* jm, const float* phi, const void* cs, float* gcp, int64_t n, int cpd,
               cudaStream_t st) {
    // enough warps to fill the machine, but long runs (up to 64 points) so that same-cell points aggregate
    int64_t ppw = n / ((int64_t)thb_num_sms() * 64);
    if (ppw < 4) ppw = 4;
    if (ppw > 64) ppw = 64;
    const int64_t warps = (n + ppw - 1) / ppw;
    const int grid = (int)((warps + kWarpsPerCta - 1) / kWarpsPerCta);
    // stride of the block permutation: about 0.618 of the block count, made coprime to it (THB_BWD_PERM=0: blocks in
    // order; any other value: that stride)
    int64_t stride = 1;
    {
        const char* e = getenv("THB_BWD_PERM");
        const long long want = e ? atoll(e) : -1;
        if (want != 0 && warps > 8) {
            stride = want > 0 ? (int64_t)want : ((int64_t)((double)warps * 0.6180339887) | 1);
            auto gcd = [](int64_t a, int64_t b) { while (b) { const int64_t t = a % b; a = b; b = t; } return a; };
            while (gcd(stride, warps) != 1) stride += 1;
        }
    }
#define THB_BWD(C)                                                                                              \
    k_csr_backward<JT, CT, C><<<grid, kWarpsPerCta * 32, 0, st>>>(go, (const JT*)jm, phi, (const CT*)cs, gcp, n, \
                                                                  (int)ppw, stride)
    switch (cpd) {
        case 1: THB_BWD(1); break;
        case 2: THB_BWD(2); break;
        case 3: THB_BWD(3); break;
        case 4: THB_BWD(4); break;
        default: return thb_set_error(THB_E_INVALID, "cp_dim must be 1..4");
    }
#undef THB_BWD
    THB_LAUNCHED();
    return THB_OK;
}

}  // namespace

extern "C" int thb_eval_forward(const float* ctrl_pts, const void* jm, int jm_is_i64, const float* phi, const void* cumsum,
                                int cumsum_is_i64, float* out, int64_t npoints, int cp_dim, void* stream) {
    THB_REQUIRE(npoints >= 0, "thb_eval_forward: negative point count");
    if (npoints == 0) return THB_OK;
    THB_REQUIRE(ctrl_pts && jm && phi && cumsum && out, "thb_eval_forward: null pointer");
    THB_REQUIRE(cp_dim >= 1 && cp_dim <= 4, "cp_dim must be 1..4");
    cudaStream_t st = (cudaStream_t)stream;
    {   // bulk-staged kernel (thb_eval_bulk.cu) whenever the arrays allow 16-byte copies
        const int rc = thb_csr_forward_bulk(ctrl_pts, jm, jm_is_i64, phi, cumsum, cumsum_is_i64, out, npoints, cp_dim, st);
        if (rc != 1) return rc;
    }
    if (jm_is_i64 && cumsum_is_i64) return launch_fwd<long long, long long>(ctrl_pts, jm, phi, cumsum, out, npoints, cp_dim, st);
    if (jm_is_i64) return launch_fwd<long long, int>(ctrl_pts, jm, phi, cumsum, out, npoints, cp_dim, st);
    if (cumsum_is_i64) return launch_fwd<int, long long>(ctrl_pts, jm, phi, cumsum, out, npoints, cp_dim, st);
    return launch_fwd<int, int>(ctrl_pts, jm, phi, cumsum, out, npoints, cp_dim, st);
}

extern "C" int thb_eval_backward(const float* grad_out, const void* jm, int jm_is_i64, const float* phi, const void* cumsum,
                                 int cumsum_is_i64, float* grad_ctrl_pts, int64_t npoints, int64_t nctrl, int cp_dim,
                                 void* stream) {
    THB_REQUIRE(npoints >= 0 && nctrl >= 0, "thb_eval_backward: negative size");
    THB_REQUIRE(cp_dim >= 1 && cp_dim <= 4, "cp_dim must be 1..4");
    cudaStream_t st = (cudaStream_t)stream;
    if (nctrl > 0) {
        THB_REQUIRE(grad_ctrl_pts, "thb_eval_backward: null gradient buffer");
        THB_CUDA(cudaMemsetAsync(grad_ctrl_pts, 0, sizeof(float) * nctrl * cp_dim, st));
    }
    if (npoints == 0) return THB_OK;
    THB_REQUIRE(grad_out && jm && phi && cumsum, "thb_eval_backward: null pointer");
    {   // bulk-staged kernel (thb_eval_bulk.cu) whenever the arrays allow 16-byte copies
        const int rc = thb_csr_backward_bulk(grad_out, jm, jm_is_i64, phi, cumsum, cumsum_is_i64, grad_ctrl_pts, npoints, cp_dim, st);
        if (rc != 1) return rc;
    }
    if (jm_is_i64 && cumsum_is_i64) return launch_bwd<long long, long long>(grad_out, jm, phi, cumsum, grad_ctrl_pts, npoints, cp_dim, st);
    if (jm_is_i64) return launch_bwd<long long, int>(grad_out, jm, phi, cumsum, grad_ctrl_pts, npoints, cp_dim, st);
    if (cumsum_is_i64) return launch_bwd<int, long long>(grad_out, jm, phi, cumsum, grad_ctrl_pts, npoints, cp_dim, st);
    return launch_bwd<int, int>(grad_out, jm, phi, cumsum, grad_ctrl_pts, npoints, cp_dim, st);
}
