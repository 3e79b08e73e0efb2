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

namespace {

constexpr int kWarpsPerCta = 8;

constexpr int kFwdStrips = 5;   // support lists up to 160 entries live in registers (one entry per lane and strip)
#ifndef THB_FWD_POINTS
#define THB_FWD_POINTS 16
#endif
constexpr int kFwdPoints = THB_FWD_POINTS;  // consecutive points per warp

// A warp owns kFwdPoints consecutive points.  Per point the lanes read the segment coalesced (one entry per lane and
// strip).  The gathered control points stay in registers for as long as consecutive points have the same support
// list (points of one cell do), so the dependent second round trip of a point -- the gather through Jm -- disappears
// inside a run.  (A variant that streams the warp's contiguous PHI / Jm run through a shared-memory ring with
// 16-byte cp.async was measured at 4.4 ms int64 / 3.2 ms int32 against 3.7 / 3.7 ms for this one on config 3: its
// per-point bookkeeping executes 3.2 G warp-instructions, so it is issue-bound before it is bandwidth-bound.)
template <typename JT, typename CT, int CPD>
__global__ void __launch_bounds__(kWarpsPerCta * 32) k_csr_forward(const float* __restrict__ cp, const JT* __restrict__ jm,
                                                                     const float* __restrict__ phi,
                                                                     const CT* __restrict__ cumsum, float* __restrict__ out,
                                                                     int64_t npoints) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    const int64_t p0 = warp * kFwdPoints;
    if (p0 >= npoints) return;
    const int np = (int)min((int64_t)kFwdPoints, npoints - p0);
    // segment bounds of the warp's points: lane l holds the END of point p0 + l
    const int64_t my_end = lane < np ? (int64_t)cumsum[p0 + lane] : 0;
    int64_t beg = p0 ? (int64_t)cumsum[p0 - 1] : 0;

    JT key[kFwdStrips];
    float cpv[kFwdStrips][CPD];
    int runS = -1;
#pragma unroll
    for (int s = 0; s < kFwdStrips; ++s) key[s] = (JT)-2;

    for (int i = 0; i < np; ++i) {
        const int64_t end = __shfl_sync(0xffffffffu, my_end, i);
        const int64_t S64 = end - beg;
        float acc[CPD];
#pragma unroll
        for (int k = 0; k < CPD; ++k) acc[k] = 0.0f;
        if (S64 <= 32 * kFwdStrips) {
            const int S = (int)S64;
            JT cj[kFwdStrips];
            float cw[kFwdStrips];
            bool same = S == runS;
#pragma unroll
            for (int s = 0; s < kFwdStrips; ++s) {
                const int o = s * 32 + lane;
                const bool ok = o < S;
                cj[s] = ok ? __ldg(jm + beg + o) : (JT)-1;
                cw[s] = ok ? __ldg(phi + beg + o) : 0.0f;
            }
#pragma unroll
            for (int s = 0; s < kFwdStrips; ++s) same = same && (cj[s] == key[s]);
            if (!__all_sync(0xffffffffu, same)) {
#pragma unroll
                for (int s = 0; s < kFwdStrips; ++s) {
                    key[s] = cj[s];
#pragma unroll
                    for (int k = 0; k < CPD; ++k) cpv[s][k] = cj[s] >= 0 ? __ldg(cp + (int64_t)cj[s] * CPD + k) : 0.0f;
                }
                runS = S;
            }
#pragma unroll
            for (int s = 0; s < kFwdStrips; ++s)
#pragma unroll
                for (int k = 0; k < CPD; ++k) acc[k] = fmaf(cw[s], cpv[s][k], acc[k]);
        } else {  // very long list: streamed, nothing cached
            for (int64_t j = beg + lane; j < end; j += 32) {
                const float w = __ldg(phi + j);
                const int64_t r = (int64_t)__ldg(jm + j);
#pragma unroll
                for (int k = 0; k < CPD; ++k) acc[k] = fmaf(w, __ldg(cp + r * CPD + k), acc[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < CPD; ++k) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
        }
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < CPD; ++k) out[(p0 + i) * CPD + k] = acc[k];
        }
        beg = end;
    }
}

constexpr int kStrips = 6;  // support lists up to 192 entries are aggregated across points

template <typename JT, typename CT, int CPD>
__global__ void __launch_bounds__(kWarpsPerCta * 32) k_csr_backward(const float* __restrict__ gout, const JT* __restrict__ jm,
                                                                      const float* __restrict__ phi,
                                                                      const CT* __restrict__ cumsum, float* __restrict__ gcp,
                                                                      int64_t npoints, int pts_per_warp) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    const int64_t p0 = warp * pts_per_warp;
    if (p0 >= npoints) return;
    const int64_t p1 = (p0 + pts_per_warp < npoints) ? p0 + pts_per_warp : npoints;

    float acc[kStrips][CPD];
    long long key[kStrips];
#pragma unroll
    for (int s = 0; s < kStrips; ++s) {
        key[s] = -1;
#pragma unroll
        for (int k = 0; k < CPD; ++k) acc[s][k] = 0.0f;
    }
    int runS = -1;

    auto flush = [&]() {
#pragma unroll
        for (int s = 0; s < kStrips; ++s) {
            if (key[s] >= 0) {
#pragma unroll
                for (int k = 0; k < CPD; ++k) atomicAdd(gcp + key[s] * CPD + k, acc[s][k]);
            }
            key[s] = -1;
#pragma unroll
            for (int k = 0; k < CPD; ++k) acc[s][k] = 0.0f;
        }
        runS = -1;
    };

    for (int64_t i = p0; i < p1; ++i) {
        const int64_t beg = i ? (int64_t)cumsum[i - 1] : 0;
        const int64_t end = (int64_t)cumsum[i];
        const int64_t S64 = end - beg;
        float g[CPD];
#pragma unroll
        for (int k = 0; k < CPD; ++k) g[k] = __ldg(gout + i * CPD + k);
        if (S64 > 32 * kStrips) {  // very long list: no aggregation
            flush();
            for (int64_t j = beg + lane; j < end; j += 32) {
                const float w = __ldg(phi + j);
                const int64_t r = (int64_t)__ldg(jm + j);
#pragma unroll
                for (int k = 0; k < CPD; ++k) atomicAdd(gcp + r * CPD + k, w * g[k]);
            }
            continue;
        }
        const int S = (int)S64;
        long long myj[kStrips];
        float myw[kStrips];
        bool same = (S == runS);
#pragma unroll
        for (int s = 0; s < kStrips; ++s) {
            const int o = s * 32 + lane;
            const bool valid = o < S;
            myj[s] = valid ? (long long)__ldg(jm + beg + o) : -1;
            myw[s] = valid ? __ldg(phi + beg + o) : 0.0f;
            same = same && (myj[s] == key[s]);
        }
        same = __all_sync(0xffffffffu, same);
        if (!same) {
            flush();
#pragma unroll
            for (int s = 0; s < kStrips; ++s) key[s] = myj[s];
            runS = S;
        }
#pragma unroll
        for (int s = 0; s < kStrips; ++s) {
#pragma unroll
            for (int k = 0; k < CPD; ++k) acc[s][k] = fmaf(myw[s], g[k], acc[s][k]);
        }
    }
    flush();
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
#define THB_BWD(C)                                                                                              \
    k_csr_backward<JT, CT, C><<<grid, kWarpsPerCta * 32, 0, st>>>(go, (const JT*)jm, phi, (const CT*)cs, gcp, n, \
                                                                  (int)ppw)
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
    cudaStream_t st = (cudaStream_t)stream;
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
    if (jm_is_i64 && cumsum_is_i64) return launch_bwd<long long, long long>(grad_out, jm, phi, cumsum, grad_ctrl_pts, npoints, cp_dim, st);
    if (jm_is_i64) return launch_bwd<long long, int>(grad_out, jm, phi, cumsum, grad_ctrl_pts, npoints, cp_dim, st);
    if (cumsum_is_i64) return launch_bwd<int, long long>(grad_out, jm, phi, cumsum, grad_ctrl_pts, npoints, cp_dim, st);
    return launch_bwd<int, int>(grad_out, jm, phi, cumsum, grad_ctrl_pts, npoints, cp_dim, st);
}
