// THB_eval.forward / backward (compute_pts_cuda_kernels.cu:3-26, 28-47) from BULK-STAGED segments.
//
// The register-path kernel of thb_eval_csr.cu (lane = entry, coalesced loads, control points cached in registers while
// consecutive points share a support list) is limited by the bytes a warp can keep in flight: its loads live in registers
// (four points at a time) and it waits for them at ~50 % issue, 0.6 of HBM on config 3.  Here the memory side is the TMA:
// a warp's next chunk of up to 16 consecutive points is ONE contiguous range of PHI and one of Jm, fetched by two
// cp.async.bulk copies (one lane issues them, completion on an mbarrier) into a two-stage shared-memory ring -- the whole
// next chunk in flight per warp, no registers, no address arithmetic, no prefetch instructions.  The compute side is the
// same point body (thb_csr_impl.cuh) reading shared memory.
//
// Measured and dropped on the way (config 3, int32 indices, register path 2.40 ms): the reference's own loop, one LANE per
// point walking its staged segment (no reductions, no run detection: 9.5 instructions per entry) -- lanes of one cell sit
// S words apart, S = 64 is the common case, so either all lanes hit one bank or, staggered, every gather touches up to 32
// control-point rows: 2.48 ms at best (4-way conflicts, 8 rows).
#include <cstdint>
#include <cstdlib>

#include "thb_common.cuh"
#include "thb_csr_impl.cuh"

namespace {

using namespace thb_csr;

#ifndef THB_BULK_CAP
#define THB_BULK_CAP 768   // entries per stage (~11 points of config 3).  Measured on config 3, forward / backward ms with int32 indices:
                           // 2304 entries, 6 warps per SM: 2.48 / -; 1152, 12 warps: 2.28 / 2.68; 768, 16 warps: 2.10 / 2.56; 640, 20 warps
                           // and two points in flight: 2.17 / 2.60 -- registers (128 at 16 warps) bound the number of warps
#endif
constexpr int kCap = THB_BULK_CAP;
constexpr int kSlack = 8;   // alignment shift (<= 3) + rounding up to 16 bytes (<= 3)
constexpr int kChunkPtsMax = 16;

template <typename JT>
struct BulkStage {
    float phi[kCap + kSlack];
    JT jm[kCap + kSlack];   // (the kernel addresses jm as phi + (kCap + kSlack) * 4 bytes when sizeof(JT) == 4)
};
template <typename JT>
struct BulkWarp {
    BulkStage<JT> st[2];
    unsigned long long bar[2];
};
// backward: the grad_out rows of the chunk travel with it (their global-memory latency would be the one left exposed)
template <typename JT>
struct BulkWarpBwd {
    BulkStage<JT> st[2];
    float g[2][kChunkPtsMax * 4 + kSlack];
    unsigned long long bar[2];
};
#ifndef THB_BULK_WARPS32
#define THB_BULK_WARPS32 16
#endif
#ifndef THB_BULK_WARPS64
#define THB_BULK_WARPS64 12
#endif
template <typename JT>
struct BulkCfg {
    // autonomous warps per CTA (one CTA per SM: 200-225 KB of stages; 128 registers per thread at 16 warps)
    static constexpr int kWarps = sizeof(JT) == 8 ? THB_BULK_WARPS64 : THB_BULK_WARPS32;
};

constexpr int kChunkPts = kChunkPtsMax;   // points per chunk at most (two reduce8 batches)

struct Chunk {
    int64_t p;       // first point
    int64_t base;    // first entry
    int m;           // points (0: none left)
    int npairs;
    int my_beg, my_end;   // this lane's segment relative to base (lane < m)
    bool bulk;
};

// cumsum[p + lane] (the end of the lane's segment) or -1 past the warp's range.  Loaded one chunk AHEAD of its use: the
// chain p -> cumsum -> number of points that fit -> next p is serial, and a load per chunk on the critical path cost a
// third of the kernel (measured)
template <typename CT>
THB_DEV int64_t load_end(const CT* __restrict__ cumsum, int64_t p, int64_t p1, int lane) {
    return p + lane < p1 ? (int64_t)__ldg(cumsum + p + lane) : (int64_t)-1;
}

// up to kChunkPts points from p (< p1) whose entries fit one stage; `end` = load_end(p).  gtot / cpd (backward): the number
// of floats of grad_out, whose rows [p, p + m) are staged too
THB_DEV Chunk describe(int64_t end, int64_t p, int64_t p1, int64_t base, int64_t nnz, int esz_jm, int lane, int64_t gtot = 0, int cpd = 0) {
    Chunk c;
    c.p = p; c.base = base; c.m = 0; c.npairs = 0; c.my_beg = 0; c.my_end = 0; c.bulk = false;
    if (p >= p1) return c;
    const bool in = end >= 0 && lane < kChunkPts;
    const int64_t rel = in ? end - base : 0;
    const bool fits = in && rel >= 0 && rel <= (int64_t)kCap;
    const unsigned mask = __ballot_sync(0xffffffffu, fits);   // cumsum is non-decreasing: a prefix of the lanes ...
    int m = __ffs(~mask) - 1;                                 // ... and if it is not (invalid input), the copies stay bounded
    c.bulk = m > 0;
    if (m == 0) m = 1;   // one over-long segment: walked from global memory by lane 0
    c.m = m;
    c.my_end = (int)min(rel, (int64_t)0x7fffffff);
    const int64_t tot = __shfl_sync(0xffffffffu, rel, m - 1);
    c.npairs = (int)min(tot, (int64_t)0x7fffffff);
    if (lane >= m) c.my_end = c.npairs;   // lanes past the chunk describe empty segments at its end: no range checks when the
                                          // kernels walk the points
    const int up = __shfl_up_sync(0xffffffffu, c.my_end, 1);
    c.my_beg = lane == 0 ? 0 : up;
    if (c.bulk) {   // the copies are rounded out to 16 bytes: they must stay inside the arrays
        const int64_t end4 = (base + tot + 3) & ~(int64_t)3;
        const int per16 = 16 / esz_jm;
        const int64_t endj = (base + tot + per16 - 1) / per16 * per16;
        if (end4 > nnz || endj > nnz || tot == 0) c.bulk = false;   // (nothing to copy: the points have no supports)
        if (gtot > 0 && (((p + m) * cpd + 3) & ~(int64_t)3) > gtot) c.bulk = false;
    }
    return c;
}

template <typename JT>
THB_DEV void issue(const Chunk& c, BulkStage<JT>& st, unsigned long long* bar, const float* __restrict__ phi, const JT* __restrict__ jm,
                   int lane, float* gst = nullptr, const float* __restrict__ gout = nullptr, int cpd = 0) {
    if (lane == 0) {
        constexpr int per16 = 16 / (int)sizeof(JT);
        const int64_t a0 = c.base & ~(int64_t)3, a1 = (c.base + c.npairs + 3) & ~(int64_t)3;
        const int64_t j0 = c.base / per16 * per16, j1 = (c.base + c.npairs + per16 - 1) / per16 * per16;
        const unsigned bp = (unsigned)(a1 - a0) * 4u, bj = (unsigned)(j1 - j0) * (unsigned)sizeof(JT);
        unsigned bg = 0;
        int64_t g0 = 0;
        if (gst) {   // rows [p, p + m) of grad_out, rounded out to 16 bytes (describe() checked that this stays inside)
            g0 = (c.p * cpd) & ~(int64_t)3;
            bg = (unsigned)((((c.p + c.m) * cpd + 3) & ~(int64_t)3) - g0) * 4u;
        }
        mbar_expect_tx(bar, bp + bj + bg);
        bulk_g2s(st.phi, phi + a0, bp, bar);
        bulk_g2s(st.jm, jm + j0, bj, bar);
        if (gst) bulk_g2s(gst, gout + g0, bg, bar);
    }
}

template <typename JT, typename CT, int CPD>
__global__ void __launch_bounds__(32 * BulkCfg<JT>::kWarps, 1) k_csr_forward_bulk(const float* __restrict__ cp, const JT* __restrict__ jm,
                                                                                const float* __restrict__ phi,
                                                                                const CT* __restrict__ cumsum, float* __restrict__ out,
                                                                                int64_t npoints, int pts_per_warp) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    BulkWarp<JT>& sw = reinterpret_cast<BulkWarp<JT>*>(smem_raw)[wid];
    const int64_t warp = (int64_t)blockIdx.x * BulkCfg<JT>::kWarps + wid;
    const int64_t p0 = warp * pts_per_warp;
    if (p0 >= npoints) return;
    const int64_t p1 = min(p0 + pts_per_warp, npoints);
    if (lane == 0) {
        mbar_init(&sw.bar[0], 1);
        mbar_init(&sw.bar[1], 1);
        mbar_init_fence();
    }
    __syncwarp();
    const int64_t nnz = (int64_t)cumsum[npoints - 1];
    constexpr int per16 = 16 / (int)sizeof(JT);
    constexpr int ej = (int)sizeof(JT);
    Chunk cur = describe(load_end<CT>(cumsum, p0, p1, lane), p0, p1, p0 ? (int64_t)cumsum[p0 - 1] : 0, nnz, ej, lane);
    if (cur.bulk) issue<JT>(cur, sw.st[0], &sw.bar[0], phi, jm, lane);
    Chunk nxt = describe(load_end<CT>(cumsum, cur.p + cur.m, p1, lane), cur.p + cur.m, p1, cur.base + cur.npairs, nnz, ej, lane);
    if (nxt.m > 0 && nxt.bulk) issue<JT>(nxt, sw.st[1], &sw.bar[1], phi, jm, lane);
    int64_t p_after = nxt.p + nxt.m;
    int64_t end_after = load_end<CT>(cumsum, p_after, p1, lane);   // consumed after the next compute phase
    int s = 0;
    unsigned parity = 0u;   // bit s: phase of stage s
#ifdef THB_BULK_NPT
    constexpr int NPT = THB_BULK_NPT;
#else
    constexpr int NPT = 4;
#endif
    FwdCache<JT, CPD> c;
    c.runS = -1;
#pragma unroll
    for (int q = 0; q < kFwdStrips; ++q) c.key[q] = (JT)-2;
    while (cur.m > 0) {
        if (cur.bulk) {
            mbar_wait(&sw.bar[s], (parity >> s) & 1u);
            parity ^= 1u << s;
        }
        // staged chunk: pointers into the stage, entry 0 = the chunk's first entry; otherwise (ends of the arrays, over-long
        // segments) the same bodies on global memory
        const float* cphi = cur.bulk ? sw.st[s].phi + (int)(cur.base & 3) : phi + cur.base;
        const JT* cjm = cur.bulk ? sw.st[s].jm + (int)(cur.base % per16) : jm + cur.base;
        const int np = cur.m;
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
                    int ob[NPT], S[NPT], smax = 0;
#pragma unroll
                    for (int h = 0; h < NPT; ++h) {
                        const int i = i0 + j0 + h;
                        const int e = __shfl_sync(0xffffffffu, cur.my_end, i);   // i < 32; lanes >= np hold the chunk's end
                        ob[h] = beg;
                        S[h] = e - beg;
                        beg = e;
                        smax = max(smax, S[h]);
                    }
                    if (cur.bulk && smax <= 64) fwd_points<JT, CPD, 2, NPT, true>(cp, cjm, cphi, 0, ob, S, c, a, lane);
                    else if (cur.bulk && smax <= 96) fwd_points<JT, CPD, 3, NPT, true>(cp, cjm, cphi, 0, ob, S, c, a, lane);
                    else if (cur.bulk && smax <= 32 * kFwdStrips) fwd_points<JT, CPD, kFwdStrips, NPT, true>(cp, cjm, cphi, 0, ob, S, c, a, lane);
                    else {   // long lists, or a chunk that was not staged: streamed from global memory, nothing cached
#pragma unroll
                        for (int h = 0; h < NPT; ++h) {
                            for (int64_t j = cur.base + ob[h] + lane; j < cur.base + ob[h] + S[h]; j += 32) {
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
            const int mine = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
#pragma unroll
            for (int k = 0; k < CPD; ++k) {
                const float r = reduce8(acc[k], lane);
                if ((lane & 3) == 0 && i0 + mine < np) out[(cur.p + i0 + mine) * CPD + k] = r;
            }
        }
        __syncwarp();
        // the stage just read is free: the chunk after next goes there
        Chunk nn = describe(end_after, p_after, p1, nxt.base + nxt.npairs, nnz, ej, lane);
        if (nn.m > 0 && nn.bulk) issue<JT>(nn, sw.st[s], &sw.bar[s], phi, jm, lane);
        p_after = nn.p + nn.m;
        end_after = load_end<CT>(cumsum, p_after, p1, lane);
        cur = nxt;
        nxt = nn;
        s ^= 1;
    }
}

// THB_eval.backward (compute_pts_cuda_kernels.cu:28-47) on the same ring: lane = entry, products accumulated in registers
// while consecutive points share a support list, atomics when the list changes (bwd_points, thb_csr_impl.cuh).  A warp walks
// its range of points in order, so the run cache carries over chunk boundaries; warps that run together work on ranges far
// apart (different control points).
template <typename JT, typename CT, int CPD>
__global__ void __launch_bounds__(32 * BulkCfg<JT>::kWarps, 1) k_csr_backward_bulk(const float* __restrict__ gout, const JT* __restrict__ jm,
                                                                                 const float* __restrict__ phi,
                                                                                 const CT* __restrict__ cumsum, float* __restrict__ gcp,
                                                                                 int64_t npoints, int pts_per_warp, int64_t nranges,
                                                                                 int64_t perm_stride) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    BulkWarpBwd<JT>& sw = reinterpret_cast<BulkWarpBwd<JT>*>(smem_raw)[wid];
    int64_t warp = (int64_t)blockIdx.x * BulkCfg<JT>::kWarps + wid;
    if (warp >= nranges) return;
    // warps that run together take ranges far apart (range index times a stride coprime to the number of ranges):
    // neighbouring ranges add into the same control points, and atomics on one address serialise in L2 (in order: 4.7 ms)
    warp = (warp * perm_stride) % nranges;
    const int64_t p0 = warp * pts_per_warp;
    if (p0 >= npoints) return;
    const int64_t p1 = min(p0 + pts_per_warp, npoints);
    if (lane == 0) {
        mbar_init(&sw.bar[0], 1);
        mbar_init(&sw.bar[1], 1);
        mbar_init_fence();
    }
    __syncwarp();
    const int64_t nnz = (int64_t)cumsum[npoints - 1];
    const int64_t gtot = npoints * CPD;
    constexpr int per16 = 16 / (int)sizeof(JT);
    constexpr int ej = (int)sizeof(JT);
#ifdef THB_BULK_NPT
    constexpr int NPT = sizeof(JT) == 8 ? 2 : THB_BULK_NPT;
#else
    constexpr int NPT = sizeof(JT) == 8 ? 2 : 4;
#endif
    Chunk cur = describe(load_end<CT>(cumsum, p0, p1, lane), p0, p1, p0 ? (int64_t)cumsum[p0 - 1] : 0, nnz, ej, lane, gtot, CPD);
    if (cur.bulk) issue<JT>(cur, sw.st[0], &sw.bar[0], phi, jm, lane, sw.g[0], gout, CPD);
    Chunk nxt = describe(load_end<CT>(cumsum, cur.p + cur.m, p1, lane), cur.p + cur.m, p1, cur.base + cur.npairs, nnz, ej, lane, gtot, CPD);
    if (nxt.m > 0 && nxt.bulk) issue<JT>(nxt, sw.st[1], &sw.bar[1], phi, jm, lane, sw.g[1], gout, CPD);
    int64_t p_after = nxt.p + nxt.m;
    int64_t end_after = load_end<CT>(cumsum, p_after, p1, lane);
    int s = 0;
    unsigned parity = 0u;
    BwdCache<JT, CPD> c;
#pragma unroll
    for (int q = 0; q < kStrips; ++q) {
        c.key[q] = (JT)-1;
#pragma unroll
        for (int k = 0; k < CPD; ++k) c.acc[q][k] = 0.0f;
    }
    c.runS = -1;
    while (cur.m > 0) {
        if (cur.bulk) {
            mbar_wait(&sw.bar[s], (parity >> s) & 1u);
            parity ^= 1u << s;
        }
        const float* cphi = sw.st[s].phi + (int)(cur.base & 3);
        const JT* cjm = sw.st[s].jm + (int)(cur.base % per16);
        const float* cg = sw.g[s] + (int)((cur.p * CPD) & 3);
        const int np = cur.m;
        int beg = 0;
        for (int i = 0; i < np; i += NPT) {
            int ob[NPT], S[NPT], smax = 0;
#pragma unroll
            for (int h = 0; h < NPT; ++h) {
                const int e = __shfl_sync(0xffffffffu, cur.my_end, i + h);   // < 32; lanes >= np hold the chunk's end
                ob[h] = beg;
                S[h] = e - beg;
                beg = e;
                smax = max(smax, S[h]);
            }
            const int nvalid = min(NPT, np - i);
            if (cur.bulk && smax <= 64) bwd_points<JT, CPD, 2, NPT, true>(cg, cjm, cphi, 0, i, nvalid, ob, S, c, gcp, lane);
            else if (cur.bulk && smax <= 96) bwd_points<JT, CPD, 3, NPT, true>(cg, cjm, cphi, 0, i, nvalid, ob, S, c, gcp, lane);
            else if (cur.bulk && smax <= 32 * kStrips) bwd_points<JT, CPD, kStrips, NPT, true>(cg, cjm, cphi, 0, i, nvalid, ob, S, c, gcp, lane);
            else {   // long lists, or a chunk that was not staged: no aggregation, global memory
                bwd_flush<JT, CPD>(c, gcp);
                for (int h = 0; h < nvalid; ++h) {
                    float g[CPD];
#pragma unroll
                    for (int k = 0; k < CPD; ++k) g[k] = __ldg(gout + (cur.p + i + h) * CPD + k);
                    for (int64_t j = cur.base + ob[h] + lane; j < cur.base + ob[h] + S[h]; j += 32) {
                        const float w = __ldg(phi + j);
                        const int64_t r = (int64_t)__ldg(jm + j);
#pragma unroll
                        for (int k = 0; k < CPD; ++k) atomicAdd(gcp + r * CPD + k, w * g[k]);
                    }
                }
            }
        }
        __syncwarp();
        Chunk nn = describe(end_after, p_after, p1, nxt.base + nxt.npairs, nnz, ej, lane, gtot, CPD);
        if (nn.m > 0 && nn.bulk) issue<JT>(nn, sw.st[s], &sw.bar[s], phi, jm, lane, sw.g[s], gout, CPD);
        p_after = nn.p + nn.m;
        end_after = load_end<CT>(cumsum, p_after, p1, lane);
        cur = nxt;
        nxt = nn;
        s ^= 1;
    }
    bwd_flush<JT, CPD>(c, gcp);
}

// THB_CSR_BULK: 0 = register-path kernels only, 1 (default) = staged kernel for 32-bit indices, 2 = for 64-bit ones too
int bulk_mode() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("THB_CSR_BULK");
        v = e ? atoi(e) : 1;
        if (v < 0 || v > 2) v = 1;
    }
    return v;
}
bool bulk_enabled() { return bulk_mode() != 0; }
// THB_CSR_BULK_BWD: the same switch for the backward kernel
int bulk_mode_bwd() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("THB_CSR_BULK_BWD");
        v = e ? atoi(e) : 2;
        if (v < 0 || v > 2) v = 2;
    }
    return v;
}

template <typename JT, typename CT>
int launch_fwd_bulk(const float* cp, const void* jm, const float* phi, const void* cs, float* out, int64_t n, int cpd, cudaStream_t st) {
    constexpr int W = BulkCfg<JT>::kWarps;
    const size_t smem = sizeof(BulkWarp<JT>) * W;
    // ranges of 2048 points per warp: many more CTAs than fit at once, dealt by the hardware as CTAs retire
    int ppw = 2048;
    while (ppw > 64 && (n + ppw - 1) / ppw < (int64_t)thb_num_sms() * W * 8) ppw >>= 1;
    const int64_t warps = (n + ppw - 1) / ppw;
    const int grid = (int)((warps + W - 1) / W);
#define THB_FWDB(C)                                                                                                        \
    do {                                                                                                                   \
        static bool attr_set[THB_MAX_DEVICES] = {};                                                                        \
        const int dev = thb_device();                                                                                      \
        if (!attr_set[dev]) {                                                                                              \
            THB_CUDA(cudaFuncSetAttribute(k_csr_forward_bulk<JT, CT, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            attr_set[dev] = true;                                                                                          \
        }                                                                                                                  \
        k_csr_forward_bulk<JT, CT, C><<<grid, 32 * W, smem, st>>>(cp, (const JT*)jm, phi, (const CT*)cs, out, n, ppw);      \
    } while (0)
    switch (cpd) {
        case 1: THB_FWDB(1); break;
        case 2: THB_FWDB(2); break;
        case 3: THB_FWDB(3); break;
        case 4: THB_FWDB(4); break;
        default: return thb_set_error(THB_E_INVALID, "cp_dim must be 1..4");
    }
#undef THB_FWDB
    THB_LAUNCHED();
    return THB_OK;
}

template <typename JT, typename CT>
int launch_bwd_bulk(const float* go, const void* jm, const float* phi, const void* cs, float* gcp, int64_t n, int cpd, cudaStream_t st) {
    constexpr int W = BulkCfg<JT>::kWarps;
    const size_t smem = sizeof(BulkWarpBwd<JT>) * W;
    int ppw = 1024;   // long ranges: the runs of one cell aggregate in registers, and the ranges of resident warps lie far apart
    while (ppw > 64 && (n + ppw - 1) / ppw < (int64_t)thb_num_sms() * W * 8) ppw >>= 1;
    const int64_t warps = (n + ppw - 1) / ppw;
    const int grid = (int)((warps + W - 1) / W);
    int64_t stride = 1;
    if (warps > 8) {
        stride = (int64_t)((double)warps * 0.6180339887) | 1;
        auto gcd = [](int64_t a, int64_t b) { while (b) { const int64_t t = a % b; a = b; b = t; } return a; };
        while (gcd(stride, warps) != 1) stride += 2;
    }
#define THB_BWDB(C)                                                                                                        \
    do {                                                                                                                   \
        static bool attr_set[THB_MAX_DEVICES] = {};                                                                        \
        const int dev = thb_device();                                                                                      \
        if (!attr_set[dev]) {                                                                                              \
            THB_CUDA(cudaFuncSetAttribute(k_csr_backward_bulk<JT, CT, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            attr_set[dev] = true;                                                                                          \
        }                                                                                                                  \
        k_csr_backward_bulk<JT, CT, C><<<grid, 32 * W, smem, st>>>(go, (const JT*)jm, phi, (const CT*)cs, gcp, n, ppw, warps, stride); \
    } while (0)
    switch (cpd) {
        case 1: THB_BWDB(1); break;
        case 2: THB_BWDB(2); break;
        case 3: THB_BWDB(3); break;
        case 4: THB_BWDB(4); break;
        default: return thb_set_error(THB_E_INVALID, "cp_dim must be 1..4");
    }
#undef THB_BWDB
    THB_LAUNCHED();
    return THB_OK;
}

}  // namespace

// grad_ctrl_pts must be zeroed by the caller (thb_eval_backward does)
int thb_csr_backward_bulk(const float* grad_out, const void* jm, int jm_is_i64, const float* phi, const void* cumsum, int cumsum_is_i64,
                          float* gcp, int64_t npoints, int cp_dim, cudaStream_t st) {
    const int mode = bulk_mode_bwd();
    if (mode == 0 || ((((uintptr_t)phi) | ((uintptr_t)jm) | ((uintptr_t)grad_out)) & 15) != 0) return 1;
    if (jm_is_i64 && mode != 2) return 1;
    if (jm_is_i64 && cumsum_is_i64) return launch_bwd_bulk<long long, long long>(grad_out, jm, phi, cumsum, gcp, npoints, cp_dim, st);
    if (jm_is_i64) return launch_bwd_bulk<long long, int>(grad_out, jm, phi, cumsum, gcp, npoints, cp_dim, st);
    if (cumsum_is_i64) return launch_bwd_bulk<int, long long>(grad_out, jm, phi, cumsum, gcp, npoints, cp_dim, st);
    return launch_bwd_bulk<int, int>(grad_out, jm, phi, cumsum, gcp, npoints, cp_dim, st);
}

int thb_csr_forward_bulk(const float* cp, const void* jm, int jm_is_i64, const float* phi, const void* cumsum, int cumsum_is_i64,
                         float* out, int64_t npoints, int cp_dim, cudaStream_t st) {
    if (!bulk_enabled() || ((((uintptr_t)phi) | ((uintptr_t)jm)) & 15) != 0) return 1;
    // Measured on config 3 (16.8 M points, 1.16 G pairs): int32 indices 2.10 ms staged against 2.40 ms on the register path
    // (instruction-bound now: 1.37 G warp-instructions at 56 % issue, the loads no longer wait); int64 indices 2.91 against
    // 2.75 ms -- the larger stages leave 12 warps per SM for a body with 64-bit key compares -- so those keep the register
    // path unless THB_CSR_BULK=2 asks for the staged kernel.
    if (jm_is_i64 && bulk_mode() != 2) return 1;
    if (jm_is_i64 && cumsum_is_i64) return launch_fwd_bulk<long long, long long>(cp, jm, phi, cumsum, out, npoints, cp_dim, st);
    if (jm_is_i64) return launch_fwd_bulk<long long, int>(cp, jm, phi, cumsum, out, npoints, cp_dim, st);
    if (cumsum_is_i64) return launch_fwd_bulk<int, long long>(cp, jm, phi, cumsum, out, npoints, cp_dim, st);
    return launch_fwd_bulk<int, int>(cp, jm, phi, cumsum, out, npoints, cp_dim, st);
}
