// Point bodies of the CSR contraction kernels (THB_eval.forward, compute_pts_cuda_kernels.cu:3-26), shared by the
// register-path kernel (thb_eval_csr.cu: loads from global memory) and the bulk-staged one (thb_eval_bulk.cu: the same
// bodies reading segments the TMA put in shared memory).
#pragma once
#include <cstdint>

#include "thb_common.cuh"

namespace thb_csr {

constexpr int kFwdStrips = 5;   // support lists up to 160 entries live in registers (one entry per lane and strip)
#ifndef THB_FWD_POINTS
#define THB_FWD_POINTS 32
#endif
constexpr int kFwdPoints = THB_FWD_POINTS;  // consecutive points per warp (a multiple of kFwdBatch)
constexpr int kFwdBatch = 8;                // points whose per-lane partial sums are reduced together

// Sum over the 32 lanes of NB = 8 values per lane at once: three exchange steps halve the number of values a lane holds
// (lanes that differ in bit 4, 3, 2 split the values between them), two butterfly steps finish.  9 shuffles per 8 points
// and component instead of 40 for eight separate butterflies.  Afterwards lane l with (l & 3) == 0 holds the total of value
// index ((l >> 4) & 1) * 4 + ((l >> 3) & 1) * 2 + ((l >> 2) & 1).
THB_DEV float reduce8(float (&v)[8], int lane) {
    float w4[4], w2[2];
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float send = b4 ? v[j] : v[j + 4];
        const float keep = b4 ? v[j + 4] : v[j];
        w4[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const float send = b3 ? w4[j] : w4[j + 2];
        const float keep = b3 ? w4[j + 2] : w4[j];
        w2[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    const float send = b2 ? w2[0] : w2[1];
    const float keep = b2 ? w2[1] : w2[0];
    float r = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    r += __shfl_xor_sync(0xffffffffu, r, 2);
    r += __shfl_xor_sync(0xffffffffu, r, 1);
    return r;
}

// A warp owns kFwdPoints consecutive points.  Per point the lanes read the segment coalesced (one entry per lane and
// strip).  The gathered control points stay in registers for as long as consecutive points have the same support list
// (points of one cell do), so the dependent second round trip of a point -- the gather through Jm -- disappears inside a
// run.  Points are taken NPT at a time: the loads of all of them are issued before any is used (bytes in flight per warp),
// and the body is specialised for the number of strips NS the longest of them needs (no per-strip branches); per-lane
// partial sums are reduced eight points at a time (reduce8).  Segment ends travel as 32-bit offsets from the warp's first
// entry.  History on config 3 (16.8 M points, 1.16 G pairs): one point at a time with five fixed strips and a butterfly
// per point and component 3.66 ms int64 / 3.71 ms int32 (2.4 G warp-instructions, issue-bound); a variant that streams
// PHI / Jm through a shared-memory ring with 16-byte cp.async 4.4 / 3.2 ms (3.2 G warp-instructions).
template <typename JT, int CPD>
struct FwdCache {
    JT key[kFwdStrips];
    float cpv[kFwdStrips][CPD];
    int runS;
};

#ifndef THB_FWD_PF_DIST
#define THB_FWD_PF_DIST 1   // groups of NPT points between the L2 prefetch and the loads (1: 2.40 ms, 2: 2.40, 3: 2.44, 4: 2.48, none: 2.50 -- config 3, int32)
#endif
THB_DEV void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// NPT points starting at entry offsets ob[h] (relative to base), lengths S[h] <= 32 * NS; a[h][k] = per-lane partial sums
// explicit shared-memory loads (volatile: they stay behind the mbarrier wait that makes the staged data visible)
// (32-bit shared-window addresses: one conversion per point, the strips are immediates)
THB_DEV unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
THB_DEV float lds_f32(unsigned a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
template <typename JT>
THB_DEV JT lds_index(unsigned a) {
    if (sizeof(JT) == 8) {
        long long v;
        asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(a));
        return (JT)v;
    } else {
        int v;
        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a));
        return (JT)v;
    }
}

// STAGED: jm / phi point into shared memory (segments staged by the TMA); otherwise streaming loads from global memory
template <typename JT, int CPD, int NS, int NPT, bool STAGED = false>
THB_DEV void fwd_points(const float* __restrict__ cp, const JT* __restrict__ jm, const float* __restrict__ phi, int64_t base,
                        const int (&ob)[NPT], const int (&S)[NPT], FwdCache<JT, CPD>& c, float (&a)[NPT][CPD], int lane) {
    JT cj[NPT][NS];
    float cw[NPT][NS];
#pragma unroll
    for (int h = 0; h < NPT; ++h) {
        const JT* pj = jm + base + ob[h] + lane;
        const float* pw = phi + base + ob[h] + lane;
        const unsigned aj = STAGED ? smem_addr(pj) : 0u, aw = STAGED ? smem_addr(pw) : 0u;
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const bool ok = s * 32 + lane < S[h];
            if (STAGED) {
                cj[h][s] = (JT)-1;
                cw[h][s] = 0.0f;
                if (ok) {
                    cj[h][s] = lds_index<JT>(aj + s * 32 * (unsigned)sizeof(JT));
                    cw[h][s] = lds_f32(aw + s * 128u);
                }
            } else {
                cj[h][s] = ok ? __ldcs(pj + s * 32) : (JT)-1;   // read once: streaming (evict-first) loads keep the control
                cw[h][s] = ok ? __ldcs(pw + s * 32) : 0.0f;    // points / the gradient rows resident in L2
            }
        }
    }
#pragma unroll
    for (int h = 0; h < NPT; ++h) {
        bool same = S[h] == c.runS;
#pragma unroll
        for (int s = 0; s < NS; ++s) same = same && (cj[h][s] == c.key[s]);
        if (!__all_sync(0xffffffffu, same)) {
#pragma unroll
            for (int s = 0; s < kFwdStrips; ++s) c.key[s] = s < NS ? cj[h][s < NS ? s : 0] : (JT)-1;
#pragma unroll
            for (int s = 0; s < kFwdStrips; ++s) {   // strips beyond NS hold zeros: a later, wider body multiplies them by 0
#pragma unroll
                for (int k = 0; k < CPD; ++k)
                    c.cpv[s][k] = (s < NS && cj[h][s < NS ? s : 0] >= 0) ? __ldg(cp + (int64_t)cj[h][s < NS ? s : 0] * CPD + k) : 0.0f;
            }
            c.runS = S[h];
        }
#pragma unroll
        for (int k = 0; k < CPD; ++k) a[h][k] = 0.0f;
#pragma unroll
        for (int s = 0; s < NS; ++s) {
#pragma unroll
            for (int k = 0; k < CPD; ++k) a[h][k] = fmaf(cw[h][s], c.cpv[s][k], a[h][k]);
        }
    }
}


constexpr int kStrips = kFwdStrips;  // support lists up to 160 entries are aggregated across points

// grad_ctrl_pts[Jm[j]] += PHI[j] * grad_out[i].  A warp owns a run of consecutive points; per point the lanes read the
// segment coalesced, NPT points at a time (all loads first), body specialised for the number of strips.  While consecutive
// points have the same support list the products are accumulated in registers; the atomics are issued only when the list
// changes (or the warp's run ends): cp_dim * S atomics per RUN of points instead of per point -- the warp-aggregated
// scatter-add.
template <typename JT, int CPD>
struct BwdCache {
    JT key[kStrips];
    float acc[kStrips][CPD];
    int runS;
};

// (Vector reductions -- red.global.add.v2 / .v4.f32, a row of three floats as one 8-byte pair plus one scalar -- were
// measured slower than three scalar reductions here, and the second code path made the kernel miss in the instruction
// cache: 4.1 instead of 3.0 ms on config 3.  The kernel is kept small on purpose.)
// Straight-line: every strip tests its own key (strips past the run hold -1) and the reductions are predicated -- the
// flush is inlined at every point of every specialised body, and a branch per strip made it ~150 instructions with the
// register moves at the joins (a quarter of the kernel's instructions on config 3, where a run is ~7 points).
template <typename JT, int CPD>
THB_DEV void bwd_flush(BwdCache<JT, CPD>& c, float* __restrict__ gcp) {
#pragma unroll
    for (int s = 0; s < kStrips; ++s) {
        if (c.key[s] >= 0) {
            float* d = gcp + (int64_t)c.key[s] * CPD;
#pragma unroll
            for (int k = 0; k < CPD; ++k) atomicAdd(d + k, c.acc[s][k]);
        }
#pragma unroll
        for (int k = 0; k < CPD; ++k) c.acc[s][k] = 0.0f;
        c.key[s] = (JT)-1;
    }
    c.runS = -1;
}

#ifndef THB_BWD_LD
#define THB_BWD_LD __ldg   // (streaming __ldcs loads measured slower here: 4.1 vs 3.0 ms on config 3, int32)
#endif
// STAGED: gout / jm / phi point into shared memory (gout: the rows of the chunk, row0 relative to it)
template <typename JT, int CPD, int NS, int NPT, bool STAGED = false>
THB_DEV void bwd_points(const float* __restrict__ gout, const JT* __restrict__ jm, const float* __restrict__ phi, int64_t base,
                        int64_t row0, int nvalid, const int (&ob)[NPT], const int (&S)[NPT], BwdCache<JT, CPD>& c,
                        float* __restrict__ gcp, int lane) {
    JT cj[NPT][NS];
    float cw[NPT][NS], g[NPT][CPD];
#pragma unroll
    for (int h = 0; h < NPT; ++h) {
        const JT* pj = jm + base + ob[h] + lane;
        const float* pw = phi + base + ob[h] + lane;
        const int64_t row = row0 + (h < nvalid ? h : 0);
#pragma unroll
        for (int k = 0; k < CPD; ++k) g[h][k] = STAGED ? lds_f32(smem_addr(gout + row * CPD) + 4u * k) : __ldg(gout + row * CPD + k);
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const bool ok = s * 32 + lane < S[h];
            if (STAGED) {
                cj[h][s] = (JT)-1;
                cw[h][s] = 0.0f;
                if (ok) {
                    cj[h][s] = lds_index<JT>(smem_addr(pj) + s * 32 * (unsigned)sizeof(JT));
                    cw[h][s] = lds_f32(smem_addr(pw) + s * 128u);
                }
            } else {
                cj[h][s] = ok ? THB_BWD_LD(pj + s * 32) : (JT)-1;
                cw[h][s] = ok ? THB_BWD_LD(pw + s * 32) : 0.0f;
            }
        }
    }
#pragma unroll
    for (int h = 0; h < NPT; ++h) {
        if (h >= nvalid) break;   // warp-uniform
        bool same = S[h] == c.runS;
#pragma unroll
        for (int s = 0; s < NS; ++s) same = same && (cj[h][s] == c.key[s]);
        if (!__all_sync(0xffffffffu, same)) {
            bwd_flush<JT, CPD>(c, gcp);
#pragma unroll
            for (int s = 0; s < NS; ++s) c.key[s] = cj[h][s];
            c.runS = S[h];
        }
#pragma unroll
        for (int s = 0; s < NS; ++s) {
#pragma unroll
            for (int k = 0; k < CPD; ++k) c.acc[s][k] = fmaf(cw[h][s], g[h][k], c.acc[s][k]);
        }
    }
}


}  // namespace thb_csr
