// Span search, 1-D basis API kernels, active-cell lookup and the plan (points grouped by active cell).
//   find_span_array_jax      THB/bspline_funcs.py:77-83      -> thb_find_spans
//   basis_fns_vmap / der_*   THB/bspline_funcs.py:185-194    -> thb_basis_1d
//   compute_active_span      THB/core.py:157-189             -> thb_active_span, thb_plan_build
//   Jm construction          THB/core.py:644-650             -> thb_expand_supports
#include "thb_common.cuh"

namespace {

constexpr int kMaxSmemKnots = 2048;

__global__ void __launch_bounds__(256) k_find_spans(const float* __restrict__ params, int64_t stride, int64_t n,
                                                    const float* __restrict__ knots, int nknots, int degree,
                                                    int32_t* __restrict__ spans) {
    __shared__ float sk[kMaxSmemKnots];
    const bool in_smem = nknots <= kMaxSmemKnots;
    if (in_smem) {
        for (int i = threadIdx.x; i < nknots; i += blockDim.x) sk[i] = knots[i];
        __syncthreads();
    }
    const float* U = in_smem ? sk : knots;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        spans[i] = span_right(U, nknots, degree, params[i * stride]);
}

__global__ void __launch_bounds__(256) k_basis_1d(const float* __restrict__ params, int64_t stride, int64_t n,
                                                  const float* __restrict__ knots, int nknots, int degree,
                                                  float* __restrict__ values, float* __restrict__ derivs) {
    __shared__ float sk[kMaxSmemKnots];
    const bool in_smem = nknots <= kMaxSmemKnots;
    if (in_smem) {
        for (int i = threadIdx.x; i < nknots; i += blockDim.x) sk[i] = knots[i];
        __syncthreads();
    }
    const float* U = in_smem ? sk : knots;
    const int nf = nknots - degree - 1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float u = params[i * stride];
        const int s = span_right(U, nknots, degree, u);
        float N[THB_MAX_DEG + 1], dN[THB_MAX_DEG + 1];
        const bool ok = s >= degree && s < nf;
        if (ok) cox_de_boor_rt(degree, s, u, U, N, derivs ? dN : nullptr);
        for (int r = 0; r <= degree; ++r) {
            values[i * (degree + 1) + r] = ok ? N[r] : 0.0f;
            if (derivs) derivs[i * (degree + 1) + r] = ok ? dN[r] : 0.0f;
        }
    }
}

__global__ void __launch_bounds__(256) k_active_span(const __grid_constant__ thb_space_desc sp, const float* __restrict__ params,
                                                     int64_t n, int32_t* __restrict__ slot_out, int32_t* __restrict__ nsupp_out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float u[3] = {0.f, 0.f, 0.f};
        for (int k = 0; k < sp.ndim; ++k) u[k] = params[i * sp.ndim + k];
        const int slot = locate_slot(sp, u);
        slot_out[i] = slot;
        if (nsupp_out) nsupp_out[i] = slot >= 0 ? __ldg(sp.cellinfo + (int64_t)slot * THB_INFO_W + 5) : 0;
    }
}

// One warp per point: copies the cell's support ids into the point's segment (coalesced both ways).
template <typename JT>
__global__ void __launch_bounds__(256) k_expand_supports(const __grid_constant__ thb_space_desc sp, const int32_t* __restrict__ slot,
                                                         const int64_t* __restrict__ offsets, int64_t n, JT* __restrict__ jm_out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp0; i < n; i += nwarps) {
        const int s = slot[i];
        if (s < 0) continue;
        const int32_t* ci = sp.cellinfo + (int64_t)s * THB_INFO_W;
        const int soff = __ldg(ci + 4), S = __ldg(ci + 5);
        const int64_t o = offsets[i];
        for (int j = lane; j < S; j += 32) jm_out[o + j] = (JT)__ldg(sp.supp_jm + soff + j);
    }
}

// ---- plan: counting sort of the points by slot ----------------------------------------------------------
// Finest-level span with the knots of that level staged in shared memory.  First guess from a uniform
// partition, corrected by walking (exact for 'uniform-ish' knots in <= 2 steps), binary search otherwise;
// the result is the same span as span_right() for every in-domain u.
// Finest-level cell of u.  B = finest breakpoints (B[i] = U[p+i]) in shared memory, nc0 level-0 cells, L levels
// above level 0.  Levels are dyadically nested with bit-identical breakpoints and the refinement inserts
// midpoints, so inside a level-0 cell the finest breakpoints are uniform up to rounding: guess the level-0 cell
// from the uniform partition (corrected by one if needed), then the sub-cell by linear interpolation, and VERIFY
// against the two bracketing breakpoints; a one-step correction and finally a binary search catch every other
// knot vector.  The verified result is exactly searchsorted(U, u, 'right') - 1 - p, with the reference's rule
// that u == U[-1] belongs to the last cell (THB/bspline_funcs.py:77-83).
THB_DEV float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

THB_DEV int finest_cell(const float* B, int nc0, int L, float scale0, float u0, float u1, float u) {
    if (!(u >= u0) || u > u1) return -1;  // also rejects NaN
    const int nc = nc0 << L;
    // level-0 cell: uniform guess, branch-free correction by one
    int c0 = min((int)((u - u0) * scale0), nc0 - 1);
    c0 = min(max(c0 + (int)(u >= B[(c0 + 1) << L]) - (int)(u < B[c0 << L]), 0), nc0 - 1);
    const float lo0 = B[c0 << L], hi0 = B[(c0 + 1) << L];
    // sub-cell by linear interpolation (the guess only has to be close: it is verified below)
    const int f = min(max((int)((u - lo0) * rcp_approx(hi0 - lo0) * (float)(1 << L)), 0), (1 << L) - 1);
    int c = (c0 << L) + f;
    c = min(max(c + (int)(u >= B[c + 1]) - (int)(u < B[c]), 0), nc - 1);  // rounding of the guess: one step at most
    if (u >= B[c] && (u < B[c + 1] || c == nc - 1)) return c;            // verified: taken by (almost) every point
    if (u == u1) return nc - 1;
    int lo = 0, hi = nc;  // arbitrary knot vectors: binary search, invariant B[lo] <= u < B[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (B[mid] <= u) lo = mid; else hi = mid;
    }
    return lo;
}

// Same result through a lookup table: bins at most 0.4 finest cells wide, lut[bin] = cell of the bin's left edge,
// so the cell of u is lut[bin] or the next one.  Verified like above; any miss falls back to finest_cell().
THB_DEV int finest_cell_lut(const float* B, const int32_t* __restrict__ lut, int nbins, float lscale, int nc, float u0, float u1, float u) {
    if (!(u >= u0) || u > u1) return -1;
    int c = __ldg(lut + min((int)((u - u0) * lscale), nbins - 1));
    c = min(c + (int)(u >= B[c + 1]), nc - 1);
    if (u >= B[c] && (u < B[c + 1] || c == nc - 1)) return c;
    return -2;  // not verified
}

constexpr int kPlanSmemKnots = 3 * 1024;

// slot of a point given the finest-level knots in shared memory; the cell_slot entries of ALL levels are
// fetched at once (independent loads) and the finest active one is selected.
struct PlanKnots {
    int off[3];  // offset of the first BREAKPOINT (knot p) of each finest knot vector inside the shared-memory copy
    int nc0[3];  // level-0 cells
    float scale[3], u0[3], u1[3];
};

THB_DEV int locate_slot_fast(const thb_space_desc& sp, const float* sk, const PlanKnots& pk, const float* u) {
    const int L = sp.nlevels - 1;
    int cf[3] = {0, 0, 0};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (k < sp.ndim) {
            int c = -2;
            if (sp.span_lut)
                c = finest_cell_lut(sk + pk.off[k], sp.span_lut + sp.lut_off[k], sp.lut_n[k], sp.lut_scale[k], pk.nc0[k] << L, pk.u0[k], pk.u1[k], u[k]);
            if (c == -2) c = finest_cell(sk + pk.off[k], pk.nc0[k], L, pk.scale[k], pk.u0[k], pk.u1[k], u[k]);
            if (c < 0) return -1;
            cf[k] = c;
        }
    }
    if (sp.finest_slot) {  // one lookup: finest cell -> slot of the active cell that contains it
        int lin = cf[0] * sp.ncells[L][1] + cf[1];
        if (sp.ndim == 3) lin = lin * sp.ncells[L][2] + cf[2];
        return __ldg(sp.finest_slot + lin);
    }
    // independent lookups at every level (issued back to back), finest active one wins
    int slot = -1;
#pragma unroll 4
    for (int lev = 0; lev <= L; ++lev) {
        const int sh = L - lev;
        int lin = (cf[0] >> sh) * sp.ncells[lev][1] + (cf[1] >> sh);
        if (sp.ndim == 3) lin = lin * sp.ncells[lev][2] + (cf[2] >> sh);
        const int cand = __ldg(sp.cell_slot + sp.cell_slot_off[lev] + lin);
        if (cand >= 0) slot = cand;
    }
    return slot;
}

// Stages the finest knots in shared memory; false when they do not fit (the caller then uses locate_slot()).
THB_DEV bool stage_finest_knots(const thb_space_desc& sp, float* sk, PlanKnots& pk) {
    const int L = sp.nlevels - 1;
    int tot = 0;
    for (int k = 0; k < sp.ndim; ++k) tot += sp.knot_len[L][k];
    const bool fits = tot <= kPlanSmemKnots;
    int pos = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        pk.off[k] = 0;
        pk.nc0[k] = 1;
        pk.scale[k] = pk.u0[k] = pk.u1[k] = 0.f;
        if (k < sp.ndim && fits) {
            const float* g = sp.knots + sp.knot_off[L][k];
            const int p = sp.degree[k], nc = sp.ncells[L][k];
            for (int i = threadIdx.x; i < sp.knot_len[L][k]; i += blockDim.x) sk[pos + i] = g[i];
            pk.off[k] = pos + p;
            pk.nc0[k] = sp.ncells[0][k];
            pk.u0[k] = g[p];
            pk.u1[k] = g[p + nc];
            pk.scale[k] = (float)pk.nc0[k] / (pk.u1[k] - pk.u0[k]);
            pos += sp.knot_len[L][k];
        }
    }
    __syncthreads();
    return fits;
}

// pass 1: slot per point + histogram.  Lanes of a warp that hit the same slot are combined with
// __match_any_sync so that one atomic is issued per distinct slot per warp (grid-ordered points share cells).
__global__ void __launch_bounds__(256) k_plan_count(const __grid_constant__ thb_space_desc sp, const float* __restrict__ params,
                                                    int64_t n, int32_t* __restrict__ slot_out, int32_t* __restrict__ count) {
    __shared__ float sk[kPlanSmemKnots];
    PlanKnots pk;
    const bool in_smem = stage_finest_knots(sp, sk, pk);
    const int lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t nround = (n + 2 * stride - 1) / (2 * stride) * (2 * stride);  // whole warps stay in the loop (match_any)
    const int d = sp.ndim;
    // two points per thread per iteration: their loads and lookups are independent and overlap
    for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < nround; i0 += 2 * stride) {
        int slot[2];
        float u[2][3];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int64_t i = i0 + h * stride;
            const int64_t j = i < n ? i : 0;
            u[h][0] = params[j * d];
            u[h][1] = params[j * d + 1];
            u[h][2] = d == 3 ? params[j * d + 2] : 0.f;
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int64_t i = i0 + h * stride;
            slot[h] = -1;
            if (i < n) {
                slot[h] = in_smem ? locate_slot_fast(sp, sk, pk, u[h]) : locate_slot(sp, u[h]);
                slot_out[i] = slot[h];
            }
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const unsigned peers = __match_any_sync(0xffffffffu, slot[h]);
            if (slot[h] >= 0 && lane == (__ffs(peers) - 1)) atomicAdd(count + slot[h], __popc(peers));
        }
    }
}

// pass 1, fast form (needs the span lookup tables, the finest-cell slot map and knots that fit in shared memory):
// straight-line code per point -- bin -> candidate cell -> one comparison -> verification against the two bracketing
// breakpoints -> one slot lookup.  A point whose span is not verified (possible only for knot vectors the tables do
// not resolve) is appended to `fail_list` and handled by k_plan_fixup with the robust search.
constexpr int kPlanSmemLut = 3 * 1024;  // span lookup tables staged in shared memory when they fit (else read through L1)

#ifndef THB_COUNT_PTS
#define THB_COUNT_PTS 2
#endif
constexpr int kCountPts = THB_COUNT_PTS;  // points per thread and iteration of the count pass

template <int D, bool SLUT, bool EXACT>
__global__ void __launch_bounds__(256) k_plan_count_fast(const __grid_constant__ thb_space_desc sp, const float* __restrict__ params,
                                                         unsigned n, int32_t* __restrict__ slot_out, int32_t* __restrict__ count,
                                                         int32_t* __restrict__ fail_list, int32_t* __restrict__ counters) {
    __shared__ float sk[kPlanSmemKnots];
    __shared__ int s_lut[SLUT ? kPlanSmemLut : 1];
    PlanKnots pk;
    stage_finest_knots(sp, sk, pk);
    const int L = sp.nlevels - 1;
    int loff[D], nbm1[D], last[D], off[D];
    float ls[D], u0[D], u1[D];
#pragma unroll
    for (int k = 0; k < D; ++k) {
        loff[k] = sp.lut_off[k];
        nbm1[k] = sp.lut_n[k] - 1;
        last[k] = sp.ncells[L][k] - 1 + pk.off[k];   // index into sk of the last cell's left breakpoint
        ls[k] = sp.lut_scale[k];
        off[k] = pk.off[k];
        u0[k] = pk.u0[k];
        u1[k] = pk.u1[k];
    }
    if (SLUT) {
        const int tot = sp.lut_off[D - 1] + sp.lut_n[D - 1];
        for (int i = threadIdx.x; i < tot; i += blockDim.x) s_lut[i] = sp.span_lut[i];
        __syncthreads();
    }
    const int n1 = sp.ncells[L][1], n2 = D == 3 ? sp.ncells[L][2] : 1;
    const int32_t* __restrict__ fslot = sp.finest_slot;
    const int32_t* __restrict__ glut = sp.span_lut;
    const unsigned lane = threadIdx.x & 31;
    const unsigned stride = gridDim.x * blockDim.x;
    const unsigned nround = (n + kCountPts * stride - 1) / (kCountPts * stride) * (kCountPts * stride);  // n < 2^31 and stride < 2^20: no overflow
    for (unsigned i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < nround; i0 += kCountPts * stride) {
        float u[kCountPts][D];
        int slot[kCountPts];
#pragma unroll
        for (int h = 0; h < kCountPts; ++h) {
            const unsigned i = i0 + h * stride;
            const float* src = params + (size_t)(i < n ? i : 0u) * D;
#pragma unroll
            for (int k = 0; k < D; ++k) u[h][k] = src[k];
        }
#pragma unroll
        for (int h = 0; h < kCountPts; ++h) {
            const unsigned i = i0 + h * stride;
            bool ok = true, inside = true;
            int c[D];
#pragma unroll
            for (int k = 0; k < D; ++k) {
                // branch-free: bin -> candidate cell c0 -> (b0, b1, b2) = its breakpoints and the next one -> the cell is c0
                // or c0 + 1 -> verified against its own two breakpoints
                const float x = u[h][k];
                const bool in = (x >= u0[k]) & (x <= u1[k]);   // false for NaN
                int bin = min((int)((x - u0[k]) * ls[k]), nbm1[k]);
                bin = in ? bin : 0;
                const int c0 = (SLUT ? s_lut[loff[k] + bin] : __ldg(glut + loff[k] + bin)) + off[k];
                if (EXACT) {
                    // the host PROVED (packed.py: same float32 bin formula evaluated at every breakpoint, at most one
                    // breakpoint per bin) that the cell of any x of this bin is lut[bin] or the next one: one comparison
                    const int cc = min(c0 + (int)(x >= sk[c0 + 1]), last[k]);
                    inside &= in;
                    c[k] = cc - off[k];
                } else {
                    const float b0 = sk[c0], b1 = sk[c0 + 1], b2 = sk[c0 + 2];   // the p repeated end knots pad the last read
                    const bool inc = x >= b1;
                    const float lo = inc ? b1 : b0, hi = inc ? b2 : b1;
                    const int cc = min(c0 + (inc ? 1 : 0), last[k]);
                    ok &= (x >= lo) & ((x < hi) | (cc == last[k]));
                    inside &= in;
                    c[k] = cc - off[k];
                }
            }
            int sl = -1;
            if (i < n) {
                if (inside & ok) {
                    int lin = c[0] * n1 + c[1];
                    if (D == 3) lin = lin * n2 + c[2];
                    sl = __ldg(fslot + lin);
                    slot_out[i] = sl;
                } else if (!inside) {
                    slot_out[i] = -1;
                } else {
                    fail_list[atomicAdd(counters + 3, 1)] = (int32_t)i;  // robust path in k_plan_fixup
                }
            }
            slot[h] = sl;
        }
#pragma unroll
        for (int h = 0; h < kCountPts; ++h) {
            const unsigned peers = __match_any_sync(0xffffffffu, slot[h]);
            if (slot[h] >= 0 && lane == (unsigned)(__ffs(peers) - 1)) atomicAdd(count + slot[h], __popc(peers));
        }
    }
}

// pass 1, vector form of the EXACT case (host-proven span table in shared memory, one finest-cell -> slot lookup):
// a thread owns FOUR CONSECUTIVE points -- 3 (2-D: 2) LDG.128 and one STG.128 instead of 12 scalar loads and 4 stores --
// and 113 instructions per point become ~70:
//   * the staged table holds indices into the staged knots, and the breakpoint after the last cell is replaced by +inf
//     in the staged copy, so the candidate needs no clamp: cell = lut[bin] + (x >= knot[lut[bin] + 1]);
//   * bin = umin(float2uint_rz((x - u0) * scale), nb - 1): the same value as the proven formula for every x of the
//     domain, and a valid index for anything else (NaN / negative -> 0);
//   * the histogram adds one RED per RUN of equal slots among the thread's four points (neighbouring points share cells;
//     no __match_any_sync -- a warp-wide match costs more issue slots than the REDs it saves).
#ifndef THB_COUNT_VEC_MIN
#define THB_COUNT_VEC_MIN 4000000   // points from which the vector histogram kernel is used
#endif
constexpr long long kCountVecMin = THB_COUNT_VEC_MIN;
#ifndef THB_COUNT_VEC_GROUPS
#define THB_COUNT_VEC_GROUPS 2
#endif
template <int D>
__global__ void __launch_bounds__(256) k_plan_count_vec(const __grid_constant__ thb_space_desc sp, const float* __restrict__ params,
                                                        unsigned n, int32_t* __restrict__ slot_out, int32_t* __restrict__ count) {
    constexpr int G = THB_COUNT_VEC_GROUPS;   // groups of four points per thread and iteration (independent loads in flight)
    constexpr int NV = D == 3 ? 3 : 2;        // float4 per group
    __shared__ float sk[kPlanSmemKnots];
    __shared__ unsigned s_lut[kPlanSmemLut];  // BYTE offset into sk of the candidate cell's left breakpoint
    PlanKnots pk;
    stage_finest_knots(sp, sk, pk);
    const int L = sp.nlevels - 1;
    unsigned nbm1[D];
    const unsigned* lutk[D];
    float ls[D], u0[D], u1[D];
#pragma unroll
    for (int k = 0; k < D; ++k) {
        nbm1[k] = (unsigned)(sp.lut_n[k] - 1);
        lutk[k] = s_lut + sp.lut_off[k];
        ls[k] = sp.lut_scale[k];
        u0[k] = pk.u0[k];
        u1[k] = pk.u1[k];
        for (int i = threadIdx.x; i < sp.lut_n[k]; i += blockDim.x)
            s_lut[sp.lut_off[k] + i] = 4u * (unsigned)(sp.span_lut[sp.lut_off[k] + i] + pk.off[k]);
        if (threadIdx.x == 0) sk[pk.off[k] + sp.ncells[L][k]] = __int_as_float(0x7f800000);
    }
    __syncthreads();
    // 4 * lin = sum_k 4 * (i_k - off_k) * m_k with i_k the index into sk: everything stays in byte offsets
    const unsigned n1 = sp.ncells[L][1], n2 = D == 3 ? sp.ncells[L][2] : 1;
    const unsigned m0 = D == 3 ? n1 * n2 : n1, m1 = D == 3 ? n2 : 1;
    const unsigned lin_base = 0u - 4u * (D == 3 ? ((unsigned)pk.off[0] * n1 + (unsigned)pk.off[1]) * n2 + (unsigned)pk.off[2]
                                                : (unsigned)pk.off[0] * n1 + (unsigned)pk.off[1]);
    const char* const skb = reinterpret_cast<const char*>(sk);
    const char* const fslot = reinterpret_cast<const char*>(sp.finest_slot);
    // byte offset of the point's finest cell in finest_slot, or 0xffffffff outside the domain (NaN included)
    auto locate = [&](const float* x) -> unsigned {
        bool inside = true;
        unsigned lin = lin_base;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            inside &= (x[k] >= u0[k]) & (x[k] <= u1[k]);
            const unsigned bin = min(__float2uint_rz((x[k] - u0[k]) * ls[k]), nbm1[k]);
            const unsigned c0 = lutk[k][bin];
            const unsigned cc = c0 + ((x[k] >= *reinterpret_cast<const float*>(skb + c0 + 4)) ? 4u : 0u);
            lin += cc * (k == 0 ? m0 : (k == 1 ? m1 : 1u));
        }
        return inside ? lin : 0xffffffffu;
    };
    const unsigned nfull = n >> 2;   // groups with four points
    const unsigned stride = gridDim.x * blockDim.x;
    const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
    const float4* __restrict__ p4 = reinterpret_cast<const float4*>(params);
    for (unsigned g0 = tid; g0 < nfull; g0 += G * stride) {
        float u[G][4 * D];
#pragma unroll
        for (int h = 0; h < G; ++h) {
            const unsigned g = min(g0 + h * stride, nfull - 1);
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const float4 q = __ldg(p4 + (size_t)g * NV + v);
                u[h][4 * v] = q.x; u[h][4 * v + 1] = q.y; u[h][4 * v + 2] = q.z; u[h][4 * v + 3] = q.w;
            }
        }
#pragma unroll
        for (int h = 0; h < G; ++h) {
            const unsigned g = g0 + h * stride;
            if (g >= nfull) break;
            unsigned lin[4];
            int sl[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) lin[q] = locate(&u[h][q * D]);
#pragma unroll
            for (int q = 0; q < 4; ++q) sl[q] = lin[q] != 0xffffffffu ? __ldg(reinterpret_cast<const int32_t*>(fslot + lin[q])) : -1;
            reinterpret_cast<int4*>(slot_out)[g] = make_int4(sl[0], sl[1], sl[2], sl[3]);
            int cur = sl[0], cnt = 1;
#pragma unroll
            for (int q = 1; q < 4; ++q) {
                if (sl[q] == cur) {
                    ++cnt;
                } else {
                    if (cur >= 0) atomicAdd(count + cur, cnt);
                    cur = sl[q];
                    cnt = 1;
                }
            }
            if (cur >= 0) atomicAdd(count + cur, cnt);
        }
    }
    // the last one to three points
    const unsigned i = 4u * nfull + tid;
    if (i < n) {
        float x[D];
#pragma unroll
        for (int k = 0; k < D; ++k) x[k] = __ldg(params + (size_t)i * D + k);
        const unsigned lin = locate(x);
        const int sl = lin != 0xffffffffu ? __ldg(reinterpret_cast<const int32_t*>(fslot + lin)) : -1;
        slot_out[i] = sl;
        if (sl >= 0) atomicAdd(count + sl, 1);
    }
}

__global__ void __launch_bounds__(256) k_plan_fixup(const __grid_constant__ thb_space_desc sp, const float* __restrict__ params,
                                                    const int32_t* __restrict__ fail_list, const int32_t* __restrict__ counters,
                                                    int32_t* __restrict__ slot_out, int32_t* __restrict__ count) {
    const int nfail = counters[3];
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < nfail; j += gridDim.x * blockDim.x) {
        const int64_t i = fail_list[j];
        float u[3] = {0.f, 0.f, 0.f};
        for (int k = 0; k < sp.ndim; ++k) u[k] = params[i * sp.ndim + k];
        const int sl = locate_slot(sp, u);
        slot_out[i] = sl;
        if (sl >= 0) atomicAdd(count + sl, 1);
    }
}

// pass 2 (single CTA): exclusive scan of the histogram -> cell_start; number of work items per cell and
// their exclusive scan -> cell_cursor is reused to hold the first item index of each cell.
// Every thread owns a CONTIGUOUS run of slots (a multiple of 4, read as int4): it sums its run, one block-wide scan
// of the 1024 run totals follows (warp shuffles, two barriers in all), then it re-reads its run (L1 hits) and writes
// the prefix sums.  The earlier round-by-round form paid two barriers per 4096 slots (13 rounds for C3's 52 k slots).
__global__ void __launch_bounds__(1024) k_plan_scan(const int32_t* __restrict__ count, int32_t* __restrict__ start,
                                                    int32_t* __restrict__ item_start, int nslots, int ppi,
                                                    int32_t* __restrict__ counters) {
    __shared__ int s_wp[32], s_wi[32];
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const int per = (((nslots + 1023) >> 10) + 3) & ~3;
    const int b0 = t * per, b1 = min(b0 + per, nslots);
    const int sh = (ppi & (ppi - 1)) == 0 ? 31 - __clz(ppi) : -1;  // points_per_item is normally a power of two
    auto items_of = [&](int c) { return sh >= 0 ? (c + ppi - 1) >> sh : (c + ppi - 1) / ppi; };
    int tp = 0, ti = 0;
#pragma unroll 4
    for (int i = b0; i < b1; i += 4) {
        int4 c = make_int4(0, 0, 0, 0);
        if (i + 3 < nslots) {
            c = *reinterpret_cast<const int4*>(count + i);  // count has nslots+1 entries in a 16-byte aligned row
        } else {
            c.x = count[i];
            if (i + 1 < nslots) c.y = count[i + 1];
            if (i + 2 < nslots) c.z = count[i + 2];
        }
        tp += c.x + c.y + c.z + c.w;
        ti += items_of(c.x) + items_of(c.y) + items_of(c.z) + items_of(c.w);
    }
    int sp_ = tp, si_ = ti;  // inclusive warp scan of the run totals
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int a = __shfl_up_sync(0xffffffffu, sp_, o), b2 = __shfl_up_sync(0xffffffffu, si_, o);
        if (lane >= o) { sp_ += a; si_ += b2; }
    }
    if (lane == 31) { s_wp[w] = sp_; s_wi[w] = si_; }
    __syncthreads();
    if (w == 0) {
        int a = s_wp[lane], b2 = s_wi[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int x = __shfl_up_sync(0xffffffffu, a, o), y = __shfl_up_sync(0xffffffffu, b2, o);
            if (lane >= o) { a += x; b2 += y; }
        }
        s_wp[lane] = a;
        s_wi[lane] = b2;
    }
    __syncthreads();
    int rp = (w ? s_wp[w - 1] : 0) + sp_ - tp;
    int ri = (w ? s_wi[w - 1] : 0) + si_ - ti;
#pragma unroll 4
    for (int i = b0; i < b1; i += 4) {
        if (i + 3 < nslots) {
            const int4 c = *reinterpret_cast<const int4*>(count + i);
            int4 ps, is;
            ps.x = rp; is.x = ri; rp += c.x; ri += items_of(c.x);
            ps.y = rp; is.y = ri; rp += c.y; ri += items_of(c.y);
            ps.z = rp; is.z = ri; rp += c.z; ri += items_of(c.z);
            ps.w = rp; is.w = ri; rp += c.w; ri += items_of(c.w);
            *reinterpret_cast<int4*>(start + i) = ps;
            *reinterpret_cast<int4*>(item_start + i) = is;
        } else {
            for (int k = i; k < nslots; ++k) {
                const int c = count[k];
                start[k] = rp;
                item_start[k] = ri;
                rp += c;
                ri += items_of(c);
            }
        }
    }
    if (t == 1023) {
        start[nslots] = s_wp[31];
        counters[0] = s_wi[31];  // number of work items
        counters[1] = 0;         // dynamic scheduler cursor
        counters[2] = s_wp[31];  // number of in-domain points
    }
}

// pass 2, two-kernel form for larger hierarchies: every CTA owns 4096 consecutive slots (int4 per thread, coalesced).
// k_plan_scan_sums writes the per-CTA totals, k_plan_scan_write adds the totals of the CTAs before it (every CTA sums
// them itself: a few hundred values at most) to its local scan.  Two short launches instead of one CTA walking all slots.
constexpr int kScanTile = 4096;

THB_DEV void scan_tile_load(const int32_t* __restrict__ count, int nslots, int i0, int ppi, int sh, int (&c)[4], int (&m)[4]) {
    int4 v = make_int4(0, 0, 0, 0);
    if (i0 + 3 < nslots) {
        v = *reinterpret_cast<const int4*>(count + i0);
    } else {
        if (i0 < nslots) v.x = count[i0];
        if (i0 + 1 < nslots) v.y = count[i0 + 1];
        if (i0 + 2 < nslots) v.z = count[i0 + 2];
    }
    c[0] = v.x; c[1] = v.y; c[2] = v.z; c[3] = v.w;
#pragma unroll
    for (int k = 0; k < 4; ++k) m[k] = sh >= 0 ? (c[k] + ppi - 1) >> sh : (c[k] + ppi - 1) / ppi;
}

// block-wide inclusive scan of one (points, items) pair per thread; returns the block totals in (tot_p, tot_i)
THB_DEV void scan_block(int& sp_, int& si_, int& tot_p, int& tot_i, int* s_wp, int* s_wi) {
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int a = __shfl_up_sync(0xffffffffu, sp_, o), b2 = __shfl_up_sync(0xffffffffu, si_, o);
        if (lane >= o) { sp_ += a; si_ += b2; }
    }
    if (lane == 31) { s_wp[w] = sp_; s_wi[w] = si_; }
    __syncthreads();
    if (w == 0) {
        const int nw = blockDim.x >> 5;
        int a = lane < nw ? s_wp[lane] : 0, b2 = lane < nw ? s_wi[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int x = __shfl_up_sync(0xffffffffu, a, o), y = __shfl_up_sync(0xffffffffu, b2, o);
            if (lane >= o) { a += x; b2 += y; }
        }
        s_wp[lane] = a;
        s_wi[lane] = b2;
    }
    __syncthreads();
    sp_ += w ? s_wp[w - 1] : 0;
    si_ += w ? s_wi[w - 1] : 0;
    tot_p = s_wp[31];
    tot_i = s_wi[31];
}

__global__ void __launch_bounds__(1024) k_plan_scan_sums(const int32_t* __restrict__ count, int nslots, int ppi,
                                                         int32_t* __restrict__ sums) {
    __shared__ int s_wp[32], s_wi[32];
    const int sh = (ppi & (ppi - 1)) == 0 ? 31 - __clz(ppi) : -1;
    int c[4], m[4];
    scan_tile_load(count, nslots, blockIdx.x * kScanTile + 4 * threadIdx.x, ppi, sh, c, m);
    int sp_ = c[0] + c[1] + c[2] + c[3], si_ = m[0] + m[1] + m[2] + m[3], tp, ti;
    scan_block(sp_, si_, tp, ti, s_wp, s_wi);
    if (threadIdx.x == 0) { sums[2 * blockIdx.x] = tp; sums[2 * blockIdx.x + 1] = ti; }
}

__global__ void __launch_bounds__(1024) k_plan_scan_write(const int32_t* __restrict__ count, int32_t* __restrict__ start,
                                                          int32_t* __restrict__ item_start, int nslots, int ppi,
                                                          const int32_t* __restrict__ sums, int32_t* __restrict__ counters) {
    __shared__ int s_wp[32], s_wi[32], s_base[2];
    const int t = threadIdx.x;
    const int sh = (ppi & (ppi - 1)) == 0 ? 31 - __clz(ppi) : -1;
    // totals of the tiles before this one
    int bp = 0, bi = 0;
    for (int j = t; j < (int)blockIdx.x; j += 1024) { bp += sums[2 * j]; bi += sums[2 * j + 1]; }
    {
        int dp, di;
        scan_block(bp, bi, dp, di, s_wp, s_wi);   // only the totals are used
        if (t == 0) { s_base[0] = dp; s_base[1] = di; }
        __syncthreads();
    }
    const int base_p = s_base[0], base_i = s_base[1];
    const int i0 = blockIdx.x * kScanTile + 4 * t;
    int c[4], m[4];
    scan_tile_load(count, nslots, i0, ppi, sh, c, m);
    const int tp0 = c[0] + c[1] + c[2] + c[3], ti0 = m[0] + m[1] + m[2] + m[3];
    int sp_ = tp0, si_ = ti0, tot_p, tot_i;
    __syncthreads();   // s_wp / s_wi are reused
    scan_block(sp_, si_, tot_p, tot_i, s_wp, s_wi);
    int rp = base_p + sp_ - tp0, ri = base_i + si_ - ti0;
    if (i0 + 3 < nslots) {
        int4 ps, is;
        ps.x = rp; is.x = ri; rp += c[0]; ri += m[0];
        ps.y = rp; is.y = ri; rp += c[1]; ri += m[1];
        ps.z = rp; is.z = ri; rp += c[2]; ri += m[2];
        ps.w = rp; is.w = ri;
        *reinterpret_cast<int4*>(start + i0) = ps;
        *reinterpret_cast<int4*>(item_start + i0) = is;
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (i0 + k < nslots) { start[i0 + k] = rp; item_start[i0 + k] = ri; }
            rp += c[k]; ri += m[k];
        }
    }
    if (blockIdx.x == gridDim.x - 1 && t == 0) {
        start[nslots] = base_p + tot_p;
        counters[0] = base_i + tot_i;  // number of work items
        counters[1] = 0;               // dynamic scheduler cursor
        counters[2] = base_p + tot_p;  // number of in-domain points
    }
}

// pass 3: work items of every cell; afterwards cell_cursor holds the scatter cursor (the cell's first place).
// One thread per cell LOADS the cell's descriptor, but the 64-byte item records are WRITTEN by groups of four lanes (one
// int4 each), eight records per warp-wide store: a thread writing its own records issued 16-byte stores 64 bytes apart
// (32 sectors per instruction) and the level-0 cells -- 16 items each, all in the first CTAs -- made that 17 us on C3.
__global__ void __launch_bounds__(256) k_plan_items(const __grid_constant__ thb_space_desc sp, const int32_t* __restrict__ count,
                                                    const int32_t* __restrict__ start, int32_t* __restrict__ item_start_then_cursor,
                                                    int nslots, int ppi, int max_items, int4* __restrict__ items) {
    __shared__ int4 s_rec[8][32][4];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    // The 32 cells of a warp lie nslots / 32 apart: cells with many work items (the coarse ones: 16 items each on config 3)
    // sit next to each other in slot order, and a warp that owned 32 of them ran 3x longer than the kernel needs (the items
    // of a cell beyond its first are written one cell after the other).
    const int per_lane = (nslots + 31) >> 5;
    const int gw = blockIdx.x * (blockDim.x >> 5) + wid;
    const int s = gw < per_lane ? lane * per_lane + gw : nslots;
    const int sh = (ppi & (ppi - 1)) == 0 ? 31 - __clz(ppi) : -1;
    int c = 0, b = 0, it0 = 0, m = 0;
    if (s < nslots) {
        c = count[s]; b = start[s]; it0 = item_start_then_cursor[s];
        item_start_then_cursor[s] = b;  // from here on the scatter cursor of the cell: next free place
    }
    if (c > 0) {   // cells without points (most cells, on a rank that holds a slab of the domain) read no descriptor
        const int4* ci = reinterpret_cast<const int4*>(sp.cellinfo + (int64_t)s * THB_INFO_W);
        const int4 a = __ldg(ci), e = __ldg(ci + 1), f = __ldg(ci + 2);  // level c0 c1 c2 | supp_off S tile_off n_direct | sep_off n_sep full_off 0
        const unsigned long long dm = __ldg(sp.dmask + s);
        m = sh >= 0 ? (c + ppi - 1) >> sh : (c + ppi - 1) / ppi;
        s_rec[wid][lane][0] = make_int4(s, b, min(ppi, c), a.x);
        s_rec[wid][lane][1] = make_int4(a.y, a.z, a.w, e.x);
        s_rec[wid][lane][2] = make_int4(e.y, e.z, e.w, f.y);
        s_rec[wid][lane][3] = make_int4((int)(unsigned)(dm & 0xffffffffull), (int)(unsigned)(dm >> 32), f.x, f.z);
    }
    __syncwarp();
    const int part = lane & 3, grp = lane >> 2;
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) {   // first item of eight cells per pass
        const int l = pass * 8 + grp;
        const int ml = __shfl_sync(0xffffffffu, m, l), itl = __shfl_sync(0xffffffffu, it0, l);
        if (ml > 0 && itl < max_items) items[(int64_t)itl * 4 + part] = s_rec[wid][l][part];
    }
    unsigned multi = __ballot_sync(0xffffffffu, m > 1);   // further items of the cells that have more than one
    while (multi) {
        const int l = __ffs(multi) - 1;
        multi &= multi - 1;
        const int ml = __shfl_sync(0xffffffffu, m, l), itl = __shfl_sync(0xffffffffu, it0, l);
        const int bl = __shfl_sync(0xffffffffu, b, l), cl = __shfl_sync(0xffffffffu, c, l);
        int4 v = s_rec[wid][l][part];
        for (int j = 1 + grp; j < ml; j += 8) {
            if (part == 0) { v.y = bl + j * ppi; v.z = min(ppi, cl - j * ppi); }
            if (itl + j < max_items) items[(int64_t)(itl + j) * 4 + part] = v;
        }
    }
}

// pass 4: scatter the points into cell order as 16-byte records (u0, u1, u2 or 0, caller index as bits) -- one
// STG.128 per point here, one 16-byte cp.async per point in the evaluation kernels.  Lanes of one warp that share a
// slot take consecutive places, which keeps runs of neighbouring points contiguous.  The cursor of a cell starts at
// the cell's first place (k_plan_items), so one atomic returns the destination.
#ifndef THB_SCATTER_PTS
#define THB_SCATTER_PTS 2
#endif
constexpr int kScatterPts = THB_SCATTER_PTS;
__global__ void __launch_bounds__(256) k_plan_scatter(const int32_t* __restrict__ slot, const float* __restrict__ params, int ndim,
                                                      int64_t n, int32_t* __restrict__ cursor, float4* __restrict__ sorted) {
    constexpr int H = kScatterPts;
    const int lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t nround = (n + H * stride - 1) / (H * stride) * (H * stride);
    for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < nround; i0 += H * stride) {
        int s[H], base[H], leader[H];
        unsigned peers[H];
        float u[H][3];
#pragma unroll
        for (int h = 0; h < H; ++h) {  // all loads of the points first
            const int64_t i = i0 + h * stride;
            const int64_t j = i < n ? i : 0;
            s[h] = i < n ? slot[j] : -1;
            u[h][0] = params[j * ndim];
            u[h][1] = params[j * ndim + 1];
            u[h][2] = ndim == 3 ? params[j * ndim + 2] : 0.f;
        }
#pragma unroll
        for (int h = 0; h < H; ++h) {
            peers[h] = __match_any_sync(0xffffffffu, s[h]);
            leader[h] = __ffs(peers[h]) - 1;
            base[h] = 0;
            if (s[h] >= 0 && lane == leader[h]) base[h] = atomicAdd(cursor + s[h], __popc(peers[h]));
        }
#pragma unroll
        for (int h = 0; h < H; ++h) {
            base[h] = __shfl_sync(0xffffffffu, base[h], leader[h]);
            if (s[h] >= 0) {
                const int64_t pos = (int64_t)base[h] + __popc(peers[h] & ((1u << lane) - 1u));
                sorted[pos] = make_float4(u[h][0], u[h][1], u[h][2], __int_as_float((int32_t)(i0 + h * stride)));
            }
        }
    }
}

// pass 4, LIGHT plan: the same placement, but only the caller index travels (4 bytes written and the slot read per point,
// instead of 16 written and slot + parameters read): the evaluation kernel gathers the parameters through the permutation.
__global__ void __launch_bounds__(256) k_plan_scatter_light(const int32_t* __restrict__ slot, int64_t n, int32_t* __restrict__ cursor,
                                                            int32_t* __restrict__ perm) {
    constexpr int H = 4;
    const int lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t nround = (n + H * stride - 1) / (H * stride) * (H * stride);
    for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < nround; i0 += H * stride) {
        int s[H], base[H], leader[H];
        unsigned peers[H];
#pragma unroll
        for (int h = 0; h < H; ++h) {
            const int64_t i = i0 + h * stride;
            s[h] = i < n ? slot[i] : -1;
        }
#pragma unroll
        for (int h = 0; h < H; ++h) {
            peers[h] = __match_any_sync(0xffffffffu, s[h]);
            leader[h] = __ffs(peers[h]) - 1;
            base[h] = 0;
            if (s[h] >= 0 && lane == leader[h]) base[h] = atomicAdd(cursor + s[h], __popc(peers[h]));
        }
#pragma unroll
        for (int h = 0; h < H; ++h) {
            base[h] = __shfl_sync(0xffffffffu, base[h], leader[h]);
            if (s[h] >= 0) perm[(int64_t)base[h] + __popc(peers[h] & ((1u << lane) - 1u))] = (int32_t)(i0 + h * stride);
        }
    }
}

// records of a light plan: sorted[j] = (params[perm[j]], perm[j]) for the in-domain points (counters[2] of them)
__global__ void __launch_bounds__(256) k_plan_records(const int32_t* __restrict__ perm, const float* __restrict__ params, int ndim,
                                                      const int32_t* __restrict__ counters, float4* __restrict__ sorted) {
    const int64_t n_in = counters[2];
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_in; j += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = perm[j];
        sorted[j] = make_float4(params[i * ndim], params[i * ndim + 1], ndim == 3 ? params[i * ndim + 2] : 0.f, __int_as_float((int32_t)i));
    }
}

// compute_tensor_product (THB/bspline_funcs.py:54-59): out[b][i][j][k] = a[b][i] * bb[b][j] * c[b][k] for a batch of vector
// triples (nc == 1 and c == nullptr for the 2-D form); one thread per output entry, coalesced stores
__global__ void __launch_bounds__(256) k_tensor_product(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c,
                                                        int64_t batch, int na, int nb, int nc, float* __restrict__ out) {
    const int64_t per = (int64_t)na * nb * nc, tot = batch * per;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t q = e / per;
        const int r = (int)(e - q * per);
        const int k = r % nc, j = (r / nc) % nb, i = r / (nc * nb);
        float v = __ldg(a + q * na + i) * __ldg(b + q * nb + j);
        if (c) v *= __ldg(c + q * nc + k);
        out[e] = v;
    }
}

int grid_for(int64_t n, int block, int per_sm) {
    int64_t want = (n + block - 1) / block;
    int64_t cap = (int64_t)thb_num_sms() * per_sm;
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}

// CTAs per SM of the two passes over the points (grid-stride loops).  Fewer than the 8 that fit leave room for a
// concurrent kernel on another stream (the local control nets do not depend on the plan): THB_PLAN_CTAS_PER_SM.
int plan_ctas_per_sm() {
    static int v = 0;
    if (v == 0) {
        const char* e = getenv("THB_PLAN_CTAS_PER_SM");
        v = e ? atoi(e) : 16;
        if (v < 1 || v > 32) v = 16;
    }
    return v;
}

// THB_COUNT_VEC=0 selects the scalar histogram kernel (k_plan_count_fast) where the vector form would run
bool count_vec_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("THB_COUNT_VEC");
        v = (e && atoi(e) == 0) ? 0 : 1;
    }
    return v != 0;
}

// the scatter pass alone (the nets can run beside it on another stream: it is HBM-bound, they are issue / latency-bound)
int scatter_ctas_per_sm() {
    static int v = 0;
    if (v == 0) {
        const char* e = getenv("THB_SCATTER_CTAS_PER_SM");
        v = e ? atoi(e) : plan_ctas_per_sm();
        if (v < 1 || v > 32) v = plan_ctas_per_sm();
    }
    return v;
}

}  // namespace

int thb_check_space(const thb_space_desc* sp, const char* who) {
    THB_REQUIRE(sp, "%s: null space descriptor", who);
    THB_REQUIRE(sp->ndim == 2 || sp->ndim == 3, "%s: ndim must be 2 or 3", who);
    THB_REQUIRE(sp->nlevels >= 1 && sp->nlevels <= THB_MAX_LEVELS, "%s: bad level count", who);
    THB_REQUIRE(sp->nslots > 0, "%s: no active cells", who);
    THB_REQUIRE(sp->knots && sp->cell_slot && sp->cellinfo && sp->dmask && sp->supp_jm && sp->tiles, "%s: null table", who);
    return THB_OK;
}

extern "C" int thb_find_spans(const float* params, int64_t stride, int64_t npoints, const float* knots, int nknots,
                              int degree, int32_t* spans, void* stream) {
    THB_REQUIRE(npoints >= 0 && degree >= 0 && nknots >= 2 * (degree + 1), "thb_find_spans: bad sizes");
    if (npoints == 0) return THB_OK;
    THB_REQUIRE(params && knots && spans && stride >= 1, "thb_find_spans: null pointer");
    k_find_spans<<<grid_for(npoints, 256, 16), 256, 0, (cudaStream_t)stream>>>(params, stride, npoints, knots, nknots, degree, spans);
    THB_LAUNCHED();
    return THB_OK;
}

extern "C" int thb_basis_1d(const float* params, int64_t stride, int64_t npoints, const float* knots, int nknots, int degree,
                            float* values, float* derivs, void* stream) {
    THB_REQUIRE(npoints >= 0 && degree >= 0 && degree <= THB_MAX_DEG && nknots >= 2 * (degree + 1), "thb_basis_1d: bad sizes");
    if (npoints == 0) return THB_OK;
    THB_REQUIRE(params && knots && values && stride >= 1, "thb_basis_1d: null pointer");
    k_basis_1d<<<grid_for(npoints, 256, 16), 256, 0, (cudaStream_t)stream>>>(params, stride, npoints, knots, nknots, degree, values, derivs);
    THB_LAUNCHED();
    return THB_OK;
}

extern "C" int thb_tensor_product(const float* a, const float* b, const float* c, int64_t batch, int na, int nb, int nc, float* out,
                                  void* stream) {
    THB_REQUIRE(batch >= 0 && na >= 1 && nb >= 1 && nc >= 1, "thb_tensor_product: bad sizes");
    THB_REQUIRE(c || nc == 1, "thb_tensor_product: nc > 1 needs the third factor");
    if (batch == 0) return THB_OK;
    THB_REQUIRE(a && b && out, "thb_tensor_product: null pointer");
    k_tensor_product<<<grid_for(batch * na * nb * nc, 256, 16), 256, 0, (cudaStream_t)stream>>>(a, b, c, batch, na, nb, nc, out);
    THB_LAUNCHED();
    return THB_OK;
}

extern "C" int thb_active_span(const thb_space_desc* space, const float* params, int64_t npoints, int32_t* slot,
                               int32_t* num_supp, void* stream) {
    int rc = thb_check_space(space, "thb_active_span");
    if (rc) return rc;
    THB_REQUIRE(npoints >= 0, "thb_active_span: negative point count");
    if (npoints == 0) return THB_OK;
    THB_REQUIRE(params && slot, "thb_active_span: null pointer");
    k_active_span<<<grid_for(npoints, 256, 16), 256, 0, (cudaStream_t)stream>>>(*space, params, npoints, slot, num_supp);
    THB_LAUNCHED();
    return THB_OK;
}

extern "C" int thb_expand_supports(const thb_space_desc* space, const int32_t* slot, const int64_t* offsets, int64_t npoints,
                                   void* jm_out, int jm_is_i64, void* stream) {
    int rc = thb_check_space(space, "thb_expand_supports");
    if (rc) return rc;
    if (npoints <= 0) return THB_OK;
    THB_REQUIRE(slot && offsets && jm_out, "thb_expand_supports: null pointer");
    const int grid = grid_for(npoints * 32, 256, 16);
    if (jm_is_i64)
        k_expand_supports<long long><<<grid, 256, 0, (cudaStream_t)stream>>>(*space, slot, offsets, npoints, (long long*)jm_out);
    else
        k_expand_supports<int><<<grid, 256, 0, (cudaStream_t)stream>>>(*space, slot, offsets, npoints, (int*)jm_out);
    THB_LAUNCHED();
    return THB_OK;
}

// phases (THB_PLAN_*): histogram (slot per point + points per cell: all the nets of the touched cells need), scan + work
// items, scatter of the points into cell order
static int plan_build_impl(const thb_space_desc* space, const float* params, const thb_plan_desc* plan, void* stream, int phases) {
    int rc = thb_check_space(space, "thb_plan_build");
    if (rc) return rc;
    THB_REQUIRE(plan && plan->npoints >= 0 && plan->npoints < (int64_t)2147483647, "thb_plan_build: bad point count");
    THB_REQUIRE(plan->points_per_item >= 32 && plan->points_per_item <= 1024, "thb_plan_build: points_per_item out of range");
    THB_REQUIRE(plan->slot && plan->scratch && plan->sorted_params && plan->cell_count && plan->cell_start && plan->cell_cursor && plan->items && plan->counters,
                "thb_plan_build: null plan buffer");
    THB_REQUIRE((int64_t)plan->max_items >= plan->npoints / plan->points_per_item + space->nslots, "thb_plan_build: max_items too small");
    THB_REQUIRE((((uintptr_t)plan->cell_count | (uintptr_t)plan->cell_start | (uintptr_t)plan->cell_cursor) & 15) == 0,
                "thb_plan_build: cell_count / cell_start / cell_cursor must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n = plan->npoints;
    const int ns = space->nslots;
    THB_REQUIRE(n == 0 || params, "thb_plan_build: null params");
    if (phases & THB_PLAN_HISTOGRAM) {
        THB_CUDA(cudaMemsetAsync(plan->cell_count, 0, sizeof(int32_t) * (ns + 1), st));
        if (n > 0) {
            int finest_knots = 0;
            for (int k = 0; k < space->ndim; ++k) finest_knots += space->knot_len[space->nlevels - 1][k];
            if (space->span_lut && space->finest_slot && finest_knots <= kPlanSmemKnots) {
                const int lut_tot = space->lut_off[space->ndim - 1] + space->lut_n[space->ndim - 1];
                const int g = grid_for(n, 256, plan_ctas_per_sm());
                const bool exact = space->lut_exact != 0;   // spans proven by the host: no verification, no fix-up pass
                // (small batches -- a rank's slab, a chunk of the host pipeline -- keep the scalar kernel, whose twice as many
                // threads hide the latency of a single iteration better: 19 against 25 us for 1.4 M points, with or without
                // the staging prologue of the vector kernel; one wave of 6 resident CTAs per SM for the vector kernel was
                // measured too: 0.220 instead of 0.204 ms for the plan of config 3)
                if (exact && lut_tot <= kPlanSmemLut && count_vec_enabled() && n >= kCountVecMin &&
                    ((((uintptr_t)params) | ((uintptr_t)plan->slot)) & 15) == 0) {
                    const int gv = grid_for((n + 3) / 4, 256, plan_ctas_per_sm());   // >= 1 CTA: the tail runs on the first threads
                    if (space->ndim == 3) k_plan_count_vec<3><<<gv, 256, 0, st>>>(*space, params, (unsigned)n, plan->slot, plan->cell_count);
                    else k_plan_count_vec<2><<<gv, 256, 0, st>>>(*space, params, (unsigned)n, plan->slot, plan->cell_count);
                    THB_LAUNCHED();
                } else {
                    // plan->scratch holds the (normally empty) list of unverified points, counters[3] their number
                    THB_CUDA(cudaMemsetAsync(plan->counters + 3, 0, sizeof(int32_t), st));
#define THB_COUNT(D_, S_)                                                                                                        \
    do {                                                                                                                        \
        if (exact) k_plan_count_fast<D_, S_, true><<<g, 256, 0, st>>>(*space, params, (unsigned)n, plan->slot, plan->cell_count, plan->scratch, plan->counters); \
        else k_plan_count_fast<D_, S_, false><<<g, 256, 0, st>>>(*space, params, (unsigned)n, plan->slot, plan->cell_count, plan->scratch, plan->counters); \
    } while (0)
                    if (space->ndim == 3) { if (lut_tot <= kPlanSmemLut) THB_COUNT(3, true); else THB_COUNT(3, false); }
                    else { if (lut_tot <= kPlanSmemLut) THB_COUNT(2, true); else THB_COUNT(2, false); }
#undef THB_COUNT
                    THB_LAUNCHED();
                    if (!exact) {
                        k_plan_fixup<<<thb_num_sms(), 256, 0, st>>>(*space, params, plan->scratch, plan->counters, plan->slot, plan->cell_count);
                        THB_LAUNCHED();
                    }
                }
            } else {
                k_plan_count<<<grid_for(n, 256, 16), 256, 0, st>>>(*space, params, n, plan->slot, plan->cell_count);
                THB_LAUNCHED();
            }
        }
    }
    if (phases & THB_PLAN_ITEMS) {
        const int ntile = (ns + kScanTile - 1) / kScanTile;
        if (ntile >= 4 && n >= 2 * (int64_t)ntile && (((uintptr_t)plan->scratch) & 3) == 0) {
            // plan->scratch is free between the fix-up pass and the end of the plan: it holds the per-tile totals
            k_plan_scan_sums<<<ntile, 1024, 0, st>>>(plan->cell_count, ns, plan->points_per_item, plan->scratch);
            THB_LAUNCHED();
            k_plan_scan_write<<<ntile, 1024, 0, st>>>(plan->cell_count, plan->cell_start, plan->cell_cursor, ns, plan->points_per_item,
                                                      plan->scratch, plan->counters);
            THB_LAUNCHED();
        } else {
            k_plan_scan<<<1, 1024, 0, st>>>(plan->cell_count, plan->cell_start, plan->cell_cursor, ns, plan->points_per_item, plan->counters);
            THB_LAUNCHED();
        }
        k_plan_items<<<(ns + 255) / 256, 256, 0, st>>>(*space, plan->cell_count, plan->cell_start, plan->cell_cursor, ns, plan->points_per_item,
                                                       plan->max_items, (int4*)plan->items);
        THB_LAUNCHED();
    }
    if ((phases & THB_PLAN_SCATTER) && n > 0) {
        if (plan->light) {   // permutation only; plan->scratch is free from here on
            k_plan_scatter_light<<<grid_for(n, 256, scatter_ctas_per_sm()), 256, 0, st>>>(plan->slot, n, plan->cell_cursor, plan->scratch);
        } else {
            k_plan_scatter<<<grid_for(n, 256, scatter_ctas_per_sm()), 256, 0, st>>>(plan->slot, params, space->ndim, n, plan->cell_cursor,
                                                                                     (float4*)plan->sorted_params);
        }
        THB_LAUNCHED();
    }
    return THB_OK;
}

extern "C" int thb_plan_build(const thb_space_desc* space, const float* params, const thb_plan_desc* plan, void* stream) {
    return plan_build_impl(space, params, plan, stream, THB_PLAN_HISTOGRAM | THB_PLAN_ITEMS | THB_PLAN_SCATTER);
}

extern "C" int thb_plan_build_phase(const thb_space_desc* space, const float* params, const thb_plan_desc* plan, int phases,
                                    void* stream) {
    THB_REQUIRE(phases >= 1 && phases <= (THB_PLAN_HISTOGRAM | THB_PLAN_ITEMS | THB_PLAN_SCATTER),
                "thb_plan_build_phase: phases must be a combination of THB_PLAN_HISTOGRAM, THB_PLAN_ITEMS, THB_PLAN_SCATTER");
    return plan_build_impl(space, params, plan, stream, phases);
}

extern "C" int thb_plan_records(const thb_space_desc* space, const float* params, const thb_plan_desc* plan, void* stream) {
    int rc = thb_check_space(space, "thb_plan_records");
    if (rc) return rc;
    THB_REQUIRE(plan && plan->scratch && plan->sorted_params && plan->counters, "thb_plan_records: null plan buffer");
    THB_REQUIRE(plan->light, "thb_plan_records: the plan already holds its records");
    if (plan->npoints == 0) return THB_OK;
    THB_REQUIRE(params, "thb_plan_records: null params");
    k_plan_records<<<grid_for(plan->npoints, 256, 16), 256, 0, (cudaStream_t)stream>>>(plan->scratch, params, space->ndim, plan->counters,
                                                                                       (float4*)plan->sorted_params);
    THB_LAUNCHED();
    return THB_OK;
}
