// Tile kernels: THB_basis_fns with PHI materialised (THB/core.py:604-701; float64 statement core.py:232-268) and the
// per-point formulations of the fused THB_basis_fns + THBEval.forward / backward (THB/torch_funcs.py:22-47).  The
// default fused evaluation (local control nets) is in thb_net_impl.cuh and reuses the pieces defined here.
//
// Work decomposition:
//   * the points of a batch are grouped by active cell (thb_plan_build); a work item = one cell and up to
//     points_per_item of its points; persistent one-warp CTAs pull 64-byte item records from an atomic cursor
//   * per item the warp stages in shared memory (cp.async, double buffered): the 2p local knots per dimension, the
//     control points of the cell's own-level window (SoA, zero where inactive), the factor vectors of the rank-one
//     coarser-level supports and -- in chunks of kWChunk supports -- the (p+1)^d coefficient tiles of the others, plus
//     the supports' control points
//   * each lane owns PPT points: register-resident Cox-de Boor (values + first derivatives), the
//     (p+1)^d tensor product kept as packed f32x2 register pairs, then
//       PHI_s = <tile_s, tp>  with broadcast LDS.128 of the tile and FFMA2          (gather-multiply)
//       out  += PHI_s * ctrl_pts[Jm_s]                                              (contraction)
//     own-level supports have one-hot tiles: their PHI is tp[t] itself, so their contraction is three
//     dot products of tp with the staged window control points.
//   * MODE_BASIS writes PHI (THB_basis_fns); MODE_FUSED writes only the evaluated geometry.
//   * cellnet (formulation tile_net): tiles are first folded into the window control points (a per-item local control
//     net Q[t] = CPwin[t] + sum_s tile_s[t] * CP_s), after which every point costs 3*(p+1)^d FMAs.
//   * k_tile_backward (CTAs of 128 threads): the per-point formulation of the fused backward.
#pragma once
#include "thb_common.cuh"

namespace thb_tile {

#ifndef THB_TILE_MINB
#define THB_TILE_MINB 1  // register budget hint (9 => 168 registers + spills in the FFMA2 loop: measured slower)
#endif
constexpr int kThreads = 32;   // forward tile kernels: one warp per CTA
constexpr int kBwdThreads = 128;
constexpr int kChunk = 64;

// (the default fused evaluation -- local control nets -- lives in thb_net_impl.cuh; MODE_FUSED here is the per-point
// formulation, the reference's gather-multiply + contraction without PHI leaving the SM)
enum { MODE_FUSED = 0, MODE_BASIS = 1 };

struct Args {
    const int4* items;
    int32_t* counters;  // [0] = number of items, [1] = scheduler cursor
    const float* params;
    const float* sorted;     // plan->sorted_params: [N][4] records in cell order (u0, u1, u2 or 0, caller index bits)
    const float* ctrl_pts;   // fused
    const int64_t* offsets;  // basis: exclusive prefix sum of num_supp (caller order)
    float* out[4];           // per channel: fused [N, cp_dim] / basis [nnz]
    int32_t channels[4];
    int32_t nch;
    int32_t cellnet;
    int32_t rows_sorted;     // rows of out / grad_out are in plan order (thb_plan_desc::rows_in_plan_order)
    // backward
    const float* grad_out;
    float* grad_cp;
    float* item_ws;          // ordered (deterministic) backward: per-ITEM net gradients [nitems][4][TS], plain stores
    const int32_t* perm;     // LIGHT plan (k_net_eval): perm[j] = caller index of the point in place j; nullptr = records in `sorted`
    int32_t ndim;            // ... and the width of `params` the points are gathered from
};

template <int N0_, int N1_, int N2_>
struct Shp {
    static constexpr int N0 = N0_, N1 = N1_, N2 = N2_;
    static constexpr int T = N0 * N1 * N2;
    static constexpr int TS = (T + 3) / 4 * 4;
    static constexpr int NP = TS / 2;  // f32x2 pairs
    static constexpr int NQ = TS / 4;  // 16-byte quads
    static constexpr int P0 = N0 - 1, P1 = N1 - 1, P2 = N2 - 1;
};

struct Item {
    int slot, begin, count;
    int lev, c0, c1, c2, soff, S, toff, ndir;
    int nsep, sep_off, full_off;  // split of the tiled supports: rank-one (12 floats) / full tiles
    unsigned long long dm;
};

// Work-item record written by thb_plan_build: 16 ints = slot, begin, count, level, c0, c1, c2, supp_off,
// S, tile_off, n_direct, n_sep, dmask lo, dmask hi, sep_off, full_off  (the cell's row of cellinfo/dmask is copied in, so one
// 64-byte read describes the item).
THB_DEV Item unpack_item(const int4 v0, const int4 v1, const int4 v2, const int4 v3) {
    Item r;
    r.slot = v0.x; r.begin = v0.y; r.count = v0.z; r.lev = v0.w;
    r.c0 = v1.x; r.c1 = v1.y; r.c2 = v1.z; r.soff = v1.w;
    r.S = v2.x; r.toff = v2.y; r.ndir = v2.z; r.nsep = v2.w;
    r.sep_off = v3.z; r.full_off = v3.w;
    r.dm = (unsigned long long)(unsigned)v3.x | ((unsigned long long)(unsigned)v3.y << 32);
    return r;
}
THB_DEV Item load_item(const thb_space_desc& sp, const int4* items, int it) {
    const int4* p = items + (int64_t)it * 4;
    return unpack_item(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(p + 3));
}

template <class SH>
THB_DEV void stage_knots(const thb_space_desc& sp, const Item& it, float (*s_knots)[8]) {
    const int tid = threadIdx.x;
    if (tid < 24) {
        const int dim = tid >> 3, m = tid & 7;
        const int P = dim == 0 ? SH::P0 : (dim == 1 ? SH::P1 : SH::P2);
        const int c = dim == 0 ? it.c0 : (dim == 1 ? it.c1 : it.c2);
        if (dim < sp.ndim && m < 2 * P) s_knots[dim][m] = __ldg(sp.knots + sp.knot_off[it.lev][dim] + c + 1 + m);
    }
}

// 1-D bases of one point in the item's cell
template <class SH, bool DER>
struct Bases {
    float n0[SH::N0], n1[SH::N1], n2[SH::N2];
    float d0[DER ? SH::N0 : 1], d1[DER ? SH::N1 : 1], d2[DER ? SH::N2 : 1];
    THB_DEV void eval(const float* u, const float (*s_knots)[8]) {
        cox_de_boor<SH::P0, DER>(u[0], s_knots[0], n0, d0);
        cox_de_boor<SH::P1, DER>(u[1], s_knots[1], n1, d1);
        if (SH::N2 > 1) cox_de_boor<SH::P2, DER>(u[2], s_knots[2], n2, d2);
        else { n2[0] = 1.0f; if (DER) d2[0] = 0.0f; }
    }
    // tensor product for a channel (bit k = derivative in dim k), C order (last dim fastest)
    THB_DEV void tensor(int ch, float* tp) const {
        float a[SH::N0], b[SH::N1], c[SH::N2];
#pragma unroll
        for (int i = 0; i < SH::N0; ++i) a[i] = (DER && (ch & 1)) ? d0[i] : n0[i];
#pragma unroll
        for (int i = 0; i < SH::N1; ++i) b[i] = (DER && (ch & 2)) ? d1[i] : n1[i];
#pragma unroll
        for (int i = 0; i < SH::N2; ++i) c[i] = (DER && (ch & 4)) ? d2[i] : n2[i];
#pragma unroll
        for (int i = 0; i < SH::N0; ++i)
#pragma unroll
            for (int j = 0; j < SH::N1; ++j) {
                const float ab = a[i] * b[j];
#pragma unroll
                for (int k = 0; k < SH::N2; ++k) tp[(i * SH::N1 + j) * SH::N2 + k] = ab * c[k];
            }
#pragma unroll
        for (int t = SH::T; t < SH::TS; ++t) tp[t] = 0.0f;
    }
};

// Reciprocal knot differences of one cell: den(j, r) = K[P+r] - K[P-j+r], stored at j(j-1)/2 + r.
// They do not depend on the point, so the Cox-de Boor recursion of every point of the item multiplies by
// these instead of dividing (one division per entry per ITEM instead of p(p+1)/2 per dimension per POINT).
template <int P, bool DER>
THB_DEV void cox_de_boor_rk(float u, const float* K, const float* RK, float* N, float* dN) {
    float left[P + 1], right[P + 1];
    N[0] = 1.0f;
#pragma unroll
    for (int j = 1; j <= P; ++j) {
        left[j] = u - K[P - j];
        right[j] = K[P - 1 + j] - u;
        float carry = 0.0f, prev_t = 0.0f;
#pragma unroll
        for (int r = 0; r < j; ++r) {
            const float t = N[r] * RK[j * (j - 1) / 2 + r];
            if (DER && j == P) {
                dN[r] = float(P) * (prev_t - t);
                prev_t = t;
            }
            N[r] = fmaf(right[r + 1], t, carry);
            carry = left[j - r] * t;
        }
        if (DER && j == P) dN[P] = float(P) * prev_t;
        N[j] = carry;
    }
    if (DER && P == 0) dN[0] = 0.0f;
}

template <class SH, bool DER>
struct BasesRK {
    float n0[SH::N0], n1[SH::N1], n2[SH::N2];
    float d0[DER ? SH::N0 : 1], d1[DER ? SH::N1 : 1], d2[DER ? SH::N2 : 1];
    THB_DEV void eval(const float* u, const float (*s_knots)[8], const float (*s_rk)[8]) {
        cox_de_boor_rk<SH::P0, DER>(u[0], s_knots[0], s_rk[0], n0, d0);
        cox_de_boor_rk<SH::P1, DER>(u[1], s_knots[1], s_rk[1], n1, d1);
        if (SH::N2 > 1) cox_de_boor_rk<SH::P2, DER>(u[2], s_knots[2], s_rk[2], n2, d2);
        else { n2[0] = 1.0f; if (DER) d2[0] = 0.0f; }
    }
    // 1-D factors of a channel (bit k = first derivative in dimension k)
    THB_DEV void select(int ch, float* a, float* b, float* c) const {
#pragma unroll
        for (int i = 0; i < SH::N0; ++i) a[i] = (DER && (ch & 1)) ? d0[i] : n0[i];
#pragma unroll
        for (int i = 0; i < SH::N1; ++i) b[i] = (DER && (ch & 2)) ? d1[i] : n1[i];
#pragma unroll
        for (int i = 0; i < SH::N2; ++i) c[i] = (DER && (ch & 4)) ? d2[i] : n2[i];
    }
    THB_DEV void tensor(int ch, float* tp) const {
        float a[SH::N0], b[SH::N1], c[SH::N2];
        select(ch, a, b, c);
#pragma unroll
        for (int i = 0; i < SH::N0; ++i)
#pragma unroll
            for (int j = 0; j < SH::N1; ++j) {
                const float ab = a[i] * b[j];
#pragma unroll
                for (int k = 0; k < SH::N2; ++k) tp[(i * SH::N1 + j) * SH::N2 + k] = ab * c[k];
            }
#pragma unroll
        for (int t = SH::T; t < SH::TS; ++t) tp[t] = 0.0f;
    }
};

THB_DEV float f4c(const float4& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w)); }

#ifndef THB_WCHUNK
#define THB_WCHUNK 16
#endif
#ifndef THB_SUNROLL
#define THB_SUNROLL 2
#endif
#define THB_PRAGMA_(x) _Pragma(#x)
#define THB_PRAGMA(x) THB_PRAGMA_(x)
constexpr int kWChunk = THB_WCHUNK;  // supports staged per chunk by one warp
constexpr int kMaxCC = 128;  // coarse-support control points gathered once per item (else per chunk)
constexpr int kRowPad = 4;   // floats of padding per staged tile row: rows s..s+3 start in different banks

// Per-warp shared memory.  Everything that can be known ahead of time is double buffered and filled by
// cp.async, so the global-memory latency of item i+1 is paid while item i is being computed and costs no
// registers.
template <class SH, int MODE, bool DER>
struct WarpSmem {
    int4 rec[3][4];                       // ring of work-item records (lookahead 2)
    float knots[2][3][8];                 // local knots of the item's cell, per dimension
    float rk[3][8];                       // reciprocal knot differences of the current item
    float cpw[2][4][SH::TS];              // own-level window control points (SoA, 0 where inactive)
    float4 prec[2][64];                   // point records of a sub-batch: (u0, u1, u2, caller-order index as bits)
    static constexpr int kCC = kMaxCC;
    float4 cc[kCC];                       // control points of the coarser-level supports (rank-one ones first)
    float4 sepv[MODE == MODE_FUSED ? kCC : 1][3];  // factor vectors of the rank-one tiled supports
    float tiles[2][kWChunk * (SH::TS + kRowPad)];  // coefficient tiles, streamed in chunks (padded rows)
    // derivative channels contract against control points RELATIVE to a reference point of the item: the
    // derivative basis values sum to zero, so the result is unchanged, but the summed terms are O(p/h * h)
    // instead of O(p/h * |CP|) -- this removes the float32 cancellation error of parametric derivatives
    float cpw_rel[(MODE == MODE_FUSED && DER) ? 4 : 1][(MODE == MODE_FUSED && DER) ? SH::TS : 1];
    float4 cc_rel[(MODE == MODE_FUSED && DER) ? kCC : 1];
    float ref[4];
    // MODE_BASIS: PHI values are transposed through here so that every point's row is written coalesced
    float stage[MODE == MODE_BASIS ? (32 * (SH::TS + 1) > 64 * (kWChunk + 1) ? 32 * (SH::TS + 1) : 64 * (kWChunk + 1)) : 1];
    long long poff[MODE == MODE_BASIS ? 64 : 1];
};

template <int PPT>
THB_DEV void prefetch_points(const Args& ar, int begin, int count, int base, float4* prec) {
    const int lane = threadIdx.x;
    const float4* src = reinterpret_cast<const float4*>(ar.sorted);
#pragma unroll
    for (int q = 0; q < PPT; ++q) {
        const int li = base + q * 32 + lane;
        if (q == 0 || base + q * 32 < count)   // warp-uniform: blocks beyond the item are not fetched
            cp_async16(prec + q * 32 + lane, src + begin + (li < count ? li : 0));
    }
}

// The same for a LIGHT plan: the caller indices of the item's points are already in shared memory (pidx, one item ahead);
// the parameters are gathered from the caller's array (three 4-byte cp.async per point), the index goes into .w.
template <int PPT>
THB_DEV void prefetch_points_gather(const Args& ar, const int* pidx, int count, int base, float4* prec) {
    const int lane = threadIdx.x;
#pragma unroll
    for (int q = 0; q < PPT; ++q) {
        const int li = base + q * 32 + lane;
        if (q == 0 || base + q * 32 < count) {
            const int idx = pidx[li < count ? li : 0];
            const float* src = ar.params + (int64_t)idx * ar.ndim;
            float* dst = reinterpret_cast<float*>(prec + q * 32 + lane);
            cp_async4(dst, src);
            cp_async4(dst + 1, src + 1);
            if (ar.ndim == 3) cp_async4(dst + 2, src + 2);
            else dst[2] = 0.0f;
            dst[3] = __int_as_float(idx);
        }
    }
}

// knots and own-level window control points of an item -> shared memory buffers (cp.async + zero fill).
// Flat id of window position t: fn_off[lev] + ravel(c + t) (the reference's numbering, THB/core.py:644-650),
// so no index table has to be read first.
template <class SH, int CPD>
THB_DEV void prefetch_item_tables(const thb_space_desc& sp, const Args& ar, const Item& it, float (*knots)[8], float (*cpw)[SH::TS],
                                  bool want_cp) {
    const int lane = threadIdx.x;
    if (lane < 24) {
        const int dim = lane >> 3, m = lane & 7;
        const int P = dim == 0 ? SH::P0 : (dim == 1 ? SH::P1 : SH::P2);
        const int c = dim == 0 ? it.c0 : (dim == 1 ? it.c1 : it.c2);
        if (dim < sp.ndim && m < 2 * P) cp_async4(&knots[dim][m], sp.knots + sp.knot_off[it.lev][dim] + c + 1 + m);
    }
    if (want_cp) {
        const int n1 = sp.nfns[it.lev][1], n2 = sp.nfns[it.lev][2];
        const int64_t f0 = sp.fn_off[it.lev];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int t = lane + 32 * h;
            if (t < SH::TS) {
                const bool on = it.ndir > 0 && t < SH::T && ((it.dm >> t) & 1ull);
                if (on) {
                    const int t2 = t % SH::N2, t1 = (t / SH::N2) % SH::N1, t0 = t / (SH::N2 * SH::N1);
                    const int64_t r = f0 + ((int64_t)(it.c0 + t0) * n1 + (it.c1 + t1)) * n2 + (it.c2 + t2);
#pragma unroll
                    for (int k = 0; k < CPD; ++k) cp_async4(&cpw[k][t], ar.ctrl_pts + r * CPD + k);
                    if (CPD < 4) cpw[3][t] = 0.f;
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) cpw[k][t] = 0.f;
                }
            }
        }
    }
}

template <int CPD>
THB_DEV void gather_ids(const int32_t* __restrict__ ids, const Args& ar, int n, float4* s_cc) {
    for (int s = threadIdx.x; s < n; s += 32) {
        const int64_t r = __ldg(ids + s);
        float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < CPD; ++k) v[k] = __ldg(ar.ctrl_pts + r * CPD + k);
        s_cc[s] = make_float4(v[0], v[1], v[2], v[3]);
    }
}

template <int CPD>
THB_DEV void gather_cc(const thb_space_desc& sp, const Args& ar, const Item& it, int first, int n, float4* s_cc) {
    for (int s = threadIdx.x; s < n; s += 32) {
        const int64_t r = __ldg(sp.supp_jm + it.soff + it.ndir + first + s);
        float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < CPD; ++k) v[k] = __ldg(ar.ctrl_pts + r * CPD + k);
        s_cc[s] = make_float4(v[0], v[1], v[2], v[3]);
    }
}

template <int NQ>
THB_DEV void issue_chunk(const float* tile_src, int c0, int ns, int TS, float* buf) {
    const float4* src = reinterpret_cast<const float4*>(tile_src + (int64_t)c0 * TS);
    for (int i = threadIdx.x; i < ns * NQ; i += 32) {
        const int row = i / NQ, quad = i - row * NQ;
        cp_async16(buf + row * (TS + kRowPad) + quad * 4, src + i);
    }
    cp_async_commit();
}

// One sub-batch of an item: PPT*32 points (lane-major), all channels.
// G > 1 (only with PPT == 1): "support split" for sub-batches of <= 32/G points -- G lane groups hold the same
// points and each takes every G-th tiled support, so small cells do not leave 3/4 of the lanes idle; the
// groups read G different (bank-staggered) tile rows with one LDS and their partial sums are combined by shuffles.
template <class SH, int PPT, int G, int CPD, int MODE, bool DER>
THB_DEV void warp_points(const thb_space_desc& sp, const Args& ar, const Item& it, int base, int ntile, bool cellnet,
                         WarpSmem<SH, MODE, DER>& sm, int tb, int pb) {
    static_assert(G == 1 || (PPT == 1 && MODE == MODE_FUSED), "support split: one point per lane, fused mode only");
    constexpr int T = SH::T, TS = SH::TS, NP = SH::NP, NQ = SH::NQ;
    constexpr int NPT = 32 / G;                // points per pass
    constexpr int TSP = TS + kRowPad;          // staged tile row stride
    const int lane_raw = threadIdx.x;
    const int grp = lane_raw / NPT;            // support group of this lane
    const int lane = lane_raw % NPT;           // point slot of this lane
    const bool use_direct = (it.ndir > 0) || (MODE == MODE_FUSED && cellnet);
    // fused per-point mode with split tables: rank-one tiled supports are evaluated from their three factor vectors
    // (12 FMAs + 2 multiplications instead of a (p+1)^d dot product), only the remaining ones stream full tiles
    const bool split = MODE == MODE_FUSED && !cellnet && sp.sep_vec != nullptr && ntile <= kMaxCC;
    const int nsep = split ? it.nsep : 0;
    const int nfl = split ? ntile - it.nsep : ntile;                                  // supports that need a full tile
    const float* tile_src = split ? sp.full_tiles + (int64_t)it.full_off * TS : sp.tiles + (int64_t)it.toff * TS;
    const bool tile_loop = !(MODE == MODE_FUSED && cellnet) && nfl > 0;
    const bool cc_all = ntile <= kMaxCC;
    if (tile_loop) issue_chunk<NQ>(tile_src, 0, min(kWChunk, nfl), TS, sm.tiles[0]);  // overlaps the basis evaluation below

    int idx[PPT];
    bool valid[PPT];
    int64_t off[PPT];
    BasesRK<SH, DER> bs[PPT];
#pragma unroll
    for (int q = 0; q < PPT; ++q) {
        valid[q] = base + q * 32 + lane < it.count;
        const float4 rec = sm.prec[pb][q * 32 + lane];
        idx[q] = __float_as_int(rec.w);
        if (MODE == MODE_BASIS) off[q] = __ldg(ar.offsets + idx[q]);
        const float u[3] = {rec.x, rec.y, rec.z};
        bs[q].eval(u, sm.knots[tb], sm.rk);
        if (MODE == MODE_BASIS) sm.poff[q * 32 + lane] = off[q];
    }
    constexpr int LDS_ = TS + 1;  // padded row of the PHI staging area
    int nval[PPT];
#pragma unroll
    for (int q = 0; q < PPT; ++q) nval[q] = min(32, max(0, it.count - base - q * 32));
    // positions of this lane's window entries (t = lane, lane + 32) among the cell's own-level supports
    int dpos[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int t = lane + 32 * h;
        dpos[h] = (it.ndir > 0 && t < T && ((it.dm >> t) & 1ull)) ? __popcll(it.dm & ((1ull << t) - 1ull)) : -1;
    }
    if (MODE == MODE_BASIS) __syncwarp();
    for (int chi = 0; chi < ar.nch; ++chi) {
        const int ch = ar.channels[chi];
        const bool use_rel = DER && MODE == MODE_FUSED && ch != 0 && !cellnet && cc_all;  // see WarpSmem::cpw_rel
        f32x2 tp2[PPT][NP];
#pragma unroll
        for (int q = 0; q < PPT; ++q) {
            float tp[TS];
            bs[q].tensor(ch, tp);
            if (MODE == MODE_BASIS && it.ndir > 0) {
                // own-level supports: PHI = tp[t].  Transpose through shared memory: lane = point on the way in,
                // lane = window entry on the way out, so each point's row goes out as one coalesced store.
#pragma unroll
                for (int t = 0; t < T; ++t) sm.stage[lane * LDS_ + t] = tp[t];
                __syncwarp();
                for (int pt = 0; pt < nval[q]; ++pt) {
                    float* dst = ar.out[chi] + sm.poff[q * 32 + pt];
#pragma unroll
                    for (int h = 0; h < 2; ++h)
                        if (dpos[h] >= 0) dst[dpos[h]] = sm.stage[pt * LDS_ + lane + 32 * h];
                }
                __syncwarp();
            }
#pragma unroll
            for (int m = 0; m < NP; ++m) tp2[q][m] = pack2(tp[2 * m], tp[2 * m + 1]);
        }
        float acc[PPT][4];
#pragma unroll
        for (int q = 0; q < PPT; ++q)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[q][k] = 0.0f;

        if (MODE == MODE_FUSED && use_direct) {
#pragma unroll
            for (int k = 0; k < CPD; ++k) {
                f32x2 a[PPT][2];
#pragma unroll
                for (int q = 0; q < PPT; ++q) a[q][0] = a[q][1] = 0ull;
                const ulonglong2* col = reinterpret_cast<const ulonglong2*>(use_rel ? sm.cpw_rel[k] : sm.cpw[tb][k]);
#pragma unroll
                for (int m = 0; m < NQ; ++m) {
                    const ulonglong2 cv = col[m];
#pragma unroll
                    for (int q = 0; q < PPT; ++q) {
                        ffma2(a[q][0], tp2[q][2 * m], cv.x);
                        ffma2(a[q][1], tp2[q][2 * m + 1], cv.y);
                    }
                }
#pragma unroll
                for (int q = 0; q < PPT; ++q) acc[q][k] = (G == 1 || grp == 0) ? hsum2(a[q][0]) + hsum2(a[q][1]) : 0.0f;
            }
        }
        if (MODE == MODE_FUSED && nsep > 0) {
            float fa[PPT][SH::N0], fb[PPT][SH::N1], fc[PPT][SH::N2];
#pragma unroll
            for (int q = 0; q < PPT; ++q) bs[q].select(ch, fa[q], fb[q], fc[q]);
            const float4* ccs = use_rel ? sm.cc_rel : sm.cc;
#pragma unroll 2
            for (int s0 = 0; s0 < nsep; s0 += G) {
                const bool s_on = G == 1 || s0 + grp < nsep;
                const int s = (G == 1) ? s0 : min(s0 + grp, nsep - 1);
                const float4 v0 = sm.sepv[s][0], v1 = sm.sepv[s][1], v2 = sm.sepv[s][2];
                const float4 cc = ccs[s];
#pragma unroll
                for (int q = 0; q < PPT; ++q) {
                    float e0 = 0.f, e1 = 0.f, e2 = 0.f;
#pragma unroll
                    for (int i = 0; i < SH::N0; ++i) e0 = fmaf(f4c(v0, i), fa[q][i], e0);
#pragma unroll
                    for (int i = 0; i < SH::N1; ++i) e1 = fmaf(f4c(v1, i), fb[q][i], e1);
#pragma unroll
                    for (int i = 0; i < SH::N2; ++i) e2 = fmaf(f4c(v2, i), fc[q][i], e2);
                    const float phi = s_on ? e0 * e1 * e2 : 0.0f;
                    acc[q][0] = fmaf(phi, cc.x, acc[q][0]);
                    acc[q][1] = fmaf(phi, cc.y, acc[q][1]);
                    acc[q][2] = fmaf(phi, cc.z, acc[q][2]);
                    if (CPD == 4) acc[q][3] = fmaf(phi, cc.w, acc[q][3]);
                }
            }
        }
        if (tile_loop) {
            // tiles stream through two shared-memory buffers: cp.async of chunk c+1 overlaps the FMAs of chunk c
            if (chi > 0) issue_chunk<NQ>(tile_src, 0, min(kWChunk, nfl), TS, sm.tiles[0]);
            int buf = 0;
            for (int c0 = 0; c0 < nfl; c0 += kWChunk) {
                const int ns = min(kWChunk, nfl - c0);
                const int c1 = c0 + kWChunk;
                if (c1 < nfl) {
                    issue_chunk<NQ>(tile_src, c1, min(kWChunk, nfl - c1), TS, sm.tiles[buf ^ 1]);
                    cp_async_wait<1>();
                } else {
                    cp_async_wait<0>();
                }
                if (MODE == MODE_FUSED && !cc_all) gather_cc<CPD>(sp, ar, it, c0, ns, sm.cc);
                __syncwarp();
                const float* tbuf = sm.tiles[buf];
                const float4* ccp = (use_rel ? sm.cc_rel : sm.cc) + (cc_all ? nsep + c0 : 0);
THB_PRAGMA(unroll THB_SUNROLL)
                for (int s0 = 0; s0 < ns; s0 += G) {
                    const bool s_on = G == 1 || s0 + grp < ns;
                    const int s = (G == 1) ? s0 : min(s0 + grp, ns - 1);
                    const ulonglong2* row = reinterpret_cast<const ulonglong2*>(tbuf + s * TSP);
                    f32x2 a[PPT][2];
#pragma unroll
                    for (int q = 0; q < PPT; ++q) a[q][0] = a[q][1] = 0ull;
#pragma unroll
                    for (int m = 0; m < NQ; ++m) {
                        const ulonglong2 tv = row[m];
#pragma unroll
                        for (int q = 0; q < PPT; ++q) {
                            ffma2(a[q][0], tp2[q][2 * m], tv.x);
                            ffma2(a[q][1], tp2[q][2 * m + 1], tv.y);
                        }
                    }
                    if (MODE == MODE_FUSED) {
                        const float4 cc = ccp[s];
#pragma unroll
                        for (int q = 0; q < PPT; ++q) {
                            const float phi = s_on ? hsum2(a[q][0]) + hsum2(a[q][1]) : 0.0f;
                            acc[q][0] = fmaf(phi, cc.x, acc[q][0]);
                            acc[q][1] = fmaf(phi, cc.y, acc[q][1]);
                            acc[q][2] = fmaf(phi, cc.z, acc[q][2]);
                            if (CPD == 4) acc[q][3] = fmaf(phi, cc.w, acc[q][3]);
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < PPT; ++q) sm.stage[(q * 32 + lane) * (kWChunk + 1) + s] = hsum2(a[q][0]) + hsum2(a[q][1]);
                    }
                }
                __syncwarp();
                if (MODE == MODE_BASIS) {
                    // coalesced write-out of the chunk: 16 consecutive supports of a point are contiguous in PHI
#pragma unroll
                    for (int q = 0; q < PPT; ++q) {
                        for (int e = lane; e < nval[q] * kWChunk; e += 32) {
                            const int pt = e / kWChunk, sc = e - pt * kWChunk;
                            if (sc < ns) ar.out[chi][sm.poff[q * 32 + pt] + it.ndir + c0 + sc] = sm.stage[(q * 32 + pt) * (kWChunk + 1) + sc];
                        }
                    }
                    __syncwarp();
                }
                buf ^= 1;
            }
        }
        if (MODE == MODE_FUSED) {
            if (G > 1) {  // combine the support groups
#pragma unroll
                for (int k = 0; k < CPD; ++k)
#pragma unroll
                    for (int o = NPT; o < 32; o <<= 1) acc[0][k] += __shfl_xor_sync(0xffffffffu, acc[0][k], o);
            }
#pragma unroll
            for (int q = 0; q < PPT; ++q) {
                if (valid[q] && (G == 1 || grp == 0)) {
                    float* o = ar.out[chi] + (int64_t)idx[q] * CPD;
#pragma unroll
                    for (int k = 0; k < CPD; ++k) o[k] = acc[q][k];
                }
            }
        }
    }
}

// Persistent, warp-autonomous kernel: every CTA is ONE warp that pulls work items from an atomic cursor.
// No CTA-wide barrier exists.  Software pipeline: while item i is computed, the records of items i+1 and
// i+2 are (being) fetched, the id of item i+3 is being claimed, and the points of the next sub-batch plus
// the knots / window control points of item i+1 stream into the other half of the shared-memory buffers.
template <int N0, int N1, int N2, int PPT, int CPD, int MODE, bool DER>
__global__ void __launch_bounds__(kThreads, THB_TILE_MINB) k_tile(const __grid_constant__ thb_space_desc sp, const __grid_constant__ Args ar) {
    using SH = Shp<N0, N1, N2>;
    constexpr int TS = SH::TS, NQ = SH::NQ;
    constexpr int SB = 32 * PPT;
    __shared__ __align__(16) WarpSmem<SH, MODE, DER> sm;

    const int lane = threadIdx.x;
    const int nitems = ar.counters[0];
    int32_t* const cursor = ar.counters + 1;
    const int d = sp.ndim;
    const bool cellnet = (MODE == MODE_FUSED) && ar.cellnet;
    const bool want_cp = (MODE == MODE_FUSED);

    // claim three items up front (ids are increasing); id3 is claimed asynchronously
    int c3 = 0;
    if (lane == 0) c3 = atomicAdd(cursor, 3);
    const int id0 = __shfl_sync(0xffffffffu, c3, 0);
    if (id0 >= nitems) return;
    int id1 = id0 + 1, id2 = id0 + 2;
    Item it = load_item(sp, ar.items, id0);
    if (lane < 4 && id1 < nitems) cp_async16(&sm.rec[1][lane], ar.items + (int64_t)id1 * 4 + lane);
    if (lane < 4 && id2 < nitems) cp_async16(&sm.rec[2][lane], ar.items + (int64_t)id2 * 4 + lane);
    prefetch_item_tables<SH, CPD>(sp, ar, it, sm.knots[0], sm.cpw[0], want_cp);
    prefetch_points<PPT>(ar, it.begin, it.count, 0, sm.prec[0]);
    cp_async_commit();
    int nn_id = 0;
    if (lane == 0) nn_id = atomicAdd(cursor, 1);
    int ring = 0, tb = 0, pb = 0;
    cp_async_wait<0>();
    __syncwarp();

    for (;;) {
        const int ntile = it.S - it.ndir;
        // reciprocal knot differences of this cell (from the staged knots)
        if (lane < 24) {
            const int dim = lane >> 3, m = lane & 7;
            const int P = dim == 0 ? SH::P0 : (dim == 1 ? SH::P1 : SH::P2);
            if (dim < d && m < P * (P + 1) / 2) {
                int j = 1, r = m;
                while (r >= j) { r -= j; ++j; }
                sm.rk[dim][m] = 1.0f / (sm.knots[tb][dim][P + r] - sm.knots[tb][dim][P - j + r]);
            }
        }
        const bool split = MODE == MODE_FUSED && !cellnet && sp.sep_vec != nullptr && ntile <= kMaxCC;
        if (split) {
            // rank-one supports: factor vectors by cp.async, their control points, then the full-tile supports' ones
            const float4* src = reinterpret_cast<const float4*>(sp.sep_vec + (int64_t)it.sep_off * 12);
            for (int i = lane; i < it.nsep * 3; i += 32) cp_async16(&sm.sepv[0][0] + i, src + i);
            cp_async_commit();
            gather_ids<CPD>(sp.sep_jm + it.sep_off, ar, it.nsep, sm.cc);
            gather_ids<CPD>(sp.full_jm + it.full_off, ar, ntile - it.nsep, sm.cc + it.nsep);
            cp_async_wait<0>();
        } else if (MODE == MODE_FUSED && ntile > 0 && ntile <= kMaxCC) {
            gather_cc<CPD>(sp, ar, it, 0, ntile, sm.cc);
        }
        if (MODE == MODE_FUSED && DER && lane < 4) {
            // reference point: the control point of the item's first support
            const int64_t r = __ldg(sp.supp_jm + it.soff);
            sm.ref[lane] = (lane < CPD && it.S > 0) ? __ldg(ar.ctrl_pts + r * CPD + lane) : 0.f;
        }
        __syncwarp();
        if (MODE == MODE_FUSED && DER) {
            for (int t = lane; t < TS; t += 32) {
                const bool on = it.ndir > 0 && t < SH::T && ((it.dm >> t) & 1ull);
#pragma unroll
                for (int k = 0; k < 4; ++k) sm.cpw_rel[k][t] = on ? sm.cpw[tb][k][t] - sm.ref[k] : 0.f;
            }
            if (ntile <= kMaxCC) {
                for (int s = lane; s < ntile; s += 32) {
                    const float4 c = sm.cc[s];
                    sm.cc_rel[s] = make_float4(c.x - sm.ref[0], c.y - sm.ref[1], c.z - sm.ref[2], c.w - sm.ref[3]);
                }
            }
            __syncwarp();
        }
        if (cellnet) {
            for (int c0 = 0; c0 < ntile; c0 += kWChunk) {
                const int ns = min(kWChunk, ntile - c0);
                const float4* src = reinterpret_cast<const float4*>(sp.tiles + (int64_t)(it.toff + c0) * TS);
                float4* dst = reinterpret_cast<float4*>(sm.tiles[0]);
                for (int i = lane; i < ns * NQ; i += 32) dst[i] = __ldg(src + i);
                if (ntile > kMaxCC) gather_cc<CPD>(sp, ar, it, c0, ns, sm.cc);
                __syncwarp();
                const float4* ccp = sm.cc + (ntile > kMaxCC ? 0 : c0);
                for (int t = lane; t < TS; t += 32) {
                    float q[4] = {0.f, 0.f, 0.f, 0.f};
                    for (int s = 0; s < ns; ++s) {
                        const float w = sm.tiles[0][s * TS + t];
                        const float4 cc = ccp[s];
                        q[0] = fmaf(w, cc.x, q[0]); q[1] = fmaf(w, cc.y, q[1]);
                        q[2] = fmaf(w, cc.z, q[2]); q[3] = fmaf(w, cc.w, q[3]);
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) sm.cpw[tb][k][t] += q[k];
                }
                __syncwarp();
            }
        }
        const bool nx_valid = id1 < nitems;
        for (int base = 0; base < it.count; base += SB) {
            // prefetch the next sub-batch: same item, or the first one of the next item together with its tables
            if (base + SB < it.count) {
                prefetch_points<PPT>(ar, it.begin, it.count, base + SB, sm.prec[pb ^ 1]);
            } else if (nx_valid) {
                const int4* r = sm.rec[(ring + 1) % 3];
                const Item nx = unpack_item(r[0], r[1], r[2], r[3]);
                prefetch_points<PPT>(ar, nx.begin, nx.count, 0, sm.prec[pb ^ 1]);
                prefetch_item_tables<SH, CPD>(sp, ar, nx, sm.knots[tb ^ 1], sm.cpw[tb ^ 1], want_cp);
            }
            cp_async_commit();
            const int rem = it.count - base;
            if (MODE == MODE_FUSED && rem <= 8 && ntile > 0)
                warp_points<SH, 1, (MODE == MODE_FUSED ? 4 : 1), CPD, MODE, DER>(sp, ar, it, base, ntile, cellnet, sm, tb, pb);
            else if (MODE == MODE_FUSED && rem <= 16 && ntile > 0)
                warp_points<SH, 1, (MODE == MODE_FUSED ? 2 : 1), CPD, MODE, DER>(sp, ar, it, base, ntile, cellnet, sm, tb, pb);
            else if (PPT == 2 && rem <= 32)
                warp_points<SH, 1, 1, CPD, MODE, DER>(sp, ar, it, base, ntile, cellnet, sm, tb, pb);
            else
                warp_points<SH, PPT, 1, CPD, MODE, DER>(sp, ar, it, base, ntile, cellnet, sm, tb, pb);
            cp_async_wait<0>();
            __syncwarp();
            pb ^= 1;
        }
        if (!nx_valid) break;
        // advance: item i+1 becomes current; fetch the record of item i+3 into the slot item i used
        ring = (ring + 1) % 3;
        {
            const int4* r = sm.rec[ring];
            it = unpack_item(r[0], r[1], r[2], r[3]);
        }
        tb ^= 1;
        const int id3 = __shfl_sync(0xffffffffu, nn_id, 0);
        id1 = id2;
        id2 = id3;
        if (lane < 4 && id3 < nitems) cp_async16(&sm.rec[(ring + 2) % 3][lane], ar.items + (int64_t)id3 * 4 + lane);
        cp_async_commit();
        if (lane == 0) nn_id = atomicAdd(cursor, 1);
    }
}

// ---- fused backward: grad_ctrl_pts = A^T grad_out without materialising A --------------------------------
// Per item: every thread writes the tensor product of its point into shared memory (point-minor layout),
// the CTA reduces gQ[t][k] = sum_points tp[point][t] * g[point][k], then scatters
//   own-level supports : grad_cp[Jm_t] += gQ[t]
//   coarser supports   : grad_cp[Jm_s] += sum_t tile_s[t] * gQ[t]
// i.e. S*cp_dim atomics per ITEM (up to 256 points) instead of per point as in the reference
// (compute_pts_cuda_kernels.cu:28-47).
template <int N0, int N1, int N2, int CPD>
__global__ void __launch_bounds__(kBwdThreads) k_tile_backward(const __grid_constant__ thb_space_desc sp,
                                                            const __grid_constant__ Args ar) {
    using SH = Shp<N0, N1, N2>;
    constexpr int T = SH::T, TS = SH::TS;
    constexpr int LD = kBwdThreads + 1;  // padded row length: conflict-free column reads
    __shared__ float s_knots[3][8];
    __shared__ float s_tp[T * LD];
    __shared__ float s_g[4][LD];
    __shared__ float s_gq[TS][4];
    __shared__ int s_item;

    const int tid = threadIdx.x;
    const int nitems = ar.counters[0];
    const int d = sp.ndim;

    for (;;) {
        if (tid == 0) s_item = atomicAdd(ar.counters + 1, 1);
        __syncthreads();
        const int iti = s_item;
        if (iti >= nitems) break;
        const Item it = load_item(sp, ar.items, iti);
        stage_knots<SH>(sp, it, s_knots);
        float gq[(T * CPD + kBwdThreads - 1) / kBwdThreads];  // this thread's share of the (t,k) outputs
#pragma unroll
        for (int o = 0; o < (T * CPD + kBwdThreads - 1) / kBwdThreads; ++o) gq[o] = 0.0f;
        __syncthreads();

        for (int base = 0; base < it.count; base += kBwdThreads) {
            const int li = base + tid;
            const bool valid = li < it.count;
            const float4 rec = __ldg(reinterpret_cast<const float4*>(ar.sorted) + it.begin + (valid ? li : 0));
            const int idx = __float_as_int(rec.w);
            const float u[3] = {rec.x, rec.y, rec.z};
            Bases<SH, false> bs;
            bs.eval(u, s_knots);
            float tp[TS];
            bs.tensor(0, tp);
#pragma unroll
            for (int t = 0; t < T; ++t) s_tp[t * LD + tid] = valid ? tp[t] : 0.0f;
#pragma unroll
            for (int k = 0; k < 4; ++k) s_g[k][tid] = (valid && k < CPD) ? __ldg(ar.grad_out + (int64_t)idx * CPD + k) : 0.0f;
            __syncthreads();
            const int np = min(kBwdThreads, it.count - base);
#pragma unroll
            for (int o = 0; o < (T * CPD + kBwdThreads - 1) / kBwdThreads; ++o) {
                const int e = o * kBwdThreads + tid;
                if (e < T * CPD) {
                    const int t = e / CPD, k = e - t * CPD;
                    const float* row = s_tp + t * LD;
                    const float* g = s_g[k];
                    float a = 0.0f;
                    for (int p = 0; p < np; ++p) a = fmaf(row[p], g[p], a);
                    gq[o] += a;
                }
            }
            __syncthreads();
        }
        // publish gQ
#pragma unroll
        for (int o = 0; o < (T * CPD + kBwdThreads - 1) / kBwdThreads; ++o) {
            const int e = o * kBwdThreads + tid;
            if (e < T * CPD) s_gq[e / CPD][e % CPD] = gq[o];
        }
        __syncthreads();
        // own-level supports
        if (it.ndir > 0) {
            for (int e = tid; e < T * CPD; e += kBwdThreads) {
                const int t = e / CPD, k = e - t * CPD;
                if ((it.dm >> t) & 1ull) {
                    const int pos = __popcll(it.dm & ((1ull << t) - 1ull));
                    const int64_t r = __ldg(sp.supp_jm + it.soff + pos);
                    atomicAdd(ar.grad_cp + r * CPD + k, s_gq[t][k]);
                }
            }
        }
        // coarser-level supports
        const int ntile = it.S - it.ndir;
        for (int e = tid; e < ntile * CPD; e += kBwdThreads) {
            const int s = e / CPD, k = e - s * CPD;
            const float4* row = reinterpret_cast<const float4*>(sp.tiles + (int64_t)(it.toff + s) * TS);
            float a = 0.0f;
#pragma unroll
            for (int m = 0; m < TS / 4; ++m) {
                const float4 w = __ldg(row + m);
                a = fmaf(w.x, s_gq[4 * m + 0][k], a);
                a = fmaf(w.y, s_gq[4 * m + 1][k], a);
                a = fmaf(w.z, s_gq[4 * m + 2][k], a);
                a = fmaf(w.w, s_gq[4 * m + 3][k], a);
            }
            const int64_t r = __ldg(sp.supp_jm + it.soff + it.ndir + s);
            atomicAdd(ar.grad_cp + r * CPD + k, a);
        }
        // s_gq / s_knots are rewritten only after the next item's first __syncthreads
    }
}

// ---- launch plumbing ---------------------------------------------------------------------------------------

template <typename K>
int persistent_grid(K kernel, int threads) {
    int per_sm = 0;
    cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
    return per_sm * thb_num_sms();
}

#define THB_TILE_LAUNCH_N(KERNEL, NT)                                             \
    do {                                                                          \
        static int grid_[THB_MAX_DEVICES] = {};                                   \
        int& grid = grid_[thb_device()];                                          \
        if (grid == 0) grid = persistent_grid(KERNEL, NT);                        \
        KERNEL<<<grid, NT, 0, st>>>(*sp, *ar);                                    \
        THB_LAUNCHED();                                                           \
        return THB_OK;                                                            \
    } while (0)
#define THB_TILE_LAUNCH(KERNEL) THB_TILE_LAUNCH_N(KERNEL, kThreads)

template <int N0, int N1, int N2>
int launch_fused(const thb_space_desc* sp, const Args* ar, int ppt, int cpd, int der, cudaStream_t st) {
    if (ppt == 2) {
        if (cpd == 3 && !der) THB_TILE_LAUNCH((k_tile<N0, N1, N2, 2, 3, MODE_FUSED, false>));
        if (cpd == 3 && der) THB_TILE_LAUNCH((k_tile<N0, N1, N2, 2, 3, MODE_FUSED, true>));
        if (cpd == 4 && !der) THB_TILE_LAUNCH((k_tile<N0, N1, N2, 2, 4, MODE_FUSED, false>));
        if (cpd == 4 && der) THB_TILE_LAUNCH((k_tile<N0, N1, N2, 2, 4, MODE_FUSED, true>));
    } else {
        if (cpd == 3 && !der) THB_TILE_LAUNCH((k_tile<N0, N1, N2, 1, 3, MODE_FUSED, false>));
        if (cpd == 3 && der) THB_TILE_LAUNCH((k_tile<N0, N1, N2, 1, 3, MODE_FUSED, true>));
        if (cpd == 4 && !der) THB_TILE_LAUNCH((k_tile<N0, N1, N2, 1, 4, MODE_FUSED, false>));
        if (cpd == 4 && der) THB_TILE_LAUNCH((k_tile<N0, N1, N2, 1, 4, MODE_FUSED, true>));
    }
    return thb_set_error(THB_E_UNSUPPORTED, "fused: unsupported variant");
}

template <int N0, int N1, int N2>
int launch_basis(const thb_space_desc* sp, const Args* ar, int ppt, int der, cudaStream_t st) {
    if (ppt == 2) {
        if (!der) THB_TILE_LAUNCH((k_tile<N0, N1, N2, 2, 3, MODE_BASIS, false>));
        THB_TILE_LAUNCH((k_tile<N0, N1, N2, 2, 3, MODE_BASIS, true>));
    }
    if (!der) THB_TILE_LAUNCH((k_tile<N0, N1, N2, 1, 3, MODE_BASIS, false>));
    THB_TILE_LAUNCH((k_tile<N0, N1, N2, 1, 3, MODE_BASIS, true>));
}

template <int N0, int N1, int N2>
int launch_backward(const thb_space_desc* sp, const Args* ar, int cpd, cudaStream_t st) {
    if (cpd == 3) THB_TILE_LAUNCH_N((k_tile_backward<N0, N1, N2, 3>), kBwdThreads);
    if (cpd == 4) THB_TILE_LAUNCH_N((k_tile_backward<N0, N1, N2, 4>), kBwdThreads);
    return thb_set_error(THB_E_UNSUPPORTED, "backward: unsupported cp_dim");
}

}  // namespace thb_tile
