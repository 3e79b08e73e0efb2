// Local-control-net evaluation: the default formulation of the fused THB_basis_fns + THBEval.forward path
// (THB/core.py:604-701 + THB/torch_funcs.py:22-33).
//
// A coarser-level support s of a cell contributes PHI_s(u) * CP_s with PHI_s(u) = <tile_s, tp(u)>, tp = the (p+1)^d
// tensor product of the cell's own-level B-splines.  Summed over the supports of the cell
//     sum_s PHI_s(u) CP_s = < tp(u), W >,      W[t] = [t active] * CP_own[t] + sum_s tile_s[t] * CP_s
// so W -- the cell's LOCAL CONTROL NET, (p+1)^d control points -- depends on the hierarchy and the control points only.
//   k_cell_nets : one warp per active cell builds W (own-level window + fold of the coarser-level supports: rank-one
//                 ones from their three factor vectors, the others from their tiles) and writes it to the workspace
//                 [nslots][planes][TS]; planes 0..3 = W, planes 4..7 (derivative channels) = W built from control
//                 points RELATIVE to the cell's first support (derivative basis values sum to zero, so the result is
//                 the same, but the summed terms are O(p/h * h) instead of O(p/h * |CP|): no float32 cancellation)
//   k_net_eval  : persistent warp-autonomous kernel over the plan's work items; a point costs Cox-de Boor plus the
//                 sum-factorised contraction with W ((p+1)^d + (p+1)^(d-1) + (p+1)^(d-2) FMAs per output component),
//                 no (p+1)^d register array and no per-point table traffic.
#pragma once
#include "thb_tile_impl.cuh"

namespace thb_tile {

#ifndef THB_NET_MINB
#define THB_NET_MINB 16  // k_net_eval: <= 128 registers, 16 one-warp CTAs per SM
#endif
#ifndef THB_NET_MINB_DER
#define THB_NET_MINB_DER 16  // derivative channels: more live values (both planes sets, values and derivatives of the bases)
#endif
#ifndef THB_NETS_MINB
#define THB_NETS_MINB 9  // k_cell_nets: <= 56 registers, 36 warps per SM (it is latency-bound: occupancy matters; prefetching
                         // the next cell's descriptor / control points one chunk ahead cost more registers than it hid latency)
#endif
constexpr int kNetChunk = 32;     // tiled supports staged at once by one warp of k_cell_nets
constexpr int kNetWarps = 4;      // warps (= cells in flight) per CTA of k_cell_nets

struct NetArgs {
    const float* ctrl_pts;
    const int32_t* cell_count;  // plan histogram: cells without points are skipped (nullptr = every cell)
    float* nets;                // [nslots][planes][TS]
    int32_t planes;             // 4, or 8 with the relative planes
    const int32_t* row_map;     // k_cell_unfold: flat function id -> row of the (compact) gradient buffer; nullptr = identity
    int32_t mono;               // k_cell_nets: 1 = store the nets in MONOMIAL form (see net_to_monomial), 0 = control points
    float* contrib;             // k_cell_unfold, ordered mode: [nsupp][cp_dim] one slot per (cell, support), plain stores
    int32_t ctas_per_sm;        // k_cell_nets: > 0 caps the resident CTAs per SM (room for a kernel on another stream)
};

template <int CPD>
THB_DEV float4 load_cp(const float* __restrict__ cp, int64_t r) {
    float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < CPD; ++k) v[k] = __ldg(cp + r * CPD + k);
    return make_float4(v[0], v[1], v[2], v[3]);
}

// control point minus the cell's reference point (monomial nets are built from RELATIVE control points: every polynomial
// coefficient but the constant one is a difference of control points, and forming it from small numbers keeps the
// first-derivative channels free of float32 cancellation; the reference point is added back to the constant term)
template <int CPD>
THB_DEV float4 load_cp_sub(const float* __restrict__ cp, int64_t r, const float* ref, bool sub) {
    float4 v = load_cp<CPD>(cp, r);
    if (sub) { v.x -= ref[0]; v.y -= ref[1]; v.z -= ref[2]; v.w -= ref[3]; }
    return v;
}

// (w01, w23)[k] += (x01, x23) * c[k]   (and the relative planes with c - ref)
template <int CPD, bool REL>
THB_DEV void fold2(f32x2 (*w2)[4], f32x2 (*wr2)[4], f32x2 x01, f32x2 x23, const float4& c, const float* ref) {
    const float ck[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
    for (int k = 0; k < CPD; ++k) {
        const f32x2 c2 = pack2(ck[k], ck[k]);
        ffma2(w2[0][k], x01, c2);
        ffma2(w2[1][k], x23, c2);
        if (REL) {
            const float r = ck[k] - ref[k];
            const f32x2 r2 = pack2(r, r);
            ffma2(wr2[0][k], x01, r2);
            ffma2(wr2[1][k], x23, r2);
        }
    }
}

// Control net -> monomial form of one cell, in place in a per-warp shared-memory buffer [planes][TS].
// Over a cell the geometry is <tp(u), W> with tp the tensor product of the cell's p+1 B-splines per dimension; with
// N_r(u_k) = sum_m C_k[m][r] s_k^m (sp.poly: s_k = (u_k - lo_k) / half_k - 1 in [-1, 1], float64 on the host)
//     <tp(u), W> = sum_{m0,m1,m2} M[m0][m1][m2] s0^m0 s1^m1 s2^m2,   M = (C_0 x C_1 x C_2) W
// so a point costs three subtractions and a nested Horner scheme (k_net_eval, MONO) instead of Cox-de Boor plus the
// contraction with tp(u).  Three separable passes, one per dimension; every lane owns the entries t = lane (+ 32).
// cmat: the three 4 x 4 matrices C_k[m][r] of the cell, staged in shared memory by stage_span_polys (cp.async issued when
// the cell is picked up, so their latency is hidden behind the fold of the supports).
THB_DEV void stage_span_polys(const thb_space_desc& sp, int lev, int c0, int c1, int c2, float4 (*cmat)[4]) {
    const int lane = threadIdx.x & 31;
    if (lane < 12) {
        const int dim = lane >> 2, m = lane & 3;
        const int c = dim == 0 ? c0 : (dim == 1 ? c1 : c2);
        if (dim < sp.ndim) cp_async16(&cmat[dim][m], sp.poly + (int64_t)(sp.poly_off[lev][dim] + c) * 20 + 4 * m);
    }
    cp_async_commit();
}

template <int N0, int N1, int N2, int CPD, bool REL>
THB_DEV void net_to_monomial(const float4 (*cmat)[4], float* buf, int TS) {
    constexpr int planes = REL ? 2 * CPD : CPD;   // plane index k -> buffer plane k (value) or 4 + k - CPD (relative)
    auto plane_of = [](int k) { return k < CPD ? k : 4 + (k - CPD); };
    constexpr int T = N0 * N1 * N2;
    constexpr int NT = (T + 31) / 32;
    const int lane = threadIdx.x & 31;
    constexpr int NN[3] = {N0, N1, N2};
    constexpr int STR[3] = {N1 * N2, N2, 1};
#pragma unroll
    for (int dim = 2; dim >= 0; --dim) {
        if (NN[dim] == 1) continue;
        float o[NT][planes];
#pragma unroll
        for (int h = 0; h < NT; ++h) {
            const int t = lane + 32 * h;
            if (t < T) {
                const int m = (t / STR[dim]) % NN[dim];
                const int base = t - m * STR[dim];
                const float4 cm = cmat[dim][m];   // C[m][0..3]
                const float cr[4] = {cm.x, cm.y, cm.z, cm.w};
#pragma unroll
                for (int k = 0; k < planes; ++k) {
                    float acc = 0.f;
#pragma unroll
                    for (int r = 0; r < NN[dim]; ++r) acc = fmaf(cr[r], buf[plane_of(k) * TS + base + r * STR[dim]], acc);
                    o[h][k] = acc;
                }
            }
        }
        __syncwarp();
#pragma unroll
        for (int h = 0; h < NT; ++h) {
            const int t = lane + 32 * h;
            if (t < T) {
#pragma unroll
                for (int k = 0; k < planes; ++k) buf[plane_of(k) * TS + t] = o[h][k];
            }
        }
        __syncwarp();
    }
}

template <int N0, int N1, int N2, int CPD, bool REL>
__global__ void __launch_bounds__(32 * kNetWarps, (REL || CPD == 4) ? 6 : THB_NETS_MINB) k_cell_nets(const __grid_constant__ thb_space_desc sp, const __grid_constant__ NetArgs na) {
    using SH = Shp<N0, N1, N2>;
    constexpr int T = SH::T, TS = SH::TS;
    constexpr int NT = (TS + 31) / 32;
    constexpr bool kSameBC = (32 % (N1 * N2)) == 0;  // t and t + 32 differ in the first index only
    __shared__ __align__(16) float4 s_sepv[kNetWarps][kNetChunk][3];
    __shared__ __align__(16) float4 s_cc[kNetWarps][kNetChunk];
    __shared__ __align__(16) float s_conv[kNetWarps][REL ? 8 : 4][TS];   // control net -> monomial form (na.mono)
    __shared__ __align__(16) float4 s_cmat[kNetWarps][3][4];              // the cell's span polynomials C_k[m][r]
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float4(*sepv)[3] = s_sepv[wid];
    float4* cc = s_cc[wid];
    const bool split = sp.sep_vec != nullptr;
    const int ib = 4 + (lane / N2) % N1, ic = 8 + lane % N2;

    for (int slot = blockIdx.x * kNetWarps + wid; slot < sp.nslots; slot += gridDim.x * kNetWarps) {
        if (na.cell_count && __ldg(na.cell_count + slot) == 0) continue;
        const int4* ci = reinterpret_cast<const int4*>(sp.cellinfo + (int64_t)slot * THB_INFO_W);
        const int4 a = __ldg(ci), e = __ldg(ci + 1), f = __ldg(ci + 2);  // level c0 c1 c2 | supp_off S tile_off n_direct | sep_off n_sep full_off
        const unsigned long long dm = __ldg(sp.dmask + slot);
        const int lev = a.x, soff = e.x, S = e.y, ndir = e.w;
        const int ntile = S - ndir;
        const int nsep = split ? f.y : 0;
        if (na.mono) stage_span_polys(sp, lev, a.y, a.z, a.w, s_cmat[wid]);
        float ref[4] = {0.f, 0.f, 0.f, 0.f};
        const bool sub = na.mono != 0;   // monomial form: W is built from control points relative to `ref`
        if ((REL || sub) && S > 0) {
            const float4 r = load_cp<CPD>(na.ctrl_pts, __ldg(sp.supp_jm + soff));
            ref[0] = r.x; ref[1] = r.y; ref[2] = r.z; ref[3] = r.w;
        }
        if constexpr (N0 == 4 && N1 * N2 == 16) {
            // Cubic-in-u shapes with a 16-entry (b, c) face: lane = (half, b, c) owns the four window entries a = 0..3 of
            // its (b, c); the two half-warps take alternate supports and are summed with one shuffle per value at the
            // end.  Per support a lane issues 4 shared-memory loads and 8 packed FP32 instructions for 12 FMAs (the
            // generic path below: 5 loads and 5 instructions for 6 FMAs, and twice the shared-memory wavefronts).
            const int half = lane >> 4, l16 = lane & 15;
            const int jb = 4 + l16 / N2, jc = 8 + l16 % N2;
            f32x2 w2[2][4], wr2[REL ? 2 : 1][4];
            {
                const int n1 = sp.nfns[lev][1], n2 = sp.nfns[lev][2];
                const int64_t f0 = sp.fn_off[lev];
                const int t1 = l16 / N2, t2 = l16 % N2;
#pragma unroll
                for (int p = 0; p < 2; ++p) {
                    float4 v[2];
                    bool on[2];
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const int t0 = 2 * p + q, t = t0 * 16 + l16;
                        on[q] = ndir > 0 && p == half && ((dm >> t) & 1ull);  // half h loads the entries a = 2h, 2h+1
                        v[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (on[q]) v[q] = load_cp_sub<CPD>(na.ctrl_pts, f0 + ((int64_t)(a.y + t0) * n1 + (a.z + t1)) * n2 + (a.w + t2), ref, sub);
                    }
                    w2[p][0] = pack2(v[0].x, v[1].x); w2[p][1] = pack2(v[0].y, v[1].y); w2[p][2] = pack2(v[0].z, v[1].z);
                    w2[p][3] = pack2(v[0].w, v[1].w);
                    if (REL) {
                        wr2[p][0] = pack2(on[0] ? v[0].x - ref[0] : 0.f, on[1] ? v[1].x - ref[0] : 0.f);
                        wr2[p][1] = pack2(on[0] ? v[0].y - ref[1] : 0.f, on[1] ? v[1].y - ref[1] : 0.f);
                        wr2[p][2] = pack2(on[0] ? v[0].z - ref[2] : 0.f, on[1] ? v[1].z - ref[2] : 0.f);
                        wr2[p][3] = pack2(on[0] ? v[0].w - ref[3] : 0.f, on[1] ? v[1].w - ref[3] : 0.f);
                    }
                }
            }
            const float* tile_src = split ? sp.full_tiles + (int64_t)f.z * TS : sp.tiles + (int64_t)e.z * TS;
            const int32_t* full_jm = split ? sp.full_jm + f.z : sp.supp_jm + soff + ndir;
            static_assert(kNetChunk == 32, "one tiled support per lane and chunk");
            for (int c0 = 0; c0 < ntile; c0 += kNetChunk) {
                const int n = min(kNetChunk, ntile - c0);
                const int s_lo = min(c0, nsep), s_n = min(c0 + n, nsep) - s_lo;
                const int f_lo = max(c0, nsep) - nsep, f_n = n - s_n;
                __syncwarp();
                if (s_n > 0) {
                    const float4* src = reinterpret_cast<const float4*>(sp.sep_vec + (int64_t)(f.x + s_lo) * 12);
#pragma unroll
                    for (int i = 0; i < 3; ++i)
                        if (lane + 32 * i < s_n * 3) cp_async16(&sepv[0][0] + lane + 32 * i, src + lane + 32 * i);
                    cp_async_commit();
                }
                {   // control point of tiled support c0 + lane (rank-one ones first)
                    const int g = c0 + lane;
                    if (g < ntile) cc[lane] = load_cp_sub<CPD>(na.ctrl_pts, __ldg(g < nsep ? sp.sep_jm + f.x + g : full_jm + (g - nsep)), ref, sub);
                }
                cp_async_wait<0>();
                __syncwarp();
#pragma unroll 4
                for (int i = half; i < s_n; i += 2) {
                    const float* sv = reinterpret_cast<const float*>(&sepv[i][0]);
                    const float4 v0 = sepv[i][0];
                    const float4 c = cc[i];
                    const float bc = sv[jb] * sv[jc];
                    const f32x2 bc2 = pack2(bc, bc);
                    const f32x2 x01 = fmul2(pack2(v0.x, v0.y), bc2), x23 = fmul2(pack2(v0.z, v0.w), bc2);
                    fold2<CPD, REL>(w2, wr2, x01, x23, c, ref);
                }
                const float* rows = tile_src + (int64_t)f_lo * TS + l16;
#pragma unroll 4
                for (int j = half; j < f_n; j += 2) {
                    const float* r = rows + (int64_t)j * TS;
                    const float4 c = cc[s_n + j];
                    const f32x2 x01 = pack2(__ldg(r), __ldg(r + 16)), x23 = pack2(__ldg(r + 32), __ldg(r + 48));
                    fold2<CPD, REL>(w2, wr2, x01, x23, c, ref);
                }
            }
            // sum the two halves; half h then writes the entries a = 2h, 2h+1 (to the workspace, or to shared memory on
            // the way to the monomial form)
            float* const gdst = na.nets + (int64_t)slot * na.planes * TS;
            float* dst = na.mono ? &s_conv[wid][0][0] : gdst;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float o[2][2], orr[2][2];
#pragma unroll
                for (int p = 0; p < 2; ++p) {
                    unpack2(w2[p][k], o[p][0], o[p][1]);
                    o[p][0] += __shfl_xor_sync(0xffffffffu, o[p][0], 16);
                    o[p][1] += __shfl_xor_sync(0xffffffffu, o[p][1], 16);
                    if (REL) {
                        unpack2(wr2[p][k], orr[p][0], orr[p][1]);
                        orr[p][0] += __shfl_xor_sync(0xffffffffu, orr[p][0], 16);
                        orr[p][1] += __shfl_xor_sync(0xffffffffu, orr[p][1], 16);
                    }
                }
                const int tb = half * 32 + l16;
                dst[k * TS + tb] = half ? o[1][0] : o[0][0];
                dst[k * TS + tb + 16] = half ? o[1][1] : o[0][1];
                if (REL) {
                    dst[(4 + k) * TS + tb] = half ? orr[1][0] : orr[0][0];
                    dst[(4 + k) * TS + tb + 16] = half ? orr[1][1] : orr[0][1];
                }
            }
            if (na.mono) {
                cp_async_wait<0>();
                __syncwarp();
                net_to_monomial<N0, N1, N2, CPD, REL>(s_cmat[wid], dst, TS);
                if (lane < CPD) dst[lane * TS] += ref[lane];   // the reference point goes back into the constant term
                __syncwarp();
                const float4* b4 = reinterpret_cast<const float4*>(dst);
                float4* d4 = reinterpret_cast<float4*>(gdst);
                for (int i = lane; i < (REL ? 8 : 4) * TS / 4; i += 32) d4[i] = b4[i];
                __syncwarp();
            }
            continue;
        }
        float w[NT][4], wr[REL ? NT : 1][4];
        int ia[NT];
        bool tv[NT];
        {
            const int n1 = sp.nfns[lev][1], n2 = sp.nfns[lev][2];
            const int64_t f0 = sp.fn_off[lev];
#pragma unroll
            for (int h = 0; h < NT; ++h) {
                const int t = lane + 32 * h;
                tv[h] = t < T;
                const int t2 = t % N2, t1 = (t / N2) % N1, t0 = t / (N2 * N1);
                ia[h] = tv[h] ? t0 : 0;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                const bool on = ndir > 0 && tv[h] && ((dm >> t) & 1ull);
                if (on) v = load_cp_sub<CPD>(na.ctrl_pts, f0 + ((int64_t)(a.y + t0) * n1 + (a.z + t1)) * n2 + (a.w + t2), ref, sub);
                w[h][0] = v.x; w[h][1] = v.y; w[h][2] = v.z; w[h][3] = v.w;
                if (REL) {
                    wr[h][0] = on ? v.x - ref[0] : 0.f; wr[h][1] = on ? v.y - ref[1] : 0.f;
                    wr[h][2] = on ? v.z - ref[2] : 0.f; wr[h][3] = on ? v.w - ref[3] : 0.f;
                }
            }
        }
        // packed accumulators of the rank-one fold (two window entries per lane), added to w / wr at the end
        f32x2 w2[4] = {0ull, 0ull, 0ull, 0ull}, wr2[REL ? 4 : 1];
        if (REL) {
#pragma unroll
            for (int k = 0; k < 4; ++k) wr2[k] = 0ull;
        }
        // tiled supports in chunks: [0, nsep) rank one, [nsep, ntile) full tiles
        const float* tile_src = split ? sp.full_tiles + (int64_t)f.z * TS : sp.tiles + (int64_t)e.z * TS;
        const int32_t* full_jm = split ? sp.full_jm + f.z : sp.supp_jm + soff + ndir;
        for (int c0 = 0; c0 < ntile; c0 += kNetChunk) {
            const int n = min(kNetChunk, ntile - c0);
            const int s_lo = min(c0, nsep), s_n = min(c0 + n, nsep) - s_lo;
            const int f_lo = max(c0, nsep) - nsep, f_n = n - s_n;
            __syncwarp();
            if (s_n > 0) {
                const float4* src = reinterpret_cast<const float4*>(sp.sep_vec + (int64_t)(f.x + s_lo) * 12);
                for (int i = lane; i < s_n * 3; i += 32) cp_async16(&sepv[0][0] + i, src + i);
                cp_async_commit();
                for (int i = lane; i < s_n; i += 32) cc[i] = load_cp_sub<CPD>(na.ctrl_pts, __ldg(sp.sep_jm + f.x + s_lo + i), ref, sub);
            }
            for (int i = lane; i < f_n; i += 32) cc[s_n + i] = load_cp_sub<CPD>(na.ctrl_pts, __ldg(full_jm + f_lo + i), ref, sub);
            cp_async_wait<0>();
            __syncwarp();
            if constexpr (NT == 2 && kSameBC && T == TS) {
                // both window entries of a lane share v1[b] * v2[c]: one FMUL2 forms the pair of tile entries, one FFMA2
                // per output component accumulates both (scalar second operand)
#pragma unroll 8
                for (int i = 0; i < s_n; ++i) {
                    const float* sv = reinterpret_cast<const float*>(&sepv[i][0]);
                    const float4 c = cc[i];
                    const float bc = sv[ib] * sv[ic];
                    const f32x2 x2 = fmul2(pack2(sv[ia[0]], sv[ia[1]]), pack2(bc, bc));
                    ffma2(w2[0], x2, pack2(c.x, c.x));
                    ffma2(w2[1], x2, pack2(c.y, c.y));
                    ffma2(w2[2], x2, pack2(c.z, c.z));
                    if (CPD == 4) ffma2(w2[3], x2, pack2(c.w, c.w));
                    if (REL) {
                        ffma2(wr2[0], x2, pack2(c.x - ref[0], c.x - ref[0]));
                        ffma2(wr2[1], x2, pack2(c.y - ref[1], c.y - ref[1]));
                        ffma2(wr2[2], x2, pack2(c.z - ref[2], c.z - ref[2]));
                        if (CPD == 4) ffma2(wr2[3], x2, pack2(c.w - ref[3], c.w - ref[3]));
                    }
                }
            } else {
#pragma unroll 4
                for (int i = 0; i < s_n; ++i) {
                    const float* sv = reinterpret_cast<const float*>(&sepv[i][0]);
                    const float4 c = cc[i];
                    const float bc = sv[ib] * sv[ic];
#pragma unroll
                    for (int h = 0; h < NT; ++h) {
                        float x;
                        if (kSameBC || h == 0) {
                            x = tv[h] ? sv[ia[h]] * bc : 0.f;
                        } else {
                            const int t = lane + 32 * h;
                            x = tv[h] ? sv[ia[h]] * sv[4 + (t / N2) % N1] * sv[8 + t % N2] : 0.f;
                        }
                        w[h][0] = fmaf(x, c.x, w[h][0]); w[h][1] = fmaf(x, c.y, w[h][1]); w[h][2] = fmaf(x, c.z, w[h][2]);
                        if (CPD == 4) w[h][3] = fmaf(x, c.w, w[h][3]);
                        if (REL) {
                            wr[h][0] = fmaf(x, c.x - ref[0], wr[h][0]); wr[h][1] = fmaf(x, c.y - ref[1], wr[h][1]);
                            wr[h][2] = fmaf(x, c.z - ref[2], wr[h][2]);
                            if (CPD == 4) wr[h][3] = fmaf(x, c.w - ref[3], wr[h][3]);
                        }
                    }
                }
            }
            const float* rows = tile_src + (int64_t)f_lo * TS + lane;
#pragma unroll 4
            for (int j = 0; j < f_n; ++j) {
                const float4 c = cc[s_n + j];
#pragma unroll
                for (int h = 0; h < NT; ++h) {
                    const float x = (lane + 32 * h < TS) ? __ldg(rows + (int64_t)j * TS + 32 * h) : 0.f;
                    w[h][0] = fmaf(x, c.x, w[h][0]); w[h][1] = fmaf(x, c.y, w[h][1]); w[h][2] = fmaf(x, c.z, w[h][2]);
                    if (CPD == 4) w[h][3] = fmaf(x, c.w, w[h][3]);
                    if (REL) {
                        wr[h][0] = fmaf(x, c.x - ref[0], wr[h][0]); wr[h][1] = fmaf(x, c.y - ref[1], wr[h][1]);
                        wr[h][2] = fmaf(x, c.z - ref[2], wr[h][2]);
                        if (CPD == 4) wr[h][3] = fmaf(x, c.w - ref[3], wr[h][3]);
                    }
                }
            }
        }
        if constexpr (NT == 2) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float lo, hi;
                unpack2(w2[k], lo, hi);
                w[0][k] += lo; w[1][k] += hi;
                if (REL) {
                    unpack2(wr2[k], lo, hi);
                    wr[0][k] += lo; wr[1][k] += hi;
                }
            }
        }
        float* dst = na.nets + (int64_t)slot * na.planes * TS;
        if (na.mono) {
            float* buf = &s_conv[wid][0][0];
#pragma unroll
            for (int h = 0; h < NT; ++h) {
                const int t = lane + 32 * h;
                if (t < TS) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        buf[k * TS + t] = w[h][k];
                        if (REL) buf[(4 + k) * TS + t] = wr[h][k];
                    }
                }
            }
            cp_async_wait<0>();
            __syncwarp();
            net_to_monomial<N0, N1, N2, CPD, REL>(s_cmat[wid], buf, TS);
            if (lane < CPD) buf[lane * TS] += ref[lane];   // the reference point goes back into the constant term
            __syncwarp();
            const float4* b4 = reinterpret_cast<const float4*>(buf);
            float4* d4 = reinterpret_cast<float4*>(dst);
            for (int i = lane; i < (REL ? 8 : 4) * TS / 4; i += 32) d4[i] = b4[i];
            __syncwarp();
            continue;
        }
#pragma unroll
        for (int h = 0; h < NT; ++h) {
            const int t = lane + 32 * h;
            if (t < TS) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    dst[k * TS + t] = w[h][k];
                    if (REL) dst[(4 + k) * TS + t] = wr[h][k];
                }
            }
        }
    }
}

// ---- evaluation against the nets --------------------------------------------------------------------------------
template <class SH, bool DER, bool MONO = false>
struct NetSmem {
    static constexpr int NPL = (DER && !MONO) ? 8 : 4;   // monomial nets: one set of planes serves every channel
    int4 rec[4][4];                  // ring of work-item records (lookahead 3)
    float knots[2][3][8];            // local knots of the item's cell, per dimension
    float rk[3][8];                  // reciprocal knot differences of the current item
    float net[2][NPL][SH::TS];       // the item's local control net (planes 4..7: relative), double buffered
    float4 prec[2][128];             // point records of a sub-batch (up to 4 per lane): (u0, u1, u2, caller-order index as bits)
    int pidx[3][256];                // light plan: caller indices of the points of the current item and the two after it (<= 256 each)
};

template <class SH, bool DER, bool MONO>
THB_DEV void prefetch_net(const thb_space_desc& sp, const float* __restrict__ nets, const Item& it, float (*knots)[8],
                          float (*net)[SH::TS]) {
    constexpr int NPL = (DER && !MONO) ? 8 : 4;
    const int lane = threadIdx.x;
    if (MONO) {   // (lo, 1 / half) of the cell in every dimension: floats 16, 17 of its row of sp.poly
        if (lane < 6) {
            const int dim = lane >> 1, m = lane & 1;
            const int c = dim == 0 ? it.c0 : (dim == 1 ? it.c1 : it.c2);
            if (dim < sp.ndim) cp_async4(&knots[dim][m], sp.poly + (int64_t)(sp.poly_off[it.lev][dim] + c) * 20 + 16 + m);
        }
    } else if (lane < 24) {
        const int dim = lane >> 3, m = lane & 7;
        const int P = dim == 0 ? SH::P0 : (dim == 1 ? SH::P1 : SH::P2);
        const int c = dim == 0 ? it.c0 : (dim == 1 ? it.c1 : it.c2);
        if (dim < sp.ndim && m < 2 * P) cp_async4(&knots[dim][m], sp.knots + sp.knot_off[it.lev][dim] + c + 1 + m);
    }
    const float4* src = reinterpret_cast<const float4*>(nets + (int64_t)it.slot * NPL * SH::TS);
    float4* dst = reinterpret_cast<float4*>(&net[0][0]);
#pragma unroll
    for (int i = 0; i < (NPL * SH::TS / 4 + 31) / 32; ++i) {
        const int j = lane + 32 * i;
        if (j < NPL * SH::TS / 4) cp_async16(dst + j, src + j);
    }
}

// One-dimensional Horner step helpers of the monomial evaluation.  A polynomial c0 + c1 s + ... + c_{N-1} s^{N-1}: value by
// the usual scheme; first derivative with respect to s as ((N-1) c_{N-1} s + (N-2) c_{N-2}) s + ... + c1, written with
// the per-step multipliers mult[j] so that value and derivative share the code: r = c_{top}; r = r * mult[j] + c_j.
//   value      : top = N-1, steps j = N-2 .. 0, mult = s
//   derivative : top = N-1, steps j = N-2 .. 1, mult_j = s * (j+1)/j ... (see horner_mults)
template <int N>
THB_DEV void horner_mults(float s, bool der, float* mult) {
    // der: d/ds sum c_m s^m = sum_{m>=1} m c_m s^{m-1}; nested: r_{N-1} = c_{N-1}; r_j = r_{j+1} * (s * (j+1)/j) + c_j for j >= 1
    // gives r_1 = sum_m (m/1) c_m s^{m-1}: exactly the derivative.
#pragma unroll
    for (int j = 0; j < N - 1; ++j) mult[j] = der ? (j >= 1 ? s * (float(j + 1) / float(j)) : 0.f) : s;
}

// PPT*32 points of an item against the MONOMIAL net M (k_cell_nets with na.mono):
//     out_k = sum_{m0,m1,m2} M[k][m0][m1][m2] s0^m0 s1^m1 s2^m2,  s = (u - lo) / half - 1
// Nested Horner: over m1 on pairs of m2 (FFMA2 with a scalar multiplier), over m0 on the same pairs, then over m2.
// 30 FFMA2 + 3 FFMA per output component for cubic 3-D against 45 FFMA2 for the B-spline form -- and no Cox-de Boor.
// Derivative channels use the derivative polynomial in the flagged dimensions (times 1 / half of that dimension).
template <class SH, int PPT, int CPD, bool DER>
THB_DEV void net_points_mono(const Args& ar, const Item& it, int base, NetSmem<SH, DER, true>& sm, int tb, int pb) {
    constexpr int N0 = SH::N0, N1 = SH::N1, N2 = SH::N2;
    const int lane = threadIdx.x;
    int idx[PPT];
    bool valid[PPT];
    float s[PPT][3];
    float ih[3];
#pragma unroll
    for (int dim = 0; dim < 3; ++dim) ih[dim] = sm.knots[tb][dim][1];
#pragma unroll
    for (int q = 0; q < PPT; ++q) {
        valid[q] = base + q * 32 + lane < it.count;
        const float4 rec = sm.prec[pb][q * 32 + lane];
        idx[q] = ar.rows_sorted ? it.begin + base + q * 32 + lane : __float_as_int(rec.w);
        s[q][0] = fmaf(rec.x - sm.knots[tb][0][0], ih[0], -1.0f);
        s[q][1] = fmaf(rec.y - sm.knots[tb][1][0], ih[1], -1.0f);
        s[q][2] = N2 > 1 ? fmaf(rec.z - sm.knots[tb][2][0], ih[2], -1.0f) : 0.0f;
    }
    for (int chi = 0; chi < ar.nch; ++chi) {
        const int ch = ar.channels[chi];
        constexpr int pl = 0;  // one set of planes: built from relative control points, reference point in the constant term
        const bool d0 = DER && (ch & 1), d1 = DER && (ch & 2), d2 = DER && (ch & 4);
        const float scale = (d0 ? ih[0] : 1.0f) * (d1 ? ih[1] : 1.0f) * (d2 ? ih[2] : 1.0f);
        float m0[PPT][N0 > 1 ? N0 - 1 : 1], m1[PPT][N1 > 1 ? N1 - 1 : 1], m2[PPT][N2 > 1 ? N2 - 1 : 1];
#pragma unroll
        for (int q = 0; q < PPT; ++q) {
            horner_mults<N0>(s[q][0], d0, m0[q]);
            horner_mults<N1>(s[q][1], d1, m1[q]);
            horner_mults<N2>(s[q][2], d2, m2[q]);
        }
        // lowest index reached by the Horner scheme of each dimension: 0 (value) or 1 (derivative)
        const int e0 = d0 ? 1 : 0, e1 = d1 ? 1 : 0, e2 = d2 ? 1 : 0;
        float acc[PPT][4];
        if constexpr (N2 == 4) {
#pragma unroll
            for (int k = 0; k < CPD; ++k) {
                const ulonglong2* col = reinterpret_cast<const ulonglong2*>(sm.net[tb][pl + k]);   // col[m0 * N1 + m1] = M[m0][m1][0..3]
                f32x2 Q[PPT][2];
#pragma unroll
                for (int a = N0 - 1; a >= 0; --a) {
                    if (a < e0) break;
                    ulonglong2 cv[N1];
#pragma unroll
                    for (int b = 0; b < N1; ++b) cv[b] = col[a * N1 + b];
#pragma unroll
                    for (int q = 0; q < PPT; ++q) {
                        f32x2 E0 = cv[N1 - 1].x, E1 = cv[N1 - 1].y;
#pragma unroll
                        for (int b = N1 - 2; b >= 0; --b) {
                            if (b < e1) break;
                            const f32x2 x = pack2(m1[q][b], m1[q][b]);
                            f32x2 t0 = cv[b].x, t1 = cv[b].y;
                            ffma2(t0, E0, x);
                            ffma2(t1, E1, x);
                            E0 = t0; E1 = t1;
                        }
                        if (a == N0 - 1) {
                            Q[q][0] = E0; Q[q][1] = E1;
                        } else {
                            const f32x2 x = pack2(m0[q][a], m0[q][a]);
                            ffma2(E0, Q[q][0], x);
                            ffma2(E1, Q[q][1], x);
                            Q[q][0] = E0; Q[q][1] = E1;
                        }
                    }
                }
#pragma unroll
                for (int q = 0; q < PPT; ++q) {
                    float c0, c1, c2, c3;
                    unpack2(Q[q][0], c0, c1);
                    unpack2(Q[q][1], c2, c3);
                    float r = c3;
                    r = fmaf(r, m2[q][2], c2);
                    r = fmaf(r, m2[q][1], c1);
                    if (!d2) r = fmaf(r, m2[q][0], c0);
                    acc[q][k] = DER ? r * scale : r;
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < CPD; ++k) {
                const float* col = sm.net[tb][pl + k];
#pragma unroll
                for (int q = 0; q < PPT; ++q) {
                    float ra = 0.0f;
#pragma unroll
                    for (int a = N0 - 1; a >= 0; --a) {
                        if (a < e0) break;
                        float rb = 0.0f;
#pragma unroll
                        for (int b = N1 - 1; b >= 0; --b) {
                            if (b < e1) break;
                            float rc = col[(a * N1 + b) * N2 + N2 - 1];
#pragma unroll
                            for (int c = N2 - 2; c >= 0; --c) {
                                if (c < e2) break;
                                rc = fmaf(rc, m2[q][c], col[(a * N1 + b) * N2 + c]);
                            }
                            if (N2 == 1 && d2) rc = 0.0f;
                            rb = b == N1 - 1 ? rc : fmaf(rb, m1[q][b], rc);
                        }
                        ra = a == N0 - 1 ? rb : fmaf(ra, m0[q][a], rb);
                    }
                    acc[q][k] = DER ? ra * scale : ra;
                }
            }
        }
#pragma unroll
        for (int q = 0; q < PPT; ++q) {
            if (valid[q]) {
                float* o = ar.out[chi] + (int64_t)idx[q] * CPD;
#pragma unroll
                for (int k = 0; k < CPD; ++k) o[k] = acc[q][k];
            }
        }
    }
}

// PPT*32 points of an item:  out_k = sum_a A[a] sum_c C[c] sum_b B[b] * W[k][a][b][c]
template <class SH, int PPT, int CPD, bool DER>
THB_DEV void net_points(const Args& ar, const Item& it, int base, NetSmem<SH, DER>& sm, int tb, int pb) {
    constexpr int N0 = SH::N0, N1 = SH::N1, N2 = SH::N2;
    const int lane = threadIdx.x;
    int idx[PPT];
    bool valid[PPT];
    BasesRK<SH, DER> bs[PPT];
    // the cell's knots and reciprocal differences are warp-uniform: four LDS.128 per dimension instead of 2p + p(p+1)/2
    // scalar loads per point
    float K[3][8], RK[3][8];
#pragma unroll
    for (int dim = 0; dim < 3; ++dim) {
        const float4 k0 = *reinterpret_cast<const float4*>(&sm.knots[tb][dim][0]), k1 = *reinterpret_cast<const float4*>(&sm.knots[tb][dim][4]);
        const float4 r0 = *reinterpret_cast<const float4*>(&sm.rk[dim][0]), r1 = *reinterpret_cast<const float4*>(&sm.rk[dim][4]);
        K[dim][0] = k0.x; K[dim][1] = k0.y; K[dim][2] = k0.z; K[dim][3] = k0.w; K[dim][4] = k1.x; K[dim][5] = k1.y; K[dim][6] = k1.z; K[dim][7] = k1.w;
        RK[dim][0] = r0.x; RK[dim][1] = r0.y; RK[dim][2] = r0.z; RK[dim][3] = r0.w; RK[dim][4] = r1.x; RK[dim][5] = r1.y; RK[dim][6] = r1.z; RK[dim][7] = r1.w;
    }
#pragma unroll
    for (int q = 0; q < PPT; ++q) {
        valid[q] = base + q * 32 + lane < it.count;
        const float4 rec = sm.prec[pb][q * 32 + lane];
        idx[q] = ar.rows_sorted ? it.begin + base + q * 32 + lane : __float_as_int(rec.w);
        const float u[3] = {rec.x, rec.y, rec.z};
        bs[q].eval(u, K, RK);
    }
    for (int chi = 0; chi < ar.nch; ++chi) {
        const int ch = ar.channels[chi];
        const int pl = (DER && ch != 0) ? 4 : 0;  // derivative channels use the relative planes
        float fa[PPT][N0], fb[PPT][N1], fc[PPT][N2];
#pragma unroll
        for (int q = 0; q < PPT; ++q) bs[q].select(ch, fa[q], fb[q], fc[q]);
        float acc[PPT][4];
        if constexpr (N2 == 4) {
            // Stage 1 (the (p+1)^3 one) is e[a][h] += B[b] * W[a][b][2h..2h+1]: FFMA2 with a scalar multiplier shared by
            // consecutive instructions (operand reuse cache), so each FFMA2 fetches two register pairs -- an FMA stream
            // with three fresh operands per instruction is register-file-bound at ~64 % of the FP32 peak (scripts/probe).
            f32x2 fc2[PPT][2];
#pragma unroll
            for (int q = 0; q < PPT; ++q) {
                fc2[q][0] = pack2(fc[q][0], fc[q][1]);
                fc2[q][1] = pack2(fc[q][2], fc[q][3]);
            }
#pragma unroll
            for (int k = 0; k < CPD; ++k) {
                const ulonglong2* col = reinterpret_cast<const ulonglong2*>(sm.net[tb][pl + k]);
                f32x2 e[PPT][N0][2];
#pragma unroll
                for (int b = 0; b < N1; ++b) {
                    ulonglong2 cv[N0];
#pragma unroll
                    for (int a = 0; a < N0; ++a) cv[a] = col[a * N1 + b];
#pragma unroll
                    for (int q = 0; q < PPT; ++q) {
                        const f32x2 bb = pack2(fb[q][b], fb[q][b]);
#pragma unroll
                        for (int a = 0; a < N0; ++a) {
                            if (b == 0) {
                                e[q][a][0] = fmul2(bb, cv[a].x);
                                e[q][a][1] = fmul2(bb, cv[a].y);
                            } else {
                                ffma2(e[q][a][0], bb, cv[a].x);
                                ffma2(e[q][a][1], bb, cv[a].y);
                            }
                        }
                    }
                }
#pragma unroll
                for (int q = 0; q < PPT; ++q) {
                    f32x2 f = 0ull;
#pragma unroll
                    for (int a = 0; a < N0; ++a) {
                        f32x2 g2 = fmul2(fc2[q][0], e[q][a][0]);
                        ffma2(g2, fc2[q][1], e[q][a][1]);
                        ffma2(f, pack2(fa[q][a], fa[q][a]), g2);
                    }
                    acc[q][k] = hsum2(f);
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < CPD; ++k) {
                const float* col = sm.net[tb][pl + k];
                float f[PPT];
#pragma unroll
                for (int q = 0; q < PPT; ++q) f[q] = 0.0f;
#pragma unroll
                for (int a = 0; a < N0; ++a) {
                    float e[PPT];
#pragma unroll
                    for (int q = 0; q < PPT; ++q) e[q] = 0.0f;
#pragma unroll
                    for (int b = 0; b < N1; ++b) {
                        float cv[N2];
#pragma unroll
                        for (int c = 0; c < N2; ++c) cv[c] = col[(a * N1 + b) * N2 + c];
#pragma unroll
                        for (int q = 0; q < PPT; ++q) {
                            float dd = fc[q][0] * cv[0];
#pragma unroll
                            for (int c = 1; c < N2; ++c) dd = fmaf(fc[q][c], cv[c], dd);
                            e[q] = fmaf(fb[q][b], dd, e[q]);
                        }
                    }
#pragma unroll
                    for (int q = 0; q < PPT; ++q) f[q] = fmaf(fa[q][a], e[q], f[q]);
                }
#pragma unroll
                for (int q = 0; q < PPT; ++q) acc[q][k] = f[q];
            }
        }
#pragma unroll
        for (int q = 0; q < PPT; ++q) {
            if (valid[q]) {
                float* o = ar.out[chi] + (int64_t)idx[q] * CPD;
#pragma unroll
                for (int k = 0; k < CPD; ++k) o[k] = acc[q][k];
            }
        }
    }
}

// Value + the three first derivatives in ONE pass over the monomial net (channels 0, 1, 2, 4 in this order: the outputs of
// BASELINE config 5).  The four polynomials share their Horner stages:
//   stage 1 over m1:  E  (value in s1)            -> recurrences over m0:  Qv (value in s0), Q0 (derivative in s0)
//                     E' (derivative in s1)       -> recurrence  over m0:  Q1 (value in s0)
//   finals over m2 :  value = H(Qv), d/du2 = H'(Qv), d/du0 = H(Q0) / half0, d/du1 = H(Q1) / half1
// 56 FFMA2 + 11 FFMA per output component for cubic 3-D instead of 4 x 33 when the channels are evaluated one by one.
template <class SH, int PPT, int CPD>
THB_DEV void net_points_mono_grad4(const Args& ar, const Item& it, int base, NetSmem<SH, true, true>& sm, int tb, int pb) {
    constexpr int N0 = SH::N0, N1 = SH::N1;
    static_assert(SH::N2 == 4 && N0 >= 2 && N1 >= 2, "cubic last dimension");
    const int lane = threadIdx.x;
    int idx[PPT];
    bool valid[PPT];
    float s[PPT][3], ih[3];
#pragma unroll
    for (int dim = 0; dim < 3; ++dim) ih[dim] = sm.knots[tb][dim][1];
#pragma unroll
    for (int q = 0; q < PPT; ++q) {
        valid[q] = base + q * 32 + lane < it.count;
        const float4 rec = sm.prec[pb][q * 32 + lane];
        idx[q] = ar.rows_sorted ? it.begin + base + q * 32 + lane : __float_as_int(rec.w);
        s[q][0] = fmaf(rec.x - sm.knots[tb][0][0], ih[0], -1.0f);
        s[q][1] = fmaf(rec.y - sm.knots[tb][1][0], ih[1], -1.0f);
        s[q][2] = fmaf(rec.z - sm.knots[tb][2][0], ih[2], -1.0f);
    }
    float m0v[PPT][N0 - 1], m0d[PPT][N0 - 1], m1v[PPT][N1 - 1], m1d[PPT][N1 - 1], m2v[PPT][3], m2d[PPT][3];
#pragma unroll
    for (int q = 0; q < PPT; ++q) {
        horner_mults<N0>(s[q][0], false, m0v[q]); horner_mults<N0>(s[q][0], true, m0d[q]);
        horner_mults<N1>(s[q][1], false, m1v[q]); horner_mults<N1>(s[q][1], true, m1d[q]);
        horner_mults<4>(s[q][2], false, m2v[q]);  horner_mults<4>(s[q][2], true, m2d[q]);
    }
    float acc[4][PPT][4];
#pragma unroll
    for (int k = 0; k < CPD; ++k) {
        const ulonglong2* col = reinterpret_cast<const ulonglong2*>(sm.net[tb][k]);
        f32x2 Qv[PPT][2], Q0[PPT][2], Q1[PPT][2];
#pragma unroll
        for (int a = N0 - 1; a >= 0; --a) {
            ulonglong2 cv[N1];
#pragma unroll
            for (int b = 0; b < N1; ++b) cv[b] = col[a * N1 + b];
#pragma unroll
            for (int q = 0; q < PPT; ++q) {
                f32x2 E0 = cv[N1 - 1].x, E1 = cv[N1 - 1].y, D0 = cv[N1 - 1].x, D1 = cv[N1 - 1].y;
#pragma unroll
                for (int b = N1 - 2; b >= 0; --b) {
                    const f32x2 x = pack2(m1v[q][b], m1v[q][b]);
                    f32x2 t0 = cv[b].x, t1 = cv[b].y;
                    ffma2(t0, E0, x); ffma2(t1, E1, x);
                    E0 = t0; E1 = t1;
                    if (b >= 1) {
                        const f32x2 y = pack2(m1d[q][b], m1d[q][b]);
                        f32x2 u0 = cv[b].x, u1 = cv[b].y;
                        ffma2(u0, D0, y); ffma2(u1, D1, y);
                        D0 = u0; D1 = u1;
                    }
                }
                if (a == N0 - 1) {
                    Qv[q][0] = E0; Qv[q][1] = E1; Q0[q][0] = E0; Q0[q][1] = E1; Q1[q][0] = D0; Q1[q][1] = D1;
                } else {
                    const f32x2 x = pack2(m0v[q][a], m0v[q][a]);
                    f32x2 t0 = E0, t1 = E1, u0 = D0, u1 = D1;
                    ffma2(t0, Qv[q][0], x); ffma2(t1, Qv[q][1], x);
                    ffma2(u0, Q1[q][0], x); ffma2(u1, Q1[q][1], x);
                    Qv[q][0] = t0; Qv[q][1] = t1; Q1[q][0] = u0; Q1[q][1] = u1;
                    if (a >= 1) {
                        const f32x2 y = pack2(m0d[q][a], m0d[q][a]);
                        f32x2 w0 = E0, w1 = E1;
                        ffma2(w0, Q0[q][0], y); ffma2(w1, Q0[q][1], y);
                        Q0[q][0] = w0; Q0[q][1] = w1;
                    }
                }
            }
        }
#pragma unroll
        for (int q = 0; q < PPT; ++q) {
            auto hv = [&](const f32x2 (&Q)[2]) {
                float c0, c1, c2, c3;
                unpack2(Q[0], c0, c1); unpack2(Q[1], c2, c3);
                float r = fmaf(c3, m2v[q][2], c2);
                r = fmaf(r, m2v[q][1], c1);
                return fmaf(r, m2v[q][0], c0);
            };
            float c0, c1, c2, c3;
            unpack2(Qv[q][0], c0, c1); unpack2(Qv[q][1], c2, c3);
            float rd = fmaf(c3, m2d[q][2], c2);
            rd = fmaf(rd, m2d[q][1], c1);
            acc[0][q][k] = hv(Qv[q]);
            acc[1][q][k] = hv(Q0[q]) * ih[0];
            acc[2][q][k] = hv(Q1[q]) * ih[1];
            acc[3][q][k] = rd * ih[2];
        }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int q = 0; q < PPT; ++q) {
            if (valid[q]) {
                float* o = ar.out[c] + (int64_t)idx[q] * CPD;
#pragma unroll
                for (int k = 0; k < CPD; ++k) o[k] = acc[c][q][k];
            }
        }
}

// Persistent, warp-autonomous (one warp per CTA, no CTA barrier): work items from an atomic cursor; while item i is
// computed the records of items i+1, i+2 are in flight, and the points of the next sub-batch plus the knots / net of
// item i+1 stream into the other half of the shared-memory buffers (cp.async).
#ifndef THB_NET_MINB_P4
#define THB_NET_MINB_P4 16  // four points per lane still fit 128 registers (106 used)
#endif
template <int N0, int N1, int N2, int PPT, int CPD, bool DER, bool MONO, bool GRAD4 = false>
__global__ void __launch_bounds__(kThreads, GRAD4 ? 12 : (PPT >= 4 ? (DER ? 12 : THB_NET_MINB_P4) : (DER ? THB_NET_MINB_DER : THB_NET_MINB))) k_net_eval(const __grid_constant__ thb_space_desc sp, const __grid_constant__ Args ar,
                                                                    const float* __restrict__ nets) {
    using SH = Shp<N0, N1, N2>;
    constexpr int SB = 32 * PPT;
    __shared__ __align__(16) NetSmem<SH, DER, MONO> sm;
    const int lane = threadIdx.x;
    const int nitems = ar.counters[0];
    int32_t* const cursor = ar.counters + 1;
    const int d = sp.ndim;

    // item records run three ahead (ring of four), the permutation of a light plan two items ahead (ring of three): the
    // indices of an item are in shared memory a whole item before its first points are gathered
    int c4 = 0;
    if (lane == 0) c4 = atomicAdd(cursor, 4);
    const int id0 = __shfl_sync(0xffffffffu, c4, 0);
    if (id0 >= nitems) return;
    int id1 = id0 + 1, id2 = id0 + 2, id3 = id0 + 3;
    Item it = load_item(sp, ar.items, id0);
    if (lane < 4 && id1 < nitems) cp_async16(&sm.rec[1][lane], ar.items + (int64_t)id1 * 4 + lane);
    if (lane < 4 && id2 < nitems) cp_async16(&sm.rec[2][lane], ar.items + (int64_t)id2 * 4 + lane);
    if (lane < 4 && id3 < nitems) cp_async16(&sm.rec[3][lane], ar.items + (int64_t)id3 * 4 + lane);
    prefetch_net<SH, DER, MONO>(sp, nets, it, sm.knots[0], sm.net[0]);
    const bool light = ar.perm != nullptr;   // light plan: parameters gathered through the permutation
    auto fetch_perm = [&](const Item& x, int* dst) {   // caller indices of an item's points (at most 256)
        for (int i = lane; i < x.count; i += 32) cp_async4(dst + i, ar.perm + x.begin + i);
    };
    if (light) fetch_perm(it, sm.pidx[0]);
    else prefetch_points<PPT>(ar, it.begin, it.count, 0, sm.prec[0]);
    cp_async_commit();
    int nn_id = 0;
    if (lane == 0) nn_id = atomicAdd(cursor, 1);
    int ring = 0, tb = 0, pb = 0, pi = 0;   // pi: buffer of sm.pidx that holds the current item's indices
    cp_async_wait<0>();
    __syncwarp();
    if (light) {
        prefetch_points_gather<PPT>(ar, sm.pidx[0], it.count, 0, sm.prec[0]);
        if (id1 < nitems) {
            const int4* r = sm.rec[1];
            fetch_perm(unpack_item(r[0], r[1], r[2], r[3]), sm.pidx[1]);
        }
        if (id2 < nitems) {
            const int4* r = sm.rec[2];
            fetch_perm(unpack_item(r[0], r[1], r[2], r[3]), sm.pidx[2]);
        }
        cp_async_commit();
        cp_async_wait<0>();
        __syncwarp();
    }

    for (;;) {
        if (!MONO && lane < 24) {  // reciprocal knot differences of this cell (from the staged knots)
            const int dim = lane >> 3, m = lane & 7;
            const int P = dim == 0 ? SH::P0 : (dim == 1 ? SH::P1 : SH::P2);
            if (dim < d && m < P * (P + 1) / 2) {
                int j = 1, r = m;
                while (r >= j) { r -= j; ++j; }
                sm.rk[dim][m] = 1.0f / (sm.knots[tb][dim][P + r] - sm.knots[tb][dim][P - j + r]);
            }
        }
        __syncwarp();
        const bool nx_valid = id1 < nitems;
        for (int base = 0; base < it.count; base += SB) {
            if (base + SB < it.count) {
                if (light) prefetch_points_gather<PPT>(ar, sm.pidx[pi], it.count, base + SB, sm.prec[pb ^ 1]);
                else prefetch_points<PPT>(ar, it.begin, it.count, base + SB, sm.prec[pb ^ 1]);
            } else if (nx_valid) {
                const int4* r = sm.rec[(ring + 1) & 3];
                const Item nx = unpack_item(r[0], r[1], r[2], r[3]);
                if (light) prefetch_points_gather<PPT>(ar, sm.pidx[pi == 2 ? 0 : pi + 1], nx.count, 0, sm.prec[pb ^ 1]);   // its indices arrived an item ago
                else prefetch_points<PPT>(ar, nx.begin, nx.count, 0, sm.prec[pb ^ 1]);
                prefetch_net<SH, DER, MONO>(sp, nets, nx, sm.knots[tb ^ 1], sm.net[tb ^ 1]);
            }
            cp_async_commit();
            if constexpr (GRAD4) {
                if (it.count - base <= 32) net_points_mono_grad4<SH, 1, CPD>(ar, it, base, sm, tb, pb);
                else net_points_mono_grad4<SH, PPT, CPD>(ar, it, base, sm, tb, pb);
            } else if constexpr (MONO) {
                // four (or two) points per lane amortise the broadcast LDS.128 of the net: with two the shared-memory pipe
                // (two wavefronts per broadcast LDS.128, one pipe per SM) is as busy as the FMA pipes
                if (PPT >= 2 && it.count - base <= 32) net_points_mono<SH, 1, CPD, DER>(ar, it, base, sm, tb, pb);
                else if (PPT >= 4 && it.count - base <= 64) net_points_mono<SH, 2, CPD, DER>(ar, it, base, sm, tb, pb);
                else net_points_mono<SH, PPT, CPD, DER>(ar, it, base, sm, tb, pb);
            } else {
                if (PPT == 2 && it.count - base <= 32) net_points<SH, 1, CPD, DER>(ar, it, base, sm, tb, pb);
                else net_points<SH, PPT, CPD, DER>(ar, it, base, sm, tb, pb);
            }
            cp_async_wait<0>();
            __syncwarp();
            pb ^= 1;
        }
        if (!nx_valid) break;
        ring = (ring + 1) & 3;
        {
            const int4* r = sm.rec[ring];
            it = unpack_item(r[0], r[1], r[2], r[3]);
        }
        tb ^= 1;
        pi = pi == 2 ? 0 : pi + 1;
        const int id4 = __shfl_sync(0xffffffffu, nn_id, 0);
        id1 = id2;
        id2 = id3;
        id3 = id4;
        if (lane < 4 && id3 < nitems) cp_async16(&sm.rec[(ring + 3) & 3][lane], ar.items + (int64_t)id3 * 4 + lane);
        if (light && id2 < nitems) {   // indices of the item two ahead (its record arrived an item ago) into the buffer the
            const int4* r = sm.rec[(ring + 2) & 3];   // item just finished used; they are gathered from at the end of the next item
            fetch_perm(unpack_item(r[0], r[1], r[2], r[3]), sm.pidx[pi == 0 ? 2 : pi - 1]);
        }
        cp_async_commit();
        if (lane == 0) nn_id = atomicAdd(cursor, 1);
    }
}

// ---- backward through the nets -------------------------------------------------------------------------------------
// grad_ctrl_pts = A^T grad_out with A the value-channel basis matrix (THBEval.backward, THB/torch_funcs.py:35-47).  With
// out = <tp(u), W_cell> the chain rule splits the same way as the forward pass:
//   k_net_backward : gW_cell[t][k] += sum_{points of the item} tp(u)[t] * g[k]      (per work item, atomics into the
//                    per-cell net-gradient workspace: T*cp_dim per ITEM instead of S*cp_dim per point)
//   k_cell_unfold  : grad_cp[own t] += gW[t],   grad_cp[Jm_s] += <tile_s, gW>          (per cell, the transposed fold)
struct NetBwdSmem {
    int4 rec[3][4];
    float knots[2][3][8];
    float rk[3][8];
    float4 prec[2][64];
    float4 fa[64], fb[64], fc[64], g[64];  // 1-D basis values (padded to 4) and output gradient of a sub-batch
};

// MONO: the per-item sums are taken in the MONOMIAL basis of the cell (powers of the cell-local coordinates instead of
// Cox-de Boor: ~12 instead of ~75 instructions per point in phase 1); k_cell_unfold turns the cell's gM into gW with the
// transposed span polynomials before the fold.  The reference point of the forward's relative form drops out of the
// gradient: d/d(ref) = gM[0] - sum_t gW[t] = 0 because the rows of C sum to (1, 0, 0, 0).
template <int N0, int N1, int N2, int CPD, bool MONO>
__global__ void __launch_bounds__(kThreads, 16) k_net_backward(const __grid_constant__ thb_space_desc sp, const __grid_constant__ Args ar,
                                                              float* __restrict__ gnets) {
    using SH = Shp<N0, N1, N2>;
    constexpr int TS = SH::TS, BC = N1 * N2;
    static_assert(BC <= 16 && N0 <= 4, "lane = (half, b, c) owns the entries a = 0..3 of its (b, c)");
    constexpr int SB = 64;
    __shared__ __align__(16) NetBwdSmem sm;
    const int lane = threadIdx.x;
    const int half = lane >> 4, l16 = lane & 15;
    const bool live = l16 < BC;
    const int jb = live ? l16 / N2 : 0, jc = live ? l16 % N2 : 0;
    const int nitems = ar.counters[0];
    int32_t* const cursor = ar.counters + 1;
    const int d = sp.ndim;

    int c3 = 0;
    if (lane == 0) c3 = atomicAdd(cursor, 3);
    const int id0 = __shfl_sync(0xffffffffu, c3, 0);
    if (id0 >= nitems) return;
    int id1 = id0 + 1, id2 = id0 + 2, cur_id = id0;
    Item it = load_item(sp, ar.items, id0);
    if (lane < 4 && id1 < nitems) cp_async16(&sm.rec[1][lane], ar.items + (int64_t)id1 * 4 + lane);
    if (lane < 4 && id2 < nitems) cp_async16(&sm.rec[2][lane], ar.items + (int64_t)id2 * 4 + lane);
    auto prefetch_knots = [&](const Item& x, float (*knots)[8]) {
        if (MONO) {   // (lo, 1 / half) of the cell in every dimension
            if (lane < 6) {
                const int dim = lane >> 1, m = lane & 1;
                const int c = dim == 0 ? x.c0 : (dim == 1 ? x.c1 : x.c2);
                if (dim < d) cp_async4(&knots[dim][m], sp.poly + (int64_t)(sp.poly_off[x.lev][dim] + c) * 20 + 16 + m);
            }
        } else if (lane < 24) {
            const int dim = lane >> 3, m = lane & 7;
            const int P = dim == 0 ? SH::P0 : (dim == 1 ? SH::P1 : SH::P2);
            const int c = dim == 0 ? x.c0 : (dim == 1 ? x.c1 : x.c2);
            if (dim < d && m < 2 * P) cp_async4(&knots[dim][m], sp.knots + sp.knot_off[x.lev][dim] + c + 1 + m);
        }
    };
    prefetch_knots(it, sm.knots[0]);
    prefetch_points<2>(ar, it.begin, it.count, 0, sm.prec[0]);
    cp_async_commit();
    int nn_id = 0;
    if (lane == 0) nn_id = atomicAdd(cursor, 1);
    int ring = 0, tb = 0, pb = 0;
    cp_async_wait<0>();
    __syncwarp();

    for (;;) {
        if (!MONO && lane < 24) {
            const int dim = lane >> 3, m = lane & 7;
            const int P = dim == 0 ? SH::P0 : (dim == 1 ? SH::P1 : SH::P2);
            if (dim < d && m < P * (P + 1) / 2) {
                int j = 1, r = m;
                while (r >= j) { r -= j; ++j; }
                sm.rk[dim][m] = 1.0f / (sm.knots[tb][dim][P + r] - sm.knots[tb][dim][P - j + r]);
            }
        }
        __syncwarp();
        f32x2 w2[2][4];
#pragma unroll
        for (int p = 0; p < 2; ++p)
#pragma unroll
            for (int k = 0; k < 4; ++k) w2[p][k] = 0ull;
        const bool nx_valid = id1 < nitems;
        for (int base = 0; base < it.count; base += SB) {
            if (base + SB < it.count) {
                prefetch_points<2>(ar, it.begin, it.count, base + SB, sm.prec[pb ^ 1]);
            } else if (nx_valid) {
                const int4* r = sm.rec[(ring + 1) % 3];
                const Item nx = unpack_item(r[0], r[1], r[2], r[3]);
                prefetch_points<2>(ar, nx.begin, nx.count, 0, sm.prec[pb ^ 1]);
                prefetch_knots(nx, sm.knots[tb ^ 1]);
            }
            cp_async_commit();
            // phase 1: lane = point (two per lane): 1-D bases and the point's output gradient -> shared memory
            const int npts = min(SB, it.count - base);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int pi = q * 32 + lane;
                const float4 rec = sm.prec[pb][pi];
                const float u[3] = {rec.x, rec.y, rec.z};
                float a4[4] = {0.f, 0.f, 0.f, 0.f}, b4[4] = {0.f, 0.f, 0.f, 0.f}, c4[4] = {0.f, 0.f, 0.f, 0.f}, g4[4] = {0.f, 0.f, 0.f, 0.f};
                if constexpr (MONO) {   // powers of the cell-local coordinates
                    const float s0 = fmaf(u[0] - sm.knots[tb][0][0], sm.knots[tb][0][1], -1.0f);
                    const float s1 = fmaf(u[1] - sm.knots[tb][1][0], sm.knots[tb][1][1], -1.0f);
                    const float s2 = N2 > 1 ? fmaf(u[2] - sm.knots[tb][2][0], sm.knots[tb][2][1], -1.0f) : 0.0f;
                    a4[0] = b4[0] = c4[0] = 1.0f;
#pragma unroll
                    for (int i = 1; i < N0; ++i) a4[i] = a4[i - 1] * s0;
#pragma unroll
                    for (int i = 1; i < N1; ++i) b4[i] = b4[i - 1] * s1;
#pragma unroll
                    for (int i = 1; i < N2; ++i) c4[i] = c4[i - 1] * s2;
                } else {
                    BasesRK<SH, false> bs;
                    bs.eval(u, sm.knots[tb], sm.rk);
#pragma unroll
                    for (int i = 0; i < N0; ++i) a4[i] = bs.n0[i];
#pragma unroll
                    for (int i = 0; i < N1; ++i) b4[i] = bs.n1[i];
#pragma unroll
                    for (int i = 0; i < N2; ++i) c4[i] = bs.n2[i];
                }
                if (pi < npts) {
                    const int64_t idx = ar.rows_sorted ? it.begin + base + pi : __float_as_int(rec.w);
#pragma unroll
                    for (int k = 0; k < CPD; ++k) g4[k] = __ldg(ar.grad_out + idx * CPD + k);
                }
                sm.fa[pi] = make_float4(a4[0], a4[1], a4[2], a4[3]);
                sm.fb[pi] = make_float4(b4[0], b4[1], b4[2], b4[3]);
                sm.fc[pi] = make_float4(c4[0], c4[1], c4[2], c4[3]);
                sm.g[pi] = make_float4(g4[0], g4[1], g4[2], g4[3]);   // zero for the slots beyond the item's points
            }
            __syncwarp();
            // phase 2: lane = (half, b, c); the halves take alternate points
            const float* fbp = reinterpret_cast<const float*>(sm.fb) + jb;
            const float* fcp = reinterpret_cast<const float*>(sm.fc) + jc;
            const int nloop = (npts + 1) >> 1;
#pragma unroll 4
            for (int j = 0; j < nloop; ++j) {
                const int pi = 2 * j + half;   // pi == npts (odd count) reads a zero gradient
                const float4 av = sm.fa[pi];
                const float4 gv = sm.g[pi];
                const float bc = fbp[pi * 4] * fcp[pi * 4];
                const f32x2 bc2 = pack2(bc, bc);
                const f32x2 x01 = fmul2(pack2(av.x, av.y), bc2), x23 = fmul2(pack2(av.z, av.w), bc2);
                ffma2(w2[0][0], x01, pack2(gv.x, gv.x)); ffma2(w2[1][0], x23, pack2(gv.x, gv.x));
                ffma2(w2[0][1], x01, pack2(gv.y, gv.y)); ffma2(w2[1][1], x23, pack2(gv.y, gv.y));
                ffma2(w2[0][2], x01, pack2(gv.z, gv.z)); ffma2(w2[1][2], x23, pack2(gv.z, gv.z));
                if (CPD == 4) { ffma2(w2[0][3], x01, pack2(gv.w, gv.w)); ffma2(w2[1][3], x23, pack2(gv.w, gv.w)); }
            }
            cp_async_wait<0>();
            __syncwarp();
            pb ^= 1;
        }
        // the item's contribution to its cell's net gradient: halves summed, half h adds the entries a = 2h, 2h+1
        {
            // ordered mode: the item's own slot of item_ws (plain stores; k_sum_items adds the items of a cell in order)
            float* dst = ar.item_ws ? ar.item_ws + (int64_t)cur_id * 4 * TS : gnets + (int64_t)it.slot * 4 * TS;
#pragma unroll
            for (int k = 0; k < CPD; ++k) {
                float o[2][2];
#pragma unroll
                for (int p = 0; p < 2; ++p) {
                    unpack2(w2[p][k], o[p][0], o[p][1]);
                    o[p][0] += __shfl_xor_sync(0xffffffffu, o[p][0], 16);
                    o[p][1] += __shfl_xor_sync(0xffffffffu, o[p][1], 16);
                }
                if (live) {
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const int a = 2 * half + q;
                        if (a < N0) {
                            if (ar.item_ws) dst[k * TS + a * BC + l16] = half ? o[1][q] : o[0][q];
                            else atomicAdd(dst + k * TS + a * BC + l16, half ? o[1][q] : o[0][q]);
                        }
                    }
                }
            }
        }
        if (!nx_valid) break;
        ring = (ring + 1) % 3;
        {
            const int4* r = sm.rec[ring];
            it = unpack_item(r[0], r[1], r[2], r[3]);
        }
        tb ^= 1;
        const int id3 = __shfl_sync(0xffffffffu, nn_id, 0);
        cur_id = id1;
        id1 = id2;
        id2 = id3;
        if (lane < 4 && id3 < nitems) cp_async16(&sm.rec[(ring + 2) % 3][lane], ar.items + (int64_t)id3 * 4 + lane);
        cp_async_commit();
        if (lane == 0) nn_id = atomicAdd(cursor, 1);
    }
}

// ---- ordered (deterministic) variant of the two reductions that use atomics above -----------------------------------
// The reference's backward is an atomicAdd per (point, support) (compute_pts_cuda_kernels.cu:43): its result depends on the
// order in which threads arrive.  Ordered mode makes every sum of the fused backward run in a fixed order:
//   k_net_backward  stores each item's net gradient in its own slot (Args::item_ws)            -- no atomics
//   k_item_first    first item id of every cell (the items of a cell are consecutive)
//   k_sum_items     per cell: the items' net gradients added in item order                     -> the per-cell workspace
//   k_cell_unfold   writes <tile_s, gW> of every (cell, support) to its own slot (NetArgs::contrib) -- no atomics
//   k_fn_gather     per active function: its slots added in a fixed order (transposed support lists built by the host,
//                   PackedHierarchy.ordered_scatter_tables) -> the gradient row
// With the points of every cell in a canonical order (engine.Plan.canonicalize) two runs give bit-identical gradients.
static __global__ void __launch_bounds__(256) k_item_first(const int4* __restrict__ items, const int32_t* __restrict__ counters,
                                                    int32_t* __restrict__ item0) {
    const int n = counters[0];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int slot = __ldg(&items[(int64_t)i * 4].x);
        if (i == 0 || __ldg(&items[(int64_t)(i - 1) * 4].x) != slot) item0[slot] = i;
    }
}

static __global__ void __launch_bounds__(256) k_sum_items(const float* __restrict__ item_ws, const int32_t* __restrict__ item0,
                                                   const int32_t* __restrict__ cell_count, int nslots, int ppi, int TS, int T, int cpd,
                                                   float* __restrict__ gnets) {
    const int row = 4 * TS;
    // one thread per (cell, entry): the items of the cell in increasing item id
    const int64_t tot = (int64_t)nslots * row;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (int64_t)gridDim.x * blockDim.x) {
        const int slot = (int)(e / row), j = (int)(e - (int64_t)slot * row);
        const int c = __ldg(cell_count + slot);
        if (c == 0) continue;
        const int m = (c + ppi - 1) / ppi, i0 = __ldg(item0 + slot);
        float acc = 0.0f;
        if (j / TS < cpd && j % TS < T)   // the other entries of an item's slot are never written
            for (int i = 0; i < m; ++i) acc += item_ws[(int64_t)(i0 + i) * row + j];
        gnets[e] = acc;
    }
}

// one warp per active function (compact row r): contributions tr_idx[tr_ptr[r] .. tr_ptr[r+1]) in order -- lane l takes the
// entries l, l + 32, ... and the 32 partial sums are combined by a fixed butterfly
template <int CPD>
__global__ void __launch_bounds__(256) k_fn_gather(const float* __restrict__ contrib, const int32_t* __restrict__ tr_ptr,
                                                   const int32_t* __restrict__ tr_idx, const int32_t* __restrict__ active_ids,
                                                   int nrows, float* __restrict__ grad) {
    const int lane = threadIdx.x & 31;
    for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < nrows; r += (gridDim.x * blockDim.x) >> 5) {
        const int b = __ldg(tr_ptr + r), e = __ldg(tr_ptr + r + 1);
        float acc[CPD];
#pragma unroll
        for (int k = 0; k < CPD; ++k) acc[k] = 0.0f;
        for (int j = b + lane; j < e; j += 32) {
            const int64_t s = __ldg(tr_idx + j);
#pragma unroll
            for (int k = 0; k < CPD; ++k) acc[k] += contrib[s * CPD + k];
        }
#pragma unroll
        for (int k = 0; k < CPD; ++k)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
        if (lane == 0) {
            const int64_t row = active_ids ? (int64_t)__ldg(active_ids + r) : (int64_t)r;
#pragma unroll
            for (int k = 0; k < CPD; ++k) grad[row * CPD + k] = acc[k];
        }
    }
}

// One warp per active cell that holds points: lanes over the window entries scatter the own-level part, then lane = one
// tiled support: <tile_s, gW> from the factor vectors (rank one) or the tile row, three atomics per support.
template <int N0, int N1, int N2, int CPD>
__global__ void __launch_bounds__(32 * kNetWarps) k_cell_unfold(const __grid_constant__ thb_space_desc sp, const __grid_constant__ NetArgs na,
                                                               float* __restrict__ grad_cp) {
    using SH = Shp<N0, N1, N2>;
    constexpr int T = SH::T, TS = SH::TS;
    __shared__ __align__(16) float4 s_gw[kNetWarps][TS];   // [t] -> (gW[0][t], gW[1][t], gW[2][t], gW[3][t])
    __shared__ __align__(16) float4 s_cmat[kNetWarps][3][4];   // span polynomials of the cell (monomial net gradients)
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float4* gw = s_gw[wid];
    const bool split = sp.sep_vec != nullptr;
    for (int slot = blockIdx.x * kNetWarps + wid; slot < sp.nslots; slot += gridDim.x * kNetWarps) {
        if (na.cell_count && __ldg(na.cell_count + slot) == 0) continue;
        const int4* ci = reinterpret_cast<const int4*>(sp.cellinfo + (int64_t)slot * THB_INFO_W);
        const int4 a = __ldg(ci), e = __ldg(ci + 1), f = __ldg(ci + 2);
        const unsigned long long dm = __ldg(sp.dmask + slot);
        const int lev = a.x, soff = e.x, S = e.y, ndir = e.w;
        const int ntile = S - ndir, nsep = split ? f.y : 0;
        const float* src = na.nets + (int64_t)slot * 4 * TS;
        __syncwarp();
        if (na.mono) stage_span_polys(sp, lev, a.y, a.z, a.w, s_cmat[wid]);
        for (int t = lane; t < TS; t += 32) gw[t] = make_float4(src[t], src[TS + t], src[2 * TS + t], src[3 * TS + t]);
        if (na.mono) {
            // the workspace holds gM (monomial basis): gW[r] = sum_m C[m][r] gM[m] along every dimension (transposed conversion)
            cp_async_wait<0>();
            __syncwarp();
            constexpr int NT = (T + 31) / 32;
            constexpr int NN[3] = {N0, N1, N2};
            constexpr int STR[3] = {N1 * N2, N2, 1};
#pragma unroll
            for (int dim = 2; dim >= 0; --dim) {
                if (NN[dim] == 1) continue;
                const float* cm = reinterpret_cast<const float*>(&s_cmat[wid][dim][0]);   // [m][r] row major, 4 floats per row
                float4 o[NT];
#pragma unroll
                for (int h = 0; h < NT; ++h) {
                    const int t = lane + 32 * h;
                    o[h] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (t < T) {
                        const int r = (t / STR[dim]) % NN[dim];
                        const int base = t - r * STR[dim];
#pragma unroll
                        for (int m = 0; m < NN[dim]; ++m) {
                            const float c = cm[m * 4 + r];
                            const float4 v = gw[base + m * STR[dim]];
                            o[h].x = fmaf(c, v.x, o[h].x); o[h].y = fmaf(c, v.y, o[h].y); o[h].z = fmaf(c, v.z, o[h].z); o[h].w = fmaf(c, v.w, o[h].w);
                        }
                    }
                }
                __syncwarp();
#pragma unroll
                for (int h = 0; h < NT; ++h) {
                    const int t = lane + 32 * h;
                    if (t < T) gw[t] = o[h];
                }
                __syncwarp();
            }
        }
        __syncwarp();
        if (ndir > 0) {
            const int n1 = sp.nfns[lev][1], n2 = sp.nfns[lev][2];
            const int64_t f0 = sp.fn_off[lev];
            for (int t = lane; t < T; t += 32) {
                if ((dm >> t) & 1ull) {
                    const int t2 = t % N2, t1 = (t / N2) % N1, t0 = t / (N2 * N1);
                    const float4 v = gw[t];
                    if (na.contrib) {   // ordered mode: slot (cell, rank of t among the cell's own-level supports)
                        float* o = na.contrib + (int64_t)(soff + __popcll(dm & ((1ull << t) - 1ull))) * CPD;
                        o[0] = v.x;
                        if (CPD > 1) o[1] = v.y;
                        if (CPD > 2) o[2] = v.z;
                        if (CPD > 3) o[3] = v.w;
                        continue;
                    }
                    int64_t r = f0 + ((int64_t)(a.y + t0) * n1 + (a.z + t1)) * n2 + (a.w + t2);
                    if (na.row_map) r = __ldg(na.row_map + r);   // compact rows: active functions only
                    atomicAdd(grad_cp + r * CPD, v.x);
                    if (CPD > 1) atomicAdd(grad_cp + r * CPD + 1, v.y);
                    if (CPD > 2) atomicAdd(grad_cp + r * CPD + 2, v.z);
                    if (CPD > 3) atomicAdd(grad_cp + r * CPD + 3, v.w);
                }
            }
        }
        const float* tile_src = split ? sp.full_tiles + (int64_t)f.z * TS : sp.tiles + (int64_t)e.z * TS;
        const int32_t* full_jm = split ? sp.full_jm + f.z : sp.supp_jm + soff + ndir;
        for (int s0 = 0; s0 < ntile; s0 += 32) {
            const int sidx = s0 + lane;
            if (sidx >= ntile) continue;
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            int64_t r;
            if (sidx < nsep) {
                const float4* fv = reinterpret_cast<const float4*>(sp.sep_vec + (int64_t)(f.x + sidx) * 12);
                const float4 v0 = __ldg(fv), v1 = __ldg(fv + 1), v2 = __ldg(fv + 2);
                r = __ldg(sp.sep_jm + f.x + sidx);
#pragma unroll
                for (int i = 0; i < N0; ++i)
#pragma unroll
                    for (int j = 0; j < N1; ++j) {
                        const float ab = f4c(v0, i) * f4c(v1, j);
#pragma unroll
                        for (int k = 0; k < N2; ++k) {
                            const float x = ab * f4c(v2, k);
                            const float4 gv = gw[(i * N1 + j) * N2 + k];
                            acc[0] = fmaf(x, gv.x, acc[0]); acc[1] = fmaf(x, gv.y, acc[1]); acc[2] = fmaf(x, gv.z, acc[2]);
                            if (CPD == 4) acc[3] = fmaf(x, gv.w, acc[3]);
                        }
                    }
            } else {
                const float* row = tile_src + (int64_t)(sidx - nsep) * TS;
                r = __ldg(full_jm + (sidx - nsep));
#pragma unroll 4
                for (int t4 = 0; t4 < TS; t4 += 4) {
                    const float4 x = __ldg(reinterpret_cast<const float4*>(row + t4));
                    const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                    for (int m = 0; m < 4; ++m) {
                        const float4 gv = gw[t4 + m];
                        acc[0] = fmaf(xs[m], gv.x, acc[0]); acc[1] = fmaf(xs[m], gv.y, acc[1]); acc[2] = fmaf(xs[m], gv.z, acc[2]);
                        if (CPD == 4) acc[3] = fmaf(xs[m], gv.w, acc[3]);
                    }
                }
            }
            if (na.contrib) {   // ordered mode: slot (cell, n_direct + position among the tiled supports: rank-one ones first)
                float* o = na.contrib + (int64_t)(soff + ndir + sidx) * CPD;
#pragma unroll
                for (int k = 0; k < CPD; ++k) o[k] = acc[k];
                continue;
            }
            if (na.row_map) r = __ldg(na.row_map + r);
#pragma unroll
            for (int k = 0; k < CPD; ++k) atomicAdd(grad_cp + r * CPD + k, acc[k]);
        }
    }
}

template <int N0, int N1, int N2>
int launch_net_backward(const thb_space_desc* sp, const Args* ar, const NetArgs* na, float* grad_cp, int cpd, cudaStream_t st) {
#define THB_NET_BWD(CPD_)                                                                              \
    do {                                                                                               \
        static int g1_[THB_MAX_DEVICES] = {}, g2_[THB_MAX_DEVICES] = {};                               \
        int &g1 = g1_[thb_device()], &g2 = g2_[thb_device()];                                          \
        if (g1 == 0) g1 = persistent_grid(k_net_backward<N0, N1, N2, CPD_, true>, kThreads);            \
        if (g2 == 0) g2 = persistent_grid(k_cell_unfold<N0, N1, N2, CPD_>, 32 * kNetWarps);             \
        if (na->mono) k_net_backward<N0, N1, N2, CPD_, true><<<g1, kThreads, 0, st>>>(*sp, *ar, na->nets);   \
        else k_net_backward<N0, N1, N2, CPD_, false><<<g1, kThreads, 0, st>>>(*sp, *ar, na->nets);           \
        THB_LAUNCHED();                                                                                \
        k_cell_unfold<N0, N1, N2, CPD_><<<min((sp->nslots + kNetWarps - 1) / kNetWarps, g2), 32 * kNetWarps, 0, st>>>(*sp, *na, grad_cp); \
        THB_LAUNCHED();                                                                                \
        return THB_OK;                                                                                 \
    } while (0)
    if (cpd == 3) THB_NET_BWD(3);
    if (cpd == 4) THB_NET_BWD(4);
#undef THB_NET_BWD
    return thb_set_error(THB_E_UNSUPPORTED, "net backward: unsupported cp_dim");
}

// ordered backward: see the comment above k_item_first.  od: the buffers of thb_ordered_desc (thb_b200.h)
template <int N0, int N1, int N2>
int launch_net_backward_ordered(const thb_space_desc* sp, const Args* ar, const NetArgs* na, const thb_ordered_desc* od,
                                const int32_t* cell_count, int ppi, int max_items, float* grad, int cpd, cudaStream_t st) {
    using SH = Shp<N0, N1, N2>;
#define THB_NET_BWD_ORD(CPD_)                                                                          \
    do {                                                                                               \
        static int g1_[THB_MAX_DEVICES] = {}, g2_[THB_MAX_DEVICES] = {};                               \
        int &g1 = g1_[thb_device()], &g2 = g2_[thb_device()];                                          \
        if (g1 == 0) g1 = persistent_grid(k_net_backward<N0, N1, N2, CPD_, true>, kThreads);            \
        if (g2 == 0) g2 = persistent_grid(k_cell_unfold<N0, N1, N2, CPD_>, 32 * kNetWarps);             \
        if (na->mono) k_net_backward<N0, N1, N2, CPD_, true><<<g1, kThreads, 0, st>>>(*sp, *ar, na->nets);   \
        else k_net_backward<N0, N1, N2, CPD_, false><<<g1, kThreads, 0, st>>>(*sp, *ar, na->nets);           \
        THB_LAUNCHED();                                                                                \
        k_item_first<<<min((max_items + 255) / 256, thb_num_sms() * 8), 256, 0, st>>>(ar->items, ar->counters, od->item0); \
        THB_LAUNCHED();                                                                                \
        k_sum_items<<<thb_num_sms() * 8, 256, 0, st>>>(od->item_ws, od->item0, cell_count, sp->nslots, ppi, SH::TS, SH::T, CPD_, na->nets); \
        THB_LAUNCHED();                                                                                \
        k_cell_unfold<N0, N1, N2, CPD_><<<min((sp->nslots + kNetWarps - 1) / kNetWarps, g2), 32 * kNetWarps, 0, st>>>(*sp, *na, grad); \
        THB_LAUNCHED();                                                                                \
        k_fn_gather<CPD_><<<min((od->nrows * 32 + 255) / 256, thb_num_sms() * 8), 256, 0, st>>>(od->contrib, od->tr_ptr, od->tr_idx, \
                                                                                              od->active_ids, od->nrows, grad);  \
        THB_LAUNCHED();                                                                                \
        return THB_OK;                                                                                 \
    } while (0)
    if (cpd == 3) THB_NET_BWD_ORD(3);
    if (cpd == 4) THB_NET_BWD_ORD(4);
#undef THB_NET_BWD_ORD
    return thb_set_error(THB_E_UNSUPPORTED, "ordered net backward: unsupported cp_dim");
}

template <int N0, int N1, int N2>
int launch_nets(const thb_space_desc* sp, const NetArgs* na, int cpd, cudaStream_t st) {
#define THB_NETS(CPD_, REL_)                                                                  \
    do {                                                                                      \
        static int full_[THB_MAX_DEVICES] = {}; /* one resident wave: no partial second wave */ \
        int& full = full_[thb_device()];                                                      \
        if (full == 0) full = persistent_grid(k_cell_nets<N0, N1, N2, CPD_, REL_>, 32 * kNetWarps); \
        int cap = full;                                                                       \
        if (na->ctas_per_sm > 0) cap = min(full, thb_num_sms() * na->ctas_per_sm);             \
        const int grid = min((sp->nslots + kNetWarps - 1) / kNetWarps, cap);                   \
        k_cell_nets<N0, N1, N2, CPD_, REL_><<<grid, 32 * kNetWarps, 0, st>>>(*sp, *na);        \
        THB_LAUNCHED();                                                                       \
        return THB_OK;                                                                        \
    } while (0)
    if (cpd == 3 && na->planes == 4) THB_NETS(3, false);
    if (cpd == 3 && na->planes == 8) THB_NETS(3, true);
    if (cpd == 4 && na->planes == 4) THB_NETS(4, false);
    if (cpd == 4 && na->planes == 8) THB_NETS(4, true);
#undef THB_NETS
    return thb_set_error(THB_E_UNSUPPORTED, "cell nets: unsupported variant");
}

template <int N0, int N1, int N2>
int launch_net_eval(const thb_space_desc* sp, const Args* ar, const float* nets, int ppt, int cpd, int der, int mono, cudaStream_t st) {
#define THB_NET_EVAL(PPT_, CPD_, DER_, MONO_, ...)                                            \
    do {                                                                                      \
        static int grid_[THB_MAX_DEVICES] = {};                                               \
        int& grid = grid_[thb_device()];                                                      \
        if (grid == 0) grid = persistent_grid(k_net_eval<N0, N1, N2, PPT_, CPD_, DER_, MONO_, ##__VA_ARGS__>, kThreads); \
        k_net_eval<N0, N1, N2, PPT_, CPD_, DER_, MONO_, ##__VA_ARGS__><<<grid, kThreads, 0, st>>>(*sp, *ar, nets); \
        THB_LAUNCHED();                                                                       \
        return THB_OK;                                                                        \
    } while (0)
    if constexpr (N2 == 4 && N0 >= 2 && N1 >= 2) {
        // value + the three first derivatives (channels 0, 1, 2, 4): one fused pass over the monomial net
        if (mono && der && ar->nch == 4 && ar->channels[0] == 0 && ar->channels[1] == 1 && ar->channels[2] == 2 && ar->channels[3] == 4 &&
            !getenv("THB_NET_NO_GRAD4")) {
            if (cpd == 3) THB_NET_EVAL(2, 3, true, true, true);
            if (cpd == 4) THB_NET_EVAL(2, 4, true, true, true);
        }
    }
    if (mono) {   // monomial nets: four or two points per lane (fewer for tails)
        if (cpd == 3 && !der && ppt == 4) THB_NET_EVAL(4, 3, false, true);
        if (cpd == 4 && !der && ppt == 4) THB_NET_EVAL(4, 4, false, true);
        if (cpd == 3 && der && ppt == 4) THB_NET_EVAL(4, 3, true, true);
        if (cpd == 4 && der && ppt == 4) THB_NET_EVAL(4, 4, true, true);
        if (cpd == 3 && !der) THB_NET_EVAL(2, 3, false, true);
        if (cpd == 3 && der) THB_NET_EVAL(2, 3, true, true);
        if (cpd == 4 && !der) THB_NET_EVAL(2, 4, false, true);
        if (cpd == 4 && der) THB_NET_EVAL(2, 4, true, true);
    } else if (ppt == 2) {
        if (cpd == 3 && !der) THB_NET_EVAL(2, 3, false, false);
        if (cpd == 3 && der) THB_NET_EVAL(2, 3, true, false);
        if (cpd == 4 && !der) THB_NET_EVAL(2, 4, false, false);
        if (cpd == 4 && der) THB_NET_EVAL(2, 4, true, false);
    } else {
        if (cpd == 3 && !der) THB_NET_EVAL(1, 3, false, false);
        if (cpd == 3 && der) THB_NET_EVAL(1, 3, true, false);
        if (cpd == 4 && !der) THB_NET_EVAL(1, 4, false, false);
        if (cpd == 4 && der) THB_NET_EVAL(1, 4, true, false);
    }
#undef THB_NET_EVAL
    return thb_set_error(THB_E_UNSUPPORTED, "net evaluation: unsupported variant");
}

// ---- THB_basis_fns with PHI materialised, factor form -------------------------------------------------------------
// PHI rows in the reference's pair order (THB/core.py:604-701): own-level supports first, then the coarser ones.
//   phase 1 (lane = point): Cox-de Boor; the 1-D values (and derivatives) go to shared memory, the row offset too
//   own-level supports (lane = window entry t): PHI = A[a_t] * B[b_t] * C[c_t] per point, written straight to the row --
//       a coalesced store per point, no (p+1)^d register array and no transpose
//   coarser supports, 32 at a time in reference order (lane = point): rank-one ones as <v0,A><v1,B><v2,C> (14 instead of
//       2*(p+1)^d FMAs), the others by a sum-factorised dot with their tile; staged [point][support] and written as
//       row pieces (lane = support).  tiled_ref says for every tiled support which of the two tables holds it.
struct BasisSmem {
    int4 rec[3][4];
    float knots[2][3][8];
    float rk[3][8];
    float4 prec[2][32];
    float4 fac[6][32];        // [2 * dim + derivative][point] -> the (up to) four 1-D values of that point
    long long poff[32];       // BYTE offset of the point's row in PHI (a store address is one IMAD.WIDE: index * 4 + row)
    float stage[32][33];      // [point][support of the chunk]
    float4 vec[32][3];        // factor vectors of the chunk's rank-one supports (lane = support on the way in)
    float tile[64];           // one full tile at a time
    int ref[32];              // tiled_ref of the chunk
};

#ifndef THB_BASIS_MINB
#define THB_BASIS_MINB 16   // one-warp CTAs per SM of k_basis_rows (11 KB of shared memory each: 20 fit)
#endif
template <int N0, int N1, int N2, bool DER>
__global__ void __launch_bounds__(kThreads, THB_BASIS_MINB) k_basis_rows(const __grid_constant__ thb_space_desc sp, const __grid_constant__ Args ar) {
    using SH = Shp<N0, N1, N2>;
    constexpr int T = SH::T, TS = SH::TS;
    constexpr int NT = (T + 31) / 32;
    constexpr bool kPairA = NT == 2 && N0 == 4 && (N1 * N2) == 16;  // t and t + 32 = entries (a, b, c) and (a + 2, b, c)
    __shared__ __align__(16) BasisSmem sm;
    const int lane = threadIdx.x;
    const int nitems = ar.counters[0];
    int32_t* const cursor = ar.counters + 1;
    const int d = sp.ndim;

    int c3 = 0;
    if (lane == 0) c3 = atomicAdd(cursor, 3);
    const int id0 = __shfl_sync(0xffffffffu, c3, 0);
    if (id0 >= nitems) return;
    int id1 = id0 + 1, id2 = id0 + 2, cur_id = id0;
    Item it = load_item(sp, ar.items, id0);
    if (lane < 4 && id1 < nitems) cp_async16(&sm.rec[1][lane], ar.items + (int64_t)id1 * 4 + lane);
    if (lane < 4 && id2 < nitems) cp_async16(&sm.rec[2][lane], ar.items + (int64_t)id2 * 4 + lane);
    auto prefetch_knots = [&](const Item& x, float (*knots)[8]) {
        if (lane < 24) {
            const int dim = lane >> 3, m = lane & 7;
            const int P = dim == 0 ? SH::P0 : (dim == 1 ? SH::P1 : SH::P2);
            const int c = dim == 0 ? x.c0 : (dim == 1 ? x.c1 : x.c2);
            if (dim < d && m < 2 * P) cp_async4(&knots[dim][m], sp.knots + sp.knot_off[x.lev][dim] + c + 1 + m);
        }
    };
    prefetch_knots(it, sm.knots[0]);
    prefetch_points<1>(ar, it.begin, it.count, 0, sm.prec[0]);
    cp_async_commit();
    int nn_id = 0;
    if (lane == 0) nn_id = atomicAdd(cursor, 1);
    int ring = 0, tb = 0, pb = 0;
    cp_async_wait<0>();
    __syncwarp();

    // this lane's window entries (own-level part): t = lane + 32 h -> (a, b, c) and its place among the active ones
    int ta[NT], tb_[NT], tc[NT];
#pragma unroll
    for (int h = 0; h < NT; ++h) {
        const int t = min(lane + 32 * h, T - 1);
        ta[h] = t / (N1 * N2); tb_[h] = (t / N2) % N1; tc[h] = t % N2;
    }

    for (;;) {
        if (lane < 24) {
            const int dim = lane >> 3, m = lane & 7;
            const int P = dim == 0 ? SH::P0 : (dim == 1 ? SH::P1 : SH::P2);
            if (dim < d && m < P * (P + 1) / 2) {
                int j = 1, r = m;
                while (r >= j) { r -= j; ++j; }
                sm.rk[dim][m] = 1.0f / (sm.knots[tb][dim][P + r] - sm.knots[tb][dim][P - j + r]);
            }
        }
        __syncwarp();
        const int ntile = it.S - it.ndir;
        int dpos[NT];
#pragma unroll
        for (int h = 0; h < NT; ++h) {
            const int t = lane + 32 * h;
            dpos[h] = (it.ndir > 0 && t < T && ((it.dm >> t) & 1ull)) ? __popcll(it.dm & ((1ull << t) - 1ull)) : -1;
        }
        const bool nx_valid = id1 < nitems;
        for (int base = 0; base < it.count; base += 32) {
            if (base + 32 < it.count) {
                prefetch_points<1>(ar, it.begin, it.count, base + 32, sm.prec[pb ^ 1]);
            } else if (nx_valid) {
                const int4* r = sm.rec[(ring + 1) % 3];
                const Item nx = unpack_item(r[0], r[1], r[2], r[3]);
                prefetch_points<1>(ar, nx.begin, nx.count, 0, sm.prec[pb ^ 1]);
                prefetch_knots(nx, sm.knots[tb ^ 1]);
            }
            cp_async_commit();
            const int nval = min(32, it.count - base);
            // phase 1: lane = point
            BasesRK<SH, DER> bs;
            {
                const float4 rec = sm.prec[pb][lane];
                const float u[3] = {rec.x, rec.y, rec.z};
                bs.eval(u, sm.knots[tb], sm.rk);
                sm.poff[lane] = lane < nval ? 4ll * (long long)__ldg(ar.offsets + __float_as_int(rec.w)) : 0;
                float v[6][4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    v[0][i] = i < N0 ? bs.n0[i] : 0.f; v[2][i] = i < N1 ? bs.n1[i] : 0.f; v[4][i] = i < N2 ? bs.n2[i] : 0.f;
                    v[1][i] = (DER && i < N0) ? bs.d0[i] : 0.f; v[3][i] = (DER && i < N1) ? bs.d1[i] : 0.f; v[5][i] = (DER && i < N2) ? bs.d2[i] : 0.f;
                }
#pragma unroll
                for (int q = 0; q < 6; ++q) {
                    if (!DER && (q & 1)) continue;
                    // first dimension stored as (v0, v2, v1, v3) when a lane owns the entries a and a + 2: one LDS.64
                    if (kPairA && q < 2) sm.fac[q][lane] = make_float4(v[q][0], v[q][2], v[q][1], v[q][3]);
                    else sm.fac[q][lane] = make_float4(v[q][0], v[q][1], v[q][2], v[q][3]);
                }
            }
            __syncwarp();
            for (int chi = 0; chi < ar.nch; ++chi) {
                const int ch = ar.channels[chi];
                char* const outp = reinterpret_cast<char*>(ar.out[chi]);
                const int q0 = (DER && (ch & 1)) ? 1 : 0, q1 = (DER && (ch & 2)) ? 3 : 2, q2 = (DER && (ch & 4)) ? 5 : 4;
                if (it.ndir > 0) {
                    // own-level supports: lane = window entry
                    const float* fa = reinterpret_cast<const float*>(sm.fac[q0]);
                    const float* fb = reinterpret_cast<const float*>(sm.fac[q1]);
                    const float* fc = reinterpret_cast<const float*>(sm.fac[q2]);
                    if constexpr (kPairA) {
                        const float2* fa2 = reinterpret_cast<const float2*>(sm.fac[q0]) + ta[0];  // (A[a], A[a + 2]) of point 0
                        const float* fbp = fb + tb_[0];
                        const float* fcp = fc + tc[0];
#pragma unroll 4
                        for (int pt = 0; pt < nval; ++pt) {
                            float* dst = reinterpret_cast<float*>(outp + sm.poff[pt]);
                            const float2 aa = fa2[pt * 2];
                            const float bc = fbp[pt * 4] * fcp[pt * 4];
                            if (dpos[0] >= 0) dst[(unsigned)dpos[0]] = aa.x * bc;
                            if (dpos[1] >= 0) dst[(unsigned)dpos[1]] = aa.y * bc;
                        }
                    } else {
#pragma unroll 4
                        for (int pt = 0; pt < nval; ++pt) {
                            float* dst = reinterpret_cast<float*>(outp + sm.poff[pt]);
#pragma unroll
                            for (int h = 0; h < NT; ++h) {
                                const float val = fa[pt * 4 + ta[h]] * fb[pt * 4 + tb_[h]] * fc[pt * 4 + tc[h]];
                                if (dpos[h] >= 0) dst[(unsigned)dpos[h]] = val;
                            }
                        }
                    }
                }
                // coarser-level supports, 32 at a time: lane = point computes, lane = support writes
                float A[N0], B[N1], Cc[N2];
                bs.select(ch, A, B, Cc);
                for (int c0 = 0; c0 < ntile; c0 += 32) {
                    const int ns = min(32, ntile - c0);
                    __syncwarp();
                    if (lane < ns) {
                        int r = sp.tiled_ref ? __ldg(sp.tiled_ref + it.toff + c0 + lane) : (((it.toff + c0 + lane) << 1) | 1);
                        sm.ref[lane] = r;
                        if (!(r & 1)) {
                            const float4* fv = reinterpret_cast<const float4*>(sp.sep_vec + (int64_t)(r >> 1) * 12);
                            sm.vec[lane][0] = __ldg(fv); sm.vec[lane][1] = __ldg(fv + 1); sm.vec[lane][2] = __ldg(fv + 2);
                        }
                    }
                    __syncwarp();
                    for (int j = 0; j < ns; ++j) {
                        const int r = sm.ref[j];
                        float e;
                        if (!(r & 1)) {
                            const float4 v0 = sm.vec[j][0], v1 = sm.vec[j][1], v2 = sm.vec[j][2];
                            float e0 = 0.f, e1 = 0.f, e2 = 0.f;
#pragma unroll
                            for (int i = 0; i < N0; ++i) e0 = fmaf(f4c(v0, i), A[i], e0);
#pragma unroll
                            for (int i = 0; i < N1; ++i) e1 = fmaf(f4c(v1, i), B[i], e1);
#pragma unroll
                            for (int i = 0; i < N2; ++i) e2 = fmaf(f4c(v2, i), Cc[i], e2);
                            e = e0 * e1 * e2;
                        } else {
                            const float* row = (sp.tiled_ref ? sp.full_tiles : sp.tiles) + (int64_t)(r >> 1) * TS;
                            __syncwarp();
                            for (int t = lane; t < TS; t += 32) sm.tile[t] = __ldg(row + t);
                            __syncwarp();
                            e = 0.f;
#pragma unroll
                            for (int i = 0; i < N0; ++i) {
                                float ei = 0.f;
#pragma unroll
                                for (int k = 0; k < N1; ++k) {
                                    float ek = 0.f;
#pragma unroll
                                    for (int m = 0; m < N2; ++m) ek = fmaf(Cc[m], sm.tile[(i * N1 + k) * N2 + m], ek);
                                    ei = fmaf(B[k], ek, ei);
                                }
                                e = fmaf(A[i], ei, e);
                            }
                        }
                        sm.stage[lane][j] = e;
                    }
                    __syncwarp();
                    if (lane < ns) {
#pragma unroll 4
                        for (int pt = 0; pt < nval; ++pt) reinterpret_cast<float*>(outp + sm.poff[pt])[(unsigned)(it.ndir + c0 + lane)] = sm.stage[pt][lane];
                    }
                }
            }
            cp_async_wait<0>();
            __syncwarp();
            pb ^= 1;
        }
        if (!nx_valid) break;
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

template <int N0, int N1, int N2>
int launch_basis_rows(const thb_space_desc* sp, const Args* ar, int der, cudaStream_t st) {
    if (!der) THB_TILE_LAUNCH((k_basis_rows<N0, N1, N2, false>));
    THB_TILE_LAUNCH((k_basis_rows<N0, N1, N2, true>));
}

// ---- per-shape launch table (one translation unit per tensor-product shape, see thb_tile_inst_*.cu) ---------------
struct ShapeOps {
    int n0, n1, n2;
    // ppt in {1,2}; cpd in {3,4}; der in {0,1}
    int (*fused)(const thb_space_desc*, const Args*, int ppt, int cpd, int der, cudaStream_t);
    int (*nets)(const thb_space_desc*, const NetArgs*, int cpd, cudaStream_t);
    int (*net_eval)(const thb_space_desc*, const Args*, const float* nets, int ppt, int cpd, int der, int mono, cudaStream_t);
    int (*net_backward)(const thb_space_desc*, const Args*, const NetArgs*, float* grad_cp, int cpd, cudaStream_t);
    int (*net_backward_ordered)(const thb_space_desc*, const Args*, const NetArgs*, const thb_ordered_desc*, const int32_t* cell_count,
                                int ppi, int max_items, float* grad, int cpd, cudaStream_t);
    int (*basis)(const thb_space_desc*, const Args*, int ppt, int der, cudaStream_t);
    int (*basis_rows)(const thb_space_desc*, const Args*, int der, cudaStream_t);
    int (*backward)(const thb_space_desc*, const Args*, int cpd, cudaStream_t);
};

#define THB_DEFINE_SHAPE(A, B, C)                                                                                    \
    namespace thb_tile {                                                                                             \
    extern const ShapeOps shape_ops_##A##B##C = {A, B, C, &launch_fused<A, B, C>, &launch_nets<A, B, C>,              \
                                                 &launch_net_eval<A, B, C>, &launch_net_backward<A, B, C>,               \
                                                 &launch_net_backward_ordered<A, B, C>, &launch_basis<A, B, C>,          \
                                                 &launch_basis_rows<A, B, C>,                                          \
                                                 &launch_backward<A, B, C>};                                         \
    }

}  // namespace thb_tile
