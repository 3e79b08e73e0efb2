/* CPU ORACLE, plain C (test infrastructure -- NOT part of the product; see oracle/thb_oracle.py header).
 *
 * Restates, in double precision on the host:
 *   oracle_find_spans     find_span_array_jax                       THB/bspline_funcs.py:77-83
 *   oracle_basis_funs     basisFun (NURBS-book A2.2) + 1st derivs   THB/bspline_funcs.py:86-101
 *   oracle_eval_forward   compute_pts_forward  (float accumulation, same loop order)   THB_extensions/compute_pts.cpp:4-29
 *   oracle_eval_backward  compute_pts_backward                                          THB_extensions/compute_pts.cpp:31-54
 *   oracle_thb_eval       compute_active_span + compute_THB_fns_tp + compute_pts_forward
 *                         (THB/core.py:157-189, 232-268) for a hierarchy given as flat tables: per point the
 *                         finest-level span, the walk to the active cell, Cox-de Boor, the (p+1)^d tensor
 *                         product, one dot product per support, then sum PHI * control point.
 * The last one takes the coefficient windows in the flat "tile" form (one (p+1)^d window per cell and
 * coarser-level support, expressed in the cell's own level; own-level supports are the tensor-product
 * entries themselves) because the reference's dense [nfns x nfns_finest] tables cannot be stored beyond toy
 * sizes.  tests/test_oracle_c.py checks it against the NumPy oracle, which is pinned to the reference's
 * own outputs.  Built by oracle/build_c.py (gcc -O2 -fopenmp).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXL 8
#define MAXD 6

/* launchers such as torchrun export OMP_NUM_THREADS=1: the benchmark's CPU baseline asks for the host's cores explicitly */
void oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

static int span_right(const float* U, int nk, int p, float u) {
    int lo = 0, hi = nk;
    while (lo < hi) {
        int mid = (lo + hi) / 2;
        if (U[mid] <= u) lo = mid + 1; else hi = mid;
    }
    int n = nk - p - 1, idx = lo - 1;
    if (idx > n) idx = n;
    if (u == U[nk - 1]) idx = n - 1;
    return idx;
}

void oracle_find_spans(const float* params, int64_t stride, int64_t n, const float* knots, int nk, int p, int32_t* out) {
    for (int64_t i = 0; i < n; ++i) out[i] = span_right(knots, nk, p, params[i * stride]);
}

/* values N[0..p] and first derivatives dN[0..p] (dN may be NULL) of the functions span-p..span */
void oracle_basis_funs(int span, double u, int p, const double* U, double* N, double* dN) {
    double left[MAXD + 1], right[MAXD + 1], low[MAXD + 1];
    N[0] = 1.0;
    for (int j = 1; j <= p; ++j) {
        if (j == p) memcpy(low, N, sizeof(double) * p);
        left[j] = u - U[span + 1 - j];
        right[j] = U[span + j] - u;
        double carry = 0.0;
        for (int r = 0; r < j; ++r) {
            double t = N[r] / (right[r + 1] + left[j - r]);
            N[r] = carry + right[r + 1] * t;
            carry = left[j - r] * t;
        }
        N[j] = carry;
    }
    if (dN) {
        if (p == 0) { dN[0] = 0.0; return; }
        double a[MAXD + 2];
        a[0] = 0.0; a[p + 1] = 0.0;
        for (int r = 1; r <= p; ++r) a[r] = low[r - 1] / (U[span + r] - U[span - p + r]);
        for (int r = 0; r <= p; ++r) dN[r] = p * (a[r] - a[r + 1]);
    }
}

void oracle_eval_forward(const float* cp, const int64_t* jm, const float* phi, const int64_t* cumsum, float* out, int64_t n, int cpd) {
    for (int64_t i = 0; i < n; ++i) {
        int64_t b = i ? cumsum[i - 1] : 0, e = cumsum[i];
        float v[4] = {0, 0, 0, 0};
        for (int64_t j = b; j < e; ++j)
            for (int k = 0; k < cpd; ++k) v[k] += phi[j] * cp[jm[j] * cpd + k];
        for (int k = 0; k < cpd; ++k) out[i * cpd + k] = v[k];
    }
}

void oracle_eval_backward(const float* go, const int64_t* jm, const float* phi, const int64_t* cumsum, float* gcp, int64_t n, int64_t ncp, int cpd) {
    memset(gcp, 0, sizeof(float) * ncp * cpd);
    for (int64_t i = 0; i < n; ++i) {
        int64_t b = i ? cumsum[i - 1] : 0, e = cumsum[i];
        for (int64_t j = b; j < e; ++j)
            for (int k = 0; k < cpd; ++k) gcp[jm[j] * cpd + k] += phi[j] * go[i * cpd + k];
    }
}

typedef struct {
    int32_t ndim, nlevels, degree[3], tile_stride;
    int32_t info_w; /* ints per cellinfo row (columns 0..7 are used here) */
    int32_t knot_off[MAXL][3], knot_len[MAXL][3], ncells[MAXL][3];
    int64_t cell_slot_off[MAXL];
    const float* knots;
    const int32_t* cell_slot;
    const int32_t* cellinfo; /* [nslots][info_w]: level, c0, c1, c2, supp_off, S, tile_off, n_direct, ... */
    const uint64_t* dmask;
    const int32_t* supp_jm;
    const float* tiles;
} oracle_space;

/* mode bit k: derivative in dim k.  Outputs (any may be NULL): out[n][cpd] (double), slot[n], phi (flat, at
 * offsets[i]), all in caller order. Returns the number of in-domain points. */
int64_t oracle_thb_eval(const oracle_space* sp, const float* params, int64_t n, int mode, const float* cp, int cpd,
                        double* out, int32_t* slot_out, const int64_t* offsets, double* phi_out) {
    const int d = sp->ndim, L = sp->nlevels - 1;
    int T = 1;
    for (int k = 0; k < d; ++k) T *= sp->degree[k] + 1;
    int64_t inside = 0;
#pragma omp parallel for schedule(dynamic, 1024) reduction(+ : inside)
    for (int64_t i = 0; i < n; ++i) {
        int cf[3] = {0, 0, 0}, ok = 1;
        for (int k = 0; k < d; ++k) {
            int s = span_right(sp->knots + sp->knot_off[L][k], sp->knot_len[L][k], sp->degree[k], params[i * d + k]);
            cf[k] = s - sp->degree[k];
            if (cf[k] < 0 || cf[k] >= sp->ncells[L][k]) ok = 0;
        }
        int slot = -1;
        if (ok)
            for (int lev = L; lev >= 0 && slot < 0; --lev) {
                int sh = L - lev;
                int64_t lin = cf[0] >> sh;
                lin = lin * sp->ncells[lev][1] + (cf[1] >> sh);
                if (d == 3) lin = lin * sp->ncells[lev][2] + (cf[2] >> sh);
                slot = sp->cell_slot[sp->cell_slot_off[lev] + lin];
            }
        if (slot_out) slot_out[i] = slot;
        if (out) for (int k = 0; k < cpd; ++k) out[i * cpd + k] = 0.0;
        if (slot < 0) continue;
        inside += 1;
        const int32_t* ci = sp->cellinfo + (int64_t)slot * sp->info_w;
        const int lev = ci[0], soff = ci[4], S = ci[5], toff = ci[6], ndir = ci[7];
        double N[3][MAXD + 1], dN[3][MAXD + 1], Ud[4 * MAXD + 4];
        for (int k = 0; k < d; ++k) {
            const int p = sp->degree[k], c = ci[1 + k];
            const float* U = sp->knots + sp->knot_off[lev][k];
            /* local copy of the knots the recursion touches: U[c+1 .. c+2p], placed so that index span works */
            for (int m = 0; m <= 2 * p + 1; ++m) Ud[m] = (double)U[c + m];
            oracle_basis_funs(p, (double)params[i * d + k], p, Ud, N[k], dN[k]);
        }
        double tp[64];
        {
            const int n1 = sp->degree[1] + 1, n2 = d == 3 ? sp->degree[2] + 1 : 1;
            for (int a = 0; a <= sp->degree[0]; ++a)
                for (int b = 0; b < n1; ++b)
                    for (int c = 0; c < n2; ++c) {
                        double v = ((mode & 1) ? dN[0][a] : N[0][a]) * ((mode & 2) ? dN[1][b] : N[1][b]);
                        if (d == 3) v *= (mode & 4) ? dN[2][c] : N[2][c];
                        tp[(a * n1 + b) * n2 + c] = v;
                    }
        }
        double acc[4] = {0, 0, 0, 0};
        int s = 0;
        if (ndir > 0) {
            const uint64_t dm = sp->dmask[slot];
            for (int t = 0; t < T; ++t)
                if ((dm >> t) & 1ull) {
                    const double phi = tp[t];
                    if (phi_out) phi_out[offsets[i] + s] = phi;
                    if (out) for (int k = 0; k < cpd; ++k) acc[k] += phi * (double)cp[(int64_t)sp->supp_jm[soff + s] * cpd + k];
                    ++s;
                }
        }
        for (; s < S; ++s) {
            const float* tile = sp->tiles + (int64_t)(toff + s - ndir) * sp->tile_stride;
            double phi = 0.0;
            for (int t = 0; t < T; ++t) phi += (double)tile[t] * tp[t];
            if (phi_out) phi_out[offsets[i] + s] = phi;
            if (out) for (int k = 0; k < cpd; ++k) acc[k] += phi * (double)cp[(int64_t)sp->supp_jm[soff + s] * cpd + k];
        }
        if (out) for (int k = 0; k < cpd; ++k) out[i * cpd + k] = acc[k];
    }
    return inside;
}

/* ---------------------------------------------------------------------------------------------------------------
 * Table-free evaluator: the reference's (single-level-)truncated functions from their DEFINITION,
 *     trunc(B^l_i)(x) = sum_{j child of i, fns[l+1][j] != 1}  prod_k Coeff[l][k][i_k, j_k] * B^{l+1}_j(x)
 * (for l == L the function itself), i.e. compute_refinement_operators + finest-level evaluation in exact arithmetic
 * [THB/core.py:43-114, 232-268] without the dense tables -- the C form of oracle/thb_oracle.py: DirectEvaluator.point,
 * so that the packed tables of the product can be pinned at 1e4..1e5 points on the BASELINE spaces.
 * Cell location: finest level downwards, first level whose cell is active [THB/core.py:165-187].
 * Supports: own level first, then ancestors via cell // 2, window in C order, fns == 1 only [THB/core.py:424-449]. */
typedef struct {
    int32_t ndim, nlevels, degree[3];
    int32_t nfn[MAXL][3], ncell[MAXL][3], nk[MAXL][3];
    const double* knots[MAXL][3];
    const uint8_t* cells[MAXL];  /* [ncell] C order, 1 = active */
    const uint8_t* fns[MAXL];    /* [nfn] C order, 1 = active */
    const double* coeff[MAXL][3]; /* level l -> l+1, [nfn[l][k]][nfn[l+1][k]] row major (l < L) */
    int64_t fn_off[MAXL];
} oracle_direct;

static int span_right_d(const double* U, int nk, int p, double u) {
    int lo = 0, hi = nk;
    while (lo < hi) {
        int mid = (lo + hi) / 2;
        if (U[mid] <= u) lo = mid + 1; else hi = mid;
    }
    int n = nk - p - 1, idx = lo - 1;
    if (idx > n) idx = n;
    if (u == U[nk - 1]) idx = n - 1;
    return idx;
}

/* x: [n][d] doubles.  Outputs per point: lev, cell[3], nsupp, jm[max_supp], phi[max_supp], dphi[max_supp][3] (may be NULL).
 * Returns the largest support count met (> max_supp means the buffers were too small for some points: nsupp is still
 * right, the lists are cut). */
int64_t oracle_direct_eval(const oracle_direct* sp, const double* x, int64_t n, int max_supp, int32_t* lev_out, int32_t* cell_out,
                           int32_t* nsupp, int64_t* jm, double* phi, double* dphi) {
    const int d = sp->ndim, L = sp->nlevels - 1;
    int64_t worst = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(max : worst)
    for (int64_t i = 0; i < n; ++i) {
        const double* xi = x + i * d;
        int lev = -1, c[3] = {0, 0, 0};
        for (int l = L; l >= 0 && lev < 0; --l) {
            int ok = 1, cc[3] = {0, 0, 0};
            for (int k = 0; k < d; ++k) {
                cc[k] = span_right_d(sp->knots[l][k], sp->nk[l][k], sp->degree[k], xi[k]) - sp->degree[k];
                if (cc[k] < 0 || cc[k] >= sp->ncell[l][k]) ok = 0;
            }
            if (!ok) continue;
            int64_t lin = cc[0];
            for (int k = 1; k < d; ++k) lin = lin * sp->ncell[l][k] + cc[k];
            if (sp->cells[l][lin] == 1) { lev = l; c[0] = cc[0]; c[1] = cc[1]; c[2] = cc[2]; }
        }
        lev_out[i] = lev;
        for (int k = 0; k < 3; ++k) cell_out[i * 3 + k] = c[k];
        nsupp[i] = 0;
        if (lev < 0) continue;
        /* 1-D values / derivatives at every level that is needed (levels lev+1 .. 1 for truncated functions, L for own) */
        int start[MAXL][3];
        double N[MAXL][3][MAXD + 1], dN[MAXL][3][MAXD + 1];
        int have[MAXL];
        for (int l = 0; l <= L; ++l) have[l] = 0;
        int s = 0, cw[3] = {c[0], c[1], c[2]};
        const int w0 = sp->degree[0] + 1, w1 = sp->degree[1] + 1, w2 = d == 3 ? sp->degree[2] + 1 : 1;
        for (int fl = lev; fl >= 0; --fl) {
            const int el = fl == L ? fl : fl + 1;
            if (!have[el]) {
                for (int k = 0; k < d; ++k) {
                    const int p = sp->degree[k];
                    const int sp_ = span_right_d(sp->knots[el][k], sp->nk[el][k], p, xi[k]);
                    start[el][k] = sp_ - p;
                    oracle_basis_funs(sp_, xi[k], p, sp->knots[el][k], N[el][k], dN[el][k]);
                }
                have[el] = 1;
            }
            for (int a = 0; a < w0; ++a)
                for (int b = 0; b < w1; ++b)
                    for (int e = 0; e < w2; ++e) {
                        int fn[3] = {cw[0] + a, cw[1] + b, cw[2] + e};
                        int64_t lin = fn[0];
                        for (int k = 1; k < d; ++k) lin = lin * sp->nfn[fl][k] + fn[k];
                        if (sp->fns[fl][lin] != 1) continue;
                        double v = 0.0, g[3] = {0.0, 0.0, 0.0};
                        if (fl == L) {
                            int ok = 1, idx[3] = {0, 0, 0};
                            for (int k = 0; k < d; ++k) {
                                idx[k] = fn[k] - start[el][k];
                                if (idx[k] < 0 || idx[k] > sp->degree[k]) ok = 0;
                            }
                            if (ok) {
                                v = 1.0;
                                for (int k = 0; k < d; ++k) v *= N[el][k][idx[k]];
                                for (int j = 0; j < d; ++j) {
                                    double t = 1.0;
                                    for (int k = 0; k < d; ++k) t *= (k == j ? dN[el][k][idx[k]] : N[el][k][idx[k]]);
                                    g[j] = t;
                                }
                            }
                        } else {
                            for (int ta = 0; ta < w0; ++ta)
                                for (int tb = 0; tb < w1; ++tb)
                                    for (int te = 0; te < w2; ++te) {
                                        const int t[3] = {ta, tb, te};
                                        int64_t cl = start[el][0] + ta;
                                        double cwt = 1.0;
                                        for (int k = 0; k < d; ++k) {
                                            const int child = start[el][k] + t[k];
                                            if (k) cl = cl * sp->nfn[el][k] + child;
                                            cwt *= sp->coeff[fl][k][(int64_t)fn[k] * sp->nfn[el][k] + child];
                                        }
                                        if (sp->fns[el][cl] == 1 || cwt == 0.0) continue;
                                        double tv = cwt;
                                        for (int k = 0; k < d; ++k) tv *= N[el][k][t[k]];
                                        v += tv;
                                        for (int j = 0; j < d; ++j) {
                                            double tg = cwt;
                                            for (int k = 0; k < d; ++k) tg *= (k == j ? dN[el][k][t[k]] : N[el][k][t[k]]);
                                            g[j] += tg;
                                        }
                                    }
                        }
                        if (s < max_supp) {
                            jm[i * max_supp + s] = sp->fn_off[fl] + lin;
                            phi[i * max_supp + s] = v;
                            if (dphi) for (int j = 0; j < 3; ++j) dphi[(i * max_supp + s) * 3 + j] = g[j];
                        }
                        ++s;
                    }
            for (int k = 0; k < 3; ++k) cw[k] /= 2;
        }
        nsupp[i] = s;
        if (s > worst) worst = s;
    }
    return worst;
}
