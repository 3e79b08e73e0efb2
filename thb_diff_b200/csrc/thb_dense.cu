// Legacy form of THB_basis_fns for callers that hold the reference's DENSE coefficient tables
// (fn_coeffs[lev] of shape [*sh_fns[lev], *sh_fns[max_lev]], THB/core.py:43-114, flattened and stacked as in
// core.py:636-643).  Evaluation exactly as the reference: finest-level span and basis per point
// (core.py:652-656, 681-684), then for every (point, support) pair the dot product of the (p+1)^d tensor
// product with the finest-level window of the support's table row (core.py:658-677 with the SURVEY Q2/Q3
// repairs; float64 statement: compute_THB_fns_tp, core.py:232-268).
//
// One thread per (point, support) pair; the point of a pair is found by binary search in the inclusive
// cumsum.  This is a compatibility path (tables of this form only exist for small spaces); the production
// path is the tile kernel in thb_tile_impl.cuh.
#include "thb_common.cuh"

namespace {

struct DenseArgs {
    int ndim;
    int degree[3];
    int knot_off[3], knot_len[3];
    int nf[3];  // finest-level function counts
};

template <typename JT>
__global__ void __launch_bounds__(256) k_basis_dense(const float* __restrict__ params, int64_t npoints, DenseArgs a,
                                                     const float* __restrict__ knots, const float* __restrict__ coeffs,
                                                     const JT* __restrict__ jm, const int64_t* __restrict__ cumsum, int channel,
                                                     float* __restrict__ phi_out) {
    const int64_t nnz = __ldg(cumsum + (npoints - 1));   // read on the device: the host never waits for it
    const int64_t row_len = (int64_t)a.nf[0] * a.nf[1] * (a.ndim == 3 ? a.nf[2] : 1);
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < nnz; j += (int64_t)gridDim.x * blockDim.x) {
        // first point i with cumsum[i] > j
        int64_t lo = 0, hi = npoints - 1;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (__ldg(cumsum + mid) > j) hi = mid; else lo = mid + 1;
        }
        const int64_t i = lo;
        float N[3][THB_MAX_DEG + 1];
        int start[3] = {0, 0, 0};
        bool ok = true;
        for (int k = 0; k < a.ndim; ++k) {
            const float u = params[i * a.ndim + k];
            const float* U = knots + a.knot_off[k];
            const int s = span_right(U, a.knot_len[k], a.degree[k], u);
            if (s < a.degree[k] || s >= a.nf[k]) { ok = false; break; }
            float v[THB_MAX_DEG + 1], dv[THB_MAX_DEG + 1];
            cox_de_boor_rt(a.degree[k], s, u, U, v, dv);
            const bool der = (channel >> k) & 1;
            for (int r = 0; r <= a.degree[k]; ++r) N[k][r] = der ? dv[r] : v[r];
            start[k] = s - a.degree[k];
        }
        float acc = 0.0f;
        if (ok) {
            const float* row = coeffs + (int64_t)jm[j] * row_len;
            const int n2 = a.ndim == 3 ? a.degree[2] + 1 : 1;
            for (int x = 0; x <= a.degree[0]; ++x)
                for (int y = 0; y <= a.degree[1]; ++y) {
                    const float xy = N[0][x] * N[1][y];
                    int64_t base = (int64_t)(start[0] + x) * a.nf[1] + (start[1] + y);
                    if (a.ndim == 3) base = base * a.nf[2] + start[2];
                    for (int z = 0; z < n2; ++z) acc = fmaf(__ldg(row + base + z), a.ndim == 3 ? xy * N[2][z] : xy, acc);
                }
        }
        phi_out[j] = acc;
    }
}

}  // namespace

extern "C" int thb_basis_dense(const float* params, int64_t npoints, int ndim, const int32_t* degree, const float* knots,
                               const int32_t* knot_off, const int32_t* knot_len, const int32_t* nfns_finest, const float* coeffs,
                               const void* jm, int jm_is_i64, const int64_t* cumsum, int channel, float* phi_out, void* stream) {
    THB_REQUIRE(ndim == 2 || ndim == 3, "thb_basis_dense: ndim must be 2 or 3");
    THB_REQUIRE(npoints >= 0, "thb_basis_dense: negative point count");
    if (npoints == 0) return THB_OK;
    THB_REQUIRE(params && degree && knots && knot_off && knot_len && nfns_finest && coeffs && jm && cumsum && phi_out,
                "thb_basis_dense: null pointer");
    THB_REQUIRE(channel >= 0 && channel < (1 << ndim), "thb_basis_dense: bad channel mask");
    DenseArgs a = {};
    a.ndim = ndim;
    for (int k = 0; k < ndim; ++k) {
        THB_REQUIRE(degree[k] >= 0 && degree[k] <= THB_MAX_DEG, "thb_basis_dense: degree must be 0..%d", THB_MAX_DEG);
        a.degree[k] = degree[k]; a.knot_off[k] = knot_off[k]; a.knot_len[k] = knot_len[k]; a.nf[k] = nfns_finest[k];
    }
    cudaStream_t st = (cudaStream_t)stream;
    // nnz = cumsum[npoints-1] lives on the device and stays there: the grid is sized from npoints (every point has
    // at least one pair, so this never exceeds the pair count by much) capped at the machine, and strides over nnz
    int64_t want = (npoints * 8 + 255) / 256, cap = (int64_t)thb_num_sms() * 16;
    const int grid = (int)(want < cap ? want : cap);
    if (jm_is_i64)
        k_basis_dense<long long><<<grid, 256, 0, st>>>(params, npoints, a, knots, coeffs, (const long long*)jm, cumsum, channel, phi_out);
    else
        k_basis_dense<int><<<grid, 256, 0, st>>>(params, npoints, a, knots, coeffs, (const int*)jm, cumsum, channel, phi_out);
    THB_LAUNCHED();
    return THB_OK;
}
