// Evaluation on a TENSOR-PRODUCT GRID of parameters (the point sets of the reference's generate_parametric_coordinates,
// THB/bspline_funcs.py:20-35, and of every BASELINE configuration): out[i0, i1, i2] = geometry(u0[i0], u1[i1], u2[i2]).
//
// The general path (thb_plan_build + thb_eval_nets) treats the points as an arbitrary batch: a counting sort by active cell,
// then (p+1)^d-term work per point.  On a grid both steps collapse:
//   * the points of an active cell are the tensor product of one index range per axis -- no sort, no per-point records: the
//     host lists (cell, index box) work items once per grid (thb_diff_b200.engine.GridPlan);
//   * over a cell the geometry is the polynomial  sum M[m0][m1][m2] s0^m0 s1^m1 s2^m2  (monomial nets, thb_cell_nets with
//     THB_NETS_MONOMIAL), and a polynomial on a box of n0 x n1 x n2 points is evaluated by sum factorisation AT THE CELL
//     LEVEL: Horner in s2 for the n2 values (N0*N1 polynomials), Horner in s1 for the n1*n2 pairs (N0 polynomials), Horner
//     in s0 per point -- (N0-1) FMAs per point and component plus O(n1*n2 + n2) per box, instead of (p+1)^d-many per point.
// One warp per work item (a box of at most 16 x 16 x 16 points of one cell); lane = (i1, i2) pair, loop over i0.
#include "thb_common.cuh"

namespace {

constexpr int kBox = 16;          // points per axis and work item
constexpr int kGridWarps = 4;     // warps (= items in flight) per CTA

struct GridArgs {
    const float* nets;            // [nslots][4][TS] monomial nets
    const int32_t* items;         // [nitems][8]: slot, i0, n0, i1, n1, i2, n2, level
    const int32_t* cells;         // [nitems][4]: c0, c1, c2, 0  (the cell's 1-D cell indices)
    int32_t nitems;
    const float* axis[3];
    int64_t stride[3];            // output row of point (i0, i1, i2) = i0 * stride[0] + i1 * stride[1] + i2 * stride[2]
    float* out;
    int32_t* cursor;              // dynamic scheduler (zeroed by the entry point)
};

// One warp per work item, items handed out by an atomic cursor (the host sorts them by size, large boxes first).
// Measured on the 256^3 grid of config 3 (54 963 boxes): 0.136 ms.  Variants that did NOT help, all between 0.145 and
// 0.19 ms: a two-stage cp.async pipeline for the nets with round-robin item assignment, a persistent single-wave grid,
// 8-byte and 16-byte stores per lane (pairs / quadruples of i2), an output slice staged through shared memory so that every
// store instruction writes whole sectors.  The kernel runs at ~110 G points/s for every grid size from 128^3 to 512^3 and
// no memory-side unit is above 35 % busy (ncu): what limits it is the serial latency of a warp's per-box phases.
template <int N0, int N1, int N2, int CPD>
__global__ void __launch_bounds__(32 * kGridWarps) k_grid_eval(const __grid_constant__ thb_space_desc sp, const __grid_constant__ GridArgs ga) {
    constexpr int T = N0 * N1 * N2, TS = (T + 3) / 4 * 4;
    __shared__ float s_m[kGridWarps][CPD][TS];                 // the cell's monomial net
    __shared__ float s_p2[kGridWarps][CPD][kBox][N0 * N1];     // after the Horner pass over m2: [i2][m0][m1]
    __shared__ float s_s[kGridWarps][3][kBox];                 // local coordinates of the box's axis values
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float (*M)[TS] = s_m[wid];
    float (*P2)[kBox][N0 * N1] = s_p2[wid];
    float (*S)[kBox] = s_s[wid];

    for (;;) {
        int item = 0;
        if (lane == 0) item = atomicAdd(ga.cursor, 1);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= ga.nitems) break;
        const int4 r0 = __ldg(reinterpret_cast<const int4*>(ga.items + (int64_t)item * 8));
        const int4 r1 = __ldg(reinterpret_cast<const int4*>(ga.items + (int64_t)item * 8) + 1);
        const int4 cc = __ldg(reinterpret_cast<const int4*>(ga.cells + (int64_t)item * 4));
        const int slot = r0.x, i0 = r0.y, n0 = r0.z, i1 = r0.w, n1 = r1.x, i2 = r1.y, n2 = r1.z, lev = r1.w;
        __syncwarp();
        // net -> shared memory; local coordinates s = (u - lo) / half - 1 of the box's axis values
        const float* src = ga.nets + (int64_t)slot * 4 * TS;
        for (int e = lane; e < CPD * TS; e += 32) M[e / TS][e % TS] = __ldg(src + e);
        {
            const int cell[3] = {cc.x, cc.y, cc.z};
            const int first[3] = {i0, i1, i2}, cnt[3] = {n0, n1, n2};
#pragma unroll
            for (int dim = 0; dim < 3; ++dim) {
                if (dim == 2 && N2 == 1) { if (lane == 0) S[2][0] = 0.0f; continue; }
                if (lane < cnt[dim]) {
                    const float* row = sp.poly + (int64_t)(sp.poly_off[lev][dim] + cell[dim]) * 20;
                    S[dim][lane] = fmaf(__ldg(ga.axis[dim] + first[dim] + lane) - __ldg(row + 16), __ldg(row + 17), -1.0f);
                }
            }
        }
        __syncwarp();
        // pass over m2: P2[k][i2][m0 m1] = sum_m2 M[k][m0][m1][m2] s2^m2       (n2 * N0 * N1 polynomials)
        for (int e = lane; e < n2 * N0 * N1; e += 32) {
            const int q = e / (N0 * N1), ab = e % (N0 * N1);
            const float s2 = S[2][q];
#pragma unroll
            for (int k = 0; k < CPD; ++k) {
                float r = M[k][ab * N2 + N2 - 1];
#pragma unroll
                for (int c = N2 - 2; c >= 0; --c) r = fmaf(r, s2, M[k][ab * N2 + c]);
                P2[k][q][ab] = r;
            }
        }
        __syncwarp();
        // lane = (i1, i2) pair: Horner over m1 -> N0 coefficients per component in registers, then Horner over m0 per point
        const int npair = n1 * n2;
        const unsigned inv = (65536u + n2 - 1) / n2;   // pr / n2 for pr < 256, n2 <= 16 by multiplication
        for (int pr = lane; pr < npair; pr += 32) {
            const int j = (int)((pr * inv) >> 16), q = pr - j * n2;
            const float s1 = S[1][j];
            float c0[CPD][N0];
#pragma unroll
            for (int k = 0; k < CPD; ++k)
#pragma unroll
                for (int a = 0; a < N0; ++a) {
                    float r = P2[k][q][a * N1 + N1 - 1];
#pragma unroll
                    for (int b = N1 - 2; b >= 0; --b) r = fmaf(r, s1, P2[k][q][a * N1 + b]);
                    c0[k][a] = r;
                }
            float* o = ga.out + ((int64_t)(i1 + j) * ga.stride[1] + (int64_t)(i2 + q) * ga.stride[2] + (int64_t)i0 * ga.stride[0]) * CPD;
            const int64_t step = ga.stride[0] * CPD;
            for (int i = 0; i < n0; ++i) {
                const float s0 = S[0][i];
#pragma unroll
                for (int k = 0; k < CPD; ++k) {
                    float r = c0[k][N0 - 1];
#pragma unroll
                    for (int a = N0 - 2; a >= 0; --a) r = fmaf(r, s0, c0[k][a]);
                    o[k] = r;
                }
                o += step;
            }
        }
    }
}

template <int N0, int N1, int N2>
int launch_grid(const thb_space_desc* sp, const GridArgs* ga, int cpd, cudaStream_t st) {
    const int grid = min((ga->nitems + kGridWarps - 1) / kGridWarps, thb_num_sms() * 8);
    if (cpd == 3) k_grid_eval<N0, N1, N2, 3><<<grid, 32 * kGridWarps, 0, st>>>(*sp, *ga);
    else k_grid_eval<N0, N1, N2, 4><<<grid, 32 * kGridWarps, 0, st>>>(*sp, *ga);
    THB_LAUNCHED();
    return THB_OK;
}

}  // namespace

int thb_check_space(const thb_space_desc* sp, const char* who);

extern "C" int thb_eval_grid(const thb_space_desc* space, const float* nets, const int32_t* items, const int32_t* item_cells, int nitems,
                             const float* axis0, const float* axis1, const float* axis2, const int64_t* row_stride, int cp_dim,
                             float* out, int32_t* cursor, void* stream) {
    int rc = thb_check_space(space, "thb_eval_grid");
    if (rc) return rc;
    THB_REQUIRE(space->poly, "thb_eval_grid: needs the span polynomials (thb_space_desc::poly) and monomial nets");
    THB_REQUIRE(cp_dim == 3 || cp_dim == 4, "thb_eval_grid: cp_dim must be 3 or 4");
    THB_REQUIRE(nitems >= 0, "thb_eval_grid: negative item count");
    if (nitems == 0) return THB_OK;
    THB_REQUIRE(nets && items && item_cells && axis0 && axis1 && (axis2 || space->ndim == 2) && row_stride && out && cursor, "thb_eval_grid: null pointer");
    GridArgs ga = {};
    ga.nets = nets; ga.items = items; ga.cells = item_cells; ga.nitems = nitems;
    ga.axis[0] = axis0; ga.axis[1] = axis1; ga.axis[2] = axis2 ? axis2 : axis0;
    ga.stride[0] = row_stride[0]; ga.stride[1] = row_stride[1]; ga.stride[2] = space->ndim == 3 ? row_stride[2] : 0;
    ga.out = out; ga.cursor = cursor;
    cudaStream_t st = (cudaStream_t)stream;
    THB_CUDA(cudaMemsetAsync(cursor, 0, sizeof(int32_t), st));
    const int n0 = space->degree[0] + 1, n1 = space->degree[1] + 1, n2 = space->ndim == 3 ? space->degree[2] + 1 : 1;
#define THB_GRID_CASE(A, B, C) if (n0 == A && n1 == B && n2 == C) return launch_grid<A, B, C>(space, &ga, cp_dim, st);
    THB_GRID_CASE(2, 2, 1) THB_GRID_CASE(3, 3, 1) THB_GRID_CASE(4, 4, 1) THB_GRID_CASE(3, 4, 1) THB_GRID_CASE(4, 3, 1)
    THB_GRID_CASE(2, 2, 2) THB_GRID_CASE(3, 3, 3) THB_GRID_CASE(4, 4, 4) THB_GRID_CASE(3, 4, 3) THB_GRID_CASE(3, 3, 4)
    THB_GRID_CASE(3, 4, 4) THB_GRID_CASE(4, 3, 3) THB_GRID_CASE(4, 3, 4) THB_GRID_CASE(4, 4, 3)
#undef THB_GRID_CASE
    return thb_set_error(THB_E_UNSUPPORTED, "thb_eval_grid: no kernel for degrees (%d,%d,%d)", space->degree[0], space->degree[1], space->degree[2]);
}
