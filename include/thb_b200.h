/* thb_b200.h -- C ABI of libthb_b200.so: the B200 (sm_100a) implementation of THB-Diff's evaluation hot path.
 *
 * Conventions (all entry points):
 *   - every pointer is a DEVICE pointer unless the name ends in _host; sizes are element counts
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*); nothing synchronises, nothing allocates
 *   - return value: 0 = ok, negative = THB_E_* (thb_last_error() gives the text for the calling thread)
 *   - no torch / C++ types cross this boundary; the reference-side bindings are shown in INTEGRATION.md
 *
 * Each function names the reference interface it replaces (paths relative to the reference repo).
 */
#ifndef THB_B200_H
#define THB_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define THB_API __attribute__((visibility("default")))
#else
#define THB_API
#endif

#define THB_MAX_LEVELS 8
#define THB_INFO_W 12 /* ints per cellinfo row: level, c0, c1, c2, supp_off, S, tile_off, n_direct, sep_off, n_sep, full_off, 0 */

enum {
    THB_OK = 0,
    THB_E_INVALID = -1,     /* bad argument (null pointer, size, unsupported degree combination ...) */
    THB_E_CUDA = -2,        /* a CUDA runtime call failed; see thb_last_error() */
    THB_E_UNSUPPORTED = -3, /* valid request the library has no kernel for */
};

/* Channel bit masks for thb_basis_tiled / thb_eval_fused: bit k set = use the first derivative of the 1-D
 * basis in dimension k.  0 = values; (1<<k) = d/du_k (true gradient component); all bits = the reference's
 * return_mode=[False,True] "derivative" (mixed partial, THB/core.py:686-701). */
#define THB_CH_VALUE 0

/* Packed hierarchy descriptor (see thb_diff_b200/packed.py for the producer). Passed by pointer from the
 * host; the library copies it into kernel arguments.  Replaces the dict-of-levels inputs
 * (knotvectors, cells, ac_cells_ac_supp, fn_coeffs, fn_shapes) of THB/core.py:157-189 and :604-701. */
typedef struct thb_space_desc {
    int32_t ndim;                           /* 2 or 3 */
    int32_t nlevels;                        /* <= THB_MAX_LEVELS */
    int32_t degree[3];                      /* degree[2] = 0 in 2-D */
    int32_t tile_stride;                    /* floats per tile row: (p+1)^d rounded up to a multiple of 4 */
    int32_t nslots;                         /* number of active cells */
    int32_t max_tiled;                      /* max number of tiled supports of any cell */
    int32_t knot_off[THB_MAX_LEVELS][3];    /* offset of each 1-D knot vector inside `knots` */
    int32_t knot_len[THB_MAX_LEVELS][3];
    int32_t ncells[THB_MAX_LEVELS][3];
    int32_t nfns[THB_MAX_LEVELS][3];
    int64_t fn_off[THB_MAX_LEVELS + 1];     /* level offsets of the flat function / control-point numbering */
    int64_t cell_slot_off[THB_MAX_LEVELS];  /* level offsets inside `cell_slot` */
    const float* knots;                     /* f32, all levels and dims */
    const int32_t* cell_slot;               /* per cell: slot id of the active cell, or -1 */
    const int32_t* cellinfo;                /* [nslots][THB_INFO_W] */
    const uint64_t* dmask;                  /* [nslots] own-level window activity bits */
    const int32_t* supp_jm;                 /* CSR values: flat ids of each cell's supports, reference order */
    const float* tiles;                     /* [ntiles][tile_stride] */
    /* Split of each cell's tiled supports for the fused kernel: supports whose tile is rank one (an outer product of
     * one (p_k+1)-vector per dimension -- every untruncated function and every function truncated across a flat face of
     * a refinement region) are stored as 12 floats (3 x 4, zero padded), the rest as full tiles.  sep_off / n_sep /
     * full_off of a cell are in cellinfo; n_full = S - n_direct - n_sep.  All four may be NULL (then `tiles` is used). */
    const float* sep_vec;                   /* [nsep_total][12] */
    const int32_t* sep_jm;                  /* [nsep_total] flat support ids */
    const float* full_tiles;                /* [nfull_total][tile_stride] */
    const int32_t* full_jm;                 /* [nfull_total] */
    const int32_t* span_lut;                /* optional (may be NULL): per dimension lut_n[k] uniform bins over the domain;
                                               entry = finest-level cell that contains the bin's left edge */
    int32_t lut_off[3];
    int32_t lut_n[3];
    float lut_scale[3];                     /* bins per unit parameter: bin = (u - U[p]) * lut_scale */
    int32_t lut_exact;                      /* 1: the producer proved, with the same float32 bin formula evaluated at every
                                               breakpoint, that the cell of any u of a bin is span_lut[bin] or the next one
                                               (the plan then skips the verification and the fix-up pass) */
    const int32_t* finest_slot;             /* optional (may be NULL): per FINEST-level cell the slot of the active cell
                                               that contains it (-1 if none) -- one lookup instead of one per level */
    const int32_t* tiled_ref;               /* optional (NULL without the split tables): for tiled support j of the global
                                               tile order (tile_off + position in the cell's reference-ordered list):
                                               (row of sep_vec << 1) or (row of full_tiles << 1) | 1 */
    const float* poly;                      /* optional (may be NULL): per level, dimension and 1-D cell a row of 20 floats --
                                               the p+1 B-splines of the cell as polynomials in s = (u - lo) / half - 1,
                                               N_r(u) = sum_m C[m][r] s^m, C as [m][r] padded to 4 x 4, then (lo, 1/half, 0, 0).
                                               Used by the monomial form of the local control nets (thb_cell_nets with THB_NETS_MONOMIAL) */
    int32_t poly_off[THB_MAX_LEVELS][3];    /* first row of each (level, dimension) */
} thb_space_desc;

/* A plan = the points of one batch grouped by active cell. All arrays are caller-allocated device memory:
 *   slot [N] i32, scratch [N] i32, sorted_params [N][4] f32 = the points in cell order as 16-byte records
 *   (u0, u1, u2 or 0, caller index i as raw bits: record j of cell c, cell_start[c] <= j < cell_start[c+1], is point i),
 *   cell_count/cell_start/cell_cursor [nslots+1] i32 (16-byte aligned: read / written as int4),
 *   items [max_items][16] i32 (slot, begin, count, level, c0, c1, c2, supp_off, S, tile_off, n_direct, n_sep,
 *   dmask lo, dmask hi, sep_off, full_off), counters [4] i32 (counters[0] = number of items).
 * max_items >= N / points_per_item + nslots. */
typedef struct thb_plan_desc {
    int64_t npoints;
    int32_t points_per_item;
    int32_t max_items;
    int32_t* slot;
    int32_t* scratch;
    float* sorted_params;
    int32_t* cell_count;
    int32_t* cell_start;
    int32_t* cell_cursor;
    int32_t* items;
    int32_t* counters;
    int32_t rows_in_plan_order;  /* 0: row i of out / grad_out belongs to point i of params (the reference's layout).
                                    1: row j belongs to the point of record j of sorted_params -- coalesced rows for callers
                                    that keep their per-point data in plan order (a fit loop permutes its targets once);
                                    honoured by thb_eval_nets / thb_eval_fused (local nets) / thb_eval_fused_backward with
                                    a workspace, an error elsewhere */
    int32_t light;               /* 0: thb_plan_build writes the 16-byte records (sorted_params).
                                    1: LIGHT plan -- thb_plan_build writes only the permutation (scratch[j] = caller index of the
                                       point in place j of the cell order, 4 bytes per point instead of 16 written and 12 read);
                                       thb_eval_nets gathers the parameters through it (needs `params` below); every other
                                       consumer of a plan needs thb_plan_records first.
                                    2: light plan whose records have been materialised by thb_plan_records (set by the caller) */
    const float* params;         /* the batch the plan was built from (same pointer as passed to thb_plan_build); read by
                                    thb_eval_nets when light == 1 */
} thb_plan_desc;

THB_API int thb_version(void);
THB_API const char* thb_last_error(void);

/* ---- L5 boundary: the torch extension THB_eval ------------------------------------------------------------
 * thb_eval_forward  replaces THB_eval.forward / cpp_forward
 *     (THB/THB_extensions/compute_pts_cuda_kernels.cu:3-26,49-71; compute_pts.cpp:4-29)
 *     out[i, :] = sum_{j in [cumsum[i-1], cumsum[i])} phi[j] * ctrl_pts[jm[j], :]   (cumsum inclusive)
 * thb_eval_backward replaces THB_eval.backward / cpp_backward
 *     (compute_pts_cuda_kernels.cu:28-47,73-97; compute_pts.cpp:31-54)
 *     grad_ctrl_pts[jm[j], :] += phi[j] * grad_out[i, :];  grad_ctrl_pts is zeroed by the callee.
 * jm / cumsum may be int64 (reference dtype) or int32; cp_dim in 1..4 (reference: 3). */
THB_API int thb_eval_forward(const float* ctrl_pts, const void* jm, int jm_is_i64, const float* phi, const void* cumsum,
                     int cumsum_is_i64, float* out, int64_t npoints, int cp_dim, void* stream);
THB_API int thb_eval_backward(const float* grad_out, const void* jm, int jm_is_i64, const float* phi, const void* cumsum,
                      int cumsum_is_i64, float* grad_ctrl_pts, int64_t npoints, int64_t nctrl, int cp_dim,
                      void* stream);

/* ---- 1-D primitives ----------------------------------------------------------------------------------------
 * thb_find_spans replaces find_span_array_jax (THB/bspline_funcs.py:77-83): searchsorted(U,u,'right')-1,
 *     >n -> n, u==U[-1] -> n-1, u<U[0] -> -1, n = nknots-degree-1.  params are read with element stride
 *     `stride` (so a column of an [N,d] array can be passed).
 * thb_basis_1d replaces basis_fns_vmap / der_basis_fns_vmap (bspline_funcs.py:185-194) using the O(p^2)
 *     recursion of basisFun (bspline_funcs.py:86-101): values [N, degree+1] and (if non-null) first
 *     derivatives of the functions span-degree .. span.  Out-of-domain points produce zeros. degree <= 5. */
THB_API int thb_find_spans(const float* params, int64_t stride, int64_t npoints, const float* knots, int nknots, int degree,
                   int32_t* spans, void* stream);
THB_API int thb_basis_1d(const float* params, int64_t stride, int64_t npoints, const float* knots, int nknots, int degree,
                 float* values, float* derivs, void* stream);

/* compute_tensor_product (THB/bspline_funcs.py:54-59): outer product of 2 or 3 vectors, batched:
 * out[q][i][j][k] = a[q][i] * b[q][j] * c[q][k] for q < batch (c == NULL and nc == 1: the 2-D form, out[q][i][j]). */
THB_API int thb_tensor_product(const float* a, const float* b, const float* c, int64_t batch, int na, int nb, int nc, float* out,
                       void* stream);

/* ---- active-cell lookup: compute_active_span (THB/core.py:157-189) ----------------------------------------
 * For every point the finest-level span per dimension, then the finest level whose cell is active.
 * slot[i] = slot id of that cell (-1 for points outside the domain); num_supp[i] (nullable) = its number of
 * supports; params is [N, ndim] row-major f32. */
THB_API int thb_active_span(const thb_space_desc* space, const float* params, int64_t npoints, int32_t* slot,
                    int32_t* num_supp, void* stream);

/* Expands the per-cell CSR to per-point flat support ids in the reference's pair order
 * (Jm of THB/core.py:644-650 and THB/torch_funcs.py:72-78). offsets = exclusive prefix sum of num_supp. */
THB_API int thb_expand_supports(const thb_space_desc* space, const int32_t* slot, const int64_t* offsets, int64_t npoints,
                        void* jm_out, int jm_is_i64, void* stream);

/* Groups the points of a batch by active cell (counting sort) and cuts the groups into work items.
 * Computes plan->slot itself (same result as thb_active_span). */
THB_API int thb_plan_build(const thb_space_desc* space, const float* params, const thb_plan_desc* plan, void* stream);

/* The same in up to three calls, so that a caller can put other work between or beside them.  `phases` is a combination of
 * THB_PLAN_HISTOGRAM (plan->slot and cell_count -- everything thb_cell_nets needs to know which cells hold points),
 * THB_PLAN_ITEMS (scan + work items) and THB_PLAN_SCATTER (the points into cell order), issued in that order over the calls.
 * engine.plan_and_nets runs the nets kernel on a second stream beside the phases after the histogram. */
#define THB_PLAN_HISTOGRAM 1
#define THB_PLAN_ITEMS 2
#define THB_PLAN_SCATTER 4
THB_API int thb_plan_build_phase(const thb_space_desc* space, const float* params, const thb_plan_desc* plan, int phases,
                         void* stream);

/* ---- THB_basis_fns (THB/core.py:604-701; float64 statement: compute_THB_fns_tp, core.py:203-268) -----------
 * phi_out[c][offsets[i] + s] for every channel c < nchannels, point i and support s of the point's cell. */
THB_API int thb_basis_tiled(const thb_space_desc* space, const thb_plan_desc* plan, const float* params,
                    const int64_t* offsets, const int32_t* channels, int nchannels, float* const* phi_out_host,
                    void* stream);

/* Legacy form of THB_basis_fns for callers that hold the reference's dense tables: coeffs is
 * [nfn_total, prod(nfns_finest)] f32 (THB/core.py:636-643), jm/cumsum as in thb_eval_forward, evaluation at the
 * finest level exactly as THB/core.py:652-677 (with the SURVEY Q2/Q3 repairs).  nnz = cumsum[npoints-1] is read by
 * the kernel itself: like every other entry point this one never synchronises. */
THB_API int thb_basis_dense(const float* params, int64_t npoints, int ndim, const int32_t* degree, const float* knots,
                    const int32_t* knot_off, const int32_t* knot_len, const int32_t* nfns_finest,
                    const float* coeffs, const void* jm, int jm_is_i64, const int64_t* cumsum, int channel,
                    float* phi_out, void* stream);

/* Materialises the 16-byte records of a LIGHT plan (light == 1) from its permutation: sorted_params[j] = (params[perm[j]],
 * perm[j]).  The caller then sets plan->light = 2. */
THB_API int thb_plan_records(const thb_space_desc* space, const float* params, const thb_plan_desc* plan, void* stream);

/* ---- fused basis + evaluation (PHI never leaves the SM) ----------------------------------------------------
 * out_host[c] is a device buffer [N, cp_dim] per channel; rows of points outside the domain (slot -1) are NOT written,
 * pass zero-filled buffers if such points can occur.  Replaces THB_basis_fns followed by THBEval.forward
 * (THB/torch_funcs.py:22-33).  formulation selects how sum_s PHI_s(u) CP_s is evaluated (same result to rounding):
 *   THB_EVAL_LOCAL_NET (0, default)  thb_cell_nets for the cells that hold points of the plan + thb_eval_nets (below);
 *                                    needs `workspace` (thb_net_workspace_floats floats), NULL for the other two
 *   THB_EVAL_TILE_NET  (1)           the per-point tile kernel with a per-item fold of all tiles
 *   THB_EVAL_PER_POINT (2)           PHI_s(u) of every support is formed per point, then contracted -- the reference's
 *                                    gather-multiply + contraction, without PHI leaving the SM */
#define THB_EVAL_LOCAL_NET 0
#define THB_EVAL_TILE_NET 1
#define THB_EVAL_PER_POINT 2
THB_API int thb_eval_fused(const thb_space_desc* space, const thb_plan_desc* plan, const float* params,
                   const float* ctrl_pts, int cp_dim, const int32_t* channels, int nchannels, float* const* out_host,
                   int formulation, float* workspace, void* stream);

/* Local control nets.  Over one active cell  sum_s PHI_s(u) CP_s = <tp(u), W>  with tp the (p+1)^d tensor product of
 * the cell's own-level B-splines and  W[t] = [t active] CP_own[t] + sum_{coarser s} tile_s[t] CP_s : W depends on the
 * hierarchy and the control points only, so it is built once per cell (and may be kept across batches of points for as
 * long as ctrl_pts do not change) and a point costs Cox-de Boor plus one sum-factorised contraction with W.
 * nets: [nslots][planes][tile_stride] f32, planes = 4 (value channel only) or 8 (THB_NETS_DERIVATIVES: planes 4..7 are W
 * built from control points relative to the cell's first support, used by the derivative channels).
 * flags (the same value goes to thb_net_workspace_floats, thb_cell_nets and thb_eval_nets):
 *   THB_NETS_DERIVATIVES (1)  eight planes, derivative channels allowed
 *   THB_NETS_MONOMIAL    (2)  the nets are stored in MONOMIAL form: with the cell's B-splines written as polynomials in the
 *                             cell-local coordinate s_k = (u_k - lo_k) / half_k - 1 (thb_space_desc::poly, required),
 *                             <tp(u), W> = sum M[m0][m1][m2] s0^m0 s1^m1 s2^m2 and a point costs three subtractions plus a
 *                             nested Horner scheme (30 packed FMAs per output component for cubic 3-D) -- no Cox-de Boor.
 *                             Monomial nets are built from control points relative to the cell's first support with that
 *                             point added back to the constant term, so ONE set of four planes serves the value and every
 *                             derivative channel (THB_NETS_DERIVATIVES changes nothing for them).
 *                             thb_eval_fused uses this form whenever `poly` is present.
 * thb_cell_nets: plan == NULL builds every active cell, otherwise only the cells that hold points of the plan.
 * Together the two calls replace THB_basis_fns (THB/core.py:604-701) + prepare_data_for_CUDA_evaluation
 * (THB/torch_funcs.py:50-80) + THBEval.forward / THB_eval.forward (THB/torch_funcs.py:22-33,
 * THB_extensions/compute_pts_cuda_kernels.cu:3-26); thb_cell_nets alone takes the place of the dense coefficient tables
 * of compute_refinement_operators (THB/core.py:43-114) on the evaluation side. */
#define THB_NETS_DERIVATIVES 1
#define THB_NETS_MONOMIAL 2
#define THB_NETS_CTAS(k) (((k) & 0xff) << 8) /* thb_cell_nets only: at most k resident CTAs per SM, so that a kernel on
                                                another stream (the plan of the batch) finds room beside it */
THB_API int64_t thb_net_workspace_floats(const thb_space_desc* space, int flags);
THB_API int thb_cell_nets(const thb_space_desc* space, const thb_plan_desc* plan, const float* ctrl_pts, int cp_dim,
                  int flags, float* nets, void* stream);
THB_API int thb_eval_nets(const thb_space_desc* space, const thb_plan_desc* plan, const float* nets, int flags,
                  int cp_dim, const int32_t* channels, int nchannels, float* const* out_host, void* stream);
/* Evaluation on a TENSOR-PRODUCT GRID of parameters u0[i0] x u1[i1] (x u2[i2]) -- the point sets the reference generates with
 * generate_parametric_coordinates (THB/bspline_funcs.py:20-35) and every BASELINE configuration uses.  No plan is built: the
 * points of an active cell are one index range per axis, so the host lists work items once per grid
 * (thb_diff_b200.engine.GridPlan): items [nitems][8] i32 = slot, i0, n0, i1, n1, i2, n2, level (index box of at most 16 values per
 * axis inside ONE active cell; n2 = 1 in 2-D), item_cells [nitems][4] i32 = the cell's 1-D cell indices.  nets: MONOMIAL nets
 * of thb_cell_nets (THB_NETS_MONOMIAL, four planes).  The geometry of point (i0, i1, i2) is written to row
 * i0 * row_stride[0] + i1 * row_stride[1] + i2 * row_stride[2] of out [*, cp_dim]; rows of points outside the domain are not
 * written.  cursor: one device i32 of scratch.  Cost per point: (p0) FMAs per component plus O(n1 n2 + n2) per box (sum
 * factorisation at the cell level) -- against (p+1)^d-many per point on the general path. */
THB_API int thb_eval_grid(const thb_space_desc* space, const float* nets, const int32_t* items, const int32_t* item_cells, int nitems,
                  const float* axis0, const float* axis1, const float* axis2, const int64_t* row_stride, int cp_dim, float* out,
                  int32_t* cursor, void* stream);

/* grad_ctrl_pts[nctrl, cp_dim] (zeroed by the callee) = A^T grad_out, A = value-channel basis matrix.
 * Replaces THBEval.backward (THB/torch_funcs.py:35-47) for the fused path.  workspace: thb_net_workspace_floats(space, 0)
 * floats (overwritten) -> the gradient goes through per-cell net gradients (one reduction per work item, one transposed
 * fold per cell); NULL -> the per-item tile kernel. */
THB_API int thb_eval_fused_backward(const thb_space_desc* space, const thb_plan_desc* plan, const float* params,
                            const float* grad_out, int cp_dim, float* grad_ctrl_pts, int64_t nctrl, float* workspace,
                            void* stream);

/* Same gradient in COMPACT rows: grad_rows[row_map[f], :] receives the gradient of flat function / control-point id f.
 * Only active functions ever get one (config 3: 42 153 of 2 598 588 rows), so a fit loop keeps, all-reduces and steps
 * nrows x cp_dim values instead of nctrl x cp_dim.  row_map: [nctrl] i32, >= 0 for every support of an active cell
 * (thb_diff_b200/packed.py: PackedHierarchy.active_rows()); grad_rows [nrows, cp_dim] is zeroed by the callee;
 * workspace as above (required).  The reference has no counterpart: its backward always fills the dense
 * [nCP, 3] tensor (THB/THB_extensions/compute_pts_cuda_kernels.cu:73-97). */
THB_API int thb_eval_fused_backward_rows(const thb_space_desc* space, const thb_plan_desc* plan, const float* params,
                                 const float* grad_out, int cp_dim, const int32_t* row_map, float* grad_rows,
                                 int64_t nrows, float* workspace, void* stream);

/* ORDERED (deterministic) form of the same gradient.  The reference's backward adds with one atomicAdd per (point,
 * support) (THB/THB_extensions/compute_pts_cuda_kernels.cu:43): the result depends on the arrival order of the threads, and
 * so does the default path above (atomics per work item and per cell).  Here every sum runs in a fixed order: per-item net
 * gradients are stored, added per cell in item order, the per-(cell, support) contributions are stored and added per
 * function in the order of a transposed support list.  Bit-identical results from run to run need, in addition, the points
 * of every cell in a canonical order inside the plan (thb_diff_b200.engine.Plan.canonicalize sorts them by caller index).
 * Buffers (device): item_ws [max_items][4][tile_stride] f32, item0 [nslots] i32, contrib [nsupp][cp_dim] f32 (nsupp = total
 * length of the cells' support lists); tr_ptr [nrows + 1], tr_idx [nsupp] i32: for compact row r the slots of `contrib`
 * that belong to its function, slot = supp_off(cell) + position with the cell's supports in the order own-level (window
 * order) | rank-one tiled | full-tile (PackedHierarchy.ordered_scatter_tables); active_ids [nrows] i32 or NULL: NULL writes
 * the compact rows grad[nrows, cp_dim], otherwise grad is the dense [nctrl, cp_dim] tensor (zeroed by the callee) and row
 * active_ids[r] receives the sum.  workspace: thb_net_workspace_floats(space, 0) floats, required. */
typedef struct thb_ordered_desc {
    float* item_ws;
    int32_t* item0;
    float* contrib;
    const int32_t* tr_ptr;
    const int32_t* tr_idx;
    const int32_t* active_ids;
    int64_t nsupp;
    int32_t nrows;
    int32_t pad_;
} thb_ordered_desc;
THB_API int thb_eval_fused_backward_ordered(const thb_space_desc* space, const thb_plan_desc* plan, const float* params,
                                    const float* grad_out, int cp_dim, const thb_ordered_desc* ordered, float* grad,
                                    int64_t grad_rows, float* workspace, void* stream);

/* ---- fit loop helpers (BASELINE config 4; the reference composes these from torch ops around THBEval,
 * THB/torch_funcs.py:8-47) ----------------------------------------------------------------------------------
 * thb_mse_grad: loss[0] = scale * sum_i (out[i] - target[i])^2 and grad_out[i] = 2 scale (out[i] - target[i]) over
 *     nvalues floats in one pass.  Deterministic (fixed-order reduction).  workspace: thb_mse_workspace_bytes() bytes,
 *     zero-filled before the FIRST call (the kernel leaves it ready for the next one); buffers 16-byte aligned.
 * thb_adam_rows: torch.optim.Adam's update (betas, eps, no weight decay, no amsgrad) applied to the rows row_ids[r] of the
 *     dense parameter params[nctrl, cp_dim] with gradient grad_rows[r, :] * grad_scale and state exp_avg / exp_avg_sq
 *     [nrows, cp_dim].  Rows outside row_ids keep a zero gradient for ever, and Adam never moves such a row, so this is
 *     the dense optimiser restricted to the rows it can change.  step: 1-based step number; when step_dev (device i32)
 *     is non-null the number is step_dev[0] + 1 and the call increments it (CUDA-graph replay).
 * thb_adam_rows_peer: the all-reduce of the gradient rows and the Adam step as ONE enqueue over NVLink peer memory, no
 *     collective library call.  peer_grads[r] / peer_flags[r] (host arrays of `world` DEVICE pointers): rank r's gradient
 *     rows [nrows, cp_dim] f32 and flag array (>= world u32, zero-initialised) as mapped into THIS process (symmetric /
 *     peer memory); entry `rank` is this rank's own buffer.  Every rank stores the step number (step_dev[0] + 1) into
 *     its slot of every peer's flags, waits for all of them, then adds the peers' rows in rank order -- bitwise the same
 *     sum on every rank -- and applies Adam; step_dev is incremented.  The caller alternates between TWO gradient buffers
 *     per rank from step to step (a rank running ahead must not overwrite rows a slow peer still reads).  err[0] is set
 *     to 1 when a peer does not arrive within ~4 s (the call never hangs the GPU). */
THB_API int64_t thb_mse_workspace_bytes(void);
THB_API int thb_mse_grad(const float* out, const float* target, int64_t nvalues, float scale, float* grad_out, float* loss,
                 void* workspace, void* stream);
THB_API int thb_adam_rows(float* params, const int32_t* row_ids, const float* grad_rows, float* exp_avg, float* exp_avg_sq,
                  int64_t nrows, int cp_dim, float lr, float beta1, float beta2, float eps, float grad_scale, int step,
                  int32_t* step_dev, void* stream);

THB_API int thb_adam_rows_peer(float* params, const int32_t* row_ids, const void* const* peer_grads, void* const* peer_flags,
                       int world, int rank, float* exp_avg, float* exp_avg_sq, int64_t nrows, int cp_dim, float lr, float beta1,
                       float beta2, float eps, float grad_scale, int32_t* step_dev, int32_t* err, void* stream);

/* ---- measurement helpers -----------------------------------------------------------------------------------
 * Dependent-chain FP32 FMA micro-kernel used by bench.py to measure the FP32 roofline denominator.
 * variant 0: scalar FFMA, 1: packed FFMA2.  Enqueues one launch; flops_out_host receives the FLOP count. */
THB_API int thb_fma_peak(int variant, int iters, float* sink, double* flops_out_host, void* stream);
/* Number of kernels this library has launched in the calling process (for bench.py's gpu_launches). */
THB_API int64_t thb_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* THB_B200_H */
