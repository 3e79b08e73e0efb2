// Dispatch of the tile kernels by tensor-product shape, and the C entry points
//   thb_basis_tiled (THB_basis_fns, THB/core.py:604-701), thb_eval_fused / thb_eval_fused_backward
//   (THB_basis_fns + THBEval, THB/torch_funcs.py:22-47).
#include <cstdlib>

#include "thb_net_impl.cuh"

int thb_check_space(const thb_space_desc* sp, const char* who);

namespace thb_tile {
#ifdef THB_ONLY_444  // tuning builds: only the cubic 3-D kernels
#define THB_SHAPES(X) X(4, 4, 4)
#else
#define THB_SHAPES(X) \
    X(2, 2, 1) X(3, 3, 1) X(4, 4, 1) X(3, 4, 1) X(4, 3, 1) X(2, 2, 2) X(3, 3, 3) X(4, 4, 4) X(3, 4, 3) X(3, 3, 4) X(3, 4, 4) X(4, 3, 3) X(4, 3, 4) X(4, 4, 3)
#endif
#define THB_DECL(A, B, C) extern const ShapeOps shape_ops_##A##B##C;
THB_SHAPES(THB_DECL)
#define THB_REF(A, B, C) &shape_ops_##A##B##C,
static const ShapeOps* const kShapes[] = {THB_SHAPES(THB_REF)};

static const ShapeOps* find_shape(const thb_space_desc* sp) {
    const int n0 = sp->degree[0] + 1, n1 = sp->degree[1] + 1, n2 = sp->ndim == 3 ? sp->degree[2] + 1 : 1;
    for (const ShapeOps* s : kShapes)
        if (s->n0 == n0 && s->n1 == n1 && s->n2 == n2) return s;
    return nullptr;
}

static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}
}  // namespace thb_tile

using namespace thb_tile;

static int check_plan(const thb_plan_desc* plan, const char* who, bool light_ok = false) {
    THB_REQUIRE(plan && plan->sorted_params && plan->items && plan->counters, "%s: null plan buffer", who);
    THB_REQUIRE(light_ok || plan->light != 1, "%s: light plan without records (thb_plan_records first)", who);
    return THB_OK;
}

static int fill_channels(Args& ar, const int32_t* channels, int nchannels, int ndim, bool* der, const char* who) {
    THB_REQUIRE(nchannels >= 1 && nchannels <= 4 && channels, "%s: 1..4 channels", who);
    *der = false;
    for (int c = 0; c < nchannels; ++c) {
        THB_REQUIRE(channels[c] >= 0 && channels[c] < (1 << ndim), "%s: bad channel mask", who);
        ar.channels[c] = channels[c];
        if (channels[c]) *der = true;
    }
    ar.nch = nchannels;
    return THB_OK;
}

extern "C" int thb_basis_tiled(const thb_space_desc* space, const thb_plan_desc* plan, const float* params,
                               const int64_t* offsets, const int32_t* channels, int nchannels, float* const* phi_out_host,
                               void* stream) {
    int rc = thb_check_space(space, "thb_basis_tiled");
    if (rc) return rc;
    if ((rc = check_plan(plan, "thb_basis_tiled"))) return rc;
    if (plan->npoints == 0) return THB_OK;
    THB_REQUIRE(params && offsets && phi_out_host, "thb_basis_tiled: null pointer");
    THB_REQUIRE(!plan->rows_in_plan_order, "thb_basis_tiled: rows_in_plan_order is not supported (PHI rows follow `offsets`)");
    const ShapeOps* ops = find_shape(space);
    if (!ops) return thb_set_error(THB_E_UNSUPPORTED, "thb_basis_tiled: no kernel for degrees (%d,%d,%d)", space->degree[0], space->degree[1], space->degree[2]);
    Args ar = {};
    bool der;
    if ((rc = fill_channels(ar, channels, nchannels, space->ndim, &der, "thb_basis_tiled"))) return rc;
    for (int c = 0; c < nchannels; ++c) {
        THB_REQUIRE(phi_out_host[c], "thb_basis_tiled: null output");
        ar.out[c] = phi_out_host[c];
    }
    ar.items = (const int4*)plan->items; ar.counters = plan->counters;
    ar.params = params; ar.sorted = plan->sorted_params; ar.offsets = offsets;
    cudaStream_t st = (cudaStream_t)stream;
    THB_CUDA(cudaMemsetAsync(plan->counters + 1, 0, sizeof(int32_t), st));
    if (!env_int("THB_BASIS_TILE_KERNEL", 0)) return ops->basis_rows(space, &ar, der ? 1 : 0, st);  // factor-form rows
    return ops->basis(space, &ar, env_int("THB_BASIS_PPT", 1), der ? 1 : 0, st);
}

extern "C" int thb_eval_fused(const thb_space_desc* space, const thb_plan_desc* plan, const float* params,
                              const float* ctrl_pts, int cp_dim, const int32_t* channels, int nchannels, float* const* out_host,
                              int formulation, float* workspace, void* stream) {
    int rc = thb_check_space(space, "thb_eval_fused");
    if (rc) return rc;
    if ((rc = check_plan(plan, "thb_eval_fused", formulation == THB_EVAL_LOCAL_NET))) return rc;
    if (plan->npoints == 0) return THB_OK;
    THB_REQUIRE(params && ctrl_pts && out_host, "thb_eval_fused: null pointer");
    THB_REQUIRE(cp_dim == 3 || cp_dim == 4, "thb_eval_fused: cp_dim must be 3 or 4 (pad other widths to 4)");
    const ShapeOps* ops = find_shape(space);
    if (!ops) return thb_set_error(THB_E_UNSUPPORTED, "thb_eval_fused: no kernel for degrees (%d,%d,%d)", space->degree[0], space->degree[1], space->degree[2]);
    Args ar = {};
    bool der;
    if ((rc = fill_channels(ar, channels, nchannels, space->ndim, &der, "thb_eval_fused"))) return rc;
    THB_REQUIRE(formulation >= 0 && formulation <= 2, "thb_eval_fused: unknown formulation");
    if (formulation == THB_EVAL_LOCAL_NET) {
        THB_REQUIRE(workspace, "thb_eval_fused: the local-net formulation needs a workspace (thb_net_workspace_floats)");
        // monomial nets (THB_NETS_MONOMIAL) whenever the hierarchy carries the span polynomials: no per-point Cox-de Boor
        const int flags = (der ? THB_NETS_DERIVATIVES : 0) | ((space->poly && env_int("THB_NET_MONO", 1)) ? THB_NETS_MONOMIAL : 0);
        if ((rc = thb_cell_nets(space, plan, ctrl_pts, cp_dim, flags, workspace, stream))) return rc;
        return thb_eval_nets(space, plan, workspace, flags, cp_dim, channels, nchannels, out_host, stream);
    }
    THB_REQUIRE(!plan->rows_in_plan_order, "thb_eval_fused: rows_in_plan_order is only supported by the local-net formulation");
    for (int c = 0; c < nchannels; ++c) {
        THB_REQUIRE(out_host[c], "thb_eval_fused: null output");
        ar.out[c] = out_host[c];
    }
    ar.items = (const int4*)plan->items; ar.counters = plan->counters;
    ar.params = params; ar.sorted = plan->sorted_params; ar.ctrl_pts = ctrl_pts; ar.cellnet = formulation == THB_EVAL_TILE_NET ? 1 : 0;
    cudaStream_t st = (cudaStream_t)stream;
    THB_CUDA(cudaMemsetAsync(plan->counters + 1, 0, sizeof(int32_t), st));
    return ops->fused(space, &ar, env_int("THB_FUSED_PPT", 2), cp_dim, der ? 1 : 0, st);
}

extern "C" int64_t thb_net_workspace_floats(const thb_space_desc* space, int flags) {
    if (!space) return 0;
    // monomial nets serve every channel from one set of planes (see thb_b200.h)
    return (int64_t)space->nslots * (((flags & THB_NETS_DERIVATIVES) && !(flags & THB_NETS_MONOMIAL)) ? 8 : 4) * space->tile_stride;
}

extern "C" int thb_cell_nets(const thb_space_desc* space, const thb_plan_desc* plan, const float* ctrl_pts, int cp_dim,
                             int flags, float* nets, void* stream) {
    int rc = thb_check_space(space, "thb_cell_nets");
    if (rc) return rc;
    THB_REQUIRE(ctrl_pts && nets, "thb_cell_nets: null pointer");
    THB_REQUIRE(cp_dim == 3 || cp_dim == 4, "thb_cell_nets: cp_dim must be 3 or 4 (pad other widths to 4)");
    THB_REQUIRE(!plan || plan->cell_count, "thb_cell_nets: plan without cell_count");
    const ShapeOps* ops = find_shape(space);
    if (!ops) return thb_set_error(THB_E_UNSUPPORTED, "thb_cell_nets: no kernel for degrees (%d,%d,%d)", space->degree[0], space->degree[1], space->degree[2]);
    NetArgs na = {};
    THB_REQUIRE(!(flags & THB_NETS_MONOMIAL) || space->poly, "thb_cell_nets: the monomial form needs the span polynomials (thb_space_desc::poly)");
    na.ctrl_pts = ctrl_pts; na.cell_count = plan ? plan->cell_count : nullptr; na.nets = nets;
    na.mono = (flags & THB_NETS_MONOMIAL) ? 1 : 0;
    na.planes = ((flags & THB_NETS_DERIVATIVES) && !na.mono) ? 8 : 4;
    na.ctas_per_sm = (flags >> 8) & 0xff;   // THB_NETS_CTAS(k): leave room for a concurrent kernel
    return ops->nets(space, &na, cp_dim, (cudaStream_t)stream);
}

extern "C" int thb_eval_nets(const thb_space_desc* space, const thb_plan_desc* plan, const float* nets, int flags,
                             int cp_dim, const int32_t* channels, int nchannels, float* const* out_host, void* stream) {
    int rc = thb_check_space(space, "thb_eval_nets");
    if (rc) return rc;
    if ((rc = check_plan(plan, "thb_eval_nets", true))) return rc;
    if (plan->npoints == 0) return THB_OK;
    THB_REQUIRE(nets && out_host, "thb_eval_nets: null pointer");
    THB_REQUIRE(cp_dim == 3 || cp_dim == 4, "thb_eval_nets: cp_dim must be 3 or 4 (pad other widths to 4)");
    const ShapeOps* ops = find_shape(space);
    if (!ops) return thb_set_error(THB_E_UNSUPPORTED, "thb_eval_nets: no kernel for degrees (%d,%d,%d)", space->degree[0], space->degree[1], space->degree[2]);
    Args ar = {};
    bool der;
    if ((rc = fill_channels(ar, channels, nchannels, space->ndim, &der, "thb_eval_nets"))) return rc;
    const int with_derivatives = (flags & THB_NETS_DERIVATIVES) ? 1 : 0, mono = (flags & THB_NETS_MONOMIAL) ? 1 : 0;
    THB_REQUIRE(!der || with_derivatives || mono, "thb_eval_nets: derivative channels need nets built with THB_NETS_DERIVATIVES (or monomial nets)");
    THB_REQUIRE(!mono || space->poly, "thb_eval_nets: the monomial form needs the span polynomials (thb_space_desc::poly)");
    for (int c = 0; c < nchannels; ++c) {
        THB_REQUIRE(out_host[c], "thb_eval_nets: null output");
        ar.out[c] = out_host[c];
    }
    ar.items = (const int4*)plan->items; ar.counters = plan->counters; ar.sorted = plan->sorted_params;
    ar.rows_sorted = plan->rows_in_plan_order ? 1 : 0;
    if (plan->light == 1) {   // light plan: the kernel gathers the parameters through the permutation (plan->scratch)
        THB_REQUIRE(plan->params && plan->scratch, "thb_eval_nets: a light plan needs thb_plan_desc::params");
        THB_REQUIRE(plan->points_per_item <= 256, "thb_eval_nets: light plans need points_per_item <= 256");
        THB_REQUIRE(mono, "thb_eval_nets: light plans are evaluated against monomial nets");
        ar.perm = plan->scratch; ar.params = plan->params; ar.ndim = space->ndim;
    }
    cudaStream_t st = (cudaStream_t)stream;
    THB_CUDA(cudaMemsetAsync(plan->counters + 1, 0, sizeof(int32_t), st));
    return ops->net_eval(space, &ar, nets, env_int("THB_NET_PPT", mono ? 4 : 2), cp_dim, mono ? (der ? 1 : 0) : with_derivatives, mono, st);
}

static int fused_backward_impl(const thb_space_desc* space, const thb_plan_desc* plan, const float* params, const float* grad_out,
                               int cp_dim, const int32_t* row_map, float* grad, int64_t nrows, float* workspace, void* stream,
                               const char* who) {
    int rc = thb_check_space(space, who);
    if (rc) return rc;
    if ((rc = check_plan(plan, who))) return rc;
    THB_REQUIRE(cp_dim == 3 || cp_dim == 4, "%s: cp_dim must be 3 or 4", who);
    THB_REQUIRE(grad && nrows > 0, "%s: null gradient buffer", who);
    cudaStream_t st = (cudaStream_t)stream;
    THB_CUDA(cudaMemsetAsync(grad, 0, sizeof(float) * nrows * cp_dim, st));
    if (plan->npoints == 0) return THB_OK;
    THB_REQUIRE(params && grad_out, "%s: null pointer", who);
    const ShapeOps* ops = find_shape(space);
    if (!ops) return thb_set_error(THB_E_UNSUPPORTED, "%s: no kernel for degrees (%d,%d,%d)", who, space->degree[0], space->degree[1], space->degree[2]);
    Args ar = {};
    ar.items = (const int4*)plan->items; ar.counters = plan->counters;
    ar.params = params; ar.sorted = plan->sorted_params; ar.grad_out = grad_out; ar.grad_cp = grad;
    THB_CUDA(cudaMemsetAsync(plan->counters + 1, 0, sizeof(int32_t), st));
    THB_REQUIRE(workspace || !plan->rows_in_plan_order, "%s: rows_in_plan_order needs the workspace (local nets) path", who);
    THB_REQUIRE(workspace || !row_map, "%s: compact gradient rows need the workspace (local nets) path", who);
    if (workspace) {  // through the per-cell net gradients (thb_net_impl.cuh); workspace: thb_net_workspace_floats(space, 0)
        THB_REQUIRE(plan->cell_count, "%s: plan without cell_count", who);
        // (a self-cleaning workspace -- k_cell_unfold zeroing the cells it consumes instead of this memset -- was measured:
        // the fused backward went from 0.55 to 0.81 ms on config 3)
        THB_CUDA(cudaMemsetAsync(workspace, 0, sizeof(float) * (size_t)thb_net_workspace_floats(space, 0), st));
        NetArgs na = {};
        na.cell_count = plan->cell_count; na.nets = workspace; na.planes = 4; na.row_map = row_map;
        na.mono = (space->poly && env_int("THB_NET_MONO", 1)) ? 1 : 0;   // per-item sums in the monomial basis of the cell
        ar.rows_sorted = plan->rows_in_plan_order ? 1 : 0;
        return ops->net_backward(space, &ar, &na, grad, cp_dim, st);
    }
    return ops->backward(space, &ar, cp_dim, st);
}

extern "C" int thb_eval_fused_backward(const thb_space_desc* space, const thb_plan_desc* plan, const float* params,
                                       const float* grad_out, int cp_dim, float* grad_ctrl_pts, int64_t nctrl, float* workspace,
                                       void* stream) {
    return fused_backward_impl(space, plan, params, grad_out, cp_dim, nullptr, grad_ctrl_pts, nctrl, workspace, stream,
                               "thb_eval_fused_backward");
}

extern "C" int thb_eval_fused_backward_rows(const thb_space_desc* space, const thb_plan_desc* plan, const float* params,
                                            const float* grad_out, int cp_dim, const int32_t* row_map, float* grad_rows,
                                            int64_t nrows, float* workspace, void* stream) {
    THB_REQUIRE(row_map, "thb_eval_fused_backward_rows: null row map");
    return fused_backward_impl(space, plan, params, grad_out, cp_dim, row_map, grad_rows, nrows, workspace, stream,
                               "thb_eval_fused_backward_rows");
}

extern "C" int thb_eval_fused_backward_ordered(const thb_space_desc* space, const thb_plan_desc* plan, const float* params,
                                               const float* grad_out, int cp_dim, const thb_ordered_desc* od, float* grad,
                                               int64_t grad_rows, float* workspace, void* stream) {
    const char* who = "thb_eval_fused_backward_ordered";
    int rc = thb_check_space(space, who);
    if (rc) return rc;
    if ((rc = check_plan(plan, who))) return rc;
    THB_REQUIRE(cp_dim == 3 || cp_dim == 4, "%s: cp_dim must be 3 or 4", who);
    THB_REQUIRE(od && od->item_ws && od->item0 && od->contrib && od->tr_ptr && od->tr_idx && od->nrows > 0 && od->nsupp > 0, "%s: incomplete ordered descriptor", who);
    THB_REQUIRE(grad && grad_rows >= (od->active_ids ? 1 : od->nrows), "%s: gradient buffer too small", who);
    THB_REQUIRE(workspace && plan->cell_count, "%s: needs the net-gradient workspace and a plan with cell_count", who);
    cudaStream_t st = (cudaStream_t)stream;
    THB_CUDA(cudaMemsetAsync(grad, 0, sizeof(float) * grad_rows * cp_dim, st));
    if (plan->npoints == 0) return THB_OK;
    THB_REQUIRE(params && grad_out, "%s: null pointer", who);
    const ShapeOps* ops = find_shape(space);
    if (!ops) return thb_set_error(THB_E_UNSUPPORTED, "%s: no kernel for degrees (%d,%d,%d)", who, space->degree[0], space->degree[1], space->degree[2]);
    Args ar = {};
    ar.items = (const int4*)plan->items; ar.counters = plan->counters;
    ar.params = params; ar.sorted = plan->sorted_params; ar.grad_out = grad_out; ar.grad_cp = grad;
    ar.rows_sorted = plan->rows_in_plan_order ? 1 : 0;
    ar.item_ws = od->item_ws;
    THB_CUDA(cudaMemsetAsync(plan->counters + 1, 0, sizeof(int32_t), st));
    THB_CUDA(cudaMemsetAsync(workspace, 0, sizeof(float) * (size_t)thb_net_workspace_floats(space, 0), st));
    THB_CUDA(cudaMemsetAsync(od->contrib, 0, sizeof(float) * (size_t)od->nsupp * cp_dim, st));   // cells without points contribute zero
    NetArgs na = {};
    na.cell_count = plan->cell_count; na.nets = workspace; na.planes = 4; na.contrib = od->contrib;
    na.mono = (space->poly && env_int("THB_NET_MONO", 1)) ? 1 : 0;
    return ops->net_backward_ordered(space, &ar, &na, od, plan->cell_count, plan->points_per_item, plan->max_items, grad, cp_dim, st);
}
