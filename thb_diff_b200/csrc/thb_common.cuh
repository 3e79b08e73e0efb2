// Shared host/device helpers of libthb_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "thb_b200.h"

// ---- host side -------------------------------------------------------------------------------------------
int thb_set_error(int code, const char* fmt, ...);
void thb_count_launch(int n = 1);
#define THB_MAX_DEVICES 64
int thb_device();   // current device ordinal (0 if it cannot be read)
int thb_num_sms();  // of the current device

#define THB_REQUIRE(cond, ...)                                  \
    do {                                                        \
        if (!(cond)) return thb_set_error(THB_E_INVALID, __VA_ARGS__); \
    } while (0)

#define THB_CUDA(call)                                                                         \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) return thb_set_error(THB_E_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

// after a kernel launch: launch-configuration errors only (never synchronises)
#define THB_LAUNCHED()                                                                              \
    do {                                                                                            \
        cudaError_t e_ = cudaGetLastError();                                                        \
        if (e_ != cudaSuccess) return thb_set_error(THB_E_CUDA, "kernel launch: %s", cudaGetErrorString(e_)); \
        thb_count_launch();                                                                         \
    } while (0)

// ---- device side -----------------------------------------------------------------------------------------
#define THB_DEV __device__ __forceinline__

// packed FP32x2 arithmetic (Blackwell FFMA2): d = a * b + d on both halves of a 64-bit register pair
typedef unsigned long long f32x2;
THB_DEV f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
THB_DEV void unpack2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
THB_DEV void ffma2(f32x2& d, f32x2 a, f32x2 b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b)); }
THB_DEV f32x2 fmul2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
THB_DEV float hsum2(f32x2 v) {
    float lo, hi;
    unpack2(v, lo, hi);
    return lo + hi;
}

// Span rule of the reference (THB/bspline_funcs.py:77-83): searchsorted(U, u, 'right') - 1, then
// idx > n -> n, u == U[last] -> n - 1; u < U[0] gives -1.  n = nknots - degree - 1.  NaN -> -1.
THB_DEV int span_right(const float* __restrict__ U, int nknots, int degree, float u) {
    int lo = 0, hi = nknots;  // count of knots <= u
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (U[mid] <= u) lo = mid + 1; else hi = mid;
    }
    int n = nknots - degree - 1;
    int idx = lo - 1;
    if (idx > n) idx = n;
    if (u == U[nknots - 1]) idx = n - 1;
    return idx;
}

// Cox-de Boor (NURBS-book A2.2, the recursion of basisFun, THB/bspline_funcs.py:86-101) on the 2P knots
// around one cell: K[m] = U[cell + 1 + m], m = 0..2P-1 (span = cell + P).  Fully unrolled, register
// resident.  dN = first derivatives via N'_{j,P} = P (N_{j,P-1}/(U[j+P]-U[j]) - N_{j+1,P-1}/(U[j+P+1]-U[j+1])),
// whose quotients are exactly the `t` values of the last recursion step.
template <int P, bool DER>
THB_DEV void cox_de_boor(float u, const float* K, float* N, float* dN) {
    float left[P + 1], right[P + 1];
    N[0] = 1.0f;
#pragma unroll
    for (int j = 1; j <= P; ++j) {
        left[j] = u - K[P - j];
        right[j] = K[P - 1 + j] - u;
        float carry = 0.0f;
        float prev_t = 0.0f;
#pragma unroll
        for (int r = 0; r < j; ++r) {
            float t = N[r] / (right[r + 1] + left[j - r]);
            if (DER && j == P) {
                dN[r] = float(P) * (prev_t - t);
                prev_t = t;
            }
            N[r] = carry + right[r + 1] * t;
            carry = left[j - r] * t;
        }
        if (DER && j == P) dN[P] = float(P) * prev_t;
        N[j] = carry;
    }
    if (DER && P == 0) dN[0] = 0.0f;
}

// Runtime-degree version (degree <= 5) used by the 1-D API kernels and the legacy dense path.
// U = full knot vector, span = knot span index.
#define THB_MAX_DEG 5
THB_DEV void cox_de_boor_rt(int p, int span, float u, const float* __restrict__ U, float* N, float* dN) {
    float left[THB_MAX_DEG + 1], right[THB_MAX_DEG + 1];
    N[0] = 1.0f;
    if (dN) dN[0] = 0.0f;
    for (int j = 1; j <= p; ++j) {
        left[j] = u - U[span + 1 - j];
        right[j] = U[span + j] - u;
        float carry = 0.0f, prev_t = 0.0f;
        for (int r = 0; r < j; ++r) {
            float t = N[r] / (right[r + 1] + left[j - r]);
            if (dN && j == p) {
                dN[r] = float(p) * (prev_t - t);
                prev_t = t;
            }
            N[r] = carry + right[r + 1] * t;
            carry = left[j - r] * t;
        }
        if (dN && j == p) dN[p] = float(p) * prev_t;
        N[j] = carry;
    }
}

// Finest-level span per dimension, then the dyadic walk to the finest level whose cell is active
// (compute_active_span, THB/core.py:157-189).  Returns the slot id or -1.
THB_DEV int locate_slot(const thb_space_desc& sp, const float* u) {
    const int L = sp.nlevels - 1;
    int cf[3] = {0, 0, 0};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (k < sp.ndim) {
            int s = span_right(sp.knots + sp.knot_off[L][k], sp.knot_len[L][k], sp.degree[k], u[k]);
            int c = s - sp.degree[k];
            if (c < 0 || c >= sp.ncells[L][k]) return -1;
            cf[k] = c;
        }
    }
    for (int lev = L; lev >= 0; --lev) {
        int sh = L - lev;
        int64_t lin = cf[0] >> sh;
        lin = lin * sp.ncells[lev][1] + (cf[1] >> sh);
        if (sp.ndim == 3) lin = lin * sp.ncells[lev][2] + (cf[2] >> sh);
        int slot = __ldg(sp.cell_slot + sp.cell_slot_off[lev] + lin);
        if (slot >= 0) return slot;
    }
    return -1;
}

// ---- cp.async (Ampere-style asynchronous global -> shared copies; no register staging) ----------------------------
THB_DEV void cp_async16(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
THB_DEV void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
THB_DEV void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

THB_DEV void cp_async4(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem) : "memory");
}
THB_DEV void cp_async16_cg(void* smem, const void* gmem) {  // streaming data: bypass L1
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}

// ---- bulk asynchronous copies (TMA, 1-D: cp.async.bulk global -> shared, SASS UBLKCP) completing on an mbarrier ---------
// One lane arms the barrier with the byte count and issues the copies; every lane of the consumer waits on the phase
// parity.  Source and destination 16-byte aligned, size a multiple of 16.
THB_DEV void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
THB_DEV void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
THB_DEV void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
THB_DEV void bulk_g2s(void* smem, const void* gmem, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (unsigned)__cvta_generic_to_shared(smem)),
                 "l"(gmem), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
}
THB_DEV bool mbar_try_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
THB_DEV void mbar_wait(unsigned long long* bar, unsigned parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// CSR contraction from bulk-staged segments (thb_eval_bulk.cu); returns THB_OK, an error, or 1 = "not applicable here"
// (misaligned arrays, switched off): the caller then runs the register-path kernels of thb_eval_csr.cu
int thb_csr_forward_bulk(const float* cp, const void* jm, int jm_is_i64, const float* phi, const void* cumsum, int cumsum_is_i64,
                         float* out, int64_t npoints, int cp_dim, cudaStream_t st);
int thb_csr_backward_bulk(const float* grad_out, const void* jm, int jm_is_i64, const float* phi, const void* cumsum, int cumsum_is_i64,
                          float* grad_ctrl_pts, int64_t npoints, int cp_dim, cudaStream_t st);
