// Fit-loop helpers of BASELINE config 4 (differentiable THB fitting: MSE loss, backward into the control points, Adam).
// The reference drives this loop from Python with torch ops around THBEval (THB/torch_funcs.py:8-47): the loss is a
// composed torch expression and the optimiser torch.optim.Adam on the dense [nCP, 3] parameter.  Only the ACTIVE functions
// of a hierarchy ever receive a gradient (42 153 of 2 598 588 rows on config 3), so here the gradient lives in compact rows
// (thb_eval_fused_backward_rows), is all-reduced in that form, and the optimiser touches those rows only:
//   thb_mse_grad  : loss = scale * sum (out - target)^2 and grad_out = 2 scale (out - target) in ONE pass over the data
//                   (deterministic: per-CTA partial sums, the last CTA adds them in a fixed order)
//   thb_adam_rows : torch.optim.Adam's update (no weight decay, no amsgrad) on the rows row_ids[] of the dense parameter
#include "thb_common.cuh"

namespace {

constexpr int kMseThreads = 256;

__global__ void __launch_bounds__(kMseThreads) k_mse_grad(const float* __restrict__ out, const float* __restrict__ target, int64_t n,
                                                          float scale, float* __restrict__ grad_out, double* __restrict__ partial,
                                                          unsigned* __restrict__ ticket, float* __restrict__ loss) {
    __shared__ double s_sum[kMseThreads / 32];
    __shared__ bool s_last;
    const int64_t n4 = n >> 2;
    const float g2 = 2.0f * scale;
    double acc = 0.0;
    const float4* o4 = reinterpret_cast<const float4*>(out);
    const float4* t4 = reinterpret_cast<const float4*>(target);
    float4* g4 = reinterpret_cast<float4*>(grad_out);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 a = __ldcs(o4 + i), b = __ldcs(t4 + i);
        const float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z, dw = a.w - b.w;
        g4[i] = make_float4(g2 * dx, g2 * dy, g2 * dz, g2 * dw);
        acc += (double)(dx * dx + dy * dy) + (double)(dz * dz + dw * dw);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {  // tail
        const int64_t i = (n4 << 2) + threadIdx.x;
        const float dd = out[i] - target[i];
        grad_out[i] = g2 * dd;
        acc += (double)(dd * dd);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double b = 0.0;
#pragma unroll
        for (int w = 0; w < kMseThreads / 32; ++w) b += s_sum[w];
        partial[blockIdx.x] = b;
        __threadfence();
        s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (s_last && threadIdx.x < 32) {  // the last CTA adds the partial sums in index order: same result every run
        __threadfence();
        double t = 0.0;
        for (unsigned b = threadIdx.x; b < gridDim.x; b += 32) t += partial[b];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (threadIdx.x == 0) {
            loss[0] = (float)(t * (double)scale);
            *ticket = 0u;  // ready for the next call
        }
    }
}

struct AdamArgs {
    float lr, beta1, beta2, eps, grad_scale;
    int32_t step;  // 1-based step number when step_dev is null
};

__global__ void __launch_bounds__(256) k_adam_rows(float* __restrict__ params, const int32_t* __restrict__ row_ids,
                                                   const float* __restrict__ grad_rows, float* __restrict__ exp_avg,
                                                   float* __restrict__ exp_avg_sq, int64_t nrows, int cpd, AdamArgs a,
                                                   const int32_t* __restrict__ step_dev) {
    const int64_t tot = nrows * cpd;
    const int t = step_dev ? step_dev[0] + 1 : a.step;
    // bias corrections in double, like the host-side scalars of torch.optim.Adam
    const double bc1 = 1.0 - pow((double)a.beta1, (double)t), bc2 = 1.0 - pow((double)a.beta2, (double)t);
    const float step_size = (float)((double)a.lr / bc1), bc2_sqrt = (float)sqrt(bc2);
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / cpd;
        const int k = (int)(e - r * cpd);
        const float g = grad_rows[e] * a.grad_scale;
        const float m = a.beta1 * exp_avg[e] + (1.0f - a.beta1) * g;          // lerp(m, g, 1 - beta1)
        const float v = a.beta2 * exp_avg_sq[e] + (1.0f - a.beta2) * g * g;
        exp_avg[e] = m;
        exp_avg_sq[e] = v;
        const float denom = sqrtf(v) / bc2_sqrt + a.eps;
        float* p = params + (int64_t)__ldg(row_ids + r) * cpd + k;
        *p = *p - step_size * (m / denom);
    }
}

__global__ void k_step_inc(int32_t* step_dev) { step_dev[0] += 1; }

// ---- gradient all-reduce over NVLink peer memory, fused with the optimiser ----------------------------------------------
// Every rank keeps its compact gradient rows in a buffer its peers can read (symmetric memory: the same allocation made on
// every GPU and mapped into every process).  One step of a fit then needs no collective library call:
//   k_peer_arrive_wait : rank r stores the step number into slot r of every peer's flag array (st.release.sys) and waits
//                        until its own array holds that number from every peer (ld.acquire.sys): all gradients are written
//   k_adam_rows_peer   : g = sum over ranks of peer_grad[rank][e], read straight from the peers' memory in rank order (so
//                        every rank forms the same sum bit for bit and the replicated parameters stay identical), then Adam
// The gradient buffers alternate between two allocations by step parity, so a rank that runs ahead writes the other buffer
// while a slow peer still reads this one; by the time a buffer is written again every peer has passed the arrive-wait of
// the step in between, i.e. finished reading it.  A wait that lasts longer than ~4 s sets err[0] = 1 and gives up (a dead
// peer must not hang the GPU).
constexpr int kMaxPeers = 16;
struct PeerPtrs {
    const float* grad[kMaxPeers];
    unsigned* flags[kMaxPeers];
};

__global__ void __launch_bounds__(32) k_peer_arrive_wait(PeerPtrs pp, int world, int rank, const int32_t* __restrict__ step_dev,
                                                         int32_t* __restrict__ err) {
    const int r = threadIdx.x;
    const unsigned seq = (unsigned)(step_dev[0] + 1);
    if (r < world) {
        __threadfence_system();   // this rank's gradient (written by the kernels before this one) is visible to the peers
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(pp.flags[r] + rank), "r"(seq) : "memory");
        const unsigned* mine = pp.flags[rank] + r;
        const long long t0 = clock64();
        unsigned v;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
            if ((int)(v - seq) >= 0) break;
            if (clock64() - t0 > 8000000000ll) { err[0] = 1; break; }
        }
    }
}

__global__ void __launch_bounds__(256) k_adam_rows_peer(float* __restrict__ params, const int32_t* __restrict__ row_ids, PeerPtrs pp,
                                                        int world, float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq,
                                                        int64_t nrows, int cpd, AdamArgs a, const int32_t* __restrict__ step_dev) {
    const int64_t tot = nrows * cpd;
    const int t = step_dev ? step_dev[0] + 1 : a.step;
    const double bc1 = 1.0 - pow((double)a.beta1, (double)t), bc2 = 1.0 - pow((double)a.beta2, (double)t);
    const float step_size = (float)((double)a.lr / bc1), bc2_sqrt = (float)sqrt(bc2);
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (int64_t)gridDim.x * blockDim.x) {
        float g = 0.0f;
        for (int r = 0; r < world; ++r) {   // rank order: the same sum on every rank
            float v;
            asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(pp.grad[r] + e) : "memory");   // never from a stale L1 line
            g += v;
        }
        g *= a.grad_scale;
        const int64_t row = e / cpd;
        const int k = (int)(e - row * cpd);
        const float m = a.beta1 * exp_avg[e] + (1.0f - a.beta1) * g;
        const float v = a.beta2 * exp_avg_sq[e] + (1.0f - a.beta2) * g * g;
        exp_avg[e] = m;
        exp_avg_sq[e] = v;
        const float denom = sqrtf(v) / bc2_sqrt + a.eps;
        float* p = params + (int64_t)__ldg(row_ids + row) * cpd + k;
        *p = *p - step_size * (m / denom);
    }
}

}  // namespace

extern "C" int64_t thb_mse_workspace_bytes(void) { return (int64_t)(1024 * sizeof(double) + 16); }

extern "C" int thb_mse_grad(const float* out, const float* target, int64_t nvalues, float scale, float* grad_out, float* loss,
                            void* workspace, void* stream) {
    THB_REQUIRE(nvalues >= 0, "thb_mse_grad: negative size");
    THB_REQUIRE(out && target && grad_out && loss && workspace, "thb_mse_grad: null pointer");
    THB_REQUIRE(((((uintptr_t)out | (uintptr_t)target | (uintptr_t)grad_out | (uintptr_t)workspace)) & 15) == 0,
                "thb_mse_grad: buffers must be 16-byte aligned");
    int64_t want = ((nvalues >> 2) + kMseThreads - 1) / kMseThreads;
    const int64_t cap = (int64_t)thb_num_sms() * 8 < 1024 ? (int64_t)thb_num_sms() * 8 : 1024;
    if (want < 1) want = 1;
    const int grid = (int)(want < cap ? want : cap);
    double* partial = (double*)workspace;
    unsigned* ticket = (unsigned*)((char*)workspace + 1024 * sizeof(double));  // zero before the first call
    k_mse_grad<<<grid, kMseThreads, 0, (cudaStream_t)stream>>>(out, target, nvalues, scale, grad_out, partial, ticket, loss);
    THB_LAUNCHED();
    return THB_OK;
}

extern "C" int thb_adam_rows(float* params, const int32_t* row_ids, const float* grad_rows, float* exp_avg, float* exp_avg_sq,
                             int64_t nrows, int cp_dim, float lr, float beta1, float beta2, float eps, float grad_scale, int step,
                             int32_t* step_dev, void* stream) {
    THB_REQUIRE(nrows >= 0 && cp_dim >= 1 && cp_dim <= 4, "thb_adam_rows: bad sizes");
    if (nrows == 0) return THB_OK;
    THB_REQUIRE(params && row_ids && grad_rows && exp_avg && exp_avg_sq, "thb_adam_rows: null pointer");
    THB_REQUIRE(step_dev || step >= 1, "thb_adam_rows: step numbers start at 1");
    AdamArgs a = {lr, beta1, beta2, eps, grad_scale, step};
    const int64_t tot = nrows * cp_dim;
    int64_t want = (tot + 255) / 256;
    const int64_t cap = (int64_t)thb_num_sms() * 8;
    const int grid = (int)(want < cap ? want : cap);
    cudaStream_t st = (cudaStream_t)stream;
    k_adam_rows<<<grid, 256, 0, st>>>(params, row_ids, grad_rows, exp_avg, exp_avg_sq, nrows, cp_dim, a, step_dev);
    THB_LAUNCHED();
    if (step_dev) {
        k_step_inc<<<1, 1, 0, st>>>(step_dev);
        THB_LAUNCHED();
    }
    return THB_OK;
}

extern "C" int thb_adam_rows_peer(float* params, const int32_t* row_ids, const void* const* peer_grads, void* const* peer_flags,
                                  int world, int rank, float* exp_avg, float* exp_avg_sq, int64_t nrows, int cp_dim, float lr,
                                  float beta1, float beta2, float eps, float grad_scale, int32_t* step_dev, int32_t* err,
                                  void* stream) {
    THB_REQUIRE(world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world, "thb_adam_rows_peer: bad world / rank");
    THB_REQUIRE(nrows >= 0 && cp_dim >= 1 && cp_dim <= 4, "thb_adam_rows_peer: bad sizes");
    THB_REQUIRE(params && row_ids && peer_grads && peer_flags && exp_avg && exp_avg_sq && step_dev && err, "thb_adam_rows_peer: null pointer");
    PeerPtrs pp = {};
    for (int r = 0; r < world; ++r) {
        THB_REQUIRE(peer_grads[r] && peer_flags[r], "thb_adam_rows_peer: null peer pointer");
        pp.grad[r] = (const float*)peer_grads[r];
        pp.flags[r] = (unsigned*)peer_flags[r];
    }
    cudaStream_t st = (cudaStream_t)stream;
    k_peer_arrive_wait<<<1, 32, 0, st>>>(pp, world, rank, step_dev, err);
    THB_LAUNCHED();
    if (nrows > 0) {
        AdamArgs a = {lr, beta1, beta2, eps, grad_scale, 0};
        const int64_t tot = nrows * cp_dim;
        const int64_t want = (tot + 255) / 256, cap = (int64_t)thb_num_sms() * 8;
        k_adam_rows_peer<<<(int)(want < cap ? want : cap), 256, 0, st>>>(params, row_ids, pp, world, exp_avg, exp_avg_sq, nrows, cp_dim, a, step_dev);
        THB_LAUNCHED();
    }
    k_step_inc<<<1, 1, 0, st>>>(step_dev);
    THB_LAUNCHED();
    return THB_OK;
}
