// Error handling, bookkeeping and the FP32 FMA micro-benchmark of libthb_b200.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "thb_common.cuh"

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

int thb_set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
void thb_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int thb_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= THB_MAX_DEVICES) dev = 0;
    return dev;
}

// cached per device ordinal: a process may drive several GPUs (kernel attributes and grids are per device too)
int thb_num_sms() {
    static int sms[THB_MAX_DEVICES] = {};
    const int dev = thb_device();
    if (sms[dev] == 0) {
        int v = 0;
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        sms[dev] = v > 0 ? v : 148;
    }
    return sms[dev];
}

extern "C" int thb_version(void) { return 100; }
extern "C" const char* thb_last_error(void) { return g_err; }
extern "C" int64_t thb_launch_count(void) { return g_launches.load(); }

// ---- FP32 peak probe ------------------------------------------------------------------------------------
// 16 independent accumulator chains per thread, 8 warps per CTA, enough CTAs to fill every SM: what the
// FMA pipe can do when nothing else competes for issue slots.  variant 0: FFMA, variant 1: FFMA2.
template <int VARIANT>
__global__ void __launch_bounds__(256) k_fma_peak(int iters, float* sink) {
    float a = 1.0f + threadIdx.x * 1e-7f, b = 0.999f;
    if (VARIANT == 0) {
        float acc[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = i;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], b, a);
        }
        float s = 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) s += acc[i];
        if (s == 12345.678f) sink[0] = s;
    } else {
        f32x2 acc[16];
        f32x2 a2 = pack2(a, a * 0.5f), b2 = pack2(b, b);
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = pack2(float(i), float(-i));
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                // acc = acc * b2 + a2  (written as d = a*b + d with the roles swapped every other step)
                asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i]) : "l"(b2), "l"(a2));
            }
        }
        float s = 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) s += hsum2(acc[i]);
        if (s == 12345.678f) sink[0] = s;
    }
}

// Synthetic replica of the tile loop of thb_tile_impl.cuh: PPT points per lane whose 64-entry tensor product
// sits in registers as 32 f32x2 pairs, 16 tile rows in shared memory read with broadcast LDS.128, FFMA2.
// One warp per CTA; `smem_pad` bytes of dynamic shared memory cap the number of resident warps per SM.
// It measures what this loop STRUCTURE can reach on the FMA pipe, without any of the surrounding work.
template <int PPT>
__global__ void __launch_bounds__(32) k_tile_loop_probe(int iters, float* sink) {
    extern __shared__ __align__(16) float dyn[];
    __shared__ __align__(16) float tiles[16 * 68];
    for (int i = threadIdx.x; i < 16 * 68; i += 32) tiles[i] = 1.0f / (1.0f + i);
    if (dyn == nullptr) tiles[0] = 0.f;
    __syncwarp();
    f32x2 tp2[PPT][32];
#pragma unroll
    for (int q = 0; q < PPT; ++q)
#pragma unroll
        for (int m = 0; m < 32; ++m) tp2[q][m] = pack2(threadIdx.x * 1e-3f + m, q + 0.5f * m);
    float acc[PPT] = {};
    for (int it = 0; it < iters; ++it) {
#pragma unroll 2
        for (int s = 0; s < 16; ++s) {
            const ulonglong2* row = reinterpret_cast<const ulonglong2*>(tiles + s * 68);
            f32x2 a[PPT][2];
#pragma unroll
            for (int q = 0; q < PPT; ++q) a[q][0] = a[q][1] = 0ull;
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                const ulonglong2 tv = row[m];
#pragma unroll
                for (int q = 0; q < PPT; ++q) {
                    ffma2(a[q][0], tp2[q][2 * m], tv.x);
                    ffma2(a[q][1], tp2[q][2 * m + 1], tv.y);
                }
            }
#pragma unroll
            for (int q = 0; q < PPT; ++q) acc[q] += hsum2(a[q][0]) + hsum2(a[q][1]);
        }
    }
    float t = 0;
#pragma unroll
    for (int q = 0; q < PPT; ++q) t += acc[q];
    if (t == 12345.678f) sink[0] = t;
}

extern "C" int thb_fma_peak(int variant, int iters, float* sink, double* flops_out_host, void* stream) {
    THB_REQUIRE(sink && iters > 0, "thb_fma_peak: bad arguments");
    const int ctas = thb_num_sms() * 8;
    cudaStream_t st = (cudaStream_t)stream;
    if (variant >= 2) {
        // 2: PPT=2 / 8 warps per SM, 3: PPT=1 / 8 warps, 4: PPT=1 / 16 warps, 5: PPT=2 / 4 warps
        const int ppt = (variant == 2 || variant == 5) ? 2 : 1;
        const int warps = variant == 4 ? 16 : (variant == 5 ? 4 : 8);
        const int pad = (200 * 1024) / warps - 5 * 1024;  // dynamic smem so that exactly `warps` CTAs fit per SM
        const int grid = thb_num_sms() * warps;
        if (ppt == 2) {
            cudaFuncSetAttribute(k_tile_loop_probe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, pad);
            k_tile_loop_probe<2><<<grid, 32, pad, st>>>(iters, sink);
        } else {
            cudaFuncSetAttribute(k_tile_loop_probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, pad);
            k_tile_loop_probe<1><<<grid, 32, pad, st>>>(iters, sink);
        }
        if (flops_out_host) *flops_out_host = 2.0 * 2 * 32 * 16 * ppt * double(iters) * 32.0 * grid;
        THB_LAUNCHED();
        return THB_OK;
    }
    if (variant == 0) {
        k_fma_peak<0><<<ctas, 256, 0, st>>>(iters, sink);
        if (flops_out_host) *flops_out_host = 2.0 * 16 * double(iters) * 256.0 * ctas;
    } else {
        k_fma_peak<1><<<ctas, 256, 0, st>>>(iters, sink);
        if (flops_out_host) *flops_out_host = 2.0 * 2 * 16 * double(iters) * 256.0 * ctas;
    }
    THB_LAUNCHED();
    return THB_OK;
}
