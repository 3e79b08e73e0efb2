// Stand-alone probe of inner-loop structures for the tile dot products (not part of the library).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o loop_probe loop_probe.cu && ./loop_probe
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long f32x2;
#define DEV __device__ __forceinline__
DEV f32x2 pack2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
DEV void ffma2(f32x2& d, f32x2 a, f32x2 b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b)); }
DEV float hsum2(f32x2 v) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); return lo + hi; }

// V0: FFMA2, PPT points x 1 support per step (the library's loop)
template <int PPT> __global__ void __launch_bounds__(32) v0(int iters, float* sink) {
    extern __shared__ __align__(16) float dyn[]; __shared__ __align__(16) float tiles[16 * 68];
    for (int i = threadIdx.x; i < 16 * 68; i += 32) tiles[i] = 1.0f / (1.0f + i);
    if (dyn == nullptr) tiles[0] = 0; __syncwarp();
    f32x2 tp2[PPT][32];
#pragma unroll
    for (int q = 0; q < PPT; ++q)
#pragma unroll
        for (int m = 0; m < 32; ++m) tp2[q][m] = pack2(threadIdx.x * 1e-3f + m, q + 0.5f * m);
    float acc[PPT] = {};
    for (int it = 0; it < iters; ++it) {
#pragma unroll 2
        for (int s = 0; s < 16; ++s) {
            const ulonglong2* row = (const ulonglong2*)(tiles + s * 68);
            f32x2 a[PPT][2];
#pragma unroll
            for (int q = 0; q < PPT; ++q) a[q][0] = a[q][1] = 0ull;
#pragma unroll
            for (int m = 0; m < 16; ++m) { const ulonglong2 tv = row[m];
#pragma unroll
                for (int q = 0; q < PPT; ++q) { ffma2(a[q][0], tp2[q][2 * m], tv.x); ffma2(a[q][1], tp2[q][2 * m + 1], tv.y); } }
#pragma unroll
            for (int q = 0; q < PPT; ++q) acc[q] += hsum2(a[q][0]) + hsum2(a[q][1]);
        }
    }
    float t = 0;
#pragma unroll
    for (int q = 0; q < PPT; ++q) t += acc[q];
    if (t == 12345.678f) sink[0] = t;
}
// V1: scalar FFMA, PPT points x 1 support
template <int PPT> __global__ void __launch_bounds__(32) v1(int iters, float* sink) {
    extern __shared__ __align__(16) float dyn[]; __shared__ __align__(16) float tiles[16 * 68];
    for (int i = threadIdx.x; i < 16 * 68; i += 32) tiles[i] = 1.0f / (1.0f + i);
    if (dyn == nullptr) tiles[0] = 0; __syncwarp();
    float tp[PPT][64];
#pragma unroll
    for (int q = 0; q < PPT; ++q)
#pragma unroll
        for (int m = 0; m < 64; ++m) tp[q][m] = threadIdx.x * 1e-3f + m + q;
    float acc[PPT] = {};
    for (int it = 0; it < iters; ++it) {
#pragma unroll 2
        for (int s = 0; s < 16; ++s) {
            const float4* row = (const float4*)(tiles + s * 68);
            float a[PPT][4];
#pragma unroll
            for (int q = 0; q < PPT; ++q) a[q][0] = a[q][1] = a[q][2] = a[q][3] = 0.f;
#pragma unroll
            for (int m = 0; m < 16; ++m) { const float4 tv = row[m];
#pragma unroll
                for (int q = 0; q < PPT; ++q) { a[q][0] = fmaf(tp[q][4 * m], tv.x, a[q][0]); a[q][1] = fmaf(tp[q][4 * m + 1], tv.y, a[q][1]);
                                                 a[q][2] = fmaf(tp[q][4 * m + 2], tv.z, a[q][2]); a[q][3] = fmaf(tp[q][4 * m + 3], tv.w, a[q][3]); } }
#pragma unroll
            for (int q = 0; q < PPT; ++q) acc[q] += (a[q][0] + a[q][1]) + (a[q][2] + a[q][3]);
        }
    }
    float t = 0;
#pragma unroll
    for (int q = 0; q < PPT; ++q) t += acc[q];
    if (t == 12345.678f) sink[0] = t;
}
// V2: scalar FFMA, 2 points x 2 supports block
__global__ void __launch_bounds__(32) v2(int iters, float* sink) {
    extern __shared__ __align__(16) float dyn[]; __shared__ __align__(16) float tiles[16 * 68];
    for (int i = threadIdx.x; i < 16 * 68; i += 32) tiles[i] = 1.0f / (1.0f + i);
    if (dyn == nullptr) tiles[0] = 0; __syncwarp();
    float tp[2][64];
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int m = 0; m < 64; ++m) tp[q][m] = threadIdx.x * 1e-3f + m + q;
    float acc[2] = {};
    for (int it = 0; it < iters; ++it) {
        for (int s = 0; s < 16; s += 2) {
            const float4* r0 = (const float4*)(tiles + s * 68); const float4* r1 = (const float4*)(tiles + (s + 1) * 68);
            float a00[2] = {0, 0}, a01[2] = {0, 0}, a10[2] = {0, 0}, a11[2] = {0, 0};
#pragma unroll
            for (int m = 0; m < 16; ++m) { const float4 t0 = r0[m], t1 = r1[m];
                const float* x0 = &t0.x; const float* x1 = &t1.x;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    a00[e & 1] = fmaf(tp[0][4 * m + e], x0[e], a00[e & 1]); a01[e & 1] = fmaf(tp[0][4 * m + e], x1[e], a01[e & 1]);
                    a11[e & 1] = fmaf(tp[1][4 * m + e], x1[e], a11[e & 1]); a10[e & 1] = fmaf(tp[1][4 * m + e], x0[e], a10[e & 1]); } }
            acc[0] += a00[0] + a00[1] + a01[0] + a01[1]; acc[1] += a10[0] + a10[1] + a11[0] + a11[1];
        }
    }
    if (acc[0] + acc[1] == 12345.678f) sink[0] = acc[0];
}
// V3: FFMA2, 2 points x 2 supports block
__global__ void __launch_bounds__(32) v3(int iters, float* sink) {
    extern __shared__ __align__(16) float dyn[]; __shared__ __align__(16) float tiles[16 * 68];
    for (int i = threadIdx.x; i < 16 * 68; i += 32) tiles[i] = 1.0f / (1.0f + i);
    if (dyn == nullptr) tiles[0] = 0; __syncwarp();
    f32x2 tp2[2][32];
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int m = 0; m < 32; ++m) tp2[q][m] = pack2(threadIdx.x * 1e-3f + m, q + 0.5f * m);
    float acc[2] = {};
    for (int it = 0; it < iters; ++it) {
        for (int s = 0; s < 16; s += 2) {
            const ulonglong2* r0 = (const ulonglong2*)(tiles + s * 68); const ulonglong2* r1 = (const ulonglong2*)(tiles + (s + 1) * 68);
            f32x2 a00[2] = {0, 0}, a01[2] = {0, 0}, a10[2] = {0, 0}, a11[2] = {0, 0};
#pragma unroll
            for (int m = 0; m < 16; ++m) { const ulonglong2 t0 = r0[m], t1 = r1[m];
                ffma2(a00[0], tp2[0][2 * m], t0.x); ffma2(a01[0], tp2[0][2 * m], t1.x); ffma2(a11[0], tp2[1][2 * m], t1.x); ffma2(a10[0], tp2[1][2 * m], t0.x);
                ffma2(a10[1], tp2[1][2 * m + 1], t0.y); ffma2(a11[1], tp2[1][2 * m + 1], t1.y); ffma2(a01[1], tp2[0][2 * m + 1], t1.y); ffma2(a00[1], tp2[0][2 * m + 1], t0.y); }
            acc[0] += hsum2(a00[0]) + hsum2(a00[1]) + hsum2(a01[0]) + hsum2(a01[1]); acc[1] += hsum2(a10[0]) + hsum2(a10[1]) + hsum2(a11[0]) + hsum2(a11[1]);
        }
    }
    if (acc[0] + acc[1] == 12345.678f) sink[0] = acc[0];
}
// V4: FFMA2 with NO shared memory (tile pairs live in registers too): pure register-operand FFMA2 rate
__global__ void __launch_bounds__(32) v4(int iters, float* sink) {
    f32x2 tp2[32], tl[16];
#pragma unroll
    for (int m = 0; m < 32; ++m) tp2[m] = pack2(threadIdx.x * 1e-3f + m, 0.5f * m);
#pragma unroll
    for (int m = 0; m < 16; ++m) tl[m] = pack2(1.0f / (1 + m), 0.25f * m);
    f32x2 a[8] = {};
    for (int it = 0; it < iters * 16; ++it) {
#pragma unroll
        for (int m = 0; m < 32; ++m) ffma2(a[m & 7], tp2[m], tl[m & 15]);
    }
    float t = 0;
#pragma unroll
    for (int m = 0; m < 8; ++m) t += hsum2(a[m]);
    if (t == 12345.678f) sink[0] = t;
}
// V5: FFMA2 with a SCALAR multiplier shared by 16 consecutive instructions (register operands only)
__global__ void __launch_bounds__(32) v5(int iters, float* sink) {
    f32x2 x[16], acc[16]; float sc[4];
#pragma unroll
    for (int i = 0; i < 16; ++i) { x[i] = pack2(threadIdx.x * 1e-3f + i, 0.5f * i); acc[i] = 0ull; }
#pragma unroll
    for (int j = 0; j < 4; ++j) sc[j] = 1.0f + 1e-3f * j + threadIdx.x * 1e-6f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { const f32x2 s2 = pack2(sc[j], sc[j]);
#pragma unroll
            for (int i = 0; i < 16; ++i) ffma2(acc[i], s2, x[i]); }
    }
    float t = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) t += hsum2(acc[i]);
    if (t == 12345.678f) sink[0] = t;
}
// V6: the first stage of the net evaluation: per b four broadcast LDS.128 (W[a][b][0..3]), 2 points, e[q][a][h] += B[q][b] * W
__global__ void __launch_bounds__(32) v6(int iters, float* sink) {
    extern __shared__ __align__(16) float dyn[]; __shared__ __align__(16) float net[64];
    for (int i = threadIdx.x; i < 64; i += 32) net[i] = 1.0f / (1.0f + i);
    if (dyn == nullptr) net[0] = 0; __syncwarp();
    float fb[2][4];
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int b = 0; b < 4; ++b) fb[q][b] = 1.0f + 1e-3f * b + q + threadIdx.x * 1e-6f;
    f32x2 e[2][4][2];
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int a = 0; a < 4; ++a) e[q][a][0] = e[q][a][1] = 0ull;
    for (int it = 0; it < iters; ++it) {
        const ulonglong2* col = (const ulonglong2*)net;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            ulonglong2 cv[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) cv[a] = col[a * 4 + b];
#pragma unroll
            for (int q = 0; q < 2; ++q) { const f32x2 bb = pack2(fb[q][b], fb[q][b]);
#pragma unroll
                for (int a = 0; a < 4; ++a) { ffma2(e[q][a][0], bb, cv[a].x); ffma2(e[q][a][1], bb, cv[a].y); } }
        }
    }
    float t = 0;
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int a = 0; a < 4; ++a) t += hsum2(e[q][a][0]) + hsum2(e[q][a][1]);
    if (t == 12345.678f) sink[0] = t;
}
template <typename K> void run(const char* name, K k, int warps, int iters, double flop_per_thread_iter, int pad_extra = 0) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int pad = (200 * 1024) / warps - 5 * 1024; cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, pad);
    float* sink; cudaMalloc(&sink, 16);
    int grid = sms * warps; cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<<<grid, 32, pad>>>(iters, sink); cudaDeviceSynchronize();
    float best = 1e9;
    for (int r = 0; r < 3; ++r) { cudaEventRecord(a); k<<<grid, 32, pad>>>(iters, sink); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
    printf("%-44s warps/SM %2d : %6.1f TFLOP/s  (%s)\n", name, warps, flop_per_thread_iter * iters * 32.0 * grid / (best * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    const int it = 2000; const double f1 = 2.0 * 64 * 16;  // flops per point per iter
    run("V0 FFMA2  2 pts x 1 supp", v0<2>, 8, it, f1 * 2); run("V0 FFMA2  1 pt  x 1 supp", v0<1>, 16, it, f1);
    run("V1 FFMA   2 pts x 1 supp", v1<2>, 8, it, f1 * 2); run("V1 FFMA   1 pt  x 1 supp", v1<1>, 16, it, f1);
    run("V1 FFMA   3 pts x 1 supp", v1<3>, 8, it, f1 * 3); run("V1 FFMA   3 pts x 1 supp", v1<3>, 6, it, f1 * 3); run("V1 FFMA   2 pts x 1 supp", v1<2>, 10, it, f1 * 2); run("V1 FFMA   2 pts x 1 supp", v1<2>, 6, it, f1 * 2);
    run("V2 FFMA   2 pts x 2 supp block", v2, 8, it, f1 * 2); run("V3 FFMA2  2 pts x 2 supp block", v3, 8, it, f1 * 2);
    run("V4 FFMA2  register operands only", v4, 8, it, 2.0 * 2 * 32 * 16); run("V4 FFMA2  register operands only", v4, 16, it, 2.0 * 2 * 32 * 16);
    run("V5 FFMA2  scalar multiplier x16, registers", v5, 8, it, 2.0 * 2 * 64); run("V5 FFMA2  scalar multiplier x16, registers", v5, 16, it, 2.0 * 2 * 64);
    run("V6 FFMA2  net stage 1 (LDS.128 + scalar)", v6, 8, it, 2.0 * 2 * 64); run("V6 FFMA2  net stage 1 (LDS.128 + scalar)", v6, 16, it, 2.0 * 2 * 64);
    return 0;
}
