// Microbenchmark: FP32 FMA pipe throughput on B200 — plain FFMA vs FFMA2 (vector and scalar-broadcast forms).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void ffma2v(float2& d, float2 a, float2 b) {
    unsigned long long dd = *reinterpret_cast<unsigned long long*>(&d), aa = *reinterpret_cast<unsigned long long*>(&a), bb = *reinterpret_cast<unsigned long long*>(&b);
    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(aa), "l"(bb));
    d = *reinterpret_cast<float2*>(&dd);
}
__device__ __forceinline__ void ffma2s(float2& d, float a, float2 b) {
    unsigned long long dd = *reinterpret_cast<unsigned long long*>(&d), bb = *reinterpret_cast<unsigned long long*>(&b), aa;
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(aa), "l"(bb));
    d = *reinterpret_cast<float2*>(&dd);
}
template <int MODE> __global__ void k(float* out, int iters, float s) {
    float2 acc[16];
    for (int i = 0; i < 16; ++i) acc[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
    float2 b = make_float2(s, s * 0.5f);
    float a = s * 0.25f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == 0) { acc[i].x = fmaf(a, b.x, acc[i].x); acc[i].y = fmaf(a, b.y, acc[i].y); }
            if (MODE == 1) ffma2v(acc[i], b, b);
            if (MODE == 2) ffma2s(acc[i], a, b);
        }
    }
    float r = 0; for (int i = 0; i < 16; ++i) r += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE> void run(const char* name, float* out) {
    const int iters = 4096, blocks = 148 * 4, threads = 512;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(out, iters, 1.0001f);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, iters, 1.0001f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fma = (double)blocks * threads * iters * 32.0;
    printf("%-28s %.3f ms  %.1f TFLOP/s  (%.1f FMA/clk/SM at 1.965 GHz)\n", name, ms, 2 * fma / ms / 1e9, fma / (ms * 1e-3) / 148 / 1.965e9);
}
int main() {
    float* out; cudaMalloc(&out, 148 * 4 * 512 * 4);
    run<0>("FFMA (scalar)", out);
    run<1>("FFMA2 (vector operands)", out);
    run<2>("FFMA2 (scalar-broadcast A)", out);
    return 0;
}
