// Node attention: softmax_j(q_n . k_j * dh^-1/2) v_j over the nodes of one sample, one head.
// Reference: Attention.forward, src/core/network/layers/attention.py:125-135.
// One warp per (sample, head); lane = query node; K/V of the head live in shared memory (fp32);
// the qkv / out tensors are fp32 (parity path) or bf16 (tensor-core path), math is always fp32.
#include "sd_internal.h"
#include <math.h>

namespace sd {

template <typename T> struct Io;
template <> struct Io<float> {
    static __device__ __forceinline__ void load4(const float* p, float (&v)[4]) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p)); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    static __device__ __forceinline__ void store4(float* p, const float (&v)[4]) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};
template <> struct Io<__nv_bfloat16> {
    static __device__ __forceinline__ void load4(const __nv_bfloat16* p, float (&v)[4]) {
        const uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
        v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xFFFF0000u);
        v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xFFFF0000u);
    }
    static __device__ __forceinline__ void store4(__nv_bfloat16* p, const float (&v)[4]) {
        __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
        uint2 t; t.x = *reinterpret_cast<uint32_t*>(&a); t.y = *reinterpret_cast<uint32_t*>(&b);
        *reinterpret_cast<uint2*>(p) = t;
    }
};

template <typename T, int DH, int NMAX>
__global__ void __launch_bounds__(128)
node_attention_kernel(const T* __restrict__ qkv, T* __restrict__ out, int B, int N, int H) {
    extern __shared__ __align__(16) float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long wid = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    if (wid >= (long long)B * H) return;
    const int b = (int)(wid / H), h = (int)(wid % H);
    float* Ks = smem + (size_t)warp * 2 * NMAX * DH;
    float* Vs = Ks + NMAX * DH;
    const int HD = H * DH;
    const long long row_stride = 3LL * HD;
    const T* base = qkv + (long long)b * N * row_stride + h * DH;
    for (int i = lane; i < N * (DH / 4); i += 32) {
        const int j = i / (DH / 4), c4 = i % (DH / 4);
        const T* r = base + j * row_stride + 4 * c4;
        float kv[4], vv[4];
        Io<T>::load4(r + HD, kv);
        Io<T>::load4(r + 2 * HD, vv);
        *reinterpret_cast<float4*>(Ks + j * DH + 4 * c4) = make_float4(kv[0], kv[1], kv[2], kv[3]);
        *reinterpret_cast<float4*>(Vs + j * DH + 4 * c4) = make_float4(vv[0], vv[1], vv[2], vv[3]);
    }
    __syncwarp();
    const float scale = rsqrtf((float)DH);
    for (int n = lane; n < N; n += 32) {
        float q[DH];
        const T* qr = base + n * row_stride;
#pragma unroll
        for (int c = 0; c < DH; c += 4) {
            float t[4];
            Io<T>::load4(qr + c, t);
            q[c] = t[0] * scale; q[c + 1] = t[1] * scale; q[c + 2] = t[2] * scale; q[c + 3] = t[3] * scale;
        }
        float sc[NMAX];
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < NMAX; ++j) {
            if (j < N) {
                float s = 0.0f;
#pragma unroll
                for (int c = 0; c < DH; c += 4) {
                    const float4 kv = *reinterpret_cast<const float4*>(Ks + j * DH + c);
                    s = fmaf(q[c], kv.x, s); s = fmaf(q[c + 1], kv.y, s);
                    s = fmaf(q[c + 2], kv.z, s); s = fmaf(q[c + 3], kv.w, s);
                }
                sc[j] = s;
                mx = fmaxf(mx, s);
            }
        }
        float acc[DH];
#pragma unroll
        for (int c = 0; c < DH; ++c) acc[c] = 0.0f;
        float sum = 0.0f;
#pragma unroll
        for (int j = 0; j < NMAX; ++j) {
            if (j < N) {
                const float pj = expf(sc[j] - mx);
                sum += pj;
#pragma unroll
                for (int c = 0; c < DH; c += 4) {
                    const float4 vv = *reinterpret_cast<const float4*>(Vs + j * DH + c);
                    acc[c] = fmaf(pj, vv.x, acc[c]); acc[c + 1] = fmaf(pj, vv.y, acc[c + 1]);
                    acc[c + 2] = fmaf(pj, vv.z, acc[c + 2]); acc[c + 3] = fmaf(pj, vv.w, acc[c + 3]);
                }
            }
        }
        const float inv = 1.0f / sum;
        T* o = out + ((long long)b * N + n) * HD + h * DH;
#pragma unroll
        for (int c = 0; c < DH; c += 4) {
            const float t[4] = {acc[c] * inv, acc[c + 1] * inv, acc[c + 2] * inv, acc[c + 3] * inv};
            Io<T>::store4(o + c, t);
        }
    }
}

template <typename T, int DH, int NMAX>
static int launch_attention(const T* qkv, T* out, int B, int N, int H, cudaStream_t st) {
    const int warps = 4;
    const size_t smem = (size_t)warps * 2 * NMAX * DH * sizeof(float);
    auto kern = node_attention_kernel<T, DH, NMAX>;
    if (smem > 48 * 1024) SD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long tasks = (long long)B * H;
    kern<<<(unsigned)((tasks + warps - 1) / warps), warps * 32, smem, st>>>(qkv, out, B, N, H);
    SD_LAUNCH_OK("node_attention_kernel");
    return SD_OK;
}

template <typename T>
static int node_attention_any(const T* qkv, T* out, int B, int N, int heads, int dh, cudaStream_t st) {
    if (B <= 0) return SD_OK;
    if (dh == 32 && N <= 32) return launch_attention<T, 32, 32>(qkv, out, B, N, heads, st);
    if (dh == 32 && N <= 64) return launch_attention<T, 32, 64>(qkv, out, B, N, heads, st);
    if (dh == 16 && N <= 32) return launch_attention<T, 16, 32>(qkv, out, B, N, heads, st);
    if (dh == 64 && N <= 32) return launch_attention<T, 64, 32>(qkv, out, B, N, heads, st);
    set_error("node_attention: dim_head %d with %d nodes unsupported (dim_head 32: N<=64; 16/64: N<=32)", dh, N);
    return SD_ERR_UNSUPPORTED;
}

int node_attention_fp32(const float* qkv, float* out, int B, int N, int heads, int dh, cudaStream_t st) {
    return node_attention_any<float>(qkv, out, B, N, heads, dh, st);
}
int node_attention_bf16(const __nv_bfloat16* qkv, __nv_bfloat16* out, int B, int N, int heads, int dh, cudaStream_t st) {
    return node_attention_any<__nv_bfloat16>(qkv, out, B, N, heads, dh, st);
}

}  // namespace sd
