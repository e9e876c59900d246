// Node attention kernel (fp32).
#include "sd_internal.h"
#include <math.h>

namespace sd {

// =============================================================================================
// node attention: softmax_j(q_n . k_j * dh^-1/2) v_j over the nodes of one sample, one head
// Reference: Attention.forward, src/core/network/layers/attention.py:125-135.
// One warp per (sample, head); lane = query node; K/V of the head live in shared memory.
// =============================================================================================
template <int DH, int NMAX>
__global__ void __launch_bounds__(128)
node_attention_kernel(const float* __restrict__ qkv, float* __restrict__ out, int B, int N, int H) {
    extern __shared__ __align__(16) float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long wid = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    if (wid >= (long long)B * H) return;
    const int b = (int)(wid / H), h = (int)(wid % H);
    float* Ks = smem + (size_t)warp * 2 * NMAX * DH;
    float* Vs = Ks + NMAX * DH;
    const int HD = H * DH;
    const long long row_stride = 3LL * HD;
    const float* base = qkv + (long long)b * N * row_stride + h * DH;
    for (int i = lane; i < N * (DH / 4); i += 32) {
        const int j = i / (DH / 4), c4 = i % (DH / 4);
        const float* r = base + j * row_stride + 4 * c4;
        *reinterpret_cast<float4*>(Ks + j * DH + 4 * c4) = __ldg(reinterpret_cast<const float4*>(r + HD));
        *reinterpret_cast<float4*>(Vs + j * DH + 4 * c4) = __ldg(reinterpret_cast<const float4*>(r + 2 * HD));
    }
    __syncwarp();
    const float scale = rsqrtf((float)DH);
    for (int n = lane; n < N; n += 32) {
        float q[DH];
        const float* qr = base + n * row_stride;
#pragma unroll
        for (int c = 0; c < DH; c += 4) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(qr + c));
            q[c] = t.x * scale; q[c + 1] = t.y * scale; q[c + 2] = t.z * scale; q[c + 3] = t.w * scale;
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
        float* o = out + ((long long)b * N + n) * HD + h * DH;
#pragma unroll
        for (int c = 0; c < DH; c += 4)
            *reinterpret_cast<float4*>(o + c) = make_float4(acc[c] * inv, acc[c + 1] * inv, acc[c + 2] * inv, acc[c + 3] * inv);
    }
}

template <int DH, int NMAX>
static int launch_attention(const float* qkv, float* out, int B, int N, int H, cudaStream_t st) {
    const int warps = 4;
    const size_t smem = (size_t)warps * 2 * NMAX * DH * sizeof(float);
    auto kern = node_attention_kernel<DH, NMAX>;
    if (smem > 48 * 1024) SD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long tasks = (long long)B * H;
    kern<<<(unsigned)((tasks + warps - 1) / warps), warps * 32, smem, st>>>(qkv, out, B, N, H);
    SD_LAUNCH_OK("node_attention_kernel");
    return SD_OK;
}

int node_attention_fp32(const float* qkv, float* out, int B, int N, int heads, int dh, cudaStream_t st) {
    if (B <= 0) return SD_OK;
    if (N > 64) { set_error("node_attention: num_nodes %d > 64", N); return SD_ERR_UNSUPPORTED; }
    if (dh == 32) return N <= 32 ? launch_attention<32, 32>(qkv, out, B, N, heads, st) : launch_attention<32, 64>(qkv, out, B, N, heads, st);
    if (dh == 16 && N <= 32) return launch_attention<16, 32>(qkv, out, B, N, heads, st);
    if (dh == 64 && N <= 32) return launch_attention<64, 32>(qkv, out, B, N, heads, st);
    set_error("node_attention: dim_head %d with %d nodes unsupported (dim_head 32: N<=64; 16/64: N<=32)", dh, N);
    return SD_ERR_UNSUPPORTED;
}

}  // namespace sd
