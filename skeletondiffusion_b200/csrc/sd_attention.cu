// Node attention: softmax_j(q_n . k_j * dh^-1/2) v_j over the nodes of one sample, one head.
// Reference: Attention.forward, src/core/network/layers/attention.py:125-135.
//
// One warp per (sample, head) task, several tasks per warp (grid-stride), lane = query node.  The head's Q/K/V rows
// ([N, dh] each) are copied into shared memory with coalesced 16-byte loads (rows padded to dh + 4 floats so that a
// lane reading ITS OWN q row is bank-conflict free; K/V rows are read as warp-wide broadcasts).  All products run
// on the packed FFMA2 pipe: q.k as float2 partial sums over channel pairs, p_j * v_j as scalar x float2.
// qkv / out are fp32 (parity path) or bf16 (tensor-core path); the math is fp32 in both.
#include "sd_internal.h"
#include <math.h>

namespace sd {

template <typename T> struct Io;
template <> struct Io<float> {
    static __device__ __forceinline__ float4 load4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
    static __device__ __forceinline__ void store4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
};
template <> struct Io<__nv_bfloat16> {
    static __device__ __forceinline__ float4 load4(const __nv_bfloat16* p) {
        const uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
        return make_float4(__uint_as_float(t.x << 16), __uint_as_float(t.x & 0xFFFF0000u), __uint_as_float(t.y << 16), __uint_as_float(t.y & 0xFFFF0000u));
    }
    static __device__ __forceinline__ void store4(__nv_bfloat16* p, float4 v) {
        __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
        uint2 t; t.x = *reinterpret_cast<uint32_t*>(&a); t.y = *reinterpret_cast<uint32_t*>(&b);
        *reinterpret_cast<uint2*>(p) = t;
    }
};

__device__ __forceinline__ void att_ffma2(float2& d, float2 a, float2 b) {
    unsigned long long dd = *reinterpret_cast<unsigned long long*>(&d), aa = *reinterpret_cast<unsigned long long*>(&a), bb = *reinterpret_cast<unsigned long long*>(&b);
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(aa), "l"(bb));
    d = *reinterpret_cast<float2*>(&dd);
}
__device__ __forceinline__ void att_ffma2s(float2& d, float a, float2 b) {
    unsigned long long dd = *reinterpret_cast<unsigned long long*>(&d), bb = *reinterpret_cast<unsigned long long*>(&b), aa;
    asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(aa), "l"(bb));
    d = *reinterpret_cast<float2*>(&dd);
}

constexpr int ATT_WARPS = 4;

// NEXACT > 0: the node count is a compile-time constant (all index arithmetic of the tile copy folds away and the
// score loops carry no guards); NEXACT == 0: any N <= NMAX at run time.
template <typename T, int DH, int NMAX, int NEXACT>
__global__ void __launch_bounds__(ATT_WARPS * 32)
node_attention_kernel(const T* __restrict__ qkv, T* __restrict__ out, long long tasks, int N_rt, int H) {
    const int N = NEXACT > 0 ? NEXACT : N_rt;
    constexpr int LD = DH + 4;                       // padded row (floats)
    extern __shared__ __align__(16) float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* Qs = smem + (size_t)warp * 3 * NMAX * LD;
    float* Ks = Qs + NMAX * LD;
    float* Vs = Ks + NMAX * LD;
    const int HD = H * DH;
    const long long row_stride = 3LL * HD;
    const float scale = rsqrtf((float)DH);
    for (long long task = (long long)blockIdx.x * ATT_WARPS + warp; task < tasks; task += (long long)gridDim.x * ATT_WARPS) {
        const long long b = task / H;
        const int h = (int)(task % H);
        const T* base = qkv + b * N * row_stride + h * DH;
        __syncwarp();                                // previous task's readers are done with the tiles
        // tile copy in batches of 8 loads per lane: all 8 are in flight before the first shared-memory store
        // (a load -> store loop issues in order and pays one DRAM latency per iteration)
        const int per_mat = N * (DH / 4), total4 = 3 * per_mat;
        for (int i0 = lane; i0 < total4; i0 += 32 * 8) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = i0 + 32 * u;
                if (i < total4) {
                    const int which = i / per_mat, rem = i - which * per_mat;
                    const int j = rem / (DH / 4), c4 = rem % (DH / 4);
                    v[u] = Io<T>::load4(base + j * row_stride + which * HD + 4 * c4);
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = i0 + 32 * u;
                if (i < total4) {
                    const int which = i / per_mat, rem = i - which * per_mat;
                    const int j = rem / (DH / 4), c4 = rem % (DH / 4);
                    float4 t = v[u];
                    if (which == 0) { t.x *= scale; t.y *= scale; t.z *= scale; t.w *= scale; }      // q * dh^-1/2 (attention.py:128)
                    *reinterpret_cast<float4*>(Qs + which * NMAX * LD + j * LD + 4 * c4) = t;
                }
            }
        }
        __syncwarp();
        for (int n = lane; n < N; n += 32) {
            float2 q[DH / 2];
#pragma unroll
            for (int c = 0; c < DH; c += 4) {
                const float4 t = *reinterpret_cast<const float4*>(Qs + n * LD + c);
                q[c / 2] = make_float2(t.x, t.y); q[c / 2 + 1] = make_float2(t.z, t.w);
            }
            float sc[NMAX];
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < NMAX; ++j) {
                if (j < N) {
                    float2 s2[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
                    for (int c = 0; c < DH; c += 4) {       // four independent accumulator chains
                        const float4 kv = *reinterpret_cast<const float4*>(Ks + j * LD + c);
                        att_ffma2(s2[(c / 4) & 1], q[c / 2], make_float2(kv.x, kv.y));
                        att_ffma2(s2[2 + ((c / 4) & 1)], q[c / 2 + 1], make_float2(kv.z, kv.w));
                    }
                    sc[j] = ((s2[0].x + s2[1].x) + (s2[2].x + s2[3].x)) + ((s2[0].y + s2[1].y) + (s2[2].y + s2[3].y));
                    mx = fmaxf(mx, sc[j]);
                }
            }
            float2 acc[DH / 2];
#pragma unroll
            for (int c = 0; c < DH / 2; ++c) acc[c] = make_float2(0.f, 0.f);
            float sum = 0.0f;
#pragma unroll
            for (int j = 0; j < NMAX; ++j) {
                if (j < N) {
                    const float pj = expf(sc[j] - mx);
                    sum += pj;
#pragma unroll
                    for (int c = 0; c < DH; c += 4) {
                        const float4 vv = *reinterpret_cast<const float4*>(Vs + j * LD + c);
                        att_ffma2s(acc[c / 2], pj, make_float2(vv.x, vv.y));
                        att_ffma2s(acc[c / 2 + 1], pj, make_float2(vv.z, vv.w));
                    }
                }
            }
            const float inv = 1.0f / sum;
            T* o = out + (b * N + n) * HD + h * DH;
#pragma unroll
            for (int c = 0; c < DH; c += 4)
                Io<T>::store4(o + c, make_float4(acc[c / 2].x * inv, acc[c / 2].y * inv, acc[c / 2 + 1].x * inv, acc[c / 2 + 1].y * inv));
        }
    }
}

template <typename T, int DH, int NMAX, int NEXACT>
static int launch_attention(const T* qkv, T* out, int B, int N, int H, cudaStream_t st) {
    const size_t smem = (size_t)ATT_WARPS * 3 * NMAX * (DH + 4) * sizeof(float);
    auto kern = node_attention_kernel<T, DH, NMAX, NEXACT>;
    static bool configured = false;
    if (!configured && smem > 48 * 1024) {
        SD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    const long long tasks = (long long)B * H;
    long long blocks = (tasks + ATT_WARPS - 1) / ATT_WARPS;
    const long long cap = 148LL * 8;                 // persistent-style grid: a few CTAs per SM, tasks grid-strided
    if (blocks > cap) blocks = cap;
    kern<<<(unsigned)blocks, ATT_WARPS * 32, smem, st>>>(qkv, out, tasks, N, H);
    SD_LAUNCH_OK("node_attention_kernel");
    return SD_OK;
}

template <typename T>
static int node_attention_any(const T* qkv, T* out, int B, int N, int heads, int dh, cudaStream_t st) {
    if (B <= 0) return SD_OK;
    if (dh == 32 && N == 21) return launch_attention<T, 32, 21, 21>(qkv, out, B, N, heads, st);    // AMASS
    if (dh == 32 && N == 16) return launch_attention<T, 32, 16, 16>(qkv, out, B, N, heads, st);    // H36M, README
    if (dh == 32 && N == 17) return launch_attention<T, 32, 17, 17>(qkv, out, B, N, heads, st);    // FreeMan
    if (dh == 32 && N <= 32) return launch_attention<T, 32, 32, 0>(qkv, out, B, N, heads, st);
    if (dh == 32 && N <= 64) return launch_attention<T, 32, 64, 0>(qkv, out, B, N, heads, st);
    if (dh == 16 && N <= 32) return launch_attention<T, 16, 32, 0>(qkv, out, B, N, heads, st);
    if (dh == 64 && N <= 32) return launch_attention<T, 64, 32, 0>(qkv, out, B, N, heads, st);
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
