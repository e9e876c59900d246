// Graph-influence matrix as a BY-VALUE kernel parameter (constant bank) and the fully unrolled N x N mix that reads it.
// Stored input-major in float4 groups, g4[m][q] = (G^[4q][m], G^[4q+1][m], G^[4q+2][m], G^[4q+3][m]): for one input node m
// the coefficients of four consecutive output nodes are one 16-byte constant load (LDCU.128 on sm_100a, where FFMA takes
// its constant operand through a uniform register): 105 + 21 uniform loads per 441 FFMAs at N = 21 instead of one each.
#pragma once
#include <cuda_runtime.h>

namespace sd {

template <int N> struct MixMat {
    static constexpr int Q = (N + 3) / 4;
    float4 g4[N][Q];
    // G_host: row-major G^[n][m] or null (identity)
    void set(const float* G_host) {
        for (int m = 0; m < N; ++m)
            for (int q = 0; q < Q; ++q) {
                float v[4];
                for (int j = 0; j < 4; ++j) {
                    const int n = 4 * q + j;
                    v[j] = n < N ? (G_host ? G_host[n * N + m] : (n == m ? 1.0f : 0.0f)) : 0.0f;
                }
                g4[m][q] = make_float4(v[0], v[1], v[2], v[3]);
            }
    }
};

// d.x += g * x.x, d.y += g * x.y in ONE issue slot: fma.rn.f32x2 whose first operand is the splat (g, g).  ptxas keeps the
// coefficient in its uniform register and encodes the splat in the instruction (FFMA2 R, R.F32x2.HI_LO, UR.F32, R.F32x2.HI_LO),
// so a column PAIR costs what one column costs with scalar FFMAs.
__device__ __forceinline__ void mix_ffma2(float2& d, float g, float2 x) {
    unsigned long long dd = *reinterpret_cast<unsigned long long*>(&d), gg, xx = *reinterpret_cast<unsigned long long*>(&x);
    asm("mov.b64 %0, {%1, %1};" : "=l"(gg) : "f"(g));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(gg), "l"(xx));
    d = *reinterpret_cast<float2*>(&dd);
}

// out[n][c] = sum_m G^[n][m] in[m][c] for COLS columns held by this thread.  An even number of columns is mixed pairwise on the
// packed FFMA2 pipe (half the issue slots of the N x N FFMA form, same IEEE fma per element).
template <int N, int COLS>
__device__ __forceinline__ void mix_nodes(const MixMat<N>& G, const float (&in)[N][COLS], float (&out)[N][COLS]) {
    if constexpr (COLS % 2 == 0) {
        float2 acc[N][COLS / 2];
#pragma unroll
        for (int n = 0; n < N; ++n)
#pragma unroll
            for (int c = 0; c < COLS / 2; ++c) acc[n][c] = make_float2(0.0f, 0.0f);
#pragma unroll
        for (int m = 0; m < N; ++m)
#pragma unroll
            for (int q = 0; q < MixMat<N>::Q; ++q) {
                const float4 g = G.g4[m][q];
#pragma unroll
                for (int c = 0; c < COLS / 2; ++c) {
                    const float2 x = make_float2(in[m][2 * c], in[m][2 * c + 1]);
                    if (4 * q + 0 < N) mix_ffma2(acc[4 * q + 0][c], g.x, x);
                    if (4 * q + 1 < N) mix_ffma2(acc[4 * q + 1][c], g.y, x);
                    if (4 * q + 2 < N) mix_ffma2(acc[4 * q + 2][c], g.z, x);
                    if (4 * q + 3 < N) mix_ffma2(acc[4 * q + 3][c], g.w, x);
                }
            }
#pragma unroll
        for (int n = 0; n < N; ++n)
#pragma unroll
            for (int c = 0; c < COLS / 2; ++c) { out[n][2 * c] = acc[n][c].x; out[n][2 * c + 1] = acc[n][c].y; }
        return;
    }
#pragma unroll
    for (int m = 0; m < N; ++m)
#pragma unroll
        for (int q = 0; q < MixMat<N>::Q; ++q) {
            const float4 g = G.g4[m][q];
#pragma unroll
            for (int c = 0; c < COLS; ++c) {
                const float x = in[m][c];
                if (m == 0) {
                    if (4 * q + 0 < N) out[4 * q + 0][c] = g.x * x;
                    if (4 * q + 1 < N) out[4 * q + 1][c] = g.y * x;
                    if (4 * q + 2 < N) out[4 * q + 2][c] = g.z * x;
                    if (4 * q + 3 < N) out[4 * q + 3][c] = g.w * x;
                } else {
                    if (4 * q + 0 < N) out[4 * q + 0][c] = fmaf(g.x, x, out[4 * q + 0][c]);
                    if (4 * q + 1 < N) out[4 * q + 1][c] = fmaf(g.y, x, out[4 * q + 1][c]);
                    if (4 * q + 2 < N) out[4 * q + 2][c] = fmaf(g.z, x, out[4 * q + 2][c]);
                    if (4 * q + 3 < N) out[4 * q + 3][c] = fmaf(g.w, x, out[4 * q + 3][c]);
                }
            }
        }
}

}  // namespace sd
