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

// out[n][c] = sum_m G^[n][m] in[m][c] for COLS columns held by this thread
template <int N, int COLS>
__device__ __forceinline__ void mix_nodes(const MixMat<N>& G, const float (&in)[N][COLS], float (&out)[N][COLS]) {
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
