// Backward kernels of the training path (forward() / p_losses with autograd: src/core/diffusion/base.py:262-307,
// src/core/trainer.py:224-234).  fp32 throughout.  The training step differentiates only the rows that carry a loss gradient
// (with the k-best-sample relaxation 1 row in k, see skeletondiffusion_b200/training.py), so these kernels are sized for row
// sets of 10^3 .. 10^5 (sample, node) rows, not for the 537 600-row evaluation batch: straightforward shared-memory tiles, every
// reduction in a fixed order (bitwise repeatable gradients, no atomics).
//
//   StaticGraphLinear  out = G^ (x W[type]^T + b[type])           graph_structural.py:30-43
//       dYm = G^T dOut            node_mix_t_kernel
//       dX  = dYm W[type]         (forward GEMM kernels on a transposed-weight plan: host side)
//       dW[type] = sum_{b, m of type} dYm[b,m]^T x[b,m]     glin_dw_kernel (per node) + reduce_by_type_kernel
//       db[type] = sum_{b, m of type} dYm[b,m]              col_sum_kernel (per node)  + reduce_by_type_kernel
//       dG^[n,m] = sum_{b,o} dOut[b,n,o] (Y[b,m,o] + b[type(m)][o])                     glin_dg_kernel
//   Block              h = tanh(y (scale[t] + 1) + shift[t])     attention.py:66-76      ss_tanh_fwd / ss_tanh_bwd
//   RMSNorm            y = x / |x| * g sqrt(C)                   attention.py:30-36      rmsnorm_fwd / rmsnorm_bwd
//   node attention     softmax(q k^T / sqrt(dh)) v per (sample, head)   attention.py:121-136   node_attention_bwd_kernel
//   loss               mean |S[t] (out - x0)|                    nonisotropic.py:176-190 mahalanobis_bwd_kernel
#include "sd_internal.h"

namespace sd {

// ------------------------------------------------------------------------------------------------ G^T dOut
// out[b,m,c] = sum_n G[n,m] in[b,n,c]   (G row-major [N][N] in global memory; N <= SD_MAX_NODES)
__global__ void __launch_bounds__(256)
node_mix_t_kernel(const float* __restrict__ G, const float* __restrict__ in, float* __restrict__ out, int B, int N, int C) {
    extern __shared__ float gs[];                                   // [N][N]
    for (int i = threadIdx.x; i < N * N; i += blockDim.x) gs[i] = __ldg(G + i);
    __syncthreads();
    const long long total = (long long)B * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / C;
        const int c = (int)(i - b * C);
        const float* src = in + b * N * C + c;
        float v[SD_MAX_NODES];
        for (int n = 0; n < N; ++n) v[n] = __ldg(src + (long long)n * C);
        float* dst = out + b * N * C + c;
        for (int m = 0; m < N; ++m) {
            float acc = 0.f;
            for (int n = 0; n < N; ++n) acc = fmaf(gs[n * N + m], v[n], acc);
            dst[(long long)m * C] = acc;
        }
    }
}

// ------------------------------------------------------------------------------------------------ dW per node
// part[m][o][k] = sum_b dY[b,m,o] x[b,m,k]: block = (node m, 64 x 64 tile of [OUT][K]), 256 threads x (4 x 4), the batch in
// steps of 16 rows through shared memory.
__global__ void __launch_bounds__(256)
glin_dw_kernel(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ part, int B, int N, int OUT, int K) {
    __shared__ float sy[16][64 + 4], sx[16][64 + 4];
    const int m = blockIdx.z, o0 = blockIdx.y * 64, k0 = blockIdx.x * 64;
    const int to = (threadIdx.x >> 4) * 4, tk = (threadIdx.x & 15) * 4;
    float acc[4][4] = {};
    for (int b0 = 0; b0 < B; b0 += 16) {
        for (int i = threadIdx.x; i < 16 * 64; i += 256) {
            const int r = i >> 6, c = i & 63, b = b0 + r;
            sy[r][c] = (b < B && o0 + c < OUT) ? __ldg(dy + ((long long)b * N + m) * OUT + o0 + c) : 0.f;
            sx[r][c] = (b < B && k0 + c < K) ? __ldg(x + ((long long)b * N + m) * K + k0 + c) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const float4 a = *reinterpret_cast<const float4*>(&sy[r][to]);
            const float4 w = *reinterpret_cast<const float4*>(&sx[r][tk]);
            const float av[4] = {a.x, a.y, a.z, a.w}, wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (o0 + to + i < OUT && k0 + tk + j < K) part[((long long)m * OUT + o0 + to + i) * K + k0 + tk + j] = acc[i][j];
}

// part[m][o] = sum_b dY[b,m,o]: block = (node, 64 columns), 4 row groups x 64 columns, fixed-order tree
__global__ void __launch_bounds__(256)
col_sum_kernel(const float* __restrict__ dy, float* __restrict__ part, int B, int N, int OUT) {
    __shared__ float red[4][64];
    const int m = blockIdx.y, o = blockIdx.x * 64 + (threadIdx.x & 63), g = threadIdx.x >> 6;
    float acc = 0.f;
    if (o < OUT)
        for (int b = g; b < B; b += 4) acc += __ldg(dy + ((long long)b * N + m) * OUT + o);
    red[g][threadIdx.x & 63] = acc;
    __syncthreads();
    if (g == 0 && o < OUT) part[(long long)m * OUT + o] = (red[0][threadIdx.x] + red[1][threadIdx.x]) + (red[2][threadIdx.x] + red[3][threadIdx.x]);
}

// out[type][i] (+)= sum over the nodes of that type, in node order, of part[node][i]
__global__ void reduce_by_type_kernel(const float* __restrict__ part, float* __restrict__ out, NodeTypes types, int N, int n_types,
                                      long long per_node, int accumulate) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= per_node) return;
    for (int t = 0; t < n_types; ++t) {
        float acc = 0.f;
        bool any = false;
        for (int m = 0; m < N; ++m)
            if (types.t[m] == t) { acc += part[(long long)m * per_node + i]; any = true; }
        float* dst = out + (long long)t * per_node + i;
        if (any || !accumulate) *dst = accumulate ? *dst + acc : acc;
    }
}

// ------------------------------------------------------------------------------------------------ dG^
// part[blk][n][m] = sum over the block's samples and all o of dOut[b,n,o] (Y[b,m,o] + bias[type(m)][o]); a second launch folds
// the block partials in block order.  One sample at a time in shared memory, OUT in slices of 96 columns.
constexpr int DG_SLICE = 96;
__global__ void __launch_bounds__(256)
glin_dg_kernel(const float* __restrict__ dout, const float* __restrict__ y, const float* __restrict__ bias, NodeTypes types,
               float* __restrict__ part, int B, int N, int OUT) {
    extern __shared__ float dg_sm[];                                // dOut and Y slices: 2 x [N][DG_SLICE + 1]
    constexpr int LD = DG_SLICE + 1;
    float* sd_ = dg_sm;
    float* sy_ = dg_sm + (size_t)N * LD;
    float acc[8] = {};                                              // pairs p = threadIdx.x + 256 i  (N * N <= 4096 pairs: up to 16; SD_MAX_NODES = 64)
    float acc2[8] = {};
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        for (int s0 = 0; s0 < OUT; s0 += DG_SLICE) {
            const int w = min(DG_SLICE, OUT - s0);
            __syncthreads();
            for (int i = threadIdx.x; i < N * w; i += 256) {
                const int n = i / w, c = i - n * w;
                const long long off = ((long long)b * N + n) * OUT + s0 + c;
                sd_[n * LD + c] = __ldg(dout + off);
                sy_[n * LD + c] = __ldg(y + off) + (bias ? __ldg(bias + (long long)types.t[n] * OUT + s0 + c) : 0.f);
            }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int p = threadIdx.x + 256 * i;
                if (p < N * N) {
                    const int n = p / N, m = p - n * N;
                    float a = 0.f;
                    for (int c = 0; c < w; ++c) a = fmaf(sd_[n * LD + c], sy_[m * LD + c], a);
                    if (i < 8) acc[i] += a; else acc2[i - 8] += a;
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int p = threadIdx.x + 256 * i;
        if (p < N * N) part[(long long)blockIdx.x * N * N + p] = i < 8 ? acc[i] : acc2[i - 8];
    }
}
__global__ void fold_partials_kernel(const float* __restrict__ part, float* __restrict__ out, int blocks, int count, int accumulate) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    float acc = 0.f;
    for (int k = 0; k < blocks; ++k) acc += part[(long long)k * count + i];
    out[i] = accumulate ? out[i] + acc : acc;
}

// ------------------------------------------------------------------------------------------------ Block: scale / shift / tanh
// h = tanh(y * (ss[t_b][c] + 1) + ss[t_b][C + c]);  ss_table [rows][2C] or null (then h = tanh(y))
__global__ void ss_tanh_fwd_kernel(const float* __restrict__ y, const float* __restrict__ ss, const int* __restrict__ t, float* __restrict__ h,
                                   long long total, int N, int C) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const long long b = i / ((long long)N * C);
        float v = y[i];
        if (ss) { const float* row = ss + (long long)__ldg(t + b) * 2 * C; v = fmaf(v, __ldg(row + c) + 1.0f, __ldg(row + C + c)); }
        h[i] = tanhf(v);
    }
}
// dy = dh (1 - h^2) (scale + 1);  dss_rows[b][c] = sum_n dh (1 - h^2) y,  dss_rows[b][C + c] = sum_n dh (1 - h^2)   (block = sample)
__global__ void __launch_bounds__(256)
ss_tanh_bwd_kernel(const float* __restrict__ dh, const float* __restrict__ h, const float* __restrict__ y, const float* __restrict__ ss,
                   const int* __restrict__ t, float* __restrict__ dy, float* __restrict__ dss_rows, int N, int C) {
    const long long b = blockIdx.x;
    const float* row = ss ? ss + (long long)__ldg(t + b) * 2 * C : nullptr;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float mul = row ? __ldg(row + c) + 1.0f : 1.0f;
        float ds = 0.f, dsh = 0.f;
        for (int n = 0; n < N; ++n) {
            const long long i = (b * N + n) * C + c;
            const float hv = h[i];
            const float dpre = dh[i] * (1.0f - hv * hv);
            dy[i] = dpre * mul;
            if (row) { ds = fmaf(dpre, y[i], ds); dsh += dpre; }
        }
        if (row) { dss_rows[b * 2 * C + c] = ds; dss_rows[b * 2 * C + C + c] = dsh; }
    }
}

// ------------------------------------------------------------------------------------------------ RMSNorm
// y = x * inv * gs,  inv = 1 / max(|x|, 1e-12),  gs[c] = g[c] sqrt(C)      (warp per row)
__global__ void __launch_bounds__(256)
rmsnorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ g, float* __restrict__ y, float* __restrict__ inv_out, long long rows, int C) {
    const long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= rows) return;
    const float* xr = x + r * C;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) { const float v = xr[c]; s = fmaf(v, v, s); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float inv = 1.0f / fmaxf(sqrtf(s), 1e-12f);
    const float sq = sqrtf((float)C);
    for (int c = lane; c < C; c += 32) y[r * C + c] = xr[c] * inv * (__ldg(g + c) * sq);
    if (lane == 0) inv_out[r] = inv;
}
// dx = inv (u - xh sum_c(u xh)), u = dy gs, xh = x inv;   dg_part[blk][c] = sqrt(C) sum over the block's rows of dy xh
__global__ void __launch_bounds__(256)
rmsnorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ inv_in, const float* __restrict__ g,
                   float* __restrict__ dx, float* __restrict__ dg_part, long long rows, int C, int rows_per_block) {
    extern __shared__ float dg_s[];                                 // [8 warps][C]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int c = lane; c < C; c += 32) dg_s[warp * C + c] = 0.f;
    const float sq = sqrtf((float)C);
    const long long r0 = (long long)blockIdx.x * rows_per_block;
    for (long long r = r0 + warp; r < r0 + rows_per_block && r < rows; r += 8) {
        const float inv = inv_in[r];
        float dot = 0.f;
        for (int c = lane; c < C; c += 32) dot = fmaf(dy[r * C + c] * (__ldg(g + c) * sq), x[r * C + c] * inv, dot);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
        for (int c = lane; c < C; c += 32) {
            const float xh = x[r * C + c] * inv, d = dy[r * C + c];
            dx[r * C + c] = inv * (d * (__ldg(g + c) * sq) - xh * dot);
            dg_s[warp * C + c] += d * xh * sq;
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float a = 0.f;
        for (int w = 0; w < 8; ++w) a += dg_s[w * C + c];
        dg_part[(long long)blockIdx.x * C + c] = a;
    }
}

// ------------------------------------------------------------------------------------------------ node attention backward
// qkv [B][N][3 H DH] (q | k | v, each [head][DH]), dout [B][N][H DH] -> dqkv.  Warp = (sample, head), lane = node (N <= 32),
// DH = 32.  P is recomputed (softmax of q k^T / sqrt(DH)); dV = P^T dO, dP = dO V^T, dS = P (dP - rowsum(dP P)),
// dQ = dS K / sqrt(DH), dK = dS^T Q / sqrt(DH).
template <int DH>
__global__ void __launch_bounds__(128)
node_attention_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ dout, float* __restrict__ dqkv, int B, int N, int H) {
    extern __shared__ float sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long item = (long long)blockIdx.x * 4 + warp;
    if (item >= (long long)B * H) return;
    const int b = (int)(item / H), h = (int)(item % H);
    float* base = sm + warp * (5 * 32 * (DH + 1) + 32 * 33);
    float (*q)[DH + 1] = reinterpret_cast<float (*)[DH + 1]>(base);
    float (*k)[DH + 1] = reinterpret_cast<float (*)[DH + 1]>(base + 32 * (DH + 1));
    float (*v)[DH + 1] = reinterpret_cast<float (*)[DH + 1]>(base + 2 * 32 * (DH + 1));
    float (*go)[DH + 1] = reinterpret_cast<float (*)[DH + 1]>(base + 3 * 32 * (DH + 1));
    float (*ds)[33] = reinterpret_cast<float (*)[33]>(base + 4 * 32 * (DH + 1));           // dS[n][j]
    float (*pp)[33] = reinterpret_cast<float (*)[33]>(base + 4 * 32 * (DH + 1) + 32 * 33); // P[n][j]  (space of the 5th tile)
    const int ROW = 3 * H * DH, OROW = H * DH;
    const float scale = rsqrtf((float)DH);
    for (int n = 0; n < N; ++n) {                                   // coalesced: lane = channel
        const float* r = qkv + ((long long)b * N + n) * ROW + h * DH;
        q[n][lane] = r[lane]; k[n][lane] = r[H * DH + lane]; v[n][lane] = r[2 * H * DH + lane];
        go[n][lane] = dout[((long long)b * N + n) * OROW + h * DH + lane];
    }
    __syncwarp();
    const int n = lane;
    if (n < N) {
        float s[32], mx = -INFINITY;
        for (int j = 0; j < N; ++j) {
            float a = 0.f;
#pragma unroll
            for (int c = 0; c < DH; ++c) a = fmaf(q[n][c] * scale, k[j][c], a);
            s[j] = a; mx = fmaxf(mx, a);
        }
        float sum = 0.f;
        for (int j = 0; j < N; ++j) { s[j] = expf(s[j] - mx); sum += s[j]; }
        const float inv = 1.0f / sum;
        float dot = 0.f, dp[32];
        for (int j = 0; j < N; ++j) {
            s[j] *= inv;
            float a = 0.f;
#pragma unroll
            for (int c = 0; c < DH; ++c) a = fmaf(go[n][c], v[j][c], a);
            dp[j] = a; dot = fmaf(a, s[j], dot);
        }
        for (int j = 0; j < N; ++j) { pp[n][j] = s[j]; ds[n][j] = s[j] * (dp[j] - dot); }
    }
    __syncwarp();
    // outputs: lane = channel, loop over rows (coalesced stores)
    for (int r = 0; r < N; ++r) {
        float dq = 0.f, dk = 0.f, dv = 0.f;
        for (int j = 0; j < N; ++j) {
            dq = fmaf(ds[r][j], k[j][lane], dq);
            dk = fmaf(ds[j][r], q[j][lane], dk);
            dv = fmaf(pp[j][r], go[j][lane], dv);
        }
        float* o = dqkv + ((long long)b * N + r) * ROW + h * DH;
        o[lane] = dq * scale; o[H * DH + lane] = dk * scale; o[2 * H * DH + lane] = dv;
    }
}

// ------------------------------------------------------------------------------------------------ loss backward
// loss_b = mean_{n,d} |S[t_b] (out_b - x0_b)|;  dout_b = S[t_b]^T sign(S[t_b] (out_b - x0_b)) g_b / (N D)      (block = sample)
__global__ void __launch_bounds__(256)
mahalanobis_bwd_kernel(const float* __restrict__ out, const float* __restrict__ x0, const int* __restrict__ t, const float* __restrict__ S,
                       const float* __restrict__ gl, float* __restrict__ dout, int N, int D) {
    extern __shared__ float sg[];                                   // sign tile [N][D]
    const int b = blockIdx.x;
    const float* St = S + (long long)__ldg(t + b) * N * N;
    const long long base = (long long)b * N * D;
    for (int i = threadIdx.x; i < N * D; i += blockDim.x) {
        const int n = i / D, d = i % D;
        float acc = 0.f;
        for (int k = 0; k < N; ++k) acc = fmaf(__ldg(St + n * N + k), out[base + k * D + d] - x0[base + k * D + d], acc);
        sg[i] = acc > 0.f ? 1.f : (acc < 0.f ? -1.f : 0.f);
    }
    __syncthreads();
    const float w = gl[b] / (float)(N * D);
    for (int i = threadIdx.x; i < N * D; i += blockDim.x) {
        const int k = i / D, d = i % D;
        float acc = 0.f;
        for (int n = 0; n < N; ++n) acc = fmaf(__ldg(St + n * N + k), sg[n * D + d], acc);
        dout[base + i] = acc * w;
    }
}

// ================================================================================================ host
static int grid_for(long long total, int block) {
    long long g = (total + block - 1) / block;
    const long long cap = (long long)sm_count() * 16;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

int node_mix_t_fp32(const float* G, const float* in, float* out, int B, int N, int C, cudaStream_t st) {
    node_mix_t_kernel<<<grid_for((long long)B * C, 256), 256, (size_t)N * N * sizeof(float), st>>>(G, in, out, B, N, C);
    SD_LAUNCH_OK("node_mix_t_kernel");
    return SD_OK;
}

size_t glin_backward_scratch_floats(int B, int N, int OUT, int K) {
    const size_t dg_blocks = (size_t)sm_count() * 2;
    size_t a = (size_t)N * OUT * K + (size_t)N * OUT;               // per-node dW / db partials
    size_t b = dg_blocks * N * N;
    (void)B;
    return (a > b ? a : b) + 64;
}

int glin_backward_params_fp32(const sd_glin* L, const float* x, const float* dym, const float* dout, const float* y_raw, const float* bias_types,
                              float* dW, float* dbias, float* dG, float* scratch, int B, int accumulate, cudaStream_t st) {
    const int N = L->N, OUT = L->OUT, K = L->K;
    if (dW) {
        dim3 grid((K + 63) / 64, (OUT + 63) / 64, N);
        glin_dw_kernel<<<grid, 256, 0, st>>>(dym, x, scratch, B, N, OUT, K);
        SD_LAUNCH_OK("glin_dw_kernel");
        const long long per = (long long)OUT * K;
        reduce_by_type_kernel<<<(int)((per + 255) / 256), 256, 0, st>>>(scratch, dW, L->types, N, L->n_types, per, accumulate);
        SD_LAUNCH_OK("reduce_by_type_kernel");
    }
    if (dbias) {
        float* part = scratch + (size_t)N * OUT * K;
        col_sum_kernel<<<dim3((OUT + 63) / 64, N), 256, 0, st>>>(dym, part, B, N, OUT);
        SD_LAUNCH_OK("col_sum_kernel");
        reduce_by_type_kernel<<<(OUT + 255) / 256, 256, 0, st>>>(part, dbias, L->types, N, L->n_types, OUT, accumulate);
        SD_LAUNCH_OK("reduce_by_type_kernel");
    }
    if (dG) {
        if (N * N > 4096) { set_error("sd_glin_backward: more than 64 nodes"); return SD_ERR_UNSUPPORTED; }
        int blocks = sm_count() * 2;
        if (blocks > B) blocks = B;
        const size_t smem = (size_t)2 * N * (DG_SLICE + 1) * sizeof(float);
        static unsigned long long configured = 0;
        if (smem > 48 * 1024) { if (int rc = opt_in_smem(glin_dg_kernel, smem, configured)) return rc; }
        glin_dg_kernel<<<blocks, 256, smem, st>>>(dout, y_raw, bias_types, L->types, scratch, B, N, OUT);
        SD_LAUNCH_OK("glin_dg_kernel");
        fold_partials_kernel<<<(N * N + 255) / 256, 256, 0, st>>>(scratch, dG, blocks, N * N, accumulate);
        SD_LAUNCH_OK("fold_partials_kernel");
    }
    return SD_OK;
}

int ss_tanh_fwd_fp32(const float* y, const float* ss, const int* t, float* h, int B, int N, int C, cudaStream_t st) {
    const long long total = (long long)B * N * C;
    ss_tanh_fwd_kernel<<<grid_for(total, 256), 256, 0, st>>>(y, ss, t, h, total, N, C);
    SD_LAUNCH_OK("ss_tanh_fwd_kernel");
    return SD_OK;
}
int ss_tanh_bwd_fp32(const float* dh, const float* h, const float* y, const float* ss, const int* t, float* dy, float* dss_rows,
                     int B, int N, int C, cudaStream_t st) {
    ss_tanh_bwd_kernel<<<B, 256, 0, st>>>(dh, h, y, ss, t, dy, dss_rows, N, C);
    SD_LAUNCH_OK("ss_tanh_bwd_kernel");
    return SD_OK;
}
int rmsnorm_fwd_fp32(const float* x, const float* g, float* y, float* inv, long long rows, int C, cudaStream_t st) {
    rmsnorm_fwd_kernel<<<(int)((rows + 7) / 8), 256, 0, st>>>(x, g, y, inv, rows, C);
    SD_LAUNCH_OK("rmsnorm_fwd_kernel");
    return SD_OK;
}
int rmsnorm_bwd_blocks(long long rows) { long long b = (rows + 255) / 256; return (int)(b < 1 ? 1 : b); }
int rmsnorm_bwd_fp32(const float* dy, const float* x, const float* inv, const float* g, float* dx, float* dg_part, long long rows, int C, cudaStream_t st) {
    const int blocks = rmsnorm_bwd_blocks(rows);
    rmsnorm_bwd_kernel<<<blocks, 256, (size_t)8 * C * sizeof(float), st>>>(dy, x, inv, g, dx, dg_part, rows, C, 256);
    SD_LAUNCH_OK("rmsnorm_bwd_kernel");
    return SD_OK;
}
int node_attention_bwd_fp32(const float* qkv, const float* dout, float* dqkv, int B, int N, int H, int DH, cudaStream_t st) {
    if (DH != 32 || N > 32) { set_error("sd_node_attention_backward: dim_head must be 32 and nodes <= 32 (got %d, %d)", DH, N); return SD_ERR_UNSUPPORTED; }
    const size_t per_warp = (size_t)(5 * 32 * 33 + 32 * 33) * sizeof(float);
    const size_t smem = 4 * per_warp;
    auto kern = node_attention_bwd_kernel<32>;
    static unsigned long long configured = 0;
    if (int rc = opt_in_smem(kern, smem, configured)) return rc;
    const long long items = (long long)B * H;
    kern<<<(int)((items + 3) / 4), 128, smem, st>>>(qkv, dout, dqkv, B, N, H);
    SD_LAUNCH_OK("node_attention_bwd_kernel");
    return SD_OK;
}
int mahalanobis_bwd_fp32(const float* out, const float* x0, const int* t, const float* S, const float* gl, float* dout, int B, int N, int D, cudaStream_t st) {
    mahalanobis_bwd_kernel<<<B, 256, (size_t)N * D * sizeof(float), st>>>(out, x0, t, S, gl, dout, N, D);
    SD_LAUNCH_OK("mahalanobis_bwd_kernel");
    return SD_OK;
}

}  // namespace sd

using namespace sd;

extern "C" {

int sd_node_mix_transposed(const float* g_dev, const float* in_dev, float* out_dev, int batch, int num_nodes, int width, void* stream) {
    if (batch == 0) return SD_OK;
    if (!g_dev || !in_dev || !out_dev || batch < 0 || num_nodes <= 0 || num_nodes > SD_MAX_NODES || width <= 0) { set_error("sd_node_mix_transposed: invalid arguments"); return SD_ERR_INVALID; }
    return node_mix_t_fp32(g_dev, in_dev, out_dev, batch, num_nodes, width, static_cast<cudaStream_t>(stream));
}

size_t sd_glin_backward_scratch_bytes(const sd_glin* L, int batch) {
    if (!L) return 0;
    return glin_backward_scratch_floats(batch, L->N, L->OUT, L->K) * sizeof(float);
}

int sd_glin_backward_params(const sd_glin* L, const float* x_dev, const float* dym_dev, const float* dout_dev, const float* y_raw_dev,
                            const float* bias_types_dev, float* dweight_dev, float* dbias_dev, float* dg_dev, float* scratch_dev,
                            int batch, int accumulate, void* stream) {
    if (batch == 0) return SD_OK;
    if (!L || !scratch_dev || batch < 0) { set_error("sd_glin_backward_params: invalid arguments"); return SD_ERR_INVALID; }
    if ((dweight_dev && (!x_dev || !dym_dev)) || (dbias_dev && !dym_dev) || (dg_dev && (!dout_dev || !y_raw_dev))) {
        set_error("sd_glin_backward_params: a requested gradient misses its operands"); return SD_ERR_INVALID;
    }
    return glin_backward_params_fp32(L, x_dev, dym_dev, dout_dev, y_raw_dev, bias_types_dev, dweight_dev, dbias_dev, dg_dev, scratch_dev,
                                     batch, accumulate, static_cast<cudaStream_t>(stream));
}

int sd_ss_tanh_forward(const float* y_dev, const float* ss_table_dev, const int32_t* t_dev, float* h_dev, int batch, int num_nodes, int width, void* stream) {
    if (batch == 0) return SD_OK;
    if (!y_dev || !h_dev || (ss_table_dev && !t_dev)) { set_error("sd_ss_tanh_forward: null argument"); return SD_ERR_INVALID; }
    return ss_tanh_fwd_fp32(y_dev, ss_table_dev, t_dev, h_dev, batch, num_nodes, width, static_cast<cudaStream_t>(stream));
}

int sd_ss_tanh_backward(const float* dh_dev, const float* h_dev, const float* y_dev, const float* ss_table_dev, const int32_t* t_dev,
                        float* dy_dev, float* dss_rows_dev, int batch, int num_nodes, int width, void* stream) {
    if (batch == 0) return SD_OK;
    if (!dh_dev || !h_dev || !dy_dev || (ss_table_dev && (!t_dev || !y_dev || !dss_rows_dev))) { set_error("sd_ss_tanh_backward: null argument"); return SD_ERR_INVALID; }
    return ss_tanh_bwd_fp32(dh_dev, h_dev, y_dev, ss_table_dev, t_dev, dy_dev, dss_rows_dev, batch, num_nodes, width, static_cast<cudaStream_t>(stream));
}

int sd_rmsnorm_forward(const float* x_dev, const float* g_dev, float* y_dev, float* inv_dev, int64_t rows, int width, void* stream) {
    if (rows == 0) return SD_OK;
    if (!x_dev || !g_dev || !y_dev || !inv_dev || rows < 0 || width <= 0) { set_error("sd_rmsnorm_forward: invalid arguments"); return SD_ERR_INVALID; }
    return rmsnorm_fwd_fp32(x_dev, g_dev, y_dev, inv_dev, rows, width, static_cast<cudaStream_t>(stream));
}

int sd_rmsnorm_backward_blocks(int64_t rows) { return rmsnorm_bwd_blocks(rows); }

int sd_rmsnorm_backward(const float* dy_dev, const float* x_dev, const float* inv_dev, const float* g_dev, float* dx_dev, float* dg_part_dev,
                        int64_t rows, int width, void* stream) {
    if (rows == 0) return SD_OK;
    if (!dy_dev || !x_dev || !inv_dev || !g_dev || !dx_dev || !dg_part_dev || width <= 0 || width > 1024) { set_error("sd_rmsnorm_backward: invalid arguments"); return SD_ERR_INVALID; }
    return rmsnorm_bwd_fp32(dy_dev, x_dev, inv_dev, g_dev, dx_dev, dg_part_dev, rows, width, static_cast<cudaStream_t>(stream));
}

int sd_node_attention_backward(const float* qkv_dev, const float* dout_dev, float* dqkv_dev, int batch, int num_nodes, int heads, int dim_head, void* stream) {
    if (batch == 0) return SD_OK;
    if (!qkv_dev || !dout_dev || !dqkv_dev || batch < 0) { set_error("sd_node_attention_backward: invalid arguments"); return SD_ERR_INVALID; }
    return node_attention_bwd_fp32(qkv_dev, dout_dev, dqkv_dev, batch, num_nodes, heads, dim_head, static_cast<cudaStream_t>(stream));
}

int sd_mahalanobis_loss_backward(const float* out_dev, const float* x0_dev, const int32_t* t_dev, const float* s_dev, const float* grad_loss_dev,
                                 float* dout_dev, int batch, int num_nodes, int latent_dim, void* stream) {
    if (batch == 0) return SD_OK;
    if (!out_dev || !x0_dev || !t_dev || !s_dev || !grad_loss_dev || !dout_dev) { set_error("sd_mahalanobis_loss_backward: null argument"); return SD_ERR_INVALID; }
    if ((size_t)num_nodes * latent_dim * sizeof(float) > 48 * 1024) { set_error("sd_mahalanobis_loss_backward: sample of %d x %d floats exceeds 48 KB", num_nodes, latent_dim); return SD_ERR_UNSUPPORTED; }
    return mahalanobis_bwd_fp32(out_dev, x0_dev, t_dev, s_dev, grad_loss_dev, dout_dev, batch, num_nodes, latent_dim, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
