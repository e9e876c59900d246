// Non-GEMM kernels of the sampling path (fp32): node attention, RMSNorm row factor, the fused
// reverse-diffusion step, GRU gates, time-conditioning table, Philox normal fill, q_sample, loss.
#include "sd_internal.h"
#include <math.h>

namespace sd {

// =============================================================================================
// RMSNorm row factor: inv[r] = 1 / max(||x[r,:]||, 1e-12)        (layers/attention.py:36)
// =============================================================================================
__global__ void row_inv_norm_kernel(const float* __restrict__ x, float* __restrict__ inv, long long rows, int width) {
    const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= rows) return;
    const float* xr = x + r * width;
    float s = 0.0f;
    for (int i = lane; i < width; i += 32) { const float v = __ldg(xr + i); s = fmaf(v, v, s); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) inv[r] = 1.0f / fmaxf(sqrtf(s), 1e-12f);
}

int row_inv_norm_fp32(const float* x, float* inv, long long rows, int width, cudaStream_t st) {
    if (rows <= 0) return SD_OK;
    row_inv_norm_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(x, inv, rows, width);
    SD_LAUNCH_OK("row_inv_norm_kernel");
    return SD_OK;
}

// =============================================================================================
// fused reverse-diffusion step
//   x_{t-1} = C1[t] clamp(x0) + C2[t] x_t + S[t] eps          (base.py:314-341, nonisotropic.py:196-210)
// thread = (sample, 4 latent channels); the three [N,N] tables of step t sit transposed in smem and
// are read as warp-wide broadcasts; each input element is read once, each output written once.
// =============================================================================================
__device__ __forceinline__ void rs_ffma2(float2& d, float a, float2 b) {
    unsigned long long dd = *reinterpret_cast<unsigned long long*>(&d), bb = *reinterpret_cast<unsigned long long*>(&b), aa;
    asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(aa), "l"(bb));
    d = *reinterpret_cast<float2*>(&dd);
}

// thread = (sample, 2 latent channels): N float2 accumulators, FFMA2 with the table entry as the broadcast scalar
// (SASS: FFMA2 Rd, Ra.F32, Rb.F32x2, Rd): N FFMA2 per input row.  Two channels per thread keep the kernel at ~90
// registers (4-5 CTAs per SM); a 4-channel version needed 192 and ran at 8 warps per SM.  Input rows are fetched in
// batches of RB float2 so that RB loads are in flight per thread before the first use.
template <int N>
__global__ void __launch_bounds__(128, 4)
reverse_step_kernel(const float* __restrict__ c1, const float* __restrict__ c2, const float* __restrict__ sm,
                    const float* __restrict__ x_t, const float* __restrict__ x0, const View eps,
                    float* __restrict__ x_out, float* __restrict__ mean_out, long long mean_sb, int D, int B, int clip) {
    constexpr int NP = (N + 3) & ~3;                 // padded column count for float4 broadcast reads
    constexpr int RB = 7;                            // rows per load batch
    __shared__ __align__(16) float T[3][N][NP];      // T[i][k][n] = table_i[n][k]
    for (int i = threadIdx.x; i < 3 * N * NP; i += blockDim.x) {
        const int which = i / (N * NP), k = (i / NP) % N, n = i % NP;
        const float* src = which == 0 ? c1 : (which == 1 ? c2 : sm);
        T[which][k][n] = (n < N && !(which == 2 && eps.ptr == nullptr)) ? __ldg(src + n * N + k) : 0.0f;
    }
    __syncthreads();
    const int d2 = D >> 1;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)B * d2) return;
    const int b = (int)(gid / d2), d = (int)(gid % d2) * 2;
    float2 acc[N];
#pragma unroll
    for (int n = 0; n < N; ++n) acc[n] = make_float2(0.f, 0.f);

    auto accumulate = [&](const float* in, long long node_stride, int which, bool clamp) {
#pragma unroll
        for (int k0 = 0; k0 < N; k0 += RB) {
            float2 v[RB];
#pragma unroll
            for (int r = 0; r < RB; ++r)
                if (k0 + r < N) v[r] = __ldg(reinterpret_cast<const float2*>(in + (k0 + r) * node_stride));
#pragma unroll
            for (int r = 0; r < RB; ++r) {
                if (k0 + r < N) {
                    float2 x = v[r];
                    if (clamp) { x.x = fminf(fmaxf(x.x, -1.f), 1.f); x.y = fminf(fmaxf(x.y, -1.f), 1.f); }
#pragma unroll
                    for (int n = 0; n < NP; n += 4) {
                        const float4 m = *reinterpret_cast<const float4*>(&T[which][k0 + r][n]);
                        const float mm[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (n + e < N) rs_ffma2(acc[n + e], mm[e], x);
                    }
                }
            }
        }
    };
    const long long off = (long long)b * N * D + d;
    accumulate(x0 + off, D, 0, clip != 0);
    accumulate(x_t + off, D, 1, false);
    if (mean_out) {
#pragma unroll
        for (int n = 0; n < N; ++n) *reinterpret_cast<float2*>(mean_out + (long long)b * mean_sb + d + (long long)n * D) = acc[n];
    }
    if (eps.ptr) accumulate(row_ptr(eps, b, 0) + d, eps.sn, 2, false);
#pragma unroll
    for (int n = 0; n < N; ++n) *reinterpret_cast<float2*>(x_out + off + (long long)n * D) = acc[n];
}

// diagonal tables (U == I, the isotropic degenerate path): pure streaming kernel
__global__ void __launch_bounds__(256)
reverse_step_diag_kernel(const float* __restrict__ c1, const float* __restrict__ c2, const float* __restrict__ sm,
                         const float* __restrict__ x_t, const float* __restrict__ x0, const View eps,
                         float* __restrict__ x_out, float* __restrict__ mean_out, long long mean_sb, int N, int D, int B, int clip) {
    const int d4 = D >> 2;
    const long long total = (long long)B * N * d4;
    for (long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x; gid < total; gid += (long long)gridDim.x * blockDim.x) {
        const int d = (int)(gid % d4) * 4;
        const long long bn = gid / d4;
        const int n = (int)(bn % N), b = (int)(bn / N);
        const float a = __ldg(c1 + n * N + n), bb = __ldg(c2 + n * N + n);
        const long long off = bn * D + d;
        float4 c = __ldg(reinterpret_cast<const float4*>(x0 + off));
        const float4 x = __ldg(reinterpret_cast<const float4*>(x_t + off));
        if (clip) {
            c.x = fminf(fmaxf(c.x, -1.f), 1.f); c.y = fminf(fmaxf(c.y, -1.f), 1.f);
            c.z = fminf(fmaxf(c.z, -1.f), 1.f); c.w = fminf(fmaxf(c.w, -1.f), 1.f);
        }
        float4 m = make_float4(fmaf(bb, x.x, a * c.x), fmaf(bb, x.y, a * c.y), fmaf(bb, x.z, a * c.z), fmaf(bb, x.w, a * c.w));
        if (mean_out) *reinterpret_cast<float4*>(mean_out + (long long)b * mean_sb + (long long)n * D + d) = m;
        if (eps.ptr) {
            const float s = __ldg(sm + n * N + n);
            const float4 e = __ldg(reinterpret_cast<const float4*>(row_ptr(eps, b, n) + d));
            m.x = fmaf(s, e.x, m.x); m.y = fmaf(s, e.y, m.y); m.z = fmaf(s, e.z, m.z); m.w = fmaf(s, e.w, m.w);
        }
        *reinterpret_cast<float4*>(x_out + off) = m;
    }
}

// any N <= 64: thread = (sample, node, 4 channels); tables read through L1
__global__ void __launch_bounds__(256)
reverse_step_generic_kernel(const float* __restrict__ c1, const float* __restrict__ c2, const float* __restrict__ sm,
                            const float* __restrict__ x_t, const float* __restrict__ x0, const View eps,
                            float* __restrict__ x_out, float* __restrict__ mean_out, long long mean_sb, int N, int D, int B, int clip) {
    const int d4 = D >> 2;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)B * N * d4) return;
    const int d = (int)(gid % d4) * 4;
    const long long bn = gid / d4;
    const int n = (int)(bn % N), b = (int)(bn / N);
    const long long base = (long long)b * N * D + d;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = 0; k < N; ++k) {
        float4 c = __ldg(reinterpret_cast<const float4*>(x0 + base + (long long)k * D));
        if (clip) {
            c.x = fminf(fmaxf(c.x, -1.f), 1.f); c.y = fminf(fmaxf(c.y, -1.f), 1.f);
            c.z = fminf(fmaxf(c.z, -1.f), 1.f); c.w = fminf(fmaxf(c.w, -1.f), 1.f);
        }
        const float a = __ldg(c1 + n * N + k);
        acc.x = fmaf(a, c.x, acc.x); acc.y = fmaf(a, c.y, acc.y); acc.z = fmaf(a, c.z, acc.z); acc.w = fmaf(a, c.w, acc.w);
    }
    for (int k = 0; k < N; ++k) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(x_t + base + (long long)k * D));
        const float a = __ldg(c2 + n * N + k);
        acc.x = fmaf(a, x.x, acc.x); acc.y = fmaf(a, x.y, acc.y); acc.z = fmaf(a, x.z, acc.z); acc.w = fmaf(a, x.w, acc.w);
    }
    if (mean_out) *reinterpret_cast<float4*>(mean_out + (long long)b * mean_sb + (long long)n * D + d) = acc;
    if (eps.ptr) {
        for (int k = 0; k < N; ++k) {
            const float4 e = __ldg(reinterpret_cast<const float4*>(row_ptr(eps, b, k) + d));
            const float a = __ldg(sm + n * N + k);
            acc.x = fmaf(a, e.x, acc.x); acc.y = fmaf(a, e.y, acc.y); acc.z = fmaf(a, e.z, acc.z); acc.w = fmaf(a, e.w, acc.w);
        }
    }
    *reinterpret_cast<float4*>(x_out + base + (long long)n * D) = acc;
}

template <int N>
static int launch_reverse_step(const float* c1, const float* c2, const float* s, const float* x_t, const float* x0,
                               const View& eps, float* x_out, float* mean_out, long long mean_sb, int D, int B, int clip, cudaStream_t st) {
    const long long total = (long long)B * (D >> 1);
    reverse_step_kernel<N><<<(unsigned)((total + 127) / 128), 128, 0, st>>>(c1, c2, s, x_t, x0, eps, x_out, mean_out, mean_sb, D, B, clip);
    SD_LAUNCH_OK("reverse_step_kernel");
    return SD_OK;
}

int reverse_step_fp32(const sd_diffusion* df, const float* x_t, const float* x0, const View* eps_in,
                      float* x_out, float* mean_out, long long mean_sb, int t, int B, int clip, cudaStream_t st) {
    if (B <= 0) return SD_OK;
    if (t < 0 || t >= df->T) { set_error("reverse_step: t=%d outside [0,%d)", t, df->T); return SD_ERR_INVALID; }
    const int N = df->N, D = df->D;
    if (D % 4 != 0) { set_error("reverse_step: latent_dim %d must be a multiple of 4", D); return SD_ERR_UNSUPPORTED; }
    View eps; eps.ptr = nullptr; eps.sb = 0; eps.sn = 0; eps.rep = 1; eps.width = D;
    if (eps_in && eps_in->ptr) eps = *eps_in;
    const float* c1 = df->c1 + (long long)t * N * N;
    const float* c2 = df->c2 + (long long)t * N * N;
    const float* s = df->s + (long long)t * N * N;
    if (df->diagonal[t]) {
        const long long total = (long long)B * N * (D >> 2);
        long long blocks = (total + 255) / 256;
        if (blocks > 148LL * 32) blocks = 148LL * 32;
        reverse_step_diag_kernel<<<(unsigned)blocks, 256, 0, st>>>(c1, c2, s, x_t, x0, eps, x_out, mean_out, mean_sb, N, D, B, clip);
        SD_LAUNCH_OK("reverse_step_diag_kernel");
        return SD_OK;
    }
    switch (N) {
        case 16: return launch_reverse_step<16>(c1, c2, s, x_t, x0, eps, x_out, mean_out, mean_sb, D, B, clip, st);
        case 17: return launch_reverse_step<17>(c1, c2, s, x_t, x0, eps, x_out, mean_out, mean_sb, D, B, clip, st);
        case 21: return launch_reverse_step<21>(c1, c2, s, x_t, x0, eps, x_out, mean_out, mean_sb, D, B, clip, st);
        default: break;
    }
    const long long total = (long long)B * N * (D >> 2);
    reverse_step_generic_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(c1, c2, s, x_t, x0, eps, x_out, mean_out, mean_sb, N, D, B, clip);
    SD_LAUNCH_OK("reverse_step_generic_kernel");
    return SD_OK;
}

// =============================================================================================
// q_sample (nonisotropic.py:152-159) and Mahalanobis l1 loss (nonisotropic.py:180-190, base.py:298)
// per-sample t: thread = (sample, node, 4 channels); the [T,N,N] table is read through L1/L2
// =============================================================================================
__global__ void __launch_bounds__(256)
q_sample_kernel(const float* __restrict__ x0, const float* __restrict__ eps, const int* __restrict__ t,
                const float* __restrict__ sqrt_ac, const float* __restrict__ M, float* __restrict__ out,
                int B, int N, int D) {
    const int d4 = D >> 2;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)B * N * d4) return;
    const int d = (int)(gid % d4) * 4;
    const long long bn = gid / d4;
    const int n = (int)(bn % N), b = (int)(bn / N);
    const int tb = __ldg(t + b);
    const float* Mt = M + ((long long)tb * N + n) * N;
    const long long base = (long long)b * N * D + d;
    const float a = __ldg(sqrt_ac + tb);
    const float4 xs = __ldg(reinterpret_cast<const float4*>(x0 + base + (long long)n * D));
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = 0; k < N; ++k) {
        const float4 e = __ldg(reinterpret_cast<const float4*>(eps + base + (long long)k * D));
        const float m = __ldg(Mt + k);
        acc.x = fmaf(m, e.x, acc.x); acc.y = fmaf(m, e.y, acc.y); acc.z = fmaf(m, e.z, acc.z); acc.w = fmaf(m, e.w, acc.w);
    }
    *reinterpret_cast<float4*>(out + base + (long long)n * D) =
        make_float4(a * xs.x + acc.x, a * xs.y + acc.y, a * xs.z + acc.z, a * xs.w + acc.w);
}

// one block per sample; loss[b] = mean_{n,d} | sum_k S[t_b][n,k] (out - x0)[b,k,d] |
__global__ void __launch_bounds__(256)
mahalanobis_loss_kernel(const float* __restrict__ out, const float* __restrict__ x0, const int* __restrict__ t,
                        const float* __restrict__ S, float* __restrict__ loss, int N, int D) {
    const int b = blockIdx.x;
    const int tb = __ldg(t + b);
    const float* St = S + (long long)tb * N * N;
    const long long base = (long long)b * N * D;
    float part = 0.0f;
    for (int i = threadIdx.x; i < N * D; i += blockDim.x) {
        const int n = i / D, d = i % D;
        float acc = 0.0f;
        for (int k = 0; k < N; ++k)
            acc = fmaf(__ldg(St + n * N + k), __ldg(out + base + k * D + d) - __ldg(x0 + base + k * D + d), acc);
        part += fabsf(acc);
    }
    __shared__ float red[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.0f;
        for (int w = 0; w < (blockDim.x >> 5); ++w) s += red[w];
        loss[b] = s / (float)(N * D);
    }
}

// =============================================================================================
// GRU gates (recurrent.py:351-358): r = sig(xr_r+hr_r); z = sig(xr_z+hr_z); n = tanh(xr_n + r*hr_n)
//                                   h' = n - n*z + z*h            (clock mask == 1, clockwork=False)
// =============================================================================================
// FAST: sigmoid / tanh through MUFU.EX2 + MUFU.RCP (~1e-7 absolute, as in the per-sample kernels of sd_mix.cu) instead of libdevice
template <bool FAST>
__global__ void __launch_bounds__(256)
gru_gates_kernel(const View xr, const float* __restrict__ xr_bias, const float* __restrict__ hr,
                 const float* __restrict__ hr_bias, const View h_in, const ViewW h_out, int B, int N, int H) {
    const int h4 = H >> 2;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)B * N * h4) return;
    const int j = (int)(gid % h4) * 4;
    const long long bn = gid / h4;
    const int n = (int)(bn % N), b = (int)(bn / N);
    const float* xrow = row_ptr(xr, b, n);
    const float* hrow = hr + bn * 3 * H;
    float4 g[6];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        g[c] = __ldg(reinterpret_cast<const float4*>(xrow + c * H + j));
        g[3 + c] = __ldg(reinterpret_cast<const float4*>(hrow + c * H + j));
        if (xr_bias) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(xr_bias + (long long)n * 3 * H + c * H + j));
            g[c].x += t.x; g[c].y += t.y; g[c].z += t.z; g[c].w += t.w;
        }
        if (hr_bias) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(hr_bias + (long long)n * 3 * H + c * H + j));
            g[3 + c].x += t.x; g[3 + c].y += t.y; g[3 + c].z += t.z; g[3 + c].w += t.w;
        }
    }
    const float4 hx = __ldg(reinterpret_cast<const float4*>(row_ptr(h_in, b, n) + j));
    auto sig = [](float v) { return FAST ? __fdividef(1.0f, 1.0f + exp2f(-1.4426950408889634f * v)) : 1.0f / (1.0f + expf(-v)); };
    auto th = [](float v) {
        if (!FAST) return tanhf(v);
        const float t = exp2f(-2.8853900817779268f * fabsf(v));
        return copysignf(__fdividef(1.0f - t, 1.0f + t), v);
    };
    auto cell = [&](float ir, float iz, float in_, float hr_, float hz, float hn, float hprev) {
        const float r = sig(ir + hr_), z = sig(iz + hz);
        const float nn = th(in_ + r * hn);
        return nn - nn * z + z * hprev;
    };
    float4 o;
    o.x = cell(g[0].x, g[1].x, g[2].x, g[3].x, g[4].x, g[5].x, hx.x);
    o.y = cell(g[0].y, g[1].y, g[2].y, g[3].y, g[4].y, g[5].y, hx.y);
    o.z = cell(g[0].z, g[1].z, g[2].z, g[3].z, g[4].z, g[5].z, hx.z);
    o.w = cell(g[0].w, g[1].w, g[2].w, g[3].w, g[4].w, g[5].w, hx.w);
    *reinterpret_cast<float4*>(row_ptr(h_out, b, n) + j) = o;
}

int gru_gates_fp32(const View& xr, const float* xr_bias, const float* hr, const float* hr_bias,
                   const View& h_in, const ViewW& h_out, int B, int N, int H, cudaStream_t st, bool fast) {
    if (B <= 0) return SD_OK;
    if (H % 4 != 0) { set_error("gru: hidden %d must be a multiple of 4", H); return SD_ERR_UNSUPPORTED; }
    const long long total = (long long)B * N * (H >> 2);
    if (fast) gru_gates_kernel<true><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(xr, xr_bias, hr, hr_bias, h_in, h_out, B, N, H);
    else gru_gates_kernel<false><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(xr, xr_bias, hr, hr_bias, h_in, h_out, B, N, H);
    SD_LAUNCH_OK("gru_gates_kernel");
    return SD_OK;
}

// =============================================================================================
// decoder output head with identity graph influence: y[b,n,:] = act(W_fc[type(n)] h[b,n,:] + bias[n])  (decoder.py:97-98)
// F <= 4 outputs per node, one warp per (sample, node) row, shuffle reduction.
// =============================================================================================
__global__ void __launch_bounds__(256)
gru_out_fc_kernel(const float* __restrict__ Wfc, const float* __restrict__ bias_node, const NodeTypes types, int N, int H, int F,
                  const float* __restrict__ h, const ViewW out, int act, long long rows) {
    const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= rows) return;
    const int n = (int)(r % N), b = (int)(r / N);
    const float* hr = h + r * H;
    const float* w = Wfc + (long long)types.t[n] * F * H;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int u = lane; u < H; u += 32) {
        const float hv = __ldg(hr + u);
#pragma unroll
        for (int c = 0; c < 4; ++c) if (c < F) acc[c] = fmaf(__ldg(w + c * H + u), hv, acc[c]);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
    if (lane < F) {
        float v = acc[0];
        if (lane == 1) v = acc[1]; else if (lane == 2) v = acc[2]; else if (lane == 3) v = acc[3];
        if (bias_node) v += __ldg(bias_node + (long long)n * F + lane);
        if (act == SD_ACT_TANH) v = tanhf(v);
        row_ptr(out, b, n)[lane] = v;
    }
}

int gru_out_fc_fp32(const float* Wfc, const float* bias_node, const NodeTypes& types, int N, int H, int F, const float* h,
                    const ViewW& out, int act, int B, cudaStream_t st) {
    if (B <= 0) return SD_OK;
    if (F > 4) { set_error("gru_out_fc: feature size %d > 4", F); return SD_ERR_UNSUPPORTED; }
    const long long rows = (long long)B * N;
    gru_out_fc_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(Wfc, bias_node, types, N, H, F, h, out, act, rows);
    SD_LAUNCH_OK("gru_out_fc_kernel");
    return SD_OK;
}

// =============================================================================================
// time conditioning table: sinusoidal embedding -> Linear -> GELU(erf) -> Linear -> per block
// Tanh -> Linear   (nn/generator.py:47-55; layers/attention.py:81-84).  Batch-invariant: one row per
// distinct time value, computed once per plan.
// =============================================================================================
__global__ void sinusoidal_kernel(const float* __restrict__ times, float* __restrict__ emb, int rows, int C, float theta) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int half = C / 2;
    if (i >= rows * half) return;
    const int r = i / half, c = i % half;
    const float step = -(float)(log((double)theta) / (double)(half - 1));
    const float freq = expf((float)c * step);
    const float ang = times[r] * freq;
    emb[(long long)r * C + c] = sinf(ang);
    emb[(long long)r * C + half + c] = cosf(ang);
}

// y[r,o] = b[o] + sum_i act(x[r,i]) * W[o,i]; one warp per output
__global__ void dense_rows_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ W,
                                  const float* __restrict__ bias, float* __restrict__ y, long long ldy,
                                  int rows, int in, int out, int pre_act) {
    const long long wid = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (wid >= (long long)rows * out) return;
    const int r = (int)(wid / out), o = (int)(wid % out);
    const float* xr = x + r * ldx;
    const float* wr = W + (long long)o * in;
    float s = 0.0f;
    for (int i = lane; i < in; i += 32) {
        float v = __ldg(xr + i);
        if (pre_act == 1) v = tanhf(v);
        else if (pre_act == 2) v = 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
        s = fmaf(v, __ldg(wr + i), s);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) y[r * ldy + o] = s + (bias ? __ldg(bias + o) : 0.0f);
}

static int dense_rows(const float* x, long long ldx, const float* W, const float* b, float* y, long long ldy,
                      int rows, int in, int out, int pre_act, cudaStream_t st) {
    const long long warps = (long long)rows * out;
    dense_rows_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, st>>>(x, ldx, W, b, y, ldy, rows, in, out, pre_act);
    SD_LAUNCH_OK("dense_rows_kernel");
    return SD_OK;
}

int time_table_fp32(const float* times, int rows, int C, float theta, int time_dim, const float* w1, const float* b1,
                    const float* w3, const float* b3, const float* const* head_w, const float* const* head_b,
                    int n_heads, float* table, float* ws, cudaStream_t st) {
    float* emb = ws;                                  // [rows][C]
    float* h1 = emb + (long long)rows * C;            // [rows][time_dim]
    float* h2 = h1 + (long long)rows * time_dim;      // [rows][time_dim]
    const int tot = rows * (C / 2);
    sinusoidal_kernel<<<(tot + 127) / 128, 128, 0, st>>>(times, emb, rows, C, theta);
    SD_LAUNCH_OK("sinusoidal_kernel");
    int rc = dense_rows(emb, C, w1, b1, h1, time_dim, rows, C, time_dim, 0, st);
    if (rc) return rc;
    rc = dense_rows(h1, time_dim, w3, b3, h2, time_dim, rows, time_dim, time_dim, 2, st);
    if (rc) return rc;
    for (int h = 0; h < n_heads; ++h) {
        rc = dense_rows(h2, time_dim, head_w[h], head_b[h], table + (long long)h * 2 * C, (long long)n_heads * 2 * C,
                        rows, time_dim, 2 * C, 1, st);
        if (rc) return rc;
    }
    return SD_OK;
}

// =============================================================================================
// N(0,1) fill: Philox4x32-10 counter RNG + Box-Muller (replaces torch.randn at base.py:156-158).
// Element i uses counter (offset + i/4); results are independent of launch geometry and rank count.
// =============================================================================================
__device__ __forceinline__ void philox_round(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t k0, uint32_t k1) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
}

__global__ void __launch_bounds__(256)
fill_normal_kernel(float* __restrict__ out, long long count, uint64_t seed, uint64_t offset) {
    const long long quads = (count + 3) / 4;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < quads; q += (long long)gridDim.x * blockDim.x) {
        const uint64_t ctr = offset + (uint64_t)q;
        uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = 0x5EED5EEDu, c3 = 0u;
        uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            philox_round(c0, c1, c2, c3, k0, k1);
            k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
        }
        const float u0 = ((float)(c0 >> 8) + 0.5f) * (1.0f / 16777216.0f);
        const float u1 = ((float)(c1 >> 8) + 0.5f) * (1.0f / 16777216.0f);
        const float u2 = ((float)(c2 >> 8) + 0.5f) * (1.0f / 16777216.0f);
        const float u3 = ((float)(c3 >> 8) + 0.5f) * (1.0f / 16777216.0f);
        const float r0 = sqrtf(-2.0f * logf(u0)), r1 = sqrtf(-2.0f * logf(u2));
        float s0, c0f, s1, c1f;
        sincosf(6.28318530717958647692f * u1, &s0, &c0f);
        sincosf(6.28318530717958647692f * u3, &s1, &c1f);
        const float v[4] = {r0 * c0f, r0 * s0, r1 * c1f, r1 * s1};
        const long long base = q * 4;
        if (base + 3 < count && ((reinterpret_cast<uintptr_t>(out) & 15u) == 0)) {
            *reinterpret_cast<float4*>(out + base) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
            for (int i = 0; i < 4; ++i) if (base + i < count) out[base + i] = v[i];
        }
    }
}

int fill_normal(float* out, long long count, uint64_t seed, uint64_t offset, cudaStream_t st) {
    if (count <= 0) return SD_OK;
    long long blocks = ((count + 3) / 4 + 255) / 256;
    if (blocks > 148LL * 16) blocks = 148LL * 16;
    fill_normal_kernel<<<(unsigned)blocks, 256, 0, st>>>(out, count, seed, offset);
    SD_LAUNCH_OK("fill_normal_kernel");
    return SD_OK;
}

int q_sample_fp32(const float* x0, const float* eps, const int* t, const float* sqrt_ac, const float* M, float* out,
                  int B, int N, int D, cudaStream_t st) {
    if (B <= 0) return SD_OK;
    if (D % 4) { set_error("q_sample: latent_dim %% 4 != 0"); return SD_ERR_UNSUPPORTED; }
    const long long total = (long long)B * N * (D >> 2);
    q_sample_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(x0, eps, t, sqrt_ac, M, out, B, N, D);
    SD_LAUNCH_OK("q_sample_kernel");
    return SD_OK;
}

int mahalanobis_loss_fp32(const float* out, const float* x0, const int* t, const float* S, float* loss,
                          int B, int N, int D, cudaStream_t st) {
    if (B <= 0) return SD_OK;
    mahalanobis_loss_kernel<<<B, 256, 0, st>>>(out, x0, t, S, loss, N, D);
    SD_LAUNCH_OK("mahalanobis_loss_kernel");
    return SD_OK;
}

}  // namespace sd
