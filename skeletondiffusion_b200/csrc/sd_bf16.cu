// bf16 activation helpers of the tensor-core path: fp32 -> bf16 cast with in-place concat / repeat,
// RMSNorm row factor on bf16 rows, node mix of fp32 raw products with bf16 residual/output.
#include "sd_internal.h"

namespace sd {

// out[b, n, :] = bf16( cat(a0[b/rep0, n, :], a1[b/rep1, n, :]) );  thread = (row, 4 channels)
__global__ void __launch_bounds__(256)
cast_concat_bf16_kernel(const View a0, const View a1, __nv_bfloat16* __restrict__ out, int B, int N, int K) {
    const int k4 = K >> 2;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)B * N * k4) return;
    const int k = (int)(gid % k4) * 4;
    const long long bn = gid / k4;
    const int n = (int)(bn % N), b = (int)(bn / N);
    const float* src = (k < a0.width) ? row_ptr(a0, b, n) + k : row_ptr(a1, b, n) + (k - a0.width);
    const float4 v = __ldg(reinterpret_cast<const float4*>(src));
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 t; t.x = *reinterpret_cast<uint32_t*>(&lo); t.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(out + bn * K + k) = t;
}

int cast_concat_bf16(const View& a0, const View& a1, __nv_bfloat16* out, int B, int N, cudaStream_t st) {
    if (B <= 0) return SD_OK;
    const int K = a0.width + (a1.ptr ? a1.width : 0);
    auto ok = [](const View& v) { return v.ptr == nullptr || ((reinterpret_cast<uintptr_t>(v.ptr) & 15u) == 0 && v.sb % 4 == 0 && v.sn % 4 == 0 && v.width % 4 == 0); };
    if (!ok(a0) || !ok(a1)) { set_error("cast_concat_bf16: operands must be 16-byte aligned with widths %% 4 == 0"); return SD_ERR_UNSUPPORTED; }
    const long long total = (long long)B * N * (K >> 2);
    cast_concat_bf16_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(a0, a1, out, B, N, K);
    SD_LAUNCH_OK("cast_concat_bf16_kernel");
    return SD_OK;
}

__global__ void row_inv_norm_bf16_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ inv, long long rows, int width) {
    const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= rows) return;
    const uint32_t* xr = reinterpret_cast<const uint32_t*>(x + r * width);
    float s = 0.0f;
    for (int i = lane; i < (width >> 1); i += 32) {
        const uint32_t w = __ldg(xr + i);
        const float a = __uint_as_float(w << 16), b = __uint_as_float(w & 0xFFFF0000u);
        s = fmaf(a, a, s); s = fmaf(b, b, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) inv[r] = 1.0f / fmaxf(sqrtf(s), 1e-12f);
}

int row_inv_norm_bf16(const __nv_bfloat16* x, float* inv, long long rows, int width, cudaStream_t st) {
    if (rows <= 0) return SD_OK;
    if (width & 1) { set_error("row_inv_norm_bf16: odd width"); return SD_ERR_UNSUPPORTED; }
    row_inv_norm_bf16_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(x, inv, rows, width);
    SD_LAUNCH_OK("row_inv_norm_bf16_kernel");
    return SD_OK;
}

// out[b,n,o4] = epilogue( sum_m G[n,m] y[b,m,o4] ) (+ bf16 residual), written as bf16 or fp32 (contiguous [B,N,OUT])
struct Mix16Params {
    const float* G; const float* y; Epilogue epi; const __nv_bfloat16* res; void* out; int out_fp32; int N, OUT, B;
};
constexpr int MIX16_NH = 32;

__global__ void __launch_bounds__(128)
node_mix16_kernel(const Mix16Params p) {
    extern __shared__ float Gs[];   // transposed: Gs[m*N + n]
    for (int i = threadIdx.x; i < p.N * p.N; i += blockDim.x) Gs[(i % p.N) * p.N + (i / p.N)] = __ldg(p.G + i);
    __syncthreads();
    const int chunks = p.OUT >> 2;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)p.B * chunks) return;
    const int b = (int)(gid / chunks), o = (int)(gid % chunks) * 4;
    const float* yb = p.y + (long long)b * p.N * p.OUT + o;
    for (int n0 = 0; n0 < p.N; n0 += MIX16_NH) {
        float acc[MIX16_NH][4];
#pragma unroll
        for (int i = 0; i < MIX16_NH; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.0f; }
        for (int m = 0; m < p.N; ++m) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(yb + (long long)m * p.OUT));
            const float* gcol = Gs + m * p.N + n0;
#pragma unroll
            for (int i = 0; i < MIX16_NH; ++i) {
                if (n0 + i < p.N) {
                    const float g = gcol[i];
                    acc[i][0] = fmaf(g, t.x, acc[i][0]); acc[i][1] = fmaf(g, t.y, acc[i][1]);
                    acc[i][2] = fmaf(g, t.z, acc[i][2]); acc[i][3] = fmaf(g, t.w, acc[i][3]);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < MIX16_NH; ++i) {
            const int n = n0 + i;
            if (n < p.N) {
                float v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = epilogue_apply(p.epi, b, n, o + j, acc[i][j]);
                const long long off = ((long long)b * p.N + n) * p.OUT + o;
                if (p.res) {
                    const uint2 r = __ldg(reinterpret_cast<const uint2*>(p.res + off));
                    v[0] += __uint_as_float(r.x << 16); v[1] += __uint_as_float(r.x & 0xFFFF0000u);
                    v[2] += __uint_as_float(r.y << 16); v[3] += __uint_as_float(r.y & 0xFFFF0000u);
                }
                if (p.out_fp32) {
                    *reinterpret_cast<float4*>(static_cast<float*>(p.out) + off) = make_float4(v[0], v[1], v[2], v[3]);
                } else {
                    __nv_bfloat162 lo = __floats2bfloat162_rn(v[0], v[1]), hi = __floats2bfloat162_rn(v[2], v[3]);
                    uint2 t; t.x = *reinterpret_cast<uint32_t*>(&lo); t.y = *reinterpret_cast<uint32_t*>(&hi);
                    *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(p.out) + off) = t;
                }
            }
        }
    }
}

int node_mix_to_bf16(const float* G, int N, int OUT, const float* y, const Epilogue& epi, const __nv_bfloat16* res,
                     void* out, int out_fp32, int B, cudaStream_t st) {
    if (B <= 0) return SD_OK;
    if (OUT % 4) { set_error("node_mix_to_bf16: OUT %% 4 != 0"); return SD_ERR_UNSUPPORTED; }
    Mix16Params p;
    p.G = G; p.y = y; p.epi = epi; p.epi.OUT = OUT; p.res = res; p.out = out; p.out_fp32 = out_fp32; p.N = N; p.OUT = OUT; p.B = B;
    const long long total = (long long)B * (OUT >> 2);
    node_mix16_kernel<<<(unsigned)((total + 127) / 128), 128, sizeof(float) * N * N, st>>>(p);
    SD_LAUNCH_OK("node_mix16_kernel");
    return SD_OK;
}

__global__ void __launch_bounds__(256)
add_residual_kernel(float* __restrict__ out, const View res, int B, int N, int OUT, long long out_sb, long long out_sn) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)B * N * OUT) return;
    const int o = (int)(gid % OUT);
    const long long bn = gid / OUT;
    const int n = (int)(bn % N), b = (int)(bn / N);
    out[(long long)b * out_sb + (long long)n * out_sn + o] += __ldg(row_ptr(res, b, n) + o);
}

int add_residual_fp32(float* out, const View& res, int B, int N, int OUT, long long out_sb, long long out_sn, cudaStream_t st) {
    if (B <= 0) return SD_OK;
    const long long total = (long long)B * N * OUT;
    add_residual_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(out, res, B, N, OUT, out_sb, out_sn);
    SD_LAUNCH_OK("add_residual_kernel");
    return SD_OK;
}

}  // namespace sd
