// One fused kernel per recurrent step of the graph GRU with identity graph influence (the decoder's 120-step loop,
// nn/decoder.py:91-100, and the encoder layers): for a tile of 128 (sample, node) rows of one node
//     hr  = h_{i-1} @ W_hh[type]^T                  (fp32, FFMA2; all 3H gate columns in this CTA)
//     h_i = GRUCell(xr + b_ih, hr + b_hh, h_{i-1})   (recurrent.py:351-358)
//     y_i = tanh(W_fc[type] h_i + b_fc)              (decoder output head, nn/decoder.py:97-98; optional)
// h_{i-1} (48 KB) is staged once and stays in shared memory for the three 96-column gate blocks; the K-major,
// gate-interleaved W_hh tiles stream through a 3-stage cp.async ring.  Replaces three launches per step
// (GEMM, gates, output-head GEMM) and the [B, N, 3H] round trip of hr through HBM.
#include "sd_internal.h"
#include <cuda_pipeline_primitives.h>

namespace sd {

constexpr int G2_BM = 128, G2_BN = 96, G2_BK = 32, G2_THREADS = 256, G2_STAGES = 3;
constexpr int G2_B_BYTES = G2_BK * G2_BN * 4;

__device__ __forceinline__ void g2_ffma2(float2& d, float a, float2 b) {
    unsigned long long dd = *reinterpret_cast<unsigned long long*>(&d), bb = *reinterpret_cast<unsigned long long*>(&b), aa;
    asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(aa), "l"(bb));
    d = *reinterpret_cast<float2*>(&dd);
}

struct G2Params {
    const float* Wt;           // [types][H][3H] K-major, every 96-column block = gates r|z|n of 32 units
    const float* bias_x;       // [N][3H] same column order
    const float* bias_h;
    View xr;                   // [.., 3H] x-side product, same column order
    View h_prev; ViewW h_out;  // [.., H]
    const float* Wfc;          // [types][F][H] or null (no output head)
    const float* bias_fc;      // [N][F] or null
    ViewW y;                   // [.., F]
    NodeTypes types;
    int H, N, B, F;
};

template <int KT>   // KT = H / 32 k-tiles (H = 96 -> 3)
__global__ void __launch_bounds__(G2_THREADS, 2)
gru_step_fused_kernel(const G2Params p) {
    extern __shared__ __align__(128) uint8_t g2_smem[];
    uint8_t* As = g2_smem;                                   // [KT][128 rows][32 k] fp32, 16-byte chunks XOR-swizzled by row
    uint8_t* Bring = g2_smem + KT * G2_BM * 128;             // [stage][32 k][96 cols]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int node = blockIdx.y, b0 = blockIdx.x * G2_BM;
    const int type = p.types.t[node];
    const float* Wt = p.Wt + (long long)type * p.H * 3 * p.H;
    constexpr int NZ = KT;                                   // 3H / 96 gate blocks == H / 32
    constexpr int NQ = NZ * KT;                              // W tiles in (z, kt) order

    {   // stage h_{i-1}: chunk f = tid + 256 i -> (row = f % 128, c4 = f / 128), c4 in [0, 8 KT)
        const int row = tid & 127;
        const bool ok = b0 + row < p.B;
        const float* src = ok ? p.h_prev.ptr + (long long)(b0 + row) * p.h_prev.sb + (long long)node * p.h_prev.sn : nullptr;
#pragma unroll
        for (int i = 0; i < 4 * KT; ++i) {
            const int c = (tid >> 7) + 2 * i, kt = c >> 3, c4 = c & 7;
            void* dst = As + kt * (G2_BM * 128) + row * 128 + ((c4 ^ (row & 7)) << 4);
            if (ok) __pipeline_memcpy_async(dst, src + c * 4, 16);
            else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    auto issue_w = [&](int q) {
        if (q < NQ) {
            const int z = q / KT, kt = q % KT;
            uint8_t* Bs = Bring + (q % G2_STAGES) * G2_B_BYTES;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const int f = tid + 256 * i, k = f / 24, c4 = f % 24;
                __pipeline_memcpy_async(Bs + k * (G2_BN * 4) + c4 * 16, Wt + (long long)(kt * G2_BK + k) * 3 * p.H + z * G2_BN + c4 * 4, 16);
            }
        }
        __pipeline_commit();
    };
    issue_w(0);          // group 0 also carries the h_{i-1} copies issued above
    issue_w(1);

    float yacc[8][3];    // output-head partial sums of this thread's units (F <= 3)
#pragma unroll
    for (int i = 0; i < 8; ++i) yacc[i][0] = yacc[i][1] = yacc[i][2] = 0.0f;
    auto cell = [](float ir, float iz, float in_, float hr, float hz, float hn, float hprev) {
        const float r = 1.0f / (1.0f + expf(-(ir + hr)));
        const float z = 1.0f / (1.0f + expf(-(iz + hz)));
        const float n = tanhf(in_ + r * hn);
        return n - n * z + z * hprev;
    };

    for (int z = 0; z < NZ; ++z) {
        float2 acc[8][3];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) acc[i][j] = make_float2(0.f, 0.f);
        // x-side gates / biases of this block are requested before the products (consumed in the gate epilogue)
        const int o0 = z * G2_BN, u0 = z * 32 + 2 * tx;
        float2 bxg[3], bhg[3];
#pragma unroll
        for (int g = 0; g < 3; ++g) {
            bxg[g] = __ldg(reinterpret_cast<const float2*>(p.bias_x + (long long)node * 3 * p.H + o0 + 32 * g + 2 * tx));
            bhg[g] = __ldg(reinterpret_cast<const float2*>(p.bias_h + (long long)node * 3 * p.H + o0 + 32 * g + 2 * tx));
        }
        for (int kt = 0; kt < KT; ++kt) {
            const int q = z * KT + kt;
            issue_w(q + 2);
            __pipeline_wait_prior(2);
            __syncthreads();
            const uint8_t* At = As + kt * (G2_BM * 128);
            const float* Bs = reinterpret_cast<const float*>(Bring + (q % G2_STAGES) * G2_B_BYTES);
#pragma unroll
            for (int k4 = 0; k4 < G2_BK / 4; ++k4) {
                float4 a[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r = ty * 8 + i;
                    a[i] = *reinterpret_cast<const float4*>(At + r * 128 + ((k4 ^ (r & 7)) << 4));
                }
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    float2 b[3];
#pragma unroll
                    for (int j = 0; j < 3; ++j) b[j] = *reinterpret_cast<const float2*>(Bs + (k4 * 4 + kk) * G2_BN + 2 * tx + 32 * j);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float av = kk == 0 ? a[i].x : (kk == 1 ? a[i].y : (kk == 2 ? a[i].z : a[i].w));
#pragma unroll
                        for (int j = 0; j < 3; ++j) g2_ffma2(acc[i][j], av, b[j]);
                    }
                }
            }
            __syncthreads();
        }
        // gate epilogue for units u0, u0+1 of rows ty*8 .. +7 (loads of a 4-row group issued before use)
        float wfc[3][2];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            wfc[c][0] = (p.Wfc && c < p.F) ? __ldg(p.Wfc + ((long long)type * p.F + c) * p.H + u0) : 0.0f;
            wfc[c][1] = (p.Wfc && c < p.F) ? __ldg(p.Wfc + ((long long)type * p.F + c) * p.H + u0 + 1) : 0.0f;
        }
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float2 xg[4][3];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int b = b0 + ty * 8 + half * 4 + i;
                const bool ok = b < p.B;
                const float* xrow = ok ? p.xr.ptr + (long long)b * p.xr.sb + (long long)node * p.xr.sn + o0 + 2 * tx : nullptr;   // rep == 1 (host check): no division
#pragma unroll
                for (int g = 0; g < 3; ++g) xg[i][g] = ok ? __ldg(reinterpret_cast<const float2*>(xrow + 32 * g)) : make_float2(0.f, 0.f);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int ri = half * 4 + i, r = ty * 8 + ri, b = b0 + r;
                if (b >= p.B) continue;
                // previous state of these two units from the staged tile (k-tile z, chunk (2 tx) / 4)
                const int c4 = tx >> 1;
                const float* hrow = reinterpret_cast<const float*>(As + z * (G2_BM * 128) + r * 128 + ((c4 ^ (r & 7)) << 4));
                const float2 hp = *reinterpret_cast<const float2*>(hrow + (tx & 1) * 2);
                float2 hy;
                hy.x = cell(xg[i][0].x + bxg[0].x, xg[i][1].x + bxg[1].x, xg[i][2].x + bxg[2].x, acc[ri][0].x + bhg[0].x,
                            acc[ri][1].x + bhg[1].x, acc[ri][2].x + bhg[2].x, hp.x);
                hy.y = cell(xg[i][0].y + bxg[0].y, xg[i][1].y + bxg[1].y, xg[i][2].y + bxg[2].y, acc[ri][0].y + bhg[0].y,
                            acc[ri][1].y + bhg[1].y, acc[ri][2].y + bhg[2].y, hp.y);
                *reinterpret_cast<float2*>(p.h_out.ptr + (long long)b * p.h_out.sb + (long long)node * p.h_out.sn + u0) = hy;
#pragma unroll
                for (int c = 0; c < 3; ++c) yacc[ri][c] = fmaf(wfc[c][0], hy.x, fmaf(wfc[c][1], hy.y, yacc[ri][c]));
            }
        }
    }
    if (p.Wfc) {
        // reduce the output-head partial sums over the 16 threads (tx) that share a row
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) yacc[i][c] += __shfl_xor_sync(0xffffffffu, yacc[i][c], o);
        if (tx < p.F) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int b = b0 + ty * 8 + i;
                if (b >= p.B) continue;
                float v = tx == 0 ? yacc[i][0] : (tx == 1 ? yacc[i][1] : yacc[i][2]);
                if (p.bias_fc) v += __ldg(p.bias_fc + (long long)node * p.F + tx);
                (p.y.ptr + (long long)b * p.y.sb + (long long)node * p.y.sn)[tx] = tanhf(v);
            }
        }
    }
}

int gru_step_fused(const float* W_hh_perm_t, int H, const NodeTypes& types, int N, const View& xr, const float* bias_x, const float* bias_h,
                   const View& h_prev, const ViewW& h_out, const float* Wfc, const float* bias_fc, const ViewW* y, int F, int B, cudaStream_t st) {
    if (B <= 0) return SD_OK;
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
    if (H != 96 || !al16(W_hh_perm_t) || !al16(h_prev.ptr) || h_prev.sb % 4 || h_prev.sn % 4 || h_prev.rep != 1 || xr.rep != 1 || h_out.rep != 1 || (y && y->rep != 1) || (Wfc && (F > 3 || !y))) {
        set_error("gru_step_fused: unsupported configuration (H=%d F=%d)", H, F);
        return SD_ERR_UNSUPPORTED;
    }
    G2Params p;
    p.Wt = W_hh_perm_t; p.bias_x = bias_x; p.bias_h = bias_h; p.xr = xr; p.h_prev = h_prev; p.h_out = h_out;
    p.Wfc = Wfc; p.bias_fc = bias_fc; p.types = types; p.H = H; p.N = N; p.B = B; p.F = F;
    if (y) p.y = *y; else { p.y.ptr = nullptr; p.y.sb = p.y.sn = 0; p.y.rep = 1; p.y.width = 0; }
    constexpr int KT = 3;
    const int smem = KT * G2_BM * 128 + G2_STAGES * G2_B_BYTES;
    auto kern = gru_step_fused_kernel<KT>;
    static unsigned long long configured = 0;      // bit d: attribute set on device d (it is per device)
    if (int rc_attr = opt_in_smem(kern, (size_t)(smem), configured)) return rc_attr;
    dim3 grid((B + G2_BM - 1) / G2_BM, N, 1);
    kern<<<grid, G2_THREADS, smem, st>>>(p);
    SD_LAUNCH_OK("gru_step_fused_kernel");
    return SD_OK;
}

}  // namespace sd
