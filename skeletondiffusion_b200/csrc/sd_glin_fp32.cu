// StaticGraphLinear, fp32 FFMA path (the <=1e-4 parity gate).
//
//   Y[b,m,:]   = A[b,m,:] @ W[type(m)]^T                  (grouped GEMM: one group per node)
//   out[b,n,:] = epilogue( sum_m G^[n,m] * rs[b,m] * Y[b,m,:] )
//
// Reference: GraphLinear.forward / gmm, src/core/network/layers/graph_structural.py:30-43, 7-8.
// When G^ == I (host passes G == nullptr) the epilogue is fused into the GEMM tile store and Y is
// never written.  Otherwise the GEMM writes raw Y to scratch and node_mix_fp32 finishes the layer.
#include "sd_internal.h"

namespace sd {

struct GemmParams {
    View a0, a1;          // K segments (a1.ptr may be null)
    const float* W;       // [types][OUT][K]
    int K, OUT, N, B;
    NodeTypes types;
    const float* row_scale;
    Epilogue epi;
    ViewW out;
    int fused;            // 1: apply epilogue, 0: raw store
    int vecA, vecW, vecO;
};

constexpr int BM = 128, BN = 64, BK = 16, TM = 8, TN = 4;
constexpr int GEMM_THREADS = (BM / TM) * (BN / TN);   // 256

__device__ __forceinline__ float load_a_elem(const GemmParams& p, const float* r0, const float* r1, int kg) {
    if (kg < p.a0.width) return __ldg(r0 + kg);
    kg -= p.a0.width;
    if (r1 != nullptr && kg < p.a1.width) return __ldg(r1 + kg);
    return 0.0f;
}

__global__ void __launch_bounds__(GEMM_THREADS)
glin_gemm_fp32_kernel(const GemmParams p) {
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];

    const int tid = threadIdx.x;
    const int node = blockIdx.y;
    const int b0 = blockIdx.x * BM;
    const int o0 = blockIdx.z * BN;
    const float* Wt = p.W + (long long)p.types.t[node] * p.OUT * p.K;

    // A loader: thread -> (row, 8 consecutive k)
    const int a_row = tid >> 1, a_kq = (tid & 1) * 8;
    const int a_b = b0 + a_row;
    const bool a_ok = a_b < p.B;
    const float* ar0 = a_ok ? row_ptr(p.a0, a_b, node) : nullptr;
    const float* ar1 = (a_ok && p.a1.ptr) ? row_ptr(p.a1, a_b, node) : nullptr;
    // W loader: thread -> (out col, 4 consecutive k)
    const int w_c = tid >> 2, w_kq = (tid & 3) * 4;
    const int w_o = o0 + w_c;
    const bool w_ok = w_o < p.OUT;
    const float* wr = Wt + (long long)w_o * p.K;

    float a_reg[8], w_reg[4];
    auto fetch = [&](int kt) {
        const int ka = kt * BK + a_kq;
        if (!a_ok) {
#pragma unroll
            for (int i = 0; i < 8; ++i) a_reg[i] = 0.0f;
        } else if (p.vecA) {
            // both segments are multiples of 8 wide and 16-byte aligned: the 8 values sit in one segment
            const float* src = nullptr;
            if (ka < p.a0.width) src = ar0 + ka;
            else if (ar1 && ka - p.a0.width < p.a1.width) src = ar1 + (ka - p.a0.width);
            if (src) {
                const float4 v0 = __ldg(reinterpret_cast<const float4*>(src));
                const float4 v1 = __ldg(reinterpret_cast<const float4*>(src) + 1);
                a_reg[0] = v0.x; a_reg[1] = v0.y; a_reg[2] = v0.z; a_reg[3] = v0.w;
                a_reg[4] = v1.x; a_reg[5] = v1.y; a_reg[6] = v1.z; a_reg[7] = v1.w;
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) a_reg[i] = 0.0f;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) a_reg[i] = load_a_elem(p, ar0, ar1, ka + i);
        }
        const int kw = kt * BK + w_kq;
        if (!w_ok) {
            w_reg[0] = w_reg[1] = w_reg[2] = w_reg[3] = 0.0f;
        } else if (p.vecW && kw + 3 < p.K) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(wr + kw));
            w_reg[0] = v.x; w_reg[1] = v.y; w_reg[2] = v.z; w_reg[3] = v.w;
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) w_reg[i] = (kw + i < p.K) ? __ldg(wr + kw + i) : 0.0f;
        }
    };
    auto stash = [&]() {
#pragma unroll
        for (int i = 0; i < 8; ++i) As[a_kq + i][a_row] = a_reg[i];
#pragma unroll
        for (int i = 0; i < 4; ++i) Bs[w_kq + i][w_c] = w_reg[i];
    };

    const int tx = tid % (BN / TN), ty = tid / (BN / TN);
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;

    const int nk = (p.K + BK - 1) / BK;
    fetch(0);
    stash();
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        if (kt + 1 < nk) fetch(kt + 1);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a_lo = *reinterpret_cast<const float4*>(&As[kk][ty * TM]);
            const float4 a_hi = *reinterpret_cast<const float4*>(&As[kk][ty * TM + 4]);
            const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * TN]);
            const float a[TM] = {a_lo.x, a_lo.y, a_lo.z, a_lo.w, a_hi.x, a_hi.y, a_hi.z, a_hi.w};
            const float bb[TN] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
        __syncthreads();
        if (kt + 1 < nk) { stash(); __syncthreads(); }
    }

    // store
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int b = b0 + ty * TM + i;
        if (b >= p.B) continue;
        const int oc = o0 + tx * TN;
        if (oc >= p.OUT) continue;
        float v[TN];
        const float rs = (p.fused && p.row_scale) ? __ldg(p.row_scale + (long long)b * p.N + node) : 1.0f;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            v[j] = acc[i][j];
            if (p.fused && oc + j < p.OUT) v[j] = epilogue_apply(p.epi, b, node, oc + j, v[j] * rs);
        }
        float* dst = row_ptr(p.out, b, node) + oc;
        if (p.vecO && oc + TN <= p.OUT) {
            *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
#pragma unroll
            for (int j = 0; j < TN; ++j) if (oc + j < p.OUT) dst[j] = v[j];
        }
    }
}

// ------------------------------------------------------------------------------------------
// node mix: out[b,n,o4] = epilogue( sum_m G[n,m] * rs[b,m] * y[b,m,o4] ); thread = (sample, 4 channels)
// ------------------------------------------------------------------------------------------
struct MixParams {
    const float* G; const float* y; long long y_sb; const float* row_scale;
    Epilogue epi; ViewW out;
    int N, OUT, B, vec;
};
constexpr int MIX_NH = 32;   // accumulator rows per pass

template <int VEC>
__global__ void __launch_bounds__(128)
node_mix_fp32_kernel(const MixParams p) {
    extern __shared__ float Gs[];   // [N][N] transposed: Gs[m*N + n]
    for (int i = threadIdx.x; i < p.N * p.N; i += blockDim.x) {
        const int n = i / p.N, m = i % p.N;
        Gs[m * p.N + n] = __ldg(p.G + i);
    }
    __syncthreads();
    const int chunks = (p.OUT + VEC - 1) / VEC;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)p.B * chunks) return;
    const int b = (int)(gid / chunks);
    const int o = (int)(gid % chunks) * VEC;
    const float* yb = p.y + (long long)b * p.y_sb + o;
    for (int n0 = 0; n0 < p.N; n0 += MIX_NH) {
        float acc[MIX_NH][VEC];
#pragma unroll
        for (int i = 0; i < MIX_NH; ++i)
#pragma unroll
            for (int j = 0; j < VEC; ++j) acc[i][j] = 0.0f;
        for (int m = 0; m < p.N; ++m) {
            float yv[VEC];
            if (VEC == 4) {
                const float4 t = __ldg(reinterpret_cast<const float4*>(yb + (long long)m * p.OUT));
                yv[0] = t.x; yv[1] = t.y; yv[2] = t.z; yv[3] = t.w;
            } else {
#pragma unroll
                for (int j = 0; j < VEC; ++j) yv[j] = (o + j < p.OUT) ? __ldg(yb + (long long)m * p.OUT + j) : 0.0f;
            }
            if (p.row_scale) {
                const float rs = __ldg(p.row_scale + (long long)b * p.N + m);
#pragma unroll
                for (int j = 0; j < VEC; ++j) yv[j] *= rs;
            }
            const float* gcol = Gs + m * p.N + n0;
#pragma unroll
            for (int i = 0; i < MIX_NH; ++i) {
                if (n0 + i < p.N) {
                    const float g = gcol[i];
#pragma unroll
                    for (int j = 0; j < VEC; ++j) acc[i][j] = fmaf(g, yv[j], acc[i][j]);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < MIX_NH; ++i) {
            const int n = n0 + i;
            if (n < p.N) {
                float* dst = row_ptr(p.out, b, n) + o;
                float v[VEC];
#pragma unroll
                for (int j = 0; j < VEC; ++j) v[j] = (o + j < p.OUT) ? epilogue_apply(p.epi, b, n, o + j, acc[i][j]) : 0.0f;
                if (VEC == 4 && p.vec) {
                    *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
                } else {
#pragma unroll
                    for (int j = 0; j < VEC; ++j) if (o + j < p.OUT) dst[j] = v[j];
                }
            }
        }
    }
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline bool view_vec_ok(const View& v, int mult) {
    return v.ptr == nullptr || (aligned16(v.ptr) && v.sb % 4 == 0 && v.sn % 4 == 0 && v.width % mult == 0);
}

int node_mix_fp32(const float* G, int N, int OUT, const float* y, long long y_sb, const float* row_scale,
                  const Epilogue& epi, const ViewW& out, int B, cudaStream_t st) {
    MixParams mp;
    mp.G = G; mp.y = y; mp.y_sb = y_sb; mp.row_scale = row_scale; mp.epi = epi; mp.out = out;
    mp.N = N; mp.OUT = OUT; mp.B = B;
    const bool vec = (OUT % 4 == 0) && (y_sb % 4 == 0) && aligned16(y) && aligned16(out.ptr) && out.sb % 4 == 0 && out.sn % 4 == 0;
    mp.vec = vec ? 1 : 0;
    const size_t smem = sizeof(float) * N * N;
    if (vec) {
        const long long total = (long long)B * (OUT / 4);
        node_mix_fp32_kernel<4><<<(unsigned)((total + 127) / 128), 128, smem, st>>>(mp);
    } else {
        const long long total = (long long)B * OUT;
        node_mix_fp32_kernel<1><<<(unsigned)((total + 127) / 128), 128, smem, st>>>(mp);
    }
    SD_LAUNCH_OK("node_mix_fp32_kernel");
    return SD_OK;
}

int glin_forward_fp32(const float* W, const float* Wt, int K, int OUT, const NodeTypes& types, int N,
                      const float* G, const GlinCall& c, cudaStream_t st) {
    if (c.B <= 0) return SD_OK;
    const int kin = c.a0.width + (c.a1.ptr ? c.a1.width : 0);
    if (kin != K) { set_error("glin: input width %d != in_features %d", kin, K); return SD_ERR_INVALID; }
    if (G != nullptr && c.scratch == nullptr) { set_error("glin: scratch required for non-identity G"); return SD_ERR_INVALID; }
    {   // FFMA2 kernel for the regular shapes (sd_glin_ffma2.cu); the generic kernel below handles ragged K / OUT
        const bool fused = (G == nullptr);
        const ViewW dst = fused ? c.out : contiguous_view_w(c.scratch, N, OUT);
        if (glin_f2_supported(c.a0, c.a1, K, OUT, Wt, dst) && c.out.rep == 1) {
            int rc = glin_f2_launch(Wt, K, OUT, types, N, c, dst, fused, st);
            if (rc || fused) return rc;
            Epilogue e = c.epi; e.OUT = OUT;
            return node_mix_fp32(G, N, OUT, c.scratch, (long long)N * OUT, c.row_scale, e, c.out, c.B, st);
        }
    }
    GemmParams p;
    p.a0 = c.a0; p.a1 = c.a1; p.W = W; p.K = K; p.OUT = OUT; p.N = N; p.B = c.B; p.types = types;
    p.row_scale = c.row_scale; p.epi = c.epi; p.epi.OUT = OUT;
    p.fused = (G == nullptr) ? 1 : 0;
    p.out = p.fused ? c.out : contiguous_view_w(c.scratch, N, OUT);
    p.vecA = (view_vec_ok(c.a0, 8) && view_vec_ok(c.a1, 8)) ? 1 : 0;
    p.vecW = (K % 4 == 0 && aligned16(W)) ? 1 : 0;
    p.vecO = (OUT % 4 == 0 && aligned16(p.out.ptr) && p.out.sb % 4 == 0 && p.out.sn % 4 == 0) ? 1 : 0;
    dim3 grid((c.B + BM - 1) / BM, N, (OUT + BN - 1) / BN);
    glin_gemm_fp32_kernel<<<grid, GEMM_THREADS, 0, st>>>(p);
    SD_LAUNCH_OK("glin_gemm_fp32_kernel");
    if (!p.fused) {
        Epilogue e = c.epi; e.OUT = OUT;
        return node_mix_fp32(G, N, OUT, c.scratch, (long long)N * OUT, c.row_scale, e, c.out, c.B, st);
    }
    return SD_OK;
}

}  // namespace sd
