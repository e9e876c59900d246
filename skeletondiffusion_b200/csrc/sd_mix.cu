// Per-sample node mixing for a general (dense) graph-influence matrix G^.
//
//   out[b,n,:] = epilogue( sum_m G^[n,m] * rs[b,m] * Y[b,m,:] )         (graph_structural.py:30-43, gmm :7-8)
//
// The grouped GEMM kernels produce the raw per-node products Y[b,m,:] = x[b,m,:] W[type(m)]^T; the mix couples all
// nodes of a sample, and a sample's Y rows ([N, OUT] floats, 16 KB for the 192-wide AMASS layers) are ONE contiguous
// block in HBM.  A persistent CTA therefore streams whole samples through a shared-memory ring with one cp.async.bulk
// per sample (copy warp), and the compute warps work on (sample, 32-column) tasks: lane = column, the N inputs of the
// column in registers, N accumulators, N*N FFMAs whose G^ operand comes from the constant bank (the matrix is passed BY
// VALUE as a __grid_constant__ kernel parameter and the loops are fully unrolled, so every FFMA reads c[0x0][imm]: no
// load instruction and no register is spent on G^).  The layer epilogue (mixed bias, (scale+1)x+shift, tanh, residual)
// is applied to the accumulators and the result is written with coalesced 128-byte stores: Y is read once, the output
// written once.  Bound: HBM (Y + residual + out = 3 x 4 B per element); the FFMA work is N FMAs per element (21 for
// AMASS: 4.3 GFLOP per 192-wide layer at B = 25 600, 0.06 ms of the FP32 pipe).
#include "sd_internal.h"
#include "sd_tc.cuh"
#include "sd_mixmat.cuh"
#include <stdlib.h>

namespace sd {

constexpr int SMIX_WARPS = 11;                 // compute warps; + 1 copy warp = 12 warps (ptxas budgets registers per 4-warp group: 168 per thread)
constexpr int SMIX_THREADS = (SMIX_WARPS + 1) * 32;
constexpr int SMIX_MAX_STAGES = 8;


struct SmixParams {
    const float* y;            // [B][N][OUT] raw products, contiguous
    const float* row_scale;    // [B*N] or null: Y row (b, m) is scaled before the mix (RMSNorm factor of the to_qkv input)
    const float* bias_node;    // [N][OUT] (already mixed: G^ @ bias[type]) or null
    const float* ss;           // resolved scale/shift row (scale at [o], shift at [OUT + o]) or null
    int act;
    View residual;             // ptr null if none
    ViewW out;
    int B, OUT, stages;
};

struct __align__(8) SmixBarriers { uint64_t full[SMIX_MAX_STAGES], done[SMIX_MAX_STAGES]; };

__device__ __forceinline__ void smix_bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(tc::smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(tc::smem_u32(bar)) : "memory");
}

// tanh with ~1e-7 absolute error in 7 instructions: (1 - t) / (1 + t), t = exp(-2|x|) (MUFU.EX2 + MUFU.RCP, no branch).
// Its consumers are linear layers, so the ABSOLUTE error is what propagates; it equals the rounding of a value of magnitude 1.
__device__ __forceinline__ float smix_tanh(float x) {
    const float t = exp2f(-2.8853900817779268f * fabsf(x));
    return copysignf(__fdividef(1.0f - t, 1.0f + t), x);
}

// COLS: columns per lane (c and c + 32 of a 64-column task when OUT is a multiple of 64: each constant load feeds two FFMAs)
// FAST: tanh / sigmoid through MUFU.EX2 + MUFU.RCP (absolute error ~1e-7) instead of libdevice (SKELDIFF_ACCURATE_EPILOGUE=1 selects libdevice)
template <bool FAST> __device__ __forceinline__ float mix_tanh(float x) { return FAST ? smix_tanh(x) : tanhf(x); }
template <bool FAST> __device__ __forceinline__ float mix_sigmoid(float v) {
    return FAST ? __fdividef(1.0f, 1.0f + exp2f(-1.4426950408889634f * v)) : 1.0f / (1.0f + expf(-v));
}
static bool fast_epilogue() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SKELDIFF_ACCURATE_EPILOGUE"); v = (e && e[0] == '1') ? 0 : 1; }
    return v == 1;
}

template <int N, int ACT, bool HAS_RES, int COLS, bool FAST>
__global__ void __launch_bounds__(SMIX_THREADS, 1)
sample_mix_kernel(const __grid_constant__ MixMat<N> G, const SmixParams p) {
    extern __shared__ __align__(128) float smix_smem[];
    const int slab = N * p.OUT;                                       // floats per sample
    float* ring = smix_smem;                                          // [stages][slab]
    SmixBarriers* bars = reinterpret_cast<SmixBarriers*>(ring + (size_t)p.stages * slab);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int chunks = p.OUT / (32 * COLS);                           // tasks per sample
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { tc::mbar_init(&bars->full[s], 1); tc::mbar_init(&bars->done[s], (uint32_t)chunks); }
        tc::fence_barrier_init();
    }
    __syncthreads();
    const int my_samples = (int)blockIdx.x < p.B ? (p.B - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const uint32_t slab_bytes = (uint32_t)slab * 4u;
    if (warp == SMIX_WARPS) {
        // ------------------------------------------------------------ copy warp: one bulk load per sample
        if (lane == 0) {
            for (int k = 0; k < my_samples; ++k) {
                const int st = k % p.stages;
                if (k >= p.stages) tc::mbar_wait(&bars->done[st], (uint32_t)(k / p.stages - 1) & 1u);
                tc::mbar_arrive_expect_tx(&bars->full[st], slab_bytes);
                smix_bulk_load(ring + (size_t)st * slab, p.y + ((long long)blockIdx.x + (long long)k * gridDim.x) * slab, slab_bytes, &bars->full[st]);
            }
        }
        return;
    }
    // ---------------------------------------------------------------- compute warps: task = (sample k, 32 * COLS columns)
    const long long tasks = (long long)my_samples * chunks;
    for (long long task = warp; task < tasks; task += SMIX_WARPS) {
        const int k = (int)(task / chunks), c = (int)(task % chunks) * (32 * COLS) + lane;
        const int st = k % p.stages;
        const int b = (int)blockIdx.x + k * (int)gridDim.x;
        float res[N][COLS];
        if (HAS_RES) {                                                // the residual row segments are in flight during the mix
            const float* rb = p.residual.ptr + (long long)(p.residual.rep == 1 ? b : b / p.residual.rep) * p.residual.sb + c;
#pragma unroll
            for (int n = 0; n < N; ++n)
#pragma unroll
                for (int j = 0; j < COLS; ++j) res[n][j] = __ldg(rb + (long long)n * p.residual.sn + 32 * j);
        }
        float mul[COLS], add[COLS];
#pragma unroll
        for (int j = 0; j < COLS; ++j) {
            mul[j] = p.ss ? __ldg(p.ss + c + 32 * j) + 1.0f : 1.0f;
            add[j] = p.ss ? __ldg(p.ss + p.OUT + c + 32 * j) : 0.0f;
        }
        tc::mbar_wait(&bars->full[st], (uint32_t)(k / p.stages) & 1u);
        const float* ys = ring + (size_t)st * slab + c;
        float in[N][COLS];
#pragma unroll
        for (int m = 0; m < N; ++m)
#pragma unroll
            for (int j = 0; j < COLS; ++j) in[m][j] = ys[m * p.OUT + 32 * j];
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&bars->done[st]);             // this task's reads of the stage are complete
        if (p.row_scale) {
            const float* rs = p.row_scale + (long long)b * N;
#pragma unroll
            for (int m = 0; m < N; ++m) {
                const float r = __ldg(rs + m);
#pragma unroll
                for (int j = 0; j < COLS; ++j) in[m][j] *= r;
            }
        }
        float acc[N][COLS];
        mix_nodes<N, COLS>(G, in, acc);
        float* ob = p.out.ptr + (long long)b * p.out.sb + c;
#pragma unroll
        for (int n = 0; n < N; ++n)
#pragma unroll
            for (int j = 0; j < COLS; ++j) {
                float v = acc[n][j];
                if (p.bias_node) v += __ldg(p.bias_node + n * p.OUT + c + 32 * j);
                v = fmaf(v, mul[j], add[j]);
                if (ACT == SD_ACT_TANH) v = mix_tanh<FAST>(v);
                if (ACT == SD_ACT_TANH_TANH) v = mix_tanh<FAST>(mix_tanh<FAST>(v));
                if (HAS_RES) v += res[n][j];
                ob[(long long)n * p.out.sn + 32 * j] = v;
            }
    }
}

template <int N, int ACT, bool HAS_RES, int COLS, bool FAST>
static int smix_launch_c(const float* G_host, const SmixParams& p0, cudaStream_t st) {
    SmixParams p = p0;
    MixMat<N> G;
    G.set(G_host);
    const size_t slab_bytes = (size_t)N * p.OUT * 4;
    int stages = (int)((200 * 1024) / slab_bytes);
    if (stages > SMIX_MAX_STAGES) stages = SMIX_MAX_STAGES;
    if (stages < 2) { set_error("sample_mix: a sample's rows (%zu bytes) do not fit a two-stage ring", slab_bytes); return SD_ERR_UNSUPPORTED; }
    p.stages = stages;
    const size_t smem = (size_t)stages * slab_bytes + sizeof(SmixBarriers) + 128;
    auto kern = sample_mix_kernel<N, ACT, HAS_RES, COLS, FAST>;
    static unsigned long long configured = 0;
    if (int rc = opt_in_smem(kern, 227 * 1024, configured)) return rc;
    const int sms = sm_count();
    const int grid = p.B < sms ? p.B : sms;
    kern<<<grid, SMIX_THREADS, smem, st>>>(G, p);
    SD_LAUNCH_OK("sample_mix_kernel");
    return SD_OK;
}

template <int N, int ACT, bool HAS_RES>
static int smix_launch_t(const float* G_host, const SmixParams& p, cudaStream_t st) {
    constexpr bool kHasAct = ACT != SD_ACT_NONE;
    const bool fast = kHasAct && fast_epilogue();
    if (p.OUT % 64 == 0) return fast ? smix_launch_c<N, ACT, HAS_RES, 2, kHasAct>(G_host, p, st) : smix_launch_c<N, ACT, HAS_RES, 2, false>(G_host, p, st);
    return fast ? smix_launch_c<N, ACT, HAS_RES, 1, kHasAct>(G_host, p, st) : smix_launch_c<N, ACT, HAS_RES, 1, false>(G_host, p, st);
}

template <int N>
static int smix_launch_n(const float* G_host, const SmixParams& p, bool has_res, cudaStream_t st) {
    switch (p.act) {
    case SD_ACT_NONE: return has_res ? smix_launch_t<N, SD_ACT_NONE, true>(G_host, p, st) : smix_launch_t<N, SD_ACT_NONE, false>(G_host, p, st);
    case SD_ACT_TANH: return has_res ? smix_launch_t<N, SD_ACT_TANH, true>(G_host, p, st) : smix_launch_t<N, SD_ACT_TANH, false>(G_host, p, st);
    case SD_ACT_TANH_TANH: return has_res ? smix_launch_t<N, SD_ACT_TANH_TANH, true>(G_host, p, st) : smix_launch_t<N, SD_ACT_TANH_TANH, false>(G_host, p, st);
    }
    set_error("sample_mix: unknown activation %d", p.act);
    return SD_ERR_INVALID;
}

bool sample_mix_supported(int N, int OUT, const float* y, const Epilogue& epi, const ViewW& out) {
    if (!(N == 16 || N == 17 || N == 21)) return false;
    if (OUT % 32 || OUT <= 0 || (size_t)N * OUT * 4 * 2 > 200 * 1024) return false;
    if (reinterpret_cast<uintptr_t>(y) & 15u) return false;
    if (epi.ss_row_idx || out.rep != 1) return false;                 // per-sample time rows: generic kernel
    return true;
}

// out = epilogue(G^ @ (rs * Y)); the caller checked sample_mix_supported
int sample_mix_fp32(const float* G_host, int N, int OUT, const float* y, const float* row_scale, const Epilogue& epi,
                    const ViewW& out, int B, cudaStream_t st) {
    if (B <= 0) return SD_OK;
    SmixParams p;
    p.y = y; p.row_scale = row_scale; p.bias_node = epi.bias_node;
    p.ss = epi.ss ? epi.ss + (long long)epi.ss_row * epi.ss_stride : nullptr;
    p.act = epi.act; p.residual = epi.residual; p.out = out; p.B = B; p.OUT = OUT; p.stages = 0;
    const bool has_res = epi.residual.ptr != nullptr;
    if (N == 21) return smix_launch_n<21>(G_host, p, has_res, st);
    if (N == 16) return smix_launch_n<16>(G_host, p, has_res, st);
    if (N == 17) return smix_launch_n<17>(G_host, p, has_res, st);
    set_error("sample_mix: %d nodes not instantiated", N);
    return SD_ERR_UNSUPPORTED;
}


// =====================================================================================================================
// Graph-GRU step after the recurrent product, general graph influence gx_i (recurrent.py:333-358):
//     xr = gx_i @ (x W_ih^T + b_ih),  hr = gx_i @ (h W_hh^T + b_hh)
//     r = sig(xr_r + hr_r), z = sig(xr_z + hr_z), n = tanh(xr_n + r * hr_n), h' = n - n z + z h
// The raw products x W_ih^T (loop invariant in the decoder, decoder.py:81,93) and h W_hh^T (tcgen05 kernel, one launch per
// step) arrive as [B, N, 3H] fp32; the mix is linear, so r and z need ONE mix each (of xr + hr) and n needs two: four
// N x N mixes per hidden unit.  Stage = (sample, 32 hidden units): 6 N row segments of 128 B from the two product tensors
// and N from h, fetched by the 32 lanes of the copy warp with one cp.async.bulk each; compute warp = one stage, lane = unit.
// gx_i @ bias is precomputed per step (plan.py).  MIX = false (every gx_i = I) skips the FFMAs: the kernel is then the plain
// gate kernel with bulk-copy staging.
// =====================================================================================================================
constexpr int GRS_WARPS = 11, GRS_THREADS = (GRS_WARPS + 1) * 32, GRS_MAX_STAGES = 12;      // 11 compute warps + 1 copy warp

struct GruSampleParams {
    const float* hr;           // [B][N][3H] raw h W_hh^T
    View xr;                   // raw x W_ih^T: row (b, n) = 3H floats
    View h_prev;               // row (b, n) = H floats
    const float* bias_x;       // [N][3H] = gx_i @ b_ih[type] or null
    const float* bias_h;       // [N][3H]
    ViewW h_out;
    int B, H, stages;
};
struct __align__(8) GrsBarriers { uint64_t full[GRS_MAX_STAGES], done[GRS_MAX_STAGES]; };


template <int N, bool MIX, bool FAST>
__global__ void __launch_bounds__(GRS_THREADS, 1)
gru_sample_kernel(const __grid_constant__ MixMat<N> G, const GruSampleParams p) {
    extern __shared__ __align__(128) float grs_smem[];
    constexpr int SEG = 32;                                           // floats per row segment (128 B)
    constexpr int STAGE_FLOATS = 7 * N * SEG;                         // [hr r|z|n][N][32] [xr r|z|n][N][32] [h][N][32]
    GrsBarriers* bars = reinterpret_cast<GrsBarriers*>(grs_smem + (size_t)p.stages * STAGE_FLOATS);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int chunks = p.H / SEG;
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { tc::mbar_init(&bars->full[s], 1); tc::mbar_init(&bars->done[s], 1); }
        tc::fence_barrier_init();
    }
    __syncthreads();
    // task t = (round r, chunk c, warp w): sample k = r * WARPS + w of this CTA's list, hidden units [32 c, 32 c + 32)
    const int my_samples = (int)blockIdx.x < p.B ? (p.B - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int rounds = (my_samples + GRS_WARPS - 1) / GRS_WARPS;
    const long long tasks = (long long)rounds * chunks * GRS_WARPS;
    const int H3 = 3 * p.H;
    if (warp == GRS_WARPS) {
        // ------------------------------------------------------------ copy warp: 7 N segments per stage, one bulk copy each
        long long q = 0;                                              // sequence number over the tasks that exist
        for (long long t = 0; t < tasks; ++t) {
            const int w = (int)(t % GRS_WARPS), c = (int)((t / GRS_WARPS) % chunks), r = (int)(t / ((long long)GRS_WARPS * chunks));
            const int k = r * GRS_WARPS + w;
            if (k >= my_samples) continue;
            const int st = (int)(q % p.stages);
            const long long use = q / p.stages;
            ++q;
            if (use > 0) tc::mbar_wait(&bars->done[st], (uint32_t)(use - 1) & 1u);
            if (lane == 0) tc::mbar_arrive_expect_tx(&bars->full[st], (uint32_t)STAGE_FLOATS * 4u);
            __syncwarp();
            const long long b = (long long)blockIdx.x + (long long)k * gridDim.x;
            float* dst = grs_smem + (size_t)st * STAGE_FLOATS;
            const float* hr_b = p.hr + b * N * H3 + c * SEG;
            const float* xr_b = p.xr.ptr + (p.xr.rep == 1 ? b : b / p.xr.rep) * p.xr.sb + c * SEG;
            const float* h_b = p.h_prev.ptr + (p.h_prev.rep == 1 ? b : b / p.h_prev.rep) * p.h_prev.sb + c * SEG;
            for (int i = lane; i < 7 * N; i += 32) {
                const int arr = i / (3 * N);                          // 0: hr, 1: xr, 2: h
                const int rem = i - arr * 3 * N;
                const int g = rem / N, n = rem - g * N;
                const float* src = arr == 0 ? hr_b + (long long)n * H3 + g * p.H
                                 : arr == 1 ? xr_b + (long long)n * p.xr.sn + g * p.H
                                            : h_b + (long long)n * p.h_prev.sn;
                smix_bulk_load(dst + i * SEG, src, SEG * 4u, &bars->full[st]);
            }
        }
        return;
    }
    for (long long t = warp; t < tasks; t += GRS_WARPS) {
        const int c = (int)((t / GRS_WARPS) % chunks), r = (int)(t / ((long long)GRS_WARPS * chunks));
        const int k = r * GRS_WARPS + warp;
        if (k >= my_samples) continue;
        // sequence number of this task among the tasks that exist (the last round may have fewer than WARPS samples)
        const int full_rounds = my_samples / GRS_WARPS, tail = my_samples - full_rounds * GRS_WARPS;
        const long long q = r < full_rounds ? t : (long long)full_rounds * chunks * GRS_WARPS + (long long)c * tail + warp;
        const int st = (int)(q % p.stages);
        tc::mbar_wait(&bars->full[st], (uint32_t)(q / p.stages) & 1u);
        const float* s = grs_smem + (size_t)st * STAGE_FLOATS + lane;
        const int u = c * SEG + lane;
        // gate g of node n: hr at s[(g*N + n)*32], xr at s[((3+g)*N + n)*32], h at s[(6*N + n)*32]
        // the four mixes run one after the other and each reads its inputs from the stage right before it (keeping all five
        // input sets in registers next to the accumulators spilled 2 KB per thread); the stage is released after the last read
        const long long b = (long long)blockIdx.x + (long long)k * gridDim.x;
        float* ob = p.h_out.ptr + b * p.h_out.sb + u;
        auto mix = [&](const float (&in)[N][1], float (&out)[N][1]) {
            if (MIX) { mix_nodes<N, 1>(G, in, out); return; }
#pragma unroll
            for (int n = 0; n < N; ++n) out[n][0] = in[n][0];
        };
        // The four mixes run as a NON-unrolled loop over the phases r, z, hr_n, xr_n: unrolled, the compiler keeps the N*N
        // coefficients of G in registers across the four copies (2 KB of spills per thread); each phase reads its inputs from
        // the stage right before its mix, and the stage is released after the last read.
        float in[N][1], acc[N][1], rg[N], zg[N], hp[N];
#pragma unroll 1
        for (int ph = 0; ph < 4; ++ph) {
            const float* sa = s + (ph < 3 ? ph : 5) * N * SEG;        // hr_r, hr_z, hr_n, xr_n
#pragma unroll
            for (int n = 0; n < N; ++n) in[n][0] = sa[n * SEG];
            if (ph < 2) {
                const float* sx = s + (3 + ph) * N * SEG;              // + xr_r, xr_z (the mix is linear)
#pragma unroll
                for (int n = 0; n < N; ++n) in[n][0] += sx[n * SEG];
            }
            if (ph == 3) {
#pragma unroll
                for (int n = 0; n < N; ++n) hp[n] = s[(6 * N + n) * SEG];
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&bars->done[st]);     // the stage may be refilled
            }
            mix(in, acc);
            if (ph == 0) {
#pragma unroll
                for (int n = 0; n < N; ++n) {
                    float v = acc[n][0];
                    if (p.bias_x) v += __ldg(p.bias_x + n * H3 + u) + __ldg(p.bias_h + n * H3 + u);
                    rg[n] = mix_sigmoid<FAST>(v);
                }
            } else if (ph == 1) {
#pragma unroll
                for (int n = 0; n < N; ++n) {
                    float v = acc[n][0];
                    if (p.bias_x) v += __ldg(p.bias_x + n * H3 + p.H + u) + __ldg(p.bias_h + n * H3 + p.H + u);
                    zg[n] = mix_sigmoid<FAST>(v);
                }
            } else if (ph == 2) {
#pragma unroll
                for (int n = 0; n < N; ++n) {
                    float v = acc[n][0];
                    if (p.bias_h) v += __ldg(p.bias_h + n * H3 + 2 * p.H + u);
                    rg[n] *= v;                                       // r * hr_n
                }
            } else {
#pragma unroll
                for (int n = 0; n < N; ++n) {
                    float v = acc[n][0] + rg[n];
                    if (p.bias_x) v += __ldg(p.bias_x + n * H3 + 2 * p.H + u);
                    const float nn = mix_tanh<FAST>(v);
                    ob[(long long)n * p.h_out.sn] = nn - nn * zg[n] + zg[n] * hp[n];
                }
            }
        }
    }
}

template <int N, bool MIX, bool FAST>
static int grs_launch_t(const float* G_host, const GruSampleParams& p0, cudaStream_t st) {
    GruSampleParams p = p0;
    MixMat<N> G;
    G.set(MIX ? G_host : nullptr);
    const size_t stage_bytes = (size_t)7 * N * 32 * 4;
    int stages = (int)((220 * 1024) / stage_bytes);
    if (stages > GRS_MAX_STAGES) stages = GRS_MAX_STAGES;
    p.stages = stages;
    const size_t smem = stages * stage_bytes + sizeof(GrsBarriers) + 128;
    auto kern = gru_sample_kernel<N, MIX, FAST>;
    static unsigned long long configured = 0;
    if (int rc = opt_in_smem(kern, 227 * 1024, configured)) return rc;
    const int sms = sm_count();
    const int grid = p.B < sms ? p.B : sms;
    kern<<<grid, GRS_THREADS, smem, st>>>(G, p);
    SD_LAUNCH_OK("gru_sample_kernel");
    return SD_OK;
}

bool gru_sample_supported(int N, int H, const float* hr, const View& xr, const View& h_prev, const ViewW& h_out) {
    if (!(N == 16 || N == 17 || N == 21)) return false;
    if (H % 32) return false;
    auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
    if (!al(hr) || !al(xr.ptr) || !al(h_prev.ptr) || xr.sb % 4 || xr.sn % 4 || h_prev.sb % 4 || h_prev.sn % 4) return false;
    return h_out.rep == 1;
}

// h' = GRU gates of one step from the raw products; G_host = gx_i ([N][N], host) or null for the identity
int gru_sample_fp32(const float* G_host, int N, int H, const float* hr, const View& xr, const float* bias_x, const float* bias_h,
                    const View& h_prev, const ViewW& h_out, int B, cudaStream_t st) {
    if (B <= 0) return SD_OK;
    GruSampleParams p;
    p.hr = hr; p.xr = xr; p.h_prev = h_prev; p.bias_x = bias_x; p.bias_h = bias_h; p.h_out = h_out; p.B = B; p.H = H; p.stages = 0;
    if ((bias_x == nullptr) != (bias_h == nullptr)) { set_error("gru_sample: both bias tables or none"); return SD_ERR_INVALID; }
    const bool fast = fast_epilogue();
#define SD_GRS(NN) if (N == NN) return G_host ? (fast ? grs_launch_t<NN, true, true>(G_host, p, st) : grs_launch_t<NN, true, false>(G_host, p, st)) \
                                              : (fast ? grs_launch_t<NN, false, true>(G_host, p, st) : grs_launch_t<NN, false, false>(G_host, p, st));
    SD_GRS(21) SD_GRS(16) SD_GRS(17)
#undef SD_GRS
    set_error("gru_sample: %d nodes not instantiated", N);
    return SD_ERR_UNSUPPORTED;
}

// =====================================================================================================================
// Decoder output head with a general graph influence: y[b,n,:] = act( sum_m G^[n,m] (W_fc[type(m)] h[b,m,:]) + bias_node[n] )
// (decoder.py:97-98 through graph_structural.py:30-43).  F <= 4 outputs per node.  One warp per sample: lane l < N owns node l
// for the H-long dot products (h rows read with coalesced float4 loads through shared memory), then the N x N mix.
// =====================================================================================================================
constexpr int GHD_WARPS = 8;

template <int N>
__global__ void __launch_bounds__(GHD_WARPS * 32)
gru_head_kernel(const __grid_constant__ MixMat<N> G, const float* __restrict__ Wfc, const float* __restrict__ bias_node, const NodeTypes types,
                int H, int F, const View h, const ViewW out, int act, int B) {
    extern __shared__ __align__(16) float ghd_smem[];                 // [warps][N][H + 1] rows + [warps][N][4] products
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int LD = H + 1;
    float* rows = ghd_smem + (size_t)warp * (N * LD + N * 4);
    float* prod = rows + N * LD;
    for (int b = blockIdx.x * GHD_WARPS + warp; b < B; b += gridDim.x * GHD_WARPS) {
        const float* hb = h.ptr + (long long)(h.rep == 1 ? b : b / h.rep) * h.sb;
        __syncwarp();
        for (int i = lane; i < N * H; i += 32) {
            const int n = i / H, u = i - n * H;
            rows[n * LD + u] = __ldg(hb + (long long)n * h.sn + u);
        }
        __syncwarp();
        if (lane < N) {
            const float* w = Wfc + (long long)types.t[lane] * F * H;
            float a[4] = {0.f, 0.f, 0.f, 0.f};
            for (int u = 0; u < H; ++u) {
                const float hv = rows[lane * LD + u];
#pragma unroll
                for (int f = 0; f < 4; ++f) if (f < F) a[f] = fmaf(__ldg(w + f * H + u), hv, a[f]);
            }
#pragma unroll
            for (int f = 0; f < 4; ++f) prod[lane * 4 + f] = a[f];
        }
        __syncwarp();
        for (int i = lane; i < N * F; i += 32) {
            const int n = i / F, f = i - n * F;
            float v = 0.0f;
            const float* gcol = reinterpret_cast<const float*>(&G.g4[0][0]);            // g4[m][q] component j = G^[4q + j][m]
            for (int m = 0; m < N; ++m) v = fmaf(gcol[(m * MixMat<N>::Q + (n >> 2)) * 4 + (n & 3)], prod[m * 4 + f], v);
            if (bias_node) v += __ldg(bias_node + n * F + f);
            if (act == SD_ACT_TANH) v = tanhf(v);
            out.ptr[(long long)b * out.sb + (long long)n * out.sn + f] = v;
        }
    }
}

template <int N>
static int ghd_launch(const float* G_host, const float* Wfc, const float* bias_node, const NodeTypes& types, int H, int F,
                      const View& h, const ViewW& out, int act, int B, cudaStream_t st) {
    MixMat<N> G;
    G.set(G_host);
    const size_t smem = (size_t)GHD_WARPS * (N * (H + 1) + N * 4) * sizeof(float);
    auto kern = gru_head_kernel<N>;
    static unsigned long long configured = 0;
    if (int rc = opt_in_smem(kern, 160 * 1024, configured)) return rc;
    int grid = (B + GHD_WARPS - 1) / GHD_WARPS;
    const int cap = sm_count() * 4;
    if (grid > cap) grid = cap;
    kern<<<grid, GHD_WARPS * 32, smem, st>>>(G, Wfc, bias_node, types, H, F, h, out, act, B);
    SD_LAUNCH_OK("gru_head_kernel");
    return SD_OK;
}

bool gru_head_supported(int N, int H, int F) { return (N == 16 || N == 17 || N == 21) && F >= 1 && F <= 4 && H <= 256; }

int gru_head_fp32(const float* G_host, const float* Wfc, const float* bias_node, const NodeTypes& types, int N, int H, int F,
                  const View& h, const ViewW& out, int act, int B, cudaStream_t st) {
    if (B <= 0) return SD_OK;
    if (N == 21) return ghd_launch<21>(G_host, Wfc, bias_node, types, H, F, h, out, act, B, st);
    if (N == 16) return ghd_launch<16>(G_host, Wfc, bias_node, types, H, F, h, out, act, B, st);
    if (N == 17) return ghd_launch<17>(G_host, Wfc, bias_node, types, H, F, h, out, act, B, st);
    set_error("gru_head: %d nodes not instantiated", N);
    return SD_ERR_UNSUPPORTED;
}

}  // namespace sd
