// Per-sample kernels for a general (dense) graph-influence matrix G^: the node mix of a StaticGraphLinear with its layer
// epilogue, and the graph-GRU step after the recurrent product.
//
//   out[b,n,:] = epilogue( sum_m G^[n,m] * rs[b,m] * Y[b,m,:] )         (graph_structural.py:30-43, gmm :7-8)
//
// The grouped GEMM kernels produce the raw per-node products Y[b,m,:] = x[b,m,:] W[type(m)]^T; the mix couples all nodes of
// a sample.  Work item = (sample, 32 or 64 columns): lane = column, the N inputs of the column in registers, N accumulators,
// N*N FFMAs whose G^ operand comes from the constant bank (the matrix is passed BY VALUE as a __grid_constant__ kernel
// parameter, sd_mixmat.cuh).  The epilogue (mixed bias, (scale+1)x+shift, tanh, residual) is applied to the accumulators and
// the result leaves with coalesced 128-byte stores: Y is read once, the output written once.  Bound: HBM (Y + residual + out
// = 3 x 4 B per element); the FFMA work is N FMAs per element (21 for AMASS: 4.3 GFLOP per 192-wide layer at B = 25 600).
//
// Staging.  Persistent CTAs of 12 warps; every warp owns a PRIVATE ring of shared-memory slots and fetches the [N rows][32 or
// 64 columns] box of its next items itself with one TMA tensor load per box (lane 0), a few boxes ahead of the one it is
// working on.  No warp ever waits for another one.  A first version shared one ring between a copy warp and eleven compute
// warps that each visited only every third phase of a slot: with mbarrier PARITY waits a warp that asks for phase u while the
// barrier is still in phase u - 1 (two bulk copies completing out of order is enough) passes immediately, reads a half-filled
// slot and signals `done` for the wrong phase -- run-to-run differences and, rarely, a lost phase (deadlock) at B = 25 600.
// With private rings every phase of a barrier is waited for by the same warp in order, so the parity is unambiguous.
#include "sd_internal.h"
#include "sd_tc.cuh"
#include "sd_mixmat.cuh"
#include <stdlib.h>
#include <cuda.h>

namespace sd {

constexpr int PSK_WARPS = 12;                  // all compute (ptxas budgets registers per 4-warp group: 168 per thread)
constexpr int PSK_THREADS = PSK_WARPS * 32;
constexpr int PSK_SMEM = 200 * 1024;           // ring bytes per CTA (all warps)

// FAST: tanh / sigmoid through MUFU.EX2 + MUFU.RCP (absolute error ~1e-7) instead of libdevice (SKELDIFF_ACCURATE_EPILOGUE=1 selects libdevice)
template <bool FAST> __device__ __forceinline__ float mix_tanh(float x) { return FAST ? tc::tanh_ex2(x) : tanhf(x); }
template <bool FAST> __device__ __forceinline__ float mix_sigmoid(float v) {
    return FAST ? __fdividef(1.0f, 1.0f + exp2f(-1.4426950408889634f * v)) : 1.0f / (1.0f + expf(-v));
}
bool fast_epilogue() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SKELDIFF_ACCURATE_EPILOGUE"); v = (e && e[0] == '1') ? 0 : 1; }
    return v == 1;
}

typedef CUresult (*PskEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// fp32 tensor map of rank 3 or 4 without swizzle (dims / box innermost first, strides in bytes for dims 1..rank-1)
static int psk_map(CUtensorMap* m, const float* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box) {
    static PskEncodeFn enc = nullptr;
    if (!enc) {
        void* fp = nullptr; cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
            set_error("cuTensorMapEncodeTiled unavailable"); return SD_ERR_CUDA;
        }
        enc = reinterpret_cast<PskEncodeFn>(fp);
    }
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<float*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("per-sample kernel: cuTensorMapEncodeTiled failed (%d)", (int)r); return SD_ERR_CUDA; }
    return SD_OK;
}

// One warp's private ring of `slots` boxes.  issue(j) is called by the whole warp for its j-th box (lane 0 arms the slot's
// barrier and starts the TMA load; the fence orders the warp's earlier generic-proxy reads of the slot before the copy
// engine's refill), acquire(j) waits until box j has landed.  Box j lives in slot j % slots, phase (j / slots) & 1.
struct WarpRing {
    float* buf; uint64_t* full; int slots, box_floats;
    __device__ __forceinline__ float* slot_ptr(int j) const { return buf + (size_t)(j % slots) * box_floats; }
    __device__ __forceinline__ void acquire(int j) const { tc::mbar_wait(&full[j % slots], (uint32_t)(j / slots) & 1u, j); }
};

struct SmixParams {
    const float* row_scale;    // [B*N] or null: Y row (b, m) is scaled before the mix (RMSNorm factor of the to_qkv input)
    const float* bias_node;    // [N][OUT] (already mixed: G^ @ bias[type]) or null
    const float* ss;           // resolved scale/shift row (scale at [o], shift at [OUT + o]) or null
    int act;
    View residual;             // ptr null if none
    ViewW out;
    int B, OUT, slots;
};

// COLS: columns per lane (c and c + 32 of a 64-column item when OUT is a multiple of 64: each constant load feeds two FFMAs)
template <int N, int ACT, bool HAS_RES, int COLS, bool FAST>
__global__ void __launch_bounds__(PSK_THREADS, 1)
sample_mix_kernel(const __grid_constant__ MixMat<N> G, const __grid_constant__ CUtensorMap map_y, const SmixParams p) {
    extern __shared__ __align__(128) float psk_smem[];
    constexpr int BOX = N * 32 * COLS;                                // floats: [N rows][32 * COLS columns]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    WarpRing ring;
    ring.slots = p.slots; ring.box_floats = BOX;
    ring.buf = psk_smem + (size_t)warp * p.slots * BOX;
    ring.full = reinterpret_cast<uint64_t*>(psk_smem + (size_t)PSK_WARPS * p.slots * BOX) + warp * p.slots;
    if (lane == 0) {
        if (warp == 0) tc::tma_prefetch_desc(&map_y);
        for (int s = 0; s < p.slots; ++s) tc::mbar_init(&ring.full[s], 1);
        tc::fence_barrier_init();
    }
    __syncwarp();
    const int chunks = p.OUT / (32 * COLS);                           // items per sample
    const int my_samples = (int)blockIdx.x < p.B ? (p.B - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int items = my_samples * chunks;                            // this CTA's items; warp w takes w, w + WARPS, ...
    const int n_my = items > warp ? (items - 1 - warp) / PSK_WARPS + 1 : 0;
    // the mixed bias [N][OUT] is read once per item and output: kept in shared memory when it fits (a global load per use
    // left the warps waiting on the L1 / L2 round trip: 33 % of the stall samples of the first version were these FADDs)
    float* bias_s = reinterpret_cast<float*>(reinterpret_cast<uint64_t*>(psk_smem + (size_t)PSK_WARPS * p.slots * BOX) + PSK_WARPS * p.slots);
    if (p.bias_node) {                                                // (the launcher guarantees that the table fits)
        for (int i = threadIdx.x; i < N * p.OUT; i += PSK_THREADS) bias_s[i] = __ldg(p.bias_node + i);
        __syncthreads();
    }
    // position of the next box to fetch, advanced incrementally (no divisions in the loop)
    int is_j = 0, is_k = warp / chunks, is_c = warp - is_k * chunks, is_slot = 0;
    const int dk = PSK_WARPS / chunks, dc = PSK_WARPS - dk * chunks;
    auto issue = [&]() {
        tc::fence_proxy_async();
        __syncwarp();
        if (is_j < n_my && lane == 0) {
            uint64_t* bar = &ring.full[is_slot];
            tc::mbar_arrive_expect_tx(bar, (uint32_t)BOX * 4u);
            tc::tma_load_3d(ring.buf + (size_t)is_slot * BOX, &map_y, bar, is_c * (32 * COLS), 0, (int)blockIdx.x + is_k * (int)gridDim.x);
        }
        ++is_j;
        is_k += dk; is_c += dc;
        if (is_c >= chunks) { is_c -= chunks; ++is_k; }
        if (++is_slot == ring.slots) is_slot = 0;
    };
    for (int j = 0; j < p.slots - 1; ++j) issue();
    int k = warp / chunks, cc = warp - k * chunks, slot = 0;
    uint32_t phase = 0;
    for (int j = 0; j < n_my; ++j) {
        issue();                                                      // refills the slot whose box was read in the previous iteration
        const int c = cc * (32 * COLS) + lane;
        const int b = (int)blockIdx.x + k * (int)gridDim.x;
        float res[N][COLS];
        if (HAS_RES) {                                                // the residual row segments are in flight during the mix
            const float* rb = p.residual.ptr + (long long)(p.residual.rep == 1 ? b : b / p.residual.rep) * p.residual.sb + c;
#pragma unroll
            for (int n = 0; n < N; ++n)
#pragma unroll
                for (int q = 0; q < COLS; ++q) res[n][q] = __ldg(rb + (long long)n * p.residual.sn + 32 * q);
        }
        float mul[COLS], add[COLS];
#pragma unroll
        for (int q = 0; q < COLS; ++q) {
            mul[q] = p.ss ? __ldg(p.ss + c + 32 * q) + 1.0f : 1.0f;
            add[q] = p.ss ? __ldg(p.ss + p.OUT + c + 32 * q) : 0.0f;
        }
        tc::mbar_wait(&ring.full[slot], phase, j);
        const float* ys = ring.buf + (size_t)slot * BOX + lane;       // box [N][32 * COLS]
        float in[N][COLS];
#pragma unroll
        for (int m = 0; m < N; ++m)
#pragma unroll
            for (int q = 0; q < COLS; ++q) in[m][q] = ys[m * (32 * COLS) + 32 * q];
        if (p.row_scale) {
            const float* rs = p.row_scale + (long long)b * N;
#pragma unroll
            for (int m = 0; m < N; ++m) {
                const float r = __ldg(rs + m);
#pragma unroll
                for (int q = 0; q < COLS; ++q) in[m][q] *= r;
            }
        }
        float acc[N][COLS];
        mix_nodes<N, COLS>(G, in, acc);
        float* ob = p.out.ptr + (long long)b * p.out.sb + c;
#pragma unroll
        for (int n = 0; n < N; ++n)
#pragma unroll
            for (int q = 0; q < COLS; ++q) {
                float v = acc[n][q];
                if (p.bias_node) v += bias_s[n * p.OUT + c + 32 * q];
                v = fmaf(v, mul[q], add[q]);
                if (ACT == SD_ACT_TANH) v = mix_tanh<FAST>(v);
                if (ACT == SD_ACT_TANH_TANH) v = mix_tanh<FAST>(mix_tanh<FAST>(v));
                if (HAS_RES) v += res[n][q];
                ob[(long long)n * p.out.sn + 32 * q] = v;
            }
        k += dk; cc += dc;
        if (cc >= chunks) { cc -= chunks; ++k; }
        if (++slot == ring.slots) { slot = 0; phase ^= 1u; }
    }
}

template <int N, int ACT, bool HAS_RES, int COLS, bool FAST>
static int smix_launch_c(const float* G_host, const float* y, const SmixParams& p0, cudaStream_t st) {
    SmixParams p = p0;
    MixMat<N> G;
    G.set(G_host);
    const size_t box_bytes = (size_t)N * 32 * COLS * 4;
    int slots = (int)(PSK_SMEM / (PSK_WARPS * box_bytes));
    if (slots > 8) slots = 8;
    if (slots < 2) { set_error("sample_mix: box of %zu bytes does not fit a two-slot ring per warp", box_bytes); return SD_ERR_UNSUPPORTED; }
    p.slots = slots;
    CUtensorMap map_y;
    {
        const cuuint64_t dims[3] = {(cuuint64_t)p.OUT, (cuuint64_t)N, (cuuint64_t)p.B};
        const cuuint64_t strides[2] = {(cuuint64_t)p.OUT * 4, (cuuint64_t)N * p.OUT * 4};
        const cuuint32_t box[3] = {(cuuint32_t)(32 * COLS), (cuuint32_t)N, 1};
        if (int rc = psk_map(&map_y, y, 3, dims, strides, box)) return rc;
    }
    const size_t bias_bytes = p.bias_node ? (size_t)N * p.OUT * 4 : 0;
    if (bias_bytes > 24 * 1024) { set_error("sample_mix: bias table of %zu bytes does not fit shared memory", bias_bytes); return SD_ERR_UNSUPPORTED; }
    const size_t smem = (size_t)PSK_WARPS * slots * (box_bytes + 8) + bias_bytes + 128;
    auto kern = sample_mix_kernel<N, ACT, HAS_RES, COLS, FAST>;
    static unsigned long long configured = 0;
    if (int rc = opt_in_smem(kern, 227 * 1024, configured)) return rc;
    const int sms = sm_count();
    const int grid = p.B < sms ? p.B : sms;
    kern<<<grid, PSK_THREADS, smem, st>>>(G, map_y, p);
    SD_LAUNCH_OK("sample_mix_kernel");
    return SD_OK;
}

template <int N, int ACT, bool HAS_RES>
static int smix_launch_t(const float* G_host, const float* y, const SmixParams& p, cudaStream_t st) {
    constexpr bool kHasAct = ACT != SD_ACT_NONE;
    const bool fast = kHasAct && fast_epilogue();
    if (p.OUT % 64 == 0) return fast ? smix_launch_c<N, ACT, HAS_RES, 2, kHasAct>(G_host, y, p, st) : smix_launch_c<N, ACT, HAS_RES, 2, false>(G_host, y, p, st);
    return fast ? smix_launch_c<N, ACT, HAS_RES, 1, kHasAct>(G_host, y, p, st) : smix_launch_c<N, ACT, HAS_RES, 1, false>(G_host, y, p, st);
}

template <int N>
static int smix_launch_n(const float* G_host, const float* y, const SmixParams& p, bool has_res, cudaStream_t st) {
    switch (p.act) {
    case SD_ACT_NONE: return has_res ? smix_launch_t<N, SD_ACT_NONE, true>(G_host, y, p, st) : smix_launch_t<N, SD_ACT_NONE, false>(G_host, y, p, st);
    case SD_ACT_TANH: return has_res ? smix_launch_t<N, SD_ACT_TANH, true>(G_host, y, p, st) : smix_launch_t<N, SD_ACT_TANH, false>(G_host, y, p, st);
    case SD_ACT_TANH_TANH: return has_res ? smix_launch_t<N, SD_ACT_TANH_TANH, true>(G_host, y, p, st) : smix_launch_t<N, SD_ACT_TANH_TANH, false>(G_host, y, p, st);
    }
    set_error("sample_mix: unknown activation %d", p.act);
    return SD_ERR_INVALID;
}

bool sample_mix_supported(int N, int OUT, const float* y, const Epilogue& epi, const ViewW& out) {
    static int off = -1;                         // SKELDIFF_NO_SAMPLE_MIX=1: generic node-mix kernel (A/B timing, bisection)
    if (off < 0) { const char* e = getenv("SKELDIFF_NO_SAMPLE_MIX"); off = (e && e[0] == '1') ? 1 : 0; }
    if (off) return false;
    if (!(N == 16 || N == 17 || N == 21)) return false;
    if (OUT % 32 || OUT <= 0) return false;
    if (epi.bias_node && (size_t)N * OUT * 4 > 24 * 1024) return false;   // the mixed bias table lives in shared memory
    if (reinterpret_cast<uintptr_t>(y) & 15u) return false;
    if (epi.ss_row_idx || out.rep != 1) return false;                 // per-sample time rows: generic kernel
    return true;
}

// out = epilogue(G^ @ (rs * Y)); the caller checked sample_mix_supported
int sample_mix_fp32(const float* G_host, int N, int OUT, const float* y, const float* row_scale, const Epilogue& epi,
                    const ViewW& out, int B, cudaStream_t st) {
    if (B <= 0) return SD_OK;
    SmixParams p;
    p.row_scale = row_scale; p.bias_node = epi.bias_node;
    p.ss = epi.ss ? epi.ss + (long long)epi.ss_row * epi.ss_stride : nullptr;
    p.act = epi.act; p.residual = epi.residual; p.out = out; p.B = B; p.OUT = OUT; p.slots = 0;
    const bool has_res = epi.residual.ptr != nullptr;
    if (N == 21) return smix_launch_n<21>(G_host, y, p, has_res, st);
    if (N == 16) return smix_launch_n<16>(G_host, y, p, has_res, st);
    if (N == 17) return smix_launch_n<17>(G_host, y, p, has_res, st);
    set_error("sample_mix: %d nodes not instantiated", N);
    return SD_ERR_UNSUPPORTED;
}

// =====================================================================================================================
// Graph-GRU step after the recurrent product, general graph influence gx_i (recurrent.py:333-358):
//     xr = gx_i @ (x W_ih^T + b_ih),  hr = gx_i @ (h W_hh^T + b_hh)
//     r = sig(xr_r + hr_r), z = sig(xr_z + hr_z), n = tanh(xr_n + r * hr_n), h' = n - n z + z h
// The raw products x W_ih^T (loop invariant in the decoder, decoder.py:81,93) and h W_hh^T (tcgen05 kernel, one launch per
// step) arrive as [B, N, 3H] fp32; the mix is linear, so r and z need ONE mix each (of xr + hr) and n needs two: four
// N x N mixes per hidden unit, run as a non-unrolled loop over the phases r, z, hr_n, xr_n.  Item = (sample, 32 hidden
// units); it consumes seven [N][32] boxes in the order hr_r, xr_r, hr_z, xr_z, hr_n, xr_n, h, each one TMA tensor load into
// the warp's private ring (see the header), fetched a few boxes ahead.  gx_i @ bias is precomputed per step (plan.py).
// MIX = false (every gx_i = I) skips the FFMAs: the kernel is then the plain gate kernel.
// =====================================================================================================================
struct GruSampleParams {
    const float* bias_x;       // [N][3H] = gx_i @ b_ih[type] or null
    const float* bias_h;       // [N][3H]
    ViewW h_out;
    int B, H, slots;
};

template <int N, bool MIX, bool FAST>
__global__ void __launch_bounds__(PSK_THREADS, 1)
gru_sample_kernel(const __grid_constant__ MixMat<N> G, const __grid_constant__ CUtensorMap map_hr, const __grid_constant__ CUtensorMap map_xr,
                  const __grid_constant__ CUtensorMap map_h, const GruSampleParams p) {
    extern __shared__ __align__(128) float psk_smem[];
    constexpr int BOX = N * 32, BPI = 7;                              // box [N][32] floats; boxes per item
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    WarpRing ring;
    ring.slots = p.slots; ring.box_floats = BOX;
    ring.buf = psk_smem + (size_t)warp * p.slots * BOX;
    ring.full = reinterpret_cast<uint64_t*>(psk_smem + (size_t)PSK_WARPS * p.slots * BOX) + warp * p.slots;
    if (lane == 0) {
        if (warp == 0) { tc::tma_prefetch_desc(&map_hr); tc::tma_prefetch_desc(&map_xr); tc::tma_prefetch_desc(&map_h); }
        for (int s = 0; s < p.slots; ++s) tc::mbar_init(&ring.full[s], 1);
        tc::fence_barrier_init();
    }
    __syncwarp();
    const int chunks = p.H / 32;
    const int my_samples = (int)blockIdx.x < p.B ? (p.B - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int items = my_samples * chunks;
    const int n_my = items > warp ? (items - 1 - warp) / PSK_WARPS + 1 : 0;
    const int H3 = 3 * p.H;
    // next box to fetch: (item k / chunk c, part); parts in consumption order hr_r, xr_r, hr_z, xr_z, hr_n, xr_n, h
    int is_it = 0, is_part = 0, is_k = warp / chunks, is_c = warp - is_k * chunks, is_slot = 0;
    const int dk = PSK_WARPS / chunks, dc = PSK_WARPS - dk * chunks;
    auto issue = [&]() {
        tc::fence_proxy_async();
        __syncwarp();
        if (is_it < n_my && lane == 0) {
            const int b = (int)blockIdx.x + is_k * (int)gridDim.x;
            uint64_t* bar = &ring.full[is_slot];
            float* dst = ring.buf + (size_t)is_slot * BOX;
            tc::mbar_arrive_expect_tx(bar, (uint32_t)BOX * 4u);
            if (is_part == 6) tc::tma_load_3d(dst, &map_h, bar, is_c * 32, 0, b);
            else tc::tma_load_4d(dst, (is_part & 1) ? &map_xr : &map_hr, bar, is_c * 32, is_part >> 1, 0, b);
        }
        if (++is_part == BPI) {
            is_part = 0; ++is_it;
            is_k += dk; is_c += dc;
            if (is_c >= chunks) { is_c -= chunks; ++is_k; }
        }
        if (++is_slot == ring.slots) is_slot = 0;
    };
    for (int j = 0; j < p.slots - 1; ++j) issue();
    int k = warp / chunks, cc = warp - k * chunks, slot = 0;
    uint32_t phase = 0;
    // consume one box: refill the slot freed one box earlier, then wait for this one
    auto take = [&]() -> const float* {
        issue();
        tc::mbar_wait(&ring.full[slot], phase, k);
        const float* ptr = ring.buf + (size_t)slot * BOX + lane;
        if (++slot == ring.slots) { slot = 0; phase ^= 1u; }
        return ptr;
    };
    for (int it = 0; it < n_my; ++it) {
        const int u = cc * 32 + lane;
        const long long b = (long long)blockIdx.x + (long long)k * gridDim.x;
        float* ob = p.h_out.ptr + b * p.h_out.sb + u;
        // Two passes over G with TWO columns each: (r, z) then (hr_n, xr_n).  One pass with four columns needs 168 registers
        // for the inputs and accumulators alone; four passes of one column re-read G four times and, unrolled or as a
        // predicated loop, cost 5 300 instructions per item (ncu) against ~3 000 here.
        float in[N][2], acc[N][2], rg[N], zg[N], hp[N];
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
            {
                const float* s0 = take();                             // hr_r | hr_n
#pragma unroll
                for (int n = 0; n < N; ++n) in[n][0] = s0[n * 32];
                const float* s1 = take();                             // xr_r | xr_n
                if (pass == 0) {
#pragma unroll
                    for (int n = 0; n < N; ++n) in[n][0] += s1[n * 32];
                    const float* s2 = take();                         // hr_z
#pragma unroll
                    for (int n = 0; n < N; ++n) in[n][1] = s2[n * 32];
                    const float* s3 = take();                         // xr_z
#pragma unroll
                    for (int n = 0; n < N; ++n) in[n][1] += s3[n * 32];
                } else {
#pragma unroll
                    for (int n = 0; n < N; ++n) in[n][1] = s1[n * 32];
                    const float* s2 = take();                         // h
#pragma unroll
                    for (int n = 0; n < N; ++n) hp[n] = s2[n * 32];
                }
            }
            if (MIX) mix_nodes<N, 2>(G, in, acc);
            else {
#pragma unroll
                for (int n = 0; n < N; ++n) { acc[n][0] = in[n][0]; acc[n][1] = in[n][1]; }
            }
            if (pass == 0) {
#pragma unroll
                for (int n = 0; n < N; ++n) {
                    float vr = acc[n][0], vz = acc[n][1];
                    if (p.bias_x) {
                        vr += __ldg(p.bias_x + n * H3 + u) + __ldg(p.bias_h + n * H3 + u);
                        vz += __ldg(p.bias_x + n * H3 + p.H + u) + __ldg(p.bias_h + n * H3 + p.H + u);
                    }
                    rg[n] = mix_sigmoid<FAST>(vr);
                    zg[n] = mix_sigmoid<FAST>(vz);
                }
            } else {
#pragma unroll
                for (int n = 0; n < N; ++n) {
                    float vh = acc[n][0], vx = acc[n][1];
                    if (p.bias_x) { vh += __ldg(p.bias_h + n * H3 + 2 * p.H + u); vx += __ldg(p.bias_x + n * H3 + 2 * p.H + u); }
                    const float nn = mix_tanh<FAST>(vx + rg[n] * vh);
                    ob[(long long)n * p.h_out.sn] = nn - nn * zg[n] + zg[n] * hp[n];
                }
            }
        }
        k += dk; cc += dc;
        if (cc >= chunks) { cc -= chunks; ++k; }
    }
}

template <int N, bool MIX, bool FAST>
static int grs_launch_t(const float* G_host, const float* hr, const View& xr, const View& h_prev, const GruSampleParams& p0, cudaStream_t st) {
    GruSampleParams p = p0;
    MixMat<N> G;
    G.set(MIX ? G_host : nullptr);
    CUtensorMap map_hr, map_xr, map_h;
    {   // gate g of node n of sample b: [b][n][g][u]; a box is [N nodes][32 units] of one gate
        const cuuint64_t H = (cuuint64_t)p.H, B = (cuuint64_t)p.B;
        const cuuint64_t d4[4] = {H, 3, (cuuint64_t)N, B};
        const cuuint32_t b4[4] = {32, 1, (cuuint32_t)N, 1};
        const cuuint64_t s_hr[3] = {H * 4, 3 * H * 4, (cuuint64_t)N * 3 * H * 4};
        const cuuint64_t s_xr[3] = {H * 4, (cuuint64_t)xr.sn * 4, (cuuint64_t)xr.sb * 4};
        const cuuint64_t d3[3] = {H, (cuuint64_t)N, B};
        const cuuint32_t b3[3] = {32, (cuuint32_t)N, 1};
        const cuuint64_t s_h[2] = {(cuuint64_t)h_prev.sn * 4, (cuuint64_t)h_prev.sb * 4};
        if (int rc = psk_map(&map_hr, hr, 4, d4, s_hr, b4)) return rc;
        if (int rc = psk_map(&map_xr, xr.ptr, 4, d4, s_xr, b4)) return rc;
        if (int rc = psk_map(&map_h, h_prev.ptr, 3, d3, s_h, b3)) return rc;
    }
    const size_t box_bytes = (size_t)N * 32 * 4;
    int slots = (int)(PSK_SMEM / (PSK_WARPS * box_bytes));
    if (slots > 7) slots = 7;
    p.slots = slots;
    const size_t smem = (size_t)PSK_WARPS * slots * (box_bytes + 8) + 128;
    auto kern = gru_sample_kernel<N, MIX, FAST>;
    static unsigned long long configured = 0;
    if (int rc = opt_in_smem(kern, 227 * 1024, configured)) return rc;
    const int sms = sm_count();
    const int grid = p.B < sms ? p.B : sms;
    kern<<<grid, PSK_THREADS, smem, st>>>(G, map_hr, map_xr, map_h, p);
    SD_LAUNCH_OK("gru_sample_kernel");
    return SD_OK;
}

bool gru_sample_supported(int N, int H, const float* hr, const View& xr, const View& h_prev, const ViewW& h_out) {
    if (!(N == 16 || N == 17 || N == 21)) return false;
    if (H % 32) return false;
    auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
    if (!al(hr) || !al(xr.ptr) || !al(h_prev.ptr) || xr.sb % 4 || xr.sn % 4 || h_prev.sb % 4 || h_prev.sn % 4) return false;
    return h_out.rep == 1 && xr.rep == 1 && h_prev.rep == 1;          // TMA boxes address sample b directly
}

// h' = GRU gates of one step from the raw products; G_host = gx_i ([N][N], host) or null for the identity
int gru_sample_fp32(const float* G_host, int N, int H, const float* hr, const View& xr, const float* bias_x, const float* bias_h,
                    const View& h_prev, const ViewW& h_out, int B, cudaStream_t st) {
    if (B <= 0) return SD_OK;
    GruSampleParams p;
    p.bias_x = bias_x; p.bias_h = bias_h; p.h_out = h_out; p.B = B; p.H = H; p.slots = 0;
    if ((bias_x == nullptr) != (bias_h == nullptr)) { set_error("gru_sample: both bias tables or none"); return SD_ERR_INVALID; }
    const bool fast = fast_epilogue();
#define SD_GRS(NN) if (N == NN) return G_host ? (fast ? grs_launch_t<NN, true, true>(G_host, hr, xr, h_prev, p, st) : grs_launch_t<NN, true, false>(G_host, hr, xr, h_prev, p, st)) \
                                              : (fast ? grs_launch_t<NN, false, true>(G_host, hr, xr, h_prev, p, st) : grs_launch_t<NN, false, false>(G_host, hr, xr, h_prev, p, st));
    SD_GRS(21) SD_GRS(16) SD_GRS(17)
#undef SD_GRS
    set_error("gru_sample: %d nodes not instantiated", N);
    return SD_ERR_UNSUPPORTED;
}

// =====================================================================================================================
// Decoder output head with a general graph influence: y[b,n,:] = act( sum_m G^[n,m] (W_fc[type(m)] h[b,m,:]) + bias_node[n] )
// (decoder.py:97-98 through graph_structural.py:30-43).  F <= 3 outputs per node, N F <= 64.  One warp per sample: lane = hidden
// unit (coalesced 128-byte reads of the h rows, W_fc and G^ in shared memory), the N F partial dot products of the lanes are
// summed with a butterfly reduce-scatter (62 shuffles for 64 values instead of 5 per value), then the N x N mix.
// HBM-bound: reads h once (206 MB at B = 25 600), writes 6 MB.
// =====================================================================================================================
constexpr int GHD_WARPS = 8;

template <int N, int F>
__global__ void __launch_bounds__(GHD_WARPS * 32)
gru_head_kernel(const float* __restrict__ G, const float* __restrict__ Wfc, const float* __restrict__ bias_node, const NodeTypes types,
                int H, int n_types, const View h, const ViewW out, int act, int B) {
    extern __shared__ __align__(16) float ghd_smem[];                 // [n_types * F * H] weights, [N * N] G^, [warps][64] products
    float* w_s = ghd_smem;
    float* g_s = w_s + n_types * F * H;
    float* p_s = g_s + N * N + (threadIdx.x >> 5) * 64;
    for (int i = threadIdx.x; i < n_types * F * H; i += blockDim.x) w_s[i] = __ldg(Wfc + i);
    for (int i = threadIdx.x; i < N * N; i += blockDim.x) g_s[i] = G ? __ldg(G + i) : ((i / N == i % N) ? 1.0f : 0.0f);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int b = blockIdx.x * GHD_WARPS + warp; b < B; b += gridDim.x * GHD_WARPS) {
        const float* hb = h.ptr + (long long)(h.rep == 1 ? b : b / h.rep) * h.sb + lane;
        float c[64];
#pragma unroll
        for (int i = 0; i < 64; ++i) c[i] = 0.0f;
        for (int u0 = 0; u0 < H; u0 += 32) {
            float hv[N];
#pragma unroll
            for (int n = 0; n < N; ++n) hv[n] = (u0 + lane < H) ? __ldg(hb + (long long)n * h.sn + u0) : 0.0f;
#pragma unroll
            for (int n = 0; n < N; ++n) {
                const float* w = w_s + types.t[n] * F * H + u0 + lane;
#pragma unroll
                for (int f = 0; f < F; ++f) c[n * F + f] = fmaf(hv[n], (u0 + lane < H) ? w[f * H] : 0.0f, c[n * F + f]);
            }
        }
        // butterfly reduce-scatter over the 32 lanes: after the step with mask m a lane keeps the half of its values selected by
        // its bit m; lane L ends with the complete sums of indices base, base + 1, base = 32 b4 + 16 b3 + 8 b2 + 4 b1 + 2 b0
#pragma unroll
        for (int half = 32, mask = 16; half >= 2; half >>= 1, mask >>= 1) {
            const bool hi = (lane & mask) != 0;
#pragma unroll
            for (int i = 0; i < half; ++i) {
                const float send = hi ? c[i] : c[i + half];
                const float keep = hi ? c[i + half] : c[i];
                c[i] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
            }
        }
        const int base = ((lane & 16) ? 32 : 0) + ((lane & 8) ? 16 : 0) + ((lane & 4) ? 8 : 0) + ((lane & 2) ? 4 : 0) + ((lane & 1) ? 2 : 0);
        __syncwarp();
        p_s[base] = c[0];
        p_s[base + 1] = c[1];
        __syncwarp();
        for (int i = lane; i < N * F; i += 32) {
            const int n = i / F, f = i - n * F;
            float v = 0.0f;
#pragma unroll
            for (int m = 0; m < N; ++m) v = fmaf(g_s[n * N + m], p_s[m * F + f], v);
            if (bias_node) v += __ldg(bias_node + n * F + f);
            if (act == SD_ACT_TANH) v = tanhf(v);
            out.ptr[(long long)b * out.sb + (long long)n * out.sn + f] = v;
        }
    }
}

template <int N, int F>
static int ghd_launch(const float* G_dev, const float* Wfc, const float* bias_node, const NodeTypes& types, int n_types, int H,
                      const View& h, const ViewW& out, int act, int B, cudaStream_t st) {
    static_assert(N * F <= 64, "the reduce-scatter holds 64 values");
    const size_t smem = ((size_t)n_types * F * H + N * N + GHD_WARPS * 64) * sizeof(float);
    auto kern = gru_head_kernel<N, F>;
    static unsigned long long configured = 0;
    if (smem > 48 * 1024) if (int rc = opt_in_smem(kern, smem, configured)) return rc;
    int grid = (B + GHD_WARPS - 1) / GHD_WARPS;
    const int cap = sm_count() * 4;
    if (grid > cap) grid = cap;
    kern<<<grid, GHD_WARPS * 32, smem, st>>>(G_dev, Wfc, bias_node, types, H, n_types, h, out, act, B);
    SD_LAUNCH_OK("gru_head_kernel");
    return SD_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Tiled variant of the output head (the default): a CTA stages the h rows of a few whole samples in shared memory with
// coalesced loads (row stride H + 1: conflict-free for the row-per-thread dots), W_fc of every node type and G^ next to them;
// thread = (row, output) dot product over H, then thread = (sample, node, output) for the N x N mix, bias, activation.
// No shuffle trees, 4-byte stores of 3 floats per row are the only uncoalesced access.  Reads h once: HBM-bound.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int GH2_THREADS = 256;
__global__ void __launch_bounds__(GH2_THREADS)
gru_head_tiled_kernel(const float* __restrict__ G, const float* __restrict__ Wfc, const float* __restrict__ bias_node, const NodeTypes types,
                      int N, int H, int F, int n_types, int spb, const View h, const ViewW out, int act, int B) {
    extern __shared__ float gh2[];
    const int LD = H + 1, rows_max = spb * N;
    float* h_s = gh2;                                   // [spb * N][H + 1]
    float* w_s = h_s + (size_t)rows_max * LD;           // [n_types * F][H + 1]
    float* g_s = w_s + (size_t)n_types * F * LD;        // [N][N] (when G)
    float* p_s = g_s + (G ? N * N : 0);                 // [spb * N][F]
    for (int i = threadIdx.x; i < n_types * F * H; i += GH2_THREADS) w_s[(i / H) * LD + i % H] = __ldg(Wfc + i);
    if (G) for (int i = threadIdx.x; i < N * N; i += GH2_THREADS) g_s[i] = __ldg(G + i);
    const int h4 = H >> 2;
    for (long long b0 = (long long)blockIdx.x * spb; b0 < B; b0 += (long long)gridDim.x * spb) {
        const int ns = (int)min((long long)spb, (long long)B - b0), rows = ns * N;
        __syncthreads();                                // previous tile consumed (also orders the W / G fill)
        for (int i = threadIdx.x; i < rows * h4; i += GH2_THREADS) {
            const int r = i / h4, q = i - r * h4;
            const int bl = r / N, n = r - bl * N;
            const float4 v = __ldg(reinterpret_cast<const float4*>(row_ptr(h, (int)(b0 + bl), n) + 4 * q));
            float* d = h_s + r * LD + 4 * q;
            d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < rows * F; i += GH2_THREADS) {
            const int f = i / rows, r = i - f * rows;   // consecutive threads = consecutive rows (distinct banks), same output
            const int n = r % N;
            const float* hv = h_s + r * LD;
            const float* wv = w_s + (types.t[n] * F + f) * LD;
            float a0 = 0.f, a1 = 0.f;
            for (int u = 0; u + 1 < H; u += 2) { a0 = fmaf(hv[u], wv[u], a0); a1 = fmaf(hv[u + 1], wv[u + 1], a1); }
            if (H & 1) a0 = fmaf(hv[H - 1], wv[H - 1], a0);
            p_s[r * F + f] = a0 + a1;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < rows * F; i += GH2_THREADS) {
            const int r = i / F, f = i - r * F;
            const int bl = r / N, n = r - bl * N;
            float v;
            if (G) {
                v = 0.f;
                for (int m = 0; m < N; ++m) v = fmaf(g_s[n * N + m], p_s[(bl * N + m) * F + f], v);
            } else {
                v = p_s[r * F + f];
            }
            if (bias_node) v += __ldg(bias_node + n * F + f);
            if (act == SD_ACT_TANH) v = tanhf(v);
            out.ptr[(b0 + bl) * out.sb + (long long)n * out.sn + f] = v;
        }
    }
}

static int gru_head_tiled(const float* G_dev, const float* Wfc, const float* bias_node, const NodeTypes& types, int n_types, int N, int H, int F,
                          const View& h, const ViewW& out, int act, int B, cudaStream_t st, bool* used) {
    *used = false;
    if (H % 4 || h.rep != 1 || (reinterpret_cast<uintptr_t>(h.ptr) & 15u) || h.sb % 4 || h.sn % 4) return SD_OK;
    int spb = 96 / N; if (spb < 1) spb = 1;
    const size_t smem = ((size_t)spb * N * (H + 1) + (size_t)n_types * F * (H + 1) + (G_dev ? N * N : 0) + (size_t)spb * N * F) * sizeof(float);
    if (smem > 100 * 1024) return SD_OK;
    static unsigned long long configured = 0;
    if (smem > 48 * 1024) if (int rc = opt_in_smem(gru_head_tiled_kernel, 100 * 1024, configured)) return rc;
    long long grid = ((long long)B + spb - 1) / spb;
    const long long cap = (long long)sm_count() * 8;
    if (grid > cap) grid = cap;
    gru_head_tiled_kernel<<<(unsigned)grid, GH2_THREADS, smem, st>>>(G_dev, Wfc, bias_node, types, N, H, F, n_types, spb, h, out, act, B);
    count_launch();
    if (check_cuda(cudaGetLastError(), "gru_head_tiled_kernel")) return SD_ERR_CUDA;
    *used = true;
    return SD_OK;
}

bool gru_head_supported(int N, int H, int F) { return (N == 16 || N == 17 || N == 21) && F >= 1 && F <= 3 && H <= 512; }

// G_dev: fc's normalised graph influence [N][N] on the device, or null for the identity
int gru_head_fp32(const float* G_dev, const float* Wfc, const float* bias_node, const NodeTypes& types, int n_types, int N, int H, int F,
                  const View& h, const ViewW& out, int act, int B, cudaStream_t st) {
    if (B <= 0) return SD_OK;
    {
        static int tiled_env = -1;           // SKELDIFF_GRU_HEAD_TILED=0: the warp-per-sample shuffle kernel (A/B timing)
        if (tiled_env < 0) { const char* e = getenv("SKELDIFF_GRU_HEAD_TILED"); tiled_env = (e && e[0] == '0') ? 0 : 1; }
        if (tiled_env) {
            bool used = false;
            if (int rc = gru_head_tiled(G_dev, Wfc, bias_node, types, n_types, N, H, F, h, out, act, B, st, &used)) return rc;
            if (used) return SD_OK;
        }
    }
#define SD_GHD(NN) if (N == NN) { \
        if (F == 3) return ghd_launch<NN, 3>(G_dev, Wfc, bias_node, types, n_types, H, h, out, act, B, st); \
        if (F == 2) return ghd_launch<NN, 2>(G_dev, Wfc, bias_node, types, n_types, H, h, out, act, B, st); \
        return ghd_launch<NN, 1>(G_dev, Wfc, bias_node, types, n_types, H, h, out, act, B, st); }
    SD_GHD(21) SD_GHD(16) SD_GHD(17)
#undef SD_GHD
    set_error("gru_head: %d nodes not instantiated", N);
    return SD_ERR_UNSUPPORTED;
}

}  // namespace sd
