// StaticGraphLinear with fp32-grade products on the 16-bit tensor cores: fp32 operands are split into planes on the way into
// shared memory, the plane products are accumulated in fp32 in TMEM, activations stay fp32 in HBM.
//
// PL = 3 ("bf16x3"): x = x0 + x1 + x2, three bf16 planes of 8 significand bits each (exact), and
//     x . w  ~=  x0 w0 + x0 w1 + x0 w2 + x1 w0 + x1 w1 + x2 w0        (terms below 2^-24 |x||w| dropped):
// six tcgen05.mma.kind::f16 per K = 16 step.  The tensor core TRUNCATES when it adds a K = 16 partial sum to the accumulator
// (measured, scratch/acc_bias.py: with all six pairs in one accumulator the result was 12.5 ulp short in magnitude at K = 192,
// 7x the rms error of an FFMA chain), so x0 w0 goes to a "main" accumulator and the five small pairs to a second one whose
// truncation error is 2^-8 of that; the epilogue adds the two with one round-to-nearest FADD.
// PL = 2 ("fp16x2", the default): x = hi + lo 2^-11 with hi = fp16(x), lo = fp16((x - hi) 2^11): 22 significand bits, three
// products (hi hi -> main | hi lo + lo hi -> corr, scaled by 2^-11 in the epilogue), 32 KB stages.
//
// Warp roles (512 threads, one CTA per SM, persistent over (node, m-tile) items):
//   warps 0-7  transform producers: fp32 activation granules (LDG.128 of any sd_view, or LDS from the TMA-fed fp32 ring "ATMA"),
//              split, STS.64 into SWIZZLE_128B K-major plane tiles, fence.proxy.async, arrive on full[stage]
//   warps 8-11 epilogue (setmaxnreg: 176 registers): tcgen05.ld (main + corr) -> swizzled staging tile (row = lane) -> transposed
//              read (8 lanes cover 128 contiguous bytes of a row) -> row scale, bias, scale/shift, MUFU tanh, residual (own loads, or
//              [128 x 32] boxes of the TMA residual ring "RTMA") -> row-major staging -> one bulk tensor store per warp and chunk.
//              Bare layers store the first staging image directly (p.direct); T3_ACT_GRU applies the GRU gates.
//   warp 12    TMEM allocator, MMA issue; in the weight-resident schedule also the TMA of the weight planes
//   warp 13    activation-stationary schedule: streams (n-tile, k-block) weight planes through a two-slot TMA ring
//   warp 14    residual / GRU operand ring (RTMA);   warp 15: fp32 activation ring (ATMA)
// Two schedules (weight-resident / activation-stationary) and the K-split of two-segment layers are described at the places they
// are chosen (glin_tc3_launch, t3_launch_one).  Roofline of the 192 -> 192 layer at B = 25 600: the three fp32 streams (in,
// residual, out) are 0.19 ms of HBM time, 3 x 2*K*OUT FLOP per row at the 16-bit rate 0.07 ms: HBM is the binding floor; measured
// 0.30 ms with tanh + residual, 0.21 ms bare (DESIGN.md section 4 table, 4.15 for what the source-level profiles showed).
//
// Reference semantics: GraphLinear.forward, src/core/network/layers/graph_structural.py:30-43.
#include "sd_internal.h"
#include "sd_tc.cuh"
#include <cuda.h>
#include <stdlib.h>

namespace sd {

using namespace sd::tc;

constexpr int T3_BM = 128, T3_BK = 64;
constexpr int T3_PRODUCERS = 256;   // warps 0-7 (two per scheduler hide each other's LDG latency)
constexpr int T3_EPI_WARP0 = 8;     // warps 8-11: epilogue (warp % 4 = TMEM lane quarter)
constexpr int T3_MMA_WARP = 12, T3_ALLOC_WARP = 12;   // warp 12: TMEM allocation, MMA issue (+ weight TMA in the weight-resident mode)
constexpr int T3_WLOAD_WARP = 13;   // warp 13: weight-slot TMA ring of the activation-stationary mode
constexpr int T3_THREADS = 512;     // 16 warps = 4 warpgroups of 128 registers per thread at launch; warps 14-15 only take part in
                                    // the register hand-over: warpgroup 3 (MMA / loader / idle) shrinks to T3_REGS_MISC per thread and
                                    // the epilogue warpgroup grows to T3_REGS_EPI (setmaxnreg), which pays for a residual prefetch
                                    // T3_RES_AHEAD chunks deep (one chunk in flight left the 4 warps waiting on L2 latency)
constexpr int T3_REGS_MISC = 80, T3_REGS_EPI = 176, T3_RES_AHEAD = 3;
constexpr int T3_MAX_STAGES = 8, T3_MAX_WSLOTS = 4, T3_MAX_RES_SLOTS = 4;
// Internal epilogue code (not an SD_ACT_* of the C ABI): fused GRU gates of the decoder's recurrent step, see the GRU block of the epilogue
constexpr int T3_ACT_GRU = 3;
constexpr int T3_PLANE_BYTES = T3_BM * 128;         // one plane tile: 128 rows x 128 B (64 k of 16-bit operands)
// PL = 3: three bf16 planes (exact split, any fp32 magnitude), six products.  PL = 2: two fp16 planes, x = hi + lo * 2^-11 with
// hi = fp16(x), lo = fp16((x - hi) * 2^11): 22 significand bits, three products (hi hi | hi lo + lo hi, the latter scaled by
// 2^-11 in the epilogue), |x| < 65 504 (larger operands become inf and the result NaN: loud, not silently wrong).
__host__ __device__ constexpr int t3_stage_bytes(int planes) { return planes * T3_PLANE_BYTES; }
__host__ __device__ constexpr int t3_chunk_cols(int planes) { return planes == 2 ? 32 : 16; }     // epilogue chunk width (columns)

struct T3Params {
    View a0, a1;
    int B, N, K, OUT, BN, NT, MT, KB, nstage, n_types, tmem_cols;
    int k_base;                 // first weight column of this launch (K-split of a two-segment layer)
    View pre;                   // fp32 partial product added before the epilogue (ptr null if none)
    int a_stationary;           // 1: the stage ring holds the whole K of an m-tile, n-tiles looped inside the CTA, weights streamed
    int wslots;                 // weight slots of the activation-stationary mode (2 .. T3_MAX_WSLOTS)
    int res_tma;                // 1: the residual tile rides a TMA ring (warp 14 -> two [128 rows][32 columns] boxes in shared memory)
    int out_tma;                // 1: the output chunk leaves through a TMA store of the warp's staging tile (two-plane kernel)
    int raw_slots;              // ATMA: fp32 activation k-blocks arrive by TMA in a ring of [128 rows][64 floats] boxes (warp 15)
    int stg2;                   // 1: two staging tiles per epilogue warp (a chunk's bulk store overlaps the next chunk)
    float* norm_out;            // RTMA variant, CTA covers whole rows: inv[b * N + node] = 1 / max(||out row||, 1e-12) (the RMSNorm factor of
                                // the NEXT layer, layers/attention.py:36) from the values the epilogue holds anyway; null: not wanted
    int direct;                 // two-plane kernel, bare layer (no bias / scale-shift / activation / residual / partial product): the staging
                                // tile after the TMEM read IS the SWIZZLE_128B image of the output box and leaves as it is
    int merge_ld;               // two-plane epilogue: the four TMEM loads of a chunk before one wait (t3_chunk32_to_stage)
    int tab_cols;               // columns of the per-node epilogue tables in shared memory (multiple of 16)
    int res_slots;              // boxes of the residual ring (2; the fused GRU step takes as many as fit, up to T3_MAX_RES_SLOTS)
    const float* gru_bias_x;    // T3_ACT_GRU: gate-interleaved biases [N][3H] of the x side and of the h side
    const float* gru_bias_h;
    NodeTypes types;
    const float* row_scale;
    const float* bias_node;
    const float* ss;            // resolved scale/shift row or null
    View residual;              // fp32, ptr null if none
    ViewW out;
};

struct __align__(8) T3Barriers {
    uint64_t full[T3_MAX_STAGES], empty[T3_MAX_STAGES];
    uint64_t w_full, w_empty;
    uint64_t ws_full[T3_MAX_WSLOTS], ws_empty[T3_MAX_WSLOTS];       // weight slots (activation-stationary mode)
    uint64_t acc_full[2], acc_empty[2];
    uint64_t res_full[T3_MAX_RES_SLOTS], res_empty[T3_MAX_RES_SLOTS];     // residual chunk ring (res_tma)
    uint64_t raw_full[4], raw_empty[4];     // fp32 activation ring (ATMA)
    uint32_t tmem_base, pad;
};

// Exact 3-way split by truncation: the 24 significand bits of x fall 8/8/8 into three bf16 values held in the
// UPPER halves of h, m, l (x == h + m + l exactly).  Pure LOP/FADD: the first version used F2F.BF16.F32
// (round-to-nearest), whose quarter-rate conversion unit made the producers the bottleneck of the kernel.
__device__ __forceinline__ void split3(float x, uint32_t& h, uint32_t& m, uint32_t& l) {
    h = __float_as_uint(x) & 0xFFFF0000u;
    const float r1 = x - __uint_as_float(h);
    m = __float_as_uint(r1) & 0xFFFF0000u;
    const float r2 = r1 - __uint_as_float(m);
    l = __float_as_uint(r2) & 0xFFFF0000u;
}
// Two-plane fp16 split of a pair: hi = fp16(x) (one F2FP for both), lo = fp16((x - hi) * 2^11); x - hi is exact in fp32.
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn((a - hf.x) * 2048.0f, (b - hf.y) * 2048.0f);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
// Streaming 16-byte load of an activation granule: read once per CTA, so it does not allocate in L1 (the in-flight
// granules of the 256 producer threads are 64 KB, most of the L1 left beside the shared-memory carve-out).
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
// Bulk L2 prefetch: brings `bytes` (multiple of 16) at p into L2 without occupying registers or shared memory.
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(p), "r"(bytes) : "memory");
}
// (hi16 of b) << 16 | (hi16 of a): two bf16 packed in element order a, b
__device__ __forceinline__ uint32_t pack_hi(uint32_t a, uint32_t b) { return __byte_perm(a, b, 0x7632); }

// FAST: tanh through MUFU.EX2 + MUFU.RCP (tc::tanh_ex2, ~1e-7 absolute) instead of libdevice tanhf (SKELDIFF_ACCURATE_EPILOGUE=1)
template <bool FAST> __device__ __forceinline__ float t3_tanh(float x) { return FAST ? tanh_ex2(x) : tanhf(x); }

// -DSD_T3_LDG_STREAM: activation granules with L1::no_allocate.  Measured within run-to-run noise of the allocating load
// (192 -> 192 bare 228 vs 217 us, tanh 247 vs 253 us, sampling loop 199.5 vs 201.9 ms identity, 273.1 vs 271.6 ms dense): not enabled.
#ifdef SD_T3_LDG_STREAM
#define T3_LDA(p) ldg_stream(p)
#else
#define T3_LDA(p) __ldg(p)
#endif
// One 32-column chunk of the two-plane accumulators (main + corr * 2^-11, times the row scale) from TMEM into the warp's swizzled
// staging tile, row = lane.  merge: all four 16-column loads are issued before ONE wait (64 registers in flight) instead of two
// load / wait / store rounds: one TMEM round trip less on the epilogue warps' critical path.
__device__ __forceinline__ void t3_chunk32_to_stage(uint32_t t_main, uint32_t t_corr, float* stg, int lane, float rs, bool merge) {
    constexpr float CS = 0.00048828125f;
    uint32_t v0[16], c0[16], v1[16], c1[16];
    auto put = [&](const uint32_t (&v)[16], const uint32_t (&vc)[16], int hf) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float4 x;
            x.x = fmaf(__uint_as_float(vc[4 * q + 0]), CS, __uint_as_float(v[4 * q + 0])) * rs;
            x.y = fmaf(__uint_as_float(vc[4 * q + 1]), CS, __uint_as_float(v[4 * q + 1])) * rs;
            x.z = fmaf(__uint_as_float(vc[4 * q + 2]), CS, __uint_as_float(v[4 * q + 2])) * rs;
            x.w = fmaf(__uint_as_float(vc[4 * q + 3]), CS, __uint_as_float(v[4 * q + 3])) * rs;
            *reinterpret_cast<float4*>(stg + lane * 32 + 4 * ((4 * hf + q) ^ (lane & 7))) = x;
        }
    };
    tmem_ld_32x16(t_main, v0);
    tmem_ld_32x16(t_corr, c0);
    if (merge) { tmem_ld_32x16(t_main + 16u, v1); tmem_ld_32x16(t_corr + 16u, c1); }
    tmem_ld_wait();
    put(v0, c0, 0);
    if (!merge) { tmem_ld_32x16(t_main + 16u, v1); tmem_ld_32x16(t_corr + 16u, c1); tmem_ld_wait(); }
    put(v1, c1, 1);
}

// NORM: the epilogue also folds the sum of squares of the finished rows into p.norm_out (its own instantiation: the eight extra
// accumulators cost the tanh + residual kernel 9 % - 304 -> 333 us - and are paid only by the launches that want the factors)
template <int ACT, bool HAS_RES, bool FAST, int PL, bool RTMA, bool ATMA, bool NORM = false>
__global__ void __launch_bounds__(T3_THREADS, 1)
glin_tc3_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_r, const __grid_constant__ CUtensorMap map_o,
                const __grid_constant__ CUtensorMap map_a, const T3Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-byte alignment (SWIZZLE_128B atoms) by pointer arithmetic on the __shared__ array: rounding the address up through
    // uintptr_t made the compiler lose the address space and emit generic LD/ST for every shared-memory access of the kernel.
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t w_block = (uint32_t)p.BN * 128u;                 // one (plane, k-block) weight tile
    uint8_t* w_smem = smem;                                        // resident: [plane][kb][BN x 128 B]; streamed: [slot][plane][BN x 128 B]
    constexpr int STAGE_BYTES = t3_stage_bytes(PL);
    uint8_t* a_smem = w_smem + (size_t)PL * (p.a_stationary ? p.wslots : p.KB) * w_block;   // [stage][plane][128 x 128 B]
    // staging tiles first: they follow the 1024-byte aligned plane stages, so every warp's [32 rows][128 bytes] tile is 1024-byte
    // aligned (what a SWIZZLE_128B bulk tensor store of the tile expects, see p.direct)
    float* epi_stage = reinterpret_cast<float*>(a_smem + (size_t)p.nstage * STAGE_BYTES);   // 4 warps x [32 rows][16 or 32 floats], swizzled (x2: stg2)
    float* epi_mul = epi_stage + (p.stg2 ? 2 : 1) * 4 * 32 * t3_chunk_cols(PL);
    float* epi_add = epi_mul + p.tab_cols;                         // tab_cols: the columns this CTA produces for one node (BN, or OUT when it loops over the n-tiles)
    float* res_buf = epi_add + p.tab_cols;                         // res_tma: res_slots x [128 rows][32 floats]
    constexpr int RES_BOX = T3_BM * 32;
    float* a_raw = res_buf + (RTMA ? p.res_slots * RES_BOX : 0);   // ATMA: raw_slots x [128 rows][64 floats]
    constexpr int RAW_BOX = T3_BM * T3_BK;
    T3Barriers* bars = reinterpret_cast<T3Barriers*>(a_raw + (ATMA ? p.raw_slots * RAW_BOX : 0));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&map_w);
        for (int s = 0; s < p.nstage; ++s) { mbar_init(&bars->full[s], T3_PRODUCERS); mbar_init(&bars->empty[s], 1); }
        mbar_init(&bars->w_full, 1);
        mbar_init(&bars->w_empty, 1);
        for (int s = 0; s < T3_MAX_WSLOTS; ++s) { mbar_init(&bars->ws_full[s], 1); mbar_init(&bars->ws_empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&bars->acc_full[s], 1); mbar_init(&bars->acc_empty[s], 128); }
        for (int s = 0; s < T3_MAX_RES_SLOTS; ++s) { mbar_init(&bars->res_full[s], 1); mbar_init(&bars->res_empty[s], 4); }
        for (int s = 0; s < 4; ++s) { mbar_init(&bars->raw_full[s], 1); mbar_init(&bars->raw_empty[s], 8); }
        fence_barrier_init();
    }
    if (warp == T3_ALLOC_WARP) tmem_alloc(&bars->tmem_base, (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    // Gang scheduling: the NT CTAs of a gang walk the same (node, m-tile) sequence at the same time, each for its own
    // 96-column n-tile, so the activation tile fetched by one of them is an L2 hit for the others (in group-major order
    // the second read came from DRAM again: 2x A traffic in the first ncu capture).  Weights change only per node.
    //
    // Activation-stationary mode (K <= 192, wide outputs): the stage ring holds the whole K of ONE m-tile, every CTA owns
    // whole m-tiles and loops over the NT n-tiles itself while warp 13 streams the (n-tile, k-block) weight planes from
    // L2 through a two-slot TMA ring.  The activations are fetched and split once instead of NT times (to_qkv: NT = 8).
    const bool as_mode = p.a_stationary != 0;
    const int my_nt = as_mode ? 0 : (int)(blockIdx.x % p.NT);
    const int nt_lo = my_nt, nt_hi = as_mode ? p.NT : my_nt + 1;
    const long long gang = as_mode ? blockIdx.x : blockIdx.x / p.NT, n_gangs = as_mode ? gridDim.x : gridDim.x / p.NT;
    const long long total = (long long)p.N * p.MT;
    const long long item_lo = total * gang / n_gangs;
    const long long item_hi = total * (gang + 1) / n_gangs;

    if (warp < 8) {
        // ================================================================ transform producers (256 threads)
        const int t = threadIdx.x;
        const int col4 = t & 15;            // which float4 of the 64-wide k-block row
        const int row0 = t >> 4;            // rows row0 + 16*i, i = 0..7
        const uint32_t chunk = (uint32_t)(col4 >> 1), half = (uint32_t)(col4 & 1) << 3;
        // Registers are the only place where loads can be "in flight" here (shared memory is full of weight and
        // plane tiles): the producers double-buffer 8 x float4 granules, so the next k-block's loads are outstanding
        // while the current one is split and stored (a 3-deep pipeline does not fit the 128-register budget).
        int stage = 0; uint32_t phase = 0;
        const long long q_end = (item_hi - item_lo) * p.KB;       // flattened (tile, k-block) sequence
        // The (node, m-tile, k-block) position of the fetch stream advances incrementally and views without an in-place
        // repeat take a division-free path: in the first ncu capture two thirds of the producers' 860 instructions per
        // thread and k-block were 64-bit / runtime-divisor integer divisions (q / KB, it % MT, b / rep per row).
        int f_kb = 0, f_mt = (int)(item_lo % p.MT), f_node = (int)(item_lo / p.MT);
        long long f_left = q_end, e_left = q_end;
        auto fetch = [&](float4 (&v)[8]) {
            if (f_left <= 0) return;
            --f_left;
            // column of the concatenated K axis this thread fetches; the segment is chosen per float4 (segment widths are
            // multiples of 4, not of 64: cat[x_cond, x] = 96 + 96, the recurrent product has K = 96).  Columns beyond the
            // last segment are zero planes (the weight planes' out-of-range columns are zero-filled by TMA).
            const int kcol = f_kb * T3_BK + col4 * 4;
            const bool seg0 = kcol < p.a0.width;
            const View& seg = seg0 ? p.a0 : p.a1;
            const int scol = seg0 ? kcol : kcol - p.a0.width;
            const float* base = seg.ptr + (long long)f_node * seg.sn + scol;
            const int b_end = scol < seg.width ? p.B : 0;
            const int b0 = f_mt * T3_BM + row0;
            if (seg.rep == 1) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int b = b0 + 16 * i;
                    v[i] = (b < b_end) ? T3_LDA(reinterpret_cast<const float4*>(base + (long long)b * seg.sb)) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int b = b0 + 16 * i;
                    v[i] = (b < b_end) ? T3_LDA(reinterpret_cast<const float4*>(base + (long long)(b / seg.rep) * seg.sb)) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            if (++f_kb == p.KB) { f_kb = 0; if (++f_mt == p.MT) { f_mt = 0; ++f_node; } }
        };
        auto emit = [&](const float4 (&v)[8]) {
            if (e_left <= 0) return;
            --e_left;
            MBAR_WAIT_AT(&bars->empty[stage], phase ^ 1);
            uint8_t* st = a_smem + (size_t)stage * STAGE_BYTES;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = row0 + 16 * i;
                const uint32_t off = (uint32_t)r * 128u + ((chunk ^ (uint32_t)(r & 7)) << 4) + half;
                if (PL == 3) {
                    uint32_t h[4], m[4], l[4];
                    split3(v[i].x, h[0], m[0], l[0]); split3(v[i].y, h[1], m[1], l[1]);
                    split3(v[i].z, h[2], m[2], l[2]); split3(v[i].w, h[3], m[3], l[3]);
                    *reinterpret_cast<uint2*>(st + off) = make_uint2(pack_hi(h[0], h[1]), pack_hi(h[2], h[3]));
                    *reinterpret_cast<uint2*>(st + T3_PLANE_BYTES + off) = make_uint2(pack_hi(m[0], m[1]), pack_hi(m[2], m[3]));
                    *reinterpret_cast<uint2*>(st + 2 * T3_PLANE_BYTES + off) = make_uint2(pack_hi(l[0], l[1]), pack_hi(l[2], l[3]));
                } else {
                    uint32_t h01, h23, l01, l23;
                    split2(v[i].x, v[i].y, h01, l01);
                    split2(v[i].z, v[i].w, h23, l23);
                    *reinterpret_cast<uint2*>(st + off) = make_uint2(h01, h23);
                    *reinterpret_cast<uint2*>(st + T3_PLANE_BYTES + off) = make_uint2(l01, l23);
                }
            }
            fence_proxy_async();                 // generic-proxy smem writes -> visible to the tensor core (async proxy)
            mbar_arrive(&bars->full[stage]);
            if (++stage == p.nstage) { stage = 0; phase ^= 1; }
        };
        if (ATMA) {
            // The fp32 k-blocks arrive by TMA (warp 15) in a ring of [128 rows][64 floats] boxes: the producers read their eight
            // granules from shared memory (a warp instruction covers two whole 256-byte rows: conflict-free), hand the box back
            // and split as before.  No global load is issued by these warps: their loads used to be the first consumer to wait
            // (10 % of the kernel's stall samples sat on the F2FP that follows the LDG batch).
            uint32_t rs = 0, rph = 0;
            for (long long q = 0; q < q_end; ++q) {
                float4 v[8];
                MBAR_WAIT_AT(&bars->raw_full[rs], rph);
                const float* box = a_raw + (size_t)rs * RAW_BOX + col4 * 4;
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = *reinterpret_cast<const float4*>(box + (row0 + 16 * i) * T3_BK);
                fence_proxy_async();                 // the reads are ordered before the copy engine's refill of the box
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->raw_empty[rs]);
                if (++rs == (uint32_t)p.raw_slots) { rs = 0; rph ^= 1u; }
                emit(v);
            }
        } else {
        float4 va[8], vb[8];
        fetch(va);
        for (long long q = 0; q < q_end; q += 2) {
            fetch(vb); emit(va);
            fetch(va); emit(vb);
        }
        }
    } else if (warp >= 12) {
      // warpgroup 3 (MMA issue, weight loader, two idle warps) hands registers over to the epilogue warpgroup
      asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" :: "n"(T3_REGS_MISC));
      if (warp == T3_MMA_WARP) {
        // ================================================================ MMA issue (+ resident-weight TMA)
        if (lane == 0) {
            const uint32_t idesc = PL == 3 ? umma_idesc_bf16(T3_BM, (uint32_t)p.BN) : umma_idesc_f16(T3_BM, (uint32_t)p.BN);
            int stage = 0; uint32_t phase = 0, acc = 0, acc_phase = 0, w_loads = 0, ws_slot = 0, ws_phase = 0;
            long long cur_g = -1;
            for (long long it = item_lo; it < item_hi; ++it) {
                const long long g = it / p.MT;
                const int node = (int)g;
                if (!as_mode && g != cur_g) {
                    // the previous group's MMAs (w_empty phase w_loads-1) must have retired before the tile is overwritten
                    if (w_loads > 0) MBAR_WAIT_AT(&bars->w_empty, (w_loads - 1) & 1u);
                    mbar_arrive_expect_tx(&bars->w_full, (uint32_t)PL * (uint32_t)p.KB * w_block);
                    for (int pl = 0; pl < PL; ++pl)
                        for (int kb = 0; kb < p.KB; ++kb)
                            tma_load_3d(w_smem + (size_t)(pl * p.KB + kb) * w_block, &map_w, &bars->w_full, p.k_base + kb * T3_BK, my_nt * p.BN,
                                        pl * p.n_types + p.types.t[node]);
                    MBAR_WAIT_AT(&bars->w_full, w_loads & 1u);
                    ++w_loads;
                    cur_g = g;
                }
                const int stage0 = stage; const uint32_t phase0 = phase;
                for (int nt = nt_lo; nt < nt_hi; ++nt) {
                    stage = stage0; phase = phase0;             // as_mode: every n-tile re-reads the same KB stages
                    MBAR_WAIT_AT(&bars->acc_empty[acc], acc_phase ^ 1);
                    tc_fence_after();
                    // two accumulators per tile: x0 w0 alone in "main", the five 2^-8 .. 2^-16 pairs in "corr" (see header)
                    const uint32_t d_main = tmem_base + acc * 2u * (uint32_t)p.BN, d_corr = d_main + (uint32_t)p.BN;
                    for (int kb = 0; kb < p.KB; ++kb) {
                        const uint8_t* w_tile = w_smem;         // plane pw of this k-block at w_tile + pw * w_plane_stride
                        uint32_t w_plane_stride = (uint32_t)p.KB * w_block;
                        if (as_mode) {
                            MBAR_WAIT_AT(&bars->ws_full[ws_slot], ws_phase);
                            w_tile = w_smem + (size_t)ws_slot * PL * w_block; w_plane_stride = w_block;
                        } else {
                            w_tile = w_smem + (size_t)kb * w_block;
                        }
                        if (nt == nt_lo) MBAR_WAIT_AT(&bars->full[stage], phase);
                        tc_fence_after();
                        const uint32_t a_base = smem_u32(a_smem + (size_t)stage * STAGE_BYTES);
                        uint32_t first_main = (kb == 0) ? 1u : 0u, first_corr = first_main;
#pragma unroll
                        for (int pa = 0; pa < PL; ++pa) {
#pragma unroll
                            for (int pw = 0; pw < PL; ++pw) {
                                if (pa + pw > PL - 1) continue;
                                const uint64_t adesc = umma_desc_sw128(a_base + (uint32_t)pa * T3_PLANE_BYTES);
                                const uint64_t bdesc = umma_desc_sw128(smem_u32(w_tile + (size_t)pw * w_plane_stride));
                                const bool main_pair = (pa + pw == 0);
#pragma unroll
                                for (int k = 0; k < T3_BK / 16; ++k) {
                                    if (kb * T3_BK + 16 * k >= p.K) continue;        // zero tail of a partly filled last k-block (K = 96)
                                    uint32_t& first = main_pair ? first_main : first_corr;
                                    umma_bf16(main_pair ? d_main : d_corr, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, first ? 0u : 1u);
                                    first = 0u;
                                }
                            }
                        }
                        if (as_mode) {
                            umma_commit(&bars->ws_empty[ws_slot]);
                            if (++ws_slot == (uint32_t)p.wslots) { ws_slot = 0; ws_phase ^= 1u; }
                        }
                        if (nt == nt_hi - 1) umma_commit(&bars->empty[stage]);     // the planes of this k-block are no longer needed
                        if (++stage == p.nstage) { stage = 0; phase ^= 1; }
                    }
                    umma_commit(&bars->acc_full[acc]);
                    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                }
                const bool last_of_group = (it + 1 == item_hi) || ((it + 1) / p.MT != g);
                if (!as_mode && last_of_group) umma_commit(&bars->w_empty);
            }
        }
        __syncwarp();
      } else if (warp == T3_WLOAD_WARP) {
        // ================================================================ weight-slot TMA ring (activation-stationary mode)
        if (as_mode && lane == 0) {
            uint32_t slot = 0, ph = 0;
            for (long long it = item_lo; it < item_hi; ++it) {
                const int type = p.types.t[(int)(it / p.MT)];
                for (int nt = 0; nt < p.NT; ++nt)
                    for (int kb = 0; kb < p.KB; ++kb) {
                        MBAR_WAIT_AT(&bars->ws_empty[slot], ph ^ 1u);
                        mbar_arrive_expect_tx(&bars->ws_full[slot], (uint32_t)PL * w_block);
                        for (int pl = 0; pl < PL; ++pl)
                            tma_load_3d(w_smem + (size_t)(slot * PL + pl) * w_block, &map_w, &bars->ws_full[slot], p.k_base + kb * T3_BK, nt * p.BN,
                                        pl * p.n_types + type);
                        if (++slot == (uint32_t)p.wslots) { slot = 0; ph ^= 1u; }
                    }
            }
        }
        __syncwarp();
      } else if (warp == T3_WLOAD_WARP + 2) {
        // ================================================================ fp32 activation ring (ATMA): one [128 samples][64 k] box per
        // (tile, k-block) in the producers' order; columns beyond K are zero-filled by the tensor map (K = 96: half of the 2nd block)
        if (ATMA && lane == 0) {
            tma_prefetch_desc(&map_a);
            uint32_t rs = 0, rph = 0;
            for (long long it = item_lo; it < item_hi; ++it) {
                const int node = (int)(it / p.MT), mt = (int)(it % p.MT);
                for (int kb = 0; kb < p.KB; ++kb) {
                    MBAR_WAIT_AT(&bars->raw_empty[rs], rph ^ 1u);
                    mbar_arrive_expect_tx(&bars->raw_full[rs], (uint32_t)RAW_BOX * 4u);
                    tma_load_3d(a_raw + (size_t)rs * RAW_BOX, &map_a, &bars->raw_full[rs], kb * T3_BK, node, mt * T3_BM);
                    if (++rs == (uint32_t)p.raw_slots) { rs = 0; rph ^= 1u; }
                }
            }
        }
        __syncwarp();
      } else if (warp == T3_WLOAD_WARP + 1) {
        // ================================================================ residual chunk ring (res_tma): one TMA box of
        // [128 samples][32 columns] per epilogue chunk, in the epilogue's order, two boxes ahead of it.  The epilogue's own LDGs kept
        // 16 KB per SM in flight at best; a box is 16 KB and two are outstanding while a third is being consumed.
        if (RTMA && lane == 0) {
            tma_prefetch_desc(&map_r);
            uint32_t slot = 0, ph = 0;
            if (ACT == T3_ACT_GRU) tma_prefetch_desc(&map_a);
            auto box = [&](const CUtensorMap* map, int col, int node, int row) {
                MBAR_WAIT_AT(&bars->res_empty[slot], ph ^ 1u);
                mbar_arrive_expect_tx(&bars->res_full[slot], (uint32_t)RES_BOX * 4u);
                tma_load_3d(res_buf + (size_t)slot * RES_BOX, map, &bars->res_full[slot], col, node, row);
                if (++slot == (uint32_t)p.res_slots) { slot = 0; ph ^= 1u; }
            };
            for (long long it = item_lo; it < item_hi; ++it) {
                const int node = (int)(it / p.MT), mt = (int)(it % p.MT);
                for (int nt = nt_lo; nt < nt_hi; ++nt) {
                    for (int c0 = 0; c0 < p.BN; c0 += 32) box(&map_r, nt * p.BN + c0, node, mt * T3_BM);
                    // fused GRU step: after the three x-side gate boxes of the n-tile's 32 units, their previous hidden state
                    if (ACT == T3_ACT_GRU) box(&map_a, nt * 32, node, mt * T3_BM);
                }
            }
        }
        __syncwarp();
      }
    } else {
        // ================================================================ epilogue (warps 8-11: warp % 4 = lane quarter)
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" :: "n"(T3_REGS_EPI));
        const int quarter = warp & 3;
        const int et = threadIdx.x - T3_EPI_WARP0 * 32;
        // Chunk geometry.  CW columns of the tile leave TMEM per step; LPR lanes cover the CW * 4 contiguous bytes of a row, so one
        // LDG / STG.128 of the warp touches RPI rows and J instructions cover the 32 rows of the warp's lane quarter.
        // CW = 32 (two planes: shared memory to spare): 128 contiguous bytes per row = whole L2 lines; CW = 16 (three planes): 64.
        constexpr int CW = t3_chunk_cols(PL), LPR = CW / 4, RPI = 32 / LPR, J = 32 / RPI, AHEAD = (CW == 32) ? 2 : T3_RES_AHEAD;
        float* stg0 = epi_stage + quarter * (32 * CW);
        float* stg = stg0;
        uint32_t stg_flip = 0;
        const int tr = lane / LPR, tcl = lane % LPR;
        uint32_t acc = 0, acc_phase = 0;
        uint32_t r_slot = 0, r_phase = 0;                          // residual chunk ring (res_tma)
        constexpr bool res_tma = RTMA;
        float nss[J];                                              // norm_out: sum of squares of this thread's 4 columns of rows tr + RPI j
        long long cur_key = -1;
        // tables for every column this CTA produces for a node (always in the weight-resident schedule: one n-tile per CTA; in the
        // activation-stationary one when the host sized them for OUT columns: two-plane kernel), else per (node, n-tile)
        const bool tab_node = !as_mode || p.tab_cols >= p.NT * p.BN;
        for (long long it = item_lo; it < item_hi; ++it)
        for (int nt = nt_lo; nt < nt_hi; ++nt) {
            const int mt = (int)(it % p.MT);
            const int node = (int)(it / p.MT);
            const int o0 = nt * p.BN;
            if constexpr (NORM) {
                if (nt == nt_lo) {
#pragma unroll
                    for (int j = 0; j < J; ++j) nss[j] = 0.0f;
                }
            }
            // The epilogue tables change with the NODE only (a CTA walks the m-tiles of a node back to back): refilling them per
            // (node, n-tile) put two named barriers and a global-load latency in front of every n-tile of the activation-stationary
            // schedule (18 % of the epilogue warps' samples in the round-2 source-level capture).
            const long long key = tab_node ? node : (long long)node * p.NT + nt;
            const int tab0 = tab_node ? nt_lo * p.BN : o0, tab_n = tab_node ? (nt_hi - nt_lo) * p.BN : p.BN;
            if (key != cur_key) {
                asm volatile("bar.sync 1, 128;" ::: "memory");
                for (int c = et; c < tab_n; c += 128) {
                    const int o = tab0 + c;
                    if (ACT == T3_ACT_GRU) {                    // gate biases of the n-tile's columns: h side / x side
                        epi_mul[c] = __ldg(p.gru_bias_h + (long long)node * p.OUT + o);
                        epi_add[c] = __ldg(p.gru_bias_x + (long long)node * p.OUT + o);
                        continue;
                    }
                    const float mul = p.ss ? (__ldg(p.ss + o) + 1.0f) : 1.0f;
                    const float bias = p.bias_node ? __ldg(p.bias_node + (long long)node * p.OUT + o) : 0.0f;
                    epi_mul[c] = mul;
                    epi_add[c] = fmaf(bias, mul, p.ss ? __ldg(p.ss + p.OUT + o) : 0.0f);
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                cur_key = key;
            }
            // TMEM hands each lane one ROW of the tile; stored that way every LDG/STG.128 of a warp would touch 32
            // different rows (32 LSU wavefronts per instruction: in the second ncu capture the residual alone cost 140 us).
            // Each CW-column chunk is therefore transposed through a swizzled staging tile: afterwards LPR lanes cover the
            // contiguous bytes of a row and one instruction touches RPI rows.
            const int b_own = mt * T3_BM + quarter * 32 + lane;
            const float rs = (b_own < p.B && p.row_scale) ? __ldg(p.row_scale + (long long)b_own * p.N + node) : 1.0f;
            const int bT0 = mt * T3_BM + quarter * 32 + tr;                  // transposed mapping: rows bT0 + RPI j
            const float* res_base = nullptr;
            if (HAS_RES) res_base = p.residual.ptr + (long long)node * p.residual.sn + o0 + 4 * tcl;
            float* out_base = p.out.ptr + (long long)node * p.out.sn + o0 + 4 * tcl;
            long long res_off[J];
            float4 rr[AHEAD][J];                                // residual of the next AHEAD chunks (in flight)
            if (HAS_RES && !res_tma) {
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    const int bj = bT0 + RPI * j;
                    res_off[j] = (long long)(p.residual.rep == 1 ? bj : bj / p.residual.rep) * p.residual.sb;
                }
#pragma unroll
                for (int a = 0; a < AHEAD; ++a)
#pragma unroll
                    for (int j = 0; j < J; ++j)
                        if (bT0 + RPI * j < p.B && CW * a < p.BN) rr[a][j] = __ldg(reinterpret_cast<const float4*>(res_base + res_off[j] + CW * a));
            }
            if (HAS_RES && !res_tma && (nt + 1 < nt_hi || it + 1 < item_hi)) {
                // The NEXT pass's residual rows (one per thread) are pulled into L2 while this one is processed.
                const long long it2 = nt + 1 < nt_hi ? it : it + 1;
                const int o2 = (nt + 1 < nt_hi ? nt + 1 : nt_lo) * p.BN;
                const int b2 = (int)(it2 % p.MT) * T3_BM + quarter * 32 + lane, node2 = (int)(it2 / p.MT);
                if (b2 < p.B)
                    prefetch_l2_bulk(p.residual.ptr + (long long)(p.residual.rep == 1 ? b2 : b2 / p.residual.rep) * p.residual.sb +
                                     (long long)node2 * p.residual.sn + o2, (uint32_t)p.BN * 4u);
            }
            MBAR_WAIT_AT(&bars->acc_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * 2u * (uint32_t)p.BN;
            const float* pre_base = p.pre.ptr ? p.pre.ptr + (long long)node * p.pre.sn + o0 + 4 * tcl : nullptr;
            constexpr float CS = PL == 3 ? 1.0f : 0.00048828125f;        // the lo planes carry a factor 2^11
            if constexpr (ACT == T3_ACT_GRU && PL == 2) {
                // ---------------------------------------------------------------- fused GRU step (recurrent.py:351-358), identity influence.
                // The weight rows are gate-interleaved (sd_gru::W_hh_perm order): the 96 columns of n-tile nt are the r | z | n gate
                // products of hidden units [32 nt, 32 nt + 32) of this node, so one n-tile's accumulators hold everything the new
                // hidden state of those units needs.  The x-side products of the three gates (loop-invariant, same column order)
                // and the previous hidden state arrive as four [128 samples][32 columns] boxes through the residual ring; the new
                // state leaves as one bulk tensor store.  hr (619 MB per frame at B = 25 600) is neither written nor read back.
                float4 gr[J], gz[J];
#pragma unroll
                for (int gch = 0; gch < 3; ++gch) {
                    const int c0 = 32 * gch;
                    if (gch == 0 && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the last store has read the tile
                    __syncwarp();
                    t3_chunk32_to_stage(t_row + (uint32_t)c0, t_row + (uint32_t)(p.BN + c0), stg, lane, 1.0f, p.merge_ld != 0);
                    __syncwarp();
                    const float4 bh = *reinterpret_cast<const float4*>(epi_mul + (o0 - tab0) + c0 + 4 * tcl);
                    const float4 bx = *reinterpret_cast<const float4*>(epi_add + (o0 - tab0) + c0 + 4 * tcl);
                    MBAR_WAIT_AT(&bars->res_full[r_slot], r_phase);        // x-side product of this gate
                    const float* xb = res_buf + (size_t)r_slot * RES_BOX + (quarter * 32 + tr) * 32 + 4 * tcl;
                    float4 o[J];
#pragma unroll
                    for (int j = 0; j < J; ++j) {
                        const int row = tr + RPI * j;
                        float4 hx = *reinterpret_cast<const float4*>(stg + row * 32 + 4 * (tcl ^ (row & 7)));
                        float4 xg = *reinterpret_cast<const float4*>(xb + RPI * j * 32);
                        hx.x += bh.x; hx.y += bh.y; hx.z += bh.z; hx.w += bh.w;
                        xg.x += bx.x; xg.y += bx.y; xg.z += bx.z; xg.w += bx.w;
                        if (gch == 0) {
                            gr[j].x = sigmoid_ex2(xg.x + hx.x); gr[j].y = sigmoid_ex2(xg.y + hx.y);
                            gr[j].z = sigmoid_ex2(xg.z + hx.z); gr[j].w = sigmoid_ex2(xg.w + hx.w);
                        } else if (gch == 1) {
                            gz[j].x = sigmoid_ex2(xg.x + hx.x); gz[j].y = sigmoid_ex2(xg.y + hx.y);
                            gz[j].z = sigmoid_ex2(xg.z + hx.z); gz[j].w = sigmoid_ex2(xg.w + hx.w);
                        } else {
                            o[j].x = tanh_ex2(fmaf(gr[j].x, hx.x, xg.x)); o[j].y = tanh_ex2(fmaf(gr[j].y, hx.y, xg.y));
                            o[j].z = tanh_ex2(fmaf(gr[j].z, hx.z, xg.z)); o[j].w = tanh_ex2(fmaf(gr[j].w, hx.w, xg.w));
                        }
                    }
                    fence_proxy_async();                            // the box reads are ordered before the copy engine's refill
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars->res_empty[r_slot]);
                    if (++r_slot == (uint32_t)p.res_slots) { r_slot = 0; r_phase ^= 1u; }
                    if (gch == 2) {
                        MBAR_WAIT_AT(&bars->res_full[r_slot], r_phase);    // previous hidden state of these units
                        const float* hb = res_buf + (size_t)r_slot * RES_BOX + (quarter * 32 + tr) * 32 + 4 * tcl;
#pragma unroll
                        for (int j = 0; j < J; ++j) {
                            const float4 hp = *reinterpret_cast<const float4*>(hb + RPI * j * 32);
                            o[j].x = fmaf(gz[j].x, hp.x, fmaf(-o[j].x, gz[j].x, o[j].x));      // n - n z + z h
                            o[j].y = fmaf(gz[j].y, hp.y, fmaf(-o[j].y, gz[j].y, o[j].y));
                            o[j].z = fmaf(gz[j].z, hp.z, fmaf(-o[j].z, gz[j].z, o[j].z));
                            o[j].w = fmaf(gz[j].w, hp.w, fmaf(-o[j].w, gz[j].w, o[j].w));
                        }
                        fence_proxy_async();
                        __syncwarp();                               // every lane has read its transposed values and its box rows
                        if (lane == 0) mbar_arrive(&bars->res_empty[r_slot]);
                        if (++r_slot == (uint32_t)p.res_slots) { r_slot = 0; r_phase ^= 1u; }
#pragma unroll
                        for (int j = 0; j < J; ++j) *reinterpret_cast<float4*>(stg + (tr + RPI * j) * 32 + 4 * tcl) = o[j];
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) {
                            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                                         :: "l"(&map_o), "r"(smem_u32(stg)), "r"(nt * 32), "r"(node), "r"(mt * T3_BM + quarter * 32) : "memory");
                            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        }
                    }
                }
            } else
            for (int c0 = 0; c0 < p.BN; c0 += CW) {
                float4 pp[J];                                   // partial product of the first K segment (K-split layers)
                if (pre_base) {
#pragma unroll
                    for (int j = 0; j < J; ++j)
                        pp[j] = (bT0 + RPI * j < p.B) ? __ldg(reinterpret_cast<const float4*>(pre_base + (long long)(bT0 + RPI * j) * p.pre.sb + c0)) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                if (PL == 2 && p.out_tma) {                    // the bulk store that last used this staging tile has read it
                    if (p.stg2) {
                        stg = stg0 + (stg_flip ? 4 * 32 * CW : 0);
                        stg_flip ^= 1u;
                        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    } else {
                        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    }
                }
                __syncwarp();                                   // the previous chunk has been read out of the staging tile
                if constexpr (PL == 2) {
                    t3_chunk32_to_stage(t_row + (uint32_t)c0, t_row + (uint32_t)(p.BN + c0), stg, lane, rs, p.merge_ld != 0);
                    if (p.direct) {
                        // Bare layer (raw products of a layer with a dense graph influence, to_qkv, first half of a K-split): nothing
                        // is left to do per column, and the tile just written - row = lane, 16-byte chunk q at position q ^ (row & 7),
                        // 1024-byte aligned - is exactly what a SWIZZLE_128B bulk tensor store reads.  No transposition, no second pass.
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) {
                            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                                         :: "l"(&map_o), "r"(smem_u32(stg)), "r"(o0 + c0), "r"(node), "r"(mt * T3_BM + quarter * 32) : "memory");
                            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        }
                        continue;
                    }
                } else
#pragma unroll
                for (int hf = 0; hf < CW / 16; ++hf) {              // 16 columns at a time: v + vc stay within 32 registers
                    uint32_t v[16], vc[16];
                    tmem_ld_32x16(t_row + (uint32_t)(c0 + 16 * hf), v);
                    tmem_ld_32x16(t_row + (uint32_t)(p.BN + c0 + 16 * hf), vc);
                    tmem_ld_wait();
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float4 x;                               // main + corr: one round-to-nearest add (FMA for PL = 2), then the row scale
                        x.x = fmaf(__uint_as_float(vc[4 * q + 0]), CS, __uint_as_float(v[4 * q + 0])) * rs;
                        x.y = fmaf(__uint_as_float(vc[4 * q + 1]), CS, __uint_as_float(v[4 * q + 1])) * rs;
                        x.z = fmaf(__uint_as_float(vc[4 * q + 2]), CS, __uint_as_float(v[4 * q + 2])) * rs;
                        x.w = fmaf(__uint_as_float(vc[4 * q + 3]), CS, __uint_as_float(v[4 * q + 3])) * rs;
                        // row = lane; the 16-byte chunk qq of the row is stored at position qq ^ f(row) (bank-conflict free both ways)
                        const int qq = 4 * hf + q;
                        const int sw = (CW == 32) ? (lane & 7) : ((lane >> 1) & 3);
                        *reinterpret_cast<float4*>(stg + lane * CW + 4 * (qq ^ sw)) = x;
                    }
                }
                __syncwarp();
                const float4 m4 = *reinterpret_cast<const float4*>(epi_mul + (o0 - tab0) + c0 + 4 * tcl);
                const float4 a4 = *reinterpret_cast<const float4*>(epi_add + (o0 - tab0) + c0 + 4 * tcl);
                float4 o[J];
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    const int row = tr + RPI * j;
                    const int sw = (CW == 32) ? (row & 7) : ((row >> 1) & 3);
                    float4 x = *reinterpret_cast<const float4*>(stg + row * CW + 4 * (tcl ^ sw));
                    if (pre_base) { x.x += pp[j].x; x.y += pp[j].y; x.z += pp[j].z; x.w += pp[j].w; }
                    o[j].x = fmaf(x.x, m4.x, a4.x); o[j].y = fmaf(x.y, m4.y, a4.y);
                    o[j].z = fmaf(x.z, m4.z, a4.z); o[j].w = fmaf(x.w, m4.w, a4.w);
                    if (ACT == SD_ACT_TANH) { o[j].x = t3_tanh<FAST>(o[j].x); o[j].y = t3_tanh<FAST>(o[j].y); o[j].z = t3_tanh<FAST>(o[j].z); o[j].w = t3_tanh<FAST>(o[j].w); }
                    if (ACT == SD_ACT_TANH_TANH) {
                        o[j].x = t3_tanh<FAST>(t3_tanh<FAST>(o[j].x)); o[j].y = t3_tanh<FAST>(t3_tanh<FAST>(o[j].y));
                        o[j].z = t3_tanh<FAST>(t3_tanh<FAST>(o[j].z)); o[j].w = t3_tanh<FAST>(t3_tanh<FAST>(o[j].w));
                    }
                    if (HAS_RES && !res_tma) { o[j].x += rr[0][j].x; o[j].y += rr[0][j].y; o[j].z += rr[0][j].z; o[j].w += rr[0][j].w; }
                }
                if (res_tma) {                                  // this chunk's residual box: rows of the warp's lane quarter
                    MBAR_WAIT_AT(&bars->res_full[r_slot], r_phase);
                    const float* rb = res_buf + (size_t)r_slot * RES_BOX + (quarter * 32 + tr) * 32 + 4 * tcl;
#pragma unroll
                    for (int j = 0; j < J; ++j) {
                        const float4 r4 = *reinterpret_cast<const float4*>(rb + RPI * j * 32);
                        o[j].x += r4.x; o[j].y += r4.y; o[j].z += r4.z; o[j].w += r4.w;
                        if constexpr (NORM) nss[j] = fmaf(o[j].x, o[j].x, fmaf(o[j].y, o[j].y, fmaf(o[j].z, o[j].z, fmaf(o[j].w, o[j].w, nss[j]))));
                    }
                    fence_proxy_async();                        // the reads above are ordered before the copy engine's refill
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars->res_empty[r_slot]);
                    if (++r_slot == (uint32_t)p.res_slots) { r_slot = 0; r_phase ^= 1u; }
                }
                if (HAS_RES && !res_tma) {                      // rotate the ring, fetch the chunk AHEAD positions ahead
#pragma unroll
                    for (int a = 0; a + 1 < AHEAD; ++a)
#pragma unroll
                        for (int j = 0; j < J; ++j) rr[a][j] = rr[a + 1][j];
                    if (c0 + CW * AHEAD < p.BN) {
#pragma unroll
                        for (int j = 0; j < J; ++j)
                            if (bT0 + RPI * j < p.B)
                                rr[AHEAD - 1][j] = __ldg(reinterpret_cast<const float4*>(res_base + res_off[j] + c0 + CW * AHEAD));
                    }
                }
                if (PL == 2 && p.out_tma) {
                    // Output through the copy engine: the results go back into the (now free) staging tile in plain row-major order
                    // (a warp instruction writes 4 rows x 128 contiguous bytes: conflict-free) and ONE bulk tensor store moves the
                    // [32 rows][32 columns] box to the 3-D output tensor (columns, node, sample); rows beyond the batch are clipped
                    // by the tensor map.  The epilogue warps issue no global stores at all.
                    __syncwarp();                               // every lane has read its transposed values
#pragma unroll
                    for (int j = 0; j < J; ++j) *reinterpret_cast<float4*>(stg + (tr + RPI * j) * CW + 4 * tcl) = o[j];
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                                     :: "l"(&map_o), "r"(smem_u32(stg)), "r"(o0 + c0), "r"(node), "r"(mt * T3_BM + quarter * 32) : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < J; ++j) {
                        const int bj = bT0 + RPI * j;
                        if (bj < p.B) *reinterpret_cast<float4*>(out_base + (long long)bj * p.out.sb + c0) = o[j];
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&bars->acc_empty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            if constexpr (NORM) if (nt == nt_hi - 1) {
                // Row norms of the finished rows (the CTA has produced every column of them): the LPR lanes that share a row fold their
                // partial sums, the first of them writes the factor.  Saves the separate 413 MB pass over the output (3.7 % of a step).
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    float v = nss[j];
#pragma unroll
                    for (int off = LPR / 2; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
                    const int bj = bT0 + RPI * j;
                    if (tcl == 0 && bj < p.B) p.norm_out[(long long)bj * p.N + node] = 1.0f / fmaxf(sqrtf(v), 1e-12f);
                }
            }
        }
        if (PL == 2 && p.out_tma && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");    // shared memory outlives the stores
    }
    tc_fence_before();
    __syncthreads();
    if (warp == T3_ALLOC_WARP) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// ------------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn3)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// tab_cols: columns of the two epilogue tables (the n-tile of a weight-resident CTA, the whole OUT of an activation-stationary one)
static int t3_tab_cols(int cols) { return (cols + 15) & ~15; }
static size_t t3_misc_smem(int tab_cols, int pl) { return 2 * (size_t)t3_tab_cols(tab_cols) * 4 + 4 * 32 * t3_chunk_cols(pl) * 4 + sizeof(T3Barriers) + 1024; }
static int t3_kb(int K) { return (K + T3_BK - 1) / T3_BK; }      // the last k-block may be partly filled (zero planes)
static size_t t3_fixed_smem(int K, int bn, int pl) { return (size_t)pl * t3_kb(K) * bn * 128 + t3_misc_smem(bn, pl); }
static int t3_env(const char* name, int dflt) { const char* e = getenv(name); return e && e[0] ? atoi(e) : dflt; }
// activation-stationary mode: K/64 (+ extra) plane stages + the weight slots; widest n-tile that fits (>= 64 columns).
// Three planes (48 KB stages): exactly K/64 stages and two slots fill the 227 KB.  Two planes (32 KB stages): the same K/64
// stages + two slots need 155 KB (K = 192) / 187 KB (K = 256).  A third or fourth weight slot and look-ahead stages of the next
// m-tile were measured (SKELDIFF_T3_WSLOTS / SKELDIFF_T3_XSTAGES): every configuration above 196 KB -- the last shared-memory
// carve-out step that leaves more than 28 KB of L1 to the producers' and the epilogue's global loads -- was 12 % SLOWER
// (192 -> 192 bare: 235 vs 207 us, K = 256: 410 vs 374 us), so the smallest configuration is the default.
struct T3AsCfg { int bn, wslots, nstage; };
static T3AsCfg t3_as_cfg(int K, int OUT, int pl) {
    const int kb = t3_kb(K);
    const int cands[] = {128, 96, 64};
    static const int want_slots = t3_env("SKELDIFF_T3_WSLOTS", 2), want_extra = t3_env("SKELDIFF_T3_XSTAGES", 0);
    for (int bn : cands) {
        if (OUT % bn || OUT / bn < 2) continue;
        auto fits = [&](int stages, int slots) {
            return (size_t)stages * t3_stage_bytes(pl) + (size_t)slots * pl * bn * 128 + t3_misc_smem(pl == 2 ? OUT : bn, pl) <= 227 * 1024;
        };
        if (pl == 3) { if (kb <= T3_MAX_STAGES && fits(kb, 2)) return {bn, 2, kb}; continue; }
        for (int slots = want_slots > T3_MAX_WSLOTS ? T3_MAX_WSLOTS : want_slots; slots >= 2; --slots)
            for (int extra = want_extra; extra >= 0; --extra)
                if (kb + extra <= T3_MAX_STAGES && fits(kb + extra, slots)) return {bn, slots, kb + extra};
    }
    return {0, 0, 0};
}
static int t3_stages(int K, int bn, int pl) {
    const size_t budget = 227 * 1024, fixed = t3_fixed_smem(K, bn, pl);
    if (fixed + 2 * (size_t)t3_stage_bytes(pl) > budget) return 0;
    const size_t n = (budget - fixed) / t3_stage_bytes(pl);
    return (int)(n > 4 ? 4 : n);
}
static int t3_pick_bn(int K, int OUT, int pl) {
    const int cands[] = {128, 96, 64, 32};    // 4 accumulators of BN fp32 columns must fit the 512 TMEM columns
    for (int bn : cands) if (OUT % bn == 0 && t3_stages(K, bn, pl) >= 2) return bn;
    return 0;
}

bool glin_tc3_supported(int K0, int K1, int OUT) {
    // segment widths: multiples of 4 (float4 granules); the weight rows (K bf16) must be 16-byte multiples for the TMA map
    if (K0 <= 0 || K0 % 4 || K1 % 4 || (K0 + K1) % 8) return false;
    return t3_pick_bn(K0 + K1, OUT, 3) != 0;     // what fits three planes fits two
}

// operand split of the fp32-grade tensor-core path for the calls of this thread: 3 = bf16 planes, 2 = fp16 planes
// (set by the C entry points from the precision argument: SD_PREC_BF16X3 / SD_PREC_F16X2)
static thread_local bool tl_norm_written = false;
bool tc3_take_norm_written() { const bool w = tl_norm_written; tl_norm_written = false; return w; }
static thread_local int tl_split_planes = 3;
int tc_split_planes() { return tl_split_planes; }
void set_tc_split_planes(int planes) { tl_split_planes = planes == 2 ? 2 : 3; }

template <int ACT, bool HAS_RES, int PL, bool FAST = false, bool RTMA = false, bool ATMA = false, bool NORM = false>
static int t3_launch_t(const CUtensorMap& mw, const CUtensorMap& mr, const CUtensorMap& mo, const CUtensorMap& ma, const T3Params& p, int grid, size_t smem, cudaStream_t st) {
    // the libdevice epilogue is kept for the three-plane kernel only (SKELDIFF_ACCURATE_EPILOGUE=1)
    if (ACT != SD_ACT_NONE && !FAST && (PL == 2 || fast_epilogue())) return t3_launch_t<ACT, HAS_RES, PL, ACT != SD_ACT_NONE, RTMA, ATMA, NORM>(mw, mr, mo, ma, p, grid, smem, st);
    if (HAS_RES && PL == 2 && !RTMA && p.res_tma) return t3_launch_t<ACT, HAS_RES, PL, FAST, HAS_RES && PL == 2, ATMA, NORM>(mw, mr, mo, ma, p, grid, smem, st);
    if (!HAS_RES && PL == 2 && !ATMA && p.raw_slots) return t3_launch_t<ACT, HAS_RES, PL, FAST, RTMA, !HAS_RES && PL == 2, NORM>(mw, mr, mo, ma, p, grid, smem, st);
    // row norms from the epilogue: one extra instantiation, tanh + residual through the ring on the two-plane kernel (what t3_launch_one admits)
    constexpr bool NORM_OK = ACT == SD_ACT_TANH && HAS_RES && PL == 2 && FAST && RTMA && !ATMA;
    if (NORM_OK && !NORM && p.norm_out) return t3_launch_t<ACT, HAS_RES, PL, FAST, RTMA, ATMA, NORM_OK>(mw, mr, mo, ma, p, grid, smem, st);
    if (!NORM && p.norm_out) { set_error("glin_tc3: row norms requested from a kernel variant that does not compute them"); return SD_ERR_INVALID; }
    auto kern = glin_tc3_kernel<ACT, HAS_RES, FAST, PL, RTMA, ATMA, NORM>;
    static unsigned long long configured = 0;      // bit d: attribute set on device d (it is per device)
    if (int rc_attr = opt_in_smem(kern, (size_t)((227 * 1024)), configured)) return rc_attr;
    kern<<<grid, T3_THREADS, smem, st>>>(mw, mr, mo, ma, p);
    SD_LAUNCH_OK("glin_tc3_kernel");
    return SD_OK;
}

// out = epilogue(A @ W^T) with fp32 views; the caller guarantees G == identity for this call
// one launch over the weight columns [k_base, k_base + K0 + K1) of the layer; `pre` (optional) is added before the epilogue
// gru (optional): fused GRU step.  L is W_hh in the gate-interleaved row order viewed as a graph-linear H -> 3H, c.a0 the previous
// hidden state, `out` the NEW hidden state [B, N, H]; xr holds the x-side products in the same column order.
struct T3Gru { View xr; const float* bias_x; const float* bias_h; };
static int t3_launch_one(const sd_glin* L, const GlinCall& c, const ViewW& out, bool apply_epilogue, int k_base, const View* pre, cudaStream_t st,
                         const T3Gru* gru = nullptr) {
    if (!L->W_bf16 || L->planes != 3) { set_error("bf16x3 path: 3-plane weights not set on this layer"); return SD_ERR_INVALID; }
    const int PL = (tl_split_planes == 2 && L->W_f16) ? 2 : 3;
    const int K0 = c.a0.width, K1 = c.a1.ptr ? c.a1.width : 0;
    const int Kuse = K0 + K1;
    if (k_base < 0 || k_base % 8 || k_base + Kuse > L->K || !glin_tc3_supported(K0, K1, L->OUT)) { set_error("bf16x3 path: unsupported shape K=%d+%d (base %d of %d) OUT=%d", K0, K1, k_base, L->K, L->OUT); return SD_ERR_UNSUPPORTED; }
    if (c.epi.ss_row_idx && apply_epilogue) { set_error("bf16x3 path: per-sample time rows are only supported on the FFMA path"); return SD_ERR_UNSUPPORTED; }
    if (c.B <= 0) return SD_OK;
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
    auto view_ok = [&](const View& v) { return v.ptr == nullptr || (al16(v.ptr) && v.sb % 4 == 0 && v.sn % 4 == 0); };
    if (!view_ok(c.a0) || !view_ok(c.a1) || !al16(out.ptr) || out.sb % 4 || out.sn % 4 ||
        (apply_epilogue && c.epi.residual.ptr && !view_ok(c.epi.residual))) {
        set_error("bf16x3 path: operands must be 16-byte aligned");
        return SD_ERR_UNSUPPORTED;
    }
    T3Params p;
    p.a0 = c.a0; p.a1 = c.a1; if (!c.a1.ptr) { p.a1 = c.a0; p.a1.width = 0; }
    p.B = c.B; p.N = L->N; p.K = Kuse; p.OUT = L->OUT; p.BN = t3_pick_bn(Kuse, L->OUT, PL); p.NT = L->OUT / p.BN;
    p.MT = (c.B + T3_BM - 1) / T3_BM; p.KB = t3_kb(Kuse); p.nstage = t3_stages(Kuse, p.BN, PL); p.n_types = L->n_types;
    p.wslots = 2;
    p.k_base = k_base;
    if (pre) p.pre = *pre; else { p.pre.ptr = nullptr; p.pre.sb = p.pre.sn = 0; p.pre.rep = 1; p.pre.width = 0; }
    p.a_stationary = 0;
    {
        static int as_env = -1;              // SKELDIFF_TC3_AS=0 disables the activation-stationary schedule (A/B timing)
        if (as_env < 0) { const char* e = getenv("SKELDIFF_TC3_AS"); as_env = (e && e[0] == '0') ? 0 : 1; }
        const T3AsCfg as = as_env ? t3_as_cfg(Kuse, L->OUT, PL) : T3AsCfg{0, 0, 0};
        if (as.bn) { p.a_stationary = 1; p.BN = as.bn; p.NT = L->OUT / as.bn; p.nstage = as.nstage; p.wslots = as.wslots; }
        if (gru) {      // one n-tile = the three gates of 32 hidden units: 96 columns, activation-stationary, whole K in the stage ring
            if (PL != 2 || K1 != 0 || L->OUT != 3 * Kuse || Kuse % 32 || t3_kb(Kuse) > T3_MAX_STAGES || out.width != Kuse) {
                set_error("fused GRU step: needs the two-plane split and a hidden size that is a multiple of 32 (H=%d, planes=%d)", Kuse, PL);
                return SD_ERR_UNSUPPORTED;
            }
            static const int gx_stages = t3_env("SKELDIFF_T3_GRU_XSTAGES", 0), g_wslots = t3_env("SKELDIFF_T3_GRU_WSLOTS", 2);
            p.a_stationary = 1; p.BN = 96; p.NT = L->OUT / 96; p.nstage = t3_kb(Kuse) + gx_stages; p.wslots = g_wslots < 2 ? 2 : (g_wslots > T3_MAX_WSLOTS ? T3_MAX_WSLOTS : g_wslots);
            if (p.nstage > T3_MAX_STAGES) p.nstage = T3_MAX_STAGES;
        }
    }
    { static const int m = t3_env("SKELDIFF_T3_MERGE_LD", 1); p.merge_ld = m; }
    p.norm_out = nullptr;
    p.gru_bias_x = gru ? gru->bias_x : nullptr; p.gru_bias_h = gru ? gru->bias_h : nullptr;
    p.res_slots = 2;
    p.tmem_cols = 4 * p.BN <= 32 ? 32 : (4 * p.BN <= 64 ? 64 : (4 * p.BN <= 128 ? 128 : (4 * p.BN <= 256 ? 256 : 512)));   // (main + corr) x 2 buffers
    p.types = L->types;
    p.out = out;
    int act = SD_ACT_NONE;
    bool has_res = false;
    if (gru) {
        p.row_scale = nullptr; p.bias_node = nullptr; p.ss = nullptr; p.residual = gru->xr; act = T3_ACT_GRU; has_res = true;
        if (gru->xr.rep != 1 || c.a0.rep != 1 || out.rep != 1 || (reinterpret_cast<uintptr_t>(gru->xr.ptr) & 15u) || gru->xr.sb % 4 || gru->xr.sn % 4 ||
            !gru->bias_x || !gru->bias_h || p.pre.ptr) {
            set_error("fused GRU step: operands must be plain 16-byte aligned [B, N, *] tensors");
            return SD_ERR_UNSUPPORTED;
        }
    } else if (apply_epilogue) {
        p.row_scale = c.row_scale; p.bias_node = c.epi.bias_node;
        p.ss = c.epi.ss ? c.epi.ss + (long long)c.epi.ss_row * c.epi.ss_stride : nullptr;
        p.residual = c.epi.residual; act = c.epi.act; has_res = c.epi.residual.ptr != nullptr;
    } else {
        p.row_scale = c.row_scale; p.bias_node = nullptr; p.ss = nullptr; p.residual.ptr = nullptr; p.residual.sb = p.residual.sn = 0; p.residual.rep = 1; p.residual.width = 0;
    }
    // weight planes: [3][types][OUT][K] bf16 -> 3-D map (K, OUT, 3*types)
    static EncodeTiledFn3 enc = nullptr;
    if (!enc) {
        void* fp = nullptr; cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
            set_error("cuTensorMapEncodeTiled unavailable"); return SD_ERR_CUDA;
        }
        enc = reinterpret_cast<EncodeTiledFn3>(fp);
    }
    CUtensorMap mw;
    cuuint64_t dims[3] = {(cuuint64_t)L->K, (cuuint64_t)L->OUT, (cuuint64_t)(PL * L->n_types)};
    cuuint64_t strides[2] = {(cuuint64_t)L->K * 2, (cuuint64_t)L->OUT * L->K * 2};
    cuuint32_t box[3] = {(cuuint32_t)T3_BK, (cuuint32_t)p.BN, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&mw, PL == 3 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3,
                     const_cast<uint16_t*>(PL == 3 ? L->W_bf16 : L->W_f16), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (weights x3) failed: %d", (int)r); return SD_ERR_CUDA; }
    // per-node tables (all OUT columns) in the two-plane activation-stationary schedule; the three-plane one has no shared memory to spare
    p.tab_cols = t3_tab_cols(p.a_stationary && PL == 2 ? L->OUT : p.BN);
    size_t smem = (p.a_stationary ? (size_t)p.wslots * PL * p.BN * 128 + t3_misc_smem(p.tab_cols, PL) : t3_fixed_smem(Kuse, p.BN, PL)) + (size_t)p.nstage * t3_stage_bytes(PL);
    // Two staging tiles per epilogue warp (two-plane kernel, output through bulk stores): the store of chunk c overlaps chunk
    // c + 1 instead of being waited for at its start.  16 KB; taken first on wide outputs (many chunks per activation tile).
    // Bare layer on the two-plane kernel: direct SWIZZLE_128B store of the staging tile (see the kernel).  SKELDIFF_T3_DIRECT=0: off.
    static const int direct_env = t3_env("SKELDIFF_T3_DIRECT", 1);
    const bool want_direct = direct_env && PL == 2 && !gru && act == SD_ACT_NONE && !has_res && !p.ss && !p.bias_node && !p.pre.ptr &&
                             out.rep == 1 && p.BN % 32 == 0;
    p.direct = 0;
    p.stg2 = 0;
    {
        static int s2_env = -1;              // SKELDIFF_T3_STG2=0/1/2: never / wide outputs only (default) / whenever it fits
        if (s2_env < 0) { const char* e = getenv("SKELDIFF_T3_STG2"); s2_env = (e && e[0]) ? atoi(e) : 1; }
        const size_t extra = (size_t)4 * 32 * 32 * sizeof(float);
        // (a direct store overlaps the next chunk's TMEM read only with a second tile; not at the price of the activation ring's two boxes)
        const bool direct_room = want_direct && smem + extra + (K1 == 0 && c.a0.rep == 1 ? 2 * (size_t)T3_BM * T3_BK * sizeof(float) : 0) <= (size_t)227 * 1024;
        if (PL == 2 && !gru && s2_env && (s2_env == 2 || p.NT >= 4 || direct_room) && smem + extra <= (size_t)227 * 1024) { p.stg2 = 1; smem += extra; }
    }
    // Residual through a TMA ring (two-plane kernel): 32 KB of boxes [128 samples][32 columns] of the 3-D tensor (columns, node,
    // sample), loaded by the otherwise idle warp 14 two chunks ahead of the epilogue.  192 -> 192 + tanh + residual: 435 -> 312 us
    // (3.97 TB/s = 61 % of the HBM copy peak).  On the K = 256 layer (to_out) the ring lifts shared memory from 195 to 227 KB, past
    // the 196 KB carve-out step that costs L1 (see t3_as_cfg), and still wins: 349 -> 316 us (SKELDIFF_T3_RES_TMA_MAX_KB=195 restores
    // the old limit, SKELDIFF_T3_RES_TMA=0 the epilogue's own loads).
    CUtensorMap mr = mw;
    p.res_tma = 0;
    {
        static int res_env = -1;             // SKELDIFF_T3_RES_TMA=0: residual by the epilogue's own loads (A/B timing)
        if (res_env < 0) { const char* e = getenv("SKELDIFF_T3_RES_TMA"); res_env = (e && e[0] == '0') ? 0 : 1; }
        const size_t ring = (size_t)2 * T3_BM * 32 * sizeof(float);
        static const int max_kb = t3_env("SKELDIFF_T3_RES_TMA_MAX_KB", 227);
        if ((res_env || gru) && has_res && PL == 2 && p.residual.rep == 1 && p.BN % 32 == 0 && smem + ring <= (size_t)max_kb * 1024) {
            cuuint64_t rdims[3] = {(cuuint64_t)L->OUT, (cuuint64_t)L->N, (cuuint64_t)c.B};
            cuuint64_t rstrides[2] = {(cuuint64_t)p.residual.sn * 4, (cuuint64_t)p.residual.sb * 4};
            cuuint32_t rbox[3] = {32, 1, (cuuint32_t)T3_BM};
            cuuint32_t restr[3] = {1, 1, 1};
            CUresult rr = enc(&mr, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(p.residual.ptr), rdims, rstrides, rbox, restr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (rr == CUDA_SUCCESS) { p.res_tma = 1; smem += ring; }
            if (rr == CUDA_SUCCESS && gru) {     // four boxes per n-tile (three gates + the previous state): a deeper ring while it fits
                static const int want = t3_env("SKELDIFF_T3_GRU_SLOTS", 3);
                while (p.res_slots < want && p.res_slots < T3_MAX_RES_SLOTS && smem + ring / 2 <= (size_t)227 * 1024) { ++p.res_slots; smem += ring / 2; }
            }
        }
    }
    if (gru && !p.res_tma) { set_error("fused GRU step: the operand ring does not fit / could not be mapped"); return SD_ERR_UNSUPPORTED; }
    // Row norms of the output from the epilogue (GlinCall::norm_out): only where the residual-ring variant runs and the CTA covers
    // whole rows (activation-stationary schedule, or a single n-tile); tl_norm_written tells the caller whether it still has to run
    // the separate pass.  SKELDIFF_T3_NORM_FUSED=0: never.
    {
        static const int nf_env = t3_env("SKELDIFF_T3_NORM_FUSED", 1);
        if (nf_env && c.norm_out && apply_epilogue && !gru && !pre && PL == 2 && act == SD_ACT_TANH && has_res && p.res_tma &&
            (p.a_stationary || p.NT == 1) && out.rep == 1 &&
            out.sb == (long long)L->N * L->OUT && out.sn == L->OUT) {
            p.norm_out = c.norm_out;
            tl_norm_written = true;
        }
    }
    // Output through TMA stores (two-plane kernel): the output is a 3-D tensor (columns, node, sample) like the residual.
    CUtensorMap mo = mw;
    p.out_tma = 0;
    {
        static int out_env = -1;             // SKELDIFF_T3_OUT_TMA=0: output by the epilogue's STG.128 (A/B timing)
        if (out_env < 0) { const char* e = getenv("SKELDIFF_T3_OUT_TMA"); out_env = (e && e[0] == '0') ? 0 : 1; }
        if ((out_env || gru || want_direct) && PL == 2 && out.rep == 1 && p.BN % 32 == 0) {
            cuuint64_t odims[3] = {(cuuint64_t)(gru ? Kuse : L->OUT), (cuuint64_t)L->N, (cuuint64_t)c.B};
            cuuint64_t ostrides[2] = {(cuuint64_t)out.sn * 4, (cuuint64_t)out.sb * 4};
            cuuint32_t obox[3] = {32, 1, 32};
            cuuint32_t oestr[3] = {1, 1, 1};
            CUresult ro = enc(&mo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, out.ptr, odims, ostrides, obox, oestr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, want_direct ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                              CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (ro == CUDA_SUCCESS) { p.out_tma = 1; p.direct = want_direct ? 1 : 0; }
        }
    }
    if (gru && !p.out_tma) { set_error("fused GRU step: the output tensor could not be mapped"); return SD_ERR_UNSUPPORTED; }
    // fp32 activations through a TMA ring (two-plane kernel without a residual, one K segment read in place): as many 32 KB
    // boxes as fit (2 .. 4); the producers then issue no global loads at all.
    CUtensorMap ma = mw;
    p.raw_slots = 0;
    {
        static int a_env = -1;               // SKELDIFF_T3_A_TMA=0: activations by the producers' own loads (A/B timing)
        if (a_env < 0) { const char* e = getenv("SKELDIFF_T3_A_TMA"); a_env = (e && e[0] == '0') ? 0 : 1; }
        const size_t box = (size_t)T3_BM * T3_BK * sizeof(float);
        int slots = (int)(((size_t)227 * 1024 - smem) / box);
        if (slots > 4) slots = 4;
        if (gru) {          // the previous hidden state as [128 samples][32 units] boxes of the residual ring (not the k-block ring)
            cuuint64_t adims[3] = {(cuuint64_t)K0, (cuuint64_t)L->N, (cuuint64_t)c.B};
            cuuint64_t astrides[2] = {(cuuint64_t)c.a0.sn * 4, (cuuint64_t)c.a0.sb * 4};
            cuuint32_t abox[3] = {32, 1, (cuuint32_t)T3_BM};
            cuuint32_t aestr[3] = {1, 1, 1};
            CUresult ra = enc(&ma, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(c.a0.ptr), adims, astrides, abox, aestr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (ra != CUDA_SUCCESS) { set_error("fused GRU step: cuTensorMapEncodeTiled (hidden state) failed: %d", (int)ra); return SD_ERR_CUDA; }
        } else if (a_env && PL == 2 && !has_res && K1 == 0 && c.a0.rep == 1 && slots >= 2) {
            cuuint64_t adims[3] = {(cuuint64_t)K0, (cuuint64_t)L->N, (cuuint64_t)c.B};
            cuuint64_t astrides[2] = {(cuuint64_t)c.a0.sn * 4, (cuuint64_t)c.a0.sb * 4};
            cuuint32_t abox[3] = {(cuuint32_t)T3_BK, 1, (cuuint32_t)T3_BM};
            cuuint32_t aestr[3] = {1, 1, 1};
            CUresult ra = enc(&ma, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(c.a0.ptr), adims, astrides, abox, aestr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (ra == CUDA_SUCCESS) { p.raw_slots = slots; smem += (size_t)slots * box; }
        }
    }
    const int sms = sm_count();
    long long gangs = p.a_stationary ? sms : sms / p.NT;
    if (gangs < 1) gangs = 1;
    if (gangs > (long long)p.N * p.MT) gangs = (long long)p.N * p.MT;
    const int grid = (int)(p.a_stationary ? gangs : gangs * p.NT);
    if (gru) return t3_launch_t<T3_ACT_GRU, true, 2, true, true, false>(mw, mr, mo, ma, p, grid, smem, st);
#define T3_DISPATCH(PLN) \
    if (act == SD_ACT_NONE) return has_res ? t3_launch_t<SD_ACT_NONE, true, PLN>(mw, mr, mo, ma, p, grid, smem, st) : t3_launch_t<SD_ACT_NONE, false, PLN>(mw, mr, mo, ma, p, grid, smem, st); \
    if (act == SD_ACT_TANH) return has_res ? t3_launch_t<SD_ACT_TANH, true, PLN>(mw, mr, mo, ma, p, grid, smem, st) : t3_launch_t<SD_ACT_TANH, false, PLN>(mw, mr, mo, ma, p, grid, smem, st); \
    if (act == SD_ACT_TANH_TANH) return has_res ? t3_launch_t<SD_ACT_TANH_TANH, true, PLN>(mw, mr, mo, ma, p, grid, smem, st) : t3_launch_t<SD_ACT_TANH_TANH, false, PLN>(mw, mr, mo, ma, p, grid, smem, st);
    if (PL == 2) { T3_DISPATCH(2) } else { T3_DISPATCH(3) }
#undef T3_DISPATCH
    set_error("bf16x3 path: unknown activation %d", act);
    return SD_ERR_INVALID;
}

// out = epilogue(A @ W^T) with fp32 views; the caller guarantees G == identity for this call when apply_epilogue is set.
// A two-segment layer whose full K cannot run activation-stationary (K = 384: 288 KB of plane stages) but whose segments can
// is split along K: launch 1 writes the raw product of segment 0 to the caller's scratch, launch 2 adds it in front of its
// epilogue.  Two stationary launches (0.28 + 0.33 ms) replace one weight-resident launch that re-transformed A six times (1.55 ms).
int glin_tc3_launch(const sd_glin* L, const GlinCall& c, const ViewW& out, bool apply_epilogue, cudaStream_t st) {
    const int K0 = c.a0.width, K1 = c.a1.ptr ? c.a1.width : 0;
    if (K0 + K1 != L->K) { set_error("bf16x3 path: operand widths %d+%d do not match the layer's K=%d", K0, K1, L->K); return SD_ERR_INVALID; }
    static int split_env = -1;               // SKELDIFF_TC3_KSPLIT=0 disables the K-split (A/B timing)
    if (split_env < 0) { const char* e = getenv("SKELDIFF_TC3_KSPLIT"); split_env = (e && e[0] == '0') ? 0 : 1; }
    const int PL = (tl_split_planes == 2 && L->W_f16) ? 2 : 3;
    auto stationary_ok = [&](int K) { return t3_as_cfg(K, L->OUT, PL).bn != 0; };
    const bool split = split_env && K1 > 0 && apply_epilogue && c.scratch && c.scratch != out.ptr && !c.row_scale && !c.epi.residual.ptr &&
                       !stationary_ok(L->K) && stationary_ok(K0) && stationary_ok(K1);
    if (!split) return t3_launch_one(L, c, out, apply_epilogue, 0, nullptr, st);
    GlinCall c1 = c; c1.a1.ptr = nullptr; c1.a1.width = 0;
    int rc = t3_launch_one(L, c1, contiguous_view_w(c.scratch, L->N, L->OUT), false, 0, nullptr, st);
    if (rc) return rc;
    GlinCall c2 = c; c2.a0 = c.a1; c2.a1.ptr = nullptr; c2.a1.width = 0;
    const View pre = contiguous_view(c.scratch, L->N, L->OUT);
    return t3_launch_one(L, c2, out, true, K0, &pre, st);
}

// One decoder / encoder GRU step with identity graph influence on the two-plane tensor-core path: h_out = GRU(xr, h_in) with
// W_hh in the gate-interleaved order (L->W_f16: [2][types][3H][H] planes of the permuted rows), xr = x-side products in the same
// column order, bias_x / bias_h: [N][3H] permuted.  h_out must not alias h_in.
int glin_tc3_gru_step(const sd_glin* L, const View& h_in, const View& xr, const float* bias_x, const float* bias_h, const ViewW& h_out,
                      int B, cudaStream_t st) {
    if (h_in.ptr == h_out.ptr) { set_error("fused GRU step: the new hidden state must not alias the previous one"); return SD_ERR_INVALID; }
    GlinCall c;
    c.a0 = h_in; c.a1.ptr = nullptr; c.a1.sb = c.a1.sn = 0; c.a1.rep = 1; c.a1.width = 0;
    c.row_scale = nullptr; c.out = h_out; c.scratch = nullptr; c.B = B;
    c.epi.bias_node = nullptr; c.epi.ss = nullptr; c.epi.ss_row_idx = nullptr; c.epi.ss_row = 0; c.epi.ss_stride = 0; c.epi.act = SD_ACT_NONE;
    c.epi.residual.ptr = nullptr; c.epi.residual.sb = c.epi.residual.sn = 0; c.epi.residual.rep = 1; c.epi.residual.width = 0; c.epi.OUT = L->OUT;
    const T3Gru g{xr, bias_x, bias_h};
    return t3_launch_one(L, c, h_out, false, 0, nullptr, st, &g);
}

}  // namespace sd
