// StaticGraphLinear on the 5th-generation tensor cores (sm_100a): tcgen05.mma with TMEM accumulators,
// operands staged by TMA into 128B-swizzled shared memory, fused epilogue out of TMEM.
//
// Persistent, warp-specialised kernel (one CTA per SM):
//   warp 0    TMA producer: the node type's weight tile W[type][n0:n0+BN, :] is loaded once per
//             (node, n-tile) group and stays resident; activation tiles A[b0:b0+128, node, k0:k0+64]
//             stream through a ring of NSTAGE 16 KB stages (3-D tensor map over [B, N, K], so the
//             sample-major global layout is kept and per-node row gathering is done by TMA).
//   warp 1    MMA issuer: one elected lane issues tcgen05.mma.kind::f16 (M=128, N=BN, K=16) into one
//             of two TMEM accumulator stages, tcgen05.commit releases smem stages / publishes the tile.
//   warp 2    TMEM allocator (512 columns).
//   warps 4-11 epilogue (two per TMEM lane quarter, alternating 32-column chunks): tcgen05.ld, row scale (RMSNorm
//             fold), each chunk transposed through a swizzled fp32 staging tile so that four lanes cover one row;
//             bias, time scale/shift, tanh, residual (coalesced, next tile's rows prefetched into L2), bf16/fp32 store.
// The layer's arithmetic intensity (K = 192: 96 FLOP/B in bf16) is below the B200 ridge
// (~250 FLOP/B at the measured peaks), so the kernel is HBM-bound by design: its job is to stream activations once.
// Measured: 0.163 ms for 192 -> 192 + tanh + residual at B = 25 600 = 58 % of the measured HBM peak.
//
// Reference semantics: GraphLinear.forward, src/core/network/layers/graph_structural.py:30-43.
#include "sd_internal.h"
#include "sd_tc.cuh"
#include <cuda.h>

namespace sd {

using namespace sd::tc;

constexpr int TC_BM = 128;        // samples per tile (UMMA M)
constexpr int TC_BK = 64;         // bf16 elements per K block = 128 bytes = one swizzle row
constexpr int TC_MAX_STAGES = 8;   // A-tile ring depth is chosen at launch from the smem left after the weight tile
constexpr int TC_EPI_WARPS = 8;
constexpr int TC_THREADS = 128 + TC_EPI_WARPS * 32;

// epilogue activation, resolved at compile time so that the epilogue loop body stays a few KB (a first
// version with runtime branches and inlined tanhf was 37 KB of SASS per chunk and stalled on instruction fetch)
constexpr int TCA_NONE = 0, TCA_TANH_FAST = 1, TCA_TANH = 2, TCA_TANH_TANH = 3;
constexpr int TC_TMEM_COLS = 512;

struct TcParams {
    int B, N, K, OUT, BN, NT, MT, KB, KB0;   // KB0 = K blocks served by segment 0
    NodeTypes types;
    const float* row_scale;     // [B*N] or null
    const float* bias_node;     // [N][OUT] or null
    const float* ss;            // scale at [o], shift at [OUT+o] (row already resolved) or null
    int act;
    const __nv_bfloat16* res; long long res_sb, res_sn;
    void* out; int out_fp32; long long out_sb, out_sn;
    int accurate_tanh;
    int nstage;
};

struct __align__(8) TcBarriers {
    uint64_t full[TC_MAX_STAGES], empty[TC_MAX_STAGES];
    uint64_t w_full, w_empty;
    uint64_t acc_full[2], acc_empty[2];
    uint32_t tmem_base;
    uint32_t pad;
};

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&t);
}

template <int ACT> __device__ __forceinline__ float epi_act(float x) {
    if (ACT == TCA_TANH_FAST) return tanh_fast(x);
    if (ACT == TCA_TANH) return tanh_acc(x);
    if (ACT == TCA_TANH_TANH) return tanh_acc(tanh_acc(x));
    return x;
}

template <int ACT, bool HAS_RES, bool OUT_FP32>
__global__ void __launch_bounds__(TC_THREADS, 1)
glin_tc_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
               const __grid_constant__ CUtensorMap map_w, const TcParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // layout: [W: KB blocks of BN x 128 B][A: NSTAGE x 16 KB][epilogue tables 2 x BN floats][epilogue staging 8 x 4 KB][barriers]
    // 1024-byte alignment (SWIZZLE_128B atoms) by pointer arithmetic on the __shared__ array: rounding the address up through
    // uintptr_t made the compiler lose the address space and emit generic LD/ST for every shared-memory access of the kernel.
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t w_block_bytes = (uint32_t)p.BN * 128u;
    uint8_t* w_smem = smem;
    uint8_t* a_smem = w_smem + (size_t)p.KB * w_block_bytes;
    float* epi_mul = reinterpret_cast<float*>(a_smem + (size_t)p.nstage * TC_BM * 128);   // 16-byte aligned
    float* epi_add = epi_mul + p.BN;
    float* epi_stage = epi_add + p.BN;                             // TC_EPI_WARPS x [32 rows][32 floats], swizzled
    TcBarriers* bars = reinterpret_cast<TcBarriers*>(epi_stage + TC_EPI_WARPS * 32 * 32);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a0);
        tma_prefetch_desc(&map_a1);
        tma_prefetch_desc(&map_w);
        for (int s = 0; s < p.nstage; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
        mbar_init(&bars->w_full, 1);
        mbar_init(&bars->w_empty, 1);
        for (int s = 0; s < 2; ++s) { mbar_init(&bars->acc_full[s], 1); mbar_init(&bars->acc_empty[s], TC_EPI_WARPS * 32); }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(&bars->tmem_base, TC_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    // contiguous share of the (group, m-tile) work list; group = (node, n-tile)
    const long long total = (long long)p.N * p.NT * p.MT;
    const long long item_lo = total * blockIdx.x / gridDim.x;
    const long long item_hi = total * (blockIdx.x + 1) / gridDim.x;

    if (warp == 0) {
        // ===================================================================== TMA producer
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0, w_phase = 0;
            long long cur_g = -1;
            for (long long it = item_lo; it < item_hi; ++it) {
                const long long g = it / p.MT;
                const int mt = (int)(it % p.MT);
                const int node = (int)(g / p.NT), nt = (int)(g % p.NT);
                if (g != cur_g) {
                    if (cur_g >= 0) { mbar_wait(&bars->w_empty, w_phase); w_phase ^= 1; }   // MMAs of the old group are done
                    mbar_arrive_expect_tx(&bars->w_full, (uint32_t)p.KB * w_block_bytes);
                    for (int kb = 0; kb < p.KB; ++kb)
                        tma_load_3d(w_smem + (size_t)kb * w_block_bytes, &map_w, &bars->w_full, kb * TC_BK, nt * p.BN, p.types.t[node]);
                    cur_g = g;
                }
                for (int kb = 0; kb < p.KB; ++kb) {
                    mbar_wait(&bars->empty[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&bars->full[stage], TC_BM * 128);
                    if (kb < p.KB0) tma_load_3d(a_smem + (size_t)stage * TC_BM * 128, &map_a0, &bars->full[stage], kb * TC_BK, node, mt * TC_BM);
                    else tma_load_3d(a_smem + (size_t)stage * TC_BM * 128, &map_a1, &bars->full[stage], (kb - p.KB0) * TC_BK, node, mt * TC_BM);
                    if (++stage == p.nstage) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================================================================== MMA issuer
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16(TC_BM, (uint32_t)p.BN);
            int stage = 0; uint32_t phase = 0, w_phase = 0, acc = 0, acc_phase = 0;
            long long cur_g = -1;
            for (long long it = item_lo; it < item_hi; ++it) {
                const long long g = it / p.MT;
                if (g != cur_g) { mbar_wait(&bars->w_full, w_phase); w_phase ^= 1; cur_g = g; }
                mbar_wait(&bars->acc_empty[acc], acc_phase ^ 1);           // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * (uint32_t)p.BN;
                for (int kb = 0; kb < p.KB; ++kb) {
                    mbar_wait(&bars->full[stage], phase);
                    tc_fence_after();
                    const uint64_t adesc = umma_desc_sw128(smem_u32(a_smem + (size_t)stage * TC_BM * 128));
                    const uint64_t bdesc = umma_desc_sw128(smem_u32(w_smem + (size_t)kb * w_block_bytes));
#pragma unroll
                    for (int k = 0; k < TC_BK / 16; ++k)      // +32 bytes (>>4 = 2) per K=16 step inside the swizzle row
                        umma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
                    umma_commit(&bars->empty[stage]);           // smem stage reusable once these MMAs retire
                    if (++stage == p.nstage) { stage = 0; phase ^= 1; }
                }
                umma_commit(&bars->acc_full[acc]);              // accumulator complete -> epilogue
                const bool last_of_group = (it + 1 == item_hi) || ((it + 1) / p.MT != g);
                if (last_of_group) umma_commit(&bars->w_empty);  // weight tile may be overwritten
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ===================================================================== epilogue (TC_EPI_WARPS x 32 threads)
        // Warp w may only touch TMEM lanes 32*(w%4)..+31; the two warps that share a lane quarter split the
        // 32-column chunks between them (eg = 0/1), which doubles the loads in flight per accumulator row.
        const int quarter = warp & 3;
        const int eg = (warp - 4) >> 2;
        const int et = threadIdx.x - 128;
        constexpr int EPI_THREADS = TC_EPI_WARPS * 32;
        float* stg = epi_stage + (warp - 4) * (32 * 32);
        const int tr = lane >> 2, tc8 = lane & 3;
        uint32_t acc = 0, acc_phase = 0;
        long long cur_g = -1;
        for (long long it = item_lo; it < item_hi; ++it) {
            const long long g = it / p.MT;
            const int mt = (int)(it % p.MT);
            const int node = (int)(g / p.NT), nt = (int)(g % p.NT);
            const int o0 = nt * p.BN;
            if (g != cur_g) {
                asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");   // everyone finished reading the old tables
                for (int c = et; c < p.BN; c += EPI_THREADS) {
                    const int o = o0 + c;
                    const float mul = p.ss ? (__ldg(p.ss + o) + 1.0f) : 1.0f;
                    const float bias = p.bias_node ? __ldg(p.bias_node + (long long)node * p.OUT + o) : 0.0f;
                    epi_mul[c] = mul;
                    epi_add[c] = fmaf(bias, mul, p.ss ? __ldg(p.ss + p.OUT + o) : 0.0f);
                }
                asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
                cur_g = g;
            }
            // TMEM hands each lane one ROW of the tile; stored that way every LDG/STG.128 of a warp touches 32 different rows
            // (32 LSU wavefronts per instruction).  Each 32-column chunk is transposed through a swizzled 4 KB fp32 staging tile:
            // afterwards four lanes cover 8 columns each of one row (64 contiguous bytes of bf16) and an instruction touches 8 rows.
            const int b_own = mt * TC_BM + quarter * 32 + lane;
            const float rs = (b_own < p.B && p.row_scale) ? __ldg(p.row_scale + (long long)b_own * p.N + node) : 1.0f;
            const int bT0 = mt * TC_BM + quarter * 32 + tr;                      // transposed mapping: rows bT0 + 8 j, columns 8 tc8 .. +7
            const long long res_base = (long long)node * p.res_sn + o0 + 8 * tc8;
            const long long out_base = (long long)node * p.out_sn + o0 + 8 * tc8;
            uint4 rr[4];
            if (HAS_RES && eg * 32 < p.BN) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (bT0 + 8 * j < p.B) rr[j] = __ldg(reinterpret_cast<const uint4*>(p.res + (long long)(bT0 + 8 * j) * p.res_sb + res_base + eg * 32));
            }
            if (HAS_RES && it + 1 < item_hi) {
                // one residual chunk per warp in flight does not cover DRAM latency (ncu: 38 % of the samples of the first capture
                // were epilogue warps waiting for it): the NEXT tile's residual rows (half a row per thread) are pulled into L2 now
                const long long it2 = it + 1, g2 = it2 / p.MT;
                const int b2 = (int)(it2 % p.MT) * TC_BM + quarter * 32 + lane, node2 = (int)(g2 / p.NT), nt2 = (int)(g2 % p.NT);
                if (b2 < p.B)
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::
                                 "l"(p.res + (long long)b2 * p.res_sb + (long long)node2 * p.res_sn + nt2 * p.BN + eg * (p.BN >> 1)), "r"((uint32_t)p.BN) : "memory");
            }
            mbar_wait(&bars->acc_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * (uint32_t)p.BN;
            for (int c0 = eg * 32; c0 < p.BN; c0 += 64) {
                uint32_t v[32];
                tmem_ld_32x32(t_row + (uint32_t)c0, v);
                tmem_ld_wait();
                __syncwarp();                                   // the previous chunk has been read out of the staging tile
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    *reinterpret_cast<float4*>(stg + lane * 32 + 4 * (q ^ (lane & 7))) =
                        make_float4(__uint_as_float(v[4 * q]) * rs, __uint_as_float(v[4 * q + 1]) * rs, __uint_as_float(v[4 * q + 2]) * rs, __uint_as_float(v[4 * q + 3]) * rs);
                __syncwarp();
                const float4 m0 = *reinterpret_cast<const float4*>(epi_mul + c0 + 8 * tc8), m1 = *reinterpret_cast<const float4*>(epi_mul + c0 + 8 * tc8 + 4);
                const float4 a0 = *reinterpret_cast<const float4*>(epi_add + c0 + 8 * tc8), a1 = *reinterpret_cast<const float4*>(epi_add + c0 + 8 * tc8 + 4);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int row = tr + 8 * j;
                    const float4 x0 = *reinterpret_cast<const float4*>(stg + row * 32 + 4 * ((2 * tc8) ^ (tr & 7)));
                    const float4 x1 = *reinterpret_cast<const float4*>(stg + row * 32 + 4 * ((2 * tc8 + 1) ^ (tr & 7)));
                    float f[8];
                    f[0] = epi_act<ACT>(fmaf(x0.x, m0.x, a0.x)); f[1] = epi_act<ACT>(fmaf(x0.y, m0.y, a0.y));
                    f[2] = epi_act<ACT>(fmaf(x0.z, m0.z, a0.z)); f[3] = epi_act<ACT>(fmaf(x0.w, m0.w, a0.w));
                    f[4] = epi_act<ACT>(fmaf(x1.x, m1.x, a1.x)); f[5] = epi_act<ACT>(fmaf(x1.y, m1.y, a1.y));
                    f[6] = epi_act<ACT>(fmaf(x1.z, m1.z, a1.z)); f[7] = epi_act<ACT>(fmaf(x1.w, m1.w, a1.w));
                    if (HAS_RES) {
                        const uint32_t w4[4] = {rr[j].x, rr[j].y, rr[j].z, rr[j].w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            f[2 * e] += __uint_as_float(w4[e] << 16);
                            f[2 * e + 1] += __uint_as_float(w4[e] & 0xFFFF0000u);
                        }
                    }
                    const int bj = bT0 + 8 * j;
                    if (bj < p.B) {
                        if (OUT_FP32) {
                            float4* op = reinterpret_cast<float4*>(static_cast<float*>(p.out) + (long long)bj * p.out_sb + out_base + c0);
                            op[0] = make_float4(f[0], f[1], f[2], f[3]);
                            op[1] = make_float4(f[4], f[5], f[6], f[7]);
                        } else {
                            *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + (long long)bj * p.out_sb + out_base + c0) =
                                make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
                        }
                    }
                }
                if (HAS_RES && c0 + 64 < p.BN) {                // next chunk's residual goes in flight behind this chunk's stores
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (bT0 + 8 * j < p.B) rr[j] = __ldg(reinterpret_cast<const uint4*>(p.res + (long long)(bT0 + 8 * j) * p.res_sb + res_base + c0 + 64));
                }
            }
            tc_fence_before();
            mbar_arrive(&bars->acc_empty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, TC_TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 3-D bf16 tensor map: dims (inner, mid, outer), strides in elements for mid/outer, box (64, box_mid, box_outer)
static int make_map(CUtensorMap* m, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1, uint64_t s2,
                    uint32_t box1, uint32_t box2) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled unavailable"); return SD_ERR_CUDA; }
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {s1 * 2, s2 * 2};
    cuuint32_t box[3] = {(cuuint32_t)TC_BK, box1, box2};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d): dims %llu,%llu,%llu strides %llu,%llu", (int)r,
                                       (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2,
                                       (unsigned long long)s1, (unsigned long long)s2); return SD_ERR_CUDA; }
    return SD_OK;
}

static int pick_bn(int OUT) {
    if (OUT <= 256) return OUT;
    for (int bn = 256; bn >= 16; bn -= 16) if (OUT % bn == 0) return bn;
    return 0;
}

static size_t tc_fixed_smem(int K, int bn) {   // everything except the A ring
    return (size_t)(K / TC_BK) * bn * 128 + 2 * (size_t)bn * 4 + (size_t)TC_EPI_WARPS * 32 * 32 * 4 + sizeof(TcBarriers) + 1024;
}
static int tc_stages(int K, int bn) {
    const size_t budget = 227 * 1024;
    const size_t fixed = tc_fixed_smem(K, bn);
    if (fixed + 2 * TC_BM * 128 > budget) return 0;
    size_t n = (budget - fixed) / (TC_BM * 128);
    return (int)(n > TC_MAX_STAGES ? TC_MAX_STAGES : n);
}

bool glin_tc_supported(int K0, int K1, int OUT) {
    const int K = K0 + K1;
    if (K0 % TC_BK || K1 % TC_BK || K <= 0) return false;
    if (OUT % 32) return false;               // epilogue reads 32-column chunks
    const int bn = pick_bn(OUT);
    if (bn == 0 || bn % 32) return false;
    return tc_stages(K, bn) >= 2;
}

template <int ACT, bool HAS_RES, bool OUT_FP32>
static int tc_launch_t(const CUtensorMap& ma0, const CUtensorMap& ma1, const CUtensorMap& mw, const TcParams& p, int grid, size_t smem, cudaStream_t st) {
    auto kern = glin_tc_kernel<ACT, HAS_RES, OUT_FP32>;
    static unsigned long long configured = 0;      // bit d: attribute set on device d (it is per device)
    if (int rc_attr = opt_in_smem(kern, (size_t)((227 * 1024)), configured)) return rc_attr;
    kern<<<grid, TC_THREADS, smem, st>>>(ma0, ma1, mw, p);
    SD_LAUNCH_OK("glin_tc_kernel");
    return SD_OK;
}

template <int ACT>
static int tc_launch_act(const CUtensorMap& ma0, const CUtensorMap& ma1, const CUtensorMap& mw, const TcParams& p, int grid, size_t smem, cudaStream_t st) {
    if (p.res) return p.out_fp32 ? tc_launch_t<ACT, true, true>(ma0, ma1, mw, p, grid, smem, st) : tc_launch_t<ACT, true, false>(ma0, ma1, mw, p, grid, smem, st);
    return p.out_fp32 ? tc_launch_t<ACT, false, true>(ma0, ma1, mw, p, grid, smem, st) : tc_launch_t<ACT, false, false>(ma0, ma1, mw, p, grid, smem, st);
}

int glin_tc_launch(const sd_glin* L, const TcCall& c, cudaStream_t st) {
    if (!L->W_bf16) { set_error("tcgen05 path: bf16 weights not set on this layer (sd_glin_set_bf16)"); return SD_ERR_INVALID; }
    const int K0 = c.a0.width, K1 = c.a1.ptr ? c.a1.width : 0;
    if (K0 + K1 != L->K || !glin_tc_supported(K0, K1, L->OUT)) {
        set_error("tcgen05 path: unsupported shape K=%d+%d OUT=%d", K0, K1, L->OUT);
        return SD_ERR_UNSUPPORTED;
    }
    if (c.B <= 0) return SD_OK;
    TcParams p;
    p.B = c.B; p.N = L->N; p.K = L->K; p.OUT = L->OUT; p.BN = pick_bn(L->OUT); p.NT = L->OUT / p.BN;
    p.MT = (c.B + TC_BM - 1) / TC_BM; p.KB = L->K / TC_BK; p.KB0 = K0 / TC_BK;
    p.types = L->types;
    p.row_scale = c.row_scale; p.bias_node = c.bias_node; p.ss = c.ss; p.act = c.act;
    p.res = c.res; p.res_sb = c.res_sb; p.res_sn = c.res_sn;
    p.out = c.out; p.out_fp32 = c.out_fp32; p.out_sb = c.out_sb; p.out_sn = c.out_sn;
    p.accurate_tanh = c.accurate_tanh;
    p.nstage = tc_stages(L->K, p.BN);
    CUtensorMap ma0, ma1, mw;
    int rc = make_map(&ma0, c.a0.ptr, (uint64_t)K0, (uint64_t)L->N, (uint64_t)c.B, (uint64_t)c.a0.sn, (uint64_t)c.a0.sb, 1, TC_BM);
    if (rc) return rc;
    if (K1) rc = make_map(&ma1, c.a1.ptr, (uint64_t)K1, (uint64_t)L->N, (uint64_t)c.B, (uint64_t)c.a1.sn, (uint64_t)c.a1.sb, 1, TC_BM);
    else ma1 = ma0;
    if (rc) return rc;
    rc = make_map(&mw, L->W_bf16, (uint64_t)L->K, (uint64_t)L->OUT, (uint64_t)L->n_types, (uint64_t)L->K, (uint64_t)L->OUT * L->K, (uint32_t)p.BN, 1);
    if (rc) return rc;
    const size_t smem = tc_fixed_smem(L->K, p.BN) + (size_t)p.nstage * TC_BM * 128;
    const int sms = sm_count();
    const long long total = (long long)p.N * p.NT * p.MT;
    const int grid = (int)(total < sms ? total : sms);
    switch (c.act) {
        case SD_ACT_NONE: return tc_launch_act<TCA_NONE>(ma0, ma1, mw, p, grid, smem, st);
        case SD_ACT_TANH: return c.accurate_tanh ? tc_launch_act<TCA_TANH>(ma0, ma1, mw, p, grid, smem, st)
                                                 : tc_launch_act<TCA_TANH_FAST>(ma0, ma1, mw, p, grid, smem, st);
        case SD_ACT_TANH_TANH: return tc_launch_act<TCA_TANH_TANH>(ma0, ma1, mw, p, grid, smem, st);
        default: set_error("tcgen05 path: unknown activation %d", c.act); return SD_ERR_INVALID;
    }
}

}  // namespace sd
