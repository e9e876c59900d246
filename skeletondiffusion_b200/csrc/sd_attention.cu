// Node attention: softmax_j(q_n . k_j * dh^-1/2) v_j over the nodes of one sample, one head.
// Reference: Attention.forward, src/core/network/layers/attention.py:125-135.
//
// Two kernels.  node_attention_bulk_kernel<N> (below, second half of the file) is the shipped fp32 configuration (8 heads x 32
// channels, N = 16 / 17 / 21): persistent CTAs, one cp.async.bulk per sample, Q/K/V read in place.  node_attention_kernel
// handles every other shape and the bf16 I/O of the bf16 mode:
// one warp per (sample, head) task, several tasks per warp (grid-stride), lane = query node.  The head's Q/K/V rows
// ([N, dh] each) are copied into shared memory with coalesced 16-byte loads (rows padded to dh + 4 floats so that a
// lane reading ITS OWN q row is bank-conflict free; K/V rows are read as warp-wide broadcasts).  All products run
// on the packed FFMA2 pipe: q.k as float2 partial sums over channel pairs, p_j * v_j as scalar x float2.
// qkv / out are fp32 (parity path) or bf16 (tensor-core path); the math is fp32 in both.
#include "sd_internal.h"
#include "sd_tc.cuh"
#include "sd_mixmat.cuh"
#include <math.h>
#include <stdlib.h>

namespace sd {

template <typename T> struct Io;
template <> struct Io<float> {
    static __device__ __forceinline__ float4 load4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
    static __device__ __forceinline__ void store4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
};
template <> struct Io<__nv_bfloat16> {
    static __device__ __forceinline__ float4 load4(const __nv_bfloat16* p) {
        const uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
        return make_float4(__uint_as_float(t.x << 16), __uint_as_float(t.x & 0xFFFF0000u), __uint_as_float(t.y << 16), __uint_as_float(t.y & 0xFFFF0000u));
    }
    static __device__ __forceinline__ void store4(__nv_bfloat16* p, float4 v) {
        __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
        uint2 t; t.x = *reinterpret_cast<uint32_t*>(&a); t.y = *reinterpret_cast<uint32_t*>(&b);
        *reinterpret_cast<uint2*>(p) = t;
    }
};

__device__ __forceinline__ void att_ffma2(float2& d, float2 a, float2 b) {
    unsigned long long dd = *reinterpret_cast<unsigned long long*>(&d), aa = *reinterpret_cast<unsigned long long*>(&a), bb = *reinterpret_cast<unsigned long long*>(&b);
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(aa), "l"(bb));
    d = *reinterpret_cast<float2*>(&dd);
}
__device__ __forceinline__ void att_ffma2s(float2& d, float a, float2 b) {
    unsigned long long dd = *reinterpret_cast<unsigned long long*>(&d), bb = *reinterpret_cast<unsigned long long*>(&b), aa;
    asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(aa), "l"(bb));
    d = *reinterpret_cast<float2*>(&dd);
}

constexpr int ATT_WARPS = 4;

// NEXACT > 0: the node count is a compile-time constant (all index arithmetic of the tile copy folds away and the
// score loops carry no guards); NEXACT == 0: any N <= NMAX at run time.
template <typename T, int DH, int NMAX, int NEXACT>
__global__ void __launch_bounds__(ATT_WARPS * 32)
node_attention_kernel(const T* __restrict__ qkv, T* __restrict__ out, long long tasks, int N_rt, int H) {
    const int N = NEXACT > 0 ? NEXACT : N_rt;
    constexpr int LD = DH + 4;                       // padded row (floats)
    extern __shared__ __align__(16) float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* Qs = smem + (size_t)warp * 3 * NMAX * LD;
    float* Ks = Qs + NMAX * LD;
    float* Vs = Ks + NMAX * LD;
    const int HD = H * DH;
    const long long row_stride = 3LL * HD;
    const float scale = rsqrtf((float)DH);
    for (long long task = (long long)blockIdx.x * ATT_WARPS + warp; task < tasks; task += (long long)gridDim.x * ATT_WARPS) {
        const long long b = task / H;
        const int h = (int)(task % H);
        const T* base = qkv + b * N * row_stride + h * DH;
        __syncwarp();                                // previous task's readers are done with the tiles
        // tile copy in batches of 8 loads per lane: all 8 are in flight before the first shared-memory store
        // (a load -> store loop issues in order and pays one DRAM latency per iteration)
        const int per_mat = N * (DH / 4), total4 = 3 * per_mat;
        for (int i0 = lane; i0 < total4; i0 += 32 * 8) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = i0 + 32 * u;
                if (i < total4) {
                    const int which = i / per_mat, rem = i - which * per_mat;
                    const int j = rem / (DH / 4), c4 = rem % (DH / 4);
                    v[u] = Io<T>::load4(base + j * row_stride + which * HD + 4 * c4);
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = i0 + 32 * u;
                if (i < total4) {
                    const int which = i / per_mat, rem = i - which * per_mat;
                    const int j = rem / (DH / 4), c4 = rem % (DH / 4);
                    float4 t = v[u];
                    if (which == 0) { t.x *= scale; t.y *= scale; t.z *= scale; t.w *= scale; }      // q * dh^-1/2 (attention.py:128)
                    *reinterpret_cast<float4*>(Qs + which * NMAX * LD + j * LD + 4 * c4) = t;
                }
            }
        }
        __syncwarp();
        for (int n = lane; n < N; n += 32) {
            float2 q[DH / 2];
#pragma unroll
            for (int c = 0; c < DH; c += 4) {
                const float4 t = *reinterpret_cast<const float4*>(Qs + n * LD + c);
                q[c / 2] = make_float2(t.x, t.y); q[c / 2 + 1] = make_float2(t.z, t.w);
            }
            float sc[NMAX];
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < NMAX; ++j) {
                if (j < N) {
                    float2 s2[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
                    for (int c = 0; c < DH; c += 4) {       // four independent accumulator chains
                        const float4 kv = *reinterpret_cast<const float4*>(Ks + j * LD + c);
                        att_ffma2(s2[(c / 4) & 1], q[c / 2], make_float2(kv.x, kv.y));
                        att_ffma2(s2[2 + ((c / 4) & 1)], q[c / 2 + 1], make_float2(kv.z, kv.w));
                    }
                    sc[j] = ((s2[0].x + s2[1].x) + (s2[2].x + s2[3].x)) + ((s2[0].y + s2[1].y) + (s2[2].y + s2[3].y));
                    mx = fmaxf(mx, sc[j]);
                }
            }
            float2 acc[DH / 2];
#pragma unroll
            for (int c = 0; c < DH / 2; ++c) acc[c] = make_float2(0.f, 0.f);
            float sum = 0.0f;
#pragma unroll
            for (int j = 0; j < NMAX; ++j) {
                if (j < N) {
                    const float pj = expf(sc[j] - mx);
                    sum += pj;
#pragma unroll
                    for (int c = 0; c < DH; c += 4) {
                        const float4 vv = *reinterpret_cast<const float4*>(Vs + j * LD + c);
                        att_ffma2s(acc[c / 2], pj, make_float2(vv.x, vv.y));
                        att_ffma2s(acc[c / 2 + 1], pj, make_float2(vv.z, vv.w));
                    }
                }
            }
            const float inv = 1.0f / sum;
            T* o = out + (b * N + n) * HD + h * DH;
#pragma unroll
            for (int c = 0; c < DH; c += 4)
                Io<T>::store4(o + c, make_float4(acc[c / 2].x * inv, acc[c / 2].y * inv, acc[c / 2 + 1].x * inv, acc[c / 2 + 1].y * inv));
        }
    }
}

template <typename T, int DH, int NMAX, int NEXACT>
static int launch_attention(const T* qkv, T* out, int B, int N, int H, cudaStream_t st) {
    const size_t smem = (size_t)ATT_WARPS * 3 * NMAX * (DH + 4) * sizeof(float);
    auto kern = node_attention_kernel<T, DH, NMAX, NEXACT>;
    static unsigned long long configured = 0;      // bit d: attribute set on device d (it is per device)
    if (smem > 48 * 1024) if (int rc_attr = opt_in_smem(kern, (size_t)(smem), configured)) return rc_attr;
    const long long tasks = (long long)B * H;
    long long blocks = (tasks + ATT_WARPS - 1) / ATT_WARPS;
    const long long cap = 148LL * 8;                 // persistent-style grid: a few CTAs per SM, tasks grid-strided
    if (blocks > cap) blocks = cap;
    kern<<<(unsigned)blocks, ATT_WARPS * 32, smem, st>>>(qkv, out, tasks, N, H);
    SD_LAUNCH_OK("node_attention_kernel");
    return SD_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// Bulk-copy pipeline (fp32, 8 heads x 32 channels): the shipped Denoiser configuration.
//
// One sample's qkv rows are ONE contiguous block in HBM (N x 768 floats = 64.5 KB for AMASS) and so is its output
// (N x 256 floats).  A persistent CTA streams whole samples through a 3-deep shared-memory ring with one
// cp.async.bulk per sample issued by a copy warp (no register staging, no per-warp DRAM latency), eight compute warps
// (warp = head, lane = query node) read Q, K and V IN PLACE, write the normalised output over their own q slice, and
// the copy warp writes it back with bulk stores; the compute warps never synchronise with each other.
// The rows of a block are 3072 B apart, i.e. every row starts at bank 0: lanes reading THEIR OWN row chunk by chunk
// would conflict 8 ways.  Lane n therefore visits the eight 16-byte chunks of its q row (and of its output row) in
// the rotated order (i + n) & 7, which 8 consecutive lanes serve from 8 different bank groups, and undoes the
// rotation in registers with a 3-stage conditional rotate (SELs, no LSU traffic).  K and V rows are read in natural
// order by all lanes at once: pure broadcasts, one LSU wavefront per LDS.128 (a first version rotated those reads as
// well, which made every one of the 336 K/V loads per head a 4-wavefront access and the kernel LSU-bound at 1.1 ms).
constexpr int AB_HEADS = 8, AB_DH = 32, AB_STAGES = 3;
constexpr int AB_GROUPS = 1;                         // warp groups (group g takes the CTA's samples k = g, g + GROUPS, ...); 2 groups = 17 warps
                                                     // = 96 registers and only one stage left to load into: 601 us vs 577 us
constexpr int AB_COPY_WARP = AB_GROUPS * AB_HEADS;
constexpr int AB_THREADS = (AB_COPY_WARP + 1) * 32;  // compute warps + 1 copy warp

// MIX variant: 8 attention warps + the mix warps + the copy warp.  First version: 8 mix warps, one column per thread and pass, scalar
// FFMAs (17 warps).  A version with 20 warps
// and a setmaxnreg hand-over (copy / idle group 24, mix 88, attention 152) passed every small test and died with an illegal
// instruction at B = 25 600 (the same kernel without the three setmaxnreg instructions is correct); not pursued, it is not needed.
// six mix warps: 192 threads x 4 columns = two FFMA2 column pairs per thread; 15 warps leave 128 registers per thread (17 warps: 96,
// and the pairwise mix - 42 + 42 live values - spilled 1.4 KB)
constexpr int AB_MIX_WARP0 = AB_HEADS, AB_MIX_WARPS = 6, AB_MIX_COPY_WARP = AB_MIX_WARP0 + AB_MIX_WARPS, AB_THREADS_MIX = (AB_MIX_COPY_WARP + 1) * 32;
struct __align__(8) AbBarriers { uint64_t full[AB_STAGES], done[AB_STAGES], mixed[AB_STAGES]; };

__device__ __forceinline__ void ab_bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(tc::smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(tc::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void ab_bulk_store(void* gdst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(gdst), "r"(tc::smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ab_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void ab_bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void ab_bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// x[i] <- x[(i - r) & 7] (DIR = -1) or x[(i + r) & 7] (DIR = +1) for every set bit r of n: rotation of 8 float4 slots by n
template <int DIR>
__device__ __forceinline__ void ab_rotate8(float4 (&x)[8], int n) {
#pragma unroll
    for (int r = 1; r < 8; r <<= 1) {
        const bool on = (n & r) != 0;
        float4 t[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) t[i] = x[(i + DIR * r) & 7];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            x[i].x = on ? t[i].x : x[i].x; x[i].y = on ? t[i].y : x[i].y;
            x[i].z = on ? t[i].z : x[i].z; x[i].w = on ? t[i].w : x[i].w;
        }
    }
}

// MIX: qkv holds the RAW per-node products of to_qkv and the graph-influence mix of that layer (graph_structural.py:41) is done
// here, in place in shared memory, before the heads read the sample: thread = column (768 columns over 256 threads), the N
// values of the column scaled by the RMSNorm row factor, N x N FFMAs with G^ in the constant bank (kernel parameter, see
// sd_mix.cu), written back to the same N slots.  The mixed qkv tensor never exists in HBM.  The mix has its own eight warps and
// runs one sample ahead of the heads (stages of the ring: loading / being mixed / being attended); a first version did mix and
// attention back to back in the same eight warps (1.16 ms at B = 25 600 against 0.59 ms for the attention alone).
template <int N, bool MIX>
__global__ void __launch_bounds__(MIX ? AB_THREADS_MIX : AB_THREADS, 1)
node_attention_bulk_kernel(const float* __restrict__ qkv, float* __restrict__ out, int B, const __grid_constant__ MixMat<N> G,
                           const float* __restrict__ row_scale) {
    constexpr int ROW = 3 * AB_HEADS * AB_DH;                       // 768 floats per (sample, node) row: q | k | v, each [head][32]
    constexpr int IN_FLOATS = N * ROW, OUT_ROW = AB_HEADS * AB_DH, OUT_FLOATS = N * OUT_ROW;
    constexpr uint32_t IN_BYTES = IN_FLOATS * 4u, OUT_BYTES = OUT_FLOATS * 4u;
    extern __shared__ __align__(128) float ab_smem[];
    float* in_buf = ab_smem;                                        // [AB_STAGES][IN_FLOATS]
    AbBarriers* bars = reinterpret_cast<AbBarriers*>(in_buf + AB_STAGES * IN_FLOATS);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < AB_STAGES; ++s) {
            tc::mbar_init(&bars->full[s], 1); tc::mbar_init(&bars->done[s], AB_HEADS); tc::mbar_init(&bars->mixed[s], AB_MIX_WARPS);
        }
        tc::fence_barrier_init();
    }
    __syncthreads();
    const int my_samples = blockIdx.x < B ? (B - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    if (MIX && warp >= AB_MIX_WARP0 && warp < AB_MIX_COPY_WARP) {
        // ------------------------------------------------------------ mix warps: sample k + 1 is mixed while the heads attend to k
        // (thread = column, the N values of the column scaled by the RMSNorm row factor, N x N FFMAs with G^ in the constant
        // bank, written back to the same N slots); mixed[stage] tells the head warps that the slab is ready.
        const int mt = (int)threadIdx.x - AB_MIX_WARP0 * 32;
        for (int k = 0; k < my_samples; ++k) {
            const int stage = k % AB_STAGES;
            tc::mbar_wait(&bars->full[stage], (uint32_t)(k / AB_STAGES) & 1u, 100000 + k);
            float* slab = in_buf + stage * IN_FLOATS;
            const long long b = (long long)blockIdx.x + (long long)k * gridDim.x;
            const float* rsp = row_scale ? row_scale + b * N : nullptr;       // re-read per pass (L1 broadcast): no N registers held
            constexpr int MT = AB_MIX_WARPS * 32;
            static_assert(ROW == 4 * MT, "four columns per mix thread: two FFMA2 pairs");
#pragma unroll 1
            for (int pr = 0; pr < 2; ++pr) {   // columns (mt, mt + 192) and (mt + 384, mt + 576): 441 issue slots per 882 products
                const int c0 = mt + 2 * pr * MT;
                float in[N][2], acc[N][2];
#pragma unroll
                for (int m = 0; m < N; ++m) {
                    const float r = rsp ? __ldg(rsp + m) : 1.0f;
                    in[m][0] = slab[m * ROW + c0] * r; in[m][1] = slab[m * ROW + c0 + MT] * r;
                }
                mix_nodes<N, 2>(G, in, acc);
#pragma unroll
                for (int n2 = 0; n2 < N; ++n2) { slab[n2 * ROW + c0] = acc[n2][0]; slab[n2 * ROW + c0 + MT] = acc[n2][1]; }
            }
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&bars->mixed[stage]);
        }
        return;
    }
    if (warp == (MIX ? AB_MIX_COPY_WARP : AB_COPY_WARP)) {
        // ------------------------------------------------------------ copy warp: bulk load per sample, bulk stores of its result
        // Head h writes its normalised output over ITS OWN q slice of the stage (a q row and an output row are both
        // [8 heads][32] floats), so the compute warps never synchronise with each other: each arrives on done[stage]
        // when its slice is written, and this thread stores the N output rows (1 KB each, 3 KB apart in shared memory),
        // waits until they have been read and re-arms the stage with the sample three positions ahead.
        if (lane == 0) {
            auto load = [&](int k) {
                const int st = k % AB_STAGES;
                tc::mbar_arrive_expect_tx(&bars->full[st], IN_BYTES);
                ab_bulk_load(in_buf + st * IN_FLOATS, qkv + ((long long)blockIdx.x + (long long)k * gridDim.x) * IN_FLOATS, IN_BYTES, &bars->full[st]);
            };
            for (int k = 0; k < AB_STAGES && k < my_samples; ++k) load(k);
            for (int k = 0; k < my_samples; ++k) {
                const int st = k % AB_STAGES;
                tc::mbar_wait(&bars->done[st], (uint32_t)(k / AB_STAGES) & 1u, 300000 + k);
                float* dst = out + ((long long)blockIdx.x + (long long)k * gridDim.x) * OUT_FLOATS;
                const float* src = in_buf + st * IN_FLOATS;
#pragma unroll 1
                for (int r = 0; r < N; ++r) ab_bulk_store(dst + r * OUT_ROW, src + r * ROW, OUT_ROW * 4u);
                ab_bulk_commit();
                ab_bulk_wait_read();                                // the stage may be overwritten
                if (k + AB_STAGES < my_samples) load(k + AB_STAGES);
            }
            ab_bulk_wait_all();                                     // shared memory must outlive the last store
        }
        return;
    }
    // ---------------------------------------------------------------- compute warps: warp = head, lane = query node
    const int h = warp % AB_HEADS, group = warp / AB_HEADS, n = lane;
    const bool active = n < N;
    const float scale = rsqrtf((float)AB_DH);
    int rot[8];                                                     // float offset of the chunk visited at position i
#pragma unroll
    for (int i = 0; i < 8; ++i) rot[i] = 4 * ((i + n) & 7);
    for (int k = group; k < my_samples; k += AB_GROUPS) {
        const int stage = k % AB_STAGES;
        tc::mbar_wait(MIX ? &bars->mixed[stage] : &bars->full[stage], (uint32_t)(k / AB_STAGES) & 1u, 200000 + k);
        float* blk = in_buf + stage * IN_FLOATS + h * AB_DH;
        if (active) {
            float2 q[16];
            float* qrow = blk + n * ROW;
            {
                float4 qs[8];                                       // slot i = chunk (i + n) & 7
#pragma unroll
                for (int i = 0; i < 8; ++i) qs[i] = *reinterpret_cast<const float4*>(qrow + rot[i]);
                ab_rotate8<-1>(qs, n);                              // slot k = chunk k
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    q[2 * i] = make_float2(qs[i].x * scale, qs[i].y * scale);   // q * dh^-1/2 (attention.py:128)
                    q[2 * i + 1] = make_float2(qs[i].z * scale, qs[i].w * scale);
                }
            }
            float sc[N];
            float mx = -INFINITY;
            const float* kbase = blk + AB_HEADS * AB_DH;
#pragma unroll
            for (int j = 0; j < N; ++j) {
                float2 s2[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
                for (int i = 0; i < 8; ++i) {                       // four independent accumulator chains
                    const float4 kv = *reinterpret_cast<const float4*>(kbase + j * ROW + 4 * i);       // broadcast
                    att_ffma2(s2[i & 1], q[2 * i], make_float2(kv.x, kv.y));
                    att_ffma2(s2[2 + (i & 1)], q[2 * i + 1], make_float2(kv.z, kv.w));
                }
                sc[j] = ((s2[0].x + s2[1].x) + (s2[2].x + s2[3].x)) + ((s2[0].y + s2[1].y) + (s2[2].y + s2[3].y));
                mx = fmaxf(mx, sc[j]);
            }
            float2 acc[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) acc[c] = make_float2(0.f, 0.f);
            float sum = 0.0f;
            const float* vbase = blk + 2 * AB_HEADS * AB_DH;
#pragma unroll
            for (int j = 0; j < N; ++j) {
                const float pj = expf(sc[j] - mx);
                sum += pj;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 vv = *reinterpret_cast<const float4*>(vbase + j * ROW + 4 * i);       // broadcast
                    att_ffma2s(acc[2 * i], pj, make_float2(vv.x, vv.y));
                    att_ffma2s(acc[2 * i + 1], pj, make_float2(vv.z, vv.w));
                }
            }
            const float inv = 1.0f / sum;
            float4 os[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) os[i] = make_float4(acc[2 * i].x * inv, acc[2 * i].y * inv, acc[2 * i + 1].x * inv, acc[2 * i + 1].y * inv);
            ab_rotate8<1>(os, n);                                   // slot i = chunk (i + n) & 7
#pragma unroll
            for (int i = 0; i < 8; ++i) *reinterpret_cast<float4*>(qrow + rot[i]) = os[i];     // in place over this head's q slice
        }
        tc::fence_proxy_async();                                    // generic-proxy writes -> visible to the bulk stores (async proxy)
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&bars->done[stage]);
    }
}

template <int N, bool MIX>
static int launch_attention_bulk(const float* qkv, float* out, int B, cudaStream_t st, const float* G_host = nullptr, const float* row_scale = nullptr) {
    constexpr size_t smem = (size_t)(AB_STAGES * N * 3) * AB_HEADS * AB_DH * sizeof(float) + sizeof(AbBarriers) + 128;
    static_assert(smem <= 227 * 1024, "attention ring does not fit shared memory");
    MixMat<N> G;
    G.set(MIX ? G_host : nullptr);
    auto kern = node_attention_bulk_kernel<N, MIX>;
    static unsigned long long configured = 0;      // bit d: attribute set on device d (it is per device)
    if (int rc_attr = opt_in_smem(kern, (size_t)(smem), configured)) return rc_attr;
    const int sms = sm_count();
    const int grid = B < sms ? B : sms;                             // persistent: one CTA per SM, samples grid-strided
    kern<<<grid, MIX ? AB_THREADS_MIX : AB_THREADS, smem, st>>>(qkv, out, B, G, row_scale);
    SD_LAUNCH_OK("node_attention_bulk_kernel");
    return SD_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// bf16 I/O variant of the bulk pipeline (bf16 precision mode): same structure, a sample is N x 1536 B, a head's slice of a
// row is 64 B = four 16-byte chunks of 8 bf16.  Rotation over 4 chunks (done on the packed words, before unpacking); with
// 8 lanes per shared-memory phase two lanes share a bank group (2-way conflict) on the q load / output store only.
constexpr int AB16_STAGES = 5;

template <int DIR>
__device__ __forceinline__ void ab_rotate4(uint4 (&x)[4], int n) {
#pragma unroll
    for (int r = 1; r < 4; r <<= 1) {
        const bool on = (n & r) != 0;
        uint4 t[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) t[i] = x[(i + DIR * r) & 3];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            x[i].x = on ? t[i].x : x[i].x; x[i].y = on ? t[i].y : x[i].y;
            x[i].z = on ? t[i].z : x[i].z; x[i].w = on ? t[i].w : x[i].w;
        }
    }
}
__device__ __forceinline__ float2 ab_unpack(uint32_t w) { return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xFFFF0000u)); }
__device__ __forceinline__ uint32_t ab_pack(float a, float b) {
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&t);
}

struct __align__(8) Ab16Barriers { uint64_t full[AB16_STAGES], done[AB16_STAGES]; };

template <int N>
__global__ void __launch_bounds__(AB_THREADS, 1)
node_attention_bulk_bf16_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int B) {
    constexpr int ROW = 3 * AB_HEADS * AB_DH;                       // 768 bf16 per (sample, node) row
    constexpr int IN_ELEMS = N * ROW, OUT_ROW = AB_HEADS * AB_DH, OUT_ELEMS = N * OUT_ROW;
    constexpr uint32_t IN_BYTES = IN_ELEMS * 2u;
    extern __shared__ __align__(128) float ab_smem[];
    __nv_bfloat16* in_buf = reinterpret_cast<__nv_bfloat16*>(ab_smem);          // [AB16_STAGES][IN_ELEMS]
    Ab16Barriers* bars = reinterpret_cast<Ab16Barriers*>(in_buf + AB16_STAGES * IN_ELEMS);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < AB16_STAGES; ++s) { tc::mbar_init(&bars->full[s], 1); tc::mbar_init(&bars->done[s], AB_HEADS); }
        tc::fence_barrier_init();
    }
    __syncthreads();
    const int my_samples = blockIdx.x < B ? (B - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    if (warp == AB_COPY_WARP) {
        if (lane == 0) {
            auto load = [&](int k) {
                const int st = k % AB16_STAGES;
                tc::mbar_arrive_expect_tx(&bars->full[st], IN_BYTES);
                ab_bulk_load(in_buf + st * IN_ELEMS, qkv + ((long long)blockIdx.x + (long long)k * gridDim.x) * IN_ELEMS, IN_BYTES, &bars->full[st]);
            };
            for (int k = 0; k < AB16_STAGES && k < my_samples; ++k) load(k);
            for (int k = 0; k < my_samples; ++k) {
                const int st = k % AB16_STAGES;
                tc::mbar_wait(&bars->done[st], (uint32_t)(k / AB16_STAGES) & 1u);
                __nv_bfloat16* dst = out + ((long long)blockIdx.x + (long long)k * gridDim.x) * OUT_ELEMS;
                const __nv_bfloat16* src = in_buf + st * IN_ELEMS;
#pragma unroll 1
                for (int r = 0; r < N; ++r) ab_bulk_store(dst + r * OUT_ROW, src + r * ROW, OUT_ROW * 2u);
                ab_bulk_commit();
                ab_bulk_wait_read();
                if (k + AB16_STAGES < my_samples) load(k + AB16_STAGES);
            }
            ab_bulk_wait_all();
        }
        return;
    }
    const int h = warp % AB_HEADS, n = lane;
    const bool active = n < N;
    const float scale = rsqrtf((float)AB_DH);
    int rot[4];                                                     // bf16 offset of the chunk visited at position i
#pragma unroll
    for (int i = 0; i < 4; ++i) rot[i] = 8 * ((i + n) & 3);
    for (int k = 0; k < my_samples; ++k) {
        const int stage = k % AB16_STAGES;
        tc::mbar_wait(&bars->full[stage], (uint32_t)(k / AB16_STAGES) & 1u);
        __nv_bfloat16* blk = in_buf + stage * IN_ELEMS + h * AB_DH;
        if (active) {
            __nv_bfloat16* qrow = blk + n * ROW;
            float2 q[16];
            {
                uint4 qs[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) qs[i] = *reinterpret_cast<const uint4*>(qrow + rot[i]);
                ab_rotate4<-1>(qs, n);                              // slot c = chunk c
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint32_t w[4] = {qs[i].x, qs[i].y, qs[i].z, qs[i].w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 t = ab_unpack(w[e]);
                        q[4 * i + e] = make_float2(t.x * scale, t.y * scale);    // q * dh^-1/2 (attention.py:128)
                    }
                }
            }
            float sc[N];
            float mx = -INFINITY;
            const __nv_bfloat16* kbase = blk + AB_HEADS * AB_DH;
#pragma unroll
            for (int j = 0; j < N; ++j) {
                float2 s2[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint4 kv = *reinterpret_cast<const uint4*>(kbase + j * ROW + 8 * i);          // broadcast
                    att_ffma2(s2[0], q[4 * i + 0], ab_unpack(kv.x));
                    att_ffma2(s2[1], q[4 * i + 1], ab_unpack(kv.y));
                    att_ffma2(s2[2], q[4 * i + 2], ab_unpack(kv.z));
                    att_ffma2(s2[3], q[4 * i + 3], ab_unpack(kv.w));
                }
                sc[j] = ((s2[0].x + s2[1].x) + (s2[2].x + s2[3].x)) + ((s2[0].y + s2[1].y) + (s2[2].y + s2[3].y));
                mx = fmaxf(mx, sc[j]);
            }
            float2 acc[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) acc[c] = make_float2(0.f, 0.f);
            float sum = 0.0f;
            const __nv_bfloat16* vbase = blk + 2 * AB_HEADS * AB_DH;
#pragma unroll
            for (int j = 0; j < N; ++j) {
                const float pj = expf(sc[j] - mx);
                sum += pj;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint4 vv = *reinterpret_cast<const uint4*>(vbase + j * ROW + 8 * i);          // broadcast
                    att_ffma2s(acc[4 * i + 0], pj, ab_unpack(vv.x));
                    att_ffma2s(acc[4 * i + 1], pj, ab_unpack(vv.y));
                    att_ffma2s(acc[4 * i + 2], pj, ab_unpack(vv.z));
                    att_ffma2s(acc[4 * i + 3], pj, ab_unpack(vv.w));
                }
            }
            const float inv = 1.0f / sum;
            uint4 os[4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
                os[i] = make_uint4(ab_pack(acc[4 * i].x * inv, acc[4 * i].y * inv), ab_pack(acc[4 * i + 1].x * inv, acc[4 * i + 1].y * inv),
                                   ab_pack(acc[4 * i + 2].x * inv, acc[4 * i + 2].y * inv), ab_pack(acc[4 * i + 3].x * inv, acc[4 * i + 3].y * inv));
            ab_rotate4<1>(os, n);                                   // slot i = chunk (i + n) & 3
#pragma unroll
            for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(qrow + rot[i]) = os[i];       // in place over this head's q slice
        }
        tc::fence_proxy_async();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&bars->done[stage]);
    }
}

template <int N>
static int launch_attention_bulk_bf16(const __nv_bfloat16* qkv, __nv_bfloat16* out, int B, cudaStream_t st) {
    constexpr size_t smem = (size_t)(AB16_STAGES * N * 3) * AB_HEADS * AB_DH * sizeof(__nv_bfloat16) + sizeof(Ab16Barriers) + 128;
    static_assert(smem <= 227 * 1024, "bf16 attention ring does not fit shared memory");
    auto kern = node_attention_bulk_bf16_kernel<N>;
    static unsigned long long configured = 0;      // bit d: attribute set on device d (it is per device)
    if (int rc_attr = opt_in_smem(kern, (size_t)(smem), configured)) return rc_attr;
    const int sms = sm_count();
    const int grid = B < sms ? B : sms;
    kern<<<grid, AB_THREADS, smem, st>>>(qkv, out, B);
    SD_LAUNCH_OK("node_attention_bulk_bf16_kernel");
    return SD_OK;
}

static bool attention_legacy_forced() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SKELDIFF_ATTENTION_LEGACY"); v = (e && e[0] == '1') ? 1 : 0; }
    return v == 1;
}

template <typename T> struct BulkAttention {
    static bool run(const T*, T*, int, int, int, int, cudaStream_t, int*) { return false; }
};
template <> struct BulkAttention<float> {
    // returns true when the shape was handled (rc holds the status)
    static bool run(const float* qkv, float* out, int B, int N, int heads, int dh, cudaStream_t st, int* rc) {
        if (heads != AB_HEADS || dh != AB_DH || attention_legacy_forced()) return false;
        if ((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(out)) & 15u) return false;
        if (N == 21) { *rc = launch_attention_bulk<21, false>(qkv, out, B, st); return true; }    // AMASS
        if (N == 16) { *rc = launch_attention_bulk<16, false>(qkv, out, B, st); return true; }    // H36M, README
        if (N == 17) { *rc = launch_attention_bulk<17, false>(qkv, out, B, st); return true; }    // FreeMan
        return false;
    }
};

template <> struct BulkAttention<__nv_bfloat16> {
    static bool run(const __nv_bfloat16* qkv, __nv_bfloat16* out, int B, int N, int heads, int dh, cudaStream_t st, int* rc) {
        if (heads != AB_HEADS || dh != AB_DH || attention_legacy_forced()) return false;
        if ((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(out)) & 15u) return false;
        if (N == 21) { *rc = launch_attention_bulk_bf16<21>(qkv, out, B, st); return true; }
        if (N == 16) { *rc = launch_attention_bulk_bf16<16>(qkv, out, B, st); return true; }
        if (N == 17) { *rc = launch_attention_bulk_bf16<17>(qkv, out, B, st); return true; }
        return false;
    }
};

template <typename T>
static int node_attention_any(const T* qkv, T* out, int B, int N, int heads, int dh, cudaStream_t st) {
    if (B <= 0) return SD_OK;
    int rc = SD_OK;
    if (BulkAttention<T>::run(qkv, out, B, N, heads, dh, st, &rc)) return rc;
    if (dh == 32 && N == 21) return launch_attention<T, 32, 21, 21>(qkv, out, B, N, heads, st);    // AMASS
    if (dh == 32 && N == 16) return launch_attention<T, 32, 16, 16>(qkv, out, B, N, heads, st);    // H36M, README
    if (dh == 32 && N == 17) return launch_attention<T, 32, 17, 17>(qkv, out, B, N, heads, st);    // FreeMan
    if (dh == 32 && N <= 32) return launch_attention<T, 32, 32, 0>(qkv, out, B, N, heads, st);
    if (dh == 32 && N <= 64) return launch_attention<T, 32, 64, 0>(qkv, out, B, N, heads, st);
    if (dh == 16 && N <= 32) return launch_attention<T, 16, 32, 0>(qkv, out, B, N, heads, st);
    if (dh == 64 && N <= 32) return launch_attention<T, 64, 32, 0>(qkv, out, B, N, heads, st);
    set_error("node_attention: dim_head %d with %d nodes unsupported (dim_head 32: N<=64; 16/64: N<=32)", dh, N);
    return SD_ERR_UNSUPPORTED;
}

bool node_attention_mix_supported(int N, int heads, int dh, const float* qkv, const float* out) {
    static int off = -1;                         // SKELDIFF_NO_ATT_MIX=1: separate mix pass (A/B timing, bisection)
    if (off < 0) { const char* e = getenv("SKELDIFF_NO_ATT_MIX"); off = (e && e[0] == '1') ? 1 : 0; }
    if (off) return false;
    if (heads != AB_HEADS || dh != AB_DH || attention_legacy_forced()) return false;
    if ((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(out)) & 15u) return false;
    return N == 21 || N == 16 || N == 17;
}
int node_attention_mix_fp32(const float* G_host, const float* row_scale, const float* qkv, float* out, int B, int N, int heads, int dh, cudaStream_t st) {
    if (B <= 0) return SD_OK;
    if (!G_host || !node_attention_mix_supported(N, heads, dh, qkv, out)) { set_error("node_attention_mix: unsupported shape"); return SD_ERR_UNSUPPORTED; }
    if (N == 21) return launch_attention_bulk<21, true>(qkv, out, B, st, G_host, row_scale);
    if (N == 16) return launch_attention_bulk<16, true>(qkv, out, B, st, G_host, row_scale);
    return launch_attention_bulk<17, true>(qkv, out, B, st, G_host, row_scale);
}

int node_attention_fp32(const float* qkv, float* out, int B, int N, int heads, int dh, cudaStream_t st) {
    return node_attention_any<float>(qkv, out, B, N, heads, dh, st);
}
int node_attention_bf16(const __nv_bfloat16* qkv, __nv_bfloat16* out, int B, int N, int heads, int dh, cudaStream_t st) {
    return node_attention_any<__nv_bfloat16>(qkv, out, B, N, heads, dh, st);
}

}  // namespace sd
