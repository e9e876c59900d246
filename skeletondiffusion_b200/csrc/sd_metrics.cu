// Evaluation metrics of the sampling path's output: ADE / FDE / APD per observed window, one launch.
// Replaces ade / fde / apd of the reference (src/metrics/multimodal.py:44-57, :60-73, :15-35) as eval.py calls them on
// skeleton.transform_to_metric_space(pred) (rescalepose.py:29-39: a multiplication by pose_box_size, which every
// one of the three metrics is linear in, so it is applied once to the three results).
//
// One CTA per window.  The window's S predictions [S, T, F] (AMASS: 50 x 120 x 63 floats = 1.5 MB) are read from HBM
// exactly once, a few frames at a time, into shared memory ([S][frames*F] with an odd row stride, so lanes that walk
// different samples hit different banks).  While a chunk is resident
//   - one thread per (sample, frame) computes ||pred[s, f] - target[f]||; sample s's owner adds them in frame order
//     (ADE sum, FDE = last frame),
//   - the S x S pair matrix (APD) is accumulated in 5 x 5 register tiles: 10 shared-memory loads per 25 pairs and column.
// Reductions are fixed-order (one owner per sample / pair, tree sum), so results are bitwise repeatable.
// Algorithmic bytes per window: 4*(S+1)*T*F read + 12 written.  See DESIGN.md 4.7.
#include "sd_internal.h"

namespace sd {

constexpr int MM_THREADS = 256;
constexpr int MM_TB      = 5;      // edge of a register tile of sample pairs
constexpr int MM_SMAX    = 91;     // 19 x 20 / 2 = 190 tiles <= MM_THREADS

__global__ void __launch_bounds__(MM_THREADS, 4)
motion_metrics_kernel(const float* __restrict__ pred, const float* __restrict__ target, const int* __restrict__ pred_index, int S, int T, int F,
                      int fc, int stride, int row_floats, float scale, float* __restrict__ ade, float* __restrict__ fde, float* __restrict__ apd) {
    extern __shared__ float sm[];
    float* rows  = sm;                          // [S][stride]
    float* tgt   = rows + row_floats;           // [fc*F]
    float* asum  = tgt + fc * F;                // [S] sum over frames of the per-frame distance
    float* flast = asum + S;                    // [S] distance at the last frame
    float* red   = flast + S;                   // [MM_THREADS]
    float* fd    = red + MM_THREADS;            // [S][fc] per-frame distances of the resident chunk
    const int w = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long D = (long long)T * F;
    // pred_index (multimodal ground truths, mmade / mmfde): target row w is compared with the predictions of window pred_index[w]
    const float* pw = pred + (long long)(pred_index ? __ldg(pred_index + w) : w) * S * D;
    const float* tw = target + (long long)w * D;
    const int pairs = S * (S - 1) / 2;

    // APD: the S x S pair matrix is cut into MM_TB x MM_TB register tiles (upper triangle, diagonal tiles included);
    // thread (tile, q) owns the tile's 25 accumulators for the columns c = q, q + Q, ...  Lanes of a warp hold different
    // tiles and the same q: their rows differ by multiples of MM_TB, which an odd row stride spreads over distinct banks.
    const int nb = (S + MM_TB - 1) / MM_TB, ntiles = nb * (nb + 1) / 2;
    const int Q = max(1, MM_THREADS / ntiles);
    const int tile = tid % ntiles, q = tid / ntiles;
    const bool pair_thread = q < Q && apd != nullptr;
    int tI = 0, tJ = tile;
    while (tJ >= nb - tI) { tJ -= nb - tI; ++tI; }
    tJ += tI;
    int ra[MM_TB], rb[MM_TB];
    float acc[MM_TB][MM_TB];
#pragma unroll
    for (int u = 0; u < MM_TB; ++u) {
        ra[u] = min(tI * MM_TB + u, S - 1) * stride;
        rb[u] = min(tJ * MM_TB + u, S - 1) * stride;
#pragma unroll
        for (int v = 0; v < MM_TB; ++v) acc[u][v] = 0.f;
    }
    for (int s = tid; s < S; s += MM_THREADS) { asum[s] = 0.f; flast[s] = 0.f; }

    for (int f0 = 0; ; f0 += fc) {              // one extra turn (f0 >= T) folds the last chunk's frame distances
        const int nf = min(fc, T - f0), len = nf * F;
        __syncthreads();                        // previous chunk fully consumed (and asum initialised)
        if (f0 > 0) {                           // sample s adds the previous chunk's frames in frame order (fixed order)
            const int pf = min(fc, T - (f0 - fc));
            for (int s = tid; s < S; s += MM_THREADS) {
                float a = asum[s];
                for (int f = 0; f < pf; ++f) a += fd[s * fc + f];
                asum[s] = a;
                if (f0 >= T) flast[s] = fd[s * fc + pf - 1];
            }
        }
        if (f0 >= T) break;
        // rows are only 4-byte aligned (F = 63): 4-byte asynchronous copies, all of a thread's loads in flight at once
        for (int s = warp; s < S; s += MM_THREADS / 32) {
            const float* src = pw + s * D + (long long)f0 * F;
            const unsigned dst = (unsigned)__cvta_generic_to_shared(rows + (size_t)s * stride);
            for (int c = lane; c < len; c += 32)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(dst + 4u * c), "l"(src + c) : "memory");
        }
        {
            const unsigned dst = (unsigned)__cvta_generic_to_shared(tgt);
            for (int c = tid; c < len; c += MM_THREADS)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(dst + 4u * c), "l"(tw + (long long)f0 * F + c) : "memory");
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncthreads();
        for (int task = tid; task < S * nf; task += MM_THREADS) {       // one (sample, frame) distance per thread
            const int s = task / nf, f = task - s * nf;
            const float* r = rows + (size_t)s * stride + f * F;
            const float* g = tgt + f * F;
            float d2 = 0.f;
            for (int c = 0; c < F; ++c) { const float d = r[c] - g[c]; d2 = fmaf(d, d, d2); }
            fd[s * fc + f] = sqrtf(d2);
        }
        if (pair_thread) {
            for (int c = q; c < len; c += Q) {
                float av[MM_TB], bv[MM_TB];
#pragma unroll
                for (int u = 0; u < MM_TB; ++u) { av[u] = rows[ra[u] + c]; bv[u] = rows[rb[u] + c]; }
#pragma unroll
                for (int u = 0; u < MM_TB; ++u)
#pragma unroll
                    for (int v = 0; v < MM_TB; ++v) { const float d = av[u] - bv[v]; acc[u][v] = fmaf(d, d, acc[u][v]); }
            }
        }
    }

    // column shares of a tile are added in the order q = 0, 1, ... by the tile's first thread (fixed order)
    __syncthreads();
    float* part = rows;                         // [MM_THREADS][MM_TB*MM_TB], the sample rows are no longer needed
    if (pair_thread) {
#pragma unroll
        for (int u = 0; u < MM_TB; ++u)
#pragma unroll
            for (int v = 0; v < MM_TB; ++v) part[tid * (MM_TB * MM_TB) + u * MM_TB + v] = acc[u][v];
    }
    __syncthreads();
    float local = 0.f;
    if (q == 0) {
#pragma unroll
        for (int u = 0; u < MM_TB; ++u)
#pragma unroll
            for (int v = 0; v < MM_TB; ++v) {
                const int i = tI * MM_TB + u, j = tJ * MM_TB + v;
                if (i < j && j < S) {
                    float d2 = 0.f;
                    for (int qq = 0; qq < Q; ++qq) d2 += part[(tile + qq * ntiles) * (MM_TB * MM_TB) + u * MM_TB + v];
                    local += sqrtf(d2);
                }
            }
    }
    red[tid] = local;
    __syncthreads();
    for (int h = MM_THREADS / 2; h > 0; h >>= 1) {
        if (tid < h) red[tid] += red[tid + h];
        __syncthreads();
    }
    if (tid == 0) {
        float amin = asum[0], fmin_ = flast[0];
        for (int s = 1; s < S; ++s) { amin = fminf(amin, asum[s]); fmin_ = fminf(fmin_, flast[s]); }
        if (ade) ade[w] = scale * (amin / (float)T);
        if (fde) fde[w] = scale * fmin_;
        if (apd) apd[w] = pairs ? scale * (red[0] / (float)pairs) : 0.f;    // one sample: no APD possible (multimodal.py:19-20)
    }
}

static int motion_metrics_launch(const float* pred, const float* target, const int* pred_index, int windows, int samples, int frames, int feat,
                                 float scale, float* ade, float* fde, float* apd, cudaStream_t st) {
    const int S = samples, T = frames, F = feat;
    if (S > MM_SMAX) { set_error("sd_motion_metrics: at most 91 samples per window (got %d)", S); return SD_ERR_UNSUPPORTED; }
    // frames per chunk: as many as keep the sample rows within 40 KB (5 CTAs per SM); one frame at least
    int fc = (40 * 1024 / 4) / (S * F);
    fc = fc < 1 ? 1 : (fc > T ? T : fc);
    const int stride = (fc * F) | 1;
    size_t row_floats = (size_t)S * stride;                                    // reused for the tiles' partial sums at the end
    if (row_floats < (size_t)MM_THREADS * MM_TB * MM_TB) row_floats = (size_t)MM_THREADS * MM_TB * MM_TB;
    const size_t smem = (row_floats + (size_t)fc * F + 2 * (size_t)S + MM_THREADS + (size_t)S * fc) * sizeof(float);
    if (smem > 200 * 1024) { set_error("sd_motion_metrics: one frame of %d samples x %d features does not fit shared memory", S, F); return SD_ERR_UNSUPPORTED; }
    if (smem > 48 * 1024 &&
        check_cuda(cudaFuncSetAttribute(motion_metrics_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "motion_metrics smem")) return SD_ERR_CUDA;
    motion_metrics_kernel<<<windows, MM_THREADS, smem, st>>>(pred, target, pred_index, S, T, F, fc, stride, (int)row_floats, scale, ade, fde, apd);
    SD_LAUNCH_OK("motion_metrics_kernel");
    return SD_OK;
}

int motion_metrics_fp32(const float* pred, const float* target, int windows, int samples, int frames, int feat, float scale,
                        float* ade, float* fde, float* apd, cudaStream_t st) {
    return motion_metrics_launch(pred, target, nullptr, windows, samples, frames, feat, scale, ade, fde, apd, st);
}

// =====================================================================================================================
// Best-sample selection of the long-term (autoregressive) evaluation: get_best_sample_idx (src/metrics/utils.py:22-30) as used
// by long_term_prediction_best_every50 (src/eval_utils.py:44-67).  Per window: the sample with the smallest mean per-JOINT
// distance to the target segment (norm over xyz, mean over joints and frames), copied out whole (best) and its last `keep`
// frames (the next observation), both multiplied by `scale` (transform_to_metric_space).  One CTA per window; the distance of
// sample s is summed by one warp in a fixed order, so the choice is repeatable.  Replaces a .cpu() round trip per segment.
// =====================================================================================================================
__global__ void __launch_bounds__(256)
best_sample_kernel(const float* __restrict__ pred, const float* __restrict__ target, int S, int T, int J, int keep, float scale,
                   float* __restrict__ best, float* __restrict__ tail, int* __restrict__ index) {
    extern __shared__ float bs_dist[];          // [S]
    __shared__ int bs_arg;
    const int w = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const long long TJ = (long long)T * J;
    const float* tw = target + (long long)w * TJ * 3;
    for (int s = warp; s < S; s += nw) {
        const float* ps = pred + ((long long)w * S + s) * TJ * 3;
        float acc = 0.f;
        for (long long i = lane; i < TJ; i += 32) {
            const float dx = ps[3 * i] - tw[3 * i], dy = ps[3 * i + 1] - tw[3 * i + 1], dz = ps[3 * i + 2] - tw[3 * i + 2];
            acc += sqrtf(dx * dx + dy * dy + dz * dz);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) bs_dist[s] = acc;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int arg = 0;
        for (int s = 1; s < S; ++s) if (bs_dist[s] < bs_dist[arg]) arg = s;      // first minimum, like torch.min
        bs_arg = arg;
        if (index) index[w] = arg;
    }
    __syncthreads();
    const float* ps = pred + ((long long)w * S + bs_arg) * TJ * 3;
    const long long n = TJ * 3, n_tail = (long long)keep * J * 3;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const float v = ps[i] * scale;
        if (best) best[(long long)w * n + i] = v;
        if (tail && i >= n - n_tail) tail[(long long)w * n_tail + (i - (n - n_tail))] = v;
    }
}

int best_sample_fp32(const float* pred, const float* target, int windows, int samples, int frames, int joints, int keep, float scale,
                     float* best, float* tail, int* index, cudaStream_t st) {
    if (keep < 0 || keep > frames) { set_error("sd_best_sample: keep %d outside [0, %d]", keep, frames); return SD_ERR_INVALID; }
    best_sample_kernel<<<windows, 256, (size_t)samples * sizeof(float), st>>>(pred, target, samples, frames, joints, keep, scale, best, tail, index);
    SD_LAUNCH_OK("best_sample_kernel");
    return SD_OK;
}

// mean over each window's ground truths, in index order (fixed order: bitwise repeatable); a window without any gives 0
__global__ void segment_mean_kernel(const float* __restrict__ a, const float* __restrict__ b, const int* __restrict__ offsets, int windows,
                                    float* __restrict__ out_a, float* __restrict__ out_b) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= windows) return;
    const int lo = offsets[w], hi = offsets[w + 1];
    float sa = 0.f, sb = 0.f;
    for (int i = lo; i < hi; ++i) { sa += a[i]; sb += b[i]; }
    const float inv = hi > lo ? 1.0f / (float)(hi - lo) : 0.0f;
    if (out_a) out_a[w] = sa * inv;
    if (out_b) out_b[w] = sb * inv;
}

// MMADE / MMFDE (src/metrics/multimodal.py:105-135): for every multimodal ground truth g of window i, the ADE / FDE of the
// window's samples against g (minimum over the samples), then the mean over the window's ground truths.
// mm_gt: [G, T, F] all ground truths, window by window; gt_window[g] = its window; gt_offsets[i .. i+1] = its range.
int multimodal_metrics_fp32(const float* pred, const float* mm_gt, const int* gt_window, const int* gt_offsets, int windows, int n_gt,
                            int samples, int frames, int feat, float scale, float* mmade, float* mmfde, float* scratch, cudaStream_t st) {
    if (n_gt > 0) {
        int rc = motion_metrics_launch(pred, mm_gt, gt_window, n_gt, samples, frames, feat, scale, scratch, scratch + n_gt, nullptr, st);
        if (rc) return rc;
    }
    segment_mean_kernel<<<(windows + 127) / 128, 128, 0, st>>>(scratch, scratch + n_gt, gt_offsets, windows, mmade, mmfde);
    SD_LAUNCH_OK("segment_mean_kernel");
    return SD_OK;
}

}  // namespace sd
