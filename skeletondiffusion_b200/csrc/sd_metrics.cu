// Evaluation metrics of the sampling path's output: ADE / FDE / APD per observed window, one launch.
// Replaces ade / fde / apd of the reference (src/metrics/multimodal.py:44-57, :60-73, :15-35) as eval.py calls them on
// skeleton.transform_to_metric_space(pred) (rescalepose.py:29-39: a multiplication by pose_box_size, which every
// one of the three metrics is linear in, so it is applied once to the three results).
//
// One CTA per window.  The window's S predictions [S, T, F] (AMASS: 50 x 120 x 63 floats = 1.5 MB) are read from HBM
// exactly once, a few frames at a time, into shared memory ([S][frames*F] with an odd row stride, so lanes that walk
// different samples hit different banks).  While a chunk is resident
//   - thread s accumulates ||pred[s, f] - target[f]|| of its frames (ADE sum, FDE = last frame),
//   - every thread accumulates the squared distance of its share of the S(S-1)/2 sample pairs (APD).
// Reductions are fixed-order (one owner per sample / pair, tree sum), so results are bitwise repeatable.
// Algorithmic bytes per window: 4*(S+1)*T*F read + 12 written.  The pair loop is shared-memory-bandwidth bound
// (two LDS per FMA), not HBM bound: see DESIGN.md 4.7.
#include "sd_internal.h"

namespace sd {

constexpr int MM_THREADS = 256;
constexpr int MM_PMAX    = 16;     // pairs per thread: S(S-1)/2 <= 4096, i.e. S <= 91

__global__ void __launch_bounds__(MM_THREADS)
motion_metrics_kernel(const float* __restrict__ pred, const float* __restrict__ target, int S, int T, int F, int fc, int stride,
                      float scale, float* __restrict__ ade, float* __restrict__ fde, float* __restrict__ apd) {
    extern __shared__ float sm[];
    float* rows  = sm;                          // [S][stride]
    float* tgt   = rows + (size_t)S * stride;   // [fc*F]
    float* asum  = tgt + fc * F;                // [S] sum over frames of the per-frame distance
    float* flast = asum + S;                    // [S] distance at the last frame
    float* red   = flast + S;                   // [MM_THREADS]
    const int w = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long D = (long long)T * F;
    const float* pw = pred + (long long)w * S * D;
    const float* tw = target + (long long)w * D;
    const int pairs = S * (S - 1) / 2;

    int oi[MM_PMAX], oj[MM_PMAX];
    float acc[MM_PMAX];
#pragma unroll
    for (int k = 0; k < MM_PMAX; ++k) {
        acc[k] = 0.f; oi[k] = oj[k] = 0;
        int p = tid + k * MM_THREADS;
        if (p < pairs) {                        // row-major upper triangle: p -> (i, j), i < j
            int i = 0;
            while (p >= S - 1 - i) { p -= S - 1 - i; ++i; }
            oi[k] = i * stride; oj[k] = (i + 1 + p) * stride;
        }
    }
    for (int s = tid; s < S; s += MM_THREADS) { asum[s] = 0.f; flast[s] = 0.f; }

    for (int f0 = 0; f0 < T; f0 += fc) {
        const int nf = min(fc, T - f0), len = nf * F;
        __syncthreads();                        // previous chunk fully consumed (and asum initialised)
        for (int s = warp; s < S; s += MM_THREADS / 32) {
            const float* src = pw + s * D + (long long)f0 * F;
            float* dst = rows + (size_t)s * stride;
            for (int c = lane; c < len; c += 32) dst[c] = __ldg(src + c);
        }
        for (int c = tid; c < len; c += MM_THREADS) tgt[c] = __ldg(tw + (long long)f0 * F + c);
        __syncthreads();
        for (int s = tid; s < S; s += MM_THREADS) {
            const float* r = rows + (size_t)s * stride;
            float a = asum[s];
            for (int f = 0; f < nf; ++f) {
                float d2 = 0.f;
                for (int c = 0; c < F; ++c) { const float d = r[f * F + c] - tgt[f * F + c]; d2 = fmaf(d, d, d2); }
                const float dist = sqrtf(d2);
                a += dist;
                if (f0 + f == T - 1) flast[s] = dist;
            }
            asum[s] = a;
        }
#pragma unroll
        for (int k = 0; k < MM_PMAX; ++k) {
            if (tid + k * MM_THREADS < pairs) {
                const float* a = rows + oi[k];
                const float* b = rows + oj[k];
                float s0 = 0.f, s1 = 0.f;
                int c = 0;
                for (; c + 1 < len; c += 2) {
                    const float d0 = a[c] - b[c], d1 = a[c + 1] - b[c + 1];
                    s0 = fmaf(d0, d0, s0); s1 = fmaf(d1, d1, s1);
                }
                if (c < len) { const float d0 = a[c] - b[c]; s0 = fmaf(d0, d0, s0); }
                acc[k] += s0 + s1;
            }
        }
    }

    float local = 0.f;
#pragma unroll
    for (int k = 0; k < MM_PMAX; ++k) if (tid + k * MM_THREADS < pairs) local += sqrtf(acc[k]);
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

int motion_metrics_fp32(const float* pred, const float* target, int windows, int samples, int frames, int feat, float scale,
                        float* ade, float* fde, float* apd, cudaStream_t st) {
    const int S = samples, T = frames, F = feat;
    if (S * (S - 1) / 2 > MM_PMAX * MM_THREADS) { set_error("sd_motion_metrics: at most 91 samples per window (got %d)", S); return SD_ERR_UNSUPPORTED; }
    // frames per chunk: as many as keep the sample rows within 40 KB (5 CTAs per SM); one frame at least
    int fc = (40 * 1024 / 4) / (S * F);
    fc = fc < 1 ? 1 : (fc > T ? T : fc);
    const int stride = (fc * F) | 1;
    const size_t smem = ((size_t)S * stride + (size_t)fc * F + 2 * (size_t)S + MM_THREADS) * sizeof(float);
    if (smem > 200 * 1024) { set_error("sd_motion_metrics: one frame of %d samples x %d features does not fit shared memory", S, F); return SD_ERR_UNSUPPORTED; }
    if (smem > 48 * 1024 &&
        check_cuda(cudaFuncSetAttribute(motion_metrics_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "motion_metrics smem")) return SD_ERR_CUDA;
    motion_metrics_kernel<<<windows, MM_THREADS, smem, st>>>(pred, target, S, T, F, fc, stride, scale, ade, fde, apd);
    SD_LAUNCH_OK("motion_metrics_kernel");
    return SD_OK;
}

}  // namespace sd
