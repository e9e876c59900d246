/*
 * skeldiff_b200.h — C ABI of libskeldiff_sm100a.so
 *
 * B200-native (sm_100a) kernels for SkeletonDiffusion's nonisotropic latent-diffusion sampling
 * path.  The reference has no FFI: its boundary is a set of Python classes.  Each
 * entry point below names the reference method(s) it replaces (file:line under /root/reference).
 * The Python classes in skeletondiffusion_b200/ keep the reference's names/signatures/state_dict
 * keys and bind these symbols with ctypes.
 *
 * Conventions
 *   - All pointers named *_dev are CUDA device pointers owned by the caller; the library never
 *     allocates device memory, never synchronises the device, and launches only on `stream`
 *     (a cudaStream_t passed as void*).  All calls are CUDA-graph capturable unless noted.
 *   - Activations are row-major fp32 [B, N, C] ("sample-major": node rows of one sample are
 *     contiguous).  A *view* (ptr, sample_stride, node_stride, rep) addresses row (b, n) at
 *     ptr + (b / rep) * sample_stride + n * node_stride, so repeat_interleave'd conditioning and
 *     time-sliced observations are read in place, never materialised.
 *   - Every function returns 0 on success, non-zero on failure; sd_last_error() returns a
 *     thread-local message.  Handles are re-entrant per stream; one handle set per device.
 */
#ifndef SKELDIFF_B200_H
#define SKELDIFF_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SD_OK 0
#define SD_ERR_INVALID 1
#define SD_ERR_CUDA 2
#define SD_ERR_UNSUPPORTED 3

#define SD_MAX_NODES 64

/* precision of the GEMM data path */
#define SD_PREC_FP32 0      /* FFMA, fp32 operands and accumulation (parity gate <= 1e-4)            */
#define SD_PREC_BF16 1      /* tcgen05 kind::f16, bf16 operands, fp32 accumulation in TMEM            */
#define SD_PREC_BF16X3 2    /* tcgen05, operands split into 3 bf16 planes (fp32-grade products)       */
#define SD_PREC_F16X2 3     /* as BF16X3 with operands split into 2 fp16 planes (hi + lo * 2^-11, 22 significand bits,
                               three products instead of six); operands must stay below 65 504 in magnitude           */

#define SD_ACT_NONE 0
#define SD_ACT_TANH 1
#define SD_ACT_TANH_TANH 2   /* tanh(tanh(v)): encoder fc + z_activation (encoder.py:81, autoencoder.py:54) */

typedef struct sd_glin sd_glin;           /* one StaticGraphLinear                                  */
typedef struct sd_denoiser sd_denoiser;   /* Denoiser forward plan                                  */
typedef struct sd_diffusion sd_diffusion; /* per-step nonisotropic tables                           */
typedef struct sd_gru sd_gru;             /* one StaticGraphGRU cell                                */

/* strided view of a [B, N, width] fp32 activation (see Conventions) */
typedef struct sd_view {
    float*  ptr;
    int64_t sample_stride;
    int64_t node_stride;
    int32_t rep;      /* >=1: row b reads sample b / rep */
    int32_t width;    /* channels addressed through this view */
} sd_view;

const char* sd_last_error(void);
int         sd_version(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches evidence) */
uint64_t    sd_launch_count(void);
/* 1 if the library carries sm_100a code and device `dev` is compute capability 10.x */
int         sd_device_supported(int dev);

/* ------------------------------------------------------------------ StaticGraphLinear ------
 * replaces GraphLinear.forward / gmm   (src/core/network/layers/graph_structural.py:30-43, 7-8)
 *   out[b,n,:] = sum_m G^[n,m] * ( x[b,m,:] @ W[type(m)]^T + bias[type(m)] )
 * Host-side packing (skeletondiffusion_b200/plan.py) passes:
 *   weight_dev    [n_types, out, in] fp32 (n_types == 1 when node_types is None)
 *   bias_node_dev [N, out] = G^ @ bias[type]   (NULL when the layer has no bias)
 *   g_dev         [N, N] row-L1-normalised G^ (NULL when G^ == I: the mix is skipped, exact)
 */
int  sd_glin_create(int num_nodes, const int32_t* node_types_host, int n_types, int in_features,
                    int out_features, const float* weight_dev, const float* bias_node_dev,
                    const float* g_dev, sd_glin** out);
/* optional bf16 weight planes [planes][n_types, out, in] for the tcgen05 paths (planes = 1 or 3) */
int  sd_glin_set_bf16(sd_glin* L, const uint16_t* weight_bf16_dev, int planes);
/* optional fp16 weight planes [2][n_types, out, in] = (fp16(w), fp16((w - fp16(w)) * 2^11)) for SD_PREC_F16X2
 * (StaticGraphLinear weight, graph_structural.py:36); without them SD_PREC_F16X2 runs as SD_PREC_BF16X3 */
int  sd_glin_set_f16x2(sd_glin* L, const uint16_t* weight_f16_dev);
/* optional K-major fp32 copy [n_types, in, out] that enables the FFMA2 kernel (out % 96 == 0, in % 32 == 0) */
int  sd_glin_set_kmajor(sd_glin* L, const float* weight_kmajor_dev);
void sd_glin_destroy(sd_glin* L);

typedef struct sd_glin_args {
    sd_view a0;                 /* first K segment (a0.width columns)                               */
    sd_view a1;                 /* optional second K segment (ptr NULL if unused): concat-free cat  */
    const float* row_scale_dev; /* optional [B*N]: multiplies row (b,m) of x@W^T before the mix    */
    const float* scale_shift_dev; /* optional rows of [2*out]: v = v*(1+scale[o]) + shift[o]        */
    const int32_t* ss_row_dev;  /* optional [B] row index into scale_shift (NULL -> ss_row)         */
    int32_t ss_row;
    int64_t ss_row_stride;      /* elements between scale_shift rows                                */
    int32_t act;                /* SD_ACT_*                                                         */
    sd_view residual;           /* optional (ptr NULL if unused), added after the activation        */
    sd_view out;
    float*  scratch_dev;        /* >= B*N*out floats, required iff the layer has a non-identity G^  */
    int32_t batch;
    int32_t precision;          /* SD_PREC_*                                                        */
} sd_glin_args;

int sd_glin_forward(const sd_glin* L, const sd_glin_args* args, void* stream);

/* Same layer on bf16 activations kept in HBM (the tensor-core path's native format; contiguous
 * [B, N, in] / [B, N, out] tensors).  ss_row_dev: one resolved scale/shift row [2*out] or NULL.
 * scratch_dev (B*N*out floats) is required iff the layer has a non-identity G^. */
int sd_glin_forward_bf16(const sd_glin* L, const uint16_t* a_dev, const float* row_scale_dev,
                         const float* ss_row_dev, int act, const uint16_t* residual_dev, void* out_dev,
                         int out_is_fp32, float* scratch_dev, int batch, void* stream);

/* ------------------------------------------------------------------ node attention ---------
 * replaces Attention.forward's softmax(q k^T) v over nodes (layers/attention.py:125-135)
 *   qkv_dev [B, N, 3*heads*dim_head] (q | k | v, head-major inside each), out [B, N, heads*dim_head]
 */
int sd_node_attention(const float* qkv_dev, float* out_dev, int batch, int num_nodes, int heads,
                      int dim_head, void* stream);

/* RMSNorm row factor (layers/attention.py:36): inv_norm[r] = 1 / max(||x[r,:]||_2, 1e-12) */
int sd_row_inv_norm(const float* x_dev, float* inv_norm_dev, int64_t rows, int width, void* stream);

/* ------------------------------------------------------------------ time conditioning ------
 * replaces SinusoidalPosEmb -> Linear -> GELU -> Linear (nn/generator.py:47-55) and the
 * ResnetBlock heads Tanh -> Linear (layers/attention.py:81-84) for `n_rows` distinct time values.
 *   table_dev [n_rows, n_heads, 2*C];  workspace >= n_rows * (C + 2*time_dim) floats
 */
int sd_time_table(const float* times_dev, int n_rows, int C, float theta, int time_dim,
                  const float* w1_dev, const float* b1_dev, const float* w3_dev, const float* b3_dev,
                  const float* const* head_w_dev_host, const float* const* head_b_dev_host,
                  int n_heads, float* table_dev, float* workspace_dev, void* stream);

/* ------------------------------------------------------------------ Denoiser ---------------
 * replaces Denoiser.forward (nn/generator.py:86-107) incl. ResnetBlock / Block / PreNorm /
 * Attention / Residual (layers/attention.py:11-136).
 * slots: 0 init_lin; for pair i in [0, 2*depth): 1+4i block1.proj, 2+4i block2.proj, 3+4i to_qkv,
 *        4+4i to_out (pair 2*depth-1 has no attention); then final_res_block.block1, .block2,
 *        .res_linear, final_glin at 1+8*depth + {0,1,2,3}.
 * to_qkv weights must be pre-multiplied by norm.g * sqrt(C) (RMSNorm gain folded at pack time).
 */
int  sd_denoiser_create(int num_nodes, int dim, int cond_dim, int out_dim, int depth, int heads,
                        int dim_head, sd_denoiser** out);
int  sd_denoiser_set_layer(sd_denoiser* d, int slot, const sd_glin* layer);
/* scale/shift table from sd_time_table: [n_rows, 2*depth+1, 2*C] */
int  sd_denoiser_set_time_table(sd_denoiser* d, const float* table_dev, int n_rows);
void sd_denoiser_destroy(sd_denoiser* d);
size_t sd_denoiser_workspace_bytes(const sd_denoiser* d, int batch, int precision);
/* t_rows_dev: optional [B] per-sample row of the time table; NULL -> all samples use t_row */
int  sd_denoiser_forward(const sd_denoiser* d, const sd_view* x, const sd_view* x_cond,
                         const int32_t* t_rows_dev, int t_row, float* out_dev, int batch,
                         void* workspace_dev, int precision, void* stream);

/* ------------------------------------------------------------------ reverse diffusion ------
 * replaces p_sample / p_mean_variance / q_posterior / p_combine_mean_var_noise
 * (diffusion/base.py:314-341, diffusion/nonisotropic.py:196-210):
 *   x_{t-1} = C1[t] clamp(x0) + C2[t] x_t + S[t] eps,  S[t] = U diag(exp(0.5 logLambda_post[t]))
 * tables [T, N, N] fp32 on device (posterior_mean_coef1_x0, posterior_mean_coef2_xt, S).
 */
int  sd_diffusion_create(int num_nodes, int latent_dim, int timesteps, const float* c1_dev,
                         const float* c2_dev, const float* s_dev, const float* c1_host,
                         const float* c2_host, const float* s_host, sd_diffusion** out);
void sd_diffusion_destroy(sd_diffusion* d);
/* eps_dev may be NULL (t == 0: no noise).  mean_out_dev optional.  eps addressed as a view so
 * sampling_noise[:, T-1-t] is read in place. */
int  sd_reverse_step(const sd_diffusion* d, const float* x_t_dev, const float* x0_dev,
                     const sd_view* eps, float* x_out_dev, float* mean_out_dev, int t, int batch,
                     int clip_denoised, void* stream);
/* forward process q_sample (nonisotropic.py:152-159): x = sqrt_ac[t_b] x0 + M[t_b] eps, M = U sqrt(Lambda_bar_t) */
int  sd_q_sample(const float* x0_dev, const float* eps_dev, const int32_t* t_dev,
                 const float* sqrt_ac_dev, const float* m_dev, float* out_dev, int batch,
                 int num_nodes, int latent_dim, void* stream);
/* Mahalanobis l1 loss (nonisotropic.py:180-190 + base.py:298): loss[b] = mean |S[t_b] (out - x0)| */
int  sd_mahalanobis_loss(const float* out_dev, const float* x0_dev, const int32_t* t_dev,
                         const float* s_dev, float* loss_dev, int batch, int num_nodes,
                         int latent_dim, void* stream);

/* whole p_sample_loop (diffusion/base.py:343-390): x_dev holds x_T on entry and x_0 on exit.
 * sampling_noise_dev [B, T-1, N, D] (required; throughput mode fills it with sd_fill_normal).
 * means_out_dev optional [B, T-1, N, D] (return_sampling_noise=True). */
size_t sd_sample_workspace_bytes(const sd_diffusion* df, const sd_denoiser* dn, int batch, int precision);
int  sd_sample_loop(const sd_diffusion* df, const sd_denoiser* dn, float* x_dev, const sd_view* x_cond,
                    const float* sampling_noise_dev, float* means_out_dev, int batch,
                    int clip_denoised, void* workspace_dev, int precision, void* stream);

/* counter-based N(0,1) generator (Philox4x32-10 + Box-Muller); replaces torch.randn at base.py:156-158 */
int  sd_fill_normal(float* out_dev, int64_t count, uint64_t seed, uint64_t offset, void* stream);

/* ------------------------------------------------------------------ graph GRU / autoencoder -
 * replaces StaticGraphGRUCell_.forward (layers/recurrent.py:321-366) and the loops of
 * Encoder.forward (nn/encoder.py:77-82) / Decoder.forward (nn/decoder.py:85-104).
 * gx_seq_dev: [steps, N, N] the data-independent sequence gx_i (NULL when every gx_i == I);
 * bias_*_seq_dev: [steps, N, 3H] = gx_i @ bias[type].
 */
int  sd_gru_create(int num_nodes, const int32_t* node_types_host, int n_types, int input_size,
                   int hidden_size, const float* w_ih_dev, const float* w_hh_dev,
                   const float* bias_ih_seq_dev, const float* bias_hh_seq_dev,
                   const float* gx_seq_dev, int steps, sd_gru** out);
/* Three bf16 planes of W_hh ([3][n_types][3H][H], w = p0 + p1 + p2 exactly) for the tcgen05 recurrent product
 * h W_hh^T (recurrent.py:339) of the bf16x3 / bf16 precisions.  Optional: without it the FFMA kernels are used. */
int  sd_gru_set_bf16x3(sd_gru* g, const uint16_t* w_hh_planes_dev);
/* Two fp16 planes of W_hh ([2][n_types][3H][H], as sd_glin_set_f16x2) for the recurrent product under SD_PREC_F16X2 (recurrent.py:339). */
int  sd_gru_set_f16x2(sd_gru* g, const uint16_t* w_hh_f16_dev);
/* Optional, only when every gx_i == I: gate-interleaved copies that enable the fused step kernels
 * (h @ W_hh^T + gates in one launch, recurrent.py:339-358).  Row c' = 96*blk + 32*g + u of the permuted tensors holds original row
 * g*H + 32*blk + u (g = gate r/z/n); w_hh_perm_dev is additionally K-major: [n_types, H, 3H].  bias_*_perm_dev: [N, 3H] = bias[type(n)] in the same order. */
int  sd_gru_set_fused(sd_gru* g, const float* w_ih_perm_dev, const float* w_hh_perm_dev,
                      const float* bias_ih_perm_dev, const float* bias_hh_perm_dev);
/* After sd_gru_set_fused: two fp16 planes ([2][n_types][3H][H], as sd_gru_set_f16x2) of W_hh with its ROWS in the gate-interleaved
 * order.  Enables the fused tensor-core step under SD_PREC_F16X2: recurrent product on tcgen05 with the gates (recurrent.py:351-358)
 * in its epilogue; the [B, N, 3H] product never reaches HBM. */
int  sd_gru_set_fused_f16x2(sd_gru* g, const uint16_t* w_hh_perm_f16_dev);
void sd_gru_destroy(sd_gru* g);
size_t sd_encode_workspace_bytes(int windows, int obs_len, int num_nodes, int hidden, int layers);
/* obs_dev [W, T, N, F] -> z_dev [W, N, latent] = final_act(fc(h_T)); final_act = SD_ACT_TANH_TANH for
 * get_past_embedding with encoder_act = z_activation = tanh (encoder.py:81, autoencoder.py:51-55) */
int  sd_encode(const sd_glin* initial_hidden, sd_gru* const* layers_host, int n_layers, const sd_glin* fc,
               const float* obs_dev, int windows, int obs_len, int feat, float* z_dev, int final_act,
               void* workspace_dev, int precision, void* stream);
size_t sd_decode_workspace_bytes(int batch, int num_nodes, int hidden);
/* x_last2 view: rows of obs[:, -2] (ptr) and obs[:, -1] (ptr + frame_stride); latent [B,N,L];
 * out_dev [B, ph, N, F]  (autoencoder.py:66-73, decoder.py:61-104) */
int  sd_decode(const sd_glin* initial_hidden, const sd_gru* cell, const sd_glin* fc,
               const sd_view* x_prev, const sd_view* x_last, const float* latent_dev, int batch,
               int ph, int feat, float* out_dev, void* workspace_dev, int precision, void* stream);

/* ------------------------------------------------------------------ evaluation metrics ------
 * replaces ade / fde / apd (src/metrics/multimodal.py:44-57, :60-73, :15-35) as eval.py applies them to
 * skeleton.transform_to_metric_space(pred) (rescalepose.py:29-39; scale = pose_box_size, 1 for unit-box poses).
 * pred_dev [W, S, T, feat], target_dev [W, T, feat] (feat = joints*3, contiguous); outputs [W] each, any may be NULL.
 * ade = min_s mean_t ||pred - target||, fde = min_s ||.|| at the last frame, apd = mean pairwise distance of the S
 * flattened samples (0 when S == 1).  S <= 91.  One launch, predictions read once. */
int  sd_motion_metrics(const float* pred_dev, const float* target_dev, int windows, int samples, int frames, int feat,
                       float scale, float* ade_dev, float* fde_dev, float* apd_dev, void* stream);

/* MMADE / MMFDE per observed window -- replaces mmade / mmfde (src/metrics/multimodal.py:105-120, :122-135) on
 * transform_to_metric_space(pred): for every multimodal ground truth g of window i the minimum over the window's samples of
 * the mean (MMADE) / last-frame (MMFDE) L2 distance to g, then the mean over the window's ground truths (0 for a window
 * without any).  mm_gt_dev [n_gt, frames, feat]: all ground truths, window by window; gt_window_dev [n_gt]: window of each;
 * gt_offsets_dev [windows + 1]: range of each window in mm_gt_dev; scratch_dev: 2 * n_gt floats.  pred_dev as for sd_motion_metrics. */
int  sd_multimodal_metrics(const float* pred_dev, const float* mm_gt_dev, const int32_t* gt_window_dev, const int32_t* gt_offsets_dev,
                           int windows, int n_gt, int samples, int frames, int feat, float scale, float* mmade_dev, float* mmfde_dev,
                           float* scratch_dev, void* stream);

/* Best-sample selection of the long-term evaluation -- replaces get_best_sample_idx (src/metrics/utils.py:22-30) in
 * long_term_prediction_best_every50 (src/eval_utils.py:44-67): per window the sample with the smallest mean per-joint L2
 * distance to the target segment.  pred_dev [windows, samples, frames, joints, 3], target_dev [windows, frames, joints, 3];
 * outputs (each optional): best_dev [windows, frames, joints, 3] = scale * chosen sample, tail_dev [windows, keep_frames, joints, 3]
 * = its last keep_frames frames (the next observation), index_dev [windows].  No host round trip. */
int  sd_best_sample(const float* pred_dev, const float* target_dev, int windows, int samples, int frames, int joints, int keep_frames,
                    float scale, float* best_dev, float* tail_dev, int32_t* index_dev, void* stream);

/* ---------------------------------------------------------------------------------------------------------------------
 * Training path: backward kernels behind forward() / p_losses with autograd (src/core/diffusion/base.py:262-307) and
 * TrainerDiffusion.loss (src/core/trainer.py:224-234).  fp32; every reduction in a fixed order (repeatable gradients).
 * ------------------------------------------------------------------------------------------------------------------- */
/* out[b, m, :] = sum_n G[n, m] in[b, n, :]: the gradient of the node mix of GraphLinear.forward (graph_structural.py:41) */
int  sd_node_mix_transposed(const float* g_dev, const float* in_dev, float* out_dev, int batch, int num_nodes, int width, void* stream);
/* Parameter gradients of one StaticGraphLinear (graph_structural.py:30-43).  dym_dev = G^T dOut [B, N, out]; x_dev [B, N, in];
 * y_raw_dev = x W[type]^T [B, N, out] (without bias), bias_types_dev [n_types, out] or null.  Each output is optional:
 * dweight_dev [n_types, out, in], dbias_dev [n_types, out], dg_dev [N, N] (gradient with respect to the NORMALISED influence
 * matrix); accumulate != 0 adds to the outputs.  scratch_dev: sd_glin_backward_scratch_bytes. */
size_t sd_glin_backward_scratch_bytes(const sd_glin* L, int batch);
int  sd_glin_backward_params(const sd_glin* L, const float* x_dev, const float* dym_dev, const float* dout_dev, const float* y_raw_dev,
                             const float* bias_types_dev, float* dweight_dev, float* dbias_dev, float* dg_dev, float* scratch_dev,
                             int batch, int accumulate, void* stream);
/* Block (attention.py:66-76) without the projection: h = tanh(y (scale[t_b] + 1) + shift[t_b]); ss_table_dev [rows, 2 width]
 * (scale | shift) indexed by t_dev [B], or null (h = tanh(y)).  Backward: dy and the per-sample sums dss_rows_dev [B, 2 width]. */
int  sd_ss_tanh_forward(const float* y_dev, const float* ss_table_dev, const int32_t* t_dev, float* h_dev, int batch, int num_nodes, int width, void* stream);
int  sd_ss_tanh_backward(const float* dh_dev, const float* h_dev, const float* y_dev, const float* ss_table_dev, const int32_t* t_dev,
                         float* dy_dev, float* dss_rows_dev, int batch, int num_nodes, int width, void* stream);
/* RMSNorm (attention.py:30-36): y = x / max(|x|, 1e-12) * g * sqrt(width); inv_dev [rows] keeps 1 / |x| for the backward, which
 * writes dx and per-block partial sums of dg (dg_part_dev [sd_rmsnorm_backward_blocks(rows), width]; the caller adds the blocks). */
int  sd_rmsnorm_forward(const float* x_dev, const float* g_dev, float* y_dev, float* inv_dev, int64_t rows, int width, void* stream);
int  sd_rmsnorm_backward_blocks(int64_t rows);
int  sd_rmsnorm_backward(const float* dy_dev, const float* x_dev, const float* inv_dev, const float* g_dev, float* dx_dev, float* dg_part_dev,
                         int64_t rows, int width, void* stream);
/* Backward of sd_node_attention (attention.py:121-136): dqkv [B, N, 3 heads dim_head] from qkv and dout [B, N, heads dim_head]. */
int  sd_node_attention_backward(const float* qkv_dev, const float* dout_dev, float* dqkv_dev, int batch, int num_nodes, int heads, int dim_head, void* stream);
/* Backward of sd_mahalanobis_loss (nonisotropic.py:176-190, 'l1'): dout = S[t]^T sign(S[t] (out - x0)) grad_loss / (N D). */
int  sd_mahalanobis_loss_backward(const float* out_dev, const float* x0_dev, const int32_t* t_dev, const float* s_dev, const float* grad_loss_dev,
                                  float* dout_dev, int batch, int num_nodes, int latent_dim, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SKELDIFF_B200_H */
