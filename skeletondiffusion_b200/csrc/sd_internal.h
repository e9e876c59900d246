// Internal (non-ABI) declarations shared between translation units of libskeldiff_sm100a.
#pragma once
#include "sd_common.cuh"
#include <cuda_bf16.h>

struct sd_glin {
    int N, n_types, K, OUT;
    sd::NodeTypes types;
    const float* W;          // [n_types][OUT][K]
    const float* Wt;         // K-major copy [n_types][K][OUT] for the FFMA2 kernel, or null
    const float* bias_node;  // [N][OUT] or null
    const float* G;          // [N][N] or null (identity)
    const uint16_t* W_bf16;  // [planes][n_types][OUT][K] or null
    int planes;
    const uint16_t* W_f16;   // [2][n_types][OUT][K] fp16 planes (hi, lo * 2^11) of the two-plane split, or null
    float* G_host;           // host copy of G (kernel-parameter constant bank of the per-sample mix kernels) or null
};

struct sd_gru {
    int N, n_types, IN, H, steps;
    sd::NodeTypes types;
    const float* W_ih;       // [n_types][3H][IN]
    const float* W_hh;       // [n_types][3H][H]
    const float* bias_ih_seq;  // [steps][N][3H]
    const float* bias_hh_seq;  // [steps][N][3H]
    const float* gx_seq;     // [steps][N][N] or null
    float* gx_host;          // host copy of gx_seq or null
    const uint16_t* W_hh_planes;   // [3][n_types][3H][H] bf16 planes of W_hh (tcgen05 recurrent product) or null
    const uint16_t* W_hh_f16;      // [2][n_types][3H][H] fp16 planes (hi, lo * 2^11) or null
    // gate-interleaved copies for the fused FFMA2 GRU step (identity graph influence only); row c' = 96*blk + 32*g + u
    // holds original row g*H + 32*blk + u, so every 96-column GEMM block carries gates r|z|n of 32 units
    const float* W_ih_perm;  // [n_types][3H][IN]
    const float* W_hh_perm;  // K-major: [n_types][H][3H]
    const float* bias_ih_perm;  // [N][3H]
    const float* bias_hh_perm;  // [N][3H]
    const uint16_t* W_hh_perm_f16;   // [2][n_types][3H][H] fp16 planes of W_hh in the gate-interleaved ROW order (fused tensor-core step) or null
};

struct sd_diffusion {
    int N, D, T;
    const float* c1; const float* c2; const float* s;   // device [T][N][N]
    // host-side classification of every step's tables (diagonal => elementwise kernel)
    unsigned char diagonal[4096];
};

#define SD_MAX_SLOTS 96
struct sd_denoiser {
    int N, dim, cond_dim, out_dim, depth, heads, dim_head, C;
    const sd_glin* slot[SD_MAX_SLOTS];
    const float* time_table; int time_rows;
};

namespace sd {

int sm_count();               // multiprocessors of the current device (cached per device)
// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute: set it once per (kernel, device)
template <typename K>
inline int opt_in_smem(K kern, size_t bytes, unsigned long long& done_mask) {
    int dev = 0;
    if (check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return SD_ERR_CUDA;
    const unsigned long long bit = 1ull << (dev & 63);
    if (done_mask & bit) return SD_OK;
    if (check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes), "cudaFuncSetAttribute")) return SD_ERR_CUDA;
    done_mask |= bit;
    return SD_OK;
}

// generic launcher used by every composite: out = epilogue(G^ @ (A @ W^T))
struct GlinCall {
    View a0, a1;
    const float* row_scale;
    Epilogue epi;
    ViewW out;
    float* scratch;   // needed iff G != null
    int B;
    float* norm_out = nullptr;   // optional request: inv[b * N + n] = 1 / max(||out[b, n, :]||, 1e-12) written by the producing kernel when it
                                 // can (glin_tc3, two-plane residual-ring variant); tc3_take_norm_written() says whether it was
};
bool tc3_take_norm_written();    // true once after a glin_tc3 launch on this thread that honoured GlinCall::norm_out
int glin_forward_fp32(const float* W, const float* Wt, int K, int OUT, const NodeTypes& types, int N,
                      const float* G, const GlinCall& c, cudaStream_t st);
int node_mix_fp32(const float* G, int N, int OUT, const float* y, long long y_sb, const float* row_scale,
                  const Epilogue& epi, const ViewW& out, int B, cudaStream_t st);
int node_attention_fp32(const float* qkv, float* out, int B, int N, int heads, int dh, cudaStream_t st);
int row_inv_norm_fp32(const float* x, float* inv, long long rows, int width, cudaStream_t st);
int reverse_step_fp32(const sd_diffusion* d, const float* x_t, const float* x0, const View* eps,
                      float* x_out, float* mean_out, long long mean_sb, int t, int B, int clip, cudaStream_t st);
int time_table_fp32(const float* times, int rows, int C, float theta, int time_dim, const float* w1, const float* b1,
                    const float* w3, const float* b3, const float* const* head_w, const float* const* head_b,
                    int n_heads, float* table, float* ws, cudaStream_t st);
int tc_split_planes();                 // operand split of the fp32-grade tensor-core path for this thread: 3 (bf16) or 2 (fp16)
void set_tc_split_planes(int planes);
bool fast_epilogue();   // tanh / sigmoid of the fp32-grade tensor-core path through MUFU.EX2 + MUFU.RCP (default) or libdevice (SKELDIFF_ACCURATE_EPILOGUE=1)
int fill_normal(float* out, long long count, uint64_t seed, uint64_t offset, cudaStream_t st);
int motion_metrics_fp32(const float* pred, const float* target, int windows, int samples, int frames, int feat, float scale,
                        float* ade, float* fde, float* apd, cudaStream_t st);
int multimodal_metrics_fp32(const float* pred, const float* mm_gt, const int* gt_window, const int* gt_offsets, int windows, int n_gt,
                            int samples, int frames, int feat, float scale, float* mmade, float* mmfde, float* scratch, cudaStream_t st);
int best_sample_fp32(const float* pred, const float* target, int windows, int samples, int frames, int joints, int keep, float scale,
                     float* best, float* tail, int* index, cudaStream_t st);
int q_sample_fp32(const float* x0, const float* eps, const int* t, const float* sqrt_ac, const float* M, float* out,
                  int B, int N, int D, cudaStream_t st);
int mahalanobis_loss_fp32(const float* out, const float* x0, const int* t, const float* S, float* loss,
                          int B, int N, int D, cudaStream_t st);
// ---- tensor-core (tcgen05) path: bf16 activations
struct TcOperand { const __nv_bfloat16* ptr; long long sb, sn; int width; };
struct TcCall {
    TcOperand a0, a1;            // K segments (a1.ptr null if unused), rows (b, n) at ptr + b*sb + n*sn
    const float* row_scale;      // [B*N] or null
    const float* bias_node;      // [N][OUT] or null
    const float* ss;             // resolved scale/shift row: scale at [o], shift at [OUT+o]; or null
    int act;
    const __nv_bfloat16* res; long long res_sb, res_sn;
    void* out; int out_fp32; long long out_sb, out_sn;
    int B;
    int accurate_tanh;
};
bool glin_f2_supported(const View& a0, const View& a1, int K, int OUT, const float* Wt, const ViewW& out);
int glin_f2_launch(const float* Wt, int K, int OUT, const NodeTypes& types, int N, const GlinCall& c, const ViewW& out, bool fused, cudaStream_t st);
int gru_step_f2(const float* W_hh_perm_t, int H, const NodeTypes& types, int N, const View& xr, const float* bias_x, const float* bias_h,
                const View& h_prev, const ViewW& h_out, int B, cudaStream_t st);
int gru_step_fused(const float* W_hh_perm_t, int H, const NodeTypes& types, int N, const View& xr, const float* bias_x, const float* bias_h,
                   const View& h_prev, const ViewW& h_out, const float* Wfc, const float* bias_fc, const ViewW* y, int F, int B, cudaStream_t st);
int gru_out_fc_fp32(const float* Wfc, const float* bias_node, const NodeTypes& types, int N, int H, int F, const float* h,
                    const ViewW& out, int act, int B, cudaStream_t st);
bool glin_tc3_supported(int K0, int K1, int OUT);
int glin_tc3_launch(const sd_glin* L, const GlinCall& c, const ViewW& out, bool apply_epilogue, cudaStream_t st);
int glin_tc3_gru_step(const sd_glin* L, const View& h_in, const View& xr, const float* bias_x, const float* bias_h, const ViewW& h_out,
                      int B, cudaStream_t st);      // fused recurrent product + gates (identity influence, two-plane split)
bool glin_tc_supported(int K0, int K1, int OUT);
int glin_tc_launch(const sd_glin* L, const TcCall& c, cudaStream_t st);
int cast_concat_bf16(const View& a0, const View& a1, __nv_bfloat16* out, int B, int N, cudaStream_t st);
int row_inv_norm_bf16(const __nv_bfloat16* x, float* inv, long long rows, int width, cudaStream_t st);
int node_attention_bf16(const __nv_bfloat16* qkv, __nv_bfloat16* out, int B, int N, int heads, int dh, cudaStream_t st);
// mix of fp32 raw products with bf16 residual / bf16-or-fp32 output
int node_mix_to_bf16(const float* G, int N, int OUT, const float* y, const Epilogue& epi, const __nv_bfloat16* res,
                     void* out, int out_fp32, int B, cudaStream_t st);
// per-sample kernels for a general graph influence (sd_mix.cu); G_host / gx: [N][N] on the HOST (passed by value to the kernel)
bool sample_mix_supported(int N, int OUT, const float* y, const Epilogue& epi, const ViewW& out);
int sample_mix_fp32(const float* G_host, int N, int OUT, const float* y, const float* row_scale, const Epilogue& epi,
                    const ViewW& out, int B, cudaStream_t st);
bool gru_sample_supported(int N, int H, const float* hr, const View& xr, const View& h_prev, const ViewW& h_out);
int gru_sample_fp32(const float* G_host, int N, int H, const float* hr, const View& xr, const float* bias_x, const float* bias_h,
                    const View& h_prev, const ViewW& h_out, int B, cudaStream_t st);
bool gru_head_supported(int N, int H, int F);
int gru_head_fp32(const float* G_dev, const float* Wfc, const float* bias_node, const NodeTypes& types, int n_types, int N, int H, int F,
                  const View& h, const ViewW& out, int act, int B, cudaStream_t st);
// attention with the to_qkv node mix fused in: qkv holds the RAW per-node products, row (b, m) is scaled by row_scale[b*N + m]
bool node_attention_mix_supported(int N, int heads, int dh, const float* qkv, const float* out);
int node_attention_mix_fp32(const float* G_host, const float* row_scale, const float* qkv, float* out, int B, int N, int heads, int dh, cudaStream_t st);
int add_residual_fp32(float* out, const View& res, int B, int N, int OUT, long long out_sb, long long out_sn, cudaStream_t st);
int gru_gates_fp32(const View& xr, const float* xr_bias, const float* hr, const float* hr_bias,
                   const View& h_in, const ViewW& h_out, int B, int N, int H, cudaStream_t st, bool fast = false);

}  // namespace sd
