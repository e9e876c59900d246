// Internal (non-ABI) declarations shared between translation units of libskeldiff_sm100a.
#pragma once
#include "sd_common.cuh"

struct sd_glin {
    int N, n_types, K, OUT;
    sd::NodeTypes types;
    const float* W;          // [n_types][OUT][K]
    const float* bias_node;  // [N][OUT] or null
    const float* G;          // [N][N] or null (identity)
    const uint16_t* W_bf16;  // [planes][n_types][OUT][K] or null
    int planes;
};

struct sd_gru {
    int N, n_types, IN, H, steps;
    sd::NodeTypes types;
    const float* W_ih;       // [n_types][3H][IN]
    const float* W_hh;       // [n_types][3H][H]
    const float* bias_ih_seq;  // [steps][N][3H]
    const float* bias_hh_seq;  // [steps][N][3H]
    const float* gx_seq;     // [steps][N][N] or null
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

// generic launcher used by every composite: out = epilogue(G^ @ (A @ W^T))
struct GlinCall {
    View a0, a1;
    const float* row_scale;
    Epilogue epi;
    ViewW out;
    float* scratch;   // needed iff G != null
    int B;
};
int glin_forward_fp32(const float* W, int K, int OUT, const NodeTypes& types, int N,
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
int fill_normal(float* out, long long count, uint64_t seed, uint64_t offset, cudaStream_t st);
int q_sample_fp32(const float* x0, const float* eps, const int* t, const float* sqrt_ac, const float* M, float* out,
                  int B, int N, int D, cudaStream_t st);
int mahalanobis_loss_fp32(const float* out, const float* x0, const int* t, const float* S, float* loss,
                          int B, int N, int D, cudaStream_t st);
int gru_gates_fp32(const View& xr, const float* xr_bias, const float* hr, const float* hr_bias,
                   const View& h_in, const ViewW& h_out, int B, int N, int H, cudaStream_t st);

}  // namespace sd
