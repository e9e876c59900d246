// C ABI of libskeldiff_sm100a.so: handle management and the composite operators
// (Denoiser forward, p_sample_loop, encode, decode).  See include/skeldiff_b200.h.
#include "sd_internal.h"
#include <stdlib.h>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <atomic>

namespace sd {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return 0;
    set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    return 1;
}

static std::atomic<unsigned long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    int& c = cached[dev & 63];
    if (c == 0) { int v = 148; cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev); c = v; }
    return c;
}

static inline size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

struct Arena {   // carves a caller-owned workspace
    char* base; size_t off;
    explicit Arena(void* p) : base(static_cast<char*>(p)), off(0) {}
    float* floats(size_t n) { float* r = reinterpret_cast<float*>(base + off); off += align_up(n * sizeof(float)); return r; }
};

static Epilogue no_epilogue(int OUT) {
    Epilogue e; e.bias_node = nullptr; e.ss = nullptr; e.ss_row_idx = nullptr; e.ss_row = 0; e.ss_stride = 0;
    e.act = SD_ACT_NONE; e.residual.ptr = nullptr; e.residual.sb = e.residual.sn = 0; e.residual.rep = 1; e.residual.width = 0;
    e.OUT = OUT;
    return e;
}
static View null_view() { View v; v.ptr = nullptr; v.sb = v.sn = 0; v.rep = 1; v.width = 0; return v; }

// One graph-linear layer with fp32 views through the tcgen05 path (ABI-level entry, used by
// sd_glin_forward): operands are cast to bf16 in scratch, the product runs on the tensor cores.
// scratch layout: [bf16 A copy: B*N*K][fp32 Y: B*N*OUT (only for a non-identity G^)]
static int glin_forward_tc(const sd_glin* L, const GlinCall& c, int precision, cudaStream_t st) {
    if (precision != SD_PREC_BF16) { set_error("precision %d is not available for this call", precision); return SD_ERR_UNSUPPORTED; }
    if (c.epi.ss_row_idx) { set_error("tcgen05 path: per-sample time rows are only supported on the fp32 path"); return SD_ERR_UNSUPPORTED; }
    if (!c.scratch) { set_error("tcgen05 path: scratch required (B*N*(2*in + 4*out) bytes)"); return SD_ERR_INVALID; }
    if (c.out.sn % 4 || c.out.sb % 4 || c.out.rep != 1) { set_error("tcgen05 path: output view must be 16-byte aligned rows"); return SD_ERR_UNSUPPORTED; }
    const int N = L->N, K = L->K, OUT = L->OUT;
    __nv_bfloat16* a16 = reinterpret_cast<__nv_bfloat16*>(c.scratch);
    float* y = reinterpret_cast<float*>(reinterpret_cast<char*>(c.scratch) + (((size_t)c.B * N * K * 2 + 255) / 256) * 256);
    int rc = cast_concat_bf16(c.a0, c.a1, a16, c.B, N, st);
    if (rc) return rc;
    TcCall t;
    t.a0.ptr = a16; t.a0.sb = (long long)N * K; t.a0.sn = K; t.a0.width = K;
    t.a1.ptr = nullptr; t.a1.sb = t.a1.sn = 0; t.a1.width = 0;
    t.row_scale = c.row_scale; t.res = nullptr; t.res_sb = t.res_sn = 0; t.B = c.B; t.accurate_tanh = 1;
    if (L->G == nullptr) {
        t.bias_node = c.epi.bias_node;
        t.ss = c.epi.ss ? c.epi.ss + (long long)c.epi.ss_row * c.epi.ss_stride : nullptr;
        t.act = c.epi.act;
        t.out = c.out.ptr; t.out_fp32 = 1; t.out_sb = c.out.sb; t.out_sn = c.out.sn;
        rc = glin_tc_launch(L, t, st);
        if (rc) return rc;
        if (c.epi.residual.ptr) return add_residual_fp32(c.out.ptr, c.epi.residual, c.B, N, OUT, c.out.sb, c.out.sn, st);
        return SD_OK;
    }
    t.bias_node = nullptr; t.ss = nullptr; t.act = SD_ACT_NONE;
    t.out = y; t.out_fp32 = 1; t.out_sb = (long long)N * OUT; t.out_sn = OUT;
    rc = glin_tc_launch(L, t, st);
    if (rc) return rc;
    Epilogue e = c.epi; e.OUT = OUT;
    return node_mix_fp32(L->G, N, OUT, y, (long long)N * OUT, nullptr, e, c.out, c.B, st);
}

// Denoiser forward with bf16 activations end to end (tcgen05 graph-linears, bf16 attention I/O).
// Only the latent input (fp32, cast once) and the x0 output (fp32, feeds the reverse-step kernel) are fp32.
static int denoiser_forward_bf16(const sd_denoiser* d, const sd_view* x, const sd_view* x_cond, const int32_t* t_rows_dev,
                                 int t_row, float* out_dev, int B, void* workspace_dev, cudaStream_t st) {
    if (t_rows_dev) { set_error("bf16 path: per-sample time rows are only supported on the fp32 path"); return SD_ERR_UNSUPPORTED; }
    const int N = d->N, C = d->C, hd = d->heads * d->dim_head, Kin = d->dim + d->cond_dim;
    const size_t rows = (size_t)B * N;
    char* base = static_cast<char*>(workspace_dev);
    size_t off = 0;
    auto take = [&](size_t bytes) { char* r = base + off; off += (bytes + 255) / 256 * 256; return r; };
    auto h16 = [&](size_t n) { return reinterpret_cast<__nv_bfloat16*>(take(n * 2)); };
    __nv_bfloat16 *xin = h16(rows * Kin), *r = h16(rows * C), *xb = h16(rows * C), *h = h16(rows * C), *res = h16(rows * C);
    __nv_bfloat16 *qkv = h16(rows * 3 * hd), *att = h16(rows * hd);
    float* inv = reinterpret_cast<float*>(take(rows * 4));
    float* y = reinterpret_cast<float*>(take(rows * (size_t)(3 * hd > C ? 3 * hd : C) * 4));
    const int n_pairs = 2 * d->depth;
    const long long ss_stride = (long long)(n_pairs + 1) * 2 * C;
    auto tc = [&](const sd_glin* L, const __nv_bfloat16* a0, int k0, const __nv_bfloat16* a1, int k1, const float* row_scale,
                  int ss_head, int act, const __nv_bfloat16* resid, void* out, int out_fp32) -> int {
        if (!L) { set_error("graph-linear layer not set"); return SD_ERR_INVALID; }
        TcCall t;
        t.a0.ptr = a0; t.a0.sb = (long long)N * k0; t.a0.sn = k0; t.a0.width = k0;
        t.a1.ptr = a1; t.a1.sb = (long long)N * k1; t.a1.sn = k1; t.a1.width = k1;
        t.row_scale = row_scale; t.B = B; t.accurate_tanh = 0;
        const float* ss = ss_head >= 0 ? d->time_table + (long long)t_row * ss_stride + (long long)ss_head * 2 * C : nullptr;
        if (L->G == nullptr) {
            t.bias_node = L->bias_node; t.ss = ss; t.act = act;
            t.res = resid; t.res_sb = (long long)N * L->OUT; t.res_sn = L->OUT;
            t.out = out; t.out_fp32 = out_fp32; t.out_sb = (long long)N * L->OUT; t.out_sn = L->OUT;
            return glin_tc_launch(L, t, st);
        }
        t.bias_node = nullptr; t.ss = nullptr; t.act = SD_ACT_NONE; t.res = nullptr; t.res_sb = t.res_sn = 0;
        t.out = y; t.out_fp32 = 1; t.out_sb = (long long)N * L->OUT; t.out_sn = L->OUT;
        int rc = glin_tc_launch(L, t, st);
        if (rc) return rc;
        Epilogue e = no_epilogue(L->OUT);
        e.bias_node = L->bias_node;
        if (ss) { e.ss = ss; e.ss_row = 0; e.ss_stride = 0; }
        e.act = act;
        return node_mix_to_bf16(L->G, N, L->OUT, y, e, resid, out, out_fp32, B, st);
    };
    View vx = make_view(*x);
    int rc = d->cond_dim > 0 ? cast_concat_bf16(make_view(*x_cond), vx, xin, B, N, st) : cast_concat_bf16(vx, null_view(), xin, B, N, st);
    if (rc) return rc;
    rc = tc(d->slot[0], xin, Kin, nullptr, 0, nullptr, -1, SD_ACT_NONE, nullptr, r, 0);
    if (rc) return rc;
    const __nv_bfloat16* cur = r;
    for (int i = 0; i < n_pairs; ++i) {
        const int s0 = 1 + 4 * i;
        rc = tc(d->slot[s0], cur, C, nullptr, 0, nullptr, i, SD_ACT_TANH, nullptr, h, 0);
        if (rc) return rc;
        rc = tc(d->slot[s0 + 1], h, C, nullptr, 0, nullptr, -1, SD_ACT_TANH, cur, xb, 0);
        if (rc) return rc;
        cur = xb;
        if (i != n_pairs - 1) {
            rc = row_inv_norm_bf16(xb, inv, (long long)rows, C, st);
            if (rc) return rc;
            rc = tc(d->slot[s0 + 2], xb, C, nullptr, 0, inv, -1, SD_ACT_NONE, nullptr, qkv, 0);
            if (rc) return rc;
            rc = node_attention_bf16(qkv, att, B, N, d->heads, d->dim_head, st);
            if (rc) return rc;
            rc = tc(d->slot[s0 + 3], att, hd, nullptr, 0, nullptr, -1, SD_ACT_NONE, xb, xb, 0);
            if (rc) return rc;
        }
    }
    const int sf = 1 + 8 * d->depth;
    rc = tc(d->slot[sf + 2], cur, C, r, C, nullptr, -1, SD_ACT_NONE, nullptr, res, 0);
    if (rc) return rc;
    rc = tc(d->slot[sf], cur, C, r, C, nullptr, n_pairs, SD_ACT_TANH, nullptr, h, 0);
    if (rc) return rc;
    rc = tc(d->slot[sf + 1], h, C, nullptr, 0, nullptr, -1, SD_ACT_TANH, res, xb, 0);
    if (rc) return rc;
    return tc(d->slot[sf + 3], xb, C, nullptr, 0, nullptr, -1, SD_ACT_NONE, nullptr, out_dev, 1);
}

static bool tc3_views_ok(const GlinCall& c) {
    auto ok = [](const void* p, long long sb, long long sn) { return p == nullptr || ((reinterpret_cast<uintptr_t>(p) & 15u) == 0 && sb % 4 == 0 && sn % 4 == 0); };
    return ok(c.a0.ptr, c.a0.sb, c.a0.sn) && ok(c.a1.ptr, c.a1.sb, c.a1.sn) && ok(c.out.ptr, c.out.sb, c.out.sn) &&
           ok(c.epi.residual.ptr, c.epi.residual.sb, c.epi.residual.sn) && c.out.rep == 1;
}

// out = epilogue(G^ @ (rs * Y)) for a layer with a dense graph influence: the per-sample bulk-copy kernel (sd_mix.cu) where the
// shape is instantiated, the generic kernel otherwise
static int mix_epilogue(const sd_glin* L, const float* y, const float* row_scale, const Epilogue& epi, const ViewW& out, int B, cudaStream_t st) {
    Epilogue e = epi; e.OUT = L->OUT;
    if (L->G_host && sample_mix_supported(L->N, L->OUT, y, e, out)) return sample_mix_fp32(L->G_host, L->N, L->OUT, y, row_scale, e, out, B, st);
    return node_mix_fp32(L->G, L->N, L->OUT, y, (long long)L->N * L->OUT, row_scale, e, out, B, st);
}

static int run_glin(const sd_glin* L, GlinCall c, int precision, cudaStream_t st) {
    if (!L) { set_error("graph-linear layer not set"); return SD_ERR_INVALID; }
    if (c.epi.bias_node == nullptr) c.epi.bias_node = L->bias_node;
    c.epi.OUT = L->OUT;
    if (precision == SD_PREC_BF16X3) {
        // fp32-grade products on the tensor cores where the shape allows it; otherwise the (exact) FFMA kernel
        const int K1 = c.a1.ptr ? c.a1.width : 0;
        if (L->planes == 3 && !c.epi.ss_row_idx && glin_tc3_supported(c.a0.width, K1, L->OUT) && tc3_views_ok(c)) {
            if (L->G == nullptr) return glin_tc3_launch(L, c, c.out, true, st);
            if (!c.scratch) { set_error("glin: scratch required for non-identity G"); return SD_ERR_INVALID; }
            // raw products on the tensor cores (no row scale: it belongs to the INPUT node of the mix), then the per-sample mix
            GlinCall raw = c; raw.row_scale = nullptr;
            int rc = glin_tc3_launch(L, raw, contiguous_view_w(c.scratch, L->N, L->OUT), false, st);
            if (rc) return rc;
            return mix_epilogue(L, c.scratch, c.row_scale, c.epi, c.out, c.B, st);
        }
        return run_glin(L, c, SD_PREC_FP32, st);
    }
    if (precision != SD_PREC_FP32) return glin_forward_tc(L, c, precision, st);
    if (L->G != nullptr && L->G_host && c.scratch && sample_mix_supported(L->N, L->OUT, c.scratch, c.epi, c.out)) {
        // exact-fp32 products (FFMA kernels) written raw, then the per-sample mix
        GlinCall raw = c; raw.row_scale = nullptr; raw.epi = no_epilogue(L->OUT); raw.out = contiguous_view_w(c.scratch, L->N, L->OUT);
        int rc = glin_forward_fp32(L->W, L->Wt, L->K, L->OUT, L->types, L->N, nullptr, raw, st);
        if (rc) return rc;
        return mix_epilogue(L, c.scratch, c.row_scale, c.epi, c.out, c.B, st);
    }
    return glin_forward_fp32(L->W, L->Wt, L->K, L->OUT, L->types, L->N, L->G, c, st);
}

// SD_PREC_F16X2 is SD_PREC_BF16X3 with the two-plane fp16 operand split: the entry points rewrite the precision and select the
// split for the calls made on this thread while they run (nested entry points see SD_PREC_BF16X3 and leave the selection alone).
struct SplitScope {
    int prev; bool active;
    explicit SplitScope(int& precision) : prev(tc_split_planes()), active(precision == SD_PREC_F16X2) {
        if (active) { set_tc_split_planes(2); precision = SD_PREC_BF16X3; }
    }
    ~SplitScope() { if (active) set_tc_split_planes(prev); }
};

}  // namespace sd

using namespace sd;

extern "C" {

const char* sd_last_error(void) { return g_err; }
int sd_version(void) { return 100; }
uint64_t sd_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int sd_device_supported(int dev) {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return 0;
    return p.major == 10 ? 1 : 0;
}

// ------------------------------------------------------------------------------- glin handle
int sd_glin_create(int num_nodes, const int32_t* node_types_host, int n_types, int in_features, int out_features,
                   const float* weight_dev, const float* bias_node_dev, const float* g_dev, sd_glin** out) {
    if (!out || num_nodes <= 0 || num_nodes > SD_MAX_NODES || in_features <= 0 || out_features <= 0 || !weight_dev || n_types <= 0) {
        set_error("sd_glin_create: invalid arguments (N=%d in=%d out=%d types=%d)", num_nodes, in_features, out_features, n_types);
        return SD_ERR_INVALID;
    }
    sd_glin* L = new (std::nothrow) sd_glin();
    if (!L) { set_error("out of host memory"); return SD_ERR_INVALID; }
    L->N = num_nodes; L->n_types = n_types; L->K = in_features; L->OUT = out_features;
    for (int n = 0; n < SD_MAX_NODES; ++n) L->types.t[n] = 0;
    for (int n = 0; n < num_nodes; ++n) {
        const int t = node_types_host ? node_types_host[n] : 0;
        if (t < 0 || t >= n_types) { delete L; set_error("sd_glin_create: node type %d outside [0,%d)", t, n_types); return SD_ERR_INVALID; }
        L->types.t[n] = (unsigned char)t;
    }
    L->W = weight_dev; L->Wt = nullptr; L->bias_node = bias_node_dev; L->G = g_dev; L->W_bf16 = nullptr; L->planes = 0; L->W_f16 = nullptr;
    L->G_host = nullptr;
    if (g_dev) {   // host copy: the per-sample mix kernels take G^ by value (constant bank)
        L->G_host = new (std::nothrow) float[(size_t)num_nodes * num_nodes];
        if (!L->G_host || cudaMemcpy(L->G_host, g_dev, sizeof(float) * num_nodes * num_nodes, cudaMemcpyDeviceToHost) != cudaSuccess) {
            delete[] L->G_host; delete L; set_error("sd_glin_create: cannot read the graph-influence matrix"); return SD_ERR_CUDA;
        }
    }
    *out = L;
    return SD_OK;
}

int sd_glin_set_bf16(sd_glin* L, const uint16_t* weight_bf16_dev, int planes) {
    if (!L || (planes != 1 && planes != 3)) { set_error("sd_glin_set_bf16: invalid arguments"); return SD_ERR_INVALID; }
    L->W_bf16 = weight_bf16_dev; L->planes = planes;
    return SD_OK;
}

int sd_glin_set_f16x2(sd_glin* L, const uint16_t* weight_f16_dev) {
    if (!L || !weight_f16_dev) { set_error("sd_glin_set_f16x2: null argument"); return SD_ERR_INVALID; }
    L->W_f16 = weight_f16_dev;
    return SD_OK;
}

int sd_glin_set_kmajor(sd_glin* L, const float* weight_kmajor_dev) {
    if (!L) { set_error("sd_glin_set_kmajor: null handle"); return SD_ERR_INVALID; }
    L->Wt = weight_kmajor_dev;
    return SD_OK;
}

void sd_glin_destroy(sd_glin* L) { if (L) delete[] L->G_host; delete L; }

int sd_glin_forward(const sd_glin* L, const sd_glin_args* a, void* stream) {
    if (!L || !a) { set_error("sd_glin_forward: null argument"); return SD_ERR_INVALID; }
    GlinCall c;
    c.a0 = make_view(a->a0);
    c.a1 = a->a1.ptr ? make_view(a->a1) : null_view();
    c.row_scale = a->row_scale_dev;
    c.epi = no_epilogue(L->OUT);
    c.epi.ss = a->scale_shift_dev; c.epi.ss_row_idx = a->ss_row_dev; c.epi.ss_row = a->ss_row; c.epi.ss_stride = a->ss_row_stride;
    c.epi.act = a->act;
    c.epi.residual = a->residual.ptr ? make_view(a->residual) : null_view();
    c.out = make_view_w(a->out);
    c.scratch = a->scratch_dev;
    c.B = a->batch;
    int precision = a->precision;
    SplitScope split(precision);
    return run_glin(L, c, precision, static_cast<cudaStream_t>(stream));
}

int sd_glin_forward_bf16(const sd_glin* L, const uint16_t* a_dev, const float* row_scale_dev, const float* ss_row_dev, int act,
                         const uint16_t* residual_dev, void* out_dev, int out_is_fp32, float* scratch_dev, int batch, void* stream) {
    if (!L || !a_dev || !out_dev) { set_error("sd_glin_forward_bf16: null argument"); return SD_ERR_INVALID; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int N = L->N, K = L->K, OUT = L->OUT;
    TcCall t;
    t.a0.ptr = reinterpret_cast<const __nv_bfloat16*>(a_dev); t.a0.sb = (long long)N * K; t.a0.sn = K; t.a0.width = K;
    t.a1.ptr = nullptr; t.a1.sb = t.a1.sn = 0; t.a1.width = 0;
    t.row_scale = row_scale_dev; t.B = batch; t.accurate_tanh = 0;
    const __nv_bfloat16* res = reinterpret_cast<const __nv_bfloat16*>(residual_dev);
    if (L->G == nullptr) {
        t.bias_node = L->bias_node; t.ss = ss_row_dev; t.act = act;
        t.res = res; t.res_sb = (long long)N * OUT; t.res_sn = OUT;
        t.out = out_dev; t.out_fp32 = out_is_fp32; t.out_sb = (long long)N * OUT; t.out_sn = OUT;
        return glin_tc_launch(L, t, st);
    }
    if (!scratch_dev) { set_error("sd_glin_forward_bf16: scratch required for non-identity G"); return SD_ERR_INVALID; }
    t.bias_node = nullptr; t.ss = nullptr; t.act = SD_ACT_NONE; t.res = nullptr; t.res_sb = t.res_sn = 0;
    t.out = scratch_dev; t.out_fp32 = 1; t.out_sb = (long long)N * OUT; t.out_sn = OUT;
    int rc = glin_tc_launch(L, t, st);
    if (rc) return rc;
    Epilogue e = no_epilogue(OUT);
    e.bias_node = L->bias_node; e.ss = ss_row_dev; e.act = act;
    return node_mix_to_bf16(L->G, N, OUT, scratch_dev, e, res, out_dev, out_is_fp32, batch, st);
}

int sd_node_attention(const float* qkv_dev, float* out_dev, int batch, int num_nodes, int heads, int dim_head, void* stream) {
    return node_attention_fp32(qkv_dev, out_dev, batch, num_nodes, heads, dim_head, static_cast<cudaStream_t>(stream));
}

int sd_row_inv_norm(const float* x_dev, float* inv_norm_dev, int64_t rows, int width, void* stream) {
    return row_inv_norm_fp32(x_dev, inv_norm_dev, rows, width, static_cast<cudaStream_t>(stream));
}

int sd_time_table(const float* times_dev, int n_rows, int C, float theta, int time_dim, const float* w1_dev,
                  const float* b1_dev, const float* w3_dev, const float* b3_dev, const float* const* head_w_dev_host,
                  const float* const* head_b_dev_host, int n_heads, float* table_dev, float* workspace_dev, void* stream) {
    if (n_rows <= 0 || C < 4 || (C & 1) || !table_dev || !workspace_dev) { set_error("sd_time_table: invalid arguments"); return SD_ERR_INVALID; }
    return time_table_fp32(times_dev, n_rows, C, theta, time_dim, w1_dev, b1_dev, w3_dev, b3_dev, head_w_dev_host,
                           head_b_dev_host, n_heads, table_dev, workspace_dev, static_cast<cudaStream_t>(stream));
}

// ------------------------------------------------------------------------------- denoiser
int sd_denoiser_create(int num_nodes, int dim, int cond_dim, int out_dim, int depth, int heads, int dim_head, sd_denoiser** out) {
    if (!out || num_nodes <= 0 || num_nodes > SD_MAX_NODES || dim <= 0 || cond_dim < 0 || depth <= 0 || 1 + 8 * depth + 4 > SD_MAX_SLOTS) {
        set_error("sd_denoiser_create: invalid arguments");
        return SD_ERR_INVALID;
    }
    sd_denoiser* d = new (std::nothrow) sd_denoiser();
    if (!d) { set_error("out of host memory"); return SD_ERR_INVALID; }
    d->N = num_nodes; d->dim = dim; d->cond_dim = cond_dim; d->out_dim = out_dim; d->depth = depth;
    d->heads = heads; d->dim_head = dim_head; d->C = dim + cond_dim;
    for (int i = 0; i < SD_MAX_SLOTS; ++i) d->slot[i] = nullptr;
    d->time_table = nullptr; d->time_rows = 0;
    *out = d;
    return SD_OK;
}

int sd_denoiser_set_layer(sd_denoiser* d, int slot, const sd_glin* layer) {
    if (!d || slot < 0 || slot >= 1 + 8 * d->depth + 4) { set_error("sd_denoiser_set_layer: bad slot %d", slot); return SD_ERR_INVALID; }
    d->slot[slot] = layer;
    return SD_OK;
}

int sd_denoiser_set_time_table(sd_denoiser* d, const float* table_dev, int n_rows) {
    if (!d) return SD_ERR_INVALID;
    d->time_table = table_dev; d->time_rows = n_rows;
    return SD_OK;
}

void sd_denoiser_destroy(sd_denoiser* d) { delete d; }

static size_t denoiser_ws_floats(const sd_denoiser* d, int B) {
    const size_t rows = (size_t)B * d->N;
    const size_t hd = (size_t)d->heads * d->dim_head;
    size_t maxo = 3 * hd > (size_t)d->C ? 3 * hd : (size_t)d->C;
    // r, x, h, res (C each) + qkv (3hd) + att (hd) + scratch (maxo) + inv-norm (1) ; each sub-buffer padded by 64 floats
    return rows * (4 * (size_t)d->C + 3 * hd + hd + maxo + 1) + 16 * 64;
}

size_t sd_denoiser_workspace_bytes(const sd_denoiser* d, int batch, int precision) {
    (void)precision;
    if (!d || batch <= 0) return 0;
    return denoiser_ws_floats(d, batch) * sizeof(float) + 16 * 256;
}

int sd_denoiser_forward(const sd_denoiser* d, const sd_view* x, const sd_view* x_cond, const int32_t* t_rows_dev, int t_row,
                        float* out_dev, int batch, void* workspace_dev, int precision, void* stream) {
    if (batch == 0) return SD_OK;   // an empty batch is a no-op (the reference returns empty tensors; zero-size device buffers have null pointers)
    if (!d || !x || !out_dev || !workspace_dev) { set_error("sd_denoiser_forward: null argument"); return SD_ERR_INVALID; }
    if (!d->time_table) { set_error("sd_denoiser_forward: time table not set"); return SD_ERR_INVALID; }
    if (!t_rows_dev && (t_row < 0 || t_row >= d->time_rows)) { set_error("sd_denoiser_forward: time row %d outside table of %d rows", t_row, d->time_rows); return SD_ERR_INVALID; }
    if ((d->cond_dim > 0) != (x_cond != nullptr && x_cond->ptr != nullptr)) { set_error("sd_denoiser_forward: x_cond presence does not match cond_dim=%d", d->cond_dim); return SD_ERR_INVALID; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    SplitScope split(precision);
    if (precision == SD_PREC_BF16) return denoiser_forward_bf16(d, x, x_cond, t_rows_dev, t_row, out_dev, batch, workspace_dev, st);
    if (precision != SD_PREC_FP32 && precision != SD_PREC_BF16X3) { set_error("sd_denoiser_forward: precision %d not available", precision); return SD_ERR_UNSUPPORTED; }
    if (precision == SD_PREC_BF16X3 && t_rows_dev) precision = SD_PREC_FP32;   // per-sample times: FFMA kernels (both are fp32-grade)
    const int B = batch, N = d->N, C = d->C, hd = d->heads * d->dim_head;
    const size_t rows = (size_t)B * N;
    Arena ar(workspace_dev);
    float* r = ar.floats(rows * C);
    float* xb = ar.floats(rows * C);
    float* h = ar.floats(rows * C);
    float* res = ar.floats(rows * C);
    float* qkv = ar.floats(rows * 3 * hd);
    float* att = ar.floats(rows * hd);
    float* scratch = ar.floats(rows * (size_t)(3 * hd > C ? 3 * hd : C));
    float* inv = ar.floats(rows);

    const int n_pairs = 2 * d->depth;
    const int n_heads_tab = n_pairs + 1;
    const long long ss_stride = (long long)n_heads_tab * 2 * C;
    auto call = [&](const sd_glin* L, View a0, View a1, const float* row_scale, int ss_head, int act, const float* resid, float* out, int out_w) -> int {
        GlinCall c;
        c.a0 = a0; c.a1 = a1; c.row_scale = row_scale;
        c.epi = no_epilogue(out_w);
        if (ss_head >= 0) { c.epi.ss = d->time_table + (long long)ss_head * 2 * C; c.epi.ss_row_idx = t_rows_dev; c.epi.ss_row = t_row; c.epi.ss_stride = ss_stride; }
        c.epi.act = act;
        c.epi.residual = resid ? contiguous_view(resid, N, out_w) : null_view();
        c.out = contiguous_view_w(out, N, out_w);
        c.scratch = scratch; c.B = B;
        return run_glin(L, c, precision, st);
    };
    int rc;
    // init_lin on cat([x_cond, x])  (generator.py:91-95); r keeps the skip copy
    View vx = make_view(*x);
    if (d->cond_dim > 0) rc = call(d->slot[0], make_view(*x_cond), vx, nullptr, -1, SD_ACT_NONE, nullptr, r, C);
    else rc = call(d->slot[0], vx, null_view(), nullptr, -1, SD_ACT_NONE, nullptr, r, C);
    if (rc) return rc;
    const float* cur = r;
    for (int i = 0; i < n_pairs; ++i) {
        const int s0 = 1 + 4 * i;
        // ResnetBlock (attention.py:91-102)
        rc = call(d->slot[s0], contiguous_view(cur, N, C), null_view(), nullptr, i, SD_ACT_TANH, nullptr, h, C);
        if (rc) return rc;
        {   // block2 + residual; when an attention block follows, its RMSNorm row factors come out of this layer's epilogue if the
            // kernel that runs it can provide them (GlinCall::norm_out), else from the separate pass below
            GlinCall c;
            c.a0 = contiguous_view(h, N, C); c.a1 = null_view(); c.row_scale = nullptr;
            c.epi = no_epilogue(C);
            c.epi.act = SD_ACT_TANH;
            c.epi.residual = contiguous_view(cur, N, C);
            c.out = contiguous_view_w(xb, N, C);
            c.scratch = scratch; c.B = B;
            c.norm_out = (i != n_pairs - 1) ? inv : nullptr;
            (void)tc3_take_norm_written();
            rc = run_glin(d->slot[s0 + 1], c, precision, st);
            if (rc) return rc;
        }
        const bool norm_fused = tc3_take_norm_written();
        cur = xb;
        if (i != n_pairs - 1) {
            // Residual(PreNorm(Attention))  (attention.py:16-17, 44-46, 122-136)
            if (!norm_fused) {
                rc = row_inv_norm_fp32(xb, inv, (long long)rows, C, st);
                if (rc) return rc;
            }
            const sd_glin* Lq = d->slot[s0 + 2];
            if (Lq && Lq->G != nullptr && Lq->G_host && !Lq->bias_node && node_attention_mix_supported(N, d->heads, d->dim_head, qkv, att)) {
                // dense graph influence on to_qkv: the GEMM writes the RAW products and the attention kernel mixes the nodes
                // (and applies the RMSNorm row factors) in shared memory: no [B, N, 768] mix pass through HBM
                sd_glin Lraw = *Lq; Lraw.G = nullptr; Lraw.G_host = nullptr;
                rc = call(&Lraw, contiguous_view(xb, N, C), null_view(), nullptr, -1, SD_ACT_NONE, nullptr, qkv, 3 * hd);
                if (rc) return rc;
                rc = node_attention_mix_fp32(Lq->G_host, inv, qkv, att, B, N, d->heads, d->dim_head, st);
                if (rc) return rc;
            } else {
                rc = call(Lq, contiguous_view(xb, N, C), null_view(), inv, -1, SD_ACT_NONE, nullptr, qkv, 3 * hd);
                if (rc) return rc;
                rc = node_attention_fp32(qkv, att, B, N, d->heads, d->dim_head, st);
                if (rc) return rc;
            }
            rc = call(d->slot[s0 + 3], contiguous_view(att, N, hd), null_view(), nullptr, -1, SD_ACT_NONE, xb, xb, C);
            if (rc) return rc;
        }
    }
    // final_res_block on cat([x, r])  (generator.py:104-106)
    const int sf = 1 + 8 * d->depth;
    View vcur = contiguous_view(cur, N, C), vr = contiguous_view(r, N, C);
    rc = call(d->slot[sf + 2], vcur, vr, nullptr, -1, SD_ACT_NONE, nullptr, res, C);
    if (rc) return rc;
    rc = call(d->slot[sf], vcur, vr, nullptr, n_pairs, SD_ACT_TANH, nullptr, h, C);
    if (rc) return rc;
    rc = call(d->slot[sf + 1], contiguous_view(h, N, C), null_view(), nullptr, -1, SD_ACT_TANH, res, xb, C);
    if (rc) return rc;
    return call(d->slot[sf + 3], contiguous_view(xb, N, C), null_view(), nullptr, -1, SD_ACT_NONE, nullptr, out_dev, d->out_dim);
}

// ------------------------------------------------------------------------------- diffusion
static bool is_diagonal(const float* m, int N) {
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j)
            if (i != j && m[i * N + j] != 0.0f) return false;
    return true;
}

int sd_diffusion_create(int num_nodes, int latent_dim, int timesteps, const float* c1_dev, const float* c2_dev, const float* s_dev,
                        const float* c1_host, const float* c2_host, const float* s_host, sd_diffusion** out) {
    if (!out || num_nodes <= 0 || num_nodes > SD_MAX_NODES || timesteps <= 0 || timesteps > 4096 || !c1_dev || !c2_dev || !s_dev) {
        set_error("sd_diffusion_create: invalid arguments (N=%d T=%d)", num_nodes, timesteps);
        return SD_ERR_INVALID;
    }
    sd_diffusion* d = new (std::nothrow) sd_diffusion();
    if (!d) { set_error("out of host memory"); return SD_ERR_INVALID; }
    d->N = num_nodes; d->D = latent_dim; d->T = timesteps; d->c1 = c1_dev; d->c2 = c2_dev; d->s = s_dev;
    for (int t = 0; t < timesteps; ++t) {
        const size_t o = (size_t)t * num_nodes * num_nodes;
        d->diagonal[t] = (c1_host && c2_host && s_host && is_diagonal(c1_host + o, num_nodes) &&
                          is_diagonal(c2_host + o, num_nodes) && is_diagonal(s_host + o, num_nodes)) ? 1 : 0;
    }
    *out = d;
    return SD_OK;
}

void sd_diffusion_destroy(sd_diffusion* d) { delete d; }

int sd_reverse_step(const sd_diffusion* d, const float* x_t_dev, const float* x0_dev, const sd_view* eps, float* x_out_dev,
                    float* mean_out_dev, int t, int batch, int clip_denoised, void* stream) {
    if (batch == 0) return SD_OK;
    if (!d || !x_t_dev || !x0_dev || !x_out_dev) { set_error("sd_reverse_step: null argument"); return SD_ERR_INVALID; }
    if (x_out_dev == x_t_dev || x_out_dev == x0_dev) { set_error("sd_reverse_step: output must not alias an input"); return SD_ERR_INVALID; }
    View e = null_view();
    if (eps && eps->ptr) e = make_view(*eps);
    return reverse_step_fp32(d, x_t_dev, x0_dev, &e, x_out_dev, mean_out_dev, (long long)d->N * d->D, t, batch, clip_denoised,
                             static_cast<cudaStream_t>(stream));
}

int sd_q_sample(const float* x0_dev, const float* eps_dev, const int32_t* t_dev, const float* sqrt_ac_dev, const float* m_dev,
                float* out_dev, int batch, int num_nodes, int latent_dim, void* stream) {
    return q_sample_fp32(x0_dev, eps_dev, t_dev, sqrt_ac_dev, m_dev, out_dev, batch, num_nodes, latent_dim, static_cast<cudaStream_t>(stream));
}

int sd_mahalanobis_loss(const float* out_dev, const float* x0_dev, const int32_t* t_dev, const float* s_dev, float* loss_dev,
                        int batch, int num_nodes, int latent_dim, void* stream) {
    return mahalanobis_loss_fp32(out_dev, x0_dev, t_dev, s_dev, loss_dev, batch, num_nodes, latent_dim, static_cast<cudaStream_t>(stream));
}

int sd_fill_normal(float* out_dev, int64_t count, uint64_t seed, uint64_t offset, void* stream) {
    return fill_normal(out_dev, count, seed, offset, static_cast<cudaStream_t>(stream));
}

int sd_motion_metrics(const float* pred_dev, const float* target_dev, int windows, int samples, int frames, int feat,
                      float scale, float* ade_dev, float* fde_dev, float* apd_dev, void* stream) {
    if (windows == 0) return SD_OK;
    if (windows < 0 || samples < 1 || frames < 1 || feat < 1) { set_error("sd_motion_metrics: bad shape [%d, %d, %d, %d]", windows, samples, frames, feat); return SD_ERR_INVALID; }
    if (!pred_dev || !target_dev || !(ade_dev || fde_dev || apd_dev)) { set_error("sd_motion_metrics: null argument"); return SD_ERR_INVALID; }
    return motion_metrics_fp32(pred_dev, target_dev, windows, samples, frames, feat, scale, ade_dev, fde_dev, apd_dev, static_cast<cudaStream_t>(stream));
}

int sd_multimodal_metrics(const float* pred_dev, const float* mm_gt_dev, const int32_t* gt_window_dev, const int32_t* gt_offsets_dev,
                          int windows, int n_gt, int samples, int frames, int feat, float scale, float* mmade_dev, float* mmfde_dev,
                          float* scratch_dev, void* stream) {
    if (windows == 0) return SD_OK;
    if (windows < 0 || n_gt < 0 || samples < 1 || frames < 1 || feat < 1) { set_error("sd_multimodal_metrics: bad shape"); return SD_ERR_INVALID; }
    if (!pred_dev || !gt_offsets_dev || !(mmade_dev || mmfde_dev) || (n_gt > 0 && (!mm_gt_dev || !gt_window_dev || !scratch_dev))) {
        set_error("sd_multimodal_metrics: null argument"); return SD_ERR_INVALID;
    }
    return multimodal_metrics_fp32(pred_dev, mm_gt_dev, gt_window_dev, gt_offsets_dev, windows, n_gt, samples, frames, feat, scale,
                                   mmade_dev, mmfde_dev, scratch_dev, static_cast<cudaStream_t>(stream));
}

int sd_best_sample(const float* pred_dev, const float* target_dev, int windows, int samples, int frames, int joints, int keep_frames,
                   float scale, float* best_dev, float* tail_dev, int32_t* index_dev, void* stream) {
    if (windows == 0) return SD_OK;
    if (windows < 0 || samples < 1 || frames < 1 || joints < 1 || !pred_dev || !target_dev) { set_error("sd_best_sample: invalid arguments"); return SD_ERR_INVALID; }
    return best_sample_fp32(pred_dev, target_dev, windows, samples, frames, joints, keep_frames, scale, best_dev, tail_dev, index_dev,
                            static_cast<cudaStream_t>(stream));
}

size_t sd_sample_workspace_bytes(const sd_diffusion* df, const sd_denoiser* dn, int batch, int precision) {
    if (!df || !dn || batch <= 0) return 0;
    const size_t lat = align_up((size_t)batch * df->N * df->D * sizeof(float));
    return 2 * lat + sd_denoiser_workspace_bytes(dn, batch, precision) + 1024;
}

int sd_sample_loop(const sd_diffusion* df, const sd_denoiser* dn, float* x_dev, const sd_view* x_cond, const float* sampling_noise_dev,
                   float* means_out_dev, int batch, int clip_denoised, void* workspace_dev, int precision, void* stream) {
    if (batch == 0) return SD_OK;
    if (!df || !dn || !x_dev || !workspace_dev) { set_error("sd_sample_loop: null argument"); return SD_ERR_INVALID; }
    if (df->N != dn->N || df->D != dn->dim || dn->out_dim != dn->dim) { set_error("sd_sample_loop: diffusion/denoiser shape mismatch"); return SD_ERR_INVALID; }
    if (dn->time_rows < df->T) { set_error("sd_sample_loop: time table has %d rows, need %d", dn->time_rows, df->T); return SD_ERR_INVALID; }
    if (df->T > 1 && !sampling_noise_dev) { set_error("sd_sample_loop: sampling noise required"); return SD_ERR_INVALID; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    SplitScope split(precision);
    const int B = batch, N = df->N, D = df->D, T = df->T;
    const size_t lat = (size_t)B * N * D;
    Arena ar(workspace_dev);
    float* x0 = ar.floats(lat);
    float* alt = ar.floats(lat);
    void* dws = ar.base + ar.off;
    float* cur = x_dev;
    float* nxt = alt;
    for (int t = T - 1; t >= 0; --t) {
        sd_view xv; xv.ptr = cur; xv.sample_stride = (int64_t)N * D; xv.node_stride = D; xv.rep = 1; xv.width = D;
        int rc = sd_denoiser_forward(dn, &xv, x_cond, nullptr, t, x0, B, dws, precision, stream);
        if (rc) return rc;
        View eps = null_view();
        float* mean = nullptr;
        if (t > 0) {   // sampling_noise[:, (T-1) - t]   (base.py:330-331)
            const long long idx = (T - 1) - t;
            eps.ptr = sampling_noise_dev + idx * N * D; eps.sb = (long long)(T - 1) * N * D; eps.sn = D; eps.rep = 1; eps.width = D;
            if (means_out_dev) mean = means_out_dev + idx * N * D;
        }
        rc = reverse_step_fp32(df, cur, x0, &eps, nxt, mean, (long long)(T - 1) * N * D, t, B, clip_denoised, st);
        if (rc) return rc;
        float* tmp = cur; cur = nxt; nxt = tmp;
    }
    if (cur != x_dev) SD_CUDA_OK(cudaMemcpyAsync(x_dev, cur, lat * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return SD_OK;
}

// ------------------------------------------------------------------------------- graph GRU
int sd_gru_create(int num_nodes, const int32_t* node_types_host, int n_types, int input_size, int hidden_size, const float* w_ih_dev,
                  const float* w_hh_dev, const float* bias_ih_seq_dev, const float* bias_hh_seq_dev, const float* gx_seq_dev, int steps,
                  sd_gru** out) {
    if (!out || num_nodes <= 0 || num_nodes > SD_MAX_NODES || input_size <= 0 || hidden_size <= 0 || !w_ih_dev || !w_hh_dev || steps <= 0 || n_types <= 0) {
        set_error("sd_gru_create: invalid arguments");
        return SD_ERR_INVALID;
    }
    sd_gru* g = new (std::nothrow) sd_gru();
    if (!g) { set_error("out of host memory"); return SD_ERR_INVALID; }
    g->N = num_nodes; g->n_types = n_types; g->IN = input_size; g->H = hidden_size; g->steps = steps;
    for (int n = 0; n < SD_MAX_NODES; ++n) g->types.t[n] = 0;
    for (int n = 0; n < num_nodes; ++n) {
        const int t = node_types_host ? node_types_host[n] : 0;
        if (t < 0 || t >= n_types) { delete g; set_error("sd_gru_create: node type %d outside [0,%d)", t, n_types); return SD_ERR_INVALID; }
        g->types.t[n] = (unsigned char)t;
    }
    g->W_ih = w_ih_dev; g->W_hh = w_hh_dev; g->bias_ih_seq = bias_ih_seq_dev; g->bias_hh_seq = bias_hh_seq_dev; g->gx_seq = gx_seq_dev;
    g->W_ih_perm = g->W_hh_perm = g->bias_ih_perm = g->bias_hh_perm = nullptr;
    g->W_hh_planes = nullptr; g->W_hh_f16 = nullptr; g->W_hh_perm_f16 = nullptr; g->gx_host = nullptr;
    if (gx_seq_dev) {   // host copy: the per-sample gate kernel takes gx_i by value (constant bank)
        const size_t n = (size_t)steps * num_nodes * num_nodes;
        g->gx_host = new (std::nothrow) float[n];
        if (!g->gx_host || cudaMemcpy(g->gx_host, gx_seq_dev, n * sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess) {
            delete[] g->gx_host; delete g; set_error("sd_gru_create: cannot read the graph-influence sequence"); return SD_ERR_CUDA;
        }
    }
    *out = g;
    return SD_OK;
}

int sd_gru_set_bf16x3(sd_gru* g, const uint16_t* w_hh_planes_dev) {
    if (!g || !w_hh_planes_dev) { set_error("sd_gru_set_bf16x3: null argument"); return SD_ERR_INVALID; }
    g->W_hh_planes = w_hh_planes_dev;
    return SD_OK;
}

int sd_gru_set_f16x2(sd_gru* g, const uint16_t* w_hh_f16_dev) {
    if (!g || !w_hh_f16_dev) { set_error("sd_gru_set_f16x2: null argument"); return SD_ERR_INVALID; }
    g->W_hh_f16 = w_hh_f16_dev;
    return SD_OK;
}

int sd_gru_set_fused(sd_gru* g, const float* w_ih_perm_dev, const float* w_hh_perm_dev, const float* bias_ih_perm_dev,
                     const float* bias_hh_perm_dev) {
    if (!g || !w_ih_perm_dev || !w_hh_perm_dev || !bias_ih_perm_dev || !bias_hh_perm_dev) { set_error("sd_gru_set_fused: null argument"); return SD_ERR_INVALID; }
    if (g->gx_seq) { set_error("sd_gru_set_fused: only for cells whose graph-influence sequence is the identity"); return SD_ERR_INVALID; }
    if (g->H % 32) { set_error("sd_gru_set_fused: hidden size must be a multiple of 32"); return SD_ERR_UNSUPPORTED; }
    g->W_ih_perm = w_ih_perm_dev; g->W_hh_perm = w_hh_perm_dev; g->bias_ih_perm = bias_ih_perm_dev; g->bias_hh_perm = bias_hh_perm_dev;
    return SD_OK;
}

int sd_gru_set_fused_f16x2(sd_gru* g, const uint16_t* w_hh_perm_f16_dev) {
    if (!g || !w_hh_perm_f16_dev) { set_error("sd_gru_set_fused_f16x2: null argument"); return SD_ERR_INVALID; }
    if (!g->W_hh_perm) { set_error("sd_gru_set_fused_f16x2: call sd_gru_set_fused first (identity graph influence only)"); return SD_ERR_INVALID; }
    g->W_hh_perm_f16 = w_hh_perm_f16_dev;
    return SD_OK;
}

void sd_gru_destroy(sd_gru* g) { if (g) delete[] g->gx_host; delete g; }

// raw recurrent product hr = h @ W_hh^T: tcgen05 3-plane kernel (fp32-grade) when the planes are set and the precision allows it,
// the exact FFMA kernel otherwise
static int gru_recurrent_product(const sd_gru* g, const View& h, float* hr, int B, int precision, cudaStream_t st);

// raw grouped product out = a @ W^T (no bias, no mix)
static int gru_product(const float* W, int K, int OUT, const NodeTypes& types, int N, View a0, View a1, ViewW out, int B, cudaStream_t st) {
    GlinCall c;
    c.a0 = a0; c.a1 = a1; c.row_scale = nullptr; c.epi = no_epilogue(OUT); c.out = out; c.scratch = nullptr; c.B = B;
    return glin_forward_fp32(W, nullptr, K, OUT, types, N, nullptr, c, st);
}

static int gru_recurrent_product(const sd_gru* g, const View& h, float* hr, int B, int precision, cudaStream_t st) {
    const int N = g->N, H = g->H;
    if (precision != SD_PREC_FP32 && g->W_hh_planes && glin_tc3_supported(H, 0, 3 * H) &&
        (reinterpret_cast<uintptr_t>(h.ptr) & 15u) == 0 && h.sb % 4 == 0 && h.sn % 4 == 0) {
        sd_glin rec;                                   // W_hh viewed as a graph-linear H -> 3H without bias or mix
        rec.N = N; rec.n_types = g->n_types; rec.K = H; rec.OUT = 3 * H; rec.types = g->types; rec.W = g->W_hh; rec.Wt = nullptr;
        rec.bias_node = nullptr; rec.G = nullptr; rec.W_bf16 = g->W_hh_planes; rec.planes = 3; rec.W_f16 = g->W_hh_f16; rec.G_host = nullptr;
        GlinCall c;
        c.a0 = h; c.a1 = null_view(); c.row_scale = nullptr; c.epi = no_epilogue(3 * H); c.scratch = nullptr; c.B = B;
        c.out = contiguous_view_w(hr, N, 3 * H);
        return glin_tc3_launch(&rec, c, c.out, false, st);
    }
    return gru_product(g->W_hh, H, 3 * H, g->types, N, h, null_view(), contiguous_view_w(hr, N, 3 * H), B, st);
}

size_t sd_encode_workspace_bytes(int windows, int obs_len, int num_nodes, int hidden, int layers) {
    (void)layers;
    const size_t wn = (size_t)windows * num_nodes;
    // h0 + hr + hr_m + xr_m (wn*3H each, h0 wn*H) + xr_all (wn*T*3H) + two sequences (wn*T*H) + glin scratch (wn*3H)
    const size_t fl = wn * hidden + 4 * wn * 3 * hidden + wn * obs_len * 3 * (size_t)hidden + 2 * wn * obs_len * (size_t)hidden;
    return fl * sizeof(float) + 16 * 256;
}

int sd_encode(const sd_glin* initial_hidden, sd_gru* const* layers_host, int n_layers, const sd_glin* fc, const float* obs_dev,
              int windows, int obs_len, int feat, float* z_dev, int final_act, void* workspace_dev, int precision, void* stream) {
    if (windows == 0) return SD_OK;
    if (!initial_hidden || !layers_host || n_layers <= 0 || !fc || !obs_dev || !z_dev || !workspace_dev) { set_error("sd_encode: null argument"); return SD_ERR_INVALID; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    SplitScope split(precision);
    const int W = windows, T = obs_len, N = initial_hidden->N, H = layers_host[0]->H;
    const size_t wn = (size_t)W * N;
    Arena ar(workspace_dev);
    float* h0 = ar.floats(wn * H);
    float* hr = ar.floats(wn * 3 * H);
    float* hr_m = ar.floats(wn * 3 * H);
    float* xr_m = ar.floats(wn * 3 * H);
    float* scratch = ar.floats(wn * 3 * H);
    float* xr_all = ar.floats(wn * T * 3 * H);
    float* seq_a = ar.floats(wn * T * H);
    float* seq_b = ar.floats(wn * T * H);
    int rc;
    {   // h0 = initial_hidden1(x[:, 0])   (encoder.py:66)
        GlinCall c;
        c.a0.ptr = obs_dev; c.a0.sb = (long long)T * N * feat; c.a0.sn = feat; c.a0.rep = 1; c.a0.width = feat;
        c.a1 = null_view(); c.row_scale = nullptr; c.epi = no_epilogue(H);
        c.out = contiguous_view_w(h0, N, H); c.scratch = scratch; c.B = W;
        rc = run_glin(initial_hidden, c, SD_PREC_FP32, st);
        if (rc) return rc;
    }
    const float* seq_in = obs_dev; int in_w = feat;
    float* seq_out = seq_a;
    for (int l = 0; l < n_layers; ++l) {
        const sd_gru* g = layers_host[l];
        if (!g || g->IN != in_w || g->H != H || g->steps < T) { set_error("sd_encode: GRU layer %d shape/steps mismatch", l); return SD_ERR_INVALID; }
        // x-side products of all frames at once: rows (w, t) -> [W*T, N, 3H]
        View a; a.ptr = seq_in; a.sb = (long long)N * in_w; a.sn = in_w; a.rep = 1; a.width = in_w;
        View xr0; xr0.ptr = xr_all; xr0.sb = (long long)T * N * 3 * H; xr0.sn = 3 * H; xr0.rep = 1; xr0.width = 3 * H;
        ViewW ho0; ho0.ptr = seq_out; ho0.sb = (long long)T * N * H; ho0.sn = H; ho0.rep = 1; ho0.width = H;
        // per-sample gate kernel (any graph influence) for dense gx, and for every cell on the tensor-core precisions
        const bool sample_cell = (g->gx_seq ? g->gx_host != nullptr : precision != SD_PREC_FP32) &&
                                 gru_sample_supported(N, H, hr, xr0, contiguous_view(h0, N, H), ho0);
        const bool fused_cell = !sample_cell && g->W_hh_perm != nullptr;
        rc = gru_product(fused_cell ? g->W_ih_perm : g->W_ih, in_w, 3 * H, g->types, N, a, null_view(), contiguous_view_w(xr_all, N, 3 * H), W * T, st);
        if (rc) return rc;
        for (int t = 0; t < T; ++t) {
            View h_in;
            if (t == 0) h_in = contiguous_view(h0, N, H);
            else { h_in.ptr = seq_out + (size_t)(t - 1) * N * H; h_in.sb = (long long)T * N * H; h_in.sn = H; h_in.rep = 1; h_in.width = H; }
            if (sample_cell) {
                rc = gru_recurrent_product(g, h_in, hr, W, precision, st);
                if (rc) return rc;
                View xr = xr0; xr.ptr = xr_all + (size_t)t * N * 3 * H;
                ViewW h_out = ho0; h_out.ptr = seq_out + (size_t)t * N * H;
                rc = gru_sample_fp32(g->gx_host ? g->gx_host + (size_t)t * N * N : nullptr, N, H, hr, xr, g->bias_ih_seq + (size_t)t * N * 3 * H,
                                     g->bias_hh_seq + (size_t)t * N * 3 * H, h_in, h_out, W, st);
                if (rc) return rc;
                continue;
            }
            if (fused_cell) {   // identity graph influence: h @ W_hh^T and the gates in one FFMA2 kernel, written straight into seq[:, t]
                View xr; xr.ptr = xr_all + (size_t)t * N * 3 * H; xr.sb = (long long)T * N * 3 * H; xr.sn = 3 * H; xr.rep = 1; xr.width = 3 * H;
                ViewW h_out; h_out.ptr = seq_out + (size_t)t * N * H; h_out.sb = (long long)T * N * H; h_out.sn = H; h_out.rep = 1; h_out.width = H;
                if (H == 96 && h_in.rep == 1) rc = gru_step_fused(g->W_hh_perm, H, g->types, N, xr, g->bias_ih_perm, g->bias_hh_perm, h_in, h_out,
                                                                  nullptr, nullptr, nullptr, 0, W, st);
                else rc = gru_step_f2(g->W_hh_perm, H, g->types, N, xr, g->bias_ih_perm, g->bias_hh_perm, h_in, h_out, W, st);
                if (rc) return rc;
                continue;
            }
            rc = gru_product(g->W_hh, H, 3 * H, g->types, N, h_in, null_view(), contiguous_view_w(hr, N, 3 * H), W, st);
            if (rc) return rc;
            View xr; xr.ptr = xr_all + (size_t)t * N * 3 * H; xr.sb = (long long)T * N * 3 * H; xr.sn = 3 * H; xr.rep = 1; xr.width = 3 * H;
            const float* hr_use = hr;
            if (g->gx_seq) {
                const float* gx = g->gx_seq + (size_t)t * N * N;
                Epilogue e = no_epilogue(3 * H);
                rc = node_mix_fp32(gx, N, 3 * H, xr.ptr, xr.sb, nullptr, e, contiguous_view_w(xr_m, N, 3 * H), W, st);
                if (rc) return rc;
                rc = node_mix_fp32(gx, N, 3 * H, hr, (long long)N * 3 * H, nullptr, e, contiguous_view_w(hr_m, N, 3 * H), W, st);
                if (rc) return rc;
                xr = contiguous_view(xr_m, N, 3 * H);
                hr_use = hr_m;
            }
            ViewW h_out; h_out.ptr = seq_out + (size_t)t * N * H; h_out.sb = (long long)T * N * H; h_out.sn = H; h_out.rep = 1; h_out.width = H;
            rc = gru_gates_fp32(xr, g->bias_ih_seq + (size_t)t * N * 3 * H, hr_use, g->bias_hh_seq + (size_t)t * N * 3 * H, h_in, h_out, W, N, H, st);
            if (rc) return rc;
        }
        seq_in = seq_out; in_w = H;
        seq_out = (seq_out == seq_a) ? seq_b : seq_a;
    }
    // z = tanh(tanh(fc(y[:, -1])))   (encoder.py:81 + autoencoder.py:54)
    GlinCall c;
    c.a0.ptr = seq_in + (size_t)(T - 1) * N * H; c.a0.sb = (long long)T * N * H; c.a0.sn = H; c.a0.rep = 1; c.a0.width = H;
    c.a1 = null_view(); c.row_scale = nullptr; c.epi = no_epilogue(fc->OUT); c.epi.act = final_act;
    c.out = contiguous_view_w(z_dev, N, fc->OUT); c.scratch = scratch; c.B = W;
    if ((size_t)fc->OUT > (size_t)3 * H) { set_error("sd_encode: latent wider than 3*hidden unsupported"); return SD_ERR_UNSUPPORTED; }
    return run_glin(fc, c, SD_PREC_FP32, st);
}

size_t sd_decode_workspace_bytes(int batch, int num_nodes, int hidden) {
    const size_t bn = (size_t)batch * num_nodes;
    return (bn * hidden + 5 * bn * 3 * hidden) * sizeof(float) + 16 * 256;
}

int sd_decode(const sd_glin* initial_hidden, const sd_gru* cell, const sd_glin* fc, const sd_view* x_prev, const sd_view* x_last,
              const float* latent_dev, int batch, int ph, int feat, float* out_dev, void* workspace_dev, int precision, void* stream) {
    if (batch == 0) return SD_OK;
    if (!initial_hidden || !cell || !fc || !x_prev || !x_last || !latent_dev || !out_dev || !workspace_dev) { set_error("sd_decode: null argument"); return SD_ERR_INVALID; }
    if (cell->steps < ph) { set_error("sd_decode: GRU plan has %d steps, ph=%d", cell->steps, ph); return SD_ERR_INVALID; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    SplitScope split(precision);
    const int B = batch, N = cell->N, H = cell->H, L = cell->IN - feat;
    if (L <= 0 || initial_hidden->K != cell->IN || fc->K != H || fc->OUT != feat) { set_error("sd_decode: layer shapes inconsistent"); return SD_ERR_INVALID; }
    const size_t bn = (size_t)B * N;
    Arena ar(workspace_dev);
    float* h = ar.floats(bn * H);
    float* xr_raw = ar.floats(bn * 3 * H);
    float* hr = ar.floats(bn * 3 * H);
    float* xr_m = ar.floats(bn * 3 * H);
    float* hr_m = ar.floats(bn * 3 * H);
    float* scratch = ar.floats(bn * 3 * H);
    View lat = contiguous_view(latent_dev, N, L);
    int rc;
    {   // h0 = initial_hidden_h(cat[x_{-2}, latent])   (decoder.py:65,73)
        GlinCall c;
        c.a0 = make_view(*x_prev); c.a1 = lat; c.row_scale = nullptr; c.epi = no_epilogue(H);
        c.out = contiguous_view_w(h, N, H); c.scratch = scratch; c.B = B;
        rc = run_glin(initial_hidden, c, SD_PREC_FP32, st);
        if (rc) return rc;
    }
    {
        // Per step: raw recurrent product (tcgen05 3-plane kernel, or FFMA in the exact-fp32 mode), then the per-sample kernel that
        // mixes the nodes with gx_i and applies the gates (sd_mix.cu), then the output head with fc's own graph influence.
        // Used for every dense gx / fc.G, and on the tensor-core precisions also for the identity (faster than the FFMA2 step).
        const View hv = contiguous_view(h, N, H);
        const ViewW hw = contiguous_view_w(h, N, H);
        const bool dense = cell->gx_seq != nullptr || fc->G != nullptr;
        const bool ok = (cell->gx_seq == nullptr || cell->gx_host) && (fc->G == nullptr || fc->G_host) && gru_head_supported(N, H, feat) &&
                        gru_sample_supported(N, H, hr, contiguous_view(xr_raw, N, 3 * H), hv, hw);
        static int tcf_env = -1;       // SKELDIFF_GRU_TC_FUSED=0: recurrent product and gates as two kernels (A/B timing)
        if (tcf_env < 0) { const char* e = getenv("SKELDIFF_GRU_TC_FUSED"); tcf_env = (e && e[0] == '0') ? 0 : 1; }
        if (ok && tcf_env && !cell->gx_seq && cell->W_hh_perm_f16 && cell->W_ih_perm && cell->W_hh_planes && tc_split_planes() == 2 &&
            precision == SD_PREC_BF16X3 && H % 32 == 0) {
            // Identity graph influence on the two-plane tensor-core precision: ONE tcgen05 kernel per frame computes h W_hh^T (rows in the
            // gate-interleaved order) and applies the gates in its epilogue (sd_glin_tc3.cu, T3_ACT_GRU).  Per frame it reads h (twice:
            // operand and z h term, the second from L2), the loop-invariant x-side products and writes the new h: the [B, N, 3H] product
            // (619 MB each way at B = 25 600) never exists.  The state ping-pongs between h and the (now unused) hr buffer.
            rc = gru_product(cell->W_ih_perm, cell->IN, 3 * H, cell->types, N, make_view(*x_last), lat, contiguous_view_w(xr_raw, N, 3 * H), B, st);
            if (rc) return rc;
            sd_glin rec;
            rec.N = N; rec.n_types = cell->n_types; rec.K = H; rec.OUT = 3 * H; rec.types = cell->types; rec.W = nullptr; rec.Wt = nullptr;
            rec.bias_node = nullptr; rec.G = nullptr; rec.W_bf16 = cell->W_hh_planes; rec.planes = 3; rec.W_f16 = cell->W_hh_perm_f16; rec.G_host = nullptr;
            float* h_cur = h;
            float* h_nxt = hr;
            for (int i = 0; i < ph; ++i) {
                rc = glin_tc3_gru_step(&rec, contiguous_view(h_cur, N, H), contiguous_view(xr_raw, N, 3 * H), cell->bias_ih_perm, cell->bias_hh_perm,
                                       contiguous_view_w(h_nxt, N, H), B, st);
                if (rc) return rc;
                ViewW o; o.ptr = out_dev + (size_t)i * N * feat; o.sb = (long long)ph * N * feat; o.sn = feat; o.rep = 1; o.width = feat;
                rc = gru_head_fp32(fc->G, fc->W, fc->bias_node, fc->types, fc->n_types, N, H, feat, contiguous_view(h_nxt, N, H), o, SD_ACT_TANH, B, st);
                if (rc) return rc;
                float* t = h_cur; h_cur = h_nxt; h_nxt = t;
            }
            return SD_OK;
        }
        if (ok && (dense || precision != SD_PREC_FP32)) {
            rc = gru_product(cell->W_ih, cell->IN, 3 * H, cell->types, N, make_view(*x_last), lat, contiguous_view_w(xr_raw, N, 3 * H), B, st);
            if (rc) return rc;
            for (int i = 0; i < ph; ++i) {
                rc = gru_recurrent_product(cell, hv, hr, B, precision, st);
                if (rc) return rc;
                // in place: a (sample, 32 units) task reads exactly the h elements it overwrites
                static int elem_env = -1;      // SKELDIFF_GRU_ELEMWISE=0: per-sample kernel also for the identity influence (A/B timing)
                if (elem_env < 0) { const char* e = getenv("SKELDIFF_GRU_ELEMWISE"); elem_env = (e && e[0] == '0') ? 0 : 1; }
                ViewW o; o.ptr = out_dev + (size_t)i * N * feat; o.sb = (long long)ph * N * feat; o.sn = feat; o.rep = 1; o.width = feat;
                if (!cell->gx_host && elem_env) {
                    // identity graph influence: nothing couples the nodes, the gates are a plain elementwise pass over (b, n, unit):
                    // the streaming kernel (13 independent 16-byte loads per thread, MUFU gates) instead of the per-sample TMA-box
                    // ring: 0.43 -> 0.31 ms per frame (5.3 TB/s = 81 % of HBM), decode 105.6 -> 86.4 ms.  Fusing the output head
                    // into this pass (one warp per row, shuffle tree) was measured at 0.61 ms against 0.31 + 0.15 and dropped.
                    rc = gru_gates_fp32(contiguous_view(xr_raw, N, 3 * H), cell->bias_ih_seq + (size_t)i * N * 3 * H, hr,
                                        cell->bias_hh_seq + (size_t)i * N * 3 * H, hv, hw, B, N, H, st, fast_epilogue());
                } else {
                    rc = gru_sample_fp32(cell->gx_host ? cell->gx_host + (size_t)i * N * N : nullptr, N, H, hr, contiguous_view(xr_raw, N, 3 * H),
                                         cell->bias_ih_seq + (size_t)i * N * 3 * H, cell->bias_hh_seq + (size_t)i * N * 3 * H, hv, hw, B, st);
                }
                if (rc) return rc;
                rc = gru_head_fp32(fc->G, fc->W, fc->bias_node, fc->types, fc->n_types, N, H, feat, hv, o, SD_ACT_TANH, B, st);
                if (rc) return rc;
            }
            return SD_OK;
        }
    }
    if (cell->W_hh_perm && fc->G == nullptr && feat <= 4) {
        // identity graph influence: one fused FFMA2 kernel per step (h @ W_hh^T + gates) and one tiny output head.
        // The h state ping-pongs between two buffers (the three 96-column blocks of a row tile read all of h_{i-1}).
        rc = gru_product(cell->W_ih_perm, cell->IN, 3 * H, cell->types, N, make_view(*x_last), lat, contiguous_view_w(xr_raw, N, 3 * H), B, st);
        if (rc) return rc;
        float* h_cur = h;
        float* h_nxt = hr;      // [bn * 3H] buffer, only bn * H used here
        for (int i = 0; i < ph; ++i) {
            ViewW o; o.ptr = out_dev + (size_t)i * N * feat; o.sb = (long long)ph * N * feat; o.sn = feat; o.rep = 1; o.width = feat;
            if (H == 96 && feat <= 3) {   // one launch per step: products, gates and the output head
                rc = gru_step_fused(cell->W_hh_perm, H, cell->types, N, contiguous_view(xr_raw, N, 3 * H), cell->bias_ih_perm,
                                    cell->bias_hh_perm, contiguous_view(h_cur, N, H), contiguous_view_w(h_nxt, N, H), fc->W, fc->bias_node,
                                    &o, feat, B, st);
                if (rc) return rc;
            } else {
                rc = gru_step_f2(cell->W_hh_perm, H, cell->types, N, contiguous_view(xr_raw, N, 3 * H), cell->bias_ih_perm, cell->bias_hh_perm,
                                 contiguous_view(h_cur, N, H), contiguous_view_w(h_nxt, N, H), B, st);
                if (rc) return rc;
                rc = gru_out_fc_fp32(fc->W, fc->bias_node, fc->types, N, H, feat, h_nxt, o, SD_ACT_TANH, B, st);
                if (rc) return rc;
            }
            float* t = h_cur; h_cur = h_nxt; h_nxt = t;
        }
        return SD_OK;
    }
    // loop-invariant x-side product of rec_input = cat[x_{-1}, latent]   (decoder.py:81,93)
    rc = gru_product(cell->W_ih, cell->IN, 3 * H, cell->types, N, make_view(*x_last), lat, contiguous_view_w(xr_raw, N, 3 * H), B, st);
    if (rc) return rc;
    for (int i = 0; i < ph; ++i) {
        rc = gru_product(cell->W_hh, H, 3 * H, cell->types, N, contiguous_view(h, N, H), null_view(), contiguous_view_w(hr, N, 3 * H), B, st);
        if (rc) return rc;
        View xr = contiguous_view(xr_raw, N, 3 * H);
        const float* hr_use = hr;
        if (cell->gx_seq) {
            const float* gx = cell->gx_seq + (size_t)i * N * N;
            Epilogue e = no_epilogue(3 * H);
            rc = node_mix_fp32(gx, N, 3 * H, xr_raw, (long long)N * 3 * H, nullptr, e, contiguous_view_w(xr_m, N, 3 * H), B, st);
            if (rc) return rc;
            rc = node_mix_fp32(gx, N, 3 * H, hr, (long long)N * 3 * H, nullptr, e, contiguous_view_w(hr_m, N, 3 * H), B, st);
            if (rc) return rc;
            xr = contiguous_view(xr_m, N, 3 * H);
            hr_use = hr_m;
        }
        rc = gru_gates_fp32(xr, cell->bias_ih_seq + (size_t)i * N * 3 * H, hr_use, cell->bias_hh_seq + (size_t)i * N * 3 * H,
                            contiguous_view(h, N, H), contiguous_view_w(h, N, H), B, N, H, st);
        if (rc) return rc;
        // y_i = tanh(fc(h))   (decoder.py:97-98), written straight into out[:, i]
        GlinCall c;
        c.a0 = contiguous_view(h, N, H); c.a1 = null_view(); c.row_scale = nullptr; c.epi = no_epilogue(feat); c.epi.act = SD_ACT_TANH;
        c.out.ptr = out_dev + (size_t)i * N * feat; c.out.sb = (long long)ph * N * feat; c.out.sn = feat; c.out.rep = 1; c.out.width = feat;
        c.scratch = scratch; c.B = B;
        rc = run_glin(fc, c, SD_PREC_FP32, st);
        if (rc) return rc;
    }
    return SD_OK;
}

}  // extern "C"
