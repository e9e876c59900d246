// Shared device/host helpers for libskeldiff_sm100a (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/skeldiff_b200.h"

namespace sd {

struct NodeTypes { unsigned char t[SD_MAX_NODES]; };

// device-side mirror of sd_view (const pointer)
struct View {
    const float* ptr;
    long long sb;   // sample stride (elements)
    long long sn;   // node stride (elements)
    int rep;        // row b reads sample b / rep
    int width;
};
struct ViewW {
    float* ptr;
    long long sb, sn;
    int rep, width;
};

__host__ __device__ inline View make_view(const sd_view& v) {
    View r; r.ptr = v.ptr; r.sb = v.sample_stride; r.sn = v.node_stride; r.rep = v.rep > 0 ? v.rep : 1; r.width = v.width; return r;
}
__host__ __device__ inline ViewW make_view_w(const sd_view& v) {
    ViewW r; r.ptr = v.ptr; r.sb = v.sample_stride; r.sn = v.node_stride; r.rep = v.rep > 0 ? v.rep : 1; r.width = v.width; return r;
}
__host__ __device__ inline View contiguous_view(const float* p, int num_nodes, int width) {
    View r; r.ptr = p; r.sb = (long long)num_nodes * width; r.sn = width; r.rep = 1; r.width = width; return r;
}
__host__ __device__ inline ViewW contiguous_view_w(float* p, int num_nodes, int width) {
    ViewW r; r.ptr = p; r.sb = (long long)num_nodes * width; r.sn = width; r.rep = 1; r.width = width; return r;
}
// rep is a kernel parameter (uniform): views without an in-place repeat skip the ~20-instruction integer division
__device__ __forceinline__ const float* row_ptr(const View& v, int b, int n) {
    const int s = v.rep == 1 ? b : b / v.rep;
    return v.ptr + (long long)s * v.sb + (long long)n * v.sn;
}
__device__ __forceinline__ float* row_ptr(const ViewW& v, int b, int n) {
    const int s = v.rep == 1 ? b : b / v.rep;
    return v.ptr + (long long)s * v.sb + (long long)n * v.sn;
}

// fused epilogue description shared by the GEMM and the node-mix kernels
struct Epilogue {
    const float* bias_node;   // [N][OUT] or null
    const float* ss;          // scale/shift rows or null: scale at [o], shift at [OUT + o]
    const int*   ss_row_idx;  // per-sample row or null
    int          ss_row;
    long long    ss_stride;
    int          act;
    View         residual;    // ptr null if none
    int          OUT;
};

__device__ __forceinline__ float epilogue_apply(const Epilogue& e, int b, int n, int o, float v) {
    if (e.bias_node) v += __ldg(e.bias_node + (long long)n * e.OUT + o);
    if (e.ss) {
        const int r = e.ss_row_idx ? __ldg(e.ss_row_idx + b) : e.ss_row;
        const float* row = e.ss + (long long)r * e.ss_stride;
        v = fmaf(v, __ldg(row + o) + 1.0f, __ldg(row + e.OUT + o));
    }
    if (e.act == SD_ACT_TANH) v = tanhf(v);
    else if (e.act == SD_ACT_TANH_TANH) v = tanhf(tanhf(v));
    if (e.residual.ptr) v += __ldg(row_ptr(e.residual, b, n) + o);
    return v;
}

void set_error(const char* fmt, ...);
int  check_cuda(cudaError_t e, const char* what);
void count_launch();   // every kernel launch of the library is counted (sd_launch_count)

}  // namespace sd

#define SD_CUDA_OK(expr) do { if (sd::check_cuda((expr), #expr)) return SD_ERR_CUDA; } while (0)
#define SD_LAUNCH_OK(what) do { sd::count_launch(); if (sd::check_cuda(cudaGetLastError(), what)) return SD_ERR_CUDA; } while (0)
