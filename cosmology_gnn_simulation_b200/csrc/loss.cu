// K6: loss reduction + gradient seeds (train.py:107-118,255-260), deterministic two-stage sums.
//   total = w_a * mean((a - y_a)^2) + w_t * mean((tau - y_t)^2) + w_m / G * sum_g || dt * sum_{i in g} a_i ||^2
#include "common.cuh"

namespace cgnn {
namespace {

constexpr int LOSS_THREADS = 256;
constexpr int LOSS_CHUNK = 4096;       // nodes per block
constexpr int MAX_OUT = 8;
constexpr int NPART = MAX_OUT + 2;     // per-chunk partial: sum a[c] (c < out_dim), sse_acc, sse_temp

__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    float s = 0.0f;
    if (threadIdx.x == 0)
        for (int i = 0; i < LOSS_THREADS / 32; ++i) s += red[i];
    return s;                           // valid in thread 0
}

__global__ void loss_partials_kernel(const float* __restrict__ acc, const float* __restrict__ temp,
                                     const float* __restrict__ y_acc, const float* __restrict__ y_temp,
                                     const int32_t* __restrict__ ptr, int64_t n, int out_dim, int chunks_per_graph,
                                     float* __restrict__ partials) {
    __shared__ float red[LOSS_THREADS / 32];
    const int g = blockIdx.y, cx = blockIdx.x;
    int64_t g0 = ptr ? ptr[g] : 0, g1 = ptr ? ptr[g + 1] : n;
    int64_t i0 = g0 + (int64_t)cx * LOSS_CHUNK;
    int64_t i1 = i0 + LOSS_CHUNK < g1 ? i0 + LOSS_CHUNK : g1;
    float sa[MAX_OUT];
#pragma unroll
    for (int c = 0; c < MAX_OUT; ++c) sa[c] = 0.0f;
    float se_a = 0.0f, se_t = 0.0f;
    for (int64_t i = i0 + threadIdx.x; i < i1; i += LOSS_THREADS) {
#pragma unroll
        for (int c = 0; c < MAX_OUT; ++c)
            if (c < out_dim) {
                float a = acc[i * out_dim + c];
                float d = a - y_acc[i * out_dim + c];
                sa[c] += a;
                se_a = fmaf(d, d, se_a);
            }
        float d = temp[i] - y_temp[i];
        se_t = fmaf(d, d, se_t);
    }
    float* out = partials + ((int64_t)g * chunks_per_graph + cx) * NPART;
#pragma unroll
    for (int c = 0; c < MAX_OUT; ++c) {
        float s = block_sum(sa[c], red);
        if (threadIdx.x == 0) out[c] = s;
    }
    float s = block_sum(se_a, red);
    if (threadIdx.x == 0) out[MAX_OUT] = s;
    s = block_sum(se_t, red);
    if (threadIdx.x == 0) out[MAX_OUT + 1] = s;
}

// one block; thread g reduces graph g's chunks in order, thread 0 finishes
__global__ void loss_finish_kernel(const float* __restrict__ partials, int num_graphs, int chunks_per_graph,
                                   int64_t n, int out_dim, float dt, float w_acc, float w_temp, float w_mom,
                                   float* __restrict__ graph_sums /* [G][MAX_OUT+2] */, float* __restrict__ losses) {
    for (int g = threadIdx.x; g < num_graphs; g += blockDim.x) {
        float s[NPART];
        for (int c = 0; c < NPART; ++c) s[c] = 0.0f;
        for (int cx = 0; cx < chunks_per_graph; ++cx)
            for (int c = 0; c < NPART; ++c) s[c] += partials[((int64_t)g * chunks_per_graph + cx) * NPART + c];
        for (int c = 0; c < MAX_OUT; ++c) s[c] *= dt;                 // velocity change summed over the graph
        for (int c = 0; c < NPART; ++c) graph_sums[g * NPART + c] = s[c];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float sse_a = 0.0f, sse_t = 0.0f, mom = 0.0f;
        for (int g = 0; g < num_graphs; ++g) {
            sse_a += graph_sums[g * NPART + MAX_OUT];
            sse_t += graph_sums[g * NPART + MAX_OUT + 1];
            for (int c = 0; c < out_dim; ++c) { float v = graph_sums[g * NPART + c]; mom = fmaf(v, v, mom); }
        }
        float acc_mse = sse_a / ((float)n * (float)out_dim);
        float temp_mse = sse_t / (float)n;
        float mom_loss = w_mom * mom / (float)num_graphs;
        losses[0] = w_acc * acc_mse + w_temp * temp_mse + mom_loss;
        losses[1] = acc_mse;
        losses[2] = temp_mse;
        losses[3] = mom_loss;
    }
}

__global__ void loss_grad_kernel(const float* __restrict__ acc, const float* __restrict__ temp,
                                 const float* __restrict__ y_acc, const float* __restrict__ y_temp,
                                 const int32_t* __restrict__ ptr, int64_t n, int out_dim, int num_graphs, float dt,
                                 float w_acc, float w_temp, float w_mom, const float* __restrict__ graph_sums,
                                 float* __restrict__ d_acc, float* __restrict__ d_temp) {
    const int g = blockIdx.y;
    int64_t g0 = ptr ? ptr[g] : 0, g1 = ptr ? ptr[g + 1] : n;
    int64_t i = g0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= g1) return;
    const float ca = 2.0f * w_acc / ((float)n * (float)out_dim);
    const float cm = 2.0f * w_mom * dt / (float)num_graphs;
    if (d_acc)
        for (int c = 0; c < out_dim; ++c)
            d_acc[i * out_dim + c] = ca * (acc[i * out_dim + c] - y_acc[i * out_dim + c]) + cm * graph_sums[g * NPART + c];
    if (d_temp) d_temp[i] = 2.0f * w_temp / (float)n * (temp[i] - y_temp[i]);
}

}  // namespace
}  // namespace cgnn

using namespace cgnn;

static int loss_chunks(int64_t n) { return (int)((n + LOSS_CHUNK - 1) / LOSS_CHUNK); }

extern "C" int64_t cgnn_loss_workspace_bytes(int64_t n, int32_t num_graphs) {
    // every graph may be as large as n
    Carver c(nullptr);
    c.take<float>((int64_t)num_graphs * loss_chunks(n) * NPART);
    c.take<float>((int64_t)num_graphs * NPART);
    return c.off;
}

extern "C" int cgnn_loss_fwd_bwd(const float* acc, const float* temp, const float* y_acc, const float* y_temp,
                                 const int32_t* graph_ptr, int64_t n, int32_t out_dim, int32_t num_graphs, float dt,
                                 float w_acc, float w_temp, float w_mom, float* losses, float* d_acc,
                                 float* d_temp, void* workspace, int64_t workspace_bytes, cgnn_stream stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CGNN_CHECK_ARG(acc && temp && y_acc && y_temp && losses && workspace, "cgnn_loss_fwd_bwd: null pointer");
    CGNN_CHECK_ARG(n >= 1 && out_dim >= 1 && out_dim <= MAX_OUT, "cgnn_loss_fwd_bwd: need n >= 1, 1 <= out_dim <= %d", MAX_OUT);
    CGNN_CHECK_ARG(num_graphs >= 1 && num_graphs <= 65535, "cgnn_loss_fwd_bwd: bad num_graphs");
    CGNN_CHECK_ARG(num_graphs == 1 || graph_ptr != nullptr, "cgnn_loss_fwd_bwd: graph_ptr required for num_graphs > 1");
    if (workspace_bytes < cgnn_loss_workspace_bytes(n, num_graphs)) {
        set_error("cgnn_loss_fwd_bwd: workspace too small");
        return CGNN_ERR_WORKSPACE;
    }
    int cpg = loss_chunks(n);
    Carver c(workspace);
    float* partials = c.take<float>((int64_t)num_graphs * cpg * NPART);
    float* graph_sums = c.take<float>((int64_t)num_graphs * NPART);
    loss_partials_kernel<<<dim3(cpg, num_graphs), LOSS_THREADS, 0, stream>>>(acc, temp, y_acc, y_temp, graph_ptr, n,
                                                                              out_dim, cpg, partials);
    CGNN_LAUNCH_CHECK();
    loss_finish_kernel<<<1, 256, 0, stream>>>(partials, num_graphs, cpg, n, out_dim, dt, w_acc, w_temp, w_mom,
                                               graph_sums, losses);
    CGNN_LAUNCH_CHECK();
    if (d_acc || d_temp) {
        loss_grad_kernel<<<dim3((unsigned)((n + 255) / 256), num_graphs), 256, 0, stream>>>(
            acc, temp, y_acc, y_temp, graph_ptr, n, out_dim, num_graphs, dt, w_acc, w_temp, w_mom, graph_sums, d_acc, d_temp);
        CGNN_LAUNCH_CHECK();
    }
    return CGNN_OK;
}
