// Task descriptors shared by the SIMT (FP32) and tensor-core implementations of the MLP tiles.
#pragma once
#include "common.cuh"

namespace cgnn {

enum MlpMode { MODE_ROWS = 0, MODE_EDGE = 1, MODE_NODE = 2 };

struct MlpDev {
    int n_layers, in_dim, hidden, out_dim;
    const float* W[CGNN_MAX_LAYERS];
    const float* b[CGNN_MAX_LAYERS];
    const float* gamma;
    const float* beta;
    int ln_dim;                 // LayerNorm width (<= out_dim; the rest is zero padding), see cgnn_mlp
};

struct GradPtrs {
    float* W[CGNN_MAX_LAYERS];
    float* b[CGNN_MAX_LAYERS];
    float* gamma;
    float* beta;
};

// Everything one fused MLP launch needs (passed by value to the kernels).
struct MlpTask {
    MlpDev mlp;
    int mode;
    int64_t n;                  // ROWS: rows; EDGE/NODE: number of nodes (receivers)
    int64_t n_nodes;            // EDGE: rows of h / dh (receivers first, then halo senders); == n on one GPU
    int k, L;                   // EDGE: in-degree; EDGE/NODE: latent width
    int k_valid;                // EDGE, tensor-core precisions: real in-degree when k is padded to a power of two (0: k itself)
    int act_stride;             // shared-memory row stride of the activation buffers
    // forward inputs
    const float* x;             // ROWS  [rows][in_dim]
    const float* h;             // EDGE/NODE [N][L]
    const float* e_in;          // EDGE  [E][L]
    const float* agg;           // NODE  [N][L]
    const int32_t* senders;     // EDGE  [E]
    const int32_t* t_rowptr;    // EDGE backward: sender-sorted transpose (rowptr [N+1], perm [E])
    const int32_t* t_perm;
    // forward outputs
    float* out;                 // ROWS [rows][out]; EDGE e_out [E][L]; NODE h_out [N][L]
    float* agg_out;             // EDGE optional [N][L]
    // backward
    const float* dout;          // ROWS: dL/dout; NODE: dL/dh_next
    const float* de_next;       // EDGE: dL/de_next (nullable)
    const float* dagg;          // EDGE: dL/dagg [N][L]
    float* dx;                  // ROWS: dL/dx (nullable)
    float* dh;                  // NODE: written; EDGE: receiver part accumulated
    float* dagg_out;            // NODE: dL/dagg
    float* gs;                  // EDGE: per-edge sender gradient [E][L]
    float* de;                  // EDGE: dL/de [E][L]
    int need_input_grad;
    int64_t w_off[CGNN_MAX_LAYERS], b_off[CGNN_MAX_LAYERS], g_off, blob_size;
};

int mlp_validate(const cgnn_mlp* mlp, const char* who);
MlpDev mlp_to_dev(const cgnn_mlp* mlp);

int simt_mlp_fwd(MlpTask& a, cudaStream_t s);
int simt_mlp_bwd(MlpTask& a, const cgnn_mlp_grad* g, void* ws, int64_t wsb, cudaStream_t s);
int64_t simt_mlp_bwd_workspace(const cgnn_mlp* mlp);
int simt_aggregate_senders(const float* h, const int32_t* senders, int64_t n, int k, int L, float* agg, cudaStream_t s);
int simt_scatter_to_senders(const float* src, int per_receiver, const int32_t* rowptr, const int32_t* perm,
                            int64_t n, int k, int L, float* dh, cudaStream_t s);

}  // namespace cgnn
