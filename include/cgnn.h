/* cgnn.h -- C ABI of libcgnn.so: the B200 (sm_100a) Interaction-Network hot path.
 *
 * The reference (mattpan-peregrinus/Cosmology_GNN_Simulation) has no FFI of its own: its
 * boundary is two Python modules.  Each entry point below names the reference lines it
 * replaces.  All functions
 *   - are extern "C", take plain device pointers + sizes + a cudaStream_t (passed as void*),
 *   - return 0 on success or a negative cgnn_status; cgnn_last_error() gives the message,
 *   - never allocate device memory: the caller owns every buffer, including workspaces whose
 *     size is given by the matching *_workspace_bytes() query,
 *   - are asynchronous on `stream` and deterministic (no floating-point atomics).
 *
 * Layouts: all matrices row-major and contiguous; node latents h[N][L], edge latents e[E][L]
 * with E = N*k and edge id = receiver*k + rank (the reference's k-NN always yields this
 * receiver-sorted fixed-in-degree layout, data_utils.py:149-152); weights are torch
 * nn.Linear layout W[out][in]; indices int32 on the device.
 */
#ifndef CGNN_H
#define CGNN_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    CGNN_OK = 0,
    CGNN_ERR_INVALID = -1,     /* bad argument (shape, null pointer, unsupported size) */
    CGNN_ERR_CUDA = -2,        /* a CUDA runtime call or launch failed */
    CGNN_ERR_WORKSPACE = -3,   /* workspace too small */
    CGNN_ERR_UNSUPPORTED = -4  /* configuration not implemented by this build */
} cgnn_status;

typedef void* cgnn_stream;    /* cudaStream_t */

#define CGNN_MAX_LAYERS 4      /* Linear layers per MLP = mlp_num_hidden_layers + 1 */

/* One `build_mlp` (+ optional LayerNorm) of the reference: graph_network.py:15-32,133-135. */
typedef struct {
    int32_t n_layers;                   /* number of Linear layers (>= 1) */
    int32_t in_dim, hidden, out_dim;    /* widths: in -> hidden x (n_layers-1) -> out */
    const float* W[CGNN_MAX_LAYERS];    /* W[l]: [out_l][in_l] */
    const float* b[CGNN_MAX_LAYERS];    /* b[l]: [out_l] */
    const float* ln_gamma;              /* [out_dim] or NULL: no LayerNorm (decoders) */
    const float* ln_beta;
    int32_t ln_dim;                     /* 0 = out_dim.  Tensor-core precisions only: the LayerNorm spans the first ln_dim of
                                         * out_dim = 128 outputs, the rest is zero padding (weights, biases, gamma, beta of the
                                         * padded rows are zero) -- how latent / hidden widths below 128 run on the 128-wide
                                         * tcgen05 tiles (graph_network.py zero-pads the parameters, README.md:59-62 shapes) */
} cgnn_mlp;

/* Gradient destinations, same shapes as cgnn_mlp; written (=), not accumulated. */
typedef struct {
    float* W[CGNN_MAX_LAYERS];
    float* b[CGNN_MAX_LAYERS];
    float* ln_gamma;
    float* ln_beta;
} cgnn_mlp_grad;

typedef enum { CGNN_MSG_SENDER = 0, CGNN_MSG_EDGE = 1 } cgnn_message;
typedef enum { CGNN_DISP_RAW = 0, CGNN_DISP_MIN_IMAGE = 1 } cgnn_disp_mode;
typedef enum {
    CGNN_PREC_FP32 = 0,        /* FP32 SIMT FMA everywhere (<= 1e-5 parity mode) */
    CGNN_PREC_BF16X3 = 1,      /* tcgen05 tensor cores, bf16 hi/lo split operands (3 MMAs), FP32 accum/storage */
    CGNN_PREC_BF16 = 2,        /* tcgen05, single bf16 pass (fastest; ~1e-2, outside the parity bar) */
    CGNN_PREC_BF16X3_G16 = 3   /* CGNN_PREC_BF16X3 with a 2-byte gradient stream, meant for LONG row streams (the edges of a large graph): in
                                * cgnn_mp_edge_bwd and in cgnn_mlp_rows_bwd of an MLP with LayerNorm
                                *   - `de_next` / `de` (edge phase) and `dout` (rows) are bfloat16 [rows][128] arrays behind the float pointers
                                *     (round to nearest even; `de` may still alias `de_next`),
                                *   - the three gradient intermediates dY, G2, G1 of the workspace are bfloat16 as well,
                                *   - the call always runs the layered one-launch-per-GEMM composition.
                                * Forward values and ReLU gates are those of CGNN_PREC_BF16X3; the rounding of the gradient stream averages
                                * out in the weight gradients (error study on the oracle: tests/study_grad_stream.py).  MLPs without
                                * LayerNorm and every other entry point treat it as CGNN_PREC_BF16X3. */
    /* The tensor-core modes cover the processor phases (cgnn_mp_edge_* / cgnn_mp_node_*) for
     * latent = hidden = 128 with 2 hidden layers and k a power of two <= 32; other shapes return
     * CGNN_ERR_UNSUPPORTED.  The row-wise encoder / decoder entry points use the tensor cores for 3-layer
     * MLPs with hidden = 128 and in, out <= 128 (zero padded), and run their FP32 kernels for any other
     * shape or when an input gradient narrower than 128 columns is requested. */
} cgnn_precision;

const char* cgnn_last_error(void);
const char* cgnn_version(void);
/* number of kernels launched by this library in this process so far (for bench `gpu_launches`) */
int64_t cgnn_launch_count(void);
/* debug / profiling hook (tools/chain_stamps.py): `stamps` is a device buffer of launches * 2 * tiles * 16 uint64;
 * the next `launches` tensor-core chain launches record clock64() of the stages of block 0's first `tiles` tiles
 * per epilogue group into consecutive slices of it.  NULL switches the recording off (the default).  Not used
 * by the product path. */
void cgnn_debug_stamps(unsigned long long* stamps, int32_t tiles, int32_t launches);

/* ---------------------------------------------------------------------------------------------
 * K1  periodic k-NN  -- replaces extend_positions_torch + torch_cluster.knn + index remap,
 *     data_utils.py:148-152 (and :9-33).
 * pos[N][3] fp32 in [0, box] (closed).  nbr_ext[N][k] int32 receives, for each query i, the k
 * nearest of the 27N ghost-extended candidates c = s*N + j (s = shift index, x slowest), sorted
 * ascending under the total order (d2, c), d2 the fp32 distance of SURVEY App. A.2.
 * Requires 1 <= k <= 32 and 27*N >= k.
 */
int64_t cgnn_knn_workspace_bytes(int64_t n);
int cgnn_knn_periodic(const float* pos, int64_t n, float box, int32_t k, int32_t* nbr_ext,
                      void* workspace, int64_t workspace_bytes, cgnn_stream stream);
/* slab sharding: all N particles are candidates, only the queries q0 <= i < q0 + nq are answered
 * (nbr_ext[nq][k]); a rank owns a contiguous index range of the x-sorted particles */
int cgnn_knn_periodic_range(const float* pos, int64_t n, float box, int32_t k, int64_t q0, int64_t nq,
                            int32_t* nbr_ext, void* workspace, int64_t workspace_bytes, cgnn_stream stream);

/* ---------------------------------------------------------------------------------------------
 * K2  graph products -- replaces data_utils.py:150-164.
 * From nbr_ext: senders[E] int32 (= c mod N); optional edge_index[2][E] int64
 * (row 0 sender, row 1 receiver; pass NULL to skip); edge_attr[E][4] = (dx,dy,dz,|d|) with
 * d = pos[sender] - pos[receiver] (RAW, the reference's behaviour) or the minimum image
 * d = fl(pos[sender] + shift_s) - pos[receiver].
 */
int cgnn_edge_features(const float* pos, const int32_t* nbr_ext, int64_t n, int32_t k, float box,
                       int32_t disp_mode, int32_t* senders, int64_t* edge_index, float* edge_attr,
                       cgnn_stream stream);

/* the same for the receivers q0 <= i < q0 + nq only (nbr_ext[nq][k]); senders / edge_index hold GLOBAL ids */
int cgnn_edge_features_range(const float* pos, const int32_t* nbr_ext, int64_t n, int32_t k, float box,
                             int32_t disp_mode, int64_t q0, int64_t nq, int32_t* senders, int64_t* edge_index,
                             float* edge_attr, cgnn_stream stream);

/* Node features and normalised targets of one sample in one launch -- data_utils.py:86-145,166-214: wrap the W frames
 * into the box (torch.remainder), minimum-image frame differences / dt, normalise, flatten (velocities time-major then
 * xyz, then the W temperatures) into x[N][3(W-1)+W]; recent_pos[N][3] = the wrapped last frame (the k-NN input);
 * optional targets -> y_acc[N][3], y_temp[N].  Single IEEE float32 operations in the reference's order: bit-identical
 * to the reference's CPU tensor arithmetic.  pos_seq[W][N][3], temp_seq[W][N] time-major; pos_noise[N][W][3] /
 * temp_noise[N][W] are the random-walk noise of data_utils.py:36-70 (NULL = none; the draws stay with the caller's
 * torch generator); stats[8] = vel_mean, vel_std, temp_mean, temp_std, acc_mean, acc_std, temp_rate_mean,
 * temp_rate_std (host array, the metadata JSON of generate_metadata.py:32-43 rounded to float32). */
int cgnn_preprocess_features(const float* pos_seq, const float* temp_seq, const float* pos_noise, const float* temp_noise,
                             const float* target_pos, const float* target_temp, int64_t n, int32_t window, float box,
                             float dt, const float* stats, float* recent_pos, float* x, float* y_acc, float* y_temp,
                             cgnn_stream stream);

/* Sender-sorted transpose of the receiver-sorted graph (needed for the deterministic d/dh[sender]):
 * rowptr[N+1], perm[E] = edge ids grouped by sender, ascending inside each group. */
int64_t cgnn_csr_transpose_workspace_bytes(int64_t n, int64_t n_edges);
int cgnn_csr_transpose(const int32_t* senders, int64_t n, int64_t n_edges, int32_t* rowptr,
                       int32_t* perm, void* workspace, int64_t workspace_bytes, cgnn_stream stream);

/* Checks that an edge_index[2][E] int64 has the receiver-sorted fixed-in-degree layout
 * (row 1 == e / k, senders in [0,N)) and writes senders as int32.  *bad_flag (device int32) is
 * set to 0, then to 1 if the layout does not hold. */
int cgnn_edge_index_to_senders(const int64_t* edge_index, int64_t n, int32_t k, int32_t* senders,
                               int32_t* bad_flag, cgnn_stream stream);

/* ---------------------------------------------------------------------------------------------
 * K3/K6  row-wise MLP (+LayerNorm) -- GraphIndependent.forward (graph_network.py:52-64) and the
 * decoders (graph_network.py:151-152,158-159).   out[r] = [LN](MLP(x[r])).
 */
/* workspace of cgnn_mlp_rows_fwd (backward = 0) / cgnn_mlp_rows_bwd (backward = 1) */
int64_t cgnn_mlp_rows_workspace_bytes(const cgnn_mlp* mlp, int64_t rows, int32_t precision, int32_t backward);
int cgnn_mlp_rows_fwd(const cgnn_mlp* mlp, const float* x, int64_t rows, float* out, void* workspace,
                      int64_t workspace_bytes, int32_t precision, cgnn_stream stream);
/* dx may be NULL.  Parameter gradients are written to `grad`. */
int64_t cgnn_mlp_bwd_workspace_bytes(const cgnn_mlp* mlp);   /* FP32 kernels only; prefer the *_workspace_bytes above */
int cgnn_mlp_rows_bwd(const cgnn_mlp* mlp, const cgnn_mlp_grad* grad, const float* x, int64_t rows,
                      const float* dout, float* dx, void* workspace, int64_t workspace_bytes,
                      int32_t precision, cgnn_stream stream);

/* ---------------------------------------------------------------------------------------------
 * K4  one message-passing step, forward -- InteractionNetwork.forward + the residuals,
 *     graph_network.py:83-101,177-183.
 *  edge phase:  u_e = LN(MLP_e([h[s_e] | h[r_e] | e]));  e_out = e + u_e  (e_out may alias e_in: in place; NULL
 *               skips the store -- the last step's e' is never read, graph_network.py:177-183);
 *               agg_edge[i] = sum over the k in-edges of i of u_e (rank order), or NULL to skip
 *  sender aggregation (reference-actual message): agg[i] = sum_r h[senders[i*k+r]]
 *  node phase:  u_n = LN(MLP_n([h | agg]));  h_out = h + u_n
 */
/* workspace of cgnn_mp_edge_fwd (0 for CGNN_PREC_FP32; the tensor-core modes stage split-bf16 weight
 * images and the per-node layer-1 partial products there) */
/* Slab sharding: a rank's node array is [n owned receivers | halo senders], n_nodes rows in all; the
 * receivers are rows 0..n-1, `senders` index all n_nodes rows.  Single GPU: n_nodes == n. */
int64_t cgnn_mp_edge_fwd_workspace_bytes(const cgnn_mlp* edge_mlp, int64_t n_nodes, int32_t precision);
/* k_valid (tensor-core precisions; 0 or k = off): the chain needs k to be a power of two, so a graph of in-degree
 * k_valid < k is padded per receiver to k rows -- rows of rank >= k_valid are dummies: they take no part in the
 * per-receiver sums and get / give no gradient (graph_network.py pads senders and edge features; README.md:59-62
 * allows 8..32 neighbours). */
int cgnn_mp_edge_fwd(const cgnn_mlp* edge_mlp, const float* h, const float* e_in,
                     const int32_t* senders, int64_t n, int64_t n_nodes, int32_t k, int32_t k_valid, float* e_out,
                     float* agg_edge, void* workspace, int64_t workspace_bytes, int32_t precision,
                     cgnn_stream stream);
int cgnn_aggregate_senders(const float* h, const int32_t* senders, int64_t n, int32_t k,
                           int32_t latent, float* agg, cgnn_stream stream);
/* Halo exchange of a slab-sharded box (the reference has no distributed code): the row copies either side of the
 * transport (NCCL point-to-point between the ranks of one node).
 *   cgnn_halo_pack:        dst[i] = src[idx[i]]      the owned rows one peer holds as halo, gathered into the send buffer
 *   cgnn_halo_unpack_add:  dst[idx[i]] += src[i]     the gradients a peer accumulated on those copies, added on the owner;
 *                                                    idx holds no duplicates, peers are applied in ascending rank order by
 *                                                    the caller: deterministic
 * Rows are `latent` floats (a multiple of 4), idx are int64 local row ids. */
int cgnn_halo_pack(const float* src, const int64_t* idx, int64_t n_idx, int32_t latent, float* dst, cgnn_stream stream);
int cgnn_halo_unpack_add(const float* src, const int64_t* idx, int64_t n_idx, int32_t latent, float* dst, cgnn_stream stream);
int64_t cgnn_mp_node_fwd_workspace_bytes(const cgnn_mlp* node_mlp, int64_t n, int32_t precision);
int cgnn_mp_node_fwd(const cgnn_mlp* node_mlp, const float* h, const float* agg, int64_t n,
                     float* h_out, void* workspace, int64_t workspace_bytes, int32_t precision,
                     cgnn_stream stream);

/* ---------------------------------------------------------------------------------------------
 * K5  one message-passing step, backward (activations recomputed inside the tile).
 *  node phase: given dh_next = dL/dh^{t+1}:  dh = dh_next + dIn[:, :L];  dagg = dIn[:, L:]
 *  edge phase (message = edge): given de_next = dL/de^{t+1} (NULL = zero) and dagg:
 *      dU = de_next + dagg[receiver];  de = de_next + dIn[:, 2L:];
 *      dh[receiver] += sum_rank dIn[:, L:2L];  dh[sender] += dIn[:, :L] summed over the sender-sorted
 *      transpose (t_rowptr, t_perm from cgnn_csr_transpose; fixed order, no atomics).  `de` may alias
 *      `de_next` (the gradient stream is updated in place).  gs[E][L] is scratch for the per-edge sender
 *      gradient of the FP32 kernels; the tensor-core precisions accumulate it chunk by chunk inside their
 *      workspace and take gs = NULL (no E-sized scratch: what lets 2 M particles, k = 32 fit one GPU).
 *  cgnn_scatter_to_senders: dh[j] += sum over edges e with sender j (perm order) of src[e] (stride L)
 *      or of src[e / k] (src_is_per_receiver: message = sender, where src = dagg).
 */
/* workspace of cgnn_mp_node_bwd (k = 0, n_nodes = n) / cgnn_mp_edge_bwd (k = in-degree) */
int64_t cgnn_mp_bwd_workspace_bytes(const cgnn_mlp* mlp, int64_t n, int64_t n_nodes, int32_t k, int32_t precision);
int cgnn_mp_node_bwd(const cgnn_mlp* node_mlp, const cgnn_mlp_grad* grad, const float* h,
                     const float* agg, const float* dh_next, int64_t n, float* dh, float* dagg,
                     void* workspace, int64_t workspace_bytes, int32_t precision, cgnn_stream stream);
int cgnn_mp_edge_bwd(const cgnn_mlp* edge_mlp, const cgnn_mlp_grad* grad, const float* h,
                     const float* e_in, const int32_t* senders, const int32_t* t_rowptr,
                     const int32_t* t_perm, int64_t n, int64_t n_nodes, int32_t k, int32_t k_valid,
                     const float* de_next, const float* dagg, float* de, float* dh, float* gs,
                     void* workspace, int64_t workspace_bytes, int32_t precision, cgnn_stream stream);
int cgnn_scatter_to_senders(const float* src, int32_t src_is_per_receiver, const int32_t* rowptr,
                            const int32_t* perm, int64_t n, int32_t k, int32_t latent, float* dh,
                            cgnn_stream stream);

/* ---------------------------------------------------------------------------------------------
 * K6  loss -- train.py:107-118,255-260.  Deterministic two-stage reductions.
 *  losses[4] = {total, acc_mse, temp_mse, momentum};  writes the gradient seeds d_acc, d_temp
 *  (of `total`) when non-NULL.  graph_ptr[G+1] int32 = node offsets of the batched graphs
 *  (PyG `Batch.ptr`; NULL = one graph).
 */
int64_t cgnn_loss_workspace_bytes(int64_t n, int32_t num_graphs);
int cgnn_loss_fwd_bwd(const float* acc, const float* temp, const float* y_acc, const float* y_temp,
                      const int32_t* graph_ptr, int64_t n, int32_t out_dim, int32_t num_graphs, float dt,
                      float w_acc, float w_temp, float w_mom, float* losses, float* d_acc,
                      float* d_temp, void* workspace, int64_t workspace_bytes, cgnn_stream stream);

/* ---------------------------------------------------------------------------------------------
 * Optimizer step -- train.py:183-187 (torch.optim.Adam(lr, weight_decay) + ExponentialLR), :263-265.
 * One fused pass over a flat FP32 parameter buffer and its gradient / moment buffers (all n floats, 16-byte
 * aligned):  g = grads * grad_scale + weight_decay * p;  m, v <- Adam moments;  p -= lr / (1 - beta1^step) *
 * m / (sqrt(v) / sqrt(1 - beta2^step) + eps).  `lr` is the scheduler's current rate (lr0 * gamma^epoch),
 * `step` counts from 1, `grad_scale` folds an all-reduce average (1 / world) into the same pass.
 */
int cgnn_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                   float beta1, float beta2, float eps, float weight_decay, int32_t step, float grad_scale,
                   cgnn_stream stream);
/* The same with a mask `live[n]` (NULL = all live): entries whose mask is 0 are left untouched -- parameter, both
 * moments, no weight decay -- as torch.optim.Adam does for parameters whose .grad is None (the dead edge stream
 * of the reference's actual message semantics).  Works on any 16-byte aligned slice of the flat buffers, which is
 * what lets the step of one gradient bucket run while the next bucket is still being all-reduced. */
int cgnn_adam_step_masked(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, const float* live,
                          int64_t n, float lr, float beta1, float beta2, float eps, float weight_decay, int32_t step,
                          float grad_scale, cgnn_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* CGNN_H */
