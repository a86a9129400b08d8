// Host-side composition of the tensor-core kernels into the processor phases (forward and backward).
//
// Forward (one launch per phase, fully fused):        mp_tc.cu   tc_chain_fwd, 3-layer chains
// Backward (round 1: tensor-core but not yet fused):  every GEMM of the recompute / dgrad / wgrad chain is
//   a tensor-core launch over FP32 row streams held in a bounded workspace (edge rows are processed in
//   chunks, so the transient activations never exceed CHUNK_ROWS x 5 buffers):
//     recompute   A1 = relu(L1(in)), A2 = relu(L2(A1)), Y = L3(A2)          1-layer chains
//     LayerNorm   dY = LNbwd(Y, dU)  (+ d gamma, d beta)                    ln_bwd_kernel
//     dgrad       G2 = (dY W3) * [A2>0],  G1 = (G2 W2) * [A1>0],  dIn = G1 W1   1-layer chains, transposed blocks
//     wgrad       dW3 = dY^T A2, dW2 = G2^T A1, dW1 = G1^T in  (+ bias columns)  tc_wgrad_kernel
//   With W1 = [W1s | W1r | W1e] the per-edge layer 1 only sees e; the sender / receiver parts act on the
//   per-node sums of G1 (sender-sorted transpose / k consecutive rows).
//
// Reference semantics: graph_network.py:83-101,177-183 and its autograd (train.py:264).
#include <stdlib.h>

#include "tc_common.cuh"

namespace cgnn {
namespace {

constexpr int64_t CHUNK_ROWS = 1 << 21;      // edge rows per backward chunk (1 GiB per FP32 buffer)

bool tc_k_ok(int k) { return k >= 1 && k <= 32 && (k & (k - 1)) == 0; }
bool tc_shape_ok(const MlpDev& m, int in_mult) {
    return m.n_layers == 3 && m.hidden == TC_H && m.out_dim == TC_H && m.in_dim == in_mult * TC_H && m.gamma != nullptr;
}

int64_t rows_bytes(int64_t rows) { return align_up(rows * TC_H * 4, 256); }
int64_t gate_bytes(int64_t rows) { return align_up(rows * 16, 256); }      // ReLU gates: 128 bits per row

struct Scratch {
    void* wg; void* lnb;
    void carve(Carver& cv) {
        wg = cv.take<uint8_t>(wgrad_workspace_bytes());
        lnb = cv.take<uint8_t>(ln_bwd_workspace_bytes());
    }
    static int64_t bytes() { return wgrad_workspace_bytes() + ln_bwd_workspace_bytes(); }
};

ChainOp base_op(int ns, const Scratch& sc, int64_t rows) {
    ChainOp op{};
    op.ns = ns; op.rows = rows; op.n_layers = 1;
    return op;
}

// P_s = h W1[:, 0:L]^T,  P_r = h W1[:, L:2L]^T + b1   (graph_network.py:89: concat order sender, receiver, edge)
// P_s covers all n_nodes rows of h (receivers + halo senders), P_r the n receivers (the first n rows)
int project_nodes(int ns, const Scratch& sc, const MlpDev& m, const float* h, int64_t n, int64_t n_nodes, float* Ps, float* Pr,
                  cudaStream_t s) {
    for (int which = 0; which < 2; ++which) {
        ChainOp op = base_op(ns, sc, which == 0 ? n_nodes : n);
        op.in0 = h;
        op.blk[0] = {m.W[0], 3 * TC_H, 0, which * TC_H, 0};
        op.bias[0] = which == 1 ? m.b[0] : nullptr;
        op.out = which == 0 ? Ps : Pr;
        int rc = run_chain(op, s);
        if (rc) return rc;
    }
    return CGNN_OK;
}

// [rows][w] -> [rows][128], zero padded  /  [rows][128] -> [rows][w]
__global__ void pad_rows_kernel(const float* __restrict__ src, int w, int64_t rows, float* __restrict__ dst) {
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx >= rows * TC_H) return;
    const int64_t r = idx / TC_H;
    const int c = (int)(idx - r * TC_H);
    dst[idx] = c < w ? src[r * w + c] : 0.0f;
}
__global__ void slice_rows_kernel(const float* __restrict__ src, int w, int64_t rows, float* __restrict__ dst) {
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx >= rows * w) return;
    const int64_t r = idx / w;
    const int c = (int)(idx - r * w);
    dst[idx] = src[r * TC_H + c];
}
int pad_rows(const float* src, int w, int64_t rows, float* dst, cudaStream_t s) {
    pad_rows_kernel<<<(unsigned)((rows * TC_H + 255) / 256), 256, 0, s>>>(src, w, rows, dst);
    CGNN_LAUNCH_CHECK();
    return CGNN_OK;
}
int slice_rows(const float* src, int w, int64_t rows, float* dst, cudaStream_t s) {
    slice_rows_kernel<<<(unsigned)((rows * w + 255) / 256), 256, 0, s>>>(src, w, rows, dst);
    CGNN_LAUNCH_CHECK();
    return CGNN_OK;
}

// ---- per-node sums of G1 by SENDER, chunk by chunk ------------------------------------------------------------------------
// dPs[j] += sum of G1[e - r0] over the out-edges e of sender j with r0 <= e < r1, in perm order (= ascending edge id, so chunk
// after chunk in ascending r0 adds every row's terms in exactly the order of a one-shot pass: deterministic and independent of the
// chunk size).  Which nodes have out-edges inside a chunk is worked out ONCE per call (round 2 scanned every node's list in every
// chunk: 574 us per 2 Mi-row chunk at 2.1 M particles, 7 % of the training step, for 1 GiB of useful reads): a node's out-edge
// list is ascending, so one walk over it yields its (chunk, first position) pairs; they are counted, offset and filled into
// per-chunk entry lists (the order of the entries inside a list is whatever the atomics give -- every entry owns its dPs row, so
// the sums do not depend on it), and a chunk's kernel visits its own entries only, a warp per entry.
constexpr int SC_WARPS = 8, SC_NODES = 64;          // a block of 8 warps walks 64 consecutive nodes
// Counts (ent == nullptr) or fills the per-chunk entry lists.  The block first counts its own entries per chunk in shared memory and
// reserves them with ONE global atomic per chunk it touches (consecutive nodes send into the same one or two chunks; a global
// atomic per entry -- four million on 32 counters at 2.1 M particles -- took 1.6 ms per pass), then walks again to place them.
__global__ void __launch_bounds__(SC_WARPS * 32)
sender_chunks_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ perm, int64_t n, int chunk, int n_chunks,
                     int32_t* __restrict__ cnt, const int32_t* __restrict__ off, int2* __restrict__ ent) {
    extern __shared__ int32_t sc_sh[];
    int32_t* lc = sc_sh;                  // entries of this block per chunk
    int32_t* lbase = sc_sh + n_chunks;    // where they start inside the chunk's global range
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < n_chunks; i += blockDim.x) lc[i] = 0;
    __syncthreads();
    for (int pass = 0; pass < (ent != nullptr ? 2 : 1); ++pass) {
        for (int q = warp; q < SC_NODES; q += SC_WARPS) {
            // a warp per node, 32 list positions at a time (coalesced reads of perm): a position opens an entry when its chunk
            // differs from its predecessor's
            const int64_t j = (int64_t)blockIdx.x * SC_NODES + q;
            if (j >= n) break;
            int last = -1;                                    // chunk of the position before this batch
            for (int base = rowptr[j], b = rowptr[j + 1]; base < b; base += 32) {
                const int p = base + lane;
                const int c = p < b ? perm[p] / chunk : -2;
                int before = __shfl_up_sync(0xFFFFFFFFu, c, 1);
                if (lane == 0) before = last;
                if (p < b && c != before) {
                    const int local = atomicAdd(lc + c, 1);
                    if (pass == 1) ent[off[c] + lbase[c] + local] = make_int2((int)j, p);
                }
                last = __shfl_sync(0xFFFFFFFFu, c, 31);
            }
        }
        __syncthreads();
        if (pass == 0) {
            for (int i = threadIdx.x; i < n_chunks; i += blockDim.x) {
                const int m = lc[i];
                if (m > 0) lbase[i] = atomicAdd(cnt + i, m);
                lc[i] = 0;
            }
            __syncthreads();
        }
    }
}
// off[c] = entries of the chunks before c (exclusive scan; at most a few thousand chunks), counters back to zero for the fill pass
__global__ void sender_chunks_offsets_kernel(int32_t* __restrict__ cnt, int32_t* __restrict__ off, int n_chunks) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    int acc = 0;
    for (int c = 0; c < n_chunks; ++c) {
        off[c] = acc;
        acc += cnt[c];
        cnt[c] = 0;
    }
    off[n_chunks] = acc;
}
constexpr int SCATTER_WARPS = 8;
// G16: G1 is the bfloat16 gradient stream (8 bytes per lane and row instead of 16).  lane <-> 4 columns.
template <bool G16>
__global__ void __launch_bounds__(SCATTER_WARPS * 32)
scatter_chunk_kernel(const void* __restrict__ G1v, int r0, int r1, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ perm,
                     const int2* __restrict__ ent, const int32_t* __restrict__ off, float4* __restrict__ dPs) {
    const int lane = threadIdx.x & 31;
    const int first = off[0], count = off[1] - first;
    for (int i = blockIdx.x * SCATTER_WARPS + (threadIdx.x >> 5); i < count; i += gridDim.x * SCATTER_WARPS) {
        const int2 en = ent[first + i];
        const int b = rowptr[en.x + 1];
        float4* dst = dPs + (int64_t)en.x * (TC_H / 4) + lane;
        float4 s = *dst;
        for (int base = en.y; base < b; base += 32) {
            // the list is ascending: the edges of this chunk are a prefix of what is left of it
            const int e = base + lane < b ? perm[base + lane] : 0x7FFFFFFF;
            const int m = __popc(__ballot_sync(0xFFFFFFFFu, e < r1));
#pragma unroll 4
            for (int t = 0; t < m; ++t) {
                const int64_t row = __shfl_sync(0xFFFFFFFFu, e, t) - r0;
                float4 v;
                if (G16) {
                    const uint2 w = reinterpret_cast<const uint2*>(G1v)[row * (TC_H / 4) + lane];
                    v = make_float4(__uint_as_float(w.x << 16), __uint_as_float(w.x & 0xFFFF0000u), __uint_as_float(w.y << 16),
                                    __uint_as_float(w.y & 0xFFFF0000u));
                } else {
                    v = reinterpret_cast<const float4*>(G1v)[row * (TC_H / 4) + lane];
                }
                s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
            }
            if (m < 32) break;
        }
        *dst = s;
    }
}
int64_t scatter_entries(int64_t E, int64_t nn, int64_t chunk) {
    const int64_t n_chunks = (E + chunk - 1) / chunk;
    return E < nn * n_chunks ? E : nn * n_chunks;          // every edge opens at most one entry, every node at most one per chunk
}

// encoder / decoder MLPs: 3 layers, hidden 128, in <= 128, out <= 128 (LayerNorm only with out == 128)
bool tc_rows_ok(const MlpDev& m) {
    return m.n_layers == 3 && m.hidden == TC_H && m.in_dim >= 1 && m.in_dim <= TC_H && m.out_dim >= 1 && m.out_dim <= TC_H &&
           (m.gamma == nullptr || m.out_dim == TC_H);
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// workspace sizes
// ------------------------------------------------------------------------------------------------
int64_t tc_edge_fwd_workspace(const cgnn_mlp* mlp, int64_t n, int precision) {
    (void)mlp; (void)precision;
    return Scratch::bytes() + 2 * rows_bytes(n);
}
int64_t tc_node_fwd_workspace(const cgnn_mlp* mlp, int64_t n, int precision) {
    (void)mlp; (void)n; (void)precision;
    return Scratch::bytes();
}
int64_t tc_rows_workspace(const cgnn_mlp* mlp, int64_t rows, int precision, int backward) {
    (void)mlp; (void)precision;
    const int64_t chunk = rows < CHUNK_ROWS ? rows : CHUNK_ROWS;
    return Scratch::bytes() + (backward ? 6 : 2) * rows_bytes(chunk) + (backward ? 2 * gate_bytes(chunk) : 0);
}
// k == 0: node phase
int64_t tc_bwd_workspace(const cgnn_mlp* mlp, int64_t n, int64_t n_nodes, int k, int precision) {
    (void)mlp; (void)precision;
    if (k == 0) return Scratch::bytes() + 5 * rows_bytes(n) + 2 * gate_bytes(n);
    int64_t chunk = n * k < CHUNK_ROWS ? n * k : CHUNK_ROWS;
    const int64_t n_chunks = (n * k + chunk - 1) / chunk;
    return Scratch::bytes() + 5 * rows_bytes(chunk) + 2 * gate_bytes(chunk) + 2 * rows_bytes(n) + 2 * rows_bytes(n_nodes) +
           align_up(scatter_entries(n * k, n_nodes, chunk) * 8, 256) + 2 * align_up((n_chunks + 1) * 4, 256);
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
int tc_mlp_fwd(MlpTask& a, int precision, void* ws, int64_t wsb, cudaStream_t s) {
    const int ns = precision == CGNN_PREC_BF16X3 || precision == CGNN_PREC_BF16X3_G16 ? 3 : 1;
    const MlpDev& m = a.mlp;
    if (a.mode == MODE_EDGE) {
        if (!tc_shape_ok(m, 3) || !tc_k_ok(a.k)) {
            set_error("tensor-core edge phase supports latent = hidden = 128, 2 hidden layers and k a power of two <= 32 (got in %d hidden %d out %d layers %d k %d)",
                      m.in_dim, m.hidden, m.out_dim, m.n_layers, a.k);
            return CGNN_ERR_UNSUPPORTED;
        }
        const int64_t nn = a.n_nodes > 0 ? a.n_nodes : a.n;
        const int64_t need = tc_edge_fwd_workspace(nullptr, nn, precision);
        if (ws == nullptr || wsb < need) {
            set_error("cgnn_mp_edge_fwd: workspace too small (%lld < %lld)", (long long)wsb, (long long)need);
            return CGNN_ERR_WORKSPACE;
        }
        Carver cv(ws);
        Scratch sc; sc.carve(cv);
        float* Ps = cv.take<float>(nn * TC_H);
        float* Pr = cv.take<float>(nn * TC_H);
        int rc = project_nodes(ns, sc, m, a.h, a.n, nn, Ps, Pr, s);
        if (rc) return rc;
        ChainOp op = base_op(ns, sc, a.n * a.k);
        op.n_layers = 3;
        op.in0 = a.e_in;
        op.blk[0] = {m.W[0], 3 * TC_H, 0, 2 * TC_H, 0};
        op.blk[1] = {m.W[1], TC_H, 0, 0, 0};
        op.blk[2] = {m.W[2], TC_H, 0, 0, 0};
        op.bias[0] = nullptr;                         // b1 is folded into P_r
        op.bias[1] = m.b[1]; op.bias[2] = m.b[2]; op.gamma = m.gamma; op.beta = m.beta; op.ln_n = m.ln_dim;
        op.k = a.k; op.k_valid = a.k_valid; op.senders = a.senders; op.Ps = Ps; op.Pr = Pr; op.ps_rows = nn;
        op.residual = a.e_in; op.agg_out = a.agg_out; op.out = a.out;
        return run_chain(op, s);
    }
    if (a.mode == MODE_NODE) {
        if (!tc_shape_ok(m, 2)) {
            set_error("tensor-core node phase supports latent = hidden = 128 and 2 hidden layers (got in %d hidden %d out %d layers %d)",
                      m.in_dim, m.hidden, m.out_dim, m.n_layers);
            return CGNN_ERR_UNSUPPORTED;
        }
        const int64_t need = tc_node_fwd_workspace(nullptr, a.n, precision);
        if (ws == nullptr || wsb < need) {
            set_error("cgnn_mp_node_fwd: workspace too small (%lld < %lld)", (long long)wsb, (long long)need);
            return CGNN_ERR_WORKSPACE;
        }
        Carver cv(ws);
        Scratch sc; sc.carve(cv);
        ChainOp op = base_op(ns, sc, a.n);
        op.n_layers = 3;
        op.in0 = a.h; op.in1 = a.agg;
        op.blk[0] = {m.W[0], 2 * TC_H, 0, 0, 0};
        op.blk[1] = {m.W[0], 2 * TC_H, 0, TC_H, 0};
        op.blk[2] = {m.W[1], TC_H, 0, 0, 0};
        op.blk[3] = {m.W[2], TC_H, 0, 0, 0};
        op.bias[0] = m.b[0]; op.bias[1] = m.b[1]; op.bias[2] = m.b[2]; op.gamma = m.gamma; op.beta = m.beta; op.ln_n = m.ln_dim;
        op.residual = a.h; op.out = a.out;
        return run_chain(op, s);
    }
    if (a.mode == MODE_ROWS) {
        if (!tc_rows_ok(m)) return CGNN_ERR_UNSUPPORTED;          // the caller runs the FP32 kernels
        const int64_t need = tc_rows_workspace(nullptr, a.n, precision, 0);
        if (ws == nullptr || wsb < need) {
            set_error("cgnn_mlp_rows_fwd: workspace too small (%lld < %lld)", (long long)wsb, (long long)need);
            return CGNN_ERR_WORKSPACE;
        }
        Carver cv(ws);
        Scratch sc; sc.carve(cv);
        const int64_t chunk = a.n < CHUNK_ROWS ? a.n : CHUNK_ROWS;
        float* Xp = cv.take<float>(chunk * TC_H);
        float* O = cv.take<float>(chunk * TC_H);
        int rc;
        for (int64_t r0 = 0; r0 < a.n; r0 += chunk) {
            const int64_t rows = a.n - r0 < chunk ? a.n - r0 : chunk;
            const float* in = a.x + r0 * m.in_dim;
            // narrow inputs whose row pitch the TMA accepts (a multiple of 16 bytes: the 4 edge features) are read in place,
            // zero-filled to 128 columns on the way into shared memory; others (17 node features) take a padded copy
            const bool narrow_tma = m.in_dim < TC_H && (m.in_dim * 4) % 16 == 0 && ((uintptr_t)in & 15) == 0;
            if (m.in_dim < TC_H && !narrow_tma) {
                if ((rc = pad_rows(in, m.in_dim, rows, Xp, s))) return rc;
                in = Xp;
            }
            ChainOp op = base_op(ns, sc, rows);
            op.n_layers = 3;
            op.in0 = in; op.in0_cols = narrow_tma ? m.in_dim : 0;
            op.blk[0] = {m.W[0], m.in_dim, 0, 0, 0, 0, m.in_dim};
            op.blk[1] = {m.W[1], TC_H, 0, 0, 0};
            op.blk[2] = {m.W[2], TC_H, 0, 0, 0, m.out_dim, 0};
            op.bias[0] = m.b[0]; op.bias[1] = m.b[1]; op.bias[2] = m.b[2]; op.gamma = m.gamma; op.beta = m.beta; op.ln_n = m.ln_dim;
            op.out_valid = m.out_dim < TC_H ? m.out_dim : 0;
            op.out = m.out_dim == TC_H ? a.out + r0 * TC_H : O;
            if ((rc = run_chain(op, s))) return rc;
            if (m.out_dim < TC_H && (rc = slice_rows(O, m.out_dim, rows, a.out + r0 * m.out_dim, s))) return rc;
        }
        return CGNN_OK;
    }
    return CGNN_ERR_UNSUPPORTED;
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
// How one MLP's backward over `rows` rows is composed:
//   0  one-layer chains (each runs at the HBM rate) with the LayerNorm backward fused into the Y chain -- the choice for
//      long row streams (edges), where the fused chains' serial per-tile latency costs more than the row-stream passes they save;
//   1  the two fused 3-layer chains of fused_backward -- the choice when every epilogue group has at most one tile (node-sized
//      streams): nothing pipelines across tiles there, so the per-tile latency is paid either way and 2 launches replace 6.
// CGNN_BWD_FUSED=0/1 forces one of them (measurements).
static int bwd_mode(int64_t rows) {
    static int forced = -2;
    if (forced == -2) {
        const char* e = getenv("CGNN_BWD_FUSED");
        forced = e ? atoi(e) : -1;
    }
    if (forced >= 0) return forced;
    const int64_t pair_tiles = (rows + 255) / 256, clusters = num_sms() / 2;
    return pair_tiles <= 2 * clusters ? 1 : 0;
}

// The 2-byte gradient stream (CGNN_PREC_BF16X3_G16), for the backward of an MLP with LayerNorm over a long row stream: the three
// gradient intermediates dY, G2, G1 AND the gradient stream the caller carries from step to step (de_next / de of
// cgnn_mp_edge_bwd, dout of cgnn_mlp_rows_bwd) are bfloat16 in HBM -- 13 of the 21 row-stream passes of a processor step become
// half passes.  Forward values (the activations that decide the ReLU gates) are untouched; what the rounding does to the parameter
// gradients is measured on the oracle by tests/study_grad_stream.py (2.9e-4 at 1 024 particles, halving with every fourfold size:
// it averages out over the rows a weight gradient sums).  The call runs the layered composition whatever its size (the caller
// picks this precision for long streams only).
static bool grad16(int precision, const MlpDev& m) { return precision == CGNN_PREC_BF16X3_G16 && m.gamma != nullptr; }
// row r0 of a [rows][128] stream of 4- or 2-byte elements
static const float* row_at(const float* base, int64_t r0, int half) {
    return base == nullptr ? nullptr : half ? reinterpret_cast<const float*>(reinterpret_cast<const uint16_t*>(base) + r0 * TC_H) : base + r0 * TC_H;
}
static float* row_at(float* base, int64_t r0, int half) { return const_cast<float*>(row_at(static_cast<const float*>(base), r0, half)); }

// Backward of one 3-layer MLP (+ LayerNorm) over a row range as TWO fused chains plus the weight gradients:
//   R  recompute:  in -> A1 -> A2 -> Y, LayerNorm backward of (Y, dU) in the final epilogue -> dY (T);
//                  A1, A2 are written from the hidden epilogues (the dgrad masks and the wgrad operands),
//                  d gamma / d beta come out of the same launch
//   D  dgrad:      dY -> G2 = (dY W3) * [A2 > 0] -> G1 = (G2 W2) * [A1 > 0] (+ per-receiver sum)
//                  -> d_in = G1 * last^T (+ residual); G2, G1 are written from the hidden epilogues
//   dW3 = dY^T A2, dW2 = G2^T A1 (+ biases) here; dW1 = G1^T in by the caller (its input layout differs per phase).
// `r` carries the input side of R (in0 / in1, their weight blocks, bias[0] or the gather); without a LayerNorm
// (decoders) the caller has put dY, zero padded to 128 columns, into T.
static int fused_backward(int ns, const Scratch& sc, const MlpDev& m, const cgnn_mlp_grad* g, int64_t rows, ChainOp r,
                          float* A1, float* A2, float* T, float* G2, uint32_t* gate1, uint32_t* gate2,
                          const float* du_rows, const float* du_recv, int k, int k_valid,
                          ChainBlock last, const float* residual, float* d_in, float* g1_out, float* g1_agg,
                          int accumulate, cudaStream_t s, int g16 = 0, int de16 = 0) {
    int rc;
    const int n_in = r.in1 ? 2 : 1;
    // g16 (see grad16()): T, G2 and g1_out hold bfloat16 rows -- written rounded to nearest even by the chain that produces them,
    // read as they are by the next dgrad chain, the weight gradient and the caller (sender scatter, dW1); du_rows is bfloat16 too,
    // and with de16 so are residual and d_in (the edge phase's gradient stream)
    if (g16 || bwd_mode(rows) == 0) {
        // A1 = relu(layer 1), A2 = relu(A1 W2^T + b2)
        r.ns = ns; r.rows = rows; r.n_layers = 1;
        r.relu_out = 1; r.out = A1; r.bits_out = gate1;          // the ReLU gates go out as 16 B per row for the dgrad chains
        if ((rc = run_chain(r, s))) return rc;
        {
            ChainOp op = base_op(ns, sc, rows);
            op.in0 = A1; op.blk[0] = {m.W[1], TC_H, 0, 0, 0}; op.bias[0] = m.b[1]; op.relu_out = 1; op.out = A2; op.bits_out = gate2;
            if ((rc = run_chain(op, s))) return rc;
        }
        if (m.gamma != nullptr) {   // dY = LayerNorm backward of (Y = A2 W3^T + b3, dU), in the chain's final epilogue
            ChainOp op = base_op(ns, sc, rows);
            op.in0 = A2; op.blk[0] = {m.W[2], TC_H, 0, 0, 0}; op.bias[0] = m.b[2];
            op.gamma = m.gamma; op.beta = m.beta; op.ln_n = m.ln_dim; op.ln_bwd = 1; op.k = k; op.k_valid = k_valid;
            op.du_rows = du_rows; op.du_recv = du_recv; op.dgamma = g->ln_gamma; op.dbeta = g->ln_beta; op.accumulate = accumulate; op.ln_ws = sc.lnb;
            op.out = T; op.out16 = g16; op.du16 = g16;
            if ((rc = run_chain(op, s))) return rc;
        }
        if ((rc = run_wgrad(ns, T, A2, rows, g->W[2], TC_H, 0, g->b[2], accumulate, sc.wg, s, m.out_dim, 0, 0, g16))) return rc;
        {   // G2 = (dY W3) * [A2 > 0]
            ChainOp op = base_op(ns, sc, rows);
            op.in0 = T; op.blk[0] = {m.W[2], TC_H, 0, 0, 1, 0, m.out_dim}; op.mask_bits = gate2; op.out = G2;
            op.in16 = g16; op.out16 = g16;
            if ((rc = run_chain(op, s))) return rc;
        }
        if ((rc = run_wgrad(ns, G2, A1, rows, g->W[1], TC_H, 0, g->b[1], accumulate, sc.wg, s, 0, 0, 0, g16))) return rc;
        {   // G1 = (G2 W2) * [A1 > 0]   (+ per-receiver sum, taken from the FP32 accumulators)
            ChainOp op = base_op(ns, sc, rows);
            op.in0 = G2; op.blk[0] = {m.W[1], TC_H, 0, 0, 1}; op.mask_bits = gate1; op.out = g1_out;
            op.k = k; op.agg_out = g1_agg;
            op.in16 = g16; op.out16 = g16;
            if ((rc = run_chain(op, s))) return rc;
        }
        if (d_in != nullptr) {   // d_in = G1 last^T (+ residual)
            ChainOp op = base_op(ns, sc, rows);
            op.in0 = g1_out; op.blk[0] = last; op.residual = residual; op.out = d_in; op.in16 = g16;
            op.res16 = de16; op.out16 = de16;
            if ((rc = run_chain(op, s))) return rc;
        }
        return CGNN_OK;
    }
    CGNN_CHECK_ARG(!g16, "the bfloat16 gradient stream belongs to the layered composition");
    r.ns = ns; r.rows = rows; r.n_layers = 3;
    r.blk[n_in] = {m.W[1], TC_H, 0, 0, 0};
    r.blk[n_in + 1] = {m.W[2], TC_H, 0, 0, 0, m.out_dim, 0};
    r.bias[1] = m.b[1]; r.bias[2] = m.b[2];
    r.out_valid = m.out_dim < TC_H ? m.out_dim : 0;
    r.hid_out[0] = A1; r.hid_out[1] = A2;
    if (m.gamma != nullptr) {
        r.gamma = m.gamma; r.beta = m.beta; r.ln_n = m.ln_dim; r.ln_bwd = 1; r.k = k; r.k_valid = k_valid;
        r.du_rows = du_rows; r.du_recv = du_recv; r.dgamma = g->ln_gamma; r.dbeta = g->ln_beta; r.accumulate = accumulate; r.ln_ws = sc.lnb;
        r.out = T;
    } else {
        r.out = G2;                                   // Y itself is not needed: parked in G2, which D overwrites
    }
    if ((rc = run_chain(r, s))) return rc;
    // dW3 = dY^T A2, db3
    if ((rc = run_wgrad(ns, T, A2, rows, g->W[2], TC_H, 0, g->b[2], accumulate, sc.wg, s, m.out_dim, 0))) return rc;
    ChainOp d = base_op(ns, sc, rows);
    d.n_layers = 3;
    d.in0 = T;
    d.blk[0] = {m.W[2], TC_H, 0, 0, 1, 0, m.out_dim};
    d.blk[1] = {m.W[1], TC_H, 0, 0, 1};
    d.blk[2] = last;
    d.hid_mask[0] = A2; d.hid_mask[1] = A1;
    d.hid_out[0] = G2; d.hid_out[1] = g1_out;
    d.k = k; d.hid_agg[1] = g1_agg;
    d.residual = residual;
    d.out = d_in ? d_in : T;                          // no input gradient wanted: the last product lands on its own input tile
    if ((rc = run_chain(d, s))) return rc;
    // dW2 = G2^T A1, db2
    return run_wgrad(ns, G2, A1, rows, g->W[1], TC_H, 0, g->b[1], accumulate, sc.wg, s);
}

int tc_mlp_bwd(MlpTask& a, const cgnn_mlp_grad* g, void* ws, int64_t wsb, int precision, cudaStream_t s) {
    const int ns = precision == CGNN_PREC_BF16X3 || precision == CGNN_PREC_BF16X3_G16 ? 3 : 1;
    const MlpDev& m = a.mlp;
    int rc;
    if (a.mode == MODE_ROWS) {
        // input gradients come out 128 wide: only produced directly when the input is 128 wide (decoders)
        if (!tc_rows_ok(m) || (a.dx != nullptr && m.in_dim != TC_H)) return CGNN_ERR_UNSUPPORTED;
        const int64_t need = tc_rows_workspace(nullptr, a.n, precision, 1);
        if (ws == nullptr || wsb < need) {
            set_error("cgnn_mlp_rows_bwd: workspace too small (%lld < %lld)", (long long)wsb, (long long)need);
            return CGNN_ERR_WORKSPACE;
        }
        Carver cv(ws);
        Scratch sc; sc.carve(cv);
        const int64_t chunk = a.n < CHUNK_ROWS ? a.n : CHUNK_ROWS;
        float* Xp = cv.take<float>(chunk * TC_H); float* A1 = cv.take<float>(chunk * TC_H); float* A2 = cv.take<float>(chunk * TC_H);
        float* T = cv.take<float>(chunk * TC_H); float* G2 = cv.take<float>(chunk * TC_H); float* G1 = cv.take<float>(chunk * TC_H);
        uint32_t* gate1 = cv.take<uint32_t>(chunk * 4); uint32_t* gate2 = cv.take<uint32_t>(chunk * 4);
        for (int64_t r0 = 0, c = 0; r0 < a.n; r0 += chunk, ++c) {
            const int64_t rows = a.n - r0 < chunk ? a.n - r0 : chunk;
            const int acc = c > 0;
            const float* in = a.x + r0 * m.in_dim;
            const bool narrow_tma = m.in_dim < TC_H && (m.in_dim * 4) % 16 == 0 && ((uintptr_t)in & 15) == 0;
            if (m.in_dim < TC_H && !narrow_tma) {
                if ((rc = pad_rows(in, m.in_dim, rows, Xp, s))) return rc;
                in = Xp;
            }
            const int g16 = grad16(precision, m);            // (LayerNorm MLPs have out_dim == 128)
            const float* dU = g16 ? row_at(a.dout, r0, 1) : a.dout + r0 * m.out_dim;
            if (m.gamma == nullptr) {                       // no LayerNorm: dY = dout, zero padded
                if (m.out_dim < TC_H) { if ((rc = pad_rows(dU, m.out_dim, rows, T, s))) return rc; }
                else CGNN_CUDA(cudaMemcpyAsync(T, dU, (size_t)rows * TC_H * 4, cudaMemcpyDeviceToDevice, s));
            }
            ChainOp r{};
            r.in0 = in; r.in0_cols = narrow_tma ? m.in_dim : 0; r.blk[0] = {m.W[0], m.in_dim, 0, 0, 0, 0, m.in_dim}; r.bias[0] = m.b[0];
            // dx = G1 W1 (in_dim == 128) comes out of the dgrad chain's last layer
            if ((rc = fused_backward(ns, sc, m, g, rows, r, A1, A2, T, G2, gate1, gate2, dU, nullptr, 1, 0, {m.W[0], m.in_dim, 0, 0, 1, m.in_dim, 0},
                                     nullptr, a.dx ? a.dx + r0 * TC_H : nullptr, G1, nullptr, acc, s, g16))) return rc;
            if ((rc = run_wgrad(ns, G1, in, rows, g->W[0], m.in_dim, 0, g->b[0], acc, sc.wg, s, 0, m.in_dim, narrow_tma ? m.in_dim : 0, g16))) return rc;
        }
        return CGNN_OK;
    }
    if (a.mode == MODE_NODE) {
        if (!tc_shape_ok(m, 2)) return CGNN_ERR_UNSUPPORTED;
        const int64_t need = tc_bwd_workspace(nullptr, a.n, a.n, 0, precision);
        if (ws == nullptr || wsb < need) {
            set_error("cgnn_mp_node_bwd: workspace too small (%lld < %lld)", (long long)wsb, (long long)need);
            return CGNN_ERR_WORKSPACE;
        }
        Carver cv(ws);
        Scratch sc; sc.carve(cv);
        const int64_t n = a.n;
        float* A1 = cv.take<float>(n * TC_H); float* A2 = cv.take<float>(n * TC_H); float* T = cv.take<float>(n * TC_H);
        float* G2 = cv.take<float>(n * TC_H); float* G1 = cv.take<float>(n * TC_H);
        uint32_t* gate1 = cv.take<uint32_t>(n * 4); uint32_t* gate2 = cv.take<uint32_t>(n * 4);
        ChainOp r{};
        r.in0 = a.h; r.in1 = a.agg;
        r.blk[0] = {m.W[0], 2 * TC_H, 0, 0, 0}; r.blk[1] = {m.W[0], 2 * TC_H, 0, TC_H, 0};
        r.bias[0] = m.b[0];
        // dh = dh_next + G1 W1[:, 0:L] is the dgrad chain's last layer
        if ((rc = fused_backward(ns, sc, m, g, n, r, A1, A2, T, G2, gate1, gate2, a.dout, nullptr, 1, 0, {m.W[0], 2 * TC_H, 0, 0, 1}, a.dout, a.dh, G1,
                                 nullptr, 0, s))) return rc;
        // dW1 = G1^T [h | agg], db1
        if ((rc = run_wgrad(ns, G1, a.h, n, g->W[0], 2 * TC_H, 0, g->b[0], 0, sc.wg, s))) return rc;
        if ((rc = run_wgrad(ns, G1, a.agg, n, g->W[0], 2 * TC_H, TC_H, nullptr, 0, sc.wg, s))) return rc;
        {   // dagg = G1 W1[:, L:2L]
            ChainOp op = base_op(ns, sc, n);
            op.in0 = G1; op.blk[0] = {m.W[0], 2 * TC_H, 0, TC_H, 1}; op.out = a.dagg_out;
            if ((rc = run_chain(op, s))) return rc;
        }
        return CGNN_OK;
    }
    if (a.mode == MODE_EDGE) {
        if (!tc_shape_ok(m, 3) || !tc_k_ok(a.k) || a.t_rowptr == nullptr || a.t_perm == nullptr) return CGNN_ERR_UNSUPPORTED;
        const int64_t n = a.n, k = a.k, E = n * k;
        const int64_t nn = a.n_nodes > 0 ? a.n_nodes : a.n;
        const int64_t need = tc_bwd_workspace(nullptr, n, nn, a.k, precision);
        if (ws == nullptr || wsb < need) {
            set_error("cgnn_mp_edge_bwd: workspace too small (%lld < %lld)", (long long)wsb, (long long)need);
            return CGNN_ERR_WORKSPACE;
        }
        Carver cv(ws);
        Scratch sc; sc.carve(cv);
        const int64_t chunk = E < CHUNK_ROWS ? E : CHUNK_ROWS;
        float* A1 = cv.take<float>(chunk * TC_H); float* A2 = cv.take<float>(chunk * TC_H); float* T = cv.take<float>(chunk * TC_H);
        float* G2 = cv.take<float>(chunk * TC_H); float* G1 = cv.take<float>(chunk * TC_H);
        uint32_t* gate1 = cv.take<uint32_t>(chunk * 4); uint32_t* gate2 = cv.take<uint32_t>(chunk * 4);
        float* Ps = cv.take<float>(nn * TC_H); float* Pr = cv.take<float>(n * TC_H);
        float* dPs = cv.take<float>(nn * TC_H); float* dPr = cv.take<float>(n * TC_H);
        const int n_chunks = (int)((E + chunk - 1) / chunk);
        int2* ent = cv.take<int2>(scatter_entries(E, nn, chunk));
        int32_t* ecnt = cv.take<int32_t>(n_chunks + 1); int32_t* eoff = cv.take<int32_t>(n_chunks + 1);
        if ((rc = project_nodes(ns, sc, m, a.h, n, nn, Ps, Pr, s))) return rc;
        // per-node sums of G1 by sender: accumulated chunk by chunk over the sender-sorted transpose (no E-sized buffer)
        CGNN_CUDA(cudaMemsetAsync(dPs, 0, (size_t)nn * TC_H * 4, s));
        CGNN_CUDA(cudaMemsetAsync(ecnt, 0, (size_t)(n_chunks + 1) * 4, s));
        const unsigned scgrid = (unsigned)((nn + SC_NODES - 1) / SC_NODES);
        const size_t scsmem = (size_t)2 * n_chunks * sizeof(int32_t);
        CGNN_CHECK_ARG(scsmem <= 48 * 1024, "cgnn_mp_edge_bwd: too many backward chunks (%d)", n_chunks);
        sender_chunks_kernel<<<scgrid, SC_WARPS * 32, scsmem, s>>>(a.t_rowptr, a.t_perm, nn, (int)chunk, n_chunks, ecnt, nullptr, nullptr);
        CGNN_LAUNCH_CHECK();
        sender_chunks_offsets_kernel<<<1, 32, 0, s>>>(ecnt, eoff, n_chunks);
        CGNN_LAUNCH_CHECK();
        sender_chunks_kernel<<<scgrid, SC_WARPS * 32, scsmem, s>>>(a.t_rowptr, a.t_perm, nn, (int)chunk, n_chunks, ecnt, eoff, ent);
        CGNN_LAUNCH_CHECK();
        for (int64_t r0 = 0, c = 0; r0 < E; r0 += chunk, ++c) {
            const int64_t rows = E - r0 < chunk ? E - r0 : chunk;       // chunk is a multiple of 256 and of k unless it is the whole graph
            const float* e_in = a.e_in + r0 * TC_H;
            const int g16 = grad16(precision, m);
            const float* de_next = row_at(a.de_next, r0, g16);
            const int acc = c > 0;
            // recompute with e W1e^T + Ps[sender] + Pr[receiver] as layer 1; dU = de_next + dagg[receiver];
            // de = de_next + G1 W1e is the dgrad chain's last layer, the per-receiver sum of G1 is d P_r
            ChainOp r{};
            r.in0 = e_in; r.blk[0] = {m.W[0], 3 * TC_H, 0, 2 * TC_H, 0};
            r.k = a.k; r.k_valid = a.k_valid; r.senders = a.senders + r0; r.Ps = Ps; r.ps_rows = nn; r.Pr = Pr + (r0 / k) * TC_H;
            if ((rc = fused_backward(ns, sc, m, g, rows, r, A1, A2, T, G2, gate1, gate2, de_next, a.dagg + (r0 / k) * TC_H, a.k, a.k_valid,
                                     {m.W[0], 3 * TC_H, 0, 2 * TC_H, 1}, de_next, row_at(a.de, r0, g16), G1, dPr + (r0 / k) * TC_H, acc, s, g16, g16))) return rc;
            // dW1e = G1^T e, db1
            if ((rc = run_wgrad(ns, G1, e_in, rows, g->W[0], 3 * TC_H, 2 * TC_H, g->b[0], acc, sc.wg, s, 0, 0, 0, g16))) return rc;
            // (grid: enough warps to cover the memory latency; every warp strides over the chunk's entries)
            const unsigned sgrid = (unsigned)(num_sms() * 8);
            if (g16) scatter_chunk_kernel<true><<<sgrid, SCATTER_WARPS * 32, 0, s>>>(
                         G1, (int)r0, (int)(r0 + rows), a.t_rowptr, a.t_perm, ent, eoff + c, reinterpret_cast<float4*>(dPs));
            else scatter_chunk_kernel<false><<<sgrid, SCATTER_WARPS * 32, 0, s>>>(
                     G1, (int)r0, (int)(r0 + rows), a.t_rowptr, a.t_perm, ent, eoff + c, reinterpret_cast<float4*>(dPs));
            CGNN_LAUNCH_CHECK();
        }
        // (the per-receiver sums dPr came out of the chunks' G1 chains)
        if ((rc = run_wgrad(ns, dPs, a.h, nn, g->W[0], 3 * TC_H, 0, nullptr, 0, sc.wg, s))) return rc;
        if ((rc = run_wgrad(ns, dPr, a.h, n, g->W[0], 3 * TC_H, TC_H, nullptr, 0, sc.wg, s))) return rc;
        if (nn == n) {   // dh += dPs W1s + dPr W1r
            ChainOp op = base_op(ns, sc, n);
            op.in0 = dPs; op.in1 = dPr;
            op.blk[0] = {m.W[0], 3 * TC_H, 0, 0, 1}; op.blk[1] = {m.W[0], 3 * TC_H, 0, TC_H, 1};
            op.residual = a.dh; op.out = a.dh;
            if ((rc = run_chain(op, s))) return rc;
        } else {         // halo rows only have the sender part
            ChainOp op = base_op(ns, sc, nn);
            op.in0 = dPs; op.blk[0] = {m.W[0], 3 * TC_H, 0, 0, 1}; op.residual = a.dh; op.out = a.dh;
            if ((rc = run_chain(op, s))) return rc;
            ChainOp op2 = base_op(ns, sc, n);
            op2.in0 = dPr; op2.blk[0] = {m.W[0], 3 * TC_H, 0, TC_H, 1}; op2.residual = a.dh; op2.out = a.dh;
            if ((rc = run_chain(op2, s))) return rc;
        }
        return CGNN_OK;
    }
    return CGNN_ERR_UNSUPPORTED;
}

}  // namespace cgnn
