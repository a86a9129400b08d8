// FP32 SIMT implementation of the fused MLP(+LayerNorm) tiles: encoder / decoder rows (K3, K6),
// the message-passing edge and node phases (K4) and their backward with in-tile activation
// recompute (K5).  This is the FP32-only parity mode (<= 1e-5 against the oracle); it supports
// every width the reference allows (latent/hidden <= 256, 1..3 hidden layers, any k <= 64).
//
// One CTA = one tile of rows.  Layer inputs are gathered straight from the latents (no [E,3L]
// concat in HBM, graph_network.py:89,94), every layer's output stays in shared memory, LayerNorm,
// the residual add and the fixed-order k-row segmented sum happen in the epilogue, and only the
// new latents are written back.
#include "common.cuh"
#include "mlp_common.cuh"

namespace cgnn {
namespace {

constexpr int NTHREADS = 256;
constexpr int KC = 32;                 // K chunk staged per iteration
constexpr int WPAD = 257;              // staged weight chunk row stride (max 256 cols + 1)
constexpr float LN_EPS = 1e-5f;

// ------------------------------------------------------------------------------------------------
// tile geometry
// ------------------------------------------------------------------------------------------------
struct TileInfo {
    int64_t row0;      // first global row (edge id / node id / row id)
    int rows;          // valid rows in this tile
    int64_t recv0;     // EDGE: first receiver
    int nrecv;         // EDGE: receivers in this tile
};

template <int TMR>
__device__ __forceinline__ TileInfo tile_info(const MlpTask& a, int64_t tile) {
    TileInfo t;
    if (a.mode == MODE_EDGE) {
        int rpt = TMR / a.k;
        t.recv0 = tile * rpt;
        int64_t left = a.n - t.recv0;
        t.nrecv = (int)(left < rpt ? left : rpt);
        t.row0 = t.recv0 * a.k;
        t.rows = t.nrecv * a.k;
    } else {
        t.row0 = tile * TMR;
        int64_t left = a.n - t.row0;
        t.rows = (int)(left < TMR ? left : TMR);
        t.recv0 = 0; t.nrecv = 0;
    }
    return t;
}

// element (row, c) of the layer-0 input, gathered according to the mode
__device__ __forceinline__ float load_input(const MlpTask& a, const TileInfo& t, const int* sSend, int row, int c) {
    int64_t g = t.row0 + row;
    if (a.mode == MODE_ROWS) return a.x[g * a.mlp.in_dim + c];
    const int L = a.L;
    if (a.mode == MODE_EDGE) {
        if (c < L) return a.h[(int64_t)sSend[row] * L + c];
        if (c < 2 * L) return a.h[(g / a.k) * L + (c - L)];
        return a.e_in[g * L + (c - 2 * L)];
    }
    // MODE_NODE
    if (c < L) return a.h[g * L + c];
    return a.agg[g * L + (c - L)];
}

// ------------------------------------------------------------------------------------------------
// out[row][n0 + col] = sum_k A[row][k] * Wt[k][n0 + col]      (register tile RPT x NT per thread)
//   A:  shared activations (sA != nullptr, row stride a_stride) or the gathered layer-0 input
//   Wt: W[n*ldw + k] (forward, torch layout)  or  W[k*ldw + n] (w_kn = true: the dIn product)
// ------------------------------------------------------------------------------------------------
template <int RPT, int NT>
__device__ __forceinline__ void gemm_tile(const MlpTask& a, const TileInfo& t, const int* sSend,
                                          const float* sA, int a_stride,
                                          const float* __restrict__ W, int ldw, bool w_kn,
                                          int K, int n0, int ncols,
                                          float* sIn, float* sW, float (&acc)[RPT][NT]) {
    constexpr int TMR = RPT * 16;
    const int tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;
#pragma unroll
    for (int i = 0; i < RPT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[i][j] = 0.0f;

    for (int k0 = 0; k0 < K; k0 += KC) {
        // stage the weight chunk: sW[kk][col] = Wt[k0+kk][n0+col]
        if (!w_kn) {
            const int kk = tid & 31;
            for (int col = tid >> 5; col < ncols; col += NTHREADS / 32) {
                float v = 0.0f;
                if (k0 + kk < K) v = W[(int64_t)(n0 + col) * ldw + k0 + kk];
                sW[kk * WPAD + col] = v;
            }
        } else {
            for (int idx = tid; idx < KC * ncols; idx += NTHREADS) {
                int kk = idx / ncols, col = idx - kk * ncols;
                float v = 0.0f;
                if (k0 + kk < K) v = W[(int64_t)(k0 + kk) * ldw + n0 + col];
                sW[kk * WPAD + col] = v;
            }
        }
        if (sA == nullptr) {
            const int kk = tid & 31;
            for (int row = tid >> 5; row < TMR; row += NTHREADS / 32) {
                float v = 0.0f;
                if (row < t.rows && k0 + kk < K) v = load_input(a, t, sSend, row, k0 + kk);
                sIn[row * (KC + 1) + kk] = v;
            }
        }
        __syncthreads();
        const int kmax = (K - k0 < KC) ? (K - k0) : KC;
        if (sA == nullptr) {
#pragma unroll 4
            for (int kk = 0; kk < kmax; ++kk) {
                float av[RPT], bv[NT];
#pragma unroll
                for (int i = 0; i < RPT; ++i) av[i] = sIn[(ty * RPT + i) * (KC + 1) + kk];
#pragma unroll
                for (int j = 0; j < NT; ++j) bv[j] = sW[kk * WPAD + tx + 16 * j];
#pragma unroll
                for (int i = 0; i < RPT; ++i)
#pragma unroll
                    for (int j = 0; j < NT; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            }
        } else {
#pragma unroll 4
            for (int kk = 0; kk < kmax; ++kk) {
                float av[RPT], bv[NT];
#pragma unroll
                for (int i = 0; i < RPT; ++i) av[i] = sA[(ty * RPT + i) * a_stride + k0 + kk];
#pragma unroll
                for (int j = 0; j < NT; ++j) bv[j] = sW[kk * WPAD + tx + 16 * j];
#pragma unroll
                for (int i = 0; i < RPT; ++i)
#pragma unroll
                    for (int j = 0; j < NT; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            }
        }
        __syncthreads();
    }
}

template <int RPT, int NT>
__device__ __forceinline__ void linear_fwd_nt(const MlpTask& a, const TileInfo& t, const int* sSend,
                                              const float* sA, int a_stride, const float* __restrict__ W,
                                              const float* __restrict__ bias, int K, int N, bool relu,
                                              float* sOut, int o_stride, float* sIn, float* sW) {
    float acc[RPT][NT];
    gemm_tile<RPT, NT>(a, t, sSend, sA, a_stride, W, K, false, K, 0, N, sIn, sW, acc);
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
#pragma unroll
    for (int i = 0; i < RPT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            int n = tx + 16 * j;
            if (n < N) {
                float v = acc[i][j] + bias[n];
                sOut[(ty * RPT + i) * o_stride + n] = relu ? fmaxf(v, 0.0f) : v;
            }
        }
}

// one Linear layer of the forward: sOut[row][n] = act(b[n] + sum_k A[row][k] W[n][k])
template <int RPT>
__device__ __forceinline__ void linear_fwd(const MlpTask& a, const TileInfo& t, const int* sSend, int layer,
                                           const float* sA, int a_stride, float* sOut, int o_stride,
                                           float* sIn, float* sW) {
    const MlpDev& m = a.mlp;
    const int K = layer == 0 ? m.in_dim : m.hidden;
    const int N = layer == m.n_layers - 1 ? m.out_dim : m.hidden;
    const bool relu = layer < m.n_layers - 1;
    const float* __restrict__ bias = m.b[layer];
    if (N <= 16)       linear_fwd_nt<RPT, 1>(a, t, sSend, sA, a_stride, m.W[layer], bias, K, N, relu, sOut, o_stride, sIn, sW);
    else if (N <= 64)  linear_fwd_nt<RPT, 4>(a, t, sSend, sA, a_stride, m.W[layer], bias, K, N, relu, sOut, o_stride, sIn, sW);
    else if (N <= 128) linear_fwd_nt<RPT, 8>(a, t, sSend, sA, a_stride, m.W[layer], bias, K, N, relu, sOut, o_stride, sIn, sW);
    else               linear_fwd_nt<RPT, 16>(a, t, sSend, sA, a_stride, m.W[layer], bias, K, N, relu, sOut, o_stride, sIn, sW);
    __syncthreads();
}

// per-row LayerNorm statistics of sY[row][0..N): warp per row
__device__ __forceinline__ void row_stats(const float* sY, int stride, int rows, int N, float* sMean, float* sRstd) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int row = warp; row < rows; row += NTHREADS / 32) {
        float s = 0.0f;
        for (int c = lane; c < N; c += 32) s += sY[row * stride + c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
        float mean = s / (float)N;
        float v = 0.0f;
        for (int c = lane; c < N; c += 32) { float d = sY[row * stride + c] - mean; v = fmaf(d, d, v); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
        if (lane == 0) { sMean[row] = mean; sRstd[row] = 1.0f / sqrtf(v / (float)N + LN_EPS); }
    }
}

__device__ __forceinline__ void load_senders(const MlpTask& a, const TileInfo& t, int* sSend, int tmr) {
    if (a.mode == MODE_EDGE)
        for (int r = threadIdx.x; r < tmr; r += NTHREADS) sSend[r] = r < t.rows ? a.senders[t.row0 + r] : 0;
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// forward kernel
// ------------------------------------------------------------------------------------------------
constexpr int TM_F = 64;
constexpr int RPT_F = TM_F / 16;

__global__ void __launch_bounds__(NTHREADS)
mlp_fwd_kernel(MlpTask a, int64_t n_tiles) {
    extern __shared__ float smem[];
    const MlpDev& m = a.mlp;
    const int AS = a.act_stride;
    float* sAct0 = smem;
    float* sAct1 = sAct0 + TM_F * AS;
    float* sIn = sAct1 + TM_F * AS;
    float* sW = sIn + TM_F * (KC + 1);
    float* sMean = sW + KC * WPAD;
    float* sRstd = sMean + TM_F;
    int* sSend = reinterpret_cast<int*>(sRstd + TM_F);

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        TileInfo t = tile_info<TM_F>(a, tile);
        load_senders(a, t, sSend, TM_F);
        float* cur = nullptr;
        float* nxt = sAct0;
        for (int l = 0; l < m.n_layers; ++l) {
            linear_fwd<RPT_F>(a, t, sSend, l, cur, AS, nxt, AS, sIn, sW);
            cur = nxt;
            nxt = (cur == sAct0) ? sAct1 : sAct0;
        }
        const int N = m.out_dim;
        const bool ln = m.gamma != nullptr;
        if (ln) { row_stats(cur, AS, t.rows, m.ln_dim, sMean, sRstd); __syncthreads(); }   // (columns beyond ln_dim are zero padding)
        // epilogue: normalise, residual, write
        for (int idx = threadIdx.x; idx < t.rows * N; idx += NTHREADS) {
            int row = idx / N, c = idx - row * N;
            float v = cur[row * AS + c];
            if (ln) v = (v - sMean[row]) * sRstd[row] * m.gamma[c] + m.beta[c];
            int64_t g = t.row0 + row;
            if (a.mode == MODE_ROWS) {
                a.out[g * N + c] = v;
            } else if (a.mode == MODE_EDGE) {
                if (a.out != nullptr) a.out[g * N + c] = a.e_in[g * N + c] + v;       // e' = e + u_e   (graph_network.py:182)
                cur[row * AS + c] = v;                          // keep u_e for the segmented sum
            } else {
                a.out[g * N + c] = a.h[g * N + c] + v;          // h' = h + u_n   (graph_network.py:181)
            }
        }
        if (a.mode == MODE_EDGE && a.agg_out != nullptr) {
            __syncthreads();
            // deterministic segmented sum: the k in-edges of a receiver are k consecutive rows
            for (int idx = threadIdx.x; idx < t.nrecv * N; idx += NTHREADS) {
                int r = idx / N, c = idx - r * N;
                float s = 0.0f;
                for (int q = 0; q < a.k; ++q) s += cur[(r * a.k + q) * AS + c];
                a.agg_out[(t.recv0 + r) * N + c] = s;
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// backward kernel (recompute inside the tile)
// ------------------------------------------------------------------------------------------------
constexpr int TM_B = 32;
constexpr int RPT_B = TM_B / 16;

// dW[n][kcol] (+)= sum_row dY[row][n] * In[row][kcol] for one K chunk of 32 input columns
__device__ __forceinline__ void wgrad_chunk(const float* sDY, int dy_stride, int N, int rows,
                                            const float* sInp, int in_stride, int in_off,
                                            float* __restrict__ P, int ldp, int k0, int K, bool first) {
    const int tk = threadIdx.x & 31, tn = threadIdx.x >> 5;      // tn is warp-uniform
    if (k0 + tk >= K) return;
    for (int nb = 0; nb < N; nb += 64) {
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.0f;
        for (int row = 0; row < rows; ++row) {
            float x = sInp[row * in_stride + in_off + tk];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                int n = nb + tn + 8 * j;
                float d = (n < N) ? sDY[row * dy_stride + n] : 0.0f;
                acc[j] = fmaf(d, x, acc[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int n = nb + tn + 8 * j;
            if (n < N) {
                float* dst = P + (int64_t)n * ldp + k0 + tk;
                *dst = first ? acc[j] : (*dst + acc[j]);
            }
        }
    }
}

__global__ void __launch_bounds__(NTHREADS)
mlp_bwd_kernel(MlpTask a, int64_t n_tiles, float* __restrict__ partials, int64_t blob) {
    extern __shared__ float smem[];
    const MlpDev& m = a.mlp;
    const int AS = a.act_stride;
    const int nl = m.n_layers;
    float* sActs = smem;                                   // [nl][TM_B][AS]  layer outputs (post-ReLU / pre-LN)
    float* sG0 = sActs + nl * TM_B * AS;                   // [TM_B][AS] gradient ping
    float* sG1 = sG0 + TM_B * AS;                          // [TM_B][AS] gradient pong
    float* sIn = sG1 + TM_B * AS;                          // [TM_B][KC+1]
    float* sW = sIn + TM_B * (KC + 1);                     // [KC][WPAD]
    float* sMean = sW + KC * WPAD;
    float* sRstd = sMean + TM_B;
    float* sVec = sRstd + TM_B;                            // [(nl + 2)][256]: db per layer, dgamma, dbeta
    int* sSend = reinterpret_cast<int*>(sVec + (nl + 2) * 256);

    float* P = partials + (int64_t)blockIdx.x * blob;      // this CTA's private partial blob
    for (int i = threadIdx.x; i < (nl + 2) * 256; i += NTHREADS) sVec[i] = 0.0f;
    __syncthreads();

    bool first = true;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, first = false) {
        TileInfo t = tile_info<TM_B>(a, tile);
        load_senders(a, t, sSend, TM_B);
        // ---- recompute the forward, keeping every layer output -------------------------------
        for (int l = 0; l < nl; ++l)
            linear_fwd<RPT_B>(a, t, sSend, l, l == 0 ? nullptr : sActs + (l - 1) * TM_B * AS, AS,
                              sActs + l * TM_B * AS, AS, sIn, sW);
        const int NO = m.out_dim;
        float* sY = sActs + (nl - 1) * TM_B * AS;
        const bool ln = m.gamma != nullptr;
        // ---- upstream gradient dZ -> sG0 ---------------------------------------------------------
        for (int idx = threadIdx.x; idx < TM_B * NO; idx += NTHREADS) {
            int row = idx / NO, c = idx - row * NO;
            float v = 0.0f;
            if (row < t.rows) {
                int64_t g = t.row0 + row;
                if (a.mode == MODE_ROWS) v = a.dout[g * NO + c];
                else if (a.mode == MODE_NODE) v = a.dout[g * NO + c];                       // dh_next
                else v = (a.de_next ? a.de_next[g * NO + c] : 0.0f) + a.dagg[(g / a.k) * NO + c];
            }
            sG0[row * AS + c] = v;
        }
        if (ln) row_stats(sY, AS, t.rows, m.ln_dim, sMean, sRstd);
        __syncthreads();
        if (ln) {
            // dgamma / dbeta: thread per column, rows in fixed order
            for (int c = threadIdx.x; c < NO; c += NTHREADS) {
                float dg = 0.0f, db = 0.0f;
                for (int row = 0; row < t.rows; ++row) {
                    float yh = (sY[row * AS + c] - sMean[row]) * sRstd[row];
                    float dz = sG0[row * AS + c];
                    dg = fmaf(dz, yh, dg);
                    db += dz;
                }
                sVec[nl * 256 + c] += dg;
                sVec[(nl + 1) * 256 + c] += db;
            }
            __syncthreads();
            // dY = (g - mean(g) - yhat * mean(g*yhat)) * rstd,  g = dZ * gamma ; warp per row, in place
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
            for (int row = warp; row < t.rows; row += NTHREADS / 32) {
                float s1 = 0.0f, s2 = 0.0f;
                for (int c = lane; c < NO; c += 32) {
                    float g = sG0[row * AS + c] * m.gamma[c];
                    float yh = (sY[row * AS + c] - sMean[row]) * sRstd[row];
                    s1 += g; s2 = fmaf(g, yh, s2);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xFFFFFFFFu, s1, o); s2 += __shfl_xor_sync(0xFFFFFFFFu, s2, o); }
                s1 /= (float)m.ln_dim; s2 /= (float)m.ln_dim;
                for (int c = lane; c < NO; c += 32) {
                    float g = sG0[row * AS + c] * m.gamma[c];
                    float yh = (sY[row * AS + c] - sMean[row]) * sRstd[row];
                    sG0[row * AS + c] = c < m.ln_dim ? (g - s1 - yh * s2) * sRstd[row] : 0.0f;
                }
            }
            __syncthreads();
        }
        // ---- layers, last to first ---------------------------------------------------------------
        float* sDY = sG0;
        float* sDN = sG1;
        for (int l = nl - 1; l >= 0; --l) {
            const int K = l == 0 ? m.in_dim : m.hidden;
            const int N = l == nl - 1 ? m.out_dim : m.hidden;
            float* PW = P + a.w_off[l];
            // bias gradient
            for (int c = threadIdx.x; c < N; c += NTHREADS) {
                float s = 0.0f;
                for (int row = 0; row < t.rows; ++row) s += sDY[row * AS + c];
                sVec[l * 256 + c] += s;
            }
            // weight gradient, chunk by chunk over the layer input
            if (l > 0) {
                const float* sInp = sActs + (l - 1) * TM_B * AS;
                for (int k0 = 0; k0 < K; k0 += KC)
                    wgrad_chunk(sDY, AS, N, t.rows, sInp, AS, k0, PW, K, k0, K, first);
                __syncthreads();
                // dIn = dY * W, then ReLU mask of the previous layer's output
                const float* W = m.W[l];
                const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
                for (int n0 = 0; n0 < K; n0 += 256) {
                    int ncols = K - n0 < 256 ? K - n0 : 256;
                    float acc[RPT_B][16];
                    gemm_tile<RPT_B, 16>(a, t, sSend, sDY, AS, W, K, true, N, n0, ncols, sIn, sW, acc);
#pragma unroll
                    for (int i = 0; i < RPT_B; ++i)
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            int c = n0 + tx + 16 * j, row = ty * RPT_B + i;
                            if (c < K) sDN[row * AS + c] = (sInp[row * AS + c] > 0.0f && row < t.rows) ? acc[i][j] : 0.0f;
                        }
                }
                __syncthreads();
                float* tmp = sDY; sDY = sDN; sDN = tmp;
            } else {
                // layer 0: the input is gathered again chunk by chunk (never stored)
                for (int k0 = 0; k0 < K; k0 += KC) {
                    const int kk = threadIdx.x & 31;
                    for (int row = threadIdx.x >> 5; row < TM_B; row += NTHREADS / 32) {
                        float v = 0.0f;
                        if (row < t.rows && k0 + kk < K) v = load_input(a, t, sSend, row, k0 + kk);
                        sIn[row * (KC + 1) + kk] = v;
                    }
                    __syncthreads();
                    wgrad_chunk(sDY, AS, N, t.rows, sIn, KC + 1, 0, PW, K, k0, K, first);
                    __syncthreads();
                }
                if (a.need_input_grad) {
                    const float* W = m.W[0];
                    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
                    const int L = a.L;
                    for (int n0 = 0; n0 < K; n0 += 256) {
                        int ncols = K - n0 < 256 ? K - n0 : 256;
                        float acc[RPT_B][16];
                        gemm_tile<RPT_B, 16>(a, t, sSend, sDY, AS, W, K, true, N, n0, ncols, sIn, sW, acc);
#pragma unroll
                        for (int i = 0; i < RPT_B; ++i)
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                int cc = tx + 16 * j, row = ty * RPT_B + i;
                                if (cc < ncols) sDN[row * AS + cc] = acc[i][j];
                            }
                        __syncthreads();
                        // consume this block of input-gradient columns
                        if (a.mode == MODE_EDGE) {
                            // receiver columns: fixed-order sum over the k rows of each receiver
                            for (int idx = threadIdx.x; idx < t.nrecv * ncols; idx += NTHREADS) {
                                int r = idx / ncols, cc = idx - r * ncols, c = n0 + cc;
                                if (c >= L && c < 2 * L) {
                                    float s = 0.0f;
                                    for (int q = 0; q < a.k; ++q) s += sDN[(r * a.k + q) * AS + cc];
                                    a.dh[(t.recv0 + r) * L + (c - L)] += s;
                                }
                            }
                        }
                        for (int idx = threadIdx.x; idx < t.rows * ncols; idx += NTHREADS) {
                            int row = idx / ncols, cc = idx - row * ncols, c = n0 + cc;
                            float v = sDN[row * AS + cc];
                            int64_t g = t.row0 + row;
                            if (a.mode == MODE_ROWS) {
                                a.dx[g * K + c] = v;
                            } else if (a.mode == MODE_NODE) {
                                if (c < L) a.dh[g * L + c] = a.dout[g * L + c] + v;       // dh = dh_next + dIn_h
                                else a.dagg_out[g * L + (c - L)] = v;
                            } else {
                                if (c < L) a.gs[g * L + c] = v;                            // per-edge sender gradient
                                else if (c >= 2 * L)
                                    a.de[g * L + (c - 2 * L)] = (a.de_next ? a.de_next[g * L + (c - 2 * L)] : 0.0f) + v;
                            }
                        }
                        __syncthreads();
                    }
                }
            }
        }
        __syncthreads();
    }
    // ---- flush the vector partials (bias, gamma, beta) -------------------------------------------
    for (int l = 0; l < nl; ++l) {
        const int N = l == nl - 1 ? m.out_dim : m.hidden;
        for (int c = threadIdx.x; c < N; c += NTHREADS) P[a.b_off[l] + c] = sVec[l * 256 + c];
    }
    if (m.gamma != nullptr)
        for (int c = threadIdx.x; c < m.out_dim; c += NTHREADS) {
            P[a.g_off + c] = sVec[nl * 256 + c];
            P[a.g_off + m.out_dim + c] = sVec[(nl + 1) * 256 + c];
        }
}

// fixed-order sum of the per-CTA partial blobs into the caller's gradient tensors
__global__ void reduce_partials_kernel(const float* __restrict__ partials, int n_cta, int64_t blob,
                                       GradPtrs gp, MlpDev m, int64_t w_off0, int64_t w_off1, int64_t w_off2,
                                       int64_t w_off3, int64_t b_off0, int64_t b_off1, int64_t b_off2,
                                       int64_t b_off3, int64_t g_off) {
    int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= blob) return;
    float s = 0.0f;
    for (int c = 0; c < n_cta; ++c) s += partials[(int64_t)c * blob + p];
    const int64_t w_off[4] = {w_off0, w_off1, w_off2, w_off3};
    const int64_t b_off[4] = {b_off0, b_off1, b_off2, b_off3};
    for (int l = 0; l < m.n_layers; ++l) {
        const int K = l == 0 ? m.in_dim : m.hidden;
        const int N = l == m.n_layers - 1 ? m.out_dim : m.hidden;
        if (p >= w_off[l] && p < w_off[l] + (int64_t)N * K) { if (gp.W[l]) gp.W[l][p - w_off[l]] = s; return; }
        if (p >= b_off[l] && p < b_off[l] + N) { if (gp.b[l]) gp.b[l][p - b_off[l]] = s; return; }
    }
    if (m.gamma != nullptr) {
        if (p >= g_off && p < g_off + m.out_dim) { if (gp.gamma) gp.gamma[p - g_off] = s; return; }
        if (p >= g_off + m.out_dim && p < g_off + 2 * m.out_dim) { if (gp.beta) gp.beta[p - g_off - m.out_dim] = s; return; }
    }
}

// ------------------------------------------------------------------------------------------------
// small memory-bound helpers
// ------------------------------------------------------------------------------------------------
// agg[i][:] = sum_r h[senders[i*k + r]][:]   (PyG default message, graph_network.py:92)
__global__ void aggregate_senders_kernel(const float* __restrict__ h, const int32_t* __restrict__ senders,
                                         int64_t n, int k, int L4, float4* __restrict__ agg) {
    int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx >= n * L4) return;
    int64_t i = idx / L4;
    int c = (int)(idx - i * L4);
    const float4* h4 = reinterpret_cast<const float4*>(h);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < k; ++r) {
        float4 v = h4[(int64_t)senders[i * k + r] * L4 + c];
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    agg[idx] = s;
}

// dh[j][:] += sum over the sender-sorted transpose row of j (perm order) of src[...]
__global__ void scatter_to_senders_kernel(const float* __restrict__ src, int per_receiver,
                                          const int32_t* __restrict__ rowptr, const int32_t* __restrict__ perm,
                                          int64_t n, int k, int L4, float4* __restrict__ dh) {
    int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx >= n * L4) return;
    int64_t j = idx / L4;
    int c = (int)(idx - j * L4);
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4 s = dh[idx];
    int a = rowptr[j], b = rowptr[j + 1];
    for (int p = a; p < b; ++p) {
        int64_t e = perm[p];
        int64_t row = per_receiver ? e / k : e;
        float4 v = s4[row * L4 + c];
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    dh[idx] = s;
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
int act_stride_for(const MlpDev& m, bool backward) {
    int w = m.hidden > m.out_dim ? m.hidden : m.out_dim;
    if (backward) {                        // gradient buffers also hold a block of <= 256 input columns
        int in_blk = m.in_dim < 256 ? m.in_dim : 256;
        if (in_blk > w) w = in_blk;
    }
    return w + 1;
}

size_t fwd_smem_bytes(const MlpDev& m) {
    int AS = act_stride_for(m, false);
    return sizeof(float) * (2 * TM_F * AS + TM_F * (KC + 1) + KC * WPAD + 2 * TM_F) + sizeof(int) * TM_F;
}
size_t bwd_smem_bytes(const MlpDev& m) {
    int AS = act_stride_for(m, true);
    return sizeof(float) * ((m.n_layers + 2) * TM_B * AS + TM_B * (KC + 1) + KC * WPAD + 2 * TM_B +
                            (m.n_layers + 2) * 256) + sizeof(int) * TM_B;
}

int validate(const cgnn_mlp* mlp, const char* who) {
    CGNN_CHECK_ARG(mlp != nullptr, "%s: null mlp", who);
    CGNN_CHECK_ARG(mlp->n_layers >= 1 && mlp->n_layers <= CGNN_MAX_LAYERS, "%s: n_layers must be in 1..%d", who, CGNN_MAX_LAYERS);
    CGNN_CHECK_ARG(mlp->in_dim >= 1 && mlp->hidden >= 1 && mlp->out_dim >= 1, "%s: bad widths", who);
    CGNN_CHECK_ARG(mlp->hidden <= 256 && mlp->out_dim <= 256, "%s: hidden/out width > 256 not supported (got %d/%d)", who, mlp->hidden, mlp->out_dim);
    for (int l = 0; l < mlp->n_layers; ++l) CGNN_CHECK_ARG(mlp->W[l] && mlp->b[l], "%s: null weight/bias at layer %d", who, l);
    CGNN_CHECK_ARG((mlp->ln_gamma == nullptr) == (mlp->ln_beta == nullptr), "%s: ln_gamma/ln_beta must both be set or both NULL", who);
    CGNN_CHECK_ARG(mlp->ln_dim >= 0 && mlp->ln_dim <= mlp->out_dim, "%s: ln_dim must be in 0..out_dim", who);
    return CGNN_OK;
}

MlpDev to_dev(const cgnn_mlp* mlp) {
    MlpDev m;
    m.n_layers = mlp->n_layers; m.in_dim = mlp->in_dim; m.hidden = mlp->hidden; m.out_dim = mlp->out_dim;
    for (int l = 0; l < CGNN_MAX_LAYERS; ++l) { m.W[l] = l < mlp->n_layers ? mlp->W[l] : nullptr; m.b[l] = l < mlp->n_layers ? mlp->b[l] : nullptr; }
    m.gamma = mlp->ln_gamma; m.beta = mlp->ln_beta;
    m.ln_dim = mlp->ln_dim > 0 ? mlp->ln_dim : mlp->out_dim;
    return m;
}

void fill_offsets(MlpTask& a) {
    const MlpDev& m = a.mlp;
    int64_t off = 0;
    for (int l = 0; l < m.n_layers; ++l) {
        const int K = l == 0 ? m.in_dim : m.hidden;
        const int N = l == m.n_layers - 1 ? m.out_dim : m.hidden;
        a.w_off[l] = off; off += (int64_t)N * K;
        a.b_off[l] = off; off += N;
    }
    a.g_off = off;
    if (m.gamma) off += 2 * m.out_dim;
    a.blob_size = off;
}

int max_bwd_ctas() { return num_sms(); }

int64_t tiles_for(const MlpTask& a, int tmr) {
    if (a.mode == MODE_EDGE) { int rpt = tmr / a.k; return (a.n + rpt - 1) / rpt; }
    return (a.n + tmr - 1) / tmr;
}

int launch_fwd(MlpTask& a, cudaStream_t stream) {
    a.act_stride = act_stride_for(a.mlp, false);
    size_t smem = fwd_smem_bytes(a.mlp);
    static size_t configured = 0;
    if (smem > configured) {
        CGNN_CUDA(cudaFuncSetAttribute(mlp_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    int64_t tiles = tiles_for(a, TM_F);
    if (tiles == 0) return CGNN_OK;
    int64_t cap = (int64_t)num_sms() * 4;
    int grid = (int)(tiles < cap ? tiles : cap);
    mlp_fwd_kernel<<<grid, NTHREADS, smem, stream>>>(a, tiles);
    CGNN_LAUNCH_CHECK();
    return CGNN_OK;
}

int launch_bwd(MlpTask& a, const cgnn_mlp_grad* grad, void* workspace, int64_t workspace_bytes, cudaStream_t stream) {
    a.act_stride = act_stride_for(a.mlp, true);
    fill_offsets(a);
    size_t smem = bwd_smem_bytes(a.mlp);
    CGNN_CHECK_ARG(smem <= 227 * 1024, "mlp backward: shared memory need %zu exceeds 227 KB", smem);
    static size_t configured = 0;
    if (smem > configured) {
        CGNN_CUDA(cudaFuncSetAttribute(mlp_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    int64_t tiles = tiles_for(a, TM_B);
    CGNN_CHECK_ARG(tiles > 0, "mlp backward: empty input");
    int cap = max_bwd_ctas();
    int grid = (int)(tiles < cap ? tiles : cap);
    int64_t need = (int64_t)grid * a.blob_size * (int64_t)sizeof(float);
    if (workspace == nullptr || workspace_bytes < need) {
        set_error("mlp backward: workspace too small (%lld < %lld)", (long long)workspace_bytes, (long long)need);
        return CGNN_ERR_WORKSPACE;
    }
    float* partials = static_cast<float*>(workspace);
    mlp_bwd_kernel<<<grid, NTHREADS, smem, stream>>>(a, tiles, partials, a.blob_size);
    CGNN_LAUNCH_CHECK();
    GradPtrs gp;
    for (int l = 0; l < CGNN_MAX_LAYERS; ++l) { gp.W[l] = grad ? grad->W[l] : nullptr; gp.b[l] = grad ? grad->b[l] : nullptr; }
    gp.gamma = grad ? grad->ln_gamma : nullptr; gp.beta = grad ? grad->ln_beta : nullptr;
    int rb = (int)((a.blob_size + 255) / 256);
    reduce_partials_kernel<<<rb, 256, 0, stream>>>(partials, grid, a.blob_size, gp, a.mlp, a.w_off[0], a.w_off[1],
                                                   a.w_off[2], a.w_off[3], a.b_off[0], a.b_off[1], a.b_off[2],
                                                   a.b_off[3], a.g_off);
    CGNN_LAUNCH_CHECK();
    return CGNN_OK;
}

}  // namespace

// exported to the dispatcher (api.cu)
int simt_mlp_fwd(MlpTask& a, cudaStream_t s) { return launch_fwd(a, s); }
int simt_mlp_bwd(MlpTask& a, const cgnn_mlp_grad* g, void* ws, int64_t wsb, cudaStream_t s) { return launch_bwd(a, g, ws, wsb, s); }
int64_t simt_mlp_bwd_workspace(const cgnn_mlp* mlp) {
    MlpTask a{};
    a.mlp = to_dev(mlp);
    fill_offsets(a);
    return (int64_t)max_bwd_ctas() * a.blob_size * (int64_t)sizeof(float);
}
int mlp_validate(const cgnn_mlp* mlp, const char* who) { return validate(mlp, who); }
MlpDev mlp_to_dev(const cgnn_mlp* mlp) { return to_dev(mlp); }

int simt_aggregate_senders(const float* h, const int32_t* senders, int64_t n, int k, int L, float* agg, cudaStream_t stream) {
    int L4 = L / 4;
    int64_t total = n * L4;
    aggregate_senders_kernel<<<(int)((total + 255) / 256), 256, 0, stream>>>(h, senders, n, k, L4, reinterpret_cast<float4*>(agg));
    CGNN_LAUNCH_CHECK();
    return CGNN_OK;
}
int simt_scatter_to_senders(const float* src, int per_receiver, const int32_t* rowptr, const int32_t* perm,
                            int64_t n, int k, int L, float* dh, cudaStream_t stream) {
    int L4 = L / 4;
    int64_t total = n * L4;
    scatter_to_senders_kernel<<<(int)((total + 255) / 256), 256, 0, stream>>>(src, per_receiver, rowptr, perm, n, k, L4,
                                                                             reinterpret_cast<float4*>(dh));
    CGNN_LAUNCH_CHECK();
    return CGNN_OK;
}

}  // namespace cgnn
