// Shared pieces of the tcgen05 kernels (mp_tc.cu: chained MLP tiles; wgrad_tc.cu: weight gradients).
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "mlp_common.cuh"
#include "tc_ptx.cuh"

namespace cgnn {

constexpr int TC_H = 128;                 // latent = hidden = out width handled by the tensor-core kernels
constexpr int CH = 16;                    // columns per streamed chunk (64-byte rows, SWIZZLE_64B)
constexpr int NCH = TC_H / CH;            // 8 chunks per 128-column tile
constexpr int CH_BYTES = 128 * CH * 4;    // 8192

__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    hi = ptx::pack_bf16x2(x0, x1);
    const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xFFFF0000u);
    lo = ptx::pack_bf16x2(x0 - h0, x1 - h1);
}

__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}

// Orders a warp's shared-memory LOADS before a later mbarrier arrival that hands the buffer back to a TMA producer: an empty
// volatile asm that READS the registers the loads feed (so every lane's loads have returned when it is passed); follow it by
// __syncwarp() and the elected lane's arrival.  Without it the compiler may sink the arithmetic on the loaded values below the
// arrival, the loads are then merely issued -- possibly queued behind scattered global accesses -- when the slot is refilled.
__device__ __forceinline__ void consume16(const uint32_t* r) {
    asm volatile("" ::"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                 "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}

// consume16 only pins the compiler's front end: where the loaded values feed arithmetic first (split2 ...) real instructions read
// them ahead of the arrival, but RAW loaded registers (a bfloat16 stream that goes to tcgen05.st / shared memory as it is, ring-fed
// row streams) are read by nothing the assembler has to keep in front of it -- the loads are then merely in flight when the slot is
// handed back, and the TMA refill lands under them (seen as a few wrong rows of one tile in the fifth tile of a CTA, only when the
// timing was tight).  fold_zero makes the arrival itself depend on them: it XORs the registers together with real instructions and
// masks the result with `z`, a run-time zero the assembler cannot know (callers pass the top bits of a 64-bit row count); the
// result (0) offsets the barrier address of the arrival.
template <int N>
__device__ __forceinline__ uint32_t fold_zero(const uint32_t* r, uint32_t z) {
    uint32_t x = r[0];
#pragma unroll
    for (int j = 1; j < N; ++j) x ^= r[j];
    return x & z;
}

// byte offset of the 16-byte piece j (0..3) of row r inside a 64-byte-swizzled chunk buffer
__device__ __forceinline__ uint32_t swz64(int r, int j) { return (uint32_t)(r * 64 + ((j ^ ((r >> 1) & 3)) << 4)); }

// 2-D tensor map over a row-major FP32 [rows][128] array: boxes of 128 rows x 16 columns, 64-byte swizzle
int make_row_map(CUtensorMap* m, const float* base, int64_t rows);
// boxes of 128 rows x 32 columns, 128-byte swizzle (the chain kernel's input ring)
int make_row_map32(CUtensorMap* m, const float* base, int64_t rows);
// boxes of box_rows rows x 32 columns, 128-byte swizzle
int make_row_map32_rows(CUtensorMap* m, const float* base, int64_t rows, int box_rows, int cols = TC_H);
// boxes of 1 row x 32 columns, 128-byte swizzle: the map of tile::gather4 loads (four arbitrary rows per instruction)
int make_gather_map32(CUtensorMap* m, const float* base, int64_t rows);
// bfloat16 [rows][128] array: boxes of box_rows rows x 64 columns, 128-byte swizzle (the 2-byte gradient stream)
int make_row_map_bf16(CUtensorMap* m, const void* base, int64_t rows, int box_rows);

// ---- building blocks used by the backward orchestration (defined in mp_tc.cu / wgrad_tc.cu) -----------
// B[n][k] = W[(row0 + n) * ld + col0 + k] (transpose: W[(row0 + k) * ld + col0 + n]); zero where n >= nmax or k >= kmax (0 = 128)
struct ChainBlock { const float* W; int ld; int row0; int col0; int transpose; int nmax; int kmax; };
struct ChainOp {
    int ns;                       // 1 (bf16) or 3 (bf16x3)
    int64_t rows;
    const float* in0;             // [rows][128]
    int in0_cols;                 // > 0: in0 is [rows][in0_cols] with in0_cols * 4 a multiple of 16 (the TMA zero-fills up to 128)
    const float* in1;             // second input phase (nullable)
    int n_layers;                 // 1 or 3
    ChainBlock blk[4];            // n_in + n_layers - 1 weight blocks
    const float* bias[3];         // per layer (nullable)
    const float* gamma; const float* beta;      // LayerNorm after the last layer (nullable)
    int ln_n;                     // LayerNorm width: the first ln_n of the 128 outputs (0 = all 128; the rest must be zero padding)
    int relu_out;                 // ReLU on the result (before mask / residual)
    int out_valid;                // > 0: the last layer has only this many outputs (bias zero-padded to 128)
    int k;                        // > 0: rows per receiver (gather and/or segmented sum)
    int k_valid;                  // 0 or the real in-degree below a padded k (rows of rank >= k_valid are dummies)
    const int32_t* senders; const float* Ps; const float* Pr;   // gather: layer-1 pre-activation += Ps[sender] + Pr[row / k]
    int64_t ps_rows;              // rows of Ps (the node table the senders index)
    const float* mask_src;        // result = mask_src > 0 ? result : 0   (nullable)
    const float* residual;        // result += residual                   (nullable)
    float* agg_out;               // [rows / k][128] = per-receiver sum of the result before the residual (nullable)
    float* out;                   // [rows][128]
    uint32_t* bits_out;           // [rows][4]: bit c = result column c > 0 (the ReLU gate, 16 B per row instead of a 512 B mask source; nullable)
    const uint32_t* mask_bits;    // [rows][4]: result = bit c ? result : 0  (nullable; applied after relu_out / mask_src)
    // hidden layers of a 3-layer chain (index 0, 1): mask instead of ReLU, extra output, per-receiver sum of the activation
    const float* hid_mask[2]; float* hid_out[2]; float* hid_agg[2];
    // LayerNorm backward instead of forward after the last layer: out = dY of (Y, dU), dU = du_rows[row] + du_recv[row / k];
    // d gamma / d beta (=|+=) their column sums (fixed-order reduction through ln_ws, ln_bwd_workspace_bytes())
    int ln_bwd; const float* du_rows; const float* du_recv; float* dgamma; float* dbeta; int accumulate; void* ln_ws;
    // the 2-byte gradient stream of the long-stream backward: in0 / out are bfloat16 [rows][128] arrays behind the float pointers
    // (one-layer chains; in16: single input, no gather; out16: no residual)
    int in16; int out16;
    // ... and the gradient stream de that is carried from processor step to processor step: du_rows (LayerNorm backward) / residual
    // are bfloat16 [rows][128]; both only through the ring-fed one-layer chains (C_T1)
    int du16; int res16;
};
int run_chain(const ChainOp& op, cudaStream_t stream);

// dW[n][col0 + c] (=|+=) sum_rows X[row][n] * A[row][c];  db[n] (=|+=) sum_rows X[row][n]   (deterministic)
int64_t wgrad_workspace_bytes();
// only the first nrows rows / ncols columns of the 128 x 128 product are written (0 = all 128)
// a_cols > 0: A is [rows][a_cols] (a_cols * 4 a multiple of 16), zero-filled up to 128 by the TMA
// x16: X is a bfloat16 [rows][128] array (the 2-byte gradient stream) -- taken as it is, no low part
int run_wgrad(int ns, const float* X, const float* A, int64_t rows, float* dW, int ld, int col0, float* db,
              int accumulate, void* ws, cudaStream_t stream, int nrows = 0, int ncols = 0, int a_cols = 0, int x16 = 0);


// dY = LayerNorm backward of (Y, dU), dU = (dU_rows ? dU_rows[row] : 0) + (dU_recv ? dU_recv[row / k] : 0); dY may alias Y
int64_t ln_bwd_workspace_bytes();
int run_ln_bwd(const float* Y, const float* dU_rows, const float* dU_recv, int k, const float* gamma, int64_t rows,
               float* dY, float* dgamma, float* dbeta, int accumulate, void* ws, cudaStream_t stream);

}  // namespace cgnn
