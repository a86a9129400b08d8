// Tensor-core (tcgen05) implementation of the fused MLP tiles (forward chains and the chains of the backward).
//
// One kernel, `tc_chain_fwd<NS>`, runs a chain of Linear layers over 256-row tiles on a CTA PAIR
// (cta_group::2, UMMA 256x128x16, BF16 operands, FP32 accumulators in TMEM):
//
//   * weights live in shared memory for the whole (persistent) kernel as BF16 images in the UMMA
//     K-major layout -- split by output row between the two CTAs of the pair, so one CTA holds
//     64 x 128 x {hi, lo} per weight block;
//   * activations never touch shared memory: the A operand of every layer is written to TMEM by the
//     epilogue warps (tcgen05.st, two BF16 per 32-bit column) and read from there by the MMA (the "TS"
//     form); the accumulator is read back with tcgen05.ld, one row per thread;
//   * NS = 3 ("bf16x3"): every FP32 value x is split into hi = bf16(x), lo = bf16(x - hi) and a product
//     is hi*hi + lo*hi + hi*lo (three MMAs, ~2^-17 relative error, FP32-like results);
//     NS = 1 ("bf16"): single pass;
//   * two tiles are in flight per CTA (TMEM columns [0,256) and [256,512)): while the tensor core works
//     on one, the four epilogue warps of the other do bias/ReLU/split or the final epilogue;
//   * the edge phase uses W1 = [W1s | W1r | W1e] (graph_network.py:89 concat order): P_s = h W1s^T and
//     P_r = h W1r^T + b1 are computed once per NODE by this same kernel (a 1-layer chain) and the
//     per-edge layer 1 is e W1e^T + P_s[sender] + P_r[receiver]  (5 L^2 -> 3 L^2 MACs per edge);
//   * the chain input streams in through TMA (2-D tensor maps, 128 rows x 32 columns, 128-byte swizzle);
//     EVERY OTHER row stream -- gathered P_s / P_r rows, ReLU-mask sources, residuals, upstream gradients,
//     and all outputs -- is read / written by the epilogue threads themselves in the "thread = row" mapping
//     with 256-bit global loads / stores (each lane touches one full 32-byte sector of its own 512-byte
//     row; measured 5.9 TB/s with 8 warps per SM, tools/micro/rowstream.cu), loads prefetched one 16-column
//     chunk ahead.  There is no shared-memory staging, no named barrier and no TMA store in the epilogue;
//   * per-receiver sums (k consecutive rows = k consecutive lanes) are a butterfly of warp shuffles that halves
//     the number of live values at each step (15 shuffles per 16 columns at k = 16), fixed order, deterministic.
//
// Reference semantics: InteractionNetwork.forward + residuals, graph_network.py:83-101,177-183, and its autograd.
#include <stdlib.h>

#include "tc_common.cuh"

namespace cgnn {
namespace {

using namespace ptx;

constexpr int TC_THREADS = 384;           // 8 epilogue warps (2 groups of 4) + a control warpgroup: producer warp, MMA warp, 2 idle
constexpr int TC_THREADS_SPLIT = 640;     // column-split chains: 16 epilogue warps (2 groups of 8: two warps share a lane quarter, 64 columns each)
// setmaxnreg: the control warpgroup hands its registers to the epilogue warps.  The pool is what the CTA was LAUNCHED with
// (registers per thread at launch x threads): 168 x 384 -> 40 / 232;  96 x 640 -> 24 / 112 (40 / 112 would ask for more than
// the pool holds and the epilogue warps would wait for registers forever).
constexpr int EPI_REGS = 232, CTL_REGS = 40, EPI_REGS_SPLIT = 104, CTL_REGS_SPLIT = 40;
constexpr int CW = 32;                    // columns per streamed input chunk (128-byte rows, SWIZZLE_128B)
constexpr int NCW = TC_H / CW;            // 4 chunks per 128-column tile
constexpr int CW_BYTES = 128 * CW * 4;    // 16384
constexpr int MAXRING = 12;               // input ring depth is chosen per launch from the shared memory left (2 .. 12)
constexpr int WIMG = 64 * TC_H * 2;       // bytes of one weight image half (64 output rows x 128 k, bf16)
constexpr int MAX_BLOCKS = 4;             // weight blocks (MMA phases) per tile
constexpr float LN_EPS = 1e-5f;

struct TcParams {
    int64_t n_rows;             // rows of the stream (edges or nodes)
    int64_t n_pair_tiles;       // ceil(n_rows / 256)
    int n_in;                   // input phases (1: in0; 2: in0 then in1)
    int n_layers;               // Linear layers (1 or 3)
    int k;                      // rows per receiver (gather / per-receiver sums), a power of two <= 32; 1 when unused
    int kshift;                 // log2(k)
    int k_valid;                // 0, or the real in-degree when the k rows of a receiver are padded up to the power of two: rows of
                                // rank >= k_valid are dummies (no part in per-receiver sums, no gradient)
    int ln_mode;                // 0: none, 1: LayerNorm forward, 2: LayerNorm backward (the result is dY)
    int ln_n;                   // LayerNorm width: columns [ln_n, 128) are zero padding (accumulator, bias, gamma, beta all zero there)
    int gather;                 // layer-1 pre-activation += Ps[sender] + Pr[receiver]
    int relu_out;               // ReLU on the result (before mask / residual)
    const float* mask_src;      // [n_rows][128]: result = mask_src > 0 ? result : 0 (nullable)
    const float* residual;      // [n_rows][128] added to the result (nullable)
    float* agg_out;             // [n_rows / k][128] per-receiver sum of the result before the residual (nullable)
    const int32_t* senders;     // gather
    const float* Ps;            // gather: [N][128]
    const float* Pr;            // gather: [N][128] (includes the layer-1 bias)
    // hidden layers of a 3-layer chain (index = hidden layer 0, 1)
    const float* hid_mask[2];   // activation = hid_mask > 0 ? v : 0 instead of ReLU (nullable)
    float* hid_out[2];          // [n_rows][128]: the activation is also written here (nullable)
    float* hid_agg[2];          // [n_rows / k][128]: per-receiver sum of the activation (nullable)
    // LayerNorm backward (ln_mode 2): dU = du_rows[row] + du_recv[row / k] (each nullable)
    const float* du_rows;
    const float* du_recv;
    float* ln_partials;         // [gridDim.x * 8][2][128]: per-warp column sums of dU * xhat and dU
    float* out;                 // [n_rows][128]
    uint32_t* bits_out;         // [n_rows][4]: bit c = (result column c > 0) -- the ReLU gate of the backward, 16 B per row (nullable)
    const uint32_t* mask_bits;  // [n_rows][4]: result = bit c ? result : 0 -- the same gate read back (nullable)
    int n_ring;                 // input ring depth
    int in16;                   // the chain input is a bfloat16 [n_rows][128] stream (the 2-byte gradient stream of the long-stream backward):
                                // two 64-column ring slots per tile, written to the A operand as they are -- no low part, two MMA passes
    int out16;                  // `out` is a bfloat16 [n_rows][128] stream (round to nearest even)
    int t116;                   // C_T1: the ring-fed epilogue stream (dU rows / residual: the gradient stream de) is bfloat16 -- two
                                // 64-column slots per tile instead of four 32-column ones
    ChainBlock blk[MAX_BLOCKS]; // weight blocks in FP32 (torch layout): every CTA builds its BF16 images in shared memory itself
    const float* vec_src[5];    // bias of layer 1, 2, 3, gamma, beta (nullable -> zeros / ones for gamma)
    int vec_len[5];             // valid entries (zero padded to 128)
    unsigned long long* stamps; // debug (cgnn_debug_stamps): clock64 of CTA 0's groups, [2][stamp_tiles][16] (nullable)
    int stamp_tiles;
};

// debug hook: when set, the next chain launches record the stage time stamps of block 0
unsigned long long* g_stamps = nullptr;
int g_stamp_tiles = 0, g_stamp_launches = 0, g_stamp_next = 0;

// shared-memory layout (dynamic, 1024-byte aligned base)
struct Smem {
    static constexpr int vec = 0;                                  // 5 * 128 floats
    static constexpr int bars = vec + 5 * TC_H * 4;                // barriers (512 bytes)
    static constexpr int lnacc = bars + 512;                       // LayerNorm-backward variants only: 8 warps * 2 * 128 floats (column sums)
    __host__ __device__ static constexpr int weights(bool lnb) { return lnacc + (lnb ? 8 * 2 * TC_H * 4 : 0); }   // n_blocks * NSI * WIMG
    // then (1024-aligned) n_ring * CW_BYTES input ring
};
static_assert(Smem::weights(false) % 128 == 0 && Smem::weights(true) % 128 == 0, "weight images need 128-byte alignment");
__host__ __device__ constexpr uint32_t ring_offset(int n_blocks, int nsi, bool lnb) {
    return (uint32_t)((Smem::weights(lnb) + n_blocks * nsi * WIMG + 1023) / 1024 * 1024);
}

struct Bars {
    // "full" barriers are per consumer group: a group only ever waits on barriers whose uses are all its own,
    // so the phase parity it tracks can never alias a phase that belongs to the other group's tiles
    uint64_t in_full[2][MAXRING], in_empty[MAXRING];
    uint64_t a_ready[2];        // used in the leader CTA: 8 arrivals (4 warps x 2 CTAs)
    uint64_t mma_done[2];       // per CTA, one tcgen05.commit arrival
    uint32_t tmem_base;
};
static_assert(sizeof(Bars) <= 512, "barrier block");

// byte offset of the 16-byte piece j (0..7) of row r inside a 128-byte-swizzled chunk buffer
__device__ __forceinline__ uint32_t swz128(int r, int j) { return (uint32_t)(r * 128 + ((j ^ (r & 7)) << 4)); }

// 256-bit global accesses: one full 32-byte sector per lane
__device__ __forceinline__ void ld256(const float* p, float* v) {
    asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p));
}
__device__ __forceinline__ void st256(float* p, const float* v) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}
// 16 columns as bfloat16 (round to nearest even): one 32-byte sector of the row's 256 bytes
__device__ __forceinline__ void st16h(uint16_t* p, const float* v) {
    uint32_t w[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) w[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
}
__device__ __forceinline__ void ld16(const float* p, float* v) { ld256(p, v); ld256(p + 8, v + 8); }
__device__ __forceinline__ void st16(float* p, const float* v) { st256(p, v); st256(p + 8, v + 8); }

// tcgen05.wait::ld that also carries a data dependency on the 16 registers of the load it completes, so the
// compiler cannot move their uses above the wait when loads are software-pipelined
__device__ __forceinline__ void tmem_ld_wait16(float* v) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]),
                   "+f"(v[8]), "+f"(v[9]), "+f"(v[10]), "+f"(v[11]), "+f"(v[12]), "+f"(v[13]), "+f"(v[14]), "+f"(v[15])
                 :: "memory");
}

// Sums 16 columns over the k lanes (rows) of a receiver, k a power of two <= 32 (m = k/2, k/4, ... 1 lane distance):
// every step exchanges half of the live values with the partner lane, so 16 -> 8 -> 4 -> 2 -> 1 values stay live.
// On return v[0 .. nv) hold complete sums, nv = max(16 / k, 1): columns (lane & (k-1)) * nv + j for k <= 16, column
// (lane & 31) >> 1 for k = 32 (both lanes of a pair hold it).  Fixed tree order: deterministic.
__device__ __forceinline__ void receiver_sum16(float* v, int k, int lane) {
#pragma unroll
    for (int s = 0; s < 5; ++s) {
        const int m = k >> (s + 1);
        if (m >= 1) {
            const bool upper = (lane & m) != 0;
            if (s < 4) {
                const int half = 8 >> s;
#pragma unroll
                for (int j = 0; j < half; ++j) {
                    const float send = upper ? v[j] : v[j + half];
                    const float keep = upper ? v[j + half] : v[j];
                    v[j] = keep + __shfl_xor_sync(0xffffffffu, send, m);
                }
            } else {
                v[0] += __shfl_xor_sync(0xffffffffu, v[0], m);
            }
        }
    }
}
// Two 16-column blocks at once: the same butterfly on both, level by level, so that the two dependency chains (select ->
// shuffle -> add, four levels at k = 16) overlap in a warp that has nobody else to hide its latencies behind.
__device__ __forceinline__ void receiver_sum16x2(float* v, float* w, int k, int lane) {
#pragma unroll
    for (int s = 0; s < 5; ++s) {
        const int m = k >> (s + 1);
        if (m >= 1) {
            const bool upper = (lane & m) != 0;
            if (s < 4) {
                const int half = 8 >> s;
                float sv[8], sw[8];
#pragma unroll
                for (int j = 0; j < half; ++j) {
                    sv[j] = __shfl_xor_sync(0xffffffffu, upper ? v[j] : v[j + half], m);
                    sw[j] = __shfl_xor_sync(0xffffffffu, upper ? w[j] : w[j + half], m);
                }
#pragma unroll
                for (int j = 0; j < half; ++j) {
                    v[j] = (upper ? v[j + half] : v[j]) + sv[j];
                    w[j] = (upper ? w[j + half] : w[j]) + sw[j];
                }
            } else {
                const float a = __shfl_xor_sync(0xffffffffu, v[0], m), b = __shfl_xor_sync(0xffffffffu, w[0], m);
                v[0] += a;
                w[0] += b;
            }
        }
    }
}
// stores the sums receiver_sum16 left in v for columns [c, c + 16) of receiver row `dst`
__device__ __forceinline__ void receiver_store16(float* dst, int c, const float* v, int k, int lane) {
    if (k == 32) {
        if ((lane & 1) == 0) dst[c + (lane >> 1)] = v[0];
    } else if (k == 16) {
        dst[c + (lane & 15)] = v[0];
    } else if (k == 8) {
        *reinterpret_cast<float2*>(dst + c + (lane & 7) * 2) = make_float2(v[0], v[1]);
    } else if (k == 4) {
        *reinterpret_cast<float4*>(dst + c + (lane & 3) * 4) = make_float4(v[0], v[1], v[2], v[3]);
    } else if (k == 2) {
        float* d = dst + c + (lane & 1) * 8;
        *reinterpret_cast<float4*>(d) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(d + 4) = make_float4(v[4], v[5], v[6], v[7]);
    } else {
#pragma unroll
        for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(dst + c + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
}

// CFG: compile-time shape of the chain, so that every instantiation carries only the code and the registers (row-stream
// buffers) it uses: C_L3 three layers (else one), C_LN / C_LNB LayerNorm forward / backward after the last layer,
// C_SA / C_SB the final epilogue reads row stream a (P_s | dU rows | mask source) / b (P_r | dU per receiver |
// residual), C_HS the hidden epilogues read row streams (gather or mask sources), C_AGG per-receiver sums, C_RIN the
// residual is the chain's own input: its tile stays in the input ring until the output pass instead of being re-read.
// C_GIN: the gather P_s[sender] + P_r[receiver] INITIALISES the layer-1 accumulator (written to TMEM in the input phase, the
// MMAs then accumulate onto it) instead of being added in an epilogue.  C_SPLIT: 16 epilogue warps -- the two warps that
// share a TMEM lane quarter take 64 columns each (LayerNorm moments combined through two spare TMEM columns).
// C_T1: one epilogue row stream of a one-layer chain (the dU rows of the LayerNorm backward, or the residual) comes through the
// TMA ring as a second stream of the tile (tensor map 1) instead of thread = row global loads: 16 KB per instruction, requested by
// the producer tiles ahead, where the scattered 32-byte sectors of the thread = row loads were the slowest part of those chains.
// C_G4 (with C_GIN): the gathered P_s rows come through the ring too -- the producer WARP issues tile::gather4 loads (lane l the
// rows 4l..4l+3 of the tile, 32 columns per slot, tensor map 1) ahead of the tile's input chunks.
constexpr int C_L3 = 1, C_LN = 2, C_LNB = 4, C_SA = 8, C_SB = 16, C_HS = 32, C_AGG = 64, C_RIN = 128, C_GIN = 256, C_SPLIT = 512, C_T1 = 1024,
              C_G4 = 2048;
template <int NS, int CFG>
__global__ void __launch_bounds__((CFG & C_SPLIT) ? TC_THREADS_SPLIT : TC_THREADS, 1)
tc_chain_fwd(const __grid_constant__ CUtensorMap tm_in0, const __grid_constant__ CUtensorMap tm_in1, const TcParams p) {
    constexpr int NSI = NS == 3 ? 2 : 1;                    // weight / activation images per value (hi [, lo])
    constexpr bool L3 = (CFG & C_L3) != 0, SA = (CFG & C_SA) != 0, SB = (CFG & C_SB) != 0, HS = (CFG & C_HS) != 0, AGG = (CFG & C_AGG) != 0;
    constexpr bool RIN = (CFG & C_RIN) != 0, GIN = (CFG & C_GIN) != 0, SPLIT = (CFG & C_SPLIT) != 0, T1 = (CFG & C_T1) != 0;
    static_assert(!T1 || (!(CFG & C_L3) && !RIN && !SPLIT), "the ring-fed epilogue stream belongs to the one-layer chains");
    constexpr bool G4 = (CFG & C_G4) != 0;
    // the bfloat16 row streams (p.in16 / p.out16) exist in the one-layer chains only: no trace of them in the fused forward kernels
    constexpr bool IO16 = !(CFG & C_L3) && !RIN && !SPLIT && !G4 && !GIN;
    static_assert(!G4 || (GIN && !T1 && !RIN && !SPLIT), "gather4 feeds the accumulator initialisation; map 1 and the ring must be free for it");
    constexpr int NEPI = SPLIT ? 16 : 8;                    // epilogue warps
    constexpr int GW = NEPI / 2;                            // warps per epilogue group (= per tile in flight)
    constexpr int NTHR = (NEPI + 4) * 32;
    constexpr int NC = SPLIT ? 64 : TC_H;                   // accumulator columns per epilogue thread
    static_assert(!SPLIT || !(CFG & (C_LNB | C_HS | C_SA)), "the column split covers the forward chains only");
    constexpr int LN = (CFG & C_LNB) ? 2 : ((CFG & C_LN) ? 1 : 0);
    constexpr bool LNB = LN == 2;
    constexpr int N_LAYERS = L3 ? 3 : 1;
    extern __shared__ __align__(1024) uint8_t smem[];
    Bars* bars = reinterpret_cast<Bars*>(smem + Smem::bars);
    float* sVec = reinterpret_cast<float*>(smem + Smem::vec);
    const int n_blocks = p.n_in + N_LAYERS - 1;             // MMA phases per tile
    uint8_t* sW = smem + Smem::weights(LNB);
    uint8_t* sRing = smem + ring_offset(n_blocks, NSI, LNB);
    const int NRING = p.n_ring;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const int64_t cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    const int64_t n_it = p.n_pair_tiles > cluster_id ? (p.n_pair_tiles - cluster_id + n_clusters - 1) / n_clusters : 0;

    // ---- one-time setup ---------------------------------------------------------------------------
    if (tid == 0) {
        for (int i = 0; i < MAXRING; ++i) {
            mbar_init(&bars->in_full[0][i], 1);
            mbar_init(&bars->in_full[1][i], 1);
            mbar_init(&bars->in_empty[i], GW);              // every warp of the group hands a slot back (owner or not)
        }
        for (int s = 0; s < 2; ++s) { mbar_init(&bars->a_ready[s], 2 * GW); mbar_init(&bars->mma_done[s], 1); }
        fence_mbar_init();
    }
    for (int i = tid; i < 5 * TC_H; i += NTHR) {
        const int v = i / TC_H, c = i % TC_H;
        // (gamma defaults to one where there is no LayerNorm weight at all, and is zero on padded columns)
        sVec[i] = p.vec_src[v] != nullptr ? (c < p.vec_len[v] ? __ldg(p.vec_src[v] + c) : 0.0f) : (v == 3 ? 1.0f : 0.0f);
    }
    // This CTA's half (64 output rows) of every weight block as K-major BF16 image(s):  B[n][k] = W[(row0 + n) * ld + col0 + k]
    // (transposed: W[(row0 + k) * ld + col0 + n]), element (n, k) at (n / 8) * 2048 + (k / 8) * 128 + (n % 8) * 16 + (k % 8) * 2.
    // 16-byte pieces (n, k8); 14 MB of L2 reads per launch instead of a separate preparation kernel.
    for (int piece = tid; piece < n_blocks * 64 * (TC_H / 8); piece += NTHR) {
        const int b = piece / (64 * (TC_H / 8)), rem = piece % (64 * (TC_H / 8));
        const int nl = rem / (TC_H / 8), k8 = rem % (TC_H / 8);
        const int n = (int)rank * 64 + nl;
        const ChainBlock& B = p.blk[b];
        const int nmax = B.nmax ? B.nmax : TC_H, kmax = B.kmax ? B.kmax : TC_H;
        float x[8];
        if (!B.transpose && n < nmax && k8 * 8 + 8 <= kmax && ((B.ld | B.col0) & 3) == 0) {
            const float4* src = reinterpret_cast<const float4*>(B.W + (size_t)(B.row0 + n) * B.ld + B.col0 + k8 * 8);
            const float4 u0 = __ldg(src), u1 = __ldg(src + 1);
            x[0] = u0.x; x[1] = u0.y; x[2] = u0.z; x[3] = u0.w; x[4] = u1.x; x[5] = u1.y; x[6] = u1.z; x[7] = u1.w;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int k = k8 * 8 + j;
                const bool ok = n < nmax && k < kmax;
                x[j] = !ok ? 0.0f : B.transpose ? __ldg(B.W + (size_t)(B.row0 + k) * B.ld + B.col0 + n) : __ldg(B.W + (size_t)(B.row0 + n) * B.ld + B.col0 + k);
            }
        }
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) split2(x[2 * j], x[2 * j + 1], hi[j], lo[j]);
        const uint32_t off = (uint32_t)((nl >> 3) * (TC_H * 16) + k8 * 128 + (nl & 7) * 16);
        *reinterpret_cast<uint4*>(sW + (b * NSI) * WIMG + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        if (NS == 3) *reinterpret_cast<uint4*>(sW + (b * NSI + 1) * WIMG + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
    fence_proxy_async_smem();                                // the MMAs (async proxy, both CTAs of the pair) read these images
    if (LNB)
        for (int i = tid; i < 8 * 2 * TC_H; i += NTHR) reinterpret_cast<float*>(smem + Smem::lnacc)[i] = 0.0f;
    if (warp == NEPI + 1) tmem_alloc<2>(&bars->tmem_base, 512);
    if (warp == NEPI && lane == 0) {
        prefetch_tmap(&tm_in0);
        prefetch_tmap(&tm_in1);
    }
    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after_sync();
    const uint32_t tmem = bars->tmem_base;

    if (warp >= NEPI) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(SPLIT ? CTL_REGS_SPLIT : CTL_REGS));
    if (warp == NEPI) {
        // ============================ producer: weights, input chunks ============================================
        if (G4 || lane == 0) {                         // (G4: the whole warp walks the ring, lane 0 alone issues the tiled loads)
            uint32_t buf = 0, use = 0;                 // ring slot of the next chunk and how often it has been used
            for (int64_t it = 0; it < n_it; ++it) {
                const int64_t row0 = ((cluster_id + it * n_clusters) * 2 + rank) * 128;
                if (G4) {
                    // the tile's gathered rows first (the accumulator is initialised before the input is converted)
                    int sidx[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int64_t row = row0 + 4 * lane + j;
                        sidx[j] = __ldg(p.senders + (row < p.n_rows ? row : p.n_rows - 1));
                    }
                    for (int q = 0; q < NCW; ++q) {
                        mbar_wait_or_trap(&bars->in_empty[buf], (use & 1) ^ 1, 100 + buf);
                        uint64_t* full = &bars->in_full[it & 1][buf];
                        if (lane == 0) mbar_expect_tx(full, CW_BYTES);
                        __syncwarp();
                        tma_gather4_2d(sRing + buf * CW_BYTES + lane * 512, &tm_in1, q * CW, sidx[0], sidx[1], sidx[2], sidx[3], full);
                        if (++buf == (uint32_t)NRING) { buf = 0; ++use; }
                    }
                }
                for (int ip = 0; ip < p.n_in + (T1 ? 1 : 0); ++ip) {    // T1: the epilogue stream follows the MMA input(s), map 1
                    // a bfloat16 input has 64 columns per slot (the same 128-byte rows): two slots per tile
                    const bool in16 = IO16 && (ip < p.n_in ? p.in16 : p.t116);
                    const int nq = in16 ? NCW / 2 : NCW, qcols = in16 ? 2 * CW : CW;
                    for (int q = 0; q < nq; ++q) {
                        if (lane == 0) {
                            mbar_wait_or_trap(&bars->in_empty[buf], (use & 1) ^ 1, 100 + buf);
                            uint64_t* full = &bars->in_full[it & 1][buf];
                            mbar_expect_tx(full, CW_BYTES);
                            tma_load_2d(sRing + buf * CW_BYTES, ip == 0 ? &tm_in0 : &tm_in1, q * qcols, (int)row0, full);
                        }
                        if (++buf == (uint32_t)NRING) { buf = 0; ++use; }
                    }
                }
                if (G4) __syncwarp();
            }
        }
    } else if (warp == NEPI + 1) {
        // ============================ MMA issuer (leader CTA, one thread) =======================================
        if (rank == 0 && lane == 0) {
            const uint32_t idesc = umma_idesc_bf16(256, TC_H);
            const uint32_t w_base = smem_u32(sW);
            int64_t it_s[2] = {0, 1};
            int ph_s[2] = {0, 0};
            uint32_t par[2] = {0, 0};
            int active = (n_it > 0) + (n_it > 1);
            uint32_t spins = 0;
            while (active > 0) {
                for (int s = 0; s < 2; ++s) {
                    if (it_s[s] >= n_it) continue;
                    // non-blocking probe: try_wait may suspend the thread for a while when the phase is not complete, and this
                    // thread serves TWO barriers -- the other slot's operands may become ready in the meantime
                    if (!mbar_test_wait(&bars->a_ready[s], par[s])) {
                        if (++spins > (1u << 28)) { printf("cgnn: MMA issuer timeout slot %d it %lld ph %d\n", s, (long long)it_s[s], ph_s[s]); __trap(); }
                        continue;
                    }
                    spins = 0;
                    tc_fence_after_sync();
                    const int ph = ph_s[s];
                    const uint32_t d = tmem + s * 256;
                    const uint32_t a_hi = d + 128, a_lo = d + 192;
                    const uint32_t w_hi = w_base + (ph * NSI) * WIMG, w_lo = w_hi + WIMG;
                    // later input phases accumulate into layer 1; with GIN the accumulator already holds P_s[sender] + P_r[receiver]
                    uint32_t acc = ((ph > 0 && ph < p.n_in) || (GIN && ph == 0)) ? 1u : 0u;
                    // rolled loops: the descriptors advance by one K step (256 bytes of image, 8 TMEM columns) per MMA
                    const uint64_t dw_hi = umma_desc(w_hi, 128, TC_H * 16), dw_lo = umma_desc(w_lo, 128, TC_H * 16);
#pragma unroll 1
                    for (int ks = 0; ks < TC_H / 16; ++ks) {
                        umma_bf16_ts<2>(d, a_hi + ks * 8, dw_hi + (uint64_t)(ks * 16), idesc, acc);
                        acc = 1u;
                    }
                    if (NS == 3) {
                        if (!(IO16 && p.in16)) {           // (a bfloat16 input has no low part)
#pragma unroll 1
                            for (int ks = 0; ks < TC_H / 16; ++ks)
                                umma_bf16_ts<2>(d, a_lo + ks * 8, dw_hi + (uint64_t)(ks * 16), idesc, 1u);
                        }
#pragma unroll 1
                        for (int ks = 0; ks < TC_H / 16; ++ks)
                            umma_bf16_ts<2>(d, a_hi + ks * 8, dw_lo + (uint64_t)(ks * 16), idesc, 1u);
                    }
                    umma_commit<2>(&bars->mma_done[s]);
                    par[s] ^= 1u;
                    if (++ph_s[s] == n_blocks) {
                        ph_s[s] = 0;
                        it_s[s] += 2;
                        if (it_s[s] >= n_it) --active;
                    }
                }
            }
        }
    }
    } else {
        // ============================ epilogue groups: thread = row ==============================================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(SPLIT ? EPI_REGS_SPLIT : EPI_REGS));
        const uint32_t zero_rt = (uint32_t)((unsigned long long)p.n_rows >> 62);     // 0, but not to the assembler (fold_zero)
        const int g = warp / GW;                        // group = TMEM slot
        const int wq = warp & 3;                        // lane quarter this warp may touch
        const int half = SPLIT ? ((warp >> 2) & 1) : 0; // which 64 columns of the row this thread owns (column split)
        const int C0 = half * 64;                       // first owned accumulator column
        const int r = wq * 32 + lane;                   // row inside the CTA's 128-row tile
        const int gt = tid - g * (GW * 32);             // thread index inside the group
        const uint32_t tslot = tmem + ((uint32_t)(wq * 32) << 16) + g * 256;
        const uint32_t tD = tslot, tAhi = tslot + 128, tAlo = tslot + 192;
        float* lnacc = reinterpret_cast<float*>(smem + Smem::lnacc) + warp * 2 * TC_H;
        uint32_t pm = 0;                                // parity of mma_done[g]
        uint32_t in_par = 0;                            // bit b: parity of this group's next use of in_full[g][b]
        const int k = p.k;
        const bool stamping = p.stamps != nullptr && blockIdx.x == 0 && gt == 0;
#define CGNN_STAMP(slot)                                                                               \
    do {                                                                                               \
        if (stamping && (it >> 1) < p.stamp_tiles) p.stamps[((size_t)g * p.stamp_tiles + (it >> 1)) * 16 + (slot)] = clock64(); \
    } while (0)

        // ring slot of this group's next tile's first chunk: advanced by the chunks of two tiles (its own and the other group's) per
        // iteration, wrapped by subtraction -- no division in the loops
        const uint32_t uring = (uint32_t)NRING, tile_chunks = (uint32_t)((IO16 && p.in16 ? NCW / 2 : NCW) * p.n_in + (T1 ? (IO16 && p.t116 ? NCW / 2 : NCW) : 0) + (G4 ? NCW : 0));
        uint32_t ring0 = (uint32_t)g * tile_chunks % uring;
        // sender index of the first tile's row; the next tile's is fetched one tile ahead so its latency never shows
        int32_t snd_next = 0;
        if (GIN && g < n_it) {
            const int64_t first = ((cluster_id + (int64_t)g * n_clusters) * 2 + rank) * 128 + r;
            snd_next = __ldg(p.senders + (first < p.n_rows ? first : p.n_rows - 1));
        }
        for (int64_t it = g; it < n_it; it += 2) {
            const int64_t row0 = ((cluster_id + it * n_clusters) * 2 + rank) * 128;
            uint32_t buf = ring0;                   // ring slot of the next input chunk of this tile
            const uint32_t rin0 = ring0;            // RIN: slot of this tile's first input chunk
            ring0 += 2 * tile_chunks;
            while (ring0 >= uring) ring0 -= uring;
            const bool valid = row0 + r < p.n_rows;
            const int64_t grow = valid ? row0 + r : p.n_rows - 1;           // clamped: loads of padded rows stay in bounds
            const int64_t recv = grow >> p.kshift;
            const bool dummy = p.k_valid > 0 && (int)(grow & (int64_t)(k - 1)) >= p.k_valid;     // padded edge of a non-power-of-two in-degree
            const size_t rowoff = (size_t)grow * TC_H, recvoff = (size_t)recv * TC_H;
            // the sender index is fetched now so that the address of the gathered P_s row is ready when its epilogue starts
            const int32_t snd = snd_next;
            if (GIN && it + 2 < n_it) {
                const int64_t nxt = ((cluster_id + (it + 2) * n_clusters) * 2 + rank) * 128 + r;
                snd_next = __ldg(p.senders + (nxt < p.n_rows ? nxt : p.n_rows - 1));
            }
            CGNN_STAMP(0);
            // ---- input phases: stream chunks, split, write the A operand -------------------------------------
            if (GIN) {
                // layer-1 accumulator := P_s[sender] + P_r[receiver] (FP32, this thread's row and columns); the MMAs of the first
                // block accumulate e W1e^T onto it.  The scattered P_s loads of a whole column range are issued together, so
                // their latency is paid once per tile, here, where the thread would otherwise wait for its input chunks.
                const float* ps = p.Ps + (size_t)snd * TC_H + C0;
                const float* pr = p.Pr + recvoff + C0;
#pragma unroll 1
                for (int c = 0; c < NC; c += 64) {
                    float a0[16], a1[16], a2[16], a3[16];
                    if (G4) {
                        // the gathered rows of this tile sit in the next ring slots (32 columns each), row r at its swizzled place
#pragma unroll
                        for (int qq = 0; qq < 2; ++qq, buf = buf + 1 == uring ? 0 : buf + 1) {
                            mbar_wait_or_trap(&bars->in_full[g][buf], (in_par >> buf) & 1u, 130 + buf);
                            in_par ^= 1u << buf;
                            const uint8_t* src = sRing + buf * CW_BYTES;
                            float* lo16 = qq == 0 ? a0 : a2;
                            float* hi16 = qq == 0 ? a1 : a3;
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float4 u = *reinterpret_cast<const float4*>(src + swz128(r, j));
                                const float4 w = *reinterpret_cast<const float4*>(src + swz128(r, 4 + j));
                                lo16[4 * j] = u.x; lo16[4 * j + 1] = u.y; lo16[4 * j + 2] = u.z; lo16[4 * j + 3] = u.w;
                                hi16[4 * j] = w.x; hi16[4 * j + 1] = w.y; hi16[4 * j + 2] = w.z; hi16[4 * j + 3] = w.w;
                            }
                            const uint32_t fz = fold_zero<16>(reinterpret_cast<const uint32_t*>(hi16), zero_rt) ^
                                                fold_zero<16>(reinterpret_cast<const uint32_t*>(lo16), zero_rt);
                            __syncwarp();
                            if (lane == 0) mbar_arrive_local(&bars->in_empty[buf] + fz);
                        }
                    } else {
                        ld16(ps + c, a0); ld16(ps + c + 16, a1); ld16(ps + c + 32, a2); ld16(ps + c + 48, a3);
                    }
#pragma unroll
                    for (int hh = 0; hh < 4; ++hh) {
                        float* ca = hh == 0 ? a0 : hh == 1 ? a1 : hh == 2 ? a2 : a3;
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {       // the P_r row is shared by the k lanes of a receiver: cached loads
                            const float4 b = __ldg(reinterpret_cast<const float4*>(pr + c + 16 * hh + j));
                            ca[j] += b.x; ca[j + 1] += b.y; ca[j + 2] += b.z; ca[j + 3] += b.w;
                        }
                        tmem_st_32x32b_x16(tD + C0 + c + 16 * hh, reinterpret_cast<const uint32_t*>(ca));
                    }
                }
            }
            for (int ip = 0; ip < p.n_in; ++ip) {
                if (ip > 0) {                            // A is still being read by the previous phase's MMA
                    mbar_wait_or_trap(&bars->mma_done[g], pm, 140);
                    pm ^= 1u;
                    tc_fence_after_sync();
                }
                if (IO16 && p.in16) {
                    // bfloat16 input: a slot holds 64 columns of this thread's row (128 bytes, swizzled like the FP32 chunks) and
                    // they ARE the A operand's 32 words
                    for (int q = 0; q < NCW / 2; ++q, buf = buf + 1 == uring ? 0 : buf + 1) {
                        mbar_wait_or_trap(&bars->in_full[g][buf], (in_par >> buf) & 1u, 150 + buf);
                        in_par ^= 1u << buf;
                        const uint8_t* src = sRing + buf * CW_BYTES;
                        uint32_t w0[16], w1[16];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const uint4 u = *reinterpret_cast<const uint4*>(src + swz128(r, j));
                            const uint4 v = *reinterpret_cast<const uint4*>(src + swz128(r, 4 + j));
                            w0[4 * j] = u.x; w0[4 * j + 1] = u.y; w0[4 * j + 2] = u.z; w0[4 * j + 3] = u.w;
                            w1[4 * j] = v.x; w1[4 * j + 1] = v.y; w1[4 * j + 2] = v.z; w1[4 * j + 3] = v.w;
                        }
                        // the raw words go to TMEM first (a real reader of every loaded register), then the slot goes back
                        tmem_st_32x32b_x16(tAhi + q * 32, w0);
                        tmem_st_32x32b_x16(tAhi + q * 32 + 16, w1);
                        const uint32_t fz = fold_zero<16>(w0, zero_rt) ^ fold_zero<16>(w1, zero_rt);
                        __syncwarp();
                        if (lane == 0) mbar_arrive_local(&bars->in_empty[buf] + fz);
                    }
                } else
                for (int q = 0; q < NCW; ++q, buf = buf + 1 == uring ? 0 : buf + 1) {
                    // (every thread waits on every chunk's barrier, also on chunks of the other column half: a parity wait
                    //  may never fall two phases behind)
                    mbar_wait_or_trap(&bars->in_full[g][buf], (in_par >> buf) & 1u, 150 + buf);
                    in_par ^= 1u << buf;
                    if (SPLIT && (q >> 1) != half) {     // not this warp's columns: the slot goes straight back
                        __syncwarp();
                        if (lane == 0) mbar_arrive_local(&bars->in_empty[buf]);
                        continue;
                    }
                    const uint8_t* src = sRing + buf * CW_BYTES;
                    uint32_t hi[16], lo[16];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 v = *reinterpret_cast<const float4*>(src + swz128(r, j));
                        split2(v.x, v.y, hi[2 * j], lo[2 * j]);
                        split2(v.z, v.w, hi[2 * j + 1], lo[2 * j + 1]);
                    }
                    if (!(RIN && ip == 0)) {             // (RIN: the chunk is released by the output pass, which reads it as the residual)
                        consume16(hi);                   // the loads of every lane have returned before the slot is handed back
                        __syncwarp();
                        if (lane == 0) mbar_arrive_local(&bars->in_empty[buf]);
                    }
                    tmem_st_32x32b_x16(tAhi + q * 16, hi);
                    if (NS == 3) tmem_st_32x32b_x16(tAlo + q * 16, lo);
                }
                CGNN_STAMP(1);
                tmem_st_wait();
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster_relaxed(&bars->a_ready[g], 0);
                CGNN_STAMP(2);
            }
            // ---- hidden layers ------------------------------------------------------------------------------------
            for (int l = 0; l + 1 < N_LAYERS; ++l) {
                // row stream of this layer, thread = row: the mask source of a dgrad chain (the gather is part of the accumulator, GIN)
                const float* pa = !HS ? nullptr : (p.hid_mask[l] ? p.hid_mask[l] + rowoff : nullptr);
                const bool masked = pa != nullptr;
                float* hout = p.hid_out[l] ? p.hid_out[l] + rowoff : nullptr;
                float* hagg = AGG && p.hid_agg[l] ? p.hid_agg[l] + recvoff : nullptr;
                const float* bias = sVec + l * TC_H;
                // operands of the row stream: four 16-column chunk buffers, each reloaded four chunks (64 columns) ahead
                float a0[16], a1[16], a2[16], a3[16];
                if (pa) { ld16(pa, a0); ld16(pa + 16, a1); ld16(pa + 32, a2); ld16(pa + 48, a3); }
                mbar_wait_or_trap(&bars->mma_done[g], pm, 160 + l);
                pm ^= 1u;
                tc_fence_after_sync();
                CGNN_STAMP(3 + 4 * l);
                float va[16], vb[16];
                tmem_ld_32x32b_x16(tD + C0, va);
#pragma unroll 1
                for (int c = C0; c < C0 + NC; c += 64) {
#pragma unroll
                    for (int hh = 0; hh < 4; ++hh) {
                        float* v = (hh & 1) == 0 ? va : vb;
                        float* ca = hh == 0 ? a0 : hh == 1 ? a1 : hh == 2 ? a2 : a3;
                        const int cc = c + 16 * hh;
                        tmem_ld_wait16(v);
                        if (cc + 16 < C0 + NC) tmem_ld_32x32b_x16(tD + cc + 16, (hh & 1) == 0 ? vb : va);
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            const float4 b = *reinterpret_cast<const float4*>(bias + cc + j);
                            v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
                        }
                        if (masked) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) v[j] = ca[j] > 0.0f ? v[j] : 0.0f;
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.0f);
                        }
                        if (pa && cc + 64 < C0 + NC) ld16(pa + cc + 64, ca);
                        if (hout && valid) st16(hout + cc, v);
                        uint32_t hi[8], lo[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) split2(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
                        tmem_st_32x32b_x8(tAhi + cc / 2, hi);
                        if (NS == 3) tmem_st_32x32b_x8(tAlo + cc / 2, lo);
                        if (hagg) {
                            receiver_sum16(v, k, lane);
                            if (valid) receiver_store16(hagg, cc, v, k, lane);
                        }
                    }
                }
                CGNN_STAMP(5 + 4 * l);
                tmem_st_wait();
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster_relaxed(&bars->a_ready[g], 0);
                CGNN_STAMP(6 + 4 * l);
            }
            // ---- final layer: bias, [gather], [LayerNorm fwd / bwd], [ReLU], [mask], per-receiver sum, residual, store ----
            constexpr bool lnb = LNB;
            // T1: the dU rows (LayerNorm backward) or the residual arrive through the ring, 32 columns per slot, after the tile's
            // MMA input; every thread reads its own row of a slot (two 16-column halves) and the slot goes back after the second
            uint32_t tbuf = buf, t1fold = 0u;
            auto t1_load16 = [&](int cc, float* dst) {
                if (IO16 && p.t116) {
                    // bfloat16 stream: a slot holds 64 columns of the row (128 bytes); 16 columns are two of its 16-byte pieces
                    if ((cc & 63) == 0) {
                        mbar_wait_or_trap(&bars->in_full[g][tbuf], (in_par >> tbuf) & 1u, 190);
                        in_par ^= 1u << tbuf;
                    }
                    const uint8_t* src = sRing + tbuf * CW_BYTES;
                    const uint4 u0 = *reinterpret_cast<const uint4*>(src + swz128(r, (cc & 63) >> 3));
                    const uint4 u1 = *reinterpret_cast<const uint4*>(src + swz128(r, ((cc & 63) >> 3) + 1));
                    const uint32_t w[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        dst[2 * j] = __uint_as_float(w[j] << 16);
                        dst[2 * j + 1] = __uint_as_float(w[j] & 0xFFFF0000u);
                    }
                    t1fold ^= fold_zero<8>(w, zero_rt);
                    if ((cc & 63) == 48) {
                        __syncwarp();
                        if (lane == 0) mbar_arrive_local(&bars->in_empty[tbuf] + t1fold);
                        t1fold = 0u;
                        tbuf = tbuf + 1 == uring ? 0 : tbuf + 1;
                    }
                    return;
                }
                if ((cc & 31) == 0) {
                    mbar_wait_or_trap(&bars->in_full[g][tbuf], (in_par >> tbuf) & 1u, 190);
                    in_par ^= 1u << tbuf;
                }
                const uint8_t* src = sRing + tbuf * CW_BYTES;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 e4 = *reinterpret_cast<const float4*>(src + swz128(r, ((cc & 31) >> 2) + j));
                    dst[4 * j] = e4.x; dst[4 * j + 1] = e4.y; dst[4 * j + 2] = e4.z; dst[4 * j + 3] = e4.w;
                }
                // raw loaded values: the arrival that hands the slot back is made to depend on both halves (fold_zero)
                t1fold ^= fold_zero<16>(reinterpret_cast<const uint32_t*>(dst), zero_rt);
                if ((cc & 31) == 16) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive_local(&bars->in_empty[tbuf] + t1fold);
                    t1fold = 0u;
                    tbuf = tbuf + 1 == uring ? 0 : tbuf + 1;
                }
            };
            // row streams, thread = row:  a = dU rows | mask source,  b = dU per receiver | residual
            const float* pa = !SA     ? nullptr
                              : lnb   ? (p.du_rows ? p.du_rows + rowoff : nullptr)
                                      : (p.mask_src ? p.mask_src + rowoff : nullptr);
            const float* pb = !SB || RIN ? nullptr
                              : lnb   ? (p.du_recv ? p.du_recv + recvoff : nullptr)
                                      : (p.residual ? p.residual + rowoff : nullptr);
            float* outp = p.out + rowoff;
            uint16_t* outh = reinterpret_cast<uint16_t*>(p.out) + rowoff;      // out16: the same rows at 2 bytes per column
            const bool vout = valid && p.out != nullptr;            // the last processor step of a forward has no use for e'
            float* aggp = AGG && p.agg_out ? p.agg_out + recvoff : nullptr;
            // four 16-column chunk buffers per stream, each reloaded four chunks (64 columns) ahead
            float a0[16], a1[16], a2[16], a3[16], b0[16], b1[16], b2[16], b3[16];
            // (LayerNorm backward: requesting the whole dU row before the MMA wait was measured -- the wait grows by what the
            //  moment loop saves, 31.6 k vs 28.4 k cycles per tile: the scattered sectors are throughput-, not latency-limited)
            if (pa) { ld16(pa + C0, a0); ld16(pa + C0 + 16, a1); ld16(pa + C0 + 32, a2); ld16(pa + C0 + 48, a3); }
            if (pb) { ld16(pb + C0, b0); ld16(pb + C0 + 16, b1); ld16(pb + C0 + 32, b2); ld16(pb + C0 + 48, b3); }
            uint4 mbits = make_uint4(0u, 0u, 0u, 0u);
            if (p.mask_bits != nullptr) mbits = __ldg(reinterpret_cast<const uint4*>(p.mask_bits) + grow);
            mbar_wait_or_trap(&bars->mma_done[g], pm, 180);
            pm ^= 1u;
            tc_fence_after_sync();
            CGNN_STAMP(11);
            const float* bias = sVec + (N_LAYERS - 1) * TC_H;
            const float* gamma = sVec + 3 * TC_H;
            const float* beta = sVec + 4 * TC_H;
            float mean = 0.0f, rstd = 1.0f, gm1 = 0.0f, gm2 = 0.0f;
            if (LN != 0) {
                // one pass over the accumulator: moments of y shifted by the row's first element y0 (robust against
                // |mean| >> std) and, for the backward, sum g and sum g (y - y0) with g = dU * gamma
                float y0 = 0.0f, s1 = 0.0f, s2 = 0.0f, g1 = 0.0f, g2 = 0.0f;
                float va[16], vb[16];
                tmem_ld_32x32b_x16(tD + C0, va);
#pragma unroll 1
                for (int c = C0; c < C0 + NC; c += 64) {
#pragma unroll
                    for (int hh = 0; hh < 4; ++hh) {
                        float* v = (hh & 1) == 0 ? va : vb;
                        float* ca = hh == 0 ? a0 : hh == 1 ? a1 : hh == 2 ? a2 : a3;
                        float* cb = hh == 0 ? b0 : hh == 1 ? b1 : hh == 2 ? b2 : b3;
                        const int cc = c + 16 * hh;
                        tmem_ld_wait16(v);
                        if (cc + 16 < C0 + NC) tmem_ld_32x32b_x16(tD + cc + 16, (hh & 1) == 0 ? vb : va);
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            const float4 b = *reinterpret_cast<const float4*>(bias + cc + j);
                            v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
                        }
                        if (cc == C0) y0 = v[0];
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float d = v[j] - y0;
                            s1 += d;
                            s2 = fmaf(d, d, s2);
                        }
                        if (lnb) {
                            // dU of this chunk; stashed in the A-operand columns of this TMEM slot (free once the last MMA is done)
                            // for the output pass, so the dU streams are read from global memory only once
                            float du[16];
#pragma unroll
                            if (T1) {
                                float t1v[16];
                                t1_load16(cc, t1v);
#pragma unroll
                                for (int j = 0; j < 16; ++j) du[j] = valid && !dummy ? t1v[j] + (pb ? cb[j] : 0.0f) : 0.0f;
                            } else {
#pragma unroll
                                for (int j = 0; j < 16; ++j) du[j] = valid && !dummy ? (pa ? ca[j] : 0.0f) + (pb ? cb[j] : 0.0f) : 0.0f;
                            }
                            tmem_st_32x32b_x16(tAhi + cc, reinterpret_cast<const uint32_t*>(du));
#pragma unroll
                            for (int j = 0; j < 16; j += 4) {
                                const float4 gmv = *reinterpret_cast<const float4*>(gamma + cc + j);
                                const float gmj[4] = {gmv.x, gmv.y, gmv.z, gmv.w};
#pragma unroll
                                for (int u = 0; u < 4; ++u) {
                                    const float t = du[j + u] * gmj[u];
                                    g1 += t;
                                    g2 = fmaf(t, v[j + u] - y0, g2);
                                }
                            }
                            if (cc + 64 < TC_H) {
                                if (pa) ld16(pa + cc + 64, ca);
                                if (pb) ld16(pb + cc + 64, cb);
                            }
                        }
                    }
                }
                if (lnb) tmem_st_wait();
                // zero-padded columns (y = 0 exactly) took part in the sums with y - y0 = -y0: take them out, then divide by the
                // LayerNorm's own width
                const float pad = (float)(TC_H - p.ln_n), inv_n = 1.0f / (float)(SPLIT ? NC : p.ln_n);
                if (!SPLIT) { s1 = fmaf(pad, y0, s1); s2 = fmaf(-pad * y0, y0, s2); }
                float m1 = s1 * inv_n;
                mean = y0 + m1;
                float var = fmaxf(s2 * inv_n - m1 * m1, 0.0f);
                if (SPLIT) {
                    // This thread has the moments of its 64 columns; the other half of the row lives in the warp that shares the
                    // lane quarter.  Exchange (mean, M2) through four spare TMEM columns of the row (the A-operand columns are free
                    // once the last MMA is done) and combine: mean = (m_a + m_b) / 2, M2 = M2_a + M2_b + 32 (m_a - m_b)^2.
                    tmem_st_32x32b_x2(tAhi + 2 * half, __float_as_uint(mean), __float_as_uint(var * (float)NC));
                    tmem_st_wait();
                    tc_fence_before_sync();
                    named_bar_sync(1 + g * 4 + wq, 64);
                    tc_fence_after_sync();
                    float mm[4];
                    tmem_ld_32x32b_x4(tAhi, mm);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" : "+f"(mm[0]), "+f"(mm[1]), "+f"(mm[2]), "+f"(mm[3]) :: "memory");
                    const float dm = mm[0] - mm[2];
                    mean = 0.5f * (mm[0] + mm[2]);
                    var = (mm[1] + mm[3] + 32.0f * dm * dm) * (1.0f / TC_H);
                    m1 = 0.0f;
                }
                rstd = 1.0f / sqrtf(var + LN_EPS);
                gm1 = g1 * inv_n;                                          // mean of g
                gm2 = (g2 - m1 * g1) * rstd * inv_n;                       // mean of g * xhat
            }
            CGNN_STAMP(12);
            {
                float va[16], vb[16];
                tmem_ld_32x32b_x16(tD + C0, va);
#pragma unroll 1
                for (int c = C0; c < C0 + NC; c += 64) {
                    uint32_t gate0 = 0u, gate1 = 0u;         // ReLU gates of columns c .. c+31, c+32 .. c+63
                    float held[16];                          // the even 16-column block of a pair, waiting for the joint per-receiver sum
#pragma unroll
                    for (int hh = 0; hh < 4; ++hh) {
                        float* v = (hh & 1) == 0 ? va : vb;
                        const int cc = c + 16 * hh;
                        tmem_ld_wait16(v);
                        if (cc + 16 < C0 + NC) tmem_ld_32x32b_x16(tD + cc + 16, (hh & 1) == 0 ? vb : va);
                        float* ca = hh == 0 ? a0 : hh == 1 ? a1 : hh == 2 ? a2 : a3;
                        float* cb = hh == 0 ? b0 : hh == 1 ? b1 : hh == 2 ? b2 : b3;
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            const float4 b = *reinterpret_cast<const float4*>(bias + cc + j);
                            v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
                        }
                        if (LN == 1) {
#pragma unroll
                            for (int j = 0; j < 16; j += 4) {
                                const float4 gmv = *reinterpret_cast<const float4*>(gamma + cc + j);
                                const float4 bt = *reinterpret_cast<const float4*>(beta + cc + j);
                                v[j] = (v[j] - mean) * rstd * gmv.x + bt.x;
                                v[j + 1] = (v[j + 1] - mean) * rstd * gmv.y + bt.y;
                                v[j + 2] = (v[j + 2] - mean) * rstd * gmv.z + bt.z;
                                v[j + 3] = (v[j + 3] - mean) * rstd * gmv.w + bt.w;
                            }
                        } else if (lnb) {
                            // dY = rstd (g - mean(g) - xhat mean(g xhat));  column sums of dU xhat (d gamma) and dU (d beta)
                            float du[16], dgx[16];
                            tmem_ld_32x32b_x16(tAhi + cc, du);
                            tmem_ld_wait16(du);
#pragma unroll
                            for (int j = 0; j < 16; j += 4) {
                                const float4 gmv = *reinterpret_cast<const float4*>(gamma + cc + j);
                                const float gmj[4] = {gmv.x, gmv.y, gmv.z, gmv.w};
#pragma unroll
                                for (int u = 0; u < 4; ++u) {
                                    const float xh = (v[j + u] - mean) * rstd;
                                    dgx[j + u] = du[j + u] * xh;
                                    v[j + u] = cc + j + u < p.ln_n ? rstd * (du[j + u] * gmj[u] - gm1 - xh * gm2) : 0.0f;   // no gradient into the padding
                                }
                            }
                            receiver_sum16x2(dgx, du, 32, lane);       // the two column sums share one butterfly (their chains overlap)
                            if ((lane & 1) == 0) {
                                lnacc[cc + (lane >> 1)] += dgx[0];
                                lnacc[TC_H + cc + (lane >> 1)] += du[0];
                            }
                        }
                        if (p.relu_out) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.0f);
                        }
                        if (!lnb && pa) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) v[j] = ca[j] > 0.0f ? v[j] : 0.0f;
                        }
                        if (p.mask_bits != nullptr) {
                            const uint32_t wsel = (cc >> 5) == 0 ? mbits.x : (cc >> 5) == 1 ? mbits.y : (cc >> 5) == 2 ? mbits.z : mbits.w;
                            const uint32_t w16 = wsel >> (cc & 16);
#pragma unroll
                            for (int j = 0; j < 16; ++j) v[j] = (w16 >> j) & 1u ? v[j] : 0.0f;
                        }
                        if (p.bits_out != nullptr) {
                            uint32_t w16 = 0u;
#pragma unroll
                            for (int j = 0; j < 16; ++j) w16 |= (v[j] > 0.0f ? 1u : 0u) << j;
                            if (hh == 0) gate0 = w16;
                            else if (hh == 1) gate0 |= w16 << 16;
                            else if (hh == 2) gate1 = w16;
                            else gate1 |= w16 << 16;
                        }
                        if (RIN) {
                            // residual = this tile's input, still in its ring slots (this thread's own row, swizzled)
                            uint32_t rbuf = rin0 + (uint32_t)(cc >> 5);
                            if (rbuf >= uring) rbuf -= uring;
                            const uint8_t* src = sRing + rbuf * CW_BYTES;
                            float o[16];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float4 e4 = *reinterpret_cast<const float4*>(src + swz128(r, ((cc & 31) >> 2) + j));
                                o[4 * j] = v[4 * j] + e4.x; o[4 * j + 1] = v[4 * j + 1] + e4.y;
                                o[4 * j + 2] = v[4 * j + 2] + e4.z; o[4 * j + 3] = v[4 * j + 3] + e4.w;
                            }
                            if (vout) st16(outp + cc, o);
                            if ((cc & 31) == 16) {
                                // Both halves of the chunk are read: its ring slot goes back to the producer.  The arrival must come
                                // after an instruction that CONSUMES the loaded values in every lane: shared-memory loads that are only
                                // issued can still be queued behind the scattered global stores when lane 0's arrival becomes visible,
                                // and the refill then lands under them (seen as rare wrong residuals).  Consuming o[] here does that
                                // (consume16 reads the registers, the warp barrier then covers all lanes).
                                consume16(reinterpret_cast<const uint32_t*>(o));
                                __syncwarp();
                                if (lane == 0) mbar_arrive_local(&bars->in_empty[rbuf]);
                            }
                        } else if (T1 && !lnb) {
                            float t1v[16];
                            t1_load16(cc, t1v);
#pragma unroll
                            for (int j = 0; j < 16; ++j) t1v[j] += v[j];
                            if (vout) {
                                if (IO16 && p.out16) st16h(outh + cc, t1v);
                                else st16(outp + cc, t1v);
                            }
                        } else if (!lnb && pb) {
                            // residual: the sum goes out from the stream's own registers, v stays free for the per-receiver sum
#pragma unroll
                            for (int j = 0; j < 16; ++j) cb[j] += v[j];
                            if (vout) st16(outp + cc, cb);
                        } else if (vout) {
                            if (IO16 && p.out16) st16h(outh + cc, v);
                            else st16(outp + cc, v);
                        }
                        if (aggp) {
                            if (dummy) {
#pragma unroll
                                for (int j = 0; j < 16; ++j) v[j] = 0.0f;
                            }
                            // per-receiver sums, two 16-column blocks per butterfly: the even block waits for the odd one
                            if ((hh & 1) == 0) {
#pragma unroll
                                for (int j = 0; j < 16; ++j) held[j] = v[j];
                            } else {
                                receiver_sum16x2(held, v, k, lane);
                                if (valid) {
                                    receiver_store16(aggp, cc - 16, held, k, lane);
                                    receiver_store16(aggp, cc, v, k, lane);
                                }
                            }
                        }
                        if (!lnb && cc + 64 < C0 + NC) {
                            if (pa) ld16(pa + cc + 64, ca);
                            if (pb) ld16(pb + cc + 64, cb);
                        }
                    }
                    if (p.bits_out != nullptr && valid) *reinterpret_cast<uint2*>(p.bits_out + (size_t)grow * 4 + (c >> 5)) = make_uint2(gate0, gate1);
                }
            }
            CGNN_STAMP(13);
            // D and A of this slot are free again: the next tile of this group starts with its input phase
        }
        if (LNB) {
            // this warp's column sums -> its row of the partials (summed in fixed order by ln_partials_reduce)
            __syncwarp();
            float* dst = p.ln_partials + ((size_t)blockIdx.x * 8 + warp) * 2 * TC_H;
            for (int i = lane; i < 2 * TC_H; i += 32) dst[i] = lnacc[i];
        }
    }

    // ---- teardown -----------------------------------------------------------------------------------------
    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();
    if (warp == NEPI + 1) tmem_dealloc<2>(tmem, 512);
}

constexpr size_t SMEM_MAX = 227 * 1024;

// sums the per-warp LayerNorm-backward column sums [n_rows_p][2][128] in fixed order: block <-> 8 columns x 32 row slices,
// combined through shared memory
__global__ void __launch_bounds__(256) ln_partials_reduce_kernel(const float* __restrict__ partials, int n_rows_p,
                                                                 float* __restrict__ dgamma, float* __restrict__ dbeta, int accumulate) {
    __shared__ float sm[32][8];
    const int cl = threadIdx.x & 7, part = threadIdx.x >> 3;
    const int c = blockIdx.x * 8 + cl;
    float s = 0.0f;
    for (int r = part; r < n_rows_p; r += 32) s += partials[(size_t)r * 2 * TC_H + c];
    sm[part][cl] = s;
    __syncthreads();
    if (part == 0) {
        float t = 0.0f;
#pragma unroll
        for (int w = 0; w < 32; ++w) t += sm[w][cl];
        float* base = c < TC_H ? dgamma : dbeta;
        if (base != nullptr) {
            float* d = base + (c < TC_H ? c : c - TC_H);
            *d = accumulate ? *d + t : t;
        }
    }
}

template <int NS>
int run_chain_t(const ChainOp& op, cudaStream_t stream) {
    constexpr int NSI = NS == 3 ? 2 : 1;
    const int n_in = op.in1 ? 2 : 1;
    const int n_blocks = n_in + op.n_layers - 1;
    CGNN_CHECK_ARG(op.n_layers == 1 || op.n_layers == 3, "tensor-core chain: 1 or 3 layers");
    CGNN_CHECK_ARG(n_blocks <= MAX_BLOCKS && op.rows >= 1 && op.in0 && (op.out || op.agg_out), "tensor-core chain: bad arguments");
    const bool gather = op.Ps != nullptr;
    const bool uses_k = gather || op.agg_out || op.du_recv || op.hid_agg[0] || op.hid_agg[1];
    int k = 1, kshift = 0;
    if (uses_k) {
        CGNN_CHECK_ARG(op.k >= 1 && op.k <= 32 && (op.k & (op.k - 1)) == 0, "tensor-core chain: k must be a power of two <= 32 (got %d)", op.k);
        k = op.k;
        while ((1 << kshift) < k) ++kshift;
    }
    if (op.ln_bwd) {
        CGNN_CHECK_ARG(op.gamma && (op.du_rows || op.du_recv) && op.ln_ws && !op.mask_src && !op.residual && !op.agg_out,
                       "tensor-core chain: bad LayerNorm-backward arguments");
    }
    TcParams p{};
    p.n_rows = op.rows; p.n_pair_tiles = (op.rows + 255) / 256;
    p.n_in = n_in; p.n_layers = op.n_layers; p.k = k; p.kshift = kshift;
    p.k_valid = (uses_k && op.k_valid > 0 && op.k_valid < k) ? op.k_valid : 0;
    p.ln_mode = op.ln_bwd ? 2 : (op.gamma != nullptr ? 1 : 0);
    p.ln_n = op.ln_n > 0 ? op.ln_n : TC_H;
    CGNN_CHECK_ARG(p.ln_n <= TC_H, "tensor-core chain: LayerNorm width above 128");
    p.gather = gather; p.relu_out = op.relu_out; p.mask_src = op.mask_src; p.residual = op.residual;
    p.agg_out = op.agg_out; p.senders = op.senders; p.Ps = op.Ps; p.Pr = op.Pr;
    for (int l = 0; l < 2; ++l) { p.hid_mask[l] = op.hid_mask[l]; p.hid_out[l] = op.hid_out[l]; p.hid_agg[l] = op.hid_agg[l]; }
    p.du_rows = op.du_rows; p.du_recv = op.du_recv; p.ln_partials = static_cast<float*>(op.ln_ws);
    p.out = op.out; p.bits_out = op.bits_out; p.mask_bits = op.mask_bits;
    p.in16 = op.in16 != 0; p.out16 = op.out16 != 0;
    CGNN_CHECK_ARG(!op.in16 || (op.n_layers == 1 && !op.in1 && !gather && op.in0_cols == 0 && op.residual != op.in0),
                   "tensor-core chain: a bfloat16 input stream feeds one-layer chains with a single input");
    CGNN_CHECK_ARG(!op.out16 || (op.out && op.n_layers == 1 && !gather),
                   "tensor-core chain: a bfloat16 output stream belongs to one-layer chains without gather");
    for (int b = 0; b < n_blocks; ++b) p.blk[b] = op.blk[b];
    if (op.n_layers == 1) {
        p.vec_src[0] = op.bias[0];
    } else {
        p.vec_src[0] = op.bias[0]; p.vec_src[1] = op.bias[1]; p.vec_src[2] = op.bias[2];
    }
    p.vec_src[3] = op.gamma; p.vec_src[4] = op.beta;
    for (int v = 0; v < 5; ++v) p.vec_len[v] = TC_H;
    p.vec_len[3] = p.vec_len[4] = p.ln_n;
    if (op.out_valid > 0) p.vec_len[op.n_layers - 1] = op.out_valid;       // bias of a narrow last layer
    if (g_stamps != nullptr && g_stamp_next < g_stamp_launches) {
        p.stamps = g_stamps + (size_t)(g_stamp_next++) * 2 * g_stamp_tiles * 16;
        p.stamp_tiles = g_stamp_tiles;
    }
    CUtensorMap m0, m1;
    int rc;
    if (op.in16) rc = make_row_map_bf16(&m0, op.in0, op.rows, 128);
    else rc = make_row_map32_rows(&m0, op.in0, op.rows, 128, op.in0_cols > 0 ? op.in0_cols : TC_H);
    if (rc) return rc;
    // the shared memory the weights leave goes to the input ring
    const size_t ring_off = ring_offset(n_blocks, NSI, op.ln_bwd != 0);
    // the residual is the chain's own (first) input and the ring can hold both groups' tiles: no re-read
    static const bool rin_allowed = getenv("CGNN_NO_RIN") == nullptr;       // measurements: force the re-read path
    const bool rin = rin_allowed && !op.ln_bwd && op.residual != nullptr && op.residual == op.in0 &&
                     ring_off + (size_t)(2 * NCW * n_in) * CW_BYTES <= SMEM_MAX;
    CGNN_CHECK_ARG(ring_off + 2 * CW_BYTES <= SMEM_MAX, "tensor-core chain: shared memory need %zu exceeds 227 KB", ring_off + 2 * CW_BYTES);
    p.n_ring = (int)((SMEM_MAX - ring_off) / CW_BYTES);
    if (p.n_ring > MAXRING) p.n_ring = MAXRING;
    const size_t smem = ring_off + (size_t)p.n_ring * CW_BYTES;
    // the instantiation for this chain's shape
    const bool any_hidden = op.n_layers == 3 && (op.hid_mask[0] || op.hid_mask[1]);
    // one-layer chains: the dU rows / the residual ride the TMA ring (CGNN_NO_T1=1: thread = row global loads, as measured before)
    // The residual chains take it by default.  The LayerNorm-backward chain does NOT: with its dU rows on the ring the edge backward
    // stopped being reproducible (tools/stress_edge_bwd.py: 17 of 39 repetitions differed, 0 of 39 without) -- a race not yet
    // found.  FOUND (round 2, second session): the hand-back of a ring slot whose loaded registers are passed on raw did not wait for
    // the loads (fold_zero, tc_common.cuh); with the arrival depending on them 0 of 39 / 29 / 39 repetitions differ, and the
    // LayerNorm-backward chain takes its dU rows through the ring by default again (CGNN_T1_LNB=0: thread = row loads).
    static const bool t1_off = getenv("CGNN_NO_T1") != nullptr && atoi(getenv("CGNN_NO_T1")) != 0;
    static const bool t1_lnb = getenv("CGNN_T1_LNB") == nullptr || atoi(getenv("CGNN_T1_LNB")) != 0;
    const float* t1_src = t1_off || op.n_layers != 1 || op.in1 || gather ? nullptr
                          : op.ln_bwd ? (t1_lnb ? op.du_rows : nullptr)
                          : (op.residual != op.in0 && !op.mask_src ? op.residual : nullptr);
    const bool t1 = t1_src != nullptr;
    // a bfloat16 gradient stream (dU rows / residual) only travels through the ring
    CGNN_CHECK_ARG(!(op.du16 && op.du_rows) || (op.ln_bwd && t1), "tensor-core chain: bfloat16 dU rows need the ring-fed LayerNorm-backward chain");
    CGNN_CHECK_ARG(!(op.res16 && op.residual) || (!op.ln_bwd && t1), "tensor-core chain: a bfloat16 residual needs the ring-fed one-layer chain");
    CGNN_CHECK_ARG(!op.out16 || !op.residual || t1, "tensor-core chain: a bfloat16 output with a residual needs the ring-fed one-layer chain");
    p.t116 = t1 && (op.ln_bwd ? op.du16 : op.res16);
    // gather chains whose ring and second tensor map are free can take the P_s rows by tile::gather4 (CGNN_GATHER4=1).  Measured
    // slower than the thread = row loads (128 four-row gathers per tile: 25.2 k vs 20.4 k cycles per tile of the A1 chain), so opt-in.
    static const bool g4_allowed = getenv("CGNN_GATHER4") != nullptr && atoi(getenv("CGNN_GATHER4")) != 0;
    const bool g4 = g4_allowed && gather && !op.in1 && !t1 && !rin && op.n_layers == 1;
    const bool fin_a = op.ln_bwd ? (op.du_rows != nullptr && !t1) : op.mask_src != nullptr;
    const bool fin_b = op.ln_bwd ? op.du_recv != nullptr : (op.residual != nullptr && !t1);
    const bool any_agg = op.agg_out || op.hid_agg[0] || op.hid_agg[1];
    // the fused forward chains (3 layers + LayerNorm) are bound by the serial epilogue of a tile, not by memory: their columns
    // are split over 16 epilogue warps.  The one-layer chains run at the memory rate either way and keep 8 warps.
    // Measured (profiles/r02_chain_stage_stamps.txt): the split shortens a tile's own time (39.1 k -> 34.0 k cycles) but not
    // the tile period (39.6 k vs 39.9 k) nor the step, so it is opt-in (CGNN_SPLIT=1) until the coupling is understood.
    static const bool split_allowed = getenv("CGNN_SPLIT") != nullptr && atoi(getenv("CGNN_SPLIT")) != 0;
    const bool split = split_allowed && p.ln_n == TC_H && op.n_layers == 3 && op.gamma != nullptr && !op.ln_bwd && !any_hidden && !fin_a &&
                       !op.hid_out[0] && !op.hid_out[1] && !op.hid_agg[0] && !op.hid_agg[1] && !op.bits_out && !op.mask_bits;
    const int cfg = (op.n_layers == 3 ? C_L3 : 0) | (op.ln_bwd ? C_LNB : (op.gamma ? C_LN : 0)) | (fin_a ? C_SA : 0) | (fin_b ? C_SB : 0) |
                    (any_hidden ? C_HS : 0) | (any_agg ? C_AGG : 0) | (rin ? C_RIN : 0) | (gather ? C_GIN : 0) | (split ? C_SPLIT : 0) |
                    (t1 ? C_T1 : 0) | (g4 ? C_G4 : 0);
    if (g4) {
        // (the node table's row count is not part of the op; the gather only ever names valid sender rows, so the map's extent is
        //  set to the int32 index range the senders can express)
        if ((rc = make_gather_map32(&m1, op.Ps, op.ps_rows))) return rc;
    } else if (p.t116) {
        if ((rc = make_row_map_bf16(&m1, t1_src, op.rows, 128))) return rc;
    } else if ((rc = make_row_map32(&m1, t1 ? t1_src : (op.in1 ? op.in1 : op.in0), op.rows))) return rc;       // (unused map when there is no second stream)
    void (*kern)(CUtensorMap, CUtensorMap, TcParams) = nullptr;
    int slot = -1;
#define CGNN_CHAIN_CFG(i, c) else if (cfg == (c)) { kern = tc_chain_fwd<NS, (c)>; slot = (i); }
    if (false) {}
    CGNN_CHAIN_CFG(0, 0)                                              // 1 layer: plain / ReLU / two inputs
    CGNN_CHAIN_CFG(1, C_GIN)                                          // 1 layer + gather (A1 of the edge backward)
    CGNN_CHAIN_CFG(2, C_SA)                                           // 1 layer + mask (dgrad)
    CGNN_CHAIN_CFG(3, C_SA | C_AGG)                                   // 1 layer + mask + per-receiver sum
    CGNN_CHAIN_CFG(4, C_SB)                                           // 1 layer + residual
    CGNN_CHAIN_CFG(5, C_L3)                                           // decoders
    CGNN_CHAIN_CFG(6, C_L3 | C_LN)                                    // encoders
    CGNN_CHAIN_CFG(7, C_L3 | C_LN | C_SB)                             // node phase forward
    CGNN_CHAIN_CFG(8, C_L3 | C_LN | C_SB | C_GIN)                     // edge phase forward without the per-receiver sum
    CGNN_CHAIN_CFG(9, C_L3 | C_LN | C_SB | C_GIN | C_AGG)             // edge phase forward
    CGNN_CHAIN_CFG(21, C_L3 | C_LN | C_SB | C_GIN | C_AGG | C_RIN)    // edge phase forward, residual from the input ring
    CGNN_CHAIN_CFG(22, C_L3 | C_LN | C_SB | C_GIN | C_RIN)
    CGNN_CHAIN_CFG(23, C_L3 | C_LN | C_SB | C_RIN)                    // node phase forward, residual from the input ring
    CGNN_CHAIN_CFG(25, C_SPLIT | C_L3 | C_LN)                         // the same forward chains with the column-split epilogue
    CGNN_CHAIN_CFG(26, C_SPLIT | C_L3 | C_LN | C_SB)
    CGNN_CHAIN_CFG(27, C_SPLIT | C_L3 | C_LN | C_SB | C_GIN)
    CGNN_CHAIN_CFG(28, C_SPLIT | C_L3 | C_LN | C_SB | C_GIN | C_AGG)
    CGNN_CHAIN_CFG(29, C_SPLIT | C_L3 | C_LN | C_SB | C_GIN | C_AGG | C_RIN)
    CGNN_CHAIN_CFG(30, C_SPLIT | C_L3 | C_LN | C_SB | C_GIN | C_RIN)
    CGNN_CHAIN_CFG(31, C_SPLIT | C_L3 | C_LN | C_SB | C_RIN)
    CGNN_CHAIN_CFG(10, C_L3 | C_LNB | C_SA | C_SB | C_GIN)            // edge backward, recompute + LayerNorm backward
    CGNN_CHAIN_CFG(11, C_L3 | C_LNB | C_SA)                           // node backward, recompute + LayerNorm backward
    CGNN_CHAIN_CFG(12, C_L3 | C_SB | C_HS | C_AGG)                    // edge backward, dgrad chain
    CGNN_CHAIN_CFG(13, C_L3 | C_HS)                                   // node backward, dgrad chain
    CGNN_CHAIN_CFG(14, C_AGG)                                         // 1 layer + per-receiver sum
    CGNN_CHAIN_CFG(15, C_L3 | C_LNB | C_SB | C_GIN)                   // edge backward recompute, no gradient on the edge output
    CGNN_CHAIN_CFG(16, C_L3 | C_HS | C_AGG)                           // edge backward dgrad, no gradient on the edge output
    CGNN_CHAIN_CFG(17, C_L3 | C_SB | C_HS)                            // node backward dgrad
    CGNN_CHAIN_CFG(18, C_LNB | C_SA | C_SB)                           // 1 layer + LayerNorm backward (dU per row and per receiver)
    CGNN_CHAIN_CFG(19, C_LNB | C_SA)
    CGNN_CHAIN_CFG(20, C_LNB | C_SB)
    CGNN_CHAIN_CFG(34, C_GIN | C_G4)                                  // 1 layer + gather through the ring (tile::gather4)
    CGNN_CHAIN_CFG(24, C_T1)                                          // 1 layer + residual through the ring
    CGNN_CHAIN_CFG(32, C_LNB | C_T1)                                  // 1 layer + LayerNorm backward, dU rows through the ring
    CGNN_CHAIN_CFG(33, C_LNB | C_SB | C_T1)                           //   ... plus dU per receiver
#undef CGNN_CHAIN_CFG
    if (kern == nullptr) {
        set_error("tensor-core chain: no kernel instantiated for configuration 0x%x", cfg);
        return CGNN_ERR_UNSUPPORTED;
    }
    static size_t configured[35] = {0};
    if (smem > configured[slot]) {
        CGNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[slot] = smem;
    }
    int64_t pairs = num_sms() / 2;
    if (p.n_pair_tiles < pairs) pairs = p.n_pair_tiles;
    cudaLaunchConfig_t lc{};
    lc.gridDim = dim3((unsigned)(pairs * 2));
    lc.blockDim = dim3(split ? TC_THREADS_SPLIT : TC_THREADS);
    lc.dynamicSmemBytes = smem;
    lc.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    lc.attrs = at;
    lc.numAttrs = 1;
    CGNN_CUDA(cudaLaunchKernelEx(&lc, kern, m0, m1, p));
    count_launch();
    if (op.ln_bwd) {
        ln_partials_reduce_kernel<<<2 * TC_H / 8, 256, 0, stream>>>(p.ln_partials, (int)(pairs * 2 * 8), op.dgamma, op.dbeta, op.accumulate);
        CGNN_LAUNCH_CHECK();
    }
    return CGNN_OK;
}

}  // namespace

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_map(CUtensorMap* m, const void* base, int64_t rows, int box_cols, CUtensorMapSwizzle swz, int box_rows = 128, int cols = TC_H,
                    bool bf16 = false);
int make_row_map(CUtensorMap* m, const float* base, int64_t rows) { return make_map(m, base, rows, CH, CU_TENSOR_MAP_SWIZZLE_64B); }

// `cols`: width of the array in global memory (row pitch cols * 4 bytes, a multiple of 16).  A box may reach beyond it: the TMA
// fills the columns outside the tensor with zeros and reads nothing for them -- how a narrow array (edge features, 4 columns)
// enters the 128-column chains without a padded copy in HBM.
static int make_map(CUtensorMap* m, const void* base, int64_t rows, int box_cols, CUtensorMapSwizzle swz, int box_rows, int cols, bool bf16) {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        CGNN_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q));
        CGNN_CHECK_ARG(f != nullptr && q == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled is not available in this driver");
        fn = reinterpret_cast<EncodeTiledFn>(f);
    }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * (bf16 ? 2 : 4)};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = fn(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (base %p, rows %lld)", (int)r, (const void*)base, (long long)rows);
        return CGNN_ERR_CUDA;
    }
    return CGNN_OK;
}

int make_row_map32(CUtensorMap* m, const float* base, int64_t rows) { return make_map(m, base, rows, 32, CU_TENSOR_MAP_SWIZZLE_128B); }
int make_row_map32_rows(CUtensorMap* m, const float* base, int64_t rows, int box_rows, int cols) { return make_map(m, base, rows, 32, CU_TENSOR_MAP_SWIZZLE_128B, box_rows, cols); }
int make_gather_map32(CUtensorMap* m, const float* base, int64_t rows) { return make_map(m, base, rows, 32, CU_TENSOR_MAP_SWIZZLE_128B, 1); }
// bfloat16 [rows][128]: boxes of box_rows rows x 64 columns -- the same 128-byte rows and swizzle as the FP32 boxes of 32 columns
int make_row_map_bf16(CUtensorMap* m, const void* base, int64_t rows, int box_rows) {
    return make_map(m, base, rows, 64, CU_TENSOR_MAP_SWIZZLE_128B, box_rows, TC_H, true);
}

void set_debug_stamps(unsigned long long* buf, int tiles, int launches) {
    g_stamps = buf; g_stamp_tiles = tiles; g_stamp_launches = launches; g_stamp_next = 0;
}

int run_chain(const ChainOp& op, cudaStream_t stream) {
    return op.ns == 3 ? run_chain_t<3>(op, stream) : run_chain_t<1>(op, stream);
}

}  // namespace cgnn
