// Tensor-core weight gradients and the LayerNorm backward used by the tensor-core backward pass.
//
//   tc_wgrad_kernel<NS>:  dW[n][c] = sum_rows X[row][n] * A[row][c]   and   db[n] = sum_rows X[row][n]
//     X = the layer's output gradient, A = the layer's input, both FP32 [rows][128] in HBM.  The contraction
//     runs over ROWS, so both operands are MN-major for the tensor core: the converter warps read 128-row
//     tiles (TMA, 64-byte swizzle), split them into BF16 hi/lo and write the UMMA MN-major shared-memory
//     images; one thread issues tcgen05.mma (cta_group::1, M = N = 128, K = 16 rows per instruction) into a
//     TMEM accumulator that lives for the whole kernel.  A constant "ones" operand gives the bias gradient as
//     16 extra accumulator columns.  Every CTA writes its partial [128][132] block; a second kernel sums the
//     partials in fixed order (deterministic, no float atomics).
//   ln_bwd_kernel: dY = LayerNorm backward of (Y, dU), plus per-block partials of d(gamma), d(beta).
//
// Reference semantics: autograd of graph_network.py:15-32,133-135 (train.py:264).
#include "tc_common.cuh"

namespace cgnn {
namespace {

using namespace ptx;

constexpr int WG_THREADS = 320;           // 8 converter warps (four sets of two: set s takes the chunks with seq % 4 == s) + producer warp + MMA warp
constexpr int WG_PROD = 8, WG_MMA = 9;
constexpr int TR = 64;                    // rows per tile (the K extent of one accumulation step).  Round 1 used 128-row tiles with a 4-slot
                                          // ring: the operand images took 128 KB, the ring held half a tile, and ncu showed the converter
                                          // warps waiting for TMA data in 46 % of their samples.  64-row tiles halve the images, the freed
                                          // shared memory goes to the ring: two tiles of input in flight per SM.
constexpr int WG_SETS = 4;
constexpr int WG_RING = 12;               // input ring: TR rows x 32 columns (128-byte rows, SWIZZLE_128B) per chunk.  A multiple of WG_SETS,
                                          // so that a ring slot is always consumed by the same converter set and the phase parity a set
                                          // waits on can never alias a phase that belongs to another set.  (16 slots until the operand
                                          // images were double-buffered: 96 KB in flight per SM still cover the memory latency.)
constexpr int CW = 32, NCW = TC_H / CW, CW_BYTES = TR * CW * 4;
constexpr int OP_BYTES = 128 * TR * 2;    // one BF16 operand image (128 mn x TR k)
constexpr int MN_STRIDE = (TR / 8) * 128; // bytes between groups of 8 mn in an MN-major image
constexpr int ONES_BYTES = 16 * TR * 2;
constexpr int PW = 132;                   // floats per partial row: 128 dW columns + db + pad
constexpr float LN_EPS = 1e-5f;

// (a ring slot's set is slot % WG_SETS because WG_RING is a multiple of WG_SETS -- whatever the number of chunks per tile: 8, or 6
//  with a bfloat16 X stream)
static_assert(WG_RING % WG_SETS == 0, "a ring slot must keep its converter set");
struct WgSmem {
    static constexpr int ring = 0;
    static constexpr int bars = ring + WG_RING * CW_BYTES;
    static constexpr int ones = bars + 512;
    static constexpr int ops = ones + ONES_BYTES;          // two sets of {X images (NSI), A images (NSI)}: tile it uses set it & 1
};
static_assert(WgSmem::ops % 128 == 0, "operand images need 128-byte alignment");

struct WgBars {
    uint64_t in_full[WG_RING], in_empty[WG_RING];
    uint64_t ops_full[2];    // per image set: 8 converter warps
    uint64_t ops_empty[2];   // per image set: tcgen05.commit of the tile that read it
    uint32_t tmem_base;
};

// MN-major image: element (mn, k) at (mn/8)*MN_STRIDE + (k/8)*128 + (k%8)*16 + (mn%8)*2
// The operand images are double-buffered (round 2, second session): with one set the converters of tile i + 1 waited for the MMAs of
// tile i, and convert -> publish -> MMA -> commit ran as one serial chain per 64-row tile (~3.1 k cycles per tile, 4.4 TB/s on the
// bfloat16 X stream: latency-, not memory-bound).
// X16: X is a bfloat16 [rows][128] stream (the 2-byte gradient stream): two 64-column chunks per tile whose 16-byte pieces ARE the
// rows of the image's 8 x 8 core matrices -- no split, no low image, two MMAs per K step instead of three.
template <int NS, bool X16>
__global__ void __launch_bounds__(WG_THREADS, 1)
tc_wgrad_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_a,
                int64_t n_tiles, float* __restrict__ partials) {
    constexpr int NSI = NS == 3 ? 2 : 1;
    extern __shared__ __align__(1024) uint8_t smem[];
    WgBars* bars = reinterpret_cast<WgBars*>(smem + WgSmem::bars);
    uint8_t* sOnes = smem + WgSmem::ones;
    constexpr int SET_BYTES = 2 * NSI * OP_BYTES;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t n_it = n_tiles > blockIdx.x ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (tid == 0) {
        for (int i = 0; i < WG_RING; ++i) { mbar_init(&bars->in_full[i], 1); mbar_init(&bars->in_empty[i], 2); }
        for (int b = 0; b < 2; ++b) { mbar_init(&bars->ops_full[b], 8); mbar_init(&bars->ops_empty[b], 1); }
        fence_mbar_init();
    }
    for (int i = tid; i < ONES_BYTES / 4; i += WG_THREADS) reinterpret_cast<uint32_t*>(sOnes)[i] = 0u;
    __syncthreads();
    for (int k = tid; k < TR; k += WG_THREADS)
        *reinterpret_cast<uint16_t*>(sOnes + (k >> 3) * 128 + (k & 7) * 16) = 0x3F80;     // bf16(1.0) at mn = 0
    if (warp == WG_MMA) tmem_alloc<1>(&bars->tmem_base, 256);
    if (warp == WG_PROD && lane == 0) { prefetch_tmap(&tm_x); prefetch_tmap(&tm_a); }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = bars->tmem_base;

    if (warp == WG_PROD) {
        // ---- producer: X chunks then A chunks of every tile --------------------------------------------
        if (lane == 0) {
            uint32_t seq = 0;
            for (int64_t it = 0; it < n_it; ++it) {
                const int64_t row0 = (blockIdx.x + it * gridDim.x) * TR;
                for (int op = 0; op < 2; ++op) {
                    const int nq = X16 && op == 0 ? NCW / 2 : NCW, qcols = X16 && op == 0 ? 2 * CW : CW;
                    for (int q = 0; q < nq; ++q, ++seq) {
                        const uint32_t buf = seq % WG_RING, use = seq / WG_RING;
                        mbar_wait_or_trap(&bars->in_empty[buf], (use & 1) ^ 1, 200 + buf);
                        mbar_expect_tx(&bars->in_full[buf], CW_BYTES);
                        tma_load_2d(smem + WgSmem::ring + buf * CW_BYTES, op == 0 ? &tm_x : &tm_a, q * qcols, (int)row0, &bars->in_full[buf]);
                    }
                }
            }
        }
    } else if (warp == WG_MMA) {
        // ---- MMA issuer ---------------------------------------------------------------------------------------
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16_major(128, TC_H, 1, 1);
            const uint32_t idesc1 = umma_idesc_bf16_major(128, 16, 1, 1);
            const uint32_t ones = smem_u32(sOnes);
            for (int64_t it = 0; it < n_it; ++it) {
                const int b = (int)(it & 1);
                const uint32_t x_hi = smem_u32(smem + WgSmem::ops + b * SET_BYTES), x_lo = x_hi + OP_BYTES;
                const uint32_t a_hi = x_hi + NSI * OP_BYTES, a_lo = a_hi + OP_BYTES;
                mbar_wait_or_trap(&bars->ops_full[b], (uint32_t)((it >> 1) & 1), 210);
                tc_fence_after_sync();
                const uint32_t first = it == 0 ? 0u : 1u;
#pragma unroll
                for (int ks = 0; ks < TR / 16; ++ks) {
                    const uint32_t acc = (ks > 0) ? 1u : first;
                    const uint64_t dxh = umma_desc(x_hi + ks * 256, 128, MN_STRIDE);
                    umma_bf16<1>(tmem, dxh, umma_desc(a_hi + ks * 256, 128, MN_STRIDE), idesc, acc);
                    umma_bf16<1>(tmem + 128, dxh, umma_desc(ones + ks * 256, 128, MN_STRIDE), idesc1, acc);
                    if (NS == 3) {
                        const uint64_t dxl = umma_desc(x_lo + ks * 256, 128, MN_STRIDE);
                        if (!X16) umma_bf16<1>(tmem, dxl, umma_desc(a_hi + ks * 256, 128, MN_STRIDE), idesc, 1u);
                        umma_bf16<1>(tmem, dxh, umma_desc(a_lo + ks * 256, 128, MN_STRIDE), idesc, 1u);
                        if (!X16) umma_bf16<1>(tmem + 128, dxl, umma_desc(ones + ks * 256, 128, MN_STRIDE), idesc1, 1u);
                    }
                }
                umma_commit<1>(&bars->ops_empty[b]);
            }
        }
    } else {
        // ---- converters: thread = row of the tile; set s (two warps) takes the chunks with seq % 4 == s, so every
        //      scheduler has two converter warps to interleave ---------------------------------------------------
        const int r = tid & (TR - 1);
        const int cset = tid >> 6;
        uint32_t seq = 0;
        for (int64_t it = 0; it < n_it; ++it) {
            const int b = (int)(it & 1);
            if (it >= 2) mbar_wait_or_trap(&bars->ops_empty[b], (uint32_t)(((it >> 1) - 1) & 1), 220);   // the MMAs of tile it - 2 read this set
            uint8_t* sX = smem + WgSmem::ops + b * SET_BYTES;
            uint8_t* sA = sX + NSI * OP_BYTES;
            for (int op = 0; op < 2; ++op) {
                uint8_t* img = op == 0 ? sX : sA;
                const int nq = X16 && op == 0 ? NCW / 2 : NCW;
                for (int q = 0; q < nq; ++q, ++seq) {
                    if ((int)(seq & (WG_SETS - 1)) != cset) continue;
                    const uint32_t buf = seq % WG_RING, use = seq / WG_RING;
                    mbar_wait_or_trap(&bars->in_full[buf], use & 1, 230 + buf);
                    const uint8_t* src = smem + WgSmem::ring + buf * CW_BYTES;
                    if (X16 && op == 0) {
                        // columns 64q .. 64q+63 of row r: piece j = columns 64q + 8j .. + 7 = the k = r row of mn group 8q + j
                        uint32_t w[32];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const uint4 v = *reinterpret_cast<const uint4*>(src + (uint32_t)(r * 128 + ((j ^ (r & 7)) << 4)));
                            w[4 * j] = v.x; w[4 * j + 1] = v.y; w[4 * j + 2] = v.z; w[4 * j + 3] = v.w;
                        }
                        // the raw words go to the image first (a real reader of every loaded register), then the slot goes back --
                        // through an arrival that depends on them (fold_zero, tc_common.cuh)
                        uint8_t* dst = img + (8 * q) * MN_STRIDE + (r >> 3) * 128 + (r & 7) * 16;
#pragma unroll
                        for (int j = 0; j < 8; ++j) *reinterpret_cast<uint4*>(dst + j * MN_STRIDE) = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
                        const uint32_t fz = fold_zero<32>(w, (uint32_t)((unsigned long long)n_tiles >> 62));
                        __syncwarp();
                        if (lane == 0) mbar_arrive_local(&bars->in_empty[buf] + fz);
                        continue;
                    }
                    uint32_t hi[16], lo[16];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 v = *reinterpret_cast<const float4*>(src + (uint32_t)(r * 128 + ((j ^ (r & 7)) << 4)));
                        split2(v.x, v.y, hi[2 * j], lo[2 * j]);
                        split2(v.z, v.w, hi[2 * j + 1], lo[2 * j + 1]);
                    }
                    consume16(hi);                       // every lane's loads have returned before the slot is handed back
                    __syncwarp();
                    if (lane == 0) mbar_arrive_local(&bars->in_empty[buf]);
                    // columns 32q .. 32q+31 = mn groups 4q .. 4q+3; k = r
                    uint8_t* dst = img + (4 * q) * MN_STRIDE + (r >> 3) * 128 + (r & 7) * 16;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        *reinterpret_cast<uint4*>(dst + j * MN_STRIDE) = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
                        if (NS == 3) *reinterpret_cast<uint4*>(dst + OP_BYTES + j * MN_STRIDE) = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
                    }
                }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive_local(&bars->ops_full[b]);
        }
        // ---- write this CTA's partial block -----------------------------------------------------------------
        if (n_it > 0) {       // (a commit covers every MMA issued before it: the last tile's is enough)
            mbar_wait_or_trap(&bars->ops_empty[(n_it - 1) & 1], (uint32_t)(((n_it - 1) >> 1) & 1), 240);
            tc_fence_after_sync();
        }
        // accumulator row = TMEM lane: warp w reads lane quarter w & 3; the two warps of a quarter alternate 16-column blocks
        const int orow = tid & 127, ocpar = tid >> 7;
        float* dst = partials + ((size_t)blockIdx.x * 128 + orow) * PW;
        const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
#pragma unroll 1
        for (int c0 = ocpar * 16; c0 < 144; c0 += 32) {
            float v[16];
            if (n_it > 0) {
                tmem_ld_32x32b_x16(trow + c0, v);
                tmem_ld_wait();
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = 0.0f;
            }
            if (c0 < 128) {
#pragma unroll
                for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(dst + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            } else {
                dst[128] = v[0];
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == WG_MMA) tmem_dealloc<1>(tmem, 256);
}

// sums the per-CTA partials in fixed order: block <-> 64 outputs x 4 slices of the CTA list, combined through shared memory
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ partials, int n_cta, float* __restrict__ dW, int ld, int col0,
                                                           float* __restrict__ db, int accumulate, int nrows, int ncols) {
    __shared__ float sm[4][64];
    const int o = threadIdx.x & 63, sl = threadIdx.x >> 6;
    const int idx = blockIdx.x * 64 + o;
    const int n = idx / 129, c = idx % 129;
    const bool live = idx < 128 * 129 && n < nrows && (c == 128 || c < ncols);
    float s = 0.0f;
    if (live)
        for (int g = sl; g < n_cta; g += 4) s += partials[((size_t)g * 128 + n) * PW + c];
    sm[sl][o] = s;
    __syncthreads();
    if (sl == 0 && live) {
        s = (sm[0][o] + sm[1][o]) + (sm[2][o] + sm[3][o]);
        if (c < 128) {
            float* d = dW + (size_t)n * ld + col0 + c;
            *d = accumulate ? *d + s : s;
        } else if (db != nullptr) {
            db[n] = accumulate ? db[n] + s : s;
        }
    }
}

// ---- LayerNorm backward --------------------------------------------------------------------------------------
// warp per row, lane <-> 4 columns.  dU = (dU_rows ? dU_rows[row] : 0) + (dU_recv ? dU_recv[row / k] : 0)
constexpr int LNB_WARPS = 8;
__global__ void __launch_bounds__(LNB_WARPS * 32)
ln_bwd_kernel(const float* __restrict__ Y, const float* __restrict__ dU_rows, const float* __restrict__ dU_recv, int k,
              const float* __restrict__ gamma, int64_t rows, float* __restrict__ dY, float* __restrict__ partials) {
    __shared__ float sg[LNB_WARPS][TC_H], sb[LNB_WARPS][TC_H];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float4 gm = reinterpret_cast<const float4*>(gamma)[lane];
    float4 ag = make_float4(0.f, 0.f, 0.f, 0.f), ab = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t row = (int64_t)blockIdx.x * LNB_WARPS + warp; row < rows; row += (int64_t)gridDim.x * LNB_WARPS) {
        const float4 y = reinterpret_cast<const float4*>(Y + row * TC_H)[lane];
        float4 du = make_float4(0.f, 0.f, 0.f, 0.f);
        if (dU_rows) du = reinterpret_cast<const float4*>(dU_rows + row * TC_H)[lane];
        if (dU_recv) {
            const float4 t = reinterpret_cast<const float4*>(dU_recv + (row / k) * TC_H)[lane];
            du.x += t.x; du.y += t.y; du.z += t.z; du.w += t.w;
        }
        float s = (y.x + y.y) + (y.z + y.w);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
        const float mean = s * (1.0f / TC_H);
        const float d0 = y.x - mean, d1 = y.y - mean, d2 = y.z - mean, d3 = y.w - mean;
        float v = fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, d3 * d3)));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
        const float rstd = 1.0f / sqrtf(v * (1.0f / TC_H) + LN_EPS);
        const float x0 = d0 * rstd, x1 = d1 * rstd, x2 = d2 * rstd, x3 = d3 * rstd;
        const float g0 = du.x * gm.x, g1 = du.y * gm.y, g2 = du.z * gm.z, g3 = du.w * gm.w;
        float s1 = (g0 + g1) + (g2 + g3);
        float s2 = fmaf(g0, x0, fmaf(g1, x1, fmaf(g2, x2, g3 * x3)));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xFFFFFFFFu, s1, o); s2 += __shfl_xor_sync(0xFFFFFFFFu, s2, o); }
        s1 *= (1.0f / TC_H); s2 *= (1.0f / TC_H);
        reinterpret_cast<float4*>(dY + row * TC_H)[lane] =
            make_float4((g0 - s1 - x0 * s2) * rstd, (g1 - s1 - x1 * s2) * rstd, (g2 - s1 - x2 * s2) * rstd, (g3 - s1 - x3 * s2) * rstd);
        ag.x = fmaf(du.x, x0, ag.x); ag.y = fmaf(du.y, x1, ag.y); ag.z = fmaf(du.z, x2, ag.z); ag.w = fmaf(du.w, x3, ag.w);
        ab.x += du.x; ab.y += du.y; ab.z += du.z; ab.w += du.w;
    }
    reinterpret_cast<float4*>(sg[warp])[lane] = ag;
    reinterpret_cast<float4*>(sb[warp])[lane] = ab;
    __syncthreads();
    for (int c = threadIdx.x; c < TC_H; c += blockDim.x) {
        float a = 0.0f, b = 0.0f;
        for (int w = 0; w < LNB_WARPS; ++w) { a += sg[w][c]; b += sb[w][c]; }
        partials[(size_t)blockIdx.x * 2 * TC_H + c] = a;
        partials[(size_t)blockIdx.x * 2 * TC_H + TC_H + c] = b;
    }
}

__global__ void ln_bwd_reduce_kernel(const float* __restrict__ partials, int n_blocks, float* __restrict__ dgamma,
                                     float* __restrict__ dbeta, int accumulate) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= 2 * TC_H) return;
    float s = 0.0f;
    for (int b = 0; b < n_blocks; ++b) s += partials[(size_t)b * 2 * TC_H + c];
    float* base = c < TC_H ? dgamma : dbeta;
    if (base == nullptr) return;
    float* d = base + (c < TC_H ? c : c - TC_H);
    *d = accumulate ? *d + s : s;
}

constexpr int LNB_BLOCKS = 592;       // 4 per SM

}  // namespace

int64_t wgrad_workspace_bytes() { return align_up((int64_t)148 * 128 * PW * 4, 256); }

int run_wgrad(int ns, const float* X, const float* A, int64_t rows, float* dW, int ld, int col0, float* db,
              int accumulate, void* ws, cudaStream_t stream, int nrows, int ncols, int a_cols, int x16) {
    CGNN_CHECK_ARG(X && A && dW && ws && rows >= 1, "tensor-core wgrad: bad arguments");
    const int nsi = ns == 3 ? 2 : 1;
    const int64_t n_tiles = (rows + TR - 1) / TR;
    int grid = num_sms() < 148 ? num_sms() : 148;
    if (n_tiles < grid) grid = (int)n_tiles;
    CUtensorMap mx, ma;
    int rc;
    if ((rc = x16 ? make_row_map_bf16(&mx, X, rows, TR) : make_row_map32_rows(&mx, X, rows, TR))) return rc;
    if ((rc = make_row_map32_rows(&ma, A, rows, TR, a_cols > 0 ? a_cols : TC_H))) return rc;
    const size_t smem = (size_t)WgSmem::ops + (size_t)2 * 2 * nsi * OP_BYTES;           // two image sets
    float* partials = static_cast<float*>(ws);
    void (*kern)(CUtensorMap, CUtensorMap, int64_t, float*) =
        ns == 3 ? (x16 ? tc_wgrad_kernel<3, true> : tc_wgrad_kernel<3, false>) : (x16 ? tc_wgrad_kernel<1, true> : tc_wgrad_kernel<1, false>);
    static bool configured[4] = {false, false, false, false};
    const int slot = (ns == 3 ? 2 : 0) + (x16 ? 1 : 0);
    if (!configured[slot]) { CGNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); configured[slot] = true; }
    kern<<<grid, WG_THREADS, smem, stream>>>(mx, ma, n_tiles, partials);
    CGNN_LAUNCH_CHECK();
    wgrad_reduce_kernel<<<(128 * 129 + 63) / 64, 256, 0, stream>>>(partials, grid, dW, ld, col0, db, accumulate,
                                                                     nrows > 0 ? nrows : 128, ncols > 0 ? ncols : 128);
    CGNN_LAUNCH_CHECK();
    return CGNN_OK;
}

// also holds the per-warp partials of the chain kernel's LayerNorm-backward epilogue (148 CTAs x 8 warps)
int64_t ln_bwd_workspace_bytes() { return align_up((int64_t)(LNB_BLOCKS > 148 * 8 ? LNB_BLOCKS : 148 * 8) * 2 * TC_H * 4, 256); }

// dY may alias Y.  dgamma / dbeta: written (or accumulated into) after a fixed-order reduction.
int run_ln_bwd(const float* Y, const float* dU_rows, const float* dU_recv, int k, const float* gamma, int64_t rows,
               float* dY, float* dgamma, float* dbeta, int accumulate, void* ws, cudaStream_t stream) {
    CGNN_CHECK_ARG(Y && gamma && dY && ws && rows >= 1 && (dU_rows || dU_recv), "LayerNorm backward: bad arguments");
    int64_t want = (rows + LNB_WARPS - 1) / LNB_WARPS;
    const int grid = (int)(want < LNB_BLOCKS ? want : LNB_BLOCKS);
    float* partials = static_cast<float*>(ws);
    ln_bwd_kernel<<<grid, LNB_WARPS * 32, 0, stream>>>(Y, dU_rows, dU_recv, k, gamma, rows, dY, partials);
    CGNN_LAUNCH_CHECK();
    ln_bwd_reduce_kernel<<<1, 256, 0, stream>>>(partials, grid, dgamma, dbeta, accumulate);
    CGNN_LAUNCH_CHECK();
    return CGNN_OK;
}

}  // namespace cgnn
