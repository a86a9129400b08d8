// Tensor-core (tcgen05) implementation of the fused processor MLP tiles, forward.
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
//     on one, the four epilogue warps of the other do bias/ReLU/split or the LayerNorm epilogue;
//   * the edge phase uses W1 = [W1s | W1r | W1e] (graph_network.py:89 concat order): P_s = h W1s^T and
//     P_r = h W1r^T + b1 are computed once per NODE by this same kernel (a 1-layer chain) and the
//     per-edge layer 1 is e W1e^T + P_s[sender] + P_r[receiver]  (5 L^2 -> 3 L^2 MACs per edge);
//   * inputs stream in through TMA (2-D tensor maps, 16-column boxes, 64-byte swizzle), the sender rows
//     of P_s through per-row bulk copies, outputs leave through TMA stores.
//
// Reference semantics: InteractionNetwork.forward + residuals, graph_network.py:83-101,177-183.
#include "tc_common.cuh"

namespace cgnn {
namespace {

using namespace ptx;

constexpr int TC_THREADS = 320;           // 8 epilogue warps (2 groups of 4) + producer warp + MMA warp
constexpr int MAXRING = 16;               // input ring depth is chosen per launch from the shared memory left (3 .. 16)
constexpr int WIMG = 64 * TC_H * 2;       // bytes of one weight image half (64 output rows x 128 k, bf16)
constexpr int PS_STRIDE = TC_H * 4 + 16;  // padded row stride of the gathered P_s rows (conflict-free row reads)
constexpr int MAX_BLOCKS = 4;             // weight blocks (MMA phases) per tile
constexpr float LN_EPS = 1e-5f;

struct TcParams {
    int64_t n_rows;             // rows of the stream (edges or nodes)
    int64_t n_pair_tiles;       // ceil(n_rows / 256)
    int n_in;                   // input phases (1: in0; 2: in0 then in1)
    int n_layers;               // Linear layers (1 or 3)
    int k;                      // > 0: rows per receiver (gather / segmented sum)
    int has_ln;
    int gather;                 // layer-1 pre-activation += Ps[sender] + Pr[receiver]
    int relu_out;               // ReLU on the result (before mask / residual)
    const float* mask_src;      // [n_rows][128]: result = mask_src > 0 ? result : 0 (nullable)
    const float* residual;      // [n_rows][128] added to the result (nullable)
    float* agg_out;             // edge mode: [n_rows / k][128] per-receiver sum of the result before the residual (nullable)
    const int32_t* senders;     // edge mode
    const float* Ps;            // edge mode: [N][128]
    const float* Pr;            // edge mode: [N][128] (includes the layer-1 bias)
    int n_ring;                 // input ring depth (3 .. MAXRING)
    int n_ps;                   // gather staging buffers: 2 (one per group) when shared memory allows, else 1
    const uint8_t* w_images;    // [n_blocks][NSI][2 halves][WIMG]
    const float* vec;           // [5][128]: bias of layer 1, 2, 3, gamma, beta
};

// shared-memory layout (dynamic, 1024-byte aligned base)
struct Smem {
    static constexpr int sbuf = 0;                                 // 2 groups * 2 * CH_BYTES (output staging)
    static constexpr int vec = sbuf + 4 * CH_BYTES;                // 5 * 128 floats
    static constexpr int bars = vec + 5 * TC_H * 4;                // barriers (512 bytes)
    static constexpr int weights = bars + 512;                     // n_blocks * NSI * WIMG
    // then: n_ps * 128 * PS_STRIDE gather staging, then (1024-aligned) n_ring * CH_BYTES input ring
};
static_assert(Smem::weights % 128 == 0 && Smem::sbuf % 1024 == 0, "chunk buffers need 1024-byte, weight images 128-byte alignment");
__host__ __device__ constexpr uint32_t ring_offset(int n_blocks, int nsi, int n_ps) {
    return (uint32_t)((Smem::weights + n_blocks * nsi * WIMG + n_ps * 128 * PS_STRIDE + 1023) / 1024 * 1024);
}

struct Bars {
    uint64_t w_full;
    // "full" barriers are per consumer group: a group only ever waits on barriers whose uses are all its own,
    // so the phase parity it tracks can never alias a phase that belongs to the other group's tiles
    uint64_t in_full[2][MAXRING], in_empty[MAXRING];
    uint64_t ps_full[2], ps_empty[2];
    uint64_t a_ready[2];        // used in the leader CTA: 8 arrivals (4 warps x 2 CTAs)
    uint64_t mma_done[2];       // per CTA, one tcgen05.commit arrival
    uint32_t tmem_base;
};

template <int NS>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_chain_fwd(const __grid_constant__ CUtensorMap tm_in0, const __grid_constant__ CUtensorMap tm_in1,
             const __grid_constant__ CUtensorMap tm_out, const TcParams p) {
    constexpr int NSI = NS == 3 ? 2 : 1;                    // weight / activation images per value (hi [, lo])
    extern __shared__ __align__(1024) uint8_t smem[];
    Bars* bars = reinterpret_cast<Bars*>(smem + Smem::bars);
    float* sVec = reinterpret_cast<float*>(smem + Smem::vec);
    const int n_blocks = p.n_in + p.n_layers - 1;           // MMA phases per tile
    uint8_t* sW = smem + Smem::weights;
    uint8_t* sPs0 = sW + n_blocks * NSI * WIMG;             // gather only: n_ps * 128 * PS_STRIDE
    uint8_t* sRing = smem + ring_offset(n_blocks, NSI, p.gather ? p.n_ps : 0);
    const int NRING = p.n_ring;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const int64_t cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    const int64_t n_it = p.n_pair_tiles > cluster_id ? (p.n_pair_tiles - cluster_id + n_clusters - 1) / n_clusters : 0;
    const bool gather_on = p.gather != 0;

    // ---- one-time setup ---------------------------------------------------------------------------
    if (tid == 0) {
        mbar_init(&bars->w_full, 1);
        for (int i = 0; i < MAXRING; ++i) {
            mbar_init(&bars->in_full[0][i], 1);
            mbar_init(&bars->in_full[1][i], 1);
            mbar_init(&bars->in_empty[i], 4);
        }
        mbar_init(&bars->ps_full[0], 1);
        mbar_init(&bars->ps_full[1], 1);
        mbar_init(&bars->ps_empty[0], 4);
        mbar_init(&bars->ps_empty[1], 4);
        for (int s = 0; s < 2; ++s) { mbar_init(&bars->a_ready[s], 8); mbar_init(&bars->mma_done[s], 1); }
        fence_mbar_init();
    }
    for (int i = tid; i < 5 * TC_H; i += TC_THREADS) sVec[i] = p.vec[i];
    if (warp == 9) tmem_alloc<2>(&bars->tmem_base, 512);
    if (warp == 8 && lane == 0) {
        prefetch_tmap(&tm_in0);
        prefetch_tmap(&tm_in1);
        prefetch_tmap(&tm_out);
    }
    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after_sync();
    const uint32_t tmem = bars->tmem_base;

    if (warp == 8) {
        // ============================ producer: weights, input chunks, sender gathers ============================
        if (lane == 0) {
            mbar_expect_tx(&bars->w_full, (uint32_t)(n_blocks * NSI * WIMG));
            for (int b = 0; b < n_blocks; ++b)
                for (int sp = 0; sp < NSI; ++sp)
                    bulk_g2s(sW + (b * NSI + sp) * WIMG, p.w_images + ((size_t)(b * NSI + sp) * 2 + rank) * WIMG, WIMG, &bars->w_full);
        }
        uint32_t seq = 0;
        for (int64_t it = 0; it < n_it; ++it) {
            const int64_t row0 = ((cluster_id + it * n_clusters) * 2 + rank) * 128;
            if (lane == 0) {
                for (int ip = 0; ip < p.n_in; ++ip)
                    for (int q = 0; q < NCH; ++q, ++seq) {
                        const uint32_t buf = seq % NRING, use = seq / NRING;
                        mbar_wait_or_trap(&bars->in_empty[buf], (use & 1) ^ 1, 100 + buf);
                        uint64_t* full = &bars->in_full[it & 1][buf];
                        mbar_expect_tx(full, CH_BYTES);
                        tma_load_2d(sRing + buf * CH_BYTES, ip == 0 ? &tm_in0 : &tm_in1, q * CH, (int)row0, full);
                    }
            }
            if (gather_on) {
                int64_t valid = p.n_rows - row0;
                valid = valid < 0 ? 0 : (valid > 128 ? 128 : valid);
                int32_t snd[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) snd[j] = lane + 32 * j < (int)valid ? p.senders[row0 + lane + 32 * j] : 0;
                const int pb = (int)(it & (p.n_ps - 1));                    // staging buffer
                const int64_t use = p.n_ps == 2 ? (it >> 1) : it;           // how often it has been used before
                if (lane == 0) {
                    mbar_wait_or_trap(&bars->ps_empty[pb], (uint32_t)((use & 1) ^ 1), 110);
                    mbar_expect_tx(&bars->ps_full[it & 1], (uint32_t)(valid * TC_H * 4));
                }
                __syncwarp();
                uint8_t* dstPs = sPs0 + pb * (128 * PS_STRIDE);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int r = lane + 32 * j;
                    if (r < (int)valid) bulk_g2s(dstPs + r * PS_STRIDE, p.Ps + (size_t)snd[j] * TC_H, TC_H * 4, &bars->ps_full[it & 1]);
                }
            }
        }
    } else if (warp == 9) {
        // ============================ MMA issuer (leader CTA, one thread) =======================================
        if (rank == 0 && lane == 0) {
            mbar_wait_or_trap(&bars->w_full, 0, 120);
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
                    if (!mbar_try_wait(&bars->a_ready[s], par[s])) {
                        if (++spins > (1u << 27)) { printf("cgnn: MMA issuer timeout slot %d it %lld ph %d\n", s, (long long)it_s[s], ph_s[s]); __trap(); }
                        continue;
                    }
                    spins = 0;
                    tc_fence_after_sync();
                    const int ph = ph_s[s];
                    const uint32_t d = tmem + s * 256;
                    const uint32_t a_hi = d + 128, a_lo = d + 192;
                    const uint32_t w_hi = w_base + (ph * NSI) * WIMG, w_lo = w_hi + WIMG;
                    uint32_t acc = (ph > 0 && ph < p.n_in) ? 1u : 0u;       // later input phases accumulate into layer 1
#pragma unroll
                    for (int ks = 0; ks < TC_H / 16; ++ks) {
                        umma_bf16_ts<2>(d, a_hi + ks * 8, umma_desc(w_hi + ks * 256, 128, TC_H * 16), idesc, acc);
                        acc = 1u;
                    }
                    if (NS == 3) {
#pragma unroll
                        for (int ks = 0; ks < TC_H / 16; ++ks)
                            umma_bf16_ts<2>(d, a_lo + ks * 8, umma_desc(w_hi + ks * 256, 128, TC_H * 16), idesc, 1u);
#pragma unroll
                        for (int ks = 0; ks < TC_H / 16; ++ks)
                            umma_bf16_ts<2>(d, a_hi + ks * 8, umma_desc(w_lo + ks * 256, 128, TC_H * 16), idesc, 1u);
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
    } else {
        // ============================ epilogue groups: thread = row ==============================================
        const int g = warp >> 2;                        // group = TMEM slot
        const int wq = warp & 3;                        // lane quarter this warp may touch
        const int r = wq * 32 + lane;                   // row inside the CTA's 128-row tile
        const int gt = tid - g * 128;                   // thread index inside the group
        const uint32_t tslot = tmem + ((uint32_t)(wq * 32) << 16) + g * 256;
        const uint32_t tD = tslot, tAhi = tslot + 128, tAlo = tslot + 192;
        uint8_t* sS = smem + Smem::sbuf + g * 2 * CH_BYTES;
        uint32_t pm = 0;                                // parity of mma_done[g]
        uint32_t in_par = 0;                            // bit b: parity of this group's next use of in_full[g][b]
        const int bar_id = 1 + g;
        mbar_wait_or_trap(&bars->w_full, 0, 130);       // a_ready from this CTA also tells the leader its weights landed

        for (int64_t it = g; it < n_it; it += 2) {
            const int64_t row0 = ((cluster_id + it * n_clusters) * 2 + rank) * 128;
            uint32_t seq = (uint32_t)it * (uint32_t)(NCH * p.n_in);
            const int pb = (int)(it & (p.n_ps - 1));
            const uint8_t* sPs = sPs0 + pb * (128 * PS_STRIDE);
            // ---- input phases: stream chunks, split, write the A operand -------------------------------------
            for (int ip = 0; ip < p.n_in; ++ip) {
                if (ip > 0) {                            // A is still being read by the previous phase's MMA
                    mbar_wait_or_trap(&bars->mma_done[g], pm, 140);
                    pm ^= 1u;
                    tc_fence_after_sync();
                }
                for (int q = 0; q < NCH; ++q, ++seq) {
                    const uint32_t buf = seq % NRING;
                    mbar_wait_or_trap(&bars->in_full[g][buf], (in_par >> buf) & 1u, 150 + buf);
                    in_par ^= 1u << buf;
                    const uint8_t* src = sRing + buf * CH_BYTES;
                    uint32_t hi[8], lo[8];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 v = *reinterpret_cast<const float4*>(src + swz64(r, j));
                        split2(v.x, v.y, hi[2 * j], lo[2 * j]);
                        split2(v.z, v.w, hi[2 * j + 1], lo[2 * j + 1]);
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive_local(&bars->in_empty[buf]);
                    tmem_st_32x32b_x8(tAhi + q * 8, hi);
                    if (NS == 3) tmem_st_32x32b_x8(tAlo + q * 8, lo);
                }
                tmem_st_wait();
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(&bars->a_ready[g], 0);
            }
            // ---- hidden layers ------------------------------------------------------------------------------------
            for (int l = 0; l + 1 < p.n_layers; ++l) {
                mbar_wait_or_trap(&bars->mma_done[g], pm, 160 + l);
                pm ^= 1u;
                tc_fence_after_sync();
                const bool gather = gather_on && l == 0;
                if (gather) mbar_wait_or_trap(&bars->ps_full[g], (uint32_t)((it >> 1) & 1), 170);
                const float* bias = sVec + l * TC_H;
                int64_t grow = row0 + r;
                const float* pr_row = nullptr;
                if (gather) {
                    if (grow >= p.n_rows) grow = p.n_rows - 1;       // clamp (results of padded rows are never stored)
                    pr_row = p.Pr + (size_t)(grow / p.k) * TC_H;
                }
#pragma unroll 1
                for (int c0 = 0; c0 < TC_H; c0 += 32) {
                    float v[32];
                    float4 prv[8];
                    if (gather) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) prv[j] = __ldg(reinterpret_cast<const float4*>(pr_row + c0 + 4 * j));
                    }
                    tmem_ld_32x32b_x32(tD + c0, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 b = *reinterpret_cast<const float4*>(bias + c0 + j);
                        v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
                    }
                    if (gather) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 a = *reinterpret_cast<const float4*>(sPs + r * PS_STRIDE + (c0 + j) * 4);
                            const float4 b = prv[j >> 2];
                            v[j] += a.x + b.x; v[j + 1] += a.y + b.y; v[j + 2] += a.z + b.z; v[j + 3] += a.w + b.w;
                        }
                    }
                    uint32_t hi[16], lo[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) split2(fmaxf(v[2 * j], 0.0f), fmaxf(v[2 * j + 1], 0.0f), hi[j], lo[j]);
                    tmem_st_32x32b_x16(tAhi + c0 / 2, hi);
                    if (NS == 3) tmem_st_32x32b_x16(tAlo + c0 / 2, lo);
                }
                tmem_st_wait();
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) {
                    if (gather) mbar_arrive_local(&bars->ps_empty[pb]);
                    mbar_arrive_cluster(&bars->a_ready[g], 0);
                }
            }
            // ---- final layer: bias, [gather], [LayerNorm], [ReLU], [mask], segmented sum, residual, store -------
            mbar_wait_or_trap(&bars->mma_done[g], pm, 180);
            pm ^= 1u;
            tc_fence_after_sync();
            const float* bias = sVec + (p.n_layers - 1) * TC_H;
            const bool fgather = gather_on && p.n_layers == 1;      // one-layer chain: the gather lands here
            const float* pr_row = nullptr;
            if (fgather) {
                mbar_wait_or_trap(&bars->ps_full[g], (uint32_t)((it >> 1) & 1), 171);
                int64_t grow = row0 + r;
                if (grow >= p.n_rows) grow = p.n_rows - 1;
                pr_row = p.Pr + (size_t)(grow / p.k) * TC_H;
            }
            float mean = 0.0f, rstd = 1.0f;
            if (p.has_ln) {
                float s = 0.0f;
#pragma unroll 1
                for (int c0 = 0; c0 < TC_H; c0 += 32) {
                    float v[32];
                    tmem_ld_32x32b_x32(tD + c0, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 b = *reinterpret_cast<const float4*>(bias + c0 + j);
                        s += (v[j] + b.x) + (v[j + 1] + b.y) + (v[j + 2] + b.z) + (v[j + 3] + b.w);
                    }
                }
                mean = s * (1.0f / TC_H);
                float q2 = 0.0f;
#pragma unroll 1
                for (int c0 = 0; c0 < TC_H; c0 += 32) {
                    float v[32];
                    tmem_ld_32x32b_x32(tD + c0, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 b = *reinterpret_cast<const float4*>(bias + c0 + j);
                        const float d0 = v[j] + b.x - mean, d1 = v[j + 1] + b.y - mean, d2 = v[j + 2] + b.z - mean, d3 = v[j + 3] + b.w - mean;
                        q2 = fmaf(d0, d0, q2); q2 = fmaf(d1, d1, q2); q2 = fmaf(d2, d2, q2); q2 = fmaf(d3, d3, q2);
                    }
                }
                rstd = 1.0f / sqrtf(q2 * (1.0f / TC_H) + LN_EPS);
            }
            const float* gamma = sVec + 3 * TC_H;
            const float* beta = sVec + 4 * TC_H;
            const int pj = gt & 3;                           // coalesced passes: 4 threads per row (16 bytes each), 32 rows per pass
#pragma unroll 1
            for (int q = 0; q < NCH; q += 2) {                // two 16-column chunks (both staging buffers) per iteration
                // operands of the coalesced passes and of the gather: issued first, consumed after the TMEM read
                float4 mk[2][4], rs[2][4], prv[8];
                if (p.mask_src != nullptr) {
#pragma unroll
                    for (int c = 0; c < 2; ++c)
#pragma unroll
                        for (int pass = 0; pass < 4; ++pass) {
                            const int64_t grow = row0 + (gt >> 2) + pass * 32;
                            mk[c][pass] = grow < p.n_rows ? __ldg(reinterpret_cast<const float4*>(p.mask_src + grow * TC_H + (q + c) * CH + pj * 4))
                                                          : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                }
                if (p.residual != nullptr) {
#pragma unroll
                    for (int c = 0; c < 2; ++c)
#pragma unroll
                        for (int pass = 0; pass < 4; ++pass) {
                            const int64_t grow = row0 + (gt >> 2) + pass * 32;
                            rs[c][pass] = grow < p.n_rows ? __ldg(reinterpret_cast<const float4*>(p.residual + grow * TC_H + (q + c) * CH + pj * 4))
                                                          : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                }
                if (fgather) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) prv[j] = __ldg(reinterpret_cast<const float4*>(pr_row + q * CH + 4 * j));
                }
                float v[32];
                tmem_ld_32x32b_x32(tD + q * CH, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 b = *reinterpret_cast<const float4*>(bias + q * CH + j);
                    v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
                }
                if (fgather) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 a = *reinterpret_cast<const float4*>(sPs + r * PS_STRIDE + (q * CH + j) * 4);
                        const float4 b = prv[j >> 2];
                        v[j] += a.x + b.x; v[j + 1] += a.y + b.y; v[j + 2] += a.z + b.z; v[j + 3] += a.w + b.w;
                    }
                }
                if (p.has_ln) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 gm = *reinterpret_cast<const float4*>(gamma + q * CH + j);
                        const float4 bt = *reinterpret_cast<const float4*>(beta + q * CH + j);
                        v[j] = (v[j] - mean) * rstd * gm.x + bt.x;
                        v[j + 1] = (v[j + 1] - mean) * rstd * gm.y + bt.y;
                        v[j + 2] = (v[j + 2] - mean) * rstd * gm.z + bt.z;
                        v[j + 3] = (v[j + 3] - mean) * rstd * gm.w + bt.w;
                    }
                }
                if (p.relu_out) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
                }
                if (gt == 0) bulk_wait_read<0>();            // the stores of the previous pair have read both buffers
                named_bar_sync(bar_id, 128);
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        *reinterpret_cast<float4*>(sS + c * CH_BYTES + swz64(r, j)) =
                            make_float4(v[16 * c + 4 * j], v[16 * c + 4 * j + 1], v[16 * c + 4 * j + 2], v[16 * c + 4 * j + 3]);
                named_bar_sync(bar_id, 128);
                if (p.mask_src != nullptr) {
#pragma unroll
                    for (int c = 0; c < 2; ++c)
#pragma unroll
                        for (int pass = 0; pass < 4; ++pass) {
                            const int rr = (gt >> 2) + pass * 32;
                            float4* dst = reinterpret_cast<float4*>(sS + c * CH_BYTES + swz64(rr, pj));
                            float4 u = *dst;
                            const float4 m = mk[c][pass];
                            u.x = m.x > 0.0f ? u.x : 0.0f; u.y = m.y > 0.0f ? u.y : 0.0f;
                            u.z = m.z > 0.0f ? u.z : 0.0f; u.w = m.w > 0.0f ? u.w : 0.0f;
                            *dst = u;
                        }
                    if (p.agg_out != nullptr) named_bar_sync(bar_id, 128);
                }
                if (p.agg_out != nullptr) {
                    // per-receiver sum of the k rows, rank order (deterministic): thread <-> (receiver, column of the pair)
                    const int c32 = gt & 31;
                    const uint8_t* Sc = sS + (c32 >> 4) * CH_BYTES;
                    const int c = c32 & 15;
                    const int nrecv = 128 / p.k;
                    for (int rv = gt >> 5; rv < nrecv; rv += 4) {
                        const int64_t recv = row0 / p.k + rv;
                        if (recv * p.k < p.n_rows) {
                            float sum = 0.0f;
                            for (int j = 0; j < p.k; ++j) {
                                const int rr = rv * p.k + j;
                                sum += *reinterpret_cast<const float*>(Sc + swz64(rr, c >> 2) + (c & 3) * 4);
                            }
                            p.agg_out[recv * TC_H + q * CH + c32] = sum;
                        }
                    }
                    if (p.residual != nullptr) named_bar_sync(bar_id, 128);   // agg reads the value before the residual lands
                }
                if (p.residual != nullptr) {
#pragma unroll
                    for (int c = 0; c < 2; ++c)
#pragma unroll
                        for (int pass = 0; pass < 4; ++pass) {
                            const int rr = (gt >> 2) + pass * 32;
                            float4* dst = reinterpret_cast<float4*>(sS + c * CH_BYTES + swz64(rr, pj));
                            float4 u = *dst;
                            const float4 e = rs[c][pass];
                            u.x += e.x; u.y += e.y; u.z += e.z; u.w += e.w;
                            *dst = u;
                        }
                }
                fence_proxy_async_smem();
                named_bar_sync(bar_id, 128);
                if (gt == 0) {
                    tma_store_2d(&tm_out, sS, q * CH, (int)row0);
                    tma_store_2d(&tm_out, sS + CH_BYTES, (q + 1) * CH, (int)row0);
                    bulk_commit();
                }
            }
            if (fgather) {
                __syncwarp();
                if (lane == 0) mbar_arrive_local(&bars->ps_empty[pb]);
            }
            // D and A of this slot are free again: the next tile of this group starts with its input phase
        }
        if (gt == 0) bulk_wait_all<0>();
    }

    // ---- teardown -----------------------------------------------------------------------------------------
    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();
    if (warp == 9) tmem_dealloc<2>(tmem, 512);
}

// Builds the shared-memory weight images of a chain: for block b the K-major BF16 image(s) of the 128 x 128
// matrix B[n][k] = W[(row0 + n) * ld + col0 + k]  (or, transposed, W[(row0 + k) * ld + col0 + n]),
// split by output row n into two halves of 64 rows (one per CTA of the pair).
struct PrepArgs {
    ChainBlock blk[MAX_BLOCKS];
    int n_blocks;
    const float* vec_src[5];      // bias1, bias2, bias3, gamma, beta (nullable -> zeros / ones for gamma)
    int vec_len[5];               // valid entries (zero padded to 128)
};

template <int NS>
__global__ void tc_prep_kernel(PrepArgs a, uint8_t* __restrict__ images, float* __restrict__ vec) {
    constexpr int NSI = NS == 3 ? 2 : 1;
    const int b = blockIdx.y;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;       // one 16-byte piece: (n, k8)
    if (b < a.n_blocks && idx < TC_H * (TC_H / 8)) {
        const int n = idx / (TC_H / 8), k8 = idx % (TC_H / 8);
        const ChainBlock& B = a.blk[b];
        float x[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = k8 * 8 + j;
            const bool valid = n < (B.nmax ? B.nmax : TC_H) && k < (B.kmax ? B.kmax : TC_H);
            x[j] = !valid ? 0.0f : B.transpose ? B.W[(size_t)(B.row0 + k) * B.ld + B.col0 + n] : B.W[(size_t)(B.row0 + n) * B.ld + B.col0 + k];
        }
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) split2(x[2 * j], x[2 * j + 1], hi[j], lo[j]);
        const int half = n >> 6, nl = n & 63;
        const size_t off = (size_t)(nl >> 3) * (TC_H * 16) + (size_t)k8 * 128 + (nl & 7) * 16;
        *reinterpret_cast<uint4*>(images + ((size_t)(b * NSI + 0) * 2 + half) * WIMG + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        if (NS == 3)
            *reinterpret_cast<uint4*>(images + ((size_t)(b * NSI + 1) * 2 + half) * WIMG + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
    if (b == 0 && idx < 5 * TC_H) {
        const int v = idx / TC_H, c = idx % TC_H;
        vec[idx] = (a.vec_src[v] && c < a.vec_len[v]) ? a.vec_src[v][c] : (v == 3 ? 1.0f : 0.0f);
    }
}

constexpr size_t SMEM_MAX = 227 * 1024;

template <int NS>
int run_chain_t(const ChainOp& op, cudaStream_t stream) {
    constexpr int NSI = NS == 3 ? 2 : 1;
    const int n_in = op.in1 ? 2 : 1;
    const int n_blocks = n_in + op.n_layers - 1;
    CGNN_CHECK_ARG(op.n_layers == 1 || op.n_layers == 3, "tensor-core chain: 1 or 3 layers");
    CGNN_CHECK_ARG(n_blocks <= MAX_BLOCKS && op.rows >= 1 && op.in0 && op.out && op.images && op.vec, "tensor-core chain: bad arguments");
    const bool gather = op.Ps != nullptr;
    if (gather || op.agg_out) CGNN_CHECK_ARG(op.k >= 1 && 128 % op.k == 0, "tensor-core chain: k must divide 128 (got %d)", op.k);
    PrepArgs pa{};
    pa.n_blocks = n_blocks;
    for (int b = 0; b < n_blocks; ++b) pa.blk[b] = op.blk[b];
    if (op.n_layers == 1) {
        pa.vec_src[0] = op.bias[0];
    } else {
        pa.vec_src[0] = op.bias[0]; pa.vec_src[1] = op.bias[1]; pa.vec_src[2] = op.bias[2];
    }
    pa.vec_src[3] = op.gamma; pa.vec_src[4] = op.beta;
    for (int v = 0; v < 5; ++v) pa.vec_len[v] = TC_H;
    if (op.out_valid > 0) pa.vec_len[op.n_layers - 1] = op.out_valid;       // bias of a narrow last layer
    dim3 pg((TC_H * (TC_H / 8) + 255) / 256, n_blocks);
    tc_prep_kernel<NS><<<pg, 256, 0, stream>>>(pa, op.images, op.vec);
    CGNN_LAUNCH_CHECK();
    TcParams p{};
    p.n_rows = op.rows; p.n_pair_tiles = (op.rows + 255) / 256;
    p.n_in = n_in; p.n_layers = op.n_layers; p.k = op.k; p.has_ln = op.gamma != nullptr;
    p.gather = gather; p.relu_out = op.relu_out; p.mask_src = op.mask_src; p.residual = op.residual;
    p.agg_out = op.agg_out; p.senders = op.senders; p.Ps = op.Ps; p.Pr = op.Pr;
    p.w_images = op.images; p.vec = op.vec;
    CUtensorMap m0, m1, mo;
    int rc;
    if ((rc = make_row_map(&m0, op.in0, op.rows))) return rc;
    if ((rc = make_row_map(&m1, op.in1 ? op.in1 : op.in0, op.rows))) return rc;
    if ((rc = make_row_map(&mo, op.out, op.rows))) return rc;
    // spend the shared memory the weights leave on a second gather buffer, then on input-ring depth
    p.n_ps = 1;
    if (gather && ring_offset(n_blocks, NSI, 2) + 3 * CH_BYTES <= SMEM_MAX) p.n_ps = 2;
    const size_t ring_off = ring_offset(n_blocks, NSI, gather ? p.n_ps : 0);
    CGNN_CHECK_ARG(ring_off + 3 * CH_BYTES <= SMEM_MAX, "tensor-core chain: shared memory need %zu exceeds 227 KB", ring_off + 3 * CH_BYTES);
    p.n_ring = (int)((SMEM_MAX - ring_off) / CH_BYTES);
    if (p.n_ring > MAXRING) p.n_ring = MAXRING;
    const size_t smem = ring_off + (size_t)p.n_ring * CH_BYTES;
    auto kern = tc_chain_fwd<NS>;
    static size_t configured = 0;
    if (smem > configured) {
        CGNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    int64_t pairs = num_sms() / 2;
    if (p.n_pair_tiles < pairs) pairs = p.n_pair_tiles;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(pairs * 2));
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    CGNN_CUDA(cudaLaunchKernelEx(&cfg, kern, m0, m1, mo, p));
    count_launch();
    return CGNN_OK;
}

}  // namespace

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_row_map(CUtensorMap* m, const float* base, int64_t rows) {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        CGNN_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q));
        CGNN_CHECK_ARG(f != nullptr && q == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled is not available in this driver");
        fn = reinterpret_cast<EncodeTiledFn>(f);
    }
    cuuint64_t dims[2] = {(cuuint64_t)TC_H, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)TC_H * 4};
    cuuint32_t box[2] = {(cuuint32_t)CH, 128};
    cuuint32_t es[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (base %p, rows %lld)", (int)r, (const void*)base, (long long)rows);
        return CGNN_ERR_CUDA;
    }
    return CGNN_OK;
}

int64_t chain_image_bytes() { return (int64_t)MAX_BLOCKS * 2 * 2 * WIMG; }
int64_t chain_vec_bytes() { return align_up(5 * TC_H * 4, 256); }
int run_chain(const ChainOp& op, cudaStream_t stream) {
    return op.ns == 3 ? run_chain_t<3>(op, stream) : run_chain_t<1>(op, stream);
}

}  // namespace cgnn
