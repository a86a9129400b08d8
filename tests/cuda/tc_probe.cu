// Hardware-fact probe for the tcgen05 path (run on a B200: `make -C tests/cuda && tests/cuda/tc_probe <test>`).
// Each test is one tiny kernel with bounded waits; results are compared with a CPU computation.
//   1  SS  cta_group::1   D = A * W^T   (A, W K-major no-swizzle in shared memory)
//   2  TS  cta_group::1   A in TMEM (tcgen05.st, two bf16 per 32-bit column)
//   3  SS  cta_group::2   M = 256 over a CTA pair, W split by output row between the CTAs
//   4  TS  cta_group::2
//   5  SS  cta_group::1   D = A * W     (MN-major descriptor over the same K-major image of W)
//   6  TMA 2-D load with SWIZZLE_128B: dumps the shared-memory image; TMA store back
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../../cosmology_gnn_simulation_b200/csrc/tc_ptx.cuh"

using namespace cgnn::ptx;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

constexpr int KD = 64;      // contraction depth of the probe GEMMs
constexpr int NW = 128;     // output features

// K-major no-swizzle image: 8-row x 16-byte core matrices; all K chunks of an 8-row group contiguous
__device__ __forceinline__ uint32_t kmajor_off(int row, int k, int K) {      // byte offset
    return (uint32_t)((row >> 3) * (K * 16) + (k >> 3) * 128 + (row & 7) * 16 + (k & 7) * 2);
}

template <int CG, bool TS, bool BT>
__global__ void __launch_bounds__(128) mma_probe(const float* __restrict__ A, const float* __restrict__ W,
                                                 float* __restrict__ out, int* __restrict__ flag) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t rank = CG == 2 ? cluster_ctarank() : 0;
    const int nb_rows = BT ? KD : NW / CG;        // rows of the B image this CTA holds
    const int nbt = NW / CG;                      // BT: output features held by this CTA
    uint8_t* sA = smem;                           // 128 x KD bf16
    uint8_t* sB = smem + 128 * KD * 2;            // B image
    // ---- fill shared memory ------------------------------------------------------------------
    for (int i = tid; i < 128 * KD; i += 128) {
        int r = i / KD, k = i % KD;
        float v = A[(size_t)(rank * 128 + r) * KD + k];
        *reinterpret_cast<__nv_bfloat16*>(sA + kmajor_off(r, k, KD)) = __float2bfloat16_rn(v);
    }
    if (!BT) {
        // W[n][k]: this CTA holds output rows n in [rank*nb_rows, +nb_rows)
        for (int i = tid; i < nb_rows * KD; i += 128) {
            int n = i / KD, k = i % KD;
            float v = W[(size_t)(rank * nb_rows + n) * KD + k];
            *reinterpret_cast<__nv_bfloat16*>(sB + kmajor_off(n, k, KD)) = __float2bfloat16_rn(v);
        }
    } else {
        // transposed use: image of Wt[kk][n] (KD rows, NW wide, "K-major" in n); D = A * Wt   (contract over kk)
        for (int i = tid; i < KD * nbt; i += 128) {
            int kk = i / nbt, n = i % nbt;
            float v = W[(size_t)kk * NW + rank * nbt + n];
            *reinterpret_cast<__nv_bfloat16*>(sB + kmajor_off(kk, n, nbt)) = __float2bfloat16_rn(v);
        }
    }
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc<CG>(&tmem_base, 256);
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    if (CG == 2) cluster_sync_all();
    tc_fence_after_sync();
    const uint32_t tb = tmem_base;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    if (TS) {
        // row r = tid: column c (32-bit) holds k = 2c (low half) and k = 2c+1 (high half)
        uint32_t regs[KD / 2];
        for (int c = 0; c < KD / 2; ++c)
            regs[c] = pack_bf16x2(A[(size_t)(rank * 128 + tid) * KD + 2 * c], A[(size_t)(rank * 128 + tid) * KD + 2 * c + 1]);
        tmem_st_32x32b_x16(tb + lane_base + 128, regs);
        tmem_st_32x32b_x16(tb + lane_base + 128 + 16, regs + 16);
        tmem_st_wait();
        tc_fence_before_sync();
        __syncthreads();
        if (CG == 2) cluster_sync_all();
        tc_fence_after_sync();
    }
    if (rank == 0 && tid == 0) {
        const uint32_t idesc = umma_idesc_bf16_major(128 * CG, NW, 0, BT ? 1 : 0);
        for (int ks = 0; ks < KD / 16; ++ks) {
            uint64_t bdesc;
            if (!BT) bdesc = umma_desc(smem_u32(sB) + ks * 256, 128, KD * 16);
            else     bdesc = umma_desc(smem_u32(sB) + ks * 2 * (nbt * 16), /*LBO(MN view)=SBO_K*/ nbt * 16, /*SBO(MN view)=LBO_K*/ 128);
            if (TS) umma_bf16_ts<CG>(tb, tb + 128 + ks * 8, bdesc, idesc, ks > 0);
            else    umma_bf16<CG>(tb, umma_desc(smem_u32(sA) + ks * 256, 128, KD * 16), bdesc, idesc, ks > 0);
        }
        umma_commit<CG>(&bar);
    }
    bool ok = mbar_wait_bounded(&bar, 0, 1u << 22);
    if (!ok && tid == 0) atomicExch(flag, 1);
    tc_fence_after_sync();
    float v[32];
    for (int c0 = 0; c0 < NW; c0 += 32) {
        tmem_ld_32x32b_x32(tb + lane_base + c0, v);
        tmem_ld_wait();
        for (int j = 0; j < 32; ++j) out[(size_t)(rank * 128 + tid) * NW + c0 + j] = v[j];
    }
    tc_fence_before_sync();
    __syncthreads();
    if (CG == 2) cluster_sync_all();
    if (warp == 0) tmem_dealloc<CG>(tb, 256);
}

__global__ void __launch_bounds__(128) tma_probe(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_out,
                                                 float* __restrict__ dump, int* __restrict__ flag) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    float* tile = reinterpret_cast<float*>(smem);          // 128 rows x 32 floats, swizzle 128B
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, 128 * 128);
        tma_load_2d(tile, &tm_in, /*col*/ 32, /*row*/ 128, &bar);
    }
    bool ok = mbar_wait_bounded(&bar, 0, 1u << 22);
    if (!ok && threadIdx.x == 0) atomicExch(flag, 1);
    for (int i = threadIdx.x; i < 128 * 32; i += 128) dump[i] = tile[i];
    // add 1000 to every element in place, then store the tile to `out` at (col 64, row 0)
    for (int i = threadIdx.x; i < 128 * 32; i += 128) tile[i] += 1000.0f;
    fence_proxy_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
        tma_store_2d(&tm_out, tile, 64, 0);
        bulk_commit();
        bulk_wait_all<0>();
    }
    __syncthreads();
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(float* base, int rows, int cols, int box_rows, int box_cols) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = ((EncodeFn)fn)(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(2); }
    return m;
}

static float bf(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

template <int CG, bool TS, bool BT>
static int run_mma(const char* name) {
    const int M = 128 * CG;
    std::vector<float> A((size_t)M * KD), W((size_t)NW * KD), ref((size_t)M * NW), got((size_t)M * NW);
    srand(1);
    for (auto& x : A) x = (rand() % 2001 - 1000) / 500.0f;
    for (auto& x : W) x = (rand() % 2001 - 1000) / 700.0f;
    // BT: W is stored as Wt[kk][n] (KD x NW); else W[n][k]
    for (int r = 0; r < M; ++r)
        for (int n = 0; n < NW; ++n) {
            double s = 0;
            for (int k = 0; k < KD; ++k) s += (double)bf(A[(size_t)r * KD + k]) * (double)bf(BT ? W[(size_t)k * NW + n] : W[(size_t)n * KD + k]);
            ref[(size_t)r * NW + n] = (float)s;
        }
    float *dA, *dW, *dO; int* dF;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dW, W.size() * 4)); CK(cudaMalloc(&dO, got.size() * 4)); CK(cudaMalloc(&dF, 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dO, 0, got.size() * 4)); CK(cudaMemset(dF, 0, 4));
    size_t smem = 128 * KD * 2 + NW * KD * 2;
    auto kern = mma_probe<CG, TS, BT>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(CG); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    CK(cudaLaunchKernelEx(&cfg, kern, (const float*)dA, (const float*)dW, dO, dF));
    CK(cudaDeviceSynchronize());
    int flag = 0;
    CK(cudaMemcpy(got.data(), dO, got.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&flag, dF, 4, cudaMemcpyDeviceToHost));
    double maxerr = 0, maxref = 0; int bad_r = -1, bad_n = -1;
    for (int r = 0; r < M; ++r)
        for (int n = 0; n < NW; ++n) {
            double e = fabs((double)got[(size_t)r * NW + n] - ref[(size_t)r * NW + n]);
            if (e > maxerr) { maxerr = e; bad_r = r; bad_n = n; }
            maxref = fmax(maxref, fabs(ref[(size_t)r * NW + n]));
        }
    printf("%s: timeout=%d max_abs_err=%.3e (max |ref| %.3f) worst at (%d,%d)  %s\n", name, flag, maxerr, maxref, bad_r, bad_n,
           (flag == 0 && maxerr < 1e-3 * maxref) ? "PASS" : "FAIL");
    if (!(flag == 0 && maxerr < 1e-3 * maxref)) {
        for (int r : {0, 1, 8, 127, M - 1}) {
            printf("  row %3d got:", r); for (int n = 0; n < 6; ++n) printf(" %9.4f", got[(size_t)r * NW + n]);
            printf("  | col64..:"); for (int n = 64; n < 68; ++n) printf(" %9.4f", got[(size_t)r * NW + n]);
            printf("\n          ref:"); for (int n = 0; n < 6; ++n) printf(" %9.4f", ref[(size_t)r * NW + n]);
            printf("  | col64..:"); for (int n = 64; n < 68; ++n) printf(" %9.4f", ref[(size_t)r * NW + n]);
            printf("\n");
        }
    }
    return (flag == 0 && maxerr < 1e-3 * maxref) ? 0 : 1;
}

static int run_tma() {
    const int R = 512, C = 128;
    std::vector<float> h((size_t)R * C), dump(128 * 32), back((size_t)R * C);
    for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) h[(size_t)r * C + c] = (float)(r * 128 + c);
    float *dIn, *dOut, *dDump; int* dF;
    CK(cudaMalloc(&dIn, h.size() * 4)); CK(cudaMalloc(&dOut, h.size() * 4)); CK(cudaMalloc(&dDump, dump.size() * 4)); CK(cudaMalloc(&dF, 4));
    CK(cudaMemcpy(dIn, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dOut, 0, h.size() * 4)); CK(cudaMemset(dF, 0, 4));
    CUtensorMap mi = make_map(dIn, R, C, 128, 32), mo = make_map(dOut, R, C, 128, 32);
    CK(cudaFuncSetAttribute(tma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 128));
    tma_probe<<<1, 128, 128 * 128>>>(mi, mo, dDump, dF);
    CK(cudaDeviceSynchronize());
    int flag = 0;
    CK(cudaMemcpy(dump.data(), dDump, dump.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(back.data(), dOut, back.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&flag, dF, 4, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int r = 0; r < 128; ++r)
        for (int c = 0; c < 32; ++c) {
            int idx = r * 32 + (((c >> 2) ^ (r & 7)) << 2) + (c & 3);
            float want = (float)((128 + r) * 128 + 32 + c);
            if (dump[idx] != want) { if (bad < 5) printf("  swizzle mismatch r=%d c=%d: smem[%d]=%.0f want %.0f\n", r, c, idx, dump[idx], want); ++bad; }
        }
    int bad2 = 0;
    for (int r = 0; r < 128; ++r)
        for (int c = 0; c < 32; ++c) {
            float want = (float)((128 + r) * 128 + 32 + c) + 1000.0f;
            if (back[(size_t)r * C + 64 + c] != want) ++bad2;
        }
    printf("tma swizzle128 load: timeout=%d mismatches=%d  store mismatches=%d  %s\n", flag, bad, bad2, (flag == 0 && bad == 0 && bad2 == 0) ? "PASS" : "FAIL");
    if (bad) { printf("  smem row0:"); for (int i = 0; i < 32; ++i) printf(" %.0f", dump[i]); printf("\n  smem row1:"); for (int i = 32; i < 64; ++i) printf(" %.0f", dump[i]); printf("\n"); }
    return (flag == 0 && bad == 0 && bad2 == 0) ? 0 : 1;
}

int main(int argc, char** argv) {
    int t = argc > 1 ? atoi(argv[1]) : 0;
    switch (t) {
        case 1: return run_mma<1, false, false>("1 SS cg1");
        case 2: return run_mma<1, true, false>("2 TS cg1");
        case 3: return run_mma<2, false, false>("3 SS cg2");
        case 4: return run_mma<2, true, false>("4 TS cg2");
        case 5: return run_mma<1, false, true>("5 SS cg1 MN-major B");
        case 6: return run_tma();
        case 7: return run_mma<2, false, true>("7 SS cg2 MN-major B");
    }
    printf("usage: tc_probe <1..7>\n");
    return 64;
}
