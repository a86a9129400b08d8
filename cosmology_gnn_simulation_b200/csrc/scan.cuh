// Exclusive prefix sum of int32 counts (three small kernels; plumbing for the cell list and CSR).
#pragma once
#include "common.cuh"

namespace cgnn {

constexpr unsigned FULL = 0xFFFFFFFFu;

// ---- exclusive scan of `count[m]` into `start[m+1]` (three small kernels) ---------------------
constexpr int SCAN_BLOCK = 1024;
constexpr int SCAN_ITEMS = 4;                         // items per thread
constexpr int SCAN_TILE = SCAN_BLOCK * SCAN_ITEMS;    // 4096 items per block

static __device__ int block_exclusive_scan(int v, int* total, int* smem /* [32] */) {
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) smem[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int w = smem[lane];
        int winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(FULL, winc, o);
            if (lane >= o) winc += t;
        }
        smem[lane] = winc - w;           // exclusive warp offsets
        if (lane == 31) *total = winc;
    }
    __syncthreads();
    int res = inc - v + smem[wid];
    __syncthreads();
    return res;
}

// `enable` (nullable): device flag; a launch whose flag is 0 returns at once (used to skip grid levels nobody needs)
static __global__ void scan_tile_sums(const int* __restrict__ count, int64_t m, int* __restrict__ tile_sum,
                                      const int* __restrict__ enable) {
    __shared__ int smem[32];
    __shared__ int total;
    if (enable != nullptr && *enable == 0) return;
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    int s = 0;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) s += (base + j < m) ? count[base + j] : 0;
    block_exclusive_scan(s, &total, smem);
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = total;
}

static __global__ void scan_tile_offsets(int* __restrict__ tile_sum, int n_tiles, const int* __restrict__ enable) {
    // single block: serial over chunks of SCAN_BLOCK tiles
    __shared__ int smem[32];
    __shared__ int total;
    if (enable != nullptr && *enable == 0) return;
    int carry = 0;
    for (int base = 0; base < n_tiles; base += SCAN_BLOCK) {
        int i = base + threadIdx.x;
        int v = (i < n_tiles) ? tile_sum[i] : 0;
        int ex = block_exclusive_scan(v, &total, smem);
        if (i < n_tiles) tile_sum[i] = ex + carry;
        carry += total;
        __syncthreads();
    }
}

static __global__ void scan_apply(const int* __restrict__ count, int64_t m, const int* __restrict__ tile_off,
                           int* __restrict__ start, int* __restrict__ cursor, const int* __restrict__ enable) {
    __shared__ int smem[32];
    __shared__ int total;
    if (enable != nullptr && *enable == 0) return;
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    int s = 0;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) { v[j] = (base + j < m) ? count[base + j] : 0; s += v[j]; }
    int ex = block_exclusive_scan(s, &total, smem) + tile_off[blockIdx.x];
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
        if (base + j < m) { start[base + j] = ex; if (cursor) cursor[base + j] = ex; }
        ex += v[j];
    }
}


static inline int64_t scan_tiles(int64_t m) { return (m + SCAN_TILE - 1) / SCAN_TILE; }

// start[i] = cursor[i] = sum(count[0..i)) for i in [0, m).  tile_sum: scratch of scan_tiles(m) ints.
// `cursor` may be NULL.
static inline int exclusive_scan_i32(const int* count, int64_t m, int* start, int* cursor, int* tile_sum,
                                     cudaStream_t stream, const int* enable = nullptr) {
    int n_tiles = (int)scan_tiles(m);
    scan_tile_sums<<<n_tiles, SCAN_BLOCK, 0, stream>>>(count, m, tile_sum, enable);
    CGNN_LAUNCH_CHECK();
    scan_tile_offsets<<<1, SCAN_BLOCK, 0, stream>>>(tile_sum, n_tiles, enable);
    CGNN_LAUNCH_CHECK();
    scan_apply<<<n_tiles, SCAN_BLOCK, 0, stream>>>(count, m, tile_sum, start, cursor, enable);
    CGNN_LAUNCH_CHECK();
    return CGNN_OK;
}

}  // namespace cgnn
