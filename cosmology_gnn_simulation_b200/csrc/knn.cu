// K1: periodic-box k-NN on a cell list, one warp per query, warp-level top-k.
//
// Replaces data_utils.py:148-152 (extend_positions_torch + torch_cluster.knn + index remap):
// instead of materialising 27N ghost points and building a KD-tree on the CPU, the N particles
// are counting-sorted into a uniform grid and each query warp walks the periodic images of the
// surrounding cells ring by ring.  Results are identical to the exhaustive search over all 27N
// candidates under the total order (d2, c) of SURVEY App. A.2:
//   * distances use the reference's fp32 recipe with explicit _rn intrinsics (no FMA contraction),
//   * 64-bit keys (float_bits(d2) << 32 | c) make the order independent of the visiting order,
//   * a ring is only skipped when its cells provably cannot hold a key below the current k-th.
// Clustered boxes: the particles are binned into a small pyramid of grids (cell width halving per level)
// and every query searches on the coarsest level whose own cell holds at most ~k particles, so a dense
// clump is walked on a fine grid and a void on the coarse one -- same exhaustive-up-to-the-bound search,
// same result, but O(k) candidates per ring instead of O(clump size).
#include "common.cuh"
#include "scan.cuh"

namespace cgnn {

namespace {

constexpr int KNN_THREADS = 256;
constexpr unsigned long long KEY_INF = 0xFFFFFFFFFFFFFFFFull;

__host__ __device__ inline int cells_per_dim(int64_t n, int k) {
    // mean cell occupancy ~0.36*k makes the first 3x3x3 block sufficient for most queries
    double occ = 0.36 * (double)k;
    if (occ < 2.0) occ = 2.0;
    double c = cbrt((double)n / occ);
    int nc = (int)c;
    if (nc < 1) nc = 1;
    if (nc > 400) nc = 400;
    return nc;
}

__device__ __forceinline__ int cell_coord(float x, float inv_w, int nc) {
    int c = (int)(x * inv_w);
    return min(max(c, 0), nc - 1);
}

__global__ void knn_count_cells(const float* __restrict__ pos, int64_t n, float inv_w, int nc,
                                int* __restrict__ cell_of, int* __restrict__ count, const int* __restrict__ enable) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n || (enable != nullptr && *enable == 0)) return;
    int cx = cell_coord(pos[3 * i + 0], inv_w, nc);
    int cy = cell_coord(pos[3 * i + 1], inv_w, nc);
    int cz = cell_coord(pos[3 * i + 2], inv_w, nc);
    int c = (cx * nc + cy) * nc + cz;
    cell_of[i] = c;
    atomicAdd(&count[c], 1);
}

__global__ void knn_scatter(const float* __restrict__ pos, int64_t n, const int* __restrict__ cell_of,
                            int* __restrict__ cursor, float4* __restrict__ sorted, const int* __restrict__ enable) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n || (enable != nullptr && *enable == 0)) return;
    int slot = atomicAdd(&cursor[cell_of[i]], 1);
    sorted[slot] = make_float4(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], __int_as_float((int)i));
}

// need_next = 1 when some cell of this level holds more than 4k particles (a finer level pays off; below that the
// brute-force scan of the 27 cells is cheaper than building another grid)
__global__ void knn_crowded(const int* __restrict__ cell_of, const int* __restrict__ count, int64_t n, int k,
                            const int* __restrict__ enable, int* __restrict__ need_next) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n || (enable != nullptr && *enable == 0)) return;
    if (count[cell_of[i]] > 4 * k) *need_next = 1;      // benign race: every writer stores the same value
}

// ---- warp-level top-k ------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long shfl64(unsigned long long v, int src) {
    return __shfl_sync(FULL, v, src);
}
__device__ __forceinline__ unsigned long long shfl_xor64(unsigned long long v, int m) {
    return __shfl_xor_sync(FULL, v, m);
}

// `best`: lane l holds the l-th smallest key seen so far (ascending over lanes).
// Merges one new key per lane (KEY_INF = none).
__device__ __forceinline__ void topk_merge(unsigned long long& best, unsigned long long key, int k, int lane) {
    unsigned long long tau = shfl64(best, k - 1);
    if (!__any_sync(FULL, key < tau)) return;
    // bitonic sort of the 32 new keys, ascending over lanes
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            unsigned long long other = shfl_xor64(key, stride);
            bool up = (lane & size) == 0;
            bool lower = (lane & stride) == 0;
            unsigned long long mn = key < other ? key : other;
            unsigned long long mx = key < other ? other : key;
            key = (lower == up) ? mn : mx;
        }
    }
    // min(ascending, descending) keeps the 32 smallest of the union as a bitonic sequence
    unsigned long long rev = shfl64(key, 31 - lane);
    unsigned long long m = best < rev ? best : rev;
#pragma unroll
    for (int stride = 16; stride > 0; stride >>= 1) {
        unsigned long long other = shfl_xor64(m, stride);
        bool lower = (lane & stride) == 0;
        unsigned long long mn = m < other ? m : other;
        unsigned long long mx = m < other ? other : m;
        m = lower ? mn : mx;
    }
    best = m;
}

// Visits the extended cells of the cube of radius R around (cx,cy,cz) whose Chebyshev distance is
// >= r_min (r_min = 0: the whole cube; r_min = R: only the outer shell).
__device__ __forceinline__ void visit_cells(const float4* __restrict__ sorted, const int* __restrict__ cell_start,
                                            int nc, int64_t n, float box, float qx, float qy, float qz,
                                            int cx, int cy, int cz, int R, int r_min, int k, int lane,
                                            unsigned long long& best) {
    const int S = 2 * R + 1;
    const int S3 = S * S * S;
    for (int base = 0; base < S3; base += 32) {
        int ci = base + lane;
        int dx = ci / (S * S) - R, dy = (ci / S) % S - R, dz = ci % S - R;
        int ex = cx + dx, ey = cy + dy, ez = cz + dz;
        bool valid = ci < S3 && max(max(abs(dx), abs(dy)), abs(dz)) >= r_min &&
                     ex >= -nc && ex < 2 * nc && ey >= -nc && ey < 2 * nc && ez >= -nc && ez < 2 * nc;
        int sx = ex < 0 ? 0 : (ex >= nc ? 2 : 1);
        int sy = ey < 0 ? 0 : (ey >= nc ? 2 : 1);
        int sz = ez < 0 ? 0 : (ez >= nc ? 2 : 1);
        int start = 0, cnt = 0;
        if (valid) {
            int rc = ((ex - (sx - 1) * nc) * nc + (ey - (sy - 1) * nc)) * nc + (ez - (sz - 1) * nc);
            start = cell_start[rc];
            cnt = cell_start[rc + 1] - start;
        }
        int shift_id = sx * 9 + sy * 3 + sz;
        // inclusive scan of cnt over lanes
        int inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(FULL, inc, o);
            if (lane >= o) inc += t;
        }
        int total = __shfl_sync(FULL, inc, 31);
        for (int t0 = 0; t0 < total; t0 += 32) {
            int t = t0 + lane;
            // find j = first lane with inc_j > t  (binary search over the warp's scan)
            int j = 0;
#pragma unroll
            for (int step = 16; step > 0; step >>= 1) {
                int probe = __shfl_sync(FULL, inc, j + step - 1);
                if (probe <= t) j += step;
            }
            int inc_j = __shfl_sync(FULL, inc, j);
            int cnt_j = __shfl_sync(FULL, cnt, j);
            int start_j = __shfl_sync(FULL, start, j);
            int sid_j = __shfl_sync(FULL, shift_id, j);
            unsigned long long key = KEY_INF;
            if (t < total) {
                int p = start_j + (t - (inc_j - cnt_j));
                float4 c = sorted[p];
                float shx = (float)(sid_j / 9 - 1) * box;
                float shy = (float)((sid_j / 3) % 3 - 1) * box;
                float shz = (float)(sid_j % 3 - 1) * box;
                // fl(fl(pos + shift) - query): exactly the reference's ghost construction + difference
                float ddx = __fsub_rn(__fadd_rn(c.x, shx), qx);
                float ddy = __fsub_rn(__fadd_rn(c.y, shy), qy);
                float ddz = __fsub_rn(__fadd_rn(c.z, shz), qz);
                float d2 = __fadd_rn(__fadd_rn(__fmul_rn(ddx, ddx), __fmul_rn(ddy, ddy)), __fmul_rn(ddz, ddz));
                unsigned cidx = (unsigned)((int64_t)sid_j * n + (int64_t)__float_as_int(c.w));
                key = ((unsigned long long)__float_as_uint(d2) << 32) | cidx;
            }
            topk_merge(best, key, k, lane);
        }
    }
}

constexpr int KNN_MAX_LEVELS = 5;
constexpr int64_t KNN_MAX_CELLS = 1ll << 23;        // per refined level (32 MiB per int array)

struct KnnLevels {
    int n_levels;
    const int* need;            // need[l] != 0: level l was built
    int nc[KNN_MAX_LEVELS];
    const float4* sorted[KNN_MAX_LEVELS];
    const int* start[KNN_MAX_LEVELS];
};

__global__ void __launch_bounds__(KNN_THREADS)
knn_query(KnnLevels lv, int64_t n, float box, int k, int64_t q0, int64_t nq, int32_t* __restrict__ nbr_ext) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = (gridDim.x * (int64_t)blockDim.x) >> 5;
    const float margin = 1e-5f * box;
    for (int64_t q = warp0; q < n; q += n_warps) {
        float4 me = lv.sorted[0][q];
        float qx = me.x, qy = me.y, qz = me.z;
        int qi = __float_as_int(me.w);
        if (qi < q0 || qi >= q0 + nq) continue;          // only the queries of this rank's slab (warp-uniform)
        // coarsest level whose own cell is not crowded (warp-uniform: every lane sees the same query)
        int L = 0, nc = lv.nc[0], cx = 0, cy = 0, cz = 0;
        float inv_w = (float)nc / box;
        for (;; ++L) {
            nc = lv.nc[L];
            inv_w = (float)nc / box;
            cx = cell_coord(qx, inv_w, nc); cy = cell_coord(qy, inv_w, nc); cz = cell_coord(qz, inv_w, nc);
            if (L + 1 >= lv.n_levels || lv.need[L + 1] == 0) break;
            const int c = (cx * nc + cy) * nc + cz;
            if (lv.start[L][c + 1] - lv.start[L][c] <= k) break;
        }
        const float4* __restrict__ sorted = lv.sorted[L];
        const int* __restrict__ cell_start = lv.start[L];
        const float w = box / (float)nc;
        unsigned long long best = KEY_INF;
        for (int R = 1;; ++R) {
            visit_cells(sorted, cell_start, nc, n, box, qx, qy, qz, cx, cy, cz, R, R == 1 ? 0 : R, k, lane, best);
            // distance from the query to the nearest face of the explored block that still has
            // unexplored cells behind it
            float fd = 3.0e38f;
            bool open = false;
            if (cx - R > -nc)      { fd = fminf(fd, qx - (float)(cx - R) * w);     open = true; }
            if (cx + R < 2 * nc - 1) { fd = fminf(fd, (float)(cx + R + 1) * w - qx); open = true; }
            if (cy - R > -nc)      { fd = fminf(fd, qy - (float)(cy - R) * w);     open = true; }
            if (cy + R < 2 * nc - 1) { fd = fminf(fd, (float)(cy + R + 1) * w - qy); open = true; }
            if (cz - R > -nc)      { fd = fminf(fd, qz - (float)(cz - R) * w);     open = true; }
            if (cz + R < 2 * nc - 1) { fd = fminf(fd, (float)(cz + R + 1) * w - qz); open = true; }
            if (!open) break;                       // all 27 images exhausted
            unsigned long long kth = shfl64(best, k - 1);
            if (kth != KEY_INF) {
                float fds = fd - margin;
                float kd2 = __uint_as_float((unsigned)(kth >> 32));
                if (fds > 0.0f && kd2 <= fds * fds * 0.99999f) break;
            }
        }
        if (lane < k) nbr_ext[((int64_t)qi - q0) * k + lane] = (int32_t)(unsigned)(best & 0xFFFFFFFFull);
    }
}

struct KnnPlan {
    int n_levels;
    int nc[KNN_MAX_LEVELS];
    int64_t n_cells[KNN_MAX_LEVELS], n_tiles_max;
};
// level 0 is the uniform-box grid of cells_per_dim(); every further level halves the cell width while the grid
// stays below KNN_MAX_CELLS cells
KnnPlan plan_for(int64_t n, int k) {
    KnnPlan p;
    int nc = cells_per_dim(n, k);
    p.n_levels = 0;
    p.n_tiles_max = 0;
    while (p.n_levels < KNN_MAX_LEVELS) {
        const int64_t cells = (int64_t)nc * nc * nc;
        if (p.n_levels > 0 && cells > KNN_MAX_CELLS) break;
        p.nc[p.n_levels] = nc;
        p.n_cells[p.n_levels] = cells;
        const int64_t t = scan_tiles(cells + 1);
        if (t > p.n_tiles_max) p.n_tiles_max = t;
        ++p.n_levels;
        if (nc * 2 > 2000) break;
        nc *= 2;
    }
    return p;
}

// workspace layout for the worst case over k (k = 1 gives the finest level-0 grid and the largest pyramid)
struct KnnCarve {
    int* count; int* cursor; int* tile_sum; int* cell_of; int* need;
    int* start[KNN_MAX_LEVELS];
    float4* sorted[KNN_MAX_LEVELS];
    int64_t bytes;
};
KnnCarve carve(void* ws, int64_t n) {
    // any k: level 0 is at most as fine as for k = 1, every refined level has at most KNN_MAX_CELLS cells
    KnnPlan pmax = plan_for(n, 1);
    int64_t max_cells = pmax.n_cells[0] > KNN_MAX_CELLS ? pmax.n_cells[0] : KNN_MAX_CELLS;
    Carver c(ws);
    KnnCarve o;
    o.count = c.take<int>(max_cells + 1);
    o.cursor = c.take<int>(max_cells + 1);
    o.tile_sum = c.take<int>(scan_tiles(max_cells + 1) + 1);
    o.cell_of = c.take<int>(n);
    o.need = c.take<int>(KNN_MAX_LEVELS + 1);
    for (int l = 0; l < KNN_MAX_LEVELS; ++l) {
        o.start[l] = c.take<int>(max_cells + 1);
        o.sorted[l] = c.take<float4>(n);
    }
    o.bytes = c.off;
    return o;
}

}  // namespace

}  // namespace cgnn

using namespace cgnn;

extern "C" int64_t cgnn_knn_workspace_bytes(int64_t n) {
    return carve(nullptr, n).bytes;
}

extern "C" int cgnn_knn_periodic(const float* pos, int64_t n, float box, int32_t k, int32_t* nbr_ext,
                                 void* workspace, int64_t workspace_bytes, cgnn_stream stream_) {
    return cgnn_knn_periodic_range(pos, n, box, k, 0, n, nbr_ext, workspace, workspace_bytes, stream_);
}

extern "C" int cgnn_knn_periodic_range(const float* pos, int64_t n, float box, int32_t k, int64_t q0, int64_t nq,
                                       int32_t* nbr_ext, void* workspace, int64_t workspace_bytes, cgnn_stream stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CGNN_CHECK_ARG(pos && nbr_ext && workspace, "cgnn_knn_periodic: null pointer");
    CGNN_CHECK_ARG(q0 >= 0 && nq >= 0 && q0 + nq <= n, "cgnn_knn_periodic_range: query range [%lld, +%lld) outside [0, %lld)", (long long)q0, (long long)nq, (long long)n);
    if (nq == 0) return CGNN_OK;
    CGNN_CHECK_ARG(n >= 1 && k >= 1 && k <= 32, "cgnn_knn_periodic: need n >= 1 and 1 <= k <= 32 (got n=%lld k=%d)", (long long)n, k);
    CGNN_CHECK_ARG(27 * n >= k, "cgnn_knn_periodic: 27*N = %lld < k = %d", (long long)(27 * n), k);
    CGNN_CHECK_ARG(27 * n < (1ll << 32), "cgnn_knn_periodic: 27*N must fit 32 bits");
    CGNN_CHECK_ARG(box > 0.0f, "cgnn_knn_periodic: box must be positive");
    if (workspace_bytes < cgnn_knn_workspace_bytes(n)) {
        set_error("cgnn_knn_periodic: workspace too small (%lld < %lld)", (long long)workspace_bytes,
                  (long long)cgnn_knn_workspace_bytes(n));
        return CGNN_ERR_WORKSPACE;
    }
    KnnPlan p = plan_for(n, k);
    KnnCarve c = carve(workspace, n);
    KnnLevels lv;
    lv.n_levels = p.n_levels;
    lv.need = c.need;
    int blocks_n = (int)((n + 255) / 256);
    CGNN_CUDA(cudaMemsetAsync(c.need, 0, sizeof(int) * (KNN_MAX_LEVELS + 1), stream));
    for (int l = 0; l < p.n_levels; ++l) {
        // a refined level is only built when the level above it has a crowded cell (device-side flag, no host sync)
        const int* enable = l == 0 ? nullptr : c.need + l;
        const float inv_w = (float)p.nc[l] / box;
        CGNN_CUDA(cudaMemsetAsync(c.count, 0, sizeof(int) * (p.n_cells[l] + 1), stream));
        knn_count_cells<<<blocks_n, 256, 0, stream>>>(pos, n, inv_w, p.nc[l], c.cell_of, c.count, enable);
        CGNN_LAUNCH_CHECK();
        if (l + 1 < p.n_levels) {
            knn_crowded<<<blocks_n, 256, 0, stream>>>(c.cell_of, c.count, n, k, enable, c.need + l + 1);
            CGNN_LAUNCH_CHECK();
        }
        // count[n_cells] == 0, so start[n_cells] = N
        int rc = exclusive_scan_i32(c.count, p.n_cells[l] + 1, c.start[l], c.cursor, c.tile_sum, stream, enable);
        if (rc != CGNN_OK) return rc;
        knn_scatter<<<blocks_n, 256, 0, stream>>>(pos, n, c.cell_of, c.cursor, c.sorted[l], enable);
        CGNN_LAUNCH_CHECK();
        lv.nc[l] = p.nc[l];
        lv.sorted[l] = c.sorted[l];
        lv.start[l] = c.start[l];
    }
    for (int l = p.n_levels; l < KNN_MAX_LEVELS; ++l) { lv.nc[l] = 0; lv.sorted[l] = nullptr; lv.start[l] = nullptr; }
    int64_t want_blocks = (n * 32 + KNN_THREADS - 1) / KNN_THREADS;
    int64_t max_blocks = (int64_t)num_sms() * 8;
    int q_blocks = (int)(want_blocks < max_blocks ? want_blocks : max_blocks);
    knn_query<<<q_blocks, KNN_THREADS, 0, stream>>>(lv, n, box, k, q0, nq, nbr_ext);
    CGNN_LAUNCH_CHECK();
    return CGNN_OK;
}
