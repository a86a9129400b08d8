/* Oracle: periodic k-NN over the 27N ghost-extended points, in plain C.
 *
 * TEST INFRASTRUCTURE ONLY -- the checker and the CPU baseline, never the product.
 *
 * Restates what the reference executes at data_utils.py:148-149:
 *   extend_positions_torch (data_utils.py:9-33)  -> 27 shifted copies, x-slowest shift order
 *   torch_cluster.knn(ext, pos, k)               -> KD-tree (nanoflann) over the 27N points,
 *                                                   one thread, exact k nearest per query
 * torch-cluster 1.6.3 is an un-vendored dependency: its contract is restated, parity unpinned.
 * Distances follow the canonical fp32 recipe of SURVEY App. A.2 (compile with
 * -ffp-contract=off) and the k results are the smallest under the total order (d2, c).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float x, y, z; uint32_t c; } pt_t;
typedef struct { float lo[3], hi[3]; int32_t left, right; int64_t begin, end; } node_t;

#define LEAF 12

static inline float coord(const pt_t *p, int d) { return d == 0 ? p->x : (d == 1 ? p->y : p->z); }

static inline uint64_t make_key(float d2, uint32_t c) {
    uint32_t b; memcpy(&b, &d2, 4);           /* d2 >= +0: bit pattern is monotone */
    return ((uint64_t)b << 32) | c;
}

static inline float dist2(const pt_t *p, float qx, float qy, float qz) {
    float dx = p->x - qx, dy = p->y - qy, dz = p->z - qz;
    float sx = dx * dx, sy = dy * dy, sz = dz * dz;
    float s = sx + sy;
    return s + sz;
}

/* ---- bounded max-heap of keys ------------------------------------------------------------ */
static inline void heap_push(uint64_t *h, int *n, int k, uint64_t key) {
    if (*n < k) {
        int i = (*n)++;
        h[i] = key;
        while (i > 0) { int p = (i - 1) / 2; if (h[p] >= h[i]) break; uint64_t t = h[p]; h[p] = h[i]; h[i] = t; i = p; }
    } else if (key < h[0]) {
        int i = 0; h[0] = key;
        for (;;) {
            int l = 2 * i + 1, r = l + 1, m = i;
            if (l < k && h[l] > h[m]) m = l;
            if (r < k && h[r] > h[m]) m = r;
            if (m == i) break;
            uint64_t t = h[m]; h[m] = h[i]; h[i] = t; i = m;
        }
    }
}

static int cmp_u64(const void *a, const void *b) {
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : (x > y ? 1 : 0);
}

/* ---- ghost extension ------------------------------------------------------------------- */
static pt_t *make_ext(const float *pos, int64_t n, float box) {
    pt_t *ext = (pt_t *)malloc(sizeof(pt_t) * 27 * (size_t)n);
    if (!ext) return NULL;
    const float v[3] = { -box, 0.0f, box };
    for (int s = 0; s < 27; ++s) {
        float sx = v[s / 9], sy = v[(s / 3) % 3], sz = v[s % 3];
        for (int64_t j = 0; j < n; ++j) {
            pt_t *p = &ext[(size_t)s * n + j];
            p->x = pos[3 * j + 0] + sx;
            p->y = pos[3 * j + 1] + sy;
            p->z = pos[3 * j + 2] + sz;
            p->c = (uint32_t)((size_t)s * n + j);
        }
    }
    return ext;
}

/* ---- KD-tree ---------------------------------------------------------------------------- */
static void select_nth(pt_t *a, int64_t lo, int64_t hi, int64_t nth, int d) {
    /* Hoare quickselect on coordinate d, range [lo,hi) */
    while (hi - lo > 1) {
        float pv = coord(&a[lo + (hi - lo) / 2], d);
        int64_t i = lo, j = hi - 1;
        while (i <= j) {
            while (coord(&a[i], d) < pv) ++i;
            while (coord(&a[j], d) > pv) --j;
            if (i <= j) { pt_t t = a[i]; a[i] = a[j]; a[j] = t; ++i; --j; }
        }
        if (nth <= j) hi = j + 1; else if (nth >= i) lo = i; else return;
    }
}

typedef struct { node_t *nodes; int32_t count, cap; pt_t *pts; } tree_t;

static int32_t build(tree_t *t, int64_t begin, int64_t end) {
    if (t->count == t->cap) return -1;
    int32_t id = t->count++;
    node_t *nd = &t->nodes[id];
    nd->begin = begin; nd->end = end; nd->left = nd->right = -1;
    for (int d = 0; d < 3; ++d) { nd->lo[d] = INFINITY; nd->hi[d] = -INFINITY; }
    for (int64_t i = begin; i < end; ++i)
        for (int d = 0; d < 3; ++d) {
            float v = coord(&t->pts[i], d);
            if (v < nd->lo[d]) nd->lo[d] = v;
            if (v > nd->hi[d]) nd->hi[d] = v;
        }
    if (end - begin <= LEAF) return id;
    int d = 0; float w = nd->hi[0] - nd->lo[0];
    for (int a = 1; a < 3; ++a) if (nd->hi[a] - nd->lo[a] > w) { w = nd->hi[a] - nd->lo[a]; d = a; }
    if (!(w > 0.0f)) return id;               /* all coincident: keep as a (large) leaf */
    int64_t mid = begin + (end - begin) / 2;
    select_nth(t->pts, begin, end, mid, d);
    int32_t l = build(t, begin, mid);
    int32_t r = build(t, mid, end);
    t->nodes[id].left = l; t->nodes[id].right = r;   /* nodes[] is preallocated: no realloc moves */
    return id;
}

static inline double box_lb(const node_t *nd, float qx, float qy, float qz) {
    double s = 0.0, q[3] = { qx, qy, qz };
    for (int d = 0; d < 3; ++d) {
        double e = 0.0;
        if (q[d] < nd->lo[d]) e = (double)nd->lo[d] - q[d];
        else if (q[d] > nd->hi[d]) e = q[d] - (double)nd->hi[d];
        s += e * e;
    }
    return s * (1.0 - 1e-6);                  /* conservative w.r.t. fp32 rounding of d2 */
}

int knn_oracle_kdtree(const float *pos, int64_t n, float box, int k, int64_t *out) {
    if (n <= 0 || k <= 0 || 27 * n < k) return -1;
    tree_t t;
    t.pts = make_ext(pos, n, box);
    if (!t.pts) return -2;
    int64_t m = 27 * n;
    t.cap = (int32_t)(4 * (m / LEAF + 2)); t.count = 0;
    t.nodes = (node_t *)malloc(sizeof(node_t) * (size_t)t.cap);
    if (!t.nodes) { free(t.pts); return -2; }
    if (build(&t, 0, m) < 0) { free(t.nodes); free(t.pts); return -3; }

    uint64_t *heap = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)k);
    int32_t stack[128];
    for (int64_t i = 0; i < n; ++i) {
        float qx = pos[3 * i], qy = pos[3 * i + 1], qz = pos[3 * i + 2];
        int hn = 0, sp = 0;
        stack[sp++] = 0;
        while (sp > 0) {
            const node_t *nd = &t.nodes[stack[--sp]];
            if (hn == k) {
                uint32_t wb = (uint32_t)(heap[0] >> 32); float worst; memcpy(&worst, &wb, 4);
                if (box_lb(nd, qx, qy, qz) > (double)worst) continue;
            }
            if (nd->left < 0) {
                for (int64_t p = nd->begin; p < nd->end; ++p)
                    heap_push(heap, &hn, k, make_key(dist2(&t.pts[p], qx, qy, qz), t.pts[p].c));
            } else {
                const node_t *l = &t.nodes[nd->left], *r = &t.nodes[nd->right];
                double dl = box_lb(l, qx, qy, qz), dr = box_lb(r, qx, qy, qz);
                if (dl <= dr) { stack[sp++] = nd->right; stack[sp++] = nd->left; }   /* near child popped first */
                else          { stack[sp++] = nd->left;  stack[sp++] = nd->right; }
            }
        }
        qsort(heap, (size_t)k, sizeof(uint64_t), cmp_u64);
        for (int r = 0; r < k; ++r) out[i * k + r] = (int64_t)(heap[r] & 0xffffffffu);
    }
    free(heap); free(t.nodes); free(t.pts);
    return 0;
}

int knn_oracle_brute(const float *pos, int64_t n, float box, int k, int64_t *out) {
    if (n <= 0 || k <= 0 || 27 * n < k) return -1;
    pt_t *ext = make_ext(pos, n, box);
    if (!ext) return -2;
    uint64_t *heap = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)k);
    int64_t m = 27 * n;
    for (int64_t i = 0; i < n; ++i) {
        float qx = pos[3 * i], qy = pos[3 * i + 1], qz = pos[3 * i + 2];
        int hn = 0;
        for (int64_t p = 0; p < m; ++p)
            heap_push(heap, &hn, k, make_key(dist2(&ext[p], qx, qy, qz), ext[p].c));
        qsort(heap, (size_t)k, sizeof(uint64_t), cmp_u64);
        for (int r = 0; r < k; ++r) out[i * k + r] = (int64_t)(heap[r] & 0xffffffffu);
    }
    free(heap); free(ext);
    return 0;
}
