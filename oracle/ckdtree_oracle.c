/*
 * oracle/ckdtree_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Restatement of scipy.spatial.cKDTree(data) (defaults: leafsize=16, compact_nodes=True,
 * balanced_tree=True) and of tree.query(x, k) (p=2, eps=0) INCLUDING the order in which it returns
 * equidistant neighbours.  pysteps calls it from utils/cleansing.py:219-220 (detect_outliers) and
 * utils/interpolate.py:78-80 (idwinterp2d); corner coordinates are integers, so exact distance
 * ties -- and coincident points -- are common, and which of the tied points cKDTree returns (and in
 * which order) decides outlier tests and the last bits of the interpolated field.  scipy (1.18.1
 * here) is a third-party binary whose source is not under /root/reference; what is restated is
 * its published algorithm, fixed empirically against the binary where versions differ:
 *   build : node = leaf if <= 16 points; bounds recomputed per node; split dimension = largest
 *           spread (first on ties); the median element is found with std::nth_element (libstdc++
 *           introselect, restated below) comparing VALUES ONLY; split = its value, points < split
 *           go left (Hoare-style swap loop); if no point is below the split (median == minimum)
 *           the split becomes nextafter(split, +inf) and the loop is run again;
 *   query : best-first search, far children queued in a binary min-heap on the squared box
 *           distance, neighbours in a max-heap of size k; a point replaces the worst only if its
 *           squared distance is STRICTLY smaller; leaves scan their points in tree order; results
 *           are popped from the heap (binary heap with scipy's own sift rules, restated below).
 * Parity status: PINNED against the scipy binary (tests/test_oracle_ckdtree.py): tree permutation
 * and thousands of tie-heavy queries (indices and distances), bit for bit.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define LEAFSIZE 16

typedef struct {
    int split_dim;      /* -1: leaf */
    double split;
    int64_t less, greater, start, end;
} kd_node;

typedef struct {
    const double *data; /* (n, 2) */
    int64_t n;
    int64_t *idx;
    kd_node *nodes;
    int64_t nnodes, cap;
    double maxes[2], mins[2];
    int error;
} kd_tree;

/* ---- std::nth_element (libstdc++ introselect), comparator: value of `dim` only ---------------- */
#define VAL(t, i, d) ((t)->data[2 * (i) + (d)])
#define LESS(t, a, b, d) (VAL(t, a, d) < VAL(t, b, d))

/* libstdc++ heap primitives on a[first ..), used by the heap_select fallback of introselect */
static void push_heap_(kd_tree *t, int64_t *a, int64_t hole, int64_t top, int64_t value, int d)
{
    int64_t parent = (hole - 1) / 2;
    while (hole > top && LESS(t, a[parent], value, d)) {
        a[hole] = a[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    a[hole] = value;
}

static void adjust_heap_(kd_tree *t, int64_t *a, int64_t hole, int64_t len, int64_t value, int d)
{
    const int64_t top = hole;
    int64_t child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (LESS(t, a[child], a[child - 1], d)) child--;
        a[hole] = a[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        a[hole] = a[child - 1];
        hole = child - 1;
    }
    push_heap_(t, a, hole, top, value, d);
}

/* std::__heap_select(first, middle, last) */
static void heap_select_(kd_tree *t, int64_t first, int64_t middle, int64_t last, int d)
{
    int64_t *a = t->idx + first;
    const int64_t len = middle - first;
    if (len >= 2) {
        int64_t parent = (len - 2) / 2;
        for (;;) {
            adjust_heap_(t, a, parent, len, a[parent], d);
            if (parent == 0) break;
            parent--;
        }
    }
    for (int64_t i = middle; i < last; i++)
        if (LESS(t, t->idx[i], a[0], d)) {
            const int64_t value = t->idx[i];
            t->idx[i] = a[0];
            adjust_heap_(t, a, 0, len, value, d);
        }
}

static void nth_element(kd_tree *t, int64_t first, int64_t nth, int64_t last, int d)
{
    int64_t *a = t->idx;
    if (first == last || nth == last) return;
    int depth = 0;
    for (int64_t m = last - first; m > 1; m >>= 1) depth++;
    depth *= 2;
    while (last - first > 3) {
        if (depth == 0) {
            heap_select_(t, first, nth + 1, last, d);
            { int64_t tmp = a[first]; a[first] = a[nth]; a[nth] = tmp; }
            return;
        }
        depth--;
        /* __unguarded_partition_pivot: median of (first+1, mid, last-1) to first */
        const int64_t mid = first + (last - first) / 2;
        const int64_t ra = first + 1, rb = mid, rc = last - 1;
        int64_t pick;
        if (LESS(t, a[ra], a[rb], d)) {
            if (LESS(t, a[rb], a[rc], d)) pick = rb;
            else if (LESS(t, a[ra], a[rc], d)) pick = rc;
            else pick = ra;
        } else if (LESS(t, a[ra], a[rc], d)) pick = ra;
        else if (LESS(t, a[rb], a[rc], d)) pick = rc;
        else pick = rb;
        { int64_t tmp = a[first]; a[first] = a[pick]; a[pick] = tmp; }
        /* __unguarded_partition(first + 1, last, pivot = first) */
        int64_t f = first + 1, l = last;
        for (;;) {
            while (LESS(t, a[f], a[first], d)) f++;
            l--;
            while (LESS(t, a[first], a[l], d)) l--;
            if (!(f < l)) break;
            { int64_t tmp = a[f]; a[f] = a[l]; a[l] = tmp; }
            f++;
        }
        if (f <= nth) first = f; else last = f;
    }
    /* __insertion_sort(first, last) */
    for (int64_t i = first + 1; i < last; i++) {
        const int64_t v = a[i];
        if (LESS(t, v, a[first], d)) {
            memmove(a + first + 1, a + first, sizeof(int64_t) * (size_t)(i - first));
            a[first] = v;
        } else {
            int64_t j = i;
            while (LESS(t, v, a[j - 1], d)) {
                a[j] = a[j - 1];
                j--;
            }
            a[j] = v;
        }
    }
}

static int64_t partition_below(kd_tree *t, int64_t start, int64_t end, int d, double split)
{
    int64_t *a = t->idx;
    int64_t p = start, q = end - 1;
    while (p <= q) {
        if (VAL(t, a[p], d) < split) p++;
        else if (VAL(t, a[q], d) >= split) q--;
        else {
            int64_t tmp = a[p]; a[p] = a[q]; a[q] = tmp;
            p++;
            q--;
        }
    }
    return p;
}

static int64_t build(kd_tree *t, int64_t start, int64_t end)
{
    if (t->nnodes == t->cap) {
        t->cap = t->cap ? 2 * t->cap : 64;
        t->nodes = (kd_node *)realloc(t->nodes, sizeof(kd_node) * (size_t)t->cap);
    }
    const int64_t me = t->nnodes++;
    t->nodes[me].start = start;
    t->nodes[me].end = end;
    t->nodes[me].split_dim = -1;
    t->nodes[me].less = t->nodes[me].greater = -1;
    if (end - start <= LEAFSIZE) return me;
    double maxes[2], mins[2];
    for (int c = 0; c < 2; c++) maxes[c] = mins[c] = VAL(t, t->idx[start], c);
    for (int64_t j = start + 1; j < end; j++)
        for (int c = 0; c < 2; c++) {
            const double v = VAL(t, t->idx[j], c);
            maxes[c] = maxes[c] > v ? maxes[c] : v;
            mins[c] = mins[c] < v ? mins[c] : v;
        }
    int d = 0;
    double size = 0.0;
    for (int c = 0; c < 2; c++)
        if (maxes[c] - mins[c] > size) {
            d = c;
            size = maxes[c] - mins[c];
        }
    if (maxes[d] == mins[d]) return me; /* all points identical: leaf */
    const int64_t i = (end - start) / 2;
    nth_element(t, start, start + i, end, d);
    double split = VAL(t, t->idx[start + i], d);
    int64_t p = partition_below(t, start, end, d, split);
    if (p == start) {
        /* the median equals the minimum: the split moves just above it */
        split = nextafter(split, INFINITY);
        p = partition_below(t, start, end, d, split);
    }
    const int64_t lo = build(t, start, p);
    const int64_t hi = build(t, p, end);
    t->nodes[me].less = lo;
    t->nodes[me].greater = hi;
    t->nodes[me].split_dim = d;
    t->nodes[me].split = split;
    return me;
}

kd_tree *ora_kd_build(const double *data, int64_t n)
{
    kd_tree *t = (kd_tree *)calloc(1, sizeof(kd_tree));
    t->data = data;
    t->n = n;
    t->idx = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
    for (int64_t i = 0; i < n; i++) t->idx[i] = i;
    for (int c = 0; c < 2; c++) {
        t->maxes[c] = t->mins[c] = n ? data[c] : 0.0;
        for (int64_t i = 1; i < n; i++) {
            const double v = data[2 * i + c];
            if (v > t->maxes[c]) t->maxes[c] = v;
            if (v < t->mins[c]) t->mins[c] = v;
        }
    }
    build(t, 0, n);
    return t;
}

/* std::nth_element(idx, idx + nth, idx + n) comparing values[idx] -- exposed for pinning against
 * libstdc++ itself (tests/test_oracle_ckdtree.py builds a small C++ program); depth_limit < 0 uses
 * the library's own 2*lg(n) */
void ora_kd_nth_element(const double *values, int64_t n, int64_t nth, int depth_limit, int64_t *idx)
{
    kd_tree t;
    memset(&t, 0, sizeof(t));
    double *pairs = (double *)malloc(sizeof(double) * 2 * (size_t)(n > 0 ? n : 1));
    for (int64_t i = 0; i < n; i++) { pairs[2 * i] = values[i]; pairs[2 * i + 1] = 0.0; }
    t.data = pairs;
    t.n = n;
    t.idx = idx;
    (void)depth_limit;
    nth_element(&t, 0, nth, n, 0);
    free(pairs);
}

void ora_kd_free(kd_tree *t)
{
    if (!t) return;
    free(t->idx);
    free(t->nodes);
    free(t);
}

int ora_kd_error(const kd_tree *t) { return t->error; }
void ora_kd_indices(const kd_tree *t, int64_t *out) { memcpy(out, t->idx, sizeof(int64_t) * (size_t)t->n); }

/* ---- scipy's binary heap (ckdtree query) -------------------------------------------------------- */
typedef struct { double priority; int64_t payload; } heap_item;
typedef struct { heap_item *h; int64_t n, cap; } heap;

static void heap_push(heap *q, heap_item it)
{
    if (q->n == q->cap) {
        q->cap = q->cap ? 2 * q->cap + 1 : 16;
        q->h = (heap_item *)realloc(q->h, sizeof(heap_item) * (size_t)q->cap);
    }
    int64_t i = q->n++;
    q->h[i] = it;
    while (i > 0 && q->h[i].priority < q->h[(i - 1) / 2].priority) {
        heap_item tmp = q->h[(i - 1) / 2];
        q->h[(i - 1) / 2] = q->h[i];
        q->h[i] = tmp;
        i = (i - 1) / 2;
    }
}

static void heap_remove(heap *q)
{
    q->h[0] = q->h[q->n - 1];
    q->n--;
    const int64_t nn = q->n;
    int64_t i = 0, j = 1, k = 2;
    while ((j < nn && q->h[i].priority > q->h[j].priority) || (k < nn && q->h[i].priority > q->h[k].priority)) {
        const int64_t l = (k < nn && q->h[j].priority > q->h[k].priority) ? k : j;
        heap_item tmp = q->h[l];
        q->h[l] = q->h[i];
        q->h[i] = tmp;
        i = l;
        j = 2 * i + 1;
        k = 2 * i + 2;
    }
}

typedef struct { int64_t node; double side[2]; double min_distance; } nodeinfo;

/* tree.query(x, k): out_idx / out_dist (k entries; missing neighbours: index n, distance inf) */
void ora_kd_query(const kd_tree *t, const double *x, int64_t kmax, int64_t *out_idx, double *out_dist)
{
    heap q = {0, 0, 0}, nb = {0, 0, 0};
    /* node infos live in a growing pool; the heap stores pool indices */
    int64_t pool_cap = 64, pool_n = 0;
    nodeinfo *pool = (nodeinfo *)malloc(sizeof(nodeinfo) * (size_t)pool_cap);
    nodeinfo cur;
    cur.node = 0;
    cur.min_distance = 0.0;
    for (int c = 0; c < 2; c++) {
        double s = x[c] - t->maxes[c];
        const double s2 = t->mins[c] - x[c];
        if (s2 > s) s = s2;
        if (s < 0.0) s = 0.0;
        cur.side[c] = s * s;
        cur.min_distance += cur.side[c];
    }
    double dub = INFINITY;
    for (;;) {
        const kd_node *node = &t->nodes[cur.node];
        if (node->split_dim == -1) {
            for (int64_t i = node->start; i < node->end; i++) {
                const int64_t pi = t->idx[i];
                const double dx = t->data[2 * pi] - x[0], dy = t->data[2 * pi + 1] - x[1];
                double d = 0.0;
                d += dx * dx;
                d += dy * dy;
                if (d < dub) {
                    if (nb.n == kmax) heap_remove(&nb);
                    heap_item it = {-d, pi};
                    heap_push(&nb, it);
                    if (nb.n == kmax) dub = -nb.h[0].priority;
                }
            }
            if (q.n == 0) break;
            cur = pool[q.h[0].payload];
            heap_remove(&q);
        } else {
            if (cur.min_distance > dub) break;
            const int sd = node->split_dim;
            nodeinfo far = cur;
            double s;
            if (x[sd] < node->split) {
                cur.node = node->less;
                far.node = node->greater;
                s = node->split - x[sd];
            } else {
                cur.node = node->greater;
                far.node = node->less;
                s = x[sd] - node->split;
            }
            s = s * s;
            far.min_distance += s - far.side[sd];
            far.side[sd] = s;
            if (cur.min_distance > far.min_distance) {
                nodeinfo tmp = cur;
                cur = far;
                far = tmp;
            }
            if (far.min_distance <= dub) {
                if (pool_n == pool_cap) {
                    pool_cap *= 2;
                    pool = (nodeinfo *)realloc(pool, sizeof(nodeinfo) * (size_t)pool_cap);
                }
                pool[pool_n] = far;
                heap_item it = {far.min_distance, pool_n};
                pool_n++;
                heap_push(&q, it);
            }
        }
    }
    const int64_t found = nb.n;
    for (int64_t i = found - 1; i >= 0; i--) {
        out_idx[i] = nb.h[0].payload;
        out_dist[i] = sqrt(-nb.h[0].priority);
        heap_remove(&nb);
    }
    for (int64_t i = found; i < kmax; i++) {
        out_idx[i] = t->n;
        out_dist[i] = INFINITY;
    }
    free(q.h);
    free(nb.h);
    free(pool);
}

/* many queries: xs (nq, 2) -> idx (nq, k), dist (nq, k) */
void ora_kd_query_many(const kd_tree *t, const double *xs, int64_t nq, int64_t k, int64_t *idx, double *dist)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nq; i++) ora_kd_query(t, xs + 2 * i, k, idx + i * k, dist + i * k);
}
