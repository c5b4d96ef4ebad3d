// knn_body.cuh -- scipy.spatial.cKDTree (build + k-NN query, 2-D, p=2) with its exact tie order,
// and the Mahalanobis outlier test of pysteps/utils/cleansing.py:216-245 on the neighbour lists it
// returns.  Host/device source (see spline_body.cuh): tests/host_kernels/ runs it on the CPU
// against the scipy binary and the oracle (oracle/ckdtree_oracle.c, whose header states the
// algorithm).  Why: corner coordinates are integers, so equidistant and coincident vectors are
// common, and WHICH of the tied vectors cKDTree returns decides outlier tests
// (DESIGN.md section 4).
#pragma once
#include <math.h>
#include <stddef.h>

#if defined(__CUDACC__)
#define KD_FN __host__ __device__ __forceinline__
#define KD_FN_NOINLINE inline __host__ __device__ __noinline__
#else
#define KD_FN inline
#define KD_FN_NOINLINE inline
#endif

namespace kd {

constexpr int LEAFSIZE = 16;

struct Node {
    int split_dim;  // -1: leaf
    int less, greater, start, end;
    double split;
};

struct Tree {
    const double *data;  // (n, 2)
    int n;
    int *idx;            // tree order of the points (scipy: tree.indices)
    Node *nodes;         // capacity >= max_nodes(n)
    int nnodes;
    double maxes[2], mins[2];
};

KD_FN int max_nodes(int n) { return 2 * n + 1; }

#define KD_VAL(t, i, d) ((t).data[2 * (size_t)(i) + (d)])
#define KD_LESS(t, a, b, d) (KD_VAL(t, a, d) < KD_VAL(t, b, d))

// ---- libstdc++ std::nth_element (introselect), comparator: value of dimension d only ------------
KD_FN void push_heap_(const Tree &t, int *a, int hole, int top, int value, int d) {
    int parent = (hole - 1) / 2;
    while (hole > top && KD_LESS(t, a[parent], value, d)) {
        a[hole] = a[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    a[hole] = value;
}

KD_FN void adjust_heap_(const Tree &t, int *a, int hole, int len, int value, int d) {
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (KD_LESS(t, a[child], a[child - 1], d)) child--;
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

KD_FN void heap_select_(const Tree &t, int first, int middle, int last, int d) {
    int *a = t.idx + first;
    const int len = middle - first;
    if (len >= 2) {
        int parent = (len - 2) / 2;
        for (;;) {
            adjust_heap_(t, a, parent, len, a[parent], d);
            if (parent == 0) break;
            parent--;
        }
    }
    for (int i = middle; i < last; i++)
        if (KD_LESS(t, t.idx[i], a[0], d)) {
            const int value = t.idx[i];
            t.idx[i] = a[0];
            adjust_heap_(t, a, 0, len, value, d);
        }
}

KD_FN_NOINLINE void nth_element(const Tree &t, int first, int nth, int last, int d) {
    int *a = t.idx;
    if (first == last || nth == last) return;
    int depth = 0;
    for (int m = last - first; m > 1; m >>= 1) depth++;
    depth *= 2;
    while (last - first > 3) {
        if (depth == 0) {
            heap_select_(t, first, nth + 1, last, d);
            const int tmp = a[first]; a[first] = a[nth]; a[nth] = tmp;
            return;
        }
        depth--;
        const int mid = first + (last - first) / 2;
        const int ra = first + 1, rb = mid, rc = last - 1;
        int pick;
        if (KD_LESS(t, a[ra], a[rb], d)) {
            if (KD_LESS(t, a[rb], a[rc], d)) pick = rb;
            else if (KD_LESS(t, a[ra], a[rc], d)) pick = rc;
            else pick = ra;
        } else if (KD_LESS(t, a[ra], a[rc], d)) pick = ra;
        else if (KD_LESS(t, a[rb], a[rc], d)) pick = rc;
        else pick = rb;
        { const int tmp = a[first]; a[first] = a[pick]; a[pick] = tmp; }
        int f = first + 1, l = last;
        for (;;) {
            while (KD_LESS(t, a[f], a[first], d)) f++;
            l--;
            while (KD_LESS(t, a[first], a[l], d)) l--;
            if (!(f < l)) break;
            { const int tmp = a[f]; a[f] = a[l]; a[l] = tmp; }
            f++;
        }
        if (f <= nth) first = f; else last = f;
    }
    for (int i = first + 1; i < last; i++) {  // __insertion_sort
        const int v = a[i];
        if (KD_LESS(t, v, a[first], d)) {
            for (int j = i; j > first; j--) a[j] = a[j - 1];
            a[first] = v;
        } else {
            int j = i;
            while (KD_LESS(t, v, a[j - 1], d)) {
                a[j] = a[j - 1];
                j--;
            }
            a[j] = v;
        }
    }
}

KD_FN int partition_below(const Tree &t, int start, int end, int d, double split) {
    int *a = t.idx;
    int p = start, q = end - 1;
    while (p <= q) {
        if (KD_VAL(t, a[p], d) < split) p++;
        else if (KD_VAL(t, a[q], d) >= split) q--;
        else {
            const int tmp = a[p]; a[p] = a[q]; a[q] = tmp;
            p++;
            q--;
        }
    }
    return p;
}

// The recursive build of scipy, with an explicit stack (children work on disjoint index ranges,
// so their order does not matter).  `stack` holds 2 ints per pending node, capacity >= 64 pairs.
KD_FN_NOINLINE void build(Tree &t, int *stack) {
    const int n = t.n;
    for (int i = 0; i < n; i++) t.idx[i] = i;
    for (int c = 0; c < 2; c++) {
        t.maxes[c] = t.mins[c] = n ? t.data[c] : 0.0;
        for (int i = 1; i < n; i++) {
            const double v = t.data[2 * (size_t)i + c];
            if (v > t.maxes[c]) t.maxes[c] = v;
            if (v < t.mins[c]) t.mins[c] = v;
        }
    }
    t.nnodes = 1;
    t.nodes[0].start = 0;
    t.nodes[0].end = n;
    int sp = 0;
    stack[sp++] = 0;
    while (sp > 0) {
        const int me = stack[--sp];
        Node &nd = t.nodes[me];
        const int start = nd.start, end = nd.end;
        nd.split_dim = -1;
        nd.less = nd.greater = -1;
        nd.split = 0.0;
        if (end - start <= LEAFSIZE) continue;
        double maxes[2], mins[2];
        for (int c = 0; c < 2; c++) maxes[c] = mins[c] = KD_VAL(t, t.idx[start], c);
        for (int j = start + 1; j < end; j++)
            for (int c = 0; c < 2; c++) {
                const double v = KD_VAL(t, t.idx[j], c);
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
        if (maxes[d] == mins[d]) continue;  // all points identical: leaf
        const int i = (end - start) / 2;
        nth_element(t, start, start + i, end, d);
        double split = KD_VAL(t, t.idx[start + i], d);
        int p = partition_below(t, start, end, d, split);
        if (p == start) {  // the median equals the minimum: the split moves just above it
            split = nextafter(split, (double)INFINITY);
            p = partition_below(t, start, end, d, split);
        }
        const int lo = t.nnodes++, hi = t.nnodes++;
        t.nodes[lo].start = start; t.nodes[lo].end = p;
        t.nodes[hi].start = p; t.nodes[hi].end = end;
        Node &self = t.nodes[me];
        self.less = lo;
        self.greater = hi;
        self.split_dim = d;
        self.split = split;
        stack[sp++] = hi;
        stack[sp++] = lo;
    }
}

// ---- the same build on position-aligned (x, y, index) triples -------------------------------------
// The parallel device build (knn.cu) keeps the coordinates of the point at tree position p in
// kx[p], ky[p] beside idx[p] and moves the three together, so a comparison is one load.  Its
// partition steps are the libstdc++ loops restated as PAIR SWAPS: with A the positions (ascending)
// where the left scan stops and B the positions (descending) where the right scan stops, the
// sequential loop swaps (A_i, B_i) for every i below the first i with A_i >= B_i, nothing else.
// The functions below are that formulation executed serially (host tests pin it against the
// sequential restatement above and, through it, against libstdc++ and scipy); knn.cu executes the
// identical formulation with one warp per tree node.
struct Tri { double *kx, *ky; int *idx; };
struct Elem { double x, y; int i; };

KD_FN Elem tri_load(const Tri &a, int p) { Elem e; e.x = a.kx[p]; e.y = a.ky[p]; e.i = a.idx[p]; return e; }
KD_FN void tri_store(const Tri &a, int p, const Elem &e) { a.kx[p] = e.x; a.ky[p] = e.y; a.idx[p] = e.i; }
KD_FN void tri_swap(const Tri &a, int p, int q) {
    const Elem e = tri_load(a, p);
    tri_store(a, p, tri_load(a, q));
    tri_store(a, q, e);
}
KD_FN double ekey(const Elem &e, int d) { return d ? e.y : e.x; }
KD_FN const double *tri_keys(const Tri &a, int d) { return d ? a.ky : a.kx; }

KD_FN void tri_push_heap(const Tri &a, int base, int hole, int top, const Elem &value, int d) {
    int parent = (hole - 1) / 2;
    while (hole > top && ekey(tri_load(a, base + parent), d) < ekey(value, d)) {
        tri_store(a, base + hole, tri_load(a, base + parent));
        hole = parent;
        parent = (hole - 1) / 2;
    }
    tri_store(a, base + hole, value);
}

KD_FN void tri_adjust_heap(const Tri &a, int base, int hole, int len, const Elem &value, int d) {
    const int top = hole;
    int child = hole;
    const double *K = tri_keys(a, d);
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (K[base + child] < K[base + child - 1]) child--;
        tri_store(a, base + hole, tri_load(a, base + child));
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        tri_store(a, base + hole, tri_load(a, base + child - 1));
        hole = child - 1;
    }
    tri_push_heap(a, base, hole, top, value, d);
}

// std::__heap_select(first, middle, last) followed by iter_swap(first, nth) is the caller's job
KD_FN_NOINLINE void tri_heap_select(const Tri &a, int first, int middle, int last, int d) {
    const int len = middle - first;
    const double *K = tri_keys(a, d);
    if (len >= 2) {
        int parent = (len - 2) / 2;
        for (;;) {
            tri_adjust_heap(a, first, parent, len, tri_load(a, first + parent), d);
            if (parent == 0) break;
            parent--;
        }
    }
    for (int i = middle; i < last; i++)
        if (K[i] < K[first]) {
            const Elem value = tri_load(a, i);
            tri_store(a, i, tri_load(a, first));
            tri_adjust_heap(a, first, 0, len, value, d);
        }
}

// std::__insertion_sort(first, last)
KD_FN void tri_insertion_sort(const Tri &a, int first, int last, int d) {
    for (int i = first + 1; i < last; i++) {
        const Elem v = tri_load(a, i);
        if (ekey(v, d) < tri_keys(a, d)[first]) {
            for (int j = i; j > first; j--) tri_store(a, j, tri_load(a, j - 1));
            tri_store(a, first, v);
        } else {
            int j = i;
            while (ekey(v, d) < tri_keys(a, d)[j - 1]) {
                tri_store(a, j, tri_load(a, j - 1));
                j--;
            }
            tri_store(a, j, v);
        }
    }
}

// std::__move_median_to_first(first, first + 1, mid, last - 1): position of the median
KD_FN int tri_median_pick(const double *K, int first, int mid, int last) {
    const int ra = first + 1, rb = mid, rc = last - 1;
    if (K[ra] < K[rb]) {
        if (K[rb] < K[rc]) return rb;
        if (K[ra] < K[rc]) return rc;
        return ra;
    }
    if (K[ra] < K[rc]) return ra;
    if (K[rb] < K[rc]) return rc;
    return rb;
}

// MODE 0: std::__unguarded_partition(lo, hi, pivot) -- left scan stops where !(v < piv), right scan
//         where !(piv < v); returns the cut.
// MODE 1: partition_below(lo, hi, split) -- left stops where !(v < split), right where !(v >= split);
//         returns the first position of the ">= split" block.
// posA / posB: scratch of hi - lo entries each.
template <int MODE, class P>
KD_FN int pair_partition_serial(const Tri &a, int d, int lo, int hi, double piv, P *posA, P *posB) {
    const double *K = tri_keys(a, d);
    int cntA = 0, cntB = 0;
    for (int p = lo; p < hi; p++)
        if (!(K[p] < piv)) posA[cntA++] = (P)p;
    for (int p = hi - 1; p >= lo; p--)
        if (MODE == 0 ? !(piv < K[p]) : !(K[p] >= piv)) posB[cntB++] = (P)p;
    const int npair = cntA < cntB ? cntA : cntB;
    int j = 0;
    while (j < npair && posA[j] < posB[j]) {
        tri_swap(a, posA[j], posB[j]);
        j++;
    }
    if (MODE == 1) return hi - cntA;
    const int fa = j < cntA ? (int)posA[j] : hi, fb = j > 0 ? (int)posB[j - 1] : hi;
    return fa < fb ? fa : fb;
}

// the tree build on triples, serial execution of the pair formulation; depth_limit < 0: libstdc++'s
// 2 * floor(log2(len)).  Node numbering is breadth-first here (sequential build: depth-first); a
// query never looks at node numbers, only at the structure.
template <class P>
KD_FN_NOINLINE void build_pairs(Tree &t, const Tri &a, P *posA, P *posB, int *queue, int depth_limit) {
    const int n = t.n;
    for (int i = 0; i < n; i++) { a.kx[i] = t.data[2 * (size_t)i]; a.ky[i] = t.data[2 * (size_t)i + 1]; a.idx[i] = i; }
    for (int c = 0; c < 2; c++) {
        t.maxes[c] = t.mins[c] = n ? t.data[c] : 0.0;
        for (int i = 1; i < n; i++) {
            const double v = t.data[2 * (size_t)i + c];
            if (v > t.maxes[c]) t.maxes[c] = v;
            if (v < t.mins[c]) t.mins[c] = v;
        }
    }
    t.nnodes = 1;
    t.nodes[0].start = 0; t.nodes[0].end = n;
    t.nodes[0].split_dim = -1; t.nodes[0].less = t.nodes[0].greater = -1; t.nodes[0].split = 0.0;
    int qh = 0, qt = 0;
    if (n > LEAFSIZE) queue[qt++] = 0;
    while (qh < qt) {
        const int me = queue[qh++];
        const int start = t.nodes[me].start, end = t.nodes[me].end;
        double mx[2], mn[2];
        mx[0] = mn[0] = a.kx[start]; mx[1] = mn[1] = a.ky[start];
        for (int j = start + 1; j < end; j++) {
            mx[0] = mx[0] > a.kx[j] ? mx[0] : a.kx[j]; mn[0] = mn[0] < a.kx[j] ? mn[0] : a.kx[j];
            mx[1] = mx[1] > a.ky[j] ? mx[1] : a.ky[j]; mn[1] = mn[1] < a.ky[j] ? mn[1] : a.ky[j];
        }
        int d = 0;
        double size = 0.0;
        for (int c = 0; c < 2; c++)
            if (mx[c] - mn[c] > size) { d = c; size = mx[c] - mn[c]; }
        if (mx[d] == mn[d]) continue;
        const double *K = tri_keys(a, d);
        const int nth = start + (end - start) / 2;
        {   // std::nth_element(start, nth, end)
            int first = start, last = end, depth = depth_limit;
            if (depth < 0) { depth = 0; for (int m = last - first; m > 1; m >>= 1) depth++; depth *= 2; }
            bool done = false;
            while (last - first > 3) {
                if (depth == 0) {
                    tri_heap_select(a, first, nth + 1, last, d);
                    tri_swap(a, first, nth);
                    done = true;
                    break;
                }
                depth--;
                const int mid = first + (last - first) / 2;
                tri_swap(a, first, tri_median_pick(K, first, mid, last));
                const int cut = pair_partition_serial<0>(a, d, first + 1, last, K[first], posA, posB);
                if (cut <= nth) first = cut; else last = cut;
            }
            if (!done) tri_insertion_sort(a, first, last, d);
        }
        double split = K[nth];
        int p = pair_partition_serial<1>(a, d, start, end, split, posA, posB);
        if (p == start) {
            split = nextafter(split, (double)INFINITY);
            p = pair_partition_serial<1>(a, d, start, end, split, posA, posB);
        }
        const int lo = t.nnodes++, hi = t.nnodes++;
        for (int c = 0; c < 2; c++) {
            Node &ch = t.nodes[c ? hi : lo];
            ch.start = c ? p : start; ch.end = c ? end : p;
            ch.split_dim = -1; ch.less = ch.greater = -1; ch.split = 0.0;
            if (ch.end - ch.start > LEAFSIZE) queue[qt++] = c ? hi : lo;
        }
        t.nodes[me].less = lo; t.nodes[me].greater = hi; t.nodes[me].split_dim = d; t.nodes[me].split = split;
    }
    for (int i = 0; i < n; i++) t.idx[i] = a.idx[i];
}

// ---- scipy's binary heap ------------------------------------------------------------------------
struct Item { double priority; int payload; };
struct NodeInfo { int node; double side[2]; double min_distance; };

// scipy keeps {priority, pointer to a nodeinfo} in its node queue with priority == min_distance; here
// the nodeinfo itself is the heap element (same comparisons, same sift order, no side pool)
KD_FN double prio(const Item &a) { return a.priority; }
KD_FN double prio(const NodeInfo &a) { return a.min_distance; }

// heap storage handle: anything indexable (a plain pointer on the host; on the device per-thread
// arrays interleaved across the threads of a CTA in shared memory)
template <class T>
struct Strided {
    T *p;
    int stride;
    KD_FN T &operator[](int i) const { return p[(size_t)i * stride]; }
};

template <class H, class T>
KD_FN void heap_push(H h, int &n, const T &it) {
    int i = n++;
    h[i] = it;
    while (i > 0 && prio(h[i]) < prio(h[(i - 1) / 2])) {
        const T tmp = h[(i - 1) / 2];
        h[(i - 1) / 2] = h[i];
        h[i] = tmp;
        i = (i - 1) / 2;
    }
}

template <class T, class H>
KD_FN void heap_remove(H h, int &n) {
    h[0] = h[n - 1];
    n--;
    int i = 0, j = 1, k = 2;
    while ((j < n && prio(h[i]) > prio(h[j])) || (k < n && prio(h[i]) > prio(h[k]))) {
        const int l = (k < n && prio(h[j]) > prio(h[k])) ? k : j;
        const T tmp = h[l];
        h[l] = h[i];
        h[i] = tmp;
        i = l;
        j = 2 * i + 1;
        k = 2 * i + 2;
    }
}

// no spill space: a full pending-node heap is an error
struct NoGrow {
    template <class Q>
    KD_FN bool operator()(Q &, int &, int) const { return false; }
};

// tree.query(x, k): the kmax nearest points in scipy's order (missing: index n).  Scratch per
// query: nb (kmax items) and q (qcap pending nodes).  A query queues at most one far child per
// internal node it visits, so qcap = number of nodes can never overflow; with a smaller heap,
// grow(q, qcap, qn) is asked for a larger one when it is full (it moves the qn entries), and the
// return value is false (result invalid) if it could not provide one.
template <class NB, class Q, class Grow>
KD_FN_NOINLINE bool query(const Tree &t, double x0, double x1, int kmax, int *out_idx, NB nb, Q q,
                          int qcap, const Grow &grow, double *out_dist = nullptr) {
    const double x[2] = {x0, x1};
    int nbn = 0, qn = 0;
    bool ok = true;
    NodeInfo cur;
    cur.node = 0;
    cur.min_distance = 0.0;
    for (int c = 0; c < 2; c++) {
        double s = x[c] - t.maxes[c];
        const double s2 = t.mins[c] - x[c];
        if (s2 > s) s = s2;
        if (s < 0.0) s = 0.0;
        cur.side[c] = s * s;
        cur.min_distance += cur.side[c];
    }
    double dub = (double)INFINITY;
    for (;;) {
        const Node &node = t.nodes[cur.node];
        if (node.split_dim == -1) {
            for (int i = node.start; i < node.end; i++) {
                const int pi = t.idx[i];
                const double dx = t.data[2 * (size_t)pi] - x[0], dy = t.data[2 * (size_t)pi + 1] - x[1];
                double d = 0.0;
                d += dx * dx;
                d += dy * dy;
                if (d < dub) {
                    if (nbn == kmax) heap_remove<Item>(nb, nbn);
                    Item it;
                    it.priority = -d;
                    it.payload = pi;
                    heap_push(nb, nbn, it);
                    if (nbn == kmax) dub = -nb[0].priority;
                }
            }
            if (qn == 0) break;
            cur = q[0];
            heap_remove<NodeInfo>(q, qn);
        } else {
            if (cur.min_distance > dub) break;
            const int sd = node.split_dim;
            NodeInfo far = cur;
            double s;
            if (x[sd] < node.split) {
                cur.node = node.less;
                far.node = node.greater;
                s = node.split - x[sd];
            } else {
                cur.node = node.greater;
                far.node = node.less;
                s = x[sd] - node.split;
            }
            s = s * s;
            far.min_distance += s - far.side[sd];
            far.side[sd] = s;
            if (cur.min_distance > far.min_distance) {
                const NodeInfo tmp = cur;
                cur = far;
                far = tmp;
            }
            if (far.min_distance <= dub) {
                if (qn == qcap && !grow(q, qcap, qn)) ok = false;
                else heap_push(q, qn, far);
            }
        }
    }
    const int found = nbn;
    for (int i = found - 1; i >= 0; i--) {
        const Item top = nb[0];
        out_idx[i] = top.payload;
        if (out_dist) out_dist[i] = sqrt(-top.priority);
        heap_remove<Item>(nb, nbn);
    }
    for (int i = found; i < kmax; i++) {
        out_idx[i] = t.n;
        if (out_dist) out_dist[i] = (double)INFINITY;
    }
    return ok;
}

// numpy's pairwise summation of n <= 128 contiguous doubles (DOUBLE_pairwise_sum): 8 running
// sums for n >= 8, combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), then the tail
KD_FN double numpy_sum(const double *a, int n) {
    if (n < 8) {
        double res = 0.0;  // numpy starts from the first element: res = a[0]; equal to 0.0 + a[0]
        for (int i = 0; i < n; i++) res = (i == 0) ? a[0] : res + a[i];
        return res;
    }
    double r[8];
    for (int j = 0; j < 8; j++) r[j] = a[j];
    int i;
    for (i = 8; i < n - (n % 8); i += 8)
        for (int j = 0; j < 8; j++) r[j] += a[i + j];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; i++) res += a[i];
    return res;
}

// pysteps/utils/interpolate.py:78-107 for one grid point: the k nearest vectors in cKDTree's order,
// w = 1 / (d / mean_res + offset)^power normalised by numpy's sum, values accumulated in
// neighbour order.  w holds the distances on entry (k <= 128).
KD_FN void idw_point(const double *vals, int nvar, const int *inds, double *w, int k, double power, double offset,
                     double mean_res, double *out, size_t out_stride) {
    for (int j = 0; j < k; j++) {
        double d = w[j] / mean_res;
        d = d + offset;
        w[j] = 1.0 / pow(d, power);
    }
    const double wsum = numpy_sum(w, k);
    for (int j = 0; j < k; j++) w[j] = w[j] / wsum;
    for (int c = 0; c < nvar; c++) {
        double acc = 0.0;
        for (int j = 0; j < k; j++) {
            const double term = vals[(size_t)inds[j] * nvar + c] * w[j];
            acc = (j == 0) ? term : acc + term;
        }
        out[(size_t)c * out_stride] = acc;
    }
}

// cleansing.py:231-245 for vector i with its k+1 nearest (inds[0] is dropped as "the vector
// itself", :233): local Mahalanobis distance > thr.  Same formulas, in a fixed order, as the
// lower-index-ties kernel of sparse.cu (np.mean row by row, np.cov with ddof=1, np.linalg.inv as
// a pivoted LU); m = number of neighbours used.
KD_FN bool mahalanobis_outlier(const double *uv, int i, const int *inds, int m, double thr) {
    double mu = 0.0, mv = 0.0;
    for (int q = 1; q <= m; q++) {
        const int j = inds[q];
        mu = (q == 1) ? uv[2 * (size_t)j] : mu + uv[2 * (size_t)j];
        mv = (q == 1) ? uv[2 * (size_t)j + 1] : mv + uv[2 * (size_t)j + 1];
    }
    mu = mu / (double)m;
    mv = mv / (double)m;
    const double zu = uv[2 * (size_t)i] - mu, zv = uv[2 * (size_t)i + 1] - mv;
    double au = 0.0, av = 0.0;
    for (int q = 1; q <= m; q++) {
        const int j = inds[q];
        au = au + (uv[2 * (size_t)j] - mu);
        av = av + (uv[2 * (size_t)j + 1] - mv);
    }
    au = au / (double)m;
    av = av / (double)m;
    double suu = 0.0, suv = 0.0, svv = 0.0;
    for (int q = 1; q <= m; q++) {
        const int j = inds[q];
        const double a = (uv[2 * (size_t)j] - mu) - au;
        const double b = (uv[2 * (size_t)j + 1] - mv) - av;
        suu = suu + a * a;
        suv = suv + a * b;
        svv = svv + b * b;
    }
    const double fact = 1.0 / (double)(m - 1);
    const double a = suu * fact, b = suv * fact, d = svv * fact;
    double MD = 0.0;
    const bool swap = fabs(b) > fabs(a);
    const double p0 = swap ? b : a, p1 = swap ? d : b;
    const double q0 = swap ? a : b, q1 = swap ? b : d;
    if (p0 != 0.0 && !(p0 != p0)) {
        const double l = q0 * (1.0 / p0);
        const double u22 = q1 - l * p1;
        if (u22 != 0.0) {
            const double r00 = swap ? 0.0 : 1.0, r10 = swap ? 1.0 : 0.0;
            const double r01 = swap ? 1.0 : 0.0, r11 = swap ? 0.0 : 1.0;
            const double y10 = r10 - l * r00, y11 = r11 - l * r01;
            const double x10 = y10 / u22, x11 = y11 / u22;
            const double x00 = (r00 - p1 * x10) / p0;
            const double x01 = (r01 - p1 * x11) / p0;
            const double t0 = zu * x00 + zv * x10;
            const double t1 = zu * x01 + zv * x11;
            MD = sqrt(t0 * zu + t1 * zv);
        }
    }
    return MD > thr;
}

}  // namespace kd
