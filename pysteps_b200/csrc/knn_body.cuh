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
#define KD_FN_NOINLINE __host__ __device__
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

// ---- scipy's binary heap ------------------------------------------------------------------------
struct Item { double priority; int payload; };

KD_FN void heap_push(Item *h, int &n, Item it) {
    int i = n++;
    h[i] = it;
    while (i > 0 && h[i].priority < h[(i - 1) / 2].priority) {
        const Item tmp = h[(i - 1) / 2];
        h[(i - 1) / 2] = h[i];
        h[i] = tmp;
        i = (i - 1) / 2;
    }
}

KD_FN void heap_remove(Item *h, int &n) {
    h[0] = h[n - 1];
    n--;
    int i = 0, j = 1, k = 2;
    while ((j < n && h[i].priority > h[j].priority) || (k < n && h[i].priority > h[k].priority)) {
        const int l = (k < n && h[j].priority > h[k].priority) ? k : j;
        const Item tmp = h[l];
        h[l] = h[i];
        h[i] = tmp;
        i = l;
        j = 2 * i + 1;
        k = 2 * i + 2;
    }
}

struct NodeInfo { int node; double side[2]; double min_distance; };

// tree.query(x, k): the kmax nearest points in scipy's order (missing: index n).  Scratch per
// query: nb (kmax items), q and pool (nnodes entries each).
KD_FN_NOINLINE void query(const Tree &t, double x0, double x1, int kmax, int *out_idx, Item *nb, Item *q,
                          NodeInfo *pool, double *out_dist = nullptr) {
    const double x[2] = {x0, x1};
    int nbn = 0, qn = 0, pooln = 0;
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
                    if (nbn == kmax) heap_remove(nb, nbn);
                    Item it;
                    it.priority = -d;
                    it.payload = pi;
                    heap_push(nb, nbn, it);
                    if (nbn == kmax) dub = -nb[0].priority;
                }
            }
            if (qn == 0) break;
            cur = pool[q[0].payload];
            heap_remove(q, qn);
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
                pool[pooln] = far;
                Item it;
                it.priority = far.min_distance;
                it.payload = pooln;
                pooln++;
                heap_push(q, qn, it);
            }
        }
    }
    const int found = nbn;
    for (int i = found - 1; i >= 0; i--) {
        out_idx[i] = nb[0].payload;
        if (out_dist) out_dist[i] = sqrt(-nb[0].priority);
        heap_remove(nb, nbn);
    }
    for (int i = found; i < kmax; i++) {
        out_idx[i] = t.n;
        if (out_dist) out_dist[i] = (double)INFINITY;
    }
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
