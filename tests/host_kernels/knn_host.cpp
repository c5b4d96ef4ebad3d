// TEST INFRASTRUCTURE: pysteps_b200/csrc/knn_body.cuh compiled as host C++ (see spline_host.cpp).
#include <stdint.h>

#include <vector>

#include "../../pysteps_b200/csrc/knn_body.cuh"
#include "../../pysteps_b200/csrc/quantise_body.cuh"

extern "C" {

// tree permutation (tree.indices) and the k nearest of every query, in cKDTree's order
void host_kd_knn(const double *data, int n, const double *queries, int nq, int k, int *perm, int *out_idx) {
    std::vector<int> idx(n > 0 ? n : 1), stack(256);
    std::vector<kd::Node> nodes(kd::max_nodes(n));
    kd::Tree t;
    t.data = data;
    t.n = n;
    t.idx = idx.data();
    t.nodes = nodes.data();
    kd::build(t, stack.data());
    for (int i = 0; i < n; i++) perm[i] = idx[i];
    std::vector<kd::Item> nb(k);
    std::vector<kd::NodeInfo> q(t.nnodes + 1);
    for (int i = 0; i < nq; i++)
        kd::query(t, queries[2 * (size_t)i], queries[2 * (size_t)i + 1], k, out_idx + (size_t)i * k, nb.data(), q.data(),
                  t.nnodes + 1, kd::NoGrow());
}

// the same through the pair-swap formulation of the build that knn.cu runs with one warp per node
// (knn_body.cuh: build_pairs); depth_limit < 0: libstdc++'s own, 0..: forces the heap-select branch.
// Returns the number of queries whose pending-node heap would have overflowed `qcap`.
int host_kd_knn_pairs(const double *data, int n, const double *queries, int nq, int k, int depth_limit, int qcap,
                      int *perm, int *out_idx) {
    std::vector<int> idx(n > 0 ? n : 1), tidx(n > 0 ? n : 1), queue(n + 1), posA(n + 1), posB(n + 1);
    std::vector<double> kx(n > 0 ? n : 1), ky(n > 0 ? n : 1);
    std::vector<kd::Node> nodes(kd::max_nodes(n));
    kd::Tree t;
    t.data = data;
    t.n = n;
    t.idx = idx.data();
    t.nodes = nodes.data();
    kd::Tri a;
    a.kx = kx.data(); a.ky = ky.data(); a.idx = tidx.data();
    kd::build_pairs(t, a, posA.data(), posB.data(), queue.data(), depth_limit);
    for (int i = 0; i < n; i++) perm[i] = idx[i];
    if (qcap <= 0) qcap = t.nnodes + 1;
    std::vector<kd::Item> nb(k);
    std::vector<kd::NodeInfo> q(qcap);
    int overflow = 0;
    for (int i = 0; i < nq; i++)
        overflow += !kd::query(t, queries[2 * (size_t)i], queries[2 * (size_t)i + 1], k, out_idx + (size_t)i * k,
                               nb.data(), q.data(), qcap, kd::NoGrow());
    return overflow;
}

// == b200_detect_outliers_ckdtree
void host_detect_outliers_ckdtree(const double *uv, const double *xy, int n, double thr, int k, uint8_t *out) {
    if (n < 2) {
        for (int i = 0; i < n; i++) out[i] = 0;
        return;
    }
    const int kk = k + 1 < n ? k + 1 : n;  // cleansing.py:197
    std::vector<int> perm(n), inds((size_t)n * kk);
    host_kd_knn(xy, n, xy, n, kk, perm.data(), inds.data());
    for (int i = 0; i < n; i++) out[i] = kd::mahalanobis_outlier(uv, i, inds.data() + (size_t)i * kk, kk - 1, thr) ? 1 : 0;
}

// == b200_idw_fill_ckdtree: out (nvar, ny, nx)
void host_idw_fill_ckdtree(const double *xy, const double *vals, int npts, int nvar, int k, double power,
                           double offset, double mean_res, const double *xgrid, int nx, const double *ygrid, int ny,
                           double *out) {
    std::vector<int> idx(npts > 0 ? npts : 1), stack(256);
    std::vector<kd::Node> nodes(kd::max_nodes(npts));
    kd::Tree t;
    t.data = xy;
    t.n = npts;
    t.idx = idx.data();
    t.nodes = nodes.data();
    kd::build(t, stack.data());
    std::vector<kd::Item> nb(k);
    std::vector<kd::NodeInfo> q(t.nnodes + 1);
    std::vector<int> inds(k);
    std::vector<double> w(k);
    const size_t N = (size_t)ny * nx;
    for (int i = ny - 1; i >= 0; i--)
        for (int j = 0; j < nx; j++) {
            kd::query(t, xgrid[j], ygrid[i], k, inds.data(), nb.data(), q.data(), t.nnodes + 1, kd::NoGrow(), w.data());
            kd::idw_point(vals, nvar, inds.data(), w.data(), k, power, offset, mean_res, out + (size_t)i * nx + j, N);
        }
}

// float32 scaling of the quantise kernel, element by element
void host_scale_f32(const double *v, int64_t count, double im_min, double im_max, double *out) {
    for (int64_t i = 0; i < count; i++) out[i] = qz::scale_f32(v[i], im_min, im_max);
}

// float64 scaling to uint8 of the fused front end: the division-free path and the division itself
void host_scale_f64(const double *v, int64_t count, double im_min, double im_max, unsigned char *out_fast,
                    unsigned char *out_exact, int64_t *n_fallback) {
    qz::ScaleF64 sc;
    sc.init(im_min, im_max);
    int64_t nf = 0;
    for (int64_t i = 0; i < count; i++) {
        out_fast[i] = sc(v[i]);
        out_exact[i] = sc.exact(v[i]);
        // (how often the fast path defers to the division: reported, not asserted)
        const double num = v[i] - im_min, qa = num * sc.r255;
        const double t = qa + 6755399441055744.0, d = qa - (t - 6755399441055744.0);
        if (!(num == 0.0) && !(sc.wide && qa > 0.0 && qa < 255.5 && (d > 1e-9 || d < -1e-9))) nf++;
    }
    *n_fallback = nf;
}

}  // extern "C"
