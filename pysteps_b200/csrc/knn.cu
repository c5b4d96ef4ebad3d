// knn.cu -- the two k-NN stages of dense_lucaskanade with scipy.spatial.cKDTree's exact neighbour
// order: the local outlier test (pysteps/utils/cleansing.py:216-245) and the inverse-distance grid
// fill (pysteps/utils/interpolate.py:67-114).
//
// The default kernel (sparse.cu) takes equidistant neighbours by lower index; cKDTree returns them
// in an order that follows from its tree (knn_body.cuh restates tree and query bit for bit), and
// with integer corner coordinates that order decides outlier tests.  Here
// the tree is built ON THE DEVICE by one thread (<= 2000 vectors: ~50 k dependent steps, it is
// sequential by definition of nth_element) and every vector then runs scipy's best-first query in
// its own thread.
#include "common.cuh"
#include "knn_body.cuh"

namespace {

struct KnnParams {
    const double *xy, *uv;
    const int *n_dev;
    int n_cap, k;
    double thr;
    int *idx;            // n_cap
    kd::Node *nodes;     // max_nodes(n_cap)
    int *meta;           // [0] = nnodes
    double *bounds;      // maxes[2], mins[2]
    int *inds;           // n_cap * (k+1)
    kd::Item *nb;        // n_cap * (k+1)
    kd::Item *q;         // n_cap * qcap
    kd::NodeInfo *pool;  // n_cap * qcap
    int qcap;
    uint8_t *out;
};

__global__ void kd_build_kernel(const __grid_constant__ KnnParams p) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    const int n = p.n_dev ? min(*p.n_dev, p.n_cap) : p.n_cap;
    int stack[128];
    kd::Tree t;
    t.data = p.xy;
    t.n = n;
    t.idx = p.idx;
    t.nodes = p.nodes;
    kd::build(t, stack);
    p.meta[0] = t.nnodes;
    for (int c = 0; c < 2; c++) {
        p.bounds[c] = t.maxes[c];
        p.bounds[2 + c] = t.mins[c];
    }
}

__global__ void __launch_bounds__(64)
outliers_ckdtree_kernel(const __grid_constant__ KnnParams p) {
    const int n = p.n_dev ? min(*p.n_dev, p.n_cap) : p.n_cap;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (n < 2) {  // cleansing.py:178-179
        p.out[i] = 0;
        return;
    }
    const int kk = min(p.k + 1, n);  // :197
    kd::Tree t;
    t.data = p.xy;
    t.n = n;
    t.idx = p.idx;
    t.nodes = p.nodes;
    t.nnodes = p.meta[0];
    for (int c = 0; c < 2; c++) {
        t.maxes[c] = p.bounds[c];
        t.mins[c] = p.bounds[2 + c];
    }
    int *inds = p.inds + (size_t)i * (p.k + 1);
    kd::query(t, p.xy[2 * (size_t)i], p.xy[2 * (size_t)i + 1], kk, inds, p.nb + (size_t)i * (p.k + 1),
              p.q + (size_t)i * p.qcap, p.pool + (size_t)i * p.qcap);
    p.out[i] = kd::mahalanobis_outlier(p.uv, i, inds, kk - 1, p.thr) ? 1 : 0;
}

struct IdwParams {
    const double *xy, *vals, *xgrid, *ygrid;
    const int *npts_dev;
    int npts_cap, nvar, k, nx, ny;
    double power, offset, mean_res;
    int *idx;
    kd::Node *nodes;
    int *meta;
    double *bounds;
    int *inds;       // nthreads * k
    double *w;       // nthreads * k
    kd::Item *nb;    // nthreads * k
    kd::Item *q;     // nthreads * qcap
    kd::NodeInfo *pool;
    int qcap;
    double *out;     // (nvar, ny, nx)
};

__global__ void kd_build_idw_kernel(const __grid_constant__ IdwParams p) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    const int n = p.npts_dev ? min(*p.npts_dev, p.npts_cap) : p.npts_cap;
    int stack[128];
    kd::Tree t;
    t.data = p.xy;
    t.n = n;
    t.idx = p.idx;
    t.nodes = p.nodes;
    kd::build(t, stack);
    p.meta[0] = t.nnodes;
    for (int c = 0; c < 2; c++) {
        p.bounds[c] = t.maxes[c];
        p.bounds[2 + c] = t.mins[c];
    }
}

// every grid point runs scipy's query and numpy's weighting (knn_body.cuh: idw_point); a fixed
// number of threads strides over the grid so that the per-thread search scratch stays bounded
__global__ void __launch_bounds__(128)
idw_ckdtree_kernel(const __grid_constant__ IdwParams p) {
    const int n = p.npts_dev ? min(*p.npts_dev, p.npts_cap) : p.npts_cap;
    const int k = min(p.k, n);
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t nthreads = (size_t)gridDim.x * blockDim.x;
    const size_t N = (size_t)p.ny * p.nx;
    kd::Tree t;
    t.data = p.xy;
    t.n = n;
    t.idx = p.idx;
    t.nodes = p.nodes;
    t.nnodes = p.meta[0];
    for (int c = 0; c < 2; c++) {
        t.maxes[c] = p.bounds[c];
        t.mins[c] = p.bounds[2 + c];
    }
    int *inds = p.inds + tid * p.k;
    double *w = p.w + tid * p.k;
    for (size_t e = tid; e < N; e += nthreads) {
        const int i = (int)(e / p.nx), j = (int)(e % p.nx);
        kd::query(t, p.xgrid[j], p.ygrid[i], k, inds, p.nb + tid * p.k, p.q + tid * p.qcap, p.pool + tid * p.qcap, w);
        kd::idw_point(p.vals, p.nvar, inds, w, k, p.power, p.offset, p.mean_res, p.out + e, N);
    }
}

}  // namespace

extern "C" int b200_idw_fill_ckdtree(const double *xy, const double *vals, const int *npts_dev, int npts_cap,
                                     int nvar, int k, double power, double dist_offset, double mean_res,
                                     const double *xgrid, int nx, const double *ygrid, int ny, double *out,
                                     void *stream) {
    B200_REQUIRE(xy && vals && xgrid && ygrid && out && npts_cap >= 1 && nvar >= 1 && nx >= 1 && ny >= 1,
                 "bad arguments");
    B200_REQUIRE(k >= 1 && k <= 128, "k must be 1..128");
    cudaStream_t s = (cudaStream_t)stream;
    IdwParams p;
    memset(&p, 0, sizeof(p));
    p.xy = xy; p.vals = vals; p.xgrid = xgrid; p.ygrid = ygrid; p.npts_dev = npts_dev; p.npts_cap = npts_cap;
    p.nvar = nvar; p.k = k; p.nx = nx; p.ny = ny; p.power = power; p.offset = dist_offset; p.mean_res = mean_res;
    p.out = out;
    const size_t N = (size_t)ny * nx;
    const int blocks = (int)std::max<size_t>(1, std::min<size_t>((N + 127) / 128, (size_t)b200::num_sms() * 8));
    const size_t nthreads = (size_t)blocks * 128;
    b200::Scratch idx, nodes, meta, bounds, inds, w, nb, q, pool;
    B200_CUDA(idx.alloc(sizeof(int) * (size_t)npts_cap, s));
    B200_CUDA(nodes.alloc(sizeof(kd::Node) * (size_t)kd::max_nodes(npts_cap), s));
    B200_CUDA(meta.alloc(sizeof(int) * 4, s));
    B200_CUDA(bounds.alloc(sizeof(double) * 4, s));
    B200_CUDA(inds.alloc(sizeof(int) * nthreads * k, s));
    B200_CUDA(w.alloc(sizeof(double) * nthreads * k, s));
    B200_CUDA(nb.alloc(sizeof(kd::Item) * nthreads * k, s));
    p.idx = (int *)idx.p; p.nodes = (kd::Node *)nodes.p; p.meta = (int *)meta.p; p.bounds = (double *)bounds.p;
    p.inds = (int *)inds.p; p.w = (double *)w.p; p.nb = (kd::Item *)nb.p;
    kd_build_idw_kernel<<<1, 1, 0, s>>>(p);
    B200_LAUNCH_CHECK();
    int nnodes = 0;
    B200_CUDA(cudaMemcpyAsync(&nnodes, p.meta, sizeof(int), cudaMemcpyDeviceToHost, s));
    B200_CUDA(cudaStreamSynchronize(s));
    p.qcap = nnodes > 0 ? nnodes : 1;
    B200_CUDA(q.alloc(sizeof(kd::Item) * nthreads * p.qcap, s));
    B200_CUDA(pool.alloc(sizeof(kd::NodeInfo) * nthreads * p.qcap, s));
    p.q = (kd::Item *)q.p; p.pool = (kd::NodeInfo *)pool.p;
    idw_ckdtree_kernel<<<blocks, 128, 0, s>>>(p);
    B200_LAUNCH_CHECK();
    return 0;
}

extern "C" int b200_detect_outliers_ckdtree(const double *uv, const double *xy, const int *n_dev, int n_cap,
                                            double thr, int k, uint8_t *out, void *stream) {
    B200_REQUIRE(uv != nullptr && xy != nullptr && out != nullptr && n_cap >= 1, "bad arguments");
    B200_REQUIRE(k >= 1 && k <= 4096, "k out of range");
    cudaStream_t s = (cudaStream_t)stream;
    KnnParams p;
    memset(&p, 0, sizeof(p));
    p.xy = xy; p.uv = uv; p.n_dev = n_dev; p.n_cap = n_cap; p.k = k; p.thr = thr; p.out = out;
    b200::Scratch idx, nodes, meta, bounds, inds, nb, q, pool;
    B200_CUDA(idx.alloc(sizeof(int) * (size_t)n_cap, s));
    B200_CUDA(nodes.alloc(sizeof(kd::Node) * (size_t)kd::max_nodes(n_cap), s));
    B200_CUDA(meta.alloc(sizeof(int) * 4, s));
    B200_CUDA(bounds.alloc(sizeof(double) * 4, s));
    B200_CUDA(inds.alloc(sizeof(int) * (size_t)n_cap * (k + 1), s));
    B200_CUDA(nb.alloc(sizeof(kd::Item) * (size_t)n_cap * (k + 1), s));
    p.idx = (int *)idx.p; p.nodes = (kd::Node *)nodes.p; p.meta = (int *)meta.p; p.bounds = (double *)bounds.p;
    p.inds = (int *)inds.p; p.nb = (kd::Item *)nb.p;
    kd_build_kernel<<<1, 1, 0, s>>>(p);
    B200_LAUNCH_CHECK();
    // A query queues at most one far child per internal node it visits, so the node count bounds
    // its queue; it is known only now (one small read-back; this entry point synchronises).
    int nnodes = 0;
    B200_CUDA(cudaMemcpyAsync(&nnodes, p.meta, sizeof(int), cudaMemcpyDeviceToHost, s));
    B200_CUDA(cudaStreamSynchronize(s));
    p.qcap = nnodes > 0 ? nnodes : 1;
    B200_CUDA(q.alloc(sizeof(kd::Item) * (size_t)n_cap * p.qcap, s));
    B200_CUDA(pool.alloc(sizeof(kd::NodeInfo) * (size_t)n_cap * p.qcap, s));
    p.q = (kd::Item *)q.p; p.pool = (kd::NodeInfo *)pool.p;
    outliers_ckdtree_kernel<<<b200::ceil_div(n_cap, 64), 64, 0, s>>>(p);
    B200_LAUNCH_CHECK();
    return 0;
}
