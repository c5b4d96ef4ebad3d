// idw.cu -- inverse-distance-weighted k-nearest-neighbour grid fill (sm_100a).
//
// Reference: pysteps/utils/interpolate.py:67-114 (idwinterp2d): cKDTree.query(grid, k) over
// every grid point, dist/mean_res + offset, w = 1/dist^power normalised, weighted sum of
// the k values.  This is 95 % of the reference's dense_lucaskanade wall time (16.7 s of
// 17.6 s at 2048^2, single-threaded tree queries).  With <= a few thousand source vectors
// an EXHAUSTIVE top-k per pixel is the GPU-natural form: the vectors are staged in shared
// memory once per CTA and every thread scans them for its own pixel, keeping the k best
// (squared distance, index) pairs sorted in registers.  It is FP64-ALU bound, not HBM
// bound (~5 FP64 ops per pixel-vector pair vs 16 B written per pixel); distances and
// weights are float64 in the reference's operation order (ties: lower index first).
#include <math_constants.h>

#include "common.cuh"

namespace {

constexpr int IDW_THREADS = 128;
constexpr int IDW_CHUNK = 2048;  // vectors staged in shared memory at a time (32 KB)

struct IDWParams {
    const double *xy;    // (npts,2)
    const double *vals;  // (npts,nvar)
    const int *npts_dev;
    int npts_cap, nvar, k;
    const double *gx, *gy;
    int nx, ny;
    double power, offset, mean_res;
    double *out;  // (nvar, ny, nx)
};

// numpy's pairwise summation for n < 128 (8 accumulators, then the remainder)
template <int K>
__device__ __forceinline__ double np_sum(const double (&w)[K], int k) {
    if (k < 8) {
        double r = w[0];
#pragma unroll
        for (int i = 1; i < K; i++)
            if (i < k) r = __dadd_rn(r, w[i]);
        return r;
    }
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; j++) r[j] = w[j];
    const int lim = k - (k % 8);
#pragma unroll
    for (int i = 8; i < K; i++)
        if (i < lim) r[i & 7] = __dadd_rn(r[i & 7], w[i]);
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                           __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
#pragma unroll
    for (int i = 8; i < K; i++)
        if (i >= lim && i < k) res = __dadd_rn(res, w[i]);
    return res;
}

template <int K>
__global__ void __launch_bounds__(IDW_THREADS) idw_kernel(const IDWParams p) {
    __shared__ double2 spt[IDW_CHUNK];
    const int npts = p.npts_dev ? min(*p.npts_dev, p.npts_cap) : p.npts_cap;
    const int k = min(min(p.k, npts), K);
    const int j = blockIdx.x * IDW_THREADS + threadIdx.x;  // column
    const int i = blockIdx.y;                               // row
    const bool active = j < p.nx;
    const double qx = active ? p.gx[j] : 0.0, qy = p.gy[i];
    double bd[K];
    int bi[K];
#pragma unroll
    for (int q = 0; q < K; q++) { bd[q] = CUDART_INF; bi[q] = 0; }
    double worst = CUDART_INF;  // bd[k-1], refreshed only when the list changes
    for (int base = 0; base < npts; base += IDW_CHUNK) {
        const int cnt = min(IDW_CHUNK, npts - base);
        __syncthreads();
        for (int t = threadIdx.x; t < cnt; t += IDW_THREADS)
            spt[t] = make_double2(p.xy[2 * (base + t)], p.xy[2 * (base + t) + 1]);
        __syncthreads();
        if (!active) continue;
        for (int t = 0; t < cnt; t++) {
            const double2 s = spt[t];
            const double dx = __dsub_rn(s.x, qx), dy = __dsub_rn(s.y, qy);
            const double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
            if (d2 < worst) {
                // replace the worst, then bubble towards the front (strict <: earlier index
                // stays first on equal distances)
#pragma unroll
                for (int q = K - 1; q >= 0; q--)
                    if (q == k - 1) { bd[q] = d2; bi[q] = base + t; }
#pragma unroll
                for (int q = K - 1; q >= 1; q--)
                    if (q <= k - 1 && bd[q] < bd[q - 1]) {
                        const double td = bd[q]; bd[q] = bd[q - 1]; bd[q - 1] = td;
                        const int ti = bi[q]; bi[q] = bi[q - 1]; bi[q - 1] = ti;
                    }
                // K is a compile-time bound, k <= K is the runtime list length
#pragma unroll
                for (int q = 0; q < K; q++)
                    if (q == k - 1) worst = bd[q];
            }
        }
    }
    if (!active || k < 1) return;
    double w[K];
#pragma unroll
    for (int q = 0; q < K; q++) {
        double d = sqrt(bd[q]);                 // exact Euclidean distance (IEEE sqrt)
        d = __ddiv_rn(d, p.mean_res);           // interpolate.py:98
        d = __dadd_rn(d, p.offset);             // :101
        const double pw = (p.power == 0.5) ? sqrt(d) : pow(d, p.power);
        w[q] = (q < k) ? __ddiv_rn(1.0, pw) : 0.0;  // :102
    }
    const double ws = np_sum<K>(w, k);          // :103
    for (int v = 0; v < p.nvar; v++) {
        double acc = 0.0;
#pragma unroll
        for (int q = 0; q < K; q++)
            if (q < k) {
                const double term = __dmul_rn(p.vals[(size_t)bi[q] * p.nvar + v], __ddiv_rn(w[q], ws));
                acc = (q == 0) ? term : __dadd_rn(acc, term);  // :106-109
            }
        p.out[((size_t)v * p.ny + i) * p.nx + j] = acc;
    }
}

}  // namespace

extern "C" int b200_idw_fill(const double *xy, const double *vals, const int *npts_dev, int npts_cap,
                             int nvar, int k, double power, double dist_offset, double mean_res,
                             const double *xgrid, int nx, const double *ygrid, int ny, double *out,
                             void *stream) {
    B200_REQUIRE(xy && vals && xgrid && ygrid && out && npts_cap >= 1 && nvar >= 1 && nx >= 1 && ny >= 1 &&
                     k >= 1, "bad arguments");
    if (k > 32) {
        b200::set_error("idw: k must be <= 32 (k=None / larger k is not implemented)");
        return B200_ENOTSUP;
    }
    IDWParams p;
    p.xy = xy; p.vals = vals; p.npts_dev = npts_dev; p.npts_cap = npts_cap; p.nvar = nvar; p.k = k;
    p.gx = xgrid; p.gy = ygrid; p.nx = nx; p.ny = ny;
    p.power = power; p.offset = dist_offset; p.mean_res = mean_res; p.out = out;
    dim3 grid(b200::ceil_div(nx, IDW_THREADS), ny);
    cudaStream_t s = (cudaStream_t)stream;
    if (k <= 8) idw_kernel<8><<<grid, IDW_THREADS, 0, s>>>(p);
    else if (k <= 20) idw_kernel<20><<<grid, IDW_THREADS, 0, s>>>(p);
    else idw_kernel<32><<<grid, IDW_THREADS, 0, s>>>(p);
    B200_LAUNCH_CHECK();
    return 0;
}
