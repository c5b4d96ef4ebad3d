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

constexpr int IDW_TX = 32, IDW_TY = 8;            // pixel tile of one CTA
constexpr int IDW_THREADS = IDW_TX * IDW_TY;
constexpr int IDW_CHUNK = 2048;                    // source vectors examined per round
constexpr int IDW_BINS = 256;                      // distance histogram of the tile centre

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

// One CTA fills a 32x8 pixel tile.  Exhaustive search is exact but wasteful (every pixel
// against every vector); the tile first bounds its search radius: with Rk >= distance from
// the tile centre c to its k-th nearest vector and r the tile's half diagonal, every pixel
// p of the tile has its k nearest within Rk + r of p, hence within Rk + 2r of c.  Vectors
// outside that disc cannot be among any pixel's k nearest and are dropped -- the result is
// identical to the exhaustive search, with ~k..3k candidates per tile instead of all.
// Candidates are compacted IN INDEX ORDER, so equal distances still resolve to the lower
// index exactly as in the exhaustive scan.
template <int K>
__global__ void __launch_bounds__(IDW_THREADS) idw_kernel(const IDWParams p) {
    __shared__ double2 spt[IDW_CHUNK];
    __shared__ int sidx[IDW_CHUNK];
    __shared__ int hist[IDW_BINS];
    __shared__ int warp_cnt[IDW_THREADS / 32];
    __shared__ double s_R2;
    __shared__ int s_n;
    const int npts = p.npts_dev ? min(*p.npts_dev, p.npts_cap) : p.npts_cap;
    const int k = min(min(p.k, npts), K);
    const int tid = threadIdx.y * IDW_TX + threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int j = blockIdx.x * IDW_TX + threadIdx.x;  // column
    const int i = blockIdx.y * IDW_TY + threadIdx.y;  // row
    const bool active = j < p.nx && i < p.ny;
    const double qx = p.gx[min(j, p.nx - 1)], qy = p.gy[min(i, p.ny - 1)];

    // ---- tile centre, half diagonal, histogram of centre distances ---------------------
    const int j0 = blockIdx.x * IDW_TX, j1 = min(j0 + IDW_TX, p.nx) - 1;
    const int i0 = blockIdx.y * IDW_TY, i1 = min(i0 + IDW_TY, p.ny) - 1;
    const double xa = p.gx[j0], xb = p.gx[j1], ya = p.gy[i0], yb = p.gy[i1];
    const double cx = 0.5 * (xa + xb), cy = 0.5 * (ya + yb);
    // grids are monotonic (np.arange in dense_lucaskanade); the tile extent bounds the radius
    const double rt = sqrt(0.25 * (xb - xa) * (xb - xa) + 0.25 * (yb - ya) * (yb - ya));
    const double binw = fmax(rt, 1e-300) * 0.5;
    for (int b = tid; b < IDW_BINS; b += IDW_THREADS) hist[b] = 0;
    __syncthreads();
    for (int t = tid; t < npts; t += IDW_THREADS) {
        const double dx = p.xy[2 * t] - cx, dy = p.xy[2 * t + 1] - cy;
        const double d = sqrt(dx * dx + dy * dy) / binw;
        const int b = d < (double)(IDW_BINS - 1) ? (int)d : IDW_BINS - 1;
        atomicAdd(&hist[b], 1);
    }
    __syncthreads();
    if (tid == 0) {
        int acc = 0, b = 0;
        for (; b < IDW_BINS; b++) {
            acc += hist[b];
            if (acc >= k) break;
        }
        // overflow bin: no bound (every vector is a candidate)
        double R = (b >= IDW_BINS - 1) ? CUDART_INF : ((double)(b + 1) * binw + 2.0 * rt) * (1.0 + 1e-9);
        s_R2 = R * R;
    }
    __syncthreads();
    const double R2 = s_R2;

    double bd[K];
    int bi[K];
#pragma unroll
    for (int q = 0; q < K; q++) { bd[q] = CUDART_INF; bi[q] = 0; }
    double worst = CUDART_INF;  // bd[k-1], refreshed only when the list changes

    for (int base = 0; base < npts; base += IDW_CHUNK) {
        const int cnt = min(IDW_CHUNK, npts - base);
        // ---- ordered compaction of the candidates of this chunk ------------------------
        __syncthreads();
        if (tid == 0) s_n = 0;
        __syncthreads();
        for (int t0 = 0; t0 < cnt; t0 += IDW_THREADS) {
            const int t = t0 + tid;
            double2 s = make_double2(0.0, 0.0);
            bool keep = false;
            if (t < cnt) {
                s = make_double2(p.xy[2 * (base + t)], p.xy[2 * (base + t) + 1]);
                const double dx = s.x - cx, dy = s.y - cy;
                keep = !(dx * dx + dy * dy > R2);
            }
            const unsigned bal = __ballot_sync(0xffffffffu, keep);
            if (lane == 0) warp_cnt[wid] = __popc(bal);
            __syncthreads();
            int before = s_n, total = 0;
#pragma unroll
            for (int w = 0; w < IDW_THREADS / 32; w++) {
                if (w < wid) before += warp_cnt[w];
                total += warp_cnt[w];
            }
            if (keep) {
                const int o = before + __popc(bal & ((1u << lane) - 1u));
                spt[o] = s;
                sidx[o] = base + t;
            }
            __syncthreads();
            if (tid == 0) s_n += total;
            __syncthreads();
        }
        const int ncand = s_n;
        if (!active) continue;
        // ---- exhaustive top-k over the candidates ------------------------------------
        for (int t = 0; t < ncand; t++) {
            const double2 s = spt[t];
            const double dx = __dsub_rn(s.x, qx), dy = __dsub_rn(s.y, qy);
            const double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
            if (d2 < worst) {
                // replace the worst, then bubble towards the front (strict <: earlier index
                // stays first on equal distances)
                const int id = sidx[t];
#pragma unroll
                for (int q = K - 1; q >= 0; q--)
                    if (q == k - 1) { bd[q] = d2; bi[q] = id; }
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
    dim3 grid(b200::ceil_div(nx, IDW_TX), b200::ceil_div(ny, IDW_TY));
    dim3 block(IDW_TX, IDW_TY);
    cudaStream_t s = (cudaStream_t)stream;
    if (k <= 8) idw_kernel<8><<<grid, block, 0, s>>>(p);
    else if (k <= 20) idw_kernel<20><<<grid, block, 0, s>>>(p);
    else idw_kernel<32><<<grid, block, 0, s>>>(p);
    B200_LAUNCH_CHECK();
    return 0;
}
