// sparse.cu -- the sparse-vector cleansing stages of dense_lucaskanade on the device
// (sm_100a): local Mahalanobis outlier detection and grid-cell declustering.
//
// Reference: pysteps/utils/cleansing.py:124-249 (detect_outliers, coord + k given) and
// pysteps/utils/cleansing.py:21-121 (decluster).  Both are Python loops over <= a few
// thousand vectors around cKDTree / np.cov / np.linalg.inv / np.median; here each is one
// small kernel (float64, deterministic order) so the vectors never leave the device.
#include <math_constants.h>

#include "common.cuh"

namespace {

constexpr int OUT_WARPS = 4;
constexpr int OUT_KMAX = 64;  // neighbours incl. self

// one warp per vector: k+1 nearest by (distance, index) via repeated warp arg-min, then the
// 2x2 sample covariance of the k neighbours and the Mahalanobis distance of the vector.
// CACHED: n <= 32 * OUT_CACHE, every lane keeps its squared distances in registers, so the
// k+1 selection rounds only compare (the distances are computed once, not k+1 times).
constexpr int OUT_CACHE = 64;

template <bool CACHED>
__global__ void __launch_bounds__(32 * OUT_WARPS)
outliers_kernel(const double *__restrict__ uv, const double *__restrict__ xy, const int *__restrict__ n_dev,
                int n_cap, double thr, int k, uint8_t *__restrict__ out) {
    __shared__ int nbr[OUT_WARPS][OUT_KMAX];
    const int n = n_dev ? min(*n_dev, n_cap) : n_cap;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int i = blockIdx.x * OUT_WARPS + wid;
    if (i >= n) return;
    if (n < 2) {  // cleansing.py:177-178
        if (lane == 0) out[i] = 0;
        return;
    }
    const int kk = min(n, k + 1);  // :198
    const double xi = xy[2 * i], yi = xy[2 * i + 1];
    double last_d = -1.0;
    int last_j = -1;
    double dc[CACHED ? OUT_CACHE : 1];
    if (CACHED) {
#pragma unroll
        for (int q = 0; q < OUT_CACHE; q++) {
            const int j = q * 32 + lane;
            dc[q] = CUDART_INF;
            if (j < n) {
                const double dx = __dsub_rn(xy[2 * j], xi), dy = __dsub_rn(xy[2 * j + 1], yi);
                dc[q] = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
            }
        }
    }
    for (int r = 0; r < kk; r++) {
        double bd = CUDART_INF;
        int bj = 0x7fffffff;
        if (CACHED) {
#pragma unroll
            for (int q = 0; q < OUT_CACHE; q++) {
                const int j = q * 32 + lane;
                const double d = dc[q];
                // ascending j within a lane: the first strict improvement is the lowest index
                const bool after = (d > last_d) || (d == last_d && j > last_j);
                if (j < n && after && d < bd) { bd = d; bj = j; }
            }
        } else {
            for (int j = lane; j < n; j += 32) {
                const double dx = __dsub_rn(xy[2 * j], xi), dy = __dsub_rn(xy[2 * j + 1], yi);
                const double d = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
                const bool after = (d > last_d) || (d == last_d && j > last_j);
                if (after && (d < bd || (d == bd && j < bj))) { bd = d; bj = j; }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double od = __shfl_xor_sync(0xffffffffu, bd, o);
            const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
            if (od < bd || (od == bd && oj < bj)) { bd = od; bj = oj; }
        }
        last_d = bd; last_j = bj;
        if (lane == 0) nbr[wid][r] = bj;
    }
    __syncwarp();
    if (lane != 0) return;
    const int m = kk - 1;  // neighbours without the nearest (normally the vector itself), :228,234
    // mean of the neighbours (row by row) and centring, :236-237
    double mu = 0.0, mv = 0.0;
    for (int q = 1; q <= m; q++) {
        const int j = nbr[wid][q];
        mu = (q == 1) ? uv[2 * j] : __dadd_rn(mu, uv[2 * j]);
        mv = (q == 1) ? uv[2 * j + 1] : __dadd_rn(mv, uv[2 * j + 1]);
    }
    mu = __ddiv_rn(mu, (double)m);
    mv = __ddiv_rn(mv, (double)m);
    const double zu = __dsub_rn(uv[2 * i], mu), zv = __dsub_rn(uv[2 * i + 1], mv);
    // np.cov of the centred neighbours: subtract their (tiny) mean again, ddof = 1
    double au = 0.0, av = 0.0;
    for (int q = 1; q <= m; q++) {
        const int j = nbr[wid][q];
        au = __dadd_rn(au, __dsub_rn(uv[2 * j], mu));
        av = __dadd_rn(av, __dsub_rn(uv[2 * j + 1], mv));
    }
    au = __ddiv_rn(au, (double)m);
    av = __ddiv_rn(av, (double)m);
    double suu = 0.0, suv = 0.0, svv = 0.0;
    for (int q = 1; q <= m; q++) {
        const int j = nbr[wid][q];
        const double a = __dsub_rn(__dsub_rn(uv[2 * j], mu), au);
        const double b = __dsub_rn(__dsub_rn(uv[2 * j + 1], mv), av);
        suu = __dadd_rn(suu, __dmul_rn(a, a));
        suv = __dadd_rn(suv, __dmul_rn(a, b));
        svv = __dadd_rn(svv, __dmul_rn(b, b));
    }
    const double fact = __ddiv_rn(1.0, (double)(m - 1));  // m == 1 -> inf/nan like np.cov
    const double a = __dmul_rn(suu, fact), b = __dmul_rn(suv, fact), d = __dmul_rn(svv, fact);
    // np.linalg.inv: LU with partial pivoting; exactly singular -> LinAlgError -> MD = 0
    double MD = 0.0;
    const bool swap = fabs(b) > fabs(a);
    const double p0 = swap ? b : a, p1 = swap ? d : b;   // pivot row
    const double q0 = swap ? a : b, q1 = swap ? b : d;   // other row
    if (p0 != 0.0 && !(isnan(p0))) {
        const double l = __dmul_rn(q0, __ddiv_rn(1.0, p0));
        const double u22 = __dsub_rn(q1, __dmul_rn(l, p1));
        if (u22 != 0.0) {
            // solve V X = I column by column (rows permuted when swap)
            // column e0, e1 of the identity after the row permutation
            const double r00 = swap ? 0.0 : 1.0, r10 = swap ? 1.0 : 0.0;  // P*e0
            const double r01 = swap ? 1.0 : 0.0, r11 = swap ? 0.0 : 1.0;  // P*e1
            const double y10 = __dsub_rn(r10, __dmul_rn(l, r00)), y11 = __dsub_rn(r11, __dmul_rn(l, r01));
            const double x10 = __ddiv_rn(y10, u22), x11 = __ddiv_rn(y11, u22);
            const double x00 = __ddiv_rn(__dsub_rn(r00, __dmul_rn(p1, x10)), p0);
            const double x01 = __ddiv_rn(__dsub_rn(r01, __dmul_rn(p1, x11)), p0);
            // MD = sqrt(z VI z^T), :241
            const double t0 = __dadd_rn(__dmul_rn(zu, x00), __dmul_rn(zv, x10));
            const double t1 = __dadd_rn(__dmul_rn(zu, x01), __dmul_rn(zv, x11));
            MD = sqrt(__dadd_rn(__dmul_rn(t0, zu), __dmul_rn(t1, zv)));
        }
    }
    out[i] = (MD > thr) ? 1 : 0;  // NaN compares false, as in NumPy
}

// keep rows whose flag is 0, preserving order (xy[~outliers], uv[~outliers])
__global__ void __launch_bounds__(1024)
compact_rows_kernel(const double *__restrict__ xy, const double *__restrict__ uv, const uint8_t *__restrict__ drop,
                    const int *__restrict__ n_dev, int n_cap, double *__restrict__ oxy, double *__restrict__ ouv,
                    int *__restrict__ out_count) {
    __shared__ int warp_tot[32];
    __shared__ int s_base;
    const int n = n_dev ? min(*n_dev, n_cap) : n_cap;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int start = 0; start < n; start += blockDim.x) {
        const int i = start + tid;
        const bool keep = i < n && !drop[i];
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) warp_tot[wid] = __popc(bal);
        __syncthreads();
        int before = 0, total = 0;
        for (int q = 0; q < 32; q++) {
            if (q < wid) before += warp_tot[q];
            total += warp_tot[q];
        }
        if (keep) {
            const int o = s_base + before + __popc(bal & ((1u << lane) - 1u));
            oxy[2 * o] = xy[2 * i]; oxy[2 * o + 1] = xy[2 * i + 1];
            ouv[2 * o] = uv[2 * i]; ouv[2 * o + 1] = uv[2 * i + 1];
        }
        __syncthreads();
        if (tid == 0) s_base += total;
        __syncthreads();
    }
    if (tid == 0) *out_count = s_base;
}

// ---- decluster ---------------------------------------------------------------------------
constexpr int DC_MAX = 4096;

__device__ __forceinline__ double median_of(const double *__restrict__ a, int stride,
                                            const unsigned long long *__restrict__ keys, int s) {
    // rank selection (s is small): the elements of rank (s-1)/2 and s/2, averaged
    double lo = 0.0, hi = 0.0;
    const int rlo = (s - 1) / 2, rhi = s / 2;
    for (int p = 0; p < s; p++) {
        const double v = a[(size_t)(keys[p] & 0x3fffff) * stride];
        int rank = 0;
        for (int q = 0; q < s; q++) {
            const double u = a[(size_t)(keys[q] & 0x3fffff) * stride];
            rank += (u < v) || (u == v && q < p);
        }
        if (rank == rlo) lo = v;
        if (rank == rhi) hi = v;
    }
    return (rlo == rhi) ? lo : __ddiv_rn(__dadd_rn(lo, hi), 2.0);
}

__global__ void __launch_bounds__(1024)
decluster_kernel(const double *__restrict__ xy, const double *__restrict__ uv, const int *__restrict__ n_dev,
                 int n_cap, double scale, int min_samples, double *__restrict__ oxy, double *__restrict__ ouv,
                 int *__restrict__ out_count) {
    __shared__ unsigned long long key[DC_MAX];
    __shared__ unsigned short seg_start[DC_MAX + 1];
    __shared__ int s_nseg;
    const int n = n_dev ? min(*n_dev, n_cap) : n_cap;
    const int tid = threadIdx.x;
    int npad = 1;
    while (npad < n) npad <<= 1;
    if (npad < 2) npad = 2;
    // key = (cell_x, cell_y, index): np.unique(axis=0) orders cells lexicographically by x then y
    for (int i = tid; i < npad; i += blockDim.x) {
        unsigned long long kv = ~0ull;
        if (i < n) {
            const long long cx = (long long)floor(__ddiv_rn(xy[2 * i], scale)) + (1 << 20);
            const long long cy = (long long)floor(__ddiv_rn(xy[2 * i + 1], scale)) + (1 << 20);
            kv = ((unsigned long long)(cx & 0x1fffff) << 43) | ((unsigned long long)(cy & 0x1fffff) << 22) |
                 (unsigned long long)i;
        }
        key[i] = kv;
    }
    __syncthreads();
    for (int k = 2; k <= npad; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < npad / 2; t += blockDim.x) {
                const int lo = 2 * t - (t & (j - 1)), hi = lo + j;
                const bool asc = (lo & k) == 0;
                const unsigned long long a = key[lo], b = key[hi];
                if ((a > b) == asc) { key[lo] = b; key[hi] = a; }
            }
            __syncthreads();
        }
    // segment heads (first vector of every occupied cell) by a block-wide exclusive scan
    {
        __shared__ int warp_sum[32];
        const int per = (npad + (int)blockDim.x - 1) / (int)blockDim.x;  // consecutive elements per thread
        const int i0 = tid * per;
        int local = 0;
        for (int q = 0; q < per; q++) {
            const int i = i0 + q;
            if (i < n && (i == 0 || (key[i] >> 22) != (key[i - 1] >> 22))) local++;
        }
        int incl = local;
        const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) warp_sum[wid] = incl;
        __syncthreads();
        int base = 0;
        for (int w = 0; w < wid; w++) base += warp_sum[w];
        int pos = base + incl - local;  // number of heads before this thread's elements
        for (int q = 0; q < per; q++) {
            const int i = i0 + q;
            if (i < n && (i == 0 || (key[i] >> 22) != (key[i - 1] >> 22))) seg_start[pos++] = (unsigned short)i;
        }
        if (tid == (int)blockDim.x - 1) {
            s_nseg = base + incl;
            seg_start[base + incl] = (unsigned short)n;
        }
    }
    __syncthreads();
    const int nseg = s_nseg;
    // kept segments keep their sorted order; with min_samples > 1 the output index is the
    // number of kept segments before this one
    for (int sgi = tid; sgi < nseg; sgi += blockDim.x) {
        const int s0 = seg_start[sgi], s = seg_start[sgi + 1] - s0;
        if (s < min_samples) continue;
        int o = sgi;  // every cell is kept when min_samples <= 1 (what dense_lucaskanade passes)
        if (min_samples > 1) {
            o = 0;
            for (int q = 0; q < sgi; q++) o += (seg_start[q + 1] - seg_start[q]) >= min_samples;
        }
        ouv[2 * o] = median_of(uv, 2, key + s0, s);
        ouv[2 * o + 1] = median_of(uv + 1, 2, key + s0, s);
        oxy[2 * o] = median_of(xy, 2, key + s0, s);
        oxy[2 * o + 1] = median_of(xy + 1, 2, key + s0, s);
    }
    __syncthreads();
    if (tid == 0) {
        int o = 0;
        for (int q = 0; q < nseg; q++) o += (seg_start[q + 1] - seg_start[q]) >= min_samples;
        *out_count = o;
    }
}

}  // namespace

extern "C" int b200_detect_outliers(const double *uv, const double *xy, const int *n_dev, int n_cap,
                                    double thr, int k, uint8_t *out, void *stream) {
    B200_REQUIRE(uv && xy && out && n_cap >= 0 && k >= 1, "bad arguments");
    if (k + 1 > OUT_KMAX) {
        b200::set_error("detect_outliers: k must be < %d", OUT_KMAX);
        return B200_ENOTSUP;
    }
    if (n_cap == 0) return 0;
    if (n_cap <= 32 * OUT_CACHE)
        outliers_kernel<true><<<b200::ceil_div(n_cap, OUT_WARPS), 32 * OUT_WARPS, 0, (cudaStream_t)stream>>>(
            uv, xy, n_dev, n_cap, thr, k, out);
    else
        outliers_kernel<false><<<b200::ceil_div(n_cap, OUT_WARPS), 32 * OUT_WARPS, 0, (cudaStream_t)stream>>>(
            uv, xy, n_dev, n_cap, thr, k, out);
    B200_LAUNCH_CHECK();
    return 0;
}

extern "C" int b200_compact_rows(const double *xy, const double *uv, const uint8_t *drop, const int *n_dev,
                                 int n_cap, double *out_xy, double *out_uv, int *out_count, void *stream) {
    B200_REQUIRE(xy && uv && drop && out_xy && out_uv && out_count && n_cap >= 0, "bad arguments");
    compact_rows_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(xy, uv, drop, n_dev, n_cap, out_xy, out_uv, out_count);
    B200_LAUNCH_CHECK();
    return 0;
}

extern "C" int b200_decluster(const double *xy, const double *uv, const int *n_dev, int n_cap, double scale,
                              int min_samples, double *out_xy, double *out_uv, int *out_count, void *stream) {
    B200_REQUIRE(xy && uv && out_xy && out_uv && out_count && n_cap >= 0 && scale > 0.0, "bad arguments");
    if (n_cap > DC_MAX) {
        b200::set_error("decluster: at most %d vectors are supported", DC_MAX);
        return B200_ENOTSUP;
    }
    decluster_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(xy, uv, n_dev, n_cap, scale, min_samples, out_xy, out_uv,
                                                          out_count);
    B200_LAUNCH_CHECK();
    return 0;
}
