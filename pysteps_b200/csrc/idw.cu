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
// weights are float64.
//
// Neighbour ORDER: an exhaustive scan has no tree, so equal distances are listed by lower index,
// scipy.spatial.cKDTree lists them in its tree's order.  Inside the list that only permutes equal
// weights (identical result up to the order of one addition); it matters where the k-th and the
// (k+1)-th neighbour are exactly equidistant, because then the two pick different vectors.  Every
// thread therefore tracks the smallest candidate it did NOT keep; a grid point whose k-th distance
// equals it is appended to a list and recomputed from scipy's own query (knn.cu: idw_fix_kernel,
// tree built on a side stream while this kernel runs).  The fill is then the reference's at every
// grid point (<= 1e-12; the weights use rsqrt where NumPy uses sqrt/power/divide).
#include <math_constants.h>

#include "common.cuh"
#include "knn_device.cuh"

namespace {

constexpr int IDW_TX = 16, IDW_TY = 16;           // pixel tile of one CTA (8 warps of 8x4 pixels)
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
    int *tie_list;   // grid points whose k-th and (k+1)-th neighbours are equidistant (or null)
    int *tie_count;
    uint8_t *tile_done;  // per pixel tile: filled by the 32-bit-key kernel (or null)
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
// The candidates are bucketed by centre distance (counting sort on the histogram that gave
// Rk), so each pixel meets its near vectors first and later ones fail the `< worst` test
// without touching the sorted list.  Order of examination is irrelevant for the result:
// the list is ordered by (squared distance, index), i.e. equal distances resolve to the
// lower index exactly as cKDTree-free exhaustive scanning in index order would.
// Batcher odd-even merge sorting networks (generated, verified exhaustively with the 0-1
// principle): compile-time comparator lists, so the sorted list never leaves registers.
__device__ constexpr int NET20[103][2] = {{0,1},{2,3},{4,5},{6,7},{8,9},{10,11},{12,13},{14,15},{16,17},{18,19},{0,2},{1,3},{4,6},{5,7},{8,10},{9,11},{12,14},{13,15},{16,18},{17,19},{1,2},{5,6},{9,10},{13,14},{17,18},{0,4},{1,5},{2,6},{3,7},{8,12},{9,13},{10,14},{11,15},{2,4},{3,5},{10,12},{11,13},{1,2},{3,4},{5,6},{9,10},{11,12},{13,14},{17,18},{0,8},{1,9},{2,10},{3,11},{4,12},{5,13},{6,14},{7,15},{4,8},{5,9},{6,10},{7,11},{2,4},{3,5},{6,8},{7,9},{10,12},{11,13},{1,2},{3,4},{5,6},{7,8},{9,10},{11,12},{13,14},{17,18},{0,16},{1,17},{2,18},{3,19},{8,16},{9,17},{10,18},{11,19},{4,8},{5,9},{6,10},{7,11},{12,16},{13,17},{14,18},{15,19},{2,4},{3,5},{6,8},{7,9},{10,12},{11,13},{14,16},{15,17},{1,2},{3,4},{5,6},{7,8},{9,10},{11,12},{13,14},{15,16},{17,18}};
__device__ constexpr int NET8[19][2] = {{0,1},{2,3},{4,5},{6,7},{0,2},{1,3},{4,6},{5,7},{1,2},{5,6},{0,4},{1,5},{2,6},{3,7},{2,4},{3,5},{1,2},{3,4},{5,6}};
template <int K> __device__ __forceinline__ constexpr int net_size() { return K == 20 ? 103 : (K == 8 ? 19 : 0); }
template <int K> __device__ __forceinline__ constexpr int net_a(int c) { return K == 20 ? NET20[c < 103 ? c : 0][0] : (K == 8 ? NET8[c < 19 ? c : 0][0] : 0); }
template <int K> __device__ __forceinline__ constexpr int net_b(int c) { return K == 20 ? NET20[c < 103 ? c : 0][1] : (K == 8 ? NET8[c < 19 ? c : 0][1] : 0); }

__device__ __forceinline__ bool key_less(unsigned long long da, int ia, unsigned long long db, int ib) {
    return da < db || (da == db && ia < ib);
}

// Sorted list of the k best (squared distance, index) pairs of one pixel.
// EXACT = true: k == K, the list lives in registers (every index is a compile-time constant).
// The first K candidates are loaded as they come and sorted once by a sorting network; every
// later candidate is tested against the current worst and, if better, inserted branch free:
// K independent "key < entry" predicates, then each slot takes its left neighbour, the key,
// or keeps its value.  Squared distances are >= 0, so their bit patterns order like the
// doubles; (bits, index) is compared as one integer key.
// EXACT = false (k < K: fewer vectors than neighbours, or an unusual k): insertion only, with
// a runtime length -- a rare, small-problem path.
// PACKED (fast path of dense_lucaskanade): all coordinates are multiples of 1/16 below 2^14
// and there are at most 2048 vectors, so a squared distance is a multiple of 1/256 below 2^29
// and the 11 low mantissa bits of its float64 pattern are zero: the vector index is stored
// there.  One 64-bit integer then carries (distance, index) in exactly the required order --
// a comparator is one compare and two selects, and the index array disappears.
template <int K, bool EXACT, bool PACKED>
__device__ __forceinline__ void topk_scan(const double2 *__restrict__ spt, const int *__restrict__ sidx,
                                          int ncand, double qx, double qy, int k, bool first,
                                          unsigned long long (&bd)[K], int (&bi)[K], unsigned long long &rej) {
    auto dist2 = [&](int t) -> unsigned long long {
        const double2 s = spt[t];
        const double dx = __dsub_rn(s.x, qx), dy = __dsub_rn(s.y, qy);
        const unsigned long long b =
            (unsigned long long)__double_as_longlong(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
        return PACKED ? (b | (unsigned long long)sidx[t]) : b;
    };
    int t0 = 0;
    if (EXACT && first && ncand >= K) {
#pragma unroll
        for (int q = 0; q < K; q++) { bd[q] = dist2(q); if (!PACKED) bi[q] = sidx[q]; }
#pragma unroll
        for (int c = 0; c < net_size<K>(); c++) {
            const int a = net_a<K>(c), b = net_b<K>(c);
            if (PACKED) {
                const unsigned long long va = bd[a], vb = bd[b];
                const bool sw = vb < va;
                bd[a] = sw ? vb : va;
                bd[b] = sw ? va : vb;
            } else {
                const bool sw = key_less(bd[b], bi[b], bd[a], bi[a]);
                const unsigned long long va = bd[a], vb = bd[b];
                const int ia = bi[a], ib = bi[b];
                bd[a] = sw ? vb : va; bd[b] = sw ? va : vb;
                bi[a] = sw ? ib : ia; bi[b] = sw ? ia : ib;
            }
        }
        t0 = K;
    }
    for (int t = t0; t < ncand; t++) {
        const unsigned long long d2 = dist2(t);
        const int last = EXACT ? K - 1 : k - 1;
        unsigned long long wd = bd[K - 1];
        int wi = PACKED ? 0 : bi[K - 1];
        if (!EXACT) {
#pragma unroll
            for (int q = 0; q < K; q++)
                if (q == last) { wd = bd[q]; wi = bi[q]; }
        }
        const int id = PACKED ? 0 : sidx[t];
        const bool better = PACKED ? (d2 < wd) : key_less(d2, id, wd, wi);
        // smallest (squared distance) key that is not in the list: the evicted worst or the
        // rejected candidate
        const unsigned long long gone = better ? wd : d2;
        rej = gone < rej ? gone : rej;
        if (better) {
            bool lt[K];
#pragma unroll
            for (int q = 0; q < K; q++) lt[q] = PACKED ? (d2 < bd[q]) : key_less(d2, id, bd[q], bi[q]);
#pragma unroll
            for (int q = K - 1; q >= 1; q--) {
                if (EXACT || q <= last) {
                    const unsigned long long nd = lt[q - 1] ? bd[q - 1] : (lt[q] ? d2 : bd[q]);
                    bd[q] = nd;
                    if (!PACKED) {
                        const int ni = lt[q - 1] ? bi[q - 1] : (lt[q] ? id : bi[q]);
                        bi[q] = ni;
                    }
                }
            }
            if (lt[0]) { bd[0] = d2; if (!PACKED) bi[0] = id; }
        }
    }
}

// Search bound of one pixel tile (see the comment above idw_kernel): histogram of the vectors'
// distances to the tile centre, bin of the k-th, candidate bins [0, bmax]; leaves the exclusive
// bin offsets in `hist`, zeroes `fill`, caches the bins of the first IDW_CHUNK vectors in `sbin`.
struct TileBound {
    int bmax, total;
    double cx, cy, rt, binw;
    float inv_binw_f;
    __device__ __forceinline__ int bin_of(const double2 *__restrict__ pts, int t) const {
        const double2 s = pts[t];
        const float dx = (float)(s.x - cx), dy = (float)(s.y - cy);
        const float d = sqrtf(dx * dx + dy * dy) * inv_binw_f;
        return d < (float)(IDW_BINS - 1) ? (int)d : IDW_BINS - 1;
    }
};

__device__ __forceinline__ TileBound tile_bound(const IDWParams &p, const double2 *__restrict__ pts, int npts, int k,
                                                int *hist, int *fill, unsigned char *sbin, int *s_bmax,
                                                int *s_total) {
    const int tid = threadIdx.x, lane = tid & 31;
    TileBound tb;
    // ---- tile centre, half diagonal, histogram of centre distances ---------------------
    const int j0 = blockIdx.x * IDW_TX, j1 = min(j0 + IDW_TX, p.nx) - 1;
    const int i0 = blockIdx.y * IDW_TY, i1 = min(i0 + IDW_TY, p.ny) - 1;
    const double xa = p.gx[j0], xb = p.gx[j1], ya = p.gy[i0], yb = p.gy[i1];
    tb.cx = 0.5 * (xa + xb);
    tb.cy = 0.5 * (ya + yb);
    const double cx = tb.cx, cy = tb.cy;
    // grids are monotonic (np.arange in dense_lucaskanade); the tile extent bounds the radius
    const double rt = sqrt(0.25 * (xb - xa) * (xb - xa) + 0.25 * (yb - ya) * (yb - ya));
    const double binw = fmax(rt, 1e-300) * 0.5;
    const double inv_binw = 1.0 / binw;
    for (int b = tid; b < IDW_BINS; b += IDW_THREADS) { hist[b] = 0; fill[b] = 0; }
    __syncthreads();
    // The bin only has to be CONSERVATIVE (never below the true distance bin) and the same in
    // both passes, so it is computed in float32 with 1e-5 slack instead of an FP64 sqrt.
    const float inv_binw_f = (float)inv_binw * (1.0f + 1e-5f);
    tb.inv_binw_f = inv_binw_f;
    auto bin_of = [&](int t) -> int { return tb.bin_of(pts, t); };
    for (int t = tid; t < npts; t += IDW_THREADS) {
        const int b = bin_of(t);
        if (t < IDW_CHUNK) sbin[t] = (unsigned char)b;
        atomicAdd(&hist[b], 1);
    }
    __syncthreads();
    if (tid < 32) {
        // exclusive prefix over the bins (8 per lane), bin of the k-th vector, search bound
        int c[IDW_BINS / 32], sum = 0;
#pragma unroll
        for (int q = 0; q < IDW_BINS / 32; q++) { c[q] = hist[lane * (IDW_BINS / 32) + q]; sum += c[q]; }
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        int run = incl - sum, bk = IDW_BINS;  // first bin whose inclusive count reaches k
#pragma unroll
        for (int q = 0; q < IDW_BINS / 32; q++) {
            hist[lane * (IDW_BINS / 32) + q] = run;
            run += c[q];
            if (run >= k && bk == IDW_BINS) bk = lane * (IDW_BINS / 32) + q;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) bk = min(bk, __shfl_xor_sync(0xffffffffu, bk, o));
        if (lane == 0) {
            int bmax = IDW_BINS - 1;  // overflow bin reached: every vector is a candidate
            if (bk < IDW_BINS - 1) {
                const double R = ((double)(bk + 1) * binw + 2.0 * rt) * (1.0 + 1e-9);
                const double bb = R * inv_binw * (1.0 + 1e-9);
                bmax = bb < (double)(IDW_BINS - 1) ? (int)bb : IDW_BINS - 1;
            }
            *s_bmax = bmax;
        }
    }
    __syncthreads();
    const int bmax = *s_bmax;
    if (tid == 0) *s_total = (bmax == IDW_BINS - 1) ? npts : hist[bmax + 1];
    __syncthreads();
    const int total = *s_total;

    tb.bmax = bmax;
    tb.total = total;
    tb.rt = rt;
    tb.binw = binw;
    tb.inv_binw_f = inv_binw_f;
    return tb;
}

// FASTW: the weighting of dense_lucaskanade's call (two variables, power 1/2, unit resolution,
// positive offset) from rsqrt; otherwise the general NumPy-order epilogue.
template <int K, bool EXACT, bool PACKED, bool FASTW>
__global__ void __launch_bounds__(IDW_THREADS, FASTW ? 3 : 1) idw_kernel(const IDWParams p) {
    __shared__ double2 spt[IDW_CHUNK];
    __shared__ int sidx[IDW_CHUNK];
    __shared__ int hist[IDW_BINS];   // counts, then exclusive offsets
    __shared__ int fill[IDW_BINS];
    __shared__ unsigned char sbin[IDW_CHUNK];  // bin of the first IDW_CHUNK vectors
    __shared__ int s_bmax, s_total;
    if (p.tile_done && p.tile_done[blockIdx.y * gridDim.x + blockIdx.x]) return;
    const int npts = p.npts_dev ? min(*p.npts_dev, p.npts_cap) : p.npts_cap;
    const int k = min(min(p.k, npts), K);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    // a warp covers a compact 8x4 pixel patch: its lanes agree on most insert decisions
    const int j = blockIdx.x * IDW_TX + (wid & 1) * 8 + (lane & 7);   // column
    const int i = blockIdx.y * IDW_TY + (wid >> 1) * 4 + (lane >> 3);  // row
    const double2 *__restrict__ pts = reinterpret_cast<const double2 *>(p.xy);
    const bool active = j < p.nx && i < p.ny;
    const double qx = p.gx[min(j, p.nx - 1)], qy = p.gy[min(i, p.ny - 1)];

    const TileBound tb = tile_bound(p, pts, npts, k, hist, fill, sbin, &s_bmax, &s_total);
    const int bmax = tb.bmax, total = tb.total;
    auto bin_of = [&](int t) -> int { return tb.bin_of(pts, t); };

    unsigned long long bd[K];  // squared distances as ordered bit patterns
    int bi[K];
#pragma unroll
    for (int q = 0; q < K; q++) { bd[q] = 0x7ff0000000000000ull; bi[q] = 0x7fffffff; }  // +inf
    unsigned long long rej = ~0ull;  // smallest key among the candidates that are not in the list

    // sorted: all candidates fit in shared memory (one round, counting sort by centre-distance
    // bin); otherwise plain exhaustive rounds over chunks of all vectors
    const bool sorted = total <= IDW_CHUNK;
    const int nrounds = sorted ? 1 : (npts + IDW_CHUNK - 1) / IDW_CHUNK;
    for (int r = 0; r < nrounds; r++) {
        int cnt;
        __syncthreads();
        if (sorted) {
            for (int t = tid; t < npts; t += IDW_THREADS) {
                const int b = t < IDW_CHUNK ? (int)sbin[t] : bin_of(t);
                if (b <= bmax) {
                    const int o = hist[b] + atomicAdd(&fill[b], 1);
                    spt[o] = pts[t];
                    sidx[o] = t;
                }
            }
            cnt = total;
        } else {
            const int base = r * IDW_CHUNK;
            cnt = min(IDW_CHUNK, npts - base);
            for (int t = tid; t < cnt; t += IDW_THREADS) {
                spt[t] = pts[base + t];
                sidx[t] = base + t;
            }
        }
        __syncthreads();
        if (active) topk_scan<K, EXACT, PACKED>(spt, sidx, cnt, qx, qy, k, r == 0, bd, bi, rej);
    }
    // ---- equidistant k-th / (k+1)-th neighbour: listed for the exact-order recomputation -----
    if (p.tie_list != nullptr) {
        unsigned long long kth = bd[K - 1];
        if (!EXACT) {
#pragma unroll
            for (int q = 0; q < K; q++)
                if (q == k - 1) kth = bd[q];
        }
        const unsigned long long dmask = PACKED ? ~2047ull : ~0ull;
        const bool tie = active && k >= 1 && ((kth & dmask) == (rej & dmask));
        const unsigned bal = __ballot_sync(0xffffffffu, tie);
        if (bal) {
            int base = 0;
            if (lane == __ffs(bal) - 1) base = atomicAdd(p.tie_count, __popc(bal));
            base = __shfl_sync(0xffffffffu, base, __ffs(bal) - 1);
            if (tie) p.tie_list[base + __popc(bal & ((1u << lane) - 1u))] = i * p.nx + j;
        }
    }
    if (!active || k < 1) return;
    if (FASTW) {
        // dense_lucaskanade's call: w = (sqrt(d2) + offset)^-1/2 from two rsqrt (FMA pipe) instead
        // of sqrt, pow and a divide per neighbour, normalised once -- relative error a few 1e-16
        double ws = 0.0, ax = 0.0, ay = 0.0;
        const double2 *__restrict__ v2 = reinterpret_cast<const double2 *>(p.vals);
#pragma unroll
        for (int q = 0; q < K; q++) {
            if (q < k) {
                const unsigned long long b = PACKED ? (bd[q] & ~2047ull) : bd[q];
                const int id = PACKED ? (int)(bd[q] & 2047ull) : bi[q];
                const double d2 = __longlong_as_double((long long)b);
                const double dist = d2 > 0.0 ? d2 * rsqrt(d2) : 0.0;
                const double w = rsqrt(dist + p.offset);
                const double2 v = v2[id];
                ws += w;
                ax = fma(w, v.x, ax);
                ay = fma(w, v.y, ay);
            }
        }
        const double inv = 1.0 / ws;
        p.out[((size_t)0 * p.ny + i) * p.nx + j] = ax * inv;
        p.out[((size_t)1 * p.ny + i) * p.nx + j] = ay * inv;
        return;
    }
    double w[K];
#pragma unroll
    for (int q = 0; q < K; q++) {
        if (PACKED) { bi[q] = (int)(bd[q] & 2047ull); bd[q] &= ~2047ull; }
        double d = sqrt(__longlong_as_double((long long)bd[q]));  // exact Euclidean distance
        if (p.mean_res != 1.0) d = __ddiv_rn(d, p.mean_res);  // interpolate.py:98 (x / 1.0 == x)
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

// ---- 32-bit integer keys ---------------------------------------------------------------------
// dense_lucaskanade's vectors sit on the half-pixel grid (medians of integer corners) and the
// target grid on integers, so 4 * squared distance is a small INTEGER: with coordinates doubled,
// (dx2*dx2 + dy2*dy2) in int32 is exact, and while it stays below 2^21 (neighbours within 724 px)
// the pair (distance, vector index < 2048) is ONE 32-bit key.  The whole search then runs on the
// integer pipe with single-register compares and selects -- half the instructions of the 64-bit
// list, and no FP64 until the weights.  A tile whose search radius does not fit leaves its
// `tile_done` flag clear and is filled by the 64-bit kernel launched behind this one.
constexpr int IDW_KEY32_LIMIT = 1 << 21;

template <int K>
__global__ void __launch_bounds__(IDW_THREADS, 4) idw32_kernel(const IDWParams p) {
    __shared__ int2 spi[IDW_CHUNK];   // doubled coordinates of the candidates
    __shared__ int sidx[IDW_CHUNK];
    __shared__ int hist[IDW_BINS];
    __shared__ int fill[IDW_BINS];
    __shared__ unsigned char sbin[IDW_CHUNK];
    __shared__ int s_bmax, s_total;
    const int npts = p.npts_cap;
    const int k = K;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int j = blockIdx.x * IDW_TX + (wid & 1) * 8 + (lane & 7);   // column
    const int i = blockIdx.y * IDW_TY + (wid >> 1) * 4 + (lane >> 3);  // row
    const double2 *__restrict__ pts = reinterpret_cast<const double2 *>(p.xy);
    const bool active = j < p.nx && i < p.ny;
    const double qx = p.gx[min(j, p.nx - 1)], qy = p.gy[min(i, p.ny - 1)];
    const TileBound tb = tile_bound(p, pts, npts, k, hist, fill, sbin, &s_bmax, &s_total);
    const int bmax = tb.bmax, total = tb.total;
    // every candidate lies within (bmax + 1) bins of the centre, every pixel within rt of it
    const double reach = 2.0 * ((double)(bmax + 1) * tb.binw + tb.rt) * (1.0 + 1e-6);
    if (total > IDW_CHUNK || bmax == IDW_BINS - 1 || !(reach * reach < (double)IDW_KEY32_LIMIT)) return;
    for (int t = tid; t < npts; t += IDW_THREADS) {
        const int b = t < IDW_CHUNK ? (int)sbin[t] : tb.bin_of(pts, t);
        if (b <= bmax) {
            const int o = hist[b] + atomicAdd(&fill[b], 1);
            const double2 s = pts[t];
            spi[o] = make_int2((int)(2.0 * s.x), (int)(2.0 * s.y));   // exact: multiples of 1/2
            sidx[o] = t;
        }
    }
    __syncthreads();
    unsigned bd[K];
    unsigned rej = ~0u;
    if (active) {
        const int qx2 = (int)(2.0 * qx), qy2 = (int)(2.0 * qy);
        auto key_of = [&](int t) -> unsigned {
            const int2 s = spi[t];
            const int dx = s.x - qx2, dy = s.y - qy2;
            return ((unsigned)(dx * dx + dy * dy) << 11) | (unsigned)sidx[t];
        };
#pragma unroll
        for (int q = 0; q < K; q++) bd[q] = key_of(q);   // total >= k == K
#pragma unroll
        for (int c = 0; c < net_size<K>(); c++) {
            const int a = net_a<K>(c), b = net_b<K>(c);
            const unsigned va = bd[a], vb = bd[b];
            bd[a] = min(va, vb);
            bd[b] = max(va, vb);
        }
        for (int t = K; t < total; t++) {
            const unsigned d2 = key_of(t);
            const unsigned wd = bd[K - 1];
            const bool better = d2 < wd;
            const unsigned gone = better ? wd : d2;
            rej = min(rej, gone);
            if (better) {
#pragma unroll
                for (int q = K - 1; q >= 1; q--) bd[q] = d2 < bd[q - 1] ? bd[q - 1] : (d2 < bd[q] ? d2 : bd[q]);
                if (d2 < bd[0]) bd[0] = d2;
            }
        }
    }
    if (p.tie_list != nullptr) {
        const bool tie = active && ((bd[K - 1] >> 11) == (rej >> 11));
        const unsigned bal = __ballot_sync(0xffffffffu, tie);
        if (bal) {
            int base = 0;
            if (lane == __ffs(bal) - 1) base = atomicAdd(p.tie_count, __popc(bal));
            base = __shfl_sync(0xffffffffu, base, __ffs(bal) - 1);
            if (tie) p.tie_list[base + __popc(bal & ((1u << lane) - 1u))] = i * p.nx + j;
        }
    }
    if (tid == 0) p.tile_done[blockIdx.y * gridDim.x + blockIdx.x] = 1;
    if (!active) return;
    double ws = 0.0, ax = 0.0, ay = 0.0;
    const double2 *__restrict__ v2 = reinterpret_cast<const double2 *>(p.vals);
#pragma unroll
    for (int q = 0; q < K; q++) {
        const double d2 = (double)(bd[q] >> 11) * 0.25;
        const double dist = d2 > 0.0 ? d2 * rsqrt(d2) : 0.0;
        const double w = rsqrt(dist + p.offset);
        const double2 v = v2[bd[q] & 2047u];
        ws += w;
        ax = fma(w, v.x, ax);
        ay = fma(w, v.y, ay);
    }
    const double inv = 1.0 / ws;
    p.out[((size_t)0 * p.ny + i) * p.nx + j] = ax * inv;
    p.out[((size_t)1 * p.ny + i) * p.nx + j] = ay * inv;
}

// ---- k = None: every vector weighs in (interpolate.py:82-88, scipy cdist) ------------------------
// One thread per grid point, the vectors staged through shared memory in chunks; weights and sums in
// float64 in index order (NumPy sums the same terms pairwise: relative differences ~1e-14).
constexpr int IDW_ALL_MAXVAR = 8;

__global__ void __launch_bounds__(256)
idw_all_kernel(const IDWParams p) {
    __shared__ double2 spt[1024];
    const int npts = p.npts_dev ? min(*p.npts_dev, p.npts_cap) : p.npts_cap;
    const int j = blockIdx.x * 32 + (threadIdx.x & 31), i = blockIdx.y * 8 + (threadIdx.x >> 5);
    const bool active = j < p.nx && i < p.ny;
    const double qx = p.gx[min(j, p.nx - 1)], qy = p.gy[min(i, p.ny - 1)];
    const double2 *__restrict__ pts = reinterpret_cast<const double2 *>(p.xy);
    double ws = 0.0, acc[IDW_ALL_MAXVAR];
#pragma unroll
    for (int v = 0; v < IDW_ALL_MAXVAR; v++) acc[v] = 0.0;
    for (int base = 0; base < npts; base += 1024) {
        const int cnt = min(1024, npts - base);
        __syncthreads();
        for (int t = threadIdx.x; t < cnt; t += 256) spt[t] = pts[base + t];
        __syncthreads();
        if (!active) continue;
        for (int t = 0; t < cnt; t++) {
            const double dx = __dsub_rn(spt[t].x, qx), dy = __dsub_rn(spt[t].y, qy);
            double d = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
            if (p.mean_res != 1.0) d = __ddiv_rn(d, p.mean_res);
            d = __dadd_rn(d, p.offset);
            const double w = __ddiv_rn(1.0, (p.power == 0.5) ? sqrt(d) : pow(d, p.power));
            ws += w;
            for (int v = 0; v < p.nvar; v++) acc[v] += p.vals[(size_t)(base + t) * p.nvar + v] * w;
        }
    }
    if (!active) return;
    for (int v = 0; v < p.nvar; v++) p.out[((size_t)v * p.ny + i) * p.nx + j] = acc[v] / ws;
}

// library-internal side stream of the calling thread's device + fork/join events: the tree build
// (one CTA, latency bound) overlaps the exhaustive fill, the recomputation of the listed grid
// points waits for both
struct SideStream {
    cudaStream_t s = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
    int dev = -1;
};

int side_stream(SideStream **out) {
    static thread_local SideStream per_dev[16];
    int dev = 0;
    B200_CUDA(cudaGetDevice(&dev));
    B200_REQUIRE(dev >= 0 && dev < 16, "device index out of range");
    SideStream &ss = per_dev[dev];
    if (ss.s == nullptr) {
        B200_CUDA(cudaStreamCreateWithFlags(&ss.s, cudaStreamNonBlocking));
        B200_CUDA(cudaEventCreateWithFlags(&ss.fork, cudaEventDisableTiming));
        B200_CUDA(cudaEventCreateWithFlags(&ss.join, cudaEventDisableTiming));
        ss.dev = dev;
    }
    *out = &ss;
    return 0;
}

}  // namespace

extern "C" int b200_idw_fill(const double *xy, const double *vals, const int *npts_dev, int npts_cap,
                             int nvar, int k, double power, double dist_offset, double mean_res,
                             const double *xgrid, int nx, const double *ygrid, int ny,
                             int coords_on_16th_grid, double *out, void *stream) {
    B200_REQUIRE(xy && vals && xgrid && ygrid && out && npts_cap >= 1 && nvar >= 1 && nx >= 1 && ny >= 1 &&
                     k >= 1, "bad arguments");
    if (k > 32) {
        b200::set_error("idw: k must be <= 32 (k=None / larger k is not implemented)");
        return B200_ENOTSUP;
    }
    cudaStream_t s = (cudaStream_t)stream;
    // ---- cKDTree of the vectors on the side stream (see the header comment) -----------------
    kdp::TreeScratch ts;
    b200::Scratch tie;
    SideStream *ss = nullptr;
    if (int rc = side_stream(&ss)) return rc;
    if (int rc = kdp::tree_alloc(ts, npts_cap, s)) return rc;
    const size_t N = (size_t)ny * nx;
    B200_REQUIRE(N < ((size_t)1 << 31), "grid too large");
    B200_CUDA(tie.alloc(sizeof(int) * (N + 1), s));
    int *tie_count = (int *)tie.p, *tie_list = tie_count + 1;
    B200_CUDA(cudaMemsetAsync(tie_count, 0, sizeof(int), s));
    B200_CUDA(cudaEventRecord(ss->fork, s));
    B200_CUDA(cudaStreamWaitEvent(ss->s, ss->fork, 0));
    if (int rc = kdp::tree_build(xy, npts_dev, npts_cap, ts.tb, ss->s)) return rc;
    B200_CUDA(cudaEventRecord(ss->join, ss->s));

    IDWParams p;
    p.xy = xy; p.vals = vals; p.npts_dev = npts_dev; p.npts_cap = npts_cap; p.nvar = nvar; p.k = k;
    p.gx = xgrid; p.gy = ygrid; p.nx = nx; p.ny = ny;
    p.power = power; p.offset = dist_offset; p.mean_res = mean_res; p.out = out;
    p.tie_list = tie_list; p.tie_count = tie_count;
    p.tile_done = nullptr;
    dim3 grid(b200::ceil_div(nx, IDW_TX), b200::ceil_div(ny, IDW_TY));
    dim3 block(IDW_THREADS);
    // the host knows npts only as a capacity when npts_dev is given; EXACT needs k == K <= npts
    const bool exact_ok = (npts_dev == nullptr) && npts_cap >= k;
    // the caller vouches that every coordinate (vectors and grid) is a multiple of 1/16 with
    // magnitude < 2^14; with <= 2048 vectors the index fits the zero low bits of the distance
    const bool packed = coords_on_16th_grid != 0 && exact_ok && npts_cap <= IDW_CHUNK;
    const bool fastw = nvar == 2 && power == 0.5 && mean_res == 1.0 && dist_offset > 0.0;
    b200::Scratch done;
    if (k == 20 && packed && fastw && coords_on_16th_grid == 2) {
        // vectors on the half-pixel grid, grid on integers: 32-bit integer keys wherever the search
        // radius allows, the 64-bit kernel behind it for the tiles it left
        const size_t ntiles = (size_t)grid.x * grid.y;
        B200_CUDA(done.alloc(ntiles, s));
        B200_CUDA(cudaMemsetAsync(done.p, 0, ntiles, s));
        p.tile_done = (uint8_t *)done.p;
        idw32_kernel<20><<<grid, block, 0, s>>>(p);
        B200_LAUNCH_CHECK();
    }
    if (k == 20 && packed && fastw) idw_kernel<20, true, true, true><<<grid, block, 0, s>>>(p);
    else if (k == 20 && packed) idw_kernel<20, true, true, false><<<grid, block, 0, s>>>(p);
    else if (k == 20 && exact_ok && fastw) idw_kernel<20, true, false, true><<<grid, block, 0, s>>>(p);
    else if (k == 20 && exact_ok) idw_kernel<20, true, false, false><<<grid, block, 0, s>>>(p);
    else if (k == 8 && exact_ok) idw_kernel<8, true, false, false><<<grid, block, 0, s>>>(p);
    else idw_kernel<32, false, false, false><<<grid, block, 0, s>>>(p);
    B200_LAUNCH_CHECK();
    B200_CUDA(cudaStreamWaitEvent(s, ss->join, 0));
    return kdp::idw_fix(xy, vals, nvar, k, power, dist_offset, mean_res, xgrid, nx, ygrid, ny, ts.tb, tie_list,
                        tie_count, out, s);
}

extern "C" int b200_idw_fill_all(const double *xy, const double *vals, const int *npts_dev, int npts_cap, int nvar,
                                 double power, double dist_offset, double mean_res, const double *xgrid, int nx,
                                 const double *ygrid, int ny, double *out, void *stream) {
    B200_REQUIRE(xy && vals && xgrid && ygrid && out && npts_cap >= 1 && nx >= 1 && ny >= 1, "bad arguments");
    B200_REQUIRE(nvar >= 1 && nvar <= IDW_ALL_MAXVAR, "at most 8 variables");
    IDWParams p;
    memset(&p, 0, sizeof(p));
    p.xy = xy; p.vals = vals; p.npts_dev = npts_dev; p.npts_cap = npts_cap; p.nvar = nvar; p.k = npts_cap;
    p.gx = xgrid; p.gy = ygrid; p.nx = nx; p.ny = ny;
    p.power = power; p.offset = dist_offset; p.mean_res = mean_res; p.out = out;
    idw_all_kernel<<<dim3(b200::ceil_div(nx, 32), b200::ceil_div(ny, 8)), 256, 0, (cudaStream_t)stream>>>(p);
    B200_LAUNCH_CHECK();
    return 0;
}
