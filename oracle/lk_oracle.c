/*
 * oracle/lk_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C) of the OpenCV arithmetic behind pysteps' sparse
 * Lucas-Kanade tracking, used only as the parity checker of the CUDA path.
 * pysteps calls cv2.calcOpticalFlowPyrLK at pysteps/tracking/lucaskanade.py:171
 * (opencv-python is an UNPINNED third-party dependency, requirements.txt:2;
 * the binary present here and on the GPU box is 4.13.0, baseline SSE3, whose
 * source is not under /root/reference).  The restatement follows OpenCV's
 * published algorithm (modules/video/src/lkpyramid.cpp, modules/imgproc/src/
 * pyramids.cpp) and was pinned against that binary: pyramid and Scharr images
 * bit-exact, tracker output bit-exact including the float32 accumulation ORDER
 * of the 128-bit SIMD build (4 lanes over x mod 4 plus a scalar tail), see
 * tests/test_oracle_lk.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline int reflect101(int i, int L)
{
    if (L == 1) return 0;
    while (i < 0 || i >= L) {
        if (i < 0) i = -i;
        if (i >= L) i = 2 * L - 2 - i;
    }
    return i;
}

/* cv::pyrDown on CV_8U, BORDER_REFLECT_101: separable [1 4 6 4 1], out =
 * (sum + 128) >> 8, output size ((w+1)/2, (h+1)/2). */
void ora_pyrdown_u8(const uint8_t *src, int h, int w, uint8_t *dst)
{
    const int dh = (h + 1) / 2, dw = (w + 1) / 2;
    int *row = (int *)malloc(sizeof(int) * (size_t)dw * 5);
    for (int y = 0; y < dh; y++) {
        for (int k = 0; k < 5; k++) {
            const uint8_t *s = src + (size_t)reflect101(2 * y + k - 2, h) * w;
            int *r = row + (size_t)k * dw;
            for (int x = 0; x < dw; x++) {
                int x0 = reflect101(2 * x - 2, w), x1 = reflect101(2 * x - 1, w),
                    x2 = reflect101(2 * x, w), x3 = reflect101(2 * x + 1, w),
                    x4 = reflect101(2 * x + 2, w);
                r[x] = s[x0] + 4 * s[x1] + 6 * s[x2] + 4 * s[x3] + s[x4];
            }
        }
        for (int x = 0; x < dw; x++) {
            int v = row[x] + 4 * row[dw + x] + 6 * row[2 * dw + x] + 4 * row[3 * dw + x] +
                    row[4 * dw + x];
            dst[(size_t)y * dw + x] = (uint8_t)((v + 128) >> 8);
        }
    }
    free(row);
}

/* calcScharrDeriv (lkpyramid.cpp): 3x3 Scharr, smooth [3 10 3], diff [-1 0 1],
 * BORDER_REFLECT_101, int16 interleaved (Ix, Iy). */
void ora_scharr_i16(const uint8_t *src, int h, int w, int16_t *dst)
{
    for (int y = 0; y < h; y++) {
        const uint8_t *r0 = src + (size_t)reflect101(y - 1, h) * w;
        const uint8_t *r1 = src + (size_t)y * w;
        const uint8_t *r2 = src + (size_t)reflect101(y + 1, h) * w;
        for (int x = 0; x < w; x++) {
            int xm = reflect101(x - 1, w), xp = reflect101(x + 1, w);
            /* vertical smooth / vertical diff per column, then horizontal */
            int t0m = (r0[xm] + r2[xm]) * 3 + r1[xm] * 10;
            int t0p = (r0[xp] + r2[xp]) * 3 + r1[xp] * 10;
            int t1m = r2[xm] - r0[xm], t1c = r2[x] - r0[x], t1p = r2[xp] - r0[xp];
            dst[((size_t)y * w + x) * 2 + 0] = (int16_t)(t0p - t0m);
            dst[((size_t)y * w + x) * 2 + 1] = (int16_t)((t1m + t1p) * 3 + t1c * 10);
        }
    }
}

static inline int cv_round_f(float v) { return (int)lrintf(v); } /* cvRound: nearest even */
static inline int cv_floor_f(float v)
{
    int i = (int)v;
    return i - (v < (float)i);
}
static inline int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

/* padded pixel access: image border = REFLECT_101 (pyramid padding), derivative
 * border = 0 (BORDER_CONSTANT), as built by buildOpticalFlowPyramid /
 * calcOpticalFlowPyrLK */
static inline int pix(const uint8_t *img, int h, int w, int y, int x)
{
    return img[(size_t)reflect101(y, h) * w + reflect101(x, w)];
}
static inline int der(const int16_t *d, int h, int w, int y, int x, int c)
{
    if (y < 0 || y >= h || x < 0 || x >= w) return 0;
    return d[((size_t)y * w + x) * 2 + c];
}

/* float32 accumulation in the order of OpenCV's 128-bit SIMD loop: lanes over
 * x mod 4 for x < 8*floor(win/8), scalar tail after; final
 * tail + ((q0 + q2) + (q1 + q3)). */
typedef struct { float q[4]; float tail; } acc4;
static inline void acc4_zero(acc4 *a) { a->q[0] = a->q[1] = a->q[2] = a->q[3] = 0.f; a->tail = 0.f; }
static inline float acc4_sum(const acc4 *a)
{
    float s = (a->q[0] + a->q[2]) + (a->q[1] + a->q[3]);
    return a->tail + s;
}

/*
 * One pyramid level of LKTrackerInvoker for npts points.
 *   I, J      level images (h, w) uint8;  dI  Scharr of I (h, w, 2) int16
 *   prev_pts  (npts,2) float32 level-0 coordinates
 *   next_pts  (npts,2) float32 in/out (as OpenCV carries them between levels)
 *   status    (npts) uint8 in/out ; err (npts) float32 in/out
 */
void ora_lk_level(const uint8_t *I, const uint8_t *J, const int16_t *dI, int h, int w,
                  const float *prev_pts, float *next_pts, uint8_t *status, float *err,
                  int npts, int win_w, int win_h, int level, int max_level, int max_count,
                  double epsilon, double min_eig_thr)
{
    const float half_x = (win_w - 1) * 0.5f, half_y = (win_h - 1) * 0.5f;
    const int W_BITS = 14;
    const float FLT_SCALE = 1.f / (1 << 20);
    const int simd_w = (win_w / 8) * 8;
    int16_t *Iw = (int16_t *)malloc(sizeof(int16_t) * (size_t)win_w * win_h);
    int16_t *dIw = (int16_t *)malloc(sizeof(int16_t) * (size_t)win_w * win_h * 2);
    for (int p = 0; p < npts; p++) {
        float px = prev_pts[2 * p] * (float)(1. / (1 << level));
        float py = prev_pts[2 * p + 1] * (float)(1. / (1 << level));
        float nx, ny;
        if (level == max_level) { nx = px; ny = py; }
        else { nx = next_pts[2 * p] * 2.f; ny = next_pts[2 * p + 1] * 2.f; }
        next_pts[2 * p] = nx; next_pts[2 * p + 1] = ny;
        px -= half_x; py -= half_y;
        int ix = cv_floor_f(px), iy = cv_floor_f(py);
        if (ix < -win_w || ix >= w || iy < -win_h || iy >= h) {
            if (level == 0) { status[p] = 0; err[p] = 0; }
            continue;
        }
        float a = px - ix, b = py - iy;
        int iw00 = cv_round_f((1.f - a) * (1.f - b) * (1 << W_BITS));
        int iw01 = cv_round_f(a * (1.f - b) * (1 << W_BITS));
        int iw10 = cv_round_f((1.f - a) * b * (1 << W_BITS));
        int iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
        acc4 A11, A12, A22;
        acc4_zero(&A11); acc4_zero(&A12); acc4_zero(&A22);
        for (int y = 0; y < win_h; y++)
            for (int x = 0; x < win_w; x++) {
                int yy = iy + y, xx = ix + x;
                int ival = descale(pix(I, h, w, yy, xx) * iw00 + pix(I, h, w, yy, xx + 1) * iw01 +
                                   pix(I, h, w, yy + 1, xx) * iw10 + pix(I, h, w, yy + 1, xx + 1) * iw11,
                                   W_BITS - 5);
                int ixv = descale(der(dI, h, w, yy, xx, 0) * iw00 + der(dI, h, w, yy, xx + 1, 0) * iw01 +
                                  der(dI, h, w, yy + 1, xx, 0) * iw10 + der(dI, h, w, yy + 1, xx + 1, 0) * iw11,
                                  W_BITS);
                int iyv = descale(der(dI, h, w, yy, xx, 1) * iw00 + der(dI, h, w, yy, xx + 1, 1) * iw01 +
                                  der(dI, h, w, yy + 1, xx, 1) * iw10 + der(dI, h, w, yy + 1, xx + 1, 1) * iw11,
                                  W_BITS);
                Iw[y * win_w + x] = (int16_t)ival;
                dIw[(y * win_w + x) * 2] = (int16_t)ixv;
                dIw[(y * win_w + x) * 2 + 1] = (int16_t)iyv;
                float fx = (float)ixv, fy = (float)iyv;
                if (x < simd_w) {
                    int l = x & 3;
                    A22.q[l] = fy * fy + A22.q[l];
                    A12.q[l] = fx * fy + A12.q[l];
                    A11.q[l] = fx * fx + A11.q[l];
                } else {
                    A11.tail += (float)(ixv * ixv);
                    A12.tail += (float)(ixv * iyv);
                    A22.tail += (float)(iyv * iyv);
                }
            }
        float fA11 = acc4_sum(&A11) * FLT_SCALE, fA12 = acc4_sum(&A12) * FLT_SCALE,
              fA22 = acc4_sum(&A22) * FLT_SCALE;
        float D = fA11 * fA22 - fA12 * fA12;
        float minEig = (fA22 + fA11 - sqrtf((fA11 - fA22) * (fA11 - fA22) + 4.f * fA12 * fA12)) /
                       (2 * win_w * win_h);
        if (minEig < min_eig_thr || D < 1.1920928955078125e-07f) {
            if (level == 0) status[p] = 0;
            continue;
        }
        D = 1.f / D;
        nx -= half_x; ny -= half_y;
        float pdx = 0.f, pdy = 0.f;
        for (int j = 0; j < max_count; j++) {
            int jx = cv_floor_f(nx), jy = cv_floor_f(ny);
            if (jx < -win_w || jx >= w || jy < -win_h || jy >= h) {
                if (level == 0) status[p] = 0;
                break;
            }
            a = nx - jx; b = ny - jy;
            iw00 = cv_round_f((1.f - a) * (1.f - b) * (1 << W_BITS));
            iw01 = cv_round_f(a * (1.f - b) * (1 << W_BITS));
            iw10 = cv_round_f((1.f - a) * b * (1 << W_BITS));
            iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
            /* SIMD lanes: qb0 = [x(0,4) y(0,4) x(1,5) y(1,5)], qb1 = [x(2,6) y(2,6) x(3,7) y(3,7)] */
            float qb0[4] = {0, 0, 0, 0}, qb1[4] = {0, 0, 0, 0}, t1 = 0.f, t2 = 0.f;
            for (int y = 0; y < win_h; y++) {
                int dv[8];
                for (int x = 0; x < win_w; x++) {
                    int yy = jy + y, xx = jx + x;
                    int diff = descale(pix(J, h, w, yy, xx) * iw00 + pix(J, h, w, yy, xx + 1) * iw01 +
                                       pix(J, h, w, yy + 1, xx) * iw10 + pix(J, h, w, yy + 1, xx + 1) * iw11,
                                       W_BITS - 5) - Iw[y * win_w + x];
                    if (x < simd_w) {
                        dv[x & 7] = diff;
                        if ((x & 7) == 7) {
                            const int16_t *g = dIw + (y * win_w + x - 7) * 2;
                            for (int q = 0; q < 2; q++) {      /* pixel pairs (q, q+4) */
                                int sx = dv[q] * g[2 * q] + dv[q + 4] * g[2 * (q + 4)];
                                int sy = dv[q] * g[2 * q + 1] + dv[q + 4] * g[2 * (q + 4) + 1];
                                qb0[2 * q] += (float)sx;
                                qb0[2 * q + 1] += (float)sy;
                            }
                            for (int q = 2; q < 4; q++) {
                                int sx = dv[q] * g[2 * q] + dv[q + 4] * g[2 * (q + 4)];
                                int sy = dv[q] * g[2 * q + 1] + dv[q + 4] * g[2 * (q + 4) + 1];
                                qb1[2 * (q - 2)] += (float)sx;
                                qb1[2 * (q - 2) + 1] += (float)sy;
                            }
                        }
                    } else {
                        t1 += (float)(diff * dIw[(y * win_w + x) * 2]);
                        t2 += (float)(diff * dIw[(y * win_w + x) * 2 + 1]);
                    }
                }
            }
            /* qf0 = [X0 X1 0 0], qf1 = [Y0 Y1 0 0] of (qb0 + qb1); reduce: (l0+l2)+(l1+l3) */
            float X0 = qb0[0] + qb1[0], Y0 = qb0[1] + qb1[1], X1 = qb0[2] + qb1[2], Y1 = qb0[3] + qb1[3];
            float ib1 = t1 + ((X0 + 0.f) + (X1 + 0.f));
            float ib2 = t2 + ((Y0 + 0.f) + (Y1 + 0.f));
            float b1 = ib1 * FLT_SCALE, b2 = ib2 * FLT_SCALE;
            float ddx = (float)((fA12 * b2 - fA22 * b1) * D);
            float ddy = (float)((fA12 * b1 - fA11 * b2) * D);
            nx += ddx; ny += ddy;
            next_pts[2 * p] = nx + half_x; next_pts[2 * p + 1] = ny + half_y;
            if ((double)ddx * ddx + (double)ddy * ddy <= epsilon) break;
            if (j > 0 && fabs(ddx + pdx) < 0.01 && fabs(ddy + pdy) < 0.01) {
                next_pts[2 * p] -= ddx * 0.5f; next_pts[2 * p + 1] -= ddy * 0.5f;
                break;
            }
            pdx = ddx; pdy = ddy;
        }
        if (status[p] && level == 0) {
            /* the error pass re-checks that the final window start is inside */
            float fx = next_pts[2 * p] - half_x, fy = next_pts[2 * p + 1] - half_y;
            int jx = cv_floor_f(fx), jy = cv_floor_f(fy);
            if (jx < -win_w || jx >= w || jy < -win_h || jy >= h) status[p] = 0;
        }
    }
    free(Iw); free(dIw);
}

/* ------------------------------------------------------------------------
 * cv::cornerMinEigenVal(src u8, blockSize=5, ksize=3, BORDER_REFLECT_101) as the
 * 4.13.0 AVX-512 build computes it (pinned bit-for-bit, tests/test_oracle_lk.py):
 *   scale s = 1/(255 * 2^(ksize-1) * blockSize) applied to the SMOOTHING taps
 *   Dx = fma(s, r[y-1] + r[y+1], (2s) * r[y])          r = horizontal [-1 0 1]
 *   v  = fma(c[x+1], s, fma(c[x], 2s, c[x-1] * s))     (x < 32*floor(W/32))
 *        (c[x-1]*s + c[x]*2s) + c[x+1]*s               (scalar tail columns)
 *   Dy = v[y+1] - v[y-1]
 *   cov = Dx*Dx, Dx*Dy, Dy*Dy (float32); 5x5 box sums accumulated in DOUBLE:
 *   row sums left to right, column sums as a running sum down the rows
 *   (add the entering row, emit, subtract the leaving row), rounded to float32;
 *   a = xx/2, b = xy, c = yy/2; eig = (a + c) - sqrt((a - c)^2 + b^2) in float32.
 * ---------------------------------------------------------------------- */
void ora_min_eig_u8(const uint8_t *src, int h, int w, float *eig)
{
    const float s = (float)(1.0 / 5100.0), s2 = (float)(2.0 / 5100.0);
    const size_t N = (size_t)h * w;
    float *cxx = (float *)malloc(sizeof(float) * N * 3);
    float *cxy = cxx + N, *cyy = cxy + N;
    float *v = (float *)malloc(sizeof(float) * (size_t)(h + 2) * w); /* rows -1..h */
    float *g = (float *)malloc(sizeof(float) * (size_t)(h + 2) * w);
    const int tail0 = (w / 32) * 32;
    for (int yy = -1; yy <= h; yy++) {
        const uint8_t *r = src + (size_t)reflect101(yy, h) * w;
        for (int x = 0; x < w; x++) {
            float c0 = r[reflect101(x - 1, w)], c1 = r[x], c2 = r[reflect101(x + 1, w)];
            g[(size_t)(yy + 1) * w + x] = c2 - c0;
            v[(size_t)(yy + 1) * w + x] =
                (x < tail0) ? fmaf(c2, s, fmaf(c1, s2, c0 * s)) : (c0 * s + c1 * s2) + c2 * s;
        }
    }
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            float dx = fmaf(s, g[(size_t)y * w + x] + g[(size_t)(y + 2) * w + x],
                            s2 * g[(size_t)(y + 1) * w + x]);
            float dy = v[(size_t)(y + 2) * w + x] - v[(size_t)y * w + x];
            cxx[(size_t)y * w + x] = dx * dx;
            cxy[(size_t)y * w + x] = dx * dy;
            cyy[(size_t)y * w + x] = dy * dy;
        }
    double *sum = (double *)malloc(sizeof(double) * 3 * w);
    double *rs = (double *)malloc(sizeof(double) * 3 * w);
    for (int ch = 0; ch < 3; ch++) {
        const float *c = cxx + (size_t)ch * N;
        double *S = sum + (size_t)ch * w;
        for (int x = 0; x < w; x++) S[x] = 0.0;
        for (int k = -2; k < 2; k++) {
            const float *r = c + (size_t)reflect101(k, h) * w;
            for (int x = 0; x < w; x++) {
                double t = (double)r[reflect101(x - 2, w)];
                for (int d = -1; d <= 2; d++) t += (double)r[reflect101(x + d, w)];
                S[x] += t;
            }
        }
    }
    float *box = (float *)malloc(sizeof(float) * 3 * w);
    for (int y = 0; y < h; y++) {
        for (int ch = 0; ch < 3; ch++) {
            const float *c = cxx + (size_t)ch * N;
            const float *rp = c + (size_t)reflect101(y + 2, h) * w;
            const float *rm = c + (size_t)reflect101(y - 2, h) * w;
            double *S = sum + (size_t)ch * w;
            for (int x = 0; x < w; x++) {
                double tp = (double)rp[reflect101(x - 2, w)], tm = (double)rm[reflect101(x - 2, w)];
                for (int d = -1; d <= 2; d++) {
                    tp += (double)rp[reflect101(x + d, w)];
                    tm += (double)rm[reflect101(x + d, w)];
                }
                double s0 = S[x] + tp;
                box[(size_t)ch * w + x] = (float)s0;
                S[x] = s0 - tm;
            }
        }
        for (int x = 0; x < w; x++) {
            float a = box[x] * 0.5f, b = box[w + x], c = box[2 * w + x] * 0.5f;
            float t = a - c;
            eig[(size_t)y * w + x] = (a + c) - sqrtf(t * t + b * b);
        }
    }
    (void)rs;
    free(rs); free(box); free(sum); free(g); free(v); free(cxx);
}

/* ------------------------------------------------------------------------
 * pysteps/utils/interpolate.py:67-114 idwinterp2d with k nearest neighbours:
 * cKDTree.query is restated as an exhaustive search (same exact Euclidean
 * distances; neighbours ascending by distance, ties by lower point index).
 *   xy (npts,2), vals (npts,nvar) float64; grid point (gx[j], gy[i]).
 *   out (nvar, ny, nx).  Weight sum: NumPy pairwise order (8 accumulators),
 *   weighted sum: sequential in k.
 * ---------------------------------------------------------------------- */
static double np_pairwise_sum(const double *a, int n)
{
    if (n < 8) {
        double r = 0.0; /* numpy starts from the first element via -0.0 identity; same value */
        r = a[0];
        for (int i = 1; i < n; i++) r += a[i];
        return r;
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

void ora_idw(const double *xy, const double *vals, int npts, int nvar, const double *gx, int nx,
             const double *gy, int ny, int k, double power, double dist_offset, double mean_res,
             double *out, uint8_t *tie)
{
    /* tie (ny,nx), optional: 1 where the k-th and (k+1)-th nearest points are exactly
     * equidistant, i.e. where cKDTree's answer depends on its traversal order */
    if (k > npts) k = npts;
#pragma omp parallel
    {
        double *bd = (double *)malloc(sizeof(double) * (k + 1));
        int *bi = (int *)malloc(sizeof(int) * (k + 1));
        double *wt = (double *)malloc(sizeof(double) * k);
#pragma omp for schedule(dynamic, 4)
        for (int i = 0; i < ny; i++)
            for (int j = 0; j < nx; j++) {
                int cnt = 0;
                double next_best = INFINITY; /* smallest distance among the points left out */
                for (int p = 0; p < npts; p++) {
                    double dx = xy[2 * p] - gx[j], dy = xy[2 * p + 1] - gy[i];
                    double d2 = dx * dx + dy * dy;
                    if (cnt == k && !(d2 < bd[k - 1])) {
                        if (d2 < next_best) next_best = d2;
                        continue;
                    }
                    if (cnt == k && bd[k - 1] < next_best) next_best = bd[k - 1];
                    int q = cnt < k ? cnt : k - 1;
                    while (q > 0 && bd[q - 1] > d2) { bd[q] = bd[q - 1]; bi[q] = bi[q - 1]; q--; }
                    bd[q] = d2; bi[q] = p;
                    if (cnt < k) cnt++;
                }
                if (tie) tie[(size_t)i * nx + j] = (cnt == k && next_best == bd[k - 1]);
                for (int q = 0; q < k; q++) {
                    double d = sqrt(bd[q]);
                    d /= mean_res;
                    d += dist_offset;
                    wt[q] = 1.0 / (power == 0.5 ? sqrt(d) : pow(d, power));
                }
                double ws = np_pairwise_sum(wt, k);
                for (int v = 0; v < nvar; v++) {
                    double acc = 0.0;
                    for (int q = 0; q < k; q++) {
                        double term = vals[(size_t)bi[q] * nvar + v] * (wt[q] / ws);
                        acc = (q == 0) ? term : acc + term;
                    }
                    out[((size_t)v * ny + i) * nx + j] = acc;
                }
            }
        free(bd); free(bi); free(wt);
    }
}
