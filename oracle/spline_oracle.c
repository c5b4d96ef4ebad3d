/*
 * oracle/spline_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of scipy.ndimage.map_coordinates for the spline orders the
 * semi-Lagrangian extrapolator can be asked for besides 1
 * (pysteps/extrapolation/semilagrangian.py:146-157,224-253: interp_order 0 and 3; the
 * only non-default user in the reference tree is
 * examples/ens_kalman_filter_blended_forecast.py:258 with interp_order=3, mode "nearest").
 * scipy (1.18.1 here, unpinned by the reference) is a third-party binary whose source is not
 * under /root/reference; the algorithm restated is the published one of its ndimage module:
 *   - spline_filter (prefilter=True): per axis (axis 0, then axis 1), every line is multiplied
 *     by the gain (1-z)(1-1/z), then run through the causal / anti-causal first-order
 *     recursions of pole z; z is the double nearest to sqrt(3)-2 (a decimal literal in scipy,
 *     NOT sqrt(3.0)-2.0 evaluated in double, which is 2 ulp away); boundary initialisation
 *     "mirror" for mode="constant", and for mode="nearest" the input is first edge-padded by
 *     12 samples and filtered with the "reflect" initialisation;
 *   - sampling: coordinate c (+12 when padded); first tap floor(c)-1 (order 3) or floor(c+0.5)
 *     (order 0); mode constant: !(0 <= c <= L-1) -> cval, taps outside [0, L) mirrored
 *     (period 2L-2); mode nearest: the coordinate is not clamped, every tap index is clamped
 *     to [0, L-1] (x86-64 float->int overflow of floor(c) -> INT64_MIN -> index 0, so +inf and
 *     1e300 read the LOW edge); cubic B-spline weights
 *     w1 = (y*y*(y-2)*3+4)/6, w2 = (z*z*(z-2)*3+4)/6, w0 = z*z*z/6, w3 = 1-w0-w1-w2 with
 *     y = c-floor(c), z = 1-y; value = sum over taps (rows outer, columns inner) of
 *     ((a*wy)*wx), accumulated from 0.0.
 * Parity status: PINNED bit for bit against the scipy binary (tests/test_oracle_spline.py:
 * filters and samples, both modes, degenerate shapes, border and integer coordinates) and
 * against outputs of the reference extrapolator (tests/golden/sl_golden.npz, order-3 cases).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORA_MODE_CONSTANT 0
#define ORA_MODE_NEAREST 1
#define ORA_NPAD 12

/* poles of the B-spline prefilters: the doubles nearest to the exact values (decimal literals in
 * scipy's ni_splines.c); orders 2..5 have order/2 poles */
static const double POLES[6][2] = {
    {0.0, 0.0}, {0.0, 0.0},
    {-0.171572875253809902396622551580603843, 0.0},                                      /* sqrt(8) - 3 */
    {-0.267949192431122706472553658494127633, 0.0},                                      /* sqrt(3) - 2 */
    {-0.361341225900220177092212841325675255, -0.013725429297339121360331226939128204},
    {-0.430575347099973791851434783493520110, -0.043096288203264653822712376822550182},
};

/* one pole of one line of n samples with stride s, in place (the gain has been applied) */
static void filter_pole(double *c, int64_t n, int64_t s, double z, int reflect)
{
    if (!reflect) {
        double z_i = z;
        const double z_n_1 = pow(z, (double)(n - 1));
        c[0] = c[0] + z_n_1 * c[(n - 1) * s];
        for (int64_t i = 1; i < n - 1; i++) {
            c[0] += z_i * (c[i * s] + z_n_1 * c[(n - 1 - i) * s]);
            z_i *= z;
        }
        c[0] /= 1 - z_n_1 * z_n_1;
    } else {
        double z_i = z;
        const double z_n = pow(z, (double)n);
        const double c0 = c[0];
        c[0] = c[0] + z_n * c[(n - 1) * s];
        for (int64_t i = 1; i < n; i++) {
            c[0] += z_i * (c[i * s] + z_n * c[(n - 1 - i) * s]);
            z_i *= z;
        }
        c[0] *= z / (1 - z_n * z_n);
        c[0] += c0;
    }
    for (int64_t i = 1; i < n; i++) c[i * s] += z * c[(i - 1) * s];
    if (!reflect)
        c[(n - 1) * s] = (z * c[(n - 2) * s] + c[(n - 1) * s]) * z / (z * z - 1);
    else
        c[(n - 1) * s] *= z / (z - 1);
    for (int64_t i = n - 2; i >= 0; i--) c[i * s] = z * (c[(i + 1) * s] - c[i * s]);
}

/* one line: gain of all poles first, then pole after pole */
static void filter_line(double *c, int64_t n, int64_t s, int order, int reflect)
{
    if (n <= 1) return;
    const int npoles = order / 2;
    double gain = 1.0;
    for (int k = 0; k < npoles; k++) {
        const double z = POLES[order][k];
        gain *= (1.0 - z) * (1.0 - 1.0 / z);
    }
    for (int64_t i = 0; i < n; i++) c[i * s] *= gain;
    for (int k = 0; k < npoles; k++) filter_pole(c, n, s, POLES[order][k], reflect);
}

/* scipy.ndimage.spline_filter(a, order, mode="mirror"|"reflect") of an (m, n) array, in place */
void ora_spline_filter(double *a, int64_t m, int64_t n, int order, int reflect)
{
    if (order < 2) return;
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < n; j++) filter_line(a + j, m, n, order, reflect);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < m; i++) filter_line(a + i * n, n, 1, order, reflect);
}

void ora_spline_filter3(double *a, int64_t m, int64_t n, int reflect)
{
    ora_spline_filter(a, m, n, 3, reflect);
}

static inline int64_t mirror_index(int64_t idx, int64_t len)
{
    if (len <= 1) return 0;
    const int64_t s2 = 2 * len - 2;
    if (idx < 0) {
        idx = s2 * (-idx / s2) + idx;
        idx = idx <= 1 - len ? idx + s2 : -idx;
    } else if (idx >= len) {
        idx -= s2 * (idx / s2);
        if (idx >= len) idx = s2 - idx;
    }
    return idx;
}

/* x86-64 cvttsd2si semantics of (npy_intp)floor(c) */
static inline int64_t cast_floor(double f)
{
    if (!(f >= -9223372036854775808.0 && f < 9223372036854775808.0)) return INT64_MIN;
    return (int64_t)f;
}

/* index of tap base+off: mirrored in mode constant (the coordinate is inside the array there),
 * clamped in mode nearest; index arithmetic wraps like the compiled scipy */
static inline int64_t tap(int64_t base, int64_t off, int64_t len, int mode)
{
    if (mode == ORA_MODE_CONSTANT) return mirror_index(base + off, len);
    /* two's-complement wrap of start = floor(c) - order/2 and start + l, as the compiled scipy
     * does for an overflowed base (INT64_MIN - 1 == INT64_MAX -> the HIGH edge for that tap) */
    const int64_t i = (int64_t)((uint64_t)base + (uint64_t)off);
    return i < 0 ? 0 : (i >= len ? len - 1 : i);
}

/* get_spline_interpolation_weights of scipy's ni_splines.c: x becomes the offset from the middle
 * knot (odd orders: c - floor(c), even orders: c - floor(c + 0.5)); the last weight is one minus
 * the others */
static inline void spline_weights(double x, int order, double *w)
{
    double y, z, t;
    if (order & 1) x -= floor(x); else x -= floor(x + 0.5);
    y = x;
    z = 1.0 - x;
    switch (order) {
    case 1:
        w[0] = 1.0 - x;
        break;
    case 2:
        w[1] = 0.75 - x * x;
        y = 0.5 - x;
        w[0] = 0.5 * y * y;
        break;
    case 3:
        w[1] = (y * y * (y - 2.0) * 3.0 + 4.0) / 6.0;
        w[2] = (z * z * (z - 2.0) * 3.0 + 4.0) / 6.0;
        w[0] = z * z * z / 6.0;
        break;
    case 4:
        t = x * x;
        w[2] = t * (t * 0.25 - 0.625) + 115.0 / 192.0;
        y = 1.0 + x;
        w[1] = y * (y * (y * (5.0 - y) / 6.0 - 1.25) + 5.0 / 24.0) + 55.0 / 96.0;
        w[3] = z * (z * (z * (5.0 - z) / 6.0 - 1.25) + 5.0 / 24.0) + 55.0 / 96.0;
        t = 0.5 - x;
        t *= t;
        w[0] = t * t / 24.0;
        break;
    case 5:
        t = y * y;
        w[2] = t * (t * (0.25 - y / 12.0) - 0.5) + 0.55;
        t = z * z;
        w[3] = t * (t * (0.25 - z / 12.0) - 0.5) + 0.55;
        y += 1.0;
        w[1] = y * (y * (y * (y * (y / 24.0 - 0.375) + 1.25) - 1.75) + 0.625) + 0.425;
        z += 1.0;
        w[4] = z * (z * (z * (z * (z / 24.0 - 0.375) + 1.25) - 1.75) + 0.625) + 0.425;
        z = 1.0 - x;
        t = z * z;
        w[0] = z * t * t / 120.0;
        break;
    }
    w[order] = 1.0;
    for (int i = 0; i < order; i++) w[order] -= w[i];
}

/* one sample of order 0 or 2..5 from the (already filtered and, for mode nearest, padded) array f
 * of shape (M, N); cy/cx are coordinates in the ORIGINAL frame, npad the padding of f */
double ora_sample_spline(const double *f, int64_t M, int64_t N, double cy, double cx, int order,
                         int mode, double cval, int64_t npad)
{
    cy += (double)npad;
    cx += (double)npad;
    if (mode == ORA_MODE_CONSTANT) {
        if (!(cy >= 0.0 && cy <= (double)(M - 1) && cx >= 0.0 && cx <= (double)(N - 1)))
            return cval;
    }
    /* mode nearest: the coordinate is NOT clamped; every tap index is (as for order 1) */
    if (order == 0) {
        const int64_t iy = tap(cast_floor(floor(cy + 0.5)), 0, M, mode);
        const int64_t ix = tap(cast_floor(floor(cx + 0.5)), 0, N, mode);
        double t = 0.0;
        t += f[iy * N + ix];
        return t;
    }
    const int64_t by = cast_floor((order & 1) ? floor(cy) : floor(cy + 0.5));
    const int64_t bx = cast_floor((order & 1) ? floor(cx) : floor(cx + 0.5));
    int64_t ys[6], xs[6];
    for (int l = 0; l <= order; l++) {
        ys[l] = tap(by, l - order / 2, M, mode);
        xs[l] = tap(bx, l - order / 2, N, mode);
    }
    double wy[6], wx[6];
    spline_weights(cy, order, wy);
    spline_weights(cx, order, wx);
    double t = 0.0;
    for (int j = 0; j <= order; j++)
        for (int k = 0; k <= order; k++) {
            double c = f[ys[j] * N + xs[k]];
            c *= wy[j];
            c *= wx[k];
            t += c;
        }
    return t;
}

/* Prepared input of map_coordinates(a, ..., order, mode, prefilter=True): returns a malloc'ed
 * (M, N) array (filtered for order 3; edge-padded by 12 for order 3 + mode nearest) */
double *ora_spline_prepare(const double *a, int64_t m, int64_t n, int order, int mode,
                           int64_t *M, int64_t *N, int64_t *npad)
{
    const int pad = (order > 1 && mode == ORA_MODE_NEAREST) ? ORA_NPAD : 0;
    *npad = pad;
    *M = m + 2 * pad;
    *N = n + 2 * pad;
    double *f = (double *)malloc(sizeof(double) * (size_t)(*M) * (size_t)(*N));
    if (!f) return NULL;
    for (int64_t i = 0; i < *M; i++) {
        int64_t si = i - pad;
        si = si < 0 ? 0 : (si >= m ? m - 1 : si);
        for (int64_t j = 0; j < *N; j++) {
            int64_t sj = j - pad;
            sj = sj < 0 ? 0 : (sj >= n ? n - 1 : sj);
            f[i * (*N) + j] = a[si * n + sj];
        }
    }
    if (order > 1) ora_spline_filter(f, *M, *N, order, mode == ORA_MODE_NEAREST);
    return f;
}

/* scipy.ndimage.map_coordinates(a, [cy, cx], order in {0, 2, 3, 4, 5}, mode, cval, prefilter=True) */
int ora_map_coordinates_spline(const double *a, int64_t m, int64_t n, const double *cy,
                               const double *cx, int64_t npts, int order, int mode, double cval,
                               double *out)
{
    int64_t M, N, npad;
    double *f = ora_spline_prepare(a, m, n, order, mode, &M, &N, &npad);
    if (!f) return -1;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < npts; i++)
        out[i] = ora_sample_spline(f, M, N, cy[i], cx[i], order, mode, cval, npad);
    free(f);
    return 0;
}
