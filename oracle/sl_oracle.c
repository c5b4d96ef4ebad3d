/*
 * oracle/sl_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, float64) of the semi-Lagrangian extrapolator of
 * pysteps, used only as the parity checker for the CUDA path (tests/,
 * __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference leg).
 * Nothing under pysteps_b200/ may import, link or call this file.
 *
 * Follows:
 *   pysteps/extrapolation/semilagrangian.py:21-266   (extrapolate)
 *   pysteps/extrapolation/semilagrangian.py:181-198  (interpolate_motion)
 * and restates the third-party arithmetic it calls,
 *   scipy.ndimage.map_coordinates(order=1, prefilter=False)  [scipy 1.18.1,
 *   unpinned by the reference: requirements.txt:5], whose source is not under
 *   /root/reference.  The restatement was pinned bit-for-bit against the scipy
 *   binary (tests/golden/gen_sl_golden.py, tests/test_oracle_sl.py):
 *     - axis coordinate c, length L, i0 = floor(c), t = c - floor(c),
 *       weights w0 = 1 - t, w1 = 1 - w0  (NOT t: the last spline weight is
 *       one minus the others; differs from t in the last bit for 0 < c < 1);
 *     - value = sum over taps in the order (y0,x0),(y0,x1),(y1,x0),(y1,x1) of
 *       ((a * wy) * wx), accumulated left to right starting from 0.0;
 *     - mode "constant": !(0 <= c <= L-1) on either axis (NaN included) -> cval;
 *       otherwise the tap i0+1 == L (only when c == L-1) is mirrored to L-2
 *       (0 if L == 1) and still read, so a NaN there poisons the result even
 *       with weight 0;
 *     - mode "nearest": no coordinate clamp; every tap index is clamped to
 *       [0, L-1]; the float->int cast of floor(c) overflows to INT64_MIN on
 *       x86-64 for |c| >= 2^63 or non-finite c (both taps -> index 0).
 *
 * Parity status: PINNED (scipy binary + the reference's own known-answer
 * tests pysteps/tests/test_extrapolation_semilagrangian.py:9-24,57-72).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORA_MODE_CONSTANT 0
#define ORA_MODE_NEAREST 1

int ora_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* launchers such as torchrun export OMP_NUM_THREADS=1; the CPU timing legs of bench.py set the
 * thread count they report explicitly */
void ora_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n >= 1) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* x86-64 cvttsd2si semantics of (npy_intp)floor(c) */
static inline int64_t cast_floor(double f)
{
    if (!(f >= -9223372036854775808.0 && f < 9223372036854775808.0))
        return INT64_MIN;
    return (int64_t)f;
}

static inline int64_t tap_index(int64_t i, int64_t L, int mode)
{
    if (mode == ORA_MODE_NEAREST) {
        if (i < 0) return 0;
        if (i >= L) return L - 1;
        return i;
    }
    /* constant mode: only i == L can be out of range; scipy mirrors it */
    if (i >= L) return (L > 1) ? (2 * L - 2 - i) : 0;
    return i;
}

/* one order-1 sample; a is (m, n) row-major */
static inline double sample_o1(const double *a, int64_t m, int64_t n, double cy,
                               double cx, int mode, double cval)
{
    if (mode == ORA_MODE_CONSTANT) {
        if (!(cy >= 0.0 && cy <= (double)(m - 1) && cx >= 0.0 &&
              cx <= (double)(n - 1)))
            return cval;
    }
    double fy = floor(cy), fx = floor(cx);
    double ty = cy - fy, tx = cx - fx;
    int64_t iy = cast_floor(fy), ix = cast_floor(fx);
    int64_t ys[2], xs[2];
    /* INT64_MIN + 1 is still negative: both taps clamp to 0 */
    ys[0] = tap_index(iy, m, mode);
    ys[1] = tap_index(iy == INT64_MIN ? iy : iy + 1, m, mode);
    xs[0] = tap_index(ix, n, mode);
    xs[1] = tap_index(ix == INT64_MIN ? ix : ix + 1, n, mode);
    double wy[2], wx[2];
    wy[0] = 1.0 - ty; wy[1] = 1.0 - wy[0];
    wx[0] = 1.0 - tx; wx[1] = 1.0 - wx[0];
    double t = 0.0;
    for (int j = 0; j < 2; j++)
        for (int k = 0; k < 2; k++) {
            double c = a[ys[j] * n + xs[k]];
            c *= wy[j];
            c *= wx[k];
            t += c;
        }
    return t;
}

/* scipy.ndimage.map_coordinates(a, [cy, cx], order=1, mode, cval, prefilter=False) */
void ora_map_coordinates_o1(const double *a, int64_t m, int64_t n,
                            const double *cy, const double *cx, int64_t npts,
                            int mode, double cval, double *out)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < npts; i++)
        out[i] = sample_o1(a, m, n, cy[i], cx[i], mode, cval);
}

/* semilagrangian.py:181-198 interpolate_motion(displacement, velocity_inc, td)
 * disp is the ARGUMENT array (may be a temporary), vinc is overwritten. */
static void interpolate_motion(const double *V, const double *xy, int64_t m,
                               int64_t n, const double *disp, double *vinc,
                               double td, double vel_timestep, int n_iter,
                               int vel_f32)
{
    const int64_t N = m * n;
    const double s = td / vel_timestep; /* :198 scalar evaluated first */
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; i++) {
        /* :182-183 coords_warped = xy_coords + displacement; [row, col] order */
        double cx = xy[i] + disp[i];
        double cy = xy[N + i] + disp[N + i];
        double vx = sample_o1(V, m, n, cy, cx, ORA_MODE_NEAREST, 0.0);     /* :185-187 */
        double vy = sample_o1(V + N, m, n, cy, cx, ORA_MODE_NEAREST, 0.0); /* :188-190 */
        if (vel_f32) {
            /* float32 velocity: map_coordinates returns the INPUT dtype, so the
             * sampled increments are rounded to float32 when stored (:192-193) */
            vx = (double)(float)vx;
            vy = (double)(float)vy;
        }
        if (n_iter > 1) { /* :195-196 */
            vx /= (double)n_iter;
            vy /= (double)n_iter;
        }
        vinc[i] = vx * s; /* :198 */
        vinc[N + i] = vy * s;
    }
}

/*
 * semilagrangian.py:21-266 extrapolate(), interp_order == 1 only.
 *   precip     (m,n) or NULL
 *   V          (2,m,n)
 *   xy         (2,m,n) float64 coordinates (xy[0]=x/cols, xy[1]=y/rows) or NULL
 *              for the default meshgrid (:174-179)
 *   tdiff      T timestep differences (:165), vel_timestep (:159-161)
 *   disp_prev  (2,m,n) or NULL
 *   vel_f32    non-zero when the caller's velocity array is float32 (values
 *              are passed here widened to float64, which is exact)
 *   out        (T,m,n) or NULL when precip is NULL
 *   disp_out   (2,m,n) (always written)
 * returns 0, or -1 on allocation failure.
 */
/* spline_oracle.c */
double *ora_spline_prepare(const double *a, int64_t m, int64_t n, int order, int mode,
                           int64_t *M, int64_t *N, int64_t *npad);
double ora_sample_spline(const double *f, int64_t M, int64_t N, double cy, double cx, int order,
                         int mode, double cval, int64_t npad);

/* As ora_sl_extrapolate, plus semilagrangian.py:144-157,224-253: interp_order 0 or 3 for the
 * precipitation field (the motion field is always sampled with order 1), with the two
 * auxiliary order-1 mask warps.  For interp_order > 1 the caller passes precip with non-finite
 * values already zeroed (:150-152), mask_min = (precip > minval) and mask_finite as float64
 * 0/1 arrays (:147-155). */
int ora_sl_extrapolate_order(const double *precip, const double *V, int64_t m,
                             int64_t n, const double *xy_in, const double *tdiff,
                             int64_t T, double vel_timestep, int n_iter,
                             const double *disp_prev, double outval, int mode,
                             int vel_f32, int interp_order, const double *mask_min,
                             const double *mask_finite, double minval,
                             double *out, double *disp_out);

int ora_sl_extrapolate(const double *precip, const double *V, int64_t m,
                       int64_t n, const double *xy_in, const double *tdiff,
                       int64_t T, double vel_timestep, int n_iter,
                       const double *disp_prev, double outval, int mode,
                       int vel_f32, double *out, double *disp_out)
{
    return ora_sl_extrapolate_order(precip, V, m, n, xy_in, tdiff, T, vel_timestep, n_iter,
                                    disp_prev, outval, mode, vel_f32, 1, NULL, NULL, 0.0, out,
                                    disp_out);
}

int ora_sl_extrapolate_order(const double *precip, const double *V, int64_t m,
                             int64_t n, const double *xy_in, const double *tdiff,
                             int64_t T, double vel_timestep, int n_iter,
                             const double *disp_prev, double outval, int mode,
                             int vel_f32, int interp_order, const double *mask_min,
                             const double *mask_finite, double minval,
                             double *out, double *disp_out)
{
    const int64_t N = m * n;
    double *xy = (double *)malloc(sizeof(double) * 2 * N);
    double *vinc = (double *)malloc(sizeof(double) * 2 * N);
    double *tmp = (double *)malloc(sizeof(double) * 2 * N);
    double *disp = disp_out;
    if (!xy || !vinc || !tmp) {
        free(xy); free(vinc); free(tmp);
        return -1;
    }
    /* spline coefficients of the field: the prefilter of every map_coordinates call of the
     * leadtime loop sees the same input, so it is evaluated once */
    double *filt = NULL;
    int64_t FM = 0, FN = 0, fpad = 0;
    if (precip && interp_order != 1) {
        filt = ora_spline_prepare(precip, m, n, interp_order, mode, &FM, &FN, &fpad);
        if (!filt) {
            free(xy); free(vinc); free(tmp);
            return -1;
        }
    }
    if (xy_in) {
        memcpy(xy, xy_in, sizeof(double) * 2 * N);
    } else {
        for (int64_t y = 0; y < m; y++)
            for (int64_t x = 0; x < n; x++) {
                xy[y * n + x] = (double)x;
                xy[N + y * n + x] = (double)y;
            }
    }
    if (!disp_prev) {
        /* :201-203 displacement = 0; velocity_inc = V * tdiff[0] / vel_timestep
         * (left to right: multiply, then divide) */
        for (int64_t i = 0; i < 2 * N; i++) {
            disp[i] = 0.0;
            vinc[i] = V[i] * tdiff[0] / vel_timestep;
        }
    } else {
        /* :205-207 */
        memcpy(disp, disp_prev, sizeof(double) * 2 * N);
        interpolate_motion(V, xy, m, n, disp, vinc, tdiff[0], vel_timestep, n_iter, vel_f32);
    }
    for (int64_t ti = 0; ti < T; ti++) {
        const double td = tdiff[ti];
        if (n_iter > 0) { /* :210-214 */
            for (int k = 0; k < n_iter; k++) {
#pragma omp parallel for schedule(static)
                for (int64_t i = 0; i < 2 * N; i++)
                    tmp[i] = disp[i] - vinc[i] / 2.0;
                interpolate_motion(V, xy, m, n, tmp, vinc, td, vel_timestep, n_iter, vel_f32);
#pragma omp parallel for schedule(static)
                for (int64_t i = 0; i < 2 * N; i++)
                    disp[i] -= vinc[i];
                interpolate_motion(V, xy, m, n, disp, vinc, td, vel_timestep, n_iter, vel_f32);
            }
        } else { /* :215-219 */
            if (ti > 0 || disp_prev)
                interpolate_motion(V, xy, m, n, disp, vinc, td, vel_timestep, n_iter, vel_f32);
            for (int64_t i = 0; i < 2 * N; i++)
                disp[i] -= vinc[i];
        }
        if (precip) { /* :221-232 */
            double *o = out + ti * N;
#pragma omp parallel for schedule(static)
            for (int64_t i = 0; i < N; i++) {
                double cx = xy[i] + disp[i];
                double cy = xy[N + i] + disp[N + i];
                if (interp_order == 1) {
                    o[i] = sample_o1(precip, m, n, cy, cx, mode, outval);
                    continue;
                }
                double v = ora_sample_spline(filt, FM, FN, cy, cx, interp_order, mode, outval, fpad);
                if (interp_order > 1) { /* :234-253 */
                    if (sample_o1(mask_min, m, n, cy, cx, mode, 0.0) < 0.5) v = minval;
                    if (sample_o1(mask_finite, m, n, cy, cx, mode, 0.0) < 0.5) v = NAN;
                }
                o[i] = v;
            }
        }
    }
    free(xy); free(vinc); free(tmp); free(filt);
    return 0;
}
