/*
 * oracle/vet_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, float64, sequential) of the native part of pysteps'
 * Variational Echo Tracking:
 *   pysteps/motion/_vet.pyx:66-232   _warp
 *   pysteps/motion/_vet.pyx:238-621  _cost_function
 * and of scipy.ndimage.zoom(order=1, mode="nearest") as vet() uses it
 * (pysteps/motion/vet.py:580-589, 621-630).  Loops, index limits (including the
 * off-by-one of _vet.pyx:466-476) and expression shapes follow the reference
 * line by line.  The reference extension is compiled with -ffast-math
 * (setup.py:27-28), so it is itself not IEEE-strict: the oracle is pinned to it
 * (built out of tree from the reference's own .pyx, tests/golden/gen_vet_golden.py)
 * at relative 1e-9 on cost/gradient evaluations, not bit for bit.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* _vet.pyx:66-232.  image (nx,ny), mask int8 (nx,ny), disp (2,nx,ny);
 * out: new_image, morphed_mask, grad (2,nx,ny) if grad != NULL */
void ora_vet_warp(const double *image, const int8_t *mask, const double *disp, int64_t nx,
                  int64_t ny, double *new_image, int8_t *morphed_mask, double *grad)
{
    const int64_t xmi = nx - 1, ymi = ny - 1;
    const double xmf = (double)xmi, ymf = (double)ymi;
    const int64_t N = nx * ny;
    for (int64_t x = 0; x < nx; x++)
        for (int64_t y = 0; y < ny; y++) {
            double xf = (double)x - disp[x * ny + y];
            double yf = (double)y - disp[N + x * ny + y];
            int64_t x0, x1, y0, y1;
            if (xf < 0) { xf = 0; x0 = 0; x1 = 0; }
            else if (xf > xmf) { xf = xmf; x0 = xmi; x1 = xmi; }
            else { x0 = (int64_t)floor(xf); x1 = x0 + 1; if (x1 > xmi) x1 = xmi; }
            if (yf < 0) { yf = 0; y0 = 0; y1 = 0; }
            else if (yf > ymf) { yf = ymf; y0 = ymi; y1 = ymi; }
            else { y0 = (int64_t)floor(yf); y1 = y0 + 1; if (y1 > ymi) y1 = ymi; }
            const double dx = xf - (double)x0, dy = yf - (double)y0;
            double f00 = image[x0 * ny + y0];
            double f10 = image[x1 * ny + y0] - image[x0 * ny + y0];
            double f01 = image[x0 * ny + y1] - image[x0 * ny + y0];
            double f11 = (image[x0 * ny + y0] - image[x1 * ny + y0] - image[x0 * ny + y1] +
                          image[x1 * ny + y1]);
            new_image[x * ny + y] = f00 + dx * f10 + dy * f01 + dx * dy * f11;
            if (grad) {
                grad[x * ny + y] = f10 + dy * f11;
                grad[N + x * ny + y] = f01 + dx * f11;
            }
            f00 = mask[x0 * ny + y0];
            f10 = mask[x1 * ny + y0] - mask[x0 * ny + y0];
            f01 = mask[x0 * ny + y1] - mask[x0 * ny + y0];
            f11 = (mask[x0 * ny + y0] - mask[x1 * ny + y0] - mask[x0 * ny + y1] + mask[x1 * ny + y1]);
            /* the interpolated mask OVERWRITES the out-of-range flag set above (:219-226) */
            int8_t mm = (int8_t)(f00 + dx * f10 + dy * f01 + dx * dy * f11);
            morphed_mask[x * ny + y] = (mm != 0) ? 1 : 0; /* :228 */
        }
}

static inline int64_t imin(int64_t a, int64_t b) { return a < b ? a : b; }
static inline int64_t imax(int64_t a, int64_t b) { return a > b ? a : b; }
/* Python floor division */
static inline int64_t fdiv(int64_t a, int64_t b)
{
    int64_t q = a / b;
    if ((a % b != 0) && ((a < 0) != (b < 0))) q--;
    return q;
}

/*
 * _vet.pyx:238-621.  sector_disp (2,xs,ys), template/input (nx,ny), mask int8.
 * gradient == 0: out[0] = residuals, out[1] = smoothness_penalty.
 * gradient != 0: out = grad_residuals + grad_smooth, shape (2,xs,ys).
 * returns 0, -1 if the sectors do not divide the image, -2 on allocation failure.
 */
int ora_vet_cost(const double *sd, const double *templ, const double *input, const int8_t *mask,
                 int64_t xs, int64_t ys, int64_t nx, int64_t ny, float smooth_gain_f, int gradient,
                 double *out)
{
    if (nx % xs != 0 || ny % ys != 0) return -1;
    const int64_t xss = (int64_t)llround((double)nx / (double)xs);
    const int64_t yss = (int64_t)llround((double)ny / (double)ys);
    const int64_t N = nx * ny;
    const double smooth_gain = (double)smooth_gain_f; /* C float parameter, :242 */
    double *disp = (double *)calloc(2 * N, sizeof(double));
    double *coef = (double *)calloc(4 * N, sizeof(double));
    double *xg = (double *)malloc(sizeof(double) * xs), *yg = (double *)malloc(sizeof(double) * ys);
    int64_t *l_i = (int64_t *)malloc(sizeof(int64_t) * nx), *m_j = (int64_t *)malloc(sizeof(int64_t) * ny);
    int64_t *i_min = (int64_t *)malloc(sizeof(int64_t) * xs), *i_max = (int64_t *)malloc(sizeof(int64_t) * xs);
    int64_t *j_min = (int64_t *)malloc(sizeof(int64_t) * ys), *j_max = (int64_t *)malloc(sizeof(int64_t) * ys);
    double *morphed = (double *)malloc(sizeof(double) * N);
    int8_t *mmask = (int8_t *)malloc(N);
    double *gd = gradient ? (double *)malloc(sizeof(double) * 2 * N) : NULL;
    if (!disp || !coef || !xg || !yg || !l_i || !m_j || !i_min || !i_max || !j_min || !j_max || !morphed ||
        !mmask || (gradient && !gd))
        return -2;
    const int64_t i_shift = xss / 2, j_shift = yss / 2;
    /* x.reshape((xs, xss)).mean(axis=1): pairwise sums of small integer runs are exact */
    for (int64_t l = 0; l < xs; l++) {
        double s = 0;
        for (int64_t k = 0; k < xss; k++) s += (double)(l * xss + k);
        xg[l] = s / (double)xss;
    }
    for (int64_t m = 0; m < ys; m++) {
        double s = 0;
        for (int64_t k = 0; k < yss; k++) s += (double)(m * yss + k);
        yg[m] = s / (double)yss;
    }
    for (int64_t l = 0; l < xs; l++) { i_min[l] = nx; i_max[l] = nx; }
    for (int64_t m = 0; m < ys; m++) { j_min[m] = ny; j_max[m] = ny; }
    for (int64_t i = 0; i < nx; i++) { /* :419-462 */
        int64_t l0 = imin(fdiv(i - i_shift, xss), xs - 2);
        l0 = imax(l0, 0);
        const int64_t l1 = l0 + 1;
        l_i[i] = l0;
        for (int64_t j = 0; j < ny; j++) {
            int64_t m0 = imin(fdiv(j - j_shift, yss), ys - 2);
            m0 = imax(m0, 0);
            const int64_t m1 = m0 + 1;
            m_j[j] = m0;
            const double xi = (double)i, yj = (double)j;
            const double area = (xg[l1] - xg[l0]) * (yg[m1] - yg[m0]);
            const double c0 = (xg[l1] * yg[m1] - xi * yg[m1] - xg[l1] * yj + xi * yj) / area;
            const double c1 = (-xg[l1] * yg[m0] + xi * yg[m0] + xg[l1] * yj - xi * yj) / area;
            const double c2 = (-xg[l0] * yg[m1] + xi * yg[m1] + xg[l0] * yj - xi * yj) / area;
            const double c3 = (xg[l0] * yg[m0] - xi * yg[m0] - xg[l0] * yj + xi * yj) / area;
            coef[i * ny + j] = c0; coef[N + i * ny + j] = c1;
            coef[2 * N + i * ny + j] = c2; coef[3 * N + i * ny + j] = c3;
            for (int a = 0; a < 2; a++)
                disp[a * N + i * ny + j] = sd[(a * xs + l0) * ys + m0] * c0 + sd[(a * xs + l0) * ys + m1] * c1 +
                                           sd[(a * xs + l1) * ys + m0] * c2 + sd[(a * xs + l1) * ys + m1] * c3;
        }
    }
    /* :466-476 np.unique(l_i, return_index, return_counts): first index and count per value */
    for (int64_t i = 0; i < nx;) {
        int64_t l = l_i[i], c = 0, s = i;
        while (i < nx && l_i[i] == l) { i++; c++; }
        i_min[l] = s; i_max[l] = s + c - 1; /* exclusive bound one short: as written */
    }
    for (int64_t j = 0; j < ny;) {
        int64_t m = m_j[j], c = 0, s = j;
        while (j < ny && m_j[j] == m) { j++; c++; }
        j_min[m] = s; j_max[m] = s + c;
    }
    double residuals = 0.0, smoothness_penalty = 0.0;
    double *gres = NULL, *gsm = NULL;
    if (gradient) {
        gres = (double *)calloc(2 * xs * ys, sizeof(double));
        gsm = (double *)calloc(2 * xs * ys, sizeof(double));
        ora_vet_warp(templ, mask, disp, nx, ny, morphed, mmask, gd);
        for (int64_t k = 0; k < N; k++) {
            if (mask[k] > 0) mmask[k] = 1; /* :493 */
            double b = 2 * (input[k] - morphed[k]);
            if (mmask[k] == 1) b = 0;
            gd[k] *= b; gd[N + k] *= b;
        }
        for (int64_t l = 0; l < xs; l++) { /* :508-527 */
            for (int64_t m = 0; m < ys; m++)
                for (int64_t i = i_min[l]; i < i_max[l]; i++)
                    for (int64_t j = j_min[m]; j < j_max[m]; j++) {
                        gres[(0 * xs + l) * ys + m] += gd[i * ny + j] * coef[i * ny + j];
                        gres[(1 * xs + l) * ys + m] += gd[N + i * ny + j] * coef[i * ny + j];
                    }
            for (int64_t m = 1; m < ys; m++)
                for (int64_t i = i_min[l]; i < i_max[l]; i++)
                    for (int64_t j = j_min[m - 1]; j < j_max[m - 1]; j++) {
                        gres[(0 * xs + l) * ys + m] += gd[i * ny + j] * coef[N + i * ny + j];
                        gres[(1 * xs + l) * ys + m] += gd[N + i * ny + j] * coef[N + i * ny + j];
                    }
        }
        for (int64_t l = 1; l < xs; l++) { /* :529-546 */
            for (int64_t m = 0; m < ys; m++)
                for (int64_t i = i_min[l - 1]; i < i_max[l - 1]; i++)
                    for (int64_t j = j_min[m]; j < j_max[m]; j++) {
                        gres[(0 * xs + l) * ys + m] += gd[i * ny + j] * coef[2 * N + i * ny + j];
                        gres[(1 * xs + l) * ys + m] += gd[N + i * ny + j] * coef[2 * N + i * ny + j];
                    }
            for (int64_t m = 1; m < ys; m++)
                for (int64_t i = i_min[l - 1]; i < i_max[l - 1]; i++)
                    for (int64_t j = j_min[m - 1]; j < j_max[m - 1]; j++) {
                        gres[(0 * xs + l) * ys + m] += gd[i * ny + j] * coef[3 * N + i * ny + j];
                        gres[(1 * xs + l) * ys + m] += gd[N + i * ny + j] * coef[3 * N + i * ny + j];
                    }
        }
    } else {
        ora_vet_warp(templ, mask, disp, nx, ny, morphed, mmask, NULL);
        for (int64_t k = 0; k < N; k++) { /* :549-556 */
            if (mask[k] > 0) mmask[k] = 1;
            if (mmask[k] == 0) {
                const double r = morphed[k] - input[k];
                residuals += r * r;
            }
        }
    }
    if (smooth_gain > 0.) { /* :567-614 */
        for (int a = 0; a < 2; a++)
            for (int64_t l = 1; l < xs - 1; l++)
                for (int64_t m = 1; m < ys - 1; m++) {
                    const double *S = sd + a * xs * ys;
                    double dx2 = S[(l + 1) * ys + m] - 2 * S[l * ys + m] + S[(l - 1) * ys + m];
                    dx2 = dx2 / (double)(xss * xss);
                    double dy2 = S[l * ys + m + 1] - 2 * S[l * ys + m] + S[l * ys + m - 1];
                    dy2 = dy2 / (double)(yss * yss);
                    double dxy = S[(l + 1) * ys + m + 1] - S[(l + 1) * ys + m - 1] - S[(l - 1) * ys + m + 1] +
                                 S[(l - 1) * ys + m - 1];
                    dxy = dxy / (double)(4 * xss * yss);
                    if (gradient) {
                        double *G = gsm + a * xs * ys;
                        G[l * ys + m] -= 2 * dx2;
                        G[(l + 1) * ys + m] += dx2;
                        G[(l - 1) * ys + m] += dx2;
                        G[l * ys + m] -= 2 * dy2;
                        G[l * ys + m - 1] += dy2;
                        G[l * ys + m + 1] += dy2;
                        G[(l - 1) * ys + m - 1] += dxy;
                        G[(l - 1) * ys + m + 1] -= dxy;
                        G[(l + 1) * ys + m - 1] -= dxy;
                        G[(l + 1) * ys + m + 1] += dxy;
                    }
                    smoothness_penalty += dx2 * dx2 + 2 * dxy * dxy + dy2 * dy2;
                }
        smoothness_penalty *= smooth_gain;
    }
    if (gradient) {
        for (int64_t k = 0; k < 2 * xs * ys; k++) out[k] = gres[k] + gsm[k] * (2 * smooth_gain);
        free(gres); free(gsm);
    } else {
        out[0] = residuals;
        out[1] = smoothness_penalty;
    }
    free(disp); free(coef); free(xg); free(yg); free(l_i); free(m_j); free(i_min); free(i_max);
    free(j_min); free(j_max); free(morphed); free(mmask); free(gd);
    return 0;
}

/*
 * scipy.ndimage.zoom(a (c,h,w), (1, zh, zw), order=1, mode="nearest") with output shape
 * (c, oh, ow) = round(shape * zoom): coordinates o * (in-1)/(out-1) per axis (1 output:
 * coordinate 0), order-1 taps floor(c), floor(c)+1 clamped, weights w0 = 1 - t, w1 = 1 - w0,
 * value = sum over (y-tap, x-tap) of ((a * wy) * wx) from 0.0 (same as map_coordinates).
 */
void ora_zoom_o1(const double *a, int64_t c, int64_t h, int64_t w, int64_t oh, int64_t ow, double *out)
{
    const double sy = (oh > 1) ? (double)(h - 1) / (double)(oh - 1) : 0.0;
    const double sx = (ow > 1) ? (double)(w - 1) / (double)(ow - 1) : 0.0;
    for (int64_t k = 0; k < c; k++)
        for (int64_t i = 0; i < oh; i++) {
            const double cy = (double)i * sy, fy = floor(cy), ty = cy - fy;
            int64_t y0 = (int64_t)fy, y1 = y0 + 1;
            if (y0 > h - 1) y0 = h - 1;
            if (y1 > h - 1) y1 = h - 1;
            const double wy0 = 1.0 - ty, wy1 = 1.0 - wy0;
            for (int64_t j = 0; j < ow; j++) {
                const double cx = (double)j * sx, fx = floor(cx), tx = cx - fx;
                int64_t x0 = (int64_t)fx, x1 = x0 + 1;
                if (x0 > w - 1) x0 = w - 1;
                if (x1 > w - 1) x1 = w - 1;
                const double wx0 = 1.0 - tx, wx1 = 1.0 - wx0;
                const double *p = a + k * h * w;
                double t = 0.0;
                t += (p[y0 * w + x0] * wy0) * wx0;
                t += (p[y0 * w + x1] * wy0) * wx1;
                t += (p[y1 * w + x0] * wy1) * wx0;
                t += (p[y1 * w + x1] * wy1) * wx1;
                out[(k * oh + i) * ow + j] = t;
            }
        }
}
