/*
 * oracle/proesmans_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, float64) of the native extension of the Proesmans optical-flow
 * method, pysteps/motion/_proesmans.pyx:1-392 (called from pysteps/motion/proesmans.py:20-94).
 * Every function below names the lines it follows.  The reference builds that extension with
 * -O3 -ffast-math (setup.py:27-28), so its own floating-point results depend on the compiler;
 * parity is therefore a tolerance (tests/test_oracle_proesmans.py pins this file against the
 * reference extension compiled out of tree with the reference's flags, to ~1e-10 px), not bit
 * equality.  What must be preserved exactly is the ORDER OF UPDATES: the relaxation sweep is a
 * raster-order Gauss-Seidel iteration (every pixel reads the already updated west and north
 * neighbours).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define INTENSITY_SCALE (1.0 / 255.0)

/* 0 (default): the mean inconsistency is accumulated row by row, the order the CUDA path
 * reproduces; 1: in one raster-order chain, the literal order of _proesmans.pyx:209-228 -- used
 * only to pin this file bit for bit against the reference source built WITHOUT -ffast-math. */
static int g_raster_sum = 0;
void ora_proesmans_raster_sum(int on) { g_raster_sum = on; }

/* _proesmans.pyx:361-392 (note: the weights use the CLAMPED tap indices) */
static double linear_interpolate(const double *I, int64_t h, int64_t w, double x, double y)
{
    int64_t x0 = (int64_t)x, x1 = x0 + 1, y0 = (int64_t)y, y1 = y0 + 1;
    if (x0 < 0) x0 = 0;
    if (x0 > w - 1) x0 = w - 1;
    if (x1 < 0) x1 = 0;
    if (x1 > w - 1) x1 = w - 1;
    if (y0 < 0) y0 = 0;
    if (y0 > h - 1) y0 = h - 1;
    if (y1 < 0) y1 = 0;
    if (y1 > h - 1) y1 = h - 1;
    const double Ia = I[y0 * w + x0], Ib = I[y1 * w + x0], Ic = I[y0 * w + x1], Id = I[y1 * w + x1];
    const double wa = (x1 - x) * (y1 - y), wb = (x1 - x) * (y - y0), wc = (x - x0) * (y1 - y),
                 wd = (x - x0) * (y - y0);
    return wa * Ia + wb * Ib + wc * Ic + wd * Id;
}

/* :46-58 one pyramid level: mean of 2x2 blocks, destination (sh/2, sw/2) */
void ora_proesmans_pyr_down(const double *src, int64_t sh, int64_t sw, double *dst)
{
    const int64_t dh = sh / 2, dw = sw / 2;
    (void)sh;
    for (int64_t y = 0; y < dh; y++)
        for (int64_t x = 0; x < dw; x++)
            dst[y * dw + x] = (src[2 * y * sw + 2 * x] + src[2 * y * sw + 2 * x + 1] +
                               src[(2 * y + 1) * sw + 2 * x] + src[(2 * y + 1) * sw + 2 * x + 1]) / 4.0;
}

/* :256-286 scipy.ndimage.convolve(I, K, mode="constant", cval=0) with the two Sobel-like kernels:
 * a true convolution, out[y,x] = sum_{i,j} K[i][j] * I[y+1-i][x+1-j] */
void ora_proesmans_gradients(const double *I, int64_t h, int64_t w, double *G /* (2,h,w) */)
{
    const double s = INTENSITY_SCALE;
    const double Kx[3][3] = {{1.0 / 8.0 * s, 0.0, -1.0 / 8.0 * s}, {2.0 / 8.0 * s, 0.0, -2.0 / 8.0 * s},
                             {1.0 / 8.0 * s, 0.0, -1.0 / 8.0 * s}};
    const double Ky[3][3] = {{1.0 / 8.0 * s, 2.0 / 8.0 * s, 1.0 / 8.0 * s}, {0.0, 0.0, 0.0},
                             {-1.0 / 8.0 * s, -2.0 / 8.0 * s, -1.0 / 8.0 * s}};
    for (int64_t y = 0; y < h; y++)
        for (int64_t x = 0; x < w; x++) {
            /* scipy flips the kernel and correlates: the terms are added in row-major order of
             * the OFFSETS (dy, dx) = (-1,-1) .. (1,1), zero weights skipped */
            double gx = 0.0, gy = 0.0;
            for (int dy = -1; dy <= 1; dy++)
                for (int dx = -1; dx <= 1; dx++) {
                    const int64_t yy = y + dy, xx = x + dx;
                    const double v = (yy < 0 || yy >= h || xx < 0 || xx >= w) ? 0.0 : I[yy * w + xx];
                    if (Kx[1 - dy][1 - dx] != 0.0) gx += v * Kx[1 - dy][1 - dx];
                    if (Ky[1 - dy][1 - dx] != 0.0) gy += v * Ky[1 - dy][1 - dx];
                }
            G[y * w + x] = gx;
            G[h * w + y * w + x] = gy;
        }
}

/* :190-254 forward-backward consistency maps; V (2,2,h,w), GAMMA (2,h,w) */
void ora_proesmans_consistency(const double *V, int64_t h, int64_t w, double *GAMMA)
{
    const int64_t N = h * w;
    for (int i = 0; i < 2; i++) {
        double c_sum = 0.0;
        int64_t c_count = 0;
        const double *V11 = V + (2 * i + 0) * N, *V12 = V + (2 * i + 1) * N;
        const double *V21 = V + (2 * (1 - i) + 0) * N, *V22 = V + (2 * (1 - i) + 1) * N;
        double *g = GAMMA + i * N;
        for (int64_t y = 0; y < h; y++) {
            /* The mean of the inconsistency is accumulated row by row (a row sum, then the sum of
             * the row sums) instead of in one raster-order chain: the reference's own order is
             * whatever its -ffast-math build vectorised it into, and this one can be reproduced
             * exactly by a parallel implementation (one chain per row). */
            double row_sum = 0.0;
            for (int64_t x = 0; x < w; x++) {
                const double xd = x + V11[y * w + x], yd = y + V12[y * w + x];
                if (xd >= 0 && yd >= 0 && xd < w && yd < h) {
                    const double ub = linear_interpolate(V21, h, w, xd, yd);
                    const double vb = linear_interpolate(V22, h, w, xd, yd);
                    const double ud = V11[y * w + x] + ub, vd = V12[y * w + x] + vb;
                    const double c = sqrt(ud * ud + vd * vd);
                    g[y * w + x] = c;
                    if (g_raster_sum) c_sum += c; else row_sum += c;
                    c_count += 1;
                } else {
                    g[y * w + x] = -1.0;
                }
            }
            if (!g_raster_sum) c_sum += row_sum;
        }
        const double K = c_count > 0 ? 0.9 * c_sum / c_count : 0.0;
        for (int64_t q = 0; q < N; q++) {
            if (K > 1e-8) {
                if (g[q] >= 0.0) {
                    const double r = g[q] / K;
                    g[q] = 1.0 / (1.0 + r * r);
                } else {
                    g[q] = 1.0;
                }
            } else {
                g[q] = 1.0;
            }
        }
    }
}

/* :166-188 */
static double laplacian(const double *gi, const double *Vc, int64_t w, int64_t x, int64_t y)
{
#define GI(dy, dx) gi[(y + (dy)) * w + x + (dx)]
#define VV(dy, dx) Vc[(y + (dy)) * w + x + (dx)]
    const double sw = (GI(-1, 0) + GI(0, -1) + GI(0, 1) + GI(1, 0)) / 6.0 +
                      (GI(-1, -1) + GI(-1, 1) + GI(1, -1) + GI(1, 1)) / 12.0;
    if (sw > 1e-8) {
        const double v = (GI(-1, 0) * VV(-1, 0) + GI(0, -1) * VV(0, -1) + GI(0, 1) * VV(0, 1) +
                          GI(1, 0) * VV(1, 0)) / 6.0 +
                         (GI(-1, -1) * VV(-1, -1) + GI(-1, 1) * VV(-1, 1) + GI(1, -1) * VV(1, -1) +
                          GI(1, 1) * VV(1, 1)) / 12.0;
        return v / sw;
    }
    return 0.0;
#undef GI
#undef VV
}

/* :288-309 */
static void fill_edges(double *Vj /* (2,h,w) */, int64_t h, int64_t w)
{
    for (int i = 0; i < 2; i++) {
        double *v = Vj + i * h * w;
        for (int64_t x = 1; x < w - 1; x++) {
            v[x] = v[w + x];
            v[(h - 1) * w + x] = v[(h - 2) * w + x];
        }
        for (int64_t y = 1; y < h - 1; y++) {
            v[y * w] = v[y * w + 1];
            v[y * w + w - 1] = v[y * w + w - 2];
        }
        v[0] = v[w + 1];
        v[w - 1] = v[w + w - 2];
        v[(h - 1) * w] = v[(h - 2) * w + 1];
        v[(h - 1) * w + w - 1] = v[(h - 2) * w + w - 2];
    }
}

/* :81-164 num_iter relaxation iterations on one pyramid level; R (2,h,w), V (2,2,h,w) in place */
int ora_proesmans_level(const double *R, int64_t h, int64_t w, double *V, int64_t num_iter, double lam)
{
    const int64_t N = h * w;
    double *G = (double *)malloc(sizeof(double) * 4 * N);
    double *GAMMA = (double *)malloc(sizeof(double) * 2 * N);
    if (!G || !GAMMA) {
        free(G); free(GAMMA);
        return -1;
    }
    ora_proesmans_gradients(R, h, w, G);
    ora_proesmans_gradients(R + N, h, w, G + 2 * N);
    for (int64_t it = 0; it < num_iter; it++) {
        ora_proesmans_consistency(V, h, w, GAMMA);
        for (int j = 0; j < 2; j++) {
            const double *R1 = R + j * N, *R2 = R + (1 - j) * N;
            const double *G1 = G + (2 * j) * N, *G2 = G + (2 * j + 1) * N;
            double *Vj = V + 2 * j * N;
            const double *gam = GAMMA + j * N;
            for (int64_t y = 1; y < h - 1; y++)
                for (int64_t x = 1; x < w - 1; x++) {
                    const double a1 = laplacian(gam, Vj, w, x, y);
                    const double a2 = laplacian(gam, Vj + N, w, x, y);
                    const double xd = x + a1, yd = y + a2;
                    double n1 = a1, n2 = a2;
                    if (xd >= 0 && xd < w - 1 && yd >= 0 && yd < h - 1) {
                        const double It = (linear_interpolate(R2, h, w, xd, yd) - R1[y * w + x]) * INTENSITY_SCALE;
                        const double gx = G1[y * w + x], gy = G2[y * w + x];
                        const double ic = lam * It / (1.0 + lam * (gx * gx + gy * gy));
                        n1 = a1 - gx * ic;
                        n2 = a2 - gy * ic;
                    }
                    Vj[y * w + x] = n1;
                    Vj[N + y * w + x] = n2;
                }
            fill_edges(Vj, h, w);
        }
    }
    free(G); free(GAMMA);
    return 0;
}

/* :311-359 prolongation of the flow to the next finer level; Vp (2,2,hp,wp) -> Vn (2,2,hn,wn) */
void ora_proesmans_next_level(const double *Vp, int64_t hp, int64_t wp, double *Vn, int64_t hn, int64_t wn)
{
    for (int c = 0; c < 4; c++) {
        const double *src = Vp + c * hp * wp;
        double *dst = Vn + c * hn * wn;
        for (int64_t yn = 0; yn < hn; yn++) {
            const double yc = yn / 2.0;
            for (int64_t xn = 0; xn < wn; xn++) {
                const double xc = xn / 2.0;
                int64_t yci = yn / 2, xci = xn / 2;
                double v;
                if (xn % 2 != 0 || yn % 2 != 0) {
                    v = linear_interpolate(src, hp, wp, xc, yc);
                } else {
                    if (xci > wp - 1) xci = wp - 1;
                    if (yci > hp - 1) yci = hp - 1;
                    v = src[yci * wp + xci];
                }
                dst[yn * wn + xn] = 2.0 * v;
            }
        }
    }
}
