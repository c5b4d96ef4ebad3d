// proesmans_body.cuh -- per-thread bodies of the Proesmans optical-flow kernels (proesmans.cu),
// restating pysteps/motion/_proesmans.pyx:1-392.  Like spline_body.cuh this source also compiles
// as plain host C++ (tests/host_kernels/, g++ -ffp-contract=off): the library is built with
// --fmad=false and IEEE division / square root, so the plain operators below round identically on
// both sides and tests/test_kernel_bodies.py can run every body, and the wavefront order of the
// relaxation sweep, on the CPU against the oracle.
#pragma once
#include <math.h>
#include <stddef.h>

#if defined(__CUDACC__)
#define PRO_FN __host__ __device__ __forceinline__
#else
#define PRO_FN inline
#endif

namespace pro {

constexpr double INTENSITY_SCALE = 1.0 / 255.0;  // _proesmans.pyx:16

// :361-392 bilinear sample; the weights use the CLAMPED tap indices (so a coordinate on the last
// row / column gets zero or negative weights), kept as is
PRO_FN double lin(const double *I, int h, int w, double x, double y) {
    long long x0 = (long long)x, x1 = x0 + 1, y0 = (long long)y, y1 = y0 + 1;
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

// pysteps/motion/proesmans.py:79-83 (im - im_min) / (im_max - im_min) * 255.0
PRO_FN double scale_value(double v, double lo, double hi, int do_scale) {
    return do_scale ? (v - lo) / (hi - lo) * 255.0 : v;
}

// scipy.ndimage.gaussian_filter (pysteps/motion/proesmans.py:85-87), one axis: correlate1d with the
// symmetric kernel w[0 .. 2 lw] (centre w[lw]) and mode "reflect" (d c b a | a b c d | d c b a);
// terms are added from the farthest pair inwards, as scipy's symmetric branch does
constexpr int GAUSS_MAX_RADIUS = 64;
struct GaussKernel {
    int lw;
    double w[2 * GAUSS_MAX_RADIUS + 1];
};

PRO_FN double gauss_reflect_at(const double *line, int n, size_t stride, long long i) {
    if (n == 1) return line[0];
    const long long p = 2LL * n;
    i %= p;
    if (i < 0) i += p;
    return i < n ? line[(size_t)i * stride] : line[(size_t)(p - 1 - i) * stride];
}

// value at position l of the line (n samples, given stride)
PRO_FN double gauss_line_value(const double *line, int n, size_t stride, int l, const GaussKernel &k) {
    double tmp = gauss_reflect_at(line, n, stride, l) * k.w[k.lw];
    for (int ii = -k.lw; ii < 0; ii++)
        tmp += (gauss_reflect_at(line, n, stride, (long long)l + ii) + gauss_reflect_at(line, n, stride, (long long)l - ii)) *
               k.w[k.lw + ii];
    return tmp;
}

// :46-58 destination pixel (y, x) of the next pyramid level, source (sh, sw)
PRO_FN double pyr_pixel(const double *src, int sw, int y, int x) {
    return (src[(size_t)(2 * y) * sw + 2 * x] + src[(size_t)(2 * y) * sw + 2 * x + 1] +
            src[(size_t)(2 * y + 1) * sw + 2 * x] + src[(size_t)(2 * y + 1) * sw + 2 * x + 1]) / 4.0;
}

// :256-286 scipy.ndimage.convolve(I, K, mode="constant", cval=0): out[y,x] = sum K[1-dy][1-dx] I[y+dy][x+dx]
PRO_FN void grad_pixel(const double *I, int h, int w, int y, int x, double &gx, double &gy) {
    const double s = INTENSITY_SCALE;
    const double Kx[3][3] = {{1.0 / 8.0 * s, 0.0, -1.0 / 8.0 * s}, {2.0 / 8.0 * s, 0.0, -2.0 / 8.0 * s},
                             {1.0 / 8.0 * s, 0.0, -1.0 / 8.0 * s}};
    const double Ky[3][3] = {{1.0 / 8.0 * s, 2.0 / 8.0 * s, 1.0 / 8.0 * s}, {0.0, 0.0, 0.0},
                             {-1.0 / 8.0 * s, -2.0 / 8.0 * s, -1.0 / 8.0 * s}};
    // scipy flips the kernel and correlates: the terms are added in row-major order of the offsets
    // (dy, dx) = (-1,-1) .. (1,1), zero weights skipped
    gx = 0.0;
    gy = 0.0;
    for (int dy = -1; dy <= 1; dy++)
        for (int dx = -1; dx <= 1; dx++) {
            const int yy = y + dy, xx = x + dx;
            const double v = (yy < 0 || yy >= h || xx < 0 || xx >= w) ? 0.0 : I[(size_t)yy * w + xx];
            if (Kx[1 - dy][1 - dx] != 0.0) gx += v * Kx[1 - dy][1 - dx];
            if (Ky[1 - dy][1 - dx] != 0.0) gy += v * Ky[1 - dy][1 - dx];
        }
}

// :209-228 forward-backward inconsistency of pixel (y, x) for direction i: the norm of
// V[i] + V[1-i] sampled at the displaced position, or -1 outside the domain.  V is (2,2,h,w).
PRO_FN double cons_pixel(const double *V, int h, int w, int i, int y, int x) {
    const size_t N = (size_t)h * w, q = (size_t)y * w + x;
    const double *V11 = V + (size_t)(2 * i) * N, *V12 = V + (size_t)(2 * i + 1) * N;
    const double *V21 = V + (size_t)(2 * (1 - i)) * N, *V22 = V + (size_t)(2 * (1 - i) + 1) * N;
    const double xd = x + V11[q], yd = y + V12[q];
    if (xd >= 0 && yd >= 0 && xd < w && yd < h) {
        const double ub = lin(V21, h, w, xd, yd), vb = lin(V22, h, w, xd, yd);
        const double ud = V11[q] + ub, vd = V12[q] + vb;
        return sqrt(ud * ud + vd * vd);
    }
    return -1.0;
}

// :229-233 the mean inconsistency is accumulated row by row (sum and count of the valid pixels of
// one row of the map; the row results are then added in row order).  The reference's own order is
// whatever its -ffast-math build makes of the raster-order loop; this one is reproducible by one
// chain per row, and the oracle uses the same.
PRO_FN void cons_row_sum(const double *grow, int w, double &sum, long long &count) {
    sum = 0.0;
    count = 0;
    for (int x = 0; x < w; x++)
        if (grow[x] >= 0.0) {
            sum += grow[x];
            count += 1;
        }
}

PRO_FN double cons_K(const double *row_sum, const long long *row_count, int h) {
    double s = 0.0;
    long long c = 0;
    for (int y = 0; y < h; y++) {
        s += row_sum[y];
        c += row_count[y];
    }
    return c > 0 ? 0.9 * s / (double)c : 0.0;
}

// :236-254 the consistency weight from the inconsistency g and K = 0.9 * mean
PRO_FN double cons_weight(double g, double K) {
    if (K > 1e-8) {
        if (g >= 0.0) {
            const double r = g / K;
            return 1.0 / (1.0 + r * r);
        }
        return 1.0;
    }
    return 1.0;
}

// :166-188 consistency-weighted average of the 8 neighbours of component plane Vc
PRO_FN double laplacian(const double *gi, const double *Vc, int w, int x, int y) {
#define PRO_G(dy, dx) gi[(long long)(y + (dy)) * w + (x + (dx))]
#define PRO_V(dy, dx) Vc[(long long)(y + (dy)) * w + (x + (dx))]
    const double sw = (PRO_G(-1, 0) + PRO_G(0, -1) + PRO_G(0, 1) + PRO_G(1, 0)) / 6.0 +
                      (PRO_G(-1, -1) + PRO_G(-1, 1) + PRO_G(1, -1) + PRO_G(1, 1)) / 12.0;
    if (sw > 1e-8) {
        const double v = (PRO_G(-1, 0) * PRO_V(-1, 0) + PRO_G(0, -1) * PRO_V(0, -1) + PRO_G(0, 1) * PRO_V(0, 1) +
                          PRO_G(1, 0) * PRO_V(1, 0)) / 6.0 +
                         (PRO_G(-1, -1) * PRO_V(-1, -1) + PRO_G(-1, 1) * PRO_V(-1, 1) +
                          PRO_G(1, -1) * PRO_V(1, -1) + PRO_G(1, 1) * PRO_V(1, 1)) / 12.0;
        return v / sw;
    }
    return 0.0;
#undef PRO_G
#undef PRO_V
}

// :126-150 one Gauss-Seidel update of pixel (y, x), 1 <= x <= w-2, 1 <= y <= h-2, in place in
// Vj (2,h,w).  It reads the 8 neighbours of (y, x): in the reference's raster order the west and
// the three north ones already hold this sweep's values.
PRO_FN void sweep_pixel(const double *R1, const double *R2, const double *G1, const double *G2,
                        const double *gam, double *Vj, int h, int w, int x, int y, double lam) {
    const size_t N = (size_t)h * w, q = (size_t)y * w + x;
    const double a1 = laplacian(gam, Vj, w, x, y);
    const double a2 = laplacian(gam, Vj + N, w, x, y);
    const double xd = x + a1, yd = y + a2;
    double n1 = a1, n2 = a2;
    if (xd >= 0 && xd < w - 1 && yd >= 0 && yd < h - 1) {
        const double It = (lin(R2, h, w, xd, yd) - R1[q]) * INTENSITY_SCALE;
        const double gx = G1[q], gy = G2[q];
        const double ic = lam * It / (1.0 + lam * (gx * gx + gy * gy));
        n1 = a1 - gx * ic;
        n2 = a2 - gy * ic;
    }
    Vj[q] = n1;
    Vj[N + q] = n2;
}

// Wavefronts of the sweep: pixels with the same t = x + 2y are mutually independent and depend
// only on smaller t (west: t-1, north-east: t-1, north: t-2, north-west: t-3), so processing t in
// increasing order -- rows of one t in any order or in parallel -- is the raster-order sweep.
PRO_FN int sweep_first_t() { return 3; }
PRO_FN int sweep_last_t(int h, int w) { return (w - 2) + 2 * (h - 2); }
PRO_FN void sweep_rows_of(int t, int h, int w, int &ylo, int &yhi) {
    ylo = t - (w - 2);
    ylo = ylo <= 2 ? 1 : (ylo + 1) / 2;  // smallest y >= 1 with t - 2y <= w-2
    yhi = (t - 1) / 2;                   // largest y with t - 2y >= 1
    if (yhi > h - 2) yhi = h - 2;
}

// :288-309 border element e of the 2(h+w)-4 border pixels of plane v (h,w) copies its interior
// neighbour; all sources are interior pixels, so the elements are independent
PRO_FN void fill_edge_element(double *v, int h, int w, int e) {
    const int top = w, bottom = 2 * w;  // [0,w): row 0, [w,2w): row h-1, then columns
    int y, x, sy, sx;
    if (e < top) { y = 0; x = e; }
    else if (e < bottom) { y = h - 1; x = e - top; }
    else {
        const int k = e - bottom;  // rows 1 .. h-2, left then right
        y = 1 + k / 2;
        x = (k & 1) ? w - 1 : 0;
    }
    sy = y == 0 ? 1 : (y == h - 1 ? h - 2 : y);
    sx = x == 0 ? 1 : (x == w - 1 ? w - 2 : x);
    v[(size_t)y * w + x] = v[(size_t)sy * w + sx];
}
PRO_FN int fill_edge_count(int h, int w) { return 2 * w + 2 * (h - 2); }

// :311-359 pixel (yn, xn) of the next finer level of plane src (hp,wp) -> value of the (hn,wn) plane
PRO_FN double next_level_pixel(const double *src, int hp, int wp, int yn, int xn) {
    const double yc = yn / 2.0, xc = xn / 2.0;
    int yci = yn / 2, xci = xn / 2;
    double v;
    if (xn % 2 != 0 || yn % 2 != 0) {
        v = lin(src, hp, wp, xc, yc);
    } else {
        if (xci > wp - 1) xci = wp - 1;
        if (yci > hp - 1) yci = hp - 1;
        v = src[(size_t)yci * wp + xci];
    }
    return 2.0 * v;
}

}  // namespace pro
