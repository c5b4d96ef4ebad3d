// vet.cu -- Variational Echo Tracking cost function / gradient, morphing and the
// bilinear zoom of the sector field (sm_100a).
//
// Replaces the native extension of the reference, pysteps/motion/_vet.pyx:
//   _warp           (:66-232)   bilinear morphing of an image by a displacement field
//   _cost_function  (:238-621)  sector -> pixel displacement, warp, masked squared
//                               residual (cost) or its sector gradient, smoothness term
// and scipy.ndimage.zoom(order=1, mode="nearest") of pysteps/motion/vet.py:580-589,621-630.
// The CG optimiser stays on the host (scipy.optimize.minimize, vet.py:593-600) and calls
// one cost or gradient evaluation ~1000 times per field: each evaluation is ONE fused
// kernel over the image (two streamed images + int8 mask, ~17 B/pixel) plus a tiny
// finalising kernel; nothing of size m x n is materialised (the reference allocates six
// float64 temporaries per call).
//
// Reductions are deterministic (fixed partition, fixed tree) so repeated evaluations are
// bit-identical run to run, like the reference's static OpenMP schedule.  Pixels are
// partitioned by the interpolation CELL (l0, m0) they fall in: all pixels of a cell feed
// the same four sector corners, so a CTA reduces eight sums (4 corners x 2 axes) and the
// finalising kernel adds, per sector, the <= 4 adjacent cells in the reference's loop order.
#include <math_constants.h>

#include <vector>

#include "common.cuh"

namespace {

constexpr int VET_THREADS = 256;

struct VetGeom {
    int nx, ny, xs, ys, xss, yss, i_shift, j_shift, ncx, ncy, strips;
};

__device__ __forceinline__ int cell_start(int l, int shift, int ss) { return l == 0 ? 0 : shift + l * ss; }
__device__ __forceinline__ int cell_end(int l, int shift, int ss, int ns, int n) {
    return l == ns - 2 ? n : shift + (l + 1) * ss;
}
// sector centre: mean of the pixel indices of the block (_vet.pyx:389-390), exact in float64
__device__ __forceinline__ double centre(int l, int ss) { return (double)(l * ss) + (double)(ss - 1) * 0.5; }

struct Warped {
    double value, gx, gy;  // morphed image and its gradient w.r.t. (x, y) displacement sign
    bool masked;
};

// _vet.pyx:160-228 for one pixel
__device__ __forceinline__ Warped warp_pixel(const double *__restrict__ image, const int8_t *__restrict__ mask,
                                             int nx, int ny, int x, int y, double dispx, double dispy) {
    const int xmi = nx - 1, ymi = ny - 1;
    double xf = __dsub_rn((double)x, dispx), yf = __dsub_rn((double)y, dispy);
    int x0, x1, y0, y1;
    if (xf < 0) { xf = 0; x0 = 0; x1 = 0; }
    else if (xf > (double)xmi) { xf = (double)xmi; x0 = xmi; x1 = xmi; }
    else { x0 = (int)floor(xf); x1 = min(x0 + 1, xmi); }
    if (yf < 0) { yf = 0; y0 = 0; y1 = 0; }
    else if (yf > (double)ymi) { yf = (double)ymi; y0 = ymi; y1 = ymi; }
    else { y0 = (int)floor(yf); y1 = min(y0 + 1, ymi); }
    const double dx = __dsub_rn(xf, (double)x0), dy = __dsub_rn(yf, (double)y0);
    const double i00 = image[(size_t)x0 * ny + y0], i10 = image[(size_t)x1 * ny + y0];
    const double i01 = image[(size_t)x0 * ny + y1], i11 = image[(size_t)x1 * ny + y1];
    const double f10 = __dsub_rn(i10, i00), f01 = __dsub_rn(i01, i00);
    const double f11 = __dadd_rn(__dsub_rn(__dsub_rn(i00, i10), i01), i11);
    Warped w;
    w.value = __dadd_rn(__dadd_rn(__dadd_rn(i00, __dmul_rn(dx, f10)), __dmul_rn(dy, f01)),
                        __dmul_rn(__dmul_rn(dx, dy), f11));
    w.gx = __dadd_rn(f10, __dmul_rn(dy, f11));
    w.gy = __dadd_rn(f01, __dmul_rn(dx, f11));
    const double m00 = (double)mask[(size_t)x0 * ny + y0], m10 = (double)mask[(size_t)x1 * ny + y0];
    const double m01 = (double)mask[(size_t)x0 * ny + y1], m11 = (double)mask[(size_t)x1 * ny + y1];
    const double g10 = m10 - m00, g01 = m01 - m00, g11 = m00 - m10 - m01 + m11;
    const double mv = __dadd_rn(__dadd_rn(__dadd_rn(m00, __dmul_rn(dx, g10)), __dmul_rn(dy, g01)),
                                __dmul_rn(__dmul_rn(dx, dy), g11));
    // <int8>(...) truncates toward zero; the interpolated mask overwrites the out-of-range flag
    w.masked = ((signed char)(int)mv) != 0;
    return w;
}

template <int NV>
__device__ __forceinline__ void block_reduce(double (&v)[NV], double *sm /* NV * 8 */) {
#pragma unroll
    for (int k = 0; k < NV; k++)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] = __dadd_rn(v[k], __shfl_xor_sync(0xffffffffu, v[k], o));
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < NV; k++) sm[k * 8 + wid] = v[k];
    __syncthreads();
    if (threadIdx.x == 0)
#pragma unroll
        for (int k = 0; k < NV; k++) {
            double s = sm[k * 8];
            for (int w = 1; w < VET_THREADS / 32; w++) s = __dadd_rn(s, sm[k * 8 + w]);
            v[k] = s;
        }
}

// grid = (strips, ncx * ncy).  MODE 0: partial[cell][strip] = sum of squared residuals.
// MODE 1: partial[cell][strip][corner k][axis a] = sum g_a * coef_k.  MODE 2: both from ONE pass
// over the images (entries 0..7 the gradient sums, entry 8 the residual sum) -- a line search
// asks for value and slope at the same point.
template <int MODE>
__global__ void __launch_bounds__(VET_THREADS)
vet_eval_kernel(const double *__restrict__ sd, const double *__restrict__ templ, const double *__restrict__ input,
                const int8_t *__restrict__ mask, const VetGeom g, double *__restrict__ partial) {
    __shared__ double sm[9 * 8];
    const int cell = blockIdx.y, strip = blockIdx.x;
    const int l0 = cell / g.ncy, m0 = cell - l0 * g.ncy, l1 = l0 + 1, m1 = m0 + 1;
    const int i_beg = cell_start(l0, g.i_shift, g.xss), i_end = cell_end(l0, g.i_shift, g.xss, g.xs, g.nx);
    const int j_beg = cell_start(m0, g.j_shift, g.yss), j_end = cell_end(m0, g.j_shift, g.yss, g.ys, g.ny);
    const int rows = i_end - i_beg, cols = j_end - j_beg;
    const int rps = (rows + g.strips - 1) / g.strips;   // rows per strip
    const int r0 = i_beg + strip * rps, r1 = min(r0 + rps, i_end);
    const double xg0 = centre(l0, g.xss), xg1 = centre(l1, g.xss);
    const double yg0 = centre(m0, g.yss), yg1 = centre(m1, g.yss);
    const double area = __dmul_rn(__dsub_rn(xg1, xg0), __dsub_rn(yg1, yg0));
    // x / area: when the area is a power of two (sector sides 64, 128, ... -- every level of a
    // 2^k-sized frame) the quotient equals x * (1 / area) exactly, bit for bit, and the four
    // divisions per pixel become multiplications
    const long long abits = __double_as_longlong(area);
    const int aexp = (int)((abits >> 52) & 0x7ff);
    const bool area_pow2 = (abits & 0x000fffffffffffffll) == 0 && aexp > 64 && aexp < 1983;
    const double inv_area = area_pow2 ? __ddiv_rn(1.0, area) : 0.0;
    auto over_area = [&](double num) -> double { return area_pow2 ? __dmul_rn(num, inv_area) : __ddiv_rn(num, area); };
    const size_t S = (size_t)g.xs * g.ys;
    double s00[2], s01[2], s10[2], s11[2];
#pragma unroll
    for (int a = 0; a < 2; a++) {
        s00[a] = sd[a * S + (size_t)l0 * g.ys + m0];
        s01[a] = sd[a * S + (size_t)l0 * g.ys + m1];
        s10[a] = sd[a * S + (size_t)l1 * g.ys + m0];
        s11[a] = sd[a * S + (size_t)l1 * g.ys + m1];
    }
    constexpr bool GRAD = MODE >= 1, COST = MODE != 1;
    constexpr int NV = MODE == 0 ? 1 : (MODE == 1 ? 8 : 9);
    constexpr int CI = MODE == 0 ? 0 : 8;  // slot of the residual sum
    double acc[NV];
#pragma unroll
    for (int k = 0; k < NV; k++) acc[k] = 0.0;
    const int npx = max(r1 - r0, 0) * cols;
    for (int p = threadIdx.x; p < npx; p += VET_THREADS) {
        const int i = r0 + p / cols, j = j_beg + p % cols;
        const double xi = (double)i, yj = (double)j;
        // _vet.pyx:436-454, expression order as written
        const double xy = __dmul_rn(xi, yj);
        const double c0 = over_area(__dadd_rn(__dsub_rn(__dsub_rn(__dmul_rn(xg1, yg1), __dmul_rn(xi, yg1)),
                                                        __dmul_rn(xg1, yj)), xy));
        const double c1 = over_area(__dsub_rn(__dadd_rn(__dadd_rn(__dmul_rn(-xg1, yg0), __dmul_rn(xi, yg0)),
                                                        __dmul_rn(xg1, yj)), xy));
        const double c2 = over_area(__dsub_rn(__dadd_rn(__dadd_rn(__dmul_rn(-xg0, yg1), __dmul_rn(xi, yg1)),
                                                        __dmul_rn(xg0, yj)), xy));
        const double c3 = over_area(__dadd_rn(__dsub_rn(__dsub_rn(__dmul_rn(xg0, yg0), __dmul_rn(xi, yg0)),
                                                        __dmul_rn(xg0, yj)), xy));
        double disp[2];
#pragma unroll
        for (int a = 0; a < 2; a++)  // :456-462
            disp[a] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(s00[a], c0), __dmul_rn(s01[a], c1)),
                                          __dmul_rn(s10[a], c2)), __dmul_rn(s11[a], c3));
        const Warped w = warp_pixel(templ, mask, g.nx, g.ny, i, j, disp[0], disp[1]);
        const size_t idx = (size_t)i * g.ny + j;
        const bool mm = w.masked || mask[idx] > 0;  // :493 / :554
        if (COST) {
            if (!mm) {
                const double r = __dsub_rn(w.value, input[idx]);
                acc[CI] = __dadd_rn(acc[CI], __dmul_rn(r, r));
            }
        }
        if (GRAD) {
            // the reference's row range [i_min, i_max) stops one row short of the band (:466-476)
            if (i < i_end - 1) {
                const double b = mm ? 0.0 : __dmul_rn(2.0, __dsub_rn(input[idx], w.value));
                const double g0 = __dmul_rn(w.gx, b), g1 = __dmul_rn(w.gy, b);
                acc[0] = __dadd_rn(acc[0], __dmul_rn(g0, c0)); acc[1] = __dadd_rn(acc[1], __dmul_rn(g1, c0));
                acc[2] = __dadd_rn(acc[2], __dmul_rn(g0, c1)); acc[3] = __dadd_rn(acc[3], __dmul_rn(g1, c1));
                acc[4] = __dadd_rn(acc[4], __dmul_rn(g0, c2)); acc[5] = __dadd_rn(acc[5], __dmul_rn(g1, c2));
                acc[6] = __dadd_rn(acc[6], __dmul_rn(g0, c3)); acc[7] = __dadd_rn(acc[7], __dmul_rn(g1, c3));
            }
        }
    }
    block_reduce<NV>(acc, sm);
    if (threadIdx.x == 0)
#pragma unroll
        for (int k = 0; k < NV; k++) partial[((size_t)cell * g.strips + strip) * NV + k] = acc[k];
}

__device__ __forceinline__ bool interior(int l, int m, int xs, int ys) {
    return l >= 1 && l <= xs - 2 && m >= 1 && m <= ys - 2;
}

struct Deriv { double dx2, dy2, dxy; };

// second differences of one displacement component at an interior sector (:573-592)
__device__ __forceinline__ Deriv second_diff(const double *__restrict__ S, int l, int m, int ys, int xss, int yss) {
    Deriv d;
    d.dx2 = __ddiv_rn(__dadd_rn(__dsub_rn(S[(l + 1) * ys + m], __dmul_rn(2.0, S[l * ys + m])), S[(l - 1) * ys + m]),
                      (double)(xss * xss));
    d.dy2 = __ddiv_rn(__dadd_rn(__dsub_rn(S[l * ys + m + 1], __dmul_rn(2.0, S[l * ys + m])), S[l * ys + m - 1]),
                      (double)(yss * yss));
    d.dxy = __ddiv_rn(__dadd_rn(__dsub_rn(__dsub_rn(S[(l + 1) * ys + m + 1], S[(l + 1) * ys + m - 1]),
                                          S[(l - 1) * ys + m + 1]), S[(l - 1) * ys + m - 1]),
                      (double)(4 * xss * yss));
    return d;
}

// cost: out[0] = residuals (fixed-order sum of the partials), out[1] = smoothness penalty
__global__ void __launch_bounds__(VET_THREADS)
vet_final_cost_kernel(const double *__restrict__ partial, int nparts, int nv, int off,
                      const double *__restrict__ sd, const VetGeom g, double smooth_gain,
                      double *__restrict__ out) {
    __shared__ double sm[2 * 8];
    double v[2] = {0.0, 0.0};
    for (int i = threadIdx.x; i < nparts; i += VET_THREADS) v[0] = __dadd_rn(v[0], partial[(size_t)i * nv + off]);
    if (smooth_gain > 0.0) {
        const int ni = max(g.xs - 2, 0) * max(g.ys - 2, 0);
        for (int t = threadIdx.x; t < 2 * ni; t += VET_THREADS) {
            const int a = t / ni, r = t - a * ni;
            const int l = 1 + r / (g.ys - 2), m = 1 + r % (g.ys - 2);
            const Deriv d = second_diff(sd + (size_t)a * g.xs * g.ys, l, m, g.ys, g.xss, g.yss);
            v[1] = __dadd_rn(v[1], __dadd_rn(__dadd_rn(__dmul_rn(d.dx2, d.dx2), __dmul_rn(__dmul_rn(2.0, d.dxy), d.dxy)),
                                             __dmul_rn(d.dy2, d.dy2)));
        }
    }
    block_reduce<2>(v, sm);
    if (threadIdx.x == 0) {
        out[0] = v[0];
        out[1] = __dmul_rn(v[1], smooth_gain);
    }
}

// gradient: out (2, xs, ys) = grad_residuals + 2 * smooth_gain * grad_smooth
__global__ void __launch_bounds__(VET_THREADS)
vet_final_grad_kernel(const double *__restrict__ partial, int nv, const double *__restrict__ sd, const VetGeom g,
                      double smooth_gain, double *__restrict__ out) {
    const int t = blockIdx.x * VET_THREADS + threadIdx.x;
    const int S = g.xs * g.ys;
    if (t >= 2 * S) return;
    const int a = t / S, r = t - a * S, l = r / g.ys, m = r - l * g.ys;
    double gr = 0.0;
    // the reference's loop order: corner 0 of cell (l,m), corner 1 of (l,m-1), corner 2 of
    // (l-1,m), corner 3 of (l-1,m-1)   (_vet.pyx:508-546)
    const int cl[4] = {l, l, l - 1, l - 1}, cm[4] = {m, m - 1, m, m - 1};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (cl[k] < 0 || cl[k] >= g.ncx || cm[k] < 0 || cm[k] >= g.ncy) continue;
        const size_t cell = (size_t)cl[k] * g.ncy + cm[k];
        for (int s = 0; s < g.strips; s++) gr = __dadd_rn(gr, partial[(cell * g.strips + s) * nv + 2 * k + a]);
    }
    double gs = 0.0;
    if (smooth_gain > 0.0) {
        const double *Sd = sd + (size_t)a * S;
        auto D = [&](int ll, int mm) -> Deriv {
            Deriv z; z.dx2 = z.dy2 = z.dxy = 0.0;
            return interior(ll, mm, g.xs, g.ys) ? second_diff(Sd, ll, mm, g.ys, g.xss, g.yss) : z;
        };
        const Deriv c = D(l, m);
        gs = __dmul_rn(-2.0, c.dx2);
        gs = __dadd_rn(gs, D(l - 1, m).dx2);
        gs = __dadd_rn(gs, D(l + 1, m).dx2);
        gs = __dadd_rn(gs, __dmul_rn(-2.0, c.dy2));
        gs = __dadd_rn(gs, D(l, m + 1).dy2);
        gs = __dadd_rn(gs, D(l, m - 1).dy2);
        gs = __dadd_rn(gs, D(l + 1, m + 1).dxy);
        gs = __dsub_rn(gs, D(l + 1, m - 1).dxy);
        gs = __dsub_rn(gs, D(l - 1, m + 1).dxy);
        gs = __dadd_rn(gs, D(l - 1, m - 1).dxy);
    }
    out[t] = __dadd_rn(gr, __dmul_rn(gs, __dmul_rn(2.0, smooth_gain)));
}

// _warp as a stand-alone operation (vet.morph, vet.py:93-153)
__global__ void __launch_bounds__(256)
vet_warp_kernel(const double *__restrict__ image, const int8_t *__restrict__ mask, const double *__restrict__ disp,
                int nx, int ny, double *__restrict__ out, int8_t *__restrict__ omask, double *__restrict__ grad) {
    const int y = blockIdx.x * 32 + threadIdx.x, x = blockIdx.y * 8 + threadIdx.y;
    if (x >= nx || y >= ny) return;
    const size_t N = (size_t)nx * ny, i = (size_t)x * ny + y;
    const Warped w = warp_pixel(image, mask, nx, ny, x, y, disp[i], disp[N + i]);
    out[i] = w.value;
    omask[i] = w.masked ? 1 : 0;
    if (grad) { grad[i] = w.gx; grad[N + i] = w.gy; }
}

// scipy.ndimage.zoom(a (c,h,w), (1, oh/h, ow/w), order=1, mode="nearest")
__global__ void __launch_bounds__(256)
zoom_kernel(const double *__restrict__ a, int c, int h, int w, int oh, int ow, double *__restrict__ out) {
    const int j = blockIdx.x * 32 + threadIdx.x, i = blockIdx.y * 8 + threadIdx.y;
    if (j >= ow || i >= oh) return;
    const double sy = oh > 1 ? __ddiv_rn((double)(h - 1), (double)(oh - 1)) : 0.0;
    const double sx = ow > 1 ? __ddiv_rn((double)(w - 1), (double)(ow - 1)) : 0.0;
    const double cy = __dmul_rn((double)i, sy), cx = __dmul_rn((double)j, sx);
    const double fy = floor(cy), fx = floor(cx);
    const double ty = __dsub_rn(cy, fy), tx = __dsub_rn(cx, fx);
    const int y0 = min((int)fy, h - 1), y1 = min((int)fy + 1, h - 1);
    const int x0 = min((int)fx, w - 1), x1 = min((int)fx + 1, w - 1);
    const double wy0 = __dsub_rn(1.0, ty), wy1 = __dsub_rn(1.0, wy0);
    const double wx0 = __dsub_rn(1.0, tx), wx1 = __dsub_rn(1.0, wx0);
    for (int k = 0; k < c; k++) {
        const double *p = a + (size_t)k * h * w;
        double t = __dadd_rn(0.0, __dmul_rn(__dmul_rn(p[y0 * w + x0], wy0), wx0));
        t = __dadd_rn(t, __dmul_rn(__dmul_rn(p[y0 * w + x1], wy0), wx1));
        t = __dadd_rn(t, __dmul_rn(__dmul_rn(p[y1 * w + x0], wy1), wx0));
        t = __dadd_rn(t, __dmul_rn(__dmul_rn(p[y1 * w + x1], wy1), wx1));
        out[((size_t)k * oh + i) * ow + j] = t;
    }
}

}  // namespace

// mode 0: out = {residuals, smoothness}; 1: out = gradient (2, xs, ys); 2: out = {residuals,
// smoothness, gradient...} from one pass over the images
static int vet_launch(int mode, const double *sector_disp, const double *templ, const double *input,
                      const int8_t *mask, int xs, int ys, int nx, int ny, float smooth_gain, double *out,
                      cudaStream_t s) {
    B200_REQUIRE(sector_disp && templ && input && mask && out, "bad arguments");
    B200_REQUIRE(xs >= 2 && ys >= 2 && nx >= 1 && ny >= 1, "need at least 2 x 2 sectors");
    if (nx % xs != 0 || ny % ys != 0) {
        // _vet.pyx:345-353
        b200::set_error("Error computing cost function. The number of sectors don't divide the image size");
        return B200_EINVAL;
    }
    VetGeom g;
    g.nx = nx; g.ny = ny; g.xs = xs; g.ys = ys;
    g.xss = nx / xs; g.yss = ny / ys;
    g.i_shift = g.xss / 2; g.j_shift = g.yss / 2;
    g.ncx = xs - 1; g.ncy = ys - 1;
    const int ncells = g.ncx * g.ncy;
    // enough CTAs to cover the chip a few times whatever the sector count (2x2 .. 32x32)
    const int want = b200::num_sms() * 8;
    g.strips = std::max(1, std::min((want + ncells - 1) / ncells, std::max(g.xss / 2, 1)));
    const int nv = mode == 0 ? 1 : (mode == 1 ? 8 : 9);
    b200::Scratch part;
    B200_CUDA(part.alloc(sizeof(double) * (size_t)ncells * g.strips * nv, s));
    dim3 grid(g.strips, ncells);
    const double gain = (double)smooth_gain;  // C float parameter of the reference (:242)
    double *P = (double *)part.p;
    if (mode == 0) vet_eval_kernel<0><<<grid, VET_THREADS, 0, s>>>(sector_disp, templ, input, mask, g, P);
    else if (mode == 1) vet_eval_kernel<1><<<grid, VET_THREADS, 0, s>>>(sector_disp, templ, input, mask, g, P);
    else vet_eval_kernel<2><<<grid, VET_THREADS, 0, s>>>(sector_disp, templ, input, mask, g, P);
    B200_LAUNCH_CHECK();
    if (mode != 1) {
        vet_final_cost_kernel<<<1, VET_THREADS, 0, s>>>(P, ncells * g.strips, nv, mode == 0 ? 0 : 8, sector_disp, g,
                                                      gain, out);
        B200_LAUNCH_CHECK();
    }
    if (mode != 0) {
        vet_final_grad_kernel<<<b200::ceil_div(2 * xs * ys, VET_THREADS), VET_THREADS, 0, s>>>(
            P, nv, sector_disp, g, gain, mode == 1 ? out : out + 2);
        B200_LAUNCH_CHECK();
    }
    return 0;
}

extern "C" int b200_vet_cost(const double *sector_disp, const double *templ, const double *input,
                             const int8_t *mask, int xs, int ys, int nx, int ny, float smooth_gain,
                             int gradient, double *out, void *stream) {
    return vet_launch(gradient ? 1 : 0, sector_disp, templ, input, mask, xs, ys, nx, ny, smooth_gain, out,
                      (cudaStream_t)stream);
}

extern "C" int b200_vet_value_and_gradient(const double *x_host, const double *images, int nframes,
                                           const int8_t *mask, int xs, int ys, int nx, int ny,
                                           float smooth_gain, double *work, double *value_host,
                                           double *gradient_host, void *stream) {
    B200_REQUIRE(x_host && images && mask && work && value_host && gradient_host, "bad arguments");
    B200_REQUIRE(nframes == 2 || nframes == 3, "two or three frames");
    cudaStream_t s = (cudaStream_t)stream;
    const size_t S = (size_t)2 * xs * ys, N = (size_t)nx * ny;
    // work: x (S) | pair 0 {residuals, smoothness, gradient (S)} | pair 1 {...}
    double *dx = work, *r0 = work + S, *r1 = r0 + 2 + S;
    B200_CUDA(cudaMemcpyAsync(dx, x_host, S * sizeof(double), cudaMemcpyHostToDevice, s));
    // vet.py:257-268: (centre, next), then -- three frames -- (previous, centre)
    const double *a = images + (nframes == 3 ? N : 0), *b = a + N;
    if (int rc = vet_launch(2, dx, a, b, mask, xs, ys, nx, ny, smooth_gain, r0, s)) return rc;
    if (nframes == 3)
        if (int rc = vet_launch(2, dx, images, images + N, mask, xs, ys, nx, ny, smooth_gain, r1, s)) return rc;
    // one read-back of everything; the sums over the pairs in the reference's order on the host
    static thread_local std::vector<double> stage;
    stage.resize((size_t)(nframes == 3 ? 2 : 1) * (2 + S));
    B200_CUDA(cudaMemcpyAsync(stage.data(), r0, stage.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
    B200_CUDA(cudaStreamSynchronize(s));
    double residuals = stage[0], smoothness = stage[1];
    const double *g0 = stage.data() + 2;
    if (nframes == 3) {
        const double *p1 = stage.data() + 2 + S;
        residuals += p1[0];    // :263-268
        smoothness += p1[1];
        for (size_t i = 0; i < S; i++) gradient_host[i] = g0[i] + p1[2 + i];  // :288-293
    } else {
        for (size_t i = 0; i < S; i++) gradient_host[i] = g0[i];
    }
    value_host[0] = residuals;
    value_host[1] = smoothness;
    return 0;
}

// One kernel per minimisation level: cleaning (vet.py:507-523), the global `padding` frame and the
// level's divisibility padding (:548-561) from the raw frames.
//   valid(t, pixel)  = user mask clear when a mask is given, else the value is finite
//   image[t]         = the frame where valid, 0 elsewhere; outside the globally padded frame the
//                      nearest edge value of it (numpy.pad "edge")
//   mask             = 1 where any frame is invalid, in the global padding ring, and in the level's
//                      own padding (constant 1)
__global__ void __launch_bounds__(256)
vet_level_kernel(const double *__restrict__ frames, const uint8_t *__restrict__ umask, int T, int m, int n,
                 int gpad, int pi0, int pj0, int M, int N, double *__restrict__ out, int8_t *__restrict__ omask) {
    const int J = blockIdx.x * 32 + threadIdx.x, I = blockIdx.y * 8 + threadIdx.y;
    if (I >= M || J >= N) return;
    const int mg = m + 2 * gpad, ng = n + 2 * gpad;
    const bool inside = I >= pi0 && I < pi0 + mg && J >= pj0 && J < pj0 + ng;
    const int ii = min(max(I - pi0, 0), mg - 1), jj = min(max(J - pj0, 0), ng - 1);
    const int si = ii - gpad, sj = jj - gpad;
    const bool in_src = si >= 0 && si < m && sj >= 0 && sj < n;
    bool any_bad = false;
    for (int t = 0; t < T; t++) {
        double v = 0.0;
        bool bad = true;
        if (in_src) {
            const size_t idx = ((size_t)t * m + si) * n + sj;
            v = frames[idx];
            bad = umask ? umask[idx] != 0 : !isfinite(v);
        }
        any_bad |= bad;
        out[((size_t)t * M + I) * N + J] = bad ? 0.0 : v;
    }
    omask[(size_t)I * N + J] = (!inside || any_bad) ? 1 : 0;
}

extern "C" int b200_vet_level_images(const double *frames, const uint8_t *user_mask, int nframes, int m,
                                     int n, int padding, int pad_i_before, int pad_j_before, int M, int N,
                                     double *images, int8_t *mask, void *stream) {
    B200_REQUIRE(frames && images && mask && nframes >= 1 && m >= 1 && n >= 1 && padding >= 0, "bad arguments");
    B200_REQUIRE(pad_i_before >= 0 && pad_j_before >= 0 && M >= m + 2 * padding + pad_i_before &&
                     N >= n + 2 * padding + pad_j_before, "level frame smaller than the padded input");
    vet_level_kernel<<<dim3(b200::ceil_div(N, 32), b200::ceil_div(M, 8)), dim3(32, 8), 0, (cudaStream_t)stream>>>(
        frames, user_mask, nframes, m, n, padding, pad_i_before, pad_j_before, M, N, images, mask);
    B200_LAUNCH_CHECK();
    return 0;
}

extern "C" int b200_vet_warp(const double *image, const int8_t *mask, const double *displacement, int nx,
                             int ny, double *out, int8_t *out_mask, double *grad, void *stream) {
    B200_REQUIRE(image && mask && displacement && out && out_mask && nx >= 1 && ny >= 1, "bad arguments");
    vet_warp_kernel<<<dim3(b200::ceil_div(ny, 32), b200::ceil_div(nx, 8)), dim3(32, 8), 0, (cudaStream_t)stream>>>(
        image, mask, displacement, nx, ny, out, out_mask, grad);
    B200_LAUNCH_CHECK();
    return 0;
}

extern "C" int b200_zoom_bilinear(const double *a, int c, int h, int w, int oh, int ow, double *out,
                                  void *stream) {
    B200_REQUIRE(a && out && c >= 1 && h >= 1 && w >= 1 && oh >= 1 && ow >= 1, "bad arguments");
    zoom_kernel<<<dim3(b200::ceil_div(ow, 32), b200::ceil_div(oh, 8)), dim3(32, 8), 0, (cudaStream_t)stream>>>(
        a, c, h, w, oh, ow, out);
    B200_LAUNCH_CHECK();
    return 0;
}
