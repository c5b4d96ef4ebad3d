// TEST INFRASTRUCTURE: the per-thread bodies of pysteps_b200/csrc/proesmans.cu compiled as host
// C++ and driven by loops standing in for the CUDA grid (see spline_host.cpp).  The relaxation
// sweep is run wavefront by wavefront with the rows of every wavefront visited in REVERSE order,
// so a mistake in the dependency analysis (t = x + 2y) would show up as a difference from the
// oracle's raster-order sweep.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../pysteps_b200/csrc/proesmans_body.cuh"

namespace {

void consistency(const double *V, int h, int w, double *gamma) {
    const size_t N = (size_t)h * w;
    std::vector<double> rs(2 * (size_t)h);
    std::vector<long long> rc(2 * (size_t)h);
    for (size_t e = 0; e < 2 * N; e++) {
        const size_t q = e % N;
        gamma[e] = pro::cons_pixel(V, h, w, (int)(e / N), (int)(q / w), (int)(q % w));
    }
    for (int r = 2 * h - 1; r >= 0; r--) pro::cons_row_sum(gamma + (size_t)r * w, w, rs[r], rc[r]);
    for (int i = 0; i < 2; i++) {
        const double K = pro::cons_K(rs.data() + (size_t)i * h, rc.data() + (size_t)i * h, h);
        for (size_t q = 0; q < N; q++) gamma[(size_t)i * N + q] = pro::cons_weight(gamma[(size_t)i * N + q], K);
    }
}

void sweep(const double *R, const double *G, const double *gamma, double *V, int h, int w, double lam, int j) {
    const size_t N = (size_t)h * w;
    const double *R1 = R + (size_t)j * N, *R2 = R + (size_t)(1 - j) * N;
    const double *G1 = G + (size_t)(2 * j) * N, *G2 = G + (size_t)(2 * j + 1) * N;
    const double *gam = gamma + (size_t)j * N;
    double *Vj = V + (size_t)(2 * j) * N;
    if (h < 3 || w < 3) return;
    for (int t = pro::sweep_first_t(); t <= pro::sweep_last_t(h, w); t++) {
        int ylo, yhi;
        pro::sweep_rows_of(t, h, w, ylo, yhi);
        for (int y = yhi; y >= ylo; y--) pro::sweep_pixel(R1, R2, G1, G2, gam, Vj, h, w, t - 2 * y, y, lam);
    }
    for (int c = 0; c < 2; c++)
        for (int e = pro::fill_edge_count(h, w) - 1; e >= 0; e--) pro::fill_edge_element(Vj + (size_t)c * N, h, w, e);
}

}  // namespace

extern "C" {

void host_proesmans_scale(const double *in, int64_t count, double lo, double hi, int do_scale, double *out) {
    for (int64_t e = 0; e < count; e++) out[e] = pro::scale_value(in[e], lo, hi, do_scale);
}

// == b200_gaussian_filter
void host_gaussian_filter(const double *in, int h, int w, const double *weights, int radius, double *out) {
    pro::GaussKernel k;
    memset(&k, 0, sizeof(k));
    k.lw = radius;
    for (int i = 0; i < 2 * radius + 1; i++) k.w[i] = weights[i];
    std::vector<double> tmp((size_t)h * w);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) tmp[(size_t)y * w + x] = pro::gauss_line_value(in + x, h, (size_t)w, y, k);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) out[(size_t)y * w + x] = pro::gauss_line_value(tmp.data() + (size_t)y * w, w, 1, x, k);
}

// == b200_proesmans_field
int host_proesmans_field(const double *frames, int m, int n, double lam, int num_iter, int num_levels,
                         double *advfield, double *quality) {
    std::vector<int> hs(num_levels), ws(num_levels);
    hs[0] = m; ws[0] = n;
    for (int l = 1; l < num_levels; l++) { hs[l] = hs[l - 1] / 2; ws[l] = ws[l - 1] / 2; }
    if (hs[num_levels - 1] < 1 || ws[num_levels - 1] < 1) return -1;
    std::vector<std::vector<double>> pyr[2];
    for (int img = 0; img < 2; img++) {
        pyr[img].resize(num_levels);
        pyr[img][0].assign(frames + (size_t)img * m * n, frames + (size_t)(img + 1) * m * n);
        for (int l = 1; l < num_levels; l++) {
            pyr[img][l].resize((size_t)hs[l] * ws[l]);
            for (int y = 0; y < hs[l]; y++)
                for (int x = 0; x < ws[l]; x++)
                    pyr[img][l][(size_t)y * ws[l] + x] = pro::pyr_pixel(pyr[img][l - 1].data(), ws[l - 1], y, x);
        }
    }
    std::vector<double> Vc((size_t)4 * hs[num_levels - 1] * ws[num_levels - 1], 0.0), Vn;
    for (int l = num_levels - 1; l >= 0; l--) {
        const int h = hs[l], w = ws[l];
        const size_t N = (size_t)h * w;
        std::vector<double> R(2 * N), G(4 * N), gamma(2 * N);
        std::copy(pyr[0][l].begin(), pyr[0][l].end(), R.begin());
        std::copy(pyr[1][l].begin(), pyr[1][l].end(), R.begin() + N);
        for (int img = 0; img < 2; img++)
            for (size_t q = 0; q < N; q++)
                pro::grad_pixel(R.data() + (size_t)img * N, h, w, (int)(q / w), (int)(q % w),
                                G[(size_t)(2 * img) * N + q], G[(size_t)(2 * img + 1) * N + q]);
        for (int it = 0; it < num_iter; it++) {
            consistency(Vc.data(), h, w, gamma.data());
            for (int j = 1; j >= 0; j--) sweep(R.data(), G.data(), gamma.data(), Vc.data(), h, w, lam, j);
        }
        if (l > 0) {
            const int hn = hs[l - 1], wn = ws[l - 1];
            Vn.assign((size_t)4 * hn * wn, 0.0);
            for (int c = 0; c < 4; c++)
                for (int yn = 0; yn < hn; yn++)
                    for (int xn = 0; xn < wn; xn++)
                        Vn[(size_t)c * hn * wn + (size_t)yn * wn + xn] =
                            pro::next_level_pixel(Vc.data() + (size_t)c * h * w, h, w, yn, xn);
            Vc.swap(Vn);
        }
    }
    consistency(Vc.data(), m, n, quality);
    memcpy(advfield, Vc.data(), sizeof(double) * 4 * (size_t)m * n);
    return 0;
}

}  // extern "C"
