// TEST INFRASTRUCTURE: the per-thread bodies of pysteps_b200/csrc/spline.cu compiled as plain host
// C++ (g++ -O2 -ffp-contract=off) and driven by loops standing in for the CUDA grid, so that the
// CPU suite can execute the kernels' arithmetic and index logic bit for bit against the oracle
// (tests/test_kernel_bodies.py), the transposes around the row pass included (block by block,
// the two phases of every thread on either side of the barrier).  Not covered on the CPU: anything
// about the GPU execution itself.
#include <stdint.h>
#include <string.h>

#include "../../pysteps_b200/csrc/spline_body.cuh"

extern "C" {

int host_spline_pad(int order, int mode) { return (order > 1 && mode == B200_MODE_NEAREST) ? spl::NPAD : 0; }

// transpose_kernel<<<dim3(ceil(C/32), ceil(R/32)), dim3(32, 8)>>>(in, out, R, C)
void host_transpose(const double *in, double *out, int R, int C) {
    double tile[32][33];
    for (int gy = 0; gy < (R + 31) / 32; gy++)
        for (int gx = 0; gx < (C + 31) / 32; gx++) {
            for (int ty = 0; ty < 8; ty++)
                for (int tx = 0; tx < 32; tx++) spl::transpose_load(tile, in, R, C, gx * 32, gy * 32, tx, ty);
            /* __syncthreads() */
            for (int ty = 0; ty < 8; ty++)
                for (int tx = 0; tx < 32; tx++) spl::transpose_store(tile, out, R, C, gx * 32, gy * 32, tx, ty);
        }
}

// == b200_spline_prepare
void host_spline_prepare(const void *precip, int precip_dtype, int m, int n, int order, int mode,
                         const double *stats, int zero_fill, const double *poles, const double *zpow_axis0,
                         const double *zpow_axis1, double *coeffs, double *mask_min, double *mask_fin) {
    const int pad = host_spline_pad(order, mode);
    const int M = m + 2 * pad, N = n + 2 * pad;
    const size_t total = (size_t)M * N;
    const int want_masks = order > 1;
    for (size_t e = 0; e < total; e++) {
        if (precip_dtype == B200_F32)
            spl::prepare_element<float>(e, (const float *)precip, m, n, pad, stats, zero_fill, want_masks, coeffs,
                                        mask_min, mask_fin);
        else
            spl::prepare_element<double>(e, (const double *)precip, m, n, pad, stats, zero_fill, want_masks, coeffs,
                                         mask_min, mask_fin);
    }
    if (order <= 1) return;
    spl::FilterParams f0;
    memset(&f0, 0, sizeof(f0));
    f0.npoles = order / 2;
    f0.reflect = mode == B200_MODE_NEAREST;
    f0.gain = 1.0;
    for (int k = 0; k < f0.npoles; k++) {
        f0.z[k] = poles[k];
        f0.gain *= (1.0 - poles[k]) * (1.0 - 1.0 / poles[k]);
    }
    spl::FilterParams f1 = f0;
    for (int k = 0; k < f0.npoles; k++) {
        f0.zpow[k] = zpow_axis0[k];
        f1.zpow[k] = zpow_axis1[k];
    }
    // the launch sequence of b200_spline_prepare: columns, transpose, columns, transpose back
    for (int j = 0; j < N; j++) spl::filter_line(coeffs + j, M, (size_t)N, f0);
    double *tr = new double[total];
    host_transpose(coeffs, tr, M, N);
    for (int j = 0; j < M; j++) spl::filter_line(tr + j, N, (size_t)M, f1);
    host_transpose(tr, coeffs, N, M);
    delete[] tr;
}

// == b200_spline_sample
void host_spline_sample(const double *coeffs, int m, int n, int order, int mode, const double *xy,
                        const double *disp_steps, int T, int row_begin, int row_count, double outval,
                        const double *mask_min, const double *mask_fin, const double *stats, int out_dtype,
                        void *out) {
    spl::SampleParams p;
    memset(&p, 0, sizeof(p));
    p.coeffs = coeffs; p.xy = xy; p.disp = disp_steps;
    p.mask_min = mask_min; p.mask_fin = mask_fin; p.stats = stats;
    p.out = out;
    p.m = m; p.n = n; p.order = order; p.mode = mode; p.T = T;
    p.pad = host_spline_pad(order, mode);
    p.row0 = row_begin; p.rows = row_count;
    p.cval = outval;
    for (int t = 0; t < T; t++)
        for (int yl = 0; yl < row_count; yl++)
            for (int x = 0; x < n; x++) {
                const double v = spl::sample_pixel(p, x, yl, t);
                const size_t o = (size_t)t * row_count * n + (size_t)yl * n + x;
                if (out_dtype == B200_F32) ((float *)out)[o] = (float)v;
                else ((double *)out)[o] = v;
            }
}

}  // extern "C"
