// quantise_body.cuh -- float32 variant of "scale between 0 and 255" (tracking/lucaskanade.py:144-160,
// feature/shitomasi.py:141-151) as host/device source (see spline_body.cuh), so that the CPU suite
// can run it against NumPy's float32 arithmetic.
#pragma once

#if defined(__CUDACC__)
#define QZ_FN __host__ __device__ __forceinline__
#else
#define QZ_FN inline
#endif

namespace qz {

// (v - im_min) / (im_max - im_min) * 255 in float32 (v, im_min, im_max are float32 values held in
// doubles), or v - im_min when the range is <= 1e-8; returned widened (exact)
QZ_FN double scale_f32(double v, double im_min, double im_max) {
    const float vf = (float)v, lo = (float)im_min, hi = (float)im_max;
#if defined(__CUDA_ARCH__)
    if (__fsub_rn(hi, lo) > 1e-8f) return (double)__fmul_rn(__fdiv_rn(__fsub_rn(vf, lo), __fsub_rn(hi, lo)), 255.0f);
    return (double)__fsub_rn(vf, lo);
#else
    const float range = hi - lo;
    if (range > 1e-8f) {
        const float d = vf - lo;
        const float r = d / range;
        return (double)(r * 255.0f);
    }
    return (double)(vf - lo);
#endif
}

// ---- float64 frames: (v - im_min) / (im_max - im_min) * 255 -> astype(uint8), mostly without the division.
// NumPy's float -> uint8 cast on x86-64: truncation toward zero through int32, low byte kept; out of
// range / NaN -> INT_MIN -> 0.
QZ_FN unsigned char cast_u8(double v) {
    if (!(v > -2147483649.0 && v < 2147483648.0)) return 0;  // cvttsd2si -> INT_MIN -> low byte 0
    return (unsigned char)((int)v & 0xff);
}

#if defined(__CUDA_ARCH__)
#define QZ_SUB(a, b) __dsub_rn(a, b)
#define QZ_ADD(a, b) __dadd_rn(a, b)
#define QZ_MUL(a, b) __dmul_rn(a, b)
#define QZ_DIV(a, b) __ddiv_rn(a, b)
#else  // the host build is compiled with -ffp-contract=off: every operation rounds once
#define QZ_SUB(a, b) ((a) - (b))
#define QZ_ADD(a, b) ((a) + (b))
#define QZ_MUL(a, b) ((a) * (b))
#define QZ_DIV(a, b) ((a) / (b))
#endif

// The product with a precomputed 255 / range is within 4 * 2^-53 * 256 = 1.2e-13 of the reference's
// two-rounding result, so its truncation is the reference's unless it lies within 1e-9 of an integer --
// only then, and outside (0, 255.5), the division itself is evaluated.
struct ScaleF64 {
    double im_min, im_max, range, r255;
    bool wide;  // range > 1e-8 (otherwise the reference returns v - im_min)
    QZ_FN void init(double lo, double hi) {
        im_min = lo;
        im_max = hi;
        range = QZ_SUB(hi, lo);
        wide = range > 1e-8;
        r255 = QZ_DIV(255.0, range);
    }
    QZ_FN unsigned char exact(double val) const {
        const double q = wide ? QZ_MUL(QZ_DIV(QZ_SUB(val, im_min), range), 255.0) : QZ_SUB(val, im_min);
        return cast_u8(q);
    }
    QZ_FN unsigned char operator()(double val) const {
        const double num = QZ_SUB(val, im_min);
        if (num == 0.0) return 0;  // 0 / range * 255, or 0 itself
        if (wide) {
            const double qa = QZ_MUL(num, r255);
            const double magic = 6755399441055744.0;          // 2^52 + 2^51: nearest integer in the low word
            const double t = QZ_ADD(qa, magic);
            const double nearest = QZ_SUB(t, magic);
            const double d = QZ_SUB(qa, nearest);                // qa - nearest, in [-0.5, 0.5]
            if (qa > 0.0 && qa < 255.5 && (d > 1e-9 || d < -1e-9)) return (unsigned char)((int)nearest - (d < 0.0 ? 1 : 0));
        }
        return exact(val);
    }
};

#undef QZ_SUB
#undef QZ_ADD
#undef QZ_MUL
#undef QZ_DIV

}  // namespace qz
