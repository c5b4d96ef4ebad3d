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

}  // namespace qz
