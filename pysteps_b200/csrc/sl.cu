// sl.cu -- semi-Lagrangian backward-trajectory extrapolation (sm_100a).
//
// Replaces the leadtime loop of pysteps/extrapolation/semilagrangian.py:181-232:
// per leadtime, two bilinear gathers of the advection field along the
// trajectory (interpolate_motion, :181-198) and one bilinear warp of the
// precipitation field (:221-232).  In the reference each of these is a
// full-array scipy.ndimage.map_coordinates call plus NumPy temporaries.  Here a
// pixel's whole trajectory is independent of every other pixel's, so ONE kernel
// carries (displacement, velocity increment) of a pixel in float64 registers
// through all T leadtimes: the only HBM traffic is the compulsory read of the
// fields (L2 resident afterwards) and the T output planes.
//
// Arithmetic contract (pinned against scipy 1.18.1, see oracle/sl_oracle.c):
// every float64 operation of the reference is issued in the same order with
// round-to-nearest and NO fused multiply-add, so trajectories -- and therefore
// the integer tap indices -- are bit-identical to the CPU path.
#include "common.cuh"

namespace {

constexpr int SL_MAX_T = 32;  // leadtimes per launch; longer sequences are chunked
constexpr int SL_BX = 32, SL_BY = 8;

enum { SL_INIT_FRESH = 0, SL_INIT_PREV = 1, SL_INIT_RESUME = 2 };

struct SLParams {
    const void *Vi;         // (m,n) interleaved (vx,vy), field dtype
    const void *precip;     // (m,n) or null
    const double *xy;       // (2,m,n) or null -> pixel grid
    const double *disp_in;  // (2,m,n) or null
    const double *vinc_in;  // (2,m,n), RESUME only
    double *disp_out;       // (2,m,n) or null
    double *vinc_out;       // (2,m,n) or null
    void *out;              // (T,m,n) planes of this chunk
    int m, n, T, n_iter, ti_offset, init_mode, mode, has_prev;
    double vts, cval;
    double td[SL_MAX_T];
};

template <typename F> struct Vec2;
template <> struct Vec2<float> { using type = float2; };
template <> struct Vec2<double> { using type = double2; };

// one axis of scipy's order-1 footprint in mode="nearest": taps floor(c),
// floor(c)+1 each clamped to [0, L-1]; weights w0 = 1 - t, w1 = 1 - w0.
__device__ __forceinline__ void axis_nearest(double c, int L, int &i0, int &i1, double &w0,
                                             double &w1) {
    const double f = floor(c);
    const double t = __dsub_rn(c, f);
    w0 = __dsub_rn(1.0, t);
    w1 = __dsub_rn(1.0, w0);
    // (npy_intp)floor(c) on x86-64: out-of-range / non-finite -> INT64_MIN (both taps 0)
    double fc = (fabs(f) < 9223372036854775808.0) ? f : -1.0;
    fc = fmin(fmax(fc, -1.0), (double)L);
    const int i = (int)fc;
    i0 = min(max(i, 0), L - 1);
    i1 = min(max(i + 1, 0), L - 1);
}

// sum_{taps} ((a * wy) * wx), left to right from 0.0 (scipy accumulation order)
__device__ __forceinline__ double bilin(double a00, double a01, double a10, double a11,
                                        double wy0, double wy1, double wx0, double wx1) {
    double t = __dadd_rn(0.0, __dmul_rn(__dmul_rn(a00, wy0), wx0));
    t = __dadd_rn(t, __dmul_rn(__dmul_rn(a01, wy0), wx1));
    t = __dadd_rn(t, __dmul_rn(__dmul_rn(a10, wy1), wx0));
    t = __dadd_rn(t, __dmul_rn(__dmul_rn(a11, wy1), wx1));
    return t;
}

// interpolate_motion (semilagrangian.py:181-198) for one pixel
template <typename F>
__device__ __forceinline__ void sample_velocity(const typename Vec2<F>::type *__restrict__ Vi,
                                                int m, int n, double cy, double cx, double scale,
                                                int n_iter, double &vx, double &vy) {
    int y0, y1, x0, x1;
    double wy0, wy1, wx0, wx1;
    axis_nearest(cy, m, y0, y1, wy0, wy1);
    axis_nearest(cx, n, x0, x1, wx0, wx1);
    const typename Vec2<F>::type *r0 = Vi + (size_t)y0 * n;
    const typename Vec2<F>::type *r1 = Vi + (size_t)y1 * n;
    const auto a00 = __ldg(r0 + x0), a01 = __ldg(r0 + x1);
    const auto a10 = __ldg(r1 + x0), a11 = __ldg(r1 + x1);
    vx = bilin((double)a00.x, (double)a01.x, (double)a10.x, (double)a11.x, wy0, wy1, wx0, wx1);
    vy = bilin((double)a00.y, (double)a01.y, (double)a10.y, (double)a11.y, wy0, wy1, wx0, wx1);
    if (sizeof(F) == 4) {
        // float32 velocity: map_coordinates returns the input dtype, so the reference
        // stores the sampled increment rounded to float32 (:192-193)
        vx = (double)__double2float_rn(vx);
        vy = (double)__double2float_rn(vy);
    }
    if (n_iter > 1) {  // :195-196
        vx = __ddiv_rn(vx, (double)n_iter);
        vy = __ddiv_rn(vy, (double)n_iter);
    }
    vx = __dmul_rn(vx, scale);  // :198
    vy = __dmul_rn(vy, scale);
}

// map_coordinates(precip, order=1, mode, cval) for one pixel (:221-232)
template <typename F>
__device__ __forceinline__ double sample_precip(const F *__restrict__ P, int m, int n, double cy,
                                                double cx, int mode, double cval) {
    int y0, y1, x0, x1;
    double wy0, wy1, wx0, wx1;
    if (mode == B200_MODE_CONSTANT) {
        if (!(cy >= 0.0 && cy <= (double)(m - 1) && cx >= 0.0 && cx <= (double)(n - 1)))
            return cval;
        const double fy = floor(cy), fx = floor(cx);
        const double ty = __dsub_rn(cy, fy), tx = __dsub_rn(cx, fx);
        wy0 = __dsub_rn(1.0, ty); wy1 = __dsub_rn(1.0, wy0);
        wx0 = __dsub_rn(1.0, tx); wx1 = __dsub_rn(1.0, wx0);
        y0 = (int)fy; x0 = (int)fx;
        // the tap one past the end (only when c == L-1) is mirrored and still read
        y1 = (y0 + 1 < m) ? y0 + 1 : (m > 1 ? m - 2 : 0);
        x1 = (x0 + 1 < n) ? x0 + 1 : (n > 1 ? n - 2 : 0);
    } else {
        axis_nearest(cy, m, y0, y1, wy0, wy1);
        axis_nearest(cx, n, x0, x1, wx0, wx1);
    }
    const F *r0 = P + (size_t)y0 * n;
    const F *r1 = P + (size_t)y1 * n;
    return bilin((double)__ldg(r0 + x0), (double)__ldg(r0 + x1), (double)__ldg(r1 + x0),
                 (double)__ldg(r1 + x1), wy0, wy1, wx0, wx1);
}

template <typename F> __device__ __forceinline__ F from_double(double v);
template <> __device__ __forceinline__ float from_double<float>(double v) { return __double2float_rn(v); }
template <> __device__ __forceinline__ double from_double<double>(double v) { return v; }

template <typename FV, typename F>
__global__ void __launch_bounds__(SL_BX *SL_BY)
sl_multistep_kernel(const __grid_constant__ SLParams p) {
    using V2 = typename Vec2<FV>::type;
    const int x = blockIdx.x * SL_BX + threadIdx.x;
    const int y = blockIdx.y * SL_BY + threadIdx.y;
    if (x >= p.n || y >= p.m) return;
    const int m = p.m, n = p.n;
    const size_t N = (size_t)m * n;
    const size_t idx = (size_t)y * n + x;
    const V2 *__restrict__ Vi = (const V2 *)p.Vi;
    const F *__restrict__ P = (const F *)p.precip;
    F *__restrict__ out = (F *)p.out;

    double gx, gy;  // xy_coords of this pixel (:174-179)
    if (p.xy) {
        gx = p.xy[idx];
        gy = p.xy[N + idx];
    } else {
        gx = (double)x;
        gy = (double)y;
    }

    double dx, dy, ux, uy;  // displacement, velocity increment
    if (p.init_mode == SL_INIT_FRESH) {
        // :201-203  displacement = 0 ; velocity_inc = V * tdiff[0] / vel_timestep
        dx = 0.0; dy = 0.0;
        const V2 v = Vi[idx];
        ux = __ddiv_rn(__dmul_rn((double)v.x, p.td[0]), p.vts);
        uy = __ddiv_rn(__dmul_rn((double)v.y, p.td[0]), p.vts);
    } else if (p.init_mode == SL_INIT_PREV) {
        // :205-207
        dx = p.disp_in[idx]; dy = p.disp_in[N + idx];
        sample_velocity<FV>(Vi, m, n, __dadd_rn(gy, dy), __dadd_rn(gx, dx),
                           __ddiv_rn(p.td[0], p.vts), p.n_iter, ux, uy);
    } else {
        dx = p.disp_in[idx]; dy = p.disp_in[N + idx];
        ux = p.vinc_in[idx]; uy = p.vinc_in[N + idx];
    }

    for (int ti = 0; ti < p.T; ti++) {
        const double scale = __ddiv_rn(p.td[ti], p.vts);  // td / vel_timestep (:198)
        if (p.n_iter > 0) {
            for (int k = 0; k < p.n_iter; k++) {  // :211-214
                const double hx = __dsub_rn(dx, __ddiv_rn(ux, 2.0));
                const double hy = __dsub_rn(dy, __ddiv_rn(uy, 2.0));
                sample_velocity<FV>(Vi, m, n, __dadd_rn(gy, hy), __dadd_rn(gx, hx), scale,
                                   p.n_iter, ux, uy);
                dx = __dsub_rn(dx, ux);
                dy = __dsub_rn(dy, uy);
                sample_velocity<FV>(Vi, m, n, __dadd_rn(gy, dy), __dadd_rn(gx, dx), scale,
                                   p.n_iter, ux, uy);
            }
        } else {  // :215-219
            if (ti + p.ti_offset > 0 || p.has_prev)
                sample_velocity<FV>(Vi, m, n, __dadd_rn(gy, dy), __dadd_rn(gx, dx), scale,
                                   p.n_iter, ux, uy);
            dx = __dsub_rn(dx, ux);
            dy = __dsub_rn(dy, uy);
        }
        if (P) {
            const double v = sample_precip<F>(P, m, n, __dadd_rn(gy, dy), __dadd_rn(gx, dx),
                                              p.mode, p.cval);
            out[(size_t)ti * N + idx] = from_double<F>(v);
        }
    }
    if (p.disp_out) {
        p.disp_out[idx] = dx;
        p.disp_out[N + idx] = dy;
    }
    if (p.vinc_out) {
        p.vinc_out[idx] = ux;
        p.vinc_out[N + idx] = uy;
    }
}

// planar (2,m,n) -> interleaved (m,n){x,y}; one pass, 16 B per thread-iteration
template <typename F>
__global__ void __launch_bounds__(256)
interleave_kernel(const F *__restrict__ V, typename Vec2<F>::type *__restrict__ Vi, size_t N) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
        typename Vec2<F>::type v;
        v.x = __ldg(V + i);
        v.y = __ldg(V + N + i);
        Vi[i] = v;
    }
}

template <typename FV, typename F>
int sl_run(const void *precip, const void *velocity, const double *xy, const double *disp_prev,
           const double *tdiff, int T, double vts, int n_iter, double outval, int mode,
           int layout, int m, int n, void *out, double *disp_out, cudaStream_t stream) {
    using V2 = typename Vec2<FV>::type;
    const size_t N = (size_t)m * n;
    b200::Scratch vi, st_disp, st_vinc;
    const void *vi_ptr = velocity;
    if (layout == B200_LAYOUT_PLANAR) {
        B200_CUDA(vi.alloc(N * sizeof(V2), stream));
        vi_ptr = vi.p;
        const int blocks = (int)std::min<size_t>((N + 255) / 256, (size_t)b200::num_sms() * 16);
        interleave_kernel<FV><<<blocks, 256, 0, stream>>>((const FV *)velocity, (V2 *)vi.p, N);
        B200_LAUNCH_CHECK();
    }
    const int nchunks = (T + SL_MAX_T - 1) / SL_MAX_T;
    if (nchunks > 1) {
        B200_CUDA(st_disp.alloc(2 * N * sizeof(double), stream));
        B200_CUDA(st_vinc.alloc(2 * N * sizeof(double), stream));
    }
    dim3 block(SL_BX, SL_BY);
    dim3 grid(b200::ceil_div(n, SL_BX), b200::ceil_div(m, SL_BY));
    for (int c = 0; c < nchunks; c++) {
        SLParams p;
        memset(&p, 0, sizeof(p));
        p.Vi = vi_ptr;
        p.precip = precip;
        p.xy = xy;
        p.m = m; p.n = n;
        p.n_iter = n_iter;
        p.mode = mode;
        p.vts = vts;
        p.cval = outval;
        p.has_prev = disp_prev != nullptr;
        p.ti_offset = c * SL_MAX_T;
        p.T = std::min(SL_MAX_T, T - p.ti_offset);
        for (int i = 0; i < p.T; i++) p.td[i] = tdiff[p.ti_offset + i];
        if (c == 0) {
            p.init_mode = disp_prev ? SL_INIT_PREV : SL_INIT_FRESH;
            p.disp_in = disp_prev;
        } else {
            p.init_mode = SL_INIT_RESUME;
            p.disp_in = (const double *)st_disp.p;
            p.vinc_in = (const double *)st_vinc.p;
        }
        const bool last = (c == nchunks - 1);
        p.disp_out = last ? disp_out : (double *)st_disp.p;
        p.vinc_out = last ? nullptr : (double *)st_vinc.p;
        p.out = precip ? (void *)((F *)out + (size_t)p.ti_offset * N) : nullptr;
        sl_multistep_kernel<FV, F><<<grid, block, 0, stream>>>(p);
        B200_LAUNCH_CHECK();
    }
    return 0;
}

}  // namespace

extern "C" int b200_sl_extrapolate(const void *precip, const void *velocity,
                                   const double *xy_coords, const double *disp_prev,
                                   const double *tdiff, int T, double vel_timestep, int n_iter,
                                   double outval, int mode, int velocity_dtype, int velocity_layout,
                                   int precip_dtype, int m, int n, void *out, double *disp_out,
                                   void *stream) {
    B200_REQUIRE(velocity_layout == B200_LAYOUT_PLANAR || velocity_layout == B200_LAYOUT_INTERLEAVED,
                 "unknown velocity layout");
    B200_REQUIRE(velocity != nullptr, "velocity is NULL");
    B200_REQUIRE(tdiff != nullptr && T >= 1, "need at least one timestep");
    B200_REQUIRE(m >= 1 && n >= 1, "empty grid");
    B200_REQUIRE(n_iter >= 0, "n_iter must be >= 0");
    B200_REQUIRE(mode == B200_MODE_CONSTANT || mode == B200_MODE_NEAREST, "unsupported mode");
    B200_REQUIRE((precip == nullptr) == (out == nullptr), "precip and out must both be given or both NULL");
    B200_REQUIRE(precip != nullptr || disp_out != nullptr, "nothing to compute");
    cudaStream_t s = (cudaStream_t)stream;
#define SL_DISPATCH(FV, FP)                                                                  \
    return sl_run<FV, FP>(precip, velocity, xy_coords, disp_prev, tdiff, T, vel_timestep, n_iter, \
                          outval, mode, velocity_layout, m, n, out, disp_out, s)
    if (velocity_dtype == B200_F32 && precip_dtype == B200_F32) SL_DISPATCH(float, float);
    if (velocity_dtype == B200_F32 && precip_dtype == B200_F64) SL_DISPATCH(float, double);
    if (velocity_dtype == B200_F64 && precip_dtype == B200_F32) SL_DISPATCH(double, float);
    if (velocity_dtype == B200_F64 && precip_dtype == B200_F64) SL_DISPATCH(double, double);
#undef SL_DISPATCH
    b200::set_error("unknown field dtypes %d / %d", velocity_dtype, precip_dtype);
    return B200_EINVAL;
}

extern "C" int b200_sl_extrapolate_host(const void *precip, const void *velocity,
                                        const double *xy_coords, const double *disp_prev,
                                        const double *tdiff, int T, double vel_timestep,
                                        int n_iter, double outval, int mode, int velocity_dtype,
                                        int precip_dtype, int m, int n, void *out,
                                        double *disp_out) {
    B200_REQUIRE(velocity_dtype == B200_F32 || velocity_dtype == B200_F64, "unknown velocity dtype");
    B200_REQUIRE(precip_dtype == B200_F32 || precip_dtype == B200_F64, "unknown precip dtype");
    B200_REQUIRE(velocity != nullptr && m >= 1 && n >= 1 && T >= 1, "bad arguments");
    const size_t N = (size_t)m * n;
    const size_t fs = precip_dtype == B200_F32 ? 4 : 8;
    const size_t vs = velocity_dtype == B200_F32 ? 4 : 8;
    cudaStream_t s = nullptr;
    b200::Scratch dP, dV, dXY, dDP, dOut, dDO;
    B200_CUDA(dV.alloc(2 * N * vs, s));
    B200_CUDA(cudaMemcpyAsync(dV.p, velocity, 2 * N * vs, cudaMemcpyHostToDevice, s));
    if (precip) {
        B200_CUDA(dP.alloc(N * fs, s));
        B200_CUDA(cudaMemcpyAsync(dP.p, precip, N * fs, cudaMemcpyHostToDevice, s));
        B200_CUDA(dOut.alloc((size_t)T * N * fs, s));
    }
    if (xy_coords) {
        B200_CUDA(dXY.alloc(2 * N * 8, s));
        B200_CUDA(cudaMemcpyAsync(dXY.p, xy_coords, 2 * N * 8, cudaMemcpyHostToDevice, s));
    }
    if (disp_prev) {
        B200_CUDA(dDP.alloc(2 * N * 8, s));
        B200_CUDA(cudaMemcpyAsync(dDP.p, disp_prev, 2 * N * 8, cudaMemcpyHostToDevice, s));
    }
    if (disp_out) B200_CUDA(dDO.alloc(2 * N * 8, s));
    int rc = b200_sl_extrapolate(precip ? dP.p : nullptr, dV.p, xy_coords ? (const double *)dXY.p : nullptr,
                                 disp_prev ? (const double *)dDP.p : nullptr, tdiff, T, vel_timestep,
                                 n_iter, outval, mode, velocity_dtype, B200_LAYOUT_PLANAR, precip_dtype,
                                 m, n, precip ? dOut.p : nullptr,
                                 disp_out ? (double *)dDO.p : nullptr, s);
    if (rc) return rc;
    if (precip && out)
        B200_CUDA(cudaMemcpyAsync(out, dOut.p, (size_t)T * N * fs, cudaMemcpyDeviceToHost, s));
    if (disp_out)
        B200_CUDA(cudaMemcpyAsync(disp_out, dDO.p, 2 * N * 8, cudaMemcpyDeviceToHost, s));
    B200_CUDA(cudaStreamSynchronize(s));
    return 0;
}

extern "C" int b200_sl_interleave_velocity(const void *velocity, int velocity_dtype, int m, int n,
                                           void *out, void *stream) {
    B200_REQUIRE(velocity != nullptr && out != nullptr && m >= 1 && n >= 1, "bad arguments");
    const size_t N = (size_t)m * n;
    const int blocks = (int)std::min<size_t>((N + 255) / 256, (size_t)b200::num_sms() * 16);
    cudaStream_t s = (cudaStream_t)stream;
    if (velocity_dtype == B200_F32)
        interleave_kernel<float><<<blocks, 256, 0, s>>>((const float *)velocity, (float2 *)out, N);
    else if (velocity_dtype == B200_F64)
        interleave_kernel<double><<<blocks, 256, 0, s>>>((const double *)velocity, (double2 *)out, N);
    else {
        b200::set_error("unknown velocity dtype %d", velocity_dtype);
        return B200_EINVAL;
    }
    B200_LAUNCH_CHECK();
    return 0;
}
