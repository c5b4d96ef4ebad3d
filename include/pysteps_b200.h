/*
 * pysteps_b200.h -- C ABI of libpysteps_b200.so (hand-written sm_100a CUDA).
 *
 * Drop-in boundary for the advection hot path of pySTEPS/pysteps (reference
 * v1.21.3).  pysteps is pure Python; the callables it exposes through
 *   pysteps/extrapolation/interface.py:107-145  (_extrapolation_methods / get_method)
 *   pysteps/motion/interface.py:36-111          (_methods / get_method)
 * are mirrored in Python by pysteps_b200.{extrapolation,motion}; every array
 * operation behind them is one of the entry points below, reached through
 * ctypes.  All functions return 0 on success or a non-zero code (a cudaError_t,
 * or a B200_E* code); b200_last_error() returns the message of the last
 * failure on the calling thread.
 *
 * Conventions
 *   - "device" pointers are CUDA device pointers on the current device,
 *     "host" pointers are ordinary host memory.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *     Device-pointer entry points only ENQUEUE work; they never synchronise.
 *   - images are row-major (m rows, n columns); vector fields are planar
 *     (2, m, n) with [0] = x / column component, [1] = y / row component, as in
 *     pysteps/motion/interface.py:11-19.
 *   - field dtype codes: B200_F32 / B200_F64 (storage of precip, velocity and
 *     outputs).  Trajectories (displacement) are always float64.
 */
#ifndef PYSTEPS_B200_H
#define PYSTEPS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_F32 0
#define B200_F64 1

#define B200_MODE_CONSTANT 0 /* scipy map_coordinates mode="constant" */
#define B200_MODE_NEAREST 1  /* scipy map_coordinates mode="nearest"  */

#define B200_LAYOUT_PLANAR 0      /* velocity (2,m,n) as in pysteps */
#define B200_LAYOUT_INTERLEAVED 1 /* velocity (m,n,2): what the kernels read */

#define B200_EINVAL 100001  /* bad argument */
#define B200_ENOTSUP 100002 /* unsupported option */

int b200_version(void);
const char *b200_last_error(void);
/* number of kernels this library has launched in this process (all threads) */
long long b200_launch_count(void);
/* number of SMs / device name of the current device (diagnostics) */
int b200_device_info(int *sm_count, int *cc_major, int *cc_minor, char *name, int name_len);

/* ------------------------------------------------------------------------
 * Semi-Lagrangian extrapolation
 * replaces pysteps/extrapolation/semilagrangian.py:181-232 (the leadtime loop:
 * interpolate_motion + map_coordinates warp), interp_order == 1.
 *
 *   precip      device (m,n) of precip_dtype, or NULL (displacement only)
 *   velocity    device (2,m,n) [B200_LAYOUT_PLANAR] or (m,n,2)
 *               [B200_LAYOUT_INTERLEAVED] of velocity_dtype
 *   xy_coords   device (2,m,n) float64, or NULL for the default pixel grid
 *               (semilagrangian.py:174-179)
 *   disp_prev   device (2,m,n) float64 or NULL (semilagrangian.py:200-207)
 *   tdiff       HOST array of T timestep differences (semilagrangian.py:165)
 *   vel_timestep, n_iter, outval, mode: as in the reference
 *   out         device (T,m,n) of precip_dtype, or NULL when precip is NULL
 *   disp_out    device (2,m,n) float64 or NULL
 * All arithmetic is float64 in the reference's operation order, whatever the
 * storage dtypes (scipy converts each tap to double too), so results are
 * bit-identical to the reference run on arrays of the same dtypes: trajectory,
 * integer tap indices and values; float32 outputs are the float64 value
 * rounded once, as scipy does.
 * ---------------------------------------------------------------------- */
int b200_sl_extrapolate(const void *precip, const void *velocity,
                        const double *xy_coords, const double *disp_prev,
                        const double *tdiff, int T, double vel_timestep,
                        int n_iter, double outval, int mode, int velocity_dtype,
                        int velocity_layout, int precip_dtype, int m, int n, void *out,
                        double *disp_out, void *stream);

/* planar (2,m,n) -> interleaved (m,n,2) copy of an advection field, so that a
 * caller reusing one field for many calls pays the re-layout once
 * (b200_sl_extrapolate does it internally for B200_LAYOUT_PLANAR). */
int b200_sl_interleave_velocity(const void *velocity, int velocity_dtype, int m, int n,
                                void *out, void *stream);

/* Same operation on HOST buffers: allocates device scratch from the stream
 * ordered pool, copies in, runs, copies out and synchronises.  This is the
 * call a non-Python binding (cgo / JNI / plain C) would make. */
int b200_sl_extrapolate_host(const void *precip, const void *velocity,
                             const double *xy_coords, const double *disp_prev,
                             const double *tdiff, int T, double vel_timestep,
                             int n_iter, double outval, int mode, int velocity_dtype,
                             int precip_dtype, int m, int n, void *out, double *disp_out);

/* Field statistics used by the input validation of the reference
 * (semilagrangian.py:112-123, 171-172).  `stats` is a device array of 4
 * doubles: [0] number of non-finite elements, [1] np.nanmin, [2] np.nanmax
 * (NaN when every element is NaN), [3] number of NaN elements. */
int b200_field_stats(const void *a, int field_dtype, int64_t count, double *stats,
                     void *stream);

/* dtype conversion on device (float64 <-> float32), count elements */
int b200_convert(const void *src, int src_dtype, void *dst, int dst_dtype,
                 int64_t count, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PYSTEPS_B200_H */
