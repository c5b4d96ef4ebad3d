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

/* Same for the output rows [row_begin, row_begin + row_count) only: precip, velocity and
 * xy_coords are full (m,n) frames, while disp_prev / out / disp_out are band shaped
 * ((2,row_count,n), (T,row_count,n), (2,row_count,n)).  A pixel's trajectory reads the fields
 * anywhere but writes only its own pixel, so a frame is partitioned over GPUs by output
 * bands with replicated inputs and no halo exchange (SURVEY.md section 8e, config[4]). */
int b200_sl_extrapolate_rows(const void *precip, const void *velocity,
                             const double *xy_coords, const double *disp_prev,
                             const double *tdiff, int T, double vel_timestep,
                             int n_iter, double outval, int mode, int velocity_dtype,
                             int velocity_layout, int precip_dtype, int m, int n,
                             int row_begin, int row_count, void *out, double *disp_out,
                             void *stream);

/* OPT-IN float32-tap variant of b200_sl_extrapolate_rows (no counterpart in the reference, whose
 * arithmetic is float64 throughout; the north star allows "a stated float32 tolerance (bit-exact
 * for the integer displacement indices)"): the trajectory stays float64, the fields are sampled from
 * float32 copies.  A per-pixel error bound certifies that every sample floors to the same tap indices
 * as the exact kernel; pixels that cannot be certified (near a cell boundary, the border, outside,
 * non-finite) are recomputed by the exact code inside the same launch.  Values: within float32
 * rounding of b200_sl_extrapolate_rows.  Restrictions: n_iter = 1, pixel-grid coordinates, a
 * precipitation field, T <= 32.  fallback_count (device uint64, optional, accumulated not reset):
 * number of recomputed pixels. */
int b200_sl_extrapolate_rows_f32(const void *precip, const void *velocity, const double *disp_prev,
                                 const double *tdiff, int T, double vel_timestep, double outval, int mode,
                                 int velocity_dtype, int velocity_layout, int precip_dtype, int m, int n,
                                 int row_begin, int row_count, void *out, double *disp_out,
                                 unsigned long long *fallback_count, void *stream);

/* Displacement field after EVERY leadtime, disp_steps (T, 2, row_count, n) float64: the
 * trajectory part of b200_sl_extrapolate_rows alone (semilagrangian.py:201-219), for samplers
 * other than the built-in order-1 warp (interp_order 0 and 2..5 below).  Same arithmetic as the fused
 * loop: the kernel is resumed once per leadtime. */
int b200_sl_trajectories(const void *velocity, const double *xy_coords, const double *disp_prev,
                         const double *tdiff, int T, double vel_timestep, int n_iter,
                         int velocity_dtype, int velocity_layout, int m, int n, int row_begin,
                         int row_count, double *disp_steps, void *stream);

/* interp_order 0 and 2..5 (semilagrangian.py:144-157,224-253; scipy.ndimage.map_coordinates with
 * prefilter).  b200_spline_prepare: coeffs (m+2p, n+2p) float64, p = 12 for order >= 2 with
 * B200_MODE_NEAREST and 0 otherwise = the field (non-finite values zeroed when zero_fill) edge
 * padded and, for order >= 2, run through scipy's B-spline prefilter.  poles (HOST array, order/2
 * entries) are the filter poles -- the doubles nearest to the exact values, e.g. sqrt(3)-2 for
 * order 3 -- and zpow_axis0/1 (HOST arrays) pole^(L-1) for MODE_CONSTANT ("mirror" boundary) or
 * pole^L for MODE_NEAREST ("reflect"), L the padded length of the axis, evaluated by the caller
 * with the host's pow so that it is the libm value scipy uses.  For order >= 2 also mask_min =
 * (precip > nanmin) and mask_finite as float64 0/1 (m, n).  stats = b200_field_stats of precip.
 * b200_spline_sample: out (T, row_count, n) of out_dtype, pixel (y, x) of leadtime t sampled at
 * xy + disp_steps[t] -- (order+1)^2 taps (mirrored / clamped) or the nearest tap (order 0) -- then,
 * for order >= 2, reset to nanmin / NaN where the order-1 warps of the masks fall below 0.5. */
int b200_spline_prepare(const void *precip, int precip_dtype, int m, int n, int order, int mode,
                        const double *stats, int zero_fill, const double *poles,
                        const double *zpow_axis0, const double *zpow_axis1, double *coeffs,
                        double *mask_min, double *mask_finite, void *stream);
int b200_spline_sample(const double *coeffs, int m, int n, int order, int mode,
                       const double *xy_coords, const double *disp_steps, int T, int row_begin,
                       int row_count, double outval, const double *mask_min,
                       const double *mask_finite, const double *stats, int out_dtype, void *out,
                       void *stream);

/* planar (2,m,n) -> interleaved (m,n,2) copy of an advection field, so that a
 * caller reusing one field for many calls pays the re-layout once
 * (b200_sl_extrapolate does it internally for B200_LAYOUT_PLANAR). */
int b200_sl_interleave_velocity(const void *velocity, int velocity_dtype, int m, int n,
                                void *out, void *stream);

/* pysteps/noise/motion.py:129-180 (initialize_bps :129-141, generate_bps :172-180) fused with
 * the call site pysteps/nowcasts/utils.py:448-451 `velocity + velocity_pert_gen[i](t)`:
 *   p = (a_par * V/|V| + a_perp * perp(V/|V|)) / vsf     at every grid node,
 * a_par = g_par(t)*eps_par, a_perp = g_perp(t)*eps_perp (host scalars).  velocity is planar
 * (2,m,n) of velocity_dtype; `what` selects the float64 output:
 *   B200_BPS_FIELD_INTERLEAVED  V + p as (m,n,2) -- what b200_sl_extrapolate takes with
 *                               B200_LAYOUT_INTERLEAVED; replaces a 64 MB host array and its
 *                               upload per member and time step
 *   B200_BPS_FIELD_PLANAR       V + p as (2,m,n)
 *   B200_BPS_PERTURBATION       p as (2,m,n)       (the value of generate_bps)
 *   B200_BPS_UNIT               V/|V| as (2,m,n)   (perturbator["V_par"]; a, b, vsf unused)
 * n_nonfinite (device, may be NULL) receives the number of non-finite output elements -- the
 * np.isfinite(velocity) check of extrapolation/semilagrangian.py:116-123 on the field the
 * reference would have been given -- without another pass over it. */
#define B200_BPS_FIELD_INTERLEAVED 0
#define B200_BPS_FIELD_PLANAR 1
#define B200_BPS_PERTURBATION 2
#define B200_BPS_UNIT 3
int b200_bps_perturb_velocity(const void *velocity, int velocity_dtype, int m, int n,
                              double a_par, double a_perp, double vsf, int what, double *out,
                              double *n_nonfinite, void *stream);

/* One lead time of EVERY ensemble member of this GPU (nowcasts/utils.py:440-458: per member a
 * single-step extrapolator call with its own BPS-perturbed motion field, precipitation field and
 * carried displacement) in one perturbation launch and one trajectory launch per 8 members.
 *   velocity   (2,m,n) planar base field (device);  pert_coefs  HOST, members x 2: the (a, b) =
 *              (g_par(t) eps_par, g_perp(t) eps_perp) of noise/motion.py:177-180 per member
 *   precip     (members,m,n);  disp_prev (members,2,m,n) or NULL at the first lead time
 *   out        (members,m,n) precip dtype;  disp_out (members,2,m,n)
 *   n_nonfinite (device, members doubles, may be NULL): non-finite elements of each perturbed field
 * Each member's results are those of b200_bps_perturb_velocity + b200_sl_extrapolate with
 * T = 1, n_iter = 1, bit for bit. */
int b200_sl_step_batched(const void *velocity, int velocity_dtype, int m, int n, int members,
                         const double *pert_coefs, double vsf, const void *precip, int precip_dtype,
                         const double *disp_prev, double tdiff, double vel_timestep, double outval,
                         int mode, void *out, double *disp_out, double *n_nonfinite, void *stream);

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
/* dst[0..count) = value (float64) */
int b200_fill_f64(double *dst, int64_t count, double value, void *stream);


/* ------------------------------------------------------------------------
 * Dense Lucas-Kanade motion estimation
 * replaces the array work of pysteps/motion/lucaskanade.py:38-279 and the helpers
 * it calls.  pysteps delegates most of it to opencv-python / scipy; each entry
 * point names the reference call it stands for.  All pointers are device
 * pointers; images are (m,n) row-major.  "n_dev" arguments are optional device
 * pointers to an element count produced by an earlier stage (NULL: use the
 * capacity argument), so the chain runs without host round trips.
 * ---------------------------------------------------------------------- */

/* np.ma.masked_invalid + MaskedArray.min()/max()  (motion/lucaskanade.py:213-219):
 * mask_out = user_mask | !isfinite(img); stats[0..2] = min, max, count of unmasked. */
int b200_mask_invalid(const double *img, const uint8_t *user_mask, int m, int n,
                      uint8_t *mask_out, double *stats, void *stream);

/* pysteps/utils/images.py:27-86 morph_opening (cv2.morphologyEx MORPH_OPEN with the
 * 3x3 cross): pixels > *thr_dev removed by the opening are set to *min_dev (device scalars,
 * e.g. the stats of b200_mask_invalid). size must be 3. */
int b200_morph_opening(const double *img, const uint8_t *mask, int m, int n, int size,
                       const double *thr_dev, const double *min_dev, double *out, void *stream);

/* stats (12 doubles): min / max / count over unmasked pixels of all rows [0..2], of rows
 * >= 1 [3..5], of rows >= 2 [6..8]; [11] = number of pixels whose dilate x dilate buffered
 * mask (cv2.dilate, feature/shitomasi.py:131-139) is clear.  The row sets reproduce the
 * integer-indexing quirk of shitomasi.py:139 (see csrc/lk_dense.cu).  stats0 (optional) is
 * the output of b200_mask_invalid: when it shows no masked pixel the dilation is skipped. */
int b200_masked_minmax(const double *img, const uint8_t *mask, int m, int n, int dilate,
                       const double *stats0, double *stats, void *stream);

/* "scale between 0 and 255" + astype(uint8).  mode 0: tracking/lucaskanade.py:144-160;
 * mode 1: feature/shitomasi.py:131-151 (buffer_mask = dilate) with valid = buffered mask
 * clear.  stats = output of b200_masked_minmax, masked pixels take *fill_dev.
 * mode | B200_QUANTISE_F32: the frames were float32 at the API (img holds them widened): scale in
 * float32 arithmetic as NumPy does for a float32 array. */
#define B200_QUANTISE_F32 2
int b200_quantise_u8(const double *img, const uint8_t *mask, int m, int n, int mode, int dilate,
                     const double *stats, const double *fill_dev, uint8_t *out, uint8_t *valid,
                     void *stream);
/* The four stages above for one frame in three passes (mask + raw min; opening + statistics;
 * opening + both uint8 images), the opened float64 image never written to HBM: the stencil
 * passes stage (64+4) x (16+4) float64 halo tiles in shared memory with TMA
 * (cp.async.bulk.tensor.2d, two-deep mbarrier ring, persistent CTAs; csrc/lk_frontend.cu).
 * Results are those of b200_mask_invalid -> b200_morph_opening -> b200_masked_minmax ->
 * b200_quantise_u8 (mode 0 into q_track; mode 1 into q_det / valid when q_det is not NULL).
 * Needs an even n (16-byte rows for the copy engine) and buffer_mask <= 5. */
int b200_lk_frontend(const double *img, const uint8_t *user_mask, int m, int n, int size_opening,
                     int buffer_mask, int flags, uint8_t *mask, double *stats0, double *stats,
                     uint8_t *q_track, uint8_t *q_det, uint8_t *valid, void *stream);

/* cv::pyrDown (uint8, BORDER_REFLECT_101) and the int16 Scharr pair image of
 * cv::calcOpticalFlowPyrLK's pyramid (dst holds (Ix, Iy) interleaved). */
int b200_pyr_down_u8(const uint8_t *src, int h, int w, uint8_t *dst, void *stream);
int b200_scharr_i16(const uint8_t *src, int h, int w, int16_t *dst, void *stream);

/* cv::cornerMinEigenVal(q, blockSize=5, ksize=3) -> float32 (m,n), bit-identical to
 * opencv-python 4.13.0 (AVX-512 build). */
int b200_min_eig(const uint8_t *q, int m, int n, float *eig, void *stream);

/* cv::goodFeaturesToTrack selection on a minimum-eigenvalue map
 * (feature/shitomasi.py:153-162): out_xy (max_corners,2) float32 (x,y), *out_count.
 * Synchronises the stream once (the sort network is sized by the candidate count). */
int b200_good_features(const float *eig, const uint8_t *valid, int m, int n, int max_corners,
                       double quality_level, double min_distance, float *out_xy,
                       int *out_count, void *stream);

/* Pyramid geometry of cv::buildOpticalFlowPyramid (HOST pointers, 8 entries each). */
int b200_lk_pyramid_layout(int h, int w, int win_w, int win_h, int max_level, int *levels_out,
                           int64_t *offsets, int *hs, int *ws, int64_t *total_pixels);
/* Gaussian pyramid (levels contiguous) and, if deriv != NULL, its Scharr pyramid. */
int b200_lk_build_pyramid(const uint8_t *img, int h, int w, int win_w, int win_h, int max_level,
                          uint8_t *pyr, int16_t *deriv, void *stream);
/* cv::calcOpticalFlowPyrLK (tracking/lucaskanade.py:171), flags = 0: next_pts (npts,2)
 * float32, status (npts) uint8; bit-identical to opencv-python 4.13.0. */
int b200_lk_track(const uint8_t *pyrI, const uint8_t *pyrJ, const int16_t *derivI, int h, int w,
                  int win_w, int win_h, int max_level, int max_count, double epsilon,
                  double min_eig_thr, const float *prev_pts, int npts, const int *npts_dev,
                  float *next_pts, uint8_t *status, void *stream);
/* keep status == 1 rows (tracking/lucaskanade.py:174-181) and append xy = p0,
 * uv = p1 - p0 (float32 arithmetic, widened) to a float64 pool at *pool_count. */
int b200_lk_compact_tracks(const float *p0, const float *p1, const uint8_t *status,
                           const int *npts_dev, int npts_cap, double *pool_xy, double *pool_uv,
                           int *pool_count, int pool_cap, void *stream);

/* pysteps/utils/cleansing.py:124-249 detect_outliers(uv, thr, xy, k), multivariate local
 * branch: out[i] = 1 where the Mahalanobis distance to the k nearest vectors exceeds thr.  The
 * k+1 nearest vectors are taken in scipy.spatial.cKDTree's own order (cleansing.py:219-221), ties
 * and coincident vectors included: the tree is built and queried on the device exactly as scipy
 * does (csrc/knn.cu, knn_body.cuh) -- with integer corner coordinates that order decides tests. */
int b200_detect_outliers(const double *uv, const double *xy, const int *n_dev, int n_cap,
                         double thr, int k, uint8_t *out, void *stream);
/* scipy.spatial.cKDTree(xy) (leafsize 16, median splits by std::nth_element) built on the device:
 * tree_indices[n_cap] = tree.indices (the tree order of the points), *node_count = number of
 * nodes; both device pointers.  Stage entry used by the parity tests of the build. */
int b200_kdtree_build(const double *xy, const int *n_dev, int n_cap, int *tree_indices,
                      int *node_count, void *stream);
/* cleansing.py:201-214, the global branch (k is None): Mahalanobis distance of every vector to the
 * mean of ALL vectors under their sample covariance; out[i] = MD > thr. */
int b200_detect_outliers_global(const double *uv, const int *n_dev, int n_cap, double thr,
                                uint8_t *out, void *stream);
/* rows with drop == 0, order preserved */
int b200_compact_rows(const double *xy, const double *uv, const uint8_t *drop, const int *n_dev,
                      int n_cap, double *out_xy, double *out_uv, int *out_count, void *stream);
/* pysteps/utils/cleansing.py:21-121 decluster(xy, uv, scale, min_samples): per-cell medians,
 * cells in np.unique(axis=0) order. */
int b200_decluster(const double *xy, const double *uv, const int *n_dev, int n_cap, double scale,
                   int min_samples, double *out_xy, double *out_uv, int *out_count, void *stream);

/* pysteps/utils/interpolate.py:26-114 idwinterp2d: k-nearest inverse-distance weighting of
 * (npts, nvar) values onto the (ny, nx) grid -> out (nvar, ny, nx). k <= 32.  Exhaustive
 * tile-culled search; grid points whose k-th and (k+1)-th neighbours are exactly equidistant are
 * recomputed from scipy.spatial.cKDTree's query order (tree built on a library-internal side
 * stream), so the neighbour SET is the reference's everywhere (interpolate.py:78-81).
 * coords_on_16th_grid != 0 is the caller's promise that every vector and grid coordinate is
 * a multiple of 1/16 with magnitude < 2^14 (true for dense_lucaskanade: pixel grids and
 * medians of integer corners); it enables a faster, result-identical key packing.  The value 2
 * promises more: vectors on multiples of 1/2 and INTEGER grid coordinates -- the search then runs
 * on exact 32-bit integer keys. */
int b200_idw_fill(const double *xy, const double *vals, const int *npts_dev, int npts_cap,
                  int nvar, int k, double power, double dist_offset, double mean_res,
                  const double *xgrid, int nx, const double *ygrid, int ny,
                  int coords_on_16th_grid, double *out, void *stream);

/* idwinterp2d with the k nearest vectors of every grid point found, ordered and weighted exactly
 * as the reference does it (scipy.spatial.cKDTree's query order, numpy's pairwise sum of the
 * weights, values accumulated in neighbour order): equal to the reference at EVERY grid point to
 * the last bits (np.power vs pow), ties included; much slower than b200_idw_fill (a tree search
 * per grid point; b200_idw_fill runs it only where the neighbour set depends on it).  k <= 128. */
int b200_idw_fill_ckdtree(const double *xy, const double *vals, const int *npts_dev, int npts_cap,
                          int nvar, int k, double power, double dist_offset, double mean_res,
                          const double *xgrid, int nx, const double *ygrid, int ny, double *out,
                          void *stream);
/* idwinterp2d with k = None (interpolate.py:82-88): every vector contributes to every grid point.
 * nvar <= 8. */
int b200_idw_fill_all(const double *xy, const double *vals, const int *npts_dev, int npts_cap, int nvar,
                      double power, double dist_offset, double mean_res, const double *xgrid, int nx,
                      const double *ygrid, int ny, double *out, void *stream);

/* ------------------------------------------------------------------------
 * Variational Echo Tracking -- replaces the native extension of the reference,
 * pysteps/motion/_vet.pyx.  The CG optimiser (scipy.optimize.minimize, vet.py:593-600)
 * stays on the host; every point it visits costs one b200_vet_value_and_gradient call.
 * Note the reference's axis naming: axis 0 of the images is "x", axis 1 is "y"
 * (_vet.pyx:129-130); sector_disp is (2, xs, ys), images are (nx, ny), mask is int8.
 * ---------------------------------------------------------------------- */

/* _vet.pyx:238-621 _cost_function.  gradient == 0: out[0] = residuals, out[1] =
 * smoothness penalty.  gradient != 0: out (2, xs, ys) = grad_residuals + grad_smooth.
 * smooth_gain is a C float in the reference (:242) and is one here. */
int b200_vet_cost(const double *sector_disp, const double *templ, const double *input,
                  const int8_t *mask, int xs, int ys, int nx, int ny, float smooth_gain,
                  int gradient, double *out, void *stream);
/* vet_cost_function AND vet_cost_function_gradient (vet.py:165-299) at the same point from one
 * pass over the images -- what a line search asks for.  x_host: sector displacements (2, xs, ys)
 * on the HOST; images (nframes, nx, ny) and mask (nx, ny) int8 on the device (2 or 3 frames: the
 * pairs of vet.py:257-268, summed in its order); work: device scratch of 6 * xs * ys + 4 doubles.
 * Returns value_host[0] = residuals, [1] = smoothness penalty (the cost is their sum) and
 * gradient_host (2, xs, ys).  Synchronises the stream (the optimiser needs the numbers). */
int b200_vet_value_and_gradient(const double *x_host, const double *images, int nframes,
                                const int8_t *mask, int xs, int ys, int nx, int ny,
                                float smooth_gain, double *work, double *value_host,
                                double *gradient_host, void *stream);
/* The image stack and mask of one minimisation level from the raw frames (vet.py:507-523 cleaning,
 * :510-517 the global `padding` frame, :548-561 the level's divisibility padding): images
 * (nframes, M, N) zero where a frame is invalid (user_mask set, or -- user_mask NULL -- not finite),
 * edge-replicated into the level padding; mask (M, N) int8 = any frame invalid | global padding
 * ring | level padding.  The level frame starts (pad_i_before, pad_j_before) before the globally
 * padded input. */
int b200_vet_level_images(const double *frames, const uint8_t *user_mask, int nframes, int m,
                          int n, int padding, int pad_i_before, int pad_j_before, int M, int N,
                          double *images, int8_t *mask, void *stream);
/* _vet.pyx:66-232 _warp (vet.morph, vet.py:93-153): out, out_mask (int8) and, if grad is
 * not NULL, the gradient (2, nx, ny). */
int b200_vet_warp(const double *image, const int8_t *mask, const double *displacement, int nx,
                  int ny, double *out, int8_t *out_mask, double *grad, void *stream);
/* scipy.ndimage.zoom(a (c,h,w), (1, oh/h, ow/w), order=1, mode="nearest") (vet.py:621-630) */
int b200_zoom_bilinear(const double *a, int c, int h, int w, int oh, int ow, double *out,
                       void *stream);

/* ------------------------------------------------------------------------
 * Proesmans et al. (1994) optical flow -- replaces the other native extension of the reference,
 * pysteps/motion/_proesmans.pyx (called from pysteps/motion/proesmans.py:88).
 * ---------------------------------------------------------------------- */

/* pysteps/motion/proesmans.py:79-83: out = (frames - im_min) / (im_max - im_min) * 255.0 when
 * do_scale, else a float64 copy. */
int b200_proesmans_scale(const void *frames, int dtype, int64_t count, double im_min, double im_max,
                         int do_scale, double *out, void *stream);
/* scipy.ndimage.gaussian_filter(in, sigma) of a float64 (h,w) image (proesmans.py:85-87): weights
 * (HOST array, 2*radius+1 entries) is scipy's normalised kernel exp(-0.5 x^2 / sigma^2), radius =
 * int(4 sigma + 0.5) <= 64; axis 0 then axis 1, "reflect" boundary, scipy's accumulation order. */
int b200_gaussian_filter(const double *in, int h, int w, const double *weights, int radius, double *out,
                         void *stream);
/* _proesmans.pyx:19-44 _compute_advection_field(R, lam, num_iter, n_levels): frames (2,m,n)
 * float64 -> advfield (2,2,m,n) (forward / backward flow, x / y component) and quality (2,m,n)
 * (the consistency maps).  The relaxation sweep keeps the reference's raster-order Gauss-Seidel
 * update order (wavefronts t = x + 2y inside one CTA per flow field). */
int b200_proesmans_field(const double *frames, int m, int n, double lam, int num_iter, int num_levels,
                         double *advfield, double *quality, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PYSTEPS_B200_H */
