/*
 * b200cs.h -- C-ABI of libb200cs.so: the B200-native flow-map + FTLE + LAVD hot path of NumbaCS.
 *
 * Every entry point below replaces one call the reference makes on this path (file:line are
 * relative to the reference tree, alb3rtjarvis/numbacs v0.1.2).  The reference has no FFI of its
 * own for this path -- it is Python calling numba-JIT loops that call the third-party
 * `numbalsoda.dop853` and `interpolation.splines` -- so the boundary is drawn exactly where the
 * reference's Python functions sit, and the ctypes binding in numbacs_b200/_lib.py (shown in
 * INTEGRATION.md) is what a maintainer would add to the reference.
 *
 * Conventions
 *   - plain C: pointers + sizes, no C++/torch types.  All arrays float64, C-order, 'ij' indexing:
 *     flowmap[i, j, :] belongs to (x[i], y[j]).
 *   - every data pointer may be a HOST pointer or a DEVICE pointer (cudaPointerGetAttributes is
 *     used to tell).  Host inputs are uploaded, host outputs are downloaded before the call
 *     returns.  With device-only pointers the call is asynchronous on `stream`.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - work runs on the CURRENT CUDA device (cudaSetDevice by the caller; one process per GPU).
 *   - return value: 0 = ok, < 0 = error (B200CS_E_*); b200cs_last_error() gives the message of
 *     the calling thread's last failure.  No exceptions cross the ABI.
 *   - ownership: the caller allocates every output; the library never frees caller memory.  Flow
 *     and scalar-field handles own device copies of their coefficient arrays.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     B200CS_E_CUDA.
 */
#ifndef B200CS_H
#define B200CS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200CS_VERSION 100 /* 0.1.0 */

#if defined(__GNUC__)
#define B200CS_API __attribute__((visibility("default")))
#else
#define B200CS_API
#endif

/* error codes */
#define B200CS_OK 0
#define B200CS_E_INVALID (-1)     /* bad argument */
#define B200CS_E_CUDA (-2)        /* CUDA runtime error / no device */
#define B200CS_E_HANDLE (-3)      /* unknown flow / scalar handle */
#define B200CS_E_UNSUPPORTED (-4) /* e.g. method != dop853 */

/* flow kinds: get_predefined_flow (src/numbacs/flows.py:1104-1295), get_flow_2D (flows.py:121-258) */
#define B200CS_FLOW_DOUBLE_GYRE 0 /* flows.py:1146-1158, 6 params  */
#define B200CS_FLOW_BICKLEY_JET 1 /* flows.py:1182-1213, 12 params */
#define B200CS_FLOW_ABC 2         /* flows.py:1249-1258, 5 params, 3-D state */
#define B200CS_FLOW_SPLINE2D 3    /* flows.py:156-253, 1 param (int_direction) */
#define B200CS_FLOW_LINEAR2D 4    /* flows.py:458-503 (get_flow_linear_2D), 1 param */

/* extrap_mode of interpolation.splines.eval_spline / eval_linear (flows.py:121, 387, 601) */
#define B200CS_EXTRAP_CONSTANT 0
#define B200CS_EXTRAP_LINEAR 1
#define B200CS_EXTRAP_NEAREST 2

/* `method` of flowmap* (src/numbacs/integration.py:8, 124).  Only DOP853 is implemented. */
#define B200CS_METHOD_DOP853 0

/* per-particle status written to the optional `status` arrays */
#define B200CS_ST_MASKED 0     /* mask[i,j] true: not integrated, output row is 0 */
#define B200CS_ST_OK 1         /* numbalsoda `success` == True */
#define B200CS_ST_NMAX (-2)    /* more than 100000 steps */
#define B200CS_ST_HSMALL (-3)  /* step size underflow */

B200CS_API const char *b200cs_last_error(void);
B200CS_API int b200cs_version(void);
/* number of visible CUDA devices; B200CS_E_CUDA if none */
B200CS_API int b200cs_device_count(int *out_count);

/* ---- flow registry: replaces the `funcptr` (address of a numba @cfunc) the reference passes
 *      around.  A GPU cannot call a CPU function pointer, so the int handle names a device-side
 *      implementation instead. ------------------------------------------------------------ */

/* get_predefined_flow(flow_str) -> funcptr               (flows.py:1104, returns 1286-1295) */
B200CS_API int b200cs_flow_create_analytic(int kind, int *out_handle);

/* get_flow_2D(grid_vel, C_eval_u, C_eval_v, spherical, extrap_mode, r) -> funcptr (flows.py:121)
 * grid9 = {t0,t1,nt, x0,x1,nx, y0,y1,ny}; Cu/Cv are the (nt+2, nx+2, ny+2) prefilter outputs.
 * The coefficients are copied (interleaved u,v) to the current device. */
B200CS_API int b200cs_flow_create_spline(const double *grid9, const double *Cu, const double *Cv,
                              int spherical, int extrap_mode, double r, int *out_handle);

/* get_flow_linear_2D(grid_vel, U, V, spherical, extrap_mode, r)   (flows.py:418-506)
 * Same as b200cs_flow_create_spline but over the RAW velocity arrays U, V [nt, nx, ny] with
 * trilinear interpolation (interpolation.splines.eval_linear) instead of the cubic spline. */
B200CS_API int b200cs_flow_create_linear(const double *grid9, const double *U, const double *V,
                              int spherical, int extrap_mode, double r, int *out_handle);

/* get_callable_scalar(grid_f, C_eval_f, extrap_mode)          (flows.py:387-415) linear = 0
 * get_callable_scalar_linear(grid_f, f, extrap_mode)          (flows.py:601-636) linear = 1
 * `data` is (nt+2, nx+2, ny+2) coefficients (cubic) or the raw (nt, nx, ny) field (linear). */
B200CS_API int b200cs_scalar_create(const double *grid9, const double *data, int linear, int extrap_mode,
                         int *out_handle);

B200CS_API int b200cs_flow_destroy(int handle);   /* flows and scalar fields share one handle space */
/* Number of right-hand-side evaluations of an interpolated flow (b200cs_flow_create_spline /
 * _linear) that fell OUTSIDE its data grid since the counter was last reset, summed over every
 * launch that used the handle; 0 for analytic flows.  Synchronises `stream`.  The reference has no
 * such diagnostic: it is the guard on the extrapolation modes of interpolation.splines, whose
 * out-of-grid behaviour no reference test pins -- parity runs assert 0. */
B200CS_API int b200cs_flow_out_of_grid(int flow, int64_t *count, int reset, void *stream);

/* kind (B200CS_FLOW_* or -1 for a scalar field), state dimension, minimum length of params */
B200CS_API int b200cs_flow_info(int handle, int *kind, int *ndim, int *min_params);

/* get_interp_arrays_2D / get_interp_arrays_scalar: interpolation.splines.prefilter(grid, f, k=3)
 * (flows.py:43-44, 116).  data (n0, n1, n2) -> coefs (n0+2, n1+2, n2+2), natural cubic B-spline. */
B200CS_API int b200cs_prefilter_3d(const double *data, int64_t n0, int64_t n1, int64_t n2, double *coefs,
                        void *stream);

/* evaluate a scalar handle at npts points (t, x, y): get_callable_scalar(...)(pts) */
B200CS_API int b200cs_scalar_eval(int handle, const double *pts /*[npts,3]*/, int64_t npts,
                       double *out /*[npts]*/, void *stream);
/* get_callable_2D(grid_vel, C_eval_u, C_eval_v, spherical, extrap_mode, r)(point)   (flows.py:261-384):
 * (u, v) of an interpolated velocity field (a handle from b200cs_flow_create_spline / _linear) at
 * pts [npts,3] = (t, x, y) -> uv [npts,2].  Callable semantics, not right-hand-side semantics: no
 * params[0], no longitude wrap; spherical == 1 applies the 180 / (pi r cos) scaling, any other
 * value returns the raw interpolant, as the reference does. */
B200CS_API int b200cs_velocity_eval(int flow, const double *pts, int64_t npts, double *uv, void *stream);

/* curl_func_tspan(fnc, t, x, y, h)   (utils.py:570-608) with fnc = the callable above:
 * curl [nt,nx,ny] of the velocity by central differences of spacing h at every (t_k, x_i, y_j)
 * (the vorticity field of examples/elliptic_lcs/plot_qge_elliptic_lcs.py:60-62). */
B200CS_API int b200cs_curl_func_tspan(int flow, const double *t, int64_t nt, const double *x, int64_t nx,
                           const double *y, int64_t ny, double h, double *curl, void *stream);

/* evaluate the RHS of a flow at npts states: dy = rhs(t[q], y[q,:], params)  (lsoda_sig cfunc) */
B200CS_API int b200cs_flow_rhs(int flow, const double *t /*[npts]*/, const double *y /*[npts,ndim]*/,
                    int64_t npts, const double *params, int nparams, double *dy /*[npts,ndim]*/,
                    void *stream);

/* ---- particle integration: numbalsoda.dop853 inside a prange over particles --------------- */

/* flowmap_grid_2D(funcptr, t0, T, x, y, params, method, rtol, atol, mask)
 *                                                    (integration.py:123-182)  n == 0
 * flowmap_n_grid_2D(funcptr, t0, T, x, y, params, n, ...)   (integration.py:467-533)  n >= 2
 *   out    : [nx, ny, 2]            if n == 0 (final position only)
 *            [nx, ny, n, 2]         if n >= 2 (n output times incl. the initial condition)
 *   tspan  : nullable [n] = params[0] * t_eval (integration.py:533)
 *   mask   : nullable [nx, ny] bytes (numpy bool_); masked particles give zeros
 *   status : nullable [nx, ny] int32 B200CS_ST_*
 *   steps  : nullable [nx, ny, 2] int32 (accepted, rejected) step counts per particle
 *   stats  : nullable [3] int64 = {sum nfev, sum accepted, sum rejected attempts}, accumulated
 *            (+=) into the caller's buffer                                                     */
B200CS_API int b200cs_flowmap_grid_2d(int flow, double t0, double T, const double *x, int64_t nx,
                           const double *y, int64_t ny, const double *params, int nparams,
                           int method, double rtol, double atol, const uint8_t *mask, int n,
                           double *out, double *tspan, int32_t *status, int32_t *steps,
                           int64_t *stats, void *stream);

/* flowmap(funcptr, t0, T, pts, params, ...)       (integration.py:7-61)    n == 0 -> out[npts, ndim]
 * flowmap_n(funcptr, t0, T, pts, params, n=, ...) (integration.py:64-120)  n >= 2 -> out[npts, n, ndim]
 * pts is [npts, ndim]; ndim must equal the flow's state dimension (2, or 3 for abc). */
B200CS_API int b200cs_flowmap_pts(int flow, double t0, double T, const double *pts, int64_t npts, int ndim,
                       const double *params, int nparams, int method, double rtol, double atol,
                       const uint8_t *mask, int n, double *out, double *tspan, int32_t *status,
                       int32_t *steps, int64_t *stats, void *stream);

/* ---- diagnostics -------------------------------------------------------------------------- */

/* ftle_grid_2D(flowmap, T, dx, dy, mask)   (diagnostics.py:21-65; utils.py:9-46, 168-189) */
B200CS_API int b200cs_ftle_grid_2d(const double *flowmap /*[nx,ny,2]*/, int64_t nx, int64_t ny, double T,
                        double dx, double dy, const uint8_t *mask, double *ftle /*[nx,ny]*/,
                        void *stream);

/* ftle_grid_2D on a row slab of a larger grid (multi-GPU row blocks).  flowmap is [nx, ny, 2]
 * INCLUDING halo rows: with halo_lo / halo_hi = 1 the first / last slab row only serves as the
 * i-1 / i+1 stencil neighbour and no ftle row is produced for it; with 0 that edge is a true
 * domain border (ftle = 0, diagnostics.py:52-53).  ftle is [nx - halo_lo - halo_hi, ny]; mask
 * (nullable) is [nx, ny] like the slab. */
B200CS_API int b200cs_ftle_slab_2d(const double *flowmap, int64_t nx, int64_t ny, double T, double dx,
                        double dy, const uint8_t *mask, int halo_lo, int halo_hi, double *ftle,
                        void *stream);

/* The two calls of the README workflow in one: flowmap_grid_2D followed by ftle_grid_2D, with the
 * flow map kept on the device in between.  `flowmap_out` is nullable (skip the 16 B/particle
 * download when only the FTLE field is wanted).  ftle rows [row_lo, row_hi) of the FULL grid are
 * produced for an x-slab x[0:nx] that already includes whatever halo rows the caller wants:
 * with halo_lo / halo_hi = 1 the first / last slab row is integrated but only used as a stencil
 * neighbour, and its ftle row is not written (multi-GPU row blocks); with 0 that edge is a true
 * domain border and gets ftle = 0 like the reference.  ftle_out is [nx - halo_lo - halo_hi, ny]. */
B200CS_API int b200cs_flowmap_ftle_grid_2d(int flow, double t0, double T, const double *x, int64_t nx,
                                const double *y, int64_t ny, const double *params, int nparams,
                                int method, double rtol, double atol, const uint8_t *mask,
                                double dx, double dy, int halo_lo, int halo_hi,
                                double *flowmap_out /*nullable [nx,ny,2]*/, double *ftle_out,
                                int32_t *status, int64_t *stats, void *stream);

/* lavd_grid_2D(flowmap_n, tspan, T, vort_interp, xrav, yrav, period_x, period_y, mask)
 *                                          (diagnostics.py:272-379; utils.py:611-655)
 * `vort` is a handle from b200cs_scalar_create.  vort_avg (nullable, [n]) returns the spatial
 * means (diagnostics.py:324-331); with vort_avg_in != 0 it is read instead of computed (the
 * multi-GPU path all-reduces it between two calls). */
B200CS_API int b200cs_lavd_grid_2d(const double *flowmap_n /*[nx,ny,n,2]*/, int64_t nx, int64_t ny, int64_t n,
                        const double *tspan /*[n]*/, int vort, const double *xrav,
                        const double *yrav, int64_t nrav, double period_x, double period_y,
                        const uint8_t *mask, double *vort_avg, int vort_avg_in,
                        double *lavd /*[nx,ny]*/, void *stream);
/* flowmap_n_grid_2D + lavd_grid_2D fused ("LAVD carried along the trajectories"): the n-time
 * trajectory array [nx,ny,n,2] (10 GB at 1024^2 x 601) is never materialised -- the integration
 * kernel's dense-output sink evaluates |vort - vort_avg[k]| as each output time is produced and
 * accumulates the composite Simpson sum.  vort_avg (nullable [n]): precomputed spatial means;
 * when NULL they are computed over the (x, y) grid itself (diagnostics.py:324-331 with
 * xrav, yrav = meshgrid(x, y).ravel()).  flowmap_out (nullable [nx,ny,2]) receives the final
 * positions, tspan (nullable [n]) the output times params[0]*t_eval.
 * The three LAVD entries evaluate the vorticity on slabs contracted over time at the n output times
 * (n x one time level of the field in scratch memory, 16 taps per evaluation instead of 64; skipped
 * when the slabs would not fit): the reference's t -> x -> y sum re-associated, a rounding-level
 * change.  Environment B200CS_LAVD_NO_SLABS=1 selects the plain 3-D evaluator (tests, A/B). */
B200CS_API int b200cs_lavd_flowmap_grid_2d(int flow, double t0, double T, const double *x, int64_t nx,
                                const double *y, int64_t ny, const double *params, int nparams,
                                int method, double rtol, double atol, const uint8_t *mask, int n,
                                int vort, double period_x, double period_y, const double *vort_avg,
                                double *lavd, double *flowmap_out, double *tspan, int32_t *status,
                                int64_t *stats, void *stream);

/* partial sums for the spatial mean: sums[k] = sum_q vort(tspan[k], xrav[q], yrav[q]) */
B200CS_API int b200cs_lavd_vort_sums(int vort, const double *tspan, int64_t n, const double *xrav,
                          const double *yrav, int64_t nrav, double *sums /*[n]*/, void *stream);

/* ---- aux grid, Cauchy-Green tensor, eigen-pairs, ridge points ------------------------------- */

/* flowmap_aux_grid_2D(funcptr, t0, T, x, y, params, h, eig_main, compute_edge, method, rtol, atol,
 *                     mask)                                   (integration.py:249-464)
 * Final positions over the auxiliary stencil (x_i +- h, y_j), (x_i, y_j +- h) and, with eig_main,
 * the grid point itself: out is [nx, ny, n_aux, 2] with n_aux = eig_main ? 5 : 4.  Entries the
 * reference does not integrate stay 0: masked cells; with eig_main the four stencil points of
 * edge cells (only their centre point is integrated, and only if compute_edge); without
 * compute_edge every edge cell.  status is [nx, ny, n_aux], steps [nx, ny, n_aux, 2]. */
B200CS_API int b200cs_flowmap_aux_grid_2d(int flow, double t0, double T, const double *x, int64_t nx,
                               const double *y, int64_t ny, const double *params, int nparams,
                               double h, int eig_main, int compute_edge, int method, double rtol,
                               double atol, const uint8_t *mask, double *out, int32_t *status,
                               int32_t *steps, int64_t *stats, void *stream);

/* FTLE time series (examples/time_series/plot_dg_time_series.py:100-112 loops flowmap_grid_2D +
 * ftle_grid_2D over t0; flowmap_composition_initial, integration.py:684-688, loops it over the
 * nT intermediate maps): all nt frames in ONE launch.  Frame f integrates the (x, y) grid over
 * [t0s[f], t0s[f] + T]; out is [nt, nx, ny, 2], status [nt, nx, ny], steps [nt, nx, ny, 2]; the
 * mask [nx, ny] applies to every frame.  Each particle is integrated exactly as by
 * b200cs_flowmap_grid_2d (bit-identical results). */
B200CS_API int b200cs_flowmap_grid_2d_series(int flow, const double *t0s, int64_t nt, double T,
                                  const double *x, int64_t nx, const double *y, int64_t ny,
                                  const double *params, int nparams, int method, double rtol,
                                  double atol, const uint8_t *mask, double *out, int32_t *status,
                                  int32_t *steps, int64_t *stats, void *stream);

/* ftle_grid_2D on every frame of flowmaps [nt, nx, ny, 2] -> ftle [nt, nx, ny], one launch. */
B200CS_API int b200cs_ftle_series_2d(const double *flowmaps, int64_t nt, int64_t nx, int64_t ny, double T,
                          double dx, double dy, const uint8_t *mask, double *ftle, void *stream);

/* C_tensor_2D(flowmap_aux, dx, dy, h, mask)    (diagnostics.py:68-112; utils.py:49-84)
 * C[nx, ny, 3] = (C11, C12, C22) on [2, nx-2) x [2, ny-2), 0 elsewhere; dx, dy unused as in the
 * reference. */
B200CS_API int b200cs_c_tensor_2d(const double *flowmap_aux /*[nx,ny,n_aux,2]*/, int64_t nx, int64_t ny,
                       int n_aux, double dx, double dy, double h, const uint8_t *mask,
                       double *C /*[nx,ny,3]*/, void *stream);

/* C_eig_2D(flowmap, dx, dy, mask)              (diagnostics.py:200-244)
 * eigvals[nx, ny, 2] ascending, eigvecs[nx, ny, 2, 2] (column c belongs to eigvals[..., c]) as
 * np.linalg.eigh returns them (LAPACK dlaev2 conventions, signs included); interior pixels only. */
B200CS_API int b200cs_c_eig_2d(const double *flowmap /*[nx,ny,2]*/, int64_t nx, int64_t ny, double dx,
                    double dy, const uint8_t *mask, double *eigvals, double *eigvecs, void *stream);

/* C_eig_2D followed by ftle_from_eig(eigvals[..., 1], T) in ONE pass (diagnostics.py:200-244 and
 * 247-269; the call pair of examples/ftle/plot_dg_ftle_ridges.py:50-56): the FTLE value is formed
 * from the eigenvalue while it is still in registers instead of re-reading the array that was just
 * written.  ftle [nx,ny] is bit-identical to b200cs_ftle_from_eig on the returned eigvals. */
B200CS_API int b200cs_c_eig_ftle_2d(const double *flowmap, int64_t nx, int64_t ny, double dx, double dy,
                         double T, const uint8_t *mask, double *eigvals, double *eigvecs, double *ftle,
                         void *stream);

/* C_eig_aux_2D(flowmap_aux, dx, dy, h, eig_main, mask)   (diagnostics.py:115-197; utils.py:49-124)
 * eig_main: eigenvalues from the main-grid tensor (centre points), eigenvectors from the aux-grid
 * tensor, on [2, nx-2) x [2, ny-2); otherwise both from the aux grid on [1, nx-1) x [1, ny-1). */
B200CS_API int b200cs_c_eig_aux_2d(const double *flowmap_aux, int64_t nx, int64_t ny, int n_aux, double dx,
                        double dy, double h, int eig_main, const uint8_t *mask, double *eigvals,
                        double *eigvecs, void *stream);

/* ftle_from_eig(eigval_max, T)                 (diagnostics.py:247-269)
 * eigval_max is read with a stride (in doubles) so that eigvals[:, :, 1] can be passed in place:
 * ftle[q] = log(eigval_max[q*stride]) / (2|T|) where > 1, else 0. */
B200CS_API int b200cs_ftle_from_eig(const double *eigval_max, int64_t n, int64_t stride, double T,
                         double *ftle /*[n]*/, void *stream);

/* ftle_ridge_pts(f, eigvec_max, x, y, sdd_thresh, percentile)          (extraction/ridges.py:9-76)
 * _ftle_ridge_pts_connect(f, eigvec_max, x, y, sdd_thresh, percentile)  (ridges.py:232-318)
 * eigvec_max[i, j, c] is read at eigvec_max[(i*ny + j)*ev_pixel_stride + c*ev_comp_stride] so that
 * eigvecs[:, :, :, 1] (strides 4, 2) can be passed in place.  dx, dy are the reference's
 * x[1] - x[0], y[1] - y[0] (ridges.py:39-40), passed explicitly so that a row slab of a larger
 * grid uses the spacing of the full grid.  f_min = 0 or np.percentile(f, p)
 * (from b200cs_order_stats).  All outputs are nullable:
 *   r_pts [nx*ny, 3], r_vec [nx*ny, 2], sdd [nx*ny]  per-pixel arrays of the _connect form
 *       (r_pts rows are -1 where there is no ridge point, the third column is the reference's
 *        ridge-number placeholder -1)
 *   pts_compact [capacity, 2] + count: the ridge points in raveled-pixel order
 *       (r_pts[ridge_bool, :]); *count receives the total number found even when it exceeds
 *       capacity (call with pts_compact = NULL first to size the buffer). */
B200CS_API int b200cs_ftle_ridge_pts(const double *ftle /*[nx,ny]*/, const double *eigvec_max,
                          int64_t ev_pixel_stride, int64_t ev_comp_stride, int64_t nx, int64_t ny,
                          const double *x, const double *y, double dx, double dy,
                          double sdd_thresh, double f_min,
                          double *r_pts, double *r_vec, double *sdd, double *pts_compact,
                          int64_t capacity, int64_t *count, void *stream);

/* ftle_ridges(f, eigvec_max, x, y, sdd_thresh, percentile, min_ridge_pts)   (ridges.py:79-229)
 * The ridge points of b200cs_ftle_ridge_pts together with the 8-connected component of ridge
 * pixels each belongs to: roots_compact[k] is the smallest raveled pixel index of point k's
 * component (so ascending roots = scipy.ndimage.label's numbering, which the reference uses).
 * Points come in raveled-pixel order; grouping them by root gives the reference's list of ridges. */
B200CS_API int b200cs_ftle_ridges(const double *ftle, const double *eigvec_max, int64_t ev_pixel_stride,
                       int64_t ev_comp_stride, int64_t nx, int64_t ny, const double *x, const double *y,
                       double dx, double dy, double sdd_thresh, double f_min,
                       double *pts_compact /*[capacity,2]*/,
                       int64_t *roots_compact /*[capacity]*/, int64_t capacity, int64_t *count,
                       void *stream);

/* _linked_ridge_pts(f, eigvec_max, x, y, sdd_thresh, percentile, c)   (ridges.py:418-603, with
 * _link_points_stepper 321-415).  HOST code (a serial greedy walk): r_pts [nx*ny,3], r_vec
 * [nx*ny,2], sdd [nx*ny] are the host copies of b200cs_ftle_ridge_pts' per-pixel outputs, h =
 * min(dx, dy).  Outputs: linked [n_pts,2] (ordered points, grouped by curve), ridge_len
 * [n_curves,2] = (index one past the curve's last point, its length), endpoints [2 n_curves,3]
 * (x, y, label: +k for the first point of curve k, -(k + 0.1) for the last), ep_tanvecs
 * [2 n_curves,2].  counts[0..1] = (n_pts, n_curves) always; nothing is written when a capacity is
 * too small (call again with room). */
B200CS_API int b200cs_link_ridge_pts(const double *r_pts, const double *r_vec, const double *sdd, int64_t nx,
                          int64_t ny, double h, double c, double sdd_thresh, double *linked,
                          int64_t linked_capacity, int32_t *ridge_len, double *endpoints,
                          double *ep_tanvecs, int64_t curve_capacity, int64_t *counts);

/* ftle_ordered_ridges(f, eigvec_max, x, y, dist_tol, ep_tan_ang, min_ridge_pts, ...)   (ridges.py:720-1054,
 * with _endpoint_distances 606-642 and _connect_endpoints 645-717): joins the curves of
 * b200cs_link_ridge_pts whose end points are within dist_tol and line up within ep_tan_ang.  HOST
 * code.  out_pts [<= n_pts, 2] holds the resulting ridges back to back, offsets [n_out + 1] their
 * boundaries (offsets needs room for n_curves + 1 entries). */
B200CS_API int b200cs_order_ridges(const double *linked, int64_t n_pts, const int32_t *ridge_len,
                        const double *endpoints, const double *ep_tanvecs, int64_t n_curves,
                        double dist_tol, double ep_tan_ang, int64_t min_ridge_pts, double *out_pts,
                        int64_t *offsets, int64_t *n_out);

/* flowmap_composition(flowmaps, grid, nT)      (integration.py:609-644)
 * flowmaps is [nT, nx, ny, 2] (the flow maps over [t0 + k h, t0 + (k+1) h]), grid6 =
 * {x0, x1, nx, y0, y1, ny} (the UCGrid tuple, host or device); composed [nx, ny, 2] =
 * flowmaps[nT-1] o ... o flowmaps[1] o flowmaps[0] by bilinear interpolation with CONSTANT
 * extrapolation (0 outside the grid), all nT-1 interpolation passes fused into one kernel. */
B200CS_API int b200cs_flowmap_composition(const double *flowmaps, const double *grid6, int64_t nT,
                               double *composed, void *stream);

/* binary_mask_dilation(mask, corners)          (utils.py:1923-1985)
 * dilated[i, j] = mask[i, j] or any of its 4 (corners != 0: 8) neighbours; bytes (numpy bool_). */
B200CS_API int b200cs_binary_mask_dilation(const uint8_t *mask, int64_t nx, int64_t ny, int corners,
                                uint8_t *dilated, void *stream);

/* The whole composition time series in one launch: flowmaps is [nT + nframes - 1, nx, ny, 2]
 * (consecutive intervals of length h), frame f composes maps f .. f + nT - 1 exactly like
 * b200cs_flowmap_composition -- what flowmap_composition_step produces frame by frame
 * (integration.py:691-737).  composed is [nframes, nx, ny, 2]. */
B200CS_API int b200cs_flowmap_composition_series(const double *flowmaps, const double *grid6, int64_t nT,
                                      int64_t nframes, double *composed, void *stream);

/* out2 = { sorted(data)[k], sorted(data)[min(k+1, n-1)] } by radix select (no sort, data is not
 * modified): the two order statistics np.percentile interpolates between (ridges.py:45, 279). */
B200CS_API int b200cs_order_stats(const double *data, int64_t n, int64_t k, double *out2, void *stream);

/* ---- measurement helper -------------------------------------------------------------------- */

/* register-resident DFMA chains on every SM for `iters` iterations; returns achieved FP64
 * TFLOP/s (2 flops per DFMA) in *out_tflops.  Used as the measured FP64 roofline denominator. */
B200CS_API int b200cs_fp64_peak(int iters, double *out_tflops, double *out_ms);

#ifdef __cplusplus
}
#endif
#endif /* B200CS_H */
