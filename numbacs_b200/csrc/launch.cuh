// launch.cuh -- host-side launch interfaces between capi.cu and the kernel translation units.
#pragma once
#include "common.cuh"
#include "flows.cuh"

namespace b200cs {

// Arguments of the flow-map kernels (passed as one __grid_constant__ struct: every field is
// warp-uniform and lands in the constant bank).
struct IntegArgs {
    RhsParams rhs;
    double x0, xend;  // solver times: params[0]*t0, params[0]*(t0+T)   (integration.py:164)
    double rtol, atol;
    // output times t_k = out_p0 * (out_t0 + k*out_step), k < n_out   (n_out == 0: final state only)
    double out_p0, out_t0, out_step;
    int n_out;
    int out_aligned16;
    // particles: 'ij' grid (x[nx], y[ny]) or a point list pts[npts, N]
    const double *x, *y;
    long long nx, ny;
    const double *pts;
    long long npts;
    const uint8_t *mask;
    // auxiliary stencil (flowmap_aux_grid_2D): n_aux = 4 or 5 particles per grid cell, offset h
    int n_aux, aux_edge;
    double aux_h;
    // time series (kModeSeries): frame f integrates from t0s[f] over series_T; frame_pts = nx*ny
    const double *t0s;
    double series_T;
    long long frame_pts;
    double *out;
    int *status;
    int *steps;
    unsigned long long *stats;
    // fused LAVD (lavd_flowmap kernels only): |vort(traj) - vort_avg| integrated by composite
    // Simpson while the trajectory is being produced, so the [nx,ny,n,2] array never exists
    ScalarDev vort;
    const double *tspan_phys;  // [n_out] physical output times params[0]*t_eval
    const double *vort_avg;    // [n_out] spatial means
    double period_x, period_y;
    double *lavd;              // [npts]
    // queue kernels (flowmap_kernel.cuh): nq pre-initialised particle slots, SoA qstate[6][nq] =
    // (y0, y1, k1_0, k1_1, h, output index as int64: -1 for a slot without work), the fetch counter
    double *qstate;
    unsigned long long *qcounter;
    long long nq;
};

void launch_flowmap(const FlowSpec &f, const IntegArgs &A, int mode /*0 pts, 1 grid, 2 aux grid, 3 time series*/,
                    cudaStream_t s);
void launch_rhs_eval(const FlowSpec &f, const RhsParams &P, const double *t, const double *y,
                     long long npts, double *dy, cudaStream_t s);

// FTLE of rows [row_lo, row_hi) of a flow-map slab fm[nx, ny, 2]; out row r - row_lo.
// `lo_is_border` / `hi_is_border`: slab row 0 / nx-1 is a true domain border (ftle = 0).
void launch_ftle(const double *fm, long long nx, long long ny, double T, double dx, double dy,
                 const uint8_t *mask, double *out, long long row_lo, long long row_hi,
                 bool lo_is_border, bool hi_is_border, cudaStream_t s, long long frames = 1);

SplineGridDev make_grid_dev(const FlowSpec &f);

void launch_scalar_eval(const FlowSpec &f, const double *pts, long long npts, double *out, cudaStream_t s);
void launch_velocity_eval(const FlowSpec &f, const double *pts, long long npts, double *out, cudaStream_t s);
void launch_curl_tspan(const FlowSpec &f, const double *t, long long nt, const double *x, long long nx,
                       const double *y, long long ny, double h, double *curl, cudaStream_t s);
// vorticity contracted over time at the LAVD output times (diag_kernels.cu); W == nullptr: not built
struct VortSlabs {
    Scratch mem;
    const double *W = nullptr;
    long long stride = 0;
};
VortSlabs build_vort_slabs(const FlowSpec &f, const double *tspan_dev, long long n, cudaStream_t s);
void launch_vort_sums(const FlowSpec &f, const double *tspan, long long n, const double *xr,
                      const double *yr, long long nrav, long long ny_grid, double *sums, cudaStream_t s,
                      const VortSlabs *slabs = nullptr);
ScalarDev make_scalar_dev(const FlowSpec &f, const VortSlabs *slabs = nullptr);
void launch_lavd_flowmap(const FlowSpec &f, const IntegArgs &A, cudaStream_t s);
void launch_lavd(const FlowSpec &f, const double *fm_n, long long npts, long long n, const double *tspan,
                 const double *vavg, double period_x, double period_y, const uint8_t *mask, double *lavd,
                 cudaStream_t s, const VortSlabs *slabs = nullptr);
void launch_prefilter3(const double *data, long long n0, long long n1, long long n2, double *coefs,
                       cudaStream_t s);
void launch_div_scalar(double *a, long long n, double d, cudaStream_t s);
void launch_interleave(const double *u, const double *v, long long count, double2 *uv, cudaStream_t s);
void run_fp64_peak(int iters, double *tflops, double *ms);

// ---- tensor_kernels.cu: Cauchy-Green tensor / eigen-pairs, ridge points, order statistics
void launch_c_tensor(const double *fm_aux, long long nx, long long ny, int n_aux, double h,
                     const uint8_t *mask, double *C, cudaStream_t s);
// eigvals from the main-grid stencil (main_vals) or the aux stencil; eigvecs from the aux stencil
// (aux_vecs) or the main grid.  fm is [nx, ny, n_aux, 2] (n_aux = 1: a plain flow map).
void launch_c_eig(const double *fm, long long nx, long long ny, int n_aux, double h, double dx, double dy,
                  bool aux_vecs, bool main_vals, const uint8_t *mask, double *eigvals, double *eigvecs,
                  cudaStream_t s, double *ftle = nullptr, double T = 1.0);
void launch_ftle_from_eig(const double *eigval_max, long long n, long long stride, double T, double *ftle,
                          cudaStream_t s);
void launch_ridge_pts(const double *f, const double *ev, long long ev_pixel_stride, long long ev_comp_stride,
                      long long nx, long long ny, const double *x, const double *y, double dx, double dy,
                      double sdd_thresh, double f_min, double *r_pts, double *r_vec, double *sdd,
                      double *pts_compact, long long capacity, long long *count, cudaStream_t s);
void launch_ridge_components(const double *f, const double *ev, long long ev_pixel_stride,
                             long long ev_comp_stride, long long nx, long long ny, const double *x,
                             const double *y, double dx, double dy, double sdd_thresh, double f_min,
                             double *pts_compact, long long *roots_compact, long long capacity,
                             long long *count, cudaStream_t s);
void launch_composition(const double *flowmaps, const double *grid6 /*host*/, long long nT, double *out,
                        cudaStream_t s, long long frames = 1);
void launch_mask_dilation(const uint8_t *mask, long long nx, long long ny, bool corners, uint8_t *out,
                          cudaStream_t s);
void launch_order_stats(const double *data, long long n, long long k, double *out2, cudaStream_t s);

}  // namespace b200cs
