// capi.cu -- the extern "C" boundary of libb200cs.so (declared in include/b200cs.h).
// Host-side only: argument checking, host<->device staging, the flow registry, kernel launches.
#include "common.cuh"
#include "launch.cuh"

#include <cstring>
#include <vector>

namespace b200cs {
// ridge_link.cu (host code)
void link_ridge_points(const double *r_pts_in, const double *r_vec, const double *sdd, long long nx, long long ny,
                       double h, double c, double sdd_thresh, std::vector<double> &linked,
                       std::vector<int32_t> &ridge_len, std::vector<double> &endpoints,
                       std::vector<double> &tanvecs);
void order_ridges(const std::vector<double> &linked, const std::vector<int32_t> &ridge_len,
                  const std::vector<double> &endpoints, const std::vector<double> &tanvecs, double dist_tol,
                  double ep_tan_ang, long long min_ridge_pts, std::vector<double> &out_pts,
                  std::vector<long long> &offsets);
}  // namespace b200cs

namespace b200cs {

// ---------------------------------------------------------------- errors
static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

bool is_device_ptr(const void *p) {
    cudaPointerAttributes attr{};
    cudaError_t e = cudaPointerGetAttributes(&attr, p);
    if (e != cudaSuccess) {
        cudaGetLastError();  // clear
        return false;
    }
    return attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged;
}

bool is_pinned_host_ptr(const void *p) {
    cudaPointerAttributes attr{};
    cudaError_t e = cudaPointerGetAttributes(&attr, p);
    if (e != cudaSuccess) {
        cudaGetLastError();  // clear
        return false;
    }
    return attr.type == cudaMemoryTypeHost;
}

// Upload of a host array into stream-ordered scratch.  From pageable memory cudaMemcpyAsync returns
// once the source has been staged, so the caller may drop its array as soon as the API call
// returns.  From PINNED memory the copy is truly asynchronous: the call would return while the DMA
// has not read the buffer yet (a pinned torch tensor that the Python wrapper releases right after
// the call).  For those the host waits for THIS copy only (an event behind it), not for the stream.
void upload_to_scratch(void *dst, const void *src, size_t bytes, cudaStream_t s) {
    B2_CHECK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s));
    if (is_pinned_host_ptr(src)) {
        cudaEvent_t ev;
        B2_CHECK_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        cudaError_t e = cudaEventRecord(ev, s);
        if (e == cudaSuccess) e = cudaEventSynchronize(ev);
        cudaEventDestroy(ev);
        B2_CHECK_CUDA(e);
    }
}

// ---------------------------------------------------------------- registry
static std::mutex g_reg_mu;
static std::unordered_map<int, std::shared_ptr<FlowSpec>> g_reg;
static int g_next_handle = 1000;  // never 0, never a plausible pointer

int registry_add(std::shared_ptr<FlowSpec> f) {
    std::lock_guard<std::mutex> lk(g_reg_mu);
    const int h = g_next_handle++;
    g_reg[h] = std::move(f);
    return h;
}

std::shared_ptr<FlowSpec> registry_get(int handle) {
    std::lock_guard<std::mutex> lk(g_reg_mu);
    auto it = g_reg.find(handle);
    if (it == g_reg.end()) {
        set_error("unknown flow handle %d (only handles returned by b200cs_flow_create_* / "
                  "b200cs_scalar_create are valid; a CPU cfunc address cannot run on the GPU)",
                  handle);
        throw Fail{B200CS_E_HANDLE};
    }
    return it->second;
}

bool registry_remove(int handle) {
    std::lock_guard<std::mutex> lk(g_reg_mu);
    return g_reg.erase(handle) > 0;
}

static void require_device() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        set_error("no CUDA device available (%s); libb200cs has no CPU fallback",
                  e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        throw Fail{B200CS_E_CUDA};
    }
    // Scratch buffers come from the stream-ordered allocator; keep freed blocks cached in the pool
    // (the default threshold of 0 hands multi-GB staging buffers back to the driver at every
    // synchronisation and re-maps them on the next call).
    static std::mutex mu;
    static bool tuned[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64) {
        std::lock_guard<std::mutex> lk(mu);
        if (!tuned[dev]) {
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
                unsigned long long keep = ~0ull;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            }
            cudaGetLastError();
            tuned[dev] = true;
        }
    }
}

static void read_grid9(const double *grid9, Axis3 &g) {
    double h[9];
    if (is_device_ptr(grid9)) B2_CHECK_CUDA(cudaMemcpy(h, grid9, sizeof(h), cudaMemcpyDeviceToHost));
    else std::memcpy(h, grid9, sizeof(h));
    for (int d = 0; d < 3; ++d) {
        g.a[d] = h[3 * d];
        g.b[d] = h[3 * d + 1];
        g.n[d] = (int)h[3 * d + 2];
        B2_REQUIRE(g.n[d] >= 2, "grid axis %d needs at least 2 points (got %d)", d, g.n[d]);
        B2_REQUIRE(g.b[d] > g.a[d], "grid axis %d must be ascending", d);
    }
}

// params are tiny: always read them on the host (they become kernel arguments)
static void read_params(const double *params, int nparams, int min_params, RhsParams &R) {
    B2_REQUIRE(params != nullptr && nparams >= min_params,
               "params needs at least %d entries for this flow (got %d)", min_params, nparams);
    B2_REQUIRE(nparams <= kMaxParams, "at most %d params supported (got %d)", kMaxParams, nparams);
    for (int i = 0; i < kMaxParams; ++i) R.p[i] = 0.0;
    if (is_device_ptr(params))
        B2_CHECK_CUDA(cudaMemcpy(R.p, params, sizeof(double) * nparams, cudaMemcpyDeviceToHost));
    else std::memcpy(R.p, params, sizeof(double) * nparams);
}

static void fill_rhs(const FlowSpec &f, RhsParams &R) {
    for (double &d : R.d) d = 0.0;
    for (double &d : R.e) d = 0.0;
    for (double &d : R.e2) d = 0.0;
    if (f.kind == B200CS_FLOW_DOUBLE_GYRE) {
        // constants of the double-gyre RHS folded once here (the same IEEE operations the kernel
        // used to repeat in every stage): see DoubleGyreT in flows.cuh
        R.d[0] = R.p[4] * R.p[0];                                // omega * p0
        R.d[1] = (0.5 * (3.141592653589793 * R.p[1])) * R.p[0];  // p0 * pi * A / 2
        R.d[2] = -(R.p[3] * R.p[0]);                             // -p0 * alpha
        R.d[3] = R.d[0] / 3.141592653589793;                     // phase of a(t) in half-turns:
        R.d[4] = R.p[5] / 3.141592653589793;                     //   u = d[3]*t + d[4]
        // eps folded into the sinpi polynomial of a(t) = eps sin(pi u) (sinpi12_scaled_v, fastmath.cuh)
        for (int k = 0; k < 8; ++k) R.e[k] = R.p[2] * kSinPiCpHost[k];
        R.e[8] = R.p[2] * 3.141592653589793;
        for (int k = 0; k < 8; ++k) R.e2[k] = R.d[1] * kSinPiCpHost[k];   // amplitude of S+ / S- folded the same way
        R.e2[8] = R.d[1] * 3.141592653589793;
    }
    if (f.kind == B200CS_FLOW_BICKLEY_JET) {
#if B200CS_BICKLEY_FOLDED
        // BickleyJet::eval_body, folded form (flows.cuh): every product of launch constants once, here
        const double p0 = R.p[0], U0 = R.p[1], L = R.p[2];
        R.d[7] = 0.5 * L;
        R.d[5] = 1.0 / R.d[7];                              // 2 / L: 2 Y = y1 * d[5], corrected with d[7]
        R.d[6] = 4.0 * (p0 * U0);
        for (int n = 0; n < 3; ++n) {
            R.e[n] = -(p0 * R.p[9 + n]);                    // -p0 c_n
            R.e[3 + n] = 2.0 * R.p[3 + n];                  // 2 A_n
            R.e[6 + n] = -(L * (R.p[3 + n] * R.p[6 + n]));  // -L A_n k_n
        }
#else
        R.d[5] = 1.0 / R.p[2];  // 1 / L_y (BickleyJet::eval)
#endif
    }
    if (f.kind == B200CS_FLOW_SPLINE2D || f.kind == B200CS_FLOW_LINEAR2D) {
        R.d[7] = 3.141592653589793 * f.r;   // pi r, the divisor of the spherical v component (flows.py:180)
        R.d[6] = 1.0 / R.d[7];
    }
    R.coef_uv = nullptr;
    R.slot = -1;
    R.oog = f.oog;
    R.r = f.r;
    std::memset(&R.grid, 0, sizeof(R.grid));
    if (f.kind == B200CS_FLOW_SPLINE2D || f.kind == B200CS_FLOW_LINEAR2D) {
        R.grid = make_grid_dev(f);
        R.grid.oog = f.oog;
        R.coef_uv = static_cast<const double2 *>(f.coef);
    }
}

static void check_device(const FlowSpec &f) {
    int dev = 0;
    B2_CHECK_CUDA(cudaGetDevice(&dev));
    B2_REQUIRE(f.coef == nullptr || f.device == dev,
               "flow handle was created on device %d but the current device is %d", f.device, dev);
}

// common body of flowmap_grid_2d / flowmap_pts
struct AuxSpec {  // flowmap_aux_grid_2D / time-series calls only
    int n_aux = 0;  // 0: not an aux-grid call
    int edge = 0;
    double h = 0.0;
    const double *t0s = nullptr;  // non-null: time series of nt frames (t0 is ignored)
    int64_t nt = 0;
};

static void run_flowmap(int flow, double t0, double T, bool grid_mode, const double *x, int64_t nx,
                        const double *y, int64_t ny, const double *pts, int64_t npts_in, int ndim,
                        const double *params, int nparams, int method, double rtol, double atol,
                        const uint8_t *mask, int n, double *out, double *tspan, int32_t *status,
                        int32_t *steps, int64_t *stats, cudaStream_t s, AuxSpec aux = AuxSpec{}) {
    require_device();
    B2_REQUIRE(method == B200CS_METHOD_DOP853,
               "only method='dop853' is implemented on the GPU (got method id %d)", method);
    B2_REQUIRE(n == 0 || n >= 2, "n must be 0 (final state only) or >= 2 (got %d)", n);
    B2_REQUIRE(rtol > 0.0 && atol >= 0.0, "rtol must be > 0 and atol >= 0");
    auto f = registry_get(flow);
    B2_REQUIRE(f->kind >= 0, "handle %d is a scalar field, not a flow", flow);
    check_device(*f);
    const long long ncell = grid_mode ? (long long)nx * ny : (long long)npts_in;
    const long long npts = aux.n_aux ? ncell * aux.n_aux : (aux.t0s ? ncell * aux.nt : ncell);
    B2_REQUIRE(npts >= 0, "negative particle count");
    if (grid_mode) B2_REQUIRE(f->ndim == 2, "grid entry points need a 2-D flow");
    else B2_REQUIRE(ndim == f->ndim, "pts has %d columns but the flow state is %d-D", ndim, f->ndim);

    IntegArgs A{};
    read_params(params, nparams, f->min_params, A.rhs);
    fill_rhs(*f, A.rhs);
    const double p0 = A.rhs.p[0];
    A.x0 = p0 * t0;            // params[0] * linspace(t0, t0+T, n)[0]   (integration.py:164, 514)
    A.xend = p0 * (t0 + T);    //                               ...[-1]
    A.rtol = rtol;
    A.atol = atol;
    A.n_out = n;
    A.out_p0 = p0;
    A.out_t0 = t0;
    A.out_step = (n > 1) ? ((t0 + T) - t0) / (double)(n - 1) : 0.0;  // numba linspace step
    A.nx = nx;
    A.ny = ny;
    A.npts = npts;
    A.n_aux = aux.n_aux;
    A.aux_edge = aux.edge;
    A.aux_h = aux.h;
    A.series_T = T;
    A.frame_pts = ncell;

    if (tspan && n >= 2) {
        // returned times are params[0] * t_eval (integration.py:120, 533)
        std::vector<double> ts(n);
        for (int k = 0; k < n; ++k) {
            const double te = (k == n - 1) ? p0 * (t0 + T) : p0 * (t0 + (double)k * A.out_step);
            ts[k] = p0 * te;
        }
        if (is_device_ptr(tspan))
            B2_CHECK_CUDA(cudaMemcpyAsync(tspan, ts.data(), sizeof(double) * n, cudaMemcpyHostToDevice, s));
        else std::memcpy(tspan, ts.data(), sizeof(double) * n);
        if (is_device_ptr(tspan)) B2_CHECK_CUDA(cudaStreamSynchronize(s));  // ts goes out of scope
    }
    if (npts == 0) return;

    const int nd = f->ndim;
    const size_t row = (n >= 2) ? (size_t)n * nd : (size_t)nd;
    In<double> dx, dy, dpts;
    if (grid_mode) {
        B2_REQUIRE(x && y, "x and y must not be null");
        dx = In<double>(x, nx, s);
        dy = In<double>(y, ny, s);
    } else {
        B2_REQUIRE(pts, "pts must not be null");
        dpts = In<double>(pts, (size_t)npts * nd, s);
    }
    In<uint8_t> dmask(mask, ncell, s);  // one byte per grid cell / point
    In<double> dt0s(aux.t0s, aux.t0s ? (size_t)aux.nt : 0, s);
    A.t0s = dt0s.dev;
    B2_REQUIRE(out, "out must not be null");
    Out<double> dout(out, (size_t)npts * row, s);
    Out<int32_t> dstatus(status, npts, s);
    Out<int32_t> dsteps(steps, (size_t)npts * 2, s);
    Out<int64_t> dstats(stats, 3, s, /*upload_first=*/true);

    A.x = dx.dev;
    A.y = dy.dev;
    A.pts = dpts.dev;
    A.mask = dmask.dev;
    A.out = dout.dev;
    A.out_aligned16 = ((reinterpret_cast<uintptr_t>(dout.dev) & 15) == 0) ? 1 : 0;
    A.status = dstatus.dev;
    A.steps = dsteps.dev;
    A.stats = reinterpret_cast<unsigned long long *>(dstats.dev);

    launch_flowmap(*f, A, aux.n_aux ? 2 : (aux.t0s ? 3 : (grid_mode ? 1 : 0)), s);

    dout.download();
    dstatus.download();
    dsteps.download();
    dstats.download();
    if (dout.staged() || dstatus.staged() || dsteps.staged() || dstats.staged())
        B2_CHECK_CUDA(cudaStreamSynchronize(s));
}

}  // namespace b200cs

using namespace b200cs;

extern "C" {

const char *b200cs_last_error(void) { return g_err; }

int b200cs_version(void) { return B200CS_VERSION; }

int b200cs_device_count(int *out_count) {
    return guarded([&] {
        B2_REQUIRE(out_count, "out_count is null");
        *out_count = 0;
        require_device();
        B2_CHECK_CUDA(cudaGetDeviceCount(out_count));
    });
}

int b200cs_flow_create_analytic(int kind, int *out_handle) {
    return guarded([&] {
        B2_REQUIRE(out_handle, "out_handle is null");
        auto f = std::make_shared<FlowSpec>();
        f->kind = kind;
        switch (kind) {
        case B200CS_FLOW_DOUBLE_GYRE: f->ndim = 2; f->min_params = 6; break;
        case B200CS_FLOW_BICKLEY_JET: f->ndim = 2; f->min_params = 12; break;
        case B200CS_FLOW_ABC: f->ndim = 3; f->min_params = 5; break;
        default:
            set_error("unknown analytic flow kind %d", kind);
            throw Fail{B200CS_E_INVALID};
        }
        *out_handle = registry_add(std::move(f));
    });
}

static void create_gridded_flow(const double *grid9, const double *Cu, const double *Cv, int spherical,
                                int extrap_mode, double r, bool linear, int *out_handle) {
    require_device();
    B2_REQUIRE(grid9 && Cu && Cv && out_handle, "null argument");
    B2_REQUIRE(spherical >= 0 && spherical <= 2, "spherical must be 0, 1 or 2 (got %d)", spherical);
    B2_REQUIRE(extrap_mode >= 0 && extrap_mode <= 2, "unknown extrap_mode %d", extrap_mode);
    auto f = std::make_shared<FlowSpec>();
    f->kind = linear ? B200CS_FLOW_LINEAR2D : B200CS_FLOW_SPLINE2D;
    f->ndim = 2;
    f->min_params = 1;
    f->spherical = spherical;
    f->extrap = extrap_mode;
    f->linear = linear ? 1 : 0;
    f->r = r;
    read_grid9(grid9, f->grid);
    B2_CHECK_CUDA(cudaGetDevice(&f->device));
    const int pad = linear ? 0 : 2;
    const size_t count = (size_t)(f->grid.n[0] + pad) * (f->grid.n[1] + pad) * (f->grid.n[2] + pad);
    f->coef_bytes = count * sizeof(double2);
    B2_CHECK_CUDA(cudaMalloc(&f->coef, f->coef_bytes));
    B2_CHECK_CUDA(cudaMalloc(&f->oog, sizeof(unsigned long long)));
    B2_CHECK_CUDA(cudaMemset(f->oog, 0, sizeof(unsigned long long)));
    cudaStream_t s = nullptr;
    {
        In<double> du(Cu, count, s), dv(Cv, count, s);
        launch_interleave(du.dev, dv.dev, (long long)count, static_cast<double2 *>(f->coef), s);
        B2_CHECK_CUDA(cudaStreamSynchronize(s));
    }
    *out_handle = registry_add(std::move(f));
}

int b200cs_flow_create_spline(const double *grid9, const double *Cu, const double *Cv, int spherical,
                              int extrap_mode, double r, int *out_handle) {
    return guarded([&] { create_gridded_flow(grid9, Cu, Cv, spherical, extrap_mode, r, false, out_handle); });
}

int b200cs_flow_create_linear(const double *grid9, const double *U, const double *V, int spherical,
                              int extrap_mode, double r, int *out_handle) {
    return guarded([&] { create_gridded_flow(grid9, U, V, spherical, extrap_mode, r, true, out_handle); });
}

int b200cs_scalar_create(const double *grid9, const double *data, int linear, int extrap_mode,
                         int *out_handle) {
    return guarded([&] {
        require_device();
        B2_REQUIRE(grid9 && data && out_handle, "null argument");
        B2_REQUIRE(extrap_mode >= 0 && extrap_mode <= 2, "unknown extrap_mode %d", extrap_mode);
        auto f = std::make_shared<FlowSpec>();
        f->kind = -1;
        f->ndim = 0;
        f->min_params = 0;
        f->linear = linear ? 1 : 0;
        f->extrap = extrap_mode;
        read_grid9(grid9, f->grid);
        B2_CHECK_CUDA(cudaGetDevice(&f->device));
        const int pad = linear ? 0 : 2;
        const size_t count = (size_t)(f->grid.n[0] + pad) * (f->grid.n[1] + pad) * (f->grid.n[2] + pad);
        f->coef_bytes = count * sizeof(double);
        B2_CHECK_CUDA(cudaMalloc(&f->coef, f->coef_bytes));
        B2_CHECK_CUDA(cudaMemcpy(f->coef, data, f->coef_bytes, cudaMemcpyDefault));
        *out_handle = registry_add(std::move(f));
    });
}

int b200cs_flow_destroy(int handle) {
    return guarded([&] {
        if (!registry_remove(handle)) {
            set_error("unknown flow handle %d", handle);
            throw Fail{B200CS_E_HANDLE};
        }
    });
}

int b200cs_flow_out_of_grid(int flow, int64_t *count, int reset, void *stream) {
    return guarded([&] {
        require_device();
        auto f = registry_get(flow);
        B2_REQUIRE(count, "count is null");
        *count = 0;
        if (!f->oog) return;    // analytic flows have no data grid
        check_device(*f);
        cudaStream_t s = static_cast<cudaStream_t>(stream);
        unsigned long long v = 0;
        B2_CHECK_CUDA(cudaMemcpyAsync(&v, f->oog, sizeof(v), cudaMemcpyDeviceToHost, s));
        if (reset) B2_CHECK_CUDA(cudaMemsetAsync(f->oog, 0, sizeof(v), s));
        B2_CHECK_CUDA(cudaStreamSynchronize(s));
        *count = (int64_t)v;
    });
}

int b200cs_flow_info(int handle, int *kind, int *ndim, int *min_params) {
    return guarded([&] {
        auto f = registry_get(handle);
        if (kind) *kind = f->kind;
        if (ndim) *ndim = f->ndim;
        if (min_params) *min_params = f->min_params;
    });
}

int b200cs_prefilter_3d(const double *data, int64_t n0, int64_t n1, int64_t n2, double *coefs,
                        void *stream) {
    return guarded([&] {
        require_device();
        B2_REQUIRE(data && coefs, "null argument");
        B2_REQUIRE(n0 >= 2 && n1 >= 2 && n2 >= 2, "every axis needs at least 2 points");
        cudaStream_t s = static_cast<cudaStream_t>(stream);
        const size_t nin = (size_t)n0 * n1 * n2, nout = (size_t)(n0 + 2) * (n1 + 2) * (n2 + 2);
        In<double> din(data, nin, s);
        Out<double> dout(coefs, nout, s);
        launch_prefilter3(din.dev, n0, n1, n2, dout.dev, s);
        dout.download();
        if (dout.staged()) B2_CHECK_CUDA(cudaStreamSynchronize(s));
    });
}

int b200cs_scalar_eval(int handle, const double *pts, int64_t npts, double *out, void *stream) {
    return guarded([&] {
        require_device();
        auto f = registry_get(handle);
        B2_REQUIRE(f->kind == -1, "handle %d is a flow, not a scalar field", handle);
        check_device(*f);
        B2_REQUIRE(pts && out, "null argument");
        cudaStream_t s = static_cast<cudaStream_t>(stream);
        In<double> dp(pts, (size_t)npts * 3, s);
        Out<double> dout(out, npts, s);
        launch_scalar_eval(*f, dp.dev, npts, dout.dev, s);
        dout.download();
        if (dout.staged()) B2_CHECK_CUDA(cudaStreamSynchronize(s));
    });
}

int b200cs_velocity_eval(int flow, const double *pts, int64_t npts, double *uv, void *stream) {
    return guarded([&] {
        require_device();
        auto f = registry_get(flow);
        B2_REQUIRE(f->kind == B200CS_FLOW_SPLINE2D || f->kind == B200CS_FLOW_LINEAR2D,
                   "handle %d is not an interpolated velocity field", flow);
        check_device(*f);
        B2_REQUIRE(npts >= 0 && (npts == 0 || (pts && uv)), "null argument");
        cudaStream_t s = static_cast<cudaStream_t>(stream);
        In<double> dp(pts, (size_t)npts * 3, s);
        Out<double> dout(uv, (size_t)npts * 2, s);
        launch_velocity_eval(*f, dp.dev, npts, dout.dev, s);
        dout.download();
        if (dout.staged()) B2_CHECK_CUDA(cudaStreamSynchronize(s));
    });
}

int b200cs_curl_func_tspan(int flow, const double *t, int64_t nt, const double *x, int64_t nx, const double *y,
                           int64_t ny, double h, double *curl, void *stream) {
    return guarded([&] {
        require_device();
        auto f = registry_get(flow);
        B2_REQUIRE(f->kind == B200CS_FLOW_SPLINE2D || f->kind == B200CS_FLOW_LINEAR2D,
                   "handle %d is not an interpolated velocity field", flow);
        check_device(*f);
        B2_REQUIRE(nt >= 0 && nx >= 0 && ny >= 0, "negative size");
        B2_REQUIRE(h != 0.0, "h must be non-zero");
        if (nt == 0 || nx == 0 || ny == 0) return;
        B2_REQUIRE(t && x && y && curl, "null argument");
        cudaStream_t s = static_cast<cudaStream_t>(stream);
        In<double> dt(t, nt, s), dx(x, nx, s), dy(y, ny, s);
        Out<double> dout(curl, (size_t)nt * nx * ny, s);
        launch_curl_tspan(*f, dt.dev, nt, dx.dev, nx, dy.dev, ny, h, dout.dev, s);
        dout.download();
        if (dout.staged()) B2_CHECK_CUDA(cudaStreamSynchronize(s));
    });
}

int b200cs_flow_rhs(int flow, const double *t, const double *y, int64_t npts, const double *params,
                    int nparams, double *dy, void *stream) {
    return guarded([&] {
        require_device();
        auto f = registry_get(flow);
        B2_REQUIRE(f->kind >= 0, "handle %d is a scalar field, not a flow", flow);
        check_device(*f);
        B2_REQUIRE(t && y && dy, "null argument");
        cudaStream_t s = static_cast<cudaStream_t>(stream);
        RhsParams R{};
        read_params(params, nparams, f->min_params, R);
        fill_rhs(*f, R);
        In<double> dt(t, npts, s), dyin(y, (size_t)npts * f->ndim, s);
        Out<double> dout(dy, (size_t)npts * f->ndim, s);
        launch_rhs_eval(*f, R, dt.dev, dyin.dev, npts, dout.dev, s);
        dout.download();
        if (dout.staged()) B2_CHECK_CUDA(cudaStreamSynchronize(s));
    });
}

int b200cs_flowmap_grid_2d(int flow, double t0, double T, const double *x, int64_t nx, const double *y,
                           int64_t ny, const double *params, int nparams, int method, double rtol,
                           double atol, const uint8_t *mask, int n, double *out, double *tspan,
                           int32_t *status, int32_t *steps, int64_t *stats, void *stream) {
    return guarded([&] {
        B2_REQUIRE(nx >= 0 && ny >= 0, "negative grid size");
        run_flowmap(flow, t0, T, true, x, nx, y, ny, nullptr, 0, 2, params, nparams, method, rtol, atol,
                    mask, n, out, tspan, status, steps, stats, static_cast<cudaStream_t>(stream));
    });
}

int b200cs_flowmap_pts(int flow, double t0, double T, const double *pts, int64_t npts, int ndim,
                       const double *params, int nparams, int method, double rtol, double atol,
                       const uint8_t *mask, int n, double *out, double *tspan, int32_t *status,
                       int32_t *steps, int64_t *stats, void *stream) {
    return guarded([&] {
        B2_REQUIRE(npts >= 0, "negative point count");
        run_flowmap(flow, t0, T, false, nullptr, 0, nullptr, 0, pts, npts, ndim, params, nparams, method,
                    rtol, atol, mask, n, out, tspan, status, steps, stats,
                    static_cast<cudaStream_t>(stream));
    });
}

int b200cs_ftle_grid_2d(const double *flowmap, int64_t nx, int64_t ny, double T, double dx, double dy,
                        const uint8_t *mask, double *ftle, void *stream) {
    return guarded([&] {
        require_device();
        B2_REQUIRE(flowmap && ftle, "null argument");
        B2_REQUIRE(nx >= 0 && ny >= 0, "negative grid size");
        B2_REQUIRE(T != 0.0 && dx != 0.0 && dy != 0.0, "T, dx and dy must be non-zero");
        if (nx == 0 || ny == 0) return;
        cudaStream_t s = static_cast<cudaStream_t>(stream);
        const size_t np = (size_t)nx * ny;
        In<double> dfm(flowmap, np * 2, s);
        In<uint8_t> dmask(mask, np, s);
        Out<double> dout(ftle, np, s);
        launch_ftle(dfm.dev, nx, ny, T, dx, dy, dmask.dev, dout.dev, 0, nx, true, true, s);
        dout.download();
        if (dout.staged()) B2_CHECK_CUDA(cudaStreamSynchronize(s));
    });
}

int b200cs_ftle_slab_2d(const double *flowmap, int64_t nx, int64_t ny, double T, double dx, double dy,
                        const uint8_t *mask, int halo_lo, int halo_hi, double *ftle, void *stream) {
    return guarded([&] {
        require_device();
        B2_REQUIRE(flowmap && ftle, "null argument");
        B2_REQUIRE(nx >= 0 && ny >= 0, "negative grid size");
        B2_REQUIRE((halo_lo == 0 || halo_lo == 1) && (halo_hi == 0 || halo_hi == 1), "halo must be 0 or 1");
        B2_REQUIRE(T != 0.0 && dx != 0.0 && dy != 0.0, "T, dx and dy must be non-zero");
        const long long rows_out = (long long)nx - halo_lo - halo_hi;
        B2_REQUIRE(rows_out >= 0, "slab smaller than its halo");
        if (rows_out == 0 || ny == 0) return;
        cudaStream_t s = static_cast<cudaStream_t>(stream);
        const size_t np = (size_t)nx * ny;
        In<double> dfm(flowmap, np * 2, s);
        In<uint8_t> dmask(mask, np, s);
        Out<double> dout(ftle, (size_t)rows_out * ny, s);
        launch_ftle(dfm.dev, nx, ny, T, dx, dy, dmask.dev, dout.dev, halo_lo, (long long)nx - halo_hi,
                    halo_lo == 0, halo_hi == 0, s);
        dout.download();
        if (dout.staged()) B2_CHECK_CUDA(cudaStreamSynchronize(s));
    });
}

int b200cs_flowmap_ftle_grid_2d(int flow, double t0, double T, const double *x, int64_t nx,
                                const double *y, int64_t ny, const double *params, int nparams,
                                int method, double rtol, double atol, const uint8_t *mask, double dx,
                                double dy, int halo_lo, int halo_hi, double *flowmap_out,
                                double *ftle_out, int32_t *status, int64_t *stats, void *stream) {
    return guarded([&] {
        require_device();
        B2_REQUIRE(ftle_out, "ftle_out is null");
        B2_REQUIRE(nx >= 0 && ny >= 0, "negative grid size");
        B2_REQUIRE((halo_lo == 0 || halo_lo == 1) && (halo_hi == 0 || halo_hi == 1), "halo must be 0 or 1");
        B2_REQUIRE(T != 0.0 && dx != 0.0 && dy != 0.0, "T, dx and dy must be non-zero");
        cudaStream_t s = static_cast<cudaStream_t>(stream);
        const long long rows_out = (long long)nx - halo_lo - halo_hi;
        B2_REQUIRE(rows_out >= 0, "slab smaller than its halo");
        if (nx == 0 || ny == 0) return;
        const size_t np = (size_t)nx * ny;
        // the flow map lives on the device between the two kernels; a host `flowmap_out` only
        // receives a copy, a device `flowmap_out` is used directly
        Scratch fm_tmp;
        double *fm_dev = nullptr;
        bool fm_host_copy = false;
        if (flowmap_out && is_device_ptr(flowmap_out)) {
            fm_dev = flowmap_out;
        } else {
            fm_tmp = Scratch(np * 2 * sizeof(double), s);
            fm_dev = static_cast<double *>(fm_tmp.ptr);
            fm_host_copy = flowmap_out != nullptr;
        }
        In<uint8_t> dmask(mask, np, s);  // staged once, used by both kernels
        const bool host_ftle = !is_device_ptr(ftle_out);
        // ~16 chunks per call, 128..1024 rows each (a multiple of 8)
        long long chunk_rows = ((nx + 15) / 16 + 7) / 8 * 8;
        chunk_rows = chunk_rows < 128 ? 128 : (chunk_rows > 1024 ? 1024 : chunk_rows);
        if ((host_ftle || fm_host_copy) && nx >= 2 * chunk_rows && np >= (size_t(1) << 22)) {
            // Host outputs on a large grid: pipeline.  Rows are integrated in chunks; as soon as a
            // chunk's successor has been integrated its FTLE rows are complete and go out over
            // PCIe while the next chunks are being integrated (the integration is compute-bound,
            // the download is free).  Consecutive chunks alternate between two streams so that the
            // tail of one chunk's kernel (SMs idling while the last blocks finish) is filled by the
            // head of the next; the FTLE kernels and the copies run on high-priority streams so
            // that they are not queued behind whole integration chunks.
            In<double> dxs(x, nx, s), dys(y, ny, s);
            Out<int32_t> dstatus(status, np, s);
            Out<int64_t> dstats(stats, 3, s, /*upload_first=*/true);
            Scratch ftle_tmp;
            double *ftle_dev = ftle_out;
            if (host_ftle) {
                ftle_tmp = Scratch((size_t)rows_out * ny * sizeof(double), s);
                ftle_dev = static_cast<double *>(ftle_tmp.ptr);
            }
            int prio_lo = 0, prio_hi = 0;
            B2_CHECK_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
            cudaStream_t sI[2] = {nullptr, nullptr}, sF = nullptr, s2 = nullptr;
            std::vector<cudaEvent_t> events;
            auto new_event = [&](cudaStream_t on) {
                cudaEvent_t ev;
                B2_CHECK_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
                events.push_back(ev);
                B2_CHECK_CUDA(cudaEventRecord(ev, on));
                return ev;
            };
            auto cleanup = [&] {
                for (cudaStream_t st : {sI[0], sI[1], sF, s2})
                    if (st) cudaStreamSynchronize(st);
                for (cudaEvent_t ev : events) cudaEventDestroy(ev);
                for (cudaStream_t st : {sI[0], sI[1], sF, s2})
                    if (st) cudaStreamDestroy(st);
            };
            long long ftle_done = halo_lo, fm_done = 0;
            try {
                B2_CHECK_CUDA(cudaStreamCreateWithPriority(&sI[0], cudaStreamNonBlocking, prio_lo));
                B2_CHECK_CUDA(cudaStreamCreateWithPriority(&sI[1], cudaStreamNonBlocking, prio_lo));
                B2_CHECK_CUDA(cudaStreamCreateWithPriority(&sF, cudaStreamNonBlocking, prio_hi));
                B2_CHECK_CUDA(cudaStreamCreateWithPriority(&s2, cudaStreamNonBlocking, prio_hi));
                const cudaEvent_t ev_start = new_event(s);  // uploads / allocations on the caller's stream
                for (cudaStream_t st : {sI[0], sI[1], sF, s2}) B2_CHECK_CUDA(cudaStreamWaitEvent(st, ev_start, 0));
                int k = 0;
                for (long long c0 = 0; c0 < nx; c0 += chunk_rows, ++k) {
                    const long long c1 = (c0 + chunk_rows < nx) ? c0 + chunk_rows : nx;
                    cudaStream_t si = sI[k & 1];
                    run_flowmap(flow, t0, T, true, dxs.dev + c0, c1 - c0, dys.dev, ny, nullptr, 0, 2, params,
                                nparams, method, rtol, atol, dmask.dev ? dmask.dev + c0 * ny : nullptr, 0,
                                fm_dev + c0 * ny * 2, nullptr, dstatus.dev ? dstatus.dev + c0 * ny : nullptr,
                                nullptr, dstats.dev, si);
                    // sF sees the chunks in order: it waited for chunk k-1 in the previous iteration
                    B2_CHECK_CUDA(cudaStreamWaitEvent(sF, new_event(si), 0));
                    const long long hi = (c1 == nx) ? (long long)nx - halo_hi : c1 - 1;  // stencil complete below hi
                    if (hi > ftle_done)
                        launch_ftle(fm_dev, nx, ny, T, dx, dy, dmask.dev, ftle_dev + (ftle_done - halo_lo) * ny,
                                    ftle_done, hi, halo_lo == 0, halo_hi == 0, sF);
                    B2_CHECK_CUDA(cudaStreamWaitEvent(s2, new_event(sF), 0));
                    if (host_ftle && hi > ftle_done)
                        B2_CHECK_CUDA(cudaMemcpyAsync(ftle_out + (ftle_done - halo_lo) * ny,
                                                      ftle_dev + (ftle_done - halo_lo) * ny,
                                                      (size_t)(hi - ftle_done) * ny * sizeof(double),
                                                      cudaMemcpyDeviceToHost, s2));
                    if (fm_host_copy)
                        B2_CHECK_CUDA(cudaMemcpyAsync(flowmap_out + fm_done * ny * 2, fm_dev + fm_done * ny * 2,
                                                      (size_t)(c1 - fm_done) * ny * 2 * sizeof(double),
                                                      cudaMemcpyDeviceToHost, s2));
                    if (hi > ftle_done) ftle_done = hi;
                    fm_done = c1;
                }
                // the caller's stream continues after everything issued above
                B2_CHECK_CUDA(cudaStreamWaitEvent(s, new_event(s2), 0));
                dstatus.download();
                dstats.download();
                B2_CHECK_CUDA(cudaStreamSynchronize(s));
            } catch (...) {
                cleanup();
                throw;
            }
            cleanup();
            return;
        }
        run_flowmap(flow, t0, T, true, x, nx, y, ny, nullptr, 0, 2, params, nparams, method, rtol, atol,
                    dmask.dev, 0, fm_dev, nullptr, status, nullptr, stats, s);
        Out<double> dftle(ftle_out, (size_t)rows_out * ny, s);
        launch_ftle(fm_dev, nx, ny, T, dx, dy, dmask.dev, dftle.dev, halo_lo, (long long)nx - halo_hi,
                    halo_lo == 0, halo_hi == 0, s);
        dftle.download();
        if (fm_host_copy)
            B2_CHECK_CUDA(cudaMemcpyAsync(flowmap_out, fm_dev, np * 2 * sizeof(double), cudaMemcpyDeviceToHost, s));
        if (dftle.staged() || fm_host_copy) B2_CHECK_CUDA(cudaStreamSynchronize(s));
    });
}

int b200cs_lavd_vort_sums(int vort, const double *tspan, int64_t n, const double *xrav, const double *yrav,
                          int64_t nrav, double *sums, void *stream) {
    return guarded([&] {
        require_device();
        auto f = registry_get(vort);
        B2_REQUIRE(f->kind == -1, "handle %d is a flow, not a scalar field", vort);
        check_device(*f);
        B2_REQUIRE(tspan && xrav && yrav && sums, "null argument");
        cudaStream_t s = static_cast<cudaStream_t>(stream);
        In<double> dts(tspan, n, s), dxr(xrav, nrav, s), dyr(yrav, nrav, s);
        Out<double> dsum(sums, n, s);
        const VortSlabs slabs = build_vort_slabs(*f, dts.dev, n, s);
        launch_vort_sums(*f, dts.dev, n, dxr.dev, dyr.dev, nrav, 0, dsum.dev, s, &slabs);
        dsum.download();
        if (dsum.staged()) B2_CHECK_CUDA(cudaStreamSynchronize(s));
    });
}

int b200cs_lavd_grid_2d(const double *flowmap_n, int64_t nx, int64_t ny, int64_t n, const double *tspan,
                        int vort, const double *xrav, const double *yrav, int64_t nrav, double period_x,
                        double period_y, const uint8_t *mask, double *vort_avg, int vort_avg_in,
                        double *lavd, void *stream) {
    return guarded([&] {
        require_device();
        auto f = registry_get(vort);
        B2_REQUIRE(f->kind == -1, "handle %d is a flow, not a scalar field", vort);
        check_device(*f);
        B2_REQUIRE(flowmap_n && tspan && lavd, "null argument");
        B2_REQUIRE(n >= 3, "lavd needs at least 3 output times (got %lld)", (long long)n);
        B2_REQUIRE(vort_avg_in ? vort_avg != nullptr : (xrav && yrav && nrav > 0),
                   "either pass vort_avg (vort_avg_in=1) or the xrav/yrav grid");
        cudaStream_t s = static_cast<cudaStream_t>(stream);
        const size_t np = (size_t)nx * ny;
        if (np == 0) return;
        In<double> dfm(flowmap_n, np * n * 2, s), dts(tspan, n, s);
        In<uint8_t> dmask(mask, np, s);
        Out<double> dlavd(lavd, np, s);
        Out<double> davg(vort_avg, n, s, /*upload_first=*/vort_avg_in != 0);
        Scratch avg_tmp;
        double *avg_dev = davg.dev;
        if (!avg_dev) {
            avg_tmp = Scratch(sizeof(double) * n, s);
            avg_dev = static_cast<double *>(avg_tmp.ptr);
        }
        // the vorticity contracted over time at the n output times, shared by the mean and the integrand
        const VortSlabs slabs = build_vort_slabs(*f, dts.dev, n, s);
        if (!vort_avg_in) {
            In<double> dxr(xrav, nrav, s), dyr(yrav, nrav, s);
            launch_vort_sums(*f, dts.dev, n, dxr.dev, dyr.dev, nrav, 0, avg_dev, s, &slabs);
            launch_div_scalar(avg_dev, n, (double)nrav, s);  // np.mean
        }
        launch_lavd(*f, dfm.dev, (long long)np, n, dts.dev, avg_dev, period_x, period_y, dmask.dev,
                    dlavd.dev, s, &slabs);
        dlavd.download();
        if (!vort_avg_in) davg.download();
        if (dlavd.staged() || davg.staged()) B2_CHECK_CUDA(cudaStreamSynchronize(s));
    });
}

int b200cs_lavd_flowmap_grid_2d(int flow, double t0, double T, const double *x, int64_t nx, const double *y,
                                int64_t ny, const double *params, int nparams, int method, double rtol,
                                double atol, const uint8_t *mask, int n, int vort, double period_x,
                                double period_y, const double *vort_avg, double *lavd, double *flowmap_out,
                                double *tspan, int32_t *status, int64_t *stats, void *stream) {
    return guarded([&] {
        require_device();
        B2_REQUIRE(method == B200CS_METHOD_DOP853, "only method='dop853' is implemented on the GPU");
        B2_REQUIRE(n >= 3, "lavd needs at least 3 output times (got %d)", n);
        B2_REQUIRE(rtol > 0.0 && atol >= 0.0, "rtol must be > 0 and atol >= 0");
        B2_REQUIRE(x && y && lavd && nx >= 0 && ny >= 0, "bad argument");
        auto f = registry_get(flow);
        B2_REQUIRE(f->kind >= 0 && f->ndim == 2, "handle %d is not a 2-D flow", flow);
        auto fv = registry_get(vort);
        B2_REQUIRE(fv->kind == -1, "handle %d is a flow, not a scalar field", vort);
        check_device(*f);
        check_device(*fv);
        cudaStream_t s = static_cast<cudaStream_t>(stream);
        const long long npts = (long long)nx * ny;

        IntegArgs A{};
        read_params(params, nparams, f->min_params, A.rhs);
        fill_rhs(*f, A.rhs);
        const double p0 = A.rhs.p[0];
        A.x0 = p0 * t0;
        A.xend = p0 * (t0 + T);
        A.rtol = rtol;
        A.atol = atol;
        A.n_out = n;
        A.out_p0 = p0;
        A.out_t0 = t0;
        A.out_step = ((t0 + T) - t0) / (double)(n - 1);
        A.nx = nx;
        A.ny = ny;
        A.npts = npts;
        std::vector<double> ts(n);  // physical output times params[0]*t_eval (integration.py:533)
        for (int k = 0; k < n; ++k)
            ts[k] = p0 * ((k == n - 1) ? p0 * (t0 + T) : p0 * (t0 + (double)k * A.out_step));
        if (tspan) {
            if (is_device_ptr(tspan)) B2_CHECK_CUDA(cudaMemcpy(tspan, ts.data(), sizeof(double) * n, cudaMemcpyHostToDevice));
            else std::memcpy(tspan, ts.data(), sizeof(double) * n);
        }
        if (npts == 0) return;
        Scratch ts_dev(sizeof(double) * n, s), avg_dev(sizeof(double) * n, s);
        B2_CHECK_CUDA(cudaMemcpyAsync(ts_dev.ptr, ts.data(), sizeof(double) * n, cudaMemcpyHostToDevice, s));
        B2_CHECK_CUDA(cudaStreamSynchronize(s));  // ts is a local
        In<double> dx(x, nx, s), dy(y, ny, s);
        In<uint8_t> dmask(mask, npts, s);
        // the vorticity contracted over time at the n output times: 16 taps per evaluation instead of 64,
        // for the spatial means and for the integrand along the trajectories
        const VortSlabs slabs = build_vort_slabs(*fv, static_cast<const double *>(ts_dev.ptr), n, s);
        if (vort_avg) {
            B2_CHECK_CUDA(cudaMemcpyAsync(avg_dev.ptr, vort_avg, sizeof(double) * n, cudaMemcpyDefault, s));
        } else {
            launch_vort_sums(*fv, static_cast<double *>(ts_dev.ptr), n, dx.dev, dy.dev, npts, ny,
                             static_cast<double *>(avg_dev.ptr), s, &slabs);
            launch_div_scalar(static_cast<double *>(avg_dev.ptr), n, (double)npts, s);  // np.mean
        }
        Out<double> dlavd(lavd, npts, s);
        Out<double> dfm(flowmap_out, (size_t)npts * 2, s);
        Out<int32_t> dstatus(status, npts, s);
        Out<int64_t> dstats(stats, 3, s, /*upload_first=*/true);
        A.x = dx.dev;
        A.y = dy.dev;
        A.mask = dmask.dev;
        A.out = dfm.dev;
        A.status = dstatus.dev;
        A.stats = reinterpret_cast<unsigned long long *>(dstats.dev);
        A.vort = make_scalar_dev(*fv, &slabs);
        A.tspan_phys = static_cast<const double *>(ts_dev.ptr);
        A.vort_avg = static_cast<const double *>(avg_dev.ptr);
        A.period_x = period_x;
        A.period_y = period_y;
        A.lavd = dlavd.dev;
        launch_lavd_flowmap(*f, A, s);
        dlavd.download();
        dfm.download();
        dstatus.download();
        dstats.download();
        if (dlavd.staged() || dfm.staged() || dstatus.staged() || dstats.staged())
            B2_CHECK_CUDA(cudaStreamSynchronize(s));
    });
}

int b200cs_flowmap_aux_grid_2d(int flow, double t0, double T, const double *x, int64_t nx, const double *y,
                               int64_t ny, const double *params, int nparams, double h, int eig_main,
                               int compute_edge, int method, double rtol, double atol,
                               const uint8_t *mask, double *out, int32_t *status, int32_t *steps,
                               int64_t *stats, void *stream) {
    return guarded([&] {
        B2_REQUIRE(nx >= 0 && ny >= 0, "negative grid size");
        AuxSpec aux;
        aux.n_aux = eig_main ? 5 : 4;
        aux.edge = compute_edge ? 1 : 0;
        aux.h = h;
        run_flowmap(flow, t0, T, true, x, nx, y, ny, nullptr, 0, 2, params, nparams, method, rtol, atol,
                    mask, 0, out, nullptr, status, steps, stats, static_cast<cudaStream_t>(stream), aux);
    });
}

int b200cs_flowmap_grid_2d_series(int flow, const double *t0s, int64_t nt, double T, const double *x,
                                  int64_t nx, const double *y, int64_t ny, const double *params, int nparams,
                                  int method, double rtol, double atol, const uint8_t *mask, double *out,
                                  int32_t *status, int32_t *steps, int64_t *stats, void *stream) {
    return guarded([&] {
        B2_REQUIRE(nx >= 0 && ny >= 0 && nt >= 0, "negative size");
        B2_REQUIRE(t0s != nullptr || nt == 0, "t0s is null");
        if (nt == 0) return;
        AuxSpec aux;
        aux.t0s = t0s;
        aux.nt = nt;
        run_flowmap(flow, 0.0, T, true, x, nx, y, ny, nullptr, 0, 2, params, nparams, method, rtol, atol,
                    mask, 0, out, nullptr, status, steps, stats, static_cast<cudaStream_t>(stream), aux);
    });
}

int b200cs_ftle_series_2d(const double *flowmaps, int64_t nt, int64_t nx, int64_t ny, double T, double dx,
                          double dy, const uint8_t *mask, double *ftle, void *stream) {
    return guarded([&] {
        require_device();
        B2_REQUIRE(flowmaps && ftle, "null argument");
        B2_REQUIRE(nt >= 0 && nx >= 0 && ny >= 0, "negative size");
        B2_REQUIRE(T != 0.0 && dx != 0.0 && dy != 0.0, "T, dx and dy must be non-zero");
        if (nt == 0 || nx == 0 || ny == 0) return;
        cudaStream_t s = static_cast<cudaStream_t>(stream);
        const size_t np = (size_t)nx * ny;
        In<double> dfm(flowmaps, np * 2 * nt, s);
        In<uint8_t> dmask(mask, np, s);
        Out<double> dout(ftle, np * nt, s);
        for (int64_t f0 = 0; f0 < nt; f0 += 65535) {
            const int64_t nf = (nt - f0 < 65535) ? nt - f0 : 65535;
            launch_ftle(dfm.dev + f0 * np * 2, nx, ny, T, dx, dy, dmask.dev, dout.dev + f0 * np, 0, nx, true,
                        true, s, nf);
        }
        dout.download();
        if (dout.staged()) B2_CHECK_CUDA(cudaStreamSynchronize(s));
    });
}

int b200cs_c_tensor_2d(const double *flowmap_aux, int64_t nx, int64_t ny, int n_aux, double dx, double dy,
                       double h, const uint8_t *mask, double *C_out, void *stream) {
    return guarded([&] {
        require_device();
        (void)dx;
        (void)dy;  // unused by the reference too (diagnostics.py:68-112)
        B2_REQUIRE(flowmap_aux && C_out, "null argument");
        B2_REQUIRE(nx >= 0 && ny >= 0, "negative grid size");
        B2_REQUIRE(n_aux == 4 || n_aux == 5, "n_aux must be 4 or 5 (got %d)", n_aux);
        B2_REQUIRE(h != 0.0, "h must be non-zero");
        if (nx == 0 || ny == 0) return;
        cudaStream_t s = static_cast<cudaStream_t>(stream);
        const size_t np = (size_t)nx * ny;
        In<double> dfa(flowmap_aux, np * n_aux * 2, s);
        In<uint8_t> dmask(mask, np, s);
        Out<double> dout(C_out, np * 3, s);
        launch_c_tensor(dfa.dev, nx, ny, n_aux, h, dmask.dev, dout.dev, s);
        dout.download();
        if (dout.staged()) B2_CHECK_CUDA(cudaStreamSynchronize(s));
    });
}

int b200cs_c_eig_2d(const double *flowmap, int64_t nx, int64_t ny, double dx, double dy, const uint8_t *mask,
                    double *eigvals, double *eigvecs, void *stream) {
    return guarded([&] {
        require_device();
        B2_REQUIRE(flowmap && eigvals && eigvecs, "null argument");
        B2_REQUIRE(nx >= 0 && ny >= 0, "negative grid size");
        B2_REQUIRE(dx != 0.0 && dy != 0.0, "dx and dy must be non-zero");
        if (nx == 0 || ny == 0) return;
        cudaStream_t s = static_cast<cudaStream_t>(stream);
        const size_t np = (size_t)nx * ny;
        In<double> dfm(flowmap, np * 2, s);
        In<uint8_t> dmask(mask, np, s);
        Out<double> dvals(eigvals, np * 2, s), dvecs(eigvecs, np * 4, s);
        launch_c_eig(dfm.dev, nx, ny, 1, 0.0, dx, dy, /*aux_vecs=*/false, /*main_vals=*/true, dmask.dev,
                     dvals.dev, dvecs.dev, s);
        dvals.download();
        dvecs.download();
        if (dvals.staged() || dvecs.staged()) B2_CHECK_CUDA(cudaStreamSynchronize(s));
    });
}

int b200cs_c_eig_ftle_2d(const double *flowmap, int64_t nx, int64_t ny, double dx, double dy, double T,
                         const uint8_t *mask, double *eigvals, double *eigvecs, double *ftle, void *stream) {
    return guarded([&] {
        require_device();
        B2_REQUIRE(flowmap && eigvals && eigvecs && ftle, "null argument");
        B2_REQUIRE(nx >= 0 && ny >= 0, "negative grid size");
        B2_REQUIRE(dx != 0.0 && dy != 0.0 && T != 0.0, "dx, dy and T must be non-zero");
        if (nx == 0 || ny == 0) return;
        cudaStream_t s = static_cast<cudaStream_t>(stream);
        const size_t np = (size_t)nx * ny;
        In<double> dfm(flowmap, np * 2, s);
        In<uint8_t> dmask(mask, np, s);
        Out<double> dvals(eigvals, np * 2, s), dvecs(eigvecs, np * 4, s), dft(ftle, np, s);
        launch_c_eig(dfm.dev, nx, ny, 1, 0.0, dx, dy, /*aux_vecs=*/false, /*main_vals=*/true, dmask.dev,
                     dvals.dev, dvecs.dev, s, dft.dev, T);
        dvals.download();
        dvecs.download();
        dft.download();
        if (dvals.staged() || dvecs.staged() || dft.staged()) B2_CHECK_CUDA(cudaStreamSynchronize(s));
    });
}

int b200cs_c_eig_aux_2d(const double *flowmap_aux, int64_t nx, int64_t ny, int n_aux, double dx, double dy,
                        double h, int eig_main, const uint8_t *mask, double *eigvals, double *eigvecs,
                        void *stream) {
    return guarded([&] {
        require_device();
        B2_REQUIRE(flowmap_aux && eigvals && eigvecs, "null argument");
        B2_REQUIRE(nx >= 0 && ny >= 0, "negative grid size");
        B2_REQUIRE(n_aux == 4 || n_aux == 5, "n_aux must be 4 or 5 (got %d)", n_aux);
        B2_REQUIRE(!eig_main || n_aux == 5, "eig_main needs the centre point (n_aux = 5)");
        B2_REQUIRE(h != 0.0 && dx != 0.0 && dy != 0.0, "h, dx and dy must be non-zero");
        if (nx == 0 || ny == 0) return;
        cudaStream_t s = static_cast<cudaStream_t>(stream);
        const size_t np = (size_t)nx * ny;
        In<double> dfa(flowmap_aux, np * n_aux * 2, s);
        In<uint8_t> dmask(mask, np, s);
        Out<double> dvals(eigvals, np * 2, s), dvecs(eigvecs, np * 4, s);
        launch_c_eig(dfa.dev, nx, ny, n_aux, h, dx, dy, /*aux_vecs=*/true, /*main_vals=*/eig_main != 0,
                     dmask.dev, dvals.dev, dvecs.dev, s);
        dvals.download();
        dvecs.download();
        if (dvals.staged() || dvecs.staged()) B2_CHECK_CUDA(cudaStreamSynchronize(s));
    });
}

int b200cs_ftle_from_eig(const double *eigval_max, int64_t n, int64_t stride, double T, double *ftle,
                         void *stream) {
    return guarded([&] {
        require_device();
        B2_REQUIRE(eigval_max && ftle, "null argument");
        B2_REQUIRE(n >= 0 && stride >= 1, "bad size / stride");
        B2_REQUIRE(T != 0.0, "T must be non-zero");
        if (n == 0) return;
        cudaStream_t s = static_cast<cudaStream_t>(stream);
        In<double> din(eigval_max, (size_t)(n - 1) * stride + 1, s);
        Out<double> dout(ftle, n, s);
        launch_ftle_from_eig(din.dev, n, stride, T, dout.dev, s);
        dout.download();
        if (dout.staged()) B2_CHECK_CUDA(cudaStreamSynchronize(s));
    });
}

int b200cs_ftle_ridge_pts(const double *ftle, const double *eigvec_max, int64_t ev_pixel_stride,
                          int64_t ev_comp_stride, int64_t nx, int64_t ny, const double *x, const double *y,
                          double dx, double dy, double sdd_thresh, double f_min, double *r_pts, double *r_vec, double *sdd,
                          double *pts_compact, int64_t capacity, int64_t *count, void *stream) {
    return guarded([&] {
        require_device();
        B2_REQUIRE(ftle && eigvec_max && x && y, "null argument");
        B2_REQUIRE(nx >= 2 && ny >= 2, "the grid needs at least 2 points per axis");
        B2_REQUIRE(ev_pixel_stride >= 1 && ev_comp_stride >= 1, "bad eigenvector strides");
        B2_REQUIRE(!pts_compact || count, "pts_compact needs count");
        B2_REQUIRE(capacity >= 0, "negative capacity");
        cudaStream_t s = static_cast<cudaStream_t>(stream);
        const size_t np = (size_t)nx * ny;
        In<double> df(ftle, np, s);
        In<double> dev(eigvec_max, (np - 1) * ev_pixel_stride + ev_comp_stride + 1, s);
        B2_REQUIRE(dx != 0.0 && dy != 0.0, "dx and dy must be non-zero");
        In<double> dxs(x, nx, s), dys(y, ny, s);
        Out<double> drp(r_pts, np * 3, s), drv(r_vec, np * 2, s), dsdd(sdd, np, s);
        // compact output: a host buffer only receives the rows that were found
        Scratch cp_tmp, cnt_tmp;
        double *cp_dev = pts_compact;
        const bool cp_host = pts_compact && !is_device_ptr(pts_compact);
        if (cp_host && capacity > 0) {
            cp_tmp = Scratch((size_t)capacity * 2 * sizeof(double), s);
            cp_dev = static_cast<double *>(cp_tmp.ptr);
        }
        long long *cnt_dev = nullptr;
        const bool cnt_host = count && !is_device_ptr(count);
        if (count) {
            if (cnt_host) {
                cnt_tmp = Scratch(sizeof(long long), s);
                cnt_dev = static_cast<long long *>(cnt_tmp.ptr);
            } else {
                cnt_dev = reinterpret_cast<long long *>(count);
            }
        }
        launch_ridge_pts(df.dev, dev.dev, ev_pixel_stride, ev_comp_stride, nx, ny, dxs.dev, dys.dev,
                         dx, dy, sdd_thresh, f_min, drp.dev, drv.dev, dsdd.dev,
                         capacity > 0 ? cp_dev : nullptr, capacity, cnt_dev, s);
        drp.download();
        drv.download();
        dsdd.download();
        if (cnt_host || cp_host) {
            long long found = 0;
            B2_CHECK_CUDA(cudaMemcpyAsync(&found, cnt_dev, sizeof(found), cudaMemcpyDeviceToHost, s));
            B2_CHECK_CUDA(cudaStreamSynchronize(s));
            if (cnt_host) *count = found;
            const long long rows = found < capacity ? found : capacity;
            if (cp_host && rows > 0)
                B2_CHECK_CUDA(cudaMemcpyAsync(pts_compact, cp_dev, (size_t)rows * 2 * sizeof(double),
                                              cudaMemcpyDeviceToHost, s));
        }
        if (drp.staged() || drv.staged() || dsdd.staged() || cp_host)
            B2_CHECK_CUDA(cudaStreamSynchronize(s));
    });
}

int b200cs_ftle_ridges(const double *ftle, const double *eigvec_max, int64_t ev_pixel_stride,
                       int64_t ev_comp_stride, int64_t nx, int64_t ny, const double *x, const double *y,
                       double dx, double dy, double sdd_thresh, double f_min, double *pts_compact, int64_t *roots_compact,
                       int64_t capacity, int64_t *count, void *stream) {
    return guarded([&] {
        require_device();
        B2_REQUIRE(ftle && eigvec_max && x && y && count, "null argument");
        B2_REQUIRE(nx >= 2 && ny >= 2, "the grid needs at least 2 points per axis");
        B2_REQUIRE(ev_pixel_stride >= 1 && ev_comp_stride >= 1, "bad eigenvector strides");
        B2_REQUIRE(capacity >= 0 && (capacity == 0 || (pts_compact && roots_compact)), "bad output buffers");
        cudaStream_t s = static_cast<cudaStream_t>(stream);
        const size_t np = (size_t)nx * ny;
        In<double> df(ftle, np, s);
        In<double> dev(eigvec_max, (np - 1) * ev_pixel_stride + ev_comp_stride + 1, s);
        B2_REQUIRE(dx != 0.0 && dy != 0.0, "dx and dy must be non-zero");
        In<double> dxs(x, nx, s), dys(y, ny, s);
        const bool host_out = capacity > 0 && !is_device_ptr(pts_compact);
        B2_REQUIRE(capacity == 0 || host_out == !is_device_ptr(roots_compact),
                   "pts_compact and roots_compact must both be host or both be device buffers");
        Scratch cp_tmp, rt_tmp, cnt_tmp;
        double *cp_dev = pts_compact;
        long long *rt_dev = reinterpret_cast<long long *>(roots_compact);
        if (host_out) {
            cp_tmp = Scratch((size_t)capacity * 2 * sizeof(double), s);
            rt_tmp = Scratch((size_t)capacity * sizeof(long long), s);
            cp_dev = static_cast<double *>(cp_tmp.ptr);
            rt_dev = static_cast<long long *>(rt_tmp.ptr);
        }
        const bool cnt_host = !is_device_ptr(count);
        long long *cnt_dev = reinterpret_cast<long long *>(count);
        if (cnt_host) {
            cnt_tmp = Scratch(sizeof(long long), s);
            cnt_dev = static_cast<long long *>(cnt_tmp.ptr);
        }
        launch_ridge_components(df.dev, dev.dev, ev_pixel_stride, ev_comp_stride, nx, ny, dxs.dev, dys.dev,
                                dx, dy, sdd_thresh, f_min, cp_dev, rt_dev, capacity,
                                cnt_dev, s);
        if (cnt_host || host_out) {
            long long found = 0;
            B2_CHECK_CUDA(cudaMemcpyAsync(&found, cnt_dev, sizeof(found), cudaMemcpyDeviceToHost, s));
            B2_CHECK_CUDA(cudaStreamSynchronize(s));
            if (cnt_host) *count = found;
            const long long rows = found < capacity ? found : capacity;
            if (host_out && rows > 0) {
                B2_CHECK_CUDA(cudaMemcpyAsync(pts_compact, cp_dev, (size_t)rows * 2 * sizeof(double),
                                              cudaMemcpyDeviceToHost, s));
                B2_CHECK_CUDA(cudaMemcpyAsync(roots_compact, rt_dev, (size_t)rows * sizeof(long long),
                                              cudaMemcpyDeviceToHost, s));
                B2_CHECK_CUDA(cudaStreamSynchronize(s));
            }
        }
    });
}

int b200cs_link_ridge_pts(const double *r_pts, const double *r_vec, const double *sdd, int64_t nx, int64_t ny,
                          double h, double c, double sdd_thresh, double *linked, int64_t linked_capacity,
                          int32_t *ridge_len, double *endpoints, double *ep_tanvecs, int64_t curve_capacity,
                          int64_t *counts) {
    return guarded([&] {
        B2_REQUIRE(r_pts && r_vec && sdd && counts, "null argument");
        B2_REQUIRE(nx >= 5 && ny >= 5, "the grid needs at least 5 points per axis (ridge points live on [2, n-2))");
        B2_REQUIRE(h > 0.0, "h must be positive");
        B2_REQUIRE(!is_device_ptr(r_pts) && !is_device_ptr(r_vec) && !is_device_ptr(sdd),
                   "the linking stage is host code: pass host arrays (download r_pts / r_vec / sdd first)");
        std::vector<double> lk, ep, tv;
        std::vector<int32_t> rl;
        link_ridge_points(r_pts, r_vec, sdd, nx, ny, h, c, sdd_thresh, lk, rl, ep, tv);
        const int64_t n_pts = (int64_t)lk.size() / 2, n_curves = (int64_t)rl.size() / 2;
        counts[0] = n_pts;
        counts[1] = n_curves;
        if (n_pts > linked_capacity || n_curves > curve_capacity) return;   // sizes only: call again with room
        B2_REQUIRE((n_pts == 0 || linked) && (n_curves == 0 || (ridge_len && endpoints && ep_tanvecs)),
                   "null output buffer");
        if (n_pts) std::memcpy(linked, lk.data(), lk.size() * sizeof(double));
        if (n_curves) {
            std::memcpy(ridge_len, rl.data(), rl.size() * sizeof(int32_t));
            std::memcpy(endpoints, ep.data(), ep.size() * sizeof(double));
            std::memcpy(ep_tanvecs, tv.data(), tv.size() * sizeof(double));
        }
    });
}

int b200cs_order_ridges(const double *linked, int64_t n_pts, const int32_t *ridge_len, const double *endpoints,
                        const double *ep_tanvecs, int64_t n_curves, double dist_tol, double ep_tan_ang,
                        int64_t min_ridge_pts, double *out_pts, int64_t *offsets, int64_t *n_out) {
    return guarded([&] {
        B2_REQUIRE(n_out && n_pts >= 0 && n_curves >= 0, "bad argument");
        *n_out = 0;
        if (n_curves == 0) return;
        B2_REQUIRE(linked && ridge_len && endpoints && ep_tanvecs && out_pts && offsets, "null argument");
        for (int64_t r = 0; r < n_curves; ++r)
            B2_REQUIRE(ridge_len[2 * r + 1] >= 2 && ridge_len[2 * r] >= ridge_len[2 * r + 1] && ridge_len[2 * r] <= n_pts,
                       "ridge_len[%lld] = (%d, %d) does not index linked[%lld]", (long long)r, ridge_len[2 * r],
                       ridge_len[2 * r + 1], (long long)n_pts);
        std::vector<double> lk(linked, linked + 2 * n_pts), ep(endpoints, endpoints + 6 * n_curves),
            tv(ep_tanvecs, ep_tanvecs + 4 * n_curves), out;
        std::vector<int32_t> rl(ridge_len, ridge_len + 2 * n_curves);
        std::vector<long long> offs;
        order_ridges(lk, rl, ep, tv, dist_tol, ep_tan_ang, min_ridge_pts, out, offs);
        // every curve is used at most once, so out never exceeds n_pts points / n_curves curves
        *n_out = (int64_t)offs.size() - 1;
        if (!out.empty()) std::memcpy(out_pts, out.data(), out.size() * sizeof(double));
        for (size_t k = 0; k < offs.size(); ++k) offsets[k] = offs[k];
    });
}

int b200cs_flowmap_composition(const double *flowmaps, const double *grid6, int64_t nT, double *composed,
                               void *stream) {
    return guarded([&] {
        require_device();
        B2_REQUIRE(flowmaps && grid6 && composed, "null argument");
        B2_REQUIRE(nT >= 1, "nT must be at least 1 (got %lld)", (long long)nT);
        double g[6];
        B2_CHECK_CUDA(cudaMemcpy(g, grid6, sizeof(g), cudaMemcpyDefault));
        const long long nx = (long long)g[2], ny = (long long)g[5];
        B2_REQUIRE(nx >= 2 && ny >= 2, "the grid needs at least 2 points per axis");
        B2_REQUIRE(g[1] > g[0] && g[4] > g[3], "grid axes must be ascending");
        cudaStream_t s = static_cast<cudaStream_t>(stream);
        const size_t np = (size_t)nx * ny;
        In<double> dfm(flowmaps, np * 2 * nT, s);
        Out<double> dout(composed, np * 2, s);
        launch_composition(dfm.dev, g, nT, dout.dev, s);
        dout.download();
        if (dout.staged()) B2_CHECK_CUDA(cudaStreamSynchronize(s));
    });
}

int b200cs_binary_mask_dilation(const uint8_t *mask, int64_t nx, int64_t ny, int corners, uint8_t *dilated,
                                void *stream) {
    return guarded([&] {
        require_device();
        B2_REQUIRE(mask && dilated, "null argument");
        B2_REQUIRE(nx >= 0 && ny >= 0, "negative grid size");
        if (nx == 0 || ny == 0) return;
        cudaStream_t s = static_cast<cudaStream_t>(stream);
        const size_t np = (size_t)nx * ny;
        In<uint8_t> din(mask, np, s);
        Out<uint8_t> dout(dilated, np, s);
        B2_REQUIRE(din.dev != dout.dev, "in-place dilation is not supported");
        launch_mask_dilation(din.dev, nx, ny, corners != 0, dout.dev, s);
        dout.download();
        if (dout.staged()) B2_CHECK_CUDA(cudaStreamSynchronize(s));
    });
}

int b200cs_flowmap_composition_series(const double *flowmaps, const double *grid6, int64_t nT, int64_t nframes,
                                      double *composed, void *stream) {
    return guarded([&] {
        require_device();
        B2_REQUIRE(flowmaps && grid6 && composed, "null argument");
        B2_REQUIRE(nT >= 1 && nframes >= 0, "nT must be >= 1 and nframes >= 0");
        if (nframes == 0) return;
        double g[6];
        B2_CHECK_CUDA(cudaMemcpy(g, grid6, sizeof(g), cudaMemcpyDefault));
        const long long nx = (long long)g[2], ny = (long long)g[5];
        B2_REQUIRE(nx >= 2 && ny >= 2, "the grid needs at least 2 points per axis");
        B2_REQUIRE(g[1] > g[0] && g[4] > g[3], "grid axes must be ascending");
        cudaStream_t s = static_cast<cudaStream_t>(stream);
        const size_t np = (size_t)nx * ny;
        In<double> dfm(flowmaps, np * 2 * (size_t)(nT + nframes - 1), s);
        Out<double> dout(composed, np * 2 * (size_t)nframes, s);
        for (int64_t f0 = 0; f0 < nframes; f0 += 65535) {
            const int64_t nf = (nframes - f0 < 65535) ? nframes - f0 : 65535;
            launch_composition(dfm.dev + f0 * np * 2, g, nT, dout.dev + f0 * np * 2, s, nf);
        }
        dout.download();
        if (dout.staged()) B2_CHECK_CUDA(cudaStreamSynchronize(s));
    });
}

int b200cs_order_stats(const double *data, int64_t n, int64_t k, double *out2, void *stream) {
    return guarded([&] {
        require_device();
        B2_REQUIRE(data && out2, "null argument");
        B2_REQUIRE(n >= 1 && k >= 0 && k < n, "need 0 <= k < n (k = %lld, n = %lld)", (long long)k, (long long)n);
        cudaStream_t s = static_cast<cudaStream_t>(stream);
        In<double> din(data, n, s);
        Out<double> dout(out2, 2, s);
        launch_order_stats(din.dev, n, k, dout.dev, s);
        dout.download();
        if (dout.staged()) B2_CHECK_CUDA(cudaStreamSynchronize(s));
    });
}

int b200cs_fp64_peak(int iters, double *out_tflops, double *out_ms) {
    return guarded([&] {
        require_device();
        B2_REQUIRE(out_tflops && iters > 0, "bad argument");
        double ms = 0.0;
        run_fp64_peak(iters, out_tflops, &ms);
        if (out_ms) *out_ms = ms;
    });
}

}  // extern "C"
