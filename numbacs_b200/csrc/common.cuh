// common.cuh -- shared plumbing of libb200cs: error handling, host/device pointer staging, the
// flow registry.  Nothing here is on the hot path.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "../../include/b200cs.h"

namespace b200cs {

// ---------------------------------------------------------------- errors
void set_error(const char *fmt, ...);

struct Fail {  // thrown internally, converted to an error code at the ABI
    int code;
};

#define B2_CHECK_CUDA(expr)                                                                    \
    do {                                                                                       \
        cudaError_t e__ = (expr);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            ::b200cs::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__),       \
                                __FILE__, __LINE__);                                           \
            throw ::b200cs::Fail{B200CS_E_CUDA};                                               \
        }                                                                                      \
    } while (0)

#define B2_REQUIRE(cond, ...)                                                                  \
    do {                                                                                       \
        if (!(cond)) {                                                                         \
            ::b200cs::set_error(__VA_ARGS__);                                                  \
            throw ::b200cs::Fail{B200CS_E_INVALID};                                            \
        }                                                                                      \
    } while (0)

template <class F>
int guarded(F &&f) {
    try {
        f();
        return B200CS_OK;
    } catch (const Fail &e) {
        return e.code;
    } catch (const std::exception &e) {
        set_error("internal error: %s", e.what());
        return B200CS_E_INVALID;
    }
}

// ---------------------------------------------------------------- pointer staging
bool is_device_ptr(const void *p);
bool is_pinned_host_ptr(const void *p);
void upload_to_scratch(void *dst, const void *src, size_t bytes, cudaStream_t s);   // capi.cu

// Stream-ordered scratch buffer (cudaMallocAsync), freed on the same stream.
struct Scratch {
    void *ptr = nullptr;
    cudaStream_t stream = nullptr;
    Scratch() = default;
    Scratch(size_t bytes, cudaStream_t s) : stream(s) {
        if (bytes) B2_CHECK_CUDA(cudaMallocAsync(&ptr, bytes, s));
    }
    Scratch(const Scratch &) = delete;
    Scratch &operator=(const Scratch &) = delete;
    Scratch(Scratch &&o) noexcept : ptr(o.ptr), stream(o.stream) { o.ptr = nullptr; }
    Scratch &operator=(Scratch &&o) noexcept {
        if (this != &o) {
            if (ptr) cudaFreeAsync(ptr, stream);
            ptr = o.ptr;
            stream = o.stream;
            o.ptr = nullptr;
        }
        return *this;
    }
    ~Scratch() {
        if (ptr) cudaFreeAsync(ptr, stream);
    }
};

// Input array: device pointers pass through, host pointers are uploaded to scratch.
template <class T>
struct In {
    const T *dev = nullptr;
    Scratch tmp;
    In() = default;
    In(const T *p, size_t count, cudaStream_t s) {
        if (!p || !count) return;
        if (is_device_ptr(p)) {
            dev = p;
        } else {
            tmp = Scratch(count * sizeof(T), s);
            upload_to_scratch(tmp.ptr, p, count * sizeof(T), s);
            dev = static_cast<const T *>(tmp.ptr);
        }
    }
    In(In &&) = default;
    In &operator=(In &&) = default;
};

// Output array: device pointers pass through; for a host pointer a device buffer is allocated and
// download() copies it back (async on the stream; the API call synchronises afterwards).
template <class T>
struct Out {
    T *dev = nullptr;
    T *host = nullptr;
    size_t count = 0;
    Scratch tmp;
    cudaStream_t stream = nullptr;
    Out() = default;
    Out(T *p, size_t n, cudaStream_t s, bool upload_first = false) : count(n), stream(s) {
        if (!p || !n) return;
        if (is_device_ptr(p)) {
            dev = p;
        } else {
            host = p;
            tmp = Scratch(n * sizeof(T), s);
            dev = static_cast<T *>(tmp.ptr);
            if (upload_first) upload_to_scratch(dev, p, n * sizeof(T), s);
        }
    }
    Out(Out &&) = default;
    Out &operator=(Out &&) = default;
    bool staged() const { return host != nullptr; }
    void download() {
        if (host)
            B2_CHECK_CUDA(cudaMemcpyAsync(host, dev, count * sizeof(T), cudaMemcpyDeviceToHost, stream));
    }
};

// ---------------------------------------------------------------- flow registry
struct Axis3 {
    double a[3], b[3];  // grid end points per axis (t, x, y)
    int n[3];           // data points per axis
};

struct FlowSpec {
    int kind = -1;  // B200CS_FLOW_* ; -1 = scalar field
    int ndim = 2;
    int min_params = 1;
    int device = 0;
    // spline / scalar payload
    Axis3 grid{};
    int spherical = 0;
    int extrap = 0;
    int linear = 0;  // scalar fields only: trilinear on raw data
    double r = 6371.0;
    void *coef = nullptr;  // device: double2 (u,v) interleaved for flows, double for scalars
    size_t coef_bytes = 0;
    // interpolated flows: device counter of right-hand-side evaluations that fell outside the data
    // grid (where the extrapolation modes, unpinned by any reference test, decide the value)
    unsigned long long *oog = nullptr;
    ~FlowSpec() {
        if (coef) cudaFree(coef);
        if (oog) cudaFree(oog);
    }
};

int registry_add(std::shared_ptr<FlowSpec> f);
std::shared_ptr<FlowSpec> registry_get(int handle);  // throws Fail{B200CS_E_HANDLE}
bool registry_remove(int handle);

}  // namespace b200cs
