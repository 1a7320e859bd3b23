// common.hpp — shared plumbing of libr4w_b200.so (error text, CUDA checks, launch counter, stream).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

#include "../../include/r4w_b200.h"

namespace r4wb {

// thrown inside the library, converted to r4wb_error at the C boundary (nothing unwinds across it)
struct Failure {
    r4wb_error code;
    std::string what;
};

[[noreturn]] inline void fail(r4wb_error code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    throw Failure{code, buf};
}

#define R4WB_CUDA(expr)                                                                          \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess)                                                                   \
            ::r4wb::fail(R4WB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                         __FILE__, __LINE__);                                                    \
    } while (0)

extern std::atomic<uint64_t> g_kernel_launches;   // every <<<>>> of this library bumps it
cudaStream_t current_stream();                     // thread-local, set by r4wb_set_stream

inline void count_launch(int n = 1) { g_kernel_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

#define R4WB_LAUNCH_CHECK()                 \
    do {                                    \
        ::r4wb::count_launch();             \
        R4WB_CUDA(cudaGetLastError());      \
    } while (0)

// One-time per-DEVICE set-up of a kernel (function attributes such as the dynamic shared-memory limit belong to the device's
// context, and a process may drive several devices: r4wb_init_devices).  `static PerDeviceOnce once; if (once.first()) ...`
struct PerDeviceOnce {
    std::atomic<unsigned long long> mask{0};
    bool first()
    {
        int d = 0;
        cudaGetDevice(&d);
        const unsigned long long bit = 1ull << (d & 63);
        return !(mask.fetch_or(bit) & bit);
    }
};

template <typename T>
struct DevBuf {   // owning device allocation, grows on demand
    T* p = nullptr;
    size_t cap = 0;
    ~DevBuf() { if (p) cudaFree(p); }
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    T* reserve(size_t n)
    {
        if (n > cap) {
            if (p) cudaFree(p);
            p = nullptr;
            cap = 0;
            cudaError_t e = cudaMalloc(&p, n * sizeof(T));
            if (e != cudaSuccess) fail(R4WB_ERR_ALLOCATION_FAILED, "cudaMalloc(%zu bytes): %s", n * sizeof(T), cudaGetErrorString(e));
            cap = n;
        }
        return p;
    }
};

}  // namespace r4wb
