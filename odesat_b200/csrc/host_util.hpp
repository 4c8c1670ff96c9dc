// host_util.hpp — error plumbing and small RAII helpers for the C-ABI implementation.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/odesat_b200.h"

namespace odesat {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

inline std::string& last_error_ref() {
    thread_local std::string s;
    return s;
}

#define ODESAT_CUDA(expr)                                                                          \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            char _b[512];                                                                          \
            std::snprintf(_b, sizeof _b, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),   \
                          __FILE__, __LINE__);                                                     \
            throw ::odesat::Error(_e == cudaErrorMemoryAllocation ? ODESAT_ENOMEM : ODESAT_ECUDA, _b); \
        }                                                                                          \
    } while (0)

#define ODESAT_REQUIRE(cond, msg)                                                                  \
    do {                                                                                           \
        if (!(cond)) throw ::odesat::Error(ODESAT_EINVAL, std::string(msg));                       \
    } while (0)

// Runs f(), mapping exceptions to status codes; nothing ever crosses the C ABI.
template <typename F> inline int guarded(F&& f) {
    try {
        f();
        return ODESAT_OK;
    } catch (const Error& e) {
        last_error_ref() = e.what();
        return e.code;
    } catch (const std::bad_alloc&) {
        last_error_ref() = "host allocation failed";
        return ODESAT_ENOMEM;
    } catch (const std::exception& e) {
        last_error_ref() = e.what();
        return ODESAT_EINVAL;
    } catch (...) {
        last_error_ref() = "unknown error";
        return ODESAT_EINVAL;
    }
}

// Device buffer with a byte ledger (reported by odesat_batch_info).
template <typename T> struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void alloc(size_t count, int64_t* ledger = nullptr) {
        release();
        if (count == 0) count = 1;
        ODESAT_CUDA(cudaMalloc((void**)&p, count * sizeof(T)));
        n = count;
        if (ledger) *ledger += (int64_t)(count * sizeof(T));
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    size_t bytes() const { return n * sizeof(T); }
};

// Opt a kernel into its large dynamic shared memory once per (kernel instantiation, device): the attribute
// belongs to the device's copy of the function, so a second GPU used from the same process needs its own call.
// `mask` is a function-local static of the caller.
template <typename K> inline void ensure_max_smem(K kernel, int bytes, uint64_t& mask) {
    int dev = 0;
    ODESAT_CUDA(cudaGetDevice(&dev));
    const uint64_t bit = uint64_t(1) << (dev & 63);
    if (mask & bit) return;
    ODESAT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    mask |= bit;
}

inline int64_t pad32(int64_t r) { return (r + 31) / 32 * 32; }

}  // namespace odesat
