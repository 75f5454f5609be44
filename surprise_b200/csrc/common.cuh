// Shared helpers for the surprise_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/surprise_b200.h"

namespace sb2 {

void set_error(const char* fmt, ...);
int64_t& launch_counter();

#define SB2_CUDA(expr)                                                                            \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            sb2::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return SB2_ERR_CUDA;                                                                  \
        }                                                                                         \
    } while (0)

#define SB2_LAUNCH_CHECK()                                                                        \
    do {                                                                                          \
        sb2::launch_counter()++;                                                                  \
        cudaError_t _e = cudaGetLastError();                                                      \
        if (_e != cudaSuccess) {                                                                  \
            sb2::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
            return SB2_ERR_CUDA;                                                                  \
        }                                                                                         \
    } while (0)

#define SB2_TRY(expr)                 \
    do {                              \
        int _rc = (expr);             \
        if (_rc != SB2_OK) return _rc; \
    } while (0)

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

int sm_count();

// RAII device buffer on a stream (cudaMallocAsync / cudaFreeAsync).
struct DevBuf {
    void* p = nullptr;
    cudaStream_t s = nullptr;
    size_t bytes = 0;
    DevBuf() {}
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    int alloc(size_t n, cudaStream_t st) {
        release();
        s = st;
        bytes = n;
        if (n == 0) return SB2_OK;
        SB2_CUDA(cudaMallocAsync(&p, n, st));
        return SB2_OK;
    }
    void release() {
        if (p) cudaFreeAsync(p, s);
        p = nullptr;
    }
    ~DevBuf() { release(); }
    template <class T>
    T* as() const { return reinterpret_cast<T*>(p); }
};

// Host->device upload helper for the host-buffer entry points.
template <class T>
inline int upload(DevBuf& b, const T* host, size_t n, cudaStream_t st) {
    SB2_TRY(b.alloc(n * sizeof(T) + 16, st));
    if (n) SB2_CUDA(cudaMemcpyAsync(b.p, host, n * sizeof(T), cudaMemcpyHostToDevice, st));
    return SB2_OK;
}

}  // namespace sb2
