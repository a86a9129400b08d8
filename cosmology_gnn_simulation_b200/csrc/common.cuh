// Shared helpers for libcgnn (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/cgnn.h"

namespace cgnn {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

// SM count of the current device (cached)
int num_sms();

#define CGNN_CHECK_ARG(cond, ...)                                   \
    do {                                                            \
        if (!(cond)) {                                              \
            ::cgnn::set_error(__VA_ARGS__);                         \
            return CGNN_ERR_INVALID;                                \
        }                                                           \
    } while (0)

#define CGNN_CUDA(call)                                                                   \
    do {                                                                                  \
        cudaError_t _e = (call);                                                          \
        if (_e != cudaSuccess) {                                                          \
            ::cgnn::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,        \
                              cudaGetErrorString(_e));                                    \
            return CGNN_ERR_CUDA;                                                         \
        }                                                                                 \
    } while (0)

#define CGNN_LAUNCH_CHECK()                                                               \
    do {                                                                                  \
        ::cgnn::count_launch();                                                           \
        cudaError_t _e = cudaGetLastError();                                              \
        if (_e != cudaSuccess) {                                                          \
            ::cgnn::set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__,    \
                              cudaGetErrorString(_e));                                    \
            return CGNN_ERR_CUDA;                                                         \
        }                                                                                 \
    } while (0)

static inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

// Carves 256-byte aligned sub-buffers out of a caller-provided workspace.
struct Carver {
    char* base;
    int64_t off = 0;
    explicit Carver(void* p) : base(static_cast<char*>(p)) {}
    template <typename T>
    T* take(int64_t count) {
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += align_up(count * (int64_t)sizeof(T), 256);
        return p;
    }
};

}  // namespace cgnn
