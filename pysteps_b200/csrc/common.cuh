// common.cuh -- shared helpers for libpysteps_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/pysteps_b200.h"

namespace b200 {

void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);
int num_sms();
void count_launch();

#define B200_CUDA(call)                                                    \
    do {                                                                   \
        cudaError_t _e = (call);                                           \
        if (_e != cudaSuccess)                                             \
            return ::b200::cuda_fail(_e, #call, __FILE__, __LINE__);       \
    } while (0)

#define B200_LAUNCH_CHECK()                                                \
    do {                                                                   \
        cudaError_t _e = cudaGetLastError();                               \
        ::b200::count_launch();                                            \
        if (_e != cudaSuccess)                                             \
            return ::b200::cuda_fail(_e, "kernel launch", __FILE__, __LINE__); \
    } while (0)

#define B200_REQUIRE(cond, msg)                                            \
    do {                                                                   \
        if (!(cond)) {                                                     \
            ::b200::set_error("%s (%s:%d)", msg, __FILE__, __LINE__);      \
            return B200_EINVAL;                                            \
        }                                                                  \
    } while (0)

// stream-ordered scratch allocation that is released on scope exit
struct Scratch {
    void *p = nullptr;
    cudaStream_t s = nullptr;
    cudaError_t alloc(size_t bytes, cudaStream_t stream) {
        s = stream;
        return cudaMallocAsync(&p, bytes ? bytes : 1, stream);
    }
    ~Scratch() {
        if (p) cudaFreeAsync(p, s);
    }
};

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace b200
