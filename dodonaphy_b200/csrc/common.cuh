// Shared helpers for the dodonaphy_b200 CUDA library (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/dodonaphy_b200.h"

namespace dodo {

void set_error(const char *fmt, ...);
void count_launch(int n = 1);
int sm_count();
void timing_events(cudaEvent_t *start, cudaEvent_t *stop);

#define DODO_CHECK_ARG(cond, ...)            \
    do {                                     \
        if (!(cond)) {                       \
            dodo::set_error(__VA_ARGS__);    \
            return DODO_EINVAL;              \
        }                                    \
    } while (0)

#define DODO_CUDA(call)                                                                     \
    do {                                                                                    \
        cudaError_t e__ = (call);                                                           \
        if (e__ != cudaSuccess) {                                                           \
            dodo::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call,                   \
                            cudaGetErrorString(e__));                                       \
            return DODO_ECUDA;                                                              \
        }                                                                                   \
    } while (0)

#define DODO_LAUNCH_CHECK()                                                                 \
    do {                                                                                    \
        dodo::count_launch();                                                               \
        cudaError_t e__ = cudaGetLastError();                                               \
        if (e__ != cudaSuccess) {                                                           \
            dodo::set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__,               \
                            cudaGetErrorString(e__));                                       \
            return DODO_ECUDA;                                                              \
        }                                                                                   \
    } while (0)

__device__ __forceinline__ double shfl_xor_d(double v, int m) {
    return __shfl_xor_sync(0xffffffffu, v, m);
}

// butterfly sum: every lane ends with the same total, fixed order -> reproducible
__device__ __forceinline__ double warp_sum(double v) {
    v += shfl_xor_d(v, 16);
    v += shfl_xor_d(v, 8);
    v += shfl_xor_d(v, 4);
    v += shfl_xor_d(v, 2);
    v += shfl_xor_d(v, 1);
    return v;
}

__device__ __forceinline__ double2 ldg2(const double *p) {
    return __ldg(reinterpret_cast<const double2 *>(p));
}

static inline uint64_t align_up(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }

}  // namespace dodo
