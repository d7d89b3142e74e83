// Shared helpers for the gapro_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/gapro_b200.h"

void gapro_set_error(const char* fmt, ...);

#define GAPRO_CUDA_TRY(expr)                                                                              \
    do {                                                                                                  \
        cudaError_t e__ = (expr);                                                                         \
        if (e__ != cudaSuccess) {                                                                         \
            gapro_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
            return GAPRO_ERR_CUDA;                                                                        \
        }                                                                                                 \
    } while (0)

#define GAPRO_KERNEL_CHECK() GAPRO_CUDA_TRY(cudaGetLastError())

#define GAPRO_REQUIRE(cond, ...)            \
    do {                                    \
        if (!(cond)) {                      \
            gapro_set_error(__VA_ARGS__);   \
            return GAPRO_ERR_INVALID;       \
        }                                   \
    } while (0)

static inline size_t gapro_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// largest s with off[s] <= v, for a non-decreasing offsets array off[0..n]
template <typename T>
__device__ __forceinline__ int gapro_find_segment(const T* __restrict__ off, int n, T v) {
    int lo = 0, hi = n;   // invariant: off[lo] <= v < off[hi]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (off[mid] <= v) lo = mid; else hi = mid;
    }
    return lo;
}
