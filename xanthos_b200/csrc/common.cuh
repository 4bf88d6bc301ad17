// Shared helpers of libxanthos_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cstdio>
#include <cstdarg>
#include <cmath>
#include <cfloat>

#include "../../include/xanthos_b200.h"

namespace xan {

void set_error(const char *fmt, ...);
// cudaMallocAsync from the device's default pool, configured to keep freed memory cached (core.cu)
cudaError_t scratch_alloc(void **p, size_t bytes, cudaStream_t s);
template <typename T>
inline cudaError_t scratch_alloc(T **p, size_t bytes, cudaStream_t s) {
    return scratch_alloc(reinterpret_cast<void **>(p), bytes, s);
}

#define XAN_CUDA_CHECK(expr)                                                                  \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            xan::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,               \
                           cudaGetErrorString(_e));                                           \
            return XAN_E_CUDA;                                                                \
        }                                                                                     \
    } while (0)

#define XAN_REQUIRE(cond, ...)                                                                \
    do {                                                                                      \
        if (!(cond)) {                                                                        \
            xan::set_error(__VA_ARGS__);                                                      \
            return XAN_E_INVALID;                                                             \
        }                                                                                     \
    } while (0)

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Gregorian month length (calendar.monthrange; penman_monteith.py:57, hargreaves_samani.py:26,
// thornthwaite.py:113).  moy in 0..11.
__host__ __device__ inline bool is_leap_gregorian(int y) {
    return (y % 4 == 0) && ((y % 100 != 0) || (y % 400 == 0));
}
__host__ __device__ inline int month_days(int moy, bool leap) {
    const int d = (moy == 1) ? 28 : ((moy == 3 || moy == 5 || moy == 8 || moy == 10) ? 30 : 31);
    return (moy == 1 && leap) ? 29 : d;
}

// numpy.nan_to_num for float64
__device__ __forceinline__ double nan_to_num(double v) {
    if (isnan(v)) return 0.0;
    if (isinf(v)) return v > 0 ? DBL_MAX : -DBL_MAX;
    return v;
}

// Streaming (read-once) global load that does not allocate in L1.
__device__ __forceinline__ double ldg_stream(const double *p) {
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float ldg_stream(const float *p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
// Streaming store (evict-first): outputs are written once and not re-read by the same kernel.
__device__ __forceinline__ void stg_stream(double *p, double v) {
    asm volatile("st.global.cs.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}

// numpy's pairwise summation order for a contiguous reduction of k <= 128 elements
// (see oracle/pet.py:numpy_pairwise_sum).  `get(i)` returns element i.
template <typename F>
__device__ __forceinline__ double numpy_pairwise_sum(int k, F get) {
    if (k < 8) {
        double r = 0.0;
        for (int i = 0; i < k; ++i) r = r + get(i);
        return r;
    }
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = get(j);
    int i = 8;
    const int kb = k - (k % 8);
    for (; i < kb; i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = r[j] + get(i + j);
    }
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < k; ++i) res = res + get(i);
    return res;
}

}  // namespace xan
