// Shared helpers of libibt.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/ibt.h"

#define IBT_API extern "C" __attribute__((visibility("default")))

namespace ibt {

void set_last_error(cudaError_t e, const char *where);

inline int check_launch(const char *where)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_last_error(e, where); return IBT_E_CUDA; }
    return IBT_OK;
}

#define IBT_CUDA_TRY(expr)                                                        \
    do {                                                                          \
        cudaError_t e__ = (expr);                                                 \
        if (e__ != cudaSuccess) { ::ibt::set_last_error(e__, #expr); return IBT_E_CUDA; } \
    } while (0)

// REFLECT_101 index for any i (period 2(n-1)).  The common cases (inside, or one reflection) cost a few compares;
// indices further out (tiny images under big windows) take the out-of-line modulo path.
__host__ __device__ __noinline__ inline int r101_far(int i, int n)
{
    if (n == 1) return 0;
    const int p = 2 * (n - 1);
    i %= p;
    if (i < 0) i += p;
    return i < n ? i : p - i;
}
__host__ __device__ __forceinline__ int r101(int i, int n)
{
    if (i >= 0 && i < n) return i;
    if (i < 0) i = -i;                   // -1 -> 1
    if (i < n) return i;
    const int j = 2 * (n - 1) - i;       // n -> n-2
    if (j >= 0) return j;
    return r101_far(i, n);
}

constexpr int kNumSMs = 148;             // B200: 2 dies x 74 SMs

} // namespace ibt
