// Shared helpers for the hy2dl_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <algorithm>
#include "../../include/hy2dl_b200.h"

namespace hy2dl {

void set_error(const char* fmt, ...);

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

#define HY_REQUIRE(cond, code, ...)                 \
  do {                                              \
    if (!(cond)) {                                  \
      ::hy2dl::set_error(__VA_ARGS__);              \
      return (code);                                \
    }                                               \
  } while (0)

#define HY_PTR(p)                                                                                   \
  HY_REQUIRE((p) != nullptr && ::hy2dl::aligned16(p), HY2DL_ERR_ALIGN, "%s: pointer '%s' is null or not 16-byte aligned", __func__, #p)

// arrays only touched with 4-byte accesses: any float alignment will do
#define HY_NONNULL(p) HY_REQUIRE((p) != nullptr, HY2DL_ERR_ALIGN, "%s: pointer '%s' is null", __func__, #p)

#define HY_PTR_OPT(p) \
  HY_REQUIRE((p) == nullptr || ::hy2dl::aligned16(p), HY2DL_ERR_ALIGN, "%s: pointer '%s' is not 16-byte aligned", __func__, #p)

#define HY_LAUNCH_CHECK()                                                                           \
  do {                                                                                              \
    cudaError_t e__ = cudaGetLastError();                                                           \
    if (e__ != cudaSuccess) {                                                                       \
      ::hy2dl::set_error("%s: launch failed: %s", __func__, cudaGetErrorString(e__));               \
      return HY2DL_ERR_LAUNCH;                                                                      \
    }                                                                                               \
  } while (0)

__host__ __device__ inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int64_t round_up(int64_t a, int64_t b) { return cdiv(a, b) * b; }

int sm_count();

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace hy2dl
