// Shared device/host helpers for the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mvpnet_b200.h"

namespace mvp {

// thread-local error text behind mvp_last_error()
void set_error(const char *fmt, ...);

// device counter of out-of-range gather/scatter indices (see mvp_index_errors_fetch_and_clear).
// One instance: the library is a single translation unit (lib.cu).
static __device__ unsigned long long g_index_errors = 0ULL;

inline int launch_status(const char *what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: CUDA launch failed: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

inline int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

#define MVP_REQUIRE(cond, code, ...)      \
  do {                                    \
    if (!(cond)) {                        \
      ::mvp::set_error(__VA_ARGS__);      \
      return (code);                      \
    }                                     \
  } while (0)

// ---- arithmetic contract -----------------------------------------------------------------------
// Squared distance, d = key - query, in the exact sequence nvcc -O2 emits for the reference's
// `dist = 0; dist += diff * diff` loops (ball_query_kernel.cu:111-116 et al.), read off the SASS of the
// reference sources built for sm_100a (oracle/build_ref.py) and checked bit-for-bit against those
// kernels in tests/test_gpu_vs_reference_kernels.py:
//   float :  fma(dz,dz, fma(dy,dy, dx*dx))        double:  fma(dz,dz, fma(dx,dx, dy*dy))
// Spelled with intrinsics so no compiler flag can change it.
__device__ __forceinline__ float sqdist3(float kx, float ky, float kz, float qx, float qy, float qz) {
  const float dx = __fsub_rn(kx, qx), dy = __fsub_rn(ky, qy), dz = __fsub_rn(kz, qz);
  return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
}
__device__ __forceinline__ double sqdist3(double kx, double ky, double kz, double qx, double qy, double qz) {
  const double dx = __dsub_rn(kx, qx), dy = __dsub_rn(ky, qy), dz = __dsub_rn(kz, qz);
  return __fma_rn(dz, dz, __fma_rn(dx, dx, __dmul_rn(dy, dy)));
}

// 2-D variant (FPS accepts (B, N, 2) points): float fma(dy,dy, dx*dx), double fma(dx,dx, dy*dy)
__device__ __forceinline__ float sqdist2(float kx, float ky, float qx, float qy) {
  const float dx = __fsub_rn(kx, qx), dy = __fsub_rn(ky, qy);
  return __fmaf_rn(dy, dy, __fmul_rn(dx, dx));
}
__device__ __forceinline__ double sqdist2(double kx, double ky, double qx, double qy) {
  const double dx = __dsub_rn(kx, qx), dy = __dsub_rn(ky, qy);
  return __fma_rn(dx, dx, __dmul_rn(dy, dy));
}

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }

template <typename T> struct Inf;
template <> struct Inf<float> { __device__ static float v() { return __int_as_float(0x7f800000); } };
template <> struct Inf<double> { __device__ static double v() { return __longlong_as_double(0x7ff0000000000000LL); } };

// Cooperative flat copy of `count` scalars global -> shared (whole CTA).
template <typename T>
__device__ __forceinline__ void stage_keys(T *s_key, const T *__restrict__ g, int count /*scalars*/) {
  // flat copy of `count` scalars; vectorised when both sides are 16-byte aligned
  constexpr int V = 16 / sizeof(T);
  if ((reinterpret_cast<uintptr_t>(g) & 15u) == 0) {
    const int nv = count / V;
    const int4 *g4 = reinterpret_cast<const int4 *>(g);
    int4 *s4 = reinterpret_cast<int4 *>(s_key);
    for (int i = threadIdx.x; i < nv; i += blockDim.x) s4[i] = __ldg(g4 + i);
    for (int i = nv * V + threadIdx.x; i < count; i += blockDim.x) s_key[i] = g[i];
  } else {
    for (int i = threadIdx.x; i < count; i += blockDim.x) s_key[i] = g[i];
  }
}

}  // namespace mvp
