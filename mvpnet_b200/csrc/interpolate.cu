// feature_interpolate forward / backward (k == 3) for sm_100a.
//
// Semantics: mvpnet/ops/cuda/interpolate_kernel.cu:25-68 (out[b,c,n] = sum_k in[b,c,idx[b,n,k]] *
// w[b,n,k], accumulated k = 0,1,2 — with nvcc's contraction that is fma(in2,w2, fma(in1,w1, in0*w0)))
// and :131-174 (backward: atomic scatter of grad_out * w into the key features).
//
// B200 design (HBM-bound gather): one thread per query point reads its three (index, weight) pairs
// ONCE and then walks a group of channels; the reference re-reads 3 x (8 + 4) bytes per output
// element.  Consecutive threads are consecutive n, so stores are coalesced; the gathered rows
// (M floats per channel) stay in L1/L2.
#include "common.cuh"

namespace mvp {

constexpr int IP_THREADS = 256;
constexpr int IP_CH_PER_BLOCK = 32;

__device__ __forceinline__ float fma3(float a0, float w0, float a1, float w1, float a2, float w2) {
  return __fmaf_rn(a2, w2, __fmaf_rn(a1, w1, __fmul_rn(a0, w0)));
}
__device__ __forceinline__ double fma3(double a0, double w0, double a1, double w1, double a2, double w2) {
  return __fma_rn(a2, w2, __fma_rn(a1, w1, __dmul_rn(a0, w0)));
}

template <typename T>
__global__ void __launch_bounds__(IP_THREADS)
interpolate_fwd_kernel(const T *__restrict__ in, int64_t sb, int64_t sc, int64_t sm,
                       const int64_t *__restrict__ index, const T *__restrict__ weight, int C, int M, int N,
                       T *__restrict__ out) {
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * IP_CH_PER_BLOCK, c1 = min(C, c0 + IP_CH_PER_BLOCK);
  const int n = blockIdx.x * IP_THREADS + threadIdx.x;
  if (n >= N) return;
  const int64_t *ip = index + ((int64_t)b * N + n) * 3;
  const T *wp = weight + ((int64_t)b * N + n) * 3;
  int64_t j[3];
  T w[3];
  unsigned bad = 0;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    j[k] = ip[k];
    w[k] = wp[k];
    if (j[k] < 0 || j[k] >= M) { ++bad; j[k] = 0; w[k] = (T)0; }
  }
  if (bad && blockIdx.y == 0) atomicAdd(&g_index_errors, (unsigned long long)bad);
  const T *src = in + (int64_t)b * sb;
  T *dst = out + (int64_t)b * C * N + n;
#pragma unroll 4
  for (int c = c0; c < c1; ++c) {
    const T *s = src + (int64_t)c * sc;
    dst[(int64_t)c * N] = fma3(__ldg(s + j[0] * sm), w[0], __ldg(s + j[1] * sm), w[1], __ldg(s + j[2] * sm), w[2]);
  }
}

template <typename T>
__global__ void __launch_bounds__(IP_THREADS)
interpolate_bwd_kernel(const T *__restrict__ gout, const int64_t *__restrict__ index, const T *__restrict__ weight,
                       int C, int M, int N, T *__restrict__ gin) {
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * IP_CH_PER_BLOCK, c1 = min(C, c0 + IP_CH_PER_BLOCK);
  const int n = blockIdx.x * IP_THREADS + threadIdx.x;
  if (n >= N) return;
  const int64_t *ip = index + ((int64_t)b * N + n) * 3;
  const T *wp = weight + ((int64_t)b * N + n) * 3;
  int64_t j[3];
  T w[3];
  bool ok[3];
  unsigned bad = 0;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    j[k] = ip[k];
    w[k] = wp[k];
    ok[k] = j[k] >= 0 && j[k] < M;
    bad += !ok[k];
  }
  if (bad && blockIdx.y == 0) atomicAdd(&g_index_errors, (unsigned long long)bad);
  for (int c = c0; c < c1; ++c) {
    const T g = gout[((int64_t)b * C + c) * N + n];
    T *d = gin + ((int64_t)b * C + c) * M;
#pragma unroll
    for (int k = 0; k < 3; ++k)
      if (ok[k]) atomicAdd(d + j[k], g * w[k]);
  }
}

}  // namespace mvp

extern "C" int mvp_interpolate_forward(const void *in, int64_t sb, int64_t sc, int64_t sm, const int64_t *index,
                                       const void *weight, int64_t B, int64_t C, int64_t M, int64_t N, int dtype,
                                       void *out, mvp_stream_t stream_) {
  using namespace mvp;
  cudaStream_t stream = (cudaStream_t)stream_;
  MVP_REQUIRE(dtype == MVP_F32 || dtype == MVP_F64, MVP_ERR_INVALID_ARG, "interpolate: bad dtype");
  MVP_REQUIRE(B >= 0 && C >= 0 && M >= 0 && N >= 0, MVP_ERR_INVALID_ARG, "interpolate: negative size");
  if (B == 0 || C == 0 || N == 0) return 0;
  MVP_REQUIRE(in && index && weight && out, MVP_ERR_NULL, "interpolate: null pointer");
  MVP_REQUIRE(B <= 65535 && M < (1LL << 31) && N < (1LL << 31), MVP_ERR_UNSUPPORTED, "interpolate: size too large");
  const int64_t gy = (C + IP_CH_PER_BLOCK - 1) / IP_CH_PER_BLOCK;
  MVP_REQUIRE(gy <= 65535, MVP_ERR_UNSUPPORTED, "interpolate: too many channels");
  dim3 grid((unsigned)((N + IP_THREADS - 1) / IP_THREADS), (unsigned)gy, (unsigned)B);
  if (dtype == MVP_F32)
    interpolate_fwd_kernel<float><<<grid, IP_THREADS, 0, stream>>>((const float *)in, sb, sc, sm, index, (const float *)weight, (int)C, (int)M, (int)N, (float *)out);
  else
    interpolate_fwd_kernel<double><<<grid, IP_THREADS, 0, stream>>>((const double *)in, sb, sc, sm, index, (const double *)weight, (int)C, (int)M, (int)N, (double *)out);
  return launch_status("interpolate_forward");
}

extern "C" int mvp_interpolate_backward(const void *grad_out, const int64_t *index, const void *weight, int64_t B,
                                        int64_t C, int64_t M, int64_t N, int dtype, void *grad_in,
                                        mvp_stream_t stream_) {
  using namespace mvp;
  cudaStream_t stream = (cudaStream_t)stream_;
  MVP_REQUIRE(dtype == MVP_F32 || dtype == MVP_F64, MVP_ERR_INVALID_ARG, "interpolate: bad dtype");
  MVP_REQUIRE(B >= 0 && C >= 0 && M >= 0 && N >= 0, MVP_ERR_INVALID_ARG, "interpolate: negative size");
  const size_t esz = dtype == MVP_F64 ? 8 : 4;
  if (B == 0 || C == 0 || M == 0) return 0;
  MVP_REQUIRE(grad_in, MVP_ERR_NULL, "interpolate: null pointer");
  cudaError_t me = cudaMemsetAsync(grad_in, 0, (size_t)B * C * M * esz, stream);
  if (me != cudaSuccess) { set_error("interpolate_backward: memset failed: %s", cudaGetErrorString(me)); return (int)me; }
  if (N == 0) return 0;
  MVP_REQUIRE(grad_out && index && weight, MVP_ERR_NULL, "interpolate: null pointer");
  MVP_REQUIRE(B <= 65535 && M < (1LL << 31) && N < (1LL << 31), MVP_ERR_UNSUPPORTED, "interpolate: size too large");
  const int64_t gy = (C + IP_CH_PER_BLOCK - 1) / IP_CH_PER_BLOCK;
  MVP_REQUIRE(gy <= 65535, MVP_ERR_UNSUPPORTED, "interpolate: too many channels");
  dim3 grid((unsigned)((N + IP_THREADS - 1) / IP_THREADS), (unsigned)gy, (unsigned)B);
  if (dtype == MVP_F32)
    interpolate_bwd_kernel<float><<<grid, IP_THREADS, 0, stream>>>((const float *)grad_out, index, (const float *)weight, (int)C, (int)M, (int)N, (float *)grad_in);
  else
    interpolate_bwd_kernel<double><<<grid, IP_THREADS, 0, stream>>>((const double *)grad_out, index, (const double *)weight, (int)C, (int)M, (int)N, (double *)grad_in);
  return launch_status("interpolate_backward");
}
