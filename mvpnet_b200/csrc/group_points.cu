// group_points forward / backward for sm_100a.
//
// Semantics: mvpnet/ops/cuda/group_points_kernel.cu:25-47 (forward = gather along the point axis,
// out[b,c,n,k] = in[b,c,index[b,n,k]]) and :50-145 (backward = scatter-add).
//
// B200 design (HBM-bound): a thread owns 4 consecutive (n,k) output slots; it reads their int64
// indices ONCE (two 128-bit loads) and then walks a group of channels, so every output element is
// written exactly once with 128-bit coalesced stores and the index tensor is read once per channel
// GROUP instead of once per channel.  The gathered rows (N1 floats per channel) are L1/L2 resident.
// Out-of-range indices (the -1 rows of ball_query) yield 0 / are skipped and are counted, instead
// of the reference's device-side assert.
#include "common.cuh"

namespace mvp {

constexpr int GP_THREADS = 256;
constexpr int GP_CH_PER_BLOCK = 16;

template <typename T>
__global__ void __launch_bounds__(GP_THREADS)
group_points_fwd_kernel(const T *__restrict__ in, int64_t sb, int64_t sc, int64_t sn,
                        const int64_t *__restrict__ index, int C, int N1, int64_t E /*N2*K*/,
                        T *__restrict__ out) {
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * GP_CH_PER_BLOCK;
  const int c1 = min(C, c0 + GP_CH_PER_BLOCK);
  const int64_t e0 = ((int64_t)blockIdx.x * GP_THREADS + threadIdx.x) * 4;
  if (e0 >= E) return;
  const int64_t *idx = index + (int64_t)b * E + e0;
  const int nvalid = (int)min((int64_t)4, E - e0);
  int64_t j[4];
  bool ok[4];
  unsigned bad = 0;
  if (nvalid == 4 && ((reinterpret_cast<uintptr_t>(idx) & 15u) == 0)) {
    const longlong2 a = __ldg(reinterpret_cast<const longlong2 *>(idx));
    const longlong2 c = __ldg(reinterpret_cast<const longlong2 *>(idx) + 1);
    j[0] = a.x, j[1] = a.y, j[2] = c.x, j[3] = c.y;
  } else {
#pragma unroll
    for (int u = 0; u < 4; ++u) j[u] = u < nvalid ? idx[u] : 0;
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    ok[u] = j[u] >= 0 && j[u] < N1;
    if (!ok[u]) { ++bad; j[u] = 0; }
  }
  if (bad && blockIdx.y == 0) atomicAdd(&g_index_errors, (unsigned long long)bad);
  const T *src = in + (int64_t)b * sb;
  T *dst = out + ((int64_t)b * C) * E + e0;
  const bool vec = nvalid == 4 && sizeof(T) == 4 && ((E & 3) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15u) == 0);
  for (int c = c0; c < c1; ++c) {
    const T *s = src + (int64_t)c * sc;
    T v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = ok[u] ? __ldg(s + j[u] * sn) : (T)0;
    T *d = dst + (int64_t)c * E;
    if (vec) {
      float4 f = make_float4((float)v[0], (float)v[1], (float)v[2], (float)v[3]);
      __stcs(reinterpret_cast<float4 *>(d), f);  // streaming store: the grouped tensor is write-once
    } else {
#pragma unroll
      for (int u = 0; u < 4; ++u) if (u < nvalid) d[u] = v[u];
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(GP_THREADS)
group_points_bwd_kernel(const T *__restrict__ gout, const int64_t *__restrict__ index, int C, int N1,
                        int64_t E, T *__restrict__ gin) {
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * GP_CH_PER_BLOCK;
  const int c1 = min(C, c0 + GP_CH_PER_BLOCK);
  const int64_t e = (int64_t)blockIdx.x * GP_THREADS + threadIdx.x;
  if (e >= E) return;
  const int64_t j = index[(int64_t)b * E + e];
  if (j < 0 || j >= N1) {
    if (blockIdx.y == 0) atomicAdd(&g_index_errors, 1ULL);
    return;
  }
  for (int c = c0; c < c1; ++c)
    atomicAdd(gin + ((int64_t)b * C + c) * N1 + j, gout[((int64_t)b * C + c) * E + e]);
}

}  // namespace mvp

extern "C" int mvp_group_points_forward(const void *in, int64_t sb, int64_t sc, int64_t sn, const int64_t *index,
                                        int64_t B, int64_t C, int64_t N1, int64_t N2, int64_t K, int dtype,
                                        void *out, mvp_stream_t stream_) {
  using namespace mvp;
  cudaStream_t stream = (cudaStream_t)stream_;
  MVP_REQUIRE(dtype == MVP_F32 || dtype == MVP_F64, MVP_ERR_INVALID_ARG, "group_points: bad dtype");
  MVP_REQUIRE(B >= 0 && C >= 0 && N1 >= 0 && N2 >= 0 && K >= 0, MVP_ERR_INVALID_ARG, "group_points: negative size");
  const int64_t E = N2 * K;
  if (B == 0 || C == 0 || E == 0) return 0;
  MVP_REQUIRE(in && index && out, MVP_ERR_NULL, "group_points: null pointer");
  MVP_REQUIRE(B <= 65535 && N1 < (1LL << 31), MVP_ERR_UNSUPPORTED, "group_points: batch > 65535 or N1 >= 2^31");
  const int64_t gx = (E + GP_THREADS * 4 - 1) / (GP_THREADS * 4);
  const int64_t gy = (C + GP_CH_PER_BLOCK - 1) / GP_CH_PER_BLOCK;
  MVP_REQUIRE(gx < (1LL << 31) && gy <= 65535, MVP_ERR_UNSUPPORTED, "group_points: tensor too large");
  dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)B);
  if (dtype == MVP_F32)
    group_points_fwd_kernel<float><<<grid, GP_THREADS, 0, stream>>>((const float *)in, sb, sc, sn, index, (int)C, (int)N1, E, (float *)out);
  else
    group_points_fwd_kernel<double><<<grid, GP_THREADS, 0, stream>>>((const double *)in, sb, sc, sn, index, (int)C, (int)N1, E, (double *)out);
  return launch_status("group_points_forward");
}

extern "C" int mvp_group_points_backward(const void *grad_out, const int64_t *index, int64_t B, int64_t C, int64_t N1,
                                         int64_t N2, int64_t K, int dtype, void *grad_in, mvp_stream_t stream_) {
  using namespace mvp;
  cudaStream_t stream = (cudaStream_t)stream_;
  MVP_REQUIRE(dtype == MVP_F32 || dtype == MVP_F64, MVP_ERR_INVALID_ARG, "group_points: bad dtype");
  MVP_REQUIRE(B >= 0 && C >= 0 && N1 >= 0 && N2 >= 0 && K >= 0, MVP_ERR_INVALID_ARG, "group_points: negative size");
  const int64_t E = N2 * K;
  const size_t esz = dtype == MVP_F64 ? 8 : 4;
  if (B == 0 || C == 0 || N1 == 0) return 0;
  MVP_REQUIRE(grad_in, MVP_ERR_NULL, "group_points: null pointer");
  cudaError_t me = cudaMemsetAsync(grad_in, 0, (size_t)B * C * N1 * esz, stream);
  if (me != cudaSuccess) { set_error("group_points_backward: memset failed: %s", cudaGetErrorString(me)); return (int)me; }
  if (E == 0) return 0;
  MVP_REQUIRE(grad_out && index, MVP_ERR_NULL, "group_points: null pointer");
  MVP_REQUIRE(B <= 65535, MVP_ERR_UNSUPPORTED, "group_points: batch > 65535");
  const int64_t gx = (E + GP_THREADS - 1) / GP_THREADS;
  const int64_t gy = (C + GP_CH_PER_BLOCK - 1) / GP_CH_PER_BLOCK;
  MVP_REQUIRE(gx < (1LL << 31) && gy <= 65535, MVP_ERR_UNSUPPORTED, "group_points: tensor too large");
  dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)B);
  if (dtype == MVP_F32)
    group_points_bwd_kernel<float><<<grid, GP_THREADS, 0, stream>>>((const float *)grad_out, index, (int)C, (int)N1, E, (float *)grad_in);
  else
    group_points_bwd_kernel<double><<<grid, GP_THREADS, 0, stream>>>((const double *)grad_out, index, (int)C, (int)N1, E, (double *)grad_in);
  return launch_status("group_points_backward");
}
