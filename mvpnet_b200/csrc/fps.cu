// Farthest point sampling for sm_100a.
//
// Semantics follow mvpnet/ops/cuda/fps_kernel.cu:60-135 (first centroid = index 0, running
// min-distance, strict comparisons) INCLUDING its tie rule, which is a by-product of the
// reference's launch geometry: thread t = j mod BLOCK keeps its first strict maximum and the
// shared-memory tree keeps the lower slot on equality, so among equal maxima the survivor is the
// point with the smallest (bit_reverse(j mod BLOCK), j / BLOCK), BLOCK = min(2^floor(log2 N), 512)
// (>= 16).  We do not copy that geometry; we order candidates by the same key.
//
// B200 design: the serial chain of M-1 arg-max steps is latency-bound, so everything an iteration
// touches lives on-chip.  One CTA per cloud; each thread keeps PPT points (x, y, z, running min)
// in REGISTERS for the whole run, xyz is mirrored in shared memory only for the broadcast read of
// the new centroid.  The arg-max is two `redux.sync` per level (max of the distance bits, then min
// of the tie rank among the maxima) and ONE block barrier per iteration (double-buffered partials).
// Clouds too large for the register file (or fp64 / 2-D inputs) take the generic kernel that
// streams points and the running minimum through L1/L2.
#include "common.cuh"

namespace mvp {

__host__ __device__ inline int ref_block_log2(long long n) {
  // == log2 of the reference block size: clamp(2^floor(log2 n), 16, 512)
  int l = 0;
  while ((2LL << l) <= n && l < 9) ++l;
  return l < 4 ? 4 : l;
}

// tie rank: smaller wins.  (bit-reversed thread id of the reference) << 22 | (j / BLOCK)
__device__ __forceinline__ unsigned tie_rank(unsigned j, int lg) {
  const unsigned t = j & ((1u << lg) - 1u);
  return ((__brev(t) >> (32 - lg)) << 22) | (j >> lg);
}
__device__ __forceinline__ unsigned rank_to_index(unsigned r, int lg) {
  const unsigned tb = r >> 22, q = r & 0x3fffffu;
  return (q << lg) | (__brev(tb) >> (32 - lg));
}

// ------------------------------------------------------------------------------------------------
// register-resident kernel: fp32, D == 3, N <= blockDim.x * PPT
// ------------------------------------------------------------------------------------------------
template <int PPT>
__global__ void __launch_bounds__(1024, 1)
fps_regs_kernel(const float *__restrict__ points, int64_t *__restrict__ index, int N, int M, int lg) {
  extern __shared__ float s_xyz[];  // [N*3] AoS mirror for the centroid broadcast
  __shared__ unsigned s_part_d[2][32];
  __shared__ unsigned s_part_r[2][32];

  const int T = blockDim.x;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
  const float *pts = points + (size_t)blockIdx.x * N * 3;
  int64_t *out = index + (size_t)blockIdx.x * M;

  float px[PPT], py[PPT], pz[PPT], md[PPT];
#pragma unroll
  for (int p = 0; p < PPT; ++p) {
    const int j = tid + p * T;
    if (j < N) {
      px[p] = pts[3 * j], py[p] = pts[3 * j + 1], pz[p] = pts[3 * j + 2];
      s_xyz[3 * j] = px[p], s_xyz[3 * j + 1] = py[p], s_xyz[3 * j + 2] = pz[p];
      md[p] = Inf<float>::v();
    } else {  // phantom slot: distance pinned at 0 can never be a strict maximum
      px[p] = py[p] = pz[p] = 0.f;
      md[p] = 0.f;
    }
  }
  if (tid == 0) out[0] = 0;
  __syncthreads();

  unsigned cur = 0;
  for (int i = 1; i < M; ++i) {
    const float cx = s_xyz[3 * cur], cy = s_xyz[3 * cur + 1], cz = s_xyz[3 * cur + 2];
    float best = 0.f;
    int bp = 0;
#pragma unroll
    for (int p = 0; p < PPT; ++p) {
      const float d = sqdist3(px[p], py[p], pz[p], cx, cy, cz);
      md[p] = fminf(md[p], d);
      if (md[p] > best) { best = md[p]; bp = p; }
    }
    unsigned db = __float_as_uint(best);
    unsigned rk = tie_rank((unsigned)(tid + bp * T), lg);
    // warp level
    unsigned m = __reduce_max_sync(0xffffffffu, db);
    unsigned r = __reduce_min_sync(0xffffffffu, db == m ? rk : 0xffffffffu);
    const int buf = i & 1;
    if (lane == 0) { s_part_d[buf][warp] = m; s_part_r[buf][warp] = r; }
    __syncthreads();
    // block level, redundantly in every warp (no second barrier)
    db = lane < nwarps ? s_part_d[buf][lane] : 0u;
    rk = lane < nwarps ? s_part_r[buf][lane] : 0xffffffffu;
    m = __reduce_max_sync(0xffffffffu, db);
    r = __reduce_min_sync(0xffffffffu, db == m ? rk : 0xffffffffu);
    if (m != 0u) cur = rank_to_index(r, lg);  // all-zero distances: the reference keeps cur_idx
    if (tid == 0) out[i] = (int64_t)cur;
  }
}

// ------------------------------------------------------------------------------------------------
// generic kernel: any N, D in {2,3}, fp32/fp64; running minimum in global workspace `temp` [B,N]
// ------------------------------------------------------------------------------------------------
template <typename T, int D>
__global__ void __launch_bounds__(1024, 1)
fps_generic_kernel(const T *__restrict__ points, int64_t *__restrict__ index, T *__restrict__ temp,
                   long long N, long long M, int lg) {
  __shared__ T s_d[2][32];
  __shared__ unsigned s_r[2][32];
  const int nthr = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5;
  const T *pts = points + (size_t)blockIdx.x * N * D;
  T *tmp = temp + (size_t)blockIdx.x * N;
  int64_t *out = index + (size_t)blockIdx.x * M;
  for (long long j = tid; j < N; j += nthr) tmp[j] = Inf<T>::v();
  if (tid == 0) out[0] = 0;
  unsigned cur = 0;
  for (long long i = 1; i < M; ++i) {
    T c[3] = {0, 0, 0};
#pragma unroll
    for (int d = 0; d < D; ++d) c[d] = pts[(size_t)cur * D + d];
    T best = 0;
    unsigned brk = 0xffffffffu;
    for (long long j = tid; j < N; j += nthr) {
      T a[3] = {0, 0, 0};
#pragma unroll
      for (int d = 0; d < D; ++d) a[d] = pts[(size_t)j * D + d];
      T dist;
      if (D == 3) {
        dist = sqdist3(a[0], a[1], a[2], c[0], c[1], c[2]);
      } else {
        dist = sqdist2(a[0], a[1], c[0], c[1]);
      }
      const T last = tmp[j];
      if (dist < last) tmp[j] = dist; else dist = last;
      const unsigned rk = tie_rank((unsigned)j, lg);
      if (dist > best || (dist == best && dist > (T)0 && rk < brk)) { best = dist; brk = rk; }
    }
    // warp arg-max on (dist desc, rank asc)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const T od = __shfl_xor_sync(0xffffffffu, best, o);
      const unsigned orr = __shfl_xor_sync(0xffffffffu, brk, o);
      if (od > best || (od == best && orr < brk)) { best = od; brk = orr; }
    }
    const int buf = (int)(i & 1);
    if (lane == 0) { s_d[buf][warp] = best; s_r[buf][warp] = brk; }
    __syncthreads();
    best = lane < nwarps ? s_d[buf][lane] : (T)0;
    brk = lane < nwarps ? s_r[buf][lane] : 0xffffffffu;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const T od = __shfl_xor_sync(0xffffffffu, best, o);
      const unsigned orr = __shfl_xor_sync(0xffffffffu, brk, o);
      if (od > best || (od == best && orr < brk)) { best = od; brk = orr; }
    }
    if (best > (T)0) cur = rank_to_index(brk, lg);
    if (tid == 0) out[i] = (int64_t)cur;
  }
}

static bool fits_regs(int64_t N, int64_t D, int dtype) {
  return dtype == MVP_F32 && D == 3 && N <= 8192;
}

}  // namespace mvp

extern "C" int64_t mvp_fps_workspace_bytes(int64_t B, int64_t N, int64_t D, int64_t M, int dtype) {
  (void)M;
  if (B <= 0 || N <= 0) return 0;
  if (mvp::fits_regs(N, D, dtype)) return 0;
  return B * N * (dtype == MVP_F64 ? 8 : 4);
}

extern "C" int mvp_fps(const void *points, int64_t B, int64_t N, int64_t D, int64_t M, int dtype,
                       int64_t *index, void *workspace, mvp_stream_t stream_) {
  using namespace mvp;
  cudaStream_t stream = (cudaStream_t)stream_;
  MVP_REQUIRE(dtype == MVP_F32 || dtype == MVP_F64, MVP_ERR_INVALID_ARG, "fps: dtype must be MVP_F32 or MVP_F64");
  MVP_REQUIRE(D == 2 || D == 3, MVP_ERR_INVALID_ARG, "Only support dim=2 or dim=3");
  MVP_REQUIRE(M > 0, MVP_ERR_INVALID_ARG, "fps: num_centroids (%lld) must be > 0", (long long)M);
  MVP_REQUIRE(N >= M, MVP_ERR_INVALID_ARG, "fps: num_points (%lld) must be >= num_centroids (%lld)", (long long)N, (long long)M);
  MVP_REQUIRE(N < (1LL << 31), MVP_ERR_UNSUPPORTED, "fps: num_points must be < 2^31");
  if (B == 0) return 0;
  MVP_REQUIRE(points && index, MVP_ERR_NULL, "fps: null pointer");
  const int lg = ref_block_log2(N);

  if (fits_regs(N, D, dtype)) {
    // points per thread: keep >= 4 warps so barrier cost stays small, <= 1024 threads
    int ppt = 1;
    while (ppt < 8 && (N + ppt - 1) / ppt > 256) ppt *= 2;
    int threads = (int)((N + ppt - 1) / ppt);
    threads = (threads + 31) / 32 * 32;
    if (threads > 1024) threads = 1024;
    const size_t smem = (size_t)N * 3 * sizeof(float);
    const float *p = (const float *)points;
#define MVP_FPS_LAUNCH(P)                                                                          \
  do {                                                                                             \
    cudaFuncSetAttribute(fps_regs_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    fps_regs_kernel<P><<<(unsigned)B, threads, smem, stream>>>(p, index, (int)N, (int)M, lg);     \
  } while (0)
    switch (ppt) {
      case 1: MVP_FPS_LAUNCH(1); break;
      case 2: MVP_FPS_LAUNCH(2); break;
      case 4: MVP_FPS_LAUNCH(4); break;
      default: MVP_FPS_LAUNCH(8); break;
    }
#undef MVP_FPS_LAUNCH
    return launch_status("fps");
  }

  MVP_REQUIRE(workspace, MVP_ERR_NULL, "fps: workspace of mvp_fps_workspace_bytes() bytes required");
  const int threads = N >= 1024 ? 1024 : (int)((N + 31) / 32 * 32);
  if (dtype == MVP_F32) {
    if (D == 3) fps_generic_kernel<float, 3><<<(unsigned)B, threads, 0, stream>>>((const float *)points, index, (float *)workspace, N, M, lg);
    else fps_generic_kernel<float, 2><<<(unsigned)B, threads, 0, stream>>>((const float *)points, index, (float *)workspace, N, M, lg);
  } else {
    if (D == 3) fps_generic_kernel<double, 3><<<(unsigned)B, threads, 0, stream>>>((const double *)points, index, (double *)workspace, N, M, lg);
    else fps_generic_kernel<double, 2><<<(unsigned)B, threads, 0, stream>>>((const double *)points, index, (double *)workspace, N, M, lg);
  }
  return launch_status("fps");
}
