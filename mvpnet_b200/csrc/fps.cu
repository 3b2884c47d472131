// Farthest point sampling for sm_100a.
//
// Semantics follow mvpnet/ops/cuda/fps_kernel.cu:60-135 (first centroid = index 0, running
// min-distance, strict comparisons) INCLUDING its tie rule, which is a by-product of the
// reference's launch geometry: thread t = j mod BLOCK keeps its first strict maximum and the
// shared-memory tree keeps the lower slot on equality, so among equal maxima the survivor is the
// point with the smallest (bit_reverse(j mod BLOCK), j / BLOCK), BLOCK = min(2^floor(log2 N), 512)
// (>= 16).  We do not copy that geometry; we order candidates by the same key.
//
// B200 design: everything an iteration touches lives on-chip.  One CTA per cloud; each thread keeps PPT points
// (x, y, z, running min, tie rank) in REGISTERS for the whole run, xyz is mirrored in shared memory only for the
// broadcast read of the new centroid.  The arg-max is two `redux.sync` per level (max of the distance bits, then
// min of the tie rank among the maxima) and ONE block barrier per iteration (double-buffered partials).  Warps own
// spatially sorted points and skip iterations that provably cannot change them (see fps_regs_kernel).
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
// register-resident, spatially bucketed kernel: fp32, D == 3, N <= blockDim.x * PPT <= 8192
//
// The profile of the plain register-resident scan (profiles/r2_fps_lines.txt) shows the iteration is bound by the
// FP32 pipe of the ONE SM a cloud runs on (8192 x 6 operations = 768 cycles), not by latency.  Most of that work
// is provably useless: a new centroid can only lower the running minimum of points that are closer to it than
// their current minimum.  So the points are counting-sorted by Morton cell once (512 cells, shared-memory
// atomics), each warp owns 32 * PPT spatially adjacent points together with their bounding box and its current
// maximum, and a warp whose box is farther from the new centroid than that maximum (conservatively: computed
// box distance * (1 - 1e-5) >= max, the fp32 error of either side is < 1e-6) skips the iteration — its candidate
// is unchanged.  The skip is lossless, so the selected indices are bit-identical to the full scan (and to the
// reference).  The intra-cell order produced by the atomics does not matter: every point carries its tie rank.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned morton3(unsigned x, unsigned y, unsigned z) {   // 3 bits per axis -> 9 bits
  unsigned v = 0;
#pragma unroll
  for (int b = 0; b < 3; ++b) v |= (((x >> b) & 1u) << (3 * b)) | (((y >> b) & 1u) << (3 * b + 1)) | (((z >> b) & 1u) << (3 * b + 2));
  return v;
}

// BUCKET = false (small clouds, where the per-iteration latency chain and not the FP32 pipe is the limit — measured:
// N = 2048 runs 0.29 us / iteration plain, 0.40 us bucketed): identity order, thread t owns j = t + p * T with T a
// multiple of the reference BLOCK, so the first strict maximum in p order IS the reference's tie order and no
// per-point rank is kept.
template <int PPT, bool BUCKET>
__global__ void __launch_bounds__(1024, 1)
fps_regs_kernel(const float *__restrict__ points, int64_t *__restrict__ index, int N, int M, int lg) {
  extern __shared__ float s_xyz[];  // [N*3] AoS mirror (ORIGINAL order) for the centroid broadcast, then u16 perm[N]
  __shared__ unsigned s_part_d[2][32];
  __shared__ unsigned s_part_r[2][32];
  __shared__ float s_red[6][32];
  __shared__ int s_cell[512];

  const int T = blockDim.x;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
  const float *pts = points + (size_t)blockIdx.x * N * 3;
  int64_t *out = index + (size_t)blockIdx.x * M;
  unsigned short *s_perm = reinterpret_cast<unsigned short *>(s_xyz + 3 * (size_t)N);

  // ---- stage the cloud, bounding box of the cloud
  const float inf = Inf<float>::v();
  float lo[3] = {inf, inf, inf}, hi[3] = {-inf, -inf, -inf};
  for (int j = tid; j < N; j += T) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const float v = pts[3 * j + d];
      s_xyz[3 * j + d] = v;
      lo[d] = fminf(lo[d], v);
      hi[d] = fmaxf(hi[d], v);
    }
  }
  if (BUCKET) {
  for (int c = tid; c < 512; c += T) s_cell[c] = 0;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[d] = fminf(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
      hi[d] = fmaxf(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
    }
    if (lane == 0) { s_red[d][warp] = lo[d]; s_red[3 + d][warp] = hi[d]; }
  }
  __syncthreads();
  float scale[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    float l = lane < nwarps ? s_red[d][lane] : inf, h = lane < nwarps ? s_red[3 + d][lane] : -inf;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      l = fminf(l, __shfl_xor_sync(0xffffffffu, l, o));
      h = fmaxf(h, __shfl_xor_sync(0xffffffffu, h, o));
    }
    lo[d] = l;
    const float ext = h - l;
    scale[d] = ext > 0.f && ext < inf ? 8.f / ext : 0.f;     // degenerate / non-finite extent: one cell along this axis
  }
  // ---- counting sort by Morton cell (order inside a cell is arbitrary)
  auto cell_of = [&](int j) -> unsigned {
    unsigned q[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const float f = (s_xyz[3 * j + d] - lo[d]) * scale[d];
      q[d] = f >= 7.f ? 7u : (f > 0.f ? (unsigned)f : 0u);    // NaN -> 0
    }
    return morton3(q[0], q[1], q[2]);
  };
  for (int j = tid; j < N; j += T) atomicAdd(&s_cell[cell_of(j)], 1);
  __syncthreads();
  if (warp == 0) {            // exclusive scan of 512 counters: 16 per lane
    int v[16], sum = 0;
#pragma unroll
    for (int q = 0; q < 16; ++q) { v[q] = s_cell[lane * 16 + q]; sum += v[q]; }
    int pre = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, pre, o);
      if (lane >= o) pre += u;
    }
    pre -= sum;
#pragma unroll
    for (int q = 0; q < 16; ++q) { s_cell[lane * 16 + q] = pre; pre += v[q]; }
  }
  __syncthreads();
  for (int j = tid; j < N; j += T) s_perm[atomicAdd(&s_cell[cell_of(j)], 1)] = (unsigned short)j;
  }
  __syncthreads();

  // ---- this thread's points (sorted position = warp * 32 * PPT + p * 32 + lane), tie ranks, the warp's box
  float px[PPT], py[PPT], pz[PPT], md[PPT];
  unsigned rk[PPT];
  float blo[3] = {inf, inf, inf}, bhi[3] = {-inf, -inf, -inf};
#pragma unroll
  for (int p = 0; p < PPT; ++p) {
    const int s = BUCKET ? warp * 32 * PPT + p * 32 + lane : tid + p * T;
    if (s < N) {
      const unsigned j = BUCKET ? (unsigned)s_perm[s] : (unsigned)s;
      px[p] = s_xyz[3 * j], py[p] = s_xyz[3 * j + 1], pz[p] = s_xyz[3 * j + 2];
      md[p] = inf;
      rk[p] = tie_rank(j, lg);
      blo[0] = fminf(blo[0], px[p]); blo[1] = fminf(blo[1], py[p]); blo[2] = fminf(blo[2], pz[p]);
      bhi[0] = fmaxf(bhi[0], px[p]); bhi[1] = fmaxf(bhi[1], py[p]); bhi[2] = fmaxf(bhi[2], pz[p]);
    } else {  // phantom slot: distance pinned at 0 can never be a strict maximum
      px[p] = py[p] = pz[p] = 0.f;
      md[p] = 0.f;
      rk[p] = 0xffffffffu;
    }
  }
#pragma unroll
  for (int d = 0; d < 3; ++d) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      blo[d] = fminf(blo[d], __shfl_xor_sync(0xffffffffu, blo[d], o));
      bhi[d] = fmaxf(bhi[d], __shfl_xor_sync(0xffffffffu, bhi[d], o));
    }
  }
  if (tid == 0) out[0] = 0;

  // the warp's cached candidate: (max of the running minima, smallest tie rank among the maxima)
  float wmax = !BUCKET || warp * 32 * PPT < N ? inf : 0.f;
  unsigned wr = 0xffffffffu;
  unsigned cur = 0;
  for (int i = 1; i < M; ++i) {
    const float cx = s_xyz[3 * cur], cy = s_xyz[3 * cur + 1], cz = s_xyz[3 * cur + 2];
    // squared distance from the centroid to the warp's box (0 inside); warp-uniform
    const float ax = fmaxf(fmaxf(blo[0] - cx, cx - bhi[0]), 0.f), ay = fmaxf(fmaxf(blo[1] - cy, cy - bhi[1]), 0.f),
                az = fmaxf(fmaxf(blo[2] - cz, cz - bhi[2]), 0.f);
    const float box2 = __fmaf_rn(az, az, __fmaf_rn(ay, ay, __fmul_rn(ax, ax)));
    const bool skip = BUCKET && (wmax == 0.f || (box2 > 1e-30f && box2 * 0.99999f >= wmax));
    if (!skip) {
      float best = 0.f;
      int bp = 0;
#pragma unroll
      for (int p = 0; p < PPT; ++p) {
        const float d = sqdist3(px[p], py[p], pz[p], cx, cy, cz);
        md[p] = fminf(md[p], d);
        if (BUCKET) best = fmaxf(best, md[p]);
        else if (md[p] > best) { best = md[p]; bp = p; }
      }
      const unsigned m = __reduce_max_sync(0xffffffffu, __float_as_uint(best));   // non-negative floats order as integers
      wmax = __uint_as_float(m);
      unsigned r = 0xffffffffu;
      if (BUCKET) {
#pragma unroll
        for (int p = 0; p < PPT; ++p)
          if (md[p] == wmax) r = min(r, rk[p]);
      } else if (best == wmax) {
        r = tie_rank((unsigned)(tid + bp * T), lg);
      }
      wr = __reduce_min_sync(0xffffffffu, r);
    }
    const int buf = i & 1;
    if (lane == 0) { s_part_d[buf][warp] = __float_as_uint(wmax); s_part_r[buf][warp] = wr; }
    __syncthreads();
    // block level, redundantly in every warp (no second barrier)
    const unsigned db = lane < nwarps ? s_part_d[buf][lane] : 0u;
    const unsigned rb = lane < nwarps ? s_part_r[buf][lane] : 0xffffffffu;
    const unsigned m = __reduce_max_sync(0xffffffffu, db);
    const unsigned r = __reduce_min_sync(0xffffffffu, db == m ? rb : 0xffffffffu);
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
    // bucketed (N >= 4096): every point carries its tie rank, the launch geometry is free: 1024 threads.
    // plain: T must be a multiple of the reference BLOCK (= 1 << lg, <= 512) for the p-order tie rule (see the kernel).
    const bool bucket = N >= 4096;
    int threads = bucket ? 1024 : ((1 << lg) < 32 ? 32 : (1 << lg));
    while (threads < 1024 && (N + threads - 1) / threads > 8) threads *= 2;
    int ppt = (int)((N + threads - 1) / threads);
    ppt = ppt <= 1 ? 1 : ppt <= 2 ? 2 : ppt <= 4 ? 4 : 8;
    const size_t smem = (size_t)N * 3 * sizeof(float) + (size_t)N * sizeof(unsigned short);
    const float *p = (const float *)points;
#define MVP_FPS_LAUNCH(P, BK)                                                                          \
  do {                                                                                                 \
    cudaFuncSetAttribute(fps_regs_kernel<P, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    fps_regs_kernel<P, BK><<<(unsigned)B, threads, smem, stream>>>(p, index, (int)N, (int)M, lg);     \
  } while (0)
    if (bucket) {
      if (ppt <= 4) MVP_FPS_LAUNCH(4, true); else MVP_FPS_LAUNCH(8, true);
    } else {
      switch (ppt) {
        case 1: MVP_FPS_LAUNCH(1, false); break;
        case 2: MVP_FPS_LAUNCH(2, false); break;
        case 4: MVP_FPS_LAUNCH(4, false); break;
        default: MVP_FPS_LAUNCH(8, false); break;
      }
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
