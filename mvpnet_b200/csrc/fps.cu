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
#include <stdlib.h>

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
template <int PPT, bool BUCKET, int MAXT = 1024>
__global__ void __launch_bounds__(MAXT, 1)
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
// slab kernel (default for fp32, D == 3, N <= 8192): shared-memory resident, pruned per 32-point slab
//
// Measured on B200 (tools/fps_prof.py + ncu source view): with warp-sized buckets (above) 85 % of the distance
// work disappears but the iteration only went from 0.70 to 0.64 us — what remains is the per-iteration "common
// path" every one of the 32 warps executes (centroid read, box test, barrier, block arg-max: ~45 instructions x 32
// warps) plus the latency chain.  So: (1) the cloud lives in SHARED memory (structure of arrays, Morton-sorted), not in
// registers, which makes the pruning granularity independent of the warp count; (2) only W <= 8 warps run, each owning
// every W-th slab of 32 consecutive sorted points, and LANE p of a warp holds the metadata (bounding box, current
// maximum, its tie rank and position) of the warp's p-th slab, so one warp instruction tests 32 slabs; (3) a warp
// then updates only its active slabs (~3 % of them per iteration on room clouds, interleaved over the warps so that
// spatially adjacent active slabs land in different warps) with two `redux.sync` each, and its cached candidate is
// recomputed only if a slab changed.  Skips are conservative (see above): results are bit-identical to the full scan.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned morton4(unsigned x, unsigned y, unsigned z) {   // 4 bits per axis -> 12 bits
  unsigned v = 0;
#pragma unroll
  for (int b = 0; b < 4; ++b) v |= (((x >> b) & 1u) << (3 * b)) | (((y >> b) & 1u) << (3 * b + 1)) | (((z >> b) & 1u) << (3 * b + 2));
  return v;
}

constexpr int FPS_CELLS = 4096;

__global__ void __launch_bounds__(256, 1)
fps_slab_kernel(const float *__restrict__ points, int64_t *__restrict__ index, int N, int M, int lg, int S /*slabs*/) {
  extern __shared__ float s_dyn[];
  const int Np = S * 32;
  float *sx = s_dyn, *sy = sx + Np, *sz = sy + Np, *smd = sz + Np;
  unsigned *srk = reinterpret_cast<unsigned *>(smd + Np);
  int *s_cell = reinterpret_cast<int *>(srk + Np);                 // [FPS_CELLS]
  __shared__ unsigned part_m[2][8], part_r[2][8], part_p[2][8];
  __shared__ float s_red[6][8];
  __shared__ unsigned s_start;

  const int T = blockDim.x, W = T >> 5;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float *pts = points + (size_t)blockIdx.x * N * 3;
  int64_t *out = index + (size_t)blockIdx.x * M;
  const float inf = Inf<float>::v();

  // ---- bounding box of the cloud
  float lo[3] = {inf, inf, inf}, hi[3] = {-inf, -inf, -inf};
  for (int j = tid; j < N; j += T) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const float v = __ldg(pts + 3 * j + d);
      lo[d] = fminf(lo[d], v);
      hi[d] = fmaxf(hi[d], v);
    }
  }
  for (int c = tid; c < FPS_CELLS; c += T) s_cell[c] = 0;
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
    float l = lane < W ? s_red[d][lane] : inf, h = lane < W ? s_red[3 + d][lane] : -inf;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      l = fminf(l, __shfl_xor_sync(0xffffffffu, l, o));
      h = fmaxf(h, __shfl_xor_sync(0xffffffffu, h, o));
    }
    lo[d] = l;
    const float ext = h - l;
    scale[d] = ext > 0.f && ext < inf ? 16.f / ext : 0.f;    // degenerate / non-finite extent: one cell along this axis
  }
  auto cell_of = [&](float x, float y, float z) -> unsigned {
    const float f[3] = {(x - lo[0]) * scale[0], (y - lo[1]) * scale[1], (z - lo[2]) * scale[2]};
    unsigned q[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) q[d] = f[d] >= 15.f ? 15u : (f[d] > 0.f ? (unsigned)f[d] : 0u);   // NaN -> 0
    return morton4(q[0], q[1], q[2]);
  };
  // ---- counting sort by Morton cell into the shared-memory arrays (order inside a cell is arbitrary: ranks decide ties)
  for (int j = tid; j < N; j += T) atomicAdd(&s_cell[cell_of(__ldg(pts + 3 * j), __ldg(pts + 3 * j + 1), __ldg(pts + 3 * j + 2))], 1);
  __syncthreads();
  if (warp == 0) {            // exclusive scan of 4096 counters: 128 per lane
    int sum = 0;
    for (int q = 0; q < FPS_CELLS / 32; ++q) sum += s_cell[lane * (FPS_CELLS / 32) + q];
    int pre = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, pre, o);
      if (lane >= o) pre += u;
    }
    pre -= sum;
    for (int q = 0; q < FPS_CELLS / 32; ++q) {
      const int v = s_cell[lane * (FPS_CELLS / 32) + q];
      s_cell[lane * (FPS_CELLS / 32) + q] = pre;
      pre += v;
    }
  }
  __syncthreads();
  for (int j = tid; j < N; j += T) {
    const float x = __ldg(pts + 3 * j), y = __ldg(pts + 3 * j + 1), z = __ldg(pts + 3 * j + 2);
    const int pos = atomicAdd(&s_cell[cell_of(x, y, z)], 1);
    sx[pos] = x; sy[pos] = y; sz[pos] = z;
    smd[pos] = inf;
    srk[pos] = tie_rank((unsigned)j, lg);
    if (j == 0) s_start = (unsigned)pos;
  }
  for (int pos = N + tid; pos < Np; pos += T) {   // padding of the last slab: distance pinned at 0 can never be a strict maximum
    sx[pos] = sy[pos] = sz[pos] = 0.f;
    smd[pos] = 0.f;
    srk[pos] = 0xffffffffu;
  }
  __syncthreads();

  // ---- slab metadata: lane q of warp w owns slab q * W + w
  float b0 = inf, b1 = inf, b2 = inf, t0 = -inf, t1 = -inf, t2 = -inf;
  float smax = 0.f;
  unsigned srank = 0xffffffffu, spos = 0;
  for (int q = 0; q < 32; ++q) {
    const int s = q * W + warp;
    if (s >= S) break;                                   // warp-uniform
    const int base = s * 32 + lane;
    const bool real = base < N;
    float l0 = real ? sx[base] : inf, l1 = real ? sy[base] : inf, l2 = real ? sz[base] : inf;
    float h0 = real ? sx[base] : -inf, h1 = real ? sy[base] : -inf, h2 = real ? sz[base] : -inf;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      l0 = fminf(l0, __shfl_xor_sync(0xffffffffu, l0, o)); l1 = fminf(l1, __shfl_xor_sync(0xffffffffu, l1, o));
      l2 = fminf(l2, __shfl_xor_sync(0xffffffffu, l2, o)); h0 = fmaxf(h0, __shfl_xor_sync(0xffffffffu, h0, o));
      h1 = fmaxf(h1, __shfl_xor_sync(0xffffffffu, h1, o)); h2 = fmaxf(h2, __shfl_xor_sync(0xffffffffu, h2, o));
    }
    if (lane == q) { b0 = l0; b1 = l1; b2 = l2; t0 = h0; t1 = h1; t2 = h2; smax = inf; spos = (unsigned)(s * 32); }
  }
  if (tid == 0) out[0] = 0;

  unsigned wm = 0u, wr = 0xffffffffu, wpos = 0u;          // the warp's cached candidate (bits of the maximum, rank, position)
  unsigned curpos = s_start;
  long long curidx = 0;
  for (int i = 1; i < M; ++i) {
    const float cx = sx[curpos], cy = sy[curpos], cz = sz[curpos];
    // squared distance from the centroid to this lane's slab box (0 inside)
    const float ax = fmaxf(fmaxf(b0 - cx, cx - t0), 0.f), ay = fmaxf(fmaxf(b1 - cy, cy - t1), 0.f), az = fmaxf(fmaxf(b2 - cz, cz - t2), 0.f);
    const float box2 = __fmaf_rn(az, az, __fmaf_rn(ay, ay, __fmul_rn(ax, ax)));
    const bool act = !(smax == 0.f || (box2 > 1e-30f && box2 * 0.99999f >= smax));
    unsigned mask = __ballot_sync(0xffffffffu, act);
    const bool changed = mask != 0u;
    while (mask) {                                        // warp-uniform: update the active slabs
      const int q = __ffs(mask) - 1;
      mask &= mask - 1;
      const int s = q * W + warp, base = s * 32 + lane;
      const float md = fminf(smd[base], sqdist3(sx[base], sy[base], sz[base], cx, cy, cz));
      smd[base] = md;                                     // padding stays 0
      const unsigned mb = __float_as_uint(md);
      const unsigned m = __reduce_max_sync(0xffffffffu, mb);            // non-negative floats order as integers
      const unsigned rk = mb == m ? srk[base] : 0xffffffffu;
      const unsigned r = __reduce_min_sync(0xffffffffu, rk);
      const unsigned b = __ballot_sync(0xffffffffu, mb == m && rk == r);
      if (lane == q) { smax = __uint_as_float(m); srank = r; spos = (unsigned)(s * 32 + __ffs(b) - 1); }
    }
    if (changed) {
      const unsigned sb = __float_as_uint(smax);
      wm = __reduce_max_sync(0xffffffffu, sb);
      wr = __reduce_min_sync(0xffffffffu, sb == wm ? srank : 0xffffffffu);
      const unsigned b = __ballot_sync(0xffffffffu, sb == wm && srank == wr);
      wpos = __shfl_sync(0xffffffffu, spos, __ffs(b) - 1);
    }
    const int buf = i & 1;
    if (lane == 0) { part_m[buf][warp] = wm; part_r[buf][warp] = wr; part_p[buf][warp] = wpos; }
    __syncthreads();
    // block level, redundantly in every warp (no second barrier)
    const unsigned pm = lane < W ? part_m[buf][lane] : 0u, pr = lane < W ? part_r[buf][lane] : 0xffffffffu, pp = lane < W ? part_p[buf][lane] : 0u;
    const unsigned m = __reduce_max_sync(0xffffffffu, pm);
    const unsigned r = __reduce_min_sync(0xffffffffu, pm == m ? pr : 0xffffffffu);
    const unsigned b = __ballot_sync(0xffffffffu, pm == m && pr == r);
    const unsigned pos = __shfl_sync(0xffffffffu, pp, __ffs(b) - 1);
    if (m != 0u) { curpos = pos; curidx = (long long)rank_to_index(r, lg); }   // all-zero distances: the reference keeps cur_idx
    if (tid == 0) out[i] = curidx;
  }
}

// ------------------------------------------------------------------------------------------------
// large clouds (fp32, D == 3, N > 8192: whole-scene PN2SSG, BASELINE config 5): the same slab pruning with the sorted
// cloud in a global-memory workspace (L2 resident: 20 B per point) and the slab metadata — box, maximum, its tie rank
// AND the coordinates of the slab's farthest point, so that the next centroid never waits on a global load — in shared
// memory.  One CTA of 32 warps per cloud; slab s belongs to warp s mod 32 (adjacent = simultaneously active slabs go to
// different warps); a slab is 32 * PPL consecutive sorted points, PPL chosen so that there are at most 4096 slabs.
// The round-1 path streamed all N points through one SM on every iteration (200 k points: 352 ms for 8192 samples).
// ------------------------------------------------------------------------------------------------
constexpr int FPS_BIG_MAX_SLABS = 4096;
constexpr int FPS_BIG_CELLS = 32768;    // 5 bits per axis: a whole scene needs finer cells than a chunk, or every slab of a cell shares one box

__device__ __forceinline__ unsigned morton5(unsigned x, unsigned y, unsigned z) {
  unsigned v = 0;
#pragma unroll
  for (int b = 0; b < 5; ++b) v |= (((x >> b) & 1u) << (3 * b)) | (((y >> b) & 1u) << (3 * b + 1)) | (((z >> b) & 1u) << (3 * b + 2));
  return v;
}
constexpr int FPS_BIG_META = 11;       // words per slab: box lo[3], hi[3], max, rank, winner xyz

__global__ void __launch_bounds__(1024, 1)
fps_big_kernel(const float *__restrict__ points, int64_t *__restrict__ index, float *__restrict__ workspace, int N, int M, int lg, int S, int PPL) {
  extern __shared__ float s_meta[];                       // [FPS_BIG_META][S]; the cell counters alias its start during the sort
  __shared__ unsigned part_m[2][32], part_r[2][32];
  __shared__ float part_x[2][32], part_y[2][32], part_z[2][32];
  __shared__ float s_red[6][32];
  __shared__ float s_c0[3];
  __shared__ int s_scan[32];
  const int SL = 32 * PPL;
  const long long Np = (long long)S * SL;
  float *gx = workspace + (size_t)blockIdx.x * 5 * Np, *gy = gx + Np, *gz = gy + Np, *gmd = gz + Np;
  unsigned *grk = reinterpret_cast<unsigned *>(gmd + Np);
  int *s_cell = reinterpret_cast<int *>(s_meta);
  const int SP = (S + 1023) / 1024 * 1024;                 // slots (slabs rounded up to whole test rounds)
  float *m_b0 = s_meta, *m_b1 = m_b0 + SP, *m_b2 = m_b1 + SP, *m_t0 = m_b2 + SP, *m_t1 = m_t0 + SP, *m_t2 = m_t1 + SP, *m_max = m_t2 + SP;
  unsigned *m_rank = reinterpret_cast<unsigned *>(m_max + SP);
  float *m_wx = reinterpret_cast<float *>(m_rank + SP), *m_wy = m_wx + SP, *m_wz = m_wy + SP;

  const int T = 1024, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float *pts = points + (size_t)blockIdx.x * N * 3;
  int64_t *out = index + (size_t)blockIdx.x * M;
  const float inf = Inf<float>::v();

  float lo[3] = {inf, inf, inf}, hi[3] = {-inf, -inf, -inf};
  for (int j = tid; j < N; j += T) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const float v = __ldg(pts + 3 * (size_t)j + d);
      lo[d] = fminf(lo[d], v);
      hi[d] = fmaxf(hi[d], v);
    }
  }
  for (int c = tid; c < FPS_BIG_CELLS; c += T) s_cell[c] = 0;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[d] = fminf(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
      hi[d] = fmaxf(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
    }
    if (lane == 0) { s_red[d][warp] = lo[d]; s_red[3 + d][warp] = hi[d]; }
  }
  if (tid < 3) s_c0[tid] = __ldg(pts + tid);              // the first centroid is point 0
  __syncthreads();
  float scale[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    float l = s_red[d][lane], h = s_red[3 + d][lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      l = fminf(l, __shfl_xor_sync(0xffffffffu, l, o));
      h = fmaxf(h, __shfl_xor_sync(0xffffffffu, h, o));
    }
    lo[d] = l;
    const float ext = h - l;
    scale[d] = ext > 0.f && ext < inf ? 32.f / ext : 0.f;
  }
  auto cell_of = [&](float x, float y, float z) -> unsigned {
    const float f[3] = {(x - lo[0]) * scale[0], (y - lo[1]) * scale[1], (z - lo[2]) * scale[2]};
    unsigned q[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) q[d] = f[d] >= 31.f ? 31u : (f[d] > 0.f ? (unsigned)f[d] : 0u);
    return morton5(q[0], q[1], q[2]);
  };
  for (int j = tid; j < N; j += T) atomicAdd(&s_cell[cell_of(__ldg(pts + 3 * (size_t)j), __ldg(pts + 3 * (size_t)j + 1), __ldg(pts + 3 * (size_t)j + 2))], 1);
  __syncthreads();
  {   // exclusive scan of the 32768 counters by the whole CTA: 32 per thread, warp scan, scan of the 32 warp totals
    constexpr int PER = FPS_BIG_CELLS / 1024;
    int sum = 0;
    for (int q = 0; q < PER; ++q) sum += s_cell[tid * PER + q];
    int inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += u;
    }
    if (lane == 31) s_scan[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      int w = s_scan[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += u;
      }
      s_scan[lane] = w;
    }
    __syncthreads();
    int pre = inc - sum + (warp > 0 ? s_scan[warp - 1] : 0);
    for (int q = 0; q < PER; ++q) {
      const int v = s_cell[tid * PER + q];
      s_cell[tid * PER + q] = pre;
      pre += v;
    }
  }
  __syncthreads();
  for (int j = tid; j < N; j += T) {
    const float x = __ldg(pts + 3 * (size_t)j), y = __ldg(pts + 3 * (size_t)j + 1), z = __ldg(pts + 3 * (size_t)j + 2);
    const int pos = atomicAdd(&s_cell[cell_of(x, y, z)], 1);
    gx[pos] = x; gy[pos] = y; gz[pos] = z;
    gmd[pos] = inf;
    grk[pos] = tie_rank((unsigned)j, lg);
  }
  for (long long pos = N + tid; pos < Np; pos += T) {
    gx[pos] = gy[pos] = gz[pos] = 0.f;
    gmd[pos] = 0.f;
    grk[pos] = 0xffffffffu;
  }
  __syncthreads();                                        // the sorted cloud is visible to the CTA; the counters are dead
  // ---- slab metadata (warp w owns slabs w, w + 32, ...).  Slab s = (k * 32 + l) * 32 + w is tested by lane l of warp w in
  //      round k; its metadata lives at SLOT (k * 32 + w) * 32 + l, so that the 32 lanes of a test read 32 consecutive
  //      words (indexing the arrays by the slab id put all lanes of a warp on ONE bank: 32-way conflicts on every read made
  //      the iteration 15.7 us at 200 k points).
  for (int s = warp; s < S; s += 32) {
    float l0 = inf, l1 = inf, l2 = inf, h0 = -inf, h1 = -inf, h2 = -inf;
    bool any = false;
    for (int pp = 0; pp < PPL; ++pp) {
      const long long base = (long long)s * SL + pp * 32 + lane;
      if (base < N) {
        const float x = gx[base], y = gy[base], z = gz[base];
        l0 = fminf(l0, x); l1 = fminf(l1, y); l2 = fminf(l2, z); h0 = fmaxf(h0, x); h1 = fmaxf(h1, y); h2 = fmaxf(h2, z);
        any = true;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      l0 = fminf(l0, __shfl_xor_sync(0xffffffffu, l0, o)); l1 = fminf(l1, __shfl_xor_sync(0xffffffffu, l1, o));
      l2 = fminf(l2, __shfl_xor_sync(0xffffffffu, l2, o)); h0 = fmaxf(h0, __shfl_xor_sync(0xffffffffu, h0, o));
      h1 = fmaxf(h1, __shfl_xor_sync(0xffffffffu, h1, o)); h2 = fmaxf(h2, __shfl_xor_sync(0xffffffffu, h2, o));
    }
    any = __any_sync(0xffffffffu, any);
    if (lane == 0) {
      const int slot = ((s >> 10) * 32 + warp) * 32 + ((s >> 5) & 31);
      m_b0[slot] = l0; m_b1[slot] = l1; m_b2[slot] = l2; m_t0[slot] = h0; m_t1[slot] = h1; m_t2[slot] = h2;
      m_max[slot] = any ? inf : 0.f; m_rank[slot] = 0xffffffffu; m_wx[slot] = m_wy[slot] = m_wz[slot] = 0.f;
    }
  }
  if (tid == 0) out[0] = 0;
  __syncthreads();

  float cx = s_c0[0], cy = s_c0[1], cz = s_c0[2];
  long long curidx = 0;
  const int K = (S + 1023) / 1024;                        // slabs tested per lane
  for (int i = 1; i < M; ++i) {
    // ---- test this lane's slabs; the warp updates the active ones of each test round
    for (int k = 0; k < K; ++k) {
      const int s = (k * 32 + lane) * 32 + warp, slot = (k * 32 + warp) * 32 + lane;
      bool act = false;
      if (s < S) {
        const float ax = fmaxf(fmaxf(m_b0[slot] - cx, cx - m_t0[slot]), 0.f), ay = fmaxf(fmaxf(m_b1[slot] - cy, cy - m_t1[slot]), 0.f),
                    az = fmaxf(fmaxf(m_b2[slot] - cz, cz - m_t2[slot]), 0.f);
        const float box2 = __fmaf_rn(az, az, __fmaf_rn(ay, ay, __fmul_rn(ax, ax))), smax = m_max[slot];
        act = !(smax == 0.f || (box2 > 1e-30f && box2 * 0.99999f >= smax));
      }
      unsigned mask = __ballot_sync(0xffffffffu, act);
      while (mask) {
        // two active slabs per trip: their (L2-latency) loads are in flight together; a lone slab is paired with itself and
        // the duplicate's results are dropped
        const int la = __ffs(mask) - 1;
        mask &= mask - 1;
        const bool two = mask != 0u;
        const int lb = two ? __ffs(mask) - 1 : la;
        mask &= mask - 1;
        const int sa = (k * 32 + la) * 32 + warp, slot_a = (k * 32 + warp) * 32 + la;
        const int sb = (k * 32 + lb) * 32 + warp, slot_b = (k * 32 + warp) * 32 + lb;
        float best_a = 0.f, ax_ = 0.f, ay_ = 0.f, az_ = 0.f, best_b = 0.f, bx_ = 0.f, by_ = 0.f, bz_ = 0.f;
        unsigned rk_a = 0xffffffffu, rk_b = 0xffffffffu;
        for (int pp = 0; pp < PPL; ++pp) {
          const long long ba = (long long)sa * SL + pp * 32 + lane, bb = (long long)sb * SL + pp * 32 + lane;
          const float xa = gx[ba], ya = gy[ba], za = gz[ba], ma = gmd[ba];
          const unsigned ra = grk[ba];
          const float xb = gx[bb], yb = gy[bb], zb = gz[bb], mb_ = gmd[bb];
          const unsigned rb = grk[bb];
          const float mda = fminf(ma, sqdist3(xa, ya, za, cx, cy, cz)), mdb = fminf(mb_, sqdist3(xb, yb, zb, cx, cy, cz));
          gmd[ba] = mda;
          if (two) gmd[bb] = mdb;
          if (mda > best_a || (mda == best_a && ra < rk_a)) { best_a = mda; rk_a = ra; ax_ = xa; ay_ = ya; az_ = za; }
          if (mdb > best_b || (mdb == best_b && rb < rk_b)) { best_b = mdb; rk_b = rb; bx_ = xb; by_ = yb; bz_ = zb; }
        }
        {
          const unsigned mb = __float_as_uint(best_a);
          const unsigned m = __reduce_max_sync(0xffffffffu, mb);
          const unsigned r = __reduce_min_sync(0xffffffffu, mb == m ? rk_a : 0xffffffffu);
          const unsigned wb = __ballot_sync(0xffffffffu, mb == m && rk_a == r);      // every lane votes (never under a condition)
          if (lane == __ffs(wb) - 1) { m_max[slot_a] = best_a; m_rank[slot_a] = r; m_wx[slot_a] = ax_; m_wy[slot_a] = ay_; m_wz[slot_a] = az_; }
        }
        if (two) {                                                                  // warp-uniform
          const unsigned mb = __float_as_uint(best_b);
          const unsigned m = __reduce_max_sync(0xffffffffu, mb);
          const unsigned r = __reduce_min_sync(0xffffffffu, mb == m ? rk_b : 0xffffffffu);
          const unsigned wb = __ballot_sync(0xffffffffu, mb == m && rk_b == r);
          if (lane == __ffs(wb) - 1) { m_max[slot_b] = best_b; m_rank[slot_b] = r; m_wx[slot_b] = bx_; m_wy[slot_b] = by_; m_wz[slot_b] = bz_; }
        }
      }
    }
    __syncwarp();
    // ---- this lane's best slab, then warp, then block
    unsigned lm = 0u, lr = 0xffffffffu;
    int ls = 0;
    for (int k = 0; k < K; ++k) {
      const int s = (k * 32 + lane) * 32 + warp, slot = (k * 32 + warp) * 32 + lane;
      if (s < S) {
        const unsigned sb = __float_as_uint(m_max[slot]), sr = m_rank[slot];
        if (sb > lm || (sb == lm && sr < lr)) { lm = sb; lr = sr; ls = slot; }
      }
    }
    const unsigned wm = __reduce_max_sync(0xffffffffu, lm);
    const unsigned wr = __reduce_min_sync(0xffffffffu, lm == wm ? lr : 0xffffffffu);
    const int buf = i & 1;
    {
      const unsigned b = __ballot_sync(0xffffffffu, lm == wm && lr == wr);
      if (lane == __ffs(b) - 1) { part_m[buf][warp] = wm; part_r[buf][warp] = wr; part_x[buf][warp] = m_wx[ls]; part_y[buf][warp] = m_wy[ls]; part_z[buf][warp] = m_wz[ls]; }
    }
    __syncthreads();
    const unsigned pm = part_m[buf][lane], pr = part_r[buf][lane];
    const unsigned m = __reduce_max_sync(0xffffffffu, pm);
    const unsigned r = __reduce_min_sync(0xffffffffu, pm == m ? pr : 0xffffffffu);
    const int wl = __ffs(__ballot_sync(0xffffffffu, pm == m && pr == r)) - 1;
    if (m != 0u) { cx = part_x[buf][wl]; cy = part_y[buf][wl]; cz = part_z[buf][wl]; curidx = (long long)rank_to_index(r, lg); }
    if (tid == 0) out[i] = curidx;
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
static bool fits_big(int64_t N, int64_t D, int dtype) {
  static const bool off = getenv("MVPNET_B200_FPS") != nullptr && getenv("MVPNET_B200_FPS")[0] == 'g';   // "generic": the streaming kernel
  return !off && dtype == MVP_F32 && D == 3 && N > 8192;
}
static void big_geometry(int64_t N, int *S, int *PPL) {
  int ppl = 1;
  while ((N + 32LL * ppl - 1) / (32LL * ppl) > FPS_BIG_MAX_SLABS) ++ppl;
  *PPL = ppl;
  *S = (int)((N + 32LL * ppl - 1) / (32LL * ppl));
}

}  // namespace mvp

extern "C" int64_t mvp_fps_workspace_bytes(int64_t B, int64_t N, int64_t D, int64_t M, int dtype) {
  (void)M;
  if (B <= 0 || N <= 0) return 0;
  if (mvp::fits_regs(N, D, dtype)) return 0;
  if (mvp::fits_big(N, D, dtype)) {
    int S, PPL;
    mvp::big_geometry(N, &S, &PPL);
    return B * 5 * (int64_t)S * 32 * PPL * 4;
  }
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

  // MVPNET_B200_FPS=slab selects the shared-memory slab kernel for small clouds too.  Measured on B200 (32 clouds): it ties
  // the bucketed register kernel at N = 8192 (0.63 vs 0.64 us / iteration) and loses below (N = 2048: 0.97 vs 0.29 us): with
  // few slabs per warp the serial "update one active slab" chain (two redux.sync + a ballot, ~200 cycles) is longer than
  // the distance arithmetic it saves.  It is the default only where the cloud does not fit the register file (fps_big_kernel).
  static const bool use_slab = getenv("MVPNET_B200_FPS") != nullptr && getenv("MVPNET_B200_FPS")[0] == 's';
  if (fits_regs(N, D, dtype) && use_slab) {
    const int S = (int)((N + 31) / 32);
    int W = (S + 31) / 32;
    W = W < 1 ? 1 : (W > 8 ? 8 : W);
    const size_t smem = (size_t)S * 32 * 20 + (size_t)FPS_CELLS * sizeof(int);
    cudaFuncSetAttribute(fps_slab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    fps_slab_kernel<<<(unsigned)B, W * 32, smem, stream>>>((const float *)points, index, (int)N, (int)M, lg, S);
    return launch_status("fps");
  }
  if (fits_regs(N, D, dtype)) {
    // Bucketed kernel: 16 points per thread, i.e. 16 warps at N = 8192.  Once the distance work is pruned, what is left of an
    // iteration is the ISSUE cost of the common path every warp executes (centroid read, box test, barrier, block arg-max:
    // ~60 instructions): measured at N = 8192, 32 clouds: 32 warps x 8 points 0.640 us / iteration, 16 x 16 0.418 us,
    // 8 x 32 0.515 us (fewer warps, but each active warp then scans 32 points per lane).
    static const long long bucket_min = getenv("MVPNET_B200_FPS_BUCKET_MIN") ? atoll(getenv("MVPNET_B200_FPS_BUCKET_MIN")) : 4096;
    const bool bucket = N >= bucket_min;
    if (bucket) {
      int threads16 = (int)((N + 15) / 16);
      threads16 = (threads16 + 31) / 32 * 32;
      const size_t smem16 = (size_t)N * 3 * sizeof(float) + (size_t)N * sizeof(unsigned short);
      cudaFuncSetAttribute(fps_regs_kernel<16, true, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem16);
      fps_regs_kernel<16, true, 512><<<(unsigned)B, threads16, smem16, stream>>>((const float *)points, index, (int)N, (int)M, lg);
      return launch_status("fps");
    }
    // plain kernel: T must be a multiple of the reference BLOCK (= 1 << lg, <= 512) for the p-order tie rule (see the kernel)
    int threads = (1 << lg) < 32 ? 32 : (1 << lg);
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
    switch (ppt) {
      case 1: MVP_FPS_LAUNCH(1, false); break;
      case 2: MVP_FPS_LAUNCH(2, false); break;
      case 4: MVP_FPS_LAUNCH(4, false); break;
      default: MVP_FPS_LAUNCH(8, false); break;
    }
#undef MVP_FPS_LAUNCH
    return launch_status("fps");
  }

  MVP_REQUIRE(workspace, MVP_ERR_NULL, "fps: workspace of mvp_fps_workspace_bytes() bytes required");
  if (fits_big(N, D, dtype)) {
    int S, PPL;
    big_geometry(N, &S, &PPL);
    const size_t slots = (size_t)(S + 1023) / 1024 * 1024;
    const size_t smem = FPS_BIG_META * slots * 4 > (size_t)FPS_BIG_CELLS * 4 ? FPS_BIG_META * slots * 4 : (size_t)FPS_BIG_CELLS * 4;
    cudaFuncSetAttribute(fps_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    fps_big_kernel<<<(unsigned)B, 1024, smem, stream>>>((const float *)points, index, (float *)workspace, (int)N, (int)M, lg, S, PPL);
    return launch_status("fps");
  }
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
