// 2D->3D k-NN: for every scene point the k nearest VALID unprojected pixels, for sm_100a.
//
// Semantics: what mvpnet/data/scannet_2d3d.py:298-313 obtains from scikit-learn's ball tree
// (float64 Euclidean, ascending) followed by the remap to flat pixel ids.  Restated as an exact
// search under the total order (squared distance in float64 = (dx*dx + dy*dy) + dz*dz without
// contraction, pixel id); see oracle/mvp_oracle.c (mvpo_knn_pixels).
//
// Two implementations with identical results:
//   * exhaustive (knn_pixels_brute_kernel): one warp per group of queries, pixels staged per CTA in
//     shared memory, per-lane sorted top-k, warp merge.  786 M fp64 distance evaluations per chunk.
//   * uniform grid (default): the valid pixels of each cloud are counting-sorted into cubic cells
//     (bbox -> count -> scan -> scatter, all on the device), then one WARP per query walks cubic shells
//     of cells around the query; the cells of one (y, z) row are contiguous in the sorted array, so the
//     lanes stream coalesced ranges.  After shell r every unvisited pixel is farther than r * cell, so
//     the walk stops as soon as the k-th best squared distance is below (r * cell)^2 — the result is
//     EXACTLY the exhaustive one (same arithmetic, same tie rule), at ~1/50 of the evaluations.
#include <stdlib.h>

#include "common.cuh"

namespace mvp {

constexpr int KP_WARPS = 8;
constexpr int KP_TILE = 1792;  // pixels per shared-memory tile: 42 KB xyz + 1.75 KB mask (static smem <= 48 KB)
constexpr int KP_QPW = 4;
constexpr int KP_CELL_CAP = 1 << 18;  // max grid cells per cloud
#ifndef MVP_KP_CELL_SCALE
#define MVP_KP_CELL_SCALE 2.0
#endif
constexpr double KP_CELL_SCALE = MVP_KP_CELL_SCALE;  // cell edge in units of the mean pixel spacing (pixels lie on surfaces: most cells are empty)

__device__ __forceinline__ double sqdist3_nofma(double kx, double ky, double kz, double qx, double qy, double qz) {
  const double dx = __dsub_rn(kx, qx), dy = __dsub_rn(ky, qy), dz = __dsub_rn(kz, qz);
  return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

// lexicographic (distance, id) sorted insertion into a per-thread list
template <int KMAX>
__device__ __forceinline__ void topk_insert(double d, int id, double (&bd)[KMAX], int (&bi)[KMAX]) {
  if (d < bd[KMAX - 1] || (d == bd[KMAX - 1] && id < bi[KMAX - 1])) {
    bd[KMAX - 1] = d; bi[KMAX - 1] = id;
#pragma unroll
    for (int s = KMAX - 1; s > 0; --s) {
      if (bd[s] < bd[s - 1] || (bd[s] == bd[s - 1] && bi[s] < bi[s - 1])) {
        const double td = bd[s]; bd[s] = bd[s - 1]; bd[s - 1] = td;
        const int ti = bi[s]; bi[s] = bi[s - 1]; bi[s - 1] = ti;
      }
    }
  }
}

// warp arg-min of (d, i), lexicographic; d >= 0 or +inf, so its bit pattern orders like an unsigned 64-bit integer:
// three `redux.sync` (high word, low word among the minima, index among those) instead of a shuffle butterfly
__device__ __forceinline__ void warp_argmin_d(double &d, int &i) {
  const unsigned long long bits = (unsigned long long)__double_as_longlong(d);
  const unsigned hi = (unsigned)(bits >> 32), lo = (unsigned)bits;
  const unsigned mh = __reduce_min_sync(0xffffffffu, hi);
  const unsigned ml = __reduce_min_sync(0xffffffffu, hi == mh ? lo : 0xffffffffu);
  const unsigned r = __reduce_min_sync(0xffffffffu, hi == mh && lo == ml ? (unsigned)i : 0xffffffffu);
  d = __longlong_as_double((long long)(((unsigned long long)mh << 32) | ml));
  i = (int)r;
}

// ------------------------------------------------------------------------------------------------
// exhaustive kernel
// ------------------------------------------------------------------------------------------------
template <int KMAX>
__global__ void __launch_bounds__(KP_WARPS * 32)
knn_pixels_brute_kernel(const double *__restrict__ query, const double *__restrict__ pix, const uint8_t *__restrict__ mask,
                        int nq, int P, int k, int blocks_per_cloud, int64_t *__restrict__ index, double *__restrict__ dist2) {
  __shared__ __align__(16) double s_xyz[KP_TILE * 3];
  __shared__ __align__(16) uint8_t s_mask[KP_TILE];
  const int b = blockIdx.x / blocks_per_cloud;
  const int qblock = blockIdx.x % blocks_per_cloud;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const double *pbase = pix + (size_t)b * P * 3;
  const uint8_t *mbase = mask + (size_t)b * P;

  int qidx[KP_QPW];
  double qx[KP_QPW], qy[KP_QPW], qz[KP_QPW], bd[KP_QPW][KMAX];
  int bi[KP_QPW][KMAX];
#pragma unroll
  for (int q = 0; q < KP_QPW; ++q) {
    qidx[q] = (qblock * KP_WARPS + warp) * KP_QPW + q;
    qx[q] = qy[q] = qz[q] = 0.0;
    if (qidx[q] < nq) {
      const double *qp = query + ((size_t)b * nq + qidx[q]) * 3;
      qx[q] = qp[0], qy[q] = qp[1], qz[q] = qp[2];
    }
#pragma unroll
    for (int j = 0; j < KMAX; ++j) { bd[q][j] = Inf<double>::v(); bi[q][j] = 0x7fffffff; }
  }

  for (int t0 = 0; t0 < P; t0 += KP_TILE) {
    const int tn = min(KP_TILE, P - t0);
    if (t0 > 0) __syncthreads();
    stage_keys(s_xyz, pbase + (size_t)t0 * 3, tn * 3);
    stage_keys(s_mask, mbase + t0, tn);
    __syncthreads();
    for (int j = lane; j < tn; j += 32) {
      if (!s_mask[j]) continue;
      const double kx = s_xyz[3 * j], ky = s_xyz[3 * j + 1], kz = s_xyz[3 * j + 2];
#pragma unroll
      for (int q = 0; q < KP_QPW; ++q)
        topk_insert<KMAX>(sqdist3_nofma(kx, ky, kz, qx[q], qy[q], qz[q]), t0 + j, bd[q], bi[q]);
    }
  }

#pragma unroll
  for (int q = 0; q < KP_QPW; ++q) {
    if (qidx[q] >= nq) continue;  // warp-uniform
    for (int r = 0; r < k; ++r) {
      double d = bd[q][0];
      int i = bi[q][0];
      warp_argmin_d(d, i);
      if (bi[q][0] == i && i != 0x7fffffff) {
#pragma unroll
        for (int s = 0; s < KMAX - 1; ++s) { bd[q][s] = bd[q][s + 1]; bi[q][s] = bi[q][s + 1]; }
        bd[q][KMAX - 1] = Inf<double>::v(); bi[q][KMAX - 1] = 0x7fffffff;
      }
      if (lane == 0) {
        const size_t o = ((size_t)b * nq + qidx[q]) * k + r;
        index[o] = i == 0x7fffffff ? (int64_t)-1 : (int64_t)i;
        if (dist2) dist2[o] = d;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// uniform grid
// ------------------------------------------------------------------------------------------------
struct KpGrid {          // one per cloud
  double ox, oy, oz;     // origin = min corner of the valid pixels
  double s, inv_s;       // cell edge
  int gx, gy, gz;        // cells per axis, gx*gy*gz <= KP_CELL_CAP
  int n_valid;
};

__device__ __forceinline__ int cell_coord(double p, double o, double inv_s, int g) {
  const int c = (int)floor(__dmul_rn(__dsub_rn(p, o), inv_s));
  return min(max(c, 0), g - 1);
}

// bbox of the valid pixels + grid parameters; one CTA per cloud
__global__ void __launch_bounds__(1024)
kp_bbox_kernel(const double *__restrict__ pix, const uint8_t *__restrict__ mask, int P, KpGrid *__restrict__ grids, double cell_scale) {
  __shared__ double s_lo[3][32], s_hi[3][32];
  __shared__ int s_cnt[32];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double *pb = pix + (size_t)b * P * 3;
  const uint8_t *mb = mask + (size_t)b * P;
  double lo[3] = {Inf<double>::v(), Inf<double>::v(), Inf<double>::v()};
  double hi[3] = {-Inf<double>::v(), -Inf<double>::v(), -Inf<double>::v()};
  int cnt = 0;
  for (int p = tid; p < P; p += 1024) {
    if (!mb[p]) continue;
    ++cnt;
#pragma unroll
    for (int a = 0; a < 3; ++a) { const double v = pb[3 * (size_t)p + a]; lo[a] = fmin(lo[a], v); hi[a] = fmax(hi[a], v); }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      lo[a] = fmin(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
      hi[a] = fmax(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
    }
  }
  if (lane == 0) { s_cnt[warp] = cnt; for (int a = 0; a < 3; ++a) { s_lo[a][warp] = lo[a]; s_hi[a][warp] = hi[a]; } }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < 32; ++w) {
      cnt += s_cnt[w];
      for (int a = 0; a < 3; ++a) { lo[a] = fmin(lo[a], s_lo[a][w]); hi[a] = fmax(hi[a], s_hi[a][w]); }
    }
    KpGrid g;
    g.n_valid = cnt;
    if (cnt == 0) {
      g.ox = g.oy = g.oz = 0.0; g.s = 1.0; g.inv_s = 1.0; g.gx = g.gy = g.gz = 1;
    } else {
      double ext[3];
      for (int a = 0; a < 3; ++a) ext[a] = fmax(hi[a] - lo[a], 1e-6);
      // cell edge ~ mean spacing of the pixels if they filled the box; grown until the grid fits the cap
      double s = cell_scale * cbrt(ext[0] * ext[1] * ext[2] / (double)min(cnt, KP_CELL_CAP / 2));
      s = fmax(s, 1e-6);
      int gx, gy, gz;
      for (;;) {
        gx = (int)fmin(floor(ext[0] / s) + 1.0, 1e6); gy = (int)fmin(floor(ext[1] / s) + 1.0, 1e6); gz = (int)fmin(floor(ext[2] / s) + 1.0, 1e6);
        if ((long long)gx * gy * gz <= KP_CELL_CAP) break;
        s *= 1.25;
      }
      g.ox = lo[0]; g.oy = lo[1]; g.oz = lo[2]; g.s = s; g.inv_s = 1.0 / s; g.gx = gx; g.gy = gy; g.gz = gz;
    }
    grids[b] = g;
  }
}

__global__ void __launch_bounds__(256)
kp_count_kernel(const double *__restrict__ pix, const uint8_t *__restrict__ mask, int P, const KpGrid *__restrict__ grids,
                int *__restrict__ cells) {
  const int b = blockIdx.y;
  const int p = blockIdx.x * 256 + threadIdx.x;
  if (p >= P || !mask[(size_t)b * P + p]) return;
  const KpGrid g = grids[b];
  const double *v = pix + ((size_t)b * P + p) * 3;
  const int c = (cell_coord(v[2], g.oz, g.inv_s, g.gz) * g.gy + cell_coord(v[1], g.oy, g.inv_s, g.gy)) * g.gx +
                cell_coord(v[0], g.ox, g.inv_s, g.gx);
  atomicAdd(cells + (size_t)b * (KP_CELL_CAP + 1) + c, 1);
}

// exclusive scan of the per-cell counts, in place; one CTA per cloud
__global__ void __launch_bounds__(1024)
kp_scan_kernel(const KpGrid *__restrict__ grids, int *__restrict__ cells) {
  __shared__ int s_warp[32];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const KpGrid g = grids[b];
  const int n = g.gx * g.gy * g.gz;
  int *c = cells + (size_t)b * (KP_CELL_CAP + 1);
  const int per = (n + 1023) / 1024;
  const int lo = min(tid * per, n), hi = min(lo + per, n);
  int sum = 0;
  for (int i = lo; i < hi; ++i) sum += c[i];
  int inc = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = s_warp[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
    s_warp[lane] = w;
  }
  __syncthreads();
  int run = inc - sum + (warp > 0 ? s_warp[warp - 1] : 0);
  for (int i = lo; i < hi; ++i) { const int t = c[i]; c[i] = run; run += t; }
}

#ifndef MVP_KP_SCAN_UNROLL
#define MVP_KP_SCAN_UNROLL 2
#endif
constexpr int KP_SCAN_UNROLL = MVP_KP_SCAN_UNROLL;   // record loads in flight per lane in the run scan of the query kernel

// one sorted pixel: 32 bytes, read by the query kernel as two 16-byte loads
struct __align__(16) KpRec { double x, y, z; long long id; };

// scatter: cells[] enters as the start offsets and leaves as the END offsets of every cell
__global__ void __launch_bounds__(256)
kp_scatter_kernel(const double *__restrict__ pix, const uint8_t *__restrict__ mask, int P, const KpGrid *__restrict__ grids,
                  int *__restrict__ cells, KpRec *__restrict__ sorted) {
  const int b = blockIdx.y;
  const int p = blockIdx.x * 256 + threadIdx.x;
  if (p >= P || !mask[(size_t)b * P + p]) return;
  const KpGrid g = grids[b];
  const double *v = pix + ((size_t)b * P + p) * 3;
  const double x = v[0], y = v[1], z = v[2];
  const int c = (cell_coord(z, g.oz, g.inv_s, g.gz) * g.gy + cell_coord(y, g.oy, g.inv_s, g.gy)) * g.gx +
                cell_coord(x, g.ox, g.inv_s, g.gx);
  const int pos = atomicAdd(cells + (size_t)b * (KP_CELL_CAP + 1) + c, 1);
  KpRec r;
  r.x = x; r.y = y; r.z = z; r.id = p;
  sorted[(size_t)b * P + pos] = r;
}

// ---- query order: the queries of a cloud are processed in the order of their grid cell (counting sort with the pixel
// grid's own cells), so that the warps of a CTA walk the SAME neighbourhoods and their record loads hit L1 instead of each
// paying an L2 round trip (chunk points arrive in random order; ncu round 1: 55 % of the stalls on those loads, L2 at 11 %).
// The order inside a cell is arbitrary: every query is answered independently and written to its own row.
__global__ void __launch_bounds__(256)
kp_qcount_kernel(const double *__restrict__ query, int nq, const KpGrid *__restrict__ grids, int *__restrict__ qcells) {
  const int b = blockIdx.y, q = blockIdx.x * 256 + threadIdx.x;
  if (q >= nq) return;
  const KpGrid g = grids[b];
  const double *v = query + ((size_t)b * nq + q) * 3;
  const int c = (cell_coord(v[2], g.oz, g.inv_s, g.gz) * g.gy + cell_coord(v[1], g.oy, g.inv_s, g.gy)) * g.gx + cell_coord(v[0], g.ox, g.inv_s, g.gx);
  atomicAdd(qcells + (size_t)b * (KP_CELL_CAP + 1) + c, 1);
}

__global__ void __launch_bounds__(256)
kp_qscatter_kernel(const double *__restrict__ query, int nq, const KpGrid *__restrict__ grids, int *__restrict__ qcells, int *__restrict__ order) {
  const int b = blockIdx.y, q = blockIdx.x * 256 + threadIdx.x;
  if (q >= nq) return;
  const KpGrid g = grids[b];
  const double *v = query + ((size_t)b * nq + q) * 3;
  const int c = (cell_coord(v[2], g.oz, g.inv_s, g.gz) * g.gy + cell_coord(v[1], g.oy, g.inv_s, g.gy)) * g.gx + cell_coord(v[0], g.ox, g.inv_s, g.gx);
  order[(size_t)b * nq + atomicAdd(qcells + (size_t)b * (KP_CELL_CAP + 1) + c, 1)] = q;
}

// 4 CTAs per SM: 64 registers without spills (98 registers unbounded = 2 CTAs; the kernel is latency-bound, ncu round 1: 30 % warps active)
template <int KMAX>
__device__ __forceinline__ void kp_query_block(int vblock, const double *__restrict__ query, const KpGrid *__restrict__ grids, const int *__restrict__ cells,
                                               const KpRec *__restrict__ sorted, const int *__restrict__ order, int nq, int P, int k,
                                               int blocks_per_cloud, int64_t *__restrict__ index, double *__restrict__ dist2) {
  const int b = vblock / blocks_per_cloud;
  const int slot = (vblock % blocks_per_cloud) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (slot >= nq) return;  // warp-uniform
  const int q = order ? __ldg(order + (size_t)b * nq + slot) : slot;
  const KpGrid g = grids[b];
  const int *ends = cells + (size_t)b * (KP_CELL_CAP + 1);
  const KpRec *srec = sorted + (size_t)b * P;
  const double *qp = query + ((size_t)b * nq + q) * 3;
  const double qx = qp[0], qy = qp[1], qz = qp[2];
  const int cx = cell_coord(qx, g.ox, g.inv_s, g.gx), cy = cell_coord(qy, g.oy, g.inv_s, g.gy), cz = cell_coord(qz, g.oz, g.inv_s, g.gz);
  const int rmax = max(max(max(cx, g.gx - 1 - cx), max(cy, g.gy - 1 - cy)), max(cz, g.gz - 1 - cz));

  double bd[KMAX];
  int bi[KMAX];
#pragma unroll
  for (int j = 0; j < KMAX; ++j) { bd[j] = Inf<double>::v(); bi[j] = 0x7fffffff; }

  // a contiguous run of the sorted pixel array, lanes in parallel.  The kernel is bound by the latency of these loads
  // (profile: 55 % of the stall samples on them, L2 at 11 %): every trip fetches the records of TWO sub-iterations
  // (two 16-byte loads each) before it evaluates either: 2.13 -> 1.79 ms with KP_SCAN_UNROLL = 2 sub-iterations in flight (4: 2.15 ms — 80 registers, spills, one CTA fewer per SM).
  auto scan = [&](int beg, int end) {
    constexpr int U = KP_SCAN_UNROLL;
    for (int p0 = beg; p0 < end; p0 += 32 * U) {
      double2 r0[U], r1[U];
      bool ok[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int p = p0 + 32 * u + lane;
        ok[u] = p < end;
        r0[u] = make_double2(0.0, 0.0);
        r1[u] = r0[u];
        if (ok[u]) { const double2 *r = reinterpret_cast<const double2 *>(srec + p); r0[u] = __ldg(r); r1[u] = __ldg(r + 1); }
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (ok[u]) topk_insert<KMAX>(sqdist3_nofma(r0[u].x, r0[u].y, r1[u].x, qx, qy, qz), (int)__double_as_longlong(r1[u].y), bd, bi);
    }
  };

  // distance from the query to the low / high face planes of its own cell (shell 0)
  const double face_lo[3] = {__dsub_rn(qx, __dadd_rn(g.ox, __dmul_rn((double)cx, g.s))), __dsub_rn(qy, __dadd_rn(g.oy, __dmul_rn((double)cy, g.s))),
                             __dsub_rn(qz, __dadd_rn(g.oz, __dmul_rn((double)cz, g.s)))};
  const double face_hi[3] = {__dsub_rn(__dadd_rn(g.ox, __dmul_rn((double)(cx + 1), g.s)), qx), __dsub_rn(__dadd_rn(g.oy, __dmul_rn((double)(cy + 1), g.s)), qy),
                             __dsub_rn(__dadd_rn(g.oz, __dmul_rn((double)(cz + 1), g.s)), qz)};
  double kth = Inf<double>::v();   // k-th best squared distance over the whole warp so far
  for (int r = 0; r <= rmax; ++r) {
    // Shell r = the (y, z) positions of a (2r+1)^2 square; a position on the square's rim contributes its
    // whole x extent [cx-r, cx+r] (one contiguous run), an interior position only its two end cells.
    // Lanes fetch the run boundaries of 32 positions at once (the dependent loads are the latency of this
    // kernel), then the warp streams the non-empty runs cooperatively.
    const int side = 2 * r + 1, npos = side * side;
    const float inv_side = __fdividef(1.0f, (float)side);
    const int x0 = max(cx - r, 0), x1 = min(cx + r, g.gx - 1);
    for (int base = 0; base < npos; base += 32) {
      const int t = base + lane;
      int beg0 = 0, end0 = 0, beg1 = 0, end1 = 0;
      if (t < npos) {
        const int tq = side <= 255 ? (int)__fmul_rn((float)t + 0.5f, inv_side) : t / side;   // floor(t / side), exact for small t
        const int dz = tq - r, dy = t - tq * side - r;
        const int z = cz + dz, y = cy + dy;
        if (z >= 0 && z < g.gz && y >= 0 && y < g.gy) {
          const int row = (z * g.gy + y) * g.gx;
          if (abs(dz) == r || abs(dy) == r) {
            beg0 = row + x0 == 0 ? 0 : ends[row + x0 - 1];
            end0 = ends[row + x1];
          } else {
            if (cx - r >= 0) { const int c = row + cx - r; beg0 = c == 0 ? 0 : ends[c - 1]; end0 = ends[c]; }
            if (cx + r < g.gx) { const int c = row + cx + r; beg1 = ends[c - 1]; end1 = ends[c]; }
          }
        }
      }
      unsigned active = __ballot_sync(0xffffffffu, end0 > beg0 || end1 > beg1);
      while (active) {
        const int src = __ffs(active) - 1;
        active &= active - 1;
        scan(__shfl_sync(0xffffffffu, beg0, src), __shfl_sync(0xffffffffu, end0, src));
        scan(__shfl_sync(0xffffffffu, beg1, src), __shfl_sync(0xffffffffu, end1, src));
      }
    }
    // k-th smallest over the 32 sorted lists (non-destructive merge)
    {
      int head = 0;
      double d = Inf<double>::v();
      for (int j = 0; j < k; ++j) {
        double cd = Inf<double>::v();
        int ci = 0x7fffffff;
#pragma unroll
        for (int s = 0; s < KMAX; ++s) if (s == head) { cd = bd[s]; ci = bi[s]; }
        d = cd;
        int i = ci;
        warp_argmin_d(d, i);
        if (i == ci && i != 0x7fffffff) ++head;
      }
      kth = d;
    }
    // Every pixel outside the visited cube [c-r, c+r]^3 lies beyond one of the cube's faces that still has
    // cells behind it; its distance is at least the distance from the query to that face plane.
    // The face planes move outwards by one cell edge per shell: distance = (distance at r = 0) + r * s.  Two operations per
    // face instead of five (this was 9 % of the kernel's instructions); the rounding of either form (~1e-13 of the bound)
    // is far inside the 1e-9 margin of the stop rule below.
    double bound = Inf<double>::v();
    {
      const double rs = __dmul_rn((double)r, g.s);
      const int cv[3] = {cx, cy, cz}, gv[3] = {g.gx, g.gy, g.gz};
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        if (cv[a] - r > 0) bound = fmin(bound, __dadd_rn(face_lo[a], rs));
        if (cv[a] + r + 1 < gv[a]) bound = fmin(bound, __dadd_rn(face_hi[a], rs));
      }
    }
    if (bound > 0.0 && kth < __dmul_rn(__dmul_rn(bound, bound), 1.0 - 1e-9)) break;
  }

  for (int j = 0; j < k; ++j) {
    double d = bd[0];
    int i = bi[0];
    warp_argmin_d(d, i);
    if (bi[0] == i && i != 0x7fffffff) {
#pragma unroll
      for (int s = 0; s < KMAX - 1; ++s) { bd[s] = bd[s + 1]; bi[s] = bi[s + 1]; }
      bd[KMAX - 1] = Inf<double>::v(); bi[KMAX - 1] = 0x7fffffff;
    }
    if (lane == 0) {
      const size_t o = ((size_t)b * nq + q) * k + j;
      index[o] = i == 0x7fffffff ? (int64_t)-1 : (int64_t)i;
      if (dist2) dist2[o] = d;
    }
  }
}

// Warps are independent (one query each, no block-level synchronisation): the kernel walks virtual blocks with a grid
// stride, so the launch can be sized.  Default: every block its own CTA (4 resident per SM = the whole register file).
// MVPNET_B200_KP_CTAS_PER_SM = n > 0 launches n CTAs per SM instead: with n = 1 the search takes a quarter of the
// registers and can stay resident NEXT TO a persistent convolution CTA of the 2D network (it is not needed before the
// 2D network ends), instead of displacing it.
template <int KMAX>
__global__ void __launch_bounds__(256, 4)
kp_query_kernel(const double *__restrict__ query, const KpGrid *__restrict__ grids, const int *__restrict__ cells,
                const KpRec *__restrict__ sorted, const int *__restrict__ order, int nq, int P, int k,
                int blocks_per_cloud, int total_blocks, int64_t *__restrict__ index, double *__restrict__ dist2) {
  for (int vb = blockIdx.x; vb < total_blocks; vb += gridDim.x)
    kp_query_block<KMAX>(vb, query, grids, cells, sorted, order, nq, P, k, blocks_per_cloud, index, dist2);
}

static inline size_t kp_align(size_t x) { return (x + 255) / 256 * 256; }

}  // namespace mvp

extern "C" int64_t mvp_knn_pixels_workspace_bytes(int64_t B, int64_t nq, int64_t P, int64_t k) {
  using namespace mvp;
  (void)k;
  if (B <= 0 || P <= 0) return 0;
  return (int64_t)(kp_align(sizeof(KpGrid) * B) + 2 * kp_align(sizeof(int) * (size_t)B * (KP_CELL_CAP + 1)) +
                   kp_align(sizeof(KpRec) * (size_t)B * P) + kp_align(sizeof(int) * (size_t)B * (nq > 0 ? nq : 1)));
}

extern "C" int mvp_knn_pixels(const double *query, const double *pix_xyz, const uint8_t *mask, int64_t B, int64_t nq,
                              int64_t P, int64_t k, int64_t *index, double *dist2, void *workspace, mvp_stream_t stream_) {
  using namespace mvp;
  cudaStream_t stream = (cudaStream_t)stream_;
  MVP_REQUIRE(k >= 1 && k <= 8, MVP_ERR_INVALID_ARG, "knn_pixels: k must be in [1, 8]");
  MVP_REQUIRE(B >= 0 && nq >= 0 && P >= 0, MVP_ERR_INVALID_ARG, "knn_pixels: negative size");
  MVP_REQUIRE(nq < (1LL << 31) && P < (1LL << 31) - 1, MVP_ERR_UNSUPPORTED, "knn_pixels: size too large");
  if (B == 0 || nq == 0) return 0;
  MVP_REQUIRE(query && index && (P == 0 || (pix_xyz && mask)), MVP_ERR_NULL, "knn_pixels: null pointer");

  if (workspace == nullptr || P == 0) {  // exhaustive search (also the cross-check of the grid search in the tests)
    const int bpc = (int)((nq + KP_WARPS * KP_QPW - 1) / (KP_WARPS * KP_QPW));
    const int64_t grid = B * bpc;
    MVP_REQUIRE(grid < (1LL << 31), MVP_ERR_UNSUPPORTED, "knn_pixels: too many queries");
    if (k <= 3)
      knn_pixels_brute_kernel<3><<<(unsigned)grid, KP_WARPS * 32, 0, stream>>>(query, pix_xyz, mask, (int)nq, (int)P, (int)k, bpc, index, dist2);
    else
      knn_pixels_brute_kernel<8><<<(unsigned)grid, KP_WARPS * 32, 0, stream>>>(query, pix_xyz, mask, (int)nq, (int)P, (int)k, bpc, index, dist2);
    return launch_status("knn_pixels");
  }

  MVP_REQUIRE(B <= 65535, MVP_ERR_UNSUPPORTED, "knn_pixels: batch > 65535");
  unsigned char *w = (unsigned char *)workspace;
  KpGrid *grids = (KpGrid *)w;                     w += kp_align(sizeof(KpGrid) * B);
  int *cells = (int *)w;                           w += kp_align(sizeof(int) * (size_t)B * (KP_CELL_CAP + 1));
  int *qcells = (int *)w;                          w += kp_align(sizeof(int) * (size_t)B * (KP_CELL_CAP + 1));
  KpRec *sorted = (KpRec *)w;                      w += kp_align(sizeof(KpRec) * (size_t)B * P);
  int *order = (int *)w;
  cudaError_t e = cudaMemsetAsync(cells, 0, 2 * kp_align(sizeof(int) * (size_t)B * (KP_CELL_CAP + 1)), stream);
  if (e != cudaSuccess) { set_error("knn_pixels: memset failed: %s", cudaGetErrorString(e)); return (int)e; }
  static const double cell_scale = [] { const char *e = getenv("MVPNET_B200_KP_CELL_SCALE"); const double v = e ? atof(e) : 0.0; return v > 0.0 ? v : KP_CELL_SCALE; }();
  kp_bbox_kernel<<<(unsigned)B, 1024, 0, stream>>>(pix_xyz, mask, (int)P, grids, cell_scale);
  dim3 pgrid((unsigned)((P + 255) / 256), (unsigned)B);
  kp_count_kernel<<<pgrid, 256, 0, stream>>>(pix_xyz, mask, (int)P, grids, cells);
  kp_scan_kernel<<<(unsigned)B, 1024, 0, stream>>>(grids, cells);
  kp_scatter_kernel<<<pgrid, 256, 0, stream>>>(pix_xyz, mask, (int)P, grids, cells, sorted);
  static const bool sort_queries = getenv("MVPNET_B200_KP_SORT_QUERIES") == nullptr || getenv("MVPNET_B200_KP_SORT_QUERIES")[0] != '0';
  if (sort_queries) {
    dim3 qgrid((unsigned)((nq + 255) / 256), (unsigned)B);
    kp_qcount_kernel<<<qgrid, 256, 0, stream>>>(query, (int)nq, grids, qcells);
    kp_scan_kernel<<<(unsigned)B, 1024, 0, stream>>>(grids, qcells);
    kp_qscatter_kernel<<<qgrid, 256, 0, stream>>>(query, (int)nq, grids, qcells, order);
  }
  const int *ord = sort_queries ? order : nullptr;
  const int bpc = (int)((nq + 7) / 8);
  const int64_t grid = B * bpc;
  MVP_REQUIRE(grid < (1LL << 31), MVP_ERR_UNSUPPORTED, "knn_pixels: too many queries");
  static const int per_sm = [] { const char *e = getenv("MVPNET_B200_KP_CTAS_PER_SM"); return e ? atoi(e) : 0; }();
  const int64_t launch = per_sm > 0 && (int64_t)per_sm * sm_count() < grid ? (int64_t)per_sm * sm_count() : grid;
  if (k <= 3)
    kp_query_kernel<3><<<(unsigned)launch, 256, 0, stream>>>(query, grids, cells, sorted, ord, (int)nq, (int)P, (int)k, bpc, (int)grid, index, dist2);
  else
    kp_query_kernel<8><<<(unsigned)launch, 256, 0, stream>>>(query, grids, cells, sorted, ord, (int)nq, (int)P, (int)k, bpc, (int)grid, index, dist2);
  return launch_status("knn_pixels");
}
