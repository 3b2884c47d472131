// Exact uniform-grid acceleration of ball_query and the 3-NN for sm_100a.
//
// The reference kernels (ball_query_kernel.cu:58-135, knn_distance_kernel.cu:35-124) test every query against
// every key: 16.8 M distance evaluations per 8192-point chunk at SA1 / FP4, 1.6 G for a 200 k-point scene.  Here
// the keys of each cloud are counting-sorted into cubic cells on the device (bbox -> count -> scan -> scatter) and
// one WARP per query visits only the cells that can hold a result.  The arithmetic of every evaluated pair is
// the reference's (common.cuh sqdist3: the exact FMA chain), and the cell walks are conservative, so the results
// are bit-identical to the exhaustive kernels — indices, order, padding and distances:
//
//   ball query  cell edge >= R = r * 1.001, so the ball of a query touches at most 3 cells per axis.  A pair
//               with |dx| >= r in floating point can never satisfy d2 < r*r (rounding is monotone), so every hit
//               lies in the walked cells.  Hits arrive in cell order; "the first K in index order" are the K
//               smallest indices, recovered by a rank sort in shared memory (a running threshold keeps the hit
//               list bounded when a ball holds more than K + 64 keys).
//   3-NN        shells of cells around the query's cell; after shell r every unvisited key is farther than the
//               distance to the nearest face of the visited cube (minus a rounding allowance), and the walk
//               stops when the current third-best squared distance is strictly below that bound, which also
//               settles the lowest-index tie rule among the visited keys.
//
// A cloud whose bounding box is not finite (inf / nan coordinates), or whose grid degenerates (radius comparable
// to the extent), is left to the exhaustive kernel: both kernels are launched and each skips the clouds of the
// other by a per-cloud device flag, so no host synchronisation is needed.
#pragma once
#include <stdlib.h>

#include "common.cuh"

namespace mvp {

constexpr int PG_CELL_CAP_MAX = 1 << 18;
constexpr int PG_WARPS = 8;

template <typename T>
struct PgGrid {            // one per cloud
  T ox, oy, oz;            // origin = min corner of the keys
  T s, inv_s;              // cell edge
  T span;                  // |origin| + extent, summed over axes: scale of the coordinate rounding allowance
  int gx, gy, gz;          // cells per axis
  int use;                 // 1: this cloud is served by the grid kernels, 0: by the exhaustive kernel
};

template <typename T> struct PgRec;
template <> struct __align__(16) PgRec<float> { float x, y, z; int i; };
template <> struct __align__(16) PgRec<double> { double x, y, z; long long i; };

template <typename T>
__device__ __forceinline__ PgRec<T> pg_load(const PgRec<T> *p);
template <>
__device__ __forceinline__ PgRec<float> pg_load<float>(const PgRec<float> *p) {
  const float4 v = __ldg(reinterpret_cast<const float4 *>(p));
  PgRec<float> r; r.x = v.x; r.y = v.y; r.z = v.z; r.i = __float_as_int(v.w);
  return r;
}
template <>
__device__ __forceinline__ PgRec<double> pg_load<double>(const PgRec<double> *p) {
  const double2 a = __ldg(reinterpret_cast<const double2 *>(p)), b = __ldg(reinterpret_cast<const double2 *>(p) + 1);
  PgRec<double> r; r.x = a.x; r.y = a.y; r.z = b.x; r.i = __double_as_longlong(b.y);
  return r;
}

template <typename T>
__device__ __forceinline__ int pg_cell(T p, T o, T inv_s, int g) {   // monotone non-decreasing in p
  const T f = floor((p - o) * inv_s);
  // the comparisons send NaN to cell 0 and keep the int conversion defined
  return f >= (T)(g - 1) ? g - 1 : (f > (T)0 ? (int)f : 0);
}

template <typename T> __device__ __forceinline__ T pg_eps();            // rounding allowance per unit of span
template <> __device__ __forceinline__ float pg_eps<float>() { return 1e-5f; }
template <> __device__ __forceinline__ double pg_eps<double>() { return 1e-13; }

// ------------------------------------------------------------------------------------------------
// build
// ------------------------------------------------------------------------------------------------
// bbox + grid parameters; one CTA per cloud.  radius > 0: ball-query grid (cell edge >= R);
// radius == 0: k-NN grid (cell edge ~ cell_scale x mean key spacing).
template <typename T>
__global__ void __launch_bounds__(1024)
pg_bbox_kernel(const T *__restrict__ key, int N, PgGrid<T> *__restrict__ grids, T R, T cell_scale, int cell_cap, int min_cells) {
  __shared__ T s_lo[3][32], s_hi[3][32];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const T *kb = key + (size_t)b * N * 3;
  T lo[3] = {Inf<T>::v(), Inf<T>::v(), Inf<T>::v()};
  T hi[3] = {-Inf<T>::v(), -Inf<T>::v(), -Inf<T>::v()};
  for (int p = tid; p < N; p += 1024) {
#pragma unroll
    for (int a = 0; a < 3; ++a) { const T v = kb[3 * (size_t)p + a]; lo[a] = fmin(lo[a], v); hi[a] = fmax(hi[a], v); }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      lo[a] = fmin(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
      hi[a] = fmax(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
    }
  }
  if (lane == 0) { for (int a = 0; a < 3; ++a) { s_lo[a][warp] = lo[a]; s_hi[a][warp] = hi[a]; } }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < 32; ++w)
      for (int a = 0; a < 3; ++a) { lo[a] = fmin(lo[a], s_lo[a][w]); hi[a] = fmax(hi[a], s_hi[a][w]); }
    PgGrid<T> g;
    g.ox = g.oy = g.oz = 0; g.s = 1; g.inv_s = 1; g.span = 0; g.gx = g.gy = g.gz = 1; g.use = 0;
    bool finite = true;
    for (int a = 0; a < 3; ++a) finite = finite && lo[a] > -Inf<T>::v() && hi[a] < Inf<T>::v() && lo[a] <= hi[a];
    if (finite && N > 0) {
      T ext[3];
      for (int a = 0; a < 3; ++a) ext[a] = hi[a] - lo[a];
      T s;
      if (R > (T)0) {
        s = R;
      } else {
        // mean spacing of N keys filling the box (degenerate axes count as one cell edge)
        T vol = 1;
        int dims = 0;
        for (int a = 0; a < 3; ++a) if (ext[a] > (T)0) { vol *= ext[a]; ++dims; }
        s = dims == 0 ? (T)1 : cell_scale * (T)pow((double)vol / (double)N, 1.0 / dims);
      }
      const T emax = fmax(ext[0], fmax(ext[1], ext[2]));
      if (!(s > emax * (T)1e-6)) s = emax * (T)1e-6;     // at most 1e6 cells per axis before the cap loop
      if (!(s > (T)0)) s = 1;                            // all keys coincide
      int gx, gy, gz;
      for (int it = 0; it < 200; ++it) {
        gx = (int)fmin(floor(ext[0] / s) + (T)1, (T)2e6); gy = (int)fmin(floor(ext[1] / s) + (T)1, (T)2e6);
        gz = (int)fmin(floor(ext[2] / s) + (T)1, (T)2e6);
        if ((long long)gx * gy * gz <= cell_cap) break;
        s *= (T)1.25;
      }
      if ((long long)gx * gy * gz <= cell_cap && s < Inf<T>::v()) {
        g.ox = lo[0]; g.oy = lo[1]; g.oz = lo[2]; g.s = s; g.inv_s = (T)1 / s; g.gx = gx; g.gy = gy; g.gz = gz;
        g.span = fabs(lo[0]) + fabs(lo[1]) + fabs(lo[2]) + ext[0] + ext[1] + ext[2];
        g.use = (gx * gy * gz >= min_cells && g.inv_s < Inf<T>::v() && g.span < Inf<T>::v()) ? 1 : 0;
      }
    }
    grids[b] = g;
  }
}

template <typename T>
__device__ __forceinline__ int pg_cell_of(const PgGrid<T> &g, T x, T y, T z) {
  return (pg_cell(z, g.oz, g.inv_s, g.gz) * g.gy + pg_cell(y, g.oy, g.inv_s, g.gy)) * g.gx + pg_cell(x, g.ox, g.inv_s, g.gx);
}

template <typename T>
__global__ void __launch_bounds__(256)
pg_count_kernel(const T *__restrict__ key, int N, const PgGrid<T> *__restrict__ grids, int *__restrict__ cells, int cell_stride) {
  const int b = blockIdx.y;
  const int p = blockIdx.x * 256 + threadIdx.x;
  if (p >= N) return;
  const PgGrid<T> g = grids[b];
  if (!g.use) return;
  const T *v = key + ((size_t)b * N + p) * 3;
  atomicAdd(cells + (size_t)b * cell_stride + pg_cell_of(g, v[0], v[1], v[2]), 1);
}

// exclusive scan of the per-cell counts, in place; one CTA per cloud
template <typename T>
__global__ void __launch_bounds__(1024)
pg_scan_kernel(const PgGrid<T> *__restrict__ grids, int *__restrict__ cells, int cell_stride) {
  __shared__ int s_warp[32];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const PgGrid<T> g = grids[b];
  if (!g.use) return;
  const int n = g.gx * g.gy * g.gz;
  int *c = cells + (size_t)b * cell_stride;
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

// scatter: cells[] enters as the start offsets and leaves as the END offsets of every cell
template <typename T>
__global__ void __launch_bounds__(256)
pg_scatter_kernel(const T *__restrict__ key, int N, const PgGrid<T> *__restrict__ grids, int *__restrict__ cells, int cell_stride,
                  PgRec<T> *__restrict__ sorted) {
  const int b = blockIdx.y;
  const int p = blockIdx.x * 256 + threadIdx.x;
  if (p >= N) return;
  const PgGrid<T> g = grids[b];
  if (!g.use) return;
  const T *v = key + ((size_t)b * N + p) * 3;
  PgRec<T> r;
  r.x = v[0]; r.y = v[1]; r.z = v[2]; r.i = p;
  const int pos = atomicAdd(cells + (size_t)b * cell_stride + pg_cell_of(g, r.x, r.y, r.z), 1);
  sorted[(size_t)b * N + pos] = r;
}

// ------------------------------------------------------------------------------------------------
// ball query
// ------------------------------------------------------------------------------------------------
// Shared memory per warp: hits[cap] (+ dist[cap]) and the finished row[K] (+ drow[K]); cap = K + 64.
template <typename T, bool WITH_DIST>
__global__ void __launch_bounds__(PG_WARPS * 32)
pg_ball_query_kernel(const T *__restrict__ query, const PgGrid<T> *__restrict__ grids, const int *__restrict__ cells, int cell_stride,
                     const PgRec<T> *__restrict__ sorted, int N1, int N2, int K, T r2, T R, int blocks_per_cloud,
                     int64_t *__restrict__ index, T *__restrict__ distance) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int b = blockIdx.x / blocks_per_cloud;
  const PgGrid<T> g = grids[b];
  if (!g.use) return;                                           // CTA-uniform: the exhaustive kernel serves this cloud
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = (blockIdx.x % blocks_per_cloud) * PG_WARPS + warp;
  if (q >= N1) return;                                          // warp-uniform
  const int cap = K + 64;
  int *hits = reinterpret_cast<int *>(smem_raw) + (size_t)warp * (cap + K);
  int *row = hits + cap;
  T *hdist = reinterpret_cast<T *>(reinterpret_cast<int *>(smem_raw) + (size_t)PG_WARPS * (cap + K)) + (size_t)warp * (cap + K);
  T *drow = hdist + cap;
  const unsigned lt_mask = (1u << lane) - 1u;

  const T *qp = query + ((size_t)b * N1 + q) * 3;
  const T qx = qp[0], qy = qp[1], qz = qp[2];
  // cells touched by [q - R, q + R], widened by a rounding allowance (the end points are rounded sums)
  const T slack = pg_eps<T>() * (g.span + fabs(qx) + fabs(qy) + fabs(qz) + R);
  const int x0 = pg_cell(qx - R - slack, g.ox, g.inv_s, g.gx), x1 = pg_cell(qx + R + slack, g.ox, g.inv_s, g.gx);
  const int y0 = pg_cell(qy - R - slack, g.oy, g.inv_s, g.gy), y1 = pg_cell(qy + R + slack, g.oy, g.inv_s, g.gy);
  const int z0 = pg_cell(qz - R - slack, g.oz, g.inv_s, g.gz), z1 = pg_cell(qz + R + slack, g.oz, g.inv_s, g.gz);
  const int *ends = cells + (size_t)b * cell_stride;
  const PgRec<T> *sp = sorted + (size_t)b * N2;

  int count = 0;                 // entries in hits[]
  int limit = 0x7fffffff;        // once K hits are known: only smaller indices can still enter the result

  // keep the K smallest of hits[0..count) (ascending) in hits[0..K): rank sort through registers
  auto compact = [&]() {
    constexpr int PER = 8;       // cap <= 32 * PER is guaranteed by the launcher
    int e[PER], rk[PER];
    T ed[PER];
#pragma unroll
    for (int u = 0; u < PER; ++u) {
      const int i = lane + 32 * u;
      e[u] = i < count ? hits[i] : 0x7fffffff;
      if (WITH_DIST) ed[u] = i < count ? hdist[i] : (T)0;
      rk[u] = 0;
    }
    for (int j = 0; j < count; ++j) {
      const int h = hits[j];     // broadcast read
#pragma unroll
      for (int u = 0; u < PER; ++u) rk[u] += h < e[u] ? 1 : 0;
    }
    __syncwarp();
#pragma unroll
    for (int u = 0; u < PER; ++u) {
      if (lane + 32 * u < count && rk[u] < K) {
        hits[rk[u]] = e[u];
        if (WITH_DIST) hdist[rk[u]] = ed[u];
      }
    }
    __syncwarp();
    count = min(count, K);
    if (count == K) limit = hits[K - 1];
  };

  // The (y, z) rows of the query's cell block are contiguous runs of the sorted keys.  Their boundaries are fetched by
  // up to 32 lanes AT ONCE (round 1 fetched them row by row: two dependent global loads per row, nine rows, were the
  // latency chain of this kernel); the runs are then streamed in the same z-major, y-minor order as before.
  const int ny = y1 - y0 + 1, nrows = ny * (z1 - z0 + 1);
  for (int rbase = 0; rbase < nrows; rbase += 32) {
    int mybeg = 0, myend = 0;
    if (rbase + lane < nrows) {
      const int t = rbase + lane, z = z0 + t / ny, y = y0 + t - (t / ny) * ny;
      const int rowc = (z * g.gy + y) * g.gx;
      mybeg = rowc + x0 == 0 ? 0 : __ldg(ends + rowc + x0 - 1);
      myend = __ldg(ends + rowc + x1);
    }
    const int nr = min(32, nrows - rbase);
    for (int t = 0; t < nr; ++t) {
      const int beg = __shfl_sync(0xffffffffu, mybeg, t), end = __shfl_sync(0xffffffffu, myend, t);
      for (int p0 = beg; p0 < end; p0 += 32) {
        const int p = p0 + lane;
        bool hit = false;
        int ki = 0;
        T d = 0;
        if (p < end) {
          const PgRec<T> k = pg_load<T>(sp + p);
          d = sqdist3(k.x, k.y, k.z, qx, qy, qz);
          ki = (int)k.i;
          hit = d < r2 && ki < limit;
        }
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (m) {
          if (hit) {
            const int pos = count + __popc(m & lt_mask);
            hits[pos] = ki;
            if (WITH_DIST) hdist[pos] = d;
          }
          count += __popc(m);
          __syncwarp();
          if (count > cap - 32) compact();
        }
      }
    }
  }
  // final order: ascending index = the reference's visiting order
  const int found = min(count, K);
  {
    int base = 0;
    // rank sort in slices of 32 candidates against all hits
    for (; base < count; base += 32) {
      const int i = base + lane;
      const int e = i < count ? hits[i] : 0x7fffffff;
      int rk = 0;
      for (int j = 0; j < count; ++j) rk += hits[j] < e ? 1 : 0;
      if (i < count && rk < K) {
        row[rk] = e;
        if (WITH_DIST) drow[rk] = hdist[i];
      }
    }
  }
  __syncwarp();
  const int64_t pad = found > 0 ? (int64_t)row[0] : (int64_t)-1;
  int64_t *orow = index + ((size_t)b * N1 + q) * K;
  for (int k = lane; k < K; k += 32) orow[k] = k < found ? (int64_t)row[k] : pad;
  if (WITH_DIST) {
    T *od = distance + ((size_t)b * N1 + q) * K;
    for (int k = lane; k < K; k += 32) od[k] = k < found ? drow[k] : (T)-1;
  }
}

// ------------------------------------------------------------------------------------------------
// 3-NN
// ------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void pg_top3_insert(T d, int j, T (&bd)[3], int (&bi)[3]) {   // lexicographic (d, j)
  if (d < bd[2] || (d == bd[2] && j < bi[2])) {
    if (d < bd[1] || (d == bd[1] && j < bi[1])) {
      bd[2] = bd[1]; bi[2] = bi[1];
      if (d < bd[0] || (d == bd[0] && j < bi[0])) { bd[1] = bd[0]; bi[1] = bi[0]; bd[0] = d; bi[0] = j; }
      else { bd[1] = d; bi[1] = j; }
    } else { bd[2] = d; bi[2] = j; }
  }
}

// Warp arg-min of (d, i), lexicographic, every lane gets the result.  d is a squared distance or +inf (non-negative, so
// its bit pattern orders like an unsigned integer), i a key index or the 0x7fffffff sentinel: two (float) / three
// (double) `redux.sync` instead of a 5-step shuffle butterfly with a two-key compare — the ncu source view of
// pg_knn3_kernel (profiles/r2_knn3_lines.txt) had 20 % of its samples and 18 % of its instructions in that butterfly
// (the kernel is issue-bound: 86 % SM busy, 1437 instructions per query).
template <typename T>
__device__ __forceinline__ void pg_warp_argmin(T &d, int &i);
template <>
__device__ __forceinline__ void pg_warp_argmin<float>(float &d, int &i) {
  const unsigned db = __float_as_uint(d);
  const unsigned m = __reduce_min_sync(0xffffffffu, db);
  const unsigned r = __reduce_min_sync(0xffffffffu, db == m ? (unsigned)i : 0xffffffffu);
  d = __uint_as_float(m);
  i = (int)r;
}
template <>
__device__ __forceinline__ void pg_warp_argmin<double>(double &d, int &i) {
  const unsigned long long bits = (unsigned long long)__double_as_longlong(d);
  const unsigned hi = (unsigned)(bits >> 32), lo = (unsigned)bits;
  const unsigned mh = __reduce_min_sync(0xffffffffu, hi);
  const unsigned ml = __reduce_min_sync(0xffffffffu, hi == mh ? lo : 0xffffffffu);
  const unsigned r = __reduce_min_sync(0xffffffffu, hi == mh && lo == ml ? (unsigned)i : 0xffffffffu);
  d = __longlong_as_double((long long)(((unsigned long long)mh << 32) | ml));
  i = (int)r;
}

template <typename T>
__global__ void __launch_bounds__(PG_WARPS * 32)
pg_knn3_kernel(const T *__restrict__ query, const PgGrid<T> *__restrict__ grids, const int *__restrict__ cells, int cell_stride,
               const PgRec<T> *__restrict__ sorted, int N1, int N2, int blocks_per_cloud, int64_t *__restrict__ index,
               T *__restrict__ distance) {
  const int b = blockIdx.x / blocks_per_cloud;
  const PgGrid<T> g = grids[b];
  if (!g.use) return;
  const int lane = threadIdx.x & 31;
  const int q = (blockIdx.x % blocks_per_cloud) * PG_WARPS + (threadIdx.x >> 5);
  if (q >= N1) return;  // warp-uniform
  const int *ends = cells + (size_t)b * cell_stride;
  const PgRec<T> *sp = sorted + (size_t)b * N2;
  const T *qp = query + ((size_t)b * N1 + q) * 3;
  const T qx = qp[0], qy = qp[1], qz = qp[2];
  const int cx = pg_cell(qx, g.ox, g.inv_s, g.gx), cy = pg_cell(qy, g.oy, g.inv_s, g.gy), cz = pg_cell(qz, g.oz, g.inv_s, g.gz);
  const int rmax = max(max(max(cx, g.gx - 1 - cx), max(cy, g.gy - 1 - cy)), max(cz, g.gz - 1 - cz));
  const T slack = pg_eps<T>() * (g.span + fabs(qx) + fabs(qy) + fabs(qz));

  T bd[3] = {Inf<T>::v(), Inf<T>::v(), Inf<T>::v()};
  int bi[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff};
  auto scan = [&](int beg, int end) {  // a contiguous run of the sorted key array, lanes in parallel
    for (int p = beg + lane; p < end; p += 32) {
      const PgRec<T> k = pg_load<T>(sp + p);
      pg_top3_insert(sqdist3(k.x, k.y, k.z, qx, qy, qz), (int)k.i, bd, bi);
    }
  };

  for (int r = 0; r <= rmax; ++r) {
    // shell r: the (y, z) positions of a (2r+1)^2 square; a rim position contributes its whole x extent (one run),
    // an interior position its two end cells.  Lanes fetch the run boundaries of 32 positions at once.
    const int side = 2 * r + 1, npos = side * side;
    const float inv_side = __fdividef(1.0f, (float)side);
    const int x0 = max(cx - r, 0), x1 = min(cx + r, g.gx - 1);
    for (int base = 0; base < npos; base += 32) {
      const int t = base + lane;
      int beg0 = 0, end0 = 0, beg1 = 0, end1 = 0;
      if (t < npos) {
        const int tq = side <= 255 ? (int)__fmul_rn((float)t + 0.5f, inv_side) : t / side;   // floor(t / side) without an integer division (exact: checked exhaustively)
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
    if (r == rmax) break;
    // third smallest (d, i) over the 32 sorted lists (non-destructive merge)
    T kth = Inf<T>::v();
    {
      int head = 0;
      for (int j = 0; j < 3; ++j) {
        T cd = head == 0 ? bd[0] : (head == 1 ? bd[1] : (head == 2 ? bd[2] : Inf<T>::v()));
        int ci = head == 0 ? bi[0] : (head == 1 ? bi[1] : (head == 2 ? bi[2] : 0x7fffffff));
        T d = cd;
        int i = ci;
        pg_warp_argmin(d, i);
        if (i == ci && i != 0x7fffffff) ++head;
        kth = d;
      }
    }
    // every key outside the visited cube [c-r, c+r]^3 lies beyond a face of the cube that still has cells behind it
    T bound = Inf<T>::v();
    {
      const T qv[3] = {qx, qy, qz}, ov[3] = {g.ox, g.oy, g.oz};
      const int cv[3] = {cx, cy, cz}, gv[3] = {g.gx, g.gy, g.gz};
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        if (cv[a] - r > 0) bound = fmin(bound, qv[a] - (ov[a] + (T)(cv[a] - r) * g.s));
        if (cv[a] + r + 1 < gv[a]) bound = fmin(bound, (ov[a] + (T)(cv[a] + r + 1) * g.s) - qv[a]);
      }
    }
    bound -= slack;
    if (bound > (T)0 && kth < bound * bound * ((T)1 - (T)64 * pg_eps<T>())) break;
  }

  T od = 0;
  int oi = 0;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    T d = bd[0];
    int i = bi[0];
    pg_warp_argmin(d, i);
    if (bi[0] == i && i != 0x7fffffff) {
      bd[0] = bd[1]; bi[0] = bi[1]; bd[1] = bd[2]; bi[1] = bi[2];
      bd[2] = Inf<T>::v(); bi[2] = 0x7fffffff;
    }
    if (lane == j) { od = d; oi = i; }
  }
  if (lane < 3) {
    const size_t o = ((size_t)b * N1 + q) * 3 + lane;
    index[o] = (int64_t)oi;
    distance[o] = od;
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static inline size_t pg_align(size_t x) { return (x + 255) / 256 * 256; }

static inline int pg_cell_cap(int64_t N2) {
  int64_t c = 4096;
  while (c < 8 * N2 && c < PG_CELL_CAP_MAX) c <<= 1;
  return (int)c;
}

// grid search pays off from a few million pair evaluations per call
static inline bool pg_worthwhile(int64_t N1, int64_t N2) { return N2 >= 1024 && N1 * N2 >= (4LL << 20); }

template <typename T>
struct PgWorkspace {
  PgGrid<T> *grids;
  int *cells;
  PgRec<T> *sorted;
  int cell_stride;
};

template <typename T>
static size_t pg_workspace_bytes(int64_t B, int64_t N2) {
  return pg_align(sizeof(PgGrid<T>) * B) + pg_align(sizeof(int) * (size_t)B * (pg_cell_cap(N2) + 1)) +
         pg_align(sizeof(PgRec<T>) * (size_t)B * N2);
}

template <typename T>
static PgWorkspace<T> pg_carve(void *workspace, int64_t B, int64_t N2) {
  PgWorkspace<T> w;
  unsigned char *p = (unsigned char *)workspace;
  w.cell_stride = pg_cell_cap(N2) + 1;
  w.grids = (PgGrid<T> *)p;    p += pg_align(sizeof(PgGrid<T>) * B);
  w.cells = (int *)p;          p += pg_align(sizeof(int) * (size_t)B * w.cell_stride);
  w.sorted = (PgRec<T> *)p;
  return w;
}

// bbox -> count -> scan -> scatter.  R > 0: ball-query grid, else k-NN grid.
template <typename T>
static int pg_build(const T *key, int64_t B, int64_t N2, T R, int min_cells, const PgWorkspace<T> &w, cudaStream_t stream) {
  // k-NN grid: cell edge in units of the mean key spacing (keys lie on surfaces: most cells are empty)
  static const double cell_scale = [] { const char *e = getenv("MVPNET_B200_PG_CELL_SCALE"); const double v = e ? atof(e) : 0.0; return v > 0.0 ? v : 1.0; }();
  cudaError_t e = cudaMemsetAsync(w.cells, 0, sizeof(int) * (size_t)B * w.cell_stride, stream);
  if (e != cudaSuccess) { set_error("point_grid: memset failed: %s", cudaGetErrorString(e)); return (int)e; }
  pg_bbox_kernel<T><<<(unsigned)B, 1024, 0, stream>>>(key, (int)N2, w.grids, R, (T)cell_scale, w.cell_stride - 1, min_cells);
  dim3 pgrid((unsigned)((N2 + 255) / 256), (unsigned)B);
  pg_count_kernel<T><<<pgrid, 256, 0, stream>>>(key, (int)N2, w.grids, w.cells, w.cell_stride);
  pg_scan_kernel<T><<<(unsigned)B, 1024, 0, stream>>>(w.grids, w.cells, w.cell_stride);
  pg_scatter_kernel<T><<<pgrid, 256, 0, stream>>>(key, (int)N2, w.grids, w.cells, w.cell_stride, w.sorted);
  return launch_status("point_grid build");
}

}  // namespace mvp
