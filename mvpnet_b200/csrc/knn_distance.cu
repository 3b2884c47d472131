// 3-NN with squared distances for sm_100a.
//
// Semantics: mvpnet/ops/cuda/knn_distance_kernel.cu:35-124 — keys visited in index order, sorted
// insertion with strict `<`, i.e. the result is the first three keys under the total order
// (squared distance, key index), ascending; squared distances are returned.
//
// B200 design: one warp per query.  Lane l keeps a private sorted top-3 over the keys
// j = l (mod 32) — it sees them in increasing j, so strict `<` keeps the earlier index exactly as
// in the reference — and the 32 partial lists are merged at the end with three warp arg-min
// rounds on (distance, index).  Because (distance, index) is a total order the merge returns the
// same triple as the reference's single-thread scan.  Keys are staged per CTA in shared memory
// (128-bit coalesced loads, conflict-free stride-3 reads); outputs are written by lanes 0..2.
#include "common.cuh"
#include "point_grid.cu"

namespace mvp {

constexpr int KNN_WARPS = 8;

template <typename T>
__device__ __forceinline__ void top3_insert(T d, int j, T (&bd)[3], int (&bi)[3]) {
  if (d < bd[2]) {
    if (d < bd[1]) {
      bd[2] = bd[1]; bi[2] = bi[1];
      if (d < bd[0]) { bd[1] = bd[0]; bi[1] = bi[0]; bd[0] = d; bi[0] = j; }
      else { bd[1] = d; bi[1] = j; }
    } else { bd[2] = d; bi[2] = j; }
  }
}

// warp arg-min over (d, idx) lexicographic; every lane gets the winner
__device__ __forceinline__ void warp_argmin(float &d, int &idx) {
  // non-negative floats (and +inf) order like their bit patterns
  const unsigned db = __float_as_uint(d);
  const unsigned m = __reduce_min_sync(0xffffffffu, db);
  const unsigned im = __reduce_min_sync(0xffffffffu, db == m ? (unsigned)idx : 0xffffffffu);
  d = __uint_as_float(m);
  idx = (int)im;
}
__device__ __forceinline__ void warp_argmin(double &d, int &idx) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double od = __shfl_xor_sync(0xffffffffu, d, o);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (od < d || (od == d && (unsigned)oi < (unsigned)idx)) { d = od; idx = oi; }
  }
}

template <typename T, int QPW>
__global__ void __launch_bounds__(KNN_WARPS * 32)
knn3_kernel(const T *__restrict__ query, const T *__restrict__ key, int64_t *__restrict__ index,
            T *__restrict__ distance, int N1, int N2, int tile_keys, int blocks_per_cloud, const PgGrid<T> *__restrict__ grids) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  if (grids != nullptr && grids[blockIdx.x / blocks_per_cloud].use) return;   // this cloud is served by the grid kernel
  T *s_key = reinterpret_cast<T *>(smem_raw);
  const int b = blockIdx.x / blocks_per_cloud;
  const int qblock = blockIdx.x % blocks_per_cloud;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const T *kbase = key + (size_t)b * N2 * 3;

  int qidx[QPW];
  T qx[QPW], qy[QPW], qz[QPW], bd[QPW][3];
  int bi[QPW][3];
#pragma unroll
  for (int q = 0; q < QPW; ++q) {
    qidx[q] = (qblock * KNN_WARPS + warp) * QPW + q;
    qx[q] = qy[q] = qz[q] = 0;
    if (qidx[q] < N1) {
      const T *qp = query + ((size_t)b * N1 + qidx[q]) * 3;
      qx[q] = qp[0], qy[q] = qp[1], qz[q] = qp[2];
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) { bd[q][k] = Inf<T>::v(); bi[q][k] = 0x7fffffff; }
  }

  for (int t0 = 0; t0 < N2; t0 += tile_keys) {
    const int tn = min(tile_keys, N2 - t0);
    if (t0 > 0) __syncthreads();
    stage_keys(s_key, kbase + (size_t)t0 * 3, tn * 3);
    __syncthreads();
    for (int j = lane; j < tn; j += 32) {
      const T kx = s_key[3 * j], ky = s_key[3 * j + 1], kz = s_key[3 * j + 2];
#pragma unroll
      for (int q = 0; q < QPW; ++q) {
        const T d = sqdist3(kx, ky, kz, qx[q], qy[q], qz[q]);
        top3_insert(d, t0 + j, bd[q], bi[q]);
      }
    }
  }

#pragma unroll
  for (int q = 0; q < QPW; ++q) {
    if (qidx[q] >= N1) continue;  // warp-uniform
    T od = 0;
    int oi = 0;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      T d = bd[q][0];
      int i = bi[q][0];
      warp_argmin(d, i);
      if (bd[q][0] == d && bi[q][0] == i) {  // this lane owned the winner: pop its head
        bd[q][0] = bd[q][1]; bi[q][0] = bi[q][1];
        bd[q][1] = bd[q][2]; bi[q][1] = bi[q][2];
        bd[q][2] = Inf<T>::v(); bi[q][2] = 0x7fffffff;
      }
      if (lane == r) { od = d; oi = i; }
    }
    if (lane < 3) {
      const size_t o = ((size_t)b * N1 + qidx[q]) * 3 + lane;
      index[o] = (int64_t)oi;
      distance[o] = od;
    }
  }
}

static inline bool knn_grid_eligible(int64_t B, int64_t N1, int64_t N2) {
  return B > 0 && B <= 65535 && pg_worthwhile(N1, N2);
}

template <typename T>
static int launch_knn3(const T *query, const T *key, int64_t B, int64_t N1, int64_t N2, int64_t *index,
                       T *distance, void *workspace, cudaStream_t stream) {
  // ---- exact uniform-grid search (point_grid.cu) for the clouds it suits; the exhaustive kernel below skips them
  const PgGrid<T> *grids = nullptr;
  if (workspace != nullptr && knn_grid_eligible(B, N1, N2)) {
    const PgWorkspace<T> w = pg_carve<T>(workspace, B, N2);
    if (int rc = pg_build<T>(key, B, N2, (T)0, /*min_cells=*/27, w, stream)) return rc;
    const int gbpc = (int)((N1 + PG_WARPS - 1) / PG_WARPS);
    MVP_REQUIRE(B * gbpc < (1LL << 31), MVP_ERR_UNSUPPORTED, "knn_distance: too many queries");
    pg_knn3_kernel<T><<<(unsigned)(B * gbpc), PG_WARPS * 32, 0, stream>>>(query, w.grids, w.cells, w.cell_stride, w.sorted, (int)N1,
                                                                          (int)N2, gbpc, index, distance);
    if (int rc = launch_status("knn_distance (grid)")) return rc;
    grids = w.grids;
  }
  const int64_t total_q = B * N1;
  int qpw = 4;
  while (qpw > 1 && (total_q + KNN_WARPS * qpw - 1) / (KNN_WARPS * qpw) < 2 * sm_count()) qpw >>= 1;
  const int64_t cap = (int64_t)(96 * 1024 / (3 * sizeof(T)));  // <= 96 KB of keys: two CTAs per SM
  const int64_t tile = N2 <= cap ? N2 : cap / 128 * 128;
  const size_t smem = ((size_t)tile * 3 * sizeof(T) + 15) / 16 * 16;
  const int bpc = (int)((N1 + KNN_WARPS * qpw - 1) / (KNN_WARPS * qpw));
  const int64_t grid = B * bpc;
  MVP_REQUIRE(grid < (1LL << 31), MVP_ERR_UNSUPPORTED, "knn_distance: too many queries");
#define MVP_KNN(Q)                                                                              \
  do {                                                                                          \
    auto kern = knn3_kernel<T, Q>;                                                              \
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);         \
    kern<<<(unsigned)grid, KNN_WARPS * 32, smem, stream>>>(query, key, index, distance, (int)N1, \
                                                           (int)N2, (int)tile, bpc, grids);     \
  } while (0)
  if (qpw == 4) MVP_KNN(4); else if (qpw == 2) MVP_KNN(2); else MVP_KNN(1);
#undef MVP_KNN
  return launch_status("knn_distance");
}

}  // namespace mvp

extern "C" int64_t mvp_knn_distance_workspace_bytes(int64_t B, int64_t N1, int64_t N2, int dtype) {
  using namespace mvp;
  if (!knn_grid_eligible(B, N1, N2)) return 0;
  return (int64_t)(dtype == MVP_F64 ? pg_workspace_bytes<double>(B, N2) : pg_workspace_bytes<float>(B, N2));
}

extern "C" int mvp_knn_distance(const void *query, const void *key, int64_t B, int64_t N1, int64_t N2, int64_t k,
                                int dtype, int64_t *index, void *distance, void *workspace, mvp_stream_t stream) {
  using namespace mvp;
  MVP_REQUIRE(dtype == MVP_F32 || dtype == MVP_F64, MVP_ERR_INVALID_ARG, "knn_distance: bad dtype");
  MVP_REQUIRE(k == 3, MVP_ERR_INVALID_ARG, "Only support 3-NN.");
  MVP_REQUIRE(N2 >= k, MVP_ERR_INVALID_ARG, "knn_distance: num_key (%lld) must be >= k", (long long)N2);
  MVP_REQUIRE(B >= 0 && N1 >= 0, MVP_ERR_INVALID_ARG, "knn_distance: negative size");
  MVP_REQUIRE(N1 < (1LL << 31) && N2 < (1LL << 31), MVP_ERR_UNSUPPORTED, "knn_distance: size too large");
  if (B == 0 || N1 == 0) return 0;
  MVP_REQUIRE(query && key && index && distance, MVP_ERR_NULL, "knn_distance: null pointer");
  if (dtype == MVP_F32)
    return launch_knn3<float>((const float *)query, (const float *)key, B, N1, N2, index, (float *)distance,
                              workspace, (cudaStream_t)stream);
  return launch_knn3<double>((const double *)query, (const double *)key, B, N1, N2, index, (double *)distance,
                             workspace, (cudaStream_t)stream);
}
