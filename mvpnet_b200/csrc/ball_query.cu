// Ball query (+ optional distances) for sm_100a.
//
// Semantics: mvpnet/ops/cuda/ball_query_kernel.cu:58-135 / ball_query_distance_kernel.cu:59-139 —
// first K keys IN INDEX ORDER with d2 < r*r (strict), tail padded with the first hit, rows without a
// hit stay -1 (host fill at ball_query_kernel.cu:164), distances of padded slots stay -1.
//
// B200 design: one WARP per query instead of the reference's one thread per query.  The 32 lanes
// test 32 consecutive keys, `ballot` + popc turn the hit mask into ordered output slots (index
// order is preserved by construction), and the scan stops as soon as K hits are known — the
// reference keeps scanning all N2 keys.  The keys of the cloud are staged once per CTA in shared
// memory (coalesced 128-bit loads when the cloud base is 16-byte aligned; stride-3 AoS reads are
// bank-conflict free), rows are assembled in shared memory and written as coalesced int64 runs.
#include "common.cuh"
#include "point_grid.cu"

namespace mvp {

constexpr int BQ_WARPS = 8;

template <typename T, bool WITH_DIST, int QPW>
__global__ void __launch_bounds__(BQ_WARPS * 32)
ball_query_kernel(const T *__restrict__ query, const T *__restrict__ key, int64_t *__restrict__ index,
                  T *__restrict__ distance, int N1, int N2, int K, T r2, int tile_keys, int blocks_per_cloud,
                  const PgGrid<T> *__restrict__ grids) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  if (grids != nullptr && grids[blockIdx.x / blocks_per_cloud].use) return;   // this cloud is served by the grid kernel
  T *s_key = reinterpret_cast<T *>(smem_raw);                               // [tile_keys*3]
  int *s_row = reinterpret_cast<int *>(s_key + (size_t)tile_keys * 3);      // [warps][QPW][K]
  T *s_dist = reinterpret_cast<T *>(s_row + BQ_WARPS * QPW * K);            // same shape (WITH_DIST)

  const int b = blockIdx.x / blocks_per_cloud;
  const int qblock = blockIdx.x % blocks_per_cloud;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  const T *kbase = key + (size_t)b * N2 * 3;

  int qidx[QPW], cnt[QPW];
  T qx[QPW], qy[QPW], qz[QPW];
#pragma unroll
  for (int q = 0; q < QPW; ++q) {
    qidx[q] = (qblock * BQ_WARPS + warp) * QPW + q;
    cnt[q] = 0;
    qx[q] = qy[q] = qz[q] = 0;
    if (qidx[q] < N1) {
      const T *qp = query + ((size_t)b * N1 + qidx[q]) * 3;
      qx[q] = qp[0], qy[q] = qp[1], qz[q] = qp[2];
    }
  }

  for (int t0 = 0; t0 < N2; t0 += tile_keys) {
    const int tn = min(tile_keys, N2 - t0);
    if (t0 > 0) __syncthreads();
    stage_keys(s_key, kbase + (size_t)t0 * 3, tn * 3);
    __syncthreads();
#pragma unroll
    for (int q = 0; q < QPW; ++q) {
      if (qidx[q] >= N1) continue;
      int *row = s_row + (warp * QPW + q) * K;
      T *drow = s_dist + (warp * QPW + q) * K;
      int c = cnt[q];
      for (int base = 0; base < tn && c < K; base += 128) {
        // 4 independent 32-key probes per step for ILP; order restored below
        T d[4];
        unsigned m[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = base + u * 32 + lane;
          const int jj = j < tn ? j : tn - 1;
          d[u] = sqdist3(s_key[3 * jj], s_key[3 * jj + 1], s_key[3 * jj + 2], qx[q], qy[q], qz[q]);
          m[u] = __ballot_sync(0xffffffffu, j < tn && d[u] < r2);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (m[u]) {
            const int pos = c + __popc(m[u] & lt_mask);
            if (((m[u] >> lane) & 1u) && pos < K) {
              row[pos] = t0 + base + u * 32 + lane;
              if (WITH_DIST) drow[pos] = d[u];
            }
            c += __popc(m[u]);
          }
        }
      }
      cnt[q] = c;
    }
  }
  __syncwarp();
#pragma unroll
  for (int q = 0; q < QPW; ++q) {
    if (qidx[q] >= N1) continue;
    const int *row = s_row + (warp * QPW + q) * K;
    const T *drow = s_dist + (warp * QPW + q) * K;
    const int c = min(cnt[q], K);
    const int64_t pad = c > 0 ? (int64_t)row[0] : (int64_t)-1;
    int64_t *orow = index + ((size_t)b * N1 + qidx[q]) * K;
    for (int k = lane; k < K; k += 32) orow[k] = k < c ? (int64_t)row[k] : pad;
    if (WITH_DIST) {
      T *od = distance + ((size_t)b * N1 + qidx[q]) * K;
      for (int k = lane; k < K; k += 32) od[k] = k < c ? drow[k] : (T)-1;
    }
  }
}

// the grid path keeps at most K + 64 <= 256 candidate hits per warp in registers while it sorts them
static inline bool bq_grid_eligible(int64_t B, int64_t N1, int64_t N2, float radius, int64_t K) {
  return B > 0 && B <= 65535 && K <= 192 && radius > 0.f && radius < 1e30f && pg_worthwhile(N1, N2);
}

template <typename T>
static int launch_ball_query(const T *query, const T *key, int64_t B, int64_t N1, int64_t N2, float radius,
                             int64_t K, int64_t *index, T *distance, void *workspace, cudaStream_t stream) {
  const T r = (T)radius;  // squared in the tensor dtype, ball_query_kernel.cu:73
  const T r2 = r * r;
  const bool wd = distance != nullptr;
  // ---- exact uniform-grid search (point_grid.cu) for the clouds it suits; the exhaustive kernel below skips them
  const PgGrid<T> *grids = nullptr;
  if (workspace != nullptr && bq_grid_eligible(B, N1, N2, radius, K)) {
    const PgWorkspace<T> w = pg_carve<T>(workspace, B, N2);
    const T R = r * (T)1.001;
    if (int rc = pg_build<T>(key, B, N2, R, /*min_cells=*/64, w, stream)) return rc;
    const int cap = (int)K + 64;
    const size_t gsm = (size_t)PG_WARPS * (cap + (int)K) * (sizeof(int) + (wd ? sizeof(T) : 0));
    const int gbpc = (int)((N1 + PG_WARPS - 1) / PG_WARPS);
    MVP_REQUIRE(B * gbpc < (1LL << 31), MVP_ERR_UNSUPPORTED, "ball_query: too many queries");
    if (wd) {
      auto kern = pg_ball_query_kernel<T, true>;
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsm);
      kern<<<(unsigned)(B * gbpc), PG_WARPS * 32, gsm, stream>>>(query, w.grids, w.cells, w.cell_stride, w.sorted, (int)N1, (int)N2,
                                                                 (int)K, r2, R, gbpc, index, distance);
    } else {
      auto kern = pg_ball_query_kernel<T, false>;
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsm);
      kern<<<(unsigned)(B * gbpc), PG_WARPS * 32, gsm, stream>>>(query, w.grids, w.cells, w.cell_stride, w.sorted, (int)N1, (int)N2,
                                                                 (int)K, r2, R, gbpc, index, distance);
    }
    if (int rc = launch_status("ball_query (grid)")) return rc;
    grids = w.grids;
  }
  // queries per warp: fewer when the grid would not fill the machine
  const int64_t total_q = B * N1;
  int qpw = 4;
  while (qpw > 1 && (total_q + BQ_WARPS * qpw - 1) / (BQ_WARPS * qpw) < 2 * sm_count()) qpw >>= 1;
  const size_t row_bytes = (size_t)BQ_WARPS * K * (sizeof(int) + (wd ? sizeof(T) : 0));
  const size_t smem_cap = 220 * 1024;
  while (qpw > 1 && row_bytes * qpw + 3 * sizeof(T) * 256 > smem_cap) qpw >>= 1;
  MVP_REQUIRE(row_bytes * qpw + 3 * sizeof(T) * 256 <= smem_cap, MVP_ERR_UNSUPPORTED,
              "ball_query: max_neighbors=%lld too large for on-chip row staging", (long long)K);
  // key tile: whole cloud when it fits next to the rows (two CTAs per SM for <=8192 fp32 keys)
  int64_t tile = N2;
  const int64_t max_tile = (int64_t)((smem_cap / 2 - row_bytes * qpw) / (3 * sizeof(T)));
  const int64_t big_tile = (int64_t)((smem_cap - row_bytes * qpw) / (3 * sizeof(T)));
  if (tile > max_tile) tile = tile <= big_tile ? tile : (max_tile > 4096 ? max_tile : big_tile);
  tile = tile >= N2 ? N2 : tile / 128 * 128;
  const size_t smem = ((size_t)tile * 3 * sizeof(T) + 15) / 16 * 16 + row_bytes * qpw;
  const int bpc = (int)((N1 + BQ_WARPS * qpw - 1) / (BQ_WARPS * qpw));
  const int64_t grid = B * bpc;
  MVP_REQUIRE(grid < (1LL << 31), MVP_ERR_UNSUPPORTED, "ball_query: too many queries");
#define MVP_BQ(WD, Q)                                                                                   \
  do {                                                                                                  \
    auto kern = ball_query_kernel<T, WD, Q>;                                                            \
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                 \
    kern<<<(unsigned)grid, BQ_WARPS * 32, smem, stream>>>(query, key, index, distance, (int)N1, (int)N2, \
                                                          (int)K, r2, (int)tile, bpc, grids);           \
  } while (0)
  if (wd) { if (qpw == 4) MVP_BQ(true, 4); else if (qpw == 2) MVP_BQ(true, 2); else MVP_BQ(true, 1); }
  else    { if (qpw == 4) MVP_BQ(false, 4); else if (qpw == 2) MVP_BQ(false, 2); else MVP_BQ(false, 1); }
#undef MVP_BQ
  return launch_status("ball_query");
}

}  // namespace mvp

extern "C" int64_t mvp_ball_query_workspace_bytes(int64_t B, int64_t N1, int64_t N2, float radius, int64_t K, int dtype) {
  using namespace mvp;
  if (!bq_grid_eligible(B, N1, N2, radius, K)) return 0;
  return (int64_t)(dtype == MVP_F64 ? pg_workspace_bytes<double>(B, N2) : pg_workspace_bytes<float>(B, N2));
}

extern "C" int mvp_ball_query(const void *query, const void *key, int64_t B, int64_t N1, int64_t N2, float radius,
                              int64_t K, int dtype, int64_t *index, void *distance, void *workspace, mvp_stream_t stream) {
  using namespace mvp;
  MVP_REQUIRE(dtype == MVP_F32 || dtype == MVP_F64, MVP_ERR_INVALID_ARG, "ball_query: bad dtype");
  MVP_REQUIRE(K > 0, MVP_ERR_INVALID_ARG, "ball_query: max_neighbors must be > 0");
  MVP_REQUIRE(B >= 0 && N1 >= 0 && N2 >= 0, MVP_ERR_INVALID_ARG, "ball_query: negative size");
  MVP_REQUIRE(N1 < (1LL << 31) && N2 < (1LL << 31) && K < (1LL << 20), MVP_ERR_UNSUPPORTED, "ball_query: size too large");
  if (B == 0 || N1 == 0) return 0;
  MVP_REQUIRE(query && index && (key || N2 == 0), MVP_ERR_NULL, "ball_query: null pointer");
  if (dtype == MVP_F32)
    return launch_ball_query<float>((const float *)query, (const float *)key, B, N1, N2, radius, K, index,
                                    (float *)distance, workspace, (cudaStream_t)stream);
  return launch_ball_query<double>((const double *)query, (const double *)key, B, N1, N2, radius, K, index,
                                   (double *)distance, workspace, (cudaStream_t)stream);
}
