// 2D->3D k-NN: for every scene point the k nearest VALID unprojected pixels, for sm_100a.
//
// Semantics: what mvpnet/data/scannet_2d3d.py:298-313 obtains from scikit-learn's ball tree
// (float64 Euclidean, ascending) followed by the remap to flat pixel ids.  Restated as an exact
// search under the total order (squared distance in float64 = (dx*dx + dy*dy) + dz*dz without
// contraction, pixel id); see oracle/mvp_oracle.c (mvpo_knn_pixels).
//
// This file holds the exhaustive kernel: one warp per group of QPW queries, pixels staged per CTA
// in shared memory (xyz as float64 + validity byte), every lane keeps a private sorted top-k over
// its residue class of pixels, lists are merged with k warp arg-min rounds.
#include "common.cuh"

namespace mvp {

constexpr int KP_WARPS = 8;
constexpr int KP_TILE = 1792;  // pixels per shared-memory tile: 42 KB xyz + 1.75 KB mask (static smem <= 48 KB)
constexpr int KP_QPW = 4;

__device__ __forceinline__ double sqdist3_nofma(double kx, double ky, double kz, double qx, double qy, double qz) {
  const double dx = __dsub_rn(kx, qx), dy = __dsub_rn(ky, qy), dz = __dsub_rn(kz, qz);
  return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

template <int KMAX>
__global__ void __launch_bounds__(KP_WARPS * 32)
knn_pixels_kernel(const double *__restrict__ query, const double *__restrict__ pix, const uint8_t *__restrict__ mask,
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
      for (int q = 0; q < KP_QPW; ++q) {
        const double d = sqdist3_nofma(kx, ky, kz, qx[q], qy[q], qz[q]);
        if (d < bd[q][KMAX - 1]) {
          // sorted insertion, strict <: an equal distance stays behind the earlier pixel id
          bd[q][KMAX - 1] = d; bi[q][KMAX - 1] = t0 + j;
#pragma unroll
          for (int s = KMAX - 1; s > 0; --s) {
            if (bd[q][s] < bd[q][s - 1]) {
              const double td = bd[q][s]; bd[q][s] = bd[q][s - 1]; bd[q][s - 1] = td;
              const int ti = bi[q][s]; bi[q][s] = bi[q][s - 1]; bi[q][s - 1] = ti;
            }
          }
        }
      }
    }
  }

#pragma unroll
  for (int q = 0; q < KP_QPW; ++q) {
    if (qidx[q] >= nq) continue;  // warp-uniform
    for (int r = 0; r < k; ++r) {
      double d = bd[q][0];
      int i = bi[q][0];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double od = __shfl_xor_sync(0xffffffffu, d, o);
        const int oi = __shfl_xor_sync(0xffffffffu, i, o);
        if (od < d || (od == d && oi < i)) { d = od; i = oi; }
      }
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

}  // namespace mvp

extern "C" int mvp_knn_pixels(const double *query, const double *pix_xyz, const uint8_t *mask, int64_t B, int64_t nq,
                              int64_t P, int64_t k, int64_t *index, double *dist2, mvp_stream_t stream) {
  using namespace mvp;
  MVP_REQUIRE(k >= 1 && k <= 8, MVP_ERR_INVALID_ARG, "knn_pixels: k must be in [1, 8]");
  MVP_REQUIRE(B >= 0 && nq >= 0 && P >= 0, MVP_ERR_INVALID_ARG, "knn_pixels: negative size");
  MVP_REQUIRE(nq < (1LL << 31) && P < (1LL << 31) - 1, MVP_ERR_UNSUPPORTED, "knn_pixels: size too large");
  if (B == 0 || nq == 0) return 0;
  MVP_REQUIRE(query && index && (P == 0 || (pix_xyz && mask)), MVP_ERR_NULL, "knn_pixels: null pointer");
  const int bpc = (int)((nq + KP_WARPS * KP_QPW - 1) / (KP_WARPS * KP_QPW));
  const int64_t grid = B * bpc;
  MVP_REQUIRE(grid < (1LL << 31), MVP_ERR_UNSUPPORTED, "knn_pixels: too many queries");
  if (k <= 3)
    knn_pixels_kernel<3><<<(unsigned)grid, KP_WARPS * 32, 0, (cudaStream_t)stream>>>(query, pix_xyz, mask, (int)nq, (int)P, (int)k, bpc, index, dist2);
  else
    knn_pixels_kernel<8><<<(unsigned)grid, KP_WARPS * 32, 0, (cudaStream_t)stream>>>(query, pix_xyz, mask, (int)nq, (int)P, (int)k, bpc, index, dist2);
  return launch_status("knn_pixels");
}
