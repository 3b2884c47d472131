// 3x3 / stride-1 / pad-1 convolution on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a, for the 2D
// network in front of FeatureAggregation (reference: mvpnet/models/unet_resnet34.py:9-125 — 92 % of its multiply-adds
// are such convolutions; MVPNet3D.forward, mvpnet_3d.py:94-99, spends ~90 % of a chunk's time there on fp32 cuDNN).
//
//   out[n,y,x,:] = act( bias + sum_{ky,kx} W[:, :, ky, kx] . in[n, y+ky-1, x+kx-1, :]  (+ residual[n,y,x,:]) )
//
// Tensors are fp32 NHWC; `in` may be the channel concatenation of two tensors (the UNet's cat([up, skip]),
// unet_resnet34.py:96-112, never materialised).  BatchNorm is folded into W / bias by the caller (eval mode).
//
// Implicit GEMM, M = pixels, N = Cout, K = 9 x Cin, with the same precision scheme as tc_mlp.cu: every fp32
// operand is split into bf16 hi + bf16 lo and each K-step issues three kind::f16 MMAs into one fp32 TMEM
// accumulator (hi*hi + hi*lo + lo*hi).
//
//   * Tile = 128 output pixels = 16 rows x 8 columns of one image (or 8 x 8 of two images when H <= 8).  The A
//     operand is the tile's input patch WITH its 1-pixel halo, staged once per 16-channel chunk in the canonical
//     K-major no-swizzle UMMA layout with the pixel's x coordinate as the fastest index: a core matrix (8 rows x
//     16 B) is 8 consecutive pixels of one image row, the stride between 8-row groups (SBO) is one halo row.  A
//     filter tap (dy, dx) is then nothing but a different START ADDRESS of the same staged patch — nine MMAs
//     groups read nine shifted views, no im2col copy exists anywhere.
//   * Weights stream through a shared-memory ring as 1-D bulk async copies (one stage = one tap of one 16-channel
//     chunk, hi | lo), and every stage is used by the TM (<= 4) pixel tiles a CTA keeps in flight (TM accumulators
//     in TMEM), which keeps the L2 -> SM weight traffic at 1/TM of the tensor pipe's appetite.
//   * Warp roles: 4 epilogue warps (TMEM -> bias / residual / ReLU -> global), 4 patch-producer warps (global fp32
//     -> bf16 hi/lo -> shared), one MMA issuer lane, one weight-producer lane; mbarrier hand-offs throughout.
#include "common.cuh"

namespace mvp {
namespace tcc {

using namespace tc;   // PTX wrappers of tc_mlp.cu

constexpr int HC = 10;                        // halo columns: 8 + 2
constexpr int HR_MAX = 20;                    // halo rows: 18 (one image) or 2 x 10 interleaved (two images)
constexpr int UNITS = HC * HR_MAX;            // 16-byte units of one 8-channel slab of a patch
constexpr int SLAB_BYTES = UNITS * 16;        // 3200: K-direction stride between core matrices (LBO)
constexpr int SLOT_HALF = 2 * SLAB_BYTES;     // hi (or lo) part of a patch: 16 channels
constexpr int SLOT_BYTES = 2 * SLOT_HALF;     // 12800
constexpr int MAX_TM = 4;
constexpr int MAX_STAGES = 8;
constexpr int EPI_THREADS = 128, PROD_THREADS = 128;
constexpr int THREADS = EPI_THREADS + PROD_THREADS + 64;

struct ConvArgs {
  const float *x1, *x2;
  int C1, C2;
  int N, H, W;
  const unsigned char *wp;     // [nb][chunk][tap][hi|lo][k8 (2)][n (Nt)][8] bf16
  const float *bias, *res;
  float *out;
  int Cout, Nt, NB;
  int relu;
  int ipt;                     // images per tile: 1 (16 rows of one image) or 2 (8 rows of two images)
  int TX, TY;                  // tiles per image along x / y
  long long ntiles, ngroups;
  int TM;                      // tiles per group (share every weight stage)
  int nchunks;                 // (C1 + C2) / 16
  int stages;
  int tmem_cols;
};

struct TileCoord { int n, y0, x0; };

__device__ __forceinline__ TileCoord tile_coord(const ConvArgs &a, long long t) {
  TileCoord c;
  if (a.ipt == 1) {
    const int per = a.TX * a.TY;
    c.n = (int)(t / per);
    const int r = (int)(t - (long long)c.n * per);
    c.y0 = (r / a.TX) * 16;
    c.x0 = (r % a.TX) * 8;
  } else {
    c.n = (int)(t / a.TX) * 2;
    c.y0 = 0;
    c.x0 = (int)(t % a.TX) * 8;
  }
  return c;
}

__global__ void __launch_bounds__(THREADS, 1)
tc_conv3x3_kernel(const ConvArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // shared memory: [2 slot sets][TM slots] patches | [stages] weight ring | barriers
  unsigned char *a_base = smem;
  const size_t stage_bytes = (size_t)64 * a.Nt;
  unsigned char *b_base = a_base + (size_t)2 * a.TM * SLOT_BYTES;
  uint64_t *bars = reinterpret_cast<uint64_t *>(b_base + (size_t)a.stages * stage_bytes);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * MAX_STAGES + 6);
  const uint32_t bar_bfull = smem_u32(bars), bar_bempty = smem_u32(bars + MAX_STAGES);
  const uint32_t bar_afull = smem_u32(bars + 2 * MAX_STAGES), bar_aempty = smem_u32(bars + 2 * MAX_STAGES + 2);
  const uint32_t bar_accfull = smem_u32(bars + 2 * MAX_STAGES + 4), bar_accempty = smem_u32(bars + 2 * MAX_STAGES + 5);

  if (tid == 0) {
    for (int s = 0; s < MAX_STAGES; ++s) { mbar_init(bar_bfull + 8 * s, 1); mbar_init(bar_bempty + 8 * s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(bar_afull + 8 * s, PROD_THREADS); mbar_init(bar_aempty + 8 * s, 1); }
    mbar_init(bar_accfull, 1);
    mbar_init(bar_accempty, EPI_THREADS);
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc(smem_u32(tmem_slot), (uint32_t)a.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const long long nworks = a.ngroups * a.NB;
  const int rows_h = a.ipt == 1 ? 18 : 20;                  // halo rows in use
  const int S = a.stages;

  if (warp < 4) {
    // =========================== epilogue: thread = TMEM lane = pixel ================================================
    const int row = warp * 32 + lane, g = row >> 3, xx = row & 7;
    const uint32_t t_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
    uint32_t it = 0;
    for (long long w = blockIdx.x; w < nworks; w += gridDim.x, ++it) {
      const int nb = (int)(w / a.ngroups);
      const long long group = w - (long long)nb * a.ngroups;
      mbar_wait(bar_accfull, it & 1u);
      tc_fence_after();
      for (int t = 0; t < a.TM; ++t) {
        const long long tile = group * a.TM + t;
        if (tile >= a.ntiles) break;
        const TileCoord tc_ = tile_coord(a, tile);
        const int img = a.ipt == 1 ? tc_.n : tc_.n + (g & 1);
        const int y = a.ipt == 1 ? tc_.y0 + g : (g >> 1);
        const int x = tc_.x0 + xx;
        const bool ok = img < a.N && y < a.H && x < a.W;
        const size_t pix = ((size_t)img * a.H + y) * a.W + x;
        const size_t obase = pix * a.Cout + (size_t)nb * a.Nt;
        uint32_t rn[16];
        tmem_ld16_issue(t_lane + (uint32_t)(t * a.Nt), rn);
        for (int c = 0; c < a.Nt; c += 16) {
          float v[16];
          tmem_ld_wait(rn);
#pragma unroll
          for (int q = 0; q < 16; ++q) v[q] = __uint_as_float(rn[q]);
          if (c + 16 < a.Nt) tmem_ld16_issue(t_lane + (uint32_t)(t * a.Nt + c + 16), rn);
          const float4 *bp = reinterpret_cast<const float4 *>(a.bias + nb * a.Nt + c);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 bq = __ldg(bp + q);
            v[4 * q] += bq.x; v[4 * q + 1] += bq.y; v[4 * q + 2] += bq.z; v[4 * q + 3] += bq.w;
          }
          if (ok) {
            if (a.res != nullptr) {
              const float4 *rp = reinterpret_cast<const float4 *>(a.res + obase + c);
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float4 rq = __ldg(rp + q);
                v[4 * q] += rq.x; v[4 * q + 1] += rq.y; v[4 * q + 2] += rq.z; v[4 * q + 3] += rq.w;
              }
            }
            if (a.relu) {
#pragma unroll
              for (int q = 0; q < 16; ++q) v[q] = fmaxf(v[q], 0.f);
            }
            float4 *op = reinterpret_cast<float4 *>(a.out + obase + c);
#pragma unroll
            for (int q = 0; q < 4; ++q) op[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(bar_accempty);
    }
  } else if (warp < 8) {
    // =========================== patch producers: global fp32 NHWC -> bf16 hi / lo patches ===========================
    const int ptid = tid - EPI_THREADS;
    const int nunits = rows_h * HC * 2;                     // (pixel, 8-channel slab) units of one patch
    uint32_t it = 0;
    for (long long w = blockIdx.x; w < nworks; w += gridDim.x) {
      const long long group = w % a.ngroups;
      for (int c = 0; c < a.nchunks; ++c, ++it) {
        const uint32_t ss = it & 1u;
        if (it >= 2) mbar_wait(bar_aempty + 8 * ss, ((it >> 1) - 1u) & 1u);
        const int k0 = c * 16;
        const float *src = k0 < a.C1 ? a.x1 : a.x2;
        const int C = k0 < a.C1 ? a.C1 : a.C2, ch = k0 < a.C1 ? k0 : k0 - a.C1;
        for (int t = 0; t < a.TM; ++t) {
          const long long tile = group * a.TM + t;
          if (tile >= a.ntiles) break;
          const TileCoord tc_ = tile_coord(a, tile);
          unsigned char *slot = a_base + (size_t)(ss * a.TM + t) * SLOT_BYTES;
          constexpr int U = 4;                               // nunits <= 400 < 4 * 128
          float4 lo4[U], hi4[U];
          int offs[U];
#pragma unroll
          for (int i = 0; i < U; ++i) {
            const int u = ptid + i * PROD_THREADS;
            offs[i] = -1;
            lo4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            hi4[i] = lo4[i];
            if (u < nunits) {
              const int k8 = u & 1, p = u >> 1, hr = p / HC, hc = p - hr * HC;
              offs[i] = k8 * SLAB_BYTES + p * 16;
              const int img = a.ipt == 1 ? tc_.n : tc_.n + (hr & 1);
              const int y = a.ipt == 1 ? tc_.y0 + hr - 1 : (hr >> 1) - 1;
              const int x = tc_.x0 + hc - 1;
              if (img < a.N && y >= 0 && y < a.H && x >= 0 && x < a.W) {
                const float4 *gp = reinterpret_cast<const float4 *>(src + (((size_t)img * a.H + y) * a.W + x) * C + ch + k8 * 8);
                lo4[i] = __ldg(gp);
                hi4[i] = __ldg(gp + 1);
              }
            }
          }
#pragma unroll
          for (int i = 0; i < U; ++i) {
            if (offs[i] >= 0) {
              uint32_t h[4], l[4];
              split_pair(lo4[i].x, lo4[i].y, h[0], l[0]);
              split_pair(lo4[i].z, lo4[i].w, h[1], l[1]);
              split_pair(hi4[i].x, hi4[i].y, h[2], l[2]);
              split_pair(hi4[i].z, hi4[i].w, h[3], l[3]);
              *reinterpret_cast<uint4 *>(slot + offs[i]) = make_uint4(h[0], h[1], h[2], h[3]);
              *reinterpret_cast<uint4 *>(slot + SLOT_HALF + offs[i]) = make_uint4(l[0], l[1], l[2], l[3]);
            }
          }
        }
        fence_proxy_async();
        mbar_arrive(bar_afull + 8 * ss);
      }
    }
  } else if (warp == 8) {
    // =========================== MMA issuer ===========================================================================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(128, a.Nt);
      uint32_t a_it = 0, b_it = 0, w_it = 0;
      for (long long w = blockIdx.x; w < nworks; w += gridDim.x, ++w_it) {
        const long long group = w % a.ngroups;
        const long long left = a.ntiles - group * a.TM;
        const int nt = left < a.TM ? (int)left : a.TM;
        if (w_it > 0) { mbar_wait(bar_accempty, (w_it - 1u) & 1u); tc_fence_after(); }
        for (int c = 0; c < a.nchunks; ++c, ++a_it) {
          const uint32_t ss = a_it & 1u;
          mbar_wait(bar_afull + 8 * ss, (a_it >> 1) & 1u);
          tc_fence_after();
          for (int tap = 0; tap < 9; ++tap, ++b_it) {
            const uint32_t s = b_it % (uint32_t)S;
            mbar_wait(bar_bfull + 8 * s, (b_it / (uint32_t)S) & 1u);
            tc_fence_after();
            const int dy = tap / 3 - 1, dx = tap - (tap / 3) * 3 - 1;
            const uint32_t aoff = (uint32_t)(((a.ipt * (1 + dy)) * HC + (1 + dx)) * 16);
            const uint32_t b_hi = smem_u32(b_base + (size_t)s * stage_bytes), b_lo = b_hi + 32u * (uint32_t)a.Nt;
            const uint64_t bh = make_desc(b_hi, (uint32_t)a.Nt * 16u, 128), bl = make_desc(b_lo, (uint32_t)a.Nt * 16u, 128);
            for (int t = 0; t < nt; ++t) {
              const uint32_t a_hi = smem_u32(a_base + (size_t)(ss * a.TM + t) * SLOT_BYTES) + aoff, a_lo = a_hi + SLOT_HALF;
              const uint64_t ah = make_desc(a_hi, SLAB_BYTES, HC * 16), al = make_desc(a_lo, SLAB_BYTES, HC * 16);
              const uint32_t d = tmem_base + (uint32_t)(t * a.Nt);
              umma_bf16(d, ah, bh, idesc, (c == 0 && tap == 0) ? 0u : 1u);
              umma_bf16(d, ah, bl, idesc, 1u);
              umma_bf16(d, al, bh, idesc, 1u);
            }
            umma_commit(bar_bempty + 8 * s);
          }
          umma_commit(bar_aempty + 8 * ss);
        }
        umma_commit(bar_accfull);
      }
    }
    __syncwarp();
  } else {
    // =========================== weight producer ======================================================================
    if (lane == 0) {
      uint32_t it = 0;
      for (long long w = blockIdx.x; w < nworks; w += gridDim.x) {
        const int nb = (int)(w / a.ngroups);
        const unsigned char *wsrc = a.wp + (size_t)nb * a.nchunks * 9 * stage_bytes;
        for (int j = 0; j < a.nchunks * 9; ++j, ++it) {
          const uint32_t s = it % (uint32_t)S;
          if (it >= (uint32_t)S) mbar_wait(bar_bempty + 8 * s, ((it / (uint32_t)S) - 1u) & 1u);
          mbar_expect_tx(bar_bfull + 8 * s, (uint32_t)stage_bytes);
          bulk_g2s(smem_u32(b_base + (size_t)s * stage_bytes), wsrc + (size_t)j * stage_bytes, (uint32_t)stage_bytes, bar_bfull + 8 * s);
        }
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
}

}  // namespace tcc
}  // namespace mvp

extern "C" int64_t mvp_tc_conv3x3_weight_bytes(int64_t Cin, int64_t Cout) { return Cin * Cout * 9 * 4; }

extern "C" int mvp_tc_conv3x3(const float *x1, int64_t C1, const float *x2, int64_t C2, int64_t N, int64_t H, int64_t W,
                              const void *w_packed, const float *bias, int64_t Cout, const float *residual, int relu,
                              float *out, mvp_stream_t stream) {
  using namespace mvp;
  MVP_REQUIRE(N >= 0 && H > 0 && W > 0, MVP_ERR_INVALID_ARG, "tc_conv3x3: bad sizes");
  MVP_REQUIRE(C1 > 0 && C1 % 16 == 0 && C2 >= 0 && C2 % 16 == 0, MVP_ERR_UNSUPPORTED, "tc_conv3x3: input channels must be multiples of 16");
  MVP_REQUIRE(Cout > 0 && Cout % 16 == 0 && (Cout <= 256 || Cout % 256 == 0), MVP_ERR_UNSUPPORTED,
              "tc_conv3x3: output channels must be a multiple of 16, and of 256 above 256");
  MVP_REQUIRE(N * H * W < (1LL << 31), MVP_ERR_UNSUPPORTED, "tc_conv3x3: more than 2^31 pixels");
  if (N == 0) return 0;
  MVP_REQUIRE(x1 && w_packed && bias && out && (x2 || C2 == 0), MVP_ERR_NULL, "tc_conv3x3: null pointer");
  MVP_REQUIRE((((uintptr_t)x1 | (uintptr_t)x2 | (uintptr_t)w_packed | (uintptr_t)bias | (uintptr_t)residual | (uintptr_t)out) & 15) == 0,
              MVP_ERR_INVALID_ARG, "tc_conv3x3: pointers must be 16-byte aligned");
  tcc::ConvArgs a = {};
  a.x1 = x1; a.x2 = x2; a.C1 = (int)C1; a.C2 = (int)C2; a.N = (int)N; a.H = (int)H; a.W = (int)W;
  a.wp = (const unsigned char *)w_packed; a.bias = bias; a.res = residual; a.out = out; a.relu = relu;
  a.Cout = (int)Cout; a.Nt = Cout <= 256 ? (int)Cout : 256; a.NB = a.Cout / a.Nt;
  a.ipt = H <= 8 ? 2 : 1;
  a.TX = (int)((W + 7) / 8);
  a.TY = a.ipt == 1 ? (int)((H + 15) / 16) : 1;
  a.ntiles = a.ipt == 1 ? N * a.TX * a.TY : ((N + 1) / 2) * a.TX;
  a.TM = 512 / a.Nt < tcc::MAX_TM ? 512 / a.Nt : tcc::MAX_TM;
  // spread over the SMs before stacking tiles on one CTA
  while (a.TM > 1 && (a.ntiles + a.TM - 1) / a.TM * a.NB < sm_count()) a.TM >>= 1;
  a.ngroups = (a.ntiles + a.TM - 1) / a.TM;
  a.nchunks = (int)((C1 + C2) / 16);
  a.tmem_cols = 32;
  while (a.tmem_cols < a.TM * a.Nt) a.tmem_cols <<= 1;
  const size_t fixed = (size_t)2 * a.TM * tcc::SLOT_BYTES + 512;
  a.stages = tcc::MAX_STAGES;
  while (a.stages > 2 && fixed + (size_t)a.stages * 64 * a.Nt > tc::SMEM_CAP) --a.stages;
  const size_t smem = fixed + (size_t)a.stages * 64 * a.Nt;
  MVP_REQUIRE(smem <= tc::SMEM_CAP, MVP_ERR_UNSUPPORTED, "tc_conv3x3: shared memory budget exceeded");
  cudaError_t e = cudaFuncSetAttribute(tcc::tc_conv3x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("tc_conv3x3: smem attribute (%zu B): %s", smem, cudaGetErrorString(e)); return (int)e; }
  const long long nworks = a.ngroups * a.NB;
  // one CTA per SM (TMEM: up to 512 columns each); smaller shared-memory footprints could co-reside, which the
  // full-width TMEM allocation of a second CTA would turn into a dead-lock, so the grid never exceeds the SM count
  long long grid = sm_count();
  if (grid > nworks) grid = nworks;
  static const bool debug = getenv("MVPNET_B200_DEBUG") != nullptr;
  if (debug)
    fprintf(stderr, "[tc_conv3x3] N=%d H=%d W=%d Cin=%d+%d Cout=%d Nt=%d ipt=%d tiles=%lld TM=%d works=%lld stages=%d smem=%zu tmem=%d\n",
            a.N, a.H, a.W, a.C1, a.C2, a.Cout, a.Nt, a.ipt, a.ntiles, a.TM, nworks, a.stages, smem, a.tmem_cols);
  tcc::tc_conv3x3_kernel<<<(unsigned)grid, tcc::THREADS, smem, (cudaStream_t)stream>>>(a);
  return launch_status("tc_conv3x3");
}
