// 3x3 / stride-1 / pad-1 convolution on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a, for the 2D
// network in front of FeatureAggregation (reference: mvpnet/models/unet_resnet34.py:9-125 — 92 % of its multiply-adds
// are such convolutions; MVPNet3D.forward, mvpnet_3d.py:94-99, spends ~90 % of a chunk's time there on fp32 cuDNN).
//
//   out[n,y,x,:] = act( bias + sum_{ky,kx} W[:, :, ky, kx] . in[n, y+ky-1, x+kx-1, :]  (+ residual[n,y,x,:]) )
//
// `in` may be the channel concatenation of two tensors (the UNet's cat([up, skip]), unet_resnet34.py:96-112, never
// materialised).  BatchNorm is folded into W / bias by the caller (eval mode).
//
// Precision: as in tc_mlp.cu every fp32 value is carried as bf16 hi + bf16 lo (lo = bf16(v - hi)) and each K-step
// issues three kind::f16 MMAs into one fp32 TMEM accumulator (hi*hi + hi*lo + lo*hi).  Activations LIVE in that form
// between layers ("split-planar"): two planes (hi, lo), each [N][C/8][H][W][8] bf16 — the same 4 bytes per value as
// fp32, but every 8-channel slab of an image is one contiguous H x W plane of 16-byte pixels, which is exactly the
// unit of the UMMA K-major no-swizzle operand layout.  (For H <= 8 two images share a tile and the planes are stored
// pair-interleaved, [N/2][C/8][H][2][W][8].)
//
// Implicit GEMM, M = pixels, N = Cout, K = 9 x Cin:
//   * Tile = 128 output pixels = 16 rows x 8 columns of one image (8 x 8 of two images when H <= 8).  ONE TMA box
//     load per (tile, 16-channel chunk, plane) brings the tile's input patch with its 1-pixel halo — zero padding is
//     the TMA out-of-bounds fill — straight into the operand layout: a core matrix (8 rows x 16 B) is 8 consecutive
//     pixels of an image row, the stride between 8-row groups (SBO) is one halo row.  A filter tap (dy, dx) is then
//     nothing but a different START ADDRESS of the same staged patch: nine groups of MMAs read nine shifted views,
//     no im2col copy exists anywhere and no thread ever touches the activations on their way in.
//   * Weights stream through a shared-memory ring as 1-D bulk copies (one stage = one tap of one 16-channel chunk,
//     hi | lo) and every stage is used by the TM (<= 4) pixel tiles a CTA keeps in flight (TM accumulators in TMEM),
//     which keeps the L2 -> SM weight traffic at 1/TM of the tensor pipe's appetite.
//   * TMEM holds two sets of accumulators where they fit (Cout <= 128): the epilogue of one tile group (TMEM ->
//     bias / residual / ReLU -> split -> global) overlaps the MMAs of the next.
//   * Warp roles: 4 epilogue warps, one MMA-issuer lane, one patch-producer lane (TMA), one weight-producer lane;
//     mbarrier hand-offs throughout.
#include <cuda.h>

#include <mutex>

#include "common.cuh"

namespace mvp {
namespace tcc {

using namespace tc;   // PTX wrappers of tc_mlp.cu

constexpr int HC = 10;                        // halo columns: 8 + 2
constexpr int SLOT_HALF = 2 * 200 * 16;       // hi (or lo) part of a patch: 2 slabs x (<= 20 halo rows x 10) x 16 B
constexpr int SLOT_BYTES = 2 * SLOT_HALF;     // 12800
constexpr int MAX_TM = 4;
constexpr int MAX_STAGES = 8;
constexpr int MAX_ASETS = 3;
constexpr int EPI_WARPS = 8;                    // two per TMEM lane quarter: alternate 16-column chunks
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int THREADS = EPI_THREADS + 96;      // + MMA issuer, patch producer, weight producer warps

struct ConvArgs {
  CUtensorMap m1h, m1l, m2h, m2l;   // split-planar inputs: hi / lo plane of x1 and x2
  int C1, C2;
  int N, H, W;
  const unsigned char *wp;          // [nb][chunk][tap][hi|lo][k8 (2)][n (Nt)][8] bf16
  const float *bias;
  const __nv_bfloat16 *res;         // split-planar residual (hi plane; lo at + plane_out) or null
  __nv_bfloat16 *out_p;             // split-planar output or null
  float *out_f;                     // fp32 NHWC output or null
  __nv_bfloat16 *out_rh, *out_rl;   // row-split output or null: bf16 hi / lo planes, each NHWC (one 2*Cout-byte row per pixel)
  long long plane_out;              // elements of one plane of res / out_p
  int Cout, Nt, NB;
  int relu;
  int ipt;                          // images per tile: 1 (16 rows of one image) or 2 (8 rows of two images)
  int TX, TY;                       // tiles per image along x / y
  long long ntiles, ngroups;
  int TM;                           // tiles per group (share every weight stage)
  int nacc;                         // accumulator sets in TMEM (1 or 2)
  int asets;                        // patch ring depth (sets of TM slots)
  int nchunks;                      // (C1 + C2) / 16
  int stages;                       // weight ring depth
  int tps;                          // taps per weight stage: 1, 3 (one filter row) or 9 (a whole chunk)
  int tmem_cols;
  int dbg;                          // MVPNET_B200_CONV_DBG experiment bits (timing studies only; results are wrong)
  unsigned int *sched;              // {next work item, retired CTAs}: zero between launches (self re-arming)
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

// element offset of pixel (n, y, x), slab c8 inside one split-planar plane
__device__ __forceinline__ size_t planar_off(int n, int c8, int y, int x, int C8, int H, int W, int pair) {
  if (pair) return (((((size_t)(n >> 1) * C8 + c8) * H + y) * 2 + (n & 1)) * W + x) * 8;
  return ((((size_t)n * C8 + c8) * H + y) * W + x) * 8;
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap *map, int c0, int c1, int c2, int c3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}

__device__ __forceinline__ void unpack8(const uint4 h, const uint4 l, float (&v)[8]) {   // hi + lo -> fp32
  const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(hw[i] << 16) + __uint_as_float(lw[i] << 16);
    v[2 * i + 1] = __uint_as_float(hw[i] & 0xffff0000u) + __uint_as_float(lw[i] & 0xffff0000u);
  }
}

// ---- dynamic work distribution ---------------------------------------------------------------------------------------
// Work items are handed out by a global counter instead of blockIdx striding: a CTA that starts late (its SM was busy
// with a kernel of the geometry stream: FPS holds 32 SMs for over a millisecond) simply takes fewer items, and the
// last wave is shared by whoever is free.  One warp of the CTA (the patch producer) draws the indices and publishes
// them to the other roles through a 4-deep shared-memory ring guarded by mbarriers.
constexpr int SCHED_DEPTH = 4;
constexpr int SCHED_CONSUMERS = 2 + EPI_WARPS;        // MMA issuer, weight producer, epilogue warps (one arrival each)

__device__ __forceinline__ int sched_produce(uint32_t k, volatile int *ring, uint32_t bar_full, uint32_t bar_empty, unsigned int *counter) {
  const uint32_t slot = k & (SCHED_DEPTH - 1);
  if (k >= SCHED_DEPTH) mbar_wait(bar_empty + 8 * slot, ((k / SCHED_DEPTH) - 1u) & 1u);
  if (elect_one()) {
    ring[slot] = (int)atomicAdd(counter, 1u);
    mbar_arrive(bar_full + 8 * slot);
  }
  __syncwarp();
  return ring[slot];
}
__device__ __forceinline__ int sched_consume(uint32_t k, volatile int *ring, uint32_t bar_full, uint32_t bar_empty) {
  const uint32_t slot = k & (SCHED_DEPTH - 1);
  mbar_wait(bar_full + 8 * slot, (k / SCHED_DEPTH) & 1u);
  const int w = ring[slot];
  __syncwarp();
  if (elect_one()) mbar_arrive(bar_empty + 8 * slot);
  __syncwarp();
  return w;
}
// the last CTA to run dry re-arms the counter pair for the next launch that uses it
__device__ __forceinline__ void sched_retire(unsigned int *counter) {
  if (elect_one()) {
    const unsigned int done = atomicAdd(counter + 1, 1u);
    if (done == gridDim.x - 1) { counter[0] = 0u; counter[1] = 0u; __threadfence(); }
  }
  __syncwarp();
}

// ---- epilogue of one work item (shared by the single-CTA and the CTA-pair kernel) ------------------------------------
// thread = TMEM lane = pixel of a tile; the two warps of a lane quarter take alternate 16-column chunks.  The (tile, chunk)
// items of a work item form one software pipeline: the residual of item i + 1 is in flight while item i is finished, and the
// first item's residual is requested BEFORE the accumulator barrier.
struct EpiThread {
  int g, xx;                // tile row / column of this thread's pixel
  uint32_t t_lane;          // TMEM address of this thread's lane (column 0)
  int C8o, c0, cpt, pair;
  size_t slab_stride;       // elements between slabs of one image
};

__device__ __forceinline__ EpiThread epi_thread(const ConvArgs &a, int warp, int lane, uint32_t tmem_base) {
  EpiThread e;
  const int quarter = warp & 3, half = warp >> 2, row = quarter * 32 + lane;
  e.g = row >> 3; e.xx = row & 7;
  e.t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
  e.C8o = a.Cout >> 3;
  e.pair = a.ipt == 2;
  e.slab_stride = (size_t)a.H * a.W * 8 * (e.pair ? 2 : 1);
  e.c0 = half * 16;
  e.cpt = (a.Nt - e.c0 + 31) / 32;                         // this thread's chunks per tile
  return e;
}

// tiles [tile0, tile0 + nt) of output block nb, accumulators in set `set`; waits for `bar_full` (phase `parity`) itself
__device__ __forceinline__ void epilogue_tiles(const ConvArgs &a, const EpiThread &e, int nb, long long tile0, int nt, uint32_t set,
                                               const float *s_bias, uint32_t bar_full, uint32_t parity) {
  const int g = e.g, xx = e.xx, pair = e.pair, c0 = e.c0, cpt = e.cpt;
  const size_t slab_stride = e.slab_stride;
  struct Item { size_t pbase, fbase; bool ok; };
  auto locate = [&](long long tile) {
    Item r;
    const TileCoord tc_ = tile_coord(a, tile);
    const int img = pair ? tc_.n + (g & 1) : tc_.n;
    const int y = pair ? (g >> 1) : tc_.y0 + g;
    const int x = tc_.x0 + xx;
    r.ok = img < a.N && y < a.H && x < a.W && !(a.dbg & 2);
    r.pbase = r.ok ? planar_off(img, nb * (a.Nt >> 3), y, x, e.C8o, a.H, a.W, pair) : 0;
    r.fbase = r.ok ? (((size_t)img * a.H + y) * a.W + x) * a.Cout + (size_t)nb * a.Nt : 0;
    return r;
  };
  auto load_res = [&](const Item &p, int c, uint4 (&rh)[2], uint4 (&rl)[2]) {
    if (a.res != nullptr && p.ok) {
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const size_t o = p.pbase + (size_t)((c >> 3) + s) * slab_stride;
        rh[s] = __ldg(reinterpret_cast<const uint4 *>(a.res + o));
        rl[s] = __ldg(reinterpret_cast<const uint4 *>(a.res + a.plane_out + o));
      }
    }
  };
  const float *bias_s = s_bias + nb * a.Nt;
  Item cur = locate(nt > 0 ? tile0 : 0);
  uint4 rh[2], rl[2];
  if (cpt > 0 && nt > 0) load_res(cur, c0, rh, rl);
  mbar_wait(bar_full, parity);
  tc_fence_after();
  for (int t = 0; t < nt && cpt > 0; ++t) {
    const uint32_t t_acc = e.t_lane + (uint32_t)((set * a.TM + t) * a.Nt);
    Item nxt = cur;
    if (t + 1 < nt) nxt = locate(tile0 + t + 1);
    uint32_t rn[16];
    tmem_ld16_issue(t_acc + (uint32_t)c0, rn);
    for (int c = c0; c < a.Nt; c += 32) {
      float v[16];
      tmem_ld_wait(rn);
#pragma unroll
      for (int q = 0; q < 16; ++q) v[q] = __uint_as_float(rn[q]);
      if (c + 32 < a.Nt) tmem_ld16_issue(t_acc + (uint32_t)(c + 32), rn);
      const float4 *bp = reinterpret_cast<const float4 *>(bias_s + c);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 bq = bp[q];
        v[4 * q] += bq.x; v[4 * q + 1] += bq.y; v[4 * q + 2] += bq.z; v[4 * q + 3] += bq.w;
      }
      if (a.res != nullptr) {
        if (cur.ok) {
          float r0[8], r1[8];
          unpack8(rh[0], rl[0], r0);
          unpack8(rh[1], rl[1], r1);
#pragma unroll
          for (int q = 0; q < 8; ++q) { v[q] += r0[q]; v[8 + q] += r1[q]; }
        }
        if (c + 32 < a.Nt) load_res(cur, c + 32, rh, rl);          // next item: same tile, next chunk ...
        else if (t + 1 < nt) load_res(nxt, c0, rh, rl);              // ... or the first chunk of the next tile
      }
      if (cur.ok) {
        if (a.relu) {
#pragma unroll
          for (int q = 0; q < 16; ++q) v[q] = fmaxf(v[q], 0.f);
        }
        if (a.out_p != nullptr) {
#pragma unroll
          for (int s = 0; s < 2; ++s) {
            uint32_t h[4], l[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) split_pair(v[8 * s + 2 * q], v[8 * s + 2 * q + 1], h[q], l[q]);
            const size_t o = cur.pbase + (size_t)((c >> 3) + s) * slab_stride;
            *reinterpret_cast<uint4 *>(a.out_p + o) = make_uint4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uint4 *>(a.out_p + a.plane_out + o) = make_uint4(l[0], l[1], l[2], l[3]);
          }
        }
        if (a.out_f != nullptr) {
          float *op = a.out_f + cur.fbase + c;
#pragma unroll
          for (int q = 0; q < 2; ++q)
            st_global_256(op + 8 * q, make_uint4(__float_as_uint(v[8 * q]), __float_as_uint(v[8 * q + 1]), __float_as_uint(v[8 * q + 2]), __float_as_uint(v[8 * q + 3])),
                          make_uint4(__float_as_uint(v[8 * q + 4]), __float_as_uint(v[8 * q + 5]), __float_as_uint(v[8 * q + 6]), __float_as_uint(v[8 * q + 7])));
        }
        if (a.out_rh != nullptr) {   // what the fused FeatureAggregation gathers: pixel-major rows, already split
          uint32_t h[8], l[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) split_pair(v[2 * q], v[2 * q + 1], h[q], l[q]);
          st_global_256(a.out_rh + cur.fbase + c, make_uint4(h[0], h[1], h[2], h[3]), make_uint4(h[4], h[5], h[6], h[7]));
          st_global_256(a.out_rl + cur.fbase + c, make_uint4(l[0], l[1], l[2], l[3]), make_uint4(l[4], l[5], l[6], l[7]));
        }
      }
    }
    cur = nxt;
  }
}

__global__ void __launch_bounds__(THREADS, 1)
tc_conv3x3_kernel(const __grid_constant__ ConvArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // shared memory: [asets][TM] patches | [stages] weight ring | barriers
  unsigned char *a_base = smem;
  const size_t tap_bytes = (size_t)64 * a.Nt, stage_bytes = tap_bytes * a.tps;
  unsigned char *b_base = a_base + (size_t)a.asets * a.TM * SLOT_BYTES;
  uint64_t *bars = reinterpret_cast<uint64_t *>(b_base + (size_t)a.stages * stage_bytes);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * MAX_STAGES + 2 * MAX_ASETS + 4);
  volatile int *s_ring = reinterpret_cast<volatile int *>(bars + 28);      // [SCHED_DEPTH] work indices
  const uint32_t bar_sfull = smem_u32(bars + 32), bar_sempty = smem_u32(bars + 32 + SCHED_DEPTH);
  float *s_bias = reinterpret_cast<float *>(bars + 48);        // [Cout]: the epilogue reads it once per chunk (L1 is a few KB here)
  const uint32_t bar_bfull = smem_u32(bars), bar_bempty = smem_u32(bars + MAX_STAGES);
  const uint32_t bar_afull = smem_u32(bars + 2 * MAX_STAGES), bar_aempty = smem_u32(bars + 2 * MAX_STAGES + MAX_ASETS);
  const uint32_t bar_accfull = smem_u32(bars + 2 * MAX_STAGES + 2 * MAX_ASETS), bar_accempty = bar_accfull + 16;

  if (tid == 0) {
    for (int s = 0; s < MAX_STAGES; ++s) { mbar_init(bar_bfull + 8 * s, 1); mbar_init(bar_bempty + 8 * s, 1); }
    for (int s = 0; s < MAX_ASETS; ++s) { mbar_init(bar_afull + 8 * s, 1); mbar_init(bar_aempty + 8 * s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(bar_accfull + 8 * s, 1); mbar_init(bar_accempty + 8 * s, EPI_THREADS); }
    for (int s = 0; s < SCHED_DEPTH; ++s) { mbar_init(bar_sfull + 8 * s, 1); mbar_init(bar_sempty + 8 * s, SCHED_CONSUMERS); }
    fence_barrier_init();
  }
  if (warp == EPI_WARPS) tmem_alloc(smem_u32(tmem_slot), (uint32_t)a.tmem_cols);
  for (int i = tid; i < a.Cout; i += THREADS) s_bias[i] = a.bias[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int ngroups = (int)a.ngroups, nworks = ngroups * a.NB;
  const int rows_h = a.ipt == 1 ? 18 : 20;                  // halo rows of a patch
  const uint32_t slab_bytes = (uint32_t)(rows_h * HC * 16); // K-direction stride between core matrices (LBO)
  const uint32_t S = (uint32_t)a.stages, AS = (uint32_t)a.asets, NA = (uint32_t)a.nacc;
  const int pair = a.ipt == 2;

  if (warp < EPI_WARPS) {
    // =========================== epilogue (epilogue_tiles above) ======================================================
    const EpiThread e = epi_thread(a, warp, lane, tmem_base);
    for (uint32_t it = 0;; ++it) {
      const int w = sched_consume(it, s_ring, bar_sfull, bar_sempty);
      if (w >= nworks) break;
      const int nb = w / ngroups;
      const long long group = w - nb * ngroups;
      const uint32_t set = it % NA;
      const long long left = a.ntiles - group * a.TM;
      const int nt = left < a.TM ? (int)left : a.TM;
      epilogue_tiles(a, e, nb, group * a.TM, nt, set, s_bias, bar_accfull + 8 * set, (it / NA) & 1u);
      tc_fence_before();
      mbar_arrive(bar_accempty + 8 * set);
    }
  } else if (warp == EPI_WARPS) {
    // =========================== MMA issuer ===========================================================================
    // The WHOLE warp walks the loops (every value is warp-uniform, so descriptors live in uniform registers) and one
    // elected lane issues; a loop under `if (lane == 0)` makes ptxas wrap every tcgen05 instruction in a
    // thread-by-thread broadcast loop, which costs more than the MMAs of a narrow layer take to execute.
    const uint32_t idesc = make_idesc(128, a.Nt);
    const uint32_t a_s = smem_u32(a_base), b_s = smem_u32(b_base);
    // descriptors as (low word, high word): the low word holds the start address (>> 4, bits 0-13) and the K-direction
    // stride; stepping to another tap / plane / tile / stage is one 32-bit add on the low word.  Ring positions and
    // phases are running counters: this warp's instruction stream is the pacing item of narrow layers (a K-step of
    // three N = 64 MMAs executes in ~100 cycles), so no division, modulo or 64-bit arithmetic is left in the loops.
    const uint64_t adesc0 = make_desc(0, slab_bytes, HC * 16), bdesc0 = make_desc(0, (uint32_t)a.Nt * 16u, 128);
    const uint32_t a_lo0 = (uint32_t)adesc0 + (a_s >> 4), a_hi32 = (uint32_t)(adesc0 >> 32);
    const uint32_t b_lo0 = (uint32_t)bdesc0 + (b_s >> 4), b_hi32 = (uint32_t)(bdesc0 >> 32);
    const uint32_t set16 = (uint32_t)(a.TM * SLOT_BYTES) >> 4, stage16 = (uint32_t)stage_bytes >> 4, tap16 = (uint32_t)tap_bytes >> 4, lo_of_hi = 2u * (uint32_t)a.Nt;
    const uint32_t row16 = (uint32_t)(a.ipt * HC);            // one image row of the patch, in 16-byte units
    auto desc64 = [](uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; };
    uint32_t ss = 0, a_ph = 0, s = 0, b_ph = 0, set = 0, acc_ph = 0;
    for (uint32_t w_it = 0;; ++w_it) {
      const int w = sched_consume(w_it, s_ring, bar_sfull, bar_sempty);
      if (w >= nworks) break;
      const long long left = a.ntiles - (long long)(w % ngroups) * a.TM;
      const int nt = left < a.TM ? (int)left : a.TM;
      if (w_it >= NA) mbar_wait(bar_accempty + 8 * set, acc_ph ^ 1u);
      tc_fence_after();
      const uint32_t d0 = tmem_base + set * (uint32_t)(a.TM * a.Nt);
      uint32_t first = 0u;
      for (int c = 0; c < a.nchunks; ++c) {
        mbar_wait(bar_afull + 8 * ss, a_ph);
        const uint32_t a_org_lo = a_lo0 + ss * set16;          // tap (-1, -1): the patch origin
        // One weight stage holds tps taps (a filter row, or the whole 3x3).  Within a stage the loops run tile-outer,
        // tap-inner: consecutive MMAs accumulate into the SAME TMEM tile (3 x tps of them) before the accumulator
        // changes — switching accumulators after every tap cost a pipeline bubble per switch on the narrow layers.
        for (int st = 0; st < 9; st += a.tps) {
          mbar_wait(bar_bfull + 8 * s, b_ph);
          const uint32_t b_lo = b_lo0 + s * stage16;
          if (elect_one()) {
            if (!(a.dbg & 16)) {
#pragma unroll
              for (int t = 0; t < MAX_TM; ++t) {
                if (t < nt) {
                  const uint32_t d = d0 + (uint32_t)(t * a.Nt);
                  const uint32_t a_t = a_org_lo + (uint32_t)(t * (SLOT_BYTES >> 4));
                  uint32_t acc = first, bt = b_lo;
                  for (int j = 0; j < a.tps; ++j, bt += tap16) {
                    const int tap = st + j, dy = tap / 3, dx = tap - dy * 3;
                    const uint32_t lo = a_t + (uint32_t)dy * row16 + (uint32_t)dx;
                    const uint64_t ah = desc64(lo, a_hi32), al = desc64(lo + (SLOT_HALF >> 4), a_hi32);
                    const uint64_t bh = desc64(bt, b_hi32), bl = desc64(bt + lo_of_hi, b_hi32);
                    umma_bf16(d, ah, bh, idesc, acc);
                    umma_bf16(d, ah, bl, idesc, 1u);
                    umma_bf16(d, al, bh, idesc, 1u);
                    acc = 1u;
                  }
                }
              }
            }
            if (a.dbg & 64) mbar_arrive(bar_bempty + 8 * s); else umma_commit(bar_bempty + 8 * s);
          }
          __syncwarp();
          first = 1u;
          if (++s == S) { s = 0; b_ph ^= 1u; }
        }
        if (elect_one()) umma_commit(bar_aempty + 8 * ss);
        __syncwarp();
        if (++ss == AS) { ss = 0; a_ph ^= 1u; }
      }
      if (elect_one()) umma_commit(bar_accfull + 8 * set);
      __syncwarp();
      if (++set == NA) { set = 0; acc_ph ^= 1u; }
    }
  } else if (warp == EPI_WARPS + 1) {
    // =========================== patch producer: one TMA box per (tile, chunk, plane) ================================
    const uint32_t box_bytes = 2u * slab_bytes;
    const uint32_t a_s = smem_u32(a_base);
    uint32_t ss = 0, ph = 0, it = 0;
    for (uint32_t k = 0;; ++k) {
      const int w = sched_produce(k, s_ring, bar_sfull, bar_sempty, a.sched);
      if (w >= nworks) break;
      const long long group = w % ngroups;
      const long long left = a.ntiles - group * a.TM;
      const int nt = left < a.TM ? (int)left : a.TM;
      int cx[MAX_TM], cy[MAX_TM], cn[MAX_TM];               // tile coordinates: once per work item
#pragma unroll
      for (int t = 0; t < MAX_TM; ++t) {
        const TileCoord tc_ = tile_coord(a, group * a.TM + (t < nt ? t : 0));
        cx[t] = (tc_.x0 - 1) * 8; cy[t] = tc_.y0 - 1; cn[t] = pair ? (tc_.n >> 1) : tc_.n;
      }
      for (int c = 0; c < a.nchunks; ++c, ++it) {
        if (it >= AS) mbar_wait(bar_aempty + 8 * ss, ph ^ 1u);
        const int k0 = c * 16;
        const bool first = k0 < a.C1;
        const CUtensorMap *mh = first ? &a.m1h : &a.m2h, *ml = first ? &a.m1l : &a.m2l;
        const int C8 = (first ? a.C1 : a.C2) >> 3, s0 = (first ? k0 : k0 - a.C1) >> 3;
        const uint32_t bar = bar_afull + 8 * ss, dst0 = a_s + ss * (uint32_t)(a.TM * SLOT_BYTES);
        if (elect_one()) {
          if (a.dbg & 32) {                                  // timing study: no patch traffic
            mbar_arrive(bar);
          } else {
            mbar_expect_tx(bar, (uint32_t)nt * 2u * box_bytes);
#pragma unroll
            for (int t = 0; t < MAX_TM; ++t) {
              if (t < nt) {
                const uint32_t dst = dst0 + (uint32_t)(t * SLOT_BYTES);
                if (!pair) {               // dims (8 * W, H, C8 * N)
                  tma_load_3d(dst, mh, cx[t], cy[t], cn[t] * C8 + s0, bar);
                  tma_load_3d(dst + SLOT_HALF, ml, cx[t], cy[t], cn[t] * C8 + s0, bar);
                } else {                   // dims (8 * W, 2, H, C8 * N/2)
                  tma_load_4d(dst, mh, cx[t], 0, -1, cn[t] * C8 + s0, bar);
                  tma_load_4d(dst + SLOT_HALF, ml, cx[t], 0, -1, cn[t] * C8 + s0, bar);
                }
              }
            }
          }
        }
        __syncwarp();
        if (++ss == AS) { ss = 0; ph ^= 1u; }
      }
    }
    sched_retire(a.sched);
  } else {
    // =========================== weight producer ======================================================================
    const uint32_t b_s = smem_u32(b_base);
    const uint32_t nbytes = (a.dbg & 4) ? 16u : (uint32_t)stage_bytes;
    const int per_work = a.nchunks * 9 / a.tps;
    uint32_t s = 0, ph = 0, it = 0;
    for (uint32_t k = 0;; ++k) {
      const int w = sched_consume(k, s_ring, bar_sfull, bar_sempty);
      if (w >= nworks) break;
      const int nb = w / ngroups;
      const unsigned char *wsrc = a.wp + (size_t)nb * per_work * stage_bytes;
      for (int j = 0; j < per_work; ++j, ++it, wsrc += stage_bytes) {
        if (it >= S) mbar_wait(bar_bempty + 8 * s, ph ^ 1u);
        if (elect_one()) {
          if (a.dbg & 128) {
            mbar_arrive(bar_bfull + 8 * s);
          } else {
            mbar_expect_tx(bar_bfull + 8 * s, nbytes);
            bulk_g2s(b_s + s * (uint32_t)stage_bytes, wsrc, nbytes, bar_bfull + 8 * s);
          }
        }
        __syncwarp();
        if (++s == S) { s = 0; ph ^= 1u; }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == EPI_WARPS) tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
}

// ---- split-planar <-> fp32 NHWC (the boundary to the layers that still run on cuDNN) --------------------------------
// thread = (n, slab, y, x), x fastest: planar side coalesced 16-byte units, NHWC side one 32-byte sector per thread
__global__ void __launch_bounds__(256)
split_planar_kernel(const float *__restrict__ src, int N, int H, int W, int C, int pair, __nv_bfloat16 *__restrict__ dst, long long plane) {
  const int C8 = C >> 3;
  const long long total = (long long)N * C8 * H * W;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int x = (int)(i % W);
    long long r = i / W;
    const int y = (int)(r % H); r /= H;
    const int c8 = (int)(r % C8);
    const int n = (int)(r / C8);
    const float4 *p = reinterpret_cast<const float4 *>(src + (((size_t)n * H + y) * W + x) * C + c8 * 8);
    const float4 u = __ldg(p), v = __ldg(p + 1);
    uint32_t h[4], l[4];
    split_pair(u.x, u.y, h[0], l[0]); split_pair(u.z, u.w, h[1], l[1]);
    split_pair(v.x, v.y, h[2], l[2]); split_pair(v.z, v.w, h[3], l[3]);
    const size_t o = planar_off(n, c8, y, x, C8, H, W, pair);
    *reinterpret_cast<uint4 *>(dst + o) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4 *>(dst + plane + o) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

__global__ void __launch_bounds__(256)
merge_planar_kernel(const __nv_bfloat16 *__restrict__ src, long long plane, int N, int H, int W, int C, int pair, float *__restrict__ dst) {
  const int C8 = C >> 3;
  const long long total = (long long)N * C8 * H * W;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int x = (int)(i % W);
    long long r = i / W;
    const int y = (int)(r % H); r /= H;
    const int c8 = (int)(r % C8);
    const int n = (int)(r / C8);
    const size_t o = planar_off(n, c8, y, x, C8, H, W, pair);
    float v[8];
    unpack8(__ldg(reinterpret_cast<const uint4 *>(src + o)), __ldg(reinterpret_cast<const uint4 *>(src + plane + o)), v);
    float4 *p = reinterpret_cast<float4 *>(dst + (((size_t)n * H + y) * W + x) * C + c8 * 8);
    p[0] = make_float4(v[0], v[1], v[2], v[3]);
    p[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
}

// ---- host -----------------------------------------------------------------------------------------------------------
// Shared-memory budget of the convolution kernels (MVPNET_B200_CONV_SMEM_KB overrides).  Lowering it to 224 KB lets one
// small CTA of a geometry-stream kernel (the pixel k-NN with MVPNET_B200_KP_CTAS_PER_SM=1: no shared memory beyond the
// 1 KB the system reserves per CTA, a quarter of the register file) stay resident next to a persistent convolution CTA.
// Measured (round 2): no gain — the step was 1.8 % slower with the search spread under the convolutions than with the
// search taking the whole GPU for 1 ms — so the default is the full 227 KB.
static size_t conv_smem_cap() {
  static const size_t cap = [] { const char *e = getenv("MVPNET_B200_CONV_SMEM_KB"); const long v = e ? atol(e) : 0; return (size_t)((v >= 96 && v <= 227) ? v : 227) * 1024; }();
  return cap;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    return (EncodeTiledFn)p;
  }();
  return fn;
}

// Device-resident {next, retired} counter pairs for the dynamic work distribution.  A pair re-arms itself when its
// launch retires, so it can serve any number of launches that do not overlap in time.
//   * eager launches draw from a ring of EAGER pairs: two of them could only share a pair with > EAGER launches in flight;
//   * a launch recorded into a CUDA graph keeps its pair for the lifetime of the graph (the pointer is baked into the
//     kernel arguments), and graphs may be replayed concurrently with each other and with eager work on other streams:
//     captured launches therefore take pairs from a separate region that is never handed out again.
// The pool is allocated on first use; the first use must not be inside a stream capture (cudaMalloc is illegal there) —
// warm the path up once before capturing, as engine.GraphedForward does.
static unsigned int *sched_pair(cudaStream_t stream) {
  constexpr int DEVICES = 64, EAGER = 4096, CAPTURED = 61440;
  static unsigned int *pool[DEVICES] = {};
  static unsigned int next_eager[DEVICES] = {}, next_captured[DEVICES] = {};
  static std::mutex mu;                              // nn.DataParallel calls in from one thread per GPU
  std::lock_guard<std::mutex> lock(mu);
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= DEVICES) return nullptr;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &cap) != cudaSuccess) { cudaGetLastError(); cap = cudaStreamCaptureStatusNone; }
  if (pool[dev] == nullptr) {
    if (cap != cudaStreamCaptureStatusNone) return nullptr;
    unsigned int *p = nullptr;
    const size_t bytes = (size_t)(EAGER + CAPTURED) * 2 * sizeof(unsigned int);
    if (cudaMalloc(&p, bytes) != cudaSuccess || cudaMemset(p, 0, bytes) != cudaSuccess) return nullptr;
    pool[dev] = p;
  }
  if (cap != cudaStreamCaptureStatusNone) {
    if (next_captured[dev] >= (unsigned)CAPTURED) return nullptr;      // 61440 captured launches per process and device
    return pool[dev] + 2 * ((size_t)EAGER + next_captured[dev]++);
  }
  return pool[dev] + 2 * (size_t)(next_eager[dev]++ % EAGER);
}

// tensor map over one plane of a split-planar activation tensor; box = one tile's patch of one 16-channel chunk
static int make_plane_map(CUtensorMap *m, const void *plane, int64_t N, int64_t H, int64_t W, int64_t C, int pair) {
  EncodeTiledFn enc = encode_tiled();
  MVP_REQUIRE(enc != nullptr, MVP_ERR_UNSUPPORTED, "tc_conv3x3: cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t C8 = (cuuint64_t)(C / 8);
  CUresult r;
  // the 8 channels of a slab and the pixels of an image row are contiguous: one dimension, so that a box row is one
  // 160-byte request (a 16-byte innermost dimension costs one TMA request per pixel and starves the tensor pipe)
  if (!pair) {
    const cuuint64_t dims[3] = {(cuuint64_t)W * 8, (cuuint64_t)H, C8 * (cuuint64_t)N};
    const cuuint64_t strides[2] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16};
    const cuuint32_t box[3] = {HC * 8, 18, 2}, es[3] = {1, 1, 1};
    r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void *>(plane), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    const cuuint64_t dims[4] = {(cuuint64_t)W * 8, 2, (cuuint64_t)H, C8 * (cuuint64_t)((N + 1) / 2)};
    const cuuint64_t strides[3] = {(cuuint64_t)W * 16, (cuuint64_t)2 * W * 16, (cuuint64_t)H * 2 * W * 16};
    const cuuint32_t box[4] = {HC * 8, 2, 10, 2}, es[4] = {1, 1, 1, 1};
    r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(plane), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  MVP_REQUIRE(r == CUDA_SUCCESS, MVP_ERR_INVALID_ARG, "tc_conv3x3: cuTensorMapEncodeTiled failed (%d) for N=%lld H=%lld W=%lld C=%lld", (int)r,
              (long long)N, (long long)H, (long long)W, (long long)C);
  return 0;
}

}  // namespace tcc
}  // namespace mvp

extern "C" int64_t mvp_tc_conv3x3_weight_bytes(int64_t Cin, int64_t Cout) { return Cin * Cout * 9 * 4; }

// width of one block of output channels (GEMM N per accumulator) — part of the packed-weight layout.  256 = the widest
// tcgen05 N: measured against 128 (twice as many, half as long work items, two accumulator sets) the wide layers run
// 20 % slower at 128 (layer3 0.190 vs 0.152 ms) because every patch is then fetched twice as often.
// MVPNET_B200_CONV3_NT overrides (experiments).
extern "C" int64_t mvp_tc_conv3x3_nt(int64_t Cout) {
  static const int64_t cap = [] { const char *e = getenv("MVPNET_B200_CONV3_NT"); const int64_t v = e ? atoll(e) : 0; return (v == 64 || v == 128 || v == 256) ? v : (int64_t)256; }();
  return Cout <= cap ? Cout : cap;
}

// elements (bf16) of a split-planar tensor: both planes
extern "C" int64_t mvp_planar_elems(int64_t N, int64_t H, int64_t W, int64_t C) {
  const int64_t n = H <= 8 ? (N + 1) / 2 * 2 : N;
  return 2 * n * C * H * W;
}

extern "C" int mvp_tc_conv3x3(const void *x1, int64_t C1, const void *x2, int64_t C2, int64_t N, int64_t H, int64_t W,
                              const void *w_packed, const float *bias, int64_t Cout, const void *residual, int relu,
                              void *out_planar, float *out_nhwc, void *out_rows, mvp_stream_t stream) {
  using namespace mvp;
  MVP_REQUIRE(N >= 0 && H > 0 && W > 0, MVP_ERR_INVALID_ARG, "tc_conv3x3: bad sizes");
  MVP_REQUIRE(C1 > 0 && C1 % 16 == 0 && C2 >= 0 && C2 % 16 == 0, MVP_ERR_UNSUPPORTED, "tc_conv3x3: input channels must be multiples of 16");
  MVP_REQUIRE(Cout > 0 && Cout % 16 == 0 && Cout % mvp_tc_conv3x3_nt(Cout) == 0, MVP_ERR_UNSUPPORTED,
              "tc_conv3x3: output channels must be a multiple of 16, and of the block width mvp_tc_conv3x3_nt() above it");
  MVP_REQUIRE(N * H * W < (1LL << 31), MVP_ERR_UNSUPPORTED, "tc_conv3x3: more than 2^31 pixels");
  if (N == 0) return 0;
  MVP_REQUIRE(x1 && w_packed && bias && (out_planar || out_nhwc || out_rows) && (x2 || C2 == 0), MVP_ERR_NULL, "tc_conv3x3: null pointer");
  MVP_REQUIRE((((uintptr_t)x1 | (uintptr_t)x2 | (uintptr_t)w_packed | (uintptr_t)bias | (uintptr_t)residual | (uintptr_t)out_planar |
                (uintptr_t)out_nhwc | (uintptr_t)out_rows) & 15) == 0, MVP_ERR_INVALID_ARG, "tc_conv3x3: pointers must be 16-byte aligned");
  MVP_REQUIRE((((uintptr_t)out_nhwc | (uintptr_t)out_rows) & 31) == 0, MVP_ERR_INVALID_ARG,
              "tc_conv3x3: the NHWC / row-split outputs must be 32-byte aligned (256-bit stores)");
  tcc::ConvArgs a = {};
  const int pair = H <= 8 ? 1 : 0;
  const int64_t Np = pair ? (N + 1) / 2 * 2 : N;
  if (int rc = tcc::make_plane_map(&a.m1h, x1, N, H, W, C1, pair)) return rc;
  if (int rc = tcc::make_plane_map(&a.m1l, (const __nv_bfloat16 *)x1 + Np * C1 * H * W, N, H, W, C1, pair)) return rc;
  if (C2 > 0) {
    if (int rc = tcc::make_plane_map(&a.m2h, x2, N, H, W, C2, pair)) return rc;
    if (int rc = tcc::make_plane_map(&a.m2l, (const __nv_bfloat16 *)x2 + Np * C2 * H * W, N, H, W, C2, pair)) return rc;
  }
  a.C1 = (int)C1; a.C2 = (int)C2; a.N = (int)N; a.H = (int)H; a.W = (int)W;
  a.wp = (const unsigned char *)w_packed; a.bias = bias; a.res = (const __nv_bfloat16 *)residual;
  a.out_p = (__nv_bfloat16 *)out_planar; a.out_f = out_nhwc;
  a.out_rh = (__nv_bfloat16 *)out_rows; a.out_rl = out_rows ? (__nv_bfloat16 *)out_rows + N * H * W * Cout : nullptr; a.plane_out = Np * Cout * H * W; a.relu = relu;
  a.Cout = (int)Cout; a.Nt = (int)mvp_tc_conv3x3_nt(Cout); a.NB = a.Cout / a.Nt;
  a.ipt = pair ? 2 : 1;
  a.TX = (int)((W + 7) / 8);
  a.TY = pair ? 1 : (int)((H + 15) / 16);
  a.ntiles = pair ? (Np / 2) * a.TX : N * a.TX * a.TY;
  // accumulators: two sets (epilogue overlaps the next group's MMAs) where 2 x TM x Nt columns fit TMEM
  a.TM = a.Nt <= 64 ? 4 : 2;
  a.nacc = 2 * a.TM * a.Nt <= 512 ? 2 : 1;
  // tiles per work item: the largest TM whose makespan (rounds of work items over the SMs x item length) is minimal —
  // few-tile layers (the 8 x 10 level: 160 tiles) balance better with shorter items, at the price of weight traffic
  {
    long long best = -1;
    int best_tm = 1;
    for (int tm = a.TM; tm >= 1; tm >>= 1) {
      const long long works = (a.ntiles + tm - 1) / tm * a.NB, span = (works + sm_count() - 1) / sm_count() * tm;
      if (best < 0 || span * 100 < best * 97) { best = span; best_tm = tm; }   // a smaller TM must buy > 3 %: it multiplies the weight traffic
    }
    a.TM = best_tm;
    a.nacc = 2 * a.TM * a.Nt <= 512 ? 2 : 1;
  }
  a.ngroups = (a.ntiles + a.TM - 1) / a.TM;
  a.nchunks = (int)((C1 + C2) / 16);
  a.tmem_cols = 32;
  while (a.tmem_cols < a.nacc * a.TM * a.Nt) a.tmem_cols <<= 1;
  a.asets = tcc::MAX_ASETS;
  a.stages = tcc::MAX_STAGES;
  // taps per weight stage: the whole 3x3 (9 taps) where two such stages fit shared memory, one filter row otherwise —
  // all MMAs of a stage on one tile accumulate back to back into the same TMEM tile (see the issuer), and every
  // stage costs the issuer one barrier round trip
  a.tps = a.Nt <= 128 ? 9 : 3;
  {
    const char *e = getenv("MVPNET_B200_CONV_DBG");
    a.dbg = e ? atoi(e) : 0;
    const char *st = getenv("MVPNET_B200_CONV_STAGES");
    if (st && atoi(st) >= 2 && atoi(st) <= tcc::MAX_STAGES) a.stages = atoi(st);
    const char *as = getenv("MVPNET_B200_CONV_ASETS");
    if (as && atoi(as) >= 2 && atoi(as) <= tcc::MAX_ASETS) a.asets = atoi(as);
    const char *tp = getenv("MVPNET_B200_CONV_TPS");
    if (tp && (atoi(tp) == 1 || atoi(tp) == 3 || atoi(tp) == 9)) a.tps = atoi(tp);
    const char *tm = getenv("MVPNET_B200_CONV_TM");
    if (tm && atoi(tm) >= 1 && atoi(tm) <= tcc::MAX_TM && atoi(tm) * a.Nt <= 512) {
      a.TM = atoi(tm);
      a.nacc = 2 * a.TM * a.Nt <= 512 ? 2 : 1;
      a.ngroups = (a.ntiles + a.TM - 1) / a.TM;
      a.tmem_cols = 32;
      while (a.tmem_cols < a.nacc * a.TM * a.Nt) a.tmem_cols <<= 1;
    }
  }
  auto smem_of = [&]() { return (size_t)a.asets * a.TM * tcc::SLOT_BYTES + (size_t)a.stages * a.tps * 64 * a.Nt + 512 + (size_t)a.Cout * 4; };
  while (a.stages > 3 && smem_of() > tcc::conv_smem_cap()) --a.stages;
  while (a.asets > 2 && smem_of() > tcc::conv_smem_cap()) --a.asets;
  while (a.stages > 2 && smem_of() > tcc::conv_smem_cap()) --a.stages;
  if (smem_of() > tcc::conv_smem_cap() && a.tps > 1) { a.tps = a.tps == 9 ? 3 : 1; a.stages = tcc::MAX_STAGES; while (a.stages > 2 && smem_of() > tcc::conv_smem_cap()) --a.stages; }
  const size_t smem = smem_of();
  MVP_REQUIRE(smem <= tcc::conv_smem_cap(), MVP_ERR_UNSUPPORTED, "tc_conv3x3: shared memory budget exceeded");
  cudaError_t e = cudaFuncSetAttribute(tcc::tc_conv3x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("tc_conv3x3: smem attribute (%zu B): %s", smem, cudaGetErrorString(e)); return (int)e; }
  const long long nworks = a.ngroups * a.NB;
  MVP_REQUIRE(nworks < (1LL << 30), MVP_ERR_UNSUPPORTED, "tc_conv3x3: too many work items");
  a.sched = tcc::sched_pair((cudaStream_t)stream);
  MVP_REQUIRE(a.sched != nullptr, MVP_ERR_UNSUPPORTED,
              "tc_conv3x3: no scheduler counters (first call inside a stream capture, or more than 61440 captured launches)");
  long long grid = sm_count();              // persistent, one CTA per SM
  if (grid > nworks) grid = nworks;
  static const bool debug = getenv("MVPNET_B200_DEBUG") != nullptr;
  if (debug)
    fprintf(stderr, "[tc_conv3x3] N=%d H=%d W=%d Cin=%d+%d Cout=%d Nt=%d ipt=%d tiles=%lld TM=%d nacc=%d works=%lld asets=%d stages=%d tps=%d smem=%zu tmem=%d\n",
            a.N, a.H, a.W, a.C1, a.C2, a.Cout, a.Nt, a.ipt, a.ntiles, a.TM, a.nacc, nworks, a.asets, a.stages, a.tps, smem, a.tmem_cols);
  tcc::tc_conv3x3_kernel<<<(unsigned)grid, tcc::THREADS, smem, (cudaStream_t)stream>>>(a);
  return launch_status("tc_conv3x3");
}

extern "C" int mvp_split_planar(const float *nhwc, int64_t N, int64_t H, int64_t W, int64_t C, void *planar, mvp_stream_t stream) {
  using namespace mvp;
  MVP_REQUIRE(C > 0 && C % 8 == 0 && N >= 0 && H > 0 && W > 0, MVP_ERR_INVALID_ARG, "split_planar: bad sizes (C must be a multiple of 8)");
  if (N == 0) return 0;
  MVP_REQUIRE(nhwc && planar, MVP_ERR_NULL, "split_planar: null pointer");
  const int pair = H <= 8 ? 1 : 0;
  const int64_t Np = pair ? (N + 1) / 2 * 2 : N, plane = Np * C * H * W;
  if (Np != N) {   // the missing partner image of the last pair reads as zeros
    cudaError_t e = cudaMemsetAsync(planar, 0, (size_t)plane * 4, (cudaStream_t)stream);
    if (e != cudaSuccess) { set_error("split_planar: memset: %s", cudaGetErrorString(e)); return (int)e; }
  }
  const long long total = N * (C / 8) * H * W;
  const unsigned grid = (unsigned)((total + 255) / 256 < (long long)sm_count() * 16 ? (total + 255) / 256 : (long long)sm_count() * 16);
  tcc::split_planar_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(nhwc, (int)N, (int)H, (int)W, (int)C, pair, (__nv_bfloat16 *)planar, plane);
  return launch_status("split_planar");
}

extern "C" int mvp_merge_planar(const void *planar, int64_t N, int64_t H, int64_t W, int64_t C, float *nhwc, mvp_stream_t stream) {
  using namespace mvp;
  MVP_REQUIRE(C > 0 && C % 8 == 0 && N >= 0 && H > 0 && W > 0, MVP_ERR_INVALID_ARG, "merge_planar: bad sizes (C must be a multiple of 8)");
  if (N == 0) return 0;
  MVP_REQUIRE(nhwc && planar, MVP_ERR_NULL, "merge_planar: null pointer");
  const int pair = H <= 8 ? 1 : 0;
  const int64_t Np = pair ? (N + 1) / 2 * 2 : N, plane = Np * C * H * W;
  const long long total = N * (C / 8) * H * W;
  const unsigned grid = (unsigned)((total + 255) / 256 < (long long)sm_count() * 16 ? (total + 255) / 256 : (long long)sm_count() * 16);
  tcc::merge_planar_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)planar, plane, (int)N, (int)H, (int)W, (int)C, pair, nhwc);
  return launch_status("merge_planar");
}
