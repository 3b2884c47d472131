// General tap-staged convolution on tcgen05 for the layers of the 2D network that tc_conv.cu's halo-patch scheme does
// not cover (unet_resnet34.py:9-125): the three stride-2 3x3 convolutions and their 1x1 stride-2 down-sample
// branches (torchvision BasicBlock), the four 2x2 stride-2 transposed convolutions of the decoder (as four 1x1
// convolutions, one per output parity, scattered by the epilogue) and the 7x7 stem (as seven row taps over a
// horizontally unfolded input, see unfold_stem_kernel).  Same arithmetic as tc_conv.cu: split-planar bf16 hi/lo
// activations, three kind::f16 MMAs per K-step into an fp32 TMEM accumulator.
//
//   rows (GEMM M)  16 x 8 pixel tiles of the OUTPUT grid (of the INPUT grid for a transposed convolution)
//   K loop         for every channel chunk (KC = 16..64 channels) and every tap (dy, dx): ONE TMA box load per plane
//                  fetches the 128 input pixels (y*s + dy, x*s + dx) x KC channels — the stride s is the tensor map's
//                  element stride, zero padding its out-of-bounds fill — straight into the UMMA operand layout
//   weights        one ring stage per (chunk, tap), shared by the TM tiles a CTA keeps in flight
#include "common.cuh"

namespace mvp {
namespace tcc {

__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap *map, int c0, int c1, int c2, int c3, int c4, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(bar)
      : "memory");
}

constexpr int G_MAX_TAPS = 9;
constexpr int G_MAX_STAGES = 6;
constexpr int G_EPI_WARPS = 8;                  // two per TMEM lane quarter: alternate 16-column chunks (as tc_conv.cu)
constexpr int G_EPI_THREADS = G_EPI_WARPS * 32, G_THREADS = G_EPI_THREADS + 96;

struct ConvGArgs {
  CUtensorMap mh, ml;               // input planes
  int Cin, N;
  int Hi, Wi;                       // input grid
  int Hr, Wr;                       // row grid (tiles): output grid, or input grid when shuffle
  int Ho, Wo;                       // output grid
  int stride;                       // input pixel = row pixel * stride + tap offset
  int ntaps;
  int tdy[G_MAX_TAPS], tdx[G_MAX_TAPS];
  int in_pair;                      // input planes are pair-interleaved (Hi <= 8)
  int merged;                       // stride 1: the tensor maps fold (channel-in-slab, x) into one dimension (128-byte box rows)
  int shuffle;                      // transposed 2x2/s2: output block nb -> parity, out pixel = 2 * row pixel + parity
  const unsigned char *wp;          // [nb][Cin/16][tap][hi|lo][2][Nt][8] bf16 (the layout of tc_conv.cu with `ntaps` taps)
  const float *bias;                // [Cout]
  __nv_bfloat16 *out_p;             // split-planar output (Ho x Wo, Cout channels)
  long long plane_out;
  int Cout, Nt, NB;                 // NB blocks of Nt GEMM columns (shuffle: 4 * Cout / Nt)
  int relu;
  int KC, nchunks;
  int TX, TY;
  long long ntiles, ngroups;
  int TM, nacc, astages, stages, tmem_cols;
  int resident;                     // the whole weight block lives in the `stages` (= K-loop length) stages: loaded once per CTA
  unsigned int *sched;              // dynamic work distribution counters (tc_conv.cu)
};

__global__ void __launch_bounds__(G_THREADS, 1)
tc_convg_kernel(const __grid_constant__ ConvGArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t a_half = (uint32_t)(a.KC >> 3) * 2048u;         // hi (or lo) of one tile's stage: KC/8 slabs x 128 rows x 16 B
  const uint32_t a_tile = 2u * a_half;
  const uint32_t a_stage = (uint32_t)a.TM * a_tile;
  const uint32_t b_stage = (uint32_t)a.KC * (uint32_t)a.Nt * 4u;
  unsigned char *a_base = smem;
  unsigned char *b_base = a_base + (size_t)a.astages * a_stage;
  uint64_t *bars = reinterpret_cast<uint64_t *>(b_base + (size_t)a.stages * b_stage);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 4 * G_MAX_STAGES + 4);
  volatile int *s_ring = reinterpret_cast<volatile int *>(bars + 30);    // [SCHED_DEPTH] work indices
  const uint32_t bar_sfull = smem_u32(bars + 32), bar_sempty = smem_u32(bars + 32 + SCHED_DEPTH);
  float *s_bias = reinterpret_cast<float *>(bars + 48);          // [Cout]
  const uint32_t bar_bfull = smem_u32(bars), bar_bempty = bar_bfull + 8 * G_MAX_STAGES;
  const uint32_t bar_afull = bar_bempty + 8 * G_MAX_STAGES, bar_aempty = bar_afull + 8 * G_MAX_STAGES;
  const uint32_t bar_accfull = bar_aempty + 8 * G_MAX_STAGES, bar_accempty = bar_accfull + 16;

  if (tid == 0) {
    for (int s = 0; s < G_MAX_STAGES; ++s) {
      mbar_init(bar_bfull + 8 * s, 1); mbar_init(bar_bempty + 8 * s, 1);
      mbar_init(bar_afull + 8 * s, 1); mbar_init(bar_aempty + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) { mbar_init(bar_accfull + 8 * s, 1); mbar_init(bar_accempty + 8 * s, G_EPI_THREADS); }
    for (int s = 0; s < SCHED_DEPTH; ++s) { mbar_init(bar_sfull + 8 * s, 1); mbar_init(bar_sempty + 8 * s, 2 + G_EPI_THREADS / 32); }
    fence_barrier_init();
  }
  if (warp == G_EPI_WARPS) tmem_alloc(smem_u32(tmem_slot), (uint32_t)a.tmem_cols);
  for (int i = tid; i < a.Cout; i += G_THREADS) s_bias[i] = a.bias[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int ngroups = (int)a.ngroups, nworks = ngroups * a.NB;
  const uint32_t S = (uint32_t)a.stages, AS = (uint32_t)a.astages, NA = (uint32_t)a.nacc;
  const int per_img = a.TX * a.TY;
  const int iters = a.nchunks * a.ntaps;                         // K-loop length of one work item

  if (warp < G_EPI_WARPS) {
    // =========================== epilogue: thread = TMEM lane = pixel; the two warps of a lane quarter take alternate
    //                             16-column chunks (the transposed convolutions are epilogue-paced: 256 columns per K = 64) ===
    const int quarter = warp & 3, c0 = (warp >> 2) * 16;
    const int row = quarter * 32 + lane, g = row >> 3, xx = row & 7;
    const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const int C8o = a.Cout >> 3, out_pair = a.Ho <= 8;
    const size_t slab_stride = (size_t)a.Ho * a.Wo * 8 * (out_pair ? 2 : 1);
    const bool paired = a.shuffle && 2 * a.Cout <= a.Nt;         // a block holds whole (kx = 0, kx = 1) parity pairs
    const int cpc = a.Cout >> 4, npair = a.Nt >> 5;              // 16-column chunks per parity; paired items per tile
    for (uint32_t it = 0;; ++it) {
      const int w = sched_consume(it, s_ring, bar_sfull, bar_sempty);
      if (w >= nworks) break;
      const int nb = w / ngroups;
      const long long group = w - nb * ngroups;
      const uint32_t set = it % NA;
      mbar_wait(bar_accfull + 8 * set, (it / NA) & 1u);
      tc_fence_after();
      for (int t = 0; t < a.TM; ++t) {
        const long long tile = group * a.TM + t;
        if (tile >= a.ntiles) break;
        const int n = (int)(tile / per_img), r = (int)(tile - (long long)n * per_img);
        const int yr = (r / a.TX) * 16 + g, xr = (r % a.TX) * 8 + xx;
        const bool ok = yr < a.Hr && xr < a.Wr;
        const uint32_t t_acc = t_lane + (uint32_t)((set * a.TM + t) * a.Nt);
        if (paired) {
          // transposed convolution, both x parities of an output row in this block: an item = 16 channels of one y
          // parity, the two x parities side by side -> 32 contiguous bytes per (slab, plane), one STG.256 each
          for (int j = c0 >> 4; j < npair; j += 2) {
            const int kyl = j / cpc, cc = j - kyl * cpc;
            const int col = kyl * 2 * a.Cout + cc * 16, gcol = nb * a.Nt + col;
            const int par = gcol / a.Cout, co = gcol - par * a.Cout;           // par even: (ky, kx = 0)
            uint32_t ra[16], rb[16];
            tmem_ld16_issue(t_acc + (uint32_t)col, ra);
            tmem_ld16_issue(t_acc + (uint32_t)(col + a.Cout), rb);
            const size_t pbase = ok ? planar_off(n, co >> 3, 2 * yr + (par >> 1), 2 * xr, C8o, a.Ho, a.Wo, out_pair) : 0;
            float bq[16];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 b4 = reinterpret_cast<const float4 *>(s_bias + co)[q];
              bq[4 * q] = b4.x; bq[4 * q + 1] = b4.y; bq[4 * q + 2] = b4.z; bq[4 * q + 3] = b4.w;
            }
            tmem_ld_wait(ra);
            tmem_ld_wait(rb);
            float va[16], vb[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) {
              va[q] = __uint_as_float(ra[q]) + bq[q];
              vb[q] = __uint_as_float(rb[q]) + bq[q];
              if (a.relu) { va[q] = fmaxf(va[q], 0.f); vb[q] = fmaxf(vb[q], 0.f); }
            }
            if (ok) {
#pragma unroll
              for (int s = 0; s < 2; ++s) {
                uint32_t ha[4], la[4], hb[4], lb[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  split_pair(va[8 * s + 2 * q], va[8 * s + 2 * q + 1], ha[q], la[q]);
                  split_pair(vb[8 * s + 2 * q], vb[8 * s + 2 * q + 1], hb[q], lb[q]);
                }
                const size_t o = pbase + (size_t)s * slab_stride;
                st_global_256(a.out_p + o, make_uint4(ha[0], ha[1], ha[2], ha[3]), make_uint4(hb[0], hb[1], hb[2], hb[3]));
                st_global_256(a.out_p + a.plane_out + o, make_uint4(la[0], la[1], la[2], la[3]), make_uint4(lb[0], lb[1], lb[2], lb[3]));
              }
            }
          }
          continue;
        }
        uint32_t rn[16];
        if (c0 < a.Nt) tmem_ld16_issue(t_acc + (uint32_t)c0, rn);
        for (int c = c0; c < a.Nt; c += 32) {
          float v[16];
          tmem_ld_wait(rn);
#pragma unroll
          for (int q = 0; q < 16; ++q) v[q] = __uint_as_float(rn[q]);
          if (c + 32 < a.Nt) tmem_ld16_issue(t_acc + (uint32_t)(c + 32), rn);
          // GEMM column -> (parity, output channel): a 16-column chunk never straddles a parity (Cout % 16 == 0)
          const int col = nb * a.Nt + c;
          const int par = a.shuffle ? col / a.Cout : 0, co = a.shuffle ? col - par * a.Cout : col;
          const int yo = a.shuffle ? 2 * yr + (par >> 1) : yr, xo = a.shuffle ? 2 * xr + (par & 1) : xr;
          const size_t pbase = ok ? planar_off(n, co >> 3, yo, xo, C8o, a.Ho, a.Wo, out_pair) : 0;
          const float4 *bp = reinterpret_cast<const float4 *>(s_bias + co);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 bq = bp[q];
            v[4 * q] += bq.x; v[4 * q + 1] += bq.y; v[4 * q + 2] += bq.z; v[4 * q + 3] += bq.w;
          }
          if (a.relu) {
#pragma unroll
            for (int q = 0; q < 16; ++q) v[q] = fmaxf(v[q], 0.f);
          }
          if (ok) {
#pragma unroll
            for (int s = 0; s < 2; ++s) {
              uint32_t h[4], l[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) split_pair(v[8 * s + 2 * q], v[8 * s + 2 * q + 1], h[q], l[q]);
              const size_t o = pbase + (size_t)s * slab_stride;
              *reinterpret_cast<uint4 *>(a.out_p + o) = make_uint4(h[0], h[1], h[2], h[3]);
              *reinterpret_cast<uint4 *>(a.out_p + a.plane_out + o) = make_uint4(l[0], l[1], l[2], l[3]);
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(bar_accempty + 8 * set);
    }
  } else if (warp == G_EPI_WARPS) {
    // =========================== MMA issuer (whole warp walks, one elected lane issues) ==============================
    // running ring positions / phases and 32-bit descriptor arithmetic only: see tc_conv.cu
    const uint32_t idesc = make_idesc(128, a.Nt);
    const uint64_t adesc0 = make_desc(0, 2048, 128), bdesc0 = make_desc(0, (uint32_t)a.Nt * 16u, 128);
    const uint32_t a_lo0 = (uint32_t)adesc0 + (smem_u32(a_base) >> 4), a_hi32 = (uint32_t)(adesc0 >> 32);
    const uint32_t b_lo0 = (uint32_t)bdesc0 + (smem_u32(b_base) >> 4), b_hi32 = (uint32_t)(bdesc0 >> 32);
    const uint32_t a_stage16 = a_stage >> 4, a_tile16 = a_tile >> 4, a_half16 = a_half >> 4, b_stage16 = b_stage >> 4;
    const uint32_t b_piece16 = 4u * (uint32_t)a.Nt, b_half16 = b_piece16 >> 1;   // one 16-channel piece: hi | lo
    const int ksteps = a.KC >> 4;
    auto desc64 = [](uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; };
    uint32_t sa = 0, a_ph = 0, sb = 0, b_ph = 0, set = 0, acc_ph = 0;
    for (uint32_t w_it = 0;; ++w_it) {
      const int w = sched_consume(w_it, s_ring, bar_sfull, bar_sempty);
      if (w >= nworks) break;
      const long long left = a.ntiles - (long long)(w % ngroups) * a.TM;
      const int nt = left < a.TM ? (int)left : a.TM;
      if (w_it >= NA) mbar_wait(bar_accempty + 8 * set, acc_ph ^ 1u);
      tc_fence_after();
      const uint32_t d0 = tmem_base + set * (uint32_t)(a.TM * a.Nt);
      uint32_t first = 0u;
      for (int i = 0; i < iters; ++i) {
        mbar_wait(bar_afull + 8 * sa, a_ph);
        mbar_wait(bar_bfull + 8 * sb, a.resident ? 0u : b_ph);       // resident: phase 0 completes once and stays complete
        const uint32_t a_lo = a_lo0 + sa * a_stage16, b_lo = b_lo0 + sb * b_stage16;
        if (elect_one()) {
#pragma unroll
          for (int t = 0; t < MAX_TM; ++t) {
            if (t < nt) {
              const uint32_t d = d0 + (uint32_t)(t * a.Nt);
              uint32_t al = a_lo + (uint32_t)t * a_tile16, bl = b_lo, acc = first;
              for (int k = 0; k < ksteps; ++k) {
                const uint64_t ah = desc64(al, a_hi32), alo = desc64(al + a_half16, a_hi32);
                const uint64_t bh = desc64(bl, b_hi32), blo = desc64(bl + b_half16, b_hi32);
                umma_bf16(d, ah, bh, idesc, acc);
                umma_bf16(d, ah, blo, idesc, 1u);
                umma_bf16(d, alo, bh, idesc, 1u);
                al += 256u;                  // two 2048-byte slabs
                bl += b_piece16;
                acc = 1u;
              }
            }
          }
          if (!a.resident) umma_commit(bar_bempty + 8 * sb);
          umma_commit(bar_aempty + 8 * sa);
          if (i == iters - 1) umma_commit(bar_accfull + 8 * set);
        }
        __syncwarp();
        first = 1u;
        if (++sa == AS) { sa = 0; a_ph ^= 1u; }
        if (++sb == S) { sb = 0; b_ph ^= 1u; }
      }
      if (++set == NA) { set = 0; acc_ph ^= 1u; }
    }
  } else if (warp == G_EPI_WARPS + 1) {
    // =========================== input producer: one TMA box per (tile, chunk, tap, plane) ===========================
    const uint32_t a_s = smem_u32(a_base);
    const int C8 = a.Cin >> 3, kslabs = a.KC >> 3;
    const int xs = a.merged ? 8 : 1;                        // merged maps address x in elements of the (8 * W) dimension
    uint32_t sa = 0, ph = 0, it = 0;
    for (uint32_t k = 0;; ++k) {
      const int w = sched_produce(k, s_ring, bar_sfull, bar_sempty, a.sched);
      if (w >= nworks) break;
      const int group = w % ngroups;
      const long long left = a.ntiles - (long long)group * a.TM;
      const int nt = left < a.TM ? (int)left : a.TM;
      int cx[MAX_TM], cy[MAX_TM], cn[MAX_TM];               // tile origins in the input grid: once per work item
#pragma unroll
      for (int t = 0; t < MAX_TM; ++t) {
        const long long tile = (long long)group * a.TM + (t < nt ? t : 0);
        const int n = (int)(tile / per_img), r = (int)(tile - (long long)n * per_img);
        cy[t] = (r / a.TX) * 16 * a.stride; cx[t] = (r % a.TX) * 8 * a.stride; cn[t] = n;
      }
      for (int c = 0; c < a.nchunks; ++c) {
        for (int tap = 0; tap < a.ntaps; ++tap, ++it) {
          if (it >= AS) mbar_wait(bar_aempty + 8 * sa, ph ^ 1u);
          const uint32_t bar = bar_afull + 8 * sa, dst0 = a_s + sa * a_stage;
          const int dy = a.tdy[tap], dx = a.tdx[tap];
          if (elect_one()) {
            mbar_expect_tx(bar, (uint32_t)nt * a_tile);
#pragma unroll
            for (int t = 0; t < MAX_TM; ++t) {
              if (t < nt) {
                const uint32_t dst = dst0 + (uint32_t)t * a_tile;
                const int x = (cx[t] + dx) * xs, y = cy[t] + dy, n = cn[t];
                if (a.merged && !a.in_pair) {         // dims (8 * W, H, C8 * N)
                  tma_load_3d(dst, &a.mh, x, y, n * C8 + c * kslabs, bar);
                  tma_load_3d(dst + a_half, &a.ml, x, y, n * C8 + c * kslabs, bar);
                } else if (a.merged) {                // dims (8 * W, 2, H, C8 * N/2)
                  tma_load_4d(dst, &a.mh, x, n & 1, y, (n >> 1) * C8 + c * kslabs, bar);
                  tma_load_4d(dst + a_half, &a.ml, x, n & 1, y, (n >> 1) * C8 + c * kslabs, bar);
                } else if (!a.in_pair) {              // dims (8, W, H, C8 * N), traversal stride on W and H
                  tma_load_4d(dst, &a.mh, 0, x, y, n * C8 + c * kslabs, bar);
                  tma_load_4d(dst + a_half, &a.ml, 0, x, y, n * C8 + c * kslabs, bar);
                } else {                              // dims (8, W, 2, H, C8 * N/2)
                  tma_load_5d(dst, &a.mh, 0, x, n & 1, y, (n >> 1) * C8 + c * kslabs, bar);
                  tma_load_5d(dst + a_half, &a.ml, 0, x, n & 1, y, (n >> 1) * C8 + c * kslabs, bar);
                }
              }
            }
          }
          __syncwarp();
          if (++sa == AS) { sa = 0; ph ^= 1u; }
        }
      }
    }
    sched_retire(a.sched);
  } else {
    // =========================== weight producer: KC/16 pieces per stage ============================================
    const uint32_t b_s = smem_u32(b_base);
    const uint32_t b_piece = 64u * (uint32_t)a.Nt;
    const int ks = a.KC >> 4, nch16 = a.Cin >> 4;
    uint32_t s = 0, ph = 0, it = 0;
    for (uint32_t k = 0;; ++k) {
      const int w = sched_consume(k, s_ring, bar_sfull, bar_sempty);
      if (w >= nworks) break;
      const int nb = w / ngroups;
      const unsigned char *wsrc = a.wp + (size_t)nb * nch16 * a.ntaps * b_piece;
      if (a.resident && k > 0) continue;                    // every stage was filled for the first work item and is never released
      for (int c = 0; c < a.nchunks; ++c) {
        for (int tap = 0; tap < a.ntaps; ++tap, ++it) {
          if (it >= S) mbar_wait(bar_bempty + 8 * s, ph ^ 1u);
          if (elect_one()) {
            mbar_expect_tx(bar_bfull + 8 * s, b_stage);
            for (int kk = 0; kk < ks; ++kk)
              bulk_g2s(b_s + s * b_stage + (uint32_t)kk * b_piece, wsrc + ((size_t)(c * ks + kk) * a.ntaps + tap) * b_piece, b_piece, bar_bfull + 8 * s);
          }
          __syncwarp();
          if (++s == S) { s = 0; ph ^= 1u; }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == G_EPI_WARPS) tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
}

// ---- stem helper: fp32 NCHW image (3 channels) -> split-planar "row-unfolded" tensor with 32 channels per pixel:
// channel kx*3 + c = image[n, c, y, x + kx - 3] (zero outside), channels 21..31 zero.  The 7x7 convolution is then
// seven row taps (dy = -3..3, dx = 0) over this tensor (unet_resnet34.py:16-21 encoder0).
__global__ void __launch_bounds__(256)
unfold_stem_kernel(const float *__restrict__ img, int N, int H, int W, __nv_bfloat16 *__restrict__ dst, long long plane) {
  // thread = pixel (n, y, x), x fastest: 21 coalesced loads (3 channel planes x 7 consecutive columns), all four slabs
  // written by the same thread (each slab plane is contiguous along x: 16-byte stores of consecutive lanes)
  const long long total = (long long)N * H * W;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int x = (int)(i % W);
    const long long r = i / W;
    const int y = (int)(r % H), n = (int)(r / H);
    float v[32];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float *row = img + (((size_t)n * 3 + c) * H + y) * W;
#pragma unroll
      for (int kx = 0; kx < 7; ++kx) {
        const int xs = x + kx - 3;
        v[kx * 3 + c] = (xs >= 0 && xs < W) ? __ldg(row + xs) : 0.f;
      }
    }
#pragma unroll
    for (int e = 21; e < 32; ++e) v[e] = 0.f;
#pragma unroll
    for (int c8 = 0; c8 < 4; ++c8) {
      uint32_t h[4], l[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) split_pair(v[c8 * 8 + 2 * q], v[c8 * 8 + 2 * q + 1], h[q], l[q]);
      const size_t o = planar_off(n, c8, y, x, 4, H, W, 0);
      *reinterpret_cast<uint4 *>(dst + o) = make_uint4(h[0], h[1], h[2], h[3]);
      *reinterpret_cast<uint4 *>(dst + plane + o) = make_uint4(l[0], l[1], l[2], l[3]);
    }
  }
}

// ---- 3x3 / stride 2 / pad 1 max-pool on split-planar tensors (unet_resnet34.py:88 maxpool): thread = (n, slab, yo, xo)
__global__ void __launch_bounds__(256)
maxpool_planar_kernel(const __nv_bfloat16 *__restrict__ src, long long plane_in, int N, int H, int W, int C, int Ho, int Wo,
                      __nv_bfloat16 *__restrict__ dst, long long plane_out) {
  const int C8 = C >> 3, in_pair = H <= 8, out_pair = Ho <= 8;
  const long long total = (long long)N * C8 * Ho * Wo;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int xo = (int)(i % Wo);
    long long r = i / Wo;
    const int yo = (int)(r % Ho); r /= Ho;
    const int c8 = (int)(r % C8);
    const int n = (int)(r / C8);
    float m[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) m[e] = -__int_as_float(0x7f800000);
    for (int dy = -1; dy <= 1; ++dy) {
      const int y = 2 * yo + dy;
      if (y < 0 || y >= H) continue;
      for (int dx = -1; dx <= 1; ++dx) {
        const int x = 2 * xo + dx;
        if (x < 0 || x >= W) continue;
        const size_t o = planar_off(n, c8, y, x, C8, H, W, in_pair);
        float v[8];
        unpack8(__ldg(reinterpret_cast<const uint4 *>(src + o)), __ldg(reinterpret_cast<const uint4 *>(src + plane_in + o)), v);
#pragma unroll
        for (int e = 0; e < 8; ++e) m[e] = fmaxf(m[e], v[e]);
      }
    }
    uint32_t h[4], l[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) split_pair(m[2 * q], m[2 * q + 1], h[q], l[q]);
    const size_t o = planar_off(n, c8, yo, xo, C8, Ho, Wo, out_pair);
    *reinterpret_cast<uint4 *>(dst + o) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4 *>(dst + plane_out + o) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

// tensor map over one plane; box = the 16 x 8 pixels x `kslabs` slabs of one tile.  stride 1: (channel-in-slab, x) folded
// into one dimension, so a box row is one 128-byte request; stride 2: unmerged dimensions with a traversal stride.
static int make_plane_map_g(CUtensorMap *m, const void *plane, int64_t N, int64_t H, int64_t W, int64_t C, int pair, int stride, int kslabs) {
  EncodeTiledFn enc = encode_tiled();
  MVP_REQUIRE(enc != nullptr, MVP_ERR_UNSUPPORTED, "tc_conv: cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t C8 = (cuuint64_t)(C / 8), NP = (cuuint64_t)((N + 1) / 2);
  const cuuint32_t st = (cuuint32_t)stride, ks = (cuuint32_t)kslabs;
  const cuuint64_t w16 = (cuuint64_t)W * 16, hw16 = (cuuint64_t)H * W * 16;
  CUresult r;
  if (stride == 1 && !pair) {
    const cuuint64_t dims[3] = {(cuuint64_t)W * 8, (cuuint64_t)H, C8 * (cuuint64_t)N}, strides[2] = {w16, hw16};
    const cuuint32_t box[3] = {64, 16, ks}, es[3] = {1, 1, 1};
    r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void *>(plane), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else if (stride == 1) {
    const cuuint64_t dims[4] = {(cuuint64_t)W * 8, 2, (cuuint64_t)H, C8 * NP}, strides[3] = {w16, 2 * w16, 2 * hw16};
    const cuuint32_t box[4] = {64, 1, 16, ks}, es[4] = {1, 1, 1, 1};
    r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(plane), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else if (!pair) {
    const cuuint64_t dims[4] = {8, (cuuint64_t)W, (cuuint64_t)H, C8 * (cuuint64_t)N}, strides[3] = {16, w16, hw16};
    const cuuint32_t box[4] = {8, 8 * st, 16 * st, ks}, es[4] = {1, st, st, 1};
    r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(plane), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    const cuuint64_t dims[5] = {8, (cuuint64_t)W, 2, (cuuint64_t)H, C8 * NP}, strides[4] = {16, w16, 2 * w16, 2 * hw16};
    const cuuint32_t box[5] = {8, 8 * st, 1, 16 * st, ks}, es[5] = {1, st, 1, st, 1};
    r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void *>(plane), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  MVP_REQUIRE(r == CUDA_SUCCESS, MVP_ERR_INVALID_ARG, "tc_conv: cuTensorMapEncodeTiled failed (%d) for N=%lld H=%lld W=%lld C=%lld stride=%d", (int)r,
              (long long)N, (long long)H, (long long)W, (long long)C, stride);
  return 0;
}

}  // namespace tcc
}  // namespace mvp

// mode 0: convolution with `ntaps` taps (dy[i], dx[i]) and stride `stride` (output grid Ho x Wo given by the caller);
// mode 1: 2x2 / stride-2 transposed convolution (ntaps must be 1, tap (0,0); GEMM columns = 4 parities x Cout)
extern "C" int mvp_tc_conv_general(const void *x, int64_t Cin, int64_t N, int64_t Hi, int64_t Wi, int mode, int stride, int ntaps,
                                   const int *dy, const int *dx, int64_t Ho, int64_t Wo, const void *w_packed, const float *bias,
                                   int64_t Cout, int relu, void *out_planar, mvp_stream_t stream) {
  using namespace mvp;
  MVP_REQUIRE(N >= 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0, MVP_ERR_INVALID_ARG, "tc_conv_general: bad sizes");
  MVP_REQUIRE(Cin > 0 && Cin % 16 == 0 && Cout > 0 && Cout % 16 == 0 && (Cout <= 256 || Cout % 256 == 0), MVP_ERR_UNSUPPORTED,
              "tc_conv_general: channels must be multiples of 16 (Cout a multiple of 256 above 256)");
  MVP_REQUIRE((mode == 0 || mode == 1) && (stride == 1 || stride == 2) && ntaps >= 1 && ntaps <= tcc::G_MAX_TAPS && dy && dx, MVP_ERR_INVALID_ARG,
              "tc_conv_general: bad mode / stride / taps");
  MVP_REQUIRE(mode == 0 || (ntaps == 1 && stride == 1 && Ho == 2 * Hi && Wo == 2 * Wi), MVP_ERR_INVALID_ARG,
              "tc_conv_general: a transposed convolution has one tap and doubles the grid");
  MVP_REQUIRE(mode == 0 || 4 * Cout <= 256 || (4 * Cout) % 256 == 0, MVP_ERR_UNSUPPORTED,
              "tc_conv_general: 4 x Cout of a transposed convolution must be <= 256 or a multiple of 256");
  MVP_REQUIRE(N * Ho * Wo < (1LL << 31) && N * Hi * Wi < (1LL << 31), MVP_ERR_UNSUPPORTED, "tc_conv_general: more than 2^31 pixels");
  if (N == 0) return 0;
  MVP_REQUIRE(x && w_packed && bias && out_planar, MVP_ERR_NULL, "tc_conv_general: null pointer");
  MVP_REQUIRE(((uintptr_t)out_planar & 31) == 0, MVP_ERR_INVALID_ARG, "tc_conv_general: the output must be 32-byte aligned (256-bit stores)");
  tcc::ConvGArgs a = {};
  a.Cin = (int)Cin; a.N = (int)N; a.Hi = (int)Hi; a.Wi = (int)Wi; a.Ho = (int)Ho; a.Wo = (int)Wo;
  a.shuffle = mode; a.stride = stride; a.ntaps = ntaps;
  for (int i = 0; i < ntaps; ++i) { a.tdy[i] = dy[i]; a.tdx[i] = dx[i]; }
  a.Hr = mode ? (int)Hi : (int)Ho; a.Wr = mode ? (int)Wi : (int)Wo;
  a.in_pair = Hi <= 8 ? 1 : 0;
  const int64_t Np_in = a.in_pair ? (N + 1) / 2 * 2 : N, Np_out = Ho <= 8 ? (N + 1) / 2 * 2 : N;
  a.wp = (const unsigned char *)w_packed; a.bias = bias; a.out_p = (__nv_bfloat16 *)out_planar; a.plane_out = Np_out * Cout * Ho * Wo;
  a.relu = relu; a.Cout = (int)Cout;
  const int64_t G = mode ? 4 * Cout : Cout;                  // GEMM columns: a transposed convolution carries its 4 parities side by side
  a.Nt = G <= 256 ? (int)G : 256; a.NB = (int)(G / a.Nt);
  a.merged = stride == 1 ? 1 : 0;
  a.TX = (a.Wr + 7) / 8; a.TY = (a.Hr + 15) / 16;
  a.ntiles = N * a.TX * a.TY;
  // tiles per work item (share every weight stage): one tile per item was measured 1.7x slower on the stem — the
  // weights are then re-fetched from L2 for every tile
  a.TM = a.Nt <= 64 ? 4 : 2;
  while (a.TM > 1 && (a.ntiles + a.TM - 1) / a.TM * a.NB < sm_count()) a.TM >>= 1;
  // Wide single-block layers whose whole weight block fits next to three input stages (the last two transposed
  // convolutions: K = 64 / 128, 256 GEMM columns) keep the weights RESIDENT and take one tile per work item: two
  // accumulator sets then fit in TMEM and the epilogue — 256 columns per pixel for K = 64, the pacing item — overlaps
  // the next tile's loads and MMAs (with TM = 2 the single accumulator set serialised load -> MMA -> epilogue).
  {
    const int64_t kc = Cin % 64 == 0 ? 64 : (Cin % 32 == 0 ? 32 : 16), it = Cin / kc * ntaps;
    static const bool allow = [] { const char *e = getenv("MVPNET_B200_CONVG_RESIDENT"); return !(e && e[0] == '0'); }();
    a.resident = allow && a.NB == 1 && a.Nt > 128 && it <= tcc::G_MAX_STAGES &&
                 (size_t)it * kc * a.Nt * 4 + 2 * (size_t)kc * 512 + 1024 + Cout * 4 <= tcc::conv_smem_cap();
    if (a.resident) a.TM = 1;
  }
  a.nacc = 2 * a.TM * a.Nt <= 512 ? 2 : 1;
  a.ngroups = (a.ntiles + a.TM - 1) / a.TM;
  a.tmem_cols = 32;
  while (a.tmem_cols < a.nacc * a.TM * a.Nt) a.tmem_cols <<= 1;
  // channels per stage: as many as keep three input stages (TM tiles x KC channels x 512 B) + three weight stages
  // (KC x Nt x 4 B) in shared memory — every stage costs the issuer a barrier round trip
  a.KC = 64;
  if (a.resident) a.KC = Cin % 64 == 0 ? 64 : (Cin % 32 == 0 ? 32 : 16);
  while (!a.resident && a.KC > 16 && (Cin % a.KC != 0 || 3 * ((size_t)a.TM * a.KC * 512 + (size_t)a.KC * a.Nt * 4) + 1024 + Cout * 4 > tcc::conv_smem_cap())) a.KC >>= 1;
  a.nchunks = (int)(Cin / a.KC);
  if (int rc = tcc::make_plane_map_g(&a.mh, x, N, Hi, Wi, Cin, a.in_pair, stride, a.KC / 8)) return rc;
  if (int rc = tcc::make_plane_map_g(&a.ml, (const __nv_bfloat16 *)x + Np_in * Cin * Hi * Wi, N, Hi, Wi, Cin, a.in_pair, stride, a.KC / 8)) return rc;
  const size_t a_stage = (size_t)a.TM * a.KC * 512, b_stage = (size_t)a.KC * a.Nt * 4;
  a.astages = a.stages = tcc::G_MAX_STAGES;
  if (a.resident) a.stages = a.nchunks * ntaps;
  auto smem_of = [&]() { return a.astages * a_stage + a.stages * b_stage + 512 + (size_t)a.Cout * 4; };
  while (smem_of() > tcc::conv_smem_cap() && (a.astages > 2 || a.stages > 2)) {
    if (a.resident) { --a.astages; continue; }
    if (a.astages >= a.stages && a.astages > 2) --a.astages; else if (a.stages > 2) --a.stages; else --a.astages;
  }
  const size_t smem = smem_of();
  MVP_REQUIRE(smem <= tcc::conv_smem_cap(), MVP_ERR_UNSUPPORTED, "tc_conv_general: shared memory budget exceeded");
  cudaError_t e = cudaFuncSetAttribute(tcc::tc_convg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("tc_conv_general: smem attribute (%zu B): %s", smem, cudaGetErrorString(e)); return (int)e; }
  const long long nworks = a.ngroups * a.NB;
  MVP_REQUIRE(nworks < (1LL << 30), MVP_ERR_UNSUPPORTED, "tc_conv_general: too many work items");
  a.sched = tcc::sched_pair((cudaStream_t)stream);
  MVP_REQUIRE(a.sched != nullptr, MVP_ERR_UNSUPPORTED,
              "tc_conv_general: no scheduler counters (first call inside a stream capture, or more than 61440 captured launches)");
  long long grid = sm_count();
  if (grid > nworks) grid = nworks;
  static const bool debug = getenv("MVPNET_B200_DEBUG") != nullptr;
  if (debug)
    fprintf(stderr, "[tc_conv_general] N=%d in=%dx%d out=%dx%d Cin=%d Cout=%d mode=%d stride=%d taps=%d Nt=%d NB=%d KC=%d tiles=%lld TM=%d nacc=%d works=%lld astages=%d stages=%d smem=%zu\n",
            a.N, a.Hi, a.Wi, a.Ho, a.Wo, a.Cin, a.Cout, mode, stride, ntaps, a.Nt, a.NB, a.KC, a.ntiles, a.TM, a.nacc, nworks, a.astages, a.stages, smem);
  tcc::tc_convg_kernel<<<(unsigned)grid, tcc::G_THREADS, smem, (cudaStream_t)stream>>>(a);
  return launch_status("tc_conv_general");
}

extern "C" int mvp_unfold_stem(const float *image_nchw, int64_t N, int64_t H, int64_t W, void *planar32, mvp_stream_t stream) {
  using namespace mvp;
  MVP_REQUIRE(N >= 0 && H > 8 && W > 0, MVP_ERR_INVALID_ARG, "unfold_stem: bad sizes (H must exceed 8)");
  if (N == 0) return 0;
  MVP_REQUIRE(image_nchw && planar32, MVP_ERR_NULL, "unfold_stem: null pointer");
  const long long total = N * H * W;
  const unsigned grid = (unsigned)((total + 255) / 256 < (long long)sm_count() * 16 ? (total + 255) / 256 : (long long)sm_count() * 16);
  tcc::unfold_stem_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(image_nchw, (int)N, (int)H, (int)W, (__nv_bfloat16 *)planar32, N * 32 * H * W);
  return launch_status("unfold_stem");
}

extern "C" int mvp_maxpool3x3s2_planar(const void *x, int64_t N, int64_t H, int64_t W, int64_t C, void *out, mvp_stream_t stream) {
  using namespace mvp;
  MVP_REQUIRE(N >= 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, MVP_ERR_INVALID_ARG, "maxpool_planar: bad sizes");
  if (N == 0) return 0;
  MVP_REQUIRE(x && out, MVP_ERR_NULL, "maxpool_planar: null pointer");
  const int64_t Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  const int64_t Np_in = H <= 8 ? (N + 1) / 2 * 2 : N, Np_out = Ho <= 8 ? (N + 1) / 2 * 2 : N;
  if (Np_out != N) {
    cudaError_t e = cudaMemsetAsync(out, 0, (size_t)Np_out * C * Ho * Wo * 4, (cudaStream_t)stream);
    if (e != cudaSuccess) { set_error("maxpool_planar: memset: %s", cudaGetErrorString(e)); return (int)e; }
  }
  const long long total = N * (C / 8) * Ho * Wo;
  const unsigned grid = (unsigned)((total + 255) / 256 < (long long)sm_count() * 16 ? (total + 255) / 256 : (long long)sm_count() * 16);
  tcc::maxpool_planar_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)x, Np_in * C * H * W, (int)N, (int)H, (int)W, (int)C,
                                                                    (int)Ho, (int)Wo, (__nv_bfloat16 *)out, Np_out * C * Ho * Wo);
  return launch_status("maxpool_planar");
}
