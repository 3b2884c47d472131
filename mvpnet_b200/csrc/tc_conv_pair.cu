// CTA-pair variant of the 3x3 convolution of tc_conv.cu (every level above 8 image rows): tcgen05.mma.cta_group::2.
//
// Why: an SS-mode MMA reads its A tile (128 pixels x 16 channels = 4 KB) and its B tile (16 x Nt x 2 B) from shared
// memory every time it is issued.  At Nt = 64 that is 6 KB per 32 tensor-pipe cycles = 192 B/cycle against the 128
// B/cycle a shared memory delivers: the 64-channel layers of the UNet (layer1, decoder1, decoder0: 40 % of the
// network's 3x3 time) ran at ~58 cycles per MMA instead of 32 (profiles/r2_conv_dbg.txt).  Two CTAs on the two SMs of
// a TPC issue ONE M = 256 MMA: each SM reads its own 128 pixels and only HALF of the weights (Nt/2 columns, the
// hardware exchanges the halves), 5 KB per SM per MMA at Nt = 64, and each CTA fetches only half of every weight
// stage from L2.  The 256-channel layers are not operand-fetch bound, but with half the weight stream per CTA they can
// take ONE tile per CTA and work item (TM = 1) without flooding L2, so that two accumulator sets fit tensor memory and
// their epilogue (residual add included) overlaps the next item's MMAs: +2-3 % on the whole step (MVPNET_B200_CONV_PAIR_NT=128
// keeps them on the single-CTA kernel).
//
// Structure (same roles, rings and arithmetic as tc_conv3x3_kernel — results are bit-identical):
//   * cluster of 2 CTAs; a work item is 2 x TM tiles, CTA r takes tiles [r*TM, r*TM + TM) of it
//   * the LEADER (cluster rank 0) issues every MMA; patches and weight halves are loaded by both CTAs with
//     cp.async.bulk.tensor...cta_group::2, whose transaction bytes all land on the leader's "full" barriers
//   * tcgen05.commit...multicast::cluster releases the "empty" barriers and publishes the accumulators in both CTAs
//   * the peer's epilogue warps release the accumulator set with a remote mbarrier arrive on the leader
//   * the leader draws work items from the global counter and publishes them into both CTAs' rings
#include "common.cuh"

namespace mvp {
namespace tcc {

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// shared::cluster address of `addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on a barrier of either CTA.  `_cl`: release at CLUSTER scope — everything this thread wrote before (the remote
// store of a work index) is visible to the other CTA's waiter; it costs a cluster-wide memory fence (MEMBAR), which behind
// the output stores of an epilogue warp showed up as 6 % of the kernel's stall samples (profiles/r2_conv_pair_lines.txt).
// The plain form (default semantics, CTA scope, what CUTLASS' ClusterBarrier::arrive(cta_id) emits) is enough where
// nothing written by generic-proxy stores is handed over: releasing an accumulator set (ordered by
// tcgen05.fence::before_thread_sync) or a ring slot that was only read.
__device__ __forceinline__ void mbar_arrive_cl(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void st_cluster_u32(uint32_t cluster_addr, uint32_t v) {
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(cluster_addr), "r"(v) : "memory");
}
// wait on a barrier of THIS CTA whose arrivals may come from the other CTA (cluster-scope acquire)
__device__ __forceinline__ void mbar_wait_cl(uint32_t bar, uint32_t parity) {
  uint32_t ok, spins = 0;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!ok && ++spins > (1u << 26)) __trap();
  } while (!ok);
}
__device__ __forceinline__ void umma_bf16_2cta(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives (when every MMA issued so far has completed) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2cta(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t smem_dst, uint32_t cols) {   // the same warp of both CTAs
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// tensor loads whose completion bytes are counted on a barrier of the LEADER CTA (`bar_cl`: shared::cluster address)
__device__ __forceinline__ void tma_load_3d_2cta(uint32_t dst, const CUtensorMap *map, int c0, int c1, int c2, uint32_t bar_cl) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(bar_cl)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_2cta(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar_cl) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar_cl)
               : "memory");
}

constexpr int P_SCHED_CONSUMERS = 2 * (2 + EPI_WARPS);   // per CTA: 8 epilogue warps, the weight producer, and the issuer (leader) / patch producer (peer)

// leader's patch-producer warp: draw the next work item and publish it in both CTAs
__device__ __forceinline__ int sched_produce2(uint32_t k, volatile int *ring, uint32_t ring_s, uint32_t bar_full, uint32_t bar_empty, unsigned int *counter) {
  const uint32_t slot = k & (SCHED_DEPTH - 1);
  if (k >= SCHED_DEPTH) mbar_wait_cl(bar_empty + 8 * slot, ((k / SCHED_DEPTH) - 1u) & 1u);
  if (elect_one()) {
    const unsigned int v = atomicAdd(counter, 1u);
    ring[slot] = (int)v;
    st_cluster_u32(mapa(ring_s + 4 * slot, 1), v);
    mbar_arrive(bar_full + 8 * slot);
    mbar_arrive_cl(mapa(bar_full + 8 * slot, 1));          // release.cluster: orders the remote store before the arrive
  }
  __syncwarp();
  return ring[slot];
}
// any other role, either CTA: the slot is released on the LEADER's "empty" barrier
__device__ __forceinline__ int sched_consume2(uint32_t k, volatile int *ring, uint32_t bar_full, uint32_t bar_empty) {
  const uint32_t slot = k & (SCHED_DEPTH - 1);
  mbar_wait_cl(bar_full + 8 * slot, (k / SCHED_DEPTH) & 1u);
  const int w = ring[slot];
  __syncwarp();
  if (elect_one()) mbar_arrive_remote(mapa(bar_empty + 8 * slot, 0));
  __syncwarp();
  return w;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
tc_conv3x3_pair_kernel(const __grid_constant__ ConvArgs a, const __grid_constant__ CUtensorMap wmap) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  // shared memory: [asets][TM] patches | [stages] weight ring (this CTA's half of the columns) | barriers
  unsigned char *a_base = smem;
  const uint32_t tap_bytes = 32u * (uint32_t)a.Nt, stage_bytes = tap_bytes * (uint32_t)a.tps;      // half of the columns
  unsigned char *b_base = a_base + (size_t)a.asets * a.TM * SLOT_BYTES;
  uint64_t *bars = reinterpret_cast<uint64_t *>(b_base + (size_t)a.stages * stage_bytes);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * MAX_STAGES + 2 * MAX_ASETS + 4);
  volatile int *s_ring = reinterpret_cast<volatile int *>(bars + 28);
  const uint32_t ring_s = smem_u32(bars + 28);
  const uint32_t bar_sfull = smem_u32(bars + 32), bar_sempty = smem_u32(bars + 32 + SCHED_DEPTH);
  float *s_bias = reinterpret_cast<float *>(bars + 48);
  const uint32_t bar_bfull = smem_u32(bars), bar_bempty = smem_u32(bars + MAX_STAGES);
  const uint32_t bar_afull = smem_u32(bars + 2 * MAX_STAGES), bar_aempty = smem_u32(bars + 2 * MAX_STAGES + MAX_ASETS);
  const uint32_t bar_accfull = smem_u32(bars + 2 * MAX_STAGES + 2 * MAX_ASETS), bar_accempty = bar_accfull + 16;

  if (tid == 0) {
    for (int s = 0; s < MAX_STAGES; ++s) { mbar_init(bar_bfull + 8 * s, 1); mbar_init(bar_bempty + 8 * s, 1); }
    for (int s = 0; s < MAX_ASETS; ++s) { mbar_init(bar_afull + 8 * s, 1); mbar_init(bar_aempty + 8 * s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(bar_accfull + 8 * s, 1); mbar_init(bar_accempty + 8 * s, 2 * EPI_WARPS); }
    for (int s = 0; s < SCHED_DEPTH; ++s) { mbar_init(bar_sfull + 8 * s, 1); mbar_init(bar_sempty + 8 * s, P_SCHED_CONSUMERS); }
    fence_barrier_init();
  }
  if (warp == EPI_WARPS) tmem_alloc_2cta(smem_u32(tmem_slot), (uint32_t)a.tmem_cols);
  for (int i = tid; i < a.Cout; i += THREADS) s_bias[i] = a.bias[i];
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();               // the other CTA's barriers exist before anything arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int ngroups = (int)a.ngroups, nworks = ngroups * a.NB;       // a group = 2 * TM tiles
  const uint32_t slab_bytes = (uint32_t)(18 * HC * 16);
  const uint32_t S = (uint32_t)a.stages, AS = (uint32_t)a.asets, NA = (uint32_t)a.nacc;
  auto tiles_of = [&](long long group, uint32_t r) {                 // tiles of CTA r in this group
    const long long left = a.ntiles - (group * 2 + r) * a.TM;
    return left <= 0 ? 0 : (left < a.TM ? (int)left : a.TM);
  };

  if (warp < EPI_WARPS) {
    // =========================== epilogue (both CTAs, own tiles): epilogue_tiles of tc_conv.cu =======================
    const EpiThread e = epi_thread(a, warp, lane, tmem_base);
    const uint32_t accempty_leader = mapa(bar_accempty, 0);
    for (uint32_t it = 0;; ++it) {
      const int w = sched_consume2(it, s_ring, bar_sfull, bar_sempty);
      if (w >= nworks) break;
      const int nb = w / ngroups;
      const long long group = w - nb * ngroups;
      const uint32_t set = it % NA;
      epilogue_tiles(a, e, nb, (group * 2 + rank) * a.TM, tiles_of(group, rank), set, s_bias, bar_accfull + 8 * set, (it / NA) & 1u);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(accempty_leader + 8 * set);  // one arrival per epilogue warp of either CTA
    }
  } else if (warp == EPI_WARPS) {
    // =========================== MMA issuer: the leader CTA only ======================================================
    if (leader) {
      const uint32_t idesc = make_idesc(256, a.Nt);
      const uint32_t a_s = smem_u32(a_base), b_s = smem_u32(b_base);
      // B: this CTA holds Nt/2 columns: [tap][hi|lo][k8 (2)][n (Nt/2)][8]
      const uint64_t adesc0 = make_desc(0, slab_bytes, HC * 16), bdesc0 = make_desc(0, (uint32_t)a.Nt * 8u, 128);
      const uint32_t a_lo0 = (uint32_t)adesc0 + (a_s >> 4), a_hi32 = (uint32_t)(adesc0 >> 32);
      const uint32_t b_lo0 = (uint32_t)bdesc0 + (b_s >> 4), b_hi32 = (uint32_t)(bdesc0 >> 32);
      const uint32_t set16 = (uint32_t)(a.TM * SLOT_BYTES) >> 4, stage16 = stage_bytes >> 4, tap16 = tap_bytes >> 4, lo_of_hi = (uint32_t)a.Nt;
      const uint32_t row16 = (uint32_t)HC;
      auto desc64 = [](uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; };
      uint32_t ss = 0, a_ph = 0, s = 0, b_ph = 0, set = 0, acc_ph = 0;
      for (uint32_t w_it = 0;; ++w_it) {
        const int w = sched_consume2(w_it, s_ring, bar_sfull, bar_sempty);
        if (w >= nworks) break;
        const int nt = tiles_of(w % ngroups, 0);                     // the leader's tiles: never fewer than the peer's
        if (w_it >= NA) mbar_wait_cl(bar_accempty + 8 * set, acc_ph ^ 1u);
        tc_fence_after();
        const uint32_t d0 = tmem_base + set * (uint32_t)(a.TM * a.Nt);
        uint32_t first = 0u;
        for (int c = 0; c < a.nchunks; ++c) {
          mbar_wait(bar_afull + 8 * ss, a_ph);
          const uint32_t a_org_lo = a_lo0 + ss * set16;
          for (int st = 0; st < 9; st += a.tps) {
            mbar_wait(bar_bfull + 8 * s, b_ph);
            tc_fence_after();
            const uint32_t b_lo = b_lo0 + s * stage16;
            if (elect_one()) {
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
                    umma_bf16_2cta(d, ah, bh, idesc, acc);
                    umma_bf16_2cta(d, ah, bl, idesc, 1u);
                    umma_bf16_2cta(d, al, bh, idesc, 1u);
                    acc = 1u;
                  }
                }
              }
              umma_commit_2cta(bar_bempty + 8 * s);
            }
            __syncwarp();
            first = 1u;
            if (++s == S) { s = 0; b_ph ^= 1u; }
          }
          if (elect_one()) umma_commit_2cta(bar_aempty + 8 * ss);
          __syncwarp();
          if (++ss == AS) { ss = 0; a_ph ^= 1u; }
        }
        if (elect_one()) umma_commit_2cta(bar_accfull + 8 * set);
        __syncwarp();
        if (++set == NA) { set = 0; acc_ph ^= 1u; }
      }
    }
  } else if (warp == EPI_WARPS + 1) {
    // =========================== patch producer (both CTAs, own tiles; bytes are counted on the leader) ===============
    const uint32_t box_bytes = 2u * slab_bytes;
    const uint32_t a_s = smem_u32(a_base);
    const uint32_t afull_leader = mapa(bar_afull, 0);
    uint32_t ss = 0, ph = 0, it = 0;
    for (uint32_t k = 0;; ++k) {
      const int w = leader ? sched_produce2(k, s_ring, ring_s, bar_sfull, bar_sempty, a.sched) : sched_consume2(k, s_ring, bar_sfull, bar_sempty);
      if (w >= nworks) break;
      const long long group = w % ngroups;
      const int nt = tiles_of(group, rank), nt_both = tiles_of(group, 0) + tiles_of(group, 1);
      const long long tile0 = (group * 2 + rank) * a.TM;
      int cx[MAX_TM], cy[MAX_TM], cn[MAX_TM];
#pragma unroll
      for (int t = 0; t < MAX_TM; ++t) {
        const TileCoord tc_ = tile_coord(a, t < nt ? tile0 + t : 0);
        cx[t] = (tc_.x0 - 1) * 8; cy[t] = tc_.y0 - 1; cn[t] = tc_.n;
      }
      for (int c = 0; c < a.nchunks; ++c, ++it) {
        if (it >= AS) mbar_wait(bar_aempty + 8 * ss, ph ^ 1u);
        const int k0 = c * 16;
        const bool first = k0 < a.C1;
        const CUtensorMap *mh = first ? &a.m1h : &a.m2h, *ml = first ? &a.m1l : &a.m2l;
        const int C8 = (first ? a.C1 : a.C2) >> 3, s0 = (first ? k0 : k0 - a.C1) >> 3;
        const uint32_t dst0 = a_s + ss * (uint32_t)(a.TM * SLOT_BYTES);
        if (elect_one()) {
          if (leader) mbar_expect_tx(bar_afull + 8 * ss, (uint32_t)nt_both * 2u * box_bytes);
#pragma unroll
          for (int t = 0; t < MAX_TM; ++t) {
            if (t < nt) {
              const uint32_t dst = dst0 + (uint32_t)(t * SLOT_BYTES);
              tma_load_3d_2cta(dst, mh, cx[t], cy[t], cn[t] * C8 + s0, afull_leader + 8 * ss);
              tma_load_3d_2cta(dst + SLOT_HALF, ml, cx[t], cy[t], cn[t] * C8 + s0, afull_leader + 8 * ss);
            }
          }
        }
        __syncwarp();
        if (++ss == AS) { ss = 0; ph ^= 1u; }
      }
    }
    if (leader) {                    // the last PAIR to run dry re-arms the counters
      if (elect_one()) {
        const unsigned int done = atomicAdd(a.sched + 1, 1u);
        if (done == gridDim.x / 2 - 1) { a.sched[0] = 0u; a.sched[1] = 0u; __threadfence(); }
      }
      __syncwarp();
    }
  } else {
    // =========================== weight producer (both CTAs: this CTA's half of the columns) ==========================
    const uint32_t b_s = smem_u32(b_base);
    const uint32_t bfull_leader = mapa(bar_bfull, 0);
    const int per_work = a.nchunks * 9 / a.tps;
    const int rows_tap = (int)(tap_bytes >> 7);                       // 128-byte rows of the weight tensor map per tap
    uint32_t s = 0, ph = 0, it = 0;
    for (uint32_t k = 0;; ++k) {
      const int w = sched_consume2(k, s_ring, bar_sfull, bar_sempty);
      if (w >= nworks) break;
      const int nb = w / ngroups;
      // packed with block width Nt/2: block 2 * nb + rank is this CTA's half of output block nb
      int wrow = (int)(((long long)(2 * nb + (int)rank) * per_work * stage_bytes) >> 7);
      for (int j = 0; j < per_work; ++j, ++it) {
        if (it >= S) mbar_wait(bar_bempty + 8 * s, ph ^ 1u);
        if (elect_one()) {
          if (leader) mbar_expect_tx(bar_bfull + 8 * s, 2u * stage_bytes);
          for (int tp = 0; tp < a.tps; ++tp)
            tma_load_2d_2cta(b_s + s * stage_bytes + (uint32_t)tp * tap_bytes, &wmap, 0, wrow + tp * rows_tap, bfull_leader + 8 * s);
        }
        __syncwarp();
        wrow += a.tps * rows_tap;
        if (++s == S) { s = 0; ph ^= 1u; }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();               // nobody leaves (or frees tensor memory) while the other CTA may still signal it
  if (warp == EPI_WARPS) tmem_dealloc_2cta(tmem_base, (uint32_t)a.tmem_cols);
}

// the packed weights as a 2-D byte tensor of 128-byte rows; box = one tap of one CTA's half block
static int make_weight_map(CUtensorMap *m, const void *wp, int64_t bytes, int rows_tap) {
  EncodeTiledFn enc = encode_tiled();
  MVP_REQUIRE(enc != nullptr, MVP_ERR_UNSUPPORTED, "tc_conv3x3: cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t dims[2] = {128, (cuuint64_t)(bytes / 128)}, strides[1] = {128};
  const cuuint32_t box[2] = {128, (cuuint32_t)rows_tap}, es[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(wp), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MVP_REQUIRE(r == CUDA_SUCCESS, MVP_ERR_INVALID_ARG, "tc_conv3x3: cuTensorMapEncodeTiled failed (%d) for the weights (%lld bytes)", (int)r, (long long)bytes);
  return 0;
}

}  // namespace tcc
}  // namespace mvp

// The pair kernel serves images of more than 8 rows (output blocks of 32..256 channels, multiples of 32).  Its weights are packed with
// block width mvp_tc_conv3x3_nt(Cout) / 2 (each CTA of a pair streams its own half block).
extern "C" int mvp_tc_conv3x3_pair_supported(int64_t Cout, int64_t H) {
  static const bool allow = [] { const char *e = getenv("MVPNET_B200_CONV_PAIR"); return !(e && e[0] == '0'); }();
  static const int64_t max_nt = [] { const char *e = getenv("MVPNET_B200_CONV_PAIR_NT"); const int64_t v = e ? atoll(e) : 0; return v >= 32 ? v : (int64_t)256; }();
  const int64_t nt = mvp_tc_conv3x3_nt(Cout);
  return allow && nt <= max_nt && nt % 32 == 0 && Cout % nt == 0 && H > 8;
}

extern "C" int mvp_tc_conv3x3_pair(const void *x1, int64_t C1, const void *x2, int64_t C2, int64_t N, int64_t H, int64_t W,
                                   const void *w_packed_half, const float *bias, int64_t Cout, const void *residual, int relu,
                                   void *out_planar, float *out_nhwc, void *out_rows, mvp_stream_t stream) {
  using namespace mvp;
  MVP_REQUIRE(N >= 0 && H > 0 && W > 0, MVP_ERR_INVALID_ARG, "tc_conv3x3_pair: bad sizes");
  MVP_REQUIRE(C1 > 0 && C1 % 16 == 0 && C2 >= 0 && C2 % 16 == 0, MVP_ERR_UNSUPPORTED, "tc_conv3x3_pair: input channels must be multiples of 16");
  MVP_REQUIRE(Cout > 0 && mvp_tc_conv3x3_pair_supported(Cout, H), MVP_ERR_UNSUPPORTED,
              "tc_conv3x3_pair: needs output blocks of 32..256 channels (multiples of 32) and H > 8 (mvp_tc_conv3x3_pair_supported)");
  MVP_REQUIRE(N * H * W < (1LL << 31), MVP_ERR_UNSUPPORTED, "tc_conv3x3_pair: more than 2^31 pixels");
  if (N == 0) return 0;
  MVP_REQUIRE(x1 && w_packed_half && bias && (out_planar || out_nhwc || out_rows) && (x2 || C2 == 0), MVP_ERR_NULL, "tc_conv3x3_pair: null pointer");
  MVP_REQUIRE((((uintptr_t)x1 | (uintptr_t)x2 | (uintptr_t)bias | (uintptr_t)residual | (uintptr_t)out_planar) & 15) == 0, MVP_ERR_INVALID_ARG,
              "tc_conv3x3_pair: pointers must be 16-byte aligned");
  MVP_REQUIRE((((uintptr_t)out_nhwc | (uintptr_t)out_rows) & 31) == 0 && ((uintptr_t)w_packed_half & 127) == 0, MVP_ERR_INVALID_ARG,
              "tc_conv3x3_pair: the NHWC / row-split outputs must be 32-byte aligned, the packed weights 128-byte aligned");
  tcc::ConvArgs a = {};
  if (int rc = tcc::make_plane_map(&a.m1h, x1, N, H, W, C1, 0)) return rc;
  if (int rc = tcc::make_plane_map(&a.m1l, (const __nv_bfloat16 *)x1 + N * C1 * H * W, N, H, W, C1, 0)) return rc;
  if (C2 > 0) {
    if (int rc = tcc::make_plane_map(&a.m2h, x2, N, H, W, C2, 0)) return rc;
    if (int rc = tcc::make_plane_map(&a.m2l, (const __nv_bfloat16 *)x2 + N * C2 * H * W, N, H, W, C2, 0)) return rc;
  }
  a.C1 = (int)C1; a.C2 = (int)C2; a.N = (int)N; a.H = (int)H; a.W = (int)W;
  a.bias = bias; a.res = (const __nv_bfloat16 *)residual;
  a.out_p = (__nv_bfloat16 *)out_planar; a.out_f = out_nhwc;
  a.out_rh = (__nv_bfloat16 *)out_rows; a.out_rl = out_rows ? (__nv_bfloat16 *)out_rows + N * H * W * Cout : nullptr; a.plane_out = N * Cout * H * W; a.relu = relu;
  a.Cout = (int)Cout; a.Nt = (int)mvp_tc_conv3x3_nt(Cout); a.NB = a.Cout / a.Nt;
  a.ipt = 1;
  a.TX = (int)((W + 7) / 8);
  a.TY = (int)((H + 15) / 16);
  a.ntiles = N * a.TX * a.TY;
  const int pairs = sm_count() / 2;
  // tiles per CTA and work item: the largest TM with the smallest makespan over the CTA pairs (see mvp_tc_conv3x3)
  a.TM = a.Nt <= 64 ? 4 : (a.Nt <= 128 ? 2 : 1);     // 256 columns: one tile per CTA, so that two accumulator sets fit tensor memory
  {
    long long best = -1;
    int best_tm = 1;
    for (int tm = a.TM; tm >= 1; tm >>= 1) {
      const long long works = (a.ntiles + 2 * tm - 1) / (2 * tm) * a.NB, span = (works + pairs - 1) / pairs * tm;
      if (best < 0 || span * 100 < best * 90) { best = span; best_tm = tm; }   // a smaller TM must buy > 10 %: it multiplies the weight traffic (layer2: TM = 2 0.123 ms, TM = 1 0.126 ms at 8 % less span)
    }
    a.TM = best_tm;
  }
  {
    const char *tm = getenv("MVPNET_B200_CONV_TM");
    if (tm && atoi(tm) >= 1 && atoi(tm) <= tcc::MAX_TM && atoi(tm) * a.Nt <= 512) a.TM = atoi(tm);
  }
  a.nacc = 2 * a.TM * a.Nt <= 512 ? 2 : 1;
  a.ngroups = (a.ntiles + 2 * a.TM - 1) / (2 * a.TM);
  a.nchunks = (int)((C1 + C2) / 16);
  a.tmem_cols = 32;
  while (a.tmem_cols < a.nacc * a.TM * a.Nt) a.tmem_cols <<= 1;
  a.asets = tcc::MAX_ASETS;
  a.stages = tcc::MAX_STAGES;
  a.tps = 9;
  {
    const char *st = getenv("MVPNET_B200_CONV_STAGES");
    if (st && atoi(st) >= 2 && atoi(st) <= tcc::MAX_STAGES) a.stages = atoi(st);
    const char *as = getenv("MVPNET_B200_CONV_ASETS");
    if (as && atoi(as) >= 2 && atoi(as) <= tcc::MAX_ASETS) a.asets = atoi(as);
    const char *tp = getenv("MVPNET_B200_CONV_TPS");
    if (tp && (atoi(tp) == 1 || atoi(tp) == 3 || atoi(tp) == 9)) a.tps = atoi(tp);
  }
  auto smem_of = [&]() { return (size_t)a.asets * a.TM * tcc::SLOT_BYTES + (size_t)a.stages * a.tps * 32 * a.Nt + 512 + (size_t)a.Cout * 4; };
  while (a.stages > 3 && smem_of() > tcc::conv_smem_cap()) --a.stages;
  while (a.asets > 2 && smem_of() > tcc::conv_smem_cap()) --a.asets;
  while (a.stages > 2 && smem_of() > tcc::conv_smem_cap()) --a.stages;
  const size_t smem = smem_of();
  MVP_REQUIRE(smem <= tcc::conv_smem_cap(), MVP_ERR_UNSUPPORTED, "tc_conv3x3_pair: shared memory budget exceeded");
  CUtensorMap wmap;
  if (int rc = tcc::make_weight_map(&wmap, w_packed_half, (C1 + C2) * Cout * 9 * 4, (32 * a.Nt) >> 7)) return rc;
  cudaError_t e = cudaFuncSetAttribute(tcc::tc_conv3x3_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("tc_conv3x3_pair: smem attribute (%zu B): %s", smem, cudaGetErrorString(e)); return (int)e; }
  const long long nworks = a.ngroups * a.NB;
  MVP_REQUIRE(nworks < (1LL << 30), MVP_ERR_UNSUPPORTED, "tc_conv3x3_pair: too many work items");
  a.sched = tcc::sched_pair((cudaStream_t)stream);
  MVP_REQUIRE(a.sched != nullptr, MVP_ERR_UNSUPPORTED,
              "tc_conv3x3_pair: no scheduler counters (first call inside a stream capture, or more than 61440 captured launches)");
  long long clusters = pairs;
  if (clusters > nworks) clusters = nworks;
  static const bool debug = getenv("MVPNET_B200_DEBUG") != nullptr;
  if (debug)
    fprintf(stderr, "[tc_conv3x3_pair] N=%d H=%d W=%d Cin=%d+%d Cout=%d Nt=%d tiles=%lld TM=%d nacc=%d works=%lld asets=%d stages=%d tps=%d smem=%zu tmem=%d clusters=%lld\n",
            a.N, a.H, a.W, a.C1, a.C2, a.Cout, a.Nt, a.ntiles, a.TM, a.nacc, nworks, a.asets, a.stages, a.tps, smem, a.tmem_cols, clusters);
  tcc::tc_conv3x3_pair_kernel<<<(unsigned)(2 * clusters), tcc::THREADS, smem, (cudaStream_t)stream>>>(a, wmap);
  return launch_status("tc_conv3x3_pair");
}
