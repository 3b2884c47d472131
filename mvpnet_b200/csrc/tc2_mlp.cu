// Second-generation fused "gather neighbourhood -> 1x1-conv chain -> reduce" kernel for sm_100a (tcgen05 / TMEM):
// SetAbstraction (mvpnet/models/pn2/modules.py:20-37, 100-108) and FeatureAggregation (mvpnet/models/mvpnet_3d.py:37-61,
// 100-109) when the gathered features arrive PRE-SPLIT (two bf16 planes hi = bf16(v), lo = bf16(v - hi), point-major
// rows of C = 64 * nchunks channels), which every producer on the fast path now writes from its epilogue.
//
// What changed against tc_mlp.cu (round 1; still used for the wide, weight-streaming chains), each point taken from
// the round-1 ncu source view (profiles/r1_ncu_fa_sa1.md: 28 % of the stalls on the row gather feeding the bf16
// split, 15 % on the accumulator barrier) or measured on the box (profiles/r2_probe_gather.md):
//   * no thread touches a gathered activation: rows go global -> shared memory with 16-byte cp.async, 8 consecutive
//     lanes = the 8 chunks of one 128-byte row (whole-line reads), written straight into the SWIZZLE_128B K-major
//     operand layout (conflict-free), and the MMA reads that tile through a swizzled descriptor.  TMA gather4
//     (cp.async.bulk.tensor.2d.tile::gather4) produces the same layout and was measured first: 2.1x slower here
//     (2327 vs 683 ns per 32 KB tile per SM), so it is not used;
//   * the gather of unit n+1 is issued as soon as the first layer's MMAs of unit n have retired and lands while
//     unit n's epilogues run; neighbour indices are fetched two units ahead, coordinates one unit ahead (registers);
//   * inner layers keep their A operand in TENSOR MEMORY: the epilogue of layer l writes relu(acc + bias) as packed
//     bf16 hi / lo pairs with tcgen05.st and layer l+1 is issued as tcgen05.mma [d], [a_tmem], b_desc — shared
//     memory is read only for the (small) weight operand, which lifts the operand-fetch bound of the N = 32 / 64
//     layers (an SS-mode N = 32 MMA reads 5 KB of shared memory for 16 cycles of math);
//   * weights and biases are resident in shared memory for the whole kernel;
//   * FeatureAggregation tiles are 128 points x one pixel slot (k passes; the running sum / max over the slots lives
//     in spare tensor-memory columns) instead of 32 points x 3 slots + 32 idle rows: no idle rows, no staging pass;
//   * outputs are written fp32 (module API / skip connections) and/or pre-split for the next gather.
// Precision scheme (bf16 hi/lo, three products per K-step, fp32 accumulation in TMEM) is unchanged.
#include <cuda_bf16.h>

#include "common.cuh"

namespace mvp {
namespace tc2 {
using namespace tc;   // PTX wrappers of tc_mlp.cu (same translation unit)

constexpr int MAXG = 6;             // tile groups per CTA: 3 of 256 threads (two warps per TMEM lane quarter) or up to 6 of 128 threads
constexpr int MAXL = 6;
constexpr int CHUNK = 16384;        // bytes of one plane of a 64-channel chunk of the gathered tile: 128 rows x 128 B
constexpr int REL_PLANE = 4096;     // bytes of one plane of the relation slab pair: 2 K-slabs x 128 rows x 16 B

struct Args {
  long long rows_out;                       // SA: B * M centroids; FA: B * Np points
  const __nv_bfloat16 *src_hi, *src_lo;     // [source rows, C]
  int C;
  const float *xyz;                         // SA: keys [B, N, 3];  FA: pixel xyz [B, P, 3]
  const float *new_xyz;                     // SA: centroids [B, M, 3];  FA: points [B, Np, 3]
  const int64_t *nbr;                       // SA: [B, M, 32];  FA: knn [B, Np, k]
  unsigned n_src, n_out;                    // per cloud: N / P, M / Np
  int k;                                    // FA: pixel slots per point (1..4)
  int reduce;                               // FA: REDUCE_SUM / REDUCE_MAX
  unsigned hw, w, hp_wp, wp, nv;            // FA: pixel j = (v, y, x) of an nv x h x w stack -> feature row (b*nv+v)*hp_wp + y*wp + x
  float *out_f32;                           // [rows_out, out_channels] or null
  __nv_bfloat16 *out_hi, *out_lo;           // [rows_out, out_channels] or null
  long long *prof;                          // debug (MVPNET_B200_TC2_PROF): per-phase clock64 sums of CTA 0 (16 words per group, issuer at 96), or null
};

struct Plan {
  int num_layers;
  int k[MAXL], n[MAXL], relu[MAXL];
  const __nv_bfloat16 *w_hi[MAXL], *w_lo[MAXL];   // [k/8][n][8]
  const float *bias[MAXL];
  int out_channels;
  int woff[MAXL], boff[MAXL], wbytes, bfloats;
  int nchunks;
  int acc_cols, ta_cols, part_cols, tg_cols, tmem_alloc;   // per group: accumulator | A hi | A lo | FA partial reduction
  int groups;
};

__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {   // K-major SWIZZLE_128B: 8-row groups 1024 B apart
  return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ uint64_t desc64(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
               "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, uint32_t src_bytes) {   // src_bytes 0 = zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// Worker-side wait: try_wait suspends the warp in hardware instead of spinning.  The fused kernels are issue-bound
// (ncu: a third of all issued warp instructions were test_wait spins of idle groups), so a waiting group must not
// compete for issue slots with the groups that have work.  A protocol error traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity) {
  uint32_t ok, spins = 0;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(20000u)      // suspend-time hint (ns): fewer wake-ups that only re-issue the wait
        : "memory");
    if (!ok && ++spins > (1u << 20)) __trap();
  } while (!ok);
}
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
template <int GT>
__device__ __forceinline__ void group_bar(int g) { asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "n"(GT) : "memory"); }

// the group's unit sequence: tiles blockIdx.x + (g + j * NG) * gridDim.x, each `passes` times (FA: one per pixel slot)
struct Cursor {
  long long tile;
  int slot;
  bool valid;
};

struct RowRef {      // what thread r (< 128) knows about its row of a unit
  int frow;          // feature row in src_hi / src_lo, -1 = no source (zero row)
  long long xrow;    // row of `xyz`
  long long orow;    // centroid / point row of `new_xyz`
};

template <int MODE>
__device__ __forceinline__ long long idx_address(const Args &a, const Cursor &c, int r) {   // element offset into nbr, -1 = none
  if (!c.valid) return -1;
  if (MODE == MODE_SA) {
    const long long gid = c.tile * 4 + (r >> 5);
    return gid < a.rows_out ? gid * 32 + (r & 31) : -1;
  }
  const long long pid = c.tile * 128 + r;
  return pid < a.rows_out ? pid * a.k + c.slot : -1;
}

template <int MODE>
__device__ __forceinline__ RowRef resolve(const Args &a, const Cursor &c, int r, long long j) {
  RowRef o;
  o.frow = -1; o.xrow = 0; o.orow = 0;
  if (!c.valid) return o;
  const long long orow = MODE == MODE_SA ? c.tile * 4 + (r >> 5) : c.tile * 128 + r;
  if (orow >= a.rows_out || j < 0 || j >= (long long)a.n_src) return o;
  const unsigned b = (unsigned)orow / a.n_out;                       // rows_out < 2^31 (checked by the host)
  o.orow = orow;
  o.xrow = (long long)b * a.n_src + j;
  if (MODE == MODE_SA) {
    o.frow = (int)o.xrow;
  } else {
    const unsigned ju = (unsigned)j, v = ju / a.hw, pix = ju - v * a.hw, y = pix / a.w, x = pix - y * a.w;
    o.frow = (int)((b * a.nv + v) * a.hp_wp + y * a.wp + x);
  }
  return o;
}

// GT = worker threads per tile group: 256 (8 warps, two per TMEM lane quarter taking alternate 16-column chunks) or 128
// (4 warps, one per quarter).  More, narrower groups keep more units in flight per SM: the phase clocks of the 3 x 256
// configuration showed a group waiting 40 % of its time on the request -> MMA -> commit round trips with the SM's issue
// slots half idle.
template <int MODE, int NG, int GT>
__global__ void __launch_bounds__(NG *GT + 32, 1)
tc2_kernel(const Args a, const Plan m, long long num_tiles) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int NW = NG * (GT / 32);
  constexpr int NTHREADS = NG * GT + 32;
  const uint32_t gbytes = (uint32_t)m.nchunks * 2 * CHUNK + 2 * REL_PLANE;
  unsigned char *wreg = smem + (size_t)NG * gbytes;
  float *bias_s = reinterpret_cast<float *>(wreg + ((m.wbytes + 127) & ~127));
  int *src_s = reinterpret_cast<int *>(bias_s + ((m.bfloats + 31) & ~31));
  uint64_t *bars = reinterpret_cast<uint64_t *>(src_s + NG * 128);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * MAXG);
  const uint32_t bar_aready0 = smem_u32(bars), bar_acc0 = smem_u32(bars + MAXG);

  if (tid == 0) {
    for (int g = 0; g < MAXG; ++g) { mbar_init(bar_aready0 + 8 * g, GT); mbar_init(bar_acc0 + 8 * g, 1); }
    fence_barrier_init();
  }
  if (warp == NW) tmem_alloc(smem_u32(tmem_slot), (uint32_t)m.tmem_alloc);
  {   // resident weights (hi | lo per layer) and biases; relation slabs start as zeros (only channels 0..3 are ever written)
    for (int l = 0; l < m.num_layers; ++l) {
      const size_t bytes = (size_t)m.k[l] * m.n[l] * 2;
      const uint4 *gh = reinterpret_cast<const uint4 *>(m.w_hi[l]), *gl = reinterpret_cast<const uint4 *>(m.w_lo[l]);
      uint4 *sh = reinterpret_cast<uint4 *>(wreg + m.woff[l]), *sl = reinterpret_cast<uint4 *>(wreg + m.woff[l] + bytes);
      for (size_t o = tid; o < bytes / 16; o += NTHREADS) { sh[o] = __ldg(gh + o); sl[o] = __ldg(gl + o); }
      for (int o = tid; o < m.n[l]; o += NTHREADS) bias_s[m.boff[l] + o] = __ldg(m.bias[l] + o);
    }
    for (int g = 0; g < NG; ++g) {
      uint4 *z = reinterpret_cast<uint4 *>(smem + (size_t)g * gbytes + (size_t)m.nchunks * 2 * CHUNK);
      for (int o = tid; o < 2 * REL_PLANE / 16; o += NTHREADS) z[o] = make_uint4(0u, 0u, 0u, 0u);
    }
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const long long n_my = blockIdx.x < num_tiles ? (num_tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;   // tiles of this CTA
  const int passes = MODE == MODE_FA ? a.k : 1;
  const int L = m.num_layers;

  if (warp < NW) {
    // =========================== workers ============================================================================
    const int g = warp / (GT / 32), ltid = tid - g * GT, lwarp = ltid >> 5;
    const uint32_t A_s = smem_u32(smem) + (uint32_t)g * gbytes;
    unsigned char *rel_hi = smem + (size_t)g * gbytes + (size_t)m.nchunks * 2 * CHUNK, *rel_lo = rel_hi + REL_PLANE;
    int *src = src_s + g * 128;
    const uint32_t bar_aready = bar_aready0 + 8 * g, bar_acc = bar_acc0 + 8 * g;
    constexpr int NHALF = GT / 128;                       // warps per TMEM lane quarter
    const int quarter = lwarp & 3, half = lwarp >> 2;
    const uint32_t t_lane = tmem_base + (uint32_t)(g * m.tg_cols) + ((uint32_t)(quarter * 32) << 16);
    const uint32_t t_ahi = t_lane + (uint32_t)m.acc_cols, t_alo = t_ahi + (uint32_t)m.ta_cols, t_part = t_alo + (uint32_t)m.ta_cols;
    const int row = quarter * 32 + lane;
    const long long my_tiles = n_my > g ? (n_my - g + NG - 1) / NG : 0;
    const long long units = my_tiles * passes;
    auto first = [&]() { Cursor c; c.tile = blockIdx.x + (long long)g * gridDim.x; c.slot = 0; c.valid = my_tiles > 0; return c; };
    auto advance = [&](Cursor c) {
      if (!c.valid) return c;
      if (++c.slot == passes) { c.slot = 0; c.tile += (long long)NG * gridDim.x; c.valid = c.tile < num_tiles; }
      return c;
    };
    auto load_idx = [&](const Cursor &c) -> long long {
      if (ltid >= 128) return -1;
      const long long o = idx_address<MODE>(a, c, ltid);
      return o >= 0 ? (long long)__ldg(a.nbr + o) : -1;
    };
    // coordinates of the row's source and of its centroid / point (issued early, consumed by build())
    auto load_xyz = [&](const RowRef &rr, float (&p)[3], float (&q)[3]) {
      if (ltid < 128 && rr.frow >= 0) {
#pragma unroll
        for (int i = 0; i < 3; ++i) { p[i] = __ldg(a.xyz + rr.xrow * 3 + i); q[i] = __ldg(a.new_xyz + rr.orow * 3 + i); }
      } else {
#pragma unroll
        for (int i = 0; i < 3; ++i) p[i] = q[i] = 0.f;
      }
    };
    // relation channels + source rows of a unit into shared memory, then the row gather (all 256 threads)
    auto build = [&](const RowRef &rr, const float (&p)[3], const float (&q)[3]) {
      if (ltid < 128) {
        float d[4];
        d[0] = __fsub_rn(p[0], q[0]); d[1] = __fsub_rn(p[1], q[1]); d[2] = __fsub_rn(p[2], q[2]);
        d[3] = MODE == MODE_FA ? __fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])) : 0.f;
        if (rr.frow < 0) d[0] = d[1] = d[2] = d[3] = 0.f;
        uint32_t h0, l0, h1, l1;
        split_pair(d[0], d[1], h0, l0);
        split_pair(d[2], d[3], h1, l1);
        const int off = (ltid >> 3) * 128 + (ltid & 7) * 16;
        *reinterpret_cast<uint4 *>(rel_hi + off) = make_uint4(h0, h1, 0u, 0u);
        *reinterpret_cast<uint4 *>(rel_lo + off) = make_uint4(l0, l1, 0u, 0u);
        src[ltid] = rr.frow;
      }
      group_bar<GT>(g);
      const int chunk = lane & 7, rsub = lane >> 3;
      constexpr int RPP = GT / 8;                          // rows per pass: every warp takes 4
#pragma unroll
      for (int ps = 0; ps < 128 / RPP; ++ps) {
        const int r = ps * RPP + lwarp * 4 + rsub;
        const int s = src[r];
        const uint32_t nbytes = s >= 0 ? 16u : 0u;
        const size_t e = (size_t)(s >= 0 ? s : 0) * a.C + chunk * 8;
        const uint32_t dst = A_s + (uint32_t)(r * 128 + ((chunk ^ (r & 7)) << 4));
        for (int qn = 0; qn < m.nchunks; ++qn) {
          cp_async16(dst + (uint32_t)(qn * 2 * CHUNK), a.src_hi + e + qn * 64, nbytes);
          cp_async16(dst + (uint32_t)(qn * 2 * CHUNK + CHUNK), a.src_lo + e + qn * 64, nbytes);
        }
      }
    };

    Cursor c0 = first(), c1 = advance(c0), c2 = advance(c1);
    float p1[3], q1[3];
    RowRef r1;
    long long raw2;
    {   // prologue: unit 0 built synchronously, unit 1 resolved, unit 2's index in flight
      const long long raw0 = load_idx(c0), raw1 = load_idx(c1);
      raw2 = load_idx(c2);
      const RowRef r0 = resolve<MODE>(a, c0, ltid & 127, raw0);
      float p0[3], q0[3];
      load_xyz(r0, p0, q0);
      if (c0.valid) build(r0, p0, q0);
      r1 = resolve<MODE>(a, c1, ltid & 127, raw1);
      load_xyz(r1, p1, q1);
    }
    uint32_t acc_phase = 0;
    const bool prof = a.prof != nullptr && blockIdx.x == 0 && ltid == 0;
    long long tp = prof ? clock64() : 0;
    auto mark = [&](int phase) {
      if (prof) { const long long t = clock64(); a.prof[g * 16 + phase] += t - tp; tp = t; }
    };
    for (long long n = 0; n < units; ++n) {
      // ---- layer 0 of unit n: its gather (issued one unit ago) has to have landed
      cp_async_wait_all();
      mark(0);
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(bar_aready);
      mbar_wait_sleep(bar_acc, acc_phase);
      acc_phase ^= 1u;
      tc_fence_after();
      mark(1);
      // ---- the gathered tile is free again: build unit n+1, move the prefetch pipeline one step
      if (c1.valid) build(r1, p1, q1);
      r1 = resolve<MODE>(a, c2, ltid & 127, raw2);
      load_xyz(r1, p1, q1);
      const Cursor cur = c0;
      c0 = c1; c1 = c2; c2 = advance(c2);
      raw2 = load_idx(c2);
      mark(2);

      // ---- epilogues: thread = row (TMEM lane 32 * quarter + lane); the two warps of a quarter take alternate 16-column
      //      chunks; the TMEM load of the next chunk is in flight while this one is processed
      for (int l = 0; l < L; ++l) {
        const int N = m.n[l];
        const bool last = l == L - 1;
        const float *bs = bias_s + m.boff[l];
        uint32_t rn[16];
        int c = half;
        if (c * 16 < N) tmem_ld16_issue(t_lane + (uint32_t)(c * 16), rn);
        while (c * 16 < N) {
          float v[16];
          tmem_ld_wait(rn);
#pragma unroll
          for (int q = 0; q < 16; ++q) v[q] = __uint_as_float(rn[q]);
          if ((c + NHALF) * 16 < N) tmem_ld16_issue(t_lane + (uint32_t)((c + NHALF) * 16), rn);
          const float4 *bp = reinterpret_cast<const float4 *>(bs + c * 16);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 bq = bp[q];
            add_pair(v[4 * q], v[4 * q + 1], bq.x, bq.y);
            add_pair(v[4 * q + 2], v[4 * q + 3], bq.z, bq.w);
          }
          if (m.relu[l]) {
#pragma unroll
            for (int q = 0; q < 16; ++q) v[q] = fmaxf(v[q], 0.f);
          }
          if (!last) {
            uint32_t h[8], lo[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) split_pair(v[2 * q], v[2 * q + 1], h[q], lo[q]);
            tmem_st8(t_ahi + (uint32_t)(c * 8), h);
            tmem_st8(t_alo + (uint32_t)(c * 8), lo);
          } else if (MODE == MODE_SA) {
            const long long gid = cur.tile * 4 + quarter;
            // max over the 32 neighbours (= lanes) of 16 columns as a halving butterfly: lane L ends up with column L >> 1
#pragma unroll
            for (int hh = 8, o = 16; hh >= 1; hh >>= 1, o >>= 1) {
              const bool up = (lane & o) != 0;
#pragma unroll
              for (int q = 0; q < hh; ++q) {
                const float send = up ? v[q] : v[q + hh], mine = up ? v[q + hh] : v[q];
                v[q] = fmaxf(mine, __shfl_xor_sync(0xffffffffu, send, o));
              }
            }
            const float keep = fmaxf(v[0], __shfl_xor_sync(0xffffffffu, v[0], 1));
            const int col = c * 16 + (lane >> 1);
            if (!(lane & 1) && col < m.out_channels && gid < a.rows_out) {
              const size_t o = (size_t)gid * m.out_channels + col;
              if (a.out_f32 != nullptr) a.out_f32[o] = keep;
              if (a.out_hi != nullptr) {
                const __nv_bfloat16 hb = __float2bfloat16_rn(keep);
                a.out_hi[o] = hb;
                a.out_lo[o] = __float2bfloat16_rn(__fsub_rn(keep, __bfloat162float(hb)));
              }
            }
          } else {
            // FA: reduce over the pixel slots through spare TMEM columns (same lane, same thread); the last slot writes the row
            if (cur.slot > 0) {
              uint32_t pr[16];
              tmem_ld16_issue(t_part + (uint32_t)(c * 16), pr);
              tmem_ld_wait(pr);
#pragma unroll
              for (int q = 0; q < 16; ++q) {
                const float prev = __uint_as_float(pr[q]);
                v[q] = a.reduce == REDUCE_SUM ? __fadd_rn(prev, v[q]) : fmaxf(prev, v[q]);
              }
            }
            if (cur.slot < passes - 1) {
              uint32_t w0[8], w1[8];
#pragma unroll
              for (int q = 0; q < 8; ++q) { w0[q] = __float_as_uint(v[q]); w1[q] = __float_as_uint(v[8 + q]); }
              tmem_st8(t_part + (uint32_t)(c * 16), w0);
              tmem_st8(t_part + (uint32_t)(c * 16 + 8), w1);
            }
            const long long pid = cur.tile * 128 + row;
            if (cur.slot == passes - 1 && pid < a.rows_out) {
              const size_t o = (size_t)pid * m.out_channels + c * 16;
              if (a.out_f32 != nullptr) {
#pragma unroll
                for (int q = 0; q < 2; ++q)       // 256-bit stores: whole L2 sectors per instruction
                  st_global_256(a.out_f32 + o + 8 * q, make_uint4(__float_as_uint(v[8 * q]), __float_as_uint(v[8 * q + 1]), __float_as_uint(v[8 * q + 2]), __float_as_uint(v[8 * q + 3])),
                                make_uint4(__float_as_uint(v[8 * q + 4]), __float_as_uint(v[8 * q + 5]), __float_as_uint(v[8 * q + 6]), __float_as_uint(v[8 * q + 7])));
              }
              if (a.out_hi != nullptr) {
                uint32_t h[8], lo[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) split_pair(v[2 * q], v[2 * q + 1], h[q], lo[q]);
                st_global_256(a.out_hi + o, make_uint4(h[0], h[1], h[2], h[3]), make_uint4(h[4], h[5], h[6], h[7]));
                st_global_256(a.out_lo + o, make_uint4(lo[0], lo[1], lo[2], lo[3]), make_uint4(lo[4], lo[5], lo[6], lo[7]));
              }
            }
          }
          c += NHALF;
        }
        if (last && MODE == MODE_FA) tmem_st_wait();   // the partial columns are re-read by this thread in the next slot
        mark(3 + 2 * l);
        if (!last) {
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(bar_aready);
          mbar_wait_sleep(bar_acc, acc_phase);
          acc_phase ^= 1u;
          tc_fence_after();
          mark(4 + 2 * l);
        }
      }
    }
    cp_async_wait_all();
  } else {
    // =========================== MMA issuer ===========================================================================
    // The whole warp polls the groups' request barriers and serves WHICHEVER group is ready (a fixed round-robin order
    // made the groups advance in lock step: all of them in their epilogues while the tensor pipe idled, then all of
    // them waiting); every value is warp-uniform, one elected lane issues.
    uint32_t ph[MAXG] = {};
    int layer[MAXG] = {};
    long long left[MAXG] = {};
    long long remaining = 0;
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      left[g] = (n_my > g ? (n_my - g + NG - 1) / NG : 0) * passes * L;
      remaining += left[g];
    }
    const uint32_t smem_s = smem_u32(smem), wreg_s = smem_u32(wreg);
    uint32_t idle = 0;
    const bool iprof = a.prof != nullptr && blockIdx.x == 0 && lane == 0;
    long long ti = iprof ? clock64() : 0;
    while (remaining > 0) {
      bool any = false;
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        if (left[g] == 0 || !mbar_test(bar_aready0 + 8 * g, ph[g])) continue;
        any = true;
        if (iprof) { const long long t = clock64(); a.prof[96] += t - ti; ti = t; }     // time spent polling
        ph[g] ^= 1u;
        tc_fence_after();
        const int l = layer[g];
        layer[g] = l + 1 == L ? 0 : l + 1;
        --left[g];
        --remaining;
        // Descriptors as (low word, high word): the low word holds the start address (>> 4) and the K-direction stride,
        // so stepping to the next K-step / plane / chunk is one 32-bit add.  This warp's instruction stream paces the
        // narrow layers (an N = 32 MMA is 16 cycles of tensor time): no 64-bit arithmetic, division or descriptor
        // re-encoding inside the loops (ncu, first version: 40 % of the worker samples were waits on this warp).
        // per-layer constants straight from the kernel parameters (constant bank, indexed by l)
        const int K = m.k[l], N = m.n[l];
        const uint32_t idesc = make_idesc(128, N), b_hi32 = (128u >> 4) | (1u << 14);
        const uint32_t wh = wreg_s + (uint32_t)m.woff[l];
        uint32_t b_h = ((wh >> 4) & 0x3fffu) | ((uint32_t)N << 16), b_l = b_h + (uint32_t)((K * N * 2) >> 4);
        const uint32_t kinc = (uint32_t)N * 2u;                 // 16 channels of the weight operand = N * 32 bytes (>> 4)
        const uint32_t t_acc = tmem_base + (uint32_t)(g * m.tg_cols);
        if (elect_one()) {
          if (l == 0) {
            constexpr uint32_t a_hi32 = (1024u >> 4) | (1u << 14) | (2u << 29);        // SWIZZLE_128B, 8-row groups 1024 B apart
            uint32_t a_h = ((smem_s + (uint32_t)g * gbytes) >> 4) | (1u << 16), a_l = a_h + (CHUNK >> 4);
            uint32_t acc = 0u;
            for (int qn = 0; qn < m.nchunks; ++qn) {
#pragma unroll
              for (int s4 = 0; s4 < 4; ++s4) {
                umma_bf16(t_acc, desc64(a_h, a_hi32), desc64(b_h, b_hi32), idesc, acc);
                umma_bf16(t_acc, desc64(a_h, a_hi32), desc64(b_l, b_hi32), idesc, 1u);
                umma_bf16(t_acc, desc64(a_l, a_hi32), desc64(b_h, b_hi32), idesc, 1u);
                acc = 1u;
                a_h += 2u; a_l += 2u; b_h += kinc; b_l += kinc;
              }
              a_h += (2 * CHUNK >> 4) - 8u; a_l += (2 * CHUNK >> 4) - 8u;
            }
            // relation slab pair (no swizzle: K-slabs 2048 B apart, 8-row groups 128 B apart)
            const uint32_t r_h = (((smem_s + (uint32_t)g * gbytes + (uint32_t)(m.nchunks * 2 * CHUNK)) >> 4) & 0x3fffu) | ((2048u >> 4) << 16);
            const uint32_t r_l = r_h + (REL_PLANE >> 4);
            umma_bf16(t_acc, desc64(r_h, b_hi32), desc64(b_h, b_hi32), idesc, 1u);
            umma_bf16(t_acc, desc64(r_h, b_hi32), desc64(b_l, b_hi32), idesc, 1u);
            umma_bf16(t_acc, desc64(r_l, b_hi32), desc64(b_h, b_hi32), idesc, 1u);
          } else {
            uint32_t t_ahi = t_acc + (uint32_t)m.acc_cols, t_alo = t_ahi + (uint32_t)m.ta_cols;
            const int ksteps = K / 16;
            umma_bf16_ts(t_acc, t_ahi, desc64(b_h, b_hi32), idesc, 0u);
            umma_bf16_ts(t_acc, t_ahi, desc64(b_l, b_hi32), idesc, 1u);
            umma_bf16_ts(t_acc, t_alo, desc64(b_h, b_hi32), idesc, 1u);
            for (int ks = 1; ks < ksteps; ++ks) {
              t_ahi += 8u; t_alo += 8u; b_h += kinc; b_l += kinc;
              umma_bf16_ts(t_acc, t_ahi, desc64(b_h, b_hi32), idesc, 1u);
              umma_bf16_ts(t_acc, t_ahi, desc64(b_l, b_hi32), idesc, 1u);
              umma_bf16_ts(t_acc, t_alo, desc64(b_h, b_hi32), idesc, 1u);
            }
          }
          umma_commit(bar_acc0 + 8 * g);
        }
        __syncwarp();
        if (iprof) { const long long t = clock64(); a.prof[97 + (l < 3 ? l : 3)] += t - ti; a.prof[101] += 1; ti = t; }   // time spent issuing layer l
      }
      if (any) idle = 0;
      else if (++idle > (1u << 26)) __trap();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == NW) tmem_dealloc(tmem_base, (uint32_t)m.tmem_alloc);
}

constexpr size_t SMEM_CAP2 = 227 * 1024;

static size_t smem_bytes2(const Plan &m, int groups) {
  const size_t g = (size_t)m.nchunks * 2 * CHUNK + 2 * REL_PLANE;
  return 1024 + groups * g + ((m.wbytes + 127) & ~127) + (size_t)((m.bfloats + 31) & ~31) * 4 + groups * 128 * 4 + 2 * MAXG * 8 + 64;
}

// derived layout; false when the chain / channel count is not one this kernel handles (caller falls back to tc_mlp.cu)
static bool make_plan(Plan &m, int mode, int64_t C) {
  if (C <= 0 || C % 64 != 0 || C > 256) return false;
  if (m.num_layers < 2 || m.num_layers > MAXL) return false;
  if (m.k[0] != C + 16) return false;
  m.nchunks = (int)(C / 64);
  int nmax = 0, kin = 0, wb = 0, bf = 0;
  for (int l = 0; l < m.num_layers; ++l) {
    if (m.n[l] % 16 != 0 || m.n[l] > 256 || m.k[l] % 16 != 0) return false;
    if (l > 0 && m.k[l] != m.n[l - 1]) return false;
    if (m.n[l] > nmax) nmax = m.n[l];
    if (l > 0 && m.k[l] > kin) kin = m.k[l];
    m.woff[l] = wb; wb += m.k[l] * m.n[l] * 4;
    m.boff[l] = bf; bf += m.n[l];
  }
  m.wbytes = wb; m.bfloats = bf;
  // TMEM column granularity of the A planes: 32 (SA1: 128 columns per group -> 4 groups of 128 threads, measured best: 0.250 ms);
  // MVPNET_B200_TC2_TA_ALIGN=16 packs the hi / lo planes back to back (SA1: 96 columns -> 5 groups, 0.263 ms; results identical)
  static const int ta_align = getenv("MVPNET_B200_TC2_TA_ALIGN") ? atoi(getenv("MVPNET_B200_TC2_TA_ALIGN")) : 32;
  const int al = ta_align == 32 ? 32 : 16;
  m.acc_cols = (nmax + 31) & ~31;
  m.ta_cols = ((kin / 2) + al - 1) & ~(al - 1);
  m.part_cols = mode == MODE_FA ? ((m.n[m.num_layers - 1] + 31) & ~31) : 0;
  m.tg_cols = m.acc_cols + 2 * m.ta_cols + m.part_cols;
  for (m.groups = MAXG; m.groups >= 1; --m.groups)
    if (m.groups * m.tg_cols <= 512 && smem_bytes2(m, m.groups) <= SMEM_CAP2) break;
  if (m.groups < 1) return false;
  return true;
}

template <int MODE, int NG, int GT>
static int launch_ng(const Args &a, Plan m, long long tiles, cudaStream_t stream) {
  auto kern = tc2_kernel<MODE, NG, GT>;
  m.groups = NG;
  m.tmem_alloc = 32;
  while (m.tmem_alloc < NG * m.tg_cols) m.tmem_alloc <<= 1;
  const size_t smem = smem_bytes2(m, NG);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("tc2: smem attribute (%zu B): %s", smem, cudaGetErrorString(e)); return (int)e; }
  long long grid = sm_count();
  const long long want = (tiles + NG - 1) / NG;
  if (grid > want) grid = want;
  static const bool debug = getenv("MVPNET_B200_DEBUG") != nullptr;
  if (debug)
    fprintf(stderr, "[tc2 mode=%d] tiles=%lld groups=%d x %d threads grid=%lld smem=%zu nchunks=%d tmem=%d/%d wbytes=%d\n", MODE, tiles, NG, GT, grid, smem,
            m.nchunks, m.tg_cols, m.tmem_alloc, m.wbytes);
  kern<<<(unsigned)grid, NG * GT + 32, smem, stream>>>(a, m, tiles);
  return launch_status("tc2_fused_mlp");
}

static long long *g_prof = nullptr;

template <int MODE>
static int launch(const Args &a_in, const Plan &m, long long tiles, cudaStream_t stream) {
  Args a = a_in;
  static const bool want_prof = getenv("MVPNET_B200_TC2_PROF") != nullptr;
  if (want_prof) {                 // debug only: phase clocks of CTA 0, printed and reset by mvp_tc2_prof_dump()
    if (g_prof == nullptr) { cudaMalloc(&g_prof, 128 * sizeof(long long)); cudaMemset(g_prof, 0, 128 * sizeof(long long)); }
    a.prof = g_prof;
  }
  // m.groups = the most tile groups TMEM and shared memory allow (<= 6).  Up to 3 groups run 256 threads each; 4..6 groups run
  // 128 threads each (MVPNET_B200_TC2_GROUPS caps the count: <= 3 selects the wide groups).
  int ng = m.groups;
  const long long per_sm = tiles / sm_count();
  if (per_sm < ng) ng = per_sm < 1 ? 1 : (int)per_sm;
  static const char *force = getenv("MVPNET_B200_TC2_GROUPS");
  if (force && atoi(force) >= 1 && atoi(force) < ng) ng = atoi(force);
  switch (ng) {
    case 6: return launch_ng<MODE, 6, 128>(a, m, tiles, stream);
    case 5: return launch_ng<MODE, 5, 128>(a, m, tiles, stream);
    case 4: return launch_ng<MODE, 4, 128>(a, m, tiles, stream);
    case 3: return launch_ng<MODE, 3, 256>(a, m, tiles, stream);
    case 2: return launch_ng<MODE, 2, 256>(a, m, tiles, stream);
    default: return launch_ng<MODE, 1, 256>(a, m, tiles, stream);
  }
}

static int to_plan(const mvp_tc_chain_t *c, int mode, int64_t C, Plan *m) {
  MVP_REQUIRE(c, MVP_ERR_NULL, "tc2: null chain");
  MVP_REQUIRE(c->num_layers >= 1 && c->num_layers <= MAXL, MVP_ERR_INVALID_ARG, "tc2: 1..6 layers");
  m->num_layers = c->num_layers;
  m->out_channels = c->out_channels;
  for (int l = 0; l < c->num_layers; ++l) {
    m->k[l] = c->k[l]; m->n[l] = c->n[l]; m->relu[l] = c->relu[l];
    m->w_hi[l] = (const __nv_bfloat16 *)c->w_hi[l]; m->w_lo[l] = (const __nv_bfloat16 *)c->w_lo[l]; m->bias[l] = c->bias[l];
  }
  MVP_REQUIRE(make_plan(*m, mode, C), MVP_ERR_UNSUPPORTED, "tc2: chain / channel count not supported (C=%lld)", (long long)C);
  for (int l = 0; l < c->num_layers; ++l) {
    MVP_REQUIRE(c->w_hi[l] && c->w_lo[l] && c->bias[l], MVP_ERR_NULL, "tc2: null weights");
    MVP_REQUIRE((((uintptr_t)c->w_hi[l] | (uintptr_t)c->w_lo[l] | (uintptr_t)c->bias[l]) & 15) == 0, MVP_ERR_INVALID_ARG,
                "tc2: weights must be 16-byte aligned");
  }
  MVP_REQUIRE(c->out_channels == c->n[c->num_layers - 1], MVP_ERR_UNSUPPORTED, "tc2: out_channels must equal the last layer's padded width");
  return 0;
}

}  // namespace tc2
}  // namespace mvp

extern "C" void mvp_tc2_prof_dump(const char *tag) {
  if (mvp::tc2::g_prof == nullptr) return;
  long long h[128];
  cudaDeviceSynchronize();
  cudaMemcpy(h, mvp::tc2::g_prof, sizeof(h), cudaMemcpyDeviceToHost);
  cudaMemset(mvp::tc2::g_prof, 0, sizeof(h));
  fprintf(stderr, "[tc2 prof %s] (cycles of CTA 0, summed over launches)\n", tag);
  for (int g = 0; g < 6; ++g)
    if (h[g * 16 + 1]) fprintf(stderr, "  group %d: wait_gather %lld  wait_L0 %lld  build %lld  epi0 %lld  wait_L1 %lld  epi1 %lld  wait_L2 %lld  epi2 %lld  wait_L3 %lld epi3 %lld\n", g,
            h[g * 16], h[g * 16 + 1], h[g * 16 + 2], h[g * 16 + 3], h[g * 16 + 4], h[g * 16 + 5], h[g * 16 + 6], h[g * 16 + 7], h[g * 16 + 8], h[g * 16 + 9]);
  fprintf(stderr, "  issuer: polling %lld  issue_L0 %lld  issue_L1 %lld  issue_L2 %lld  issue_L3+ %lld  requests %lld\n", h[96], h[97], h[98], h[99], h[100], h[101]);
}

extern "C" int mvp_tc2_supported(const mvp_tc_chain_t *c, int mode, int64_t C) {
  mvp::tc2::Plan m;
  if (!c || c->num_layers < 1 || c->num_layers > mvp::tc2::MAXL) return 0;
  m.num_layers = c->num_layers;
  m.out_channels = c->out_channels;
  for (int l = 0; l < c->num_layers; ++l) { m.k[l] = c->k[l]; m.n[l] = c->n[l]; }
  if (c->out_channels != c->n[c->num_layers - 1]) return 0;
  return mvp::tc2::make_plan(m, mode, C) ? 1 : 0;
}

extern "C" int mvp_tc2_set_abstraction(const void *feat_hi, const void *feat_lo, int64_t C, const float *xyz, const float *new_xyz,
                                       const int64_t *nbr, int64_t B, int64_t N, int64_t M, int64_t K, const mvp_tc_chain_t *chain,
                                       float *out_f32, void *out_hi, void *out_lo, mvp_stream_t stream) {
  using namespace mvp;
  MVP_REQUIRE(K == 32, MVP_ERR_UNSUPPORTED, "tc2_set_abstraction: max_neighbors must be 32");
  MVP_REQUIRE(B >= 0 && N > 0 && M >= 0, MVP_ERR_INVALID_ARG, "tc2_set_abstraction: bad sizes");
  MVP_REQUIRE(B * N < (1LL << 31) && B * M * 32 < (1LL << 31), MVP_ERR_UNSUPPORTED, "tc2_set_abstraction: more than 2^31 rows");
  tc2::Plan m;
  if (int rc = tc2::to_plan(chain, MODE_SA, C, &m)) return rc;
  if (B * M == 0) return 0;
  MVP_REQUIRE(feat_hi && feat_lo && xyz && new_xyz && nbr && (out_f32 || out_hi) && (!out_hi == !out_lo), MVP_ERR_NULL,
              "tc2_set_abstraction: null pointer");
  MVP_REQUIRE((((uintptr_t)feat_hi | (uintptr_t)feat_lo) & 15) == 0, MVP_ERR_INVALID_ARG, "tc2_set_abstraction: features must be 16-byte aligned");
  tc2::Args a = {};
  a.rows_out = B * M; a.src_hi = (const __nv_bfloat16 *)feat_hi; a.src_lo = (const __nv_bfloat16 *)feat_lo; a.C = (int)C;
  a.xyz = xyz; a.new_xyz = new_xyz; a.nbr = nbr; a.n_src = (unsigned)N; a.n_out = (unsigned)M; a.k = 32;
  a.out_f32 = out_f32; a.out_hi = (__nv_bfloat16 *)out_hi; a.out_lo = (__nv_bfloat16 *)out_lo;
  return tc2::launch<MODE_SA>(a, m, (a.rows_out + 3) / 4, (cudaStream_t)stream);
}

extern "C" int mvp_tc2_feature_aggregation(const void *pix_hi, const void *pix_lo, int64_t C, int64_t nv, int64_t h, int64_t w, int64_t hp,
                                           int64_t wp, const float *pix_xyz, const float *points, const int64_t *knn, int64_t B, int64_t Np,
                                           int64_t K, int reduce_sum, const mvp_tc_chain_t *chain, float *out_f32, void *out_hi, void *out_lo,
                                           mvp_stream_t stream) {
  using namespace mvp;
  MVP_REQUIRE(K >= 1 && K <= 4, MVP_ERR_UNSUPPORTED, "tc2_feature_aggregation: k must be in [1, 4]");
  MVP_REQUIRE(B >= 0 && Np >= 0 && nv > 0 && h > 0 && w > 0 && hp >= h && wp >= w, MVP_ERR_INVALID_ARG, "tc2_feature_aggregation: bad sizes");
  MVP_REQUIRE(B * nv * hp * wp < (1LL << 31) && B * Np * K < (1LL << 31), MVP_ERR_UNSUPPORTED, "tc2_feature_aggregation: more than 2^31 rows");
  tc2::Plan m;
  if (int rc = tc2::to_plan(chain, MODE_FA, C, &m)) return rc;
  if (B * Np == 0) return 0;
  MVP_REQUIRE(pix_hi && pix_lo && pix_xyz && points && knn && (out_f32 || out_hi) && (!out_hi == !out_lo), MVP_ERR_NULL,
              "tc2_feature_aggregation: null pointer");
  MVP_REQUIRE((((uintptr_t)pix_hi | (uintptr_t)pix_lo | (uintptr_t)out_f32 | (uintptr_t)out_hi | (uintptr_t)out_lo) & 15) == 0, MVP_ERR_INVALID_ARG,
              "tc2_feature_aggregation: pointers must be 16-byte aligned");
  MVP_REQUIRE((((uintptr_t)out_f32 | (uintptr_t)out_hi | (uintptr_t)out_lo) & 31) == 0, MVP_ERR_INVALID_ARG,
              "tc2_feature_aggregation: outputs must be 32-byte aligned (256-bit stores)");
  tc2::Args a = {};
  a.rows_out = B * Np; a.src_hi = (const __nv_bfloat16 *)pix_hi; a.src_lo = (const __nv_bfloat16 *)pix_lo; a.C = (int)C;
  a.xyz = pix_xyz; a.new_xyz = points; a.nbr = knn; a.n_src = (unsigned)(nv * h * w); a.n_out = (unsigned)Np; a.k = (int)K;
  a.reduce = reduce_sum ? REDUCE_SUM : REDUCE_MAX;
  a.hw = (unsigned)(h * w); a.w = (unsigned)w; a.hp_wp = (unsigned)(hp * wp); a.wp = (unsigned)wp; a.nv = (unsigned)nv;
  a.out_f32 = out_f32; a.out_hi = (__nv_bfloat16 *)out_hi; a.out_lo = (__nv_bfloat16 *)out_lo;
  return tc2::launch<MODE_FA>(a, m, (a.rows_out + 127) / 128, (cudaStream_t)stream);
}
