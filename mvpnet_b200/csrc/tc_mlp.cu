// Tensor-core (tcgen05 / TMEM) version of the fused "build rows -> MLP chain -> reduce" kernel, sm_100a.
//
// Same contract as fused_mlp.cu (MODE_SA / MODE_FA / MODE_FP, reference lines cited there); this file
// moves the 1x1-conv chain onto the 5th-generation tensor cores:
//
//   * tile = 128 rows (UMMA M = 128, cta_group::1): SA 4 centroids x 32 neighbours, FA 32 points x 3
//     pixels (+32 idle rows), FP 128 points.  Accumulators live in TMEM (128 lanes x Cout columns fp32).
//   * precision: every fp32 operand is split into bf16 hi + bf16 lo and each K-step issues THREE
//     tcgen05.mma.kind::f16 into the same accumulator: a_hi*w_hi + a_hi*w_lo + a_lo*w_hi (fp32 accumulate).
//     On the reference modules this reproduces the golden logits to 1.6e-5 (single-pass bf16/tf32: 4e-3..8e-3;
//     the bar is 1e-4) at twice the MMA rate and half the operand bytes of a 3xTF32 scheme.
//   * operands are K-major, no-swizzle canonical UMMA layout: 8-row x 16-byte core matrices, a K-slab
//     (8 channels x all rows) is contiguous, so the epilogue of layer l writes layer l+1's A operand with
//     conflict-free 16-byte stores (one thread = one row = one TMEM lane), in place.
//   * warp roles.  A CTA runs NG (1..3) independent TILE GROUPS of 8 worker warps; each group owns an
//     activation tile in shared memory and a TMEM accumulator and walks its own tiles: build rows -> [MMA] ->
//     epilogue -> [MMA] -> ... .  One extra warp is the MMA ISSUER (one elected lane serves the groups' requests
//     round-robin) and one is the WEIGHT PRODUCER (bulk async copies global -> shared-memory ring, running ahead of
//     the issuer across layers, groups and tiles).  Hand-offs are mbarriers: a_ready[g] (256 worker arrivals ->
//     issuer), acc_ready[g] (tcgen05.commit -> workers), full/empty per ring stage.  While one group's MMAs run,
//     the other groups build / run epilogues, so the tensor pipe and the CUDA cores overlap inside one CTA and
//     the weights (resident for narrow chains, ring otherwise) are shared by the groups.
//   * SA epilogue: one warp reads the 32 TMEM lanes of one centroid (tcgen05.ld 32x32b), so the max over
//     the K = 32 neighbours is a halving butterfly of warp shuffles.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"

namespace mvp {
namespace tc {

constexpr int ROWS = 128;
constexpr int GROUP_THREADS = 256;   // 8 worker warps per tile group
constexpr int MAX_GROUPS = 3;
constexpr int CTRL_THREADS = 64;     // warp 8*NG: MMA issuer (+ TMEM owner), warp 8*NG+1: weight producer
constexpr int MAX_LAYERS = 6;
constexpr int MAX_STAGES = 6;
constexpr int SLAB = ROWS * 16;      // bytes of one K-slab (8 channels) of the A operand

struct Chain {
  int num_layers;
  int k[MAX_LAYERS];      // padded to a multiple of 16
  int n[MAX_LAYERS];      // padded to a multiple of 16, <= 512; k[l+1] == n[l]
  int relu[MAX_LAYERS];
  const __nv_bfloat16 *w_hi[MAX_LAYERS];  // per 256-wide N block: [k/8][nb][8]
  const __nv_bfloat16 *w_lo[MAX_LAYERS];
  const float *bias[MAX_LAYERS];          // [n]
  int out_channels;
  int kc;                 // K elements per weight chunk (16 or 32)
  int kmax;               // max k over layers (capacity of a group's activation tile)
  int nbmax;              // max N-block width over layers (<= 256)
  int tmem_cols;          // per group: power of two >= max n, >= 32
  int tmem_alloc;         // power of two >= groups * tmem_cols
  int resident;           // 1: all weights live in shared memory for the whole kernel (narrow chains)
  int stages;             // ring depth when not resident (2..MAX_STAGES)
  int groups;             // tile groups per CTA (1..MAX_GROUPS)
  int res_off[MAX_LAYERS];  // resident mode: byte offset of layer l's hi block (lo follows at + k*n*2)
  int res_bytes;
  int panel;              // K elements of the first layer built per pass (== k[0] unless the row is too wide for shared memory)
  int bias_off[MAX_LAYERS];  // float offset of layer l's bias in the shared-memory copy
  int bias_floats;
};

// ---- PTX wrappers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Polls with test_wait (returns at once) rather than try_wait (suspends the thread for a hardware-chosen quantum: the
// wake-up after a barrier that completes mid-suspension cost ~500 cycles per blocking wait in the convolution's
// stage ring).  MVPNET_B200_TRYWAIT=1 at build time restores try_wait.  A protocol error traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok, spins = 0;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
#ifdef MVPNET_B200_TRYWAIT
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
#else
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
#endif
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!ok && ++spins > (1u << 26)) __trap();
  } while (!ok);
}
// Worker-side wait on the accumulator: try_wait suspends the warp in hardware.  The fused kernels are issue-bound (ncu,
// round 2: a third of all issued warp instructions were test_wait spins of idle tile groups), so a waiting group must
// not compete for issue slots with the groups that have work.
__device__ __forceinline__ void mbar_wait_suspend(uint32_t bar, uint32_t parity) {
  uint32_t ok, spins = 0;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(20000u)      // suspend-time hint (ns)
        : "memory");
    if (!ok && ++spins > (1u << 20)) __trap();
  } while (!ok);
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// 1-D bulk async copy global -> shared (TMA unit, no tensor map); completion is signalled on the mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// one lane of the (converged) warp; ptxas then knows the guarded tcgen05 / bulk-copy instructions have one issuer
__device__ __forceinline__ bool elect_one() {
  uint32_t p;
  asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}" : "=r"(p));
  return p != 0;
}
__device__ __forceinline__ void group_sync(int g) {   // the 256 workers of tile group g (barrier 0 is __syncthreads)
  asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "n"(GROUP_THREADS) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t cols) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {  // the same warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}

// K-major, no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1 = sm100):
// start address >> 4 [0,14), leading (K-direction) byte offset >> 4 [16,30), stride (M/N-direction) byte
// offset >> 4 [32,46), version [46,48), layout type [61,64) = 0.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 [4,6)=1, A = B = BF16 [7,10),[10,13)=1,
// both K-major, N>>3 at [17,23), M>>4 at [24,29).
__device__ __forceinline__ uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// TMEM -> registers, 16 consecutive columns of this warp's 32 lanes.  Issue and wait are separate so that the load
// of the next chunk is in flight while the current one is processed; the wait names the registers as read-write
// operands so that no use of them can be scheduled above it.
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}

// ---- operand packing ---------------------------------------------------------------------------
// two fp32 values -> packed bf16 hi pair and packed bf16 lo pair (lo = rn(v - hi)), 5 instructions per pair:
// one packing convert, two integer ops to re-expand hi, one packed fp32 subtract (FADD2), one packing convert
// one 32-byte store (STG.256): a thread that owns 32 contiguous bytes must write them with ONE instruction — two 16-byte
// stores reach L2 as two half-written sectors each, and the sector-write rate of L2 is what bounds a store-heavy epilogue
__device__ __forceinline__ void st_global_256(void *p, uint4 a, uint4 b) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y),
               "r"(b.z), "r"(b.w)
               : "memory");
}

__device__ __forceinline__ void split_pair(float v0, float v1, uint32_t &h, uint32_t &l) {
  const __nv_bfloat162 hh = __floats2bfloat162_rn(v0, v1);
  h = *reinterpret_cast<const uint32_t *>(&hh);
  uint64_t vv, hf, d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(vv) : "f"(v0), "f"(v1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(hf) : "r"(h << 16), "r"(h & 0xffff0000u));
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(vv), "l"(hf));
  float d0, d1;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d));
  const __nv_bfloat162 ll = __floats2bfloat162_rn(d0, d1);
  l = *reinterpret_cast<const uint32_t *>(&ll);
}

__device__ __forceinline__ void add_pair(float &a, float &b, float x, float y) {   // packed fp32 add (FADD2), rn like the scalar add
  uint64_t u, w;
  asm("mov.b64 %0, {%1, %2};" : "=l"(u) : "f"(a), "f"(b));
  asm("mov.b64 %0, {%1, %2};" : "=l"(w) : "f"(x), "f"(y));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(u) : "l"(u), "l"(w));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(u));
}

// 8 consecutive channels of one row -> one 16-byte unit of the hi operand and one of the lo operand
__device__ __forceinline__ void store8(unsigned char *a_hi, unsigned char *a_lo, int row, int kb, const float (&v)[8]) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) split_pair(v[2 * i], v[2 * i + 1], h[i], l[i]);
  const size_t off = (size_t)kb * SLAB + (size_t)(row >> 3) * 128 + (size_t)(row & 7) * 16;
  *reinterpret_cast<uint4 *>(a_hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4 *>(a_lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
}

__device__ __forceinline__ void load8(const float *p, bool ok, float (&v)[8]) {
  if (ok) {
    const float4 a = __ldg(reinterpret_cast<const float4 *>(p)), b = __ldg(reinterpret_cast<const float4 *>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0.f;
  }
}

// ---- row builders ---------------------------------------------------------------------------------
// Phase A: one thread per row resolves the row's source address(es) and its relation / weight scalars into
// shared memory (coalesced index reads).  Phase B: the tile is cut into (row, 8-channel) units, row-fastest
// so that consecutive lanes write consecutive 16-byte slots of one K-slab (conflict-free); every thread
// issues the global loads of several units before it converts and stores any of them, which is what hides
// the gather latency (the two dependent loads index -> row used to be serialised per row).
struct Aux {                 // per-group, per-tile scratch in shared memory
  long long src[ROWS * 3];   // element offset of the source row(s); < 0 = no source (zero row)
  float w[ROWS * 3];         // FP: interpolation weights
  float rel[ROWS * 4];       // SA: xyz - centroid; FA: dx, dy, dz, |d|^2
};

__device__ __forceinline__ void zero8(float (&v)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = 0.f;
}

template <int MODE>
__device__ __forceinline__ void build_rows(const BuildArgs &a, Aux &x, unsigned char *a_hi, unsigned char *a_lo, int kb_begin,
                                           int kblocks, long long tile, int tid /* within the group */, int group) {
  // ---------------- phase A (once per tile: the scratch survives the K panels of a wide first layer)
  if (kb_begin == 0 && tid < ROWS) {
    const int r = tid;
    if (MODE == MODE_SA) {
      const long long gid = tile * 4 + (r >> 5);
      long long j = -1, b = 0;
      if (gid < a.rows_out) { b = gid / a.n_out; j = a.nbr[gid * 32 + (r & 31)]; }
      const bool ok = j >= 0 && j < a.n_src;
      x.src[r] = ok ? b * a.n_src + j : -1;
#pragma unroll
      for (int i = 0; i < 3; ++i)
        x.rel[r * 4 + i] = ok ? __fsub_rn(__ldg(a.xyz + ((size_t)b * a.n_src + j) * 3 + i), __ldg(a.new_xyz + gid * 3 + i)) : 0.f;
      x.rel[r * 4 + 3] = 0.f;
    } else if (MODE == MODE_FA) {
      const int i = r >> 5, p = r & 31;
      const long long pid = tile * 32 + p;
      long long j = -1, b = 0;
      if (pid < a.rows_out && i < a.k) { b = pid / a.n_out; j = a.nbr[pid * a.k + i]; }
      const bool ok = j >= 0 && j < a.n_src;
      float d[4] = {0.f, 0.f, 0.f, 0.f};
      long long off = -1;
      if (ok) {
        const unsigned ju = (unsigned)j;                       // n_src = nv * h * w < 2^31 (checked by the caller)
        const unsigned v_ = ju / (unsigned)a.hw, pix = ju - v_ * (unsigned)a.hw;
        const unsigned y = pix / (unsigned)a.w, xx = pix - y * (unsigned)a.w;
        off = ((long long)b * a.nv + v_) * a.s_n + (long long)y * a.s_h + (long long)xx * a.s_w;
        const float *s = a.xyz + ((size_t)b * a.n_src + j) * 3;
        const float *t = a.new_xyz + pid * 3;
        d[0] = __fsub_rn(__ldg(s), __ldg(t)); d[1] = __fsub_rn(__ldg(s + 1), __ldg(t + 1)); d[2] = __fsub_rn(__ldg(s + 2), __ldg(t + 2));
        d[3] = __fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2]));
      }
      x.src[r] = off;
      *reinterpret_cast<float4 *>(&x.rel[r * 4]) = make_float4(d[0], d[1], d[2], d[3]);
    } else {
      const long long pid = tile * ROWS + r;
      const bool live = pid < a.rows_out;
      const long long b = live ? pid / a.n_out : 0;
      float w[3] = {0.f, 0.f, 0.f};
      long long j[3] = {-1, -1, -1};
      if (live) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          j[k] = a.nbr[pid * 3 + k];
          w[k] = __fdiv_rn(1.0f, fmaxf(__ldg(a.dist + pid * 3 + k), a.eps));
          if (j[k] < 0 || j[k] >= a.n_src) { j[k] = -1; w[k] = 0.f; }
        }
        const float norm = __fadd_rn(__fadd_rn(w[0], w[1]), w[2]);
#pragma unroll
        for (int k = 0; k < 3; ++k) w[k] = __fdiv_rn(w[k], norm);
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) { x.src[r * 3 + k] = j[k] >= 0 ? b * a.n_src + j[k] : -1; x.w[r * 3 + k] = w[k]; }
    }
  }
  group_sync(group);
  // ---------------- phase B: a warp covers 8 rows x 4 K-slabs per step: lane = (slab sub-index << 3) | row sub-index.
  // Loads: 8 rows x 128 contiguous bytes (whole lines); stores: per slab 8 rows x 16 B = 128 contiguous bytes
  // (all 32 banks, conflict-free).
  const int warp = tid >> 5, lane = tid & 31, rs = lane & 7, ks = lane >> 3;
  constexpr int NWARP = GROUP_THREADS / 32;
  // FA: rows 32*k .. 127 carry no pixel (k < 4): 4 row blocks of 8 per live 32-row quarter
  const int rblocks = MODE == MODE_FA ? 4 * a.k : 16;
  const int nblk = rblocks * ((kblocks + 3) >> 2);
  if (MODE == MODE_FP) {
    const int Cs = a.feat_channels, Cd = a.skip_channels, sb = Cs >> 3, db = Cd >> 3;
    constexpr int U = 2;
    for (int wb0 = warp; wb0 < nblk; wb0 += NWARP * U) {
      float v0[U][8], v1[U][8], v2[U][8];
#pragma unroll
      for (int i = 0; i < U; ++i) {
        const int wb = wb0 + i * NWARP;
        const int r = (wb & 15) * 8 + rs, kl = (wb >> 4) * 4 + ks, kb = kb_begin + kl;
        zero8(v0[i]); zero8(v1[i]); zero8(v2[i]);
        if (wb < nblk && kl < kblocks) {
          if (kb < sb) {
            const long long s0 = x.src[r * 3], s1 = x.src[r * 3 + 1], s2 = x.src[r * 3 + 2];
            load8(a.feat + (size_t)(s0 < 0 ? 0 : s0) * Cs + kb * 8, s0 >= 0, v0[i]);
            load8(a.feat + (size_t)(s1 < 0 ? 0 : s1) * Cs + kb * 8, s1 >= 0, v1[i]);
            load8(a.feat + (size_t)(s2 < 0 ? 0 : s2) * Cs + kb * 8, s2 >= 0, v2[i]);
          } else if (kb < sb + db) {
            const long long pid = tile * ROWS + r;
            load8(a.skip + (size_t)(pid < a.rows_out ? pid : 0) * Cd + (kb - sb) * 8, pid < a.rows_out, v0[i]);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < U; ++i) {
        const int wb = wb0 + i * NWARP;
        const int r = (wb & 15) * 8 + rs, kl = (wb >> 4) * 4 + ks, kb = kb_begin + kl;
        if (wb < nblk && kl < kblocks) {
          if (kb < sb) {
            const float w0 = x.w[r * 3], w1 = x.w[r * 3 + 1], w2 = x.w[r * 3 + 2];
#pragma unroll
            for (int c = 0; c < 8; ++c)  // interpolate_kernel.cu:54-61 accumulation order
              v0[i][c] = __fmaf_rn(v2[i][c], w2, __fmaf_rn(v1[i][c], w1, __fmul_rn(v0[i][c], w0)));
          }
          store8(a_hi, a_lo, r, kl, v0[i]);
        }
      }
    }
  } else {
    const int cb = a.feat_channels >> 3;
    constexpr int U = 3;
    for (int wb0 = warp; wb0 < nblk; wb0 += NWARP * U) {
      float v[U][8];
#pragma unroll
      for (int i = 0; i < U; ++i) {
        const int wb = wb0 + i * NWARP;
        const int rb = MODE == MODE_FA ? wb % rblocks : (wb & 15), kq = MODE == MODE_FA ? wb / rblocks : (wb >> 4);
        const int r = rb * 8 + rs, kl = kq * 4 + ks, kb = kb_begin + kl;
        zero8(v[i]);
        if (wb < nblk && kl < kblocks) {
          const long long s0 = x.src[r];
          if (kb < cb) {
            if (MODE == MODE_SA) {
              load8(a.feat + (size_t)(s0 < 0 ? 0 : s0) * a.feat_channels + kb * 8, s0 >= 0, v[i]);
            } else if (a.s_c == 1) {
              load8(a.feat + (s0 < 0 ? 0 : s0) + kb * 8, s0 >= 0, v[i]);
            } else if (s0 >= 0) {
              const float *p = a.feat + s0 + (long long)(kb * 8) * a.s_c;
#pragma unroll
              for (int c = 0; c < 8; ++c) v[i][c] = __ldg(p + (long long)c * a.s_c);
            }
          } else if (kb == cb) {
            const float4 q = *reinterpret_cast<const float4 *>(&x.rel[r * 4]);
            v[i][0] = q.x; v[i][1] = q.y; v[i][2] = q.z; v[i][3] = q.w;
          }
        }
      }
#pragma unroll
      for (int i = 0; i < U; ++i) {
        const int wb = wb0 + i * NWARP;
        const int rb = MODE == MODE_FA ? wb % rblocks : (wb & 15), kq = MODE == MODE_FA ? wb / rblocks : (wb >> 4);
        const int r = rb * 8 + rs, kl = kq * 4 + ks;
        if (wb < nblk && kl < kblocks) store8(a_hi, a_lo, r, kl, v[i]);
      }
    }
  }
}

// ---- the kernel ----------------------------------------------------------------------------------
__device__ __forceinline__ void issue_kstep(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t w_hi, uint32_t w_lo, uint32_t nb,
                                            uint32_t idesc, bool first) {
  const uint64_t ah = make_desc(a_hi, SLAB, 128), al = make_desc(a_lo, SLAB, 128);
  const uint64_t wh = make_desc(w_hi, nb * 16, 128), wl = make_desc(w_lo, nb * 16, 128);
  umma_bf16(d_tmem, ah, wh, idesc, first ? 0u : 1u);
  umma_bf16(d_tmem, ah, wl, idesc, 1u);
  umma_bf16(d_tmem, al, wh, idesc, 1u);
}

// FA staging of the last layer: [pixel slot i][column c][point p], point stride padded to 33 words so that both the
// row-owner writes (lanes = points) and the reducing reads (lanes = columns) are bank-conflict free
constexpr int FA_PSTRIDE = 33;

template <int MODE, int NG>
// single-group CTAs are compiled for TWO per SM (<= 96 registers): two independent CTAs de-phase naturally, so one's MMA
// phase overlaps the other's build / epilogue (the groups of one CTA advance in lock step around the shared weight ring)
__global__ void __launch_bounds__(NG *GROUP_THREADS + CTRL_THREADS, NG == 1 ? 2 : 1)
tc_fused_mlp_kernel(const BuildArgs a, const Chain m, float *__restrict__ out, long long num_tiles) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int NW = NG * (GROUP_THREADS / 32);              // worker warps
  constexpr int NTHREADS = NG * GROUP_THREADS + CTRL_THREADS;
  const size_t a_bytes = (size_t)(m.kmax >> 3) * SLAB;       // one of {hi, lo} of one group's activation tile
  const size_t stage_half = (size_t)(m.kc >> 3) * m.nbmax * 16;  // one of {hi, lo} of one ring stage
  unsigned char *wreg = smem + (size_t)NG * 2 * a_bytes;     // resident weights, or the ring [stage][hi|lo]
  const size_t wreg_bytes = m.resident ? (size_t)m.res_bytes : (size_t)m.stages * 2 * stage_half;
  constexpr size_t AUX_BYTES = (sizeof(Aux) + 127) & ~(size_t)127;
  unsigned char *aux_base = wreg + ((wreg_bytes + 127) & ~(size_t)127);
  uint64_t *bars = reinterpret_cast<uint64_t *>(aux_base + NG * AUX_BYTES);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * MAX_STAGES + 2 * MAX_GROUPS);
  // biases in shared memory (round 2: the per-chunk __ldg of the bias was 5 % of FP4's stall samples, all long-scoreboard)
  float *bias_s = reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(bars) + 256);
  const uint32_t bar_full0 = smem_u32(bars), bar_empty0 = smem_u32(bars + MAX_STAGES);
  const uint32_t bar_aready0 = smem_u32(bars + 2 * MAX_STAGES), bar_acc0 = smem_u32(bars + 2 * MAX_STAGES + MAX_GROUPS);

  if (tid == 0) {
    for (int s = 0; s < MAX_STAGES; ++s) { mbar_init(bar_full0 + 8 * s, 1); mbar_init(bar_empty0 + 8 * s, 1); }
    for (int g = 0; g < MAX_GROUPS; ++g) { mbar_init(bar_aready0 + 8 * g, GROUP_THREADS); mbar_init(bar_acc0 + 8 * g, 1); }
    fence_barrier_init();
  }
  if (warp == NW) tmem_alloc(smem_u32(tmem_slot), (uint32_t)m.tmem_alloc);
  for (int l = 0; l < m.num_layers; ++l)
    for (int o = tid; o < m.n[l]; o += NTHREADS) bias_s[m.bias_off[l] + o] = __ldg(m.bias[l] + o);
  if (m.resident) {  // one cooperative copy of every layer's W_hi | W_lo for the lifetime of the CTA
    for (int l = 0; l < m.num_layers; ++l) {
      const size_t bytes = (size_t)m.k[l] * m.n[l] * 2;
      const uint4 *gh = reinterpret_cast<const uint4 *>(m.w_hi[l]), *gl = reinterpret_cast<const uint4 *>(m.w_lo[l]);
      uint4 *sh = reinterpret_cast<uint4 *>(wreg + m.res_off[l]), *sl = reinterpret_cast<uint4 *>(wreg + m.res_off[l] + bytes);
      for (size_t o = tid; o < bytes / 16; o += NTHREADS) { sh[o] = __ldg(gh + o); sl[o] = __ldg(gl + o); }
    }
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // ---- segments: the first layer is cut into K panels of m.panel channels (one panel unless the built row is too
  //      wide for shared memory); every later layer is one segment.  seg -> (layer, k_begin, k_len)
  const int npanels = (m.k[0] + m.panel - 1) / m.panel;
  const int nsegs = npanels + m.num_layers - 1;
  auto seg_info = [&](int sg, int &l, int &kb, int &kl) {
    if (sg < npanels) { l = 0; kb = sg * m.panel; kl = min(m.panel, m.k[0] - kb); }
    else { l = sg - npanels + 1; kb = 0; kl = m.k[l]; }
  };
  // tiles of this CTA: blockIdx.x + i * gridDim.x, i in [0, n_my); group g owns i == g (mod NG).  The issuer and the
  // producer walk the requests in the fixed order  round -> segment -> group.
  const long long n_my = blockIdx.x < num_tiles ? (num_tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
  const long long rounds = (n_my + NG - 1) / NG;
  const uint32_t S = (uint32_t)m.stages;

  if (warp < NW) {
    // =========================== workers: build rows, run the epilogues =============================================
    const int g = warp / (GROUP_THREADS / 32), ltid = tid - g * GROUP_THREADS, lwarp = ltid >> 5;
    unsigned char *a_hi = smem + (size_t)g * 2 * a_bytes, *a_lo = a_hi + a_bytes;
    Aux &aux = *reinterpret_cast<Aux *>(aux_base + g * AUX_BYTES);
    const uint32_t bar_aready = bar_aready0 + 8 * g, bar_acc = bar_acc0 + 8 * g;
    const uint32_t t_group = tmem_base + (uint32_t)(g * m.tmem_cols);
    const int quarter = lwarp & 3, half = lwarp >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t t_lane = t_group + ((uint32_t)(quarter * 32) << 16);
    const bool live_rows = MODE != MODE_FA || quarter < a.k;     // FA: the rows of quarter >= k carry no pixel
    uint32_t acc_phase = 0;
    for (long long i = g; i < n_my; i += NG) {
      const long long tile = blockIdx.x + i * gridDim.x;
      for (int sg = 0; sg < nsegs; ++sg) {
        int l, kbeg, klen;
        seg_info(sg, l, kbeg, klen);
        const int N = m.n[l];
        const bool last = l == m.num_layers - 1;
        const bool layer_done = sg >= npanels - 1;          // the last panel of layer 0, or any later layer
        if (l == 0) build_rows<MODE>(a, aux, a_hi, a_lo, kbeg >> 3, klen >> 3, tile, ltid, g);
        // the activation tile of this segment is complete (built above, or written by the previous epilogue) and
        // this thread no longer reads the accumulator: hand both to the issuer
        fence_proxy_async();
        tc_fence_before();
        mbar_arrive(bar_aready);
        mbar_wait_suspend(bar_acc, acc_phase);
        acc_phase ^= 1u;
        tc_fence_after();
        if (!layer_done) continue;                          // next K panel of the first layer: rebuild the activation tile

        // ---- epilogue: thread = row (TMEM lane 32 * quarter + lane); the two warps of a quarter take alternate
        //      16-column chunks; the TMEM load of the next chunk is in flight while this one is processed
        float *fbuf = reinterpret_cast<float *>(a_hi);      // MODE_FA last layer: fp32 staging over the (dead) A operand
        if (live_rows) {
          uint32_t rn[16];
          int c = half;
          if (c * 16 < N) tmem_ld16_issue(t_lane + (uint32_t)(c * 16), rn);
          while (c * 16 < N) {
            float v[16];
            tmem_ld_wait(rn);
#pragma unroll
            for (int q = 0; q < 16; ++q) v[q] = __uint_as_float(rn[q]);
            if ((c + 2) * 16 < N) tmem_ld16_issue(t_lane + (uint32_t)((c + 2) * 16), rn);
            const float4 *bp = reinterpret_cast<const float4 *>(bias_s + m.bias_off[l] + c * 16);
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
              float lo8[8], hi8[8];
#pragma unroll
              for (int q = 0; q < 8; ++q) { lo8[q] = v[q]; hi8[q] = v[8 + q]; }
              store8(a_hi, a_lo, row, 2 * c, lo8);
              store8(a_hi, a_lo, row, 2 * c + 1, hi8);
            } else if (MODE == MODE_SA) {
              const long long gid = tile * 4 + quarter;
              // max over the 32 neighbours (= lanes) of 16 columns as a halving butterfly: each exchange keeps half of
              // the columns, so 8+4+2+1+1 = 16 shuffles instead of 16 x 5; lane L ends up with column L >> 1
#pragma unroll
              for (int h = 8, o = 16; h >= 1; h >>= 1, o >>= 1) {
                const bool up = (lane & o) != 0;
#pragma unroll
                for (int q = 0; q < h; ++q) {
                  const float send = up ? v[q] : v[q + h], mine = up ? v[q + h] : v[q];
                  v[q] = fmaxf(mine, __shfl_xor_sync(0xffffffffu, send, o));
                }
              }
              const float keep = fmaxf(v[0], __shfl_xor_sync(0xffffffffu, v[0], 1));
              const int col = c * 16 + (lane >> 1);
              if (!(lane & 1) && col < m.out_channels && gid < a.rows_out) out[gid * m.out_channels + col] = keep;
            } else if (MODE == MODE_FP) {
              const long long pid = tile * ROWS + row;
              if (pid < a.rows_out) {
                float *op = out + pid * m.out_channels + c * 16;
                if ((m.out_channels & 7) == 0) {             // rows are 32-byte aligned: 256-bit stores (whole L2 sectors)
#pragma unroll
                  for (int q = 0; q < 2; ++q)
                    if (c * 16 + 8 * q < m.out_channels)
                      st_global_256(op + 8 * q, make_uint4(__float_as_uint(v[8 * q]), __float_as_uint(v[8 * q + 1]), __float_as_uint(v[8 * q + 2]), __float_as_uint(v[8 * q + 3])),
                                    make_uint4(__float_as_uint(v[8 * q + 4]), __float_as_uint(v[8 * q + 5]), __float_as_uint(v[8 * q + 6]), __float_as_uint(v[8 * q + 7])));
                } else if ((m.out_channels & 3) == 0) {      // rows are 16-byte aligned: 128-bit stores
#pragma unroll
                  for (int q = 0; q < 4; ++q)
                    if (c * 16 + 4 * q < m.out_channels)
                      *reinterpret_cast<float4 *>(op + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                } else {
#pragma unroll
                  for (int q = 0; q < 16; ++q)
                    if (c * 16 + q < m.out_channels) op[q] = v[q];
                }
              }
            } else {
#pragma unroll
              for (int q = 0; q < 16; ++q) fbuf[(size_t)(quarter * N + c * 16 + q) * FA_PSTRIDE + lane] = v[q];
            }
            c += 2;
          }
        }
        if (last && MODE == MODE_FA) {
          group_sync(g);
          // reduce over the k pixel slots; consecutive threads take consecutive channels (coalesced rows of `out`)
          const int oc = m.out_channels;
          int p = ltid / oc, c = ltid - p * oc;
          const int dp = GROUP_THREADS / oc, dc = GROUP_THREADS - dp * oc;
          for (; p < 32; p += dp, c += dc) {
            if (c >= oc) { c -= oc; ++p; if (p >= 32) break; }
            const long long pid = tile * 32 + p;
            if (pid >= a.rows_out) continue;
            float v = fbuf[(size_t)c * FA_PSTRIDE + p];
            for (int q = 1; q < a.k; ++q) {
              const float u = fbuf[(size_t)(q * N + c) * FA_PSTRIDE + p];
              v = a.reduce == REDUCE_SUM ? __fadd_rn(v, u) : fmaxf(v, u);
            }
            out[pid * oc + c] = v;
          }
          group_sync(g);                                      // the staging area is the next tile's A operand
        }
      }
    }
  } else if (warp == NW) {
    // =========================== MMA issuer ==========================================================================
    // The whole warp walks the request order (all values warp-uniform -> descriptors in uniform registers) and one
    // elected lane issues: under `if (lane == 0)` ptxas wraps every tcgen05 instruction in a thread-by-thread
    // broadcast loop that costs more than a narrow layer's MMA takes to execute.
    uint32_t ph[MAX_GROUPS] = {0u, 0u, 0u};
    uint32_t q_cons = 0;
    const uint32_t smem_s = smem_u32(smem), wreg_s = smem_u32(wreg);
    for (long long r = 0; r < rounds; ++r) {
      for (int sg = 0; sg < nsegs; ++sg) {
        int l, kbeg, klen;
        seg_info(sg, l, kbeg, klen);
        const int K = m.k[l], N = m.n[l];
        if (m.resident) {
#pragma unroll
          for (int g = 0; g < NG; ++g) {
            if (r * NG + g >= n_my) continue;
            mbar_wait(bar_aready0 + 8 * g, ph[g]);
            ph[g] ^= 1u;
            tc_fence_after();
            const uint32_t a_hi_s = smem_s + (uint32_t)((size_t)g * 2 * a_bytes), a_lo_s = a_hi_s + (uint32_t)a_bytes;
            const uint32_t t_group = tmem_base + (uint32_t)(g * m.tmem_cols);
            const uint32_t wbase = wreg_s + (uint32_t)m.res_off[l];
            if (elect_one()) {
              for (int n0 = 0; n0 < N; n0 += 256) {
                const uint32_t nb = (uint32_t)min(256, N - n0);
                const uint32_t idesc = make_idesc(ROWS, (int)nb);
                const uint32_t wh = wbase + (uint32_t)n0 * K * 2, wl = wh + (uint32_t)K * N * 2;
                for (int k0 = kbeg; k0 < kbeg + klen; k0 += 16) {
                  const uint32_t ks = (uint32_t)(k0 >> 3), ka = (uint32_t)((k0 - kbeg) >> 3);
                  issue_kstep(t_group + (uint32_t)n0, a_hi_s + ka * SLAB, a_lo_s + ka * SLAB, wh + ks * nb * 16, wl + ks * nb * 16, nb,
                              idesc, k0 == 0);
                }
              }
              umma_commit(bar_acc0 + 8 * g);
            }
            __syncwarp();
          }
        } else {
          // Streamed weights: every ring stage is used by ALL tile groups of the round before it is released (chunk-major,
          // group-minor).  Round 1 streamed a layer's weights once per group: the wide chains (FP4: 278 KB per 128 rows)
          // were bound by that L2 -> shared-memory traffic; sharing a stage divides it by the group count.
#pragma unroll
          for (int g = 0; g < NG; ++g) {
            if (r * NG + g >= n_my) continue;
            mbar_wait(bar_aready0 + 8 * g, ph[g]);
            ph[g] ^= 1u;
          }
          tc_fence_after();
          const int kchunks = (klen + m.kc - 1) / m.kc, nblocks = (N + 255) / 256, total = kchunks * nblocks;
          for (int c = 0; c < total; ++c) {
            const uint32_t s = q_cons % S;
            const int n0 = (c / kchunks) * 256, k0 = kbeg + (c % kchunks) * m.kc;
            const uint32_t nb = (uint32_t)min(256, N - n0);
            const int kc = min(m.kc, kbeg + klen - k0);
            const uint32_t idesc = make_idesc(ROWS, (int)nb);
            mbar_wait(bar_full0 + 8 * s, (q_cons / S) & 1u);                       // the chunk has landed
            const uint32_t wh = wreg_s + s * (uint32_t)(2 * stage_half), wl = wh + (uint32_t)stage_half;
            if (elect_one()) {
#pragma unroll
              for (int g = 0; g < NG; ++g) {
                if (r * NG + g >= n_my) continue;
                const uint32_t a_hi_s = smem_s + (uint32_t)((size_t)g * 2 * a_bytes), a_lo_s = a_hi_s + (uint32_t)a_bytes;
                const uint32_t t_group = tmem_base + (uint32_t)(g * m.tmem_cols);
                for (int j = 0; j < kc; j += 16) {
                  const uint32_t ka = (uint32_t)((k0 + j - kbeg) >> 3), js = (uint32_t)(j >> 3);
                  issue_kstep(t_group + (uint32_t)n0, a_hi_s + ka * SLAB, a_lo_s + ka * SLAB, wh + js * nb * 16, wl + js * nb * 16, nb,
                              idesc, k0 + j == 0);
                }
              }
              umma_commit(bar_empty0 + 8 * s);                                     // frees the stage when these MMAs retire
              if (c == total - 1) {
#pragma unroll
                for (int g = 0; g < NG; ++g)
                  if (r * NG + g < n_my) umma_commit(bar_acc0 + 8 * g);
              }
            }
            __syncwarp();
            ++q_cons;
          }
        }
      }
    }
  } else if (!m.resident) {
    // =========================== weight producer: same request order, runs ahead ======================================
    uint32_t q_prod = 0;
    const uint32_t wreg_s = smem_u32(wreg);
    for (long long r = 0; r < rounds; ++r) {
      for (int sg = 0; sg < nsegs; ++sg) {
        int l, kbeg, klen;
        seg_info(sg, l, kbeg, klen);
        const int K = m.k[l], N = m.n[l];
        const int kchunks = (klen + m.kc - 1) / m.kc, nblocks = (N + 255) / 256, total = kchunks * nblocks;
        if (r * NG >= n_my) continue;
        for (int c = 0; c < total; ++c) {      // once per round: the stage is shared by the round's tile groups
          const int n0 = (c / kchunks) * 256, k0 = kbeg + (c % kchunks) * m.kc;
          const uint32_t nb = (uint32_t)min(256, N - n0), kc = (uint32_t)min(m.kc, kbeg + klen - k0);
          const uint32_t s = q_prod % S, bytes = (kc >> 3) * nb * 16;
          if (q_prod >= S) mbar_wait(bar_empty0 + 8 * s, ((q_prod / S) - 1) & 1u);  // MMAs of the previous use have retired
          const unsigned char *gh = reinterpret_cast<const unsigned char *>(m.w_hi[l]) + (size_t)n0 * K * 2 + (size_t)(k0 >> 3) * nb * 16;
          const unsigned char *gl = reinterpret_cast<const unsigned char *>(m.w_lo[l]) + (size_t)n0 * K * 2 + (size_t)(k0 >> 3) * nb * 16;
          const uint32_t dst = wreg_s + s * (uint32_t)(2 * stage_half);
          if (elect_one()) {
            mbar_expect_tx(bar_full0 + 8 * s, 2 * bytes);
            bulk_g2s(dst, gh, bytes, bar_full0 + 8 * s);
            bulk_g2s(dst + (uint32_t)stage_half, gl, bytes, bar_full0 + 8 * s);
          }
          __syncwarp();
          ++q_prod;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == NW) tmem_dealloc(tmem_base, (uint32_t)m.tmem_alloc);
}

static size_t smem_bytes(const Chain &m) {
  const size_t a_bytes = (size_t)m.groups * (m.kmax >> 3) * SLAB * 2;
  const size_t w = m.resident ? (size_t)m.res_bytes : (size_t)m.stages * 2 * (m.kc >> 3) * m.nbmax * 16;
  return a_bytes + ((w + 127) & ~(size_t)127) + m.groups * ((sizeof(Aux) + 127) & ~(size_t)127) + 256 + (((size_t)m.bias_floats * 4 + 127) & ~(size_t)127);
}

constexpr size_t SMEM_CAP = 227 * 1024;

// fills the derived fields; returns false when the chain does not fit this kernel
static bool finalize(Chain &m, int mode, int fa_k) {
  m.nbmax = 0;
  int nmax = 0, kmax_rest = 0;
  size_t wbytes = 0;
  m.bias_floats = 0;
  for (int l = 0; l < m.num_layers; ++l) {
    m.bias_off[l] = m.bias_floats;
    m.bias_floats += m.n[l];
    if (l > 0 && m.k[l] > kmax_rest) kmax_rest = m.k[l];
    const int nb = m.n[l] < 256 ? m.n[l] : 256;
    if (nb > m.nbmax) m.nbmax = nb;
    if (m.n[l] > nmax) nmax = m.n[l];
    m.res_off[l] = (int)wbytes;
    wbytes += (size_t)m.k[l] * m.n[l] * 4;   // hi + lo
  }
  if (nmax > 512) return false;
  m.tmem_cols = 32;
  while (m.tmem_cols < nmax) m.tmem_cols <<= 1;
  m.res_bytes = (int)wbytes;
  // First-layer K panel: the whole built row when it fits, else 256 / 128 channels per pass.
  const int panels[3] = {m.k[0], 256, 128};
  for (int pi = 0; pi < 3; ++pi) {
    if (pi > 0 && panels[pi] >= m.k[0]) continue;
    m.panel = panels[pi];
    m.kmax = m.panel > kmax_rest ? m.panel : kmax_rest;   // capacity of the activation tile
    // FA: the fp32 staging [k][N][33] of the last layer must fit in the group's (dead) A operand (kmax * 512 B)
    if (mode == MODE_FA && (size_t)fa_k * m.n[m.num_layers - 1] * FA_PSTRIDE * 4 > (size_t)m.kmax * 512) continue;
    // Most tile groups first (they overlap MMAs with builds / epilogues inside the CTA and share the weights);
    // per group count: resident weights if they fit, else the deepest ring of 32- or 16-channel chunks.
    for (m.groups = MAX_GROUPS; m.groups >= 1; --m.groups) {
      if (m.groups * m.tmem_cols > 512) continue;
      m.tmem_alloc = 32;
      while (m.tmem_alloc < m.groups * m.tmem_cols) m.tmem_alloc <<= 1;
      m.kc = 32; m.stages = 2; m.resident = 1;
      if (pi == 0 && smem_bytes(m) <= SMEM_CAP) return true;
      m.resident = 0;
      int best_kc = 0, best_stages = 0;
      for (int kc = 32; kc >= 16; kc >>= 1) {
        m.kc = kc;
        for (m.stages = MAX_STAGES; m.stages >= 2; --m.stages)
          if (smem_bytes(m) <= SMEM_CAP) break;
        if (m.stages >= 2 && m.stages * kc > best_stages * best_kc) { best_kc = kc; best_stages = m.stages; }
      }
      // a ring shallower than 64 channels in flight starves the issuer: try fewer groups first
      if (best_kc && (best_stages * best_kc >= 64 || m.groups == 1)) { m.kc = best_kc; m.stages = best_stages; return true; }
    }
  }
  return false;
}

template <int MODE, int NG>
static int launch_ng(const BuildArgs &a, const Chain &m, float *out, long long tiles, cudaStream_t stream) {
  auto kern = tc_fused_mlp_kernel<MODE, NG>;
  const size_t smem = smem_bytes(m);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("tc_fused_mlp: smem attribute (%zu B): %s", smem, cudaGetErrorString(e)); return (int)e; }
  // persistent: as many CTAs as are resident: shared memory, registers, threads, TMEM columns
  constexpr int THREADS = NG * GROUP_THREADS + CTRL_THREADS;
  cudaFuncAttributes fa;
  int per_sm = (int)(SMEM_CAP / (smem + 1024));
  if (cudaFuncGetAttributes(&fa, kern) == cudaSuccess && fa.numRegs > 0) {
    const int by_regs = 65536 / (((fa.numRegs + 7) / 8 * 8) * THREADS);
    if (by_regs < per_sm) per_sm = by_regs;
  }
  if (per_sm > 2048 / THREADS) per_sm = 2048 / THREADS;
  if (per_sm * m.tmem_alloc > 512) per_sm = 512 / m.tmem_alloc;
  if (per_sm < 1) per_sm = 1;
  static const bool debug = getenv("MVPNET_B200_DEBUG") != nullptr;
  if (debug)
    fprintf(stderr, "[tc_fused_mlp mode=%d] tiles=%lld groups=%d smem=%zu regs=%d per_sm=%d resident=%d kc=%d stages=%d panel=%d kmax=%d tmem=%d/%d\n",
            MODE, tiles, NG, smem, fa.numRegs, per_sm, m.resident, m.kc, m.stages, m.panel, m.kmax, m.tmem_cols, m.tmem_alloc);
  long long grid = (long long)sm_count() * per_sm;
  const long long want = (tiles + NG - 1) / NG;
  if (grid > want) grid = want;
  kern<<<(unsigned)grid, THREADS, smem, stream>>>(a, m, out, tiles);
  return launch_status("tc_fused_mlp");
}

template <int MODE>
static int launch(const BuildArgs &a, Chain m, float *out, long long tiles, cudaStream_t stream) {
  // fewer groups when there are not enough tiles to give every SM NG of them: spread over the SMs first
  int ng = m.groups;
  const long long per_sm_tiles = tiles / sm_count();
  if (per_sm_tiles < ng) ng = per_sm_tiles < 1 ? 1 : (int)per_sm_tiles;
  static const char *force = getenv("MVPNET_B200_TC_GROUPS");
  if (force && atoi(force) >= 1 && atoi(force) < ng) ng = atoi(force);
  if (ng != m.groups) {
    m.groups = ng;
    m.tmem_alloc = 32;
    while (m.tmem_alloc < m.groups * m.tmem_cols) m.tmem_alloc <<= 1;
    if (!m.resident) {                         // the freed activation tiles deepen the ring
      for (int s = MAX_STAGES; s > m.stages; --s) {
        Chain t = m;
        t.stages = s;
        if (smem_bytes(t) <= SMEM_CAP) { m.stages = s; break; }
      }
    }
  }
  // Streamed-weight chains: TWO single-group CTAs per SM where a tile + a two-stage ring fit twice (FP4: 107 KB, 128 TMEM
  // columns, 96 registers) instead of one CTA with two lock-stepped groups.  Independent CTAs de-phase, so one's MMA phase
  // overlaps the other's gather / epilogue: FP4 0.288 -> 0.239 ms (each CTA streams its own weights; the ring depth was
  // measured irrelevant: 2 or 5 stages, same time).  MVPNET_B200_TC_TWO_CTAS=0 keeps the grouped form.
  static const bool two_ctas = getenv("MVPNET_B200_TC_TWO_CTAS") == nullptr || getenv("MVPNET_B200_TC_TWO_CTAS")[0] != '0';
  if (two_ctas && !m.resident && ng >= 2 && m.tmem_cols <= 256 && tiles >= 4LL * sm_count()) {
    Chain t = m;
    t.groups = 1;
    t.stages = 2;
    t.tmem_alloc = 32;
    while (t.tmem_alloc < t.tmem_cols) t.tmem_alloc <<= 1;
    if (2 * (smem_bytes(t) + 1024) <= SMEM_CAP) return launch_ng<MODE, 1>(a, t, out, tiles, stream);
  }
  static const char *cap = getenv("MVPNET_B200_TC_STAGES_CAP");     // experiment knob
  if (cap && !m.resident && atoi(cap) >= 2 && atoi(cap) < m.stages) m.stages = atoi(cap);
  if (ng == 3) return launch_ng<MODE, 3>(a, m, out, tiles, stream);
  if (ng == 2) return launch_ng<MODE, 2>(a, m, out, tiles, stream);
  return launch_ng<MODE, 1>(a, m, out, tiles, stream);
}

}  // namespace tc
}  // namespace mvp

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
static int mvp_tc_to_chain(const mvp_tc_chain_t *c, int k0_min, int mode, int fa_k, mvp::tc::Chain *m) {
  using namespace mvp;
  MVP_REQUIRE(c, MVP_ERR_NULL, "tc_fused_mlp: null chain");
  MVP_REQUIRE(c->num_layers >= 1 && c->num_layers <= tc::MAX_LAYERS, MVP_ERR_INVALID_ARG, "tc_fused_mlp: 1..6 layers");
  MVP_REQUIRE(c->k[0] >= k0_min, MVP_ERR_INVALID_ARG, "tc_fused_mlp: k[0]=%d < %d input channels", c->k[0], k0_min);
  m->num_layers = c->num_layers;
  m->out_channels = c->out_channels;
  for (int l = 0; l < c->num_layers; ++l) {
    MVP_REQUIRE(c->k[l] % 16 == 0 && c->n[l] % 16 == 0 && c->k[l] > 0 && c->n[l] > 0, MVP_ERR_INVALID_ARG,
                "tc_fused_mlp: k and n must be positive multiples of 16");
    MVP_REQUIRE(l == 0 || c->k[l] == c->n[l - 1], MVP_ERR_INVALID_ARG, "tc_fused_mlp: k[l] must equal n[l-1]");
    MVP_REQUIRE(c->w_hi[l] && c->w_lo[l] && c->bias[l], MVP_ERR_NULL, "tc_fused_mlp: null weights");
    MVP_REQUIRE((((uintptr_t)c->w_hi[l] | (uintptr_t)c->w_lo[l] | (uintptr_t)c->bias[l]) & 15) == 0, MVP_ERR_INVALID_ARG,
                "tc_fused_mlp: weights must be 16-byte aligned");
    m->k[l] = c->k[l]; m->n[l] = c->n[l]; m->relu[l] = c->relu[l];
    m->w_hi[l] = (const __nv_bfloat16 *)c->w_hi[l]; m->w_lo[l] = (const __nv_bfloat16 *)c->w_lo[l]; m->bias[l] = c->bias[l];
  }
  MVP_REQUIRE(c->out_channels > 0 && c->out_channels <= c->n[c->num_layers - 1], MVP_ERR_INVALID_ARG, "tc_fused_mlp: bad out_channels");
  MVP_REQUIRE(tc::finalize(*m, mode, fa_k), MVP_ERR_UNSUPPORTED, "tc_fused_mlp: chain too wide for shared memory / TMEM (kmax=%d)", m->kmax);
  return 0;
}

extern "C" int mvp_tc_chain_supported(const mvp_tc_chain_t *c, int mode) {
  mvp::tc::Chain m;
  if (!c || c->num_layers < 1 || c->num_layers > mvp::tc::MAX_LAYERS) return 0;
  m.num_layers = c->num_layers;
  for (int l = 0; l < c->num_layers; ++l) { m.k[l] = c->k[l]; m.n[l] = c->n[l]; }
  return mvp::tc::finalize(m, mode, 4) ? 1 : 0;
}

extern "C" int mvp_tc_fused_set_abstraction(const float *feat, int64_t C, const float *xyz, const float *new_xyz,
                                            const int64_t *nbr, int64_t B, int64_t N, int64_t M, int64_t K,
                                            const mvp_tc_chain_t *chain, float *out, mvp_stream_t stream) {
  using namespace mvp;
  MVP_REQUIRE(K == 32, MVP_ERR_UNSUPPORTED, "tc_fused_set_abstraction: max_neighbors must be 32");
  MVP_REQUIRE(C % 8 == 0 && C >= 0, MVP_ERR_UNSUPPORTED, "tc_fused_set_abstraction: feature channels must be a multiple of 8");
  MVP_REQUIRE(B >= 0 && N > 0 && M >= 0, MVP_ERR_INVALID_ARG, "tc_fused_set_abstraction: bad sizes");
  tc::Chain m;
  if (int rc = mvp_tc_to_chain(chain, (int)C + 3, MODE_SA, 0, &m)) return rc;
  if (B * M == 0) return 0;
  MVP_REQUIRE(xyz && new_xyz && nbr && out && (feat || C == 0), MVP_ERR_NULL, "tc_fused_set_abstraction: null pointer");
  MVP_REQUIRE(((uintptr_t)feat & 15) == 0, MVP_ERR_INVALID_ARG, "tc_fused_set_abstraction: feat must be 16-byte aligned");
  BuildArgs a = {};
  a.rows_out = B * M; a.feat_channels = (int)C; a.feat = feat; a.xyz = xyz; a.new_xyz = new_xyz; a.nbr = nbr;
  a.n_src = N; a.n_out = M; a.k = 32;
  return tc::launch<MODE_SA>(a, m, out, (a.rows_out + 3) / 4, (cudaStream_t)stream);
}

extern "C" int mvp_tc_fused_feature_aggregation(const float *feat2d, int64_t s_n, int64_t s_c, int64_t s_h, int64_t s_w,
                                                int64_t C, int64_t nv, int64_t h, int64_t w, const float *pix_xyz,
                                                const float *points, const int64_t *knn, int64_t B, int64_t Np, int64_t K,
                                                int reduce_sum, const mvp_tc_chain_t *chain, float *out, mvp_stream_t stream) {
  using namespace mvp;
  MVP_REQUIRE(K >= 1 && K <= 4, MVP_ERR_UNSUPPORTED, "tc_fused_feature_aggregation: k must be in [1, 4]");
  MVP_REQUIRE(C % 8 == 0 && C > 0, MVP_ERR_UNSUPPORTED, "tc_fused_feature_aggregation: feature channels must be a multiple of 8");
  MVP_REQUIRE(B >= 0 && Np >= 0 && nv > 0 && h > 0 && w > 0, MVP_ERR_INVALID_ARG, "tc_fused_feature_aggregation: bad sizes");
  MVP_REQUIRE(nv * h * w < (1LL << 31), MVP_ERR_UNSUPPORTED, "tc_fused_feature_aggregation: more than 2^31 pixels per cloud");
  tc::Chain m;
  if (int rc = mvp_tc_to_chain(chain, (int)C + 4, MODE_FA, (int)K, &m)) return rc;
  if (B * Np == 0) return 0;
  MVP_REQUIRE(feat2d && pix_xyz && points && knn && out, MVP_ERR_NULL, "tc_fused_feature_aggregation: null pointer");
  if (s_c == 1)
    MVP_REQUIRE((((uintptr_t)feat2d) & 15) == 0 && s_n % 4 == 0 && s_h % 4 == 0 && s_w % 4 == 0, MVP_ERR_UNSUPPORTED,
                "tc_fused_feature_aggregation: channels-last feature map must be 16-byte aligned per pixel");
  BuildArgs a = {};
  a.rows_out = B * Np; a.feat_channels = (int)C; a.feat = feat2d; a.xyz = pix_xyz; a.new_xyz = points; a.nbr = knn;
  a.n_src = nv * h * w; a.n_out = Np; a.k = (int)K; a.reduce = reduce_sum ? REDUCE_SUM : REDUCE_MAX;
  a.s_n = s_n; a.s_c = s_c; a.s_h = s_h; a.s_w = s_w; a.hw = (int)(h * w); a.w = (int)w; a.nv = (int)nv;
  return tc::launch<MODE_FA>(a, m, out, (a.rows_out + 31) / 32, (cudaStream_t)stream);
}

extern "C" int mvp_tc_fused_feature_propagation(const float *sparse_feat, int64_t Cs, const int64_t *idx, const float *dist2,
                                                const float *skip, int64_t Cd, int64_t B, int64_t Ns, int64_t Nd, float eps,
                                                const mvp_tc_chain_t *chain, float *out, mvp_stream_t stream) {
  using namespace mvp;
  MVP_REQUIRE(Cs % 8 == 0 && Cd % 8 == 0 && Cs > 0 && Cd >= 0, MVP_ERR_UNSUPPORTED,
              "tc_fused_feature_propagation: channel counts must be multiples of 8");
  MVP_REQUIRE(B >= 0 && Ns > 0 && Nd >= 0, MVP_ERR_INVALID_ARG, "tc_fused_feature_propagation: bad sizes");
  tc::Chain m;
  if (int rc = mvp_tc_to_chain(chain, (int)(Cs + Cd), MODE_FP, 0, &m)) return rc;
  if (B * Nd == 0) return 0;
  MVP_REQUIRE(sparse_feat && idx && dist2 && out && (skip || Cd == 0), MVP_ERR_NULL, "tc_fused_feature_propagation: null pointer");
  MVP_REQUIRE((((uintptr_t)sparse_feat | (uintptr_t)skip) & 15) == 0, MVP_ERR_INVALID_ARG,
              "tc_fused_feature_propagation: features must be 16-byte aligned");
  MVP_REQUIRE(((uintptr_t)out & 31) == 0 || (chain->out_channels & 7) != 0, MVP_ERR_INVALID_ARG,
              "tc_fused_feature_propagation: the output must be 32-byte aligned (256-bit stores)");
  BuildArgs a = {};
  a.rows_out = B * Nd; a.feat_channels = (int)Cs; a.feat = sparse_feat; a.nbr = idx; a.dist = dist2; a.skip = skip;
  a.skip_channels = (int)Cd; a.n_src = Ns; a.n_out = Nd; a.k = 3; a.eps = eps;
  return tc::launch<MODE_FP>(a, m, out, (a.rows_out + tc::ROWS - 1) / tc::ROWS, (cudaStream_t)stream);
}
