// Hardware probe for the round-2 fused-MLP redesign (run on the B200 box, not part of the product):
//   A. cp.async.bulk.tensor.2d.tile::gather4 into a SWIZZLE_128B K-major tile: which box shape the tensor map needs and
//      where the bytes land;
//   B. tcgen05.mma reading that gathered tile through a SWIZZLE_128B descriptor (A) x a no-swizzle weight tile (B);
//   C. tcgen05.mma with the A operand in TENSOR MEMORY (written by tcgen05.st as packed bf16 pairs);
//   D. issue throughput of gather4 (32 lanes x 2, or one lane x 64) against 16-byte cp.async (LDGSTS) for a
//      128-row x 64-channel x {hi, lo} neighbourhood tile, 148 persistent CTAs, 4-stage ring.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/probe_tma_gather.bin tools/probe_tma_gather.cu
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess) {                                                                    \
      printf("CUDA error '%s' at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);         \
      return 1;                                                                                 \
    }                                                                                           \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {   // false = timed out
  uint32_t ok, spins = 0;
  do {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!ok && ++spins > (1u << 22)) return false;
  } while (!ok);
  return true;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void gather4(uint32_t dst, const CUtensorMap *map, int col, int r0, int r1, int r2, int r3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(dst),
      "l"(map), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar)
      : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ uint64_t desc_noswz(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) | ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {   // K-major SWIZZLE_128B: SBO = 1024 (8 rows x 128 B), LBO unused (1)
  return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b),
               "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a_tmem),
               "l"(b), "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\ntcgen05.wait::st.sync.aligned;" ::"r"(taddr),
               "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

// ---------------------------------------------------------------------------------------------- A: where gather4 lands
__global__ void probe_a(const __grid_constant__ CUtensorMap map, const int *idx, uint4 *out, int *status) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  unsigned char *tile = (unsigned char *)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
  const uint32_t b = smem_u32(&bar);
  for (int i = threadIdx.x; i < 16384 / 16; i += blockDim.x) ((uint4 *)tile)[i] = make_uint4(0xdeadbeefu, 0xdeadbeefu, 0xdeadbeefu, 0xdeadbeefu);
  if (threadIdx.x == 0) { mbar_init(b, 1); fence_barrier_init(); }
  fence_proxy_async();
  __syncthreads();
  if (threadIdx.x < 32) {
    const int l = threadIdx.x;
    if (l == 0) mbar_expect_tx(b, 16384);
    __syncwarp();
    gather4(smem_u32(tile) + l * 512, &map, 0, idx[4 * l], idx[4 * l + 1], idx[4 * l + 2], idx[4 * l + 3], b);
  }
  const bool ok = mbar_wait(b, 0);
  if (threadIdx.x == 0) *status = ok ? 1 : -1;
  __syncthreads();
  for (int i = threadIdx.x; i < 16384 / 16; i += blockDim.x) out[i] = ((uint4 *)tile)[i];
}

// ---------------------------------------------------------------------------------------------- B: SW128 A (gathered) x no-swizzle B
__global__ void __launch_bounds__(128, 1) probe_b(const __grid_constant__ CUtensorMap map, const int *idx, const __nv_bfloat16 *wB /*[32][64]*/,
                                                   float *out /*[128][32]*/, int *status) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bars[2];
  __shared__ uint32_t tslot;
  unsigned char *tile = (unsigned char *)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
  unsigned char *wt = tile + 16384;                         // [K/8 = 8][N = 32][8] bf16 = 4 KB
  const uint32_t b0 = smem_u32(&bars[0]), b1 = smem_u32(&bars[1]);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { mbar_init(b0, 1); mbar_init(b1, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&tslot), 32);
  for (int i = tid; i < 32 * 64; i += 128) {
    const int n = i / 64, k = i % 64;
    ((__nv_bfloat16 *)wt)[(k >> 3) * 32 * 8 + n * 8 + (k & 7)] = wB[i];
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tslot;
  if (warp == 0) {
    if (lane == 0) mbar_expect_tx(b0, 16384);
    __syncwarp();
    gather4(smem_u32(tile) + lane * 512, &map, 0, idx[4 * lane], idx[4 * lane + 1], idx[4 * lane + 2], idx[4 * lane + 3], b0);
    const bool ok = mbar_wait(b0, 0);
    if (!ok && lane == 0) *status = -1;
    tc_fence_after();
    if (lane == 0) {
      const uint32_t idesc = make_idesc(128, 32);
      for (int k = 0; k < 4; ++k)
        umma_ss(tbase, desc_sw128(smem_u32(tile) + k * 32), desc_noswz(smem_u32(wt) + k * 2 * 32 * 16, 32 * 16, 128), idesc, k > 0);
      umma_commit(b1);
    }
    __syncwarp();
  }
  const bool ok = mbar_wait(b1, 0);
  tc_fence_after();
  uint32_t r[16];
  for (int c = 0; c < 2; ++c) {
    tmem_ld16(tbase + ((uint32_t)(warp * 32) << 16) + c * 16, r);
    for (int q = 0; q < 16; ++q) out[(warp * 32 + lane) * 32 + c * 16 + q] = __uint_as_float(r[q]);
  }
  if (tid == 0 && *status == 0) *status = ok ? 1 : -2;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 32);
}

// ---------------------------------------------------------------------------------------------- C: A operand in tensor memory
__global__ void __launch_bounds__(128, 1) probe_c(const __nv_bfloat16 *A /*[128][16]*/, const __nv_bfloat16 *wB /*[32][16]*/, float *out, int *status) {
  __shared__ __align__(128) unsigned char wt[2 * 32 * 16];  // [K/8 = 2][32][8]
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t b1 = smem_u32(&bar);
  if (tid == 0) { mbar_init(b1, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&tslot), 64);
  for (int i = tid; i < 32 * 16; i += 128) {
    const int n = i / 16, k = i % 16;
    ((__nv_bfloat16 *)wt)[(k >> 3) * 32 * 8 + n * 8 + (k & 7)] = wB[i];
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tslot;
  // row = tid: pack (k, k+1) into one 32-bit column; 8 columns at tbase + 32
  uint32_t v[8];
  for (int c = 0; c < 8; ++c) {
    const __nv_bfloat162 p = __halves2bfloat162(A[tid * 16 + 2 * c], A[tid * 16 + 2 * c + 1]);
    v[c] = *reinterpret_cast<const uint32_t *>(&p);
  }
  tmem_st8(tbase + ((uint32_t)(warp * 32) << 16) + 32, v);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) {
    if (lane == 0) {
      umma_ts(tbase, tbase + 32, desc_noswz(smem_u32(wt), 32 * 16, 128), make_idesc(128, 32), 0);
      umma_commit(b1);
    }
    __syncwarp();
  }
  const bool ok = mbar_wait(b1, 0);
  tc_fence_after();
  uint32_t r[16];
  for (int c = 0; c < 2; ++c) {
    tmem_ld16(tbase + ((uint32_t)(warp * 32) << 16) + c * 16, r);
    for (int q = 0; q < 16; ++q) out[(warp * 32 + lane) * 32 + c * 16 + q] = __uint_as_float(r[q]);
  }
  if (tid == 0) *status = ok ? 1 : -2;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 64);
}

// ---------------------------------------------------------------------------------------------- D: gather throughput
constexpr int STAGES = 4;
constexpr int TILE_BYTES = 32768;   // 128 rows x 128 B x {hi, lo}

// mode 0: 32 lanes x 2 gather4;  mode 1: lane 0 x 64 gather4;  mode 2: 4 producer warps, 16-byte cp.async
__global__ void __launch_bounds__(288, 1) probe_d(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
                                                   const __nv_bfloat16 *hi, const __nv_bfloat16 *lo, const int *idx, int tiles_per_cta, int mode,
                                                   unsigned long long *sink, int *status) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t full[STAGES], empty[STAGES];
  unsigned char *ring = (unsigned char *)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nprod = mode == 2 || mode == 3 ? 128 : mode == 4 ? 256 : 1;     // arrivals on `full` besides the tx bytes
  const int cons_warp = (int)(blockDim.x >> 5) - 1;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(smem_u32(&full[s]), nprod); mbar_init(smem_u32(&empty[s]), 1); }
    fence_barrier_init();
  }
  __syncthreads();
  const int *my = idx + (size_t)blockIdx.x * tiles_per_cta * 128;
  if (warp == cons_warp) {                   // consumer
    unsigned long long acc = 0;
    for (int t = 0; t < tiles_per_cta; ++t) {
      const int s = t % STAGES;
      if (!mbar_wait(smem_u32(&full[s]), (t / STAGES) & 1)) { if (lane == 0) *status = -3; return; }
      acc += *(const unsigned long long *)(ring + s * TILE_BYTES + lane * 1024);
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&empty[s]));
    }
    if (acc == 0x1234567ull) *sink = acc;
  } else if (mode < 2 && warp == 0) {
    for (int t = 0; t < tiles_per_cta; ++t) {
      const int s = t % STAGES;
      const int4 r = *reinterpret_cast<const int4 *>(my + t * 128 + 4 * lane);
      if (t >= STAGES && !mbar_wait(smem_u32(&empty[s]), ((t / STAGES) - 1) & 1)) { if (lane == 0) *status = -4; return; }
      const uint32_t dst = smem_u32(ring) + s * TILE_BYTES, fb = smem_u32(&full[s]);
      if (lane == 0) mbar_expect_tx(fb, TILE_BYTES);
      __syncwarp();
      if (mode == 0) {
        gather4(dst + lane * 512, &map_hi, 0, r.x, r.y, r.z, r.w, fb);
        gather4(dst + 16384 + lane * 512, &map_lo, 0, r.x, r.y, r.z, r.w, fb);
      } else {
        for (int l = 0; l < 32; ++l) {
          const int a = __shfl_sync(0xffffffffu, r.x, l), b = __shfl_sync(0xffffffffu, r.y, l), c = __shfl_sync(0xffffffffu, r.z, l),
                    d = __shfl_sync(0xffffffffu, r.w, l);
          if (lane == 0) {
            gather4(dst + l * 512, &map_hi, 0, a, b, c, d, fb);
            gather4(dst + 16384 + l * 512, &map_lo, 0, a, b, c, d, fb);
          }
        }
      }
    }
  } else if (mode >= 3 && warp < nprod / 32) {
    // SWIZZLE_128B K-major tile: row r = 128 contiguous bytes, 16-byte chunk c at position c ^ (r & 7).  8 consecutive
    // lanes take the 8 chunks of one row: every warp instruction reads 4 whole 128-byte lines and writes 512
    // contiguous (permuted) bytes of shared memory.
    const int chunk = lane & 7, rsub = lane >> 3, rows_per_pass = nprod / 8;
    for (int t = 0; t < tiles_per_cta; ++t) {
      const int s = t % STAGES;
      if (t >= STAGES && !mbar_wait(smem_u32(&empty[s]), ((t / STAGES) - 1) & 1)) { if (tid == 0) *status = -4; return; }
      const uint32_t base = smem_u32(ring) + s * TILE_BYTES;
#pragma unroll 4
      for (int r0 = 0; r0 < 128; r0 += rows_per_pass) {
        const int r = r0 + warp * 4 + rsub;
        const int row = my[t * 128 + r];
        const uint32_t dst = base + r * 128 + ((chunk ^ (r & 7)) << 4);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"((const unsigned char *)(hi + (size_t)row * 64) + chunk * 16) : "memory");
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16384), "l"((const unsigned char *)(lo + (size_t)row * 64) + chunk * 16) : "memory");
      }
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&full[s])) : "memory");
    }
  } else if (mode == 2 && warp < 4) {
    // thread = one row (128 rows): 8 + 8 cp.async of 16 B into the no-swizzle K-slab layout [k/8][row][16 B]
    for (int t = 0; t < tiles_per_cta; ++t) {
      const int s = t % STAGES;
      const int row = my[t * 128 + tid];
      if (t >= STAGES && !mbar_wait(smem_u32(&empty[s]), ((t / STAGES) - 1) & 1)) { if (tid == 0) *status = -4; return; }
      const uint32_t dst = smem_u32(ring) + s * TILE_BYTES + tid * 16;
      const unsigned char *gh = (const unsigned char *)(hi + (size_t)row * 64), *gl = (const unsigned char *)(lo + (size_t)row * 64);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + c * 2048), "l"(gh + c * 16) : "memory");
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16384 + c * 2048), "l"(gl + c * 16) : "memory");
      }
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&full[s])) : "memory");
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static float bf(__nv_bfloat16 x) { return __bfloat162float(x); }

int main() {
  void *p = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  EncodeTiledFn enc = (EncodeTiledFn)p;
  const int R = 32 * 8192, C = 64;
  std::vector<__nv_bfloat16> h((size_t)R * C);
  for (int r = 0; r < R; ++r)
    for (int c = 0; c < C; ++c) h[(size_t)r * C + c] = __float2bfloat16((float)(((r * 7 + c * 3) % 255) - 127) / 64.f);
  __nv_bfloat16 *d_hi, *d_lo;
  CK(cudaMalloc(&d_hi, (size_t)R * C * 2));
  CK(cudaMalloc(&d_lo, (size_t)R * C * 2));
  CK(cudaMemcpy(d_hi, h.data(), (size_t)R * C * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_lo, h.data(), (size_t)R * C * 2, cudaMemcpyHostToDevice));
  int *d_status;
  CK(cudaMalloc(&d_status, 4));

  std::vector<int> idx(128);
  srand(1);
  for (int i = 0; i < 128; ++i) idx[i] = rand() % R;
  int *d_idx;
  CK(cudaMalloc(&d_idx, 128 * 4));
  CK(cudaMemcpy(d_idx, idx.data(), 128 * 4, cudaMemcpyHostToDevice));

  CUtensorMap maps[2];
  int good_variant = -1;
  for (int variant = 0; variant < 1; ++variant) {   // box {64, 4} encodes but the copy faults (illegal instruction): gather4 wants box rows == 1
    const cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)R};
    const cuuint64_t strides[1] = {(cuuint64_t)C * 2};
    const cuuint32_t box[2] = {(cuuint32_t)C, variant == 0 ? 1u : 4u}, es[2] = {1, 1};
    CUresult r = enc(&maps[variant], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d_hi, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("[A] variant box={64,%d}: encode rc=%d\n", variant == 0 ? 1 : 4, (int)r);
    if (r != CUDA_SUCCESS) continue;
    uint4 *d_out;
    CK(cudaMalloc(&d_out, 16384));
    CK(cudaMemset(d_status, 0, 4));
    CK(cudaFuncSetAttribute(probe_a, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 1024));
    probe_a<<<1, 128, 16384 + 1024>>>(maps[variant], d_idx, d_out, d_status);
    cudaError_t e = cudaDeviceSynchronize();
    int st = 0;
    if (e != cudaSuccess) { printf("[A] variant %d: kernel error %s\n", variant, cudaGetErrorString(e)); return 1; }
    CK(cudaMemcpy(&st, d_status, 4, cudaMemcpyDeviceToHost));
    std::vector<uint16_t> o(8192);
    CK(cudaMemcpy(o.data(), d_out, 16384, cudaMemcpyDeviceToHost));
    // expected SW128: row i at byte i*128, 16-byte chunk c stored at chunk position c ^ (i & 7)
    int bad_swz = 0, bad_lin = 0;
    for (int i = 0; i < 128; ++i)
      for (int c = 0; c < 64; ++c) {
        const uint16_t want = *reinterpret_cast<const uint16_t *>(&h[(size_t)idx[i] * C + c]);
        const int chunk = c / 8, e8 = c % 8;
        if (o[i * 64 + ((chunk ^ (i & 7)) * 8) + e8] != want) ++bad_swz;
        if (o[i * 64 + c] != want) ++bad_lin;
      }
    printf("[A] variant %d: status=%d mismatches vs SW128 layout=%d, vs linear layout=%d\n", variant, st, bad_swz, bad_lin);
    if (st == 1 && bad_swz == 0 && good_variant < 0) good_variant = variant;
    cudaFree(d_out);
  }
  if (good_variant < 0) { printf("[A] no working gather4 variant\n"); }
  else printf("[A] OK: gather4 works with box variant %d\n", good_variant);

  // ---- B
  std::vector<__nv_bfloat16> wB(32 * 64);
  for (int i = 0; i < 32 * 64; ++i) wB[i] = __float2bfloat16((float)((i * 11) % 61 - 30) / 32.f);
  __nv_bfloat16 *d_w;
  float *d_o;
  CK(cudaMalloc(&d_w, 32 * 64 * 2));
  CK(cudaMalloc(&d_o, 128 * 32 * 4));
  CK(cudaMemcpy(d_w, wB.data(), 32 * 64 * 2, cudaMemcpyHostToDevice));
  std::vector<float> o(128 * 32);
  if (good_variant >= 0) {
    CK(cudaMemset(d_status, 0, 4));
    CK(cudaFuncSetAttribute(probe_b, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 4096 + 1024));
    probe_b<<<1, 128, 16384 + 4096 + 1024>>>(maps[good_variant], d_idx, d_w, d_o, d_status);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("[B] kernel error %s\n", cudaGetErrorString(e)); return 1; }
    int st;
    CK(cudaMemcpy(&st, d_status, 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(o.data(), d_o, 128 * 32 * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0;
    for (int i = 0; i < 128; ++i)
      for (int n = 0; n < 32; ++n) {
        double ref = 0;
        for (int k = 0; k < 64; ++k) ref += (double)bf(h[(size_t)idx[i] * C + k]) * bf(wB[n * 64 + k]);
        maxerr = fmax(maxerr, fabs(ref - o[i * 32 + n]));
      }
    printf("[B] SW128 gathered A x no-swizzle B: status=%d max|err|=%g %s\n", st, maxerr, maxerr < 1e-3 ? "OK" : "MISMATCH");
  }
  // ---- C
  {
    std::vector<__nv_bfloat16> A(128 * 16), w2(32 * 16);
    for (int i = 0; i < 128 * 16; ++i) A[i] = __float2bfloat16((float)((i * 13) % 97 - 48) / 32.f);
    for (int i = 0; i < 32 * 16; ++i) w2[i] = __float2bfloat16((float)((i * 5) % 53 - 26) / 16.f);
    __nv_bfloat16 *d_a, *d_w2;
    CK(cudaMalloc(&d_a, 128 * 16 * 2));
    CK(cudaMalloc(&d_w2, 32 * 16 * 2));
    CK(cudaMemcpy(d_a, A.data(), 128 * 16 * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_w2, w2.data(), 32 * 16 * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_status, 0, 4));
    probe_c<<<1, 128>>>(d_a, d_w2, d_o, d_status);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("[C] kernel error %s\n", cudaGetErrorString(e)); return 1; }
    int st;
    CK(cudaMemcpy(&st, d_status, 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(o.data(), d_o, 128 * 32 * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0;
    for (int i = 0; i < 128; ++i)
      for (int n = 0; n < 32; ++n) {
        double ref = 0;
        for (int k = 0; k < 16; ++k) ref += (double)bf(A[i * 16 + k]) * bf(w2[n * 16 + k]);
        maxerr = fmax(maxerr, fabs(ref - o[i * 32 + n]));
      }
    printf("[C] A in tensor memory (packed bf16 pairs, 32x32b st): status=%d max|err|=%g %s\n", st, maxerr, maxerr < 1e-3 ? "OK" : "MISMATCH");
  }
  // ---- D
  if (good_variant >= 0) {
    const int tiles = 512, ctas = 148;
    std::vector<int> big((size_t)ctas * tiles * 128);
    for (size_t t = 0; t < (size_t)ctas * tiles; ++t) {
      const int cloud = (int)(t % 32), centre = rand() % 8192;
      for (int i = 0; i < 128; ++i) {
        int j = centre + (rand() % 1024) - 512;       // a neighbourhood: rows within +-512 of the centroid's row
        j = j < 0 ? 0 : (j > 8191 ? 8191 : j);
        big[t * 128 + i] = cloud * 8192 + j;
      }
    }
    int *d_big;
    unsigned long long *d_sink;
    CK(cudaMalloc(&d_big, big.size() * 4));
    CK(cudaMalloc(&d_sink, 8));
    CK(cudaMemcpy(d_big, big.data(), big.size() * 4, cudaMemcpyHostToDevice));
    CUtensorMap mh, ml;
    const cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)R};
    const cuuint64_t strides[1] = {(cuuint64_t)C * 2};
    const cuuint32_t box[2] = {(cuuint32_t)C, good_variant == 0 ? 1u : 4u}, es[2] = {1, 1};
    enc(&mh, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d_hi, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    enc(&ml, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d_lo, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    const size_t smem = STAGES * TILE_BYTES + 1024;
    CK(cudaFuncSetAttribute(probe_d, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int mode = 0; mode < 5; ++mode) {
      CK(cudaMemset(d_status, 0, 4));
      for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        probe_d<<<ctas, mode == 4 ? 288 : 160, smem>>>(mh, ml, d_hi, d_lo, d_big, tiles, mode, d_sink, d_status);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("[D] mode %d kernel error %s\n", mode, cudaGetErrorString(e)); return 1; }
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        int st;
        CK(cudaMemcpy(&st, d_status, 4, cudaMemcpyDeviceToHost));
        if (rep == 2)
          printf("[D] mode %d (%s): status=%d %.3f ms for %d tiles/CTA x %d CTAs -> %.1f ns/tile/SM, %.0f GB/s aggregate\n", mode,
                 mode == 0 ? "gather4, 32 lanes x 2" : mode == 1 ? "gather4, lane 0 x 64" : mode == 2 ? "cp.async 16 B, thread = row, no-swizzle slabs" : mode == 3 ? "cp.async 16 B, SW128, 8 lanes = 1 row, 128 threads" : "cp.async 16 B, SW128, 8 lanes = 1 row, 256 threads", st, ms, tiles, ctas,
                 ms * 1e6 / tiles, (double)ctas * tiles * TILE_BYTES / (ms * 1e-3) / 1e9);
      }
    }
  }
  printf("probe done\n");
  return 0;
}
