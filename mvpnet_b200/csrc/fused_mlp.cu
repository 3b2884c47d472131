// Fused "build rows -> pointwise MLP chain -> reduce over neighbours" kernel for sm_100a (fp32 SIMT).
//
// One kernel covers the three places where the reference runs a SharedMLP over gathered rows and
// round-trips every intermediate (B, C, M, K) tensor through HBM:
//   MODE_SA  SetAbstraction       modules.py:20-37,106-108: group_points(xyz, feature) by ball_query
//            index, subtract the centroid, cat([feature, xyz]), 3 x (1x1 conv + BN + ReLU), max over K
//   MODE_FA  FeatureAggregation   mvpnet_3d.py:37-61 (+ the two group_points of :100-109): gather k pixel
//            features and pixel xyz, relation [dxyz, |dxyz|^2], cat([feature, relation]), MLP, sum|max over k
//   MODE_FP  FeaturePropagation   modules.py:122-149,178-186 (+ the segmentation head pn2ssg.py:112-115):
//            inverse-squared-distance weights from the 3-NN, interpolate, cat([interp, skip]), MLP
// The grouped tensor never exists in HBM: rows are built in shared memory, every layer reads its
// input tile from shared memory and writes the next tile to shared memory, and only the reduced
// output leaves the SM.  BatchNorm (eval) is folded into the weights on the host.
//
// Data layout between kernels of the fast path is POINT-MAJOR ([B, N, C]: one point's channels are
// contiguous) so that a row gather is one coalesced 4*C-byte read.
//
// Tile: R = 32 * RM rows per CTA.  Lane = row within a 32-row slab, each thread owns RM rows x 8
// output channels (RM*8 accumulators); the NW warps of the CTA split the output channels.  x is read
// from shared memory as float4 along k (conflict-free: row stride = 4 * odd), weights are k-major
// ([Cin][Cout], 32 contiguous bytes per warp per k) and come through L1 as warp-broadcast loads.
#include "common.cuh"

namespace mvp {

constexpr int MLP_MAX_LAYERS = 6;
enum { MODE_SA = 0, MODE_FA = 1, MODE_FP = 2 };
enum { REDUCE_MAX = 0, REDUCE_SUM = 1 };

struct MlpChain {
  int num_layers;
  int cin[MLP_MAX_LAYERS];    // padded to a multiple of 4
  int cout[MLP_MAX_LAYERS];   // padded to a multiple of 8; cin[l+1] == cout[l]
  int relu[MLP_MAX_LAYERS];
  const float *wt[MLP_MAX_LAYERS];    // [cin][cout] k-major, BN folded, zero padded
  const float *bias[MLP_MAX_LAYERS];  // [cout]
  int out_channels;           // true channel count of the last layer (<= cout[last])
};

struct BuildArgs {
  // common
  long long rows_out;         // number of output rows (centroids / points) over the whole batch
  int feat_channels;          // C of the gathered / interpolated features
  // MODE_SA: feat [B,N,C] (or null), xyz [B,N,3], new_xyz [B,M,3], nbr [B,M,32]
  // MODE_FA: feat2d (strided 4-D), pix_xyz [B,P,3], points [B,Np,3], nbr = knn [B,Np,K]
  // MODE_FP: feat = sparse feats [B,Ns,Cs], nbr = idx [B,Nd,3], dist [B,Nd,3], skip [B,Nd,Cd]
  const float *feat;
  const float *xyz;
  const float *new_xyz;
  const int64_t *nbr;
  const float *dist;
  const float *skip;
  int skip_channels;
  long long n_src;            // N (SA), P = nv*h*w (FA), Ns (FP): source points per cloud
  long long n_out;            // M (SA), Np (FA), Nd (FP): output rows per cloud
  int k;                      // neighbours per output row (SA: 32, FA: <= 4, FP: 3)
  int reduce;                 // REDUCE_MAX / REDUCE_SUM (FA); SA is always max
  float eps;                  // FP: clamp of the squared distance
  // FA: feat2d addressing: element (b, v, c, y, x) at feat + (b*nv+v)*s_n + c*s_c + y*s_h + x*s_w
  long long s_n, s_c, s_h, s_w;
  int hw, w, nv;
};

__host__ __device__ inline int row_stride(int c) { return ((c / 4) & 1) ? c : c + 4; }  // 4 * odd

// ---------------------------------------------------------------------------------------------
// input builders: fill `in` ([R][S0], zero padded to cin[0]) for the tile starting at row_base
// ---------------------------------------------------------------------------------------------
template <int RM, int NW>
__device__ __forceinline__ void build_sa(const BuildArgs &a, float *in, int S0, int cin0, long long tile) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = a.feat_channels;
  for (int r = warp; r < 32 * RM; r += NW) {
    const int i = r >> 5, k = r & 31;
    const long long gid = tile * RM + i;         // centroid over the whole batch
    float *row = in + (size_t)(i * 32 + k) * S0;
    if (gid >= a.rows_out) {
      for (int c = lane; c < cin0; c += 32) row[c] = 0.f;
      continue;
    }
    const long long b = gid / a.n_out;
    const long long j = a.nbr[gid * 32 + k];
    const bool ok = j >= 0 && j < a.n_src;
    if (C > 0) {
      const float *src = a.feat + ((size_t)b * a.n_src + (ok ? j : 0)) * C;
      if ((C & 3) == 0) {
        for (int c = lane * 4; c < C; c += 128) {
          float4 v = ok ? __ldg(reinterpret_cast<const float4 *>(src + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
          *reinterpret_cast<float4 *>(row + c) = v;
        }
      } else {
        for (int c = lane; c < C; c += 32) row[c] = ok ? __ldg(src + c) : 0.f;
      }
    }
    if (lane < cin0 - C) {
      float v = 0.f;
      if (lane < 3 && ok) v = __fsub_rn(__ldg(a.xyz + ((size_t)b * a.n_src + j) * 3 + lane), __ldg(a.new_xyz + gid * 3 + lane));
      row[C + lane] = v;
    }
  }
}

template <int RM, int NW>
__device__ __forceinline__ void build_fa(const BuildArgs &a, float *in, int S0, int cin0, long long tile) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = a.feat_channels;
  for (int r = warp; r < 32 * RM; r += NW) {
    const int i = r >> 5, p = r & 31;            // neighbour i of point p
    const long long pid = tile * 32 + p;
    float *row = in + (size_t)(i * 32 + p) * S0;
    if (pid >= a.rows_out || i >= a.k) {
      for (int c = lane; c < cin0; c += 32) row[c] = 0.f;
      continue;
    }
    const long long b = pid / a.n_out;
    const long long j = a.nbr[pid * a.k + i];
    const bool ok = j >= 0 && j < a.n_src;
    const long long jj = ok ? j : 0;
    const int v = (int)(jj / a.hw), pix = (int)(jj - (long long)v * a.hw);
    const int y = pix / a.w, x = pix - y * a.w;
    const float *src = a.feat + ((size_t)b * a.nv + v) * a.s_n + (size_t)y * a.s_h + (size_t)x * a.s_w;
    for (int c = lane; c < C; c += 32) row[c] = ok ? __ldg(src + (size_t)c * a.s_c) : 0.f;
    if (lane == 0) {
      float dx = 0.f, dy = 0.f, dz = 0.f;
      if (ok) {
        const float *s = a.xyz + ((size_t)b * a.n_src + j) * 3;
        const float *t = a.new_xyz + pid * 3;
        dx = __fsub_rn(__ldg(s), __ldg(t)); dy = __fsub_rn(__ldg(s + 1), __ldg(t + 1)); dz = __fsub_rn(__ldg(s + 2), __ldg(t + 2));
      }
      row[C] = dx; row[C + 1] = dy; row[C + 2] = dz;
      // torch.sum(diff ** 2, dim=1): plain products and adds, x, y, z order
      row[C + 3] = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    }
    for (int c = C + 4 + lane; c < cin0; c += 32) row[c] = 0.f;
  }
}

template <int RM, int NW>
__device__ __forceinline__ void build_fp(const BuildArgs &a, float *in, int S0, int cin0, long long tile) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Cs = a.feat_channels, Cd = a.skip_channels;
  for (int r = warp; r < 32 * RM; r += NW) {
    const long long pid = tile * (32 * RM) + r;
    float *row = in + (size_t)r * S0;
    if (pid >= a.rows_out) {
      for (int c = lane; c < cin0; c += 32) row[c] = 0.f;
      continue;
    }
    const long long b = pid / a.n_out;
    // inverse squared-distance weights, modules.py:135-140: 1/clamp(d, eps), normalised
    float w[3];
    long long j[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      j[k] = a.nbr[pid * 3 + k];
      w[k] = __fdiv_rn(1.0f, fmaxf(__ldg(a.dist + pid * 3 + k), a.eps));
      if (j[k] < 0 || j[k] >= a.n_src) { j[k] = 0; w[k] = 0.f; }
    }
    const float norm = __fadd_rn(__fadd_rn(w[0], w[1]), w[2]);
#pragma unroll
    for (int k = 0; k < 3; ++k) w[k] = __fdiv_rn(w[k], norm);
    const float *s0 = a.feat + ((size_t)b * a.n_src + j[0]) * Cs;
    const float *s1 = a.feat + ((size_t)b * a.n_src + j[1]) * Cs;
    const float *s2 = a.feat + ((size_t)b * a.n_src + j[2]) * Cs;
    if ((Cs & 3) == 0) {
      for (int c = lane * 4; c < Cs; c += 128) {
        const float4 v0 = __ldg(reinterpret_cast<const float4 *>(s0 + c));
        const float4 v1 = __ldg(reinterpret_cast<const float4 *>(s1 + c));
        const float4 v2 = __ldg(reinterpret_cast<const float4 *>(s2 + c));
        float4 o;  // interpolate_kernel.cu:54-61 accumulation order
        o.x = __fmaf_rn(v2.x, w[2], __fmaf_rn(v1.x, w[1], __fmul_rn(v0.x, w[0])));
        o.y = __fmaf_rn(v2.y, w[2], __fmaf_rn(v1.y, w[1], __fmul_rn(v0.y, w[0])));
        o.z = __fmaf_rn(v2.z, w[2], __fmaf_rn(v1.z, w[1], __fmul_rn(v0.z, w[0])));
        o.w = __fmaf_rn(v2.w, w[2], __fmaf_rn(v1.w, w[1], __fmul_rn(v0.w, w[0])));
        *reinterpret_cast<float4 *>(row + c) = o;
      }
    } else {
      for (int c = lane; c < Cs; c += 32)
        row[c] = __fmaf_rn(__ldg(s2 + c), w[2], __fmaf_rn(__ldg(s1 + c), w[1], __fmul_rn(__ldg(s0 + c), w[0])));
    }
    if (Cd > 0) {
      const float *sk = a.skip + (size_t)pid * Cd;
      for (int c = lane; c < Cd; c += 32) row[Cs + c] = __ldg(sk + c);
    }
    for (int c = Cs + Cd + lane; c < cin0; c += 32) row[c] = 0.f;
  }
}

// ---------------------------------------------------------------------------------------------
// one layer on the tile: out[r][c] = act(bias[c] + sum_k in[r][k] * wt[k][c])
// ---------------------------------------------------------------------------------------------
template <int RM, int NW>
__device__ __forceinline__ void mlp_layer(const float *__restrict__ in, int Sin, float *__restrict__ out, int Sout,
                                          const float *__restrict__ wt, const float *__restrict__ bias, int cin,
                                          int cout, bool relu) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int cb = warp * 8; cb < cout; cb += NW * 8) {
    float acc[RM][8];
    {
      const float4 b0 = __ldg(reinterpret_cast<const float4 *>(bias + cb));
      const float4 b1 = __ldg(reinterpret_cast<const float4 *>(bias + cb + 4));
#pragma unroll
      for (int i = 0; i < RM; ++i) {
        acc[i][0] = b0.x; acc[i][1] = b0.y; acc[i][2] = b0.z; acc[i][3] = b0.w;
        acc[i][4] = b1.x; acc[i][5] = b1.y; acc[i][6] = b1.z; acc[i][7] = b1.w;
      }
    }
    const float *wp = wt + cb;
#pragma unroll 2
    for (int k = 0; k < cin; k += 4) {
      float4 xv[RM];
#pragma unroll
      for (int i = 0; i < RM; ++i) xv[i] = *reinterpret_cast<const float4 *>(in + (size_t)(i * 32 + lane) * Sin + k);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const float4 w0 = __ldg(reinterpret_cast<const float4 *>(wp + (size_t)(k + kk) * cout));
        const float4 w1 = __ldg(reinterpret_cast<const float4 *>(wp + (size_t)(k + kk) * cout + 4));
#pragma unroll
        for (int i = 0; i < RM; ++i) {
          const float x = kk == 0 ? xv[i].x : kk == 1 ? xv[i].y : kk == 2 ? xv[i].z : xv[i].w;
          acc[i][0] = fmaf(x, w0.x, acc[i][0]); acc[i][1] = fmaf(x, w0.y, acc[i][1]);
          acc[i][2] = fmaf(x, w0.z, acc[i][2]); acc[i][3] = fmaf(x, w0.w, acc[i][3]);
          acc[i][4] = fmaf(x, w1.x, acc[i][4]); acc[i][5] = fmaf(x, w1.y, acc[i][5]);
          acc[i][6] = fmaf(x, w1.z, acc[i][6]); acc[i][7] = fmaf(x, w1.w, acc[i][7]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < RM; ++i) {
      float4 o0 = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      float4 o1 = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
      if (relu) {
        o0.x = fmaxf(o0.x, 0.f); o0.y = fmaxf(o0.y, 0.f); o0.z = fmaxf(o0.z, 0.f); o0.w = fmaxf(o0.w, 0.f);
        o1.x = fmaxf(o1.x, 0.f); o1.y = fmaxf(o1.y, 0.f); o1.z = fmaxf(o1.z, 0.f); o1.w = fmaxf(o1.w, 0.f);
      }
      float *o = out + (size_t)(i * 32 + lane) * Sout + cb;
      *reinterpret_cast<float4 *>(o) = o0;
      *reinterpret_cast<float4 *>(o + 4) = o1;
    }
  }
}

template <int MODE, int RM, int NW>
__global__ void __launch_bounds__(NW * 32)
fused_mlp_kernel(const BuildArgs a, const MlpChain m, float *__restrict__ out, int buf0_floats) {
  extern __shared__ __align__(16) float smem_f[];
  float *const bufA = smem_f, *const bufB = smem_f + buf0_floats;
  const long long tile = blockIdx.x;
  const int S0 = row_stride(m.cin[0]);
  if (MODE == MODE_SA) build_sa<RM, NW>(a, bufA, S0, m.cin[0], tile);
  else if (MODE == MODE_FA) build_fa<RM, NW>(a, bufA, S0, m.cin[0], tile);
  else build_fp<RM, NW>(a, bufA, S0, m.cin[0], tile);
  __syncthreads();
  for (int l = 0; l < m.num_layers; ++l) {
    mlp_layer<RM, NW>((l & 1) ? bufB : bufA, row_stride(m.cin[l]), (l & 1) ? bufA : bufB, row_stride(m.cout[l]), m.wt[l], m.bias[l],
                      m.cin[l], m.cout[l], m.relu[l] != 0);
    __syncthreads();
  }
  const float *fin = (m.num_layers & 1) ? bufB : bufA;
  const int Sf = row_stride(m.cout[m.num_layers - 1]);
  const int Co = m.out_channels;
  if (MODE == MODE_SA) {
    // max over the 32 neighbours (rows i*32 .. i*32+31) of centroid i
    for (int e = threadIdx.x; e < RM * Co; e += NW * 32) {
      const int i = e / Co, c = e - i * Co;
      const long long gid = tile * RM + i;
      if (gid >= a.rows_out) continue;
      const float *col = fin + (size_t)(i * 32) * Sf + c;
      float v = col[0];
#pragma unroll 8
      for (int k = 1; k < 32; ++k) v = fmaxf(v, col[(size_t)k * Sf]);
      out[gid * Co + c] = v;
    }
  } else if (MODE == MODE_FA) {
    for (int e = threadIdx.x; e < 32 * Co; e += NW * 32) {
      const int p = e / Co, c = e - p * Co;
      const long long pid = tile * 32 + p;
      if (pid >= a.rows_out) continue;
      float v = fin[(size_t)p * Sf + c];
      for (int i = 1; i < a.k; ++i) {
        const float u = fin[(size_t)(i * 32 + p) * Sf + c];
        v = a.reduce == REDUCE_SUM ? __fadd_rn(v, u) : fmaxf(v, u);
      }
      out[pid * Co + c] = v;
    }
  } else {
    for (int e = threadIdx.x; e < 32 * RM * Co; e += NW * 32) {
      const int r = e / Co, c = e - r * Co;
      const long long pid = tile * (32 * RM) + r;
      if (pid >= a.rows_out) continue;
      out[pid * Co + c] = fin[(size_t)r * Sf + c];
    }
  }
}

template <int MODE, int RM, int NW>
static int launch_fused(const BuildArgs &a, const MlpChain &m, float *out, long long tiles, size_t buf0, size_t buf1,
                        cudaStream_t stream) {
  auto kern = fused_mlp_kernel<MODE, RM, NW>;
  const size_t smem = (buf0 + buf1) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("fused_mlp: smem attribute (%zu B): %s", smem, cudaGetErrorString(e)); return (int)e; }
  kern<<<(unsigned)tiles, NW * 32, smem, stream>>>(a, m, out, (int)buf0);
  return launch_status("fused_mlp");
}

}  // namespace mvp

// ---------------------------------------------------------------------------------------------
// C ABI (declared in include/mvpnet_b200.h)
// ---------------------------------------------------------------------------------------------
static int mvp_check_chain(const mvp_mlp_chain_t *c, int cin_true) {
  using namespace mvp;
  MVP_REQUIRE(c, MVP_ERR_NULL, "fused_mlp: null chain");
  MVP_REQUIRE(c->num_layers >= 1 && c->num_layers <= MLP_MAX_LAYERS, MVP_ERR_INVALID_ARG, "fused_mlp: 1..6 layers");
  MVP_REQUIRE(c->cin[0] >= cin_true && c->cin[0] % 4 == 0, MVP_ERR_INVALID_ARG,
              "fused_mlp: cin[0]=%d must be >= %d and a multiple of 4", c->cin[0], cin_true);
  for (int l = 0; l < c->num_layers; ++l) {
    MVP_REQUIRE(c->cout[l] % 8 == 0 && c->cout[l] > 0, MVP_ERR_INVALID_ARG, "fused_mlp: cout must be a multiple of 8");
    MVP_REQUIRE(c->wt[l] && c->bias[l], MVP_ERR_NULL, "fused_mlp: null weights");
    MVP_REQUIRE((((uintptr_t)c->wt[l]) & 15) == 0 && (((uintptr_t)c->bias[l]) & 15) == 0, MVP_ERR_INVALID_ARG,
                "fused_mlp: weights must be 16-byte aligned");
    if (l > 0) MVP_REQUIRE(c->cin[l] == c->cout[l - 1], MVP_ERR_INVALID_ARG, "fused_mlp: cin[l] must equal cout[l-1]");
  }
  MVP_REQUIRE(c->out_channels > 0 && c->out_channels <= c->cout[c->num_layers - 1], MVP_ERR_INVALID_ARG,
              "fused_mlp: bad out_channels");
  return 0;
}

static void mvp_to_chain(const mvp_mlp_chain_t *c, mvp::MlpChain *m, size_t *buf0, size_t *buf1, int rows) {
  m->num_layers = c->num_layers;
  m->out_channels = c->out_channels;
  size_t b0 = 0, b1 = 0;
  for (int l = 0; l < c->num_layers; ++l) {
    m->cin[l] = c->cin[l]; m->cout[l] = c->cout[l]; m->relu[l] = c->relu[l];
    m->wt[l] = c->wt[l]; m->bias[l] = c->bias[l];
    const size_t in_sz = (size_t)rows * mvp::row_stride(c->cin[l]), out_sz = (size_t)rows * mvp::row_stride(c->cout[l]);
    if (l & 1) { if (in_sz > b1) b1 = in_sz; if (out_sz > b0) b0 = out_sz; }
    else { if (in_sz > b0) b0 = in_sz; if (out_sz > b1) b1 = out_sz; }
  }
  *buf0 = b0; *buf1 = b1;
}

#define MVP_SMEM_LIMIT (227 * 1024)

extern "C" int mvp_fused_set_abstraction(const float *feat, int64_t C, const float *xyz, const float *new_xyz,
                                         const int64_t *nbr, int64_t B, int64_t N, int64_t M, int64_t K,
                                         const mvp_mlp_chain_t *chain, float *out, mvp_stream_t stream) {
  using namespace mvp;
  MVP_REQUIRE(K == 32, MVP_ERR_UNSUPPORTED, "fused_set_abstraction: max_neighbors must be 32 (got %lld)", (long long)K);
  MVP_REQUIRE(B >= 0 && N > 0 && M >= 0 && C >= 0, MVP_ERR_INVALID_ARG, "fused_set_abstraction: bad sizes");
  if (int rc = mvp_check_chain(chain, (int)C + 3)) return rc;
  if (B * M == 0) return 0;
  MVP_REQUIRE(xyz && new_xyz && nbr && out && (feat || C == 0), MVP_ERR_NULL, "fused_set_abstraction: null pointer");
  MVP_REQUIRE(C == 0 || ((uintptr_t)feat & 15) == 0, MVP_ERR_INVALID_ARG, "fused_set_abstraction: feat must be 16-byte aligned");
  BuildArgs a = {};
  a.rows_out = B * M; a.feat_channels = (int)C; a.feat = feat; a.xyz = xyz; a.new_xyz = new_xyz; a.nbr = nbr;
  a.n_src = N; a.n_out = M; a.k = 32;
  MlpChain m;
  size_t b0, b1;
  // widest tile whose two ping-pong buffers fit
  mvp_to_chain(chain, &m, &b0, &b1, 128);
  if ((b0 + b1) * 4 <= MVP_SMEM_LIMIT) {
    const long long tiles = (a.rows_out + 3) / 4;
    return launch_fused<MODE_SA, 4, 16>(a, m, out, tiles, b0, b1, (cudaStream_t)stream);
  }
  mvp_to_chain(chain, &m, &b0, &b1, 64);
  if ((b0 + b1) * 4 <= MVP_SMEM_LIMIT) {
    const long long tiles = (a.rows_out + 1) / 2;
    return launch_fused<MODE_SA, 2, 16>(a, m, out, tiles, b0, b1, (cudaStream_t)stream);
  }
  mvp_to_chain(chain, &m, &b0, &b1, 32);
  MVP_REQUIRE((b0 + b1) * 4 <= MVP_SMEM_LIMIT, MVP_ERR_UNSUPPORTED, "fused_set_abstraction: layers too wide for shared memory");
  return launch_fused<MODE_SA, 1, 16>(a, m, out, a.rows_out, b0, b1, (cudaStream_t)stream);
}

extern "C" int mvp_fused_feature_aggregation(const float *feat2d, int64_t s_n, int64_t s_c, int64_t s_h, int64_t s_w,
                                             int64_t C, int64_t nv, int64_t h, int64_t w, const float *pix_xyz,
                                             const float *points, const int64_t *knn, int64_t B, int64_t Np, int64_t K,
                                             int reduce_sum, const mvp_mlp_chain_t *chain, float *out,
                                             mvp_stream_t stream) {
  using namespace mvp;
  MVP_REQUIRE(K >= 1 && K <= 4, MVP_ERR_UNSUPPORTED, "fused_feature_aggregation: k must be in [1, 4] (got %lld)", (long long)K);
  MVP_REQUIRE(B >= 0 && Np >= 0 && C > 0 && nv > 0 && h > 0 && w > 0, MVP_ERR_INVALID_ARG, "fused_feature_aggregation: bad sizes");
  if (int rc = mvp_check_chain(chain, (int)C + 4)) return rc;
  if (B * Np == 0) return 0;
  MVP_REQUIRE(feat2d && pix_xyz && points && knn && out, MVP_ERR_NULL, "fused_feature_aggregation: null pointer");
  BuildArgs a = {};
  a.rows_out = B * Np; a.feat_channels = (int)C; a.feat = feat2d; a.xyz = pix_xyz; a.new_xyz = points; a.nbr = knn;
  a.n_src = nv * h * w; a.n_out = Np; a.k = (int)K; a.reduce = reduce_sum ? REDUCE_SUM : REDUCE_MAX;
  a.s_n = s_n; a.s_c = s_c; a.s_h = s_h; a.s_w = s_w; a.hw = (int)(h * w); a.w = (int)w; a.nv = (int)nv;
  MlpChain m;
  size_t b0, b1;
  const long long tiles = (a.rows_out + 31) / 32;
  if (K == 3) {
    mvp_to_chain(chain, &m, &b0, &b1, 96);
    MVP_REQUIRE((b0 + b1) * 4 <= MVP_SMEM_LIMIT, MVP_ERR_UNSUPPORTED, "fused_feature_aggregation: layers too wide");
    return launch_fused<MODE_FA, 3, 8>(a, m, out, tiles, b0, b1, (cudaStream_t)stream);
  }
  mvp_to_chain(chain, &m, &b0, &b1, 128);
  MVP_REQUIRE((b0 + b1) * 4 <= MVP_SMEM_LIMIT, MVP_ERR_UNSUPPORTED, "fused_feature_aggregation: layers too wide");
  return launch_fused<MODE_FA, 4, 8>(a, m, out, tiles, b0, b1, (cudaStream_t)stream);
}

extern "C" int mvp_fused_feature_propagation(const float *sparse_feat, int64_t Cs, const int64_t *idx, const float *dist2,
                                             const float *skip, int64_t Cd, int64_t B, int64_t Ns, int64_t Nd, float eps,
                                             const mvp_mlp_chain_t *chain, float *out, mvp_stream_t stream) {
  using namespace mvp;
  MVP_REQUIRE(B >= 0 && Ns > 0 && Nd >= 0 && Cs > 0 && Cd >= 0, MVP_ERR_INVALID_ARG, "fused_feature_propagation: bad sizes");
  if (int rc = mvp_check_chain(chain, (int)(Cs + Cd))) return rc;
  if (B * Nd == 0) return 0;
  MVP_REQUIRE(sparse_feat && idx && dist2 && out && (skip || Cd == 0), MVP_ERR_NULL, "fused_feature_propagation: null pointer");
  MVP_REQUIRE(((uintptr_t)sparse_feat & 15) == 0, MVP_ERR_INVALID_ARG, "fused_feature_propagation: features must be 16-byte aligned");
  BuildArgs a = {};
  a.rows_out = B * Nd; a.feat_channels = (int)Cs; a.feat = sparse_feat; a.nbr = idx; a.dist = dist2; a.skip = skip;
  a.skip_channels = (int)Cd; a.n_src = Ns; a.n_out = Nd; a.k = 3; a.eps = eps;
  MlpChain m;
  size_t b0, b1;
  mvp_to_chain(chain, &m, &b0, &b1, 128);
  if ((b0 + b1) * 4 <= MVP_SMEM_LIMIT)
    return launch_fused<MODE_FP, 4, 16>(a, m, out, (a.rows_out + 127) / 128, b0, b1, (cudaStream_t)stream);
  mvp_to_chain(chain, &m, &b0, &b1, 64);
  if ((b0 + b1) * 4 <= MVP_SMEM_LIMIT)
    return launch_fused<MODE_FP, 2, 16>(a, m, out, (a.rows_out + 63) / 64, b0, b1, (cudaStream_t)stream);
  mvp_to_chain(chain, &m, &b0, &b1, 32);
  MVP_REQUIRE((b0 + b1) * 4 <= MVP_SMEM_LIMIT, MVP_ERR_UNSUPPORTED, "fused_feature_propagation: layers too wide for shared memory");
  return launch_fused<MODE_FP, 1, 16>(a, m, out, (a.rows_out + 31) / 32, b0, b1, (cudaStream_t)stream);
}
