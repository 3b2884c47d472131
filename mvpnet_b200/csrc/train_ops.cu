// Training-step kernels for sm_100a: what train_mvpnet_3d.py:158-180 runs around the network.
//
//  * DETERMINISTIC backward of group_points / feature_interpolate.  The reference scatters with atomicAdd
//    (group_points_kernel.cu:50-89, interpolate_kernel.cu:131-174): the summation order, and therefore the gradient
//    bits, change from run to run.  Here the (row, neighbour) entries of every cloud are bucketed by DESTINATION point
//    (count -> scan -> fill), the incoming gradient is transposed once to row-major [entry row][channel] so that a
//    destination reads whole 256-byte rows, and one warp per destination adds its entries in ascending entry order:
//    the result is the same bits on every run and equals a sequential loop over (n, k) — which is what oracle/
//    restates.  No atomics touch floating-point data.
//  * SegLoss (mvpnet/models/loss.py:5-21: weighted cross entropy, ignore_index, mean over the non-ignored points),
//    forward + backward, and the SegAccuracy / SegIoU statistics (mvpnet/models/metric.py:26-73: argmax, confusion
//    matrix) in one pass over the logits; the reductions are two-stage with a fixed order (deterministic).
#include "common.cuh"

namespace mvp {
namespace trn {

// ------------------------------------------------------------------------------------------------ destination lists
__global__ void dl_count_kernel(const int64_t *__restrict__ index, int64_t E, int N1, int *__restrict__ cnt /*[B][N1]*/) {
  const int b = blockIdx.y;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int64_t j = index[(int64_t)b * E + e];
  if (j < 0 || j >= N1) { atomicAdd(&g_index_errors, 1ULL); return; }
  atomicAdd(cnt + (size_t)b * N1 + j, 1);
}

// exclusive scan of cnt[b][0..N1) -> off[b][0..N1]; cnt is reset to zero (it becomes the fill cursor); one CTA per cloud
__global__ void __launch_bounds__(1024) dl_scan_kernel(int *__restrict__ cnt, int *__restrict__ off, int N1) {
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int *c = cnt + (size_t)b * N1, *o = off + (size_t)b * (N1 + 1);
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < N1; base += 1024) {
    const int i = base + tid;
    const int v = i < N1 ? c[i] : 0;
    int x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, x, d);
      if (lane >= d) x += u;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    if (warp == 0) {
      int w = s_warp[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, w, d);
        if (lane >= d) w += u;
      }
      s_warp[lane] = w;
    }
    __syncthreads();
    const int carry = s_carry, before = warp > 0 ? s_warp[warp - 1] : 0;
    if (i < N1) { o[i] = carry + before + x - v; c[i] = 0; }
    __syncthreads();
    if (tid == 1023) s_carry = carry + before + x;
    __syncthreads();
  }
  if (tid == 0) o[N1] = s_carry;
}

__global__ void dl_fill_kernel(const int64_t *__restrict__ index, int64_t E, int N1, const int *__restrict__ off, int *__restrict__ cur,
                               int *__restrict__ list /*[B][E]*/) {
  const int b = blockIdx.y;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int64_t j = index[(int64_t)b * E + e];
  if (j < 0 || j >= N1) return;
  const int pos = off[(size_t)b * (N1 + 1) + j] + atomicAdd(cur + (size_t)b * N1 + j, 1);   // order fixed later by the consumer
  list[(size_t)b * E + pos] = (int)e;
}

// (B, C, R) -> (B, R, C), 32 x 32 tiles
__global__ void __launch_bounds__(256) transpose_kernel(const float *__restrict__ in, int C, int64_t R, float *__restrict__ out) {
  __shared__ float t[32][33];
  const int b = blockIdx.z;
  const int64_t r0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float *ib = in + (size_t)b * C * R;
  float *ob = out + (size_t)b * R * C;
  for (int i = ty; i < 32; i += 8)
    if (c0 + i < C && r0 + tx < R) t[i][tx] = __ldcs(ib + (size_t)(c0 + i) * R + r0 + tx);   // read once
  __syncthreads();
  for (int i = ty; i < 32; i += 8)
    if (r0 + i < R && c0 + tx < C) ob[(size_t)(r0 + i) * C + c0 + tx] = t[tx][i];
}

// One warp per destination point j, 64 channels per CTA pass: adds the entries of j in ascending entry order.
// WEIGHTED: entry e = (n, k) reads row n of gT with weight w[e] (feature_interpolate); otherwise row e, weight 1 (group_points).
template <bool WEIGHTED>
__global__ void __launch_bounds__(1024) dl_gather_kernel(const float *__restrict__ gT /*[B][R][C]*/, const float *__restrict__ weight /*[B][E]*/,
                                                         const int *__restrict__ off, const int *__restrict__ list, int C, int N1, int64_t E,
                                                         int64_t R, int K, float *__restrict__ grad_in /*[B][C][N1]*/) {
  __shared__ float tile[64][33];
  const int b = blockIdx.z, c0 = blockIdx.y * 64;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + warp;
  float a0 = 0.f, a1 = 0.f;
  if (j < N1) {
    const int beg = off[(size_t)b * (N1 + 1) + j], L = off[(size_t)b * (N1 + 1) + j + 1] - beg;
    const int *lst = list + (size_t)b * E + beg;
    const float *gb = gT + (size_t)b * R * C;
    const bool ok0 = c0 + lane < C, ok1 = c0 + 32 + lane < C;
    int last = -1;
    if (L <= 32) {
      const int mine = lane < L ? lst[lane] : 0x7fffffff;
      for (int t = 0; t < L; ++t) {
        const int e = __reduce_min_sync(0xffffffffu, mine > last ? mine : 0x7fffffff);
        last = e;
        const int64_t row = WEIGHTED ? e / K : e;
        const float w = WEIGHTED ? __ldg(weight + (size_t)b * E + e) : 1.f;
        const float *g = gb + (size_t)row * C + c0;
        if (ok0) a0 = __fadd_rn(a0, WEIGHTED ? __fmul_rn(__ldg(g + lane), w) : __ldg(g + lane));
        if (ok1) a1 = __fadd_rn(a1, WEIGHTED ? __fmul_rn(__ldg(g + 32 + lane), w) : __ldg(g + 32 + lane));
      }
    } else {
      for (int t = 0; t < L; ++t) {   // long lists (rare): selection by repeated minimum
        int cand = 0x7fffffff;
        for (int i = lane; i < L; i += 32) {
          const int e = lst[i];
          if (e > last && e < cand) cand = e;
        }
        const int e = __reduce_min_sync(0xffffffffu, cand);
        last = e;
        const int64_t row = WEIGHTED ? e / K : e;
        const float w = WEIGHTED ? __ldg(weight + (size_t)b * E + e) : 1.f;
        const float *g = gb + (size_t)row * C + c0;
        if (ok0) a0 = __fadd_rn(a0, WEIGHTED ? __fmul_rn(__ldg(g + lane), w) : __ldg(g + lane));
        if (ok1) a1 = __fadd_rn(a1, WEIGHTED ? __fmul_rn(__ldg(g + 32 + lane), w) : __ldg(g + 32 + lane));
      }
    }
  }
  tile[lane][warp] = a0;
  tile[32 + lane][warp] = a1;
  __syncthreads();
  // coalesced along j: warp w writes channels w and w + 32 of the 32 destinations of this CTA
  const int jj = blockIdx.x * 32 + lane;
  if (jj < N1) {
    if (c0 + warp < C) grad_in[((size_t)b * C + c0 + warp) * N1 + jj] = tile[warp][lane];
    if (c0 + 32 + warp < C) grad_in[((size_t)b * C + c0 + 32 + warp) * N1 + jj] = tile[32 + warp][lane];
  }
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct DetWs {
  int *cnt, *off, *list;
  float *gT;
  size_t bytes;
};
static DetWs det_layout(void *ws, int64_t B, int64_t C, int64_t N1, int64_t E, int64_t R) {
  DetWs w;
  size_t o = 0;
  unsigned char *p = (unsigned char *)ws;
  w.cnt = (int *)(p + o); o += align256((size_t)B * N1 * 4);
  w.off = (int *)(p + o); o += align256((size_t)B * (N1 + 1) * 4);
  w.list = (int *)(p + o); o += align256((size_t)B * E * 4);
  w.gT = (float *)(p + o); o += align256((size_t)B * R * C * 4);
  w.bytes = o;
  return w;
}

static int det_backward(const float *grad_out, const int64_t *index, const float *weight, int64_t B, int64_t C, int64_t N1, int64_t N2,
                        int64_t K, bool weighted, float *grad_in, void *workspace, cudaStream_t stream, const char *what) {
  const int64_t E = N2 * K, R = weighted ? N2 : E;
  MVP_REQUIRE(B >= 0 && C >= 0 && N1 >= 0 && N2 >= 0 && K >= 0, MVP_ERR_INVALID_ARG, "%s: negative size", what);
  if (B == 0 || C == 0 || N1 == 0) return 0;
  MVP_REQUIRE(grad_in, MVP_ERR_NULL, "%s: null pointer", what);
  if (E == 0) {
    cudaMemsetAsync(grad_in, 0, (size_t)B * C * N1 * 4, stream);
    return launch_status(what);
  }
  MVP_REQUIRE(grad_out && index && workspace && (!weighted || weight), MVP_ERR_NULL, "%s: null pointer", what);
  MVP_REQUIRE(B <= 65535 && E < (1LL << 31) && N1 < (1LL << 31) && C <= 65535 * 32, MVP_ERR_UNSUPPORTED, "%s: tensor too large", what);
  DetWs w = det_layout(workspace, B, C, N1, E, R);
  cudaMemsetAsync(w.cnt, 0, (size_t)B * N1 * 4, stream);
  const dim3 ge((unsigned)((E + 255) / 256), (unsigned)B);
  dl_count_kernel<<<ge, 256, 0, stream>>>(index, E, (int)N1, w.cnt);
  dl_scan_kernel<<<(unsigned)B, 1024, 0, stream>>>(w.cnt, w.off, (int)N1);
  dl_fill_kernel<<<ge, 256, 0, stream>>>(index, E, (int)N1, w.off, w.cnt, w.list);
  const dim3 gt((unsigned)((R + 31) / 32), (unsigned)((C + 31) / 32), (unsigned)B);
  transpose_kernel<<<gt, 256, 0, stream>>>(grad_out, (int)C, R, w.gT);
  const dim3 gg((unsigned)((N1 + 31) / 32), (unsigned)((C + 63) / 64), (unsigned)B);
  if (weighted) dl_gather_kernel<true><<<gg, 1024, 0, stream>>>(w.gT, weight, w.off, w.list, (int)C, (int)N1, E, R, (int)K, grad_in);
  else dl_gather_kernel<false><<<gg, 1024, 0, stream>>>(w.gT, nullptr, w.off, w.list, (int)C, (int)N1, E, R, (int)K, grad_in);
  return launch_status(what);
}

// ------------------------------------------------------------------------------------------------ SegLoss / metrics
constexpr int SL_THREADS = 256;
constexpr int SL_MAXC = 64;

// per point: log-sum-exp, weighted NLL, argmax; per CTA: partial sums (fixed-order tree) and a confusion matrix
template <bool WITH_LOSS>
__global__ void __launch_bounds__(SL_THREADS) seg_stats_kernel(const float *__restrict__ logit /*[B][C][N]*/, const int64_t *__restrict__ label,
                                                               const float *__restrict__ weight, int C, int64_t N, int64_t total /*B*N*/,
                                                               long long ignore_index, float *__restrict__ lse_out,
                                                               double *__restrict__ partial /*[grid][2]*/, unsigned long long *__restrict__ conf /*[C][C]*/) {
  extern __shared__ unsigned int s_conf[];     // [C*C]
  __shared__ double s_a[SL_THREADS / 32], s_b[SL_THREADS / 32];
  for (int i = threadIdx.x; i < C * C; i += SL_THREADS) s_conf[i] = 0u;
  __syncthreads();
  const int64_t p = (int64_t)blockIdx.x * SL_THREADS + threadIdx.x;
  double wnll = 0.0, wsum = 0.0;
  if (p < total) {
    const int64_t b = p / N, n = p - b * N;
    const float *x = logit + (size_t)b * C * N + n;
    float mx = __ldg(x);
    int am = 0;
    for (int c = 1; c < C; ++c) {
      const float v = __ldg(x + (size_t)c * N);
      if (v > mx) { mx = v; am = c; }           // first maximum, like torch.argmax
    }
    const long long y = label[p];
    const bool valid = y != ignore_index && y >= 0 && y < C;
    if (WITH_LOSS) {
      float s = 0.f;
      for (int c = 0; c < C; ++c) s += expf(__ldg(x + (size_t)c * N) - mx);
      const float lse = mx + logf(s);
      lse_out[p] = lse;
      if (valid) {
        const float w = weight ? __ldg(weight + y) : 1.f;
        wnll = (double)w * (double)(lse - __ldg(x + (size_t)y * N));
        wsum = (double)w;
      }
    }
    if (valid) atomicAdd(&s_conf[(int)y * C + am], 1u);
  }
  if (WITH_LOSS) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      wnll += __shfl_xor_sync(0xffffffffu, wnll, o);
      wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
    }
    if ((threadIdx.x & 31) == 0) { s_a[threadIdx.x >> 5] = wnll; s_b[threadIdx.x >> 5] = wsum; }
  }
  __syncthreads();
  if (WITH_LOSS && threadIdx.x == 0) {
    double a = 0.0, bsum = 0.0;
    for (int w = 0; w < SL_THREADS / 32; ++w) { a += s_a[w]; bsum += s_b[w]; }
    partial[2 * (size_t)blockIdx.x] = a;
    partial[2 * (size_t)blockIdx.x + 1] = bsum;
  }
  if (conf != nullptr)
    for (int i = threadIdx.x; i < C * C; i += SL_THREADS)
      if (s_conf[i]) atomicAdd(conf + i, (unsigned long long)s_conf[i]);   // integer: order-independent
}

// fixed-order second stage: out[0] = mean weighted NLL, out[1] = sum of weights over the non-ignored points
__global__ void __launch_bounds__(1024) seg_loss_finish_kernel(const double *__restrict__ partial, int nblocks, float *__restrict__ out) {
  __shared__ double s_a[1024], s_b[1024];
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < nblocks; i += 1024) { a += partial[2 * (size_t)i]; b += partial[2 * (size_t)i + 1]; }
  s_a[threadIdx.x] = a; s_b[threadIdx.x] = b;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (threadIdx.x < o) { s_a[threadIdx.x] += s_a[threadIdx.x + o]; s_b[threadIdx.x] += s_b[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { out[0] = (float)(s_a[0] / s_b[0]); out[1] = (float)s_b[0]; }
}

// d loss / d logit[b,c,n] = g * w[y] * (softmax_c - [c == y]) / sum_w   (0 for ignored points)
__global__ void __launch_bounds__(SL_THREADS) seg_loss_bwd_kernel(const float *__restrict__ logit, const int64_t *__restrict__ label,
                                                                  const float *__restrict__ weight, const float *__restrict__ lse,
                                                                  const float *__restrict__ loss_out /*[1] = sum_w*/, const float *__restrict__ gscale,
                                                                  int C, int64_t N, int64_t total, long long ignore_index, float *__restrict__ grad) {
  const int64_t p = (int64_t)blockIdx.x * SL_THREADS + threadIdx.x;
  if (p >= total) return;
  const int64_t b = p / N, n = p - b * N;
  const float *x = logit + (size_t)b * C * N + n;
  float *g = grad + (size_t)b * C * N + n;
  const long long y = label[p];
  const bool valid = y != ignore_index && y >= 0 && y < C;
  float k = 0.f;
  if (valid) k = __ldg(gscale) * (weight ? __ldg(weight + y) : 1.f) / __ldg(loss_out + 1);
  const float l = lse[p];
  for (int c = 0; c < C; ++c) {
    const float sm = expf(__ldg(x + (size_t)c * N) - l);
    g[(size_t)c * N] = valid ? k * (sm - (c == (int)y ? 1.f : 0.f)) : 0.f;
  }
}

}  // namespace trn
}  // namespace mvp

extern "C" int64_t mvp_scatter_det_workspace_bytes(int64_t B, int64_t C, int64_t N1, int64_t N2, int64_t K, int weighted) {
  if (B <= 0 || C <= 0 || N1 <= 0 || N2 * K <= 0) return 256;
  return (int64_t)mvp::trn::det_layout(nullptr, B, C, N1, N2 * K, weighted ? N2 : N2 * K).bytes;
}

extern "C" int mvp_group_points_backward_det(const float *grad_out, const int64_t *index, int64_t B, int64_t C, int64_t N1, int64_t N2,
                                             int64_t K, float *grad_in, void *workspace, mvp_stream_t stream) {
  return mvp::trn::det_backward(grad_out, index, nullptr, B, C, N1, N2, K, false, grad_in, workspace, (cudaStream_t)stream,
                                "group_points_backward_det");
}

extern "C" int mvp_interpolate_backward_det(const float *grad_out, const int64_t *index, const float *weight, int64_t B, int64_t C, int64_t N1,
                                            int64_t N2, float *grad_in, void *workspace, mvp_stream_t stream) {
  return mvp::trn::det_backward(grad_out, index, weight, B, C, N1, N2, 3, true, grad_in, workspace, (cudaStream_t)stream,
                                "interpolate_backward_det");
}

extern "C" int64_t mvp_seg_loss_workspace_bytes(int64_t B, int64_t N) {
  const int64_t blocks = (B * N + mvp::trn::SL_THREADS - 1) / mvp::trn::SL_THREADS;
  return (blocks > 0 ? blocks : 1) * 16;
}

extern "C" int mvp_seg_loss_forward(const float *logit, const int64_t *label, const float *weight, int64_t B, int64_t C, int64_t N,
                                    int64_t ignore_index, float *lse, float *loss_out, uint64_t *conf, void *workspace, mvp_stream_t stream_) {
  using namespace mvp;
  cudaStream_t stream = (cudaStream_t)stream_;
  MVP_REQUIRE(B >= 0 && N >= 0 && C >= 1 && C <= trn::SL_MAXC, MVP_ERR_UNSUPPORTED, "seg_loss: 1..%d classes", trn::SL_MAXC);
  MVP_REQUIRE(loss_out, MVP_ERR_NULL, "seg_loss: null pointer");
  const int64_t total = B * N;
  if (total == 0) { cudaMemsetAsync(loss_out, 0, 8, stream); return launch_status("seg_loss"); }
  MVP_REQUIRE(logit && label && lse && workspace, MVP_ERR_NULL, "seg_loss: null pointer");
  const int64_t blocks = (total + trn::SL_THREADS - 1) / trn::SL_THREADS;
  MVP_REQUIRE(blocks < (1LL << 31), MVP_ERR_UNSUPPORTED, "seg_loss: too many points");
  trn::seg_stats_kernel<true><<<(unsigned)blocks, trn::SL_THREADS, (size_t)C * C * 4, stream>>>(logit, label, weight, (int)C, N, total, ignore_index, lse,
                                                                                                 (double *)workspace, (unsigned long long *)conf);
  trn::seg_loss_finish_kernel<<<1, 1024, 0, stream>>>((const double *)workspace, (int)blocks, loss_out);
  return launch_status("seg_loss");
}

extern "C" int mvp_seg_loss_backward(const float *logit, const int64_t *label, const float *weight, const float *lse, const float *loss_out,
                                     const float *grad_scale, int64_t B, int64_t C, int64_t N, int64_t ignore_index, float *grad_logit,
                                     mvp_stream_t stream) {
  using namespace mvp;
  MVP_REQUIRE(B >= 0 && N >= 0 && C >= 1 && C <= trn::SL_MAXC, MVP_ERR_UNSUPPORTED, "seg_loss: 1..%d classes", trn::SL_MAXC);
  const int64_t total = B * N;
  if (total == 0) return 0;
  MVP_REQUIRE(logit && label && lse && loss_out && grad_scale && grad_logit, MVP_ERR_NULL, "seg_loss_backward: null pointer");
  const int64_t blocks = (total + trn::SL_THREADS - 1) / trn::SL_THREADS;
  trn::seg_loss_bwd_kernel<<<(unsigned)blocks, trn::SL_THREADS, 0, (cudaStream_t)stream>>>(logit, label, weight, lse, loss_out, grad_scale, (int)C, N, total,
                                                                                           ignore_index, grad_logit);
  return launch_status("seg_loss_backward");
}

extern "C" int mvp_seg_confusion(const float *logit, const int64_t *label, int64_t B, int64_t C, int64_t N, int64_t ignore_index, uint64_t *conf,
                                 mvp_stream_t stream) {
  using namespace mvp;
  MVP_REQUIRE(B >= 0 && N >= 0 && C >= 1 && C <= trn::SL_MAXC, MVP_ERR_UNSUPPORTED, "seg_confusion: 1..%d classes", trn::SL_MAXC);
  const int64_t total = B * N;
  if (total == 0) return 0;
  MVP_REQUIRE(logit && label && conf, MVP_ERR_NULL, "seg_confusion: null pointer");
  const int64_t blocks = (total + trn::SL_THREADS - 1) / trn::SL_THREADS;
  trn::seg_stats_kernel<false><<<(unsigned)blocks, trn::SL_THREADS, (size_t)C * C * 4, (cudaStream_t)stream>>>(logit, label, nullptr, (int)C, N, total,
                                                                                                                ignore_index, nullptr, nullptr,
                                                                                                                (unsigned long long *)conf);
  return launch_status("seg_confusion");
}
