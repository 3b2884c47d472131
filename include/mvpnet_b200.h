/*
 * mvpnet_b200.h — C ABI of the B200-native (sm_100a) MVPNet hot path.
 *
 * The reference (maxjaritz/mvpnet) has no C ABI: its lower boundary is six pybind11 torch
 * extension modules (mvpnet/ops/cuda/*.cpp) taking at::Tensor.  Each entry point below replaces
 * the at::Tensor-typed function named in its comment with plain device pointers + sizes, so any
 * host (the torch shim in mvpnet_b200/csrc/torch_ext.cpp, ctypes, or another runtime) can bind it.
 *
 * Conventions
 *   - all data pointers are DEVICE pointers on the current CUDA device; outputs are
 *     caller-allocated; nothing here allocates or synchronises
 *   - `stream` is a cudaStream_t (NULL = legacy default stream, which is what the reference's
 *     `<<<grid, block>>>` launches use)
 *   - dtype: MVP_F32 or MVP_F64 (the reference dispatches AT_DISPATCH_FLOATING_TYPES)
 *   - indices are int64 in and out (API-visible dtype of the reference)
 *   - return value: 0 on success; MVP_ERR_* (<0) for argument errors (the reference raises
 *     RuntimeError via TORCH_CHECK / CHECK_EQ there); >0 is a cudaError_t from the launch
 *   - mvp_last_error() returns a thread-local message for the last non-zero return
 *   - arithmetic contract: squared distances with d = key - query in the input dtype are
 *     float fma(dz,dz, fma(dy,dy, dx*dx)), double fma(dz,dz, fma(dx,dx, dy*dy)) — what nvcc -O2 emits
 *     for the reference loops on sm_100a (checked against the built reference kernels); comparisons strict
 */
#ifndef MVPNET_B200_H_
#define MVPNET_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MVP_F32 0
#define MVP_F64 1

#define MVP_ERR_INVALID_ARG (-1)
#define MVP_ERR_UNSUPPORTED (-2)
#define MVP_ERR_NULL (-3)

typedef void *mvp_stream_t; /* cudaStream_t */

const char *mvp_last_error(void);
/* ABI version of this header; bumped on any signature change. */
int mvp_abi_version(void);

/* Out-of-range gather/scatter indices (e.g. the -1 rows ball_query emits for a query with no
 * neighbour) do not fault: forward gathers produce 0, backward scatters skip the element, and a
 * device-side counter is incremented.  (Reference: device-side assert, group_points_kernel.cu:85,
 * interpolate_kernel.cu:59,170.)  This call copies the counter to the host (synchronises the
 * stream) and resets it. */
int mvp_index_errors_fetch_and_clear(mvp_stream_t stream, uint64_t *count);

/* ---- farthest point sampling ------------------------------------------------------------------
 * replaces fps_cuda.farthest_point_sample  (mvpnet/ops/cuda/fps.cpp:7-13, fps_kernel.cu:144-180)
 * points [B,N,D] contiguous, D in {2,3}; index out [B,M] int64; requires 0 < M <= N.
 * workspace: mvp_fps_workspace_bytes() bytes (may be 0 -> pass NULL). */
int64_t mvp_fps_workspace_bytes(int64_t B, int64_t N, int64_t D, int64_t M, int dtype);
int mvp_fps(const void *points, int64_t B, int64_t N, int64_t D, int64_t M, int dtype,
            int64_t *index, void *workspace, mvp_stream_t stream);

/* ---- ball query --------------------------------------------------------------------------------
 * replaces ball_query_cuda.ball_query (ball_query.cpp:7-15, ball_query_kernel.cu:147-187) and,
 * with distance != NULL, ball_query_distance_cuda.ball_query_distance
 * (ball_query_distance.cpp:7-15).  query [B,N1,3], key [B,N2,3] contiguous; index [B,N1,K];
 * distance [B,N1,K] in dtype or NULL.  First K keys in index order with d2 < r*r; tail padded with
 * the first hit (index only, distance pad = -1); no hit -> row of -1.
 * workspace: mvp_ball_query_workspace_bytes() bytes (0 -> pass NULL) -> exact uniform-grid search
 * (device-built, csrc/point_grid.cu) for the clouds it suits; NULL -> exhaustive search.  Identical results. */
int64_t mvp_ball_query_workspace_bytes(int64_t B, int64_t N1, int64_t N2, float radius, int64_t K, int dtype);
int mvp_ball_query(const void *query, const void *key, int64_t B, int64_t N1, int64_t N2,
                   float radius, int64_t K, int dtype, int64_t *index, void *distance,
                   void *workspace, mvp_stream_t stream);

/* ---- 3-NN --------------------------------------------------------------------------------------
 * replaces knn_distance_cuda.knn_distance (knn_distance.cpp:8-16, knn_distance_kernel.cu:154-196)
 * k must be 3 and N2 >= 3.  index [B,N1,3] int64, distance [B,N1,3] squared, ascending, lowest
 * key index first among equal distances.
 * workspace: mvp_knn_distance_workspace_bytes() bytes (0 -> pass NULL) -> exact uniform-grid search
 * (csrc/point_grid.cu); NULL -> exhaustive search.  Identical results. */
int64_t mvp_knn_distance_workspace_bytes(int64_t B, int64_t N1, int64_t N2, int dtype);
int mvp_knn_distance(const void *query, const void *key, int64_t B, int64_t N1, int64_t N2,
                     int64_t k, int dtype, int64_t *index, void *distance, void *workspace,
                     mvp_stream_t stream);

/* ---- group points ------------------------------------------------------------------------------
 * replaces group_points_cuda.group_points_forward / _backward (group_points.cpp:7-19,
 * group_points_kernel.cu:25-47, 99-145).
 * forward: out[b,c,n,k] = in[b,c,index[b,n,k]]; `in` may be strided (element strides sb,sc,sn —
 * the reference gathers through expand()ed views), index [B,N2,K] and out [B,C,N2,K] contiguous.
 * backward: grad_in[B,C,N1] (contiguous, fully overwritten) = scatter-add of grad_out. */
int mvp_group_points_forward(const void *in, int64_t sb, int64_t sc, int64_t sn,
                             const int64_t *index, int64_t B, int64_t C, int64_t N1, int64_t N2,
                             int64_t K, int dtype, void *out, mvp_stream_t stream);
int mvp_group_points_backward(const void *grad_out, const int64_t *index, int64_t B, int64_t C,
                              int64_t N1, int64_t N2, int64_t K, int dtype, void *grad_in,
                              mvp_stream_t stream);

/* ---- feature interpolate -----------------------------------------------------------------------
 * replaces interpolate_cuda.interpolate_forward / _backward (interpolate.cpp:8-22,
 * interpolate_kernel.cu:78-124, 184-230).  k == 3.
 * forward: out[b,c,n] = sum_k in[b,c,index[b,n,k]] * weight[b,n,k]; in [B,C,M] with element
 * strides (sb,sc,sm); index/weight [B,N,3] contiguous; out [B,C,N].
 * backward: grad_in [B,C,M] (overwritten) = scatter-add of grad_out[b,c,n]*weight[b,n,k]. */
int mvp_interpolate_forward(const void *in, int64_t sb, int64_t sc, int64_t sm,
                            const int64_t *index, const void *weight, int64_t B, int64_t C,
                            int64_t M, int64_t N, int dtype, void *out, mvp_stream_t stream);
int mvp_interpolate_backward(const void *grad_out, const int64_t *index, const void *weight,
                             int64_t B, int64_t C, int64_t M, int64_t N, int dtype, void *grad_in,
                             mvp_stream_t stream);

/* ---- unprojection of depth maps (data side of FeatureAggregation) ------------------------------
 * replaces depth2xyz + pose + masks, mvpnet/data/scannet_2d3d.py:33-39, 255-262, 273-281.
 * depth [B,nv,h,w] f32 metres; cam_inv [B,nv,3,3] f32 = inverse intrinsics (np.linalg.inv in the
 * reference); pose [B,nv,4,4] f32; chunk_box [B,4] f64 {x0,y0,x1,y1} or NULL (margin 0.1 applied
 * as in the reference).  Outputs (any may be NULL): xyz64 [B,nv*h*w,3] f64 (k-NN input),
 * xyz32 [B,nv,h,w,3] f32 (`image_xyz`), mask [B,nv*h*w] u8. */
int mvp_unproject(const float *depth, const float *cam_inv, const float *pose,
                  const double *chunk_box, int64_t B, int64_t nv, int64_t h, int64_t w,
                  double *xyz64, float *xyz32, uint8_t *mask, mvp_stream_t stream);

/* ---- decoding of the stored input formats on the device (what the dataset holds is what crosses PCIe) ----------
 * replaces the host-side conversions of mvpnet/data/scannet_2d3d.py:229-251: colour uint8 HWC [N,H,W,3] -> float32
 * CHW [N,3,H,W] as (u8 / 255 - mean[c]) / std[c] (float32, in that order; mean3 / std3 are HOST arrays of 3 floats);
 * depth uint16 millimetres -> float32 metres as float32(mm) / 1000. */
int mvp_decode_rgb_u8(const uint8_t *rgb_hwc, int64_t N, int64_t H, int64_t W, const float *mean3, const float *std3,
                      float *out_chw, mvp_stream_t stream);
int mvp_decode_depth_u16(const uint16_t *depth_mm, int64_t count, float *depth_m, mvp_stream_t stream);

/* ---- 2D->3D k-NN over valid pixels -------------------------------------------------------------
 * replaces sklearn NearestNeighbors(k,'ball_tree').fit(valid).kneighbors(points) + the remap to
 * flat pixel ids, mvpnet/data/scannet_2d3d.py:298-313.  query [B,nq,3] f64 (chunk points),
 * pix_xyz [B,P,3] f64, mask [B,P] u8; index out [B,nq,k] int64 flat pixel ids ascending by
 * distance, ties -> lowest id; dist2 [B,nq,k] f64 or NULL.  1 <= k <= 8.  A cloud with fewer than
 * k valid pixels yields -1 in the missing slots (sklearn raises there).
 * workspace: mvp_knn_pixels_workspace_bytes() bytes -> exact uniform-grid search (device-built);
 * NULL -> exhaustive search.  Both return identical results. */
int64_t mvp_knn_pixels_workspace_bytes(int64_t B, int64_t nq, int64_t P, int64_t k);
int mvp_knn_pixels(const double *query, const double *pix_xyz, const uint8_t *mask, int64_t B,
                   int64_t nq, int64_t P, int64_t k, int64_t *index, double *dist2, void *workspace,
                   mvp_stream_t stream);


/* ==== fused inference kernels (fp32) =============================================================
 * Layout between fused kernels is POINT-MAJOR: features [B, N, C] with one point's channels
 * contiguous (the reference's public tensors are channel-major [B, C, N]; the host transposes at the
 * two ends of the fast path only).
 *
 * mvp_mlp_chain_t: a SharedMLP with eval-mode BatchNorm folded in (common/nn/modules/mlp.py:38-75,
 * conv.py:29-51): layer l computes act(bias[l] + x . wt[l]) with wt[l] stored k-major [cin[l]][cout[l]],
 * zero padded; cin[0] is the (padded, multiple of 4) width of the built input row, cout[l] multiples
 * of 8, cin[l+1] == cout[l]; out_channels = true width of the last layer. */
#define MVP_MLP_MAX_LAYERS 6
typedef struct {
  int32_t num_layers;
  int32_t cin[MVP_MLP_MAX_LAYERS];
  int32_t cout[MVP_MLP_MAX_LAYERS];
  int32_t relu[MVP_MLP_MAX_LAYERS];
  const float *wt[MVP_MLP_MAX_LAYERS];
  const float *bias[MVP_MLP_MAX_LAYERS];
  int32_t out_channels;
} mvp_mlp_chain_t;

/* replaces QueryGrouper.forward + SharedMLP(ndim=2) + torch.max(dim=3) of SetAbstraction
 * (mvpnet/models/pn2/modules.py:20-37, 106-108) given the ball_query index:
 * row(b,m,k) = cat[feat[b,nbr[b,m,k],:], xyz[b,nbr[b,m,k],:] - new_xyz[b,m,:]]  (features first),
 * out[b,m,:] = max_k MLP(row).  feat [B,N,C] or NULL (C = 0), xyz [B,N,3], new_xyz [B,M,3],
 * nbr [B,M,K] int64 with K == 32, out [B,M,out_channels].  chain->cin[0] >= C + 3. */
int mvp_fused_set_abstraction(const float *feat, int64_t C, const float *xyz, const float *new_xyz,
                              const int64_t *nbr, int64_t B, int64_t N, int64_t M, int64_t K,
                              const mvp_mlp_chain_t *chain, float *out, mvp_stream_t stream);

/* replaces the two group_points of MVPNet3D.forward + FeatureAggregation.forward
 * (mvpnet/models/mvpnet_3d.py:100-109, 37-61): for point p and neighbour pixel j = knn[b,p,i]:
 * row = cat[feat2d[b, j, :], d = pix_xyz[b,j,:] - points[b,p,:], |d|^2]; out[b,p,:] = sum_i|max_i MLP(row).
 * feat2d is the 2D network output addressed in place: element (b, view, c, y, x) at
 * feat2d[(b*nv+view)*s_n + c*s_c + y*s_h + x*s_w] (NCHW or channels-last); pix_xyz [B,nv*h*w,3];
 * points [B,Np,3]; knn [B,Np,K] flat pixel ids, 1 <= K <= 4; out [B,Np,out_channels]. */
int mvp_fused_feature_aggregation(const float *feat2d, int64_t s_n, int64_t s_c, int64_t s_h, int64_t s_w,
                                  int64_t C, int64_t nv, int64_t h, int64_t w, const float *pix_xyz,
                                  const float *points, const int64_t *knn, int64_t B, int64_t Np, int64_t K,
                                  int reduce_sum, const mvp_mlp_chain_t *chain, float *out,
                                  mvp_stream_t stream);

/* replaces FeatureInterpolator.forward + SharedMLP(ndim=1) of FeaturePropagation
 * (mvpnet/models/pn2/modules.py:122-149, 178-186), optionally followed by the segmentation head
 * (pn2ssg.py:112-115) as extra chain layers, given the 3-NN (index, squared distance):
 * w_k = (1/max(d_k,eps)) / sum_k(1/max(d_k,eps)); row = cat[sum_k w_k * sparse[b,idx_k,:], skip[b,n,:]].
 * sparse_feat [B,Ns,Cs], idx/dist2 [B,Nd,3], skip [B,Nd,Cd] or NULL, out [B,Nd,out_channels]. */
int mvp_fused_feature_propagation(const float *sparse_feat, int64_t Cs, const int64_t *idx, const float *dist2,
                                  const float *skip, int64_t Cd, int64_t B, int64_t Ns, int64_t Nd, float eps,
                                  const mvp_mlp_chain_t *chain, float *out, mvp_stream_t stream);

/* ==== tensor-core variants of the fused kernels (tcgen05 / TMEM, bf16 hi/lo split x 3 products) ======
 * Same contracts as the three mvp_fused_* calls above.  mvp_tc_chain_t: layer l has K = k[l] input
 * channels (multiple of 16; k[0] >= width of the built row, zero padded; k[l+1] == n[l]) and N = n[l]
 * output channels (multiple of 16, <= 512).  Weights (BatchNorm folded) are split on the host into
 * bf16 w_hi = bf16(w), w_lo = bf16(w - w_hi) and stored in the kernel's operand order: for every
 * 256-wide block of output channels n0: [k/8][min(256, n - n0)][8] bf16 (K-slab major).  bias fp32 [n].
 * Feature channel counts must be multiples of 8.  mvp_tc_chain_supported() tells whether a chain's
 * activation tile fits shared memory / TMEM (mode 0 = set abstraction, 1 = aggregation, 2 = propagation). */
typedef struct {
  int32_t num_layers;
  int32_t k[MVP_MLP_MAX_LAYERS];
  int32_t n[MVP_MLP_MAX_LAYERS];
  int32_t relu[MVP_MLP_MAX_LAYERS];
  const void *w_hi[MVP_MLP_MAX_LAYERS];
  const void *w_lo[MVP_MLP_MAX_LAYERS];
  const float *bias[MVP_MLP_MAX_LAYERS];
  int32_t out_channels;
} mvp_tc_chain_t;

int mvp_tc_chain_supported(const mvp_tc_chain_t *chain, int mode);
int mvp_tc_fused_set_abstraction(const float *feat, int64_t C, const float *xyz, const float *new_xyz,
                                 const int64_t *nbr, int64_t B, int64_t N, int64_t M, int64_t K,
                                 const mvp_tc_chain_t *chain, float *out, mvp_stream_t stream);
int mvp_tc_fused_feature_aggregation(const float *feat2d, int64_t s_n, int64_t s_c, int64_t s_h, int64_t s_w,
                                     int64_t C, int64_t nv, int64_t h, int64_t w, const float *pix_xyz,
                                     const float *points, const int64_t *knn, int64_t B, int64_t Np, int64_t K,
                                     int reduce_sum, const mvp_tc_chain_t *chain, float *out,
                                     mvp_stream_t stream);
int mvp_tc_fused_feature_propagation(const float *sparse_feat, int64_t Cs, const int64_t *idx, const float *dist2,
                                     const float *skip, int64_t Cd, int64_t B, int64_t Ns, int64_t Nd, float eps,
                                     const mvp_tc_chain_t *chain, float *out, mvp_stream_t stream);

/* ==== training step (csrc/train_ops.cu) ================================================
 * Deterministic backward of group_points / feature_interpolate: same result as mvp_group_points_backward /
 * mvp_interpolate_backward (reference group_points_kernel.cu:50-89, interpolate_kernel.cu:131-174) but the entries
 * of every destination point are added in ascending (n, k) order instead of by atomicAdd, so the gradient is the same
 * bits on every run (fp32 only).  workspace: mvp_scatter_det_workspace_bytes() bytes of device memory. */
int64_t mvp_scatter_det_workspace_bytes(int64_t B, int64_t C, int64_t N1, int64_t N2, int64_t K, int weighted);
int mvp_group_points_backward_det(const float *grad_out, const int64_t *index, int64_t B, int64_t C, int64_t N1,
                                  int64_t N2, int64_t K, float *grad_in, void *workspace, mvp_stream_t stream);
int mvp_interpolate_backward_det(const float *grad_out, const int64_t *index, const float *weight, int64_t B, int64_t C,
                                 int64_t N1, int64_t N2, float *grad_in, void *workspace, mvp_stream_t stream);
/* SegLoss (mvpnet/models/loss.py:5-21): weighted cross entropy over logit [B, C, N] / label [B, N] (int64), mean over
 * the points whose label != ignore_index, C <= 64.  forward writes loss_out[0] = loss, loss_out[1] = sum of weights,
 * lse [B*N] (kept for backward) and, when conf != NULL, ADDS the confusion matrix conf[label * C + argmax] (uint64
 * [C, C]) that SegAccuracy / SegIoU (mvpnet/models/metric.py:26-73) are computed from.  Reductions are two-stage in a
 * fixed order: deterministic.  workspace: mvp_seg_loss_workspace_bytes() bytes.  backward: grad_logit [B, C, N] =
 * grad_scale[0] * d loss / d logit.  mvp_seg_confusion: the statistics alone (evaluation). */
int64_t mvp_seg_loss_workspace_bytes(int64_t B, int64_t N);
int mvp_seg_loss_forward(const float *logit, const int64_t *label, const float *weight, int64_t B, int64_t C, int64_t N,
                         int64_t ignore_index, float *lse, float *loss_out, uint64_t *conf, void *workspace,
                         mvp_stream_t stream);
int mvp_seg_loss_backward(const float *logit, const int64_t *label, const float *weight, const float *lse,
                          const float *loss_out, const float *grad_scale, int64_t B, int64_t C, int64_t N,
                          int64_t ignore_index, float *grad_logit, mvp_stream_t stream);
int mvp_seg_confusion(const float *logit, const int64_t *label, int64_t B, int64_t C, int64_t N, int64_t ignore_index,
                      uint64_t *conf, mvp_stream_t stream);

/* ==== second-generation fused gather kernels (csrc/tc2_mlp.cu): PRE-SPLIT inputs =====================
 * Same arithmetic and chain format as mvp_tc_fused_set_abstraction / _feature_aggregation, but the gathered
 * features arrive as two bf16 planes (hi = bf16(v), lo = bf16(v - hi)), rows of C = 64 / 128 / 256 channels
 * (feat_hi / feat_lo [B*N, C]; pix_hi / pix_lo [(B*nv), hp, wp, C] with hp >= h, wp >= w the padded image the 2D
 * network writes), are staged by cp.async straight into the swizzled MMA operand layout, inner layers keep their
 * activations in tensor memory, and the result is written fp32 (out_f32 [rows, out_channels], may be NULL) and / or
 * pre-split for the next gather (out_hi / out_lo, both or neither).  k[0] must be C + 16, every n <= 256, all weights
 * resident in shared memory: mvp_tc2_supported() tells; callers fall back to the mvp_tc_fused_* entry otherwise. */
int mvp_tc2_supported(const mvp_tc_chain_t *chain, int mode, int64_t C);
void mvp_tc2_prof_dump(const char *tag);   /* debug: prints the phase clocks collected under MVPNET_B200_TC2_PROF=1 */
int mvp_tc2_set_abstraction(const void *feat_hi, const void *feat_lo, int64_t C, const float *xyz, const float *new_xyz,
                            const int64_t *nbr, int64_t B, int64_t N, int64_t M, int64_t K, const mvp_tc_chain_t *chain,
                            float *out_f32, void *out_hi, void *out_lo, mvp_stream_t stream);
int mvp_tc2_feature_aggregation(const void *pix_hi, const void *pix_lo, int64_t C, int64_t nv, int64_t h, int64_t w,
                                int64_t hp, int64_t wp, const float *pix_xyz, const float *points, const int64_t *knn,
                                int64_t B, int64_t Np, int64_t K, int reduce_sum, const mvp_tc_chain_t *chain,
                                float *out_f32, void *out_hi, void *out_lo, mvp_stream_t stream);

/* ==== 3x3 / stride 1 / pad 1 convolution on the tensor cores (csrc/tc_conv.cu) ========================
 * The convolutions of the 2D network in front of FeatureAggregation (mvpnet/models/unet_resnet34.py:9-125;
 * called from MVPNet3D.forward, mvpnet_3d.py:94-99):
 *   out[n,y,x,:] = act(bias + sum_{ky,kx} W[:,:,ky,kx] . cat(x1, x2)[n, y+ky-1, x+kx-1, :] (+ residual[n,y,x,:]))
 * Activations are "split-planar": every fp32 value v is carried as bf16 hi = bf16(v) and lo = bf16(v - hi); a tensor
 * is two planes (hi, then lo), each [N][C/8][H][W][8] bf16 — or, when H <= 8, pair-interleaved
 * [ceil(N/2)][C/8][H][2][W][8] (image 2p and 2p+1 share rows; a missing partner must read as zeros).
 * mvp_planar_elems() gives the bf16 element count of both planes; mvp_split_planar / mvp_merge_planar convert from /
 * to fp32 NHWC.  x2 may be NULL (C2 = 0): cat([up, skip]) without the copy.  residual (split-planar, Cout channels)
 * may be NULL.  The result is written split-planar (out_planar), as fp32 NHWC (out_nhwc) and / or row-split (out_rows:
 * two bf16 planes hi, then lo, each NHWC — the pre-split pixel rows mvp_tc2_feature_aggregation gathers); at least one.
 * C1, C2, Cout multiples of 16 (Cout a multiple of Nt above it).  Weights (BatchNorm folded by the caller) are
 * split the same way and stored in the kernel's operand order [Cout/Nt][Cin/16][tap = ky*3+kx][hi|lo][2][Nt][8],
 * Nt = mvp_tc_conv3x3_nt(Cout) (= min(Cout, 256)), element (nb, c, tap, hl, k8, n, e) = W_hl[nb*Nt + n, c*16 + k8*8 + e, ky, kx]:
 * mvp_tc_conv3x3_weight_bytes() bytes. */
int64_t mvp_tc_conv3x3_weight_bytes(int64_t Cin, int64_t Cout);
int64_t mvp_tc_conv3x3_nt(int64_t Cout);
int64_t mvp_planar_elems(int64_t N, int64_t H, int64_t W, int64_t C);
int mvp_tc_conv3x3(const void *x1, int64_t C1, const void *x2, int64_t C2, int64_t N, int64_t H, int64_t W,
                   const void *w_packed, const float *bias, int64_t Cout, const void *residual, int relu,
                   void *out_planar, float *out_nhwc, void *out_rows, mvp_stream_t stream);
/* CTA-pair variant (csrc/tc_conv_pair.cu: tcgen05.mma.cta_group::2, two SMs per 256-pixel tile pair) for images of more
 * than 8 rows (output blocks of 32..256 channels) — mvp_tc_conv3x3_pair_supported(Cout, H).  Same arguments and
 * bit-identical results; the weights are packed with block width mvp_tc_conv3x3_nt(Cout) / 2 (each CTA of a pair streams
 * its own half block): [2 * Cout/Nt][Cin/16][tap][hi|lo][2][Nt/2][8], 128-byte aligned. */
int mvp_tc_conv3x3_pair_supported(int64_t Cout, int64_t H);
int mvp_tc_conv3x3_pair(const void *x1, int64_t C1, const void *x2, int64_t C2, int64_t N, int64_t H, int64_t W,
                        const void *w_packed_half, const float *bias, int64_t Cout, const void *residual, int relu,
                        void *out_planar, float *out_nhwc, void *out_rows, mvp_stream_t stream);
int mvp_split_planar(const float *nhwc, int64_t N, int64_t H, int64_t W, int64_t C, void *planar, mvp_stream_t stream);
int mvp_merge_planar(const void *planar, int64_t N, int64_t H, int64_t W, int64_t C, float *nhwc, mvp_stream_t stream);

/* ==== the other layers of the 2D network on the tensor cores (csrc/tc_convg.cu) =====================
 * mvp_tc_conv_general: x split-planar (N, Hi, Wi, Cin) -> out split-planar (N, Ho, Wo, Cout).
 *   mode 0  convolution with `ntaps` taps: out[n,y,x,:] = act(bias + sum_t W_t . x[n, y*stride + dy[t], x*stride + dx[t], :])
 *           (zero outside the input); stride 1 or 2; the caller passes the output grid.  Covers ResNet-34's three
 *           stride-2 3x3 convolutions and 1x1 stride-2 down-samples (torchvision BasicBlock; unet_resnet34.py:17-28)
 *           and, with mvp_unfold_stem, the 7x7 stem as seven row taps.
 *   mode 1  2x2 / stride-2 transposed convolution (unet_resnet34.py:31-60 deconv*): ntaps = 1, tap (0,0),
 *           Ho = 2 Hi, Wo = 2 Wi; GEMM column (2*ky + kx) * Cout + co.
 * Weights: [G/Nt][Cin/16][tap][hi|lo][2][Nt][8] bf16 with G = Cout (mode 0) or 4*Cout (mode 1), Nt = min(G, 256).
 * mvp_unfold_stem: fp32 NCHW 3-channel image -> split-planar (N, H, W, 32): channel kx*3 + c = image[n, c, y, x+kx-3].
 * mvp_maxpool3x3s2_planar: 3x3 / stride 2 / pad 1 max-pool, split-planar in and out ((H-1)/2+1 x (W-1)/2+1). */
int mvp_tc_conv_general(const void *x, int64_t Cin, int64_t N, int64_t Hi, int64_t Wi, int mode, int stride, int ntaps,
                        const int *dy, const int *dx, int64_t Ho, int64_t Wo, const void *w_packed, const float *bias,
                        int64_t Cout, int relu, void *out_planar, mvp_stream_t stream);
int mvp_unfold_stem(const float *image_nchw, int64_t N, int64_t H, int64_t W, void *planar32, mvp_stream_t stream);
int mvp_maxpool3x3s2_planar(const void *x, int64_t N, int64_t H, int64_t W, int64_t C, void *out, mvp_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MVPNET_B200_H_ */
