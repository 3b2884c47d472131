// Torch shim over the C ABI (include/mvpnet_b200.h).
//
// Exposes, as sub-modules of one extension, the six pybind11 modules the reference builds in
// mvpnet/ops/setup.py:13-67 with the same function names and argument order
// (mvpnet/ops/cuda/{fps,ball_query,ball_query_distance,group_points,knn_distance,interpolate}.cpp).
// The shim only validates (TORCH_CHECK -> RuntimeError like the reference), allocates outputs, takes
// the CURRENT stream and a device guard (the reference launches on the legacy default stream without
// a guard; this is stricter and compatible), and forwards raw pointers to the C ABI.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include <cstdlib>
#include <vector>

#include "../../include/mvpnet_b200.h"

namespace {

inline void check_rc(int rc) { TORCH_CHECK(rc == 0, mvp_last_error(), " (mvpnet_b200 rc=", rc, ")"); }

inline int dtype_of(const at::Tensor &t, const char *name) {
  if (t.scalar_type() == at::kFloat) return MVP_F32;
  if (t.scalar_type() == at::kDouble) return MVP_F64;
  TORCH_CHECK(false, name, " must be float32 or float64, got ", t.scalar_type());
  return -1;
}

#define CHECK_CUDA(x) TORCH_CHECK((x).is_cuda(), #x " must be a CUDA tensor")
#define CHECK_CONTIGUOUS(x) TORCH_CHECK((x).is_contiguous(), #x " must be contiguous")
#define CHECK_INPUT(x) \
  CHECK_CUDA(x);       \
  CHECK_CONTIGUOUS(x)

inline mvp_stream_t cur_stream() { return (mvp_stream_t)at::cuda::getCurrentCUDAStream().stream(); }

// fps.cpp:7-13
at::Tensor farthest_point_sample(const at::Tensor points, const int64_t num_centroids) {
  CHECK_INPUT(points);
  TORCH_CHECK(points.dim() == 3, "points must be (B, N, D)");
  const auto B = points.size(0), N = points.size(1), D = points.size(2);
  TORCH_CHECK(D == 2 || D == 3, "Only support dim=2 or dim=3");
  TORCH_CHECK(num_centroids > 0, "Check failed: num_centroids > 0");
  TORCH_CHECK(N >= num_centroids, "Check failed: num_points >= num_centroids");
  const int dt = dtype_of(points, "points");
  c10::cuda::CUDAGuard guard(points.device());
  auto index = at::empty({B, num_centroids}, points.options().dtype(at::kLong));
  const int64_t ws = mvp_fps_workspace_bytes(B, N, D, num_centroids, dt);
  at::Tensor work;
  if (ws > 0) work = at::empty({ws}, points.options().dtype(at::kByte));
  check_rc(mvp_fps(points.data_ptr(), B, N, D, num_centroids, dt, index.data_ptr<int64_t>(),
                   ws > 0 ? work.data_ptr() : nullptr, cur_stream()));
  return index;
}

void check_query_key(const at::Tensor &query, const at::Tensor &key) {
  CHECK_INPUT(query);
  CHECK_INPUT(key);
  TORCH_CHECK(query.dim() == 3 && key.dim() == 3, "query/key must be (B, N, 3)");
  TORCH_CHECK(query.size(2) == 3, "Check failed: query.size(2) == 3");
  TORCH_CHECK(key.size(2) == 3, "Check failed: key.size(2) == 3");
  TORCH_CHECK(key.size(0) == query.size(0), "Check failed: key.size(0) == batch_size");
  TORCH_CHECK(query.scalar_type() == key.scalar_type(), "query and key must have the same dtype");
  TORCH_CHECK(query.device() == key.device(), "query and key must be on the same device");
}

// workspace of the exact uniform-grid searches (csrc/point_grid.cu); MVPNET_B200_GRID=0 forces the exhaustive kernels
static bool &grid_enabled() {
  static bool enabled = [] { const char *e = getenv("MVPNET_B200_GRID"); return !(e && e[0] == '0'); }();
  return enabled;
}
// tests flip this to compare the grid search with the exhaustive kernels in one process; returns the previous setting
bool set_grid_search(bool on) { const bool was = grid_enabled(); grid_enabled() = on; return was; }

at::Tensor grid_workspace(const at::Tensor &like, int64_t bytes) {
  if (!grid_enabled() || bytes <= 0) return at::Tensor();
  return at::empty({bytes}, like.options().dtype(at::kByte));
}

// ball_query.cpp:7-15
at::Tensor ball_query(const at::Tensor query, const at::Tensor key, const float radius, const int64_t max_neighbors) {
  check_query_key(query, key);
  TORCH_CHECK(max_neighbors > 0, "max_neighbors must be positive");
  const int dt = dtype_of(query, "query");
  c10::cuda::CUDAGuard guard(query.device());
  auto index = at::empty({query.size(0), query.size(1), max_neighbors}, query.options().dtype(at::kLong));
  at::Tensor work = grid_workspace(query, mvp_ball_query_workspace_bytes(query.size(0), query.size(1), key.size(1), radius, max_neighbors, dt));
  check_rc(mvp_ball_query(query.data_ptr(), key.data_ptr(), query.size(0), query.size(1), key.size(1), radius,
                          max_neighbors, dt, index.data_ptr<int64_t>(), nullptr, work.defined() ? work.data_ptr() : nullptr,
                          cur_stream()));
  return index;
}

// ball_query_distance.cpp:7-15
std::vector<at::Tensor> ball_query_distance(const at::Tensor query, const at::Tensor key, const float radius,
                                            const int64_t max_neighbors) {
  check_query_key(query, key);
  TORCH_CHECK(max_neighbors > 0, "max_neighbors must be positive");
  const int dt = dtype_of(query, "query");
  c10::cuda::CUDAGuard guard(query.device());
  auto index = at::empty({query.size(0), query.size(1), max_neighbors}, query.options().dtype(at::kLong));
  auto distance = at::empty({query.size(0), query.size(1), max_neighbors}, query.options());
  at::Tensor work = grid_workspace(query, mvp_ball_query_workspace_bytes(query.size(0), query.size(1), key.size(1), radius, max_neighbors, dt));
  check_rc(mvp_ball_query(query.data_ptr(), key.data_ptr(), query.size(0), query.size(1), key.size(1), radius,
                          max_neighbors, dt, index.data_ptr<int64_t>(), distance.data_ptr(),
                          work.defined() ? work.data_ptr() : nullptr, cur_stream()));
  return {index, distance};
}

// knn_distance.cpp:8-16
std::vector<at::Tensor> knn_distance(const at::Tensor query, const at::Tensor key, const int64_t k) {
  check_query_key(query, key);
  TORCH_CHECK(key.size(1) >= k, "Check failed: num_key >= k");
  TORCH_CHECK(k == 3, "Only support 3-NN.");
  const int dt = dtype_of(query, "query");
  c10::cuda::CUDAGuard guard(query.device());
  auto index = at::empty({query.size(0), query.size(1), k}, query.options().dtype(at::kLong));
  auto distance = at::empty({query.size(0), query.size(1), k}, query.options());
  at::Tensor work = grid_workspace(query, mvp_knn_distance_workspace_bytes(query.size(0), query.size(1), key.size(1), dt));
  check_rc(mvp_knn_distance(query.data_ptr(), key.data_ptr(), query.size(0), query.size(1), key.size(1), k, dt,
                            index.data_ptr<int64_t>(), distance.data_ptr(), work.defined() ? work.data_ptr() : nullptr,
                            cur_stream()));
  return {index, distance};
}

// group_points.cpp:7-19
at::Tensor group_points_forward(const at::Tensor input, const at::Tensor index) {
  CHECK_CUDA(input);
  CHECK_CUDA(index);
  TORCH_CHECK(input.dim() == 3, "Check failed: input.dim() == 3");
  TORCH_CHECK(index.dim() == 3, "Check failed: index.dim() == 3");
  TORCH_CHECK(index.size(0) == input.size(0), "Check failed: index.size(0) == batch_size");
  TORCH_CHECK(index.scalar_type() == at::kLong, "index must be int64");
  const int dt = dtype_of(input, "input");
  c10::cuda::CUDAGuard guard(input.device());
  const auto idx = index.contiguous();
  const auto B = input.size(0), C = input.size(1), N1 = input.size(2), N2 = idx.size(1), K = idx.size(2);
  auto out = at::empty({B, C, N2, K}, input.options());
  check_rc(mvp_group_points_forward(input.data_ptr(), input.stride(0), input.stride(1), input.stride(2),
                                    idx.data_ptr<int64_t>(), B, C, N1, N2, K, dt, out.data_ptr(), cur_stream()));
  return out;
}

at::Tensor group_points_backward(const at::Tensor grad_output, const at::Tensor index, const int64_t num_points) {
  CHECK_CUDA(grad_output);
  CHECK_CUDA(index);
  TORCH_CHECK(grad_output.dim() == 4, "Check failed: grad_output.dim() == 4");
  TORCH_CHECK(index.dim() == 3, "Check failed: index.dim() == 3");
  TORCH_CHECK(index.size(0) == grad_output.size(0), "Check failed: index.size(0) == batch_size");
  TORCH_CHECK(index.size(1) == grad_output.size(2), "Check failed: index.size(1) == num_select");
  TORCH_CHECK(index.size(2) == grad_output.size(3), "Check failed: index.size(2) == k");
  TORCH_CHECK(index.scalar_type() == at::kLong, "index must be int64");
  const int dt = dtype_of(grad_output, "grad_output");
  c10::cuda::CUDAGuard guard(grad_output.device());
  const auto g = grad_output.contiguous();
  const auto idx = index.contiguous();
  const auto B = g.size(0), C = g.size(1), N2 = g.size(2), K = g.size(3);
  auto grad_input = at::empty({B, C, num_points}, g.options());
  check_rc(mvp_group_points_backward(g.data_ptr(), idx.data_ptr<int64_t>(), B, C, num_points, N2, K, dt,
                                     grad_input.data_ptr(), cur_stream()));
  return grad_input;
}

void check_interp(const at::Tensor &x, const at::Tensor &index, const at::Tensor &weight, int64_t num_select) {
  CHECK_CUDA(x);
  CHECK_CUDA(index);
  CHECK_CUDA(weight);
  TORCH_CHECK(x.dim() == 3 && index.dim() == 3 && weight.dim() == 3, "interpolate: tensors must be 3-D");
  TORCH_CHECK(index.size(0) == x.size(0), "Check failed: index.size(0) == batch_size");
  TORCH_CHECK(index.size(2) == 3, "Check failed: k == 3");
  TORCH_CHECK(index.size(1) == num_select, "Check failed: index.size(1) == num_select");
  TORCH_CHECK(weight.size(0) == x.size(0), "Check failed: weight.size(0) == batch_size");
  TORCH_CHECK(weight.size(1) == num_select, "Check failed: weight.size(1) == num_select");
  TORCH_CHECK(weight.size(2) == 3, "Check failed: weight.size(2) == k");
  TORCH_CHECK(index.scalar_type() == at::kLong, "index must be int64");
  TORCH_CHECK(weight.scalar_type() == x.scalar_type(), "weight must have the dtype of the features");
}

// interpolate.cpp:8-22
at::Tensor interpolate_forward(const at::Tensor input, const at::Tensor index, const at::Tensor weight) {
  check_interp(input, index, weight, index.size(1));
  const int dt = dtype_of(input, "input");
  c10::cuda::CUDAGuard guard(input.device());
  const auto idx = index.contiguous();
  const auto w = weight.contiguous();
  const auto B = input.size(0), C = input.size(1), M = input.size(2), N = idx.size(1);
  auto out = at::empty({B, C, N}, input.options());
  check_rc(mvp_interpolate_forward(input.data_ptr(), input.stride(0), input.stride(1), input.stride(2),
                                   idx.data_ptr<int64_t>(), w.data_ptr(), B, C, M, N, dt, out.data_ptr(), cur_stream()));
  return out;
}

at::Tensor interpolate_backward(const at::Tensor grad_output, const at::Tensor index, const at::Tensor weight,
                                const int64_t num_inst) {
  check_interp(grad_output, index, weight, grad_output.size(2));
  const int dt = dtype_of(grad_output, "grad_output");
  c10::cuda::CUDAGuard guard(grad_output.device());
  const auto g = grad_output.contiguous();
  const auto idx = index.contiguous();
  const auto w = weight.contiguous();
  const auto B = g.size(0), C = g.size(1), N = g.size(2);
  auto grad_input = at::empty({B, C, num_inst}, g.options());
  check_rc(mvp_interpolate_backward(g.data_ptr(), idx.data_ptr<int64_t>(), w.data_ptr(), B, C, num_inst, N, dt,
                                    grad_input.data_ptr(), cur_stream()));
  return grad_input;
}

#define CHECK_F32(x) TORCH_CHECK((x).scalar_type() == at::kFloat, #x " must be float32")
// ---- stored input formats -> network inputs on the device (csrc/unproject.cu)
at::Tensor decode_rgb_u8(const at::Tensor rgb, const std::vector<double> mean, const std::vector<double> stddev) {
  CHECK_INPUT(rgb);
  TORCH_CHECK(rgb.scalar_type() == at::kByte && rgb.dim() >= 3 && rgb.size(-1) == 3, "decode_rgb_u8: uint8 (..., H, W, 3)");
  TORCH_CHECK(mean.size() == 3 && stddev.size() == 3, "decode_rgb_u8: three means and three standard deviations");
  const float m[3] = {(float)mean[0], (float)mean[1], (float)mean[2]}, sd[3] = {(float)stddev[0], (float)stddev[1], (float)stddev[2]};
  const auto H = rgb.size(-3), W = rgb.size(-2), N = rgb.numel() / (3 * H * W);
  c10::cuda::CUDAGuard guard(rgb.device());
  auto sizes = rgb.sizes().vec();
  sizes[sizes.size() - 3] = 3; sizes[sizes.size() - 2] = H; sizes[sizes.size() - 1] = W;
  auto out = at::empty(sizes, rgb.options().dtype(at::kFloat));
  check_rc(mvp_decode_rgb_u8(rgb.data_ptr<uint8_t>(), N, H, W, m, sd, out.data_ptr<float>(), cur_stream()));
  return out;
}

// depth_mm: int16 tensor holding the uint16 bit patterns of the depth PNGs (torch has no first-class uint16)
at::Tensor decode_depth_u16(const at::Tensor depth_mm) {
  CHECK_INPUT(depth_mm);
  TORCH_CHECK(depth_mm.scalar_type() == at::kShort || depth_mm.scalar_type() == at::kUInt16, "decode_depth_u16: int16 / uint16 bit patterns");
  c10::cuda::CUDAGuard guard(depth_mm.device());
  auto out = at::empty(depth_mm.sizes(), depth_mm.options().dtype(at::kFloat));
  check_rc(mvp_decode_depth_u16((const uint16_t *)depth_mm.data_ptr(), depth_mm.numel(), out.data_ptr<float>(), cur_stream()));
  return out;
}

// ---- deterministic backward variants (csrc/train_ops.cu): same signatures as the reference's backward functions
at::Tensor group_points_backward_det(const at::Tensor grad_output, const at::Tensor index, const int64_t num_points) {
  CHECK_CUDA(grad_output);
  CHECK_CUDA(index);
  TORCH_CHECK(grad_output.dim() == 4 && index.dim() == 3, "group_points_backward: grad_output (B,C,N2,K), index (B,N2,K)");
  TORCH_CHECK(index.size(0) == grad_output.size(0) && index.size(1) == grad_output.size(2) && index.size(2) == grad_output.size(3),
              "group_points_backward: shape mismatch");
  TORCH_CHECK(index.scalar_type() == at::kLong, "index must be int64");
  TORCH_CHECK(grad_output.scalar_type() == at::kFloat, "deterministic backward is float32 only");
  c10::cuda::CUDAGuard guard(grad_output.device());
  const auto g = grad_output.contiguous();
  const auto idx = index.contiguous();
  const auto B = g.size(0), C = g.size(1), N2 = g.size(2), K = g.size(3);
  auto grad_input = at::empty({B, C, num_points}, g.options());
  auto ws = at::empty({mvp_scatter_det_workspace_bytes(B, C, num_points, N2, K, 0)}, g.options().dtype(at::kByte));
  check_rc(mvp_group_points_backward_det(g.data_ptr<float>(), idx.data_ptr<int64_t>(), B, C, num_points, N2, K, grad_input.data_ptr<float>(),
                                         ws.data_ptr(), cur_stream()));
  return grad_input;
}

at::Tensor interpolate_backward_det(const at::Tensor grad_output, const at::Tensor index, const at::Tensor weight, const int64_t num_inst) {
  check_interp(grad_output, index, weight, grad_output.size(2));
  TORCH_CHECK(grad_output.scalar_type() == at::kFloat, "deterministic backward is float32 only");
  c10::cuda::CUDAGuard guard(grad_output.device());
  const auto g = grad_output.contiguous();
  const auto idx = index.contiguous();
  const auto w = weight.contiguous();
  const auto B = g.size(0), C = g.size(1), N = g.size(2);
  auto grad_input = at::empty({B, C, num_inst}, g.options());
  auto ws = at::empty({mvp_scatter_det_workspace_bytes(B, C, num_inst, N, 3, 1)}, g.options().dtype(at::kByte));
  check_rc(mvp_interpolate_backward_det(g.data_ptr<float>(), idx.data_ptr<int64_t>(), w.data_ptr<float>(), B, C, num_inst, N,
                                        grad_input.data_ptr<float>(), ws.data_ptr(), cur_stream()));
  return grad_input;
}

// ---- SegLoss / SegAccuracy / SegIoU statistics (csrc/train_ops.cu; mvpnet/models/loss.py, metric.py)
// returns (loss_out float[2] = {loss, sum of weights}, lse float (B, N), conf int64 (C, C) = confusion matrix of this call)
std::vector<at::Tensor> seg_loss_forward(const at::Tensor logit, const at::Tensor label, const c10::optional<at::Tensor> weight,
                                         int64_t ignore_index) {
  CHECK_INPUT(logit); CHECK_INPUT(label); CHECK_F32(logit);
  TORCH_CHECK(logit.dim() == 3 && label.dim() == 2 && label.size(0) == logit.size(0) && label.size(1) == logit.size(2),
              "seg_loss: logit (B, C, N), label (B, N)");
  TORCH_CHECK(label.scalar_type() == at::kLong, "seg_loss: label must be int64");
  const float *wp = nullptr;
  if (weight.has_value() && weight->defined()) {
    CHECK_INPUT((*weight)); CHECK_F32((*weight));
    TORCH_CHECK(weight->numel() == logit.size(1), "seg_loss: weight must have one entry per class");
    wp = weight->data_ptr<float>();
  }
  c10::cuda::CUDAGuard guard(logit.device());
  const auto B = logit.size(0), C = logit.size(1), N = logit.size(2);
  auto out = at::empty({2}, logit.options());
  auto lse = at::empty({B, N}, logit.options());
  auto conf = at::zeros({C, C}, logit.options().dtype(at::kLong));
  auto ws = at::empty({mvp_seg_loss_workspace_bytes(B, N)}, logit.options().dtype(at::kByte));
  check_rc(mvp_seg_loss_forward(logit.data_ptr<float>(), label.data_ptr<int64_t>(), wp, B, C, N, ignore_index, lse.data_ptr<float>(),
                                out.data_ptr<float>(), (uint64_t *)conf.data_ptr<int64_t>(), ws.data_ptr(), cur_stream()));
  return {out, lse, conf};
}

at::Tensor seg_loss_backward(const at::Tensor logit, const at::Tensor label, const c10::optional<at::Tensor> weight, const at::Tensor lse,
                             const at::Tensor loss_out, const at::Tensor grad_scale, int64_t ignore_index) {
  CHECK_INPUT(logit); CHECK_INPUT(label); CHECK_INPUT(lse); CHECK_INPUT(loss_out); CHECK_INPUT(grad_scale);
  CHECK_F32(logit); CHECK_F32(lse); CHECK_F32(loss_out); CHECK_F32(grad_scale);
  const float *wp = nullptr;
  if (weight.has_value() && weight->defined()) wp = weight->data_ptr<float>();
  c10::cuda::CUDAGuard guard(logit.device());
  auto grad = at::empty_like(logit);
  check_rc(mvp_seg_loss_backward(logit.data_ptr<float>(), label.data_ptr<int64_t>(), wp, lse.data_ptr<float>(), loss_out.data_ptr<float>(),
                                 grad_scale.data_ptr<float>(), logit.size(0), logit.size(1), logit.size(2), ignore_index,
                                 grad.data_ptr<float>(), cur_stream()));
  return grad;
}

// adds the confusion matrix of (argmax(logit), label) into conf (int64 (C, C)) in place
void seg_confusion(const at::Tensor logit, const at::Tensor label, int64_t ignore_index, at::Tensor conf) {
  CHECK_INPUT(logit); CHECK_INPUT(label); CHECK_INPUT(conf); CHECK_F32(logit);
  TORCH_CHECK(logit.dim() == 3 && label.dim() == 2 && label.size(0) == logit.size(0) && label.size(1) == logit.size(2),
              "seg_confusion: logit (B, C, N), label (B, N)");
  TORCH_CHECK(label.scalar_type() == at::kLong && conf.scalar_type() == at::kLong && conf.numel() == logit.size(1) * logit.size(1),
              "seg_confusion: label int64, conf int64 (C, C)");
  c10::cuda::CUDAGuard guard(logit.device());
  check_rc(mvp_seg_confusion(logit.data_ptr<float>(), label.data_ptr<int64_t>(), logit.size(0), logit.size(1), logit.size(2), ignore_index,
                             (uint64_t *)conf.data_ptr<int64_t>(), cur_stream()));
}

// ---- data side of FeatureAggregation (no reference extension; scannet_2d3d.py:33-39,255-313) ----
std::vector<at::Tensor> unproject(const at::Tensor depth, const at::Tensor cam_inv, const at::Tensor pose,
                                  const c10::optional<at::Tensor> chunk_box, bool want_xyz64) {
  CHECK_INPUT(depth);
  CHECK_INPUT(cam_inv);
  CHECK_INPUT(pose);
  TORCH_CHECK(depth.dim() == 4, "depth must be (B, nv, h, w)");
  const auto B = depth.size(0), nv = depth.size(1), h = depth.size(2), w = depth.size(3);
  TORCH_CHECK(depth.scalar_type() == at::kFloat && cam_inv.scalar_type() == at::kFloat && pose.scalar_type() == at::kFloat,
              "depth, cam_inv and pose must be float32");
  TORCH_CHECK(cam_inv.numel() == B * nv * 9, "cam_inv must be (B, nv, 3, 3)");
  TORCH_CHECK(pose.numel() == B * nv * 16, "pose must be (B, nv, 4, 4)");
  const double *box = nullptr;
  at::Tensor boxc;
  if (chunk_box.has_value() && chunk_box->defined()) {
    boxc = chunk_box->contiguous();
    CHECK_CUDA(boxc);
    TORCH_CHECK(boxc.scalar_type() == at::kDouble && boxc.numel() == B * 4, "chunk_box must be float64 (B, 4)");
    box = boxc.data_ptr<double>();
  }
  c10::cuda::CUDAGuard guard(depth.device());
  auto xyz32 = at::empty({B, nv, h, w, 3}, depth.options());
  auto mask = at::empty({B, nv, h, w}, depth.options().dtype(at::kByte));
  at::Tensor xyz64;
  if (want_xyz64) xyz64 = at::empty({B, nv * h * w, 3}, depth.options().dtype(at::kDouble));
  check_rc(mvp_unproject(depth.data_ptr<float>(), cam_inv.data_ptr<float>(), pose.data_ptr<float>(), box, B, nv, h, w,
                         want_xyz64 ? xyz64.data_ptr<double>() : nullptr, xyz32.data_ptr<float>(),
                         mask.data_ptr<uint8_t>(), cur_stream()));
  return {xyz32, mask, want_xyz64 ? xyz64 : at::Tensor()};
}

std::vector<at::Tensor> knn_pixels(const at::Tensor query, const at::Tensor pix_xyz, const at::Tensor mask, int64_t k, bool exhaustive) {
  CHECK_INPUT(query);
  CHECK_INPUT(pix_xyz);
  CHECK_INPUT(mask);
  TORCH_CHECK(query.dim() == 3 && query.size(2) == 3, "query must be (B, nq, 3)");
  TORCH_CHECK(pix_xyz.dim() == 3 && pix_xyz.size(2) == 3, "pix_xyz must be (B, P, 3)");
  TORCH_CHECK(query.scalar_type() == at::kDouble && pix_xyz.scalar_type() == at::kDouble, "query/pix_xyz must be float64");
  TORCH_CHECK(mask.scalar_type() == at::kByte || mask.scalar_type() == at::kBool, "mask must be uint8/bool");
  TORCH_CHECK(mask.numel() == pix_xyz.size(0) * pix_xyz.size(1), "mask must be (B, P)");
  TORCH_CHECK(query.size(0) == pix_xyz.size(0), "batch mismatch");
  c10::cuda::CUDAGuard guard(query.device());
  auto index = at::empty({query.size(0), query.size(1), k}, query.options().dtype(at::kLong));
  auto dist2 = at::empty({query.size(0), query.size(1), k}, query.options());
  at::Tensor work;
  void *wp = nullptr;
  if (!exhaustive) {
    const int64_t ws = mvp_knn_pixels_workspace_bytes(query.size(0), query.size(1), pix_xyz.size(1), k);
    if (ws > 0) {
      work = at::empty({ws}, query.options().dtype(at::kByte));
      wp = work.data_ptr();
    }
  }
  check_rc(mvp_knn_pixels(query.data_ptr<double>(), pix_xyz.data_ptr<double>(), (const uint8_t *)mask.data_ptr(),
                          query.size(0), query.size(1), pix_xyz.size(1), k, index.data_ptr<int64_t>(),
                          dist2.data_ptr<double>(), wp, cur_stream()));
  return {index, dist2};
}


// ---- fused inference kernels (point-major layout; see include/mvpnet_b200.h) ----------------------
struct Chain {
  mvp_mlp_chain_t c;
  Chain(const std::vector<at::Tensor> &wts, const std::vector<at::Tensor> &biases, const std::vector<int64_t> &relu,
        int64_t out_channels, const at::Device &dev) {
    TORCH_CHECK(!wts.empty() && wts.size() <= MVP_MLP_MAX_LAYERS, "fused mlp: 1..", MVP_MLP_MAX_LAYERS, " layers");
    TORCH_CHECK(wts.size() == biases.size() && wts.size() == relu.size(), "fused mlp: list lengths differ");
    c.num_layers = (int32_t)wts.size();
    c.out_channels = (int32_t)out_channels;
    for (size_t l = 0; l < wts.size(); ++l) {
      const auto &w = wts[l];
      const auto &b = biases[l];
      TORCH_CHECK(w.is_cuda() && b.is_cuda() && w.device() == dev && b.device() == dev, "fused mlp: weights on wrong device");
      TORCH_CHECK(w.is_contiguous() && b.is_contiguous() && w.dim() == 2 && b.dim() == 1, "fused mlp: weights must be contiguous [cin][cout] / [cout]");
      TORCH_CHECK(w.scalar_type() == at::kFloat && b.scalar_type() == at::kFloat, "fused mlp: weights must be float32");
      TORCH_CHECK(b.size(0) == w.size(1), "fused mlp: bias/weight mismatch");
      c.cin[l] = (int32_t)w.size(0);
      c.cout[l] = (int32_t)w.size(1);
      c.relu[l] = (int32_t)relu[l];
      c.wt[l] = w.data_ptr<float>();
      c.bias[l] = b.data_ptr<float>();
    }
  }
};


at::Tensor fused_set_abstraction(const c10::optional<at::Tensor> feat, const at::Tensor xyz, const at::Tensor new_xyz,
                                 const at::Tensor nbr, const std::vector<at::Tensor> wts, const std::vector<at::Tensor> biases,
                                 const std::vector<int64_t> relu, int64_t out_channels) {
  CHECK_INPUT(xyz); CHECK_INPUT(new_xyz); CHECK_INPUT(nbr);
  CHECK_F32(xyz); CHECK_F32(new_xyz);
  TORCH_CHECK(xyz.dim() == 3 && xyz.size(2) == 3 && new_xyz.dim() == 3 && new_xyz.size(2) == 3, "xyz/new_xyz must be (B, N, 3)");
  TORCH_CHECK(nbr.scalar_type() == at::kLong && nbr.dim() == 3, "nbr must be int64 (B, M, K)");
  const auto B = xyz.size(0), N = xyz.size(1), M = new_xyz.size(1), K = nbr.size(2);
  TORCH_CHECK(new_xyz.size(0) == B && nbr.size(0) == B && nbr.size(1) == M, "fused_set_abstraction: shape mismatch");
  int64_t C = 0;
  const float *fp = nullptr;
  if (feat.has_value() && feat->defined()) {
    CHECK_INPUT((*feat)); CHECK_F32((*feat));
    TORCH_CHECK(feat->dim() == 3 && feat->size(0) == B && feat->size(1) == N, "feat must be (B, N, C)");
    C = feat->size(2);
    fp = feat->data_ptr<float>();
  }
  c10::cuda::CUDAGuard guard(xyz.device());
  Chain ch(wts, biases, relu, out_channels, xyz.device());
  auto out = at::empty({B, M, out_channels}, xyz.options());
  check_rc(mvp_fused_set_abstraction(fp, C, xyz.data_ptr<float>(), new_xyz.data_ptr<float>(), nbr.data_ptr<int64_t>(), B, N, M,
                                     K, &ch.c, out.data_ptr<float>(), cur_stream()));
  return out;
}

at::Tensor fused_feature_aggregation(const at::Tensor feat2d, const at::Tensor pix_xyz, const at::Tensor points,
                                     const at::Tensor knn, bool reduce_sum, const std::vector<at::Tensor> wts,
                                     const std::vector<at::Tensor> biases, const std::vector<int64_t> relu, int64_t out_channels) {
  CHECK_CUDA(feat2d); CHECK_F32(feat2d);
  CHECK_INPUT(pix_xyz); CHECK_INPUT(points); CHECK_INPUT(knn);
  CHECK_F32(pix_xyz); CHECK_F32(points);
  TORCH_CHECK(feat2d.dim() == 5, "feat2d must be (B, nv, C, h, w) (any strides with a uniform view stride)");
  const auto B = feat2d.size(0), nv = feat2d.size(1), C = feat2d.size(2), h = feat2d.size(3), w = feat2d.size(4);
  TORCH_CHECK(feat2d.stride(0) == nv * feat2d.stride(1), "feat2d: batch and view axes must be collapsible");
  TORCH_CHECK(pix_xyz.dim() == 3 && pix_xyz.size(0) == B && pix_xyz.size(1) == nv * h * w && pix_xyz.size(2) == 3,
              "pix_xyz must be (B, nv*h*w, 3)");
  TORCH_CHECK(points.dim() == 3 && points.size(0) == B && points.size(2) == 3, "points must be (B, Np, 3)");
  TORCH_CHECK(knn.scalar_type() == at::kLong && knn.dim() == 3 && knn.size(0) == B && knn.size(1) == points.size(1),
              "knn must be int64 (B, Np, K)");
  c10::cuda::CUDAGuard guard(feat2d.device());
  Chain ch(wts, biases, relu, out_channels, feat2d.device());
  auto out = at::empty({B, points.size(1), out_channels}, points.options());
  check_rc(mvp_fused_feature_aggregation(feat2d.data_ptr<float>(), feat2d.stride(1), feat2d.stride(2), feat2d.stride(3),
                                         feat2d.stride(4), C, nv, h, w, pix_xyz.data_ptr<float>(), points.data_ptr<float>(),
                                         knn.data_ptr<int64_t>(), B, points.size(1), knn.size(2), reduce_sum ? 1 : 0, &ch.c,
                                         out.data_ptr<float>(), cur_stream()));
  return out;
}

at::Tensor fused_feature_propagation(const at::Tensor sparse_feat, const at::Tensor idx, const at::Tensor dist2,
                                     const c10::optional<at::Tensor> skip, double eps, const std::vector<at::Tensor> wts,
                                     const std::vector<at::Tensor> biases, const std::vector<int64_t> relu, int64_t out_channels) {
  CHECK_INPUT(sparse_feat); CHECK_INPUT(idx); CHECK_INPUT(dist2);
  CHECK_F32(sparse_feat); CHECK_F32(dist2);
  TORCH_CHECK(sparse_feat.dim() == 3 && idx.dim() == 3 && dist2.dim() == 3 && idx.size(2) == 3 && dist2.size(2) == 3,
              "fused_feature_propagation: sparse_feat (B,Ns,Cs), idx/dist2 (B,Nd,3)");
  TORCH_CHECK(idx.scalar_type() == at::kLong, "idx must be int64");
  const auto B = sparse_feat.size(0), Ns = sparse_feat.size(1), Cs = sparse_feat.size(2), Nd = idx.size(1);
  TORCH_CHECK(idx.size(0) == B && dist2.size(0) == B && dist2.size(1) == Nd, "fused_feature_propagation: shape mismatch");
  int64_t Cd = 0;
  const float *sp = nullptr;
  if (skip.has_value() && skip->defined()) {
    CHECK_INPUT((*skip)); CHECK_F32((*skip));
    TORCH_CHECK(skip->dim() == 3 && skip->size(0) == B && skip->size(1) == Nd, "skip must be (B, Nd, Cd)");
    Cd = skip->size(2);
    sp = skip->data_ptr<float>();
  }
  c10::cuda::CUDAGuard guard(sparse_feat.device());
  Chain ch(wts, biases, relu, out_channels, sparse_feat.device());
  auto out = at::empty({B, Nd, out_channels}, sparse_feat.options());
  check_rc(mvp_fused_feature_propagation(sparse_feat.data_ptr<float>(), Cs, idx.data_ptr<int64_t>(), dist2.data_ptr<float>(), sp,
                                         Cd, B, Ns, Nd, (float)eps, &ch.c, out.data_ptr<float>(), cur_stream()));
  return out;
}

// ---- tensor-core (tcgen05) variants ------------------------------------------------------------------
struct TcChain {
  mvp_tc_chain_t c;
  TcChain(const std::vector<at::Tensor> &w_hi, const std::vector<at::Tensor> &w_lo, const std::vector<at::Tensor> &biases,
          const std::vector<int64_t> &ks, const std::vector<int64_t> &ns, const std::vector<int64_t> &relu,
          int64_t out_channels, const at::Device &dev) {
    const size_t L = w_hi.size();
    TORCH_CHECK(L >= 1 && L <= MVP_MLP_MAX_LAYERS, "tc mlp: 1..", MVP_MLP_MAX_LAYERS, " layers");
    TORCH_CHECK(w_lo.size() == L && biases.size() == L && ks.size() == L && ns.size() == L && relu.size() == L,
                "tc mlp: list lengths differ");
    c.num_layers = (int32_t)L;
    c.out_channels = (int32_t)out_channels;
    for (size_t l = 0; l < L; ++l) {
      TORCH_CHECK(w_hi[l].is_cuda() && w_lo[l].is_cuda() && biases[l].is_cuda() && w_hi[l].device() == dev,
                  "tc mlp: weights on wrong device");
      TORCH_CHECK(w_hi[l].scalar_type() == at::kBFloat16 && w_lo[l].scalar_type() == at::kBFloat16 &&
                      biases[l].scalar_type() == at::kFloat,
                  "tc mlp: w_hi/w_lo must be bfloat16, bias float32");
      TORCH_CHECK(w_hi[l].is_contiguous() && w_lo[l].is_contiguous() && biases[l].is_contiguous(), "tc mlp: contiguous weights");
      TORCH_CHECK(w_hi[l].numel() == ks[l] * ns[l] && w_lo[l].numel() == ks[l] * ns[l] && biases[l].numel() == ns[l],
                  "tc mlp: weight sizes do not match k x n");
      c.k[l] = (int32_t)ks[l];
      c.n[l] = (int32_t)ns[l];
      c.relu[l] = (int32_t)relu[l];
      c.w_hi[l] = w_hi[l].data_ptr();
      c.w_lo[l] = w_lo[l].data_ptr();
      c.bias[l] = biases[l].data_ptr<float>();
    }
  }
};

// ---- 2D network on the tensor cores (csrc/tc_conv.cu).  Split-planar activations are flat bf16 tensors of
// mvp_planar_elems(N, H, W, C) elements; their logical shape travels beside them.
at::Tensor split_planar(const at::Tensor nhwc) {
  CHECK_INPUT(nhwc); CHECK_F32(nhwc);
  TORCH_CHECK(nhwc.dim() == 4, "split_planar: input must be fp32 (N, H, W, C)");
  const auto N = nhwc.size(0), H = nhwc.size(1), W = nhwc.size(2), C = nhwc.size(3);
  c10::cuda::CUDAGuard guard(nhwc.device());
  auto out = at::empty({mvp_planar_elems(N, H, W, C)}, nhwc.options().dtype(at::kBFloat16));
  check_rc(mvp_split_planar(nhwc.data_ptr<float>(), N, H, W, C, out.data_ptr(), cur_stream()));
  return out;
}

at::Tensor merge_planar(const at::Tensor planar, int64_t N, int64_t H, int64_t W, int64_t C) {
  CHECK_INPUT(planar);
  TORCH_CHECK(planar.scalar_type() == at::kBFloat16 && planar.numel() == mvp_planar_elems(N, H, W, C), "merge_planar: wrong planar size");
  c10::cuda::CUDAGuard guard(planar.device());
  auto out = at::empty({N, H, W, C}, planar.options().dtype(at::kFloat));
  check_rc(mvp_merge_planar(planar.data_ptr(), N, H, W, C, out.data_ptr<float>(), cur_stream()));
  return out;
}

// x1 (C1 channels) [+ x2 (C2 channels)] split-planar, packed weights, bias (Cout), optional split-planar residual
// -> split-planar (N,H,W,Cout), or fp32 NHWC when nhwc_out
at::Tensor tc_conv3x3(const at::Tensor x1, int64_t C1, const c10::optional<at::Tensor> x2, int64_t C2, int64_t N, int64_t H, int64_t W,
                      const at::Tensor w_packed, const at::Tensor bias, const c10::optional<at::Tensor> residual, bool relu, int64_t out_mode, bool pair) {
  // out_mode 0: split-planar; 1: fp32 NHWC; 2: row-split (2, N, H, W, Cout) bf16 (hi plane, lo plane)
  CHECK_INPUT(x1); CHECK_INPUT(w_packed); CHECK_INPUT(bias); CHECK_F32(bias);
  const auto Cout = bias.size(0);
  TORCH_CHECK(x1.scalar_type() == at::kBFloat16 && x1.numel() == mvp_planar_elems(N, H, W, C1), "tc_conv3x3: x1 is not split-planar (N,H,W,C1)");
  const void *p2 = nullptr, *pr = nullptr;
  if (x2.has_value() && x2->defined()) {
    CHECK_INPUT((*x2));
    TORCH_CHECK(x2->scalar_type() == at::kBFloat16 && x2->numel() == mvp_planar_elems(N, H, W, C2), "tc_conv3x3: x2 is not split-planar (N,H,W,C2)");
    p2 = x2->data_ptr();
  } else {
    TORCH_CHECK(C2 == 0, "tc_conv3x3: C2 given without x2");
  }
  if (residual.has_value() && residual->defined()) {
    CHECK_INPUT((*residual));
    TORCH_CHECK(residual->scalar_type() == at::kBFloat16 && residual->numel() == mvp_planar_elems(N, H, W, Cout),
                "tc_conv3x3: residual is not split-planar (N,H,W,Cout)");
    pr = residual->data_ptr();
  }
  TORCH_CHECK(w_packed.numel() * w_packed.element_size() == mvp_tc_conv3x3_weight_bytes(C1 + C2, Cout),
              "tc_conv3x3: packed weights have the wrong size for ", C1 + C2, " -> ", Cout, " channels");
  c10::cuda::CUDAGuard guard(x1.device());
  at::Tensor out;
  // pair: weights packed with half the block width, CTA-pair kernel (mvp_tc_conv3x3_pair)
  auto conv = pair ? mvp_tc_conv3x3_pair : mvp_tc_conv3x3;
  if (out_mode == 1) {
    out = at::empty({N, H, W, Cout}, x1.options().dtype(at::kFloat));
    check_rc(conv(x1.data_ptr(), C1, p2, C2, N, H, W, w_packed.data_ptr(), bias.data_ptr<float>(), Cout, pr, relu ? 1 : 0,
                            nullptr, out.data_ptr<float>(), nullptr, cur_stream()));
  } else if (out_mode == 2) {
    out = at::empty({2, N, H, W, Cout}, x1.options());
    check_rc(conv(x1.data_ptr(), C1, p2, C2, N, H, W, w_packed.data_ptr(), bias.data_ptr<float>(), Cout, pr, relu ? 1 : 0,
                            nullptr, nullptr, out.data_ptr(), cur_stream()));
  } else {
    // an odd image count leaves a partner-less image in the last pair of a pair-interleaved tensor: keep it zero
    out = (H <= 8 && (N & 1)) ? at::zeros({mvp_planar_elems(N, H, W, Cout)}, x1.options()) : at::empty({mvp_planar_elems(N, H, W, Cout)}, x1.options());
    check_rc(conv(x1.data_ptr(), C1, p2, C2, N, H, W, w_packed.data_ptr(), bias.data_ptr<float>(), Cout, pr, relu ? 1 : 0,
                            out.data_ptr(), nullptr, nullptr, cur_stream()));
  }
  return out;
}

at::Tensor planar_empty(const at::Tensor &like, int64_t N, int64_t H, int64_t W, int64_t C) {
  // an odd image count leaves a partner-less image in the last pair of a pair-interleaved tensor: keep it zero
  auto opt = like.options().dtype(at::kBFloat16);
  return (H <= 8 && (N & 1)) ? at::zeros({mvp_planar_elems(N, H, W, C)}, opt) : at::empty({mvp_planar_elems(N, H, W, C)}, opt);
}

// mode 0: taps (dy, dx) convolution with stride; mode 1: 2x2 / stride-2 transposed convolution
at::Tensor tc_conv_general(const at::Tensor x, int64_t Cin, int64_t N, int64_t Hi, int64_t Wi, int64_t mode, int64_t stride,
                           const std::vector<int64_t> dy, const std::vector<int64_t> dx, int64_t Ho, int64_t Wo, const at::Tensor w_packed,
                           const at::Tensor bias, bool relu) {
  CHECK_INPUT(x); CHECK_INPUT(w_packed); CHECK_INPUT(bias); CHECK_F32(bias);
  TORCH_CHECK(x.scalar_type() == at::kBFloat16 && x.numel() == mvp_planar_elems(N, Hi, Wi, Cin), "tc_conv_general: x is not split-planar (N,Hi,Wi,Cin)");
  TORCH_CHECK(dy.size() == dx.size() && !dy.empty() && dy.size() <= 9, "tc_conv_general: 1..9 taps");
  const auto Cout = bias.size(0);
  const int64_t G = mode ? 4 * Cout : Cout;
  TORCH_CHECK(w_packed.numel() * w_packed.element_size() == (int64_t)dy.size() * Cin * G * 4, "tc_conv_general: packed weights have the wrong size");
  std::vector<int> idy(dy.begin(), dy.end()), idx(dx.begin(), dx.end());
  c10::cuda::CUDAGuard guard(x.device());
  auto out = planar_empty(x, N, Ho, Wo, Cout);
  check_rc(mvp_tc_conv_general(x.data_ptr(), Cin, N, Hi, Wi, (int)mode, (int)stride, (int)dy.size(), idy.data(), idx.data(), Ho, Wo,
                               w_packed.data_ptr(), bias.data_ptr<float>(), Cout, relu ? 1 : 0, out.data_ptr(), cur_stream()));
  return out;
}

at::Tensor unfold_stem(const at::Tensor image) {
  CHECK_INPUT(image); CHECK_F32(image);
  TORCH_CHECK(image.dim() == 4 && image.size(1) == 3, "unfold_stem: image must be fp32 (N, 3, H, W)");
  c10::cuda::CUDAGuard guard(image.device());
  auto out = at::empty({mvp_planar_elems(image.size(0), image.size(2), image.size(3), 32)}, image.options().dtype(at::kBFloat16));
  check_rc(mvp_unfold_stem(image.data_ptr<float>(), image.size(0), image.size(2), image.size(3), out.data_ptr(), cur_stream()));
  return out;
}

at::Tensor maxpool3x3s2_planar(const at::Tensor x, int64_t N, int64_t H, int64_t W, int64_t C) {
  CHECK_INPUT(x);
  TORCH_CHECK(x.scalar_type() == at::kBFloat16 && x.numel() == mvp_planar_elems(N, H, W, C), "maxpool_planar: x is not split-planar (N,H,W,C)");
  c10::cuda::CUDAGuard guard(x.device());
  auto out = at::empty({mvp_planar_elems(N, (H - 1) / 2 + 1, (W - 1) / 2 + 1, C)}, x.options());
  check_rc(mvp_maxpool3x3s2_planar(x.data_ptr(), N, H, W, C, out.data_ptr(), cur_stream()));
  return out;
}

bool tc_chain_supported(const std::vector<int64_t> ks, const std::vector<int64_t> ns, int64_t mode) {
  mvp_tc_chain_t c = {};
  if (ks.empty() || ks.size() > MVP_MLP_MAX_LAYERS || ks.size() != ns.size()) return false;
  c.num_layers = (int32_t)ks.size();
  for (size_t l = 0; l < ks.size(); ++l) { c.k[l] = (int32_t)ks[l]; c.n[l] = (int32_t)ns[l]; }
  return mvp_tc_chain_supported(&c, (int)mode) != 0;
}

at::Tensor tc_set_abstraction(const c10::optional<at::Tensor> feat, const at::Tensor xyz, const at::Tensor new_xyz,
                              const at::Tensor nbr, const std::vector<at::Tensor> w_hi, const std::vector<at::Tensor> w_lo,
                              const std::vector<at::Tensor> biases, const std::vector<int64_t> ks, const std::vector<int64_t> ns,
                              const std::vector<int64_t> relu, int64_t out_channels) {
  CHECK_INPUT(xyz); CHECK_INPUT(new_xyz); CHECK_INPUT(nbr);
  CHECK_F32(xyz); CHECK_F32(new_xyz);
  TORCH_CHECK(xyz.dim() == 3 && xyz.size(2) == 3 && new_xyz.dim() == 3 && new_xyz.size(2) == 3, "xyz/new_xyz must be (B, N, 3)");
  TORCH_CHECK(nbr.scalar_type() == at::kLong && nbr.dim() == 3, "nbr must be int64 (B, M, K)");
  const auto B = xyz.size(0), N = xyz.size(1), M = new_xyz.size(1), K = nbr.size(2);
  TORCH_CHECK(new_xyz.size(0) == B && nbr.size(0) == B && nbr.size(1) == M, "tc_set_abstraction: shape mismatch");
  int64_t C = 0;
  const float *fp = nullptr;
  if (feat.has_value() && feat->defined()) {
    CHECK_INPUT((*feat)); CHECK_F32((*feat));
    TORCH_CHECK(feat->dim() == 3 && feat->size(0) == B && feat->size(1) == N, "feat must be (B, N, C)");
    C = feat->size(2);
    fp = feat->data_ptr<float>();
  }
  c10::cuda::CUDAGuard guard(xyz.device());
  TcChain ch(w_hi, w_lo, biases, ks, ns, relu, out_channels, xyz.device());
  auto out = at::empty({B, M, out_channels}, xyz.options());
  check_rc(mvp_tc_fused_set_abstraction(fp, C, xyz.data_ptr<float>(), new_xyz.data_ptr<float>(), nbr.data_ptr<int64_t>(), B, N, M,
                                        K, &ch.c, out.data_ptr<float>(), cur_stream()));
  return out;
}

// ---- second-generation kernels (csrc/tc2_mlp.cu): features travel pre-split as one bf16 tensor (2, ..., C) = (hi, lo)
bool tc2_supported(const std::vector<int64_t> ks, const std::vector<int64_t> ns, int64_t mode, int64_t C) {
  mvp_tc_chain_t c = {};
  if (ks.empty() || ks.size() > MVP_MLP_MAX_LAYERS || ks.size() != ns.size()) return false;
  c.num_layers = (int32_t)ks.size();
  for (size_t l = 0; l < ks.size(); ++l) { c.k[l] = (int32_t)ks[l]; c.n[l] = (int32_t)ns[l]; }
  c.out_channels = c.n[c.num_layers - 1];
  return mvp_tc2_supported(&c, (int)mode, C) != 0;
}

static void check_split(const at::Tensor &t, const char *what) {
  TORCH_CHECK(t.is_cuda() && t.is_contiguous() && t.scalar_type() == at::kBFloat16 && t.dim() >= 3 && t.size(0) == 2, what,
              " must be a contiguous CUDA bfloat16 tensor (2, ..., C): hi plane, lo plane");
}

// returns (out_f32 (B, M, oc) or an empty tensor, out_split (2, B, M, oc) or an empty tensor)
std::tuple<at::Tensor, at::Tensor> tc2_set_abstraction(const at::Tensor feat_split, const at::Tensor xyz, const at::Tensor new_xyz,
                                                       const at::Tensor nbr, const std::vector<at::Tensor> w_hi, const std::vector<at::Tensor> w_lo,
                                                       const std::vector<at::Tensor> biases, const std::vector<int64_t> ks,
                                                       const std::vector<int64_t> ns, const std::vector<int64_t> relu, int64_t out_channels,
                                                       bool want_f32, bool want_split) {
  CHECK_INPUT(xyz); CHECK_INPUT(new_xyz); CHECK_INPUT(nbr);
  CHECK_F32(xyz); CHECK_F32(new_xyz);
  check_split(feat_split, "feat_split");
  TORCH_CHECK(xyz.dim() == 3 && xyz.size(2) == 3 && new_xyz.dim() == 3 && new_xyz.size(2) == 3, "xyz/new_xyz must be (B, N, 3)");
  TORCH_CHECK(nbr.scalar_type() == at::kLong && nbr.dim() == 3, "nbr must be int64 (B, M, K)");
  const auto B = xyz.size(0), N = xyz.size(1), M = new_xyz.size(1), K = nbr.size(2);
  TORCH_CHECK(new_xyz.size(0) == B && nbr.size(0) == B && nbr.size(1) == M, "tc2_set_abstraction: shape mismatch");
  TORCH_CHECK(feat_split.dim() == 4 && feat_split.size(1) == B && feat_split.size(2) == N, "feat_split must be (2, B, N, C)");
  TORCH_CHECK(want_f32 || want_split, "tc2_set_abstraction: no output requested");
  const auto C = feat_split.size(3);
  c10::cuda::CUDAGuard guard(xyz.device());
  TcChain ch(w_hi, w_lo, biases, ks, ns, relu, out_channels, xyz.device());
  at::Tensor out = want_f32 ? at::empty({B, M, out_channels}, xyz.options()) : at::empty({0}, xyz.options());
  at::Tensor sp = want_split ? at::empty({2, B, M, out_channels}, feat_split.options()) : at::empty({0}, feat_split.options());
  const auto *fh = (const at::BFloat16 *)feat_split.data_ptr();
  auto *oh = want_split ? (at::BFloat16 *)sp.data_ptr() : nullptr;
  check_rc(mvp_tc2_set_abstraction(fh, fh + B * N * C, C, xyz.data_ptr<float>(), new_xyz.data_ptr<float>(), nbr.data_ptr<int64_t>(), B, N, M, K,
                                   &ch.c, want_f32 ? out.data_ptr<float>() : nullptr, oh, oh ? oh + B * M * out_channels : nullptr, cur_stream()));
  return std::make_tuple(out, sp);
}

// pix_split (2, B*nv, hp, wp, C): the row-split output of the 2D network (padded image geometry)
std::tuple<at::Tensor, at::Tensor> tc2_feature_aggregation(const at::Tensor pix_split, int64_t nv, int64_t h, int64_t w, const at::Tensor pix_xyz,
                                                           const at::Tensor points, const at::Tensor knn, bool reduce_sum,
                                                           const std::vector<at::Tensor> w_hi, const std::vector<at::Tensor> w_lo,
                                                           const std::vector<at::Tensor> biases, const std::vector<int64_t> ks,
                                                           const std::vector<int64_t> ns, const std::vector<int64_t> relu, int64_t out_channels,
                                                           bool want_f32, bool want_split) {
  CHECK_INPUT(pix_xyz); CHECK_INPUT(points); CHECK_INPUT(knn);
  CHECK_F32(pix_xyz); CHECK_F32(points);
  check_split(pix_split, "pix_split");
  TORCH_CHECK(pix_split.dim() == 5, "pix_split must be (2, B*nv, hp, wp, C)");
  const auto hp = pix_split.size(2), wp = pix_split.size(3), C = pix_split.size(4);
  TORCH_CHECK(nv > 0 && pix_split.size(1) % nv == 0 && hp >= h && wp >= w, "tc2_feature_aggregation: image geometry mismatch");
  const auto B = pix_split.size(1) / nv;
  TORCH_CHECK(pix_xyz.dim() == 3 && pix_xyz.size(0) == B && pix_xyz.size(1) == nv * h * w && pix_xyz.size(2) == 3, "pix_xyz must be (B, nv*h*w, 3)");
  TORCH_CHECK(points.dim() == 3 && points.size(0) == B && points.size(2) == 3, "points must be (B, Np, 3)");
  TORCH_CHECK(knn.scalar_type() == at::kLong && knn.dim() == 3 && knn.size(0) == B && knn.size(1) == points.size(1), "knn must be int64 (B, Np, K)");
  TORCH_CHECK(want_f32 || want_split, "tc2_feature_aggregation: no output requested");
  const auto Np = points.size(1);
  c10::cuda::CUDAGuard guard(points.device());
  TcChain ch(w_hi, w_lo, biases, ks, ns, relu, out_channels, points.device());
  at::Tensor out = want_f32 ? at::empty({B, Np, out_channels}, points.options()) : at::empty({0}, points.options());
  at::Tensor sp = want_split ? at::empty({2, B, Np, out_channels}, pix_split.options()) : at::empty({0}, pix_split.options());
  const auto *ph = (const at::BFloat16 *)pix_split.data_ptr();
  auto *oh = want_split ? (at::BFloat16 *)sp.data_ptr() : nullptr;
  check_rc(mvp_tc2_feature_aggregation(ph, ph + pix_split.numel() / 2, C, nv, h, w, hp, wp, pix_xyz.data_ptr<float>(), points.data_ptr<float>(),
                                       knn.data_ptr<int64_t>(), B, Np, knn.size(2), reduce_sum ? 1 : 0, &ch.c,
                                       want_f32 ? out.data_ptr<float>() : nullptr, oh, oh ? oh + B * Np * out_channels : nullptr, cur_stream()));
  return std::make_tuple(out, sp);
}

at::Tensor tc_feature_aggregation(const at::Tensor feat2d, const at::Tensor pix_xyz, const at::Tensor points, const at::Tensor knn,
                                  bool reduce_sum, const std::vector<at::Tensor> w_hi, const std::vector<at::Tensor> w_lo,
                                  const std::vector<at::Tensor> biases, const std::vector<int64_t> ks, const std::vector<int64_t> ns,
                                  const std::vector<int64_t> relu, int64_t out_channels) {
  CHECK_CUDA(feat2d); CHECK_F32(feat2d);
  CHECK_INPUT(pix_xyz); CHECK_INPUT(points); CHECK_INPUT(knn);
  CHECK_F32(pix_xyz); CHECK_F32(points);
  TORCH_CHECK(feat2d.dim() == 5, "feat2d must be (B, nv, C, h, w)");
  const auto B = feat2d.size(0), nv = feat2d.size(1), C = feat2d.size(2), h = feat2d.size(3), w = feat2d.size(4);
  TORCH_CHECK(feat2d.stride(0) == nv * feat2d.stride(1), "feat2d: batch and view axes must be collapsible");
  TORCH_CHECK(pix_xyz.dim() == 3 && pix_xyz.size(0) == B && pix_xyz.size(1) == nv * h * w && pix_xyz.size(2) == 3,
              "pix_xyz must be (B, nv*h*w, 3)");
  TORCH_CHECK(points.dim() == 3 && points.size(0) == B && points.size(2) == 3, "points must be (B, Np, 3)");
  TORCH_CHECK(knn.scalar_type() == at::kLong && knn.dim() == 3 && knn.size(0) == B && knn.size(1) == points.size(1),
              "knn must be int64 (B, Np, K)");
  c10::cuda::CUDAGuard guard(feat2d.device());
  TcChain ch(w_hi, w_lo, biases, ks, ns, relu, out_channels, feat2d.device());
  auto out = at::empty({B, points.size(1), out_channels}, points.options());
  check_rc(mvp_tc_fused_feature_aggregation(feat2d.data_ptr<float>(), feat2d.stride(1), feat2d.stride(2), feat2d.stride(3),
                                            feat2d.stride(4), C, nv, h, w, pix_xyz.data_ptr<float>(), points.data_ptr<float>(),
                                            knn.data_ptr<int64_t>(), B, points.size(1), knn.size(2), reduce_sum ? 1 : 0, &ch.c,
                                            out.data_ptr<float>(), cur_stream()));
  return out;
}

at::Tensor tc_feature_propagation(const at::Tensor sparse_feat, const at::Tensor idx, const at::Tensor dist2,
                                  const c10::optional<at::Tensor> skip, double eps, const std::vector<at::Tensor> w_hi,
                                  const std::vector<at::Tensor> w_lo, const std::vector<at::Tensor> biases,
                                  const std::vector<int64_t> ks, const std::vector<int64_t> ns, const std::vector<int64_t> relu,
                                  int64_t out_channels) {
  CHECK_INPUT(sparse_feat); CHECK_INPUT(idx); CHECK_INPUT(dist2);
  CHECK_F32(sparse_feat); CHECK_F32(dist2);
  TORCH_CHECK(sparse_feat.dim() == 3 && idx.dim() == 3 && dist2.dim() == 3 && idx.size(2) == 3 && dist2.size(2) == 3,
              "tc_feature_propagation: sparse_feat (B,Ns,Cs), idx/dist2 (B,Nd,3)");
  TORCH_CHECK(idx.scalar_type() == at::kLong, "idx must be int64");
  const auto B = sparse_feat.size(0), Ns = sparse_feat.size(1), Cs = sparse_feat.size(2), Nd = idx.size(1);
  TORCH_CHECK(idx.size(0) == B && dist2.size(0) == B && dist2.size(1) == Nd, "tc_feature_propagation: shape mismatch");
  int64_t Cd = 0;
  const float *sp = nullptr;
  if (skip.has_value() && skip->defined()) {
    CHECK_INPUT((*skip)); CHECK_F32((*skip));
    TORCH_CHECK(skip->dim() == 3 && skip->size(0) == B && skip->size(1) == Nd, "skip must be (B, Nd, Cd)");
    Cd = skip->size(2);
    sp = skip->data_ptr<float>();
  }
  c10::cuda::CUDAGuard guard(sparse_feat.device());
  TcChain ch(w_hi, w_lo, biases, ks, ns, relu, out_channels, sparse_feat.device());
  auto out = at::empty({B, Nd, out_channels}, sparse_feat.options());
  check_rc(mvp_tc_fused_feature_propagation(sparse_feat.data_ptr<float>(), Cs, idx.data_ptr<int64_t>(), dist2.data_ptr<float>(), sp,
                                            Cd, B, Ns, Nd, (float)eps, &ch.c, out.data_ptr<float>(), cur_stream()));
  return out;
}

int64_t index_errors_fetch_and_clear() {
  uint64_t n = 0;
  check_rc(mvp_index_errors_fetch_and_clear(cur_stream(), &n));
  return (int64_t)n;
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "mvpnet_b200: sm_100a kernels behind the mvpnet.ops extension interface";
  m.def("abi_version", &mvp_abi_version);
  m.def("index_errors_fetch_and_clear", &index_errors_fetch_and_clear);
  m.def("set_grid_search", &set_grid_search);
  auto fps = m.def_submodule("fps_cuda");
  fps.def("farthest_point_sample", &farthest_point_sample, "Farthest point sampling (CUDA)");
  auto bq = m.def_submodule("ball_query_cuda");
  bq.def("ball_query", &ball_query, "Ball query (CUDA)");
  auto bqd = m.def_submodule("ball_query_distance_cuda");
  bqd.def("ball_query_distance", &ball_query_distance, "Ball query with distance (CUDA)");
  auto gp = m.def_submodule("group_points_cuda");
  gp.def("group_points_forward", &group_points_forward, "Group points forward (CUDA)");
  gp.def("group_points_backward", &group_points_backward, "Group points backward (CUDA)");
  gp.def("group_points_backward_det", &group_points_backward_det, "Group points backward, deterministic summation order (CUDA)");
  auto knn = m.def_submodule("knn_distance_cuda");
  knn.def("knn_distance", &knn_distance, "k-nearest neighbor with distance (CUDA)");
  auto ip = m.def_submodule("interpolate_cuda");
  ip.def("interpolate_forward", &interpolate_forward, "Interpolate feature forward (CUDA)");
  ip.def("interpolate_backward", &interpolate_backward, "Interpolate feature backward (CUDA)");
  ip.def("interpolate_backward_det", &interpolate_backward_det, "Interpolate feature backward, deterministic summation order (CUDA)");
  auto ds = m.def_submodule("unproject_cuda");
  ds.def("unproject", &unproject, "Depth unprojection (CUDA)");
  ds.def("decode_rgb_u8", &decode_rgb_u8, "uint8 HWC colour -> normalised float32 CHW (CUDA)");
  ds.def("decode_depth_u16", &decode_depth_u16, "uint16 millimetre depth -> float32 metres (CUDA)");
  ds.def("knn_pixels", &knn_pixels, "2D->3D k-NN over valid pixels (CUDA)", py::arg("query"), py::arg("pix_xyz"),
         py::arg("mask"), py::arg("k"), py::arg("exhaustive") = false);
  auto tr = m.def_submodule("train_cuda");
  tr.def("seg_loss_forward", &seg_loss_forward, "weighted cross entropy + confusion matrix, deterministic (CUDA)");
  tr.def("seg_loss_backward", &seg_loss_backward, "gradient of the weighted cross entropy w.r.t. the logits (CUDA)");
  tr.def("seg_confusion", &seg_confusion, "confusion matrix of argmax(logit) vs label, accumulated in place (CUDA)");
  auto fz = m.def_submodule("fused_cuda");
  fz.def("set_abstraction", &fused_set_abstraction, "gather + MLP + max (CUDA)");
  fz.def("feature_aggregation", &fused_feature_aggregation, "pixel gather + relation + MLP + sum/max (CUDA)");
  fz.def("feature_propagation", &fused_feature_propagation, "3-NN interpolate + concat + MLP (CUDA)");
  fz.def("tc_chain_supported", &tc_chain_supported, "does the chain fit the tcgen05 kernel");
  fz.def("tc_set_abstraction", &tc_set_abstraction, "gather + MLP (tcgen05) + max");
  fz.def("tc2_prof_dump", [](const std::string &tag) { mvp_tc2_prof_dump(tag.c_str()); }, "debug: print tc2 phase clocks (MVPNET_B200_TC2_PROF=1)");
  fz.def("tc2_supported", &tc2_supported, "does the chain fit the pre-split-input tcgen05 kernel (csrc/tc2_mlp.cu)");
  fz.def("tc2_set_abstraction", &tc2_set_abstraction, "pre-split gather (cp.async, swizzled operand) + MLP (tcgen05, TMEM activations) + max");
  fz.def("tc2_feature_aggregation", &tc2_feature_aggregation, "pre-split pixel gather + relation + MLP (tcgen05, TMEM activations) + reduce over k");
  fz.def("tc_conv3x3", &tc_conv3x3, "3x3 conv on split-planar activations, tcgen05 bf16 hi/lo x3 (+bias, residual, ReLU)");
  fz.def("tc_conv3x3_pair_supported", [](int64_t cout, int64_t h) { return mvp_tc_conv3x3_pair_supported(cout, h) != 0; }, "CTA-pair 3x3 kernel usable for this layer");
  fz.def("tc_conv3x3_nt", &mvp_tc_conv3x3_nt, "output-channel block width of the packed 3x3 weights");
  fz.def("tc_conv_general", &tc_conv_general, "tap-staged conv / 2x2 transposed conv on split-planar activations (tcgen05)");
  fz.def("unfold_stem", &unfold_stem, "fp32 NCHW image -> row-unfolded split-planar (32 channels) for the 7x7 stem");
  fz.def("maxpool3x3s2_planar", &maxpool3x3s2_planar, "3x3/s2/p1 max-pool on split-planar activations");
  fz.def("split_planar", &split_planar, "fp32 NHWC -> split-planar bf16 hi/lo");
  fz.def("merge_planar", &merge_planar, "split-planar bf16 hi/lo -> fp32 NHWC");
  fz.def("tc_feature_aggregation", &tc_feature_aggregation, "pixel gather + relation + MLP (tcgen05) + sum/max");
  fz.def("tc_feature_propagation", &tc_feature_propagation, "3-NN interpolate + concat + MLP (tcgen05)");
}
