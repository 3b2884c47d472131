// Unprojection of depth maps into world-space points + validity mask, for sm_100a.
//
// Semantics: mvpnet/data/scannet_2d3d.py:33-39 (depth2xyz), :255-262 (z_cam > 0, camera -> world),
// :273-281 (chunk box with 0.1 m margin), :317 (float32 cast of the emitted image_xyz).  The
// reference evaluates this in float64 (int64 pixel grid x float32 inverse intrinsics promotes), and
// feeds the float64 points to its k-NN; both the float64 points and the float32 `image_xyz` are
// produced here.  Operation order is the one stated in oracle/mvp_oracle.c (mvpo_unproject).
//
// HBM-bound elementwise kernel: 4 B read, 12 + 24 + 1 B written per pixel; one thread per pixel,
// intrinsics/pose broadcast from shared memory.
#include "common.cuh"

namespace mvp {

__global__ void __launch_bounds__(256)
unproject_kernel(const float *__restrict__ depth, const float *__restrict__ cam_inv, const float *__restrict__ pose,
                 const double *__restrict__ chunk_box, int nv, int hw, int w, double *__restrict__ xyz64,
                 float *__restrict__ xyz32, uint8_t *__restrict__ mask) {
  __shared__ double s_ki[9], s_p[12], s_box[4];
  const int view = blockIdx.y;                 // b * nv + f
  const int b = view / nv;
  if (threadIdx.x < 9) s_ki[threadIdx.x] = (double)cam_inv[(size_t)view * 9 + threadIdx.x];
  if (threadIdx.x >= 32 && threadIdx.x < 44) s_p[threadIdx.x - 32] = (double)pose[(size_t)view * 16 + (threadIdx.x - 32)];
  if (chunk_box && threadIdx.x >= 64 && threadIdx.x < 68) s_box[threadIdx.x - 64] = chunk_box[(size_t)b * 4 + (threadIdx.x - 64)];
  __syncthreads();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= hw) return;
  const int v = p / w, u = p - v * w;
  const size_t g = (size_t)view * hw + p;
  const double d = (double)__ldg(depth + g);
  double cam[3], wld[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    double c = __dmul_rn(s_ki[i * 3 + 0], (double)u);
    c = __dadd_rn(c, __dmul_rn(s_ki[i * 3 + 1], (double)v));
    c = __dadd_rn(c, s_ki[i * 3 + 2]);
    cam[i] = __dmul_rn(c, d);
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    double a = __dmul_rn(cam[0], s_p[i * 4 + 0]);
    a = __fma_rn(cam[1], s_p[i * 4 + 1], a);
    a = __fma_rn(cam[2], s_p[i * 4 + 2], a);
    wld[i] = __dadd_rn(a, s_p[i * 4 + 3]);
  }
  bool ok = cam[2] > 0.0;
  if (chunk_box) {
    const double margin = 0.1;
    ok = ok && wld[0] > __dsub_rn(s_box[0], margin) && wld[0] < __dadd_rn(s_box[2], margin) &&
         wld[1] > __dsub_rn(s_box[1], margin) && wld[1] < __dadd_rn(s_box[3], margin);
  }
  if (xyz64) { xyz64[g * 3] = wld[0]; xyz64[g * 3 + 1] = wld[1]; xyz64[g * 3 + 2] = wld[2]; }
  if (xyz32) { xyz32[g * 3] = (float)wld[0]; xyz32[g * 3 + 1] = (float)wld[1]; xyz32[g * 3 + 2] = (float)wld[2]; }
  if (mask) mask[g] = ok ? 1 : 0;
}

}  // namespace mvp

extern "C" int mvp_unproject(const float *depth, const float *cam_inv, const float *pose, const double *chunk_box,
                             int64_t B, int64_t nv, int64_t h, int64_t w, double *xyz64, float *xyz32, uint8_t *mask,
                             mvp_stream_t stream) {
  using namespace mvp;
  MVP_REQUIRE(B >= 0 && nv >= 0 && h >= 0 && w >= 0, MVP_ERR_INVALID_ARG, "unproject: negative size");
  if (B * nv * h * w == 0) return 0;
  MVP_REQUIRE(depth && cam_inv && pose, MVP_ERR_NULL, "unproject: null pointer");
  MVP_REQUIRE(B * nv <= 65535 && h * w < (1LL << 31), MVP_ERR_UNSUPPORTED, "unproject: too many views or pixels");
  dim3 grid((unsigned)((h * w + 255) / 256), (unsigned)(B * nv));
  unproject_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(depth, cam_inv, pose, chunk_box, (int)nv, (int)(h * w), (int)w,
                                                            xyz64, xyz32, mask);
  return launch_status("unproject");
}
