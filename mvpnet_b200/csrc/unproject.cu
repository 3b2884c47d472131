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

// ---- decoding of the STORED input formats on the device (the host uploads what the dataset stores, 4x / 2x fewer bytes):
//   colour  uint8 HWC (PIL, scannet_2d3d.py:229-246) -> float32 CHW: (u8 / 255 - mean[c]) / std[c], each step rounded to
//           float32 in that order (np.float32 image / 255., then the normaliser's subtract and divide);
//   depth   uint16 millimetres (scannet_2d3d.py:249-251) -> float32 metres: float32(mm) / 1000.
__global__ void __launch_bounds__(256)
decode_rgb_kernel(const uint8_t *__restrict__ rgb /*[N][H][W][3]*/, long long total /*N*H*W*/, int hw, float m0, float m1, float m2, float s0,
                  float s1, float s2, float *__restrict__ out /*[N][3][H][W]*/) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const long long n = i / hw;
    const int p = (int)(i - n * hw);
    const uint8_t *s = rgb + i * 3;
    float *o = out + n * 3 * hw + p;
    o[0] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)s[0], 255.f), m0), s0);
    o[hw] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)s[1], 255.f), m1), s1);
    o[2 * (size_t)hw] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)s[2], 255.f), m2), s2);
  }
}

__global__ void __launch_bounds__(256)
decode_depth_kernel(const uint16_t *__restrict__ mm, long long total, float *__restrict__ out) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) out[i] = __fdiv_rn((float)mm[i], 1000.f);
}

}  // namespace mvp

extern "C" int mvp_decode_rgb_u8(const uint8_t *rgb_hwc, int64_t N, int64_t H, int64_t W, const float *mean3, const float *std3,
                                 float *out_chw, mvp_stream_t stream) {
  using namespace mvp;
  MVP_REQUIRE(N >= 0 && H >= 0 && W >= 0 && H * W < (1LL << 31), MVP_ERR_INVALID_ARG, "decode_rgb_u8: bad sizes");
  if (N * H * W == 0) return 0;
  MVP_REQUIRE(rgb_hwc && mean3 && std3 && out_chw, MVP_ERR_NULL, "decode_rgb_u8: null pointer (mean3 / std3 are HOST arrays of 3 floats)");
  const long long total = N * H * W;
  const long long want = (total + 255) / 256, cap = (long long)sm_count() * 16;
  decode_rgb_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(rgb_hwc, total, (int)(H * W), mean3[0], mean3[1], mean3[2], std3[0],
                                                                                           std3[1], std3[2], out_chw);
  return launch_status("decode_rgb_u8");
}

extern "C" int mvp_decode_depth_u16(const uint16_t *depth_mm, int64_t count, float *depth_m, mvp_stream_t stream) {
  using namespace mvp;
  MVP_REQUIRE(count >= 0, MVP_ERR_INVALID_ARG, "decode_depth_u16: bad size");
  if (count == 0) return 0;
  MVP_REQUIRE(depth_mm && depth_m, MVP_ERR_NULL, "decode_depth_u16: null pointer");
  const long long want = (count + 255) / 256, cap = (long long)sm_count() * 16;
  decode_depth_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(depth_mm, count, depth_m);
  return launch_status("decode_depth_u16");
}

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
