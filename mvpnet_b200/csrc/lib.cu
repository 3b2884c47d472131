// Unity translation unit of libmvpnet_b200.so: the device-side index-error counter is shared by
// several kernels, and a single TU avoids relocatable device code (-rdc) and its link step.
#include "runtime.cu"
#include "fps.cu"
#include "ball_query.cu"
#include "knn_distance.cu"
#include "group_points.cu"
#include "interpolate.cu"
#include "unproject.cu"
#include "knn_pixels.cu"
#include "fused_mlp.cu"
#include "tc_mlp.cu"
#include "tc_conv.cu"
#include "tc_convg.cu"
