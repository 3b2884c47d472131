/* Compatibility shim so that the UNMODIFIED reference extension sources
 * (/root/reference/mvpnet/ops/cuda/*.cu, written for PyTorch 1.2) compile against PyTorch 2.x.
 * Injected with -I by oracle/build_ref.py; the reference sources are compiled where they lie and
 * are never copied.  Only the names the reference uses are provided:
 *   <THC/THC.h> itself (removed from PyTorch), THCudaCheck, THArgCheck, and the glog-style
 *   CHECK_EQ / CHECK_GE / CHECK_GT macros that c10 no longer defines.
 * TEST INFRASTRUCTURE (cross-check of the oracle on the GPU box), not product. */
#pragma once
#include <c10/util/Exception.h>
#include <c10/cuda/CUDAException.h>
#include <ATen/cuda/CUDAContext.h>

#ifndef THCudaCheck
#define THCudaCheck(expr) C10_CUDA_CHECK(expr)
#endif
#ifndef THArgCheck
#define THArgCheck(cond, argn, ...) TORCH_CHECK(cond, __VA_ARGS__)
#endif
#ifndef CHECK_EQ
#define CHECK_EQ(a, b) TORCH_CHECK((a) == (b), "Check failed: " #a " == " #b)
#endif
#ifndef CHECK_GE
#define CHECK_GE(a, b) TORCH_CHECK((a) >= (b), "Check failed: " #a " >= " #b)
#endif
#ifndef CHECK_GT
#define CHECK_GT(a, b) TORCH_CHECK((a) > (b), "Check failed: " #a " > " #b)
#endif
