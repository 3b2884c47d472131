// Library-level state of the C ABI: error text, ABI version, index-error counter.
#include <stdarg.h>

#include "common.cuh"

namespace mvp {

static thread_local char t_error[512] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_error, sizeof(t_error), fmt, ap);
  va_end(ap);
}

}  // namespace mvp

extern "C" const char *mvp_last_error(void) { return mvp::t_error; }

extern "C" int mvp_abi_version(void) { return 2; }

extern "C" int mvp_index_errors_fetch_and_clear(mvp_stream_t stream_, uint64_t *count) {
  using namespace mvp;
  cudaStream_t stream = (cudaStream_t)stream_;
  MVP_REQUIRE(count, MVP_ERR_NULL, "index_errors: null pointer");
  unsigned long long host = 0, zero = 0;
  cudaError_t e = cudaMemcpyFromSymbolAsync(&host, g_index_errors, sizeof(host), 0, cudaMemcpyDeviceToHost, stream);
  if (e == cudaSuccess) e = cudaMemcpyToSymbolAsync(g_index_errors, &zero, sizeof(zero), 0, cudaMemcpyHostToDevice, stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  if (e != cudaSuccess) { set_error("index_errors: %s", cudaGetErrorString(e)); return (int)e; }
  *count = (uint64_t)host;
  return 0;
}
