// Error state, version and launch accounting of libdac_b200.so.
#include <atomic>
#include <stdarg.h>
#include <stdio.h>

#include "../../include/dac_b200.h"
#include "common.h"

namespace dac {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};
static thread_local int g_pdl = 0;
bool pdl_enabled() { return g_pdl != 0; }

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(-100 - (int)e, "%s: launch failed: %s", what, cudaGetErrorString(e));
  return 0;
}

}  // namespace dac

extern "C" int dac_version(void) { return 100; }
extern "C" void dac_set_pdl(int32_t on) { dac::g_pdl = on; }
extern "C" int dac_abi_sizes(int32_t* conv_desc_bytes, int32_t* embed_weights_bytes) {
  if (conv_desc_bytes) *conv_desc_bytes = (int32_t)sizeof(dac_conv_desc);
  if (embed_weights_bytes) *embed_weights_bytes = (int32_t)sizeof(dac_embed_weights);
  return 0;
}
extern "C" const char* dac_last_error(void) { return dac::g_err; }
extern "C" int64_t dac_launch_count(void) { return dac::g_launches.load(); }
extern "C" void dac_reset_launch_count(void) { dac::g_launches.store(0); }
