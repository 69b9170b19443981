// Host-side helpers shared by the translation units of libdac_b200.so: thread-local error string,
// launch accounting, launch-error check.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dac {
int set_error(int code, const char* fmt, ...);
// Counts the launch and turns a launch-time CUDA error into a negative return code.
int check_launch(const char* what);
inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Programmatic dependent launch (dac_set_pdl): while the flag is up, launch sites that go through launch_k() add
// cudaLaunchAttributeProgrammaticStreamSerialization, so the grid may be scheduled - barrier / TMEM set-up, resident
// weight loads - while the previous kernel of the stream drains.  Only kernels that execute griddepcontrol.wait before
// touching anything an earlier kernel wrote (or still reads) may be launched this way.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  if (pdl_enabled()) {
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
  }
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
// ... the same for a kernel that runs in clusters of `cluster` CTAs along x (1 = no cluster attribute)
template <typename... KArgs, typename... Args>
inline void launch_k_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster,
                             Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  int n = 0;
  if (pdl_enabled()) {
    at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster > 1) {
    at[n].id = cudaLaunchAttributeClusterDimension;
    at[n].val.clusterDim.x = static_cast<unsigned>(cluster);
    at[n].val.clusterDim.y = 1;
    at[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = n ? at : nullptr;
  cfg.numAttrs = n;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
}  // namespace dac
