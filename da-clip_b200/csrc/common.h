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
}  // namespace dac
