// Shared by every translation unit of libembodied_b200.so: error reporting and
// the launch counter behind emb_launch_count().
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>

namespace emb {
int fail(int code, const char* fmt, ...);
int fail_cuda(const char* who);   // formats cudaGetLastError()
void count_launch();
}  // namespace emb
