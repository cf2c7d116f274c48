// hop_common.cuh -- host-side helpers shared by the .cu translation units of libhop_b200.so.
#pragma once
#include <cuda_runtime.h>

#include <cstdio>

namespace hop {
void set_last_error(const char* msg);
int report_cuda(cudaError_t e, const char* where);
inline int check_launch(const char* where) { return report_cuda(cudaGetLastError(), where); }
}  // namespace hop
