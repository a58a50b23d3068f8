// Error plumbing shared by the extern "C" translation units.
#pragma once
#include <cuda_runtime.h>

#include <string>

namespace r3m {
extern thread_local std::string g_last_error;
int fail(int code, const std::string& msg);
int fail_cuda(cudaError_t e, const char* what);
}  // namespace r3m
