// Internals shared by the C-ABI translation units (not part of the public header).
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stddef.h>
#include <stdint.h>

namespace jacks {

// Records the message jacks_last_error() returns and passes `code` through.
int fail(int code, const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

// Grow-only device buffer (avoids cudaMalloc/cudaFree on every call).
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes);
    void release();
};

}  // namespace jacks

struct JacksCtx {
    int device;
    int n_sm;
    int64_t launches;
    cudaStream_t stream;        // compute + default copies of the host entry points
    cudaStream_t copy_stream;   // overlapped copies
    // data cubes kept resident between host calls (jacks_em_batched_host keep_resident)
    jacks::DevBuf res_test, res_ctrl, res_fx;
    int64_t res_N = 0;
    int res_L = 0, res_ctrl_lines = 0;
    bool res_valid = false;
    jacks::DevBuf out[8];       // device-side outputs of the host entry points
    jacks::DevBuf tmp;          // scratch of the p-value / preprocessing entry points
};
