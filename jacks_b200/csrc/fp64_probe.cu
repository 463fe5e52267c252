// FP64 FMA-pipe peak microbenchmark (roofline denominator of the EM and KDE kernels).
// MEASURED_PEAKS.json carries HBM and bf16 tensor peaks only, so bench.py measures this in-run.
#include <cuda_runtime.h>

#include "../../include/jacks_b200.h"
#include "capi_internal.h"

namespace {

// 16 independent accumulators per thread, fully unrolled: no dependent-issue stalls, no memory.
__global__ void __launch_bounds__(256) dfma_chain_kernel(double* out, int iters, double a, double b) {
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = (double)(threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    if (s == 123.456) out[0] = s;      // keeps the chain alive without a store in practice
}

}  // namespace

extern "C" int jacks_fp64_peak_flops(JacksCtx* ctx, double* flops_out, double* seconds_out) {
    using namespace jacks;
    if (!ctx || !flops_out) return fail(JACKS_E_INVALID, "jacks_fp64_peak_flops: NULL argument");
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    int rc = ctx->tmp.reserve(64);
    if (rc != JACKS_OK) return rc;
    const int iters = 8192, threads = 256, ctas = ctx->n_sm * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 1e30;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0, ctx->stream);
        dfma_chain_kernel<<<ctas, threads, 0, ctx->stream>>>((double*)ctx->tmp.p, iters, 0.999999, 1e-9);
        cudaEventRecord(e1, ctx->stream);
        if ((e = cudaEventSynchronize(e1)) != cudaSuccess) { cudaEventDestroy(e0); cudaEventDestroy(e1); return cuda_fail(e, "dfma_chain_kernel"); }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        ctx->launches++;
        if (rep > 0 && ms * 1e-3 < best) best = ms * 1e-3;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    const double flops = 2.0 * 16.0 * (double)iters * (double)threads * (double)ctas;
    *flops_out = flops / best;
    if (seconds_out) *seconds_out = best;
    return JACKS_OK;
}
