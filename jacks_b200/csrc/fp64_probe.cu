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

// ---- developer probe: dependent-issue latencies of DFMA and MUFU.RCP64H (cycles) ----------------
namespace {
__global__ void latency_kernel(double* out, long long* cyc, double a, double b) {
    double x = (double)threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) x = fma(x, a, b);
    }
    long long t1 = clock64();
    double r = x + 1.5;
#pragma unroll 1
    for (int i = 0; i < 64; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) { double q; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(q) : "d"(r)); r = q; }
    }
    long long t2 = clock64();
    // 4 independent chains: throughput-limited issue interval for one warp
    double y0 = x, y1 = x + 1, y2 = x + 2, y3 = x + 3;
#pragma unroll 1
    for (int i = 0; i < 64; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { y0 = fma(y0, a, b); y1 = fma(y1, a, b); y2 = fma(y2, a, b); y3 = fma(y3, a, b); }
    }
    long long t3 = clock64();
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; }
    out[threadIdx.x] = x + r + y0 + y1 + y2 + y3;
}
}  // namespace

// throughput of DFMA with three DISTINCT register-pair operands per instruction (register-file bandwidth)
__global__ void __launch_bounds__(256) dfma_3op_kernel(double* out, int iters, const double* in) {
    double acc[8], b[8], c[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { acc[i] = in[threadIdx.x + i]; b[i] = in[threadIdx.x + 8 + i]; c[i] = in[threadIdx.x + 16 + i]; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fma(acc[i], b[i], c[i]);
#pragma unroll
        for (int i = 0; i < 8; ++i) b[i] = fma(b[i], c[i], acc[i]);
#pragma unroll
        for (int i = 0; i < 8; ++i) c[i] = fma(c[i], acc[i], b[i]);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i] + b[i] + c[i];
    if (s == 123.456) out[0] = s;
}

extern "C" __attribute__((visibility("default"))) int jacks_debug_3op_probe(JacksCtx* ctx, double* flops_out) {
    using namespace jacks;
    cudaSetDevice(ctx->device);
    int rc = ctx->tmp.reserve(8192);
    if (rc != JACKS_OK) return rc;
    cudaMemsetAsync(ctx->tmp.p, 0, 8192, ctx->stream);
    const int iters = 4096, threads = 256, ctas = ctx->n_sm * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 1e30;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0, ctx->stream);
        dfma_3op_kernel<<<ctas, threads, 0, ctx->stream>>>((double*)ctx->tmp.p, iters, (const double*)ctx->tmp.p);
        cudaEventRecord(e1, ctx->stream);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms * 1e-3 < best) best = ms * 1e-3;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *flops_out = 2.0 * 24.0 * (double)iters * (double)threads * (double)ctas / best;
    return JACKS_OK;
}

extern "C" __attribute__((visibility("default"))) int jacks_debug_latency_probe(JacksCtx* ctx, double* dfma_cycles,
                                                                               double* rcp_cycles, double* dfma4_cycles) {
    using namespace jacks;
    cudaSetDevice(ctx->device);
    int rc = ctx->tmp.reserve(4096);
    if (rc != JACKS_OK) return rc;
    double* out = (double*)ctx->tmp.p;
    long long* cyc = (long long*)((char*)ctx->tmp.p + 2048);
    latency_kernel<<<1, 32, 0, ctx->stream>>>(out, cyc, 0.999999, 1e-9);
    latency_kernel<<<1, 32, 0, ctx->stream>>>(out, cyc, 0.999999, 1e-9);
    long long h[3];
    cudaError_t e = cudaMemcpyAsync(h, cyc, sizeof h, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) return cuda_fail(e, "latency_kernel");
    *dfma_cycles = h[0] / 1024.0; *rcp_cycles = h[1] / 1024.0; *dfma4_cycles = h[2] / 1024.0;
    return JACKS_OK;
}
