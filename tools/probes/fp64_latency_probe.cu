// Developer probe (standalone; NOT part of libjacks_b200.so): dependent-issue latencies of DFMA and MUFU.RCP64H, and the
// throughput of DFMA with three DISTINCT register-pair operands (register-file bandwidth) against the shared-operand peak.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/_build/fp64_latency_probe tools/probes/fp64_latency_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

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


int main() {
    int n_sm = 0;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0);
    double* buf; long long* cyc;
    cudaMalloc(&buf, 8192); cudaMalloc(&cyc, 64);
    cudaMemset(buf, 0, 8192);
    latency_kernel<<<1, 32>>>(buf, cyc, 0.999999, 1e-9);
    latency_kernel<<<1, 32>>>(buf, cyc, 0.999999, 1e-9);
    long long h[3];
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    printf("dependent DFMA %.2f cycles, dependent MUFU.RCP64H %.2f cycles, 4 independent DFMA chains %.2f cycles per DFMA\n",
           h[0] / 1024.0, h[1] / 1024.0, h[2] / 1024.0);
    const int iters = 4096, threads = 256, ctas = n_sm * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 1e30;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        dfma_3op_kernel<<<ctas, threads>>>(buf, iters, buf);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms * 1e-3 < best) best = ms * 1e-3;
    }
    printf("DFMA with 3 distinct register operands: %.2f TFLOP/s\n", 2.0 * 24.0 * iters * (double)threads * ctas / best / 1e12);
    return cudaDeviceSynchronize() == cudaSuccess ? 0 : 1;
}
