// Developer probe: accuracy of rcp.approx.ftz.f64 (MUFU.RCP64H) and of one / two Newton refinements.
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
__global__ void k(double* out, int n) {
    double m0 = 0, m1 = 0, m2 = 0, m3 = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double d = 0.5 + 1.5 * (double)i / n + 1e-9 * (i % 977);        // [0.5, 2)
        double r;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
        const double exact = 1.0 / d;
        m0 = fmax(m0, fabs(r - exact) / exact);
        double e = fma(-d, r, 1.0);
        const double q = fma(r, e, r);                     // quadratic
        m1 = fmax(m1, fabs(q - exact) / exact);
        const double t = fma(e, e, e);
        const double c = fma(r, t, r);                     // cubic (what the kernels use)
        m2 = fmax(m2, fabs(c - exact) / exact);
        double e2 = fma(-d, q, 1.0);
        const double qq = fma(q, e2, q);                   // two quadratic steps
        m3 = fmax(m3, fabs(qq - exact) / exact);
    }
    atomicMax((unsigned long long*)&out[0], __double_as_longlong(m0));
    atomicMax((unsigned long long*)&out[1], __double_as_longlong(m1));
    atomicMax((unsigned long long*)&out[2], __double_as_longlong(m2));
    atomicMax((unsigned long long*)&out[3], __double_as_longlong(m3));
}
int main() {
    double* d; cudaMalloc(&d, 32); cudaMemset(d, 0, 32);
    k<<<1024, 256>>>(d, 1 << 28);
    double h[4]; cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
    printf("max relative error: seed %.3g (2^%.1f), +quadratic %.3g, +cubic %.3g, +2 quadratic %.3g\n", h[0], log2(h[0]), h[1], h[2], h[3]);
    return 0;
}
