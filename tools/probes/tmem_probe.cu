// Developer probe: tensor memory (TMEM) as a lane-private FP64 scratchpad.
// Each warp of a 4-warp CTA owns 32 TMEM lanes; a thread stores doubles as pairs of 32-bit columns of its lane with
// tcgen05.st.32x32b and reads them back with tcgen05.ld.32x32b.  Checks the round trip and times it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_probe tmem_probe.cu && ./tmem_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tmem_st8(uint32_t addr, const double (&v)[4]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(addr),
                 "r"(__double2loint(v[0])), "r"(__double2hiint(v[0])), "r"(__double2loint(v[1])), "r"(__double2hiint(v[1])),
                 "r"(__double2loint(v[2])), "r"(__double2hiint(v[2])), "r"(__double2loint(v[3])), "r"(__double2hiint(v[3])));
}
__device__ __forceinline__ void tmem_ld8(uint32_t addr, double (&v)[4]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(addr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = __hiloint2double((int)r[2 * i + 1], (int)r[2 * i]);
}

__global__ void __launch_bounds__(128) probe(int ncols, int reps, int* bad, long long* cycles) {
    __shared__ uint32_t s_base;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        uint32_t dst = (uint32_t)__cvta_generic_to_shared(&s_base);
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst), "r"(ncols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = s_base + ((uint32_t)(warp * 32) << 16);     // this warp's lane quarter
    int errors = 0;
    // fill: column pair j holds f(block, thread, j)
    for (int c = 0; c < ncols; c += 8) {
        double v[4];
        for (int i = 0; i < 4; ++i) v[i] = 1000.0 * blockIdx.x + threadIdx.x + 1e-3 * (c / 2 + i) + 1e-9;
        tmem_st8(base + c, v);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
    for (int c = 0; c < ncols; c += 8) {
        double v[4];
        tmem_ld8(base + c, v);
        for (int i = 0; i < 4; ++i) errors += (v[i] != 1000.0 * blockIdx.x + threadIdx.x + 1e-3 * (c / 2 + i) + 1e-9);
    }
    // timing: read-modify-write sweeps over the columns
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        for (int c = 0; c < ncols; c += 8) {
            double v[4];
            tmem_ld8(base + c, v);
            for (int i = 0; i < 4; ++i) v[i] = v[i] * 1.0000001 + 1e-12;
            tmem_st8(base + c, v);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
    }
    long long t1 = clock64();
    if (errors) atomicAdd(bad, errors);
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_base), "r"(ncols));
}

int main() {
    int* bad; long long* cyc;
    cudaMalloc(&bad, 4); cudaMalloc(&cyc, 8); cudaMemset(bad, 0, 4);
    const int ncols = 128, reps = 200;
    int occ = -1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, probe, 128, 0);
    printf("occupancy API for the TMEM kernel (128 threads, no smem): %d CTAs/SM\n", occ);
    for (int ctas_per_sm = 1; ctas_per_sm <= 3; ++ctas_per_sm) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        probe<<<148 * ctas_per_sm, 128>>>(ncols, reps, bad, cyc);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        printf("  kernel %.3f ms; ", ms);
        int hb; long long hc;
        cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost); cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
        const double bytes_per_sm = (double)ctas_per_sm * 128 * ncols * 4 * reps;      // read (and as much written)
        printf("grid %d x SMs: %s, mismatches %d, %lld cycles for %d sweeps -> %.1f B/clk/SM read + as much written\n",
               ctas_per_sm, cudaGetErrorString(e), hb, hc, reps, bytes_per_sm / hc);
    }
    return 0;
}
