// Instantiations of the fast EM kernel for one guide count (compile with -DJACKS_G=<1..8>).
// One translation unit per G keeps the build parallel; em_dispatch.cu switches over them.
#include "em_fast.cuh"

#ifndef JACKS_G
#error "compile with -DJACKS_G=<guides per gene>"
#endif

namespace jacks {

namespace {

template <int G, int W, bool UNIT_C, bool PLAIN>
cudaError_t launch_inst(const EmLaunch& a, int n_sm, int grid_limit, cudaStream_t st, int* grid_out) {
    auto k = em_fast_kernel<G, W, UNIT_C, PLAIN>;
    const size_t smem = (size_t)em_fast_smem_doubles(G, W, a.L) * sizeof(double);
    if (smem > kMaxDynSmem) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, W * 32, smem);
    if (e != cudaSuccess) return e;
    if (occ < 1) return cudaErrorLaunchOutOfResources;
    if (occ > kMaxFastCtasPerSm) occ = kMaxFastCtasPerSm;     // tau_pr_den slabs are provisioned for this many
    long long grid = (long long)occ * n_sm;
    if (grid > a.n_work) grid = a.n_work;
    if (grid_limit > 0 && grid > grid_limit) grid = grid_limit;
    if (grid < 1) grid = 1;
    if (grid_out) *grid_out = (int)grid;
    k<<<(int)grid, W * 32, smem, st>>>(a);
    return cudaGetLastError();
}

template <int G, int W>
cudaError_t launch_w(const EmLaunch& a, int n_sm, int grid_limit, cudaStream_t st, int* grid_out) {
    // tau = (tps + 0.5) / (...): with the default tps = 0.5 the numerator is exactly 1
    const bool plain = !a.apply_w_hp && !(a.ctrl_lines == 1 && a.ctrl_rowmean_fill);
    if (a.tps + 0.5 == 1.0) {
        if (plain) return launch_inst<G, W, true, true>(a, n_sm, grid_limit, st, grid_out);
        return launch_inst<G, W, true, false>(a, n_sm, grid_limit, st, grid_out);
    }
    return launch_inst<G, W, false, false>(a, n_sm, grid_limit, st, grid_out);
}

}  // namespace

#define JACKS_CAT2(a, b) a##b
#define JACKS_CAT(a, b) JACKS_CAT2(a, b)

cudaError_t JACKS_CAT(launch_em_fast_g, JACKS_G)(int W, const EmLaunch& a, int n_sm, int grid_limit, cudaStream_t st,
                                                 int* grid_out) {
    switch (W) {
        case 1: return launch_w<JACKS_G, 1>(a, n_sm, grid_limit, st, grid_out);
        case 2: return launch_w<JACKS_G, 2>(a, n_sm, grid_limit, st, grid_out);
        case 4: return launch_w<JACKS_G, 4>(a, n_sm, grid_limit, st, grid_out);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace jacks
