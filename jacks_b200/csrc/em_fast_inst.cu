// Instantiations of the fast EM kernel for one guide count (compile with -DJACKS_G=<1..8>).
// One translation unit per G keeps the build parallel; em_dispatch.cu switches over them.
#include <algorithm>
#include <cstdlib>

#include "em_fast.cuh"

#ifndef JACKS_G
#error "compile with -DJACKS_G=<guides per gene>"
#endif

namespace jacks {

namespace {

// Tensor-memory variant: four one-warp teams per CTA, tau in TMEM (em_fast.cuh).  The resident CTAs per SM are the
// smaller of what shared memory / registers and what the 512 TMEM columns allow.
template <int G, bool UNIT_C, bool PLAIN, int TM, bool WS = false>
cudaError_t launch_tmem(const EmLaunch& a0, int n_sm, int grid_limit, cudaStream_t st, int* grid_out) {
    auto k = em_fast_kernel<G, 1, UNIT_C, PLAIN, TM, WS>;
    EmLaunch a = a0;
    a.tmem_cols = tmem_cols_for(G, a.L, TM);
    a.tmem_pd_off = (TM == 2) ? tmem_array_cols(G, a.L) : 0;
    const size_t smem = (size_t)em_fast_tmem_team_doubles(G, a.L, WS) * kTmemTeams * sizeof(double);
    if (a.tmem_cols == 0 || smem > kMaxDynSmem) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    // cudaOccupancyMaxActiveBlocksPerMultiprocessor answers 1 for any kernel that allocates tensor memory, although the
    // hardware co-schedules such CTAs as long as their columns fit (tools/probes/tmem_probe.cu): count the resources here.
    cudaFuncAttributes fa;
    if ((e = cudaFuncGetAttributes(&fa, k)) != cudaSuccess) return e;
    const int regs_per_cta = ((fa.numRegs + 7) / 8 * 8) * kTmemTeams * 32;
    int occ = std::min<int>(65536 / regs_per_cta, (int)((228 * 1024) / (smem + fa.sharedSizeBytes + 1024)));
    if (occ < 1) return cudaErrorLaunchOutOfResources;
    occ = std::min(occ, 512 / a.tmem_cols);
    occ = std::min(occ, kMaxFastCtasPerSm / kTmemTeams);        // tau_pr_den slabs are provisioned per team
    long long grid = (long long)occ * n_sm;
    // grid_limit < 0: leave |grid_limit| thousandths of the resident CTA slots to a kernel that runs beside this one
    if (grid_limit < 0) grid = std::max<long long>(1, grid * (1000 + grid_limit) / 1000);
    const long long want = ((long long)a.n_work + kTmemTeams - 1) / kTmemTeams;
    if (grid > want) grid = want;
    if (grid_limit > 0 && grid > grid_limit) grid = grid_limit;
    if (grid < 1) grid = 1;
    if (grid_out) *grid_out = (int)grid;
    k<<<(int)grid, kTmemTeams * 32, smem, st>>>(a);
    return cudaGetLastError();
}

// Which variant for this shape?  0: shared memory; 1: tau in TMEM; 2: tau and tau_pr_den in TMEM.
// The TMEM variants run at most two CTAs (8 genes) per SM (registers).  Measured per pass at 400 lines, 39,000 genes
// (ms; smem / mode 1 / mode 2): G=3 1.63 / 1.64 / 1.45, G=4 2.6 / 2.7 / 2.3, G=5 5.09 / 3.60 / 4.79, G=6 7.73 / 7.36 / 8.28,
// G=8 10.2 / 13.2 / 10.6; and mode 2 against shared memory at G=4 with 64 / 128 / 256 / 320 / 800 lines: 0.99 / 1.09,
// 1.57 / 1.60, 2.76 / 2.99, 3.36 / 3.76, (mode 1) 5.1 / 6.9.  Hence: mode 2 whenever both arrays of two CTAs fit the 512
// columns (it wins even against 12 shared-memory genes: a pass makes no L2 round trip); otherwise mode 1 when it puts
// more genes on an SM than shared memory does; otherwise shared memory.  JACKS_EM_TMEM=0/1/2 forces the choice where
// eligible (tests, tuning).
template <int G>
int tmem_mode(const EmLaunch& a) {
    static const int force = []() { const char* e = getenv("JACKS_EM_TMEM"); return e ? atoi(e) : -1; }();
    if (force == 0) return 0;
    const size_t team = (size_t)em_fast_tmem_team_doubles(G, a.L) * sizeof(double);
    const long long by_smem = kMaxDynSmem / (team * kTmemTeams);
    const int cols1 = tmem_cols_for(G, a.L, 1), cols2 = tmem_cols_for(G, a.L, 2);
    if (by_smem < 1 || cols1 == 0) return 0;
    const long long genes_smem = std::min<long long>(kMaxDynSmem / ((size_t)em_fast_smem_doubles(G, 1, a.L) * sizeof(double)), kMaxFastCtasPerSm);
    const long long genes1 = std::min<long long>(std::min<long long>(by_smem, 512 / cols1), JACKS_TM_CTAS) * kTmemTeams;
    const long long genes2 = cols2 ? std::min<long long>(std::min<long long>(by_smem, 512 / cols2), JACKS_TM_CTAS) * kTmemTeams : 0;
    if (force == 2) return genes2 ? 2 : 1;
    if (force == 1) return 1;
    if (genes2 >= 2 * kTmemTeams && a.L >= 32) return 2;
    if (genes1 > genes_smem) return 1;
    return 0;
}

template <int G, int W, bool UNIT_C, bool PLAIN>
cudaError_t launch_inst(const EmLaunch& a, int n_sm, int grid_limit, cudaStream_t st, int* grid_out) {
    if (W == 1) {
        const int tm = tmem_mode<G>(a);
        if (tm == 2) {
            // E[w], E[w^2] of the running pass in shared memory whenever two CTAs still fit an SM with it (L <= ~600 at
            // G = 4); JACKS_EM_WSMEM=0/1 forces the choice where it fits (tests, tuning)
            static const int force_ws = []() { const char* e = getenv("JACKS_EM_WSMEM"); return e ? atoi(e) : -1; }();
            const size_t smem_ws = (size_t)em_fast_tmem_team_doubles(G, a.L, true) * kTmemTeams * sizeof(double);
            const bool fits = JACKS_TM_CTAS * (smem_ws + 2048) <= 228 * 1024;
            if (fits && force_ws != 0) return launch_tmem<G, UNIT_C, PLAIN, 2, true>(a, n_sm, grid_limit, st, grid_out);
            return launch_tmem<G, UNIT_C, PLAIN, 2>(a, n_sm, grid_limit, st, grid_out);
        }
        if (tm == 1) return launch_tmem<G, UNIT_C, PLAIN, 1>(a, n_sm, grid_limit, st, grid_out);
    }
    auto k = em_fast_kernel<G, W, UNIT_C, PLAIN, 0>;
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
    if (grid_limit < 0) grid = std::max<long long>(1, grid * (1000 + grid_limit) / 1000);
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
