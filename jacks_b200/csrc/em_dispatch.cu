// Switch from a runtime guide count to the per-G translation units (em_fast_inst.cu).
#include "em_common.cuh"
#include <algorithm>

#include "em_fast.cuh"

namespace jacks {

#define JACKS_DECL(g)                                                                                       \
    cudaError_t launch_em_fast_g##g(int W, const EmLaunch& a, int n_sm, int grid_limit, cudaStream_t st,    \
                                    int* grid_out);
JACKS_DECL(1) JACKS_DECL(2) JACKS_DECL(3) JACKS_DECL(4) JACKS_DECL(5) JACKS_DECL(6) JACKS_DECL(7) JACKS_DECL(8)
JACKS_DECL(9) JACKS_DECL(10) JACKS_DECL(11) JACKS_DECL(12)
#undef JACKS_DECL

cudaError_t launch_em_fast(int G, int W, const EmLaunch& a, int n_sm, int grid_limit, cudaStream_t st, int* grid_out) {
    switch (G) {
        case 1: return launch_em_fast_g1(W, a, n_sm, grid_limit, st, grid_out);
        case 2: return launch_em_fast_g2(W, a, n_sm, grid_limit, st, grid_out);
        case 3: return launch_em_fast_g3(W, a, n_sm, grid_limit, st, grid_out);
        case 4: return launch_em_fast_g4(W, a, n_sm, grid_limit, st, grid_out);
        case 5: return launch_em_fast_g5(W, a, n_sm, grid_limit, st, grid_out);
        case 6: return launch_em_fast_g6(W, a, n_sm, grid_limit, st, grid_out);
        case 7: return launch_em_fast_g7(W, a, n_sm, grid_limit, st, grid_out);
        case 8: return launch_em_fast_g8(W, a, n_sm, grid_limit, st, grid_out);
        case 9: return launch_em_fast_g9(W, a, n_sm, grid_limit, st, grid_out);
        case 10: return launch_em_fast_g10(W, a, n_sm, grid_limit, st, grid_out);
        case 11: return launch_em_fast_g11(W, a, n_sm, grid_limit, st, grid_out);
        case 12: return launch_em_fast_g12(W, a, n_sm, grid_limit, st, grid_out);
        default: return cudaErrorInvalidValue;
    }
}

bool em_fast_available(int G, int W) { return G >= 1 && G <= kMaxFastG && (W == 1 || W == 2 || W == 4); }

int em_fast_single_warp_genes(int G, int L) {
    const long long smem1 = em_fast_smem_doubles(G, 1, L) * (long long)sizeof(double);
    long long genes = std::min<long long>((long long)kMaxDynSmem / smem1, kMaxFastCtasPerSm);
    const int cols = tmem_cols_for(G, L, 1);
    const long long team = em_fast_tmem_team_doubles(G, L) * (long long)sizeof(double);
    if (cols > 0 && team * kTmemTeams <= (long long)kMaxDynSmem) {
        const long long ctas = std::min<long long>(std::min<long long>((long long)kMaxDynSmem / (team * kTmemTeams), 512 / cols), JACKS_TM_CTAS);
        genes = std::max(genes, ctas * kTmemTeams);
    }
    return (int)genes;
}

size_t em_fast_smem_bytes(int G, int W, int L) { return (size_t)em_fast_smem_doubles(G, W, L) * sizeof(double); }

}  // namespace jacks
