// Instantiations of the fast EM kernel for one guide count (compile with -DJACKS_G=<1..8>).
// One translation unit per G keeps the build parallel; em_dispatch.cu switches over them.
#include "em_fast.cuh"

#ifndef JACKS_G
#error "compile with -DJACKS_G=<guides per gene>"
#endif

namespace jacks {

namespace {

// Register estimate of an instantiation: tau (KL*G) + w (2*KL) + x (2G) + sums (2G+1) + y strip (G)
// doubles, two 32-bit registers each, plus addressing / temporaries.
constexpr int est_regs(int G, int KL) { return 2 * (KL * G + 2 * KL + 5 * G + 1) + 80; }

constexpr int min_blocks(int G, int KL, int W) {
    int by_regs = 65536 / (W * 32 * est_regs(G, KL));
    int by_smem = (int)((227 * 1024) / (2 * G * KL * W * 32 * 8 + 1024));
    int by_warps = 64 / W;
    int m = by_regs < by_smem ? by_regs : by_smem;
    m = m < by_warps ? m : by_warps;
    m = m < 32 ? m : 32;
    return m < 1 ? 1 : m;
}

template <int G, int KL, int W>
struct Inst {
    static constexpr int MINB = min_blocks(G, KL, W);
    static auto kernel() { return em_fast_kernel<G, KL, W, MINB>; }
    static cudaError_t prepare(int* occ) {
        auto k = kernel();
        const size_t smem = em_fast_smem_bytes<G, KL, W>();
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, k, W * 32, smem);
    }
    static cudaError_t launch(const EmLaunch& a, int n_sm, int grid_limit, cudaStream_t st, int* grid_out) {
        int occ = 0;
        cudaError_t e = prepare(&occ);
        if (e != cudaSuccess) return e;
        if (occ < 1) return cudaErrorLaunchOutOfResources;
        long long grid = (long long)occ * n_sm;
        if (grid > a.n_work) grid = a.n_work;
        if (grid_limit > 0 && grid > grid_limit) grid = grid_limit;
        if (grid < 1) grid = 1;
        if (grid_out) *grid_out = (int)grid;
        kernel()<<<(int)grid, W * 32, em_fast_smem_bytes<G, KL, W>(), st>>>(a);
        return cudaGetLastError();
    }
};

}  // namespace

#define JACKS_CAT2(a, b) a##b
#define JACKS_CAT(a, b) JACKS_CAT2(a, b)

// Shapes compiled for every G: (KL, W) = (1,1) L<=32, (4,1) L<=128, (4,4) L<=512, (7,2) L<=448;
// (13,1) L<=416 only while its G*13 tau doubles still fit the register file (G <= 4).
#define JACKS_FOR_SHAPES(X) \
    X(1, 1) X(4, 1) X(4, 4) X(7, 2) JACKS_SHAPE_13(X)
#if JACKS_G <= 4
#define JACKS_SHAPE_13(X) X(13, 1)
#else
#define JACKS_SHAPE_13(X)
#endif

cudaError_t JACKS_CAT(launch_em_fast_g, JACKS_G)(int KL, int W, const EmLaunch& a, int n_sm, int grid_limit,
                                                 cudaStream_t st, int* grid_out) {
#define X(kl, w) if (KL == kl && W == w) return Inst<JACKS_G, kl, w>::launch(a, n_sm, grid_limit, st, grid_out);
    JACKS_FOR_SHAPES(X)
#undef X
    return cudaErrorInvalidValue;
}

int JACKS_CAT(em_fast_occupancy_g, JACKS_G)(int KL, int W) {
    int occ = 0;
#define X(kl, w) if (KL == kl && W == w) { if (Inst<JACKS_G, kl, w>::prepare(&occ) != cudaSuccess) return 0; return occ; }
    JACKS_FOR_SHAPES(X)
#undef X
    return 0;
}

bool JACKS_CAT(em_fast_available_g, JACKS_G)(int KL, int W) {
#define X(kl, w) if (KL == kl && W == w) return true;
    JACKS_FOR_SHAPES(X)
#undef X
    return false;
}

}  // namespace jacks
