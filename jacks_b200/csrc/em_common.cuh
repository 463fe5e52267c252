// Shared declarations of the batched EM kernels (host dispatcher <-> kernel translation units).
//
// The EM being computed is the per-gene variational loop of the reference,
// jacks/jacks/infer.py:55-125 (inferJACKSGene, updateX, updateW, updateTau, lowerBound).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace jacks {

// Everything one EM launch needs.  Passed by value as a kernel argument.
struct EmLaunch {
    // inputs
    const double2* test;          // [N][L]           (mean, sd)
    const double2* ctrl;          // [N][ctrl_lines]  (mean, sd)
    int L;
    int ctrl_lines;               // L or 1
    int ctrl_rowmean_fill;        // 1-D control NaN-sd rule (infer.py:63)
    const long long* gene_offsets;  // [n_genes+1] CSR
    const int* guide_rows;        // [sum G]
    const int* gene_ids;          // [n_work] genes of this bucket (indices into the CSR)
    const long long* slot_off;    // [n_work] gene_offsets[gene_ids[slot]]      (fast kernels)
    const int* slot_rows;         // [n_work][G] guide rows of the bucket's genes (fast kernels)
    int n_work;
    const int* n_work_dev;        // if non-null the work count is read from here (deferred queue)
    const double* fixed_x1;       // by row, or null
    const double* fixed_x2;
    // outputs (nullable except w1)
    double* x1; double* x2;       // [sum G] CSR order
    double* w1; double* w2;       // [n_genes][L]
    double* y;  double* tau;      // [sum G][L]
    int* n_iters; double* bound;  // [n_genes]
    // work distribution
    int* work_counter;            // zeroed before the launch; teams pull bucket slots from it
    int* deferred_ids;            // fast kernels push genes holding NaN here ...
    int* deferred_count;          // ... and count them here; the generic kernel drains the list
    // hyper-parameters (infer.py:55 defaults)
    int n_iter; int apply_w_hp;
    double tol, mu0_x, var0_x, mu0_w, var0_w, tps;
    // fast kernels: per-CTA slab [G][tpd_stride] for 2*tau_pr_den (global memory, stays L2 resident);
    // tpd_stride = em_fast_tpd_stride(L): rows padded so that look-ahead loads need no bounds logic
    double* tpd_scratch; int tpd_stride;
    // TMEM variant of the fast kernel: columns each CTA allocates (tmem_cols_for(G, L)); 0 selects the shared-memory variant
    int tmem_cols; int tmem_pd_off;        // tmem_pd_off: first column of tau_pr_den when it lives in TMEM too
    // generic kernel only: global scratch when a gene's working set exceeds shared memory
    double* scratch; long long scratch_stride;   // doubles per warp
    // Per-line mode (every (gene, line) pair an independent EM with one line; run_JACKS_single.py:83-85).
    // The cubes keep their L_full lines per row; outputs gain a line axis: w1, w2, n_iters, bound
    // [n_genes][L_full]; x1, x2, y, tau [sum G][L_full].  lines_mode 0: off (joint run over all lines);
    // 1: work item w is (bucket slot w / L_full, line w % L_full); 2: gene_ids[] holds gene*L_full + line
    // (the deferred queue of the per-line kernels).  In the generic kernel L is 1 in modes 1 and 2.
    int lines_mode; int L_full;
};

// Largest guide count with compiled fast kernels (TKOv1-style libraries have 12 guides per gene).
constexpr int kMaxFastG = 12;
// Fast path: compile-time guide count G (1..kMaxFastG), W (1, 2 or 4) warps per gene, finite data only.
// Returns cudaErrorInvalidValue if (G, W) was not instantiated.
cudaError_t launch_em_fast(int G, int W, const EmLaunch& a, int n_sm, int grid_limit, cudaStream_t st,
                           int* grid_out);
// Genes one SM can hold with ONE warp per gene: the larger of the shared-memory variant's and the TMEM variants' count.
int em_fast_single_warp_genes(int G, int L);
// Is (G, W) compiled in?
bool em_fast_available(int G, int W);
// Bytes of dynamic shared memory one gene team needs (y, tau_pr_den, tau, w1, w2 + reduction buffer).
size_t em_fast_smem_bytes(int G, int W, int L);
// Largest dynamic shared memory a CTA may request on sm_100a (227 KB minus the static words we use).
constexpr size_t kMaxDynSmem = 227 * 1024 - 64;
// Row stride (doubles) of the fast kernels' tau_pr_den slab: L plus one full look-ahead step (4 lines x 128 lanes).
inline int em_fast_tpd_stride(int L) { return ((L + 15) / 16) * 16 + 512; }
// Upper bound on resident fast-kernel CTAs per SM the plan provisions tau_pr_den slabs for.
constexpr int kMaxFastCtasPerSm = 12;

// Per-line fast path: one LANE per (gene, line) problem of a G-guide bucket (G 1..8), finite data only;
// problems holding NaN/inf go to the deferred queue as gene*L + line.  `a.L` is the cubes' line count.
cudaError_t launch_em_lines(int G, const EmLaunch& a, int n_sm, cudaStream_t st);

// Generic path: any G and L, NaN-safe, one warp per gene, working set in shared memory when it
// fits and in `scratch` (EmLaunch::scratch, one slab per warp of the grid) otherwise.
struct EmGenericConfig {
    int warps_per_cta;
    int grid;
    long long smem_doubles_per_warp;
    size_t smem_bytes;
    long long scratch_doubles_per_warp;   // 0: everything fits shared memory
};
// Working-set size (doubles) of one gene in the generic kernel.
__host__ __device__ long long em_generic_work_doubles(int G, int L);
// Genes at least this big (guides x lines) run one per CTA (em_big_kernel) instead of one per warp.
constexpr long long kBigGeneElements = 4096;
cudaError_t launch_em_big(const EmLaunch& a, int grid, long long slab_doubles, cudaStream_t st);
// max_smem: shared-memory budget of a CTA (a smaller one lets the kernel's CTAs move in beside a resident fast kernel).
cudaError_t em_generic_configure(int maxG, int L, long long n_work_cap, int n_sm, EmGenericConfig* cfg, size_t max_smem = 226 * 1024);
cudaError_t launch_em_generic(const EmGenericConfig& cfg, const EmLaunch& a, cudaStream_t st);

}  // namespace jacks
