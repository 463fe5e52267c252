// KDE p-values of gene effects against the pseudo-gene null.
//
// Reference: computeW1PvalsAndFDRs with pseudo=True (jacks/jacks/jacks_io.py:69-89) -- per cell line a
// scipy.stats.gaussian_kde over the non-NaN pseudo-gene w1 values and, per gene,
// kernel.integrate_box_1d(-100, w1) (negative selection) or integrate_box_1d(w1, 100) (positive).
// scipy 1.18.1 (stats/_kde.py) evaluates that as
//     h   = sqrt( cov(s, ddof=1, aweights=1/n) * neff^(-2/5) ),  neff = 1/sum(w^2) = n      (Scott's rule)
//     p   = sum_i (1/n) * [ ndtr((hi - s_i)/h) - ndtr((lo - s_i)/h) ]
// so the work is n_genes * L * n_pseudo normal-CDF evaluations (7.2e11 at the Avana shape with 1e5
// pseudo-genes): FP64-pipe bound.  The term that does not depend on the gene (lo = -100 or hi = +100) is
// evaluated once per line.
//
// Kernels:
//   kde_line_stats   one CTA per line: compacts the non-NaN pseudo values into a transposed, +inf-padded
//                    row sT[l][.], two-pass mean / variance, bandwidth, the gene-independent term.
//   kde_pvalues      grid (gene tiles, lines): a CTA stages chunks of sT[l] in shared memory; each thread
//                    owns GT genes of the line and streams the chunk through registers (one broadcast
//                    LDS feeds GT independent normcdf chains).
#include <cuda_runtime.h>
#include <math.h>
#include <stdlib.h>

#include "../../include/jacks_b200.h"
#include "capi_internal.h"

namespace jacks {

namespace {

constexpr int kStatsThreads = 256;
constexpr int kPvThreads = 128;
constexpr int kGT = 4;              // genes per thread
constexpr int kChunk = 1024;        // pseudo values staged per iteration (8 KB)

__device__ __forceinline__ double block_sum(double v, double* sh) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sh[w];
    return s;
}

// stats[l*8 + 0] = h, +1 = n (as double), +2 = gene-independent term, +3 = 1/n, +4 = mean, +5 = sum of squared deviations
__global__ void __launch_bounds__(kStatsThreads) kde_line_stats(const double* __restrict__ w1_pseudo, long long n_pseudo,
                                                                int L, long long ld, int positive,
                                                                double* __restrict__ sT, double* __restrict__ stats) {
    __shared__ double sh[kStatsThreads / 32];
    __shared__ int s_base;
    __shared__ int s_cnt[kStatsThreads / 32];
    // Positive selection is computed as negative selection of the mirrored problem (s -> -s, w -> -w):
    // Phi((100-s)/h) - Phi((w-s)/h) = Phi((-w+s)/h) - Phi((-100+s)/h), which also avoids the 1 - Phi cancellation.
    const double sgn = positive ? -1.0 : 1.0;
    const int l = blockIdx.x;
    double* row = sT + (long long)l * ld;
    // Order-preserving compaction (a block-wide scan of the per-warp counts, no atomics): the row order fixes the order
    // of the sums below, so mean, bandwidth and the gene-independent term are bit-reproducible from run to run and do
    // not depend on which genes are evaluated with them.
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    double sum = 0.0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long long i0 = 0; i0 < n_pseudo; i0 += kStatsThreads) {
        const long long i = i0 + threadIdx.x;
        double v = 0.0;
        bool ok = false;
        if (i < n_pseudo) { v = w1_pseudo[i * L + l]; ok = (v == v); }      // jacks_io.py:78 drops NaN
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) s_cnt[warp] = __popc(m);
        __syncthreads();
        int wbase = s_base, total = 0;
        for (int w = 0; w < kStatsThreads / 32; ++w) { const int c = s_cnt[w]; if (w < warp) wbase += c; total += c; }
        if (ok) { v *= sgn; row[wbase + __popc(m & ((1u << lane) - 1))] = v; sum += v; }
        __syncthreads();
        if (threadIdx.x == 0) s_base += total;
        __syncthreads();
    }
    const int n = s_base;
    for (long long i = n + threadIdx.x; i < ld; i += kStatsThreads) row[i] = __longlong_as_double(0x7ff0000000000000LL);
    const double mean = block_sum(sum, sh) / (double)n;
    double ss = 0.0;
    for (int i = threadIdx.x; i < n; i += kStatsThreads) { const double d = row[i] - mean; ss += d * d; }
    ss = block_sum(ss, sh);
    // cov with aweights = 1/n: sum(w d^2) / (1 - sum(w^2)) = (ss/n) / (1 - 1/n); neff = n
    const double wgt = 1.0 / (double)n;
    const double cov = (ss * wgt) / (1.0 - wgt);
    const double factor = pow((double)n, -0.2);
    const double h = sqrt(cov * factor * factor);
    double c = 0.0;
    for (int i = threadIdx.x; i < n; i += kStatsThreads) c += normcdf((-100.0 - row[i]) / h);
    c = block_sum(c, sh);
    if (threadIdx.x == 0) {
        stats[l * 8 + 0] = h;
        stats[l * 8 + 1] = (double)n;
        stats[l * 8 + 2] = c;
        stats[l * 8 + 3] = wgt;
        stats[l * 8 + 4] = mean;
        stats[l * 8 + 5] = ss;
    }
}

__global__ void __launch_bounds__(kPvThreads) kde_pvalues(const double* __restrict__ w1_genes, long long n_genes, int L,
                                                          const double* __restrict__ sT, long long ld,
                                                          const double* __restrict__ stats, int positive,
                                                          const long long* __restrict__ cell_off, double* __restrict__ pvals) {
    __shared__ double chunk[kChunk];
    const int l = blockIdx.y;
    if (cell_off && cell_off[l + 1] > cell_off[l]) return;        // this line was done by the fast path
    const double h = stats[l * 8 + 0];
    const int n = (int)stats[l * 8 + 1];
    const double cterm = stats[l * 8 + 2];
    const double wgt = stats[l * 8 + 3];
    const double rh = 1.0 / h;
    const double sgn = positive ? -1.0 : 1.0;          // mirrored problem for positive selection (see kde_line_stats)
    const long long g0 = ((long long)blockIdx.x * kPvThreads + threadIdx.x) * kGT;
    double w[kGT], acc[kGT];
#pragma unroll
    for (int k = 0; k < kGT; ++k) {
        const long long g = g0 + k;
        w[k] = (g < n_genes) ? sgn * w1_genes[g * L + l] : 0.0;
        acc[k] = 0.0;
    }
    const double* row = sT + (long long)l * ld;
    for (int base = 0; base < n; base += kChunk) {
        __syncthreads();
        for (int i = threadIdx.x; i < kChunk; i += kPvThreads) chunk[i] = row[base + i];   // row is padded to a multiple of kChunk with +inf
        __syncthreads();
        const int m = min(kChunk, n - base);
        for (int i = 0; i < m; ++i) {
            const double s = chunk[i];
#pragma unroll
            for (int k = 0; k < kGT; ++k) acc[k] += normcdf((w[k] - s) * rh);
        }
    }
#pragma unroll
    for (int k = 0; k < kGT; ++k) {
        const long long g = g0 + k;
        if (g < n_genes) {
            // negative: mean_i[Phi((w-s)/h) - Phi((-100-s)/h)]
            // positive: mean_i[Phi((100-s)/h) - Phi((w-s)/h)] = mean_i[Phi(-(w-s)/h) - Phi(-(100-s)/h)]
            pvals[g * L + l] = (acc[k] - cterm) * wgt;
        }
    }
}

// ---- fast path: sorted null values, cells of width h/4, truncated Taylor (Hermite) expansion per cell -------
// F(w) = sum_i Phi((w - s_i)/h).  Sort the null values of a line; cut the axis into cells of width delta = h/4
// with centres c_j; a point of cell j is s_i = c_j + h*u_i, |u_i| <= 1/8.  With z = (w - c_j)/h
//     sum_{i in j} Phi(z - u_i) = m_0 Phi(z) - phi(z) * sum_{k=1..K} He_{k-1}(z) M_k,    M_k = sum_i u_i^k / k!
// (Taylor in u; Phi^(k) = (-1)^(k-1) He_{k-1} phi).  Cells whose points all have z >= 8.3 contribute their count
// (Phi = 1 in double), cells below z_lo = -sqrt(min(z_max,0)^2 + 2 ln(1e16 n)) cannot change the sum at 1e-16
// relative, and when even the nearest point has z_max < -7.5 the few contributing points are summed directly.
// K = 16 keeps the truncation below 1e-14 relative for every |z| in the window (numpy prototype: 5e-15 worst case
// over bulk, tails and edges at n = 2e3 and 1e5).  Work per p-value drops from n normal CDFs to ~45 cells.
constexpr int kK = 16;                  // expansion order
constexpr int kCellRec = kK + 2;        // doubles per cell: M_0..M_K, index of the cell's first point
constexpr double kCellWidth = 0.25;     // in units of h
constexpr int kSortTileD = 4096;

__global__ void __launch_bounds__(1024) sort_rows_local(double* __restrict__ rows, long long ld, int k_lo, int k_hi) {
    __shared__ double sm[kSortTileD];
    const long long base = (long long)blockIdx.y * ld + (long long)blockIdx.x * kSortTileD;
    for (int t = threadIdx.x; t < kSortTileD; t += blockDim.x) sm[t] = rows[base + t];
    __syncthreads();
    const long long tile0 = (long long)blockIdx.x * kSortTileD;
    for (long long kk = k_lo; kk <= k_hi; kk <<= 1) {
        for (int j = (int)((kk >> 1) < kSortTileD ? (kk >> 1) : (kSortTileD >> 1)); j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < kSortTileD / 2; t += blockDim.x) {
                const int a = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int b = a | j;
                const bool up = (((tile0 + a) & kk) == 0);
                const double va = sm[a], vb = sm[b];
                if ((vb < va) == up) { sm[a] = vb; sm[b] = va; }
            }
            __syncthreads();
        }
    }
    for (int t = threadIdx.x; t < kSortTileD; t += blockDim.x) rows[base + t] = sm[t];
}

__global__ void __launch_bounds__(256) sort_rows_global(double* __restrict__ rows, long long ld, long long kk, long long j) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ld / 2) return;
    double* r = rows + (long long)blockIdx.y * ld;
    const long long a = ((t & ~(j - 1)) << 1) | (t & (j - 1));
    const long long b = a | j;
    const bool up = ((a & kk) == 0);
    const double va = r[a], vb = r[b];
    if ((vb < va) == up) { r[a] = vb; r[b] = va; }
}

// cell_off[l] .. cell_off[l+1]: cells of line l in the cell table; a line whose cells do not fit is marked
// (cell_off[l+1] == cell_off[l]) and evaluated by the brute-force kernel.  One thread, L is small.
__global__ void kde_cell_layout(const double* __restrict__ sT, long long ld, const double* __restrict__ stats, int L,
                                long long cell_cap, long long* __restrict__ cell_off) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    long long off = 0;
    cell_off[0] = 0;
    for (int l = 0; l < L; ++l) {
        const double h = stats[l * 8 + 0];
        const int n = (int)stats[l * 8 + 1];
        long long nc = 0;
        if (n >= 2 && h > 0.0 && h == h) {
            const double* row = sT + (long long)l * ld;
            const double span = (row[n - 1] - row[0]) / (h * kCellWidth);
            if (span == span && span < 1.0e7) nc = (long long)floor(span) + 1;
        }
        if (nc > 0 && off + nc <= cell_cap) off += nc;        // else: no cells -> brute force for this line
        cell_off[l + 1] = off;
    }
}

__device__ __forceinline__ long long kde_cell_of(double s, double smin, double inv_delta, long long nc) {
    long long j = (long long)floor((s - smin) * inv_delta);
    return j < 0 ? 0 : (j >= nc ? nc - 1 : j);
}

// One thread per cell: moments of the cell's points (sorted row => the points are a contiguous run).
__global__ void __launch_bounds__(128) kde_cell_moments(const double* __restrict__ sT, long long ld,
                                                        const double* __restrict__ stats, int L,
                                                        const long long* __restrict__ cell_off, double* __restrict__ cells) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cell_off[L]) return;
    int lo_l = 0, hi_l = L;                       // line of this cell: cell_off[l] <= c < cell_off[l+1]
    while (hi_l - lo_l > 1) { const int mid = (lo_l + hi_l) >> 1; if (cell_off[mid] <= c) lo_l = mid; else hi_l = mid; }
    const int l = lo_l;
    const long long j = c - cell_off[l], nc = cell_off[l + 1] - cell_off[l];
    const double h = stats[l * 8 + 0];
    const int n = (int)stats[l * 8 + 1];
    const double* row = sT + (long long)l * ld;
    const double smin = row[0];
    const double inv_delta = 1.0 / (h * kCellWidth);
    // first point with cell index >= j, first point with cell index > j (cell index is monotone along the row)
    int a = 0, b = n;
    while (a < b) { const int mid = (a + b) >> 1; if (kde_cell_of(row[mid], smin, inv_delta, nc) < j) a = mid + 1; else b = mid; }
    const int first = a;
    b = n;
    while (a < b) { const int mid = (a + b) >> 1; if (kde_cell_of(row[mid], smin, inv_delta, nc) <= j) a = mid + 1; else b = mid; }
    const int last = a;
    const double centre = smin + ((double)j + 0.5) * (h * kCellWidth);
    const double rh = 1.0 / h;
    double M[kK + 1];
#pragma unroll
    for (int k = 0; k <= kK; ++k) M[k] = 0.0;
    for (int i = first; i < last; ++i) {
        const double u = (row[i] - centre) * rh;
        double t = 1.0;
        M[0] += 1.0;
#pragma unroll
        for (int k = 1; k <= kK; ++k) { t *= u * (1.0 / k); M[k] += t; }       // u^k / k!
    }
    double* rec = cells + c * kCellRec;
#pragma unroll
    for (int k = 0; k <= kK; ++k) rec[k] = M[k];
    rec[kK + 1] = (double)first;
}

// One CTA evaluates kFastGenes genes of one line.  The line's cell table (a few hundred 144-byte records that every
// gene of the line walks a ~45-cell window of) is staged in shared memory first: read from L2 by every thread it was
// 47 GB per Avana-shaped call, i.e. the kernel ran at L2 bandwidth.  A table too large for the CTA's shared memory
// is read in place.
constexpr int kFastThreads = 256;
constexpr int kFastPerThread = 4;
constexpr int kFastGenes = kFastThreads * kFastPerThread;
constexpr int kFastSmemCells = 480;          // 67.5 KB: three CTAs per SM

__global__ void __launch_bounds__(kFastThreads) kde_pvalues_fast(const double* __restrict__ w1_genes, long long n_genes, int L,
                                                                 const double* __restrict__ sT, long long ld,
                                                                 const double* __restrict__ stats, const long long* __restrict__ cell_off,
                                                                 const double* __restrict__ cells, int positive,
                                                                 double* __restrict__ pvals) {
    extern __shared__ double s_tab[];
    const int l = blockIdx.y;
    const long long nc = cell_off[l + 1] - cell_off[l];
    if (nc == 0) return;                                   // this line is left to the brute-force kernel
    const double* ctab = cells + cell_off[l] * kCellRec;
    if (nc <= kFastSmemCells) {
        for (long long i = threadIdx.x; i < nc * kCellRec; i += kFastThreads) s_tab[i] = ctab[i];
        __syncthreads();
        ctab = s_tab;
    }
    const double h = stats[l * 8 + 0];
    const int n = (int)stats[l * 8 + 1];
    const double cterm = stats[l * 8 + 2], wgt = stats[l * 8 + 3];
    const double* row = sT + (long long)l * ld;
  for (int k = 0; k < kFastPerThread; ++k) {
    const long long g = ((long long)blockIdx.x * kFastPerThread + k) * kFastThreads + threadIdx.x;
    if (g >= n_genes) break;
    const double w = positive ? -w1_genes[g * L + l] : w1_genes[g * L + l];
    if (!(w == w)) { pvals[g * L + l] = w; continue; }
    const double rh = 1.0 / h;
    const double smin = row[0];
    const double zmax = (w - smin) * rh;
    double tot = 0.0;
    if (zmax < -7.5) {
        // far below every null value: only the nearest few points contribute
        for (int i = 0; i < n; ++i) {
            const double term = normcdf((w - row[i]) * rh);
            tot += term;
            if (term <= 1e-18 * tot) break;          // also stops at the first exact zero (z < -38.5): later points are farther
        }
    } else {
        const double zneg = fmin(zmax, 0.0);
        const double zlo = -sqrt(zneg * zneg + 2.0 * log(1e16 * (double)n));
        const double delta = h * kCellWidth;
        // z_j = (w - centre_j)/h = zmax - (j + 0.5)*kCellWidth decreases with j
        // saturated: z_j - kCellWidth/2 >= 8.3  <=>  j <= (zmax - 8.3)/kCellWidth - 1
        long long jsat = (long long)floor((zmax - 8.3) / kCellWidth) - 1;
        if (jsat >= nc) jsat = nc - 1;
        if (jsat >= 0) tot = (jsat + 1 < nc) ? ctab[(jsat + 1) * kCellRec + kK + 1] : (double)n;    // points before cell jsat+1
        long long jend = (long long)ceil((zmax - zlo) / kCellWidth);                                 // z_j + w/2 >= zlo
        if (jend >= nc) jend = nc - 1;
        (void)delta;
        for (long long j = (jsat < 0 ? 0 : jsat + 1); j <= jend; ++j) {
            const double* rec = ctab + j * kCellRec;
            const double m0 = rec[0];
            if (m0 == 0.0) continue;
            const double z = zmax - ((double)j + 0.5) * kCellWidth;
            double he_prev = 0.0, he = 1.0, S = 0.0;
#pragma unroll
            for (int k = 1; k <= kK; ++k) {
                S = fma(he, rec[k], S);
                const double nxt = fma(z, he, -(double)(k - 1) * he_prev);
                he_prev = he; he = nxt;
            }
            const double phi = exp(-0.5 * z * z) * 0.3989422804014327;
            tot += m0 * normcdf(z) - phi * S;
        }
    }
    pvals[g * L + l] = (tot - cterm) * wgt;
  }
}

// General KDE evaluation for the branches the command line does not reach (jacks_io.py:76-88):
//   kind 0/1: cdf as in kde_pvalues;  kind 2: density  sum_i exp(-z^2/2) / (n h sqrt(2 pi))  (kernel.evaluate)
//   loo != 0: every evaluated value is itself one of the null points and is left out of its own kernel
//   (jacks_io.py:81-82): n' = n-1, the bandwidth comes from the downdated mean / variance, the point's own term
//   (Phi(0) = 1/2, or exp(0) = 1) is subtracted, and the gene-independent cdf term is re-evaluated per gene.
__global__ void __launch_bounds__(kPvThreads) kde_eval_kernel(const double* __restrict__ w_eval, long long m, int L,
                                                              const double* __restrict__ sT, long long ld,
                                                              const double* __restrict__ stats, int kind, int loo,
                                                              double* __restrict__ out) {
    __shared__ double chunk[kChunk];
    const int l = blockIdx.y;
    const int n = (int)stats[l * 8 + 1];
    const double mean = stats[l * 8 + 4], ss = stats[l * 8 + 5];
    const long long g = (long long)blockIdx.x * kPvThreads + threadIdx.x;
    const double w = (g < m) ? ((kind == 1) ? -w_eval[g * L + l] : w_eval[g * L + l]) : 0.0;
    double h = stats[l * 8 + 0];
    double cnt = (double)n;
    if (loo) {
        // remove the point w from (n, mean, ss): ss' = ss - (w-mean)^2 * n/(n-1)
        cnt = (double)(n - 1);
        const double d = w - mean;
        const double ss1 = ss - d * d * (double)n / cnt;
        const double wg = 1.0 / cnt;
        const double cov = (ss1 * wg) / (1.0 - wg);
        const double f = pow(cnt, -0.2);
        h = sqrt(cov * f * f);
    }
    const double rh = 1.0 / h;
    const double edge = -100.0;
    double acc = 0.0, cacc = 0.0;
    const double* row = sT + (long long)l * ld;
    for (int base = 0; base < n; base += kChunk) {
        __syncthreads();
        for (int i = threadIdx.x; i < kChunk; i += kPvThreads) chunk[i] = row[base + i];
        __syncthreads();
        const int mm = min(kChunk, n - base);
        if (kind == 2) {
            for (int i = 0; i < mm; ++i) { const double z = (w - chunk[i]) * rh; acc += exp(-0.5 * z * z); }
        } else {
            for (int i = 0; i < mm; ++i) {
                acc += normcdf((w - chunk[i]) * rh);
                cacc += normcdf((edge - chunk[i]) * rh);
            }
        }
    }
    if (g < m) {
        double v;
        if (kind == 2) {
            if (loo) acc -= 1.0;
            v = acc / (cnt * h * 2.5066282746310002);          // sqrt(2 pi)
        } else {
            if (loo) { acc -= 0.5; cacc -= normcdf((edge - w) * rh); }
            v = (acc - cacc) / cnt;
        }
        out[g * L + l] = v;
    }
}

}  // namespace
}  // namespace jacks

using namespace jacks;

// Null density of every line from the pseudo-gene effects (jacks_io.py:76-79): compaction, bandwidth, sorted rows and
// the cell table, all in ctx->tmp.  `out` stays valid until the next KDE call on this context.
int jacks::kde_null_prepare(JacksCtx* ctx, const double* w1_pseudo, int64_t n_pseudo, int32_t n_lines, int32_t positive,
                            cudaStream_t st, KdeNull* out) {
    // rows padded to a power of two (>= one sort tile) so the bitonic network needs no bounds checks
    long long ld = kSortTileD;
    while (ld < n_pseudo) ld <<= 1;
    const long long cell_cap = (long long)n_lines * 2048;
    const size_t bytes = (size_t)ld * n_lines * sizeof(double) + (size_t)n_lines * 8 * sizeof(double) +
                         (size_t)(n_lines + 1) * sizeof(long long) + (size_t)cell_cap * kCellRec * sizeof(double);
    int rc = ctx->tmp.reserve(bytes);
    if (rc != JACKS_OK) return rc;
    cudaError_t e;
    double* sT = (double*)ctx->tmp.p;
    double* stats = sT + ld * n_lines;
    long long* cell_off = (long long*)(stats + (size_t)n_lines * 8);
    double* cells = (double*)(cell_off + n_lines + 1);
    const bool brute = getenv("JACKS_KDE_BRUTE") != nullptr;          // validation: force the O(n) sums
    kde_line_stats<<<n_lines, kStatsThreads, 0, st>>>(w1_pseudo, n_pseudo, n_lines, ld, positive ? 1 : 0, sT, stats);
    ctx->launches++;
    if (!brute) {
        dim3 lg((unsigned)(ld / kSortTileD), (unsigned)n_lines);
        sort_rows_local<<<lg, 1024, 0, st>>>(sT, ld, 2, kSortTileD);
        ctx->launches++;
        for (long long kk = 2LL * kSortTileD; kk <= ld; kk <<= 1) {
            for (long long j = kk >> 1; j >= kSortTileD; j >>= 1) {
                dim3 gg((unsigned)((ld / 2 + 255) / 256), (unsigned)n_lines);
                sort_rows_global<<<gg, 256, 0, st>>>(sT, ld, kk, j);
                ctx->launches++;
            }
            sort_rows_local<<<lg, 1024, 0, st>>>(sT, ld, (int)kk, (int)kk);
            ctx->launches++;
        }
        kde_cell_layout<<<1, 1, 0, st>>>(sT, ld, stats, n_lines, cell_cap, cell_off);
        kde_cell_moments<<<(unsigned)((cell_cap + 127) / 128), 128, 0, st>>>(sT, ld, stats, n_lines, cell_off, cells);
        ctx->launches += 2;
    }
    if ((e = cudaGetLastError()) != cudaSuccess) return cuda_fail(e, "kde null kernels launch");
    out->sT = sT; out->stats = stats; out->cell_off = cell_off; out->cells = cells;
    out->ld = ld; out->n_lines = n_lines; out->positive = positive ? 1 : 0; out->brute = brute ? 1 : 0;
    return JACKS_OK;
}

// p-values of `n_genes` rows of w1_genes [n_genes, n_lines] under a prepared null.
int jacks::kde_eval(JacksCtx* ctx, const KdeNull& nl, const double* w1_genes, int64_t n_genes, double* pvals, cudaStream_t st) {
    if (n_genes == 0) return JACKS_OK;
    cudaError_t e;
    const int n_lines = nl.n_lines;
    if (!nl.brute) {
        dim3 fg((unsigned)((n_genes + kFastGenes - 1) / kFastGenes), (unsigned)n_lines);
        const size_t fsm = (size_t)kFastSmemCells * kCellRec * sizeof(double);
        if (!ctx->kde_attr_set) {          // per device: one process may hold contexts on several
            if ((e = cudaFuncSetAttribute(kde_pvalues_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsm)) != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute");
            ctx->kde_attr_set = true;
        }
        kde_pvalues_fast<<<fg, kFastThreads, fsm, st>>>(w1_genes, n_genes, n_lines, nl.sT, nl.ld, nl.stats, nl.cell_off, nl.cells, nl.positive, pvals);
        ctx->launches++;
    }
    // lines the cell table could not hold (pathological ranges), or everything when forced
    const long long per_cta = (long long)kPvThreads * kGT;
    dim3 grid((unsigned)((n_genes + per_cta - 1) / per_cta), (unsigned)n_lines);
    kde_pvalues<<<grid, kPvThreads, 0, st>>>(w1_genes, n_genes, n_lines, nl.sT, nl.ld, nl.stats, nl.positive,
                                             nl.brute ? nullptr : nl.cell_off, pvals);
    if ((e = cudaGetLastError()) != cudaSuccess) return cuda_fail(e, "kde kernels launch");
    ctx->launches++;
    return JACKS_OK;
}

extern "C" int jacks_kde_pvalues_device(JacksCtx* ctx, const double* w1_genes, int64_t n_genes, const double* w1_pseudo,
                                        int64_t n_pseudo, int32_t n_lines, int32_t positive, double* pvals, void* stream) {
    if (!ctx || !pvals || (n_genes > 0 && !w1_genes) || !w1_pseudo) return fail(JACKS_E_INVALID, "jacks_kde_pvalues_device: NULL argument");
    if (n_genes < 0 || n_pseudo < 2 || n_lines < 1) return fail(JACKS_E_INVALID, "need n_pseudo >= 2 and n_lines >= 1 (got %lld, %d)", (long long)n_pseudo, n_lines);
    if (n_genes == 0) return JACKS_OK;
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    KdeNull nl;
    int rc = kde_null_prepare(ctx, w1_pseudo, n_pseudo, n_lines, positive, (cudaStream_t)stream, &nl);
    if (rc != JACKS_OK) return rc;
    return kde_eval(ctx, nl, w1_genes, n_genes, pvals, (cudaStream_t)stream);
}

extern "C" int jacks_kde_pvalues_host(JacksCtx* ctx, const double* w1_genes, int64_t n_genes, const double* w1_pseudo,
                                      int64_t n_pseudo, int32_t n_lines, int32_t positive, double* pvals) {
    if (!ctx || !pvals || (n_genes > 0 && !w1_genes) || !w1_pseudo) return fail(JACKS_E_INVALID, "jacks_kde_pvalues_host: NULL argument");
    if (n_genes < 0 || n_pseudo < 2 || n_lines < 1) return fail(JACKS_E_INVALID, "need n_pseudo >= 2 and n_lines >= 1 (got %lld, %d)", (long long)n_pseudo, n_lines);
    if (n_genes == 0) return JACKS_OK;
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    const size_t gb = (size_t)n_genes * n_lines * sizeof(double), pb = (size_t)n_pseudo * n_lines * sizeof(double);
    int rc;
    if ((rc = ctx->out[0].reserve(gb)) != JACKS_OK) return rc;
    if ((rc = ctx->out[1].reserve(pb)) != JACKS_OK) return rc;
    if ((rc = ctx->out[2].reserve(gb)) != JACKS_OK) return rc;
    cudaStream_t st = ctx->stream;
    if ((rc = ctx->stager.h2d(ctx->out[0].p, w1_genes, gb, st)) != JACKS_OK) return rc;
    if ((rc = ctx->stager.h2d(ctx->out[1].p, w1_pseudo, pb, st)) != JACKS_OK) return rc;
    rc = jacks_kde_pvalues_device(ctx, (const double*)ctx->out[0].p, n_genes, (const double*)ctx->out[1].p, n_pseudo,
                                  n_lines, positive, (double*)ctx->out[2].p, st);
    if (rc != JACKS_OK) return rc;
    if ((rc = ctx->stager.d2h(pvals, ctx->out[2].p, gb, st)) != JACKS_OK) return rc;
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return cuda_fail(e, "kde kernels");
    return JACKS_OK;
}

extern "C" int jacks_kde_eval_host(JacksCtx* ctx, const double* w_eval, int64_t n_eval, const double* w_null, int64_t n_null,
                                   int32_t n_lines, int32_t kind, int32_t leave_one_out, double* out) {
    if (!ctx || !out || (n_eval > 0 && !w_eval) || !w_null) return fail(JACKS_E_INVALID, "jacks_kde_eval_host: NULL argument");
    if (n_eval < 0 || n_null < (leave_one_out ? 3 : 2) || n_lines < 1) return fail(JACKS_E_INVALID, "too few null points (%lld)", (long long)n_null);
    if (kind < 0 || kind > 2) return fail(JACKS_E_INVALID, "kind must be 0 (cdf, negative), 1 (cdf, positive) or 2 (density)");
    if (n_eval == 0) return JACKS_OK;
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    const size_t gb = (size_t)n_eval * n_lines * sizeof(double), pb = (size_t)n_null * n_lines * sizeof(double);
    int rc;
    if ((rc = ctx->out[0].reserve(gb)) != JACKS_OK) return rc;
    if ((rc = ctx->out[1].reserve(pb)) != JACKS_OK) return rc;
    if ((rc = ctx->out[2].reserve(gb)) != JACKS_OK) return rc;
    const long long ld = ((long long)n_null + kChunk - 1) / kChunk * kChunk;
    if ((rc = ctx->tmp.reserve((size_t)ld * n_lines * sizeof(double) + (size_t)n_lines * 8 * sizeof(double))) != JACKS_OK) return rc;
    cudaStream_t st = ctx->stream;
    if ((rc = ctx->stager.h2d(ctx->out[0].p, w_eval, gb, st)) != JACKS_OK) return rc;
    if ((rc = ctx->stager.h2d(ctx->out[1].p, w_null, pb, st)) != JACKS_OK) return rc;
    double* sT = (double*)ctx->tmp.p;
    double* stats = sT + ld * n_lines;
    kde_line_stats<<<n_lines, kStatsThreads, 0, st>>>((const double*)ctx->out[1].p, n_null, n_lines, ld, kind == 1 ? 1 : 0, sT, stats);
    dim3 grid((unsigned)((n_eval + kPvThreads - 1) / kPvThreads), (unsigned)n_lines);
    kde_eval_kernel<<<grid, kPvThreads, 0, st>>>((const double*)ctx->out[0].p, n_eval, n_lines, sT, ld, stats, kind,
                                                 leave_one_out ? 1 : 0, (double*)ctx->out[2].p);
    ctx->launches += 2;
    if ((e = cudaGetLastError()) != cudaSuccess) return cuda_fail(e, "kde_eval launch");
    if ((rc = ctx->stager.d2h(out, ctx->out[2].p, gb, st)) != JACKS_OK) return rc;
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return cuda_fail(e, "kde kernels");
    return JACKS_OK;
}
