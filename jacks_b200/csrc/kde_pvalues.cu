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
                                                          int skip_fast, double* __restrict__ pvals) {
    __shared__ double chunk[kChunk];
    const int l = blockIdx.y;
    if (skip_fast && stats[l * 8 + 7] > 0.0) return;              // this line was done by the fast path
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

// ---- fast path: local expansions (a one-dimensional fast Gauss transform for the normal CDF) ------------------------
// F(w) = sum_i Phi((w - s_i)/h).  The axis is cut into cells of width h/4 from the smallest null value; cell c has its
// centre at zc(c) = (c + 1/2)/4 (in units of h from smin), a source point of it is s_i = smin + h (zc(c) + u_i), |u| <= 1/8,
// and an evaluation point in target cell j (any integer) is w = smin + h (zc(j) + t), |t| <= 1/8.  Then
//     (w - s_i)/h = (j - c)/4 + (t - u_i)                 and, Taylor in (t - u) around Z = (j - c)/4,
//     F(w) = sum_m t^m A_m(j),
//     A_m(j) = [m = 0] #{points in cells c <= j - KSAT - 1}
//              + (1/m!) sum_{k = -KNEG..KSAT} sum_{q = 0..Q} (-1)^q D_{m+q}(k/4) M_q(j - k),
//     M_q(c) = sum_{i in c} u_i^q / q!        D_0 = Phi,  D_n(Z) = (-1)^(n-1) He_{n-1}(Z) phi(Z).
// Source cells further than 8.5 bandwidths below the target contribute their count (every other term of the series is
// below 1e-16), cells more than 12 bandwidths above it cannot change the sum at 1e-16 relative, and a target more than
// 5 bandwidths below every null value sums its nearest points directly (F is then tiny and its relative accuracy rests
// on those few terms).  With P = Q = 12 the truncation is <= 6e-15 relative over bulk, edges and tails
// (tools/kde_fgt_prototype.py).  Per p-value this is one degree-12 polynomial; per line the table costs a convolution of
// ~80 source cells per target cell.  Preparing the null needs no sort: ONE stable counting-sort pass by cell index
// puts each cell's points next to each other in their original order, so every sum below runs in a fixed order and the
// p-values are bit-reproducible, independent of which genes are evaluated and of how the lines are sharded.
constexpr int kP = 12;                   // orders of the source moments and of the local series
constexpr int kNM = kP + 1;              // coefficients per cell
constexpr double kCellWidth = 0.25;      // in units of h
constexpr int kKsat = 34;                // (j - c)/4 >= 8.5: the cell counts fully
constexpr int kKneg = 48;                // (j - c)/4 <= -12: negligible
constexpr int kJmin = -20;               // targets below smin - 5 h: direct sum
constexpr int kMaxCells = 2048;          // source cells per line; a null whose range needs more keeps its far values outside the table
// Heavy-tailed nulls.  A range of more than kMaxCells cells (512 bandwidths, i.e. > 54 standard deviations) means a few
// values far from the rest: the table then covers the window mean +- kWinHalf bandwidths only and the values outside it
// -- by Chebyshev fewer than n / 27^2, in practice a handful -- are kept, in their original order, in a per-line list
// whose Phi terms every p-value of the line adds directly.  (Such lines used to fall back to the O(n) sums: 1.3 s instead
// of 2 ms for the p-values of a screen in which a third of the lines had one such pseudo-gene.)  Only a list that
// overflows kMaxOut sends its line to the O(n) kernel.
constexpr double kWinHalf = 255.0;       // in units of h: 2,040 cells
constexpr int kMaxOut = 2048;            // values outside the window kept per line
constexpr int kMaxTargets = kMaxCells + kKsat - kJmin;
constexpr int kPartChunk = 4096;         // null values per CTA of the partition pass
constexpr int kPartThreads = 256;
constexpr int kNK = kKneg + kKsat + 1;

// D_n(k/4), n = 0..2P, k = -KNEG..KSAT: filled once per context by kde_deriv_table (device arithmetic, so every rank
// of a sharded run holds identical bits).
__constant__ double c_dtab[kNK * (2 * kP + 1)];

__global__ void kde_deriv_table(double* __restrict__ tab) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= kNK) return;
    const double Z = (double)(k - kKneg) * kCellWidth;
    double* row = tab + k * (2 * kP + 1);
    row[0] = normcdf(Z);
    const double phi = exp(-0.5 * Z * Z) * 0.3989422804014327;
    double he_prev = 0.0, he = 1.0;                           // He_{-1}, He_0
    for (int n = 1; n <= 2 * kP; ++n) {
        row[n] = ((n & 1) ? 1.0 : -1.0) * he * phi;           // (-1)^(n-1) He_{n-1} phi
        const double nxt = Z * he - (double)(n - 1) * he_prev;
        he_prev = he; he = nxt;
    }
}

// Transpose [n_pseudo, L] -> rows sT[l][.] (mirrored for positive selection), coalesced both ways; counts the NaNs of
// every line.  (The first version read the matrix with stride L per thread: 0.9 ms of uncoalesced loads.)
__global__ void __launch_bounds__(256) kde_transpose(const double* __restrict__ w1_pseudo, long long n_pseudo, int L, long long ld,
                                                     int positive, double* __restrict__ sT, int* __restrict__ nan_count) {
    __shared__ double tile[32][33];
    __shared__ int s_nan[32];
    const long long i0 = (long long)blockIdx.x * 32;
    const int l0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const double sgn = positive ? -1.0 : 1.0;
    if (threadIdx.x < 32) s_nan[threadIdx.x] = 0;
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const long long i = i0 + r;
        const int l = l0 + tx;
        double v = __longlong_as_double(0x7ff8000000000000LL);
        if (i < n_pseudo && l < L) v = sgn * w1_pseudo[i * L + l];
        tile[r][tx] = v;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int l = l0 + r;
        const long long i = i0 + tx;
        if (l < L && i < n_pseudo) {
            const double v = tile[tx][r];
            sT[(long long)l * ld + i] = v;
            if (v != v) atomicAdd(&s_nan[r], 1);
        }
    }
    __syncthreads();
    if (threadIdx.x < 32 && s_nan[threadIdx.x] && l0 + (int)threadIdx.x < L) atomicAdd(&nan_count[l0 + threadIdx.x], s_nan[threadIdx.x]);
}

__device__ __forceinline__ double block_min(double v, double* sh) {
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double s = sh[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) s = fmin(s, sh[w]);
    return s;
}

// One CTA per line on its transposed row: drops NaNs in place (order preserving; rare), then mean, bandwidth, the
// gene-independent term, range and cell count.  Every sum runs over the row in a fixed order.
// stats[l*8 + 0] = h, +1 = n, +2 = gene-independent term, +3 = 1/n, +4 = mean, +5 = sum of squared deviations,
//                +6 = smallest null value, +7 = number of cells (0: the line is evaluated by the O(n) kernel)
constexpr int kRowThreads = 1024;
__global__ void __launch_bounds__(kRowThreads) kde_row_stats(double* __restrict__ sT, long long n_pseudo, long long ld,
                                                             const int* __restrict__ nan_count, double* __restrict__ stats,
                                                             double* __restrict__ win, double* __restrict__ outv) {
    constexpr int kStatsThreads = kRowThreads;           // (shadows the width of the older kde_line_stats kernel)
    __shared__ double sh[kStatsThreads / 32];
    __shared__ int s_base;
    __shared__ int s_cnt[kStatsThreads / 32];
    const int l = blockIdx.x;
    double* row = sT + (long long)l * ld;
    int n = (int)n_pseudo;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (nan_count[l] > 0) {
        if (threadIdx.x == 0) s_base = 0;
        __syncthreads();
        for (long long i0 = 0; i0 < n_pseudo; i0 += kStatsThreads) {
            const long long i = i0 + threadIdx.x;
            double v = 0.0;
            bool ok = false;
            if (i < n_pseudo) { v = row[i]; ok = (v == v); }                 // jacks_io.py:78 drops NaN
            const unsigned m = __ballot_sync(0xffffffffu, ok);
            if (lane == 0) s_cnt[warp] = __popc(m);
            __syncthreads();
            int wbase = s_base, total = 0;
            for (int w = 0; w < kStatsThreads / 32; ++w) { const int c = s_cnt[w]; if (w < warp) wbase += c; total += c; }
            if (ok) row[wbase + __popc(m & ((1u << lane) - 1))] = v;          // destination <= source: in place is safe
            __syncthreads();
            if (threadIdx.x == 0) s_base += total;
            __syncthreads();
        }
        n = s_base;
    }
    for (long long i = n + threadIdx.x; i < ld; i += kStatsThreads) row[i] = __longlong_as_double(0x7ff0000000000000LL);
    __syncthreads();
    double sum = 0.0, lo = __longlong_as_double(0x7ff0000000000000LL), hi = -lo;
    for (int i = threadIdx.x; i < n; i += kStatsThreads) { const double v = row[i]; sum += v; lo = fmin(lo, v); hi = fmax(hi, v); }
    const double mean = block_sum(sum, sh) / (double)n;
    const double smin = block_min(lo, sh);
    const double smax = -block_min(-hi, sh);
    double ss = 0.0;
    for (int i = threadIdx.x; i < n; i += kStatsThreads) { const double d = row[i] - mean; ss += d * d; }
    ss = block_sum(ss, sh);
    // cov with aweights = 1/n: sum(w d^2) / (1 - sum(w^2)) = (ss/n) / (1 - 1/n); neff = n
    const double wgt = 1.0 / (double)n;
    const double cov = (ss * wgt) / (1.0 - wgt);
    const double factor = pow((double)n, -0.2);
    const double h = sqrt(cov * factor * factor);
    // gene-independent term sum_i Phi((-100 - s_i)/h): every s_i >= smin, so when even the smallest null value puts
    // -100 more than 38.6 bandwidths away every term is exactly 0 in double (the usual case by a factor of thousands)
    double c = 0.0;
    if (!((-100.0 - smin) / h < -38.6)) {
        for (int i = threadIdx.x; i < n; i += kStatsThreads) c += normcdf((-100.0 - row[i]) / h);
        c = block_sum(c, sh);
    }
    // the cell table covers [lo, hi]: the whole range, or the window around the mean when the range needs too many cells
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    double lo_t = smin, hi_t = smax;
    bool usable = (n >= 2 && h > 0.0 && h == h), windowed = false;
    if (usable) {
        const double span = (smax - smin) / (h * kCellWidth);
        if (!(span == span)) usable = false;
        else if (!(span < (double)(kMaxCells - 1))) {
            windowed = true;
            lo_t = fmax(smin, mean - kWinHalf * h);
            hi_t = fmin(smax, mean + kWinHalf * h);
        }
    }
    int n_out = 0;
    if (usable && windowed) {
        // the values outside the window, in their original order (a fixed summation order keeps the p-values
        // bit-reproducible): block-wide ordered compaction, as for the NaNs above
        if (threadIdx.x == 0) s_base = 0;
        __syncthreads();
        double* ov = outv + (long long)l * kMaxOut;
        for (long long i0 = 0; i0 < n; i0 += kStatsThreads) {
            const long long i = i0 + threadIdx.x;
            double v = 0.0;
            bool out = false;
            if (i < n) { v = row[i]; out = (v < lo_t || v > hi_t); }
            const unsigned m = __ballot_sync(0xffffffffu, out);
            if (lane == 0) s_cnt[warp] = __popc(m);
            __syncthreads();
            int wbase = s_base, total = 0;
            for (int w = 0; w < kStatsThreads / 32; ++w) { const int cc = s_cnt[w]; if (w < warp) wbase += cc; total += cc; }
            const int at = wbase + __popc(m & ((1u << lane) - 1));
            if (out && at < kMaxOut) ov[at] = v;
            __syncthreads();
            if (threadIdx.x == 0) s_base += total;
            __syncthreads();
        }
        n_out = s_base;
        if (n_out > kMaxOut) usable = false;               // (cannot happen below 1.5 M pseudo-genes) -> the O(n) kernel
    }
    if (threadIdx.x == 0) {
        double nc = 0.0;
        if (usable) nc = floor((hi_t - lo_t) / (h * kCellWidth)) + 1.0;
        if (!(nc >= 1.0 && nc <= (double)kMaxCells)) { nc = 0.0; }
        const bool w_on = (nc > 0.0 && windowed);
        stats[l * 8 + 0] = h;
        stats[l * 8 + 1] = (double)n;
        stats[l * 8 + 2] = c;
        stats[l * 8 + 3] = wgt;
        stats[l * 8 + 4] = mean;
        stats[l * 8 + 5] = ss;
        stats[l * 8 + 6] = w_on ? lo_t : smin;             // origin of the cells
        stats[l * 8 + 7] = nc;
        win[l * 4 + 0] = w_on ? lo_t : -inf;               // values outside [win0, win1] are not in the table ...
        win[l * 4 + 1] = w_on ? hi_t : inf;
        win[l * 4 + 2] = w_on ? (double)n_out : 0.0;       // ... but in the line's list outv[l][0 .. n_out)
        win[l * 4 + 3] = 0.0;
    }
}

// Sum of the Phi terms of the null values of a line that lie outside its cell table (none for an ordinary null).
__device__ __forceinline__ double kde_outlier_sum(double w, double rh, const double* __restrict__ win, const double* __restrict__ outv, int l) {
    const int n_out = (int)win[l * 4 + 2];
    double tot = 0.0;
    const double* ov = outv + (long long)l * kMaxOut;
    for (int i = 0; i < n_out; ++i) tot += normcdf((w - ov[i]) * rh);
    return tot;
}

__device__ __forceinline__ int kde_cell_of(double s, double smin, double inv_delta, int nc) {
    const int j = (int)floor((s - smin) * inv_delta);
    return j < 0 ? 0 : (j >= nc ? nc - 1 : j);
}

// Partition pass 1: per (line, chunk) histogram of the cell indices.
__global__ void __launch_bounds__(kPartThreads) kde_cell_hist(const double* __restrict__ sT, long long ld, const double* __restrict__ stats,
                                                              const double* __restrict__ win, int nchunk, unsigned* __restrict__ hist) {
    __shared__ unsigned sh[kMaxCells];
    const int l = blockIdx.y, chunk = blockIdx.x;
    const int nc = (int)stats[l * 8 + 7];
    if (nc == 0) return;
    const int n = (int)stats[l * 8 + 1];
    const double smin = stats[l * 8 + 6];
    const double inv_delta = 1.0 / (stats[l * 8 + 0] * kCellWidth);
    for (int c = threadIdx.x; c < nc; c += kPartThreads) sh[c] = 0;
    __syncthreads();
    const double* row = sT + (long long)l * ld;
    const int i1 = min(n, (chunk + 1) * kPartChunk);
    const double wlo = win[l * 4 + 0], whi = win[l * 4 + 1];
    for (int i = chunk * kPartChunk + threadIdx.x; i < i1; i += kPartThreads) {
        const double v = row[i];
        if (v < wlo || v > whi) continue;                 // outside the table: in the line's list (kde_row_stats)
        atomicAdd(&sh[kde_cell_of(v, smin, inv_delta, nc)], 1u);
    }
    __syncthreads();
    unsigned* out = hist + ((long long)l * nchunk + chunk) * kMaxCells;
    for (int c = threadIdx.x; c < nc; c += kPartThreads) out[c] = sh[c];
}

// Partition pass 2 (one CTA per line): cell_first[c] = points in cells below c; hist[chunk][c] becomes the position of
// the first point of (chunk, c) in the partitioned row.
__global__ void __launch_bounds__(1024) kde_cell_scan(const double* __restrict__ stats, const double* __restrict__ win, int nchunk,
                                                      unsigned* __restrict__ hist, int* __restrict__ cell_first) {
    __shared__ unsigned tot[kMaxCells + 1];
    __shared__ unsigned wsum[32];
    const int l = blockIdx.x;
    const int nc = (int)stats[l * 8 + 7];
    if (nc == 0) return;
    unsigned* h = hist + (long long)l * nchunk * kMaxCells;
    for (int c = threadIdx.x; c < nc; c += 1024) {
        unsigned run = 0;
        for (int k = 0; k < nchunk; ++k) { const unsigned t = h[(long long)k * kMaxCells + c]; h[(long long)k * kMaxCells + c] = run; run += t; }
        tot[c] = run;
    }
    __syncthreads();
    // exclusive scan of tot[0..nc) (nc <= 2048: two elements per thread)
    const int c0 = 2 * threadIdx.x;
    const unsigned a = c0 < nc ? tot[c0] : 0u, b = c0 + 1 < nc ? tot[c0 + 1] : 0u;
    unsigned v = a + b;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
    if (lane == 31) wsum[warp] = v;
    __syncthreads();
    if (warp == 0) {
        unsigned w = wsum[lane];
        for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
        wsum[lane] = w;
    }
    __syncthreads();
    const unsigned excl = v - (a + b) + (warp ? wsum[warp - 1] : 0u);
    __syncthreads();
    if (c0 < nc) tot[c0] = excl;
    if (c0 + 1 < nc) tot[c0 + 1] = excl + a;
    __syncthreads();
    int* cf = cell_first + (long long)l * (kMaxCells + 1);
    for (int c = threadIdx.x; c < nc; c += 1024) {
        const unsigned f = tot[c];
        cf[c] = (int)f;
        for (int k = 0; k < nchunk; ++k) h[(long long)k * kMaxCells + c] += f;
    }
    if (threadIdx.x == 0) cf[nc] = (int)stats[l * 8 + 1] - (int)win[l * 4 + 2];       // the values inside the table
}

// Lanes of the warp whose `key` (BITS significant bits) equals this lane's: what __match_any_sync returns, from BITS
// ballots.  Measured in the stable scatter below (12 bits, ~25 distinct cells among 32 lanes): 1.84 -> 1.72 ms for the
// whole KDE stage; the radix sort of the preprocessing chain (preprocess_kernels.cu, ~30 distinct digits) is 3 % faster
// with MATCH.ANY and keeps it.  (A variant that partitions the chunk inside shared memory first and writes it out in
// destination order -- coalesced stores, 112 KB per CTA -- measured slower: 1.96 ms.)
template <int BITS>
__device__ __forceinline__ unsigned warp_match_bits(unsigned key) {
#ifdef JACKS_MATCH_HW
    return __match_any_sync(0xffffffffu, key);
#else
    unsigned peers = 0xffffffffu;
#pragma unroll
    for (int b = 0; b < BITS; ++b) {
        const unsigned m = __ballot_sync(0xffffffffu, (key >> b) & 1u);
        peers &= ((key >> b) & 1u) ? m : ~m;
    }
    return peers;
#endif
}

// Partition pass 3: STABLE scatter.  Warp w of the CTA owns a contiguous eighth of the chunk and walks it in order, 32
// consecutive values at a time; the lanes of a group that fall into the same cell take consecutive positions in lane
// order (match_any), and the warp's running count per cell (shared memory, one row per warp) carries on to the next group.
// Shared memory: 16-bit per-warp counts, two warps to a word (a warp owns 512 values, a chunk 4,096; both halves are
// updated with shared-memory atomics and a half never carries into its neighbour: its final value is at most 4,096),
// plus the chunk's first positions: 40 KB instead of 64, and with the registers capped at 64 four CTAs per SM instead
// of three for a pass that waits on memory: 1.77 -> 1.68 ms for the KDE stage.  (The same construction in the radix
// scatter of preprocess_kernels.cu, which carries 12-byte pairs and spills at 64 registers, measured 1 % slower.)
constexpr size_t kScatterSmem = (size_t)(kPartThreads / 64) * kMaxCells * sizeof(unsigned) + (size_t)kMaxCells * sizeof(unsigned);
__global__ void __launch_bounds__(kPartThreads, 4) kde_cell_scatter(const double* __restrict__ sT, long long ld, const double* __restrict__ stats,
                                                                 const double* __restrict__ win, int nchunk, const unsigned* __restrict__ hist,
                                                                 double* __restrict__ srt) {
    extern __shared__ unsigned cnt[];                    // [4 warp pairs][kMaxCells] packed counts + [kMaxCells] first positions
    const int l = blockIdx.y, chunk = blockIdx.x;
    const int nc = (int)stats[l * 8 + 7];
    if (nc == 0) return;
    const int n = (int)stats[l * 8 + 1];
    const double smin = stats[l * 8 + 6];
    const double inv_delta = 1.0 / (stats[l * 8 + 0] * kCellWidth);
    constexpr int W = kPartThreads / 32, SEG = kPartChunk / W;
    static_assert(W % 2 == 0 && kPartChunk <= 65535, "two 16-bit counters per word");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned* first = cnt + (W / 2) * kMaxCells;
    const unsigned* hb = hist + ((long long)l * nchunk + chunk) * kMaxCells;
    for (int c = threadIdx.x; c < (W / 2) * kMaxCells; c += kPartThreads) if ((c & (kMaxCells - 1)) < nc) cnt[c] = 0;
    for (int c = threadIdx.x; c < nc; c += kPartThreads) first[c] = hb[c];
    __syncthreads();
    const double* row = sT + (long long)l * ld;
    const double wlo = win[l * 4 + 0], whi = win[l * 4 + 1];
    const int base = chunk * kPartChunk + warp * SEG;
    unsigned* mine = cnt + (warp >> 1) * kMaxCells;
    const int sh = 16 * (warp & 1);
    double v[SEG / 32];
    int cell[SEG / 32];
#pragma unroll
    for (int k = 0; k < SEG / 32; ++k) {
        const int i = base + k * 32 + lane;
        cell[k] = -1 - lane;                              // past the end: a key of its own, never written
        v[k] = 0.0;
        if (i < n) {
            v[k] = row[i];
            if (!(v[k] < wlo || v[k] > whi)) { cell[k] = kde_cell_of(v[k], smin, inv_delta, nc); atomicAdd(&mine[cell[k]], 1u << sh); }
        }
    }
    __syncthreads();
    // per cell: exclusive prefix of the warps' counts (offsets inside the chunk's run of that cell)
    for (int c = threadIdx.x; c < nc; c += kPartThreads) {
        unsigned run = 0;
#pragma unroll
        for (int p = 0; p < W / 2; ++p) {
            const unsigned t = cnt[p * kMaxCells + c];
            const unsigned lo = run; run += t & 0xffffu;
            const unsigned hi = run; run += t >> 16;
            cnt[p * kMaxCells + c] = lo | (hi << 16);
        }
    }
    __syncthreads();
    double* dst = srt + (long long)l * ld;
#pragma unroll
    for (int k = 0; k < SEG / 32; ++k) {
        const unsigned peers = warp_match_bits<12>(cell[k] >= 0 ? (unsigned)cell[k] : (unsigned)kMaxCells);
        const int leader = __ffs(peers) - 1;
        const int rank = __popc(peers & ((1u << lane) - 1));
        unsigned old = 0;
        if (lane == leader && cell[k] >= 0)
            old = first[cell[k]] + ((atomicAdd(&mine[cell[k]], (unsigned)__popc(peers) << sh) >> sh) & 0xffffu);
        old = __shfl_sync(0xffffffffu, old, leader);
        if (cell[k] >= 0) dst[old + rank] = v[k];
        __syncwarp();
    }
}

// Moments of every cell: one warp per (line, cell); the lanes stride over the cell's points, a fixed butterfly adds up.
__global__ void __launch_bounds__(256) kde_cell_moments(const double* __restrict__ srt, long long ld, const double* __restrict__ stats,
                                                        const int* __restrict__ cell_first, double* __restrict__ mom) {
    const int l = blockIdx.y;
    const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const int nc = (int)stats[l * 8 + 7];
    if (c >= nc) return;
    const int* cf = cell_first + (long long)l * (kMaxCells + 1);
    const int first = cf[c], last = cf[c + 1];
    const double smin = stats[l * 8 + 6];
    const double rh = 1.0 / stats[l * 8 + 0];
    const double centre = ((double)c + 0.5) * kCellWidth;
    const double* row = srt + (long long)l * ld;
    double M[kNM];
#pragma unroll
    for (int k = 0; k < kNM; ++k) M[k] = 0.0;
    for (int i = first + lane; i < last; i += 32) {
        const double u = (row[i] - smin) * rh - centre;
        double t = 1.0;
        M[0] += 1.0;
#pragma unroll
        for (int k = 1; k < kNM; ++k) { t *= u * (1.0 / k); M[k] += t; }       // u^k / k!
    }
#pragma unroll
    for (int k = 0; k < kNM; ++k) {
        double s = M[k];
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        M[k] = s;
    }
    if (lane == 0) {
        double* rec = mom + ((long long)l * kMaxCells + c) * kNM;
#pragma unroll
        for (int k = 0; k < kNM; ++k) rec[k] = M[k];
    }
}

// Local expansions: one thread per (line, target cell) forms its kNM coefficients from the moments of the source cells
// in its window; the derivative table is the same for every thread at every step (constant-bank operands).
constexpr int kLocThreads = 128;
__global__ void __launch_bounds__(kLocThreads) kde_local_expansions(const double* __restrict__ stats, const int* __restrict__ cell_first,
                                                                    const double* __restrict__ mom, double* __restrict__ loc) {
    const int l = blockIdx.y;
    const int nc = (int)stats[l * 8 + 7];
    if (nc == 0) return;
    const int nt = nc + kKsat - kJmin;
    const int jj = blockIdx.x * kLocThreads + threadIdx.x;
    if (jj >= nt) return;
    const int j = jj + kJmin;
    double A[kNM];
#pragma unroll
    for (int m = 0; m < kNM; ++m) A[m] = 0.0;
    const double* M = mom + (long long)l * kMaxCells * kNM;
    // source cells c = j - k inside [0, nc), k = KSAT down to -KNEG (ascending c: a fixed order); k is the same for every
    // thread of the CTA at every step, so the derivative table is read through the constant bank without divergence
    for (int k = kKsat; k >= -kKneg; --k) {
        const int c = j - k;
        if (c < 0 || c >= nc) continue;
        const double* rec = M + (long long)c * kNM;
        if (rec[0] == 0.0) continue;
        const double* D = c_dtab + (k + kKneg) * (2 * kP + 1);
        double b[kNM];
#pragma unroll
        for (int q = 0; q < kNM; ++q) b[q] = (q & 1) ? -rec[q] : rec[q];
#pragma unroll
        for (int m = 0; m < kNM; ++m) {
            double s = 0.0;
#pragma unroll
            for (int q = 0; q < kNM; ++q) s = fma(D[m + q], b[q], s);
            A[m] += s;
        }
    }
    double fact = 1.0;
#pragma unroll
    for (int m = 1; m < kNM; ++m) { fact *= (double)m; A[m] /= fact; }
    const int csat = j - kKsat - 1;                                    // cells that count fully
    if (csat >= 0) A[0] += (double)cell_first[(long long)l * (kMaxCells + 1) + min(csat, nc - 1) + 1];
    double* out = loc + ((long long)l * kMaxTargets + jj) * kNM;
#pragma unroll
    for (int m = 0; m < kNM; ++m) out[m] = A[m];
}

// Sum over the nearest null values for a point more than 5 bandwidths below all of them (F is tiny there and its relative
// accuracy rests on those few terms).  The points of cells 0, 1, ... are added in stored order up to the cell from which
// on nothing can change the sum at 1e-18 relative: Phi(z - a)/Phi(z) <= exp(-((|z| + a)^2 - z^2)/2) for z <= -5, a >= 0,
// and the smallest null value itself (a = 0) is in the sum, so a >= sqrt(z^2 + 2 ln(1e18 n)) - |z| is negligible.
__device__ __forceinline__ double kde_tail_sum(double w, double z, double rh, int n, int nc, const int* __restrict__ cf,
                                               const double* __restrict__ row) {
    const double reach = sqrt(z * z + 2.0 * log(1e18 * (double)n)) + z;            // z < 0
    int c_stop = (int)ceil(reach * (1.0 / kCellWidth)) + 1;
    if (c_stop > nc) c_stop = nc;
    const int last = cf[c_stop];
    double tot = 0.0;
    for (int i = 0; i < last; ++i) tot += normcdf((w - row[i]) * rh);
    return tot;
}

// p-values: one thread per (gene, line), consecutive threads on consecutive lines (coalesced loads and stores); a degree-P
// polynomial from the line's table of local expansions (L2 resident: ~450 cells x 104 bytes per line).  Points in the
// far lower tail are queued (one aggregated atomic per warp) for kde_pvalues_tail: evaluated here they would hold up
// the 31 other lanes of their warp for a few thousand instructions each.
__global__ void __launch_bounds__(256) kde_pvalues_fast(const double* __restrict__ w1_genes, long long n_genes, int L,
                                                        const double* __restrict__ stats, const double* __restrict__ win,
                                                        const double* __restrict__ outv, const double* __restrict__ loc, int positive,
                                                        double* __restrict__ pvals, unsigned* __restrict__ tail_count,
                                                        long long* __restrict__ tail_idx) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    bool queue = false;
    if (idx < n_genes * L) {
        const int l = (int)(idx % L);
        const int nc = (int)stats[l * 8 + 7];
        if (nc > 0) {                                      // (nc == 0: this line is left to the O(n) kernel)
            const double wv = w1_genes[idx];
            if (!(wv == wv)) pvals[idx] = wv;
            else {
                const double w = positive ? -wv : wv;
                const int n = (int)stats[l * 8 + 1] - (int)win[l * 4 + 2];          // null values inside the cell table
                const double cterm = stats[l * 8 + 2], wgt = stats[l * 8 + 3], smin = stats[l * 8 + 6];
                const double z = (w - smin) / stats[l * 8 + 0];
                const double jf = floor(z * (1.0 / kCellWidth));
                double tot = 0.0;
                if (z < -38.6) tot = 0.0;                  // Phi underflows to exactly 0 for every null value
                else if (jf < (double)kJmin) queue = true;
                else if (jf >= (double)(nc + kKsat)) tot = (double)n;
                else {
                    const int j = (int)jf;
                    const double t = z - ((double)j + 0.5) * kCellWidth;
                    const double* a = loc + ((long long)l * kMaxTargets + (j - kJmin)) * kNM;
                    tot = a[kP];
#pragma unroll
                    for (int m = kP - 1; m >= 0; --m) tot = fma(tot, t, a[m]);
                }
                if (!queue) {
                    if (win[l * 4 + 2] != 0.0) tot += kde_outlier_sum(w, 1.0 / stats[l * 8 + 0], win, outv, l);     // heavy-tailed null
                    pvals[idx] = (tot - cterm) * wgt;
                }
            }
        }
    }
    const unsigned m = __ballot_sync(0xffffffffu, queue);
    if (m) {
        const int lane = threadIdx.x & 31;
        const int leader = __ffs(m) - 1;
        unsigned base = 0;
        if (lane == leader) base = atomicAdd(tail_count, __popc(m));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (queue) tail_idx[base + __popc(m & ((1u << lane) - 1))] = idx;
    }
}

__global__ void __launch_bounds__(256) kde_pvalues_tail(const double* __restrict__ w1_genes, int L, const double* __restrict__ srt,
                                                        long long ld, const double* __restrict__ stats, const double* __restrict__ win,
                                                        const double* __restrict__ outv, const int* __restrict__ cell_first,
                                                        int positive, double* __restrict__ pvals, const unsigned* __restrict__ tail_count,
                                                        const long long* __restrict__ tail_idx) {
    const unsigned count = *tail_count;
    for (unsigned q = blockIdx.x * blockDim.x + threadIdx.x; q < count; q += gridDim.x * blockDim.x) {
        const long long idx = tail_idx[q];
        const int l = (int)(idx % L);
        const double wv = w1_genes[idx];
        const double w = positive ? -wv : wv;
        const double rh = 1.0 / stats[l * 8 + 0];
        const double z = (w - stats[l * 8 + 6]) / stats[l * 8 + 0];
        double tot = kde_tail_sum(w, z, rh, (int)stats[l * 8 + 1], (int)stats[l * 8 + 7], cell_first + (long long)l * (kMaxCells + 1),
                                  srt + (long long)l * ld);
        if (win[l * 4 + 2] != 0.0) tot += kde_outlier_sum(w, rh, win, outv, l);
        pvals[idx] = (tot - stats[l * 8 + 2]) * stats[l * 8 + 3];
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

// Null density of every line from the pseudo-gene effects (jacks_io.py:76-79): transposed rows, bandwidth, the rows
// partitioned by cell, cell moments and the table of local expansions, all in ctx->tmp.  `out` stays valid until the
// next KDE call on this context.
int jacks::kde_null_prepare(JacksCtx* ctx, const double* w1_pseudo, int64_t n_pseudo, int32_t n_lines, int32_t positive,
                            cudaStream_t st, KdeNull* out) {
    if (n_pseudo >= (1LL << 31) - kPartChunk) return fail(JACKS_E_INVALID, "too many pseudo-genes (%lld)", (long long)n_pseudo);
    const long long ld = (n_pseudo + kChunk - 1) / kChunk * kChunk;            // rows padded with +inf for the O(n) kernels
    const int nchunk = (int)((n_pseudo + kPartChunk - 1) / kPartChunk);
    const size_t L = (size_t)n_lines;
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_sT = take((size_t)ld * L * sizeof(double)), o_srt = take((size_t)ld * L * sizeof(double));
    const size_t o_stats = take(L * 8 * sizeof(double)), o_nan = take(L * sizeof(int));
    const size_t o_win = take(L * 4 * sizeof(double)), o_outv = take(L * kMaxOut * sizeof(double));
    const size_t o_hist = take(L * nchunk * kMaxCells * sizeof(unsigned)), o_first = take(L * (kMaxCells + 1) * sizeof(int));
    const size_t o_mom = take(L * kMaxCells * kNM * sizeof(double)), o_loc = take(L * kMaxTargets * kNM * sizeof(double));
    const size_t o_tab = take((size_t)kNK * (2 * kP + 1) * sizeof(double));
    int rc = ctx->tmp.reserve(off);
    if (rc != JACKS_OK) return rc;
    cudaError_t e;
    char* base = (char*)ctx->tmp.p;
    double* sT = (double*)(base + o_sT);
    double* srt = (double*)(base + o_srt);
    double* stats = (double*)(base + o_stats);
    double* win = (double*)(base + o_win);
    double* outv = (double*)(base + o_outv);
    int* nan_count = (int*)(base + o_nan);
    unsigned* hist = (unsigned*)(base + o_hist);
    int* cell_first = (int*)(base + o_first);
    double* mom = (double*)(base + o_mom);
    double* loc = (double*)(base + o_loc);
    const bool brute = getenv("JACKS_KDE_BRUTE") != nullptr;          // validation: force the O(n) sums
    if (!ctx->kde_attr_set) {          // once per device: the derivative table and the scatter kernel's shared memory
        double* tab = (double*)(base + o_tab);
        kde_deriv_table<<<(kNK + 127) / 128, 128, 0, st>>>(tab);
        if ((e = cudaMemcpyToSymbolAsync(c_dtab, tab, (size_t)kNK * (2 * kP + 1) * sizeof(double), 0, cudaMemcpyDeviceToDevice, st)) != cudaSuccess) return cuda_fail(e, "cudaMemcpyToSymbolAsync");
        if ((e = cudaFuncSetAttribute(kde_cell_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kScatterSmem)) != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute");
        ctx->kde_attr_set = true;
        ctx->launches++;
    }
    if ((e = cudaMemsetAsync(nan_count, 0, L * sizeof(int), st)) != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync");
    kde_transpose<<<dim3((unsigned)((n_pseudo + 31) / 32), (unsigned)((n_lines + 31) / 32)), 256, 0, st>>>(w1_pseudo, n_pseudo, n_lines, ld, positive ? 1 : 0, sT, nan_count);
    kde_row_stats<<<n_lines, kRowThreads, 0, st>>>(sT, n_pseudo, ld, nan_count, stats, win, outv);
    ctx->launches += 2;
    if (!brute) {
        kde_cell_hist<<<dim3((unsigned)nchunk, (unsigned)n_lines), kPartThreads, 0, st>>>(sT, ld, stats, win, nchunk, hist);
        kde_cell_scan<<<n_lines, 1024, 0, st>>>(stats, win, nchunk, hist, cell_first);
        kde_cell_scatter<<<dim3((unsigned)nchunk, (unsigned)n_lines), kPartThreads, kScatterSmem, st>>>(sT, ld, stats, win, nchunk, hist, srt);
        kde_cell_moments<<<dim3(kMaxCells / 8, (unsigned)n_lines), 256, 0, st>>>(srt, ld, stats, cell_first, mom);
        kde_local_expansions<<<dim3((kMaxTargets + kLocThreads - 1) / kLocThreads, (unsigned)n_lines), kLocThreads, 0, st>>>(stats, cell_first, mom, loc);
        ctx->launches += 5;
    }
    if ((e = cudaGetLastError()) != cudaSuccess) return cuda_fail(e, "kde null kernels launch");
    out->sT = sT; out->srt = srt; out->stats = stats; out->cell_first = cell_first; out->loc = loc;
    out->win = win; out->outv = outv;
    out->ld = ld; out->n_lines = n_lines; out->positive = positive ? 1 : 0; out->brute = brute ? 1 : 0;
    return JACKS_OK;
}

// p-values of `n_genes` rows of w1_genes [n_genes, n_lines] under a prepared null.
int jacks::kde_eval(JacksCtx* ctx, const KdeNull& nl, const double* w1_genes, int64_t n_genes, double* pvals, cudaStream_t st) {
    if (n_genes == 0) return JACKS_OK;
    cudaError_t e;
    const int n_lines = nl.n_lines;
    if (!nl.brute) {
        const long long total = (long long)n_genes * n_lines;
        int rc = ctx->kde_tail.reserve(256 + (size_t)total * sizeof(long long));
        if (rc != JACKS_OK) return rc;
        unsigned* tail_count = (unsigned*)ctx->kde_tail.p;
        long long* tail_idx = (long long*)((char*)ctx->kde_tail.p + 256);
        if ((e = cudaMemsetAsync(tail_count, 0, sizeof(unsigned), st)) != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync");
        kde_pvalues_fast<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(w1_genes, n_genes, n_lines, nl.stats, nl.win, nl.outv, nl.loc, nl.positive, pvals, tail_count, tail_idx);
        kde_pvalues_tail<<<4 * ctx->n_sm, 256, 0, st>>>(w1_genes, n_lines, nl.srt, nl.ld, nl.stats, nl.win, nl.outv, nl.cell_first, nl.positive, pvals, tail_count, tail_idx);
        ctx->launches += 2;
    }
    // lines whose range needs more cells than the table holds (far outliers), or everything when forced
    const long long per_cta = (long long)kPvThreads * kGT;
    dim3 grid((unsigned)((n_genes + per_cta - 1) / per_cta), (unsigned)n_lines);
    kde_pvalues<<<grid, kPvThreads, 0, st>>>(w1_genes, n_genes, n_lines, nl.sT, nl.ld, nl.stats, nl.positive, nl.brute ? 0 : 1, pvals);
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
