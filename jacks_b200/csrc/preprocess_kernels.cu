// Preprocessing error model on the GPU (reference jacks/jacks/preprocess.py).
//
//   jacks_lognorm_median_host   log2(count + prior) (preprocess.py:147) and per-column centring,
//                               normalizeLogCounts 'median' (preprocess.py:80-81): transpose to
//                               column-major, radix-select the nanmedian of each column, transpose back.
//   jacks_condense_host         condense_normalised_counts (preprocess.py:61-67): per sample the mean over
//                               its replicate columns and calc_posterior_sd (preprocess.py:25-58):
//                                 means, vars = data.mean(axis=1), nanstd(data, axis=1)**2
//                                 sort guides by (mean, var, index)                       (:39-40)
//                                 window = min(max(int(N/100), 30), 800)                  (:35)
//                                 sliding mean of the sorted vars, the reference's shifted recurrence (:16-22)
//                                 reverse running maximum (monotonize, :8-13; NaN restarts the run)
//                                 sd = sqrt(m) scattered back to guide order              (:55-56)
//
// Everything is HBM/L2-bound integer-and-double streaming work: coalesced tiles, shared-memory staging,
// no tensor cores.  The sort is a segmented stable LSD radix sort (11-bit digits: variance, then mean, from index
// order).  Device entry points (jacks_normalize_device, jacks_condense_device) chain on one stream without leaving
// the GPU; jacks_preprocess_host is the whole chain for one count table with ONE upload and ONE download.
// The sliding mean uses prefix sums: for w+1 <= k <= N-w-2 the reference's recurrence
// m[k] = m[k-1] + (x[k+w+1] - x[k-w])/(2w+1) telescopes to
// m[w] + (P[k+w+2] - P[2w+2] - P[k-w+1] + P[1])/(2w+1) with P[i] = sum_{t<i} x[t] (agreement ~1e-14).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <algorithm>
#include <vector>

#include "../../include/jacks_b200.h"
#include "capi_internal.h"

namespace jacks {
namespace {

__device__ __forceinline__ double d_nan() { return __longlong_as_double(0x7ff8000000000000LL); }
__device__ __forceinline__ double d_inf() { return __longlong_as_double(0x7ff0000000000000LL); }

// order-preserving map double -> uint64 (all NaNs are excluded by the callers)
__device__ __forceinline__ unsigned long long sortable(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b & 0x8000000000000000ULL) ? ~b : (b | 0x8000000000000000ULL);
}
__device__ __forceinline__ double unsortable(unsigned long long k) {
    const unsigned long long b = (k & 0x8000000000000000ULL) ? (k & 0x7fffffffffffffffULL) : ~k;
    return __longlong_as_double((long long)b);
}

// ---- log2 + transpose ----------------------------------------------------------------------------
// T[c][i] = log2(counts[i][c] + prior); 32x32 tiles through shared memory, coalesced on both sides.
//
// log2_table (optional): host-computed log2(k + prior) for k = 0..table_len-1.  Integer counts take their
// logarithm from it, which makes them bit-identical to the numpy log2 of the calling host; device log2 is
// <= 1 ulp too, but not the same ulp, and a last-bit difference can swap two guides whose means tie to
// within an ulp in the (mean, var, index) sort downstream -- a visible change of their smoothed sd.
__global__ void __launch_bounds__(256) log_transpose_kernel(const double* __restrict__ counts, long long N, int C,
                                                            double prior, int take_log, const double* __restrict__ log2_table,
                                                            long long table_len, double* __restrict__ T, long long ldT) {
    __shared__ double tile[32][33];
    const long long i0 = (long long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const long long i = i0 + r;
        const int c = c0 + tx;
        double v = 0.0;
        if (i < N && c < C) {
            const double cnt = counts[i * C + c];
            if (!take_log) v = cnt;
            else if (log2_table && cnt >= 0.0 && cnt < (double)table_len && cnt == floor(cnt)) v = log2_table[(long long)cnt];
            else v = log2(cnt + prior);
        }
        tile[r][tx] = v;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int c = c0 + r;
        const long long i = i0 + tx;
        if (c < C && i < N) T[(long long)c * ldT + i] = tile[tx][r];
    }
}

// out[i][c] = (T[c][i] - shift[c]) * scale[c]
__global__ void __launch_bounds__(256) untranspose_shift_kernel(const double* __restrict__ T, long long ldT, long long N, int C,
                                                                const double* __restrict__ shift,
                                                                const double* __restrict__ scale, double* __restrict__ out) {
    __shared__ double tile[32][33];
    const long long i0 = (long long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const int c = c0 + r;
        const long long i = i0 + tx;
        double v = 0.0;
        if (c < C && i < N) {
            v = T[(long long)c * ldT + i] - shift[c];
            if (scale) v = v / scale[c];
        }
        tile[r][tx] = v;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const long long i = i0 + r;
        const int c = c0 + tx;
        if (i < N && c < C) out[i * C + c] = tile[tx][r];
    }
}

// ---- per-column nanmedian by radix select --------------------------------------------------------
// One CTA per column of the column-major T.  mode 0: median of T[c][.]; mode 1: median of |T[c][.] - shift[c]|.
// Rows can be restricted to a subset (row_mask != nullptr: use rows with mask != 0).
// Six passes over 11-bit digits of the order-preserving key find the lower middle element (the first pass also counts the
// valid entries); ONE more pass finds the upper middle one -- the smallest key above it unless enough entries tie with
// it.  (The first version made 8-bit passes for both order statistics separately: 17 sweeps of the column, 2.3 ms at the
// Avana shape.)
constexpr int kSelBits = 11;
constexpr int kSelBins = 1 << kSelBits;

// Which bin holds the k-th (0-based) entry?  Block-wide scan over kSelBins counts with 1,024 threads (two bins each).
__device__ __forceinline__ void block_find_bin(const unsigned* hist, long long k, int* s_bin, long long* s_before, unsigned* wsum) {
    const int c0 = 2 * threadIdx.x;
    const unsigned a = hist[c0], b = hist[c0 + 1];
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
    const long long excl = (long long)(v - (a + b)) + (warp ? (long long)wsum[warp - 1] : 0LL);
    if (a && excl <= k && k < excl + a) { *s_bin = c0; *s_before = excl; }
    else if (b && excl + a <= k && k < excl + a + b) { *s_bin = c0 + 1; *s_before = excl + a; }
    __syncthreads();
}

__global__ void __launch_bounds__(1024) column_median_kernel(const double* __restrict__ T, long long ldT, long long N,
                                                             const unsigned char* __restrict__ row_mask,
                                                             const double* __restrict__ shift, int mode,
                                                             double* __restrict__ med) {
    __shared__ unsigned hist[kSelBins];
    __shared__ unsigned wsum[32];
    __shared__ int s_bin;
    __shared__ long long s_before;
    __shared__ unsigned long long s_min;
    __shared__ unsigned s_le;
    const int c = blockIdx.x;
    const double* col = T + (long long)c * ldT;
    const double sh = (mode == 1) ? shift[c] : 0.0;
    unsigned long long prefix = 0;
    long long n = 0, k = 0;
    // digits from the top: 11 bits at shifts 53, 42, 31, 20, 9, then the last 9 bits
    for (int pass = 0; pass < 6; ++pass) {
        const int shift_bits = pass < 5 ? 53 - kSelBits * pass : 0;
        const int bits = pass < 5 ? kSelBits : 9;
        for (int bnum = threadIdx.x; bnum < kSelBins; bnum += blockDim.x) hist[bnum] = 0;
        __syncthreads();
        const unsigned long long himask = (pass == 0) ? 0ULL : (~0ULL << (shift_bits + bits));
        for (long long i = threadIdx.x; i < N; i += blockDim.x) {
            double v = col[i];
            if (mode == 1) v = fabs(v - sh);
            if (!(v == v) || (row_mask && !row_mask[i])) continue;
            const unsigned long long key = sortable(v);
            if ((key & himask) == prefix) atomicAdd(&hist[(unsigned)((key >> shift_bits) & ((1u << bits) - 1))], 1u);
        }
        __syncthreads();
        if (pass == 0) {
            // the valid entries: total of the first histogram
            unsigned part = hist[2 * threadIdx.x] + hist[2 * threadIdx.x + 1];
            for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = part;
            __syncthreads();
            unsigned tot = 0;
            for (int w = 0; w < 32; ++w) tot += wsum[w];
            __syncthreads();
            n = tot;
            if (n == 0) { if (threadIdx.x == 0) med[c] = d_nan(); return; }
            k = (n - 1) / 2;                                    // 0-based rank of the lower middle element
        }
        block_find_bin(hist, k, &s_bin, &s_before, wsum);
        prefix |= (unsigned long long)s_bin << shift_bits;
        k -= s_before;
        __syncthreads();
    }
    const double lo = unsortable(prefix);
    double hi = lo;
    if (!(n & 1)) {
        // upper middle element: rank n/2 = lower rank + 1.  Entries <= lo number `le`; if le > n/2 it ties with lo,
        // otherwise it is the smallest entry above lo.
        if (threadIdx.x == 0) { s_min = ~0ULL; s_le = 0; }
        __syncthreads();
        unsigned le = 0;
        unsigned long long mn = ~0ULL;
        for (long long i = threadIdx.x; i < N; i += blockDim.x) {
            double v = col[i];
            if (mode == 1) v = fabs(v - sh);
            if (!(v == v) || (row_mask && !row_mask[i])) continue;
            const unsigned long long key = sortable(v);
            if (key <= prefix) ++le; else if (key < mn) mn = key;
        }
        for (int o = 16; o > 0; o >>= 1) {
            le += __shfl_xor_sync(0xffffffffu, le, o);
            const unsigned long long t = __shfl_xor_sync(0xffffffffu, mn, o);
            if (t < mn) mn = t;
        }
        if ((threadIdx.x & 31) == 0) { atomicAdd(&s_le, le); atomicMin(&s_min, mn); }
        __syncthreads();
        if ((long long)s_le <= n / 2) hi = unsortable(s_min);
    }
    if (threadIdx.x == 0) med[c] = (n & 1) ? lo : 0.5 * (lo + hi);
}

__global__ void scale_mad_kernel(double* mad, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < C) mad[c] = 1.4826 * mad[c];          // preprocess.py:84
}

// normtype 'mode' (preprocess.py:85-91): np.histogram(x, bins=100) of the column, a 5-tap smoothing of the counts,
// and the centre of the fullest smoothed bin as the shift.  numpy's uniform-bin path is restated step for step
// (numpy/lib/_histograms_impl.py): edges = linspace(min, max, 101) as i*step + min with the last edge forced to
// max, a first guess trunc((x-min)/(max-min)*100), then the one-step corrections against the edges -- so every value
// lands in the bin numpy puts it in.  No FMA contraction.  One CTA per column of the column-major T.
__global__ void __launch_bounds__(1024) column_mode_kernel(const double* __restrict__ T, long long ldT, long long N,
                                                           double* __restrict__ shift) {
    __shared__ double s_edges[101];
    __shared__ unsigned int s_hist[100];
    __shared__ double s_red[64];
    const int c = blockIdx.x;
    const double* col = T + (long long)c * ldT;
    double lo = d_inf(), hi = -d_inf();
    for (long long i = threadIdx.x; i < N; i += blockDim.x) { const double v = col[i]; lo = fmin(lo, v); hi = fmax(hi, v); }
    for (int o = 16; o > 0; o >>= 1) { lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { s_red[warp] = lo; s_red[32 + warp] = hi; }
    for (int b = threadIdx.x; b < 100; b += blockDim.x) s_hist[b] = 0;
    __syncthreads();
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { lo = fmin(lo, s_red[w]); hi = fmax(hi, s_red[32 + w]); }
    if (lo == hi) { lo = lo - 0.5; hi = hi + 0.5; }                          // numpy widens a zero-width range
    const double delta = __dadd_rn(hi, -lo);
    const double step = delta / 100.0;
    for (int b = threadIdx.x; b <= 100; b += blockDim.x)
        s_edges[b] = (b == 100) ? hi : __dadd_rn(__dmul_rn((double)b, step), lo);
    __syncthreads();
    for (long long i = threadIdx.x; i < N; i += blockDim.x) {
        const double v = col[i];
        if (!(v >= lo && v <= hi)) continue;
        int idx = (int)__dmul_rn(__dadd_rn(v, -lo) / delta, 100.0);
        if (idx == 100) idx = 99;
        if (v < s_edges[idx]) idx -= 1;
        if (v >= s_edges[idx + 1] && idx != 99) idx += 1;
        atomicAdd(&s_hist[idx], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double best = -1.0;
        int arg = 0;
        for (int i = 0; i < 96; ++i) {
            double v = __dmul_rn(0.1, (double)s_hist[i]);
            v = __dadd_rn(v, __dmul_rn(0.2, (double)s_hist[i + 1]));
            v = __dadd_rn(v, __dmul_rn(0.4, (double)s_hist[i + 2]));
            v = __dadd_rn(v, __dmul_rn(0.2, (double)s_hist[i + 3]));
            v = __dadd_rn(v, __dmul_rn(0.1, (double)s_hist[i + 4]));
            if (v > best) { best = v; arg = i; }                              // np.argmax: first maximum
        }
        shift[c] = __dadd_rn(__dmul_rn(0.5, s_edges[arg + 3]), __dmul_rn(0.5, s_edges[arg + 2]));
    }
}

// ---- per-sample mean / variance and sort keys ----------------------------------------------------
// numpy semantics, no FMA contraction: mean = (x0 + x1 + ...)/R in column order; nanstd uses the NaN-aware
// mean, population variance, and the reference squares the sqrt (nanstd(...)**2, preprocess.py:38).
__global__ void __launch_bounds__(256) sample_stats_kernel(const double* __restrict__ lc, long long N, int C,
                                                           const int* __restrict__ col_off, const int* __restrict__ cols,
                                                           int S, long long Np, double* __restrict__ data,
                                                           double* __restrict__ kmean, double* __restrict__ kvar,
                                                           int* __restrict__ kidx) {
    const int s = blockIdx.y;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Np) return;
    const int c0 = col_off[s], R = col_off[s + 1] - c0;
    const long long o = (long long)s * Np + i;
    if (i >= N) { kmean[o] = d_inf(); kvar[o] = d_inf(); kidx[o] = 0x7fffffff; return; }
    double sum = 0.0, nsum = 0.0;
    int cnt = 0;
    for (int r = 0; r < R; ++r) {
        const double v = lc[i * C + cols[c0 + r]];
        sum = __dadd_rn(sum, v);
        if (v == v) { nsum = __dadd_rn(nsum, v); ++cnt; }
    }
    const double mean = sum / (double)R;
    double var = d_nan();
    if (cnt > 0) {
        const double nmean = nsum / (double)cnt;
        double ss = 0.0;
        for (int r = 0; r < R; ++r) {
            const double v = lc[i * C + cols[c0 + r]];
            if (v == v) { const double d = __dadd_rn(v, -nmean); ss = __dadd_rn(ss, __dmul_rn(d, d)); }
        }
        const double sd = sqrt(ss / (double)cnt);
        var = __dmul_rn(sd, sd);
    }
    data[(i * S + s) * 2 + 0] = mean;
    if (R < 2) data[(i * S + s) * 2 + 1] = d_nan();          // single replicate: undefined sd, preprocess.py:32
    // NaN keys sort last (Python tuple comparison with NaN is not a total order; such inputs are out of contract)
    kmean[o] = (mean == mean) ? mean : d_inf();
    kvar[o] = (var == var) ? var : d_inf();
    kidx[o] = (int)i;
}

// The same for many samples at once, tiled 32 guides x 32 samples: consecutive threads take consecutive SAMPLES of one guide
// (their replicate columns are neighbours in the row, so the reads of the [N][C] matrix and the writes of the cube's mean
// plane are coalesced; one thread per guide read 8 bytes out of every 6.4 KB row), the keys cross a shared-memory tile and
// leave coalesced along the guide axis.  Same arithmetic, same order.
__global__ void __launch_bounds__(256) sample_stats_tiled_kernel(const double* __restrict__ lc, long long N, int C,
                                                                 const int* __restrict__ col_off, const int* __restrict__ cols,
                                                                 int S, long long Np, double* __restrict__ data,
                                                                 double* __restrict__ kmean, double* __restrict__ kvar,
                                                                 int* __restrict__ kidx) {
    __shared__ double t_mean[32][33], t_var[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8
    const long long i0 = (long long)blockIdx.x * 32;
    const int s0 = blockIdx.y * 32;
    const int s = s0 + tx;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int ti = ty + 8 * k;
        const long long i = i0 + ti;
        double kmv = d_inf(), kvv = d_inf();
        if (s < S && i < N) {
            const int c0 = col_off[s], R = col_off[s + 1] - c0;
            double sum = 0.0, nsum = 0.0;
            int cnt = 0;
            for (int r = 0; r < R; ++r) {
                const double v = lc[i * C + cols[c0 + r]];
                sum = __dadd_rn(sum, v);
                if (v == v) { nsum = __dadd_rn(nsum, v); ++cnt; }
            }
            const double mean = sum / (double)R;
            double var = d_nan();
            if (cnt > 0) {
                const double nmean = nsum / (double)cnt;
                double ss = 0.0;
                for (int r = 0; r < R; ++r) {
                    const double v = lc[i * C + cols[c0 + r]];
                    if (v == v) { const double d = __dadd_rn(v, -nmean); ss = __dadd_rn(ss, __dmul_rn(d, d)); }
                }
                const double sd = sqrt(ss / (double)cnt);
                var = __dmul_rn(sd, sd);
            }
            data[(i * S + s) * 2 + 0] = mean;
            if (R < 2) data[(i * S + s) * 2 + 1] = d_nan();
            kmv = (mean == mean) ? mean : d_inf();
            kvv = (var == var) ? var : d_inf();
        }
        t_mean[ti][tx] = kmv;
        t_var[ti][tx] = kvv;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int ts = ty + 8 * k;
        const long long i = i0 + tx;
        if (s0 + ts < S && i < Np) {
            const long long o = (long long)(s0 + ts) * Np + i;
            kmean[o] = t_mean[tx][ts];
            kvar[o] = t_var[tx][ts];
            kidx[o] = (i < N) ? (int)i : 0x7fffffff;
        }
    }
}

// ---- segmented stable LSD radix sort of (mean, variance, index) keys ---------------------------------------------------
// Every sample is one segment of N keys.  The order wanted is Python's tuple order (mean, var, index), preprocess.py:39-40:
// a stable sort by the variance followed by a stable sort by the mean, starting from index order, gives exactly that.
// Each of the two is 6 passes over 11-bit digits of the order-preserving 64-bit image of the double; a pass is
// histogram (per segment and chunk) -> scan -> stable scatter of (key, index) pairs, 12 bytes per element each way.
// The first version ran a bitonic network on padded power-of-two segments: 136 compare-exchange sweeps of 20-byte
// records over 1.8x the data, 17 ms at the Avana shape.
constexpr int kRsBits = 11;
constexpr int kRsBins = 1 << kRsBits;
constexpr int kRsChunk = 4096;           // keys per CTA
constexpr int kRsThreads = 256;

__device__ __forceinline__ unsigned long long sortable_key(double v) {
    if (v == 0.0) v = 0.0;               // -0.0 and +0.0 compare equal in the reference's tuple sort
    return sortable(v);
}

// keys[s][i] = image of src[s][i] (or of src[s][idx[s][i]] when idx is given); idx_out[s][i] = i when asked for
// `gate` (all sort kernels): when given, the launch does nothing unless *gate != 0 (the exact two-key fallback of
// sort_mean_var is enqueued unconditionally and runs only if the tie repair asked for it -- no host round trip).
__global__ void __launch_bounds__(256) rs_load_keys(const double* __restrict__ src, const int* __restrict__ idx, long long ld, long long N,
                                                    unsigned long long* __restrict__ keys, int* __restrict__ idx_out,
                                                    const int* __restrict__ gate) {
    if (gate && *gate == 0) return;
    const long long o = (long long)blockIdx.y * ld;
    // grid-stride along the segment: one trip on the product path; the gated launches take a few CTAs per segment
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
        const long long from = idx ? idx[o + i] : i;
        keys[o + i] = sortable_key(src[o + from]);
        if (idx_out) idx_out[o + i] = (int)i;
    }
}

__device__ __forceinline__ void rs_hist_item(const unsigned long long* __restrict__ keys, long long ld, long long N, int shift,
                                             int nchunk, unsigned* __restrict__ hist, int s, int chunk, unsigned* sh) {
    for (int c = threadIdx.x; c < kRsBins; c += kRsThreads) sh[c] = 0;
    __syncthreads();
    const unsigned long long* row = keys + (long long)s * ld;
    const long long i1 = min(N, (long long)(chunk + 1) * kRsChunk);
    for (long long i = (long long)chunk * kRsChunk + threadIdx.x; i < i1; i += kRsThreads)
        atomicAdd(&sh[(unsigned)(row[i] >> shift) & (kRsBins - 1)], 1u);
    __syncthreads();
    unsigned* out = hist + ((long long)s * nchunk + chunk) * kRsBins;
    for (int c = threadIdx.x; c < kRsBins; c += kRsThreads) out[c] = sh[c];
}
__global__ void __launch_bounds__(kRsThreads) rs_hist(const unsigned long long* __restrict__ keys, long long ld, long long N, int shift,
                                                      int nchunk, unsigned* __restrict__ hist) {
    __shared__ unsigned sh[kRsBins];
    rs_hist_item(keys, ld, N, shift, nchunk, hist, blockIdx.y, blockIdx.x, sh);
}
// The gated twins (fallback of sort_mean_var): a small grid that loops over the (chunk, segment) items.  Their launches
// are almost always empty, and 7,000 CTAs that return at once -- each with 64 KB of shared memory in the scatter's case --
// cost 7..22 us a launch, 0.4 ms over the 36 launches of a fallback that does not run.
__global__ void __launch_bounds__(kRsThreads) rs_hist_gated(const unsigned long long* __restrict__ keys, long long ld, long long N, int shift,
                                                            int nchunk, int n_seg, unsigned* __restrict__ hist, const int* __restrict__ gate) {
    __shared__ unsigned sh[kRsBins];
    if (*gate == 0) return;
    for (long long wi = blockIdx.x; wi < (long long)nchunk * n_seg; wi += gridDim.x) {
        rs_hist_item(keys, ld, N, shift, nchunk, hist, (int)(wi / nchunk), (int)(wi % nchunk), sh);
        __syncthreads();
    }
}

// one CTA per segment: hist[chunk][bin] becomes the position of the first key of (chunk, bin) in the sorted segment
__global__ void __launch_bounds__(1024) rs_scan(int nchunk, unsigned* __restrict__ hist, const int* __restrict__ gate) {
    __shared__ unsigned tot[kRsBins];
    __shared__ unsigned wsum[32];
    if (gate && *gate == 0) return;
    unsigned* h = hist + (long long)blockIdx.x * nchunk * kRsBins;
    constexpr int KMAX = 32;                      // chunks held in registers (131,072 keys per segment); more: serial path
    if (nchunk <= KMAX) {
        // all loads of a bin first (independent), prefix in registers, then the stores: interleaved, every store to h
        // fences the next load and the chain costs a global round trip per chunk
        for (int c = threadIdx.x; c < kRsBins; c += 1024) {
            unsigned held[KMAX];
#pragma unroll
            for (int k = 0; k < KMAX; ++k) held[k] = k < nchunk ? h[(long long)k * kRsBins + c] : 0u;
            unsigned run = 0;
#pragma unroll
            for (int k = 0; k < KMAX; ++k) { const unsigned t = held[k]; held[k] = run; run += t; }
#pragma unroll
            for (int k = 0; k < KMAX; ++k) if (k < nchunk) h[(long long)k * kRsBins + c] = held[k];
            tot[c] = run;
        }
    } else {
        for (int c = threadIdx.x; c < kRsBins; c += 1024) {
            unsigned run = 0;
            for (int k = 0; k < nchunk; ++k) { const unsigned t = h[(long long)k * kRsBins + c]; h[(long long)k * kRsBins + c] = run; run += t; }
            tot[c] = run;
        }
    }
    __syncthreads();
    const int c0 = 2 * threadIdx.x;
    const unsigned a = tot[c0], b = tot[c0 + 1];
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
    tot[c0] = excl;
    tot[c0 + 1] = excl + a;
    __syncthreads();
    for (int c = threadIdx.x; c < kRsBins; c += 1024) {
        const unsigned f = tot[c];
        if (f) for (int k = 0; k < nchunk; ++k) h[(long long)k * kRsBins + c] += f;        // independent read-modify-writes
    }
}

// Stable scatter: warp w of the CTA owns a contiguous eighth of the chunk and walks it in order, 32 consecutive keys at
// a time; keys of a group with equal digit take consecutive positions in lane order (match_any); the warp's running
// count per digit carries on to the next group.
// The pass waits on memory (ncu: long_scoreboard, 37 % of the warp slots in use with three CTAs per SM), so it is built
// for occupancy: the per-warp counts are 16-bit, two warps to a word (a warp owns 512 keys, a chunk 4,096; both halves
// are updated with shared-memory atomics and a half never carries into its neighbour: its final value is at most 4,096)
// -- 32 KB instead of 64, plus the chunk's 8 KB of first positions -- and the keys are READ TWICE (once to count, once to
// move; the second time from L2) instead of being held in 64 registers between the two phases.
constexpr int kRsScatterSmem = (kRsThreads / 64) * kRsBins * 4 + kRsBins * 4;
__device__ __forceinline__ void rs_scatter_item(const unsigned long long* __restrict__ kin, const int* __restrict__ iin,
                                                long long ld, long long N, int shift, int nchunk,
                                                const unsigned* __restrict__ hist, unsigned long long* __restrict__ kout,
                                                int* __restrict__ iout, int s, int chunk, unsigned* cnt) {
    constexpr int W = kRsThreads / 32, SEG = kRsChunk / W;
    static_assert(W % 2 == 0 && kRsChunk <= 65535, "two 16-bit counters per word");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned* first = cnt + (W / 2) * kRsBins;            // [kRsBins] position of the chunk's first key of a digit
    const unsigned* hb = hist + ((long long)s * nchunk + chunk) * kRsBins;
    for (int c = threadIdx.x; c < (W / 2) * kRsBins; c += kRsThreads) cnt[c] = 0;
    for (int c = threadIdx.x; c < kRsBins; c += kRsThreads) first[c] = hb[c];
    __syncthreads();
    const long long o = (long long)s * ld;
    const long long base = (long long)chunk * kRsChunk + warp * SEG;
    unsigned* mine = cnt + (warp >> 1) * kRsBins;         // my half: bits 16*(warp & 1) ...
    const int sh = 16 * (warp & 1);
#pragma unroll 8
    for (int k = 0; k < SEG / 32; ++k) {
        const long long i = base + k * 32 + lane;
        if (i < N) atomicAdd(&mine[(unsigned)(kin[o + i] >> shift) & (kRsBins - 1)], 1u << sh);
    }
    __syncthreads();
    // per digit: exclusive prefix of the warps' counts (offsets inside the chunk's run of that digit)
    for (int c = threadIdx.x; c < kRsBins; c += kRsThreads) {
        unsigned run = 0;
#pragma unroll
        for (int p = 0; p < W / 2; ++p) {
            const unsigned t = cnt[p * kRsBins + c];
            const unsigned lo = run; run += t & 0xffffu;
            const unsigned hi = run; run += t >> 16;
            cnt[p * kRsBins + c] = lo | (hi << 16);
        }
    }
    __syncthreads();
#pragma unroll 4
    for (int k = 0; k < SEG / 32; ++k) {
        const long long i = base + k * 32 + lane;
        unsigned long long key = 0;
        int pay = 0, dig = -1 - lane;                     // past the end: a digit of its own, never written
        if (i < N) { key = kin[o + i]; pay = iin[o + i]; dig = (int)((unsigned)(key >> shift) & (kRsBins - 1)); }
        // (equal digits from kRsBits + 1 ballots instead of MATCH.ANY, which pays in the KDE scatter, measured 3 % slower here)
        const unsigned peers = __match_any_sync(0xffffffffu, dig);
        const int leader = __ffs(peers) - 1;
        const int rank = __popc(peers & ((1u << lane) - 1));
        unsigned old = 0;
        if (lane == leader && dig >= 0)
            old = first[dig] + ((atomicAdd(&mine[dig], (unsigned)__popc(peers) << sh) >> sh) & 0xffffu);
        old = __shfl_sync(0xffffffffu, old, leader);
        if (dig >= 0) { kout[o + old + rank] = key; iout[o + old + rank] = pay; }
        __syncwarp();
    }
}
__global__ void __launch_bounds__(kRsThreads, 5) rs_scatter(const unsigned long long* __restrict__ kin, const int* __restrict__ iin,
                                                         long long ld, long long N, int shift, int nchunk,
                                                         const unsigned* __restrict__ hist, unsigned long long* __restrict__ kout,
                                                         int* __restrict__ iout) {
    extern __shared__ unsigned cnt[];                    // [4 warp pairs][kRsBins] packed counts + [kRsBins] first positions
    rs_scatter_item(kin, iin, ld, N, shift, nchunk, hist, kout, iout, blockIdx.y, blockIdx.x, cnt);
}
__global__ void __launch_bounds__(kRsThreads) rs_scatter_gated(const unsigned long long* __restrict__ kin, const int* __restrict__ iin,
                                                               long long ld, long long N, int shift, int nchunk, int n_seg,
                                                               const unsigned* __restrict__ hist, unsigned long long* __restrict__ kout,
                                                               int* __restrict__ iout, const int* __restrict__ gate) {
    extern __shared__ unsigned cnt[];
    if (*gate == 0) return;
    for (long long wi = blockIdx.x; wi < (long long)nchunk * n_seg; wi += gridDim.x) {
        rs_scatter_item(kin, iin, ld, N, shift, nchunk, hist, kout, iout, (int)(wi / nchunk), (int)(wi % nchunk), cnt);
        __syncthreads();
    }
}

// sorted copies: dm[k] = km[idx[k]], dv[k] = kv[idx[k]]
__global__ void __launch_bounds__(256) rs_gather_sorted(const double* __restrict__ km, const double* __restrict__ kv, const int* __restrict__ idx,
                                                        long long ld, long long N, double* __restrict__ dm, double* __restrict__ dv,
                                                        const int* __restrict__ gate) {
    if (gate && *gate == 0) return;
    const long long o = (long long)blockIdx.y * ld;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
        const int from = idx[o + i];
        dm[o + i] = km[o + from];
        dv[o + i] = kv[o + from];
    }
}

// Tie repair after a stable sort by the MEAN alone (from index order).  The order wanted is (mean, var, index): only the
// runs of equal means can be out of order, and inside a run the stable sort has left index order, so a stable sort of
// the run by the variance finishes the job.  Screens have many such runs (integer counts: two guides with the same
// unordered pair of counts tie in the mean, and their variances can differ in the last bit) but most are short.
//   rs_tie_find, one thread per element: the thread at the head of a run measures it (bounded scan) and
//     * insertion-sorts a run of 2..kTieSelf elements itself (stable; an ordered run costs one comparison per element),
//     * queues a run of up to kTieWarp elements for a warp, of up to kTieCta elements for a CTA (rs_tie_sort_runs:
//       rank counting on the variance keys held in shared memory, stable through the position tie-break);
//     a thread that sees an inversion (same mean, larger variance first) inside a run LONGER than kTieCta raises *gate and
//     the exact two-key radix sort runs instead (sort_mean_var).  A long run that is already in order -- thousands of
//     zero-count guides with identical mean and variance -- costs nothing.
// Keys are compared through their sortable images, exactly as the radix passes see them (-0.0 == +0.0, NaN by bits).
// The means of a run are equal under that image; they stay where they are (only (var, index) move).
constexpr int kTieSelf = 16, kTieWarp = 256, kTieCta = 2048;
struct TieRun { long long at; int len; int pad; };       // at: position in the [S][ld] arrays
// ctl[0]: gate, ctl[1]: warp-sized runs queued, ctl[2]: CTA-sized runs queued
__global__ void __launch_bounds__(256) rs_tie_find(const double* __restrict__ km, double* __restrict__ kv, int* __restrict__ idx,
                                                   long long ld, long long N, int* __restrict__ ctl,
                                                   TieRun* __restrict__ warp_runs, TieRun* __restrict__ cta_runs, long long warp_cap, long long cta_cap) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const long long o = (long long)blockIdx.y * ld;
    const double* m = km + o; double* v = kv + o; int* ix = idx + o;
    const unsigned long long key = sortable_key(m[i]);
    const bool tie_next = (i + 1 < N) && sortable_key(m[i + 1]) == key;
    if (!tie_next) return;                                         // last element of its run (or a run of one)
    const bool head = (i == 0) || sortable_key(m[i - 1]) != key;
    if (head) {
        long long e = i + 2;
        while (e < N && e - i <= kTieCta && sortable_key(m[e]) == key) ++e;
        const long long len = e - i;
        if (len > kTieCta) return;                                 // a long run: its inversions (if any) raise the gate below
        if (len > kTieSelf) {
            const bool big = len > kTieWarp;
            const long long slot = atomicAdd(&ctl[big ? 2 : 1], 1);
            if (slot < (big ? cta_cap : warp_cap)) { TieRun r; r.at = o + i; r.len = (int)len; r.pad = 0; (big ? cta_runs : warp_runs)[slot] = r; }
            else ctl[0] = 1;                                       // cannot happen (the lists hold every possible run); stay exact anyway
            return;
        }
        for (long long a = i + 1; a < e; ++a) {                    // stable insertion sort by the variance
            const double va = v[a];
            const int ia = ix[a];
            const unsigned long long ka = sortable_key(va);
            long long b = a;
            while (b > i && sortable_key(v[b - 1]) > ka) { v[b] = v[b - 1]; ix[b] = ix[b - 1]; --b; }
            if (b != a) { v[b] = va; ix[b] = ia; }
        }
        return;
    }
    // not a head: does the pair (i, i+1) stand in the wrong order inside a run too long to be repaired?  (Inside a shorter
    // run the head or its helpers may be permuting these very elements: whatever is read here, the run is found short.)
    if (sortable_key(v[i]) <= sortable_key(v[i + 1])) return;
    long long h = i;
    while (h > 0 && i - h <= kTieCta && sortable_key(m[h - 1]) == key) {
        --h;
        if (((i - h) & 255) == 0 && *(volatile int*)ctl) return;   // someone else has already asked for the fallback
    }
    long long e = i + 2;
    while (e < N && e - h <= kTieCta && sortable_key(m[e]) == key) ++e;
    if (e - h > kTieCta) ctl[0] = 1;
}

// The common case in one pass: a CTA takes a tile of kTieTile consecutive sorted elements into shared memory, finds the
// runs of equal means with two block-wide scans (start of my run: max-scan of the head positions; end: reverse min-scan),
// and every element of a run that lies INSIDE the tile counts its rank among the run's variance keys (shared-memory
// reads) and moves straight to its place.  No thread walks a run serially -- screens whose counts tie by the thousand
// (every element of the bench's Poisson(300) library sits in a run) cost ~30 shared-memory compares per element.
// Runs that cross a tile border are left to their head: measured, then self-sorted or queued as in rs_tie_find; the
// gate logic for over-long runs is the same.
constexpr int kTieTile = 1024;          // 4 consecutive elements per thread
__global__ void __launch_bounds__(256) rs_tie_tile(const double* __restrict__ km, double* __restrict__ kv, int* __restrict__ idx,
                                                   long long ld, long long N, int* __restrict__ ctl,
                                                   TieRun* __restrict__ warp_runs, TieRun* __restrict__ cta_runs,
                                                   long long warp_cap, long long cta_cap) {
    __shared__ unsigned long long s_key[kTieTile];
    __shared__ double s_val[kTieTile];
    __shared__ int s_idx[kTieTile];
    __shared__ int s_wmax[8], s_wmin[8];
    constexpr int NONE = -1, INF = 30000;
    const long long o = (long long)blockIdx.y * ld;
    const long long t0 = (long long)blockIdx.x * kTieTile;
    if (t0 >= N) return;
    const int n = (int)min((long long)kTieTile, N - t0);
    const double* m = km + o + t0; double* v = kv + o + t0; int* ix = idx + o + t0;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int j0 = 4 * t;
    // heads: element j starts a run if it is the first of the segment or its mean differs from its predecessor's
    unsigned long long mk[4], prev = 0;
    bool head[4];
    if (j0 < n) prev = (t0 + j0 > 0) ? sortable_key(m[j0 - 1]) : 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int j = j0 + q;
        mk[q] = (j < n) ? sortable_key(m[j]) : 0;
        head[q] = (j < n) && ((t0 + j == 0) || mk[q] != (q ? mk[q - 1] : prev));
    }
    // does a run end exactly at the tile's end?  (position n acts as a head then)
    const bool end_is_head = (t0 + n >= N) || sortable_key(m[n]) != sortable_key(m[n - 1]);
    // start of my run: inclusive max-scan of (head ? j : NONE)
    int hs[4], run = NONE;
#pragma unroll
    for (int q = 0; q < 4; ++q) { if (head[q]) run = j0 + q; hs[q] = run; }
    int inc = run;
    for (int d = 1; d < 32; d <<= 1) { const int u = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc = max(inc, u); }
    if (lane == 31) s_wmax[warp] = inc;
    // end of my run: first head position AFTER j (reverse exclusive min-scan), INF if none inside the tile
    int he[4], nxt = INF;
#pragma unroll
    for (int q = 3; q >= 0; --q) { he[q] = nxt; if (head[q]) nxt = j0 + q; }
    if (j0 <= n - 1 && n - 1 < j0 + 4 && end_is_head) {
        // the thread owning the last element: position n is a head for everything after the last head in my segment
#pragma unroll
        for (int q = 0; q < 4; ++q) if (j0 + q < n && he[q] == INF) he[q] = n;
        if (nxt == INF) nxt = n;
    }
    int dec = nxt;
    for (int d = 1; d < 32; d <<= 1) { const int u = __shfl_down_sync(0xffffffffu, dec, d); if (lane + d < 32) dec = min(dec, u); }
    if (lane == 0) s_wmin[warp] = dec;
    __syncthreads();
    int before = NONE, after = INF;
    for (int w2 = 0; w2 < warp; ++w2) before = max(before, s_wmax[w2]);
    for (int w2 = 7; w2 > warp; --w2) after = min(after, s_wmin[w2]);
    {
        const int up = __shfl_up_sync(0xffffffffu, inc, 1);             // inclusive max of the lanes before me
        const int carry = max(before, lane ? up : NONE);
        const int dn = __shfl_down_sync(0xffffffffu, dec, 1);           // inclusive min of the lanes after me
        const int carry_r = min(after, lane < 31 ? dn : INF);
#pragma unroll
        for (int q = 0; q < 4; ++q) { hs[q] = max(hs[q], carry); if (he[q] == INF) he[q] = carry_r; }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int j = j0 + q;
        if (j < n) {
            const double x = v[j];
            s_val[j] = x; s_key[j] = sortable_key(x); s_idx[j] = ix[j];
        }
    }
    __syncthreads();
#pragma unroll 1
    for (int q = 0; q < 4; ++q) {
        const int j = j0 + q;
        if (j >= n) break;
        const int h = hs[q], e = he[q];
        if (h >= 0 && e <= n) {                                        // my run lies inside the tile
            if (e - h < 2) continue;
            const unsigned long long kj = s_key[j];
            int rank = 0;
            for (int k = h; k < e; ++k) {
                const unsigned long long kk = s_key[k];
                rank += (kk < kj || (kk == kj && k < j)) ? 1 : 0;
            }
            if (h + rank != j) { v[h + rank] = s_val[j]; ix[h + rank] = s_idx[j]; }
            continue;
        }
        // my run crosses a tile border
        const long long i = t0 + j;
        const double* mg = km + o; double* vg = kv + o; int* ig = idx + o;
        const unsigned long long key = mk[q];
        const bool tie_next = (i + 1 < N) && sortable_key(mg[i + 1]) == key;
        if (!tie_next) continue;
        if (head[q]) {                                                  // its head sits in this tile: measure, sort or queue it
            long long ee = i + 2;
            while (ee < N && ee - i <= kTieCta && sortable_key(mg[ee]) == key) ++ee;
            const long long len = ee - i;
            if (len > kTieCta) continue;
            if (len > kTieSelf) {
                const bool big = len > kTieWarp;
                const long long slot = atomicAdd(&ctl[big ? 2 : 1], 1);
                if (slot < (big ? cta_cap : warp_cap)) { TieRun r; r.at = o + i; r.len = (int)len; r.pad = 0; (big ? cta_runs : warp_runs)[slot] = r; }
                else ctl[0] = 1;
                continue;
            }
            for (long long a2 = i + 1; a2 < ee; ++a2) {
                const double va = vg[a2];
                const int ia = ig[a2];
                const unsigned long long ka = sortable_key(va);
                long long b2 = a2;
                while (b2 > i && sortable_key(vg[b2 - 1]) > ka) { vg[b2] = vg[b2 - 1]; ig[b2] = ig[b2 - 1]; --b2; }
                if (b2 != a2) { vg[b2] = va; ig[b2] = ia; }
            }
            continue;
        }
        // not the head: an inversion inside a run too long to be repaired raises the gate
        if (sortable_key(vg[i]) <= sortable_key(vg[i + 1])) continue;
        long long hh = i;
        while (hh > 0 && i - hh <= kTieCta && sortable_key(mg[hh - 1]) == key) {
            --hh;
            if (((i - hh) & 255) == 0 && *(volatile int*)ctl) break;
        }
        long long ee = i + 2;
        while (ee < N && ee - hh <= kTieCta && sortable_key(mg[ee]) == key) ++ee;
        if (ee - hh > kTieCta) ctl[0] = 1;
    }
}

// Stable rank-counting sort of the queued runs by the variance key: TEAM threads (a warp or the whole CTA) per run.
template <bool CTA>
__global__ void __launch_bounds__(256) rs_tie_sort_runs(double* __restrict__ kv, int* __restrict__ idx, const int* __restrict__ ctl,
                                                        const TieRun* __restrict__ runs, long long cap) {
    constexpr int LEN = CTA ? kTieCta : kTieWarp, TEAMS = CTA ? 1 : 8, TEAM = CTA ? 256 : 32;
    __shared__ unsigned long long s_key[TEAMS][LEN];
    __shared__ double s_val[TEAMS][LEN];
    __shared__ int s_idx[TEAMS][LEN];
    if (ctl[0]) return;                                           // the exact fallback redoes everything
    const long long n_runs = min((long long)ctl[CTA ? 2 : 1], cap);
    const int team = CTA ? 0 : (int)(threadIdx.x >> 5), t = CTA ? (int)threadIdx.x : (int)(threadIdx.x & 31);
    for (long long r = (long long)blockIdx.x * TEAMS + team; r < n_runs; r += (long long)gridDim.x * TEAMS) {
        const TieRun run = runs[r];
        double* v = kv + run.at; int* ix = idx + run.at;
        for (int j = t; j < run.len; j += TEAM) {
            const double x = v[j];
            s_val[team][j] = x; s_idx[team][j] = ix[j]; s_key[team][j] = sortable_key(x);
        }
        if (CTA) __syncthreads(); else __syncwarp();
        for (int j = t; j < run.len; j += TEAM) {
            const unsigned long long kj = s_key[team][j];
            int rank = 0;
            for (int k = 0; k < run.len; ++k) {
                const unsigned long long kk = s_key[team][k];
                rank += (kk < kj || (kk == kj && k < j)) ? 1 : 0;
            }
            v[rank] = s_val[team][j]; ix[rank] = s_idx[team][j];
        }
        if (CTA) __syncthreads(); else __syncwarp();
    }
}

// ---- sliding mean (shifted, as the reference), reverse running max, sqrt, scatter ----------------
// One CTA per sample.  P = exclusive prefix sums of the sorted variances (global scratch, chunked scan).
constexpr int kScanThreads = 512;        // four CTAs per SM: the 401 samples of an Avana-shaped screen are resident at once
constexpr int kScanItems = 8;          // consecutive elements per thread in the chunked scans

__device__ double block_exclusive_scan(double v, double* sh, double* total) {
    // inclusive warp scan
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double x = v;
    for (int o = 1; o < 32; o <<= 1) { const double y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    __syncthreads();
    if (lane == 31) sh[warp] = x;
    __syncthreads();
    if (warp == 0) {
        double w = (lane < (int)(blockDim.x >> 5)) ? sh[lane] : 0.0;
        for (int o = 1; o < 32; o <<= 1) { const double y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += y; }
        sh[lane] = w;
    }
    __syncthreads();
    const double before = (warp == 0) ? 0.0 : sh[warp - 1];
    *total = sh[(blockDim.x >> 5) - 1];
    return before + x - v;
}

// m[k] = the reference's shifted sliding mean of x[0..N) with half-width `window` (preprocess.py:16-22); p is
// scratch for N+1 prefix sums.  Whole CTA (kScanThreads threads) cooperates.
__device__ void dev_window_smooth(const double* __restrict__ x, long long N, long long window, double* __restrict__ p,
                                  double* __restrict__ m, double* sh) {
    // exclusive prefix sums p[i] = sum_{t<i} x[t], i = 0..N.  A thread owns kScanItems consecutive elements (64 contiguous
    // bytes), so a chunk of 8,192 elements costs ONE block-wide scan: 9 instead of 71 rounds of barriers for 72,000 guides.
    double carry = 0.0;
    for (long long c0 = 0; c0 < N; c0 += (long long)kScanThreads * kScanItems) {
        const long long base = c0 + (long long)threadIdx.x * kScanItems;
        double v[kScanItems], run = 0.0;
#pragma unroll
        for (int j = 0; j < kScanItems; ++j) { v[j] = (base + j < N) ? x[base + j] : 0.0; }
#pragma unroll
        for (int j = 0; j < kScanItems; ++j) { const double t = v[j]; v[j] = run; run += t; }      // v[j]: sum of my elements before j
        double total;
        const double ex = block_exclusive_scan(run, sh, &total) + carry;
#pragma unroll
        for (int j = 0; j < kScanItems; ++j)
            if (base + j < N) p[base + j] = ex + v[j];
        carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) p[N] = carry;
    __syncthreads();
    // numpy slices clip, so the head is the mean of x[0:min(2w+1,N)] and the tail slice x[N-2w-1:] wraps when
    // N < 2w+1; the tail assignment comes last and wins where the ranges overlap.
    const long long w = window;
    const double span = (double)(2 * w + 1);
    const long long head_hi = (2 * w + 1 < N) ? 2 * w + 1 : N;
    long long tail_lo = N - 2 * w - 1;
    if (tail_lo < 0) { tail_lo += N; if (tail_lo < 0) tail_lo = 0; }
    const double head = (p[head_hi] - p[0]) / (double)head_hi;              // np.mean(x[0:2w+1])
    const double tail = (p[N] - p[tail_lo]) / (double)(N - tail_lo);        // np.mean(x[N-2w-1:])
    for (long long k = threadIdx.x; k < N; k += kScanThreads) {
        double v;
        if (k >= N - w - 1) v = tail;
        else if (k <= w) v = head;
        else v = head + ((p[k + w + 2] - p[2 * w + 2]) - (p[k - w + 1] - p[1])) / span;
        m[k] = v;
    }
    __syncthreads();
}

// In-place reverse running maximum (monotonize, preprocess.py:8-13): a NaN entry stays NaN and restarts the run.
// Chunked from the end; inside a chunk a segmented max-scan (warp shuffles, then a serial pass over the 32 warp
// summaries).  Whole CTA cooperates.
__device__ void dev_reverse_running_max(double* __restrict__ m, long long N) {
    // A thread owns kScanItems consecutive elements and walks them from the top down; what crosses thread boundaries is
    // the pair (run leaving my segment at its low end, did my segment hold a NaN), scanned over the threads in reverse
    // element order exactly as single elements were before.
    __shared__ double s_carry;
    __shared__ double w_last[32];
    __shared__ int w_blocked[32];
    __shared__ double w_in[32];
    if (threadIdx.x == 0) s_carry = d_nan();         // "no carry" marker: the last element has no successor
    __syncthreads();
    constexpr long long kChunk = (long long)kScanThreads * kScanItems;
    for (long long hi = N; hi > 0; hi -= kChunk) {
        // thread 0 takes the top kScanItems elements of the chunk: [top - kScanItems + 1, top]
        const long long top = hi - 1 - (long long)threadIdx.x * kScanItems;
        double out[kScanItems];
        double run = d_nan();                        // the run inside my segment (NaN = none yet / cut)
        bool seg_nan = false;
#pragma unroll
        for (int j = 0; j < kScanItems; ++j) {       // j = 0 is the TOP element of my segment
            const long long k = top - j;
            double v = (k >= 0) ? m[k] : d_nan();
            if (k < 0) { out[j] = v; continue; }     // below the array: nothing to do (and nothing below it reads my run)
            if (v == v) {
                if (run == run && run > v) v = run;
                run = v;
            } else {
                seg_nan = true;
                run = d_nan();
            }
            out[j] = v;
        }
        // dead threads (wholly below the array) pass the incoming run through untouched: treat them as NaN-free, empty
        const bool dead = (top < 0);
        // inclusive scan of (run, blocked) over the threads; a thread whose segment held a NaN neither receives the incoming
        // run at its low end nor lets it pass
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        double cur = run;
        bool blocked = seg_nan;
        for (int o = 1; o < 32; o <<= 1) {
            const double pv = __shfl_up_sync(0xffffffffu, cur, o);
            const int pb = __shfl_up_sync(0xffffffffu, (int)blocked, o);
            if (lane >= o) {
                if (!blocked && pv == pv && (!(cur == cur) || pv > cur)) cur = pv;
                blocked = blocked || (pb != 0);
            }
        }
        __syncthreads();
        if (lane == 31) { w_last[warp] = cur; w_blocked[warp] = (int)blocked; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double r = s_carry;                       // running max entering the chunk (NaN = none)
            for (int q = 0; q < (int)(blockDim.x >> 5); ++q) {
                w_in[q] = r;
                const double last = w_last[q];
                if (w_blocked[q]) r = last;            // the warp saw a NaN: the incoming run is cut there
                else r = (r == r && (!(last == last) || r > last)) ? r : last;
            }
            s_carry = r;
        }
        __syncthreads();
        // the run entering MY segment from above: the previous thread's inclusive value, raised by the warp's incoming run
        // unless a NaN sits between the warp's start and that thread's low end
        double prev = __shfl_up_sync(0xffffffffu, cur, 1);
        const int prev_blocked = __shfl_up_sync(0xffffffffu, (int)blocked, 1);
        const double win = w_in[warp];
        double incoming;
        if (lane == 0) incoming = win;
        else if (prev_blocked) incoming = prev;
        else incoming = (win == win && (!(prev == prev) || win > prev)) ? win : prev;
        if (!dead) {
            bool open = (incoming == incoming);
#pragma unroll
            for (int j = 0; j < kScanItems; ++j) {
                const long long k = top - j;
                if (k < 0) break;
                const double v = out[j];
                if (!(v == v)) { open = false; continue; }        // a NaN stays NaN and ends the incoming run
                m[k] = (open && incoming > v) ? incoming : v;
            }
        }
        __syncthreads();
    }
}

// One CTA per sample: smooth the sorted variances, monotonize, sd = sqrt(m) scattered back to guide order.
__global__ void __launch_bounds__(kScanThreads) smooth_monotone_scatter_kernel(const double* __restrict__ kv, const int* __restrict__ ki,
                                                                               long long Np, long long N, int S, int window,
                                                                               int monotonize, const int* __restrict__ col_off,
                                                                               double* __restrict__ P, double* __restrict__ M,
                                                                               double* __restrict__ data) {
    __shared__ double sh[32];
    const int s = blockIdx.x;
    if (col_off[s + 1] - col_off[s] < 2) return;             // single replicate: sd stays NaN
    const int* idx = ki + (long long)s * Np;
    double* m = M + (long long)s * Np;
    dev_window_smooth(kv + (long long)s * Np, N, window, P + (long long)s * (Np + 1), m, sh);
    if (monotonize) dev_reverse_running_max(m, N);
    for (long long k = threadIdx.x; k < N; k += kScanThreads)
        data[((long long)idx[k] * S + s) * 2 + 1] = sqrt(m[k]);       // sqrt(NaN) = NaN
}

// calc_posterior_sd with a guide subset (preprocess.py:44-52): the mean-variance curve is fitted on the subset
// members (in sorted order), then linearly interpolated (np.interp) at every guide's mean.  One CTA.
//   km, kv, ki: keys sorted by (mean, var, index), N real entries; in_set[guide] != 0 marks subset members.
//   sub_mean, sub_var, P, M: scratch of at least N+1 doubles each.
__global__ void __launch_bounds__(kScanThreads) subset_sd_kernel(const double* __restrict__ km, const double* __restrict__ kv,
                                                                 const int* __restrict__ ki, long long N,
                                                                 const unsigned char* __restrict__ in_set, int window,
                                                                 int monotonize, double* __restrict__ sub_mean,
                                                                 double* __restrict__ sub_var, double* __restrict__ P,
                                                                 double* __restrict__ M, double* __restrict__ sd_out) {
    __shared__ double sh[32];
    __shared__ long long s_M;
    // order-preserving compaction of the subset members
    double carry = 0.0;
    for (long long c0 = 0; c0 < N; c0 += kScanThreads) {
        const long long k = c0 + threadIdx.x;
        const bool member = (k < N) && in_set[ki[k]];
        double total;
        const double ex = block_exclusive_scan(member ? 1.0 : 0.0, sh, &total);
        if (member) {
            const long long pos = (long long)(carry + ex);
            sub_mean[pos] = km[k];
            sub_var[pos] = kv[k];
        }
        carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) s_M = (long long)carry;
    __syncthreads();
    const long long Ms = s_M;
    dev_window_smooth(sub_var, Ms, window, P, M, sh);
    if (monotonize) dev_reverse_running_max(M, Ms);
    // np.interp(x = all sorted means, xp = subset means, fp = M)  (numpy/_core/src/multiarray/compiled_base.c)
    for (long long k = threadIdx.x; k < N; k += kScanThreads) {
        const double x = km[k];
        double r;
        if (x != x) r = x;
        else if (x < sub_mean[0]) r = M[0];                       // left = fp[0]
        else if (x >= sub_mean[Ms - 1]) r = M[Ms - 1];            // x == xp[-1] -> fp[-1]; beyond: right = fp[-1]
        else {
            long long lo = 0, hi = Ms - 1;                // invariant: sub_mean[lo] <= x < sub_mean[hi]
            while (hi - lo > 1) {
                const long long mid = (lo + hi) >> 1;
                if (sub_mean[mid] <= x) lo = mid; else hi = mid;
            }
            if (sub_mean[lo] == x) r = M[lo];
            else {
                const double slope = (M[lo + 1] - M[lo]) / (sub_mean[lo + 1] - sub_mean[lo]);
                r = __dadd_rn(__dmul_rn(slope, x - sub_mean[lo]), M[lo]);
                if (r != r) {
                    r = __dadd_rn(__dmul_rn(slope, x - sub_mean[lo + 1]), M[lo + 1]);
                    if (r != r && M[lo] == M[lo + 1]) r = M[lo];
                }
            }
        }
        sd_out[ki[k]] = sqrt(r);
    }
}

}  // namespace
}  // namespace jacks

using namespace jacks;

namespace jacks {
namespace {

// Bump allocator over a grow-only device buffer (256-byte aligned pieces).
struct Arena {
    size_t off = 0;
    size_t take(size_t bytes) { const size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; }
};

__global__ void __launch_bounds__(256) permute_rows_kernel(const double* __restrict__ in, const long long* __restrict__ order, long long N,
                                                           int C, double* __restrict__ out) {
    const long long r = blockIdx.x;
    const double* src = in + order[r] * C;
    double* dst = out + r * C;
    for (int c = threadIdx.x; c < C; c += blockDim.x) dst[c] = src[c];
}

// normalizeLogCounts on device buffers (preprocess.py:77-96), optionally fused with log2(count + prior) (:147).
// Scratch: ctx->tmp ([C][N] transposed copy + per-column statistics).
int normalize_core(JacksCtx* ctx, const double* d_values, long long N, int C, int take_log, double prior, const double* d_table,
                   long long table_len, int normtype, const unsigned char* d_mask, double* d_out, cudaStream_t st) {
    int rc;
    if ((rc = ctx->tmp.reserve((size_t)N * C * sizeof(double) + 2 * (size_t)C * sizeof(double))) != JACKS_OK) return rc;
    double* T = (double*)ctx->tmp.p;
    double* med = T + (size_t)N * C;
    double* mad = med + C;
    dim3 tg((unsigned)((N + 31) / 32), (unsigned)((C + 31) / 32));
    log_transpose_kernel<<<tg, 256, 0, st>>>(d_values, N, C, prior, take_log ? 1 : 0, table_len ? d_table : nullptr, table_len, T, N);
    if (normtype == JACKS_NORM_MODE) column_mode_kernel<<<C, 1024, 0, st>>>(T, N, N, med);          // preprocess.py:85-91
    else column_median_kernel<<<C, 1024, 0, st>>>(T, N, N, d_mask, nullptr, 0, med);             // preprocess.py:81 / :93
    ctx->launches += 2;
    const double* scale = nullptr;
    if (normtype == JACKS_NORM_ZMAD) {                                                            // preprocess.py:82-84
        column_median_kernel<<<C, 1024, 0, st>>>(T, N, N, nullptr, med, 1, mad);
        scale_mad_kernel<<<(C + 255) / 256, 256, 0, st>>>(mad, C);
        scale = mad;
        ctx->launches += 2;
    }
    untranspose_shift_kernel<<<tg, 256, 0, st>>>(T, N, N, C, med, scale, d_out);
    ctx->launches++;
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "normalize kernels");
    return JACKS_OK;
}

// Sort every segment's (mean, var, index) keys: on return km_s / kv_s / ki_s hold the sorted keys.  Scratch from `base`.
struct SortScratch { unsigned long long* kA; unsigned long long* kB; int* iA; int* iB; unsigned* hist; int* gate; TieRun* warp_runs; TieRun* cta_runs; };
__global__ void set_flag_kernel(int* f) { *f = 1; }
inline long long tie_warp_cap(int S, long long ld) { return (long long)S * (ld / kTieSelf + 1); }
inline long long tie_cta_cap(int S, long long ld) { return (long long)S * (ld / kTieWarp + 1); }
size_t sort_scratch_bytes(int S, long long ld, Arena* a, size_t o[8]) {
    const int nchunk = (int)((ld + kRsChunk - 1) / kRsChunk);
    o[0] = a->take((size_t)S * ld * 8); o[1] = a->take((size_t)S * ld * 8);
    o[2] = a->take((size_t)S * ld * 4); o[3] = a->take((size_t)S * ld * 4);
    o[4] = a->take((size_t)S * nchunk * kRsBins * 4);
    o[5] = a->take(256);
    o[6] = a->take((size_t)tie_warp_cap(S, ld) * sizeof(TieRun));
    o[7] = a->take((size_t)tie_cta_cap(S, ld) * sizeof(TieRun));
    return a->off;
}
int sort_mean_var(JacksCtx* ctx, cudaStream_t st, const double* km, const double* kv, int S, long long ld, long long N,
                  const SortScratch& w, double* km_s, double* kv_s, int** ki_sorted) {
    cudaError_t e;
    if (!ctx->pre_attr_set) {              // per device
        if ((e = cudaFuncSetAttribute(rs_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, kRsScatterSmem)) != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute");
        if ((e = cudaFuncSetAttribute(rs_scatter_gated, cudaFuncAttributeMaxDynamicSharedMemorySize, kRsScatterSmem)) != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute");
        ctx->pre_attr_set = true;
    }
    const int nchunk = (int)((N + kRsChunk - 1) / kRsChunk);
    dim3 eg((unsigned)((N + 255) / 256), (unsigned)S), cg((unsigned)nchunk, (unsigned)S);
    dim3 eg_gated((unsigned)std::min<long long>((N + 255) / 256, 4), (unsigned)S);      // grids of the (normally empty) gated launches
    unsigned long long* kin = w.kA; unsigned long long* kout = w.kB;
    int* iin = w.iA; int* iout = w.iB;
    // `which`: 0 stable by the variance from index order, 1 stable by the mean.  An even number of passes: the sorted
    // keys and indices end in the buffers they started in.
    auto radix = [&](int which, bool from_index_order, const int* gate) {
        rs_load_keys<<<gate ? eg_gated : eg, 256, 0, st>>>(which == 0 ? kv : km, from_index_order ? nullptr : iin, ld, N, kin, from_index_order ? iin : nullptr, gate);
        ctx->launches++;
        for (int shift = 0; shift < 64; shift += kRsBits) {
            if (!gate) {
                rs_hist<<<cg, kRsThreads, 0, st>>>(kin, ld, N, shift, nchunk, w.hist);
                rs_scan<<<S, 1024, 0, st>>>(nchunk, w.hist, nullptr);
                rs_scatter<<<cg, kRsThreads, kRsScatterSmem, st>>>(kin, iin, ld, N, shift, nchunk, w.hist, kout, iout);
            } else {
                const unsigned wg = (unsigned)std::min<long long>((long long)nchunk * S, 592);
                rs_hist_gated<<<wg, kRsThreads, 0, st>>>(kin, ld, N, shift, nchunk, S, w.hist, gate);
                rs_scan<<<S, 1024, 0, st>>>(nchunk, w.hist, gate);
                rs_scatter_gated<<<wg, kRsThreads, kRsScatterSmem, st>>>(kin, iin, ld, N, shift, nchunk, S, w.hist, kout, iout, gate);
            }
            ctx->launches += 3;
            std::swap(kin, kout); std::swap(iin, iout);
        }
    };
    static const bool two_key = []() { const char* v = getenv("JACKS_SORT_TWO_KEY"); return v && atoi(v) != 0; }();   // tests: force the fallback
    // Fast path: ONE stable sort by the mean, then the runs of equal means are put in (var, index) order in place
    // (rs_tie_repair).  Six radix passes instead of twelve.
    if ((e = cudaMemsetAsync(w.gate, 0, 4 * sizeof(int), st)) != cudaSuccess) return cuda_fail(e, "memset");
    if (!two_key) {
        radix(1, true, nullptr);
        rs_gather_sorted<<<eg, 256, 0, st>>>(km, kv, iin, ld, N, km_s, kv_s, nullptr);
        static const bool tie_find = []() { const char* v = getenv("JACKS_TIE_FIND"); return v && atoi(v) != 0; }();      // A/B: the per-head version
        if (tie_find) rs_tie_find<<<eg, 256, 0, st>>>(km_s, kv_s, iin, ld, N, w.gate, w.warp_runs, w.cta_runs, tie_warp_cap(S, ld), tie_cta_cap(S, ld));
        else rs_tie_tile<<<dim3((unsigned)((N + kTieTile - 1) / kTieTile), (unsigned)S), 256, 0, st>>>(km_s, kv_s, iin, ld, N, w.gate, w.warp_runs, w.cta_runs, tie_warp_cap(S, ld), tie_cta_cap(S, ld));
        rs_tie_sort_runs<false><<<4 * 148, 256, 0, st>>>(kv_s, iin, w.gate, w.warp_runs, tie_warp_cap(S, ld));
        rs_tie_sort_runs<true><<<148, 256, 0, st>>>(kv_s, iin, w.gate, w.cta_runs, tie_cta_cap(S, ld));
        ctx->launches += 4;
    } else {
        set_flag_kernel<<<1, 1, 0, st>>>(w.gate);
    }
    // Exact fallback, enqueued always, run only when the repair met an out-of-order run longer than it handles:
    // stable by the variance, then stable by the mean (Python's tuple order, preprocess.py:39-40).
    static const bool no_fallback = []() { const char* v = getenv("JACKS_SORT_NO_FALLBACK"); return v && atoi(v) != 0; }();   // MEASUREMENT ONLY: what the gated launches cost
    if (!no_fallback || two_key) {
        radix(0, true, w.gate);
        radix(1, false, w.gate);
        rs_gather_sorted<<<eg_gated, 256, 0, st>>>(km, kv, iin, ld, N, km_s, kv_s, w.gate);
        ctx->launches++;
    }
    if ((e = cudaGetLastError()) != cudaSuccess) return cuda_fail(e, "sort kernels");
    *ki_sorted = iin;
    return JACKS_OK;
}

// condense_normalised_counts on device buffers (preprocess.py:61-67); h_off / h_cols are host arrays (small).
int condense_core(JacksCtx* ctx, const double* d_lc, long long N, int C, const int32_t* h_off, const int32_t* h_cols, int S,
                  long long window, int monotonize, int max_rep, double* d_data, cudaStream_t st) {
    const long long ld = (N + 31) / 32 * 32;
    const size_t nidx = (size_t)h_off[S];
    Arena a;
    const size_t o_km = a.take((size_t)S * ld * 8), o_kv = a.take((size_t)S * ld * 8), o_ki = a.take((size_t)S * ld * 4);
    const size_t o_kms = a.take((size_t)S * ld * 8), o_kvs = a.take((size_t)S * ld * 8);
    const size_t o_P = a.take((size_t)S * (ld + 1) * 8), o_M = a.take((size_t)S * ld * 8);
    const size_t o_idx = a.take((S + 1 + nidx) * sizeof(int));
    size_t so[8];
    sort_scratch_bytes(S, ld, &a, so);
    int rc;
    if ((rc = ctx->pre.reserve(a.off)) != JACKS_OK) return rc;
    char* base = (char*)ctx->pre.p;
    double* km = (double*)(base + o_km); double* kv = (double*)(base + o_kv); int* ki = (int*)(base + o_ki);
    double* kms = (double*)(base + o_kms); double* kvs = (double*)(base + o_kvs);
    int* d_off = (int*)(base + o_idx);
    int* d_cols = d_off + S + 1;
    cudaError_t e;
    if ((e = cudaMemcpyAsync(d_off, h_off, (S + 1) * sizeof(int), cudaMemcpyHostToDevice, st)) != cudaSuccess) return cuda_fail(e, "H2D offsets");
    if ((e = cudaMemcpyAsync(d_cols, h_cols, nidx * sizeof(int), cudaMemcpyHostToDevice, st)) != cudaSuccess) return cuda_fail(e, "H2D cols");
    if (S >= 8) {
        dim3 tg((unsigned)((ld + 31) / 32), (unsigned)((S + 31) / 32));
        sample_stats_tiled_kernel<<<tg, 256, 0, st>>>(d_lc, N, C, d_off, d_cols, S, ld, d_data, km, kv, ki);
    } else {
        dim3 sg((unsigned)((ld + 255) / 256), (unsigned)S);
        sample_stats_kernel<<<sg, 256, 0, st>>>(d_lc, N, C, d_off, d_cols, S, ld, d_data, km, kv, ki);
    }
    ctx->launches++;
    if (max_rep >= 2) {
        SortScratch w{(unsigned long long*)(base + so[0]), (unsigned long long*)(base + so[1]), (int*)(base + so[2]), (int*)(base + so[3]),
                      (unsigned*)(base + so[4]), (int*)(base + so[5]), (TieRun*)(base + so[6]), (TieRun*)(base + so[7])};
        int* kis = nullptr;
        if ((rc = sort_mean_var(ctx, st, km, kv, S, ld, N, w, kms, kvs, &kis)) != JACKS_OK) return rc;
        smooth_monotone_scatter_kernel<<<S, kScanThreads, 0, st>>>(kvs, kis, ld, N, S, (int)window, monotonize ? 1 : 0, d_off,
                                                                   (double*)(base + o_P), (double*)(base + o_M), d_data);
        ctx->launches++;
    }
    if ((e = cudaGetLastError()) != cudaSuccess) return cuda_fail(e, "condense kernels");
    return JACKS_OK;
}

int check_samples(const int32_t* off, const int32_t* cols, int S, int n_cols, long long N, int window_override, long long* window, int* max_rep) {
    *max_rep = 0;
    for (int s = 0; s < S; ++s) {
        const int R = off[s + 1] - off[s];
        if (R < 1) return fail(JACKS_E_INVALID, "sample %d has no replicate columns", s);
        for (int r = off[s]; r < off[s + 1]; ++r)
            if (cols[r] < 0 || cols[r] >= n_cols) return fail(JACKS_E_INVALID, "column index out of range");
        *max_rep = R > *max_rep ? R : *max_rep;
    }
    // window = min(max(int(N/100), 30), 800), preprocess.py:35
    long long w = N / 100;
    if (w < 30) w = 30;
    if (w > 800) w = 800;
    if (window_override > 0) w = window_override;
    if (*max_rep >= 2 && N < w + 1)      // the reference's window_smooth indexes m[0..window]: IndexError below that
        return fail(JACKS_E_INVALID, "calc_posterior_sd needs at least %lld guides for its smoothing window (got %lld)", w + 1, N);
    *window = w;
    return JACKS_OK;
}

}  // namespace
}  // namespace jacks

extern "C" int jacks_normalize_device(JacksCtx* ctx, const double* values, int64_t n_guides, int32_t n_cols, int32_t take_log,
                                      double prior, const double* log2_table, int64_t table_len, int32_t normtype,
                                      const uint8_t* row_mask, double* out, void* stream) {
    if (!ctx || !values || !out) return fail(JACKS_E_INVALID, "jacks_normalize_device: NULL argument");
    if (n_guides < 1 || n_cols < 1) return fail(JACKS_E_INVALID, "empty matrix");
    if (normtype < JACKS_NORM_MEDIAN || normtype > JACKS_NORM_MODE) return fail(JACKS_E_INVALID, "Unrecognised normalisation type %d", normtype);
    if (normtype == JACKS_NORM_CTRL_GUIDES && !row_mask) return fail(JACKS_E_INVALID, "No guides specified for ctrl guide normalization");
    if (!log2_table || !take_log) table_len = 0;
    if (table_len < 0) return fail(JACKS_E_INVALID, "table_len < 0");
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    return normalize_core(ctx, values, n_guides, n_cols, take_log, prior, log2_table, table_len, normtype,
                          normtype == JACKS_NORM_CTRL_GUIDES ? row_mask : nullptr, out, (cudaStream_t)stream);
}

extern "C" int jacks_condense_device(JacksCtx* ctx, const double* logcounts, int64_t n_guides, int32_t n_cols,
                                     const int32_t* sample_col_offsets, const int32_t* sample_cols, int32_t n_samples,
                                     int32_t window_override, int32_t monotonize, double* data, void* stream) {
    if (!ctx || !logcounts || !sample_col_offsets || !sample_cols || !data) return fail(JACKS_E_INVALID, "jacks_condense_device: NULL argument");
    if (n_guides < 1 || n_cols < 1 || n_samples < 1) return fail(JACKS_E_INVALID, "empty input");
    long long window; int max_rep;
    int rc = check_samples(sample_col_offsets, sample_cols, n_samples, n_cols, n_guides, window_override, &window, &max_rep);
    if (rc != JACKS_OK) return rc;
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    return condense_core(ctx, logcounts, n_guides, n_cols, sample_col_offsets, sample_cols, n_samples, window, monotonize, max_rep, data,
                         (cudaStream_t)stream);
}

extern "C" int jacks_preprocess_host(JacksCtx* ctx, const double* counts, int64_t n_guides, int32_t n_cols, int32_t take_log,
                                     double prior, const double* log2_table, int64_t table_len, int32_t normtype,
                                     const uint8_t* row_mask, const int64_t* row_order, const int32_t* sample_col_offsets,
                                     const int32_t* sample_cols, int32_t n_samples, int32_t window_override, int32_t monotonize,
                                     double* logcounts_out, double* data) {
    if (!ctx || !counts || !sample_col_offsets || !sample_cols || !data) return fail(JACKS_E_INVALID, "jacks_preprocess_host: NULL argument");
    if (n_guides < 1 || n_cols < 1 || n_samples < 1) return fail(JACKS_E_INVALID, "empty input");
    if (normtype < JACKS_NORM_MEDIAN || normtype > JACKS_NORM_MODE) return fail(JACKS_E_INVALID, "Unrecognised normalisation type %d", normtype);
    if (normtype == JACKS_NORM_CTRL_GUIDES && !row_mask) return fail(JACKS_E_INVALID, "No guides specified for ctrl guide normalization");
    if (!log2_table || !take_log) table_len = 0;
    if (table_len < 0) return fail(JACKS_E_INVALID, "table_len < 0");
    const long long N = n_guides;
    const int C = n_cols, S = n_samples;
    long long window; int max_rep;
    int rc = check_samples(sample_col_offsets, sample_cols, S, C, N, window_override, &window, &max_rep);
    if (rc != JACKS_OK) return rc;
    if (row_order) for (long long i = 0; i < N; ++i) if (row_order[i] < 0 || row_order[i] >= N) return fail(JACKS_E_INVALID, "row_order out of range");
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    const size_t mb = (size_t)N * C * sizeof(double), db = (size_t)N * S * 2 * sizeof(double);
    if ((rc = ctx->out[0].reserve(mb)) != JACKS_OK) return rc;                        // counts, later the row-sorted log counts
    if ((rc = ctx->out[1].reserve(mb)) != JACKS_OK) return rc;                        // normalised log counts
    if ((rc = ctx->out[2].reserve((size_t)table_len * sizeof(double) + 8)) != JACKS_OK) return rc;
    if ((rc = ctx->out[3].reserve((size_t)N * 9 + 16)) != JACKS_OK) return rc;        // row order (8 B) + row mask (1 B)
    if ((rc = ctx->out[4].reserve(db)) != JACKS_OK) return rc;
    cudaStream_t st = ctx->stream;
    if ((rc = ctx->stager.h2d(ctx->out[0].p, counts, mb, st)) != JACKS_OK) return rc;          // the ONE upload
    if (table_len && (e = cudaMemcpyAsync(ctx->out[2].p, log2_table, (size_t)table_len * sizeof(double), cudaMemcpyHostToDevice, st)) != cudaSuccess) return cuda_fail(e, "H2D log2 table");
    long long* d_order = (long long*)ctx->out[3].p;
    unsigned char* d_mask = (unsigned char*)(d_order + N);
    if (normtype == JACKS_NORM_CTRL_GUIDES && (e = cudaMemcpyAsync(d_mask, row_mask, (size_t)N, cudaMemcpyHostToDevice, st)) != cudaSuccess) return cuda_fail(e, "H2D row mask");
    if ((rc = normalize_core(ctx, (const double*)ctx->out[0].p, N, C, take_log, prior, (const double*)ctx->out[2].p, table_len, normtype,
                             normtype == JACKS_NORM_CTRL_GUIDES ? d_mask : nullptr, (double*)ctx->out[1].p, st)) != JACKS_OK) return rc;
    const double* d_lc = (const double*)ctx->out[1].p;
    if (row_order) {                      // the loader's sort by guide id (preprocess.py:154-157), after the normalisation
        if ((e = cudaMemcpyAsync(d_order, row_order, (size_t)N * sizeof(long long), cudaMemcpyHostToDevice, st)) != cudaSuccess) return cuda_fail(e, "H2D row order");
        permute_rows_kernel<<<(unsigned)N, 256, 0, st>>>(d_lc, d_order, N, C, (double*)ctx->out[0].p);
        ctx->launches++;
        d_lc = (const double*)ctx->out[0].p;
    }
    if ((rc = condense_core(ctx, d_lc, N, C, sample_col_offsets, sample_cols, S, window, monotonize, max_rep, (double*)ctx->out[4].p, st)) != JACKS_OK) return rc;
    if (logcounts_out && (rc = ctx->stager.d2h(logcounts_out, d_lc, mb, st)) != JACKS_OK) return rc;
    if ((rc = ctx->stager.d2h(data, ctx->out[4].p, db, st)) != JACKS_OK) return rc;
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return cuda_fail(e, "preprocess kernels");
    return JACKS_OK;
}

extern "C" int jacks_normalize_host(JacksCtx* ctx, const double* values, int64_t n_guides, int32_t n_cols, int32_t take_log,
                                    double prior, const double* log2_table, int64_t table_len, int32_t normtype,
                                    const uint8_t* row_mask, double* out) {
    if (!ctx || !values || !out) return fail(JACKS_E_INVALID, "jacks_normalize_host: NULL argument");
    if (n_guides < 1 || n_cols < 1) return fail(JACKS_E_INVALID, "empty matrix");
    if (normtype < JACKS_NORM_MEDIAN || normtype > JACKS_NORM_MODE) return fail(JACKS_E_INVALID, "Unrecognised normalisation type %d", normtype);
    if (normtype == JACKS_NORM_CTRL_GUIDES && !row_mask) return fail(JACKS_E_INVALID, "No guides specified for ctrl guide normalization");
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    const long long N = n_guides;
    const int C = n_cols;
    const size_t mb = (size_t)N * C * sizeof(double);
    if (!log2_table || !take_log) table_len = 0;
    if (table_len < 0) return fail(JACKS_E_INVALID, "table_len < 0");
    int rc;
    if ((rc = ctx->out[0].reserve(mb)) != JACKS_OK) return rc;
    if ((rc = ctx->out[1].reserve(mb)) != JACKS_OK) return rc;
    if ((rc = ctx->out[2].reserve((size_t)table_len * sizeof(double) + 8)) != JACKS_OK) return rc;
    if ((rc = ctx->out[3].reserve((size_t)N + 8)) != JACKS_OK) return rc;
    cudaStream_t st = ctx->stream;
    if ((rc = ctx->stager.h2d(ctx->out[0].p, values, mb, st)) != JACKS_OK) return rc;
    if (table_len && (e = cudaMemcpyAsync(ctx->out[2].p, log2_table, (size_t)table_len * sizeof(double), cudaMemcpyHostToDevice, st)) != cudaSuccess) return cuda_fail(e, "H2D log2 table");
    const unsigned char* d_mask = nullptr;
    if (normtype == JACKS_NORM_CTRL_GUIDES) {
        if ((e = cudaMemcpyAsync(ctx->out[3].p, row_mask, (size_t)N, cudaMemcpyHostToDevice, st)) != cudaSuccess) return cuda_fail(e, "H2D row mask");
        d_mask = (const unsigned char*)ctx->out[3].p;
    }
    if ((rc = normalize_core(ctx, (const double*)ctx->out[0].p, N, C, take_log, prior, (const double*)ctx->out[2].p, table_len, normtype, d_mask,
                             (double*)ctx->out[1].p, st)) != JACKS_OK) return rc;
    if ((rc = ctx->stager.d2h(out, ctx->out[1].p, mb, st)) != JACKS_OK) return rc;
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return cuda_fail(e, "normalize kernels");
    return JACKS_OK;
}

extern "C" int jacks_condense_host(JacksCtx* ctx, const double* logcounts, int64_t n_guides, int32_t n_cols,
                                   const int32_t* sample_col_offsets, const int32_t* sample_cols, int32_t n_samples,
                                   int32_t window_override, int32_t monotonize, double* data) {
    if (!ctx || !logcounts || !sample_col_offsets || !sample_cols || !data) return fail(JACKS_E_INVALID, "jacks_condense_host: NULL argument");
    if (n_guides < 1 || n_cols < 1 || n_samples < 1) return fail(JACKS_E_INVALID, "empty input");
    const long long N = n_guides;
    const int S = n_samples;
    long long window; int max_rep;
    int rc = check_samples(sample_col_offsets, sample_cols, S, n_cols, N, window_override, &window, &max_rep);
    if (rc != JACKS_OK) return rc;
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    const size_t lb = (size_t)N * n_cols * sizeof(double), db = (size_t)N * S * 2 * sizeof(double);
    if ((rc = ctx->out[0].reserve(lb)) != JACKS_OK) return rc;
    if ((rc = ctx->out[1].reserve(db)) != JACKS_OK) return rc;
    cudaStream_t st = ctx->stream;
    if ((rc = ctx->stager.h2d(ctx->out[0].p, logcounts, lb, st)) != JACKS_OK) return rc;
    if ((rc = condense_core(ctx, (const double*)ctx->out[0].p, N, n_cols, sample_col_offsets, sample_cols, S, window, monotonize, max_rep,
                            (double*)ctx->out[1].p, st)) != JACKS_OK) return rc;
    if ((rc = ctx->stager.d2h(data, ctx->out[1].p, db, st)) != JACKS_OK) return rc;
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return cuda_fail(e, "condense kernels");
    return JACKS_OK;
}

extern "C" int jacks_posterior_sd_subset_host(JacksCtx* ctx, const double* data, int64_t n_guides, int32_t n_rep,
                                              const uint8_t* in_set, int32_t window, int32_t monotonize, double* sd) {
    if (!ctx || !data || !in_set || !sd) return fail(JACKS_E_INVALID, "jacks_posterior_sd_subset_host: NULL argument");
    if (n_guides < 1 || n_rep < 2) return fail(JACKS_E_INVALID, "need at least one guide and two columns");
    const long long N = n_guides;
    long long members = 0;
    for (long long i = 0; i < N; ++i) members += in_set[i] ? 1 : 0;
    if (window < 1) return fail(JACKS_E_INVALID, "window < 1");
    if (members < (long long)window + 1)
        return fail(JACKS_E_INVALID, "calc_posterior_sd needs at least %d subset guides for its smoothing window (got %lld)", window + 1, members);
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    const long long Np = (N + 31) / 32 * 32;
    int rc;
    const size_t lb = (size_t)N * n_rep * sizeof(double);
    if ((rc = ctx->out[0].reserve(lb)) != JACKS_OK) return rc;
    if ((rc = ctx->out[1].reserve((size_t)N * 2 * sizeof(double))) != JACKS_OK) return rc;       // data plane of sample_stats
    if ((rc = ctx->out[2].reserve((size_t)Np * 2 * sizeof(double))) != JACKS_OK) return rc;      // means: keys, sorted
    if ((rc = ctx->out[3].reserve((size_t)Np * 2 * sizeof(double))) != JACKS_OK) return rc;      // variances: keys, sorted
    if ((rc = ctx->out[4].reserve((size_t)Np * sizeof(int))) != JACKS_OK) return rc;
    if ((rc = ctx->out[5].reserve((size_t)(N + 1) * 4 * sizeof(double))) != JACKS_OK) return rc;  // sub_mean, sub_var, P, M
    if ((rc = ctx->out[6].reserve((size_t)N * sizeof(double))) != JACKS_OK) return rc;            // sd
    if ((rc = ctx->out[7].reserve((size_t)(2 + n_rep) * sizeof(int) + (size_t)N + 16)) != JACKS_OK) return rc;
    Arena ar;
    size_t so[8];
    sort_scratch_bytes(1, Np, &ar, so);
    if ((rc = ctx->pre.reserve(ar.off)) != JACKS_OK) return rc;
    cudaStream_t st = ctx->stream;
    std::vector<int> h(2 + n_rep);
    h[0] = 0; h[1] = n_rep;
    for (int r = 0; r < n_rep; ++r) h[2 + r] = r;
    int* d_off = (int*)ctx->out[7].p;
    int* d_cols = d_off + 2;
    unsigned char* d_set = (unsigned char*)(d_cols + n_rep);
    if ((rc = ctx->stager.h2d(ctx->out[0].p, data, lb, st)) != JACKS_OK) return rc;
    if ((e = cudaMemcpyAsync(d_off, h.data(), h.size() * sizeof(int), cudaMemcpyHostToDevice, st)) != cudaSuccess) return cuda_fail(e, "H2D index");
    if ((e = cudaMemcpyAsync(d_set, in_set, (size_t)N, cudaMemcpyHostToDevice, st)) != cudaSuccess) return cuda_fail(e, "H2D subset");
    double* km0 = (double*)ctx->out[2].p; double* kv0 = (double*)ctx->out[3].p;
    double* km = km0 + Np; double* kv = kv0 + Np;
    int* ki = nullptr;
    dim3 sg((unsigned)((Np + 255) / 256), 1);
    sample_stats_kernel<<<sg, 256, 0, st>>>((const double*)ctx->out[0].p, N, n_rep, d_off, d_cols, 1, Np, (double*)ctx->out[1].p, km0, kv0, (int*)ctx->out[4].p);
    ctx->launches++;
    {
        char* base = (char*)ctx->pre.p;
        SortScratch w{(unsigned long long*)(base + so[0]), (unsigned long long*)(base + so[1]), (int*)(base + so[2]), (int*)(base + so[3]),
                      (unsigned*)(base + so[4]), (int*)(base + so[5]), (TieRun*)(base + so[6]), (TieRun*)(base + so[7])};
        if ((rc = sort_mean_var(ctx, st, km0, kv0, 1, Np, N, w, km, kv, &ki)) != JACKS_OK) return rc;
    }
    double* scratch = (double*)ctx->out[5].p;
    subset_sd_kernel<<<1, kScanThreads, 0, st>>>(km, kv, ki, N, d_set, window, monotonize ? 1 : 0, scratch, scratch + (N + 1),
                                                  scratch + 2 * (N + 1), scratch + 3 * (N + 1), (double*)ctx->out[6].p);
    ctx->launches++;
    if ((e = cudaGetLastError()) != cudaSuccess) return cuda_fail(e, "subset sd kernels");
    if ((rc = ctx->stager.d2h(sd, ctx->out[6].p, (size_t)N * sizeof(double), st)) != JACKS_OK) return rc;
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return cuda_fail(e, "subset sd kernels");
    return JACKS_OK;
}
