// Generic batched EM kernel: any guide count G and line count L, NaN-safe, one warp per gene.
//
// This is the catch-all behind the fast kernels (em_fast.cuh): it runs genes whose shape has no
// fast instantiation (G > 12, or a line count too large for the fast variants) and genes with missing measurements (NaN in y), following
// numpy's NaN rules as the reference uses them:
//   sums are nansum (NaN terms count as 0)                   infer.py:39,42,44,125
//   w1 init and the x normaliser use nanmedian               infer.py:81,108
//   x1.mean()/max()/min() are NOT NaN-aware                  infer.py:108
//   updateTau propagates NaN                                 infer.py:119-120
// The working set (y, tau_pr_den, tau: [G][Lp]; w1, w2: [Lp]; x1, x2, S1, S2: [G]) lives in
// shared memory when it fits and in a global scratch slab otherwise; the code addresses it
// through generic pointers so both cases share one body.
#include "em_common.cuh"

namespace jacks {

#define JACKS_FULL_MASK 0xffffffffu

__device__ __forceinline__ double gwarp_allsum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(JACKS_FULL_MASK, v, o);
    return v;
}
__device__ __forceinline__ double nz(double v) { return (v == v) ? v : 0.0; }
__device__ __forceinline__ double dnan() { return __longlong_as_double(0x7ff8000000000000LL); }

__host__ __device__ long long em_generic_work_doubles(int G, int L) {
    const long long Lp = ((long long)L + 31) / 32 * 32;
    return (3LL * G + 2) * Lp + 4LL * G;
}

// Order-preserving map of a (non-NaN) double onto unsigned 64-bit keys.
__device__ __forceinline__ unsigned long long dkey(double x) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ULL);
}
__device__ __forceinline__ double dunkey(unsigned long long k) {
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffULL) : ~k;
    return __longlong_as_double((long long)b);
}

// nanmedian of n values at v[0], v[stride], ... computed by ONE thread (numpy semantics).
// Small n: rank counting, O(n^2) compares.  Large n (a gene with hundreds of guides, e.g. a library's non-targeting
// controls under one name): the k-th smallest key is built bit by bit from the top, 64 counting passes -- O(64 n)
// instead of O(n^2), which at n = 1,000 took 1.9 s per gene.
__device__ double nanmedian_strided(const double* v, int n, long long stride) {
    int valid = 0;
    for (int i = 0; i < n; ++i) { const double x = v[i * stride]; valid += (x == x); }
    if (valid == 0) return dnan();
    const int k_lo = (valid - 1) / 2, k_hi = valid / 2;
    if (n > 48) {
        unsigned long long prefix = 0;
        for (int bit = 63; bit >= 0; --bit) {
            const unsigned long long cand = prefix | (1ULL << bit);
            int below = 0;                                   // keys < cand
            for (int i = 0; i < n; ++i) { const double x = v[i * stride]; below += (x == x) && (dkey(x) < cand); }
            if (below <= k_lo) prefix = cand;                // the k_lo-th smallest (0-based) has this bit set
        }
        const double lo = dunkey(prefix);
        if (k_hi == k_lo) return lo;
        int le = 0;
        unsigned long long next = ~0ULL;
        for (int i = 0; i < n; ++i) {
            const double x = v[i * stride];
            if (!(x == x)) continue;
            const unsigned long long k = dkey(x);
            le += (k <= prefix);
            if (k > prefix && k < next) next = k;
        }
        const double hi = (le > k_hi) ? lo : dunkey(next);
        return 0.5 * (lo + hi);
    }
    double lo = 0.0, hi = 0.0;
    for (int i = 0; i < n; ++i) {
        const double xi = v[i * stride];
        if (!(xi == xi)) continue;
        int r = 0;
        for (int j = 0; j < n; ++j) {
            const double xj = v[j * stride];
            if (!(xj == xj) || j == i) continue;
            r += (j < i) ? (xj <= xi) : (xj < xi);
        }
        if (r == k_lo) lo = xi;
        if (r == k_hi) hi = xi;
    }
    return (k_lo == k_hi) ? lo : 0.5 * (lo + hi);
}

__global__ void __launch_bounds__(256) em_generic_kernel(const EmLaunch a, const int warps_per_cta,
                                                        const long long smem_doubles_per_warp) {
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    // per-line mode (em_common.cuh): one line of the cubes per work item, outputs carry a line axis
    const int L = a.lines_mode ? 1 : a.L;
    const long long row_stride = a.lines_mode ? a.L_full : a.L;      // lines per cube / output row
    const long long Lp = ((long long)L + 31) / 32 * 32;
    const double tps = a.tps;
    const double c_tau = a.tps + 0.5;
    const double ivx = 1.0 / a.var0_x;
    const double mx = a.mu0_x / a.var0_x;
    const bool fixed = (a.fixed_x1 != nullptr);
    const long long n_work = (a.n_work_dev ? (long long)*a.n_work_dev : (long long)a.n_work) *
                             (a.lines_mode == 1 ? a.L_full : 1);
    const long long gwarp = (long long)blockIdx.x * warps_per_cta + warp;

    for (;;) {
        int slot = 0;
        if (lane == 0) slot = atomicAdd(a.work_counter, 1);
        slot = __shfl_sync(JACKS_FULL_MASK, slot, 0);
        if (slot >= n_work) break;
        int gid, line = 0;
        if (a.lines_mode == 1) { const int s = slot / a.L_full; line = slot - s * a.L_full; gid = a.gene_ids[s]; }
        else if (a.lines_mode == 2) { const int pid = a.gene_ids[slot]; gid = pid / a.L_full; line = pid - gid * a.L_full; }
        else gid = a.gene_ids[slot];
        const long long off = a.gene_offsets[gid];
        const int G = (int)(a.gene_offsets[gid + 1] - off);

        double* wk = (em_generic_work_doubles(G, L) <= smem_doubles_per_warp)
                         ? (smem + (long long)warp * smem_doubles_per_warp)
                         : (a.scratch + gwarp * a.scratch_stride);
        double* ys = wk;
        double* tpds = ys + (long long)G * Lp;
        double* taus = tpds + (long long)G * Lp;
        double* w1s = taus + (long long)G * Lp;
        double* w2s = w1s + Lp;
        double* x1s = w2s + Lp;
        double* x2s = x1s + G;
        double* S1s = x2s + G;
        double* S2s = S1s + G;

        // ---- set-up, infer.py:58-70 ----------------------------------------------------------
        for (int g = 0; g < G; ++g) {
            const int row = a.guide_rows[off + g];
            const double2* tr = a.test + (long long)row * row_stride + line;
            const double2* cr = a.ctrl + (long long)row * a.ctrl_lines + (a.ctrl_lines == 1 ? 0 : line);
            double2 c0 = make_double2(0.0, 0.0);
            double fill = 0.0;
            if (a.ctrl_lines == 1) {
                c0 = cr[0];
                if (a.ctrl_rowmean_fill && c0.y != c0.y) {
                    double s = 0.0;
                    for (int l = lane; l < L; l += 32) { const double e = tr[l].y; s += (e != e) ? 2.0 : e; }
                    fill = gwarp_allsum(s) / (double)L;
                }
            }
            for (int l = lane; l < L; l += 32) {
                const double2 t = tr[l];
                const double2 c = (a.ctrl_lines == 1) ? c0 : cr[l];
                const double te = (t.y != t.y) ? 2.0 : t.y;
                double ce = c.y;
                if (ce != ce) ce = (a.ctrl_lines == 1 && a.ctrl_rowmean_fill) ? fill : te;
                ys[g * Lp + l] = t.x - c.x;
                tpds[g * Lp + l] = tps * (te * te + ce * ce + 1e-2);
            }
            if (lane == 0) {
                if (fixed) { x1s[g] = a.fixed_x1[row]; x2s[g] = a.fixed_x2[row]; }
                else { x1s[g] = a.mu0_x; x2s[g] = a.mu0_x * a.mu0_x; }
            }
        }
        __syncwarp();

        // ---- init, infer.py:81-86 ------------------------------------------------------------
        for (int l = lane; l < L; l += 32) {
            const double w1 = nanmedian_strided(ys + l, G, Lp);
            w1s[l] = w1;
            w2s[l] = w1 * w1;
        }
        double bound_p = 0.0;
        for (int g = 0; g < G; ++g) {
            const double x1g = x1s[g], x2g = x2s[g];
            double s1 = 0.0, s2 = 0.0;
            for (int l = lane; l < L; l += 32) {
                const double yv = ys[g * Lp + l];
                const double t = tps / tpds[g * Lp + l];
                taus[g * Lp + l] = t;
                const double w1 = w1s[l], w2 = w2s[l];
                const double xw = x1g * w1;
                bound_p += nz(t * (yv * yv + x2g * w2 - 2.0 * xw * yv));
                s1 += nz(yv * t * w1);
                s2 += nz(t * w2);
            }
            s1 = gwarp_allsum(s1);
            s2 = gwarp_allsum(s2);
            if (lane == 0) { S1s[g] = s1; S2s[g] = s2; }
        }
        double bound = gwarp_allsum(bound_p);
        double mu0_w = a.mu0_w, var0_w = a.var0_w;
        int iters = 0;
        __syncwarp();

        for (int it = 0; it < a.n_iter; ++it) {
            const double last = bound;
            if (!fixed) {
                // X step, infer.py:103-110
                for (int g = lane; g < G; g += 32) {
                    const double den = S2s[g] + ivx;
                    const double x1u = (mx + S1s[g]) / den;
                    x1s[g] = x1u;
                    x2s[g] = x1u * x1u + 1.0 / den;
                }
                __syncwarp();
                double sum = 0.0, hi = x1s[0], lo = x1s[0];
                bool anynan = false;
                for (int g = 0; g < G; ++g) {
                    const double v = x1s[g];
                    sum += v;
                    anynan |= !(v == v);
                    hi = fmax(hi, v); lo = fmin(lo, v);
                }
                if (anynan) { hi = dnan(); lo = dnan(); }      // ndarray.max/min propagate NaN
                // nanmedian of x1 by distributed rank selection
                int valid = 0;
                for (int g = 0; g < G; ++g) { const double v = x1s[g]; valid += (v == v); }
                double med;
                if (valid == 0) med = dnan();
                else {
                    const int k_lo = (valid - 1) / 2, k_hi = valid / 2;
                    double plo = 0.0, phi = 0.0;
                    for (int i = lane; i < G; i += 32) {
                        const double xi = x1s[i];
                        if (!(xi == xi)) continue;
                        int r = 0;
                        for (int j = 0; j < G; ++j) {
                            const double xj = x1s[j];
                            if (!(xj == xj) || j == i) continue;
                            r += (j < i) ? (xj <= xi) : (xj < xi);
                        }
                        if (r == k_lo) plo = xi;
                        if (r == k_hi) phi = xi;
                    }
                    plo = gwarp_allsum(plo);
                    phi = gwarp_allsum(phi);
                    med = (k_lo == k_hi) ? plo : 0.5 * (plo + phi);
                }
                const double wadj = 0.5 / G;
                const double x1m = sum / G + 2 * wadj * med - wadj * hi - wadj * lo;
                __syncwarp();
                for (int g = lane; g < G; g += 32) {
                    x1s[g] = x1s[g] / x1m;
                    x2s[g] = x2s[g] / x1m / x1m;
                }
                __syncwarp();
            }
            if (a.apply_w_hp && L > 1) {
                // infer.py:92: plain (NaN-propagating) mean and population variance of w1
                double s = 0.0;
                for (int l = lane; l < L; l += 32) s += w1s[l];
                const double mean = gwarp_allsum(s) / (double)L;
                double v = 0.0;
                for (int l = lane; l < L; l += 32) { const double d = w1s[l] - mean; v += d * d; }
                mu0_w = mean;
                var0_w = (gwarp_allsum(v) / (double)L) * 3 + 1e-4;
            }
            const double ivw = 1.0 / var0_w;
            const double mw = mu0_w / var0_w;
            // W step, infer.py:112-116
            for (int l = lane; l < L; l += 32) {
                double sa = 0.0, sb = 0.0;
                for (int g = 0; g < G; ++g) {
                    const double t = taus[g * Lp + l];
                    sa += nz(x2s[g] * t);
                    sb += nz(x1s[g] * (ys[g * Lp + l] * t));
                }
                const double den = sa + ivw;
                const double w1 = (mw + sb) / den;
                w1s[l] = w1;
                w2s[l] = w1 * w1 + 1.0 / den;
            }
            // tau step (infer.py:118-121), bound (infer.py:123-125), sums for the next X step
            bound_p = 0.0;
            for (int g = 0; g < G; ++g) {
                const double x1g = x1s[g], x2g = x2s[g];
                double s1 = 0.0, s2 = 0.0;
                for (int l = lane; l < L; l += 32) {
                    const double yv = ys[g * Lp + l];
                    const double w1 = w1s[l], w2 = w2s[l];
                    const double xw = x1g * w1;
                    const double xw2 = x2g * w2;
                    const double y2 = yv * yv;
                    const double bstar = y2 - 2 * yv * xw + xw2;
                    const double t = c_tau / (tpds[g * Lp + l] + 0.5 * bstar);
                    taus[g * Lp + l] = t;
                    bound_p += nz(t * (y2 + xw2 - 2 * xw * yv));
                    s1 += nz(yv * t * w1);
                    s2 += nz(t * w2);
                }
                s1 = gwarp_allsum(s1);
                s2 = gwarp_allsum(s2);
                if (lane == 0) { S1s[g] = s1; S2s[g] = s2; }
            }
            bound = gwarp_allsum(bound_p);
            __syncwarp();
            ++iters;
            if (fabs(last - bound) < a.tol) break;
        }

        // ---- results ---------------------------------------------------------------------------
        for (int l = lane; l < L; l += 32) {
            a.w1[(long long)gid * row_stride + line + l] = w1s[l];
            if (a.w2) a.w2[(long long)gid * row_stride + line + l] = w2s[l];
        }
        if (a.tau || a.y) {
            for (int g = 0; g < G; ++g)
                for (int l = lane; l < L; l += 32) {
                    if (a.tau) a.tau[(off + g) * row_stride + line + l] = taus[g * Lp + l];
                    if (a.y) a.y[(off + g) * row_stride + line + l] = ys[g * Lp + l];
                }
        }
        for (int g = lane; g < G; g += 32) {
            const long long q = a.lines_mode ? (off + g) * row_stride + line : off + g;
            if (a.x1) a.x1[q] = x1s[g];
            if (a.x2) a.x2[q] = x2s[g];
        }
        if (lane == 0) {
            const long long q = a.lines_mode ? (long long)gid * row_stride + line : gid;
            if (a.n_iters) a.n_iters[q] = iters;
            if (a.bound) a.bound[q] = bound;
        }
        __syncwarp();
    }
}

// ---- big genes: one CTA per gene -------------------------------------------------------------------------------------
// A gene with hundreds of guides (a library's ~1,000 non-targeting controls under one gene name is the usual case) is a
// 400,000-element problem; on ONE warp of the kernel above its passes crawl through global scratch at memory latency
// (110 ms for G = 1,000, L = 400 -- longer than the rest of the screen).  Here the same arithmetic and NaN rules run on a
// whole CTA: warps take guides in turn wherever the warp kernel loops over guides, threads take lines wherever it loops
// over lines, and the few cross-gene quantities are block reductions with a fixed summation order.  The working set
// always lives in a per-CTA global slab (L2).

constexpr int kBigThreads = 512;

struct BlockRed {
    double* red;     // [32] shared
    __device__ double sum(double v) {
        v = gwarp_allsum(v);
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
        __syncthreads();
        if (lane == 0) red[warp] = v;
        __syncthreads();
        double s = 0.0;
        for (int w = 0; w < nw; ++w) s += red[w];
        return s;
    }
    __device__ double max(double v) {      // NaN-free inputs expected (callers handle NaN separately)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(JACKS_FULL_MASK, v, o));
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
        __syncthreads();
        if (lane == 0) red[warp] = v;
        __syncthreads();
        double s = red[0];
        for (int w = 1; w < nw; ++w) s = fmax(s, red[w]);
        return s;
    }
};

__global__ void __launch_bounds__(kBigThreads) em_big_kernel(const EmLaunch a, const long long slab_doubles) {
    __shared__ double s_red[32];
    __shared__ int s_slot;
    BlockRed br{s_red};
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
    const int L = a.L;
    const long long Lp = ((long long)L + 31) / 32 * 32;
    const double tps = a.tps;
    const double c_tau = a.tps + 0.5;
    const double ivx = 1.0 / a.var0_x;
    const double mx = a.mu0_x / a.var0_x;
    const bool fixed = (a.fixed_x1 != nullptr);

    for (;;) {
        __syncthreads();
        if (tid == 0) s_slot = atomicAdd(a.work_counter, 1);
        __syncthreads();
        const int slot = s_slot;
        if (slot >= a.n_work) break;
        const int gid = a.gene_ids[slot];
        const long long off = a.gene_offsets[gid];
        const int G = (int)(a.gene_offsets[gid + 1] - off);
        double* wk = a.scratch + (long long)blockIdx.x * slab_doubles;
        double* ys = wk;
        double* tpds = ys + (long long)G * Lp;
        double* taus = tpds + (long long)G * Lp;
        double* w1s = taus + (long long)G * Lp;
        double* w2s = w1s + Lp;
        double* x1s = w2s + Lp;
        double* x2s = x1s + G;
        double* S1s = x2s + G;
        double* S2s = S1s + G;

        // ---- set-up, infer.py:58-70: one warp per guide --------------------------------------------------------------
        for (int g = warp; g < G; g += nwarps) {
            const int row = a.guide_rows[off + g];
            const double2* tr = a.test + (long long)row * L;
            const double2* cr = a.ctrl + (long long)row * a.ctrl_lines;
            double2 c0 = make_double2(0.0, 0.0);
            double fill = 0.0;
            if (a.ctrl_lines == 1) {
                c0 = cr[0];
                if (a.ctrl_rowmean_fill && c0.y != c0.y) {
                    double s = 0.0;
                    for (int l = lane; l < L; l += 32) { const double e = tr[l].y; s += (e != e) ? 2.0 : e; }
                    fill = gwarp_allsum(s) / (double)L;
                }
            }
            for (int l = lane; l < L; l += 32) {
                const double2 t = tr[l];
                const double2 c = (a.ctrl_lines == 1) ? c0 : cr[l];
                const double te = (t.y != t.y) ? 2.0 : t.y;
                double ce = c.y;
                if (ce != ce) ce = (a.ctrl_lines == 1 && a.ctrl_rowmean_fill) ? fill : te;
                ys[g * Lp + l] = t.x - c.x;
                tpds[g * Lp + l] = tps * (te * te + ce * ce + 1e-2);
            }
            if (lane == 0) {
                if (fixed) { x1s[g] = a.fixed_x1[row]; x2s[g] = a.fixed_x2[row]; }
                else { x1s[g] = a.mu0_x; x2s[g] = a.mu0_x * a.mu0_x; }
            }
        }
        __syncthreads();

        // ---- init, infer.py:81-86: one thread per line for the median, one warp per guide for the sums -----------------
        for (int l = tid; l < L; l += T) {
            const double w1 = nanmedian_strided(ys + l, G, Lp);
            w1s[l] = w1;
            w2s[l] = w1 * w1;
        }
        __syncthreads();
        double bound_p = 0.0;
        for (int g = warp; g < G; g += nwarps) {
            const double x1g = x1s[g], x2g = x2s[g];
            double s1 = 0.0, s2 = 0.0;
            for (int l = lane; l < L; l += 32) {
                const double yv = ys[g * Lp + l];
                const double t = tps / tpds[g * Lp + l];
                taus[g * Lp + l] = t;
                const double w1 = w1s[l], w2 = w2s[l];
                const double xw = x1g * w1;
                bound_p += nz(t * (yv * yv + x2g * w2 - 2.0 * xw * yv));
                s1 += nz(yv * t * w1);
                s2 += nz(t * w2);
            }
            s1 = gwarp_allsum(s1);
            s2 = gwarp_allsum(s2);
            if (lane == 0) { S1s[g] = s1; S2s[g] = s2; }
        }
        double bound = br.sum(bound_p);
        double mu0_w = a.mu0_w, var0_w = a.var0_w;
        int iters = 0;
        __syncthreads();

        for (int it = 0; it < a.n_iter; ++it) {
            const double last = bound;
            if (!fixed) {
                // X step, infer.py:103-110
                for (int g = tid; g < G; g += T) {
                    const double den = S2s[g] + ivx;
                    const double x1u = (mx + S1s[g]) / den;
                    x1s[g] = x1u;
                    x2s[g] = x1u * x1u + 1.0 / den;
                }
                __syncthreads();
                double psum = 0.0, phi_ = -1.79769313486231570e308, plo_ = 1.79769313486231570e308;
                int pnan = 0, pvalid = 0;
                for (int g = tid; g < G; g += T) {
                    const double v = x1s[g];
                    psum += v;                              // plain sum: NaN propagates like ndarray.mean
                    if (v == v) { phi_ = fmax(phi_, v); plo_ = fmin(plo_, v); ++pvalid; } else ++pnan;
                }
                const double sum = br.sum(psum);
                const int valid = (int)(br.sum((double)pvalid) + 0.5);
                const bool anynan = br.sum((double)pnan) > 0.5;
                double hi = br.max(phi_), lo = -br.max(-plo_);
                if (anynan) { hi = dnan(); lo = dnan(); }      // ndarray.max/min propagate NaN
                double med;
                if (valid == 0) med = dnan();
                else {
                    // nanmedian of x1 by rank selection distributed over the threads
                    const int k_lo = (valid - 1) / 2, k_hi = valid / 2;
                    double plo = 0.0, phi = 0.0;
                    for (int i = tid; i < G; i += T) {
                        const double xi = x1s[i];
                        if (!(xi == xi)) continue;
                        int r = 0;
                        for (int j = 0; j < G; ++j) {
                            const double xj = x1s[j];
                            if (!(xj == xj) || j == i) continue;
                            r += (j < i) ? (xj <= xi) : (xj < xi);
                        }
                        if (r == k_lo) plo = xi;
                        if (r == k_hi) phi = xi;
                    }
                    plo = br.sum(plo);
                    phi = br.sum(phi);
                    med = (k_lo == k_hi) ? plo : 0.5 * (plo + phi);
                }
                const double wadj = 0.5 / G;
                const double x1m = sum / G + 2 * wadj * med - wadj * hi - wadj * lo;
                __syncthreads();
                for (int g = tid; g < G; g += T) {
                    x1s[g] = x1s[g] / x1m;
                    x2s[g] = x2s[g] / x1m / x1m;
                }
                __syncthreads();
            }
            if (a.apply_w_hp && L > 1) {
                // infer.py:92: plain (NaN-propagating) mean and population variance of w1
                double s = 0.0;
                for (int l = tid; l < L; l += T) s += w1s[l];
                const double mean = br.sum(s) / (double)L;
                double v = 0.0;
                for (int l = tid; l < L; l += T) { const double d = w1s[l] - mean; v += d * d; }
                mu0_w = mean;
                var0_w = (br.sum(v) / (double)L) * 3 + 1e-4;
            }
            const double ivw = 1.0 / var0_w;
            const double mw = mu0_w / var0_w;
            __syncthreads();
            // W step, infer.py:112-116: one thread per line
            for (int l = tid; l < L; l += T) {
                double sa = 0.0, sb = 0.0;
                for (int g = 0; g < G; ++g) {
                    const double t = taus[g * Lp + l];
                    sa += nz(x2s[g] * t);
                    sb += nz(x1s[g] * (ys[g * Lp + l] * t));
                }
                const double den = sa + ivw;
                const double w1 = (mw + sb) / den;
                w1s[l] = w1;
                w2s[l] = w1 * w1 + 1.0 / den;
            }
            __syncthreads();
            // tau step (infer.py:118-121), bound (infer.py:123-125), sums for the next X step: one warp per guide
            bound_p = 0.0;
            for (int g = warp; g < G; g += nwarps) {
                const double x1g = x1s[g], x2g = x2s[g];
                double s1 = 0.0, s2 = 0.0;
                for (int l = lane; l < L; l += 32) {
                    const double yv = ys[g * Lp + l];
                    const double w1 = w1s[l], w2 = w2s[l];
                    const double xw = x1g * w1;
                    const double xw2 = x2g * w2;
                    const double y2 = yv * yv;
                    const double bstar = y2 - 2 * yv * xw + xw2;
                    const double t = c_tau / (tpds[g * Lp + l] + 0.5 * bstar);
                    taus[g * Lp + l] = t;
                    bound_p += nz(t * (y2 + xw2 - 2 * xw * yv));
                    s1 += nz(yv * t * w1);
                    s2 += nz(t * w2);
                }
                s1 = gwarp_allsum(s1);
                s2 = gwarp_allsum(s2);
                if (lane == 0) { S1s[g] = s1; S2s[g] = s2; }
            }
            bound = br.sum(bound_p);
            __syncthreads();
            ++iters;
            if (fabs(last - bound) < a.tol) break;
        }

        // ---- results -----------------------------------------------------------------------------------------------------
        for (int l = tid; l < L; l += T) {
            a.w1[(long long)gid * L + l] = w1s[l];
            if (a.w2) a.w2[(long long)gid * L + l] = w2s[l];
        }
        if (a.tau || a.y) {
            for (int g = warp; g < G; g += nwarps)
                for (int l = lane; l < L; l += 32) {
                    if (a.tau) a.tau[(off + g) * L + l] = taus[g * Lp + l];
                    if (a.y) a.y[(off + g) * L + l] = ys[g * Lp + l];
                }
        }
        for (int g = tid; g < G; g += T) {
            if (a.x1) a.x1[off + g] = x1s[g];
            if (a.x2) a.x2[off + g] = x2s[g];
        }
        if (tid == 0) {
            if (a.n_iters) a.n_iters[gid] = iters;
            if (a.bound) a.bound[gid] = bound;
        }
    }
}

// Launch the CTA-per-gene kernel for a bucket of big genes; `scratch` must hold grid * slab_doubles doubles.
cudaError_t launch_em_big(const EmLaunch& a, int grid, long long slab_doubles, cudaStream_t st) {
    em_big_kernel<<<grid, kBigThreads, 0, st>>>(a, slab_doubles);
    return cudaGetLastError();
}

cudaError_t em_generic_configure(int maxG, int L, long long n_work_cap, int n_sm, EmGenericConfig* cfg, size_t max_smem) {
    // Shared-memory budget: as many warps per CTA (<= 8) as fit the largest gene; if not even one
    // fits, keep a slab per warp for the smaller genes of the batch and spill the big ones to the
    // global scratch.
    const long long need = em_generic_work_doubles(maxG, L);
    int wpc = 8;
    long long per_warp = need;
    while (wpc > 1 && (size_t)(per_warp * wpc) * sizeof(double) > max_smem) wpc >>= 1;
    if ((size_t)(per_warp * wpc) * sizeof(double) > max_smem) {
        wpc = 4;
        per_warp = (long long)(max_smem / sizeof(double) / wpc);
    }
    const size_t smem = (size_t)(per_warp * wpc) * sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(em_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, em_generic_kernel, wpc * 32, smem);
    if (e != cudaSuccess) return e;
    if (occ < 1) occ = 1;
    long long ctas = (long long)occ * n_sm;
    const long long want = (n_work_cap + wpc - 1) / wpc;
    if (want < ctas) ctas = want;
    if (ctas < 1) ctas = 1;
    cfg->warps_per_cta = wpc;
    cfg->smem_doubles_per_warp = per_warp;
    cfg->smem_bytes = smem;
    cfg->grid = (int)ctas;
    cfg->scratch_doubles_per_warp = (need > per_warp) ? need : 0;
    return cudaSuccess;
}

cudaError_t launch_em_generic(const EmGenericConfig& cfg, const EmLaunch& a, cudaStream_t st) {
    // the attribute is per kernel, and two configurations (generic bucket, deferred queue) share it
    cudaError_t e = cudaFuncSetAttribute(em_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem_bytes);
    if (e != cudaSuccess) return e;
    em_generic_kernel<<<cfg.grid, cfg.warps_per_cta * 32, cfg.smem_bytes, st>>>(a, cfg.warps_per_cta,
                                                                               cfg.smem_doubles_per_warp);
    return cudaGetLastError();
}

}  // namespace jacks
