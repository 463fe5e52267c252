// Fast batched EM kernel: one gene per team (CTA) of W warps, compile-time guide count G.
//
// Restates the reference's per-gene loop (jacks/jacks/infer.py:55-99) for finite data:
//   set-up   y = test - ctrl, tau_pr_den = tps*(sd_t^2 + sd_c^2 + 1e-2)          infer.py:58-70
//   init     x1 = mu0_x (or fixed_x), w1 = median_g y, tau = tps/tau_pr_den      infer.py:73-86
//   loop     X step (skipped with fixed_x) -> optional w hyper-prior -> W step
//            -> tau step -> bound; stop when |bound - last_bound| < tol          infer.py:89-98
//   X step   infer.py:103-110    W step  infer.py:112-116
//   tau step infer.py:118-121    bound   infer.py:123-125
//
// Layout.  The gene's working set -- y, tau_pr_den, tau as [G][L] and w1, w2 as [L] -- is staged in
// shared memory once ((3G+2)*L doubles, 44.8 KB at G=4, L=400) and never leaves the SM until the
// gene has converged.  Lane `lane` of warp `wv` owns lines l = wv*32 + lane + k*32*W; every line's
// W step, tau step and bound contribution are lane-local (G is a compile-time constant, so the G
// values of a line sit in registers while the line is processed), and per EM pass there is exactly
// ONE cross-lane operation: an all-reduce of 2G+1 doubles carrying the bound of this pass and the
// 2G row sums the next X step needs, both formed from the same fresh tau in the same sweep.
// The sweep is a rolled loop over lines (a few hundred instructions), so the hot code stays inside
// the instruction cache; the FP64 pipe is the intended limiter ("EM kernel").
//
// Arithmetic notes (all FP64).  b* = y^2 - 2*y*x1*w1 + x2*w2 is formed as fma(y, y - 2*x1*w1,
// x2*w2); the bound term tau*(y^2 + x2*w2 - 2*x1*w1*y) (infer.py:125) is the same quantity, so it
// is accumulated as tau*b*.  Reciprocals (1/den of the W step, tau = c/(...)) use MUFU.RCP64H plus
// a cubic Newton step (<= 1 ulp), not IEEE division: agreement with the reference stays ~1e-13,
// far inside the 1e-6 tolerance, and the convergence test |d bound| < 0.1 is insensitive to it
// (closest approach over all reference runs: 6.6e-5).
//
// Genes whose y block holds a NaN/inf (missing measurements) are not handled here: the team
// appends the gene to the deferred queue and the NaN-safe generic kernel (em_generic.cu) runs it.
#pragma once
#include "em_common.cuh"

namespace jacks {

#define JACKS_FULL_MASK 0xffffffffu

__device__ __forceinline__ double warp_allsum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(JACKS_FULL_MASK, v, o);
    return v;
}

// 1/d for normal, finite d: hardware seed (>= 20 bits) + one cubic step r*(1 + e + e^2), e = 1 - d*r.
__device__ __forceinline__ double fast_rcp(double d) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    const double e = fma(-d, r, 1.0);
    const double t = fma(e, e, e);
    return fma(r, t, r);
}

// Median of G finite values held in registers (numpy semantics: mean of the two middle values
// for even G).  Rank selection with index tie-break, O(G^2) compares, no data movement.
template <int G>
__device__ __forceinline__ double median_small(const double (&v)[G]) {
    if (G == 1) return v[0];
    if (G == 2) return 0.5 * (v[0] + v[1]);
    const int k_lo = (G - 1) / 2, k_hi = G / 2;
    double lo = v[0], hi = v[0];
#pragma unroll
    for (int i = 0; i < G; ++i) {
        int r = 0;
#pragma unroll
        for (int j = 0; j < G; ++j) {
            if (j < i) r += (v[j] <= v[i]);
            else if (j > i) r += (v[j] < v[i]);
        }
        if (r == k_lo) lo = v[i];
        if (r == k_hi) hi = v[i];
    }
    return (k_lo == k_hi) ? lo : 0.5 * (lo + hi);
}

// All-reduce of NV doubles over the team (W warps).  Every thread of the team ends up with
// bit-identical sums (xor butterfly inside the warp, fixed-order sum over warps), which keeps the
// convergence test uniform.  `red` is a [2][W][NV_MAX] shared buffer, `parity` flips per call.
template <int NV, int NV_MAX, int W>
__device__ __forceinline__ void team_allsum(double (&v)[NV], double* red, int& parity, int warp, int lane) {
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = warp_allsum(v[i]);
    if (W > 1) {
        double* buf = red + parity * (W * NV_MAX);
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < NV; ++i) buf[warp * NV_MAX + i] = v[i];
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            double s = buf[i];
#pragma unroll
            for (int w = 1; w < W; ++w) s += buf[w * NV_MAX + i];
            v[i] = s;
        }
        parity ^= 1;
    }
}

// Shared-memory doubles of one team: y, tau_pr_den, tau [G][L]; w1, w2 [L]; reduction buffer.
__host__ __device__ constexpr long long em_fast_smem_doubles(int G, int W, int L) {
    return (3LL * G + 2) * L + (W > 1 ? 2LL * W * (2 * G + 1) : 0);
}

template <int G, int W, bool UNIT_C>
__global__ void __launch_bounds__(W * 32) em_fast_kernel(const EmLaunch a) {
    constexpr int NV = 2 * G + 1;
    constexpr int STEP = 32 * W;
    extern __shared__ double smem[];
    const int L = a.L;
    double* ys = smem;                       // [G][L]
    double* tpds = ys + (size_t)G * L;       // [G][L]
    double* taus = tpds + (size_t)G * L;     // [G][L]
    double* w1s = taus + (size_t)G * L;      // [L]
    double* w2s = w1s + L;                   // [L]
    double* red = w2s + L;                   // [2][W][NV]   (W > 1 only)
    __shared__ int s_slot;

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const double tps = a.tps;
    const double c_tau = a.tps + 0.5;            // infer.py:120
    const double ivx = 1.0 / a.var0_x;           // infer.py:104
    const double mx = a.mu0_x / a.var0_x;
    const bool fixed = (a.fixed_x1 != nullptr);
    int parity = 0;

    for (;;) {
        // ---- fetch the next gene of this bucket --------------------------------------------
        int slot;
        if (W == 1) {
            slot = 0;
            if (lane == 0) slot = atomicAdd(a.work_counter, 1);
            slot = __shfl_sync(JACKS_FULL_MASK, slot, 0);
        } else {
            if (threadIdx.x == 0) s_slot = atomicAdd(a.work_counter, 1);
            __syncthreads();
            slot = s_slot;
        }
        if (slot >= a.n_work) break;
        const int gid = a.gene_ids[slot];
        const long long off = a.gene_offsets[gid];
        // rotate which warp owns the (possibly partial) last strip so the SM sub-partitions stay even
        const int l0 = ((W == 1) ? 0 : ((warp + slot) % W) * 32) + lane;

        // ---- stage y and tau_pr_den (infer.py:58-70) ---------------------------------------
        bool bad = false;
        double x1r[G], x2r[G];
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const int row = a.guide_rows[off + g];
            const double2* tr = a.test + (long long)row * L;
            const double2* cr = a.ctrl + (long long)row * a.ctrl_lines;
            if (fixed) { x1r[g] = a.fixed_x1[row]; x2r[g] = a.fixed_x2[row]; }
            else { x1r[g] = a.mu0_x; x2r[g] = a.mu0_x * a.mu0_x; }
            double2 c0 = make_double2(0.0, 0.0);
            double fill = 0.0;
            if (a.ctrl_lines == 1) {
                c0 = cr[0];
                if (a.ctrl_rowmean_fill && c0.y != c0.y) {
                    // 1-D control with an undefined sd: row mean of the (NaN -> 2.0) test sds, infer.py:63
                    double one[1] = {0.0};
                    for (int l = l0; l < L; l += STEP) {
                        const double e = tr[l].y;
                        one[0] += (e != e) ? 2.0 : e;
                    }
                    team_allsum<1, NV, W>(one, red, parity, warp, lane);
                    fill = one[0] / (double)L;
                }
            }
#pragma unroll 4
            for (int l = l0; l < L; l += STEP) {
                const double2 t = tr[l];
                const double2 c = (a.ctrl_lines == 1) ? c0 : cr[l];
                const double te = (t.y != t.y) ? 2.0 : t.y;                     // infer.py:58
                double ce = c.y;
                if (ce != ce) ce = (a.ctrl_lines == 1 && a.ctrl_rowmean_fill) ? fill : te;   // :63 / :68
                const double yv = t.x - c.x;
                const double pd = tps * (te * te + ce * ce + 1e-2);
                bad |= !(fabs(yv) <= 1.79769313486231570e308) || !(pd <= 1.79769313486231570e308);
                ys[g * L + l] = yv;
                tpds[g * L + l] = pd;
            }
        }
        const int any_bad = (W == 1) ? __any_sync(JACKS_FULL_MASK, bad) : __syncthreads_or(bad);
        if (any_bad) {
            if (threadIdx.x == 0) {
                const int p = atomicAdd(a.deferred_count, 1);
                a.deferred_ids[p] = gid;
            }
            continue;
        }
        // every lane reads back only the shared-memory words it wrote itself: no barrier needed

        // ---- initial w, tau, bound and the first X-step sums (infer.py:81-86) --------------
        double acc[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) acc[i] = 0.0;
        for (int l = l0; l < L; l += STEP) {
            double yv[G];
#pragma unroll
            for (int g = 0; g < G; ++g) yv[g] = ys[g * L + l];
            const double w1 = median_small<G>(yv);
            const double w2 = w1 * w1;
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const double t = tps / tpds[g * L + l];
                taus[g * L + l] = t;
                const double xw = x1r[g] * w1;
                acc[2 * G] += t * (yv[g] * yv[g] + x2r[g] * w2 - 2.0 * xw * yv[g]);
                acc[g] += yv[g] * t * w1;
                acc[G + g] += t * w2;
            }
            w1s[l] = w1; w2s[l] = w2;
        }
        team_allsum<NV, NV, W>(acc, red, parity, warp, lane);
        double bound = acc[2 * G];
        double mu0_w = a.mu0_w, var0_w = a.var0_w;
        int iters = 0;

        // ---- EM passes (infer.py:89-98) ------------------------------------------------------
        for (int it = 0; it < a.n_iter; ++it) {
            const double last = bound;
            if (!fixed) {
                // X step, infer.py:103-110.  Lane g (< G) owns guide g; results are broadcast.
                double s1 = acc[0], s2 = acc[G];
#pragma unroll
                for (int g = 1; g < G; ++g) if (lane == g) { s1 = acc[g]; s2 = acc[G + g]; }
                const double rden = fast_rcp(s2 + ivx);
                const double x1u = (mx + s1) * rden;
                const double x2u = fma(x1u, x1u, rden);
                double xa[G];
#pragma unroll
                for (int g = 0; g < G; ++g) xa[g] = __shfl_sync(JACKS_FULL_MASK, x1u, g);
                double sum = xa[0], hi = xa[0], lo = xa[0];
#pragma unroll
                for (int g = 1; g < G; ++g) { sum += xa[g]; hi = fmax(hi, xa[g]); lo = fmin(lo, xa[g]); }
                const double wadj = 0.5 / G;
                const double x1m = sum * (1.0 / G) + 2 * wadj * median_small<G>(xa) - wadj * hi - wadj * lo;
                const double rm = 1.0 / x1m;         // x1m may be tiny or negative: IEEE division
                const double x1n = x1u * rm;
                const double x2n = x2u * rm * rm;
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    x1r[g] = __shfl_sync(JACKS_FULL_MASK, x1n, g);
                    x2r[g] = __shfl_sync(JACKS_FULL_MASK, x2n, g);
                }
            }
            if (a.apply_w_hp && L > 1) {
                // hierarchical prior from the current w1, infer.py:92 (population variance)
                double m[1] = {0.0};
                for (int l = l0; l < L; l += STEP) m[0] += w1s[l];
                team_allsum<1, NV, W>(m, red, parity, warp, lane);
                const double mean = m[0] / (double)L;
                double v[1] = {0.0};
                for (int l = l0; l < L; l += STEP) { const double d = w1s[l] - mean; v[0] += d * d; }
                team_allsum<1, NV, W>(v, red, parity, warp, lane);
                mu0_w = mean;
                var0_w = (v[0] / (double)L) * 3 + 1e-4;
            }
            const double ivw = 1.0 / var0_w;        // infer.py:113
            const double mw = mu0_w / var0_w;
#pragma unroll
            for (int i = 0; i < NV; ++i) acc[i] = 0.0;
#pragma unroll 2
            for (int l = l0; l < L; l += STEP) {
                double yv[G], tv[G];
#pragma unroll
                for (int g = 0; g < G; ++g) { yv[g] = ys[g * L + l]; tv[g] = taus[g * L + l]; }
                // W step for this line, infer.py:112-116 (old tau, new x)
                double sa = ivw, sb = mw;
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    sa = fma(x2r[g], tv[g], sa);
                    sb = fma(x1r[g], yv[g] * tv[g], sb);
                }
                const double r = fast_rcp(sa);
                const double w1 = sb * r;
                const double w2 = fma(w1, w1, r);
                const double m2w1 = -2.0 * w1;
                // tau step (infer.py:118-121), bound (infer.py:123-125) and next X-step sums
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    const double t1 = fma(x1r[g], m2w1, yv[g]);            // y - 2*x1*w1
                    const double bstar = fma(yv[g], t1, x2r[g] * w2);      // y^2 - 2*y*x1*w1 + x2*w2
                    const double rd = fast_rcp(fma(0.5, bstar, tpds[g * L + l]));
                    const double t = UNIT_C ? rd : c_tau * rd;
                    taus[g * L + l] = t;
                    acc[2 * G] = fma(t, bstar, acc[2 * G]);
                    acc[g] = fma(yv[g] * t, w1, acc[g]);
                    acc[G + g] = fma(t, w2, acc[G + g]);
                }
                w1s[l] = w1; w2s[l] = w2;
            }
            team_allsum<NV, NV, W>(acc, red, parity, warp, lane);
            bound = acc[2 * G];
            ++iters;
            if (fabs(last - bound) < a.tol) break;
        }

        // ---- write results ------------------------------------------------------------------
        for (int l = l0; l < L; l += STEP) {
            a.w1[(long long)gid * L + l] = w1s[l];
            if (a.w2) a.w2[(long long)gid * L + l] = w2s[l];
            if (a.tau) {
#pragma unroll
                for (int g = 0; g < G; ++g) a.tau[(off + g) * L + l] = taus[g * L + l];
            }
            if (a.y) {
#pragma unroll
                for (int g = 0; g < G; ++g) a.y[(off + g) * L + l] = ys[g * L + l];
            }
        }
        if (warp == 0) {
            double xo1 = x1r[0], xo2 = x2r[0];
#pragma unroll
            for (int g = 1; g < G; ++g) if (lane == g) { xo1 = x1r[g]; xo2 = x2r[g]; }
            if (lane < G) {
                if (a.x1) a.x1[off + lane] = xo1;
                if (a.x2) a.x2[off + lane] = xo2;
            }
            if (lane == 0) {
                if (a.n_iters) a.n_iters[gid] = iters;
                if (a.bound) a.bound[gid] = bound;
            }
        }
    }
}

}  // namespace jacks
