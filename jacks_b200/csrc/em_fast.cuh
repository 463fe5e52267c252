// Fast batched EM kernel: one gene per team of W warps, compile-time guide count G.
//
// Restates the reference's per-gene loop (jacks/jacks/infer.py:55-99) for finite data:
//   set-up   y = test - ctrl, tau_pr_den = tps*(sd_t^2 + sd_c^2 + 1e-2)          infer.py:58-70
//   init     x1 = mu0_x (or fixed_x), w1 = median_g y, tau = tps/tau_pr_den      infer.py:73-86
//   loop     X step (skipped with fixed_x) -> optional w hyper-prior -> W step
//            -> tau step -> bound; stop when |bound - last_bound| < tol          infer.py:89-98
//   X step   infer.py:103-110    W step  infer.py:112-116
//   tau step infer.py:118-121    bound   infer.py:123-125
//
// Data layout.  Each lane owns KL "strips": line l = (k*W + warp')*32 + lane for k < KL.  The
// gene's y and tau_pr_den blocks ([G][Lp], Lp = KL*W*32) are staged once in shared memory; the
// evolving state (tau for the lane's G*KL elements, w1/w2 for its KL lines, x1/x2 replicated) lives
// in registers.  Per EM pass the only cross-lane traffic is ONE all-reduce of 2G+1 doubles: the
// bound of this pass and the 2G row sums the next X step needs are formed from the same fresh
// tau in the same sweep.  Lines beyond L are padded with y = 0, tau_pr_den = +inf, which makes
// every padded element contribute an exact 0 everywhere (tau = 0), so the sweeps need no masks.
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

template <int G, int KL, int W, int MINB>
__global__ void __launch_bounds__(W * 32, MINB) em_fast_kernel(const EmLaunch a) {
    constexpr int Lp = KL * W * 32;
    constexpr int NV = 2 * G + 1;
    extern __shared__ double smem[];
    double* ys = smem;                 // [G][Lp]
    double* tpds = smem + G * Lp;      // [G][Lp]
    double* red = tpds + G * Lp;       // [2][W][NV]   (W > 1 only)
    __shared__ int s_slot;

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int L = a.L;
    const double tps = a.tps;
    const double c_tau = a.tps + 0.5;            // infer.py:120
    const double ivx = 1.0 / a.var0_x;           // infer.py:104
    const double mx = a.mu0_x / a.var0_x;
    const bool fixed = (a.fixed_x1 != nullptr);
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
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
        const int wv = (W == 1) ? 0 : (warp + slot) % W;   // rotate the warp that owns the partial strip

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
                    double acc = 0.0;
                    for (int l = threadIdx.x; l < L; l += W * 32) {
                        double e = tr[l].y;
                        acc += (e != e) ? 2.0 : e;
                    }
                    double one[1] = {acc};
                    team_allsum<1, NV, W>(one, red, parity, warp, lane);
                    fill = one[0] / (double)L;
                }
            }
#pragma unroll
            for (int k = 0; k < KL; ++k) {
                const int l = (k * W + wv) * 32 + lane;
                double yv = 0.0, pd = inf;
                if (l < L) {
                    const double2 t = tr[l];
                    const double2 c = (a.ctrl_lines == 1) ? c0 : cr[l];
                    const double te = (t.y != t.y) ? 2.0 : t.y;                     // infer.py:58
                    double ce = c.y;
                    if (ce != ce) ce = (a.ctrl_lines == 1 && a.ctrl_rowmean_fill) ? fill : te;   // :63 / :68
                    yv = t.x - c.x;
                    pd = tps * (te * te + ce * ce + 1e-2);
                    bad |= !(fabs(yv) < inf) || !(pd == pd);
                }
                ys[g * Lp + l] = yv;
                tpds[g * Lp + l] = pd;
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
        double tau_r[KL][G], w1_r[KL], w2_r[KL];
        double acc[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) acc[i] = 0.0;
#pragma unroll
        for (int k = 0; k < KL; ++k) {
            const int s = k * W + wv;
            w1_r[k] = 0.0; w2_r[k] = 0.0;
            if (s * 32 < L) {
                const int l = s * 32 + lane;
                double yv[G];
#pragma unroll
                for (int g = 0; g < G; ++g) yv[g] = ys[g * Lp + l];
                const double w1 = median_small<G>(yv);
                const double w2 = w1 * w1;
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    const double t = tps / tpds[g * Lp + l];
                    tau_r[k][g] = t;
                    const double xw = x1r[g] * w1;
                    acc[2 * G] += t * (yv[g] * yv[g] + x2r[g] * w2 - 2.0 * xw * yv[g]);
                    acc[g] += yv[g] * t * w1;
                    acc[G + g] += t * w2;
                }
                w1_r[k] = w1; w2_r[k] = w2;
            } else {
#pragma unroll
                for (int g = 0; g < G; ++g) tau_r[k][g] = 0.0;
            }
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
                const double den = s2 + ivx;
                const double x1u = (mx + s1) / den;
                const double x2u = x1u * x1u + 1.0 / den;
                double xa[G];
#pragma unroll
                for (int g = 0; g < G; ++g) xa[g] = __shfl_sync(JACKS_FULL_MASK, x1u, g);
                double sum = xa[0], hi = xa[0], lo = xa[0];
#pragma unroll
                for (int g = 1; g < G; ++g) { sum += xa[g]; hi = fmax(hi, xa[g]); lo = fmin(lo, xa[g]); }
                const double wadj = 0.5 / G;
                const double x1m = sum / G + 2 * wadj * median_small<G>(xa) - wadj * hi - wadj * lo;
                const double x1n = x1u / x1m;
                const double x2n = x2u / x1m / x1m;
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    x1r[g] = __shfl_sync(JACKS_FULL_MASK, x1n, g);
                    x2r[g] = __shfl_sync(JACKS_FULL_MASK, x2n, g);
                }
            }
            if (a.apply_w_hp && L > 1) {
                // hierarchical prior from the current w1, infer.py:92 (population variance)
                double m[1] = {0.0};
#pragma unroll
                for (int k = 0; k < KL; ++k) {
                    const int l = (k * W + wv) * 32 + lane;
                    if (l < L) m[0] += w1_r[k];
                }
                team_allsum<1, NV, W>(m, red, parity, warp, lane);
                const double mean = m[0] / (double)L;
                double v[1] = {0.0};
#pragma unroll
                for (int k = 0; k < KL; ++k) {
                    const int l = (k * W + wv) * 32 + lane;
                    if (l < L) { const double d = w1_r[k] - mean; v[0] += d * d; }
                }
                team_allsum<1, NV, W>(v, red, parity, warp, lane);
                mu0_w = mean;
                var0_w = (v[0] / (double)L) * 3 + 1e-4;
            }
            const double ivw = 1.0 / var0_w;        // infer.py:113
            const double mw = mu0_w / var0_w;
#pragma unroll
            for (int i = 0; i < NV; ++i) acc[i] = 0.0;
#pragma unroll
            for (int k = 0; k < KL; ++k) {
                const int s = k * W + wv;
                if (s * 32 < L) {
                    const int l = s * 32 + lane;
                    double yv[G];
                    double sa = 0.0, sb = 0.0;
                    // W step for this line, infer.py:112-116 (old tau, new x)
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        yv[g] = ys[g * Lp + l];
                        const double t = tau_r[k][g];
                        sa += x2r[g] * t;
                        sb += x1r[g] * (yv[g] * t);
                    }
                    const double den = sa + ivw;
                    const double w1 = (mw + sb) / den;
                    const double w2 = w1 * w1 + 1.0 / den;
                    // tau step (infer.py:118-121), bound (infer.py:123-125) and next X-step sums
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        const double xw = x1r[g] * w1;
                        const double xw2 = x2r[g] * w2;
                        const double y2 = yv[g] * yv[g];
                        const double bstar = y2 - 2 * yv[g] * xw + xw2;
                        const double t = c_tau / (tpds[g * Lp + l] + 0.5 * bstar);
                        tau_r[k][g] = t;
                        acc[2 * G] += t * (y2 + xw2 - 2 * xw * yv[g]);
                        acc[g] += yv[g] * t * w1;
                        acc[G + g] += t * w2;
                    }
                    w1_r[k] = w1; w2_r[k] = w2;
                }
            }
            team_allsum<NV, NV, W>(acc, red, parity, warp, lane);
            bound = acc[2 * G];
            ++iters;
            if (fabs(last - bound) < a.tol) break;
        }

        // ---- write results ------------------------------------------------------------------
#pragma unroll
        for (int k = 0; k < KL; ++k) {
            const int l = (k * W + wv) * 32 + lane;
            if (l < L) {
                a.w1[(long long)gid * L + l] = w1_r[k];
                if (a.w2) a.w2[(long long)gid * L + l] = w2_r[k];
                if (a.tau) {
#pragma unroll
                    for (int g = 0; g < G; ++g) a.tau[(off + g) * L + l] = tau_r[k][g];
                }
                if (a.y) {
#pragma unroll
                    for (int g = 0; g < G; ++g) a.y[(off + g) * L + l] = ys[g * Lp + l];
                }
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

template <int G, int KL, int W>
constexpr size_t em_fast_smem_bytes() {
    return (size_t)(2 * G * KL * W * 32 + (W > 1 ? 2 * W * (2 * G + 1) : 0)) * sizeof(double);
}

}  // namespace jacks
