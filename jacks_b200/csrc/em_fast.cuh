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
// Layout, shared-memory variant (TM = 0).  Shared memory holds what is read AND written every pass: y and tau as
// [G][L] (25.6 KB at G=4, L=400) plus the reduction tiles -- 8 genes, i.e. two warps per scheduler, per SM.
// tau_pr_den is read once per pass and never written after staging: it lives in a per-team global slab that stays L2
// resident and is fetched one sweep step ahead.  E[w], E[w^2] are write-only within a pass: straight to the gene's
// output rows.  Lane `lane` of warp `wv` owns lines l = wv*32 + lane + k*32*W; every line's W step, tau step and
// bound contribution are lane-local (G is a compile-time constant, so the G values of a line sit in registers
// while the line is processed), and per EM pass there is exactly ONE cross-lane operation: an all-reduce of
// 2G+1 doubles carrying the bound of this pass and the 2G row sums the next X step needs, both formed from the
// same fresh tau in the same sweep.  The sweep is a rolled loop over line steps (a few hundred instructions), so
// the hot code stays inside the instruction cache.
//
// Tensor-memory variants (TM = 1, 2; one-warp teams, four independent teams per CTA).  Blackwell's TMEM is used as a
// lane-private FP64 scratchpad (tmem.cuh): tau (TM >= 1) and tau_pr_den (TM = 2) live in column pairs [line][g] of each
// thread's own TMEM lane and move with one LDTM / STTM per sweep step; only y stays in shared memory.  TM = 2 needs
// no global slab and makes no L2 round trip inside a pass; staging then also initialises the gene.  The launcher
// (em_fast_inst.cu) picks the variant by shape; DESIGN.md section 4.1 has the measurements.
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
#include "tmem.cuh"

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

// Reciprocal seed for a Newton refinement: MUFU.RCP64H delivers the HIGH word of 1/d (>= 20 bits); as PTX defines it the
// low word is zero, which costs a register move per seed.  The low word only carries mantissa bits 2^-21 and below, so
// any bit pattern will do for a seed: it is taken from `junk`, a value that is dead at this point, whose register pair
// the seed can then occupy without a move.
__device__ __forceinline__ double rcp_seed(double d, double junk) {
    double t;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(t) : "d"(d));
#ifdef JACKS_SEED_ZERO_LOW
    (void)junk;
    return t;
#else
    return __hiloint2double(__double2hiint(t), __double2loint(junk));
#endif
}

// Median of G finite values held in registers (numpy semantics: mean of the two middle values for
// even G).  G <= 5 use min/max networks (the median stays an exact element / exact pair mean);
// larger G use rank selection with index tie-break, O(G^2) compares, no data movement.
template <int G>
__device__ __forceinline__ double median_small(const double (&v)[G]) {
    if (G == 1) return v[0];
    if (G == 2) return 0.5 * (v[0] + v[1]);
    if (G == 3) return fmax(fmin(v[0], v[1]), fmin(fmax(v[0], v[1]), v[2]));
    if (G == 4) {
        const double lo1 = fmin(v[0], v[1]), hi1 = fmax(v[0], v[1]);
        const double lo2 = fmin(v[2], v[3]), hi2 = fmax(v[2], v[3]);
        return 0.5 * (fmax(lo1, lo2) + fmin(hi1, hi2));
    }
    if (G == 5) {
        // median of five in 6 compare-exchange layers
        double a = fmin(v[0], v[1]), b = fmax(v[0], v[1]);
        double c = fmin(v[2], v[3]), d = fmax(v[2], v[3]);
        const double lo = fmax(a, c);         // drop the smaller of the two minima
        const double hi = fmin(b, d);         // drop the larger of the two maxima
        // median of {lo, hi, v[4]} -- lo and hi may be in either order
        const double p = fmin(lo, hi), q = fmax(lo, hi);
        return fmax(p, fmin(q, v[4]));
    }
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

// Maximum, minimum and median of G FINITE values in one go (the X step's normaliser, infer.py:108).  Plain compares
// and selects: fmax/fmin spend instructions on NaN operands the fast path never sees; for G = 4 one five-comparator
// network yields all three (the median stays the exact mean of the two middle elements, as numpy forms it).
__device__ __forceinline__ double sel_max(double a, double b) { return a > b ? a : b; }
__device__ __forceinline__ double sel_min(double a, double b) { return a < b ? a : b; }
template <int G>
__device__ __forceinline__ void max_min_median(const double (&v)[G], double& hi, double& lo, double& med) {
    if (G == 4) {
        const double lo1 = sel_min(v[0], v[1]), hi1 = sel_max(v[0], v[1]);
        const double lo2 = sel_min(v[2], v[3]), hi2 = sel_max(v[2], v[3]);
        hi = sel_max(hi1, hi2);
        lo = sel_min(lo1, lo2);
        med = 0.5 * (sel_max(lo1, lo2) + sel_min(hi1, hi2));
        return;
    }
    hi = v[0]; lo = v[0];
#pragma unroll
    for (int g = 1; g < G; ++g) { hi = sel_max(hi, v[g]); lo = sel_min(lo, v[g]); }
    med = median_small<G>(v);
}

// All-reduce of NV doubles over the team (W warps); every thread ends up with bit-identical sums
// (fixed summation order), which keeps the convergence test uniform.
// A butterfly over 32 lanes costs 5*NV shuffle pairs + 5*NV DADDs plus register moves (~320
// instructions at NV = 9) of mostly exposed latency for a lone warp.  Instead the lanes park their
// partials in a padded [NV][33] shared tile, H = 1, 2 or 4 lanes per value sum 32/H of them, one or
// two shuffles finish the value, and the NV totals are broadcast back through shared memory:
// ~60 instructions.  `scr`: this warp's [NV_MAX][33] tile; `tot`: [2][W][NV_MAX]; `parity` flips per call.
template <int NV, int NV_MAX, int W>
__device__ __forceinline__ void team_allsum(double (&v)[NV], double* scr, double* tot, int& parity, int warp, int lane) {
    constexpr int H = (NV <= 8) ? 4 : ((NV <= 16) ? 2 : 1);
    constexpr int CH = 32 / H;
#pragma unroll
    for (int i = 0; i < NV; ++i) scr[i * 33 + lane] = v[i];
    __syncwarp();
    const int q = (H == 4) ? (lane & 7) : ((H == 2) ? (lane & 15) : lane);
    const int h = (H == 4) ? (lane >> 3) : ((H == 2) ? (lane >> 4) : 0);
    double s0 = 0.0, s1 = 0.0;
    if (q < NV) {
        const double* row = scr + q * 33 + h * CH;
#pragma unroll
        for (int k = 0; k < CH; k += 2) { s0 += row[k]; s1 += row[k + 1]; }
    }
    s0 += s1;
    if (H >= 2) s0 += __shfl_xor_sync(JACKS_FULL_MASK, s0, (H == 4) ? 8 : 16);
    if (H == 4) s0 += __shfl_xor_sync(JACKS_FULL_MASK, s0, 16);
    double* mine = tot + parity * (W * NV_MAX) + warp * NV_MAX;
    if (q < NV && h == 0) mine[q] = s0;
    if (W > 1) __syncthreads(); else __syncwarp();
    const double* all = tot + parity * (W * NV_MAX);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double s = all[i];
#pragma unroll
        for (int w = 1; w < W; ++w) s += all[w * NV_MAX + i];
        v[i] = s;
    }
    parity ^= 1;
}

// One sweep step over U lines of a lane (l, l+step, ...): W step, tau step, bound and next-X sums.
// Written stage by stage over the U*G independent elements so that consecutive instructions are
// independent (a dependent DFMA issues every ~9 cycles on sm_100a, an independent one every 2):
// with U*G >= 8 each stage is long enough to cover the latency of the previous one.
//
// Scaled state (all scalings are exact powers of two): shared memory holds th = tau/2 and the slab
// holds pd2 = 2*tau_pr_den, so that the tau step (infer.py:118-121) is
//     den2 = 2*(tau_pr_den + b*/2) = y*(y - 2*x1*w1) + (x2*w2 + pd2),      th = c/den2,
// three FMAs per element, and the bound needs no b* at all: tau*b* = 2c - 2*tau_pr_den*tau, i.e.
//     sum tau*b* = 2*c*G*L - 2*sum pd2*th                                   (infer.py:123-125).
// The W step and the X sums use the doubled guide efficacies (tx = 2*x) / are doubled by the caller.
// 13 FP64 instructions per element and pass, 6 more per line.
//
// MASKED: the step's last line (u = U-1) may lie beyond L for some lanes (`valid` false).  Those
// lanes still execute the arithmetic on whatever the shared-memory words past the line hold (the
// allocation carries one step of slack, see em_fast_smem_doubles), but neither store nor
// accumulate it: the stores and the three accumulating FMAs of that line are predicated.
// reload: each pd2 register is refilled with the next step's value (`pd_next`, L2) as soon as its
// only use (t1 = x2*w2 + pd2) has issued -- the load has the rest of this step and the W step of
// the next to land.  The bound therefore accumulates t1*th and the caller removes x2*sum th*w2.
template <int G> struct EmX { double x1[G], x2[G], tx1[G], tx2[G]; };
// What the sweep needs of x: E[x^2] and the doubled moments (the state is th = tau/2).  E[x] itself enters only as
// -2*x1*w1, formed from tx1 with a (free) operand negation.
template <int G> struct EmXs { double x2[G], tx1[G], tx2[G]; };

// TM: th lives in tensor memory (tmem.cuh) instead of shared memory: the step's U*G values are U*G consecutive
// column pairs of the thread's own TMEM lane at `tcol`, moved by one tcgen05.ld and one tcgen05.st per step.
// A lane's dead line (MASKED, !valid) then holds garbage that is loaded, computed on and stored back, never used.
// TM == 2: pd2 lives in TMEM as well (columns `pcol`...), loaded at the top of the step that uses it: no L2 round
// trip, no look-ahead registers, no global slab.
template <int G, int U, int UP, bool UNIT_C, bool MASKED, int TM, bool WS>
__device__ __forceinline__ void em_sweep(const uint32_t tcol, const uint32_t pcol, const double* __restrict__ ys, double (&pd)[UP][G],
                                         const double* __restrict__ pd_next, const int Lp,
                                         double* __restrict__ ths, double* __restrict__ w1p,
                                         double* __restrict__ w2g, const int L, const int l, const int step,
                                         const EmXs<G>& x, const double ivw, const double mw, const double c_tau,
                                         double (&a1)[G], double (&a2)[G], double (&ab)[G], const bool valid, const bool reload) {
    double yv[U][G], tv[U][G];
    uint32_t traw[TM ? 2 * U * G : 2];
    uint32_t praw[TM == 2 ? 2 * U * G : 2];
    if (TM) tmem_ld_issue<(TM ? U * G : 1)>(tcol, reinterpret_cast<uint32_t(&)[2 * (TM ? U * G : 1)]>(traw));
#ifndef JACKS_PD_LATE
    if (TM == 2) tmem_ld_issue<(TM == 2 ? U * G : 1)>(pcol, reinterpret_cast<uint32_t(&)[2 * (TM == 2 ? U * G : 1)]>(praw));
#endif
#pragma unroll
    for (int u = 0; u < U; ++u) {
        // the masked line of a lane that has run out of lines reads line 0 instead (any valid word)
        const int lu = (MASKED && u == U - 1 && !valid) ? 0 : l + u * step;
#pragma unroll
        for (int g = 0; g < G; ++g) {
            yv[u][g] = ys[g * L + lu];
            if (!TM) tv[u][g] = ths[g * L + lu];
        }
    }
    if (TM) {
        tmem_wait_ld();
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int g = 0; g < G; ++g)
                tv[u][g] = __hiloint2double((int)traw[2 * (u * G + g) + 1], (int)traw[2 * (u * G + g)]);
#ifndef JACKS_PD_LATE
        if (TM == 2) {
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int g = 0; g < G; ++g)
                    pd[u][g] = __hiloint2double((int)praw[2 * (u * G + g) + 1], (int)praw[2 * (u * G + g)]);
        }
#endif
    }
    // W step, infer.py:112-116 (old tau, new x): den = 1/var0_w + sum_g x2*tau, num = mu0_w/var0_w + sum_g x1*y*tau
    double yt[U][G];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
        for (int g = 0; g < G; ++g) yt[u][g] = yv[u][g] * tv[u][g];
    double sa[U], sb[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { sa[u] = ivw; sb[u] = mw; }
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
        for (int u = 0; u < U; ++u) {
            sa[u] = fma(x.tx2[g], tv[u][g], sa[u]);
            sb[u] = fma(x.tx1[g], yt[u][g], sb[u]);
        }
#ifdef JACKS_PD_LATE
    // tau_pr_den comes in only now, after the last use of the old tau: both arrays can pass through the SAME block of
    // 2*U*G consecutive registers (LDTM needs them contiguous) without the first being copied out of the way
    if (TM == 2) {
        tmem_ld_issue<(TM == 2 ? U * G : 1)>(pcol, reinterpret_cast<uint32_t(&)[2 * (TM == 2 ? U * G : 1)]>(praw));
        tmem_wait_ld();
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int g = 0; g < G; ++g)
                pd[u][g] = __hiloint2double((int)praw[2 * (u * G + g) + 1], (int)praw[2 * (u * G + g)]);
    }
#endif
    // JACKS_RCP2 (developer A/B, off): one quadratic Newton step instead of the cubic one, 12 instead of 13 FP64
    // instructions per element: 6.46 -> 6.16 ms per Avana pass with identical pass counts, but the seed carries only
    // 20 bits, tau comes out 1e-12 off and gene scores that cancel to |w1| ~ 1e-6 miss the 1e-6 RELATIVE bar
    // (the full-size Avana parity test: worst 3.3e-6; 1.2e-6 with a zeroed low seed word).  Rejected.
    double r[U], e[U];
#pragma unroll
    for (int u = 0; u < U; ++u) asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r[u]) : "d"(sa[u]));
#pragma unroll
    for (int u = 0; u < U; ++u) e[u] = fma(-sa[u], r[u], 1.0);
#ifndef JACKS_RCP2
#pragma unroll
    for (int u = 0; u < U; ++u) e[u] = fma(e[u], e[u], e[u]);
#endif
#pragma unroll
    for (int u = 0; u < U; ++u) r[u] = fma(r[u], e[u], r[u]);
    double w1[U], w2[U];
#pragma unroll
    for (int u = 0; u < U; ++u) w1[u] = sb[u] * r[u];
#pragma unroll
    for (int u = 0; u < U; ++u) w2[u] = fma(w1[u], w1[u], r[u]);
    // tau step, infer.py:118-121
    double d[U][G], t1[U][G], rd[U][G], ee[U][G];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
        for (int g = 0; g < G; ++g) {
            d[u][g] = fma(-x.tx1[g], w1[u], yv[u][g]);
            t1[u][g] = fma(x.x2[g], w2[u], pd[u][g]);
            if (TM != 2 && reload) pd[u][g] = __ldg(pd_next + g * Lp + u * step);
        }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
        for (int g = 0; g < G; ++g) d[u][g] = fma(yv[u][g], d[u][g], t1[u][g]);
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
        for (int g = 0; g < G; ++g) rd[u][g] = rcp_seed(d[u][g], tv[u][g]);       // the old tau is dead by now
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
        for (int g = 0; g < G; ++g) ee[u][g] = fma(-d[u][g], rd[u][g], 1.0);
#ifndef JACKS_RCP2
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
        for (int g = 0; g < G; ++g) ee[u][g] = fma(ee[u][g], ee[u][g], ee[u][g]);
#endif
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
        for (int g = 0; g < G; ++g) {
            rd[u][g] = fma(rd[u][g], ee[u][g], rd[u][g]);
            if (!UNIT_C) rd[u][g] *= c_tau;
        }
    if (TM) {
        tmem_store<U * G>(tcol, &rd[0][0]);
    } else {
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int g = 0; g < G; ++g)
                if (!MASKED || u < U - 1 || valid) ths[g * L + l + u * step] = rd[u][g];
    }
    // bound (sum pd2*th = sum t1*th - x2*sum th*w2, the caller subtracts) and the next X step's row sums
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
        for (int g = 0; g < G; ++g) {
            if (!MASKED || u < U - 1 || valid) ab[g] = fma(t1[u][g], rd[u][g], ab[g]);
            yt[u][g] = yv[u][g] * rd[u][g];
        }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
        for (int g = 0; g < G; ++g) {
            if (!MASKED || u < U - 1 || valid) {
                a1[g] = fma(yt[u][g], w1[u], a1[g]);
                a2[g] = fma(rd[u][g], w2[u], a2[g]);
            }
        }
#pragma unroll
    for (int u = 0; u < U; ++u)
        if (!MASKED || u < U - 1 || valid) {
            if (WS) {
                // E[w], E[w^2] of the pass as one 16-byte shared-memory store per line; they reach the gene's output rows
                // once, after the last pass (every pass used to write them to global memory: two STG and their
                // descriptor moves per line, and 0.8 GB of DRAM writes per Avana pass for values that are overwritten)
                reinterpret_cast<double2*>(w1p)[l + u * step] = make_double2(w1[u], w2[u]);
            } else {
                w1p[l + u * step] = w1[u];                // E[w], E[w^2]: write-only state within a pass, kept in the
                if (w2g) w2g[l + u * step] = w2[u];      // gene's output rows (L2 resident), not in shared memory
            }
        }
}

// The sums a pass returns (register-passed struct): a1 = sum_l y*th*w1, a2 = sum_l th*w2 per guide
// (HALF the X-step sums) and ab = sum pd2*th.
template <int G> struct EmSums { double a1[G], a2[G], ab; };

// Lines per sweep step: U*G >= 8 independent elements, and few enough (<= 20) to stay in registers
// (measured at 400 lines: G = 5 wants 3 lines from shared memory, 4 from TMEM; G = 6 wants 3 either way).
template <int G, int TM> __host__ __device__ constexpr int em_sweep_lines() {
#ifdef JACKS_UMAX_TM
    if (TM && G <= 4) return JACKS_UMAX_TM;
#endif
    return (G >= 9) ? 1 : (G >= 7) ? 2 : ((G >= (TM ? 6 : 5)) ? 3 : 4);
}

// One EM pass over all lines of a warp.  Deliberately NOT inlined: compiled on its own, ptxas
// interleaves the U*G independent element chains of a step (a lone warp then keeps the FP64 pipe
// ~75 % busy, tools/sass_sim.py); inlined into the kernel it serialises them line by line.
//
// Step plan (warp-uniform, no divergence).  `n` = lines of the warp's fullest lane: `n_big` steps of UMAX lines, then one
// final MASKED step of B = 1..UMAX lines whose last line is live only in the lanes with l < L.  B is a template
// parameter (the kernel switches over it once per pass), so only the two step bodies a launch actually uses occupy the
// instruction cache.  (Measured at 400 lines, 13 = 4+4+4+1 against the balanced 4+3+3+3: 6.70 against 6.88 ms -- the
// kernel is bound by instruction dispatch, FP64 instructions weighing ~2.6 slots each, and a 3-line step carries more
// bookkeeping per line than a 4-line one; the short tail step is cheap in those terms.)
//
// pd2 (TM != 2) lives in an L2-resident global slab [G][Lp] (read once per pass, never written after staging, so
// it does not have to occupy shared memory); rows are padded (Lp >= L + U*STEP) so that the look-ahead
// loads need no bounds logic.  A step refills its pd registers for the NEXT step, which is never taller than itself.
template <int G, int STEP, bool UNIT_C, int TM, int B, bool WS>
__device__ __noinline__ EmSums<G> em_sweep_all(const uint32_t tbase, const uint32_t pdoff, const double* ys, const double* tpdg, const int Lp, double* ths,
                                               double* w1p, double* w2g, const int L, const int l0,
                                               const int n_big, const EmXs<G> x_in, const double ivw, const double mw,
                                               const double c_tau) {
    constexpr int UMAX = em_sweep_lines<G, TM>();
    constexpr int UP = UMAX;
    const EmXs<G>& x = x_in;
    double a1[G], a2[G], ab[G];
#pragma unroll
    for (int g = 0; g < G; ++g) { a1[g] = 0.0; a2[g] = 0.0; ab[g] = 0.0; }
    int l = l0;
    double pd[UP][G];
    if (TM != 2) {
#pragma unroll
        for (int u = 0; u < UP; ++u)
#pragma unroll
            for (int g = 0; g < G; ++g) pd[u][g] = __ldg(tpdg + g * Lp + l + u * STEP);
    }
    // the previous pass's (or the initialisation's) TMEM stores must have landed before these columns are read
    // back; waiting here rather than after the last store hides the store latency behind the reduction and X step
    if (TM) tmem_wait_st();
    uint32_t tcol = tbase;
    for (int i = 0; i < n_big; ++i, l += UP * STEP, tcol += 2 * UP * G)
        em_sweep<G, UP, UP, UNIT_C, false, TM, WS>(tcol, tcol + pdoff, ys, pd, tpdg + l + UP * STEP, Lp, ths, w1p, w2g, L, l, STEP, x, ivw, mw, c_tau, a1, a2, ab, true, true);
    em_sweep<G, B, UP, UNIT_C, true, TM, WS>(tcol, tcol + pdoff, ys, pd, nullptr, Lp, ths, w1p, w2g, L, l, STEP, x, ivw, mw, c_tau, a1, a2, ab, (l + (B - 1) * STEP) < L, false);
    EmSums<G> r;
    // sum pd2*th = sum (t1 - x2*w2)*th: the accumulators hold sum t1*th
    r.ab = fma(-x.x2[0], a2[0], ab[0]);
#pragma unroll
    for (int g = 1; g < G; ++g) r.ab += fma(-x.x2[g], a2[g], ab[g]);
#pragma unroll
    for (int g = 0; g < G; ++g) { r.a1[g] = a1[g]; r.a2[g] = a2[g]; }
    return r;
}

// The plan for `n` lines per lane (see em_sweep_all): n = n_big*umax + B, 1 <= B <= umax.
__device__ __forceinline__ void em_step_plan(int n, int umax, int& B, int& n_big) {
    if (n < 1) n = 1;          // an empty strip (W > 1, short line axis) runs one fully masked step
    n_big = (n - 1) / umax;
    B = n - n_big * umax;
}

// Stage a gene: y = test - ctrl (shared) and pd2 = 2*tau_pr_den (global slab) for the lane's lines, infer.py:58-70.
// All G rows of a line batch are loaded before any is consumed (G*4 independent 16-byte loads in flight per
// lane): staging is pure memory latency for a team that has only its own warps to hide it behind.
// Not inlined (runs once per gene).  Returns true if a staged value is not finite.
// TMEM variants: the loop has a warp-uniform trip count (the TMEM stores are warp-collective) and dead lines are
// predicated; with TM == 2, pd2 goes to the thread's TMEM lane at `tpd` ([line][g] column pairs) instead of the slab.
// In that mode the pass also INITIALISES the gene (infer.py:81-86), since every value it needs is in registers:
// th = tau/2 = tps/pd2 to the TMEM columns at `tth`, the initial E[w] (median over the guides) and E[w^2] to the
// gene's output rows, and the lane's partial sums for the initial bound and the first X step (formed with th,
// i.e. HALF sums) in the returned struct.
template <int G> struct EmStaged { double acc[2 * G + 1]; bool bad; };
// Staging also initialises the gene in the TMEM variants -- mode 2 always; mode 1 up to four guides (measured at
// 400..800 lines: G = 4 gains 5 %, G = 5..7 lose 1..4 % to register pressure in the staging loop).
template <int G, int TM> __host__ __device__ constexpr bool em_stage_initialises() { return TM == 2 || (TM == 1 && G <= 4); }

template <int G, int STEP, int TM, bool WS>
__device__ __noinline__ EmStaged<G> em_stage_gene(const double2* __restrict__ test, const double2* __restrict__ ctrl,
                                                  const int* rows, double* __restrict__ ys, double* __restrict__ tpdg,
                                                  const int L, const int Lp, const int l0, const int ctrl_lines,
                                                  const int rowmean_fill, const double* fill, const double tps,
                                                  const uint32_t tpd, const uint32_t tth, const int n_warp,
                                                  const EmX<G> x, double* __restrict__ w1g, double* __restrict__ w2g) {
    EmStaged<G> r;
#pragma unroll
    for (int i = 0; i < 2 * G + 1; ++i) r.acc[i] = 0.0;
    bool bad = false;
    const double2* tr[G];
    const double2* cr[G];
    double2 c0[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
        tr[g] = test + (long long)rows[g] * L;
        cr[g] = ctrl + (long long)rows[g] * ctrl_lines;
        c0[g] = make_double2(0.0, 0.0);
    }
    if (ctrl_lines == 1) {
#pragma unroll
        for (int g = 0; g < G; ++g) c0[g] = cr[g][0];
    }
#ifdef JACKS_STAGE_V2
    constexpr int B = (G >= 9) ? 1 : 2;       // lines per batch
    if (em_stage_initialises<G, TM>()) {
        // Per-guide constants of the initial sums; with one control column its mean, its (NaN-patched) variance term
        // and the 2*tps factor leave the line loop as well.
        const bool one_ctrl = (ctrl_lines == 1);
        const double two_tps = 2.0 * tps;
        double tx1[G], cm[G], cc[G];
        bool cnan[G];
#pragma unroll
        for (int g = 0; g < G; ++g) {
            tx1[g] = 2.0 * x.x1[g];
            cm[g] = c0[g].x;
            double ce = c0[g].y;
            cnan[g] = one_ctrl && (ce != ce) && !rowmean_fill;         // matched rule: a NaN control sd takes the test sd (:68)
            if (one_ctrl && (ce != ce) && rowmean_fill) ce = fill[g];   // 1-D control rule (:63)
            cc[g] = fma(ce, ce, 1e-2);
        }
        double chk = 0.0;                       // finite <=> every staged value is: checked once, after the loop
#pragma unroll 1
        for (int i = 0; i < n_warp; i += B) {
            double2 t[B][G], c[B][G];
            bool live[B];
#pragma unroll
            for (int b = 0; b < B; ++b) {
                const int l = l0 + (i + b) * STEP;
                live[b] = l < L;
                const int lb = live[b] ? l : 0;
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    t[b][g] = tr[g][lb];
                    if (!one_ctrl) c[b][g] = cr[g][lb];
                }
            }
            double pdv[B][G], th[B][G];
#pragma unroll
            for (int b = 0; b < B; ++b) {
                const int l = l0 + (i + b) * STEP;
                double yv[G];
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    const double te = (t[b][g].y != t[b][g].y) ? 2.0 : t[b][g].y;               // infer.py:58
                    double base;
                    if (one_ctrl) {
                        yv[g] = t[b][g].x - cm[g];
                        base = cnan[g] ? fma(te, te, 1e-2) : cc[g];
                    } else {
                        double ce = c[b][g].y;
                        if (ce != ce) ce = te;                                                   // :68
                        yv[g] = t[b][g].x - c[b][g].x;
                        base = fma(ce, ce, 1e-2);
                    }
                    pdv[b][g] = two_tps * fma(te, te, base);                                     // pd2 = 2*tau_pr_den, infer.py:70
                    th[b][g] = tps * fast_rcp(pdv[b][g]);                                        // tau/2 = tps/pd2, infer.py:83
                    if (live[b]) {
                        chk += fabs(yv[g]) + pdv[b][g];
                        ys[g * L + l] = yv[g];
                    }
                }
                double w1, hi_, lo_;
                max_min_median<G>(yv, hi_, lo_, w1);                                             // infer.py:81 (finite genes only:
                const double w2 = w1 * w1;                                                       //  the others are deferred)
                if (live[b]) {
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        // tau*(y^2 + x2*w2 - 2*x1*w1*y) = tau*(y*(y - 2*x1*w1) + x2*w2), infer.py:125
                        const double bb = fma(yv[g], fma(-tx1[g], w1, yv[g]), x.x2[g] * w2);
                        r.acc[2 * G] = fma(th[b][g], bb, r.acc[2 * G]);
                        r.acc[g] = fma(yv[g] * th[b][g], w1, r.acc[g]);
                        r.acc[G + g] = fma(th[b][g], w2, r.acc[G + g]);
                    }
                    if (WS) reinterpret_cast<double2*>(w1g)[l] = make_double2(w1, w2);
                    else {
                        w1g[l] = w1;
                        if (w2g) w2g[l] = w2;
                    }
                }
            }
            if (TM == 2) {
                tmem_store<B * G>(tpd + 2 * G * i, &pdv[0][0]);
            } else {
#pragma unroll
                for (int b = 0; b < B; ++b)
                    if (live[b]) {
#pragma unroll
                        for (int g = 0; g < G; ++g) tpdg[g * Lp + l0 + (i + b) * STEP] = pdv[b][g];
                    }
            }
            tmem_store<B * G>(tth + 2 * G * i, &th[0][0]);
        }
        tmem_wait_st();
        r.bad = !(chk <= 1.79769313486231570e308);
        return r;
    }
#else
    constexpr int B = (G >= 9) ? 1 : 2;       // lines per batch
    if (em_stage_initialises<G, TM>()) {
#pragma unroll 1
        for (int i = 0; i < n_warp; i += B) {
            double2 t[B][G], c[B][G];
            bool live[B];
#pragma unroll
            for (int b = 0; b < B; ++b) {
                const int l = l0 + (i + b) * STEP;
                live[b] = l < L;
                const int lb = live[b] ? l : 0;
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    t[b][g] = tr[g][lb];
                    c[b][g] = (ctrl_lines == 1) ? c0[g] : cr[g][lb];
                }
            }
            double pdv[B][G], th[B][G];
#pragma unroll
            for (int b = 0; b < B; ++b) {
                const int l = l0 + (i + b) * STEP;
                double yv[G];
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    const double te = (t[b][g].y != t[b][g].y) ? 2.0 : t[b][g].y;               // infer.py:58
                    double ce = c[b][g].y;
                    if (ce != ce) ce = (ctrl_lines == 1 && rowmean_fill) ? fill[g] : te;         // :63 / :68
                    yv[g] = t[b][g].x - c[b][g].x;
                    pdv[b][g] = 2.0 * (tps * (te * te + ce * ce + 1e-2));                        // pd2 = 2*tau_pr_den (exact)
                    th[b][g] = tps * fast_rcp(pdv[b][g]);                                        // tau/2 = tps/pd2, infer.py:83
                    if (live[b]) {
                        bad |= !(fabs(yv[g]) <= 1.79769313486231570e308) || !(pdv[b][g] <= 1.79769313486231570e308);
                        ys[g * L + l] = yv[g];
                    }
                }
                const double w1 = median_small<G>(yv);                                           // infer.py:81
                const double w2 = w1 * w1;
                if (live[b]) {
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        const double xw = x.x1[g] * w1;
                        r.acc[2 * G] += th[b][g] * (yv[g] * yv[g] + x.x2[g] * w2 - 2.0 * xw * yv[g]);
                        r.acc[g] += yv[g] * th[b][g] * w1;
                        r.acc[G + g] += th[b][g] * w2;
                    }
                    if (WS) reinterpret_cast<double2*>(w1g)[l] = make_double2(w1, w2);
                    else {
                        w1g[l] = w1;
                        if (w2g) w2g[l] = w2;
                    }
                }
            }
            if (TM == 2) {
                tmem_store<B * G>(tpd + 2 * G * i, &pdv[0][0]);
            } else {
#pragma unroll
                for (int b = 0; b < B; ++b)
                    if (live[b]) {
#pragma unroll
                        for (int g = 0; g < G; ++g) tpdg[g * Lp + l0 + (i + b) * STEP] = pdv[b][g];
                    }
            }
            tmem_store<B * G>(tth + 2 * G * i, &th[0][0]);
        }
        tmem_wait_st();
        r.bad = bad;
        return r;
    }
#endif
    for (int l = l0; l < L; l += B * STEP) {
        double2 t[B][G], c[B][G];
#pragma unroll
        for (int b = 0; b < B; ++b) {
            const int lb = (l + b * STEP < L) ? l + b * STEP : l;
#pragma unroll
            for (int g = 0; g < G; ++g) {
                t[b][g] = tr[g][lb];
                c[b][g] = (ctrl_lines == 1) ? c0[g] : cr[g][lb];
            }
        }
#pragma unroll
        for (int b = 0; b < B; ++b) {
            if (l + b * STEP < L) {
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    const double te = (t[b][g].y != t[b][g].y) ? 2.0 : t[b][g].y;               // infer.py:58
                    double ce = c[b][g].y;
                    if (ce != ce) ce = (ctrl_lines == 1 && rowmean_fill) ? fill[g] : te;         // :63 / :68
                    const double yv = t[b][g].x - c[b][g].x;
                    const double pd = 2.0 * (tps * (te * te + ce * ce + 1e-2));   // pd2 = 2*tau_pr_den (exact)
                    bad |= !(fabs(yv) <= 1.79769313486231570e308) || !(pd <= 1.79769313486231570e308);
                    ys[g * L + l + b * STEP] = yv;
                    tpdg[g * Lp + l + b * STEP] = pd;
                }
            }
        }
    }
    r.bad = bad;
    return r;
}

// Lean staging + initialisation for the common launch: PLAIN (no hyper-prior, no 1-D-control fill), one-warp team,
// the initialising tensor-memory variants.  Same arithmetic as the initialising branch of em_stage_gene, but everything
// that is uniform per launch is a template parameter and everything that is uniform per gene is hoisted:
//   ONE_CTRL  one control column serves all lines ([N][1] control; its sd is finite -- the kernel checks and sends genes
//             with an undefined control sd through em_stage_gene): the control's mean and its variance term
//             sd_c^2 + 1e-2 are per-guide constants;
//   the lines every lane owns (l0 + k*32 < L for all lanes, k < n_full) run without predicates, two at a time, through
//   per-guide pointers that advance by a constant; the last, partial line runs once through the predicated one-line tail;
//   a staged value that is not finite shows in the initial bound sum (th > 0 for finite data, and y^2 is a term of every
//   element's contribution), so there is no per-element check.
// em_stage_gene executed ~560 instructions per two lines at G = 4 (run-time control layout, per-element flags, address
// arithmetic); this one ~260.
template <int G> struct EmCtl { double cm[G], cc[G]; };
template <int G> struct EmX12 { double x1[G], x2[G]; };

template <int G, bool ONE_CTRL, int NB>
__device__ __forceinline__ void em_stage_lines(const double2 (&t)[NB][G], const double2 (&c)[NB][G], const EmCtl<G>& k,
                                               const double (&tx1)[G], const double (&x2)[G], const double tps,
                                               const double two_tps, double (&yv)[NB][G], double (&pdv)[NB][G],
                                               double (&th)[NB][G], double (&w1)[NB], double (&w2)[NB]) {
#pragma unroll
    for (int b = 0; b < NB; ++b)
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const double te = (t[b][g].y != t[b][g].y) ? 2.0 : t[b][g].y;                    // infer.py:58
            double base;
            if (ONE_CTRL) {
                yv[b][g] = t[b][g].x - k.cm[g];
                base = k.cc[g];
            } else {
                double ce = c[b][g].y;
                if (ce != ce) ce = te;                                                        // infer.py:68
                yv[b][g] = t[b][g].x - c[b][g].x;
                base = fma(ce, ce, 1e-2);
            }
            pdv[b][g] = two_tps * fma(te, te, base);                                          // pd2 = 2*tau_pr_den, infer.py:70
        }
    // th = tau/2 = tps/pd2 (infer.py:83), stage by stage over the NB*G independent elements
    double r[NB][G], e[NB][G];
#pragma unroll
    for (int b = 0; b < NB; ++b)
#pragma unroll
        for (int g = 0; g < G; ++g) asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r[b][g]) : "d"(pdv[b][g]));
#pragma unroll
    for (int b = 0; b < NB; ++b)
#pragma unroll
        for (int g = 0; g < G; ++g) e[b][g] = fma(-pdv[b][g], r[b][g], 1.0);
#pragma unroll
    for (int b = 0; b < NB; ++b)
#pragma unroll
        for (int g = 0; g < G; ++g) e[b][g] = fma(e[b][g], e[b][g], e[b][g]);
#pragma unroll
    for (int b = 0; b < NB; ++b)
#pragma unroll
        for (int g = 0; g < G; ++g) th[b][g] = tps * fma(r[b][g], e[b][g], r[b][g]);
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        double hi_, lo_;
        max_min_median<G>(yv[b], hi_, lo_, w1[b]);                                            // infer.py:81 (finite genes only)
        w2[b] = w1[b] * w1[b];
    }
    (void)tx1; (void)x2;
}

// The initial sums of one line: the bound sum tau*(y^2 + x2*w2 - 2*x1*w1*y) = tau*(y*(y - 2*x1*w1) + x2*w2)
// (infer.py:125) and the first X step's row sums, all formed with th = tau/2.
template <int G>
__device__ __forceinline__ void em_stage_sums(const double (&yv)[G], const double (&th)[G], const double w1, const double w2,
                                              const double (&tx1)[G], const double (&x2)[G], double (&acc)[2 * G + 1]) {
#pragma unroll
    for (int g = 0; g < G; ++g) {
        const double bb = fma(yv[g], fma(-tx1[g], w1, yv[g]), x2[g] * w2);
        acc[2 * G] = fma(th[g], bb, acc[2 * G]);
        acc[g] = fma(yv[g] * th[g], w1, acc[g]);
        acc[G + g] = fma(th[g], w2, acc[G + g]);
    }
}

template <int G, int TM, bool WS, bool ONE_CTRL>
__device__ __noinline__ EmStaged<G> em_stage_plain(const double2* __restrict__ test, const double2* __restrict__ ctrl,
                                                   const int* rows, const EmCtl<G> k, double* __restrict__ ys,
                                                   double* __restrict__ tpdg, const int L, const int Lp, const int lane,
                                                   const double tps, uint32_t tpd, uint32_t tth, const int n_warp,
                                                   const EmX12<G> x, double* __restrict__ w1g, double* __restrict__ w2g) {
    constexpr int B = (G >= 9) ? 1 : 2;       // lines per batch
    constexpr int STEP = 32;
    EmStaged<G> r;
#pragma unroll
    for (int i = 0; i < 2 * G + 1; ++i) r.acc[i] = 0.0;
    const double two_tps = 2.0 * tps;
    double tx1[G], x2[G];
    const double2* tp[G];
    const double2* cp[G];
    double* yp[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
        tx1[g] = 2.0 * x.x1[g];
        x2[g] = x.x2[g];
        tp[g] = test + (long long)rows[g] * L + lane;
        cp[g] = ONE_CTRL ? ctrl : ctrl + (long long)rows[g] * L + lane;
        yp[g] = ys + g * L + lane;
    }
    double* wp = w1g + (WS ? 2 * lane : lane);
    double* w2p = w2g ? w2g + lane : nullptr;
    double* pp[G];
    if (TM != 2) {
#pragma unroll
        for (int g = 0; g < G; ++g) pp[g] = tpdg + g * Lp + lane;
    }
    const int n_full = L / STEP;              // lines that every lane of the warp owns
    int i = 0;
#pragma unroll 1
    for (; i + B <= n_full; i += B) {
        double2 t[B][G], c[B][G];
#pragma unroll
        for (int b = 0; b < B; ++b)
#pragma unroll
            for (int g = 0; g < G; ++g) {
                t[b][g] = tp[g][b * STEP];
                if (!ONE_CTRL) c[b][g] = cp[g][b * STEP];
            }
        double yv[B][G], pdv[B][G], th[B][G], w1[B], w2[B];
        em_stage_lines<G, ONE_CTRL, B>(t, c, k, tx1, x2, tps, two_tps, yv, pdv, th, w1, w2);
#pragma unroll
        for (int b = 0; b < B; ++b) {
#pragma unroll
            for (int g = 0; g < G; ++g) yp[g][b * STEP] = yv[b][g];
            em_stage_sums<G>(yv[b], th[b], w1[b], w2[b], tx1, x2, r.acc);
            if (WS) reinterpret_cast<double2*>(wp)[b * STEP] = make_double2(w1[b], w2[b]);
            else {
                wp[b * STEP] = w1[b];
                if (w2p) w2p[b * STEP] = w2[b];
            }
            if (TM != 2) {
#pragma unroll
                for (int g = 0; g < G; ++g) pp[g][b * STEP] = pdv[b][g];
            }
        }
        if (TM == 2) tmem_store<B * G>(tpd, &pdv[0][0]);
        tmem_store<B * G>(tth, &th[0][0]);
        tpd += 2 * G * B;
        tth += 2 * G * B;
#pragma unroll
        for (int g = 0; g < G; ++g) {
            tp[g] += B * STEP;
            if (!ONE_CTRL) cp[g] += B * STEP;
            yp[g] += B * STEP;
            if (TM != 2) pp[g] += B * STEP;
        }
        wp += (WS ? 2 : 1) * B * STEP;
        if (w2p) w2p += B * STEP;
    }
    // the left-over lines, one at a time: a full line when n_full is odd, then the warp's partial line
#pragma unroll 1
    for (; i < n_warp; ++i) {
        const bool live = lane + i * STEP < L;
        double2 t[1][G], c[1][G];
#pragma unroll
        for (int g = 0; g < G; ++g) {
            // a lane without this line reads its own line 0 instead (mode 2 runs at L >= 32) and discards the result
            t[0][g] = live ? tp[g][0] : test[(long long)rows[g] * L + lane];
            if (!ONE_CTRL) c[0][g] = live ? cp[g][0] : ctrl[(long long)rows[g] * L + lane];
        }
        double yv[1][G], pdv[1][G], th[1][G], w1[1], w2[1];
        em_stage_lines<G, ONE_CTRL, 1>(t, c, k, tx1, x2, tps, two_tps, yv, pdv, th, w1, w2);
        if (live) {
#pragma unroll
            for (int g = 0; g < G; ++g) yp[g][0] = yv[0][g];
            em_stage_sums<G>(yv[0], th[0], w1[0], w2[0], tx1, x2, r.acc);
            if (WS) reinterpret_cast<double2*>(wp)[0] = make_double2(w1[0], w2[0]);
            else {
                wp[0] = w1[0];
                if (w2p) w2p[0] = w2[0];
            }
            if (TM != 2) {
#pragma unroll
                for (int g = 0; g < G; ++g) pp[g][0] = pdv[0][g];
            }
        }
        if (TM == 2) tmem_store<G>(tpd, &pdv[0][0]);
        tmem_store<G>(tth, &th[0][0]);
        tpd += 2 * G;
        tth += 2 * G;
#pragma unroll
        for (int g = 0; g < G; ++g) {
            tp[g] += STEP;
            if (!ONE_CTRL) cp[g] += STEP;
            yp[g] += STEP;
            if (TM != 2) pp[g] += STEP;
        }
        wp += (WS ? 2 : 1) * STEP;
        if (w2p) w2p += STEP;
    }
    tmem_wait_st();
    // finite <=> every staged value of this lane is (see above)
    r.bad = !(fabs(r.acc[2 * G]) <= 1.79769313486231570e308);
    return r;
}

// Shared memory of one team: y and tau [G][L], the reduction tiles and totals.  tau_pr_den, E[w] and E[w^2]
// live in global memory (L2): 27.5 KB per gene at G = 4, L = 400 -> 8 genes (two warps per scheduler) per SM.
__host__ __device__ constexpr long long em_fast_smem_doubles(int G, int W, int L) {
    return 2LL * G * L + (long long)W * (2 * G + 1) * 33 + 2LL * W * (2 * G + 1);
}
// TMEM variant: four independent one-warp teams per CTA, tau in tensor memory; per team y [G][L] plus the reduction
// tile and totals, rounded to 128 bytes.  Two CTAs (eight genes) per SM: with three, the 168-register budget makes
// ptxas spill in the sweep, and 12 genes at 3 lines per step measured no faster than 8 at 4 (7.9 ms at L = 400).
constexpr int kTmemTeams = 4;
#ifndef JACKS_TM_CTAS
#define JACKS_TM_CTAS 2
#endif
__host__ __device__ constexpr long long em_fast_tmem_team_doubles(int G, int L, bool ws = false) {
    // ws: plus (E[w], E[w^2]) pairs of the running pass, [L][2]
    return ((long long)G * L + (2 * G + 1) * 33 + 2 * (2 * G + 1) + (ws ? 2LL * L : 0) + 15) / 16 * 16;
}

// L2 prefetch of the G guide rows (test + control) of the gene a team will run next.  One rolled loop over
// all (row, 128-byte chunk) pairs: this runs once per gene and must stay small in code.
template <int G>
__device__ __forceinline__ void prefetch_gene_l2(const double2* test, const double2* ctrl, const int* slot_rows, int L,
                                                 int ctrl_lines, int tid, int nthreads) {
    const int nchunk = (L * 16 + 127) >> 7;
#pragma unroll 1
    for (int i = tid; i < G * nchunk; i += nthreads) {
        const int g = i / nchunk;
        const long long row = slot_rows[g];
        const int b = (i - g * nchunk) << 7;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(test + row * L) + b));
        if (ctrl_lines > 1) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(ctrl + row * ctrl_lines) + b));
    }
}

// PLAIN: the launch uses neither the w hyper-prior (infer.py:92) nor the 1-D-control sd fill (infer.py:63); both
// are uniform per launch, and compiling them out keeps the common kernel's code (and its instruction-cache
// footprint) small.
// TM: the tensor-memory variant (W must be 1): a CTA is kTmemTeams independent one-warp teams, each with its own
// gene, shared-memory block and TMEM lane quarter; th = tau/2 lives in TMEM columns [line][g] of the thread's lane.
// Nothing synchronises the teams between the TMEM allocation at the start and its release at the end.
// TM: 0 shared-memory variant; 1 tau in TMEM; 2 tau and tau_pr_den in TMEM (no global slab, no L2 traffic in a pass).
template <int G, int W, bool UNIT_C, bool PLAIN, int TM, bool WS = false>
__global__ void __launch_bounds__(TM ? kTmemTeams * 32 : W * 32, TM ? JACKS_TM_CTAS : 0) em_fast_kernel(const EmLaunch a) {
    static_assert(!TM || W == 1, "the TMEM variant runs one-warp teams");
    static_assert(!WS || TM == 2, "E[w] is kept in shared memory by the tensor-memory mode 2 kernel only");
    constexpr int NV = 2 * G + 1;
    constexpr int STEP = 32 * W;
    constexpr int UMAX = em_sweep_lines<G, TM>();
    extern __shared__ double smem[];
    const int L = a.L;
    const int lane = threadIdx.x & 31;
    const int cta_warp = threadIdx.x >> 5;
    const int warp = TM ? 0 : cta_warp;                 // warp index inside the team
    const int tid = TM ? lane : (int)threadIdx.x;       // thread index inside the team
    const int team = TM ? blockIdx.x * kTmemTeams + cta_warp : blockIdx.x;
    double* ys = smem + (TM ? cta_warp * em_fast_tmem_team_doubles(G, L, WS) : 0);   // [G][L]
    double* taus = TM ? nullptr : ys + (size_t)G * L;   // [G][L]  th = tau/2 (see em_sweep); TMEM columns if TM
    double* tile = TM ? ys + (size_t)G * L : taus + (size_t)G * L;
    double* scr = tile + warp * (NV * 33);              // this warp's reduction tile
    double* red = tile + W * (NV * 33);                 // [2][W][NV] totals
    double* wb = red + 2 * W * NV;                      // WS: (E[w], E[w^2]) of the running pass, [L][2]
    wb += (reinterpret_cast<unsigned long long>(wb) & 8) ? 1 : 0;       // 16-byte aligned pairs
    const int Lp = a.tpd_stride;
    double* tpdg = a.tpd_scratch + (size_t)team * G * Lp;                  // this team's pd2 slab [G][Lp] (global, L2)
    __shared__ int s_slot[2];
    __shared__ uint32_t s_tmem;
    uint32_t tbase = 0;
    if (TM) {
        if (cta_warp == 0) tmem_alloc(&s_tmem, (uint32_t)a.tmem_cols);
        tmem_fence_before_sync();
        __syncthreads();
        tmem_fence_after_sync();
        tbase = s_tmem + ((uint32_t)(cta_warp * 32) << 16);      // this warp's lane quarter
    }
    const uint32_t pdoff = (uint32_t)a.tmem_pd_off;              // columns of pd2 behind those of tau (TM == 2)

    const double tps = a.tps;
    const double c_tau = a.tps + 0.5;            // infer.py:120
    const double ivx = 1.0 / a.var0_x;           // infer.py:104
    const double mx = a.mu0_x / a.var0_x;
    const bool fixed = (a.fixed_x1 != nullptr);
    int parity = 0;

    // Work distribution: teams pull bucket slots from a global counter one gene AHEAD of the one they
    // are computing.  The atomic is issued at the top of a gene and consumed after its staging phase;
    // the next gene's index entries are loaded then and consumed after the init phase, where an L2
    // prefetch of its rows is issued -- a team has only its own warps to hide latency behind.
    int slot, gid = 0, rows[G];
    long long off = 0;
    {
        int first = 0;
        if (W == 1) {
            if (lane == 0) first = atomicAdd(a.work_counter, 1);
            first = __shfl_sync(JACKS_FULL_MASK, first, 0);
        } else {
            if (tid == 0) s_slot[0] = atomicAdd(a.work_counter, 1);
            __syncthreads();
            first = s_slot[0];
        }
        slot = first;
        if (slot < a.n_work) {
            gid = a.gene_ids[slot];
            off = a.slot_off[slot];
#pragma unroll
            for (int g = 0; g < G; ++g) rows[g] = a.slot_rows[(long long)slot * G + g];
        }
    }
    int pipe = 1;

    while (slot < a.n_work) {
        int pending = 0;                       // raw result of the look-ahead atomic (thread 0 only)
        if (tid == 0) pending = atomicAdd(a.work_counter, 1);
        // rotate which warp owns the (possibly partial) last strip so the SM sub-partitions stay even
        const int wv = (W == 1) ? 0 : (warp + slot) % W;
        const int l0 = wv * 32 + lane;
        const int n_warp = (L - wv * 32 + STEP - 1) / STEP;     // lines of this warp's fullest lane (lane 0)
        int plan_B, plan_big;
        em_step_plan(n_warp, UMAX, plan_B, plan_big);

        // ---- stage y and tau_pr_den (infer.py:58-70) ---------------------------------------
        double x1r[G], x2r[G], fill[G];
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const int row = rows[g];
            if (fixed) { x1r[g] = a.fixed_x1[row]; x2r[g] = a.fixed_x2[row]; }
            else { x1r[g] = a.mu0_x; x2r[g] = a.mu0_x * a.mu0_x; }
            fill[g] = 0.0;
            if (!PLAIN && a.ctrl_lines == 1 && a.ctrl_rowmean_fill) {
                const double cs = a.ctrl[(long long)row].y;
                if (cs != cs) {
                    // 1-D control with an undefined sd: row mean of the (NaN -> 2.0) test sds, infer.py:63
                    const double2* tr = a.test + (long long)row * L;
                    double one[1] = {0.0};
#pragma unroll 1
                    for (int l = l0; l < L; l += STEP) {
                        const double e = tr[l].y;
                        one[0] += (e != e) ? 2.0 : e;
                    }
                    team_allsum<1, NV, W>(one, scr, red, parity, warp, lane);
                    fill[g] = one[0] / (double)L;
                }
            }
        }
        double* const w2o = a.w2 ? a.w2 + (long long)gid * L : nullptr;       // the gene's output rows
        double* const w1o = a.w1 + (long long)gid * L;
        double* const w2g = WS ? wb + 1 : w2o;                                  // where the passes keep E[w^2], E[w]
        double* const w1g = WS ? wb : w1o;
        double acc[NV];
        bool bad;
        {
            EmX<G> xv;
#pragma unroll
            for (int g = 0; g < G; ++g) { xv.x1[g] = x1r[g]; xv.x2[g] = x2r[g]; xv.tx1[g] = 0.0; xv.tx2[g] = 0.0; }
            EmStaged<G> sg;
            bool lean = false;
#ifndef JACKS_STAGE_OLD
            if (PLAIN && W == 1 && em_stage_initialises<G, TM>()) {
                // the common launch: lean staging (em_stage_plain); a gene whose single control column has an undefined sd
                // (matched rule, infer.py:68: it takes each line's test sd) goes through the general function
                EmX12<G> xs;
                EmCtl<G> k;
#pragma unroll
                for (int g = 0; g < G; ++g) { xs.x1[g] = x1r[g]; xs.x2[g] = x2r[g]; k.cm[g] = 0.0; k.cc[g] = 0.0; }
                if (a.ctrl_lines == 1) {
                    bool cn = false;
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        const double2 c0 = a.ctrl[rows[g]];
                        k.cm[g] = c0.x;
                        k.cc[g] = fma(c0.y, c0.y, 1e-2);
                        cn |= (c0.y != c0.y);
                    }
                    if (!cn) {
                        sg = em_stage_plain<G, TM, WS, true>(a.test, a.ctrl, rows, k, ys, tpdg, L, Lp, lane, tps, tbase + pdoff, tbase, n_warp, xs, w1g, w2g);
                        lean = true;
                    }
                } else {
                    sg = em_stage_plain<G, TM, WS, false>(a.test, a.ctrl, rows, k, ys, tpdg, L, Lp, lane, tps, tbase + pdoff, tbase, n_warp, xs, w1g, w2g);
                    lean = true;
                }
            }
#endif
            if (!lean)
                sg = em_stage_gene<G, STEP, TM, WS>(a.test, a.ctrl, rows, ys, tpdg, L, Lp, l0, a.ctrl_lines, a.ctrl_rowmean_fill,
                                                    fill, tps, tbase + pdoff, tbase, n_warp, xv, w1g, w2g);
            bad = sg.bad;
#pragma unroll
            for (int i = 0; i < NV; ++i) acc[i] = sg.acc[i];       // the initial sums if staging initialises; otherwise zeros
        }
        const int any_bad = (W == 1) ? __any_sync(JACKS_FULL_MASK, bad) : __syncthreads_or(bad);
        // the look-ahead atomic has landed by now: share it and start loading the next gene's index
        int nslot;
        if (W == 1) {
            nslot = __shfl_sync(JACKS_FULL_MASK, pending, 0);
        } else {
            if (tid == 0) s_slot[pipe] = pending;
            __syncthreads();
            nslot = s_slot[pipe];
            pipe ^= 1;
        }
        int ngid = 0, nrows[G];
        long long noff = 0;
#pragma unroll
        for (int g = 0; g < G; ++g) nrows[g] = 0;
        if (nslot < a.n_work) {
            ngid = a.gene_ids[nslot];
            noff = a.slot_off[nslot];
#pragma unroll
            for (int g = 0; g < G; ++g) nrows[g] = a.slot_rows[(long long)nslot * G + g];
        }
        if (any_bad) {
            if (tid == 0) {
                const int p = atomicAdd(a.deferred_count, 1);
                a.deferred_ids[p] = gid;
            }
        } else {
            // every lane reads back only the words (shared and global) it wrote itself: no barrier needed

            // ---- initial w, tau, bound and the first X-step sums (infer.py:81-86) --------------
            if (em_stage_initialises<G, TM>()) {
                // staged and initialised in one pass (em_stage_gene)
            } else if (TM) {
                // warp-uniform trip count: the TMEM stores are warp-collective; a lane past its last line
                // computes on line 0 and stores a value nobody uses
#pragma unroll 1
                for (int i = 0; i < n_warp; ++i) {
                    const int l = l0 + i * STEP;
                    const bool live = l < L;
                    const int lu = live ? l : 0;
                    double yv[G], th[G], pdv[G];
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        yv[g] = ys[g * L + lu];
                        pdv[g] = tpdg[g * Lp + lu];
                    }
                    const double w1 = median_small<G>(yv);
                    const double w2 = w1 * w1;
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        th[g] = tps * fast_rcp(pdv[g]);                    // th = tau/2 = tps/pd2, infer.py:83
                        if (live) {
                            const double xw = x1r[g] * w1;
                            acc[2 * G] += th[g] * (yv[g] * yv[g] + x2r[g] * w2 - 2.0 * xw * yv[g]);
                            acc[g] += yv[g] * th[g] * w1;
                            acc[G + g] += th[g] * w2;
                        }
                    }
                    tmem_store<G>(tbase + 2 * G * i, th);
                    if (live) {
                        w1g[l] = w1;
                        if (w2g) w2g[l] = w2;
                    }
                }
            } else {
#pragma unroll 2
                for (int l = l0; l < L; l += STEP) {
                    double yv[G];
#pragma unroll
                    for (int g = 0; g < G; ++g) yv[g] = ys[g * L + l];
                    const double w1 = median_small<G>(yv);
                    const double w2 = w1 * w1;
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        const double t = tps * fast_rcp(tpdg[g * Lp + l]);      // th = tau/2 = tps/pd2, infer.py:83
                        taus[g * L + l] = t;
                        const double xw = x1r[g] * w1;
                        acc[2 * G] += t * (yv[g] * yv[g] + x2r[g] * w2 - 2.0 * xw * yv[g]);
                        acc[g] += yv[g] * t * w1;
                        acc[G + g] += t * w2;
                    }
                    if (WS) reinterpret_cast<double2*>(w1g)[l] = make_double2(w1, w2);
                    else {
                        w1g[l] = w1;
                        if (w2g) w2g[l] = w2;
                    }
                }
            }
            team_allsum<NV, NV, W>(acc, scr, red, parity, warp, lane);
            // every sum was formed with th = tau/2: acc[0..2G) stay HALF sums (the X step doubles them in its
            // first FMA), the bound is doubled here (exact)
            double bound = 2.0 * acc[2 * G];
            const double bound_k = 2.0 * c_tau * (double)(G * L);   // sum tau*b* = 2c*G*L - 2*sum pd2*th
            double mu0_w = a.mu0_w, var0_w = a.var0_w;
            double ivw = 1.0 / var0_w;              // infer.py:113
            double mw = mu0_w / var0_w;
            int iters = 0;

            // next gene: its index entries have arrived during the init sweep; pull its rows towards L2
            if (nslot < a.n_work)
                prefetch_gene_l2<G>(a.test, a.ctrl, a.slot_rows + (long long)nslot * G, L, a.ctrl_lines, tid, STEP);

            // ---- EM passes (infer.py:89-98) ------------------------------------------------------
            for (int it = 0; it < a.n_iter; ++it) {
                const double last = bound;
                if (!fixed) {
                    // X step, infer.py:103-110, evaluated redundantly by every lane (all lanes hold all sums)
                    double xu[G], x2u[G];
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        const double rden = fast_rcp(fma(2.0, acc[G + g], ivx));
                        xu[g] = fma(2.0, acc[g], mx) * rden;
                        x2u[g] = fma(xu[g], xu[g], rden);
                    }
                    double sum = xu[0], hi, lo, med;
#pragma unroll
                    for (int g = 1; g < G; ++g) sum += xu[g];
                    max_min_median<G>(xu, hi, lo, med);
                    const double wadj = 0.5 / G;
                    const double x1m = sum * (1.0 / G) + 2 * wadj * med - wadj * hi - wadj * lo;
                    const double rm = fast_rcp(x1m);
                    const double rm2 = rm * rm;
#pragma unroll
                    for (int g = 0; g < G; ++g) { x1r[g] = xu[g] * rm; x2r[g] = x2u[g] * rm2; }
                }
                if (!PLAIN && a.apply_w_hp && L > 1) {
                    // hierarchical prior from the current w1, infer.py:92 (population variance)
                    double m[1] = {0.0};
                    for (int l = l0; l < L; l += STEP) m[0] += w1g[WS ? 2 * l : l];
                    team_allsum<1, NV, W>(m, scr, red, parity, warp, lane);
                    const double mean = m[0] / (double)L;
                    double v[1] = {0.0};
                    for (int l = l0; l < L; l += STEP) { const double d = w1g[WS ? 2 * l : l] - mean; v[0] += d * d; }
                    team_allsum<1, NV, W>(v, scr, red, parity, warp, lane);
                    mu0_w = mean;
                    var0_w = (v[0] / (double)L) * 3 + 1e-4;
                    ivw = 1.0 / var0_w;
                    mw = mu0_w / var0_w;
                }
                {
                    EmXs<G> xv;
#pragma unroll
                    for (int g = 0; g < G; ++g) { xv.x2[g] = x2r[g]; xv.tx1[g] = 2.0 * x1r[g]; xv.tx2[g] = 2.0 * x2r[g]; }
                    EmSums<G> r;
#define JACKS_SWEEP(BB) r = em_sweep_all<G, STEP, UNIT_C, TM, (BB <= UMAX ? BB : 1), WS>(tbase, pdoff, ys, tpdg, Lp, taus, w1g, w2g, L, l0, plan_big, xv, ivw, mw, c_tau)
                    if (UMAX >= 4 && plan_B == 4) JACKS_SWEEP(4);
                    else if (UMAX >= 3 && plan_B == 3) JACKS_SWEEP(3);
                    else if (UMAX >= 2 && plan_B == 2) JACKS_SWEEP(2);
                    else JACKS_SWEEP(1);
#undef JACKS_SWEEP
#pragma unroll
                    for (int g = 0; g < G; ++g) { acc[g] = r.a1[g]; acc[G + g] = r.a2[g]; }
                    acc[2 * G] = r.ab;
                }
                team_allsum<NV, NV, W>(acc, scr, red, parity, warp, lane);
                bound = fma(-2.0, acc[2 * G], bound_k);
                ++iters;
                if (fabs(last - bound) < a.tol) break;
            }

            // ---- write results ------------------------------------------------------------------
            // results are written once and never read back here: streaming stores (evict-first) keep them from pushing the
            // pseudo-genes' control rows and the prefetched rows of the next genes out of L2 (6.44 -> 6.31 ms per Avana pass)
            if (WS) {
                for (int l = l0; l < L; l += STEP) {
                    const double2 ww = reinterpret_cast<const double2*>(wb)[l];
                    __stcs(&w1o[l], ww.x);
                    if (w2o) __stcs(&w2o[l], ww.y);
                }
            }
            if (TM && a.tau) {
                tmem_wait_st();
#pragma unroll 1
                for (int i = 0; i < n_warp; ++i) {
                    const int l = l0 + i * STEP;
                    double th[G];
                    tmem_load<G>(tbase + 2 * G * i, th);
                    if (l < L) {
#pragma unroll
                        for (int g = 0; g < G; ++g) {
                            __stcs(&a.tau[(off + g) * L + l], 2.0 * th[g]);
                        }
                    }
                }
            }
            for (int l = l0; l < L; l += STEP) {
                if (!TM && a.tau) {
#pragma unroll
                    for (int g = 0; g < G; ++g) a.tau[(off + g) * L + l] = 2.0 * taus[g * L + l];
                }
                if (a.y) {
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        __stcs(&a.y[(off + g) * L + l], ys[g * L + l]);
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
        // ---- rotate the software pipeline: the prefetched gene becomes current -----------------
        if (W > 1) __syncthreads();            // lanes change line ownership between genes (rotation)
        slot = nslot; gid = ngid; off = noff;
#pragma unroll
        for (int g = 0; g < G; ++g) rows[g] = nrows[g];
    }
    if (TM) {
        tmem_fence_before_sync();
        __syncthreads();
        if (cta_warp == 0) tmem_dealloc(s_tmem, (uint32_t)a.tmem_cols);
    }
}

}  // namespace jacks
