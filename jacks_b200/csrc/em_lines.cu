// Per-line EM ("single JACKS", reference 2018_paper_materials/scripts/jacks/run_JACKS_single.py:83-85:
// inferJACKS(gene_index, testdata[:,[ts],:], ctrldata[:,[ts],:]) for every line ts): each (gene, line) pair is
// an independent EM with L = 1, i.e. G elements of state.  A warp-per-gene team would idle 31 lanes, so here
// one LANE owns one problem with everything in registers; consecutive lanes take consecutive lines of a gene,
// so the 16-byte (mean, sd) loads of a warp are contiguous.  Lanes of a warp run different pass counts; the
// hardware masks the finished ones.  Same arithmetic as infer.py:55-125 with L = 1 (the hierarchical prior
// never applies: infer.py:92 requires L > 1); finite data only -- a problem holding NaN/inf is appended to
// the deferred queue and run by the NaN-safe generic kernel (em_generic.cu, lines_mode 2).
#include "em_fast.cuh"

namespace jacks {

template <int G>
__global__ void __launch_bounds__(128) em_lines_kernel(const EmLaunch a) {
    const int L = a.L;
    const long long total = (long long)a.n_work * L;
    const double tps = a.tps;
    const double c_tau = a.tps + 0.5;            // infer.py:120
    const double ivx = 1.0 / a.var0_x;           // infer.py:104
    const double mx = a.mu0_x / a.var0_x;
    const double ivw = 1.0 / a.var0_w;           // infer.py:113
    const double mw = a.mu0_w / a.var0_w;
    const bool fixed = (a.fixed_x1 != nullptr);
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (long long)gridDim.x * blockDim.x) {
        const int slot = (int)(p / L);
        const int line = (int)(p - (long long)slot * L);
        const int gid = a.gene_ids[slot];
        const long long off = a.slot_off[slot];
        // ---- set-up, infer.py:58-70 (matched control only: the per-line slices always have equal shapes) ----
        double y[G], pd[G], x1[G], x2[G], tau[G];
        bool bad = false;
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const int row = a.slot_rows[(long long)slot * G + g];
            const double2 t = a.test[(long long)row * L + line];
            const double2 c = a.ctrl[(long long)row * L + line];
            const double te = (t.y != t.y) ? 2.0 : t.y;               // infer.py:58
            const double ce = (c.y != c.y) ? te : c.y;                // infer.py:68
            y[g] = t.x - c.x;
            pd[g] = tps * (te * te + ce * ce + 1e-2);
            bad |= !(fabs(y[g]) <= 1.79769313486231570e308) || !(pd[g] <= 1.79769313486231570e308);
            if (fixed) { x1[g] = a.fixed_x1[row]; x2[g] = a.fixed_x2[row]; }
            else { x1[g] = a.mu0_x; x2[g] = a.mu0_x * a.mu0_x; }
        }
        if (bad) {
            const int q = atomicAdd(a.deferred_count, 1);
            a.deferred_ids[q] = (int)((long long)gid * L + line);
            continue;
        }
        // ---- init, infer.py:81-86 ----------------------------------------------------------------------
        double w1 = median_small<G>(y);
        double w2 = w1 * w1;
        double bound = 0.0, s1[G], s2[G];
#pragma unroll
        for (int g = 0; g < G; ++g) {
            tau[g] = tps * fast_rcp(pd[g]);
            const double xw = x1[g] * w1;
            bound += tau[g] * (y[g] * y[g] + x2[g] * w2 - 2.0 * xw * y[g]);
            s1[g] = y[g] * tau[g] * w1;
            s2[g] = tau[g] * w2;
        }
        int iters = 0;
        // ---- EM passes, infer.py:89-98 -------------------------------------------------------------------
        for (int it = 0; it < a.n_iter; ++it) {
            const double last = bound;
            if (!fixed) {
                // X step, infer.py:103-110
                double xu[G], x2u[G];
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    const double rden = fast_rcp(s2[g] + ivx);
                    xu[g] = (mx + s1[g]) * rden;
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
                for (int g = 0; g < G; ++g) { x1[g] = xu[g] * rm; x2[g] = x2u[g] * rm2; }
            }
            // W step, infer.py:112-116
            double sa = ivw, sb = mw;
#pragma unroll
            for (int g = 0; g < G; ++g) {
                sa = fma(x2[g], tau[g], sa);
                sb = fma(x1[g], y[g] * tau[g], sb);
            }
            const double r = fast_rcp(sa);
            w1 = sb * r;
            w2 = fma(w1, w1, r);
            // tau step (infer.py:118-121), bound (infer.py:123-125), sums of the next X step
            bound = 0.0;
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const double bs = fma(y[g], fma(x1[g], -2.0 * w1, y[g]), x2[g] * w2);
                double t = fast_rcp(fma(0.5, bs, pd[g]));
                t *= c_tau;
                tau[g] = t;
                bound = fma(t, bs, bound);
                s1[g] = y[g] * t * w1;
                s2[g] = t * w2;
            }
            ++iters;
            if (fabs(last - bound) < a.tol) break;
        }
        // ---- results: the joint layouts plus a line axis ---------------------------------------------------
        const long long pid = (long long)gid * L + line;
        a.w1[pid] = w1;
        if (a.w2) a.w2[pid] = w2;
        if (a.n_iters) a.n_iters[pid] = iters;
        if (a.bound) a.bound[pid] = bound;
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const long long q = (off + g) * L + line;
            if (a.x1) a.x1[q] = x1[g];
            if (a.x2) a.x2[q] = x2[g];
            if (a.y) a.y[q] = y[g];
            if (a.tau) a.tau[q] = tau[g];
        }
    }
}

template <int G>
static cudaError_t launch_lines(const EmLaunch& a, int n_sm, cudaStream_t st) {
    const long long total = (long long)a.n_work * a.L;
    if (total <= 0) return cudaSuccess;
    long long grid = (total + 127) / 128;
    const long long cap = (long long)n_sm * 32;       // grid-stride beyond two waves of 16 CTAs per SM
    if (grid > cap) grid = cap;
    em_lines_kernel<G><<<(int)grid, 128, 0, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_em_lines(int G, const EmLaunch& a, int n_sm, cudaStream_t st) {
    switch (G) {
        case 1: return launch_lines<1>(a, n_sm, st);
        case 2: return launch_lines<2>(a, n_sm, st);
        case 3: return launch_lines<3>(a, n_sm, st);
        case 4: return launch_lines<4>(a, n_sm, st);
        case 5: return launch_lines<5>(a, n_sm, st);
        case 6: return launch_lines<6>(a, n_sm, st);
        case 7: return launch_lines<7>(a, n_sm, st);
        case 8: return launch_lines<8>(a, n_sm, st);
        case 9: return launch_lines<9>(a, n_sm, st);
        case 10: return launch_lines<10>(a, n_sm, st);
        case 11: return launch_lines<11>(a, n_sm, st);
        case 12: return launch_lines<12>(a, n_sm, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace jacks
