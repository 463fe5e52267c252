/*
 * C restatement of the JACKS per-gene EM -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Follows felicityallen/JACKS jacks/jacks/infer.py line by line (inferJACKSGene :55-99, updateX
 * :103-110, updateW :112-116, updateTau :118-121, lowerBound :123-125, nandot :36-45) with numpy's
 * NaN rules (nansum / nanmedian ignore NaN; mean/max/min do not).  Sums run in index order, so
 * results agree with numpy to rounding (~1e-15 relative), not bit for bit; the numpy restatement
 * oracle/jacks_oracle.py is the bit-exact one and tests/test_oracle_golden.py pins this file to it
 * and to the reference's golden outputs.  Used only by tests/, smoke() and bench.py's CPU baseline
 * (full-size checks where the numpy oracle would take minutes).  Built by oracle/build_oracle.py:
 *     gcc -O2 -fPIC -shared -fopenmp -ffp-contract=off -o oracle/libem_oracle.so oracle/em_oracle.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int32_t n_iter, apply_w_hp;
    double tol, mu0_x, var0_x, mu0_w, var0_w, tau_prior_strength;
} OracleEmParams;

static double nz(double v) { return v == v ? v : 0.0; }

static int cmp_double(const void* a, const void* b) {
    double x = *(const double*)a, y = *(const double*)b;
    return (x > y) - (x < y);
}

/* numpy.nanmedian of n values with the given stride */
static double nanmedian(const double* v, int n, long stride, double* tmp) {
    int m = 0;
    for (int i = 0; i < n; ++i) if (v[i * stride] == v[i * stride]) tmp[m++] = v[i * stride];
    if (m == 0) return NAN;
    qsort(tmp, m, sizeof(double), cmp_double);
    return (m & 1) ? tmp[m / 2] : 0.5 * (tmp[m / 2 - 1] + tmp[m / 2]);
}

/* One gene.  test/ctrl: [N][L][2] / [N][ctrl_lines][2]; rows[G].  Outputs may be NULL. */
static void em_gene(const double* test, const double* ctrl, int L, int ctrl_lines, int rowmean_fill,
                    const int32_t* rows, int G, const double* fx1, const double* fx2, const OracleEmParams* p,
                    double* x1o, double* x2o, double* w1o, double* w2o, double* yo, double* tauo,
                    int32_t* iters_o, double* bound_o) {
    const double tps = p->tau_prior_strength;
    size_t GL = (size_t)G * L;
    double* y = (double*)malloc(sizeof(double) * (3 * GL + 2 * (size_t)L + 3 * (size_t)G + (size_t)(G > L ? G : L)));
    double* tpd = y + GL; double* tau = tpd + GL; double* w1 = tau + GL; double* w2 = w1 + L;
    double* x1 = w2 + L; double* x2 = x1 + G; double* tmp = x2 + G;
    for (int g = 0; g < G; ++g) {
        const double* tr = test + (size_t)rows[g] * L * 2;
        const double* cr = ctrl + (size_t)rows[g] * ctrl_lines * 2;
        double fill = 0.0;
        if (ctrl_lines == 1 && rowmean_fill && cr[1] != cr[1]) {           /* infer.py:63 */
            for (int l = 0; l < L; ++l) { double e = tr[2 * l + 1]; fill += (e != e) ? 2.0 : e; }
            fill /= L;
        }
        for (int l = 0; l < L; ++l) {
            double te = tr[2 * l + 1];
            if (te != te) te = 2.0;                                         /* :58 */
            const double* c = (ctrl_lines == 1) ? cr : cr + 2 * l;
            double ce = c[1];
            if (ce != ce) ce = (ctrl_lines == 1 && rowmean_fill) ? fill : te;   /* :63 / :68 */
            y[g * (size_t)L + l] = tr[2 * l] - c[0];
            tpd[g * (size_t)L + l] = tps * 1.0 * (te * te + ce * ce + 1e-2);
        }
        if (fx1) { x1[g] = fx1[rows[g]]; x2[g] = fx2[rows[g]]; }
        else { x1[g] = p->mu0_x; x2[g] = x1[g] * x1[g]; }
    }
    for (int l = 0; l < L; ++l) { w1[l] = nanmedian(y + l, G, L, tmp); w2[l] = w1[l] * w1[l]; }   /* :81,:85 */
    double bound = 0.0;
    for (size_t i = 0; i < GL; ++i) {
        int g = (int)(i / L), l = (int)(i % L);
        tau[i] = tps * 1.0 / tpd[i];                                        /* :83 */
        bound += nz(tau[i] * (y[i] * y[i] + x2[g] * w2[l] - 2 * (x1[g] * w1[l]) * y[i]));
    }
    double mu0_w = p->mu0_w, var0_w = p->var0_w;
    int iters = 0;
    for (int it = 0; it < p->n_iter; ++it) {
        double last = bound;
        if (!fx1) {                                                         /* updateX :103-110 */
            double sum = 0.0, hi = -INFINITY, lo = INFINITY;
            int anynan = 0;
            for (int g = 0; g < G; ++g) {
                double s1 = 0.0, s2 = 0.0;
                for (int l = 0; l < L; ++l) {
                    size_t i = g * (size_t)L + l;
                    s1 += nz((y[i] * tau[i]) * w1[l]);
                    s2 += nz(tau[i] * w2[l]);
                }
                double den = s2 + 1.0 / p->var0_x;
                x1[g] = (p->mu0_x / p->var0_x + s1) / den;
                x2[g] = x1[g] * x1[g] + 1.0 / den;
                sum += x1[g];
                if (x1[g] != x1[g]) anynan = 1;
                if (x1[g] > hi) hi = x1[g];
                if (x1[g] < lo) lo = x1[g];
            }
            if (anynan) { hi = NAN; lo = NAN; }
            double wadj = 0.5 / G;
            double x1m = sum / G + 2 * wadj * nanmedian(x1, G, 1, tmp) - wadj * hi - wadj * lo;
            for (int g = 0; g < G; ++g) { x1[g] = x1[g] / x1m; x2[g] = x2[g] / x1m / x1m; }
        }
        if (p->apply_w_hp && L > 1) {                                       /* :92 */
            double m = 0.0, v = 0.0;
            for (int l = 0; l < L; ++l) m += w1[l];
            m /= L;
            for (int l = 0; l < L; ++l) v += (w1[l] - m) * (w1[l] - m);
            mu0_w = m; var0_w = (v / L) * 3 + 1e-4;
        }
        for (int l = 0; l < L; ++l) {                                       /* updateW :112-116 */
            double a = 0.0, b = 0.0;
            for (int g = 0; g < G; ++g) {
                size_t i = g * (size_t)L + l;
                a += nz(x2[g] * tau[i]);
                b += nz(x1[g] * (y[i] * tau[i]));
            }
            double den = a + 1.0 / var0_w;
            w1[l] = (mu0_w / var0_w + b) / den;
            w2[l] = w1[l] * w1[l] + 1.0 / den;
        }
        bound = 0.0;
        for (size_t i = 0; i < GL; ++i) {                                   /* updateTau :118-121, lowerBound :123-125 */
            int g = (int)(i / L), l = (int)(i % L);
            double xw = x1[g] * w1[l], xw2 = x2[g] * w2[l];
            double bstar = y[i] * y[i] - 2 * y[i] * xw + xw2;
            tau[i] = (tps + 0.5) / (tpd[i] + 0.5 * bstar);
            bound += nz(tau[i] * (y[i] * y[i] + xw2 - 2 * xw * y[i]));
        }
        ++iters;
        if (fabs(last - bound) < p->tol) break;                             /* :97 */
    }
    if (x1o) memcpy(x1o, x1, sizeof(double) * G);
    if (x2o) memcpy(x2o, x2, sizeof(double) * G);
    if (w1o) memcpy(w1o, w1, sizeof(double) * L);
    if (w2o) memcpy(w2o, w2, sizeof(double) * L);
    if (yo) memcpy(yo, y, sizeof(double) * GL);
    if (tauo) memcpy(tauo, tau, sizeof(double) * GL);
    if (iters_o) *iters_o = iters;
    if (bound_o) *bound_o = bound;
    free(y);
}

/* Gene loop of inferJACKS (infer.py:18-30) over a CSR index; OpenMP over genes when n_threads > 1. */
int oracle_em_batched(const double* test, const double* ctrl, int64_t n_rows, int32_t L, int32_t ctrl_lines,
                      int32_t rowmean_fill, const int64_t* offsets, const int32_t* rows, int64_t n_genes,
                      const double* fx1, const double* fx2, const OracleEmParams* p, double* x1, double* x2,
                      double* w1, double* w2, double* y, double* tau, int32_t* iters, double* bound,
                      int32_t n_threads) {
    (void)n_rows;
    if (ctrl_lines != L && ctrl_lines != 1) return 4;
    int64_t g;
#pragma omp parallel for schedule(dynamic, 16) num_threads(n_threads > 0 ? n_threads : 1)
    for (g = 0; g < n_genes; ++g) {
        int64_t o = offsets[g];
        int G = (int)(offsets[g + 1] - o);
        em_gene(test, ctrl, L, ctrl_lines, rowmean_fill, rows + o, G, fx1, fx2, p, x1 ? x1 + o : 0, x2 ? x2 + o : 0,
                w1 ? w1 + g * (int64_t)L : 0, w2 ? w2 + g * (int64_t)L : 0, y ? y + o * L : 0, tau ? tau + o * L : 0,
                iters ? iters + g : 0, bound ? bound + g : 0);
    }
    return 0;
}
