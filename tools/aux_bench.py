"""Developer probe: timings of the p-value and preprocessing paths at the BASELINE shape (host entry points)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jacks_b200 import _native


def main():
    ctx = _native.context(0)
    rng = np.random.default_rng(1)
    L = int(os.environ.get('AB_LINES', 400))
    n_genes = int(os.environ.get('AB_GENES', 18000))
    for n_pseudo in [int(x) for x in os.environ.get('AB_PSEUDO', '2000,100000').split(',')]:
        pseudo = rng.normal(0.0, 0.2, size=(n_pseudo, L))
        genes = rng.normal(-0.3, 0.7, size=(n_genes, L))
        _native.kde_pvalues_host(ctx, genes[:64], pseudo[:64])
        t0 = time.perf_counter()
        pv = _native.kde_pvalues_host(ctx, genes, pseudo)
        dt = time.perf_counter() - t0
        print('kde p-values  n_genes=%d L=%d n_pseudo=%d : %.3f s  (%.3g p-values/s, %.3g cdf evals/s)' %
              (n_genes, L, n_pseudo, dt, n_genes * L / dt, n_genes * L * n_pseudo / dt), flush=True)
    N, C = 72000, 802
    counts = rng.poisson(300.0, size=(N, C)).astype(np.float64)
    _native.normalize_host(ctx, counts[:2048], 'median', take_log=True, prior=32.0)
    t0 = time.perf_counter()
    lc = _native.normalize_host(ctx, counts, 'median', take_log=True, prior=32.0)
    t1 = time.perf_counter()
    Iscreens = [[2 * s, 2 * s + 1] for s in range(C // 2)]
    data = _native.condense_host(ctx, lc, Iscreens)
    t2 = time.perf_counter()
    print('preprocess N=%d C=%d: normalize %.3f s, condense (%d samples) %.3f s' % (N, C, t1 - t0, len(Iscreens), t2 - t1))


if __name__ == '__main__':
    main()
