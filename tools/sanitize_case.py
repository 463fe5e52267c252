"""Small end-to-end case for compute-sanitizer: every kernel family once (EM fast/generic/deferred, KDE fast+brute,
preprocess incl. subset fit and mode normalisation)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jacks_b200 import _native, synth, preprocess
ctx = _native.context(0)
counts = np.array([4, 4, 1, 5, 4, 9, 4, 3, 4, 12, 4, 4, 2, 8, 6, 7] * 4, dtype=np.int64)
for L in (37, 400, 5):
    gi, test, ctrl, _ = synth.make_condensed_screen(n_genes=len(counts), n_lines=L, seed=5, guide_counts=counts, n_ctrl_genes=2)
    test[3, 1, 0] = np.nan
    names, offs, rows = _native.as_csr(gi)
    r = _native.em_batched_host(ctx, offs, rows, test, ctrl)
    r = _native.em_batched_host(ctx, offs, rows, test, np.ascontiguousarray(ctrl[:, :1]), params=_native.default_params(apply_w_hp=1))
    print('em L=%d iters' % L, r['n_iters'][:8].tolist())
rng = np.random.default_rng(0)
pv = _native.kde_pvalues_host(ctx, rng.normal(-0.3, 0.7, (300, 3)), rng.normal(0, 0.2, (5000, 3)))
os.environ['JACKS_KDE_BRUTE'] = '1'
pb = _native.kde_pvalues_host(ctx, rng.normal(-0.3, 0.7, (300, 3)), rng.normal(0, 0.2, (5000, 3)), positive=True)
del os.environ['JACKS_KDE_BRUTE']
d = _native.kde_eval_host(ctx, rng.normal(0, 0.2, (50, 3)), rng.normal(0, 0.2, (50, 3)), 'pdf', leave_one_out=False)
cnt = rng.poisson(200.0, size=(3000, 6)).astype(np.float64)
for nt in ('median', 'zmad', 'mode'):
    lc = _native.normalize_host(ctx, cnt, nt, take_log=True, prior=32.0)
data = _native.condense_host(ctx, lc, [[0, 1], [2, 3, 4], [5]])
sd = preprocess.calc_posterior_sd(lc[:, :2], guideset_indexs=list(range(0, 3000, 3)))
print('ok', float(pv.mean()), float(pb.mean()), float(np.nanmean(data)), float(sd.mean()))
