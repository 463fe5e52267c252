import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from jacks_b200 import _native, synth
for G in (150, 1000):
    counts = np.array([G] + [4] * 200, dtype=np.int64)
    gi, test, ctrl, _ = synth.make_condensed_screen(n_genes=len(counts), n_lines=400, seed=3, guide_counts=counts, n_ctrl_genes=2)
    names, offs, rows = _native.as_csr(gi)
    ctx = _native.context(0)
    _native.em_batched_host(ctx, offs, rows, test, ctrl, want=('w1', 'n_iters'))
    t0 = time.perf_counter()
    r = _native.em_batched_host(ctx, offs, rows, test, ctrl, want=('w1', 'n_iters'))
    t1 = time.perf_counter()
    r2 = _native.em_batched_host(ctx, offs[1:] - offs[1], rows[offs[1]:], test, ctrl, want=('w1', 'n_iters'))
    t2 = time.perf_counter()
    print('G=%d: with the big gene %.2f ms, without %.2f ms, big gene iterations %d' % (G, 1e3 * (t1 - t0), 1e3 * (t2 - t1), r['n_iters'][0]))

# per-line mode with the 1,000-guide gene, and a screen whose genes all hold a missing line (every gene deferred to the NaN-safe kernel)
t0 = time.perf_counter(); r = _native.em_lines_host(ctx, offs, rows, test, ctrl, want=('w1', 'n_iters')); t1 = time.perf_counter()
print('per-line mode with the G=1000 gene: %.2f ms' % (1e3 * (t1 - t0)))
gi, test, ctrl, _ = synth.make_condensed_screen(n_genes=18000, n_lines=400)
names, offs, rows = _native.as_csr(gi)
_native.em_batched_host(ctx, offs, rows, test, ctrl, want=('w1', 'n_iters'))
t0 = time.perf_counter(); _native.em_batched_host(ctx, offs, rows, test, ctrl, want=('w1', 'n_iters')); t1 = time.perf_counter()
test[:, 7, 0] = np.nan
t2 = time.perf_counter(); r = _native.em_batched_host(ctx, offs, rows, test, ctrl, want=('w1', 'n_iters')); t3 = time.perf_counter()
print('18,000 genes x 400 lines through the host call: %.1f ms clean, %.1f ms with one line missing everywhere (all genes on the NaN-safe kernel)' % (1e3 * (t1 - t0), 1e3 * (t3 - t2)))
