import sys, time, os
sys.path.insert(0, '/root/repo')
import numpy as np
import bench
from jacks_b200 import _native
w = bench.build_workload(seed_offset=0)
ctx = _native.context(0); ctx2 = _native.Context(0)
params = _native.default_params()
N, L = w['test'].shape[0], w['test'].shape[1]
pin_test = _native.PinnedArray(w['test'].shape); pin_test.array[...] = w['test']
pin_ctrl = _native.PinnedArray(w['ctrl'].shape); pin_ctrl.array[...] = w['ctrl']
r_offs, r_rows, p_offs, p_rows = w['offs'], w['rows'], w['poffs'], w['prows']
Tr = int(r_offs[-1]); nr = len(r_offs) - 1; npz = len(p_offs) - 1
out_r = {k: _native.PinnedArray(s, dt) for k, (s, dt) in dict(
    x1=((Tr,), np.float64), x2=((Tr,), np.float64), w1=((nr, L), np.float64), w2=((nr, L), np.float64),
    y=((Tr, L), np.float64), tau=((Tr, L), np.float64), n_iters=((nr,), np.int32)).items()}
out_p = {k: _native.PinnedArray((npz, L), np.float64) for k in ('w1', 'w2')}
def step(mode):
    t0 = time.perf_counter()
    if mode == 'two_ctx':
        a = _native.em_batched_host(ctx, r_offs, r_rows, pin_test.array, pin_ctrl.array, params=params, want=tuple(out_r), out={k: v.array for k, v in out_r.items()}, sync=False)
        t1 = time.perf_counter()
        b = _native.em_batched_host(ctx2, p_offs, p_rows, pin_test.array, pin_ctrl.array, params=params, want=('w1', 'w2'), out={k: v.array for k, v in out_p.items()}, sync=False, used_rows_only=True)
        t2 = time.perf_counter()
        ctx2.synchronize(); t3 = time.perf_counter(); ctx.synchronize(); t4 = time.perf_counter()
    elif mode == 'pseudo_first':
        b = _native.em_batched_host(ctx2, p_offs, p_rows, pin_test.array, pin_ctrl.array, params=params, want=('w1', 'w2'), out={k: v.array for k, v in out_p.items()}, sync=False, used_rows_only=True)
        t1 = time.perf_counter()
        a = _native.em_batched_host(ctx, r_offs, r_rows, pin_test.array, pin_ctrl.array, params=params, want=tuple(out_r), out={k: v.array for k, v in out_r.items()}, sync=False)
        t2 = time.perf_counter()
        ctx2.synchronize(); t3 = time.perf_counter(); ctx.synchronize(); t4 = time.perf_counter()
    else:
        a = _native.em_batched_host(ctx, r_offs, r_rows, pin_test.array, pin_ctrl.array, params=params, want=tuple(out_r), out={k: v.array for k, v in out_r.items()}, keep_resident=True, sync=False)
        t1 = time.perf_counter()
        b = _native.em_batched_host(ctx, p_offs, p_rows, None, None, params=params, n_rows=N, n_lines=L, ctrl_lines=L, want=('w1', 'w2'), out={k: v.array for k, v in out_p.items()}, sync=False)
        t2 = time.perf_counter()
        ctx.synchronize(); t3 = t4 = time.perf_counter()
    return [1e3 * (x - t0) for x in (t1, t2, t3, t4)]
for mode in ('one_ctx', 'pseudo_first', 'only_pseudo', ):
    if mode == 'only_pseudo':
        for _ in range(2):
            t0 = time.perf_counter()
            b = _native.em_batched_host(ctx2, p_offs, p_rows, pin_test.array, pin_ctrl.array, params=params, want=('w1', 'w2'), out={k: v.array for k, v in out_p.items()}, sync=False, used_rows_only=True)
            t1 = time.perf_counter(); ctx2.synchronize(); t2 = time.perf_counter()
            print('only_pseudo: enqueue %.1f ms, done %.1f ms' % (1e3 * (t1 - t0), 1e3 * (t2 - t0)))
        continue
    step(mode)
    for _ in range(3):
        print(mode, ['%.1f' % x for x in step(mode)])
