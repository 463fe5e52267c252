"""Developer probe: host-side timeline of jacks_run_screen_host at the BASELINE shape (JACKS_TRACE=1 prints stamps)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from jacks_b200 import _native

w = bench.build_workload()
ctx = _native.context(0)
L = w['test'].shape[1]
pin_test = _native.PinnedArray(w['test'].shape); pin_test.array[...] = w['test']
pin_ctrl = _native.PinnedArray(w['ctrl1'].shape); pin_ctrl.array[...] = w['ctrl1']
Tr = int(w['offs'][-1])
outs = {k: _native.PinnedArray(s, dt) for k, (s, dt) in dict(
    x1=((Tr,), np.float64), x2=((Tr,), np.float64), w1=((bench.N_GENES, L), np.float64), w2=((bench.N_GENES, L), np.float64),
    n_iters=((bench.N_GENES,), np.int32), pvals=((bench.N_GENES, L), np.float64)).items()}
for i in range(4):
    t0 = time.perf_counter()
    _native.run_screen_host(ctx, w['offs'], w['rows'], w['poffs'], w['prows'], pin_test.array, pin_ctrl.array,
                            want=tuple(outs), out={k: v.array for k, v in outs.items()})
    print('call %d: %.3f ms' % (i, 1e3 * (time.perf_counter() - t0)), flush=True)
