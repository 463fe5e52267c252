"""Developer probe: device-resident timing of the batched EM on the Avana shape for each kernel shape."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jacks_b200 import _native, synth


def main():
    n_genes = int(os.environ.get('QB_GENES', 18000))
    n_pseudo = int(os.environ.get('QB_PSEUDO', 100000))
    shapes = os.environ.get('QB_SHAPES', 'default;1;2;4').split(';')
    dev = torch.device('cuda:0')
    ctx = _native.context(0)
    print('fp64 peak TFLOP/s', ctx.fp64_peak_flops() / 1e12, flush=True)
    n_lines = int(os.environ.get('QB_LINES', 400))
    guides = int(os.environ.get('QB_GUIDES', 4))
    gi, test, ctrl, ctrl_genes = synth.make_condensed_screen(n_genes=n_genes, n_lines=n_lines, guides_per_gene=guides)
    names, offs, rows = _native.as_csr(gi)
    ctrl_ids = [names.index(g) for g in ctrl_genes]
    poffs, prows = synth.make_pseudo_index_csr(offs, ctrl_ids, n_pseudo)
    # one index: real genes followed by pseudo-genes
    aoffs = np.concatenate([offs, offs[-1] + poffs[1:]])
    arows = np.concatenate([rows.astype(np.int32), prows])
    N, L = test.shape[0], test.shape[1]
    d_test = torch.from_numpy(test).to(dev)
    if os.environ.get('QB_ONE_CTRL', '1') != '0':          # one control column for all lines, as the bench passes it
        ctrl = np.ascontiguousarray(ctrl[:, :1, :])
    d_ctrl = torch.from_numpy(ctrl).to(dev)
    n_all = len(aoffs) - 1
    T = int(aoffs[-1])
    w1 = torch.empty((n_all, L), dtype=torch.float64, device=dev)
    w2 = torch.empty((n_all, L), dtype=torch.float64, device=dev)
    x1 = torch.empty(T, dtype=torch.float64, device=dev)
    x2 = torch.empty(T, dtype=torch.float64, device=dev)
    iters = torch.empty(n_all, dtype=torch.int32, device=dev)
    params = _native.default_params()
    stream = torch.cuda.current_stream().cuda_stream
    for shape in shapes:
        if shape == 'default':
            os.environ.pop('JACKS_EM_W', None)
        else:
            os.environ['JACKS_EM_W'] = shape
        plan = _native.EmPlan(ctx, aoffs, arows, N, L)
        io = _native.EmDeviceIO()
        io.test, io.ctrl, io.ctrl_lines = d_test.data_ptr(), d_ctrl.data_ptr(), int(d_ctrl.shape[1])
        io.w1, io.w2, io.x1, io.x2, io.n_iters = w1.data_ptr(), w2.data_ptr(), x1.data_ptr(), x2.data_ptr(), iters.data_ptr()
        for _ in range(2):
            plan.run_device(params, io, stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = int(os.environ.get('QB_REPS', 5))
        e0.record()
        for _ in range(reps):
            plan.run_device(params, io, stream)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        it = iters.cpu().numpy()
        flops = (19.0 * it.astype(np.float64) + 12.0).sum() * guides * L
        print('shape %-8s  %8.3f ms  %10.0f genes/s  mean_iters %.2f  algorithmic %.2f TFLOP/s' %
              (shape, ms, n_all / ms * 1e3, it.mean(), flops / ms / 1e9), flush=True)
        plan.close()


if __name__ == '__main__':
    main()
