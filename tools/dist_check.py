"""Developer check: the gene-sharded drivers over NCCL (run under torchrun, one rank per GPU) give bit-identical results
to the unsharded calls on rank 0's GPU: joint EM, per-line EM and KDE p-values.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jacks_b200 import _native, dist as jd, infer, synth


def main():
    rank, local = int(os.environ['RANK']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    counts = np.r_[np.full(3000, 4), [1, 2, 3, 5, 6, 9, 12]].astype(np.int64)
    gi, test, ctrl, ctrl_genes = synth.make_condensed_screen(n_genes=len(counts), n_lines=96, seed=21, guide_counts=counts, n_ctrl_genes=50)
    names, offs, rows = _native.as_csr(gi)
    got = jd.infer_sharded(offs, rows, test, ctrl, device=local)
    ref = infer.infer_batched(offs, rows, test, ctrl, want=('x1', 'x2', 'w1', 'w2'), device=local)
    ok = all(np.array_equal(got[k], ref[k]) for k in ('w1', 'w2', 'x1', 'x2', 'n_iters'))
    gl = jd.infer_sharded(offs, rows, test, ctrl, device=local, per_line=True)
    rl = _native.em_lines_host(_native.context(local), offs, rows, test, ctrl, want=('x1', 'x2', 'w1', 'w2', 'n_iters'))
    ok_lines = all(np.array_equal(gl[k], rl[k]) for k in ('w1', 'w2', 'x1', 'x2', 'n_iters'))
    pseudo = np.random.default_rng(4).normal(0.0, 0.2, size=(5000, 96))
    pv = jd.pvalues_sharded(ref['w1'], pseudo, device=local)
    pr = _native.kde_pvalues_host(_native.context(local), ref['w1'], pseudo)
    ok_pv = np.array_equal(pv, pr, equal_nan=True)
    print('rank %d: joint %s, per-line %s, p-values %s' % (rank, ok, ok_lines, ok_pv), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if (ok and ok_lines and ok_pv) else 1)


if __name__ == '__main__':
    main()
