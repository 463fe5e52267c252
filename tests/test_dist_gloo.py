"""N > 1 path on CPU: two gloo ranks shard the gene index, each computes its block with the C oracle standing in
for the GPU kernels, one all-gather reassembles -- the result must equal the unsharded run bit for bit."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from jacks_b200 import dist as jd, synth, _native
        from oracle import c_oracle
        counts = np.array([4, 1, 7, 4, 4, 2, 9, 4, 3, 4, 4, 5, 4], dtype=np.int64)
        gi, test, ctrl, _ = synth.make_condensed_screen(n_genes=len(counts), n_lines=11, seed=3, guide_counts=counts, n_ctrl_genes=2)
        names, offs, rows = _native.as_csr(gi)

        def compute(o, r, t, c, **kw):
            return c_oracle.em_batched(o, r, t, c, want_y_tau=False, **kw)
        got = jd.infer_sharded(offs, rows, test, ctrl, compute=compute)
        ref = compute(offs, rows, test, ctrl)
        ok = all(np.array_equal(got[k], ref[k]) for k in ('w1', 'w2', 'x1', 'x2', 'n_iters'))

        # per-line mode (paper-script batch driver): every output carries a line axis
        def compute_lines(o, r, t, c, **kw):
            cols = [c_oracle.em_batched(o, r, np.ascontiguousarray(t[:, [l], :]), np.ascontiguousarray(c[:, [l], :]),
                                        want_y_tau=False, **kw) for l in range(t.shape[1])]
            return dict(w1=np.concatenate([k['w1'] for k in cols], axis=1), w2=np.concatenate([k['w2'] for k in cols], axis=1),
                        x1=np.stack([k['x1'] for k in cols], axis=1), x2=np.stack([k['x2'] for k in cols], axis=1),
                        n_iters=np.stack([k['n_iters'] for k in cols], axis=1))
        got = jd.infer_sharded(offs, rows, test, ctrl, compute=compute_lines, per_line=True)
        ref = compute_lines(offs, rows, test, ctrl)
        ok = ok and all(got[k].shape == ref[k].shape and np.array_equal(got[k], ref[k]) for k in ('w1', 'w2', 'x1', 'x2', 'n_iters'))
        # p-values of every gene against a pseudo-gene null, gene-sharded (BASELINE config 5)
        from oracle import jacks_oracle as O
        rng = np.random.default_rng(9)
        wg, wp = rng.normal(-0.2, 0.6, size=(23, 3)), rng.normal(0.0, 0.2, size=(57, 3))
        wg[4, 1] = np.nan
        pv = jd.pvalues_sharded(wg, wp, positive=False, compute=lambda blk, ps, pos: O.kde_pvalues(blk, ps, positive=pos))
        ok = ok and np.array_equal(pv, O.kde_pvalues(wg, wp), equal_nan=True)
        q.put((rank, bool(ok), jd.shard_bounds(offs, world).tolist()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_sharded_equals_unsharded_bit_for_bit(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in out), out
    assert len(set(tuple(b) for _, _, b in out)) == 1


def test_shard_bounds_balance_and_edges():
    from jacks_b200.dist import local_index, shard_bounds
    offs = np.concatenate([[0], np.cumsum([4] * 10)])
    assert shard_bounds(offs, 1).tolist() == [0, 10]
    assert shard_bounds(offs, 2).tolist() == [0, 5, 10]
    assert shard_bounds(offs, 4).tolist() == [0, 3, 5, 8, 10]
    b = shard_bounds(offs, 16)                      # more ranks than genes: empty shards allowed, monotone bounds
    assert b[0] == 0 and b[-1] == 10 and (np.diff(b) >= 0).all()
    ragged = np.concatenate([[0], np.cumsum([100, 1, 1, 1, 1, 1, 1, 94])])
    b = shard_bounds(ragged, 2)
    assert b.tolist() == [0, 1, 8]                  # balanced by guide count, not gene count
    o, r = local_index(ragged, np.arange(200), 1, 8)
    assert o[0] == 0 and o[-1] == 100 and r[0] == 100
    assert shard_bounds(np.array([0]), 4).tolist() == [0, 0, 0, 0, 0]
