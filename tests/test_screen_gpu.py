"""GPU parity of the fused screen job (C-ABI jacks_run_screen_host / jacks_run_screen_device): real-gene EM, pseudo-gene EM
and KDE p-values in one call must equal, bit for bit, the three separate calls it replaces
(jacks/jacks/jacks_io.py:425-441: inferJACKS, inferJACKS on the pseudo-genes, computeW1PvalsAndFDRs) -- and those are
checked against the oracle elsewhere (test_em_gpu.py, test_kde_gpu.py); here additionally against the C restatement of
the EM and the numpy restatement of the p-values on the same seeded screen."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def nat():
    from jacks_b200 import _native
    return _native


def make_screen(n_genes=3000, n_lines=96, n_pseudo=5000, seed=21, ragged=True):
    from jacks_b200 import _native, synth
    counts = np.full(n_genes, 4, dtype=np.int64)
    if ragged:
        counts[-7:] = [1, 2, 3, 5, 6, 9, 13]
    gi, test, ctrl, ctrl_genes = synth.make_condensed_screen(n_genes=n_genes, n_lines=n_lines, seed=seed, guide_counts=counts,
                                                             n_ctrl_genes=50)
    names, offs, rows = _native.as_csr(gi)
    pos = {g: i for i, g in enumerate(names)}
    poffs, prows = synth.make_pseudo_index_csr(offs, [pos[g] for g in ctrl_genes], n_pseudo, seed=seed + 1)
    return offs, rows.astype(np.int32), poffs, prows, test, ctrl


@pytest.mark.parametrize('common_ctrl', [False, True])
@pytest.mark.parametrize('positive', [False, True])
def test_screen_host_equals_separate_calls(nat, common_ctrl, positive):
    offs, rows, poffs, prows, test, ctrl = make_screen()
    if common_ctrl:
        ctrl = np.ascontiguousarray(ctrl[:, :1, :])          # one control column for every line: [N, 1, 2]
    ctx = nat.context(0)
    got = nat.run_screen_host(ctx, offs, rows, poffs, prows, test, ctrl, positive=positive,
                              want=('x1', 'x2', 'w1', 'w2', 'y', 'tau', 'n_iters', 'bound', 'pvals'),
                              want_pseudo=('w1', 'w2', 'n_iters', 'pvals'))
    real = nat.em_batched_host(ctx, offs, rows, test, ctrl)
    pseudo = nat.em_batched_host(ctx, poffs, prows, test, ctrl, want=('w1', 'w2', 'n_iters'))
    for k in ('x1', 'x2', 'w1', 'w2', 'y', 'tau', 'n_iters', 'bound'):
        assert np.array_equal(got[k], real[k], equal_nan=True), k
    for k in ('w1', 'w2', 'n_iters'):
        assert np.array_equal(got['pseudo_' + k], pseudo[k], equal_nan=True), k
    pv = nat.kde_pvalues_host(ctx, real['w1'], pseudo['w1'], positive=positive)
    assert np.array_equal(got['pvals'], pv, equal_nan=True)
    ppv = nat.kde_pvalues_host(ctx, pseudo['w1'], pseudo['w1'], positive=positive)
    assert np.array_equal(got['pseudo_pvals'], ppv, equal_nan=True)
    assert np.all((got['pvals'] >= 0) & (got['pvals'] <= 1))


def test_screen_host_vs_oracles(nat):
    from oracle import c_oracle, jacks_oracle as O
    offs, rows, poffs, prows, test, ctrl = make_screen(n_genes=400, n_lines=23, n_pseudo=700, seed=5)
    ctx = nat.context(0)
    got = nat.run_screen_host(ctx, offs, rows, poffs, prows, test, ctrl, want=('x1', 'x2', 'w1', 'w2', 'n_iters', 'pvals'),
                              want_pseudo=('w1', 'n_iters'))
    ref = c_oracle.em_batched(offs, rows, test, ctrl, want_y_tau=False)
    refp = c_oracle.em_batched(poffs, prows, test, ctrl, want_y_tau=False)
    assert got['n_iters'].tolist() == ref['n_iters'].tolist()
    assert got['pseudo_n_iters'].tolist() == refp['n_iters'].tolist()
    for k in ('x1', 'x2', 'w1', 'w2'):
        err = np.max(np.abs(got[k] - ref[k]) / np.maximum(np.abs(ref[k]), 1e-12))
        assert err < 1e-6, (k, err)
    # p-values of the reference restatement from the oracle's own E[w]: the EM's 1e-13 differences are amplified by the
    # slope of the null density, hence 1e-8 here (test_kde_gpu.py holds the kernel itself to 1e-10 on identical inputs)
    pv = O.kde_pvalues(ref['w1'], refp['w1'])
    m = pv > 1e-200
    assert np.max(np.abs(got['pvals'][m] - pv[m]) / pv[m]) < 1e-8


def test_screen_host_without_pseudo_genes_and_errors(nat):
    offs, rows, poffs, prows, test, ctrl = make_screen(n_genes=64, n_lines=9, n_pseudo=10, seed=2, ragged=False)
    ctx = nat.context(0)
    got = nat.run_screen_host(ctx, offs, rows, None, None, test, ctrl)          # no pseudo-genes: plain inferJACKS
    real = nat.em_batched_host(ctx, offs, rows, test, ctrl, want=('w1', 'x1'))
    assert 'pvals' not in got and np.array_equal(got['w1'], real['w1']) and np.array_equal(got['x1'], real['x1'])
    with pytest.raises(nat.JacksError):                                          # one pseudo-gene is no null distribution
        nat.run_screen_host(ctx, offs, rows, poffs[:2], prows[:poffs[1]], test, ctrl, want=('w1', 'pvals'))
    with pytest.raises(nat.JacksError, match='matched size'):
        nat.run_screen_host(ctx, offs, rows, poffs, prows, test, ctrl[:, :3, :])
    # fixed x for the real genes only (--reffile, jacks_io.py:425 vs :430)
    fx1, fx2 = np.full(test.shape[0], 0.9), np.full(test.shape[0], 0.9 ** 2 + 0.01)
    got = nat.run_screen_host(ctx, offs, rows, poffs, prows, test, ctrl, fixed_x1=fx1, fixed_x2=fx2, want=('x1', 'w1', 'pvals'),
                              want_pseudo=('w1',))
    real = nat.em_batched_host(ctx, offs, rows, test, ctrl, fixed_x1=fx1, fixed_x2=fx2, want=('w1', 'x1'))
    pseudo = nat.em_batched_host(ctx, poffs, prows, test, ctrl, want=('w1',))
    assert np.array_equal(got['w1'], real['w1']) and np.array_equal(got['x1'], real['x1'])
    assert np.array_equal(got['pseudo_w1'], pseudo['w1'])


def test_screen_device_equals_host(nat):
    torch = pytest.importorskip('torch')
    offs, rows, poffs, prows, test, ctrl = make_screen(n_genes=2500, n_lines=64, n_pseudo=3000, seed=8)
    ctx = nat.context(0)
    host = nat.run_screen_host(ctx, offs, rows, poffs, prows, test, ctrl, want=('x1', 'x2', 'w1', 'w2', 'n_iters', 'pvals'),
                               want_pseudo=('w1',))
    dev = torch.device('cuda', 0)
    N, L = test.shape[0], test.shape[1]
    d_test, d_ctrl = torch.from_numpy(test).to(dev), torch.from_numpy(ctrl).to(dev)
    n, npz, T = len(offs) - 1, len(poffs) - 1, int(offs[-1])
    f64 = dict(dtype=torch.float64, device=dev)
    o = dict(w1=torch.empty((n, L), **f64), w2=torch.empty((n, L), **f64), x1=torch.empty(T, **f64), x2=torch.empty(T, **f64),
             n_iters=torch.empty(n, dtype=torch.int32, device=dev), pvals=torch.empty((n, L), **f64),
             pseudo_w1=torch.empty((npz, L), **f64))
    pr, pp = nat.EmPlan(ctx, offs, rows, N, L), nat.EmPlan(ctx, poffs, prows, N, L)
    io = nat.ScreenDeviceIO()
    io.real.test, io.real.ctrl, io.real.ctrl_lines = d_test.data_ptr(), d_ctrl.data_ptr(), L
    for k in ('w1', 'w2', 'x1', 'x2', 'n_iters'):
        setattr(io.real, k, o[k].data_ptr())
    io.pvals, io.pseudo_w1 = o['pvals'].data_ptr(), o['pseudo_w1'].data_ptr()
    nat.run_screen_device(ctx, pr, pp, nat.default_params(), io, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    for k in ('w1', 'w2', 'x1', 'x2', 'n_iters', 'pvals', 'pseudo_w1'):
        assert np.array_equal(o[k].cpu().numpy(), host[k], equal_nan=True), k
    pr.close(); pp.close()


def test_back_to_back_async_compact_calls_do_not_race(nat):
    """Two asynchronous JACKS_UPLOAD_USED_ROWS calls on ONE context reuse the compact device cube: the second upload must
    wait for the first call's kernels (round-1 advisor finding)."""
    offs, rows, poffs, prows, test, ctrl = make_screen(n_genes=1500, n_lines=128, n_pseudo=6000, seed=12)
    poffs2, prows2 = poffs[:3001].copy(), prows[:poffs[3000]][::-1].copy()        # another index over other rows
    ctx = nat.Context(0)
    ref1 = nat.em_batched_host(ctx, poffs, prows, test, ctrl, want=('w1',))
    ref2 = nat.em_batched_host(ctx, poffs2, prows2, test, ctrl, want=('w1',))
    for _ in range(3):
        a = nat.em_batched_host(ctx, poffs, prows, test, ctrl, want=('w1',), sync=False, used_rows_only=True)
        b = nat.em_batched_host(ctx, poffs2, prows2, test, ctrl, want=('w1',), sync=False, used_rows_only=True)
        ctx.synchronize()
        assert np.array_equal(a['w1'], ref1['w1']) and np.array_equal(b['w1'], ref2['w1'])
    ctx.close()


@pytest.mark.parametrize('common_ctrl', [False, True])
def test_screen_host_pinned_cubes_nan_genes(nat, common_ctrl):
    """The call the bench times: PAGE-LOCKED cubes (the screen goes up first, the gene indexes and the control rows are
    pulled by kernels, the real-gene EM runs beside the one-launch pseudo-gene EM in the SM slots it leaves, p-values on
    their own download stream) -- with genes that hold NaN, so that the deferred-queue kernel runs in its co-residable
    configuration beside the resident pseudo-gene kernel.  Bit for bit against the separate calls on pageable arrays."""
    offs, rows, poffs, prows, test, ctrl = make_screen(n_genes=5000, n_lines=160, n_pseudo=20000, seed=33)
    rng = np.random.default_rng(3)
    for g in rng.choice(len(offs) - 1, size=40, replace=False):           # NaN measurements in 40 real genes
        r = rows[offs[g]]
        test[r, rng.integers(0, test.shape[1]), 0] = np.nan
    if common_ctrl:
        ctrl = np.ascontiguousarray(ctrl[:, :1, :])
    ctx = nat.context(0)
    pt, pc = nat.PinnedArray(test.shape), nat.PinnedArray(ctrl.shape)
    pt.array[...] = test
    pc.array[...] = ctrl
    want = ('x1', 'x2', 'w1', 'w2', 'n_iters', 'pvals')
    for _ in range(2):                                                    # twice: recycled plan buffers and staging
        got = nat.run_screen_host(ctx, offs, rows, poffs, prows, pt.array, pc.array, want=want)
    real = nat.em_batched_host(ctx, offs, rows, test, ctrl, want=('x1', 'x2', 'w1', 'w2', 'n_iters'))
    pseudo = nat.em_batched_host(ctx, poffs, prows, test, ctrl, want=('w1',))
    for k in ('x1', 'x2', 'w1', 'w2', 'n_iters'):
        assert np.array_equal(got[k], real[k], equal_nan=True), k
    pv = nat.kde_pvalues_host(ctx, real['w1'], pseudo['w1'])
    assert np.array_equal(got['pvals'], pv, equal_nan=True)
    pt.free(); pc.free()
