"""GPU parity of the batched EM (through the C-ABI) against the golden fixtures of the unmodified
reference and against the oracle on seeded inputs.  Tolerance: 1e-6 relative on x, w, tau (the
north-star tolerance; observed agreement is ~1e-12) and IDENTICAL EM iteration counts."""
import os

import numpy as np
import pytest

from conftest import edge_case_tags, load_golden, small_screen

pytestmark = pytest.mark.gpu

RTOL = 1e-6


def close(a, b, rtol=RTOL, atol=1e-300):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if a.shape != b.shape:
        return False, 'shape %s vs %s' % (a.shape, b.shape)
    nan_a, nan_b = np.isnan(a), np.isnan(b)
    if not (nan_a == nan_b).all():
        return False, 'NaN pattern differs at %s' % (np.argwhere(nan_a != nan_b)[:5].tolist(),)
    m = ~nan_a
    if not m.any():
        return True, 0.0
    err = np.abs(a[m] - b[m]) / np.maximum(np.abs(b[m]), 1e-12)
    worst = float(err.max())
    return worst <= rtol, worst


def report(problems):
    assert not problems, '\n'.join(str(p) for p in problems[:40])


@pytest.fixture(scope='module')
def nat():
    from jacks_b200 import _native
    assert _native.lib().jacks_device_count() >= 1
    return _native


def run(nat, offsets, rows, test, ctrl, **kw):
    hyper = {k: kw.pop(k) for k in list(kw) if k in ('n_iter', 'apply_w_hp', 'tol', 'mu0_x', 'var0_x', 'mu0_w', 'var0_w', 'tau_prior_strength')}
    if 'apply_w_hp' in hyper:
        hyper['apply_w_hp'] = int(bool(hyper['apply_w_hp']))
    if 'n_iter' in hyper:
        hyper['n_iter'] = int(hyper['n_iter'])
    return nat.em_batched_host(nat.context(0), offsets, rows, test, ctrl, params=nat.default_params(**hyper), **kw)


def test_edge_cases_vs_reference_golden(nat):
    from jacks_b200.infer import inferJACKSGene
    z = load_golden('edge_em.npz')
    problems = []
    for tag in edge_case_tags(z):
        kw = {k.split('.kw_')[1]: float(z[k]) for k in z.files if k.startswith(tag + '.kw_')}
        if 'apply_w_hp' in kw:
            kw['apply_w_hp'] = bool(kw['apply_w_hp'])
        if tag + '.fx1' in z.files:
            kw['fixed_x'] = {'X1': z[tag + '.fx1'], 'X2': z[tag + '.fx2']}
        d, de, c, ce = (z[tag + s].copy() for s in ('.data', '.data_err', '.ctrl', '.ctrl_err'))
        out = inferJACKSGene(d, de, c, ce, int(z[tag + '.n_iter']), **kw)
        for name, val in zip(('y', 'tau', 'x1', 'x2', 'w1', 'w2'), out):
            ok, worst = close(val, z[tag + '.' + name])
            if not ok:
                problems.append((tag, name, worst))
        # iteration count through the array-level entry point
        G, L = d.shape
        common = z[tag + '.ctrl'].shape != d.shape
        T = np.stack([z[tag + '.data'], z[tag + '.data_err']], axis=2)
        Cc = (np.stack([z[tag + '.ctrl'].reshape(G, 1), z[tag + '.ctrl_err'].reshape(G, 1)], axis=2) if common
              else np.stack([z[tag + '.ctrl'], z[tag + '.ctrl_err']], axis=2))
        fx = kw.pop('fixed_x', None)
        r = run(nat, [0, G], np.arange(G), T, Cc, n_iter=int(z[tag + '.n_iter']), ctrl_rowmean_fill=common,
                fixed_x1=None if fx is None else fx['X1'], fixed_x2=None if fx is None else fx['X2'], **kw)
        if int(r['n_iters'][0]) != int(z[tag + '.iters']):
            problems.append((tag, 'iters', int(r['n_iters'][0]), int(z[tag + '.iters'])))
    report(problems)


def test_example_small_vs_reference_golden(nat):
    gene_index, testdata, ctrldata, pre = small_screen()
    em = load_golden('small_em.npz')
    offs, rows = pre['gene_offsets'], pre['gene_rows']
    problems = []
    r = run(nat, offs, rows, testdata, ctrldata)
    if not (r['n_iters'] == em['iters']).all():
        problems.append(('joint iters differ at', np.flatnonzero(r['n_iters'] != em['iters'])[:10].tolist()))
    for k in ('x1', 'x2', 'w1', 'w2', 'tau'):
        ok, worst = close(r[k], em[k])
        if not ok:
            problems.append(('joint', k, worst))
    ok, worst = close(r['y'], (testdata[:, :, 0] - ctrldata[:, :, 0])[rows])
    if not ok:
        problems.append(('joint', 'y', worst))
    r = run(nat, offs, rows, testdata, ctrldata, apply_w_hp=True)
    if not (r['n_iters'] == em['hp_iters']).all():
        problems.append(('hp iters differ', np.flatnonzero(r['n_iters'] != em['hp_iters'])[:10].tolist()))
    for k in ('x1', 'w1', 'w2'):
        ok, worst = close(r[k], em['hp_' + k])
        if not ok:
            problems.append(('hp', k, worst))
    r = run(nat, offs, rows, testdata, ctrldata, fixed_x1=em['fx_X1'], fixed_x2=em['fx_X2'])
    if not (r['n_iters'] == em['fx_iters']).all():
        problems.append(('fixed_x iters differ', np.flatnonzero(r['n_iters'] != em['fx_iters'])[:10].tolist()))
    for k in ('w1', 'w2'):
        ok, worst = close(r[k], em['fx_' + k])
        if not ok:
            problems.append(('fixed_x', k, worst))
    ok, worst = close(r['x1'], em['fx_X1'][rows])
    if not ok:
        problems.append(('fixed_x echo', worst))
    # common control uploaded once ([N,1,2]) must give the same numbers as the replicated cube
    r1 = run(nat, offs, rows, testdata, np.ascontiguousarray(ctrldata[:, :1, :]))
    for k in ('x1', 'w1', 'w2', 'tau'):
        ok, worst = close(r1[k], em[k])
        if not ok:
            problems.append(('common-control', k, worst))
    report(problems)


def test_dropin_inferJACKS_example_small(nat):
    from jacks_b200.infer import inferJACKS
    gene_index, testdata, ctrldata, pre = small_screen()
    em = load_golden('small_em.npz')
    offs = pre['gene_offsets']
    res = inferJACKS(gene_index, testdata, ctrldata)
    assert list(res) == list(gene_index)
    problems = []
    for i, g in enumerate(gene_index):
        y, tau, x1, x2, w1, w2 = res[g]
        G = len(gene_index[g])
        assert y.shape == (G, 5) and tau.shape == (G, 5) and x1.shape == (G,) and w1.shape == (5,)
        for name, val, ref in (('x1', x1, em['x1'][offs[i]:offs[i + 1]]), ('w1', w1, em['w1'][i]), ('w2', w2, em['w2'][i]),
                               ('tau', tau, em['tau'][offs[i]:offs[i + 1]])):
            ok, worst = close(val, ref)
            if not ok:
                problems.append((g, name, worst))
    report(problems)
    res = inferJACKS({g: gene_index[g] for g in list(gene_index)[:7]}, testdata, ctrldata, w_only=True)
    assert all(v[:4] == (None, None, None, None) for v in res.values())
    with pytest.raises(Exception, match='matched size'):
        inferJACKS(gene_index, testdata, ctrldata[:, :3])
    assert inferJACKS({}, testdata, ctrldata[:, :3]) == {}


def test_avana_shape_vs_reference_golden(nat):
    from jacks_b200 import synth
    z = load_golden('avana_em.npz')
    gene_index, testdata, ctrldata, _ = synth.make_condensed_screen(n_genes=48, n_ctrl_genes=8)
    names, offs, rows = nat.as_csr(gene_index)
    problems = []
    for shape in (None, '1', '2', '4'):
        if shape is None:
            os.environ.pop('JACKS_EM_W', None)
        else:
            os.environ['JACKS_EM_W'] = shape
        try:
            r = run(nat, offs, rows, testdata, ctrldata)
            if r['n_iters'].tolist() != z['iters'].tolist():
                problems.append((shape, 'iters', r['n_iters'].tolist(), z['iters'].tolist()))
            for k in ('x1', 'x2', 'w1', 'w2'):
                ok, worst = close(r[k], z[k])
                if not ok:
                    problems.append((shape, k, worst))
            ok, worst = close(r['tau'][:16], z['tau_head'])
            if not ok:
                problems.append((shape, 'tau', worst))
            ok, worst = close(r['tau'].sum(axis=1), z['tau_rowsum'])
            if not ok:
                problems.append((shape, 'tau_rowsum', worst))
            r = run(nat, offs, rows, testdata, ctrldata, fixed_x1=z['fx_X1'], fixed_x2=z['fx_X2'], want=('w1', 'w2', 'n_iters'))
            if r['n_iters'].tolist() != z['fx_iters'].tolist():
                problems.append((shape, 'fixed_x iters'))
            for k in ('w1', 'w2'):
                ok, worst = close(r[k], z['fx_' + k])
                if not ok:
                    problems.append((shape, 'fixed_x', k, worst))
        finally:
            os.environ.pop('JACKS_EM_W', None)
    report(problems)


def _random_screen(rng, Gs, L, nan_frac=0.0, sd_nan_frac=0.0):
    Gs = np.asarray(Gs)
    N = int(Gs.sum())
    w = rng.normal(-0.4, 0.8, size=(len(Gs), L))
    x = rng.normal(1.0, 0.4, size=N)
    gene_of = np.repeat(np.arange(len(Gs)), Gs)
    sd_t = rng.uniform(0.1, 0.6, size=(N, L)); sd_c = rng.uniform(0.05, 0.3, size=(N, L))
    cm = rng.normal(0, 1, size=(N, L))
    tm = cm + x[:, None] * w[gene_of] + rng.normal(0, 1, size=(N, L)) * np.sqrt(sd_t ** 2 + sd_c ** 2)
    if nan_frac:
        tm[rng.random((N, L)) < nan_frac] = np.nan
    if sd_nan_frac:
        sd_t[rng.random((N, L)) < sd_nan_frac] = np.nan
        sd_c[rng.random((N, L)) < sd_nan_frac] = np.nan
    test = np.stack([tm, sd_t], axis=2); ctrl = np.stack([cm, sd_c], axis=2)
    offs = np.concatenate([[0], np.cumsum(Gs)]).astype(np.int64)
    perm = rng.permutation(N).astype(np.int32)          # rows in arbitrary order, like a pseudo-gene index
    return offs, perm, np.ascontiguousarray(test[np.argsort(perm)]), np.ascontiguousarray(ctrl[np.argsort(perm)])


@pytest.mark.parametrize('L', [1, 2, 5, 31, 32, 33, 64, 100, 128, 129, 257, 400, 416, 417, 448, 449, 512, 513, 700, 800, 1100])
def test_shape_sweep_vs_oracle(nat, L):
    """Every kernel path (fast shapes, generic, deferred NaN genes, scratch spill) against the C oracle,
    which tests/test_oracle_golden.py pins to the reference."""
    from oracle import c_oracle
    rng = np.random.default_rng(1000 + L)
    Gs = list(range(1, 13)) * 2 + [17, 40]
    offs, rows, test, ctrl = _random_screen(rng, Gs, L, sd_nan_frac=0.01)
    problems = []
    for label, kw in (('joint', {}), ('hp', dict(apply_w_hp=True)), ('cap', dict(n_iter=4)),
                      ('hyper', dict(mu0_x=0.9, var0_x=1.5, mu0_w=-0.05, var0_w=100.0, tau_prior_strength=0.6, tol=0.01))):
        ref = c_oracle.em_batched(offs, rows, test, ctrl, **kw)
        r = run(nat, offs, rows, test, ctrl, **kw)
        if r['n_iters'].tolist() != ref['n_iters'].tolist():
            problems.append((label, 'iters', np.flatnonzero(r['n_iters'] != ref['n_iters']).tolist()))
        for k in ('x1', 'x2', 'w1', 'w2', 'y', 'tau', 'bound'):
            ok, worst = close(r[k], ref[k])
            if not ok:
                problems.append((label, k, worst))
    # missing measurements: NaN in y -> deferred to the NaN-safe kernel
    offs, rows, test, ctrl = _random_screen(rng, Gs, L, nan_frac=0.03)
    ref = c_oracle.em_batched(offs, rows, test, ctrl)
    r = run(nat, offs, rows, test, ctrl)
    if r['n_iters'].tolist() != ref['n_iters'].tolist():
        problems.append(('nan', 'iters', np.flatnonzero(r['n_iters'] != ref['n_iters']).tolist()))
    for k in ('x1', 'x2', 'w1', 'w2', 'y', 'tau'):
        ok, worst = close(r[k], ref[k])
        if not ok:
            problems.append(('nan', k, worst))
    report(problems)


def test_huge_gene_spills_to_global_scratch(nat):
    from oracle import c_oracle
    rng = np.random.default_rng(77)
    offs, rows, test, ctrl = _random_screen(rng, [150, 3, 90], 300)
    ref = c_oracle.em_batched(offs, rows, test, ctrl)
    r = run(nat, offs, rows, test, ctrl)
    assert r['n_iters'].tolist() == ref['n_iters'].tolist()
    for k in ('x1', 'x2', 'w1', 'w2', 'tau'):
        ok, worst = close(r[k], ref[k])
        assert ok, (k, worst)


def test_full_avana_shape_vs_c_oracle(nat):
    """BASELINE.json size: 18,000 genes x 4 x 400 plus pseudo-genes drawn from 1,000 control genes."""
    from jacks_b200 import synth
    from oracle import c_oracle
    gene_index, testdata, ctrldata, ctrl_genes = synth.make_condensed_screen()
    names, offs, rows = nat.as_csr(gene_index)
    r = run(nat, offs, rows, testdata, ctrldata, want=('x1', 'x2', 'w1', 'w2', 'n_iters'), keep_resident=True)
    ref = c_oracle.em_batched(offs, rows, testdata, ctrldata, want_y_tau=False, n_threads=os.cpu_count() or 1)
    assert (r['n_iters'] == ref['n_iters']).all(), np.flatnonzero(r['n_iters'] != ref['n_iters'])[:10]
    for k in ('x1', 'x2', 'w1', 'w2'):
        ok, worst = close(r[k], ref[k])
        assert ok, (k, worst)
    ctrl_ids = [names.index(g) for g in ctrl_genes]
    poffs, prows = synth.make_pseudo_index_csr(offs, ctrl_ids, 20000)
    rp = nat.em_batched_host(nat.context(0), poffs, prows, None, None, n_rows=testdata.shape[0], n_lines=testdata.shape[1],
                             ctrl_lines=ctrldata.shape[1], want=('w1', 'n_iters'))
    refp = c_oracle.em_batched(poffs, prows, testdata, ctrldata, want_y_tau=False, n_threads=os.cpu_count() or 1)
    assert (rp['n_iters'] == refp['n_iters']).all()
    ok, worst = close(rp['w1'], refp['w1'])
    assert ok, worst


@pytest.mark.parametrize('mode', ['1', '2'])
def test_tensor_memory_variant_forced(mode):
    """The TMEM variants of the fast kernel (1: tau in tensor memory, 2: tau and tau_pr_den) are chosen by shape; force
    each for every eligible shape (JACKS_EM_TMEM, read once per process) and re-run the golden and shape-sweep parity
    tests in a child process."""
    import subprocess
    import sys
    env = dict(os.environ, JACKS_EM_TMEM=mode)
    r = subprocess.run([sys.executable, '-m', 'pytest', os.path.abspath(__file__), '-m', 'gpu', '-q', '-x', '-k',
                        'golden or shape_sweep or fixed_x'], env=env, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


def test_used_rows_only_upload_is_identical(nat):
    """JACKS_UPLOAD_USED_ROWS (compact cube of the rows an index references, re-based index) changes what is uploaded,
    not what is computed: bit-identical outputs for a pseudo-gene-like index over scattered rows, with and without
    fixed x, on a second context, synchronously and asynchronously."""
    rng = np.random.default_rng(5)
    Gs = [4] * 300 + [3, 5, 9, 1]
    offs, rows, test, ctrl = _random_screen(rng, Gs, 64, sd_nan_frac=0.01)
    N = test.shape[0]
    pool = rng.choice(N, size=90, replace=False)                   # "control guides"
    p_rows = rng.choice(pool, size=4 * 500).astype(np.int32)       # pseudo-genes draw them with replacement
    p_offs = (4 * np.arange(501)).astype(np.int64)
    ctx2 = nat.Context(0)
    fx1, fx2 = rng.uniform(0.5, 1.5, N), rng.uniform(1.0, 3.0, N)
    for kw in ({}, dict(fixed_x1=fx1, fixed_x2=fx2)):
        full = nat.em_batched_host(nat.context(0), p_offs, p_rows, test, ctrl, **kw)
        part = nat.em_batched_host(ctx2, p_offs, p_rows, test, ctrl, used_rows_only=True, **kw)
        for k in full:
            assert np.array_equal(full[k], part[k], equal_nan=True), k
    # pinned caller memory: the GPU gathers the rows itself (zero-copy reads), no host gather
    pt, pc = nat.PinnedArray(test.shape), nat.PinnedArray(ctrl.shape)
    pt.array[...] = test; pc.array[...] = ctrl
    full = nat.em_batched_host(nat.context(0), p_offs, p_rows, test, ctrl, fixed_x1=fx1, fixed_x2=fx2)
    part = nat.em_batched_host(ctx2, p_offs, p_rows, pt.array, pc.array, used_rows_only=True, fixed_x1=fx1, fixed_x2=fx2)
    for k in full:
        assert np.array_equal(full[k], part[k], equal_nan=True), ('pinned', k)
    # asynchronous, interleaved with a full call on the first context
    a = nat.em_batched_host(nat.context(0), offs, rows, test, ctrl, sync=False)
    b = nat.em_batched_host(ctx2, p_offs, p_rows, test, ctrl, used_rows_only=True, sync=False)
    ctx2.synchronize(); nat.context(0).synchronize()
    ref = nat.em_batched_host(nat.context(0), p_offs, p_rows, test, ctrl)
    for k in ref:
        assert np.array_equal(ref[k], b[k], equal_nan=True), k
    assert np.isfinite(a['w1']).all()
    empty = nat.em_batched_host(ctx2, np.zeros(1, np.int64), np.zeros(0, np.int32), test, ctrl, used_rows_only=True)
    assert empty['w1'].shape == (0, test.shape[1])
    with pytest.raises(IndexError):
        nat.em_batched_host(ctx2, np.array([0, 1], np.int64), np.array([N], np.int32), test, ctrl, used_rows_only=True)
    ctx2.close()


def test_big_gene_kernel_vs_oracle(nat):
    """Genes without a fast kernel of 4,096 elements and more run one per CTA (em_big_kernel): a 600-guide and a 1,000-guide gene among small
    ones, with missing means and sds, the hyper-prior, fixed x and an odd and an even number of valid values per line
    (both nanmedian branches of the bit-wise selection)."""
    from oracle import c_oracle
    rng = np.random.default_rng(2024)
    offs, rows, test, ctrl = _random_screen(rng, [600, 4, 1000, 7, 33], 41, nan_frac=0.002, sd_nan_frac=0.01)
    N = test.shape[0]
    problems = []
    fx1 = rng.uniform(0.5, 1.5, N)
    fx2 = fx1 ** 2 + rng.uniform(0.01, 0.5, N)             # second moments of a real posterior: E[x^2] >= E[x]^2
    for label, kw in (('joint', {}), ('hp', dict(apply_w_hp=True)), ('fixed', dict(fixed_x1=fx1, fixed_x2=fx2))):
        ref = c_oracle.em_batched(offs, rows, test, ctrl, **kw)
        r = run(nat, offs, rows, test, ctrl, **kw)
        if r['n_iters'].tolist() != ref['n_iters'].tolist():
            problems.append((label, 'iters', r['n_iters'].tolist(), ref['n_iters'].tolist()))
        for k in ('x1', 'x2', 'w1', 'w2', 'tau'):
            ok, worst = close(r[k], ref[k])
            if not ok:
                problems.append((label, k, worst))
    report(problems)
