"""GPU parity of the per-line ("single JACKS") EM -- every (gene, line) pair an independent one-line EM, the batch
form of run_JACKS_single.py:83-85 -- against golden fixtures produced by the unmodified reference's own per-line
inferJACKS calls (tests/golden/make_golden_modes.py).  1e-6 relative on x, w, tau; IDENTICAL pass counts."""
import numpy as np
import pytest

from conftest import load_golden, small_screen as _small_screen
from test_em_gpu import close, report


def small_screen():
    gene_index, testdata, ctrldata, pre = _small_screen()
    return list(gene_index), pre['gene_offsets'], pre['gene_rows'], testdata, ctrldata

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def nat():
    from jacks_b200 import _native
    assert _native.lib().jacks_device_count() >= 1
    return _native


def compare(tag, r, z, prefix, keys, problems):
    for k in keys:
        ok, worst = close(r[k], z[prefix + k])
        if not ok:
            problems.append('%s %s: %s' % (tag, k, worst))
    if not np.array_equal(r['n_iters'], z[prefix + 'iters']):
        bad = np.argwhere(r['n_iters'] != z[prefix + 'iters'])
        problems.append('%s pass counts differ at %s' % (tag, bad[:5].tolist()))


def test_example_small_every_line_vs_reference(nat):
    """All 1,579 genes (G = 1..18: per-lane kernels and, above 12 guides, the generic kernel on one-line views)."""
    genes, offs, rows, testdata, ctrldata = small_screen()
    z = load_golden('lines_em.npz')
    r = nat.em_lines_host(nat.context(0), offs, rows, testdata, ctrldata)
    problems = []
    compare('joint', r, z, 'small_', ('w1', 'w2', 'x1', 'x2', 'tau'), problems)
    e = load_golden('small_em.npz')
    r = nat.em_lines_host(nat.context(0), offs, rows, testdata, ctrldata, fixed_x1=e['fx_X1'], fixed_x2=e['fx_X2'])
    compare('fixed_x', r, z, 'smallfx_', ('w1', 'w2'), problems)
    report(problems)


def test_avana_shape_lines_vs_reference(nat):
    from jacks_b200 import synth
    z = load_golden('lines_em.npz')
    gi, test, ctrl, _ = synth.make_condensed_screen(n_genes=18000, n_lines=400)
    names = [str(n) for n in z['av_names']]
    rows = np.concatenate([gi[n] for n in names]).astype(np.int32)
    assert np.array_equal(rows, z['av_rows'])
    offs = np.cumsum([0] + [len(gi[n]) for n in names]).astype(np.int64)
    r = nat.em_lines_host(nat.context(0), offs, rows, test, ctrl)
    problems = []
    compare('avana', r, z, 'av_', ('w1', 'w2', 'x1', 'x2'), problems)
    report(problems)


def test_missing_values_lines_vs_reference(nat):
    """NaN means / sds: the per-lane kernel hands those problems to the NaN-safe generic kernel."""
    genes, offs, rows, _, _ = small_screen()
    z = load_golden('lines_em.npz')
    n = len(z['nan_names'])
    assert [str(g) for g in z['nan_names']] == genes[:n]
    o = offs[:n + 1]
    r = nat.em_lines_host(nat.context(0), o, rows[:o[-1]], z['nan_test'], z['nan_ctrl'])
    problems = []
    compare('nan', r, z, 'nan_', ('w1', 'w2', 'x1', 'x2', 'tau', 'y'), problems)
    report(problems)


def test_python_api_matches_script_semantics(nat):
    """inferJACKSLines == [inferJACKS(..., [:, [ts], :]) for ts] column by column (through this package's own joint API),
    w_only leaves the first four entries None, an empty index returns {} and unmatched shapes raise."""
    from jacks_b200.infer import inferJACKS, inferJACKSLines
    genes, offs, rows, testdata, ctrldata = small_screen()
    gi = {g: rows[offs[i]:offs[i + 1]].tolist() for i, g in enumerate(genes[:60])}
    lines = inferJACKSLines(gi, testdata, ctrldata)
    for ts in range(testdata.shape[1]):
        one = inferJACKS(gi, testdata[:, [ts], :], ctrldata[:, [ts], :])
        for g in gi:
            for k in range(4):
                ok, worst = close(lines[g][k][:, ts], np.asarray(one[g][k]).reshape(len(gi[g])), rtol=1e-9)
                assert ok, (g, ts, k, worst)
            assert abs(lines[g][4][ts] - one[g][4][0]) <= 1e-9 * max(1.0, abs(one[g][4][0]))
    w = inferJACKSLines(gi, testdata, ctrldata, w_only=True)
    assert all(v[0] is None and v[3] is None and v[4].shape == (testdata.shape[1],) for v in w.values())
    assert inferJACKSLines({}, testdata, ctrldata) == {}
    with pytest.raises(Exception, match='Expecting matched size'):
        inferJACKSLines(gi, testdata, ctrldata[:, :2, :])


def test_lines_full_scale_properties(nat):
    """Avana shape, every gene and line (7.2 M one-line EMs): finite, pass counts within the cap, and the result of a
    line does not depend on the other lines (a permutation of the line axis permutes the outputs)."""
    from jacks_b200 import synth
    gi, test, ctrl, _ = synth.make_condensed_screen(n_genes=18000, n_lines=400)
    names, offs, rows = nat.as_csr(gi)
    r = nat.em_lines_host(nat.context(0), offs, rows, test, ctrl, want=('w1', 'w2', 'n_iters'))
    assert np.isfinite(r['w1']).all() and np.isfinite(r['w2']).all()
    assert r['n_iters'].min() >= 1 and r['n_iters'].max() <= 50
    assert (r['w2'] >= r['w1'] ** 2).all()
    perm = np.random.default_rng(3).permutation(test.shape[1])
    r2 = nat.em_lines_host(nat.context(0), offs, rows, np.ascontiguousarray(test[:, perm, :]),
                           np.ascontiguousarray(ctrl[:, perm, :]), want=('w1', 'n_iters'))
    assert np.array_equal(r2['w1'], r['w1'][:, perm]) and np.array_equal(r2['n_iters'], r['n_iters'][:, perm])


def test_lines_edge_shapes_vs_oracle(nat):
    """One line, one guide, nine and 150 guides (generic kernel on one-line views), a zero pass cap, and agreement
    with this library's JOINT run when the cube has a single line."""
    from oracle import jacks_oracle as O
    from jacks_b200 import synth
    counts = np.array([1, 2, 3, 4, 5, 6, 7, 8, 9, 150, 4, 1], dtype=np.int64)
    problems = []
    for L in (1, 3, 37):
        gi, test, ctrl, _ = synth.make_condensed_screen(n_genes=len(counts), n_lines=L, seed=11 + L, guide_counts=counts, n_ctrl_genes=2)
        names, offs, rows = nat.as_csr(gi)
        r = nat.em_lines_host(nat.context(0), offs, rows, test, ctrl)
        it = {}
        ref = O.em_lines(gi, test, ctrl, iters_out=it)
        for i, g in enumerate(names):
            a, b = int(offs[i]), int(offs[i + 1])
            for k, val in (('w1', ref[g][4]), ('w2', ref[g][5])):
                ok, worst = close(r[k][i], val)
                if not ok:
                    problems.append((L, g, k, worst))
            for k, val in (('x1', ref[g][2]), ('x2', ref[g][3]), ('tau', ref[g][1]), ('y', ref[g][0])):
                ok, worst = close(r[k][a:b], val)
                if not ok:
                    problems.append((L, g, k, worst))
            if not np.array_equal(r['n_iters'][i], it[g]):
                problems.append((L, g, 'iters', r['n_iters'][i].tolist(), it[g].tolist()))
        if L == 1:
            joint = nat.em_batched_host(nat.context(0), offs, rows, test, ctrl)
            for k in ('w1', 'w2'):
                ok, worst = close(r[k], joint[k], rtol=1e-9)
                if not ok:
                    problems.append(('joint-vs-lines', k, worst))
            if not np.array_equal(r['n_iters'][:, 0], joint['n_iters']):
                problems.append('joint-vs-lines iters')
        r0 = nat.em_lines_host(nat.context(0), offs, rows, test, ctrl, params=nat.default_params(n_iter=0), want=('w1', 'n_iters'))
        assert (r0['n_iters'] == 0).all()
        med = np.array([[np.median(test[gi[g], l, 0] - ctrl[gi[g], l, 0]) for l in range(L)] for g in names])
        ok, worst = close(r0['w1'], med, rtol=1e-12)
        if not ok:
            problems.append((L, 'n_iter=0 w1 is the initial median', worst))
    report(problems)
