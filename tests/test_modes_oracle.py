"""The oracle's restatement of the paper-script batch modes (per-line inference) against fixtures produced by the
unmodified reference (tests/golden/make_golden_modes.py).  CPU only, bit-exact."""
import warnings

import numpy as np

from conftest import load_golden, small_screen
from oracle import jacks_oracle as O
from jacks_b200 import synth


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and np.array_equal(a, b, equal_nan=True)


def check(res, iters, genes, z, prefix, offs, keys):
    for i, g in enumerate(genes):
        a, b = int(offs[i]), int(offs[i + 1])
        y, tau, x1, x2, w1, w2 = res[g]
        got = dict(y=y, tau=tau, x1=x1, x2=x2)
        assert same(w1, z[prefix + 'w1'][i]) and same(w2, z[prefix + 'w2'][i]), g
        for k in keys:
            assert same(got[k], z[prefix + k][a:b]), (g, k)
        assert np.array_equal(iters[g], z[prefix + 'iters'][i]), g


def test_lines_example_small_sample_bit_exact():
    gene_index, testdata, ctrldata, pre = small_screen()
    z = load_golden('lines_em.npz')
    genes = list(gene_index)
    pick = list(range(0, len(genes), 9))               # 176 genes, every guide count class
    sub = {genes[i]: gene_index[genes[i]] for i in pick}
    offs = pre['gene_offsets']
    it = {}
    res = O.em_lines(sub, testdata, ctrldata, iters_out=it)
    for i in pick:
        g = genes[i]
        a, b = int(offs[i]), int(offs[i + 1])
        assert same(res[g][4], z['small_w1'][i]) and same(res[g][5], z['small_w2'][i]), g
        assert same(res[g][2], z['small_x1'][a:b]) and same(res[g][1], z['small_tau'][a:b]), g
        assert np.array_equal(it[g], z['small_iters'][i]), g
    e = load_golden('small_em.npz')
    it = {}
    res = O.em_lines(sub, testdata, ctrldata, fixed_x={'X1': e['fx_X1'], 'X2': e['fx_X2']}, iters_out=it)
    for i in pick:
        g = genes[i]
        assert same(res[g][4], z['smallfx_w1'][i]) and np.array_equal(it[g], z['smallfx_iters'][i]), g


def test_lines_avana_shape_bit_exact():
    z = load_golden('lines_em.npz')
    gi, test, ctrl, _ = synth.make_condensed_screen(n_genes=18000, n_lines=400)
    names = [str(n) for n in z['av_names']][:4]
    sub = {n: gi[n] for n in names}
    offs = np.cumsum([0] + [len(gi[n]) for n in [str(n) for n in z['av_names']]])
    it = {}
    res = O.em_lines(sub, test, ctrl, iters_out=it)
    check(res, it, names, z, 'av_', offs, ('x1', 'x2'))


def test_lines_missing_values_bit_exact():
    gene_index, _, _, pre = small_screen()
    z = load_golden('lines_em.npz')
    names = [str(g) for g in z['nan_names']]
    sub = {n: gene_index[n] for n in names}
    it = {}
    with warnings.catch_warnings(), np.errstate(all='ignore'):
        warnings.simplefilter('ignore')
        res = O.em_lines(sub, z['nan_test'], z['nan_ctrl'], iters_out=it)
    check(res, it, names, z, 'nan_', pre['gene_offsets'], ('y', 'tau', 'x1', 'x2'))
