"""GPU checks of the paper-script batch modes (jacks_b200/paper_modes.py): the bootstrap driver against
draws made by the unmodified reference's sample_jacks_screen.py loop (tests/golden/bootstrap.npz: identical random
selections, preprocessing and inference results), leave-one-out against the oracle, the single-line mode through
the script-level wrapper."""
import random

import numpy as np
import pytest

from conftest import load_golden, small_screen
from test_cli_gpu import write_inputs
from test_em_gpu import close

pytestmark = pytest.mark.gpu


def test_bootstrap_draws_match_reference(tmp_path):
    from jacks_b200 import jacks_io, paper_modes
    z = load_golden('bootstrap.npz')
    cfile, rfile, _ = write_inputs(str(tmp_path))
    sample_spec, ctrl_spec, _ = jacks_io.createSampleSpec(cfile, rfile, 'Replicate', 'Sample', 'CTRL', None)
    gene_spec = jacks_io.createGeneSpec(cfile, 'sgRNA', 'Gene', ignore_blank_genes=False)
    random.seed(int(z['seed']))
    details = []
    ws = paper_modes.sampleJacksScreen(sample_spec, gene_spec, ctrl_spec, str(z['test_line']), 2, -1, 2, 3,
                                       num_ctrl_reps=2, details=details)
    assert list(ws) == [str(g) for g in z['genes']]
    for bs in range(2):
        d = details[bs]
        assert np.array_equal(d['offsets'], z['offsets_%d' % bs]) and np.array_equal(d['rows'], z['rows_%d' % bs]), 'guide draws'
        assert d['cell_lines'] == [str(c) for c in z['cell_lines_%d' % bs]], 'replicate / line draws'
        ok, worst = close(d['data'], z['data_%d' % bs], rtol=1e-9)
        assert ok, ('preprocessed cube', worst)
        for k in ('w1', 'w2'):
            ok, worst = close(d[k], z['%s_%d' % (k, bs)])
            assert ok, (bs, k, worst)
    g0 = str(z['genes'][0])
    assert len(ws[g0]) == 2 and abs(ws[g0][0][0] - z['w1_0'][0]) <= 1e-6 * abs(z['w1_0'][0])
    out = str(tmp_path / 'boot.txt')
    paper_modes.writeSampledScreen(out, ws, 2)
    lines = open(out).read().splitlines()
    assert lines[0].split('\t') == ['Gene'] + ['WMean', 'WStd', 'Pval'] * 2 and len(lines) == len(ws) + 1


def test_leave_one_out_vs_oracle():
    from jacks_b200 import paper_modes
    from oracle import jacks_oracle as O
    gene_index, testdata, ctrldata, pre = small_screen()
    genes = list(gene_index)[:50]
    gi = {g: gene_index[g] for g in genes}
    got, lines = paper_modes.leaveOneOutReference(gi, testdata, ctrldata)
    L = testdata.shape[1]
    assert lines == list(range(L))
    N = testdata.shape[0]
    for held in range(L):
        keep = [l for l in range(L) if l != held]
        ref = O.em_genes(gi, testdata[:, keep, :], ctrldata[:, keep, :])
        fx = {'X1': np.zeros(N), 'X2': np.zeros(N)}
        for g in gi:
            fx['X1'][gi[g]] = ref[g][2]
            fx['X2'][gi[g]] = ref[g][3]
        one = O.em_genes(gi, testdata[:, [held], :], ctrldata[:, [held], :], fixed_x=fx, w_only=True)
        for g in gi:
            assert abs(got[g][4][held] - one[g][4][0]) <= 1e-6 * max(abs(one[g][4][0]), 1e-12), (g, held)
            assert abs(got[g][5][held] - one[g][5][0]) <= 1e-6 * max(abs(one[g][5][0]), 1e-12), (g, held)


def test_single_mode_wrapper_and_writer(tmp_path):
    """inferSingleJACKS feeds writeJacksWResults like the script's combined per-line results do."""
    from jacks_b200 import jacks_io, paper_modes
    gene_index, testdata, ctrldata, pre = small_screen()
    z = load_golden('lines_em.npz')
    res = paper_modes.inferSingleJACKS(gene_index, testdata, ctrldata)
    genes = list(gene_index)
    w1 = np.stack([res[g][4] for g in genes])
    ok, worst = close(w1, z['small_w1'])
    assert ok, worst
    assert all(res[g][0] is None for g in genes)
    names = ['L%d' % i for i in range(testdata.shape[1])]
    prefix = str(tmp_path / 'single')
    jacks_io.writeJacksWResults(prefix, res, names, write_types=['', '_std'])
    lines = open(prefix + '_gene_JACKS_results.txt').read().splitlines()
    assert lines[0].split('\t') == ['Gene'] + names and len(lines) == len(genes) + 1
