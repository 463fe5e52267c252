"""Host-side logic of the drop-in modules that needs no GPU: specification readers, CSR packing, pseudo-gene draws,
ranking, BH sets, writers and the command-line parser -- against the reference's behaviour (golden fixtures) and
its documented defaults."""
import io
import os
import random

import numpy as np
import pytest

from conftest import load_golden, small_screen
from jacks_b200 import _native, jacks_io


def test_as_csr_roundtrip_and_order():
    gi = {'b': [5, 2], 'a': [7], 'c': [1, 1, 3]}
    names, offs, rows = _native.as_csr(gi)
    assert names == ['b', 'a', 'c'] and offs.tolist() == [0, 2, 3, 6] and rows.tolist() == [5, 2, 7, 1, 1, 3]
    names, offs, rows = _native.as_csr({})
    assert names == [] and offs.tolist() == [0] and rows.size == 0


def test_parser_flags_and_defaults():
    ap = jacks_io.getJacksParser()
    a = ap.parse_args(['c.tab', 'r.tab', 'g.tab'])
    assert (a.rep_hdr, a.sample_hdr, a.common_ctrl_sample, a.ctrl_sample_hdr, a.sgrna_hdr, a.gene_hdr) == \
        ('Replicate', 'Sample', 'CONTROL', None, 'sgRNA', 'Gene')
    assert (a.ignore_blank_genes, a.outprefix, a.reffile, a.fdr, a.fdr_thresh_type, a.positive, a.apply_w_hp) == \
        (False, '', None, None, 'REGULAR', False, False)
    assert (a.norm_type, a.ctrl_genes, a.n_pseudo, a.count_prior) == ('median', None, 2000, 32)
    a = ap.parse_args(['c', 'r', 'g', '--fdr', '0.1', '--positive', '--n_pseudo', '5', '--count_prior', '8'])
    assert a.fdr == 0.1 and a.positive and a.n_pseudo == 5 and a.count_prior == 8


def test_spec_readers(tmp_path):
    rep = tmp_path / 'rep.csv'
    rep.write_text(u'# comment line\nReplicate,Sample,Ctrl\nr1,A,C\nr2,A,C\nc1,C,C\n')
    spec, ctrl, nrep = jacks_io.createSampleSpec('counts.tab', str(rep), 'Replicate', 'Sample', 'C')
    assert spec == {'counts.tab': [('A', 'r1'), ('A', 'r2'), ('C', 'c1')]} and ctrl == {'A': 'C', 'C': 'C'} and nrep == {'A': 2, 'C': 1}
    spec, ctrl, _ = jacks_io.createSampleSpec('counts.tab', str(rep), 'Replicate', 'Sample', 'ignored', ctrl_sample_hdr='Ctrl')
    assert ctrl == {'A': 'C', 'C': 'C'}
    with pytest.raises(Exception, match='Could not find control sample'):
        jacks_io.createSampleSpec('counts.tab', str(rep), 'Replicate', 'Sample', 'MISSING')
    with pytest.raises(Exception, match='Could not find column'):
        jacks_io.createSampleSpec('counts.tab', str(rep), 'Replicate', 'Sample', 'C', ctrl_sample_hdr='Nope')
    with pytest.raises(Exception, match='Could not find line with header'):
        jacks_io.createSampleSpec('counts.tab', str(rep), 'NoSuchHeader', 'Sample', 'C')
    bad = tmp_path / 'bad.tab'
    bad.write_text(u'Replicate\tSample\tCtrl\nr1\tA\tC\nr2\tA\tD\n')
    with pytest.raises(Exception, match='Different controls'):
        jacks_io.createSampleSpec('x', str(bad), 'Replicate', 'Sample', 'C', ctrl_sample_hdr='Ctrl')
    gm = tmp_path / 'map.tab'
    gm.write_text(u'sgRNA\tGene\ng1\tX\ng2\t\ng3\tY\n')
    assert jacks_io.createGeneSpec(str(gm), 'sgRNA', 'Gene', ignore_blank_genes=True) == {'g1': 'X', 'g3': 'Y'}
    assert jacks_io.createGeneSpec(str(gm), 'sgRNA', 'Gene', ignore_blank_genes=False) == {'g1': 'X', 'g2': '', 'g3': 'Y'}
    spec = {'g1': 'X', 'g3': 'Y'}
    cg = tmp_path / 'ctrl.txt'
    cg.write_text(u'X extra words\nZ\n')
    assert jacks_io.readControlGeneset(str(cg), spec) == {'X'}
    assert jacks_io.readControlGeneset('Y', spec) == {'Y'}
    with pytest.raises(Exception, match='Not a file or unrecognised control gene'):
        jacks_io.readControlGeneset('nope', spec)
    ref = tmp_path / 'ref.txt'
    ref.write_text(u'sgrna\tX1\tX2\ng1\t1.5\t2.5\n')
    assert jacks_io.loadSgrnaReference(str(ref))['g1']['X2'] == '2.5'
    with pytest.raises(Exception, match='has no sgrna reference'):
        cnt = tmp_path / 'cnt.tab'
        cnt.write_text(u'sgRNA\tGene\tr1\tc1\ng1\tX\t1\t2\ng3\tY\t3\t4\n')
        jacks_io.preprocess(str(cnt), str(rep), str(gm), common_ctrl_sample='C', reffile=str(ref))


def test_pseudo_genes_follow_the_reference_random_stream():
    z = load_golden('pvals.npz')
    gene_index, _, _, _ = small_screen()
    random.seed(int(z['seed']))
    pidx = jacks_io.createPseudoNonessGenes(gene_index, [str(g) for g in z['ctrl_sorted']], 50)
    offs, rows = z['pseudo_offsets'], z['pseudo_rows']
    assert list(pidx) == ['JACKS_PSEUDO_GENE_%d' % i for i in range(50)]
    for i, name in enumerate(pidx):
        assert pidx[name] == rows[offs[i]:offs[i + 1]].tolist()


def _interpreter_draws(gene_index, ctrl_geneset, n):
    """The reference's loop, call for call (jacks/jacks/jacks_io.py:256-264)."""
    out = {}
    ctrl_lst = [x for x in ctrl_geneset if x in gene_index]
    all_lst = [x for x in gene_index]
    for i in range(n):
        num_guides = len(gene_index[random.choice(all_lst)])
        out['JACKS_PSEUDO_GENE_%d' % i] = [random.choice(gene_index[random.choice(ctrl_lst)]) for j in range(num_guides)]
    return out


@pytest.mark.parametrize('seed', [0, 1, 12345])
def test_native_pseudo_draws_equal_the_interpreter_stream(seed):
    """jacks_pseudo_draw (C-ABI) against random.choice: ragged guide counts incl. one-guide genes (a one-bit draw with
    rejection), sequence lengths on both sides of a power of two, a stream that crosses several 624-word refills, and
    the generator left where the interpreter loop leaves it."""
    rng = np.random.RandomState(seed)
    gene_index, r = {}, 0
    for g in range(257 + seed % 7):
        G = int(rng.choice([1, 1, 2, 3, 4, 5, 8, 16, 17]))
        gene_index['gene%d' % g] = list(range(r, r + G))
        r += G
    ctrl = ['gene%d' % g for g in rng.choice(len(gene_index), 33, replace=False)] + ['not_in_the_index']
    random.seed(seed)
    random.random()                               # an odd stream position
    want = _interpreter_draws(gene_index, ctrl, 700)
    after_want = [random.random(), random.getrandbits(70), random.gauss(0, 1)]
    random.seed(seed)
    random.random()
    got = jacks_io.createPseudoNonessGenes(gene_index, ctrl, 700)
    after_got = [random.random(), random.getrandbits(70), random.gauss(0, 1)]
    assert list(got) == list(want) and got == want
    assert after_got == after_want
    names, offs, rows = _native.as_csr(got)
    assert int(offs[-1]) > 624 * 3


def test_native_pseudo_draws_retry_errors_and_fallback():
    gene_index = {'a': [0, 1, 2], 'b': list(range(3, 3 + 40)), 'c': [43]}
    # the first buffer guess (mean guide count) is too small when most picks land on the 40-guide gene: the call retries
    random.seed(3)
    want = _interpreter_draws(gene_index, ['b'], 2000)
    random.seed(3)
    assert jacks_io.createPseudoNonessGenes(gene_index, ['b'], 2000) == want
    st = np.array(random.getstate()[1], dtype=np.uint32)
    offs = np.array([0, 3, 43, 44], dtype=np.int64)
    with pytest.raises(_native.JacksError):          # too small a buffer: JACKS_E_SHAPE, state untouched, rows_needed set
        import ctypes
        need = ctypes.c_int64(0)
        before = st.copy()
        out_o = np.zeros(11, dtype=np.int64)
        out_r = np.zeros(2, dtype=np.int32)
        rows = np.arange(44, dtype=np.int32)
        cg = np.array([1], dtype=np.int64)
        rc = _native.lib().jacks_pseudo_draw(st.ctypes.data, offs.ctypes.data, rows.ctypes.data, 3, cg.ctypes.data, 1, 10,
                                             out_o.ctypes.data, out_r.ctypes.data, 2, ctypes.byref(need))
        assert rc == _native.JACKS_E_SHAPE and need.value > 2 and (st == before).all()
        _native.check(rc)
    # what random.choice raises in the reference: no control gene in the index / no gene at all
    with pytest.raises(IndexError):
        jacks_io.createPseudoNonessGenes(gene_index, ['zzz'], 5)
    with pytest.raises(IndexError):
        jacks_io.createPseudoNonessGenes({}, [], 5)
    assert jacks_io.createPseudoNonessGenes(gene_index, ['a'], 0) == {}
    # entries that are not plain integers keep the interpreter loop (same stream either way)
    odd = {'a': ['r0', 'r1'], 'b': ['r2']}
    random.seed(9)
    want = _interpreter_draws(odd, ['a'], 20)
    random.seed(9)
    assert jacks_io.createPseudoNonessGenes(odd, ['a'], 20) == want


def test_ranking_and_fdr_sets_match_reference():
    z = load_golden('pvals.npz')
    pv = {str(g): z['pv_neg'][i] for i, g in enumerate(z['order'])}
    sets = jacks_io.getFDRGeneSets(pv, 0.2)
    assert ['|'.join(sorted(s)) for s in sets] == [str(s) for s in z['fdr_sets']]
    res = {'b': (None,) * 4 + (np.array([0.1, np.nan]), None), 'a': (None,) * 4 + (np.array([0.1, 0.1]), None),
           'c': (None,) * 4 + (np.array([-1.0, 0.0]), None)}
    assert [g for _, g in jacks_io.getSortedGenes(res, False)] == ['c', 'a', 'b']       # ties break on the name
    assert [g for _, g in jacks_io.getSortedGenes(res, True)] == ['b', 'a', 'c']


def test_writers_format_and_filtering(tmp_path):
    res = {'G2': (np.zeros((1, 2)), np.ones((1, 2)), np.array([1.0]), np.array([1.25]), np.array([0.5, -0.25]), np.array([0.5, 0.25])),
           'G1': (np.zeros((2, 2)), np.ones((2, 2)), np.array([0.9, 1.1]), np.array([1.0, 1.3]), np.array([-1.5, -2.0]), np.array([2.5, 4.25])),
           'JACKS_PSEUDO_GENE_0': (None, None, None, None, np.array([0.0, 0.0]), np.array([1.0, 1.0]))}
    prefix = str(tmp_path / 'o')
    jacks_io.writeJacksWResults(prefix, res, ['L1', 'L2'], write_types=['', '_std'])
    assert open(prefix + '_gene_JACKS_results.txt').read() == 'Gene\tL1\tL2\nG1\t-1.500000e+00\t-2.000000e+00\nG2\t5.000000e-01\t-2.500000e-01\n'
    assert open(prefix + '_gene_std_JACKS_results.txt').read() == 'Gene\tL1\tL2\nG1\t5.000000e-01\t5.000000e-01\nG2\t5.000000e-01\t4.330127e-01\n'
    real = {g: res[g] for g in ('G1', 'G2')}
    jacks_io.writeJacksXResults(prefix + '_x.txt', real, {'G1': ['a', 'b'], 'G2': ['c']})
    assert open(prefix + '_x.txt').read() == 'sgrna\tX1\tX2\na\t9.000000e-01\t1.000000e+00\nb\t1.100000e+00\t1.300000e+00\nc\t1.000000e+00\t1.250000e+00\n'
    jacks_io.pickleJacksFullResults(prefix + '.pickle', real, ['L1', 'L2'], {'G1': ['a', 'b'], 'G2': ['c']})
    r, lines, grnas = jacks_io.loadJacksFullResultsFromPickle(prefix + '.pickle')
    assert lines == ['L1', 'L2'] and grnas['G1'] == ['a', 'b'] and np.array_equal(r['G1'][4], res['G1'][4])
    with pytest.raises(Exception, match='Unrecognised write type'):
        jacks_io.writeJacksWResults(prefix, real, ['L1', 'L2'], write_types=['_bogus'])
    meta = np.array([['a', 'G1'], ['b', 'G1']])
    test = np.zeros((2, 1, 2)); ctrl = np.zeros((2, 1, 2))
    test[:, 0, 0] = [1.0, 2.0]; ctrl[:, 0, 0] = [0.5, 0.5]; test[:, 0, 1] = [3.0, 0.0]; ctrl[:, 0, 1] = [4.0, 1.0]
    jacks_io.writeFoldChanges(prefix + '_lfc.txt', test, ctrl, meta, ['S'])
    jacks_io.writeFoldChanges(prefix + '_lfcs.txt', test, ctrl, meta, ['S'], write_std=True)
    assert open(prefix + '_lfc.txt').read() == 'gRNA\tgene\tS\na\tG1\t5.000000e-01\nb\tG1\t1.500000e+00\n'
    assert open(prefix + '_lfcs.txt').read() == 'gRNA\tgene\tS\na\tG1\t5.000000e+00\nb\tG1\t1.000000e+00\n'


def test_bad_arguments_raise_before_touching_the_gpu():
    from jacks_b200 import preprocess
    with pytest.raises(Exception, match='Unrecognised normalisation type'):
        preprocess.normalizeLogCounts(np.zeros((40, 2)), normtype='bogus')
    assert np.isnan(preprocess.calc_posterior_sd(np.ones((5, 1)))).all()


def test_dropin_shim_exposes_reference_names():
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'dropin'))
    try:
        import jacks.infer as I
        import jacks.jacks_io as J
        import jacks.preprocess as P
        for name in ('inferJACKS', 'inferJACKSGene', 'LOG'):
            assert hasattr(I, name)
        for name in ('runJACKS', 'runJACKSFromArgs', 'preprocess', 'load_data_and_run', 'getJacksParser', 'createSampleSpec',
                     'createGeneSpec', 'readControlGeneset', 'loadSgrnaReference', 'createPseudoNonessGenes',
                     'computeW1PvalsAndFDRs', 'writeJacksWResults', 'writeJacksXResults', 'getSortedGenes', 'getGeneWs',
                     'loadJacksFullResultsFromPickle', 'REP_HDR_DEFAULT', 'SAMPLE_HDR_DEFAULT', 'SGRNA_HDR_DEFAULT',
                     'GENE_HDR_DEFAULT', 'PICKLE_FILENAME', 'GENE_FILENAME', 'GRNA_FILENAME'):
            assert hasattr(J, name), name
        for name in ('loadDataAndPreprocess', 'collateTestControlSamples', 'calc_posterior_sd', 'subsample_and_preprocess'):
            assert hasattr(P, name), name
    finally:
        sys.path.pop(0)
        for m in [k for k in sys.modules if k == 'jacks' or k.startswith('jacks.')]:
            del sys.modules[m]


def test_native_table_writer_formats_like_python(tmp_path):
    rng = np.random.default_rng(4)
    vals = np.concatenate([rng.normal(size=200) * 10.0 ** rng.integers(-300, 300, size=200),
                           [0.0, -0.0, np.nan, np.inf, -np.inf, 5e-324, 1.7976931348623157e308, 9.9999995e-5, 1e-259, 0.5]])
    vals = vals.reshape(-1, 5)
    labels = ['row%d\tg' % i for i in range(vals.shape[0])]
    blank = np.zeros(vals.shape, dtype=bool); blank[3, 2] = True
    path = str(tmp_path / 't.txt')
    for width, fmt in ((5, '%5e'), (6, '%6e')):
        _native.table_write(path, 'h1\th2\n', labels, vals, width, blank=blank)
        want = 'h1\th2\n' + ''.join('%s\t%s\n' % (labels[i], '\t'.join('' if blank[i, j] else fmt % vals[i, j]
                                                                      for j in range(vals.shape[1]))) for i in range(len(labels)))
        assert open(path).read() == want
    _native.table_write(path, 'only header\n', [], np.zeros((0, 3)), 5)
    assert open(path).read() == 'only header\n'


def test_native_table_reader_matches_csv_module(tmp_path):
    import csv
    text = ('sgRNA\tGene\ta\tb\tc\r\n'
            'g1\tX\t10\t2.5\t3\r\n'
            '\r\n'
            '"g 2"\tY\t 7 \t1e3\t-4\r\n'
            '"g""q"\tZ\t0\t12/4\t5\r\n'
            'g4\tX\t1\t2\t3')
    p = tmp_path / 'c.tab'
    p.write_bytes(text.encode())
    labels, values, bad = _native.table_read(str(p), '\t', [2, 4, 3])
    rows = list(csv.DictReader(io.open(str(p)), delimiter='\t'))
    assert [r['sgRNA'] for r in rows][:2] == labels[:2] == ['g1', 'g 2']
    assert labels[2] == 'g""q' and labels[3] == 'g4'            # raw bytes of a field with an escaped quote
    assert values.shape == (4, 3)
    assert values[0].tolist() == [10.0, 3.0, 2.5] and values[1].tolist() == [7.0, -4.0, 1000.0]
    assert bad[2].tolist() == [False, False, True] and np.isnan(values[2, 2])       # "12/4" is left to eval()
    assert values[3].tolist() == [1.0, 3.0, 2.0] and not bad[[0, 1, 3]].any()
    pc = tmp_path / 'c.csv'
    pc.write_text(u'id,v\nq1,5\n"a,b",6\n')
    labels, values, bad = _native.table_read(str(pc), ',', [1])
    assert labels == ['q1', 'a,b'] and values[:, 0].tolist() == [5.0, 6.0]
    pairs, _, _ = _native.table_read(str(p), '\t', [], str_cols=(0, 1))
    assert pairs[0] == ('g1', 'X') and pairs[3] == ('g4', 'X')


def test_native_table_reader_number_forms_and_short_records(tmp_path):
    """The digit fast path, the strtod path and the early stop at the last requested column: integers of every length a
    double holds exactly and beyond, signs, floats, exponents, padded cells, empty cells, records shorter than the selected
    columns, quoted fields beyond the last requested column, and records with both line terminators."""
    lines = ['id\tgene\ta\tb\tc\td',
             'r0\tG\t0\t007\t123456789012345\t1234567890123456789',       # 15 digits: fast path; 19 digits: strtod
             'r1\tG\t+5\t-3\t2.50\t1e-3',
             'r2\tG\t\t 4\t5 \t.5',                                         # an empty cell, padded cells
             'r3\tG\t1\t2',                                                  # short record: c and d are missing
             'r4\tG\t9\t8\t7\t"6,\t""x"""',                                  # a quoted field with a delimiter, last column
             'r5\tG\t0x10\t1_0\tnan\tinf']                                   # not plain decimal literals: left to eval()
    p = tmp_path / 't.tab'
    p.write_bytes(('\r\n'.join(lines[:4]) + '\n' + '\n'.join(lines[4:]) + '\n').encode())
    labels, v, bad = _native.table_read(str(p), '\t', [2, 3, 4, 5])
    assert labels == ['r0', 'r1', 'r2', 'r3', 'r4', 'r5']
    assert v[0].tolist() == [0.0, 7.0, 123456789012345.0, float(1234567890123456789)] and not bad[0].any()
    assert v[1].tolist() == [5.0, -3.0, 2.5, 1e-3] and not bad[1].any()
    assert bad[2].tolist() == [True, False, False, False] and v[2, 1:].tolist() == [4.0, 5.0, 0.5]
    assert bad[3].tolist() == [False, False, True, True] and v[3, :2].tolist() == [1.0, 2.0]
    assert bad[4].tolist() == [False, False, False, True] and v[4, :3].tolist() == [9.0, 8.0, 7.0]
    assert bad[5].all()
    # only the first columns asked for: the rest of every record (incl. the quoted field) is never looked at
    labels, v, bad = _native.table_read(str(p), '\t', [2], str_cols=(0, 1))
    assert labels[4] == ('r4', 'G') and v[:, 0].tolist()[:2] == [0.0, 5.0] and bad[:, 0].tolist() == [False, False, True, False, False, True]
    pairs, v0, _ = _native.table_read(str(p), '\t', [], str_cols=(0, 1))
    assert pairs == [('r%d' % i, 'G') for i in range(6)] and v0.shape == (6, 0)


def test_count_reader_native_equals_python_reader(tmp_path):
    from jacks_b200 import preprocess
    inp = load_golden('small_inputs.npz')
    counts = inp['counts'][:400]
    p = tmp_path / 'counts.tab'
    with io.open(str(p), 'w', newline='') as f:
        f.write(u'sgRNA\tGene\t%s\r\n' % '\t'.join(str(c) for c in inp['columns']))
        for i in range(counts.shape[0]):
            f.write(u'%s\t%s\t%s\r\n' % (inp['guides'][i], inp['genes'][i], '\t'.join(str(int(v)) for v in counts[i])))
    gene_spec = {str(g): str(n) for g, n in zip(inp['guides'][:400:2], inp['genes'][:400:2])}      # every other guide mapped
    cols = [str(c) for c in inp['columns'][[3, 0, 7]]]
    a_raw, a_meta = preprocess._read_counts_native(str(p), '\t', cols, gene_spec)
    b_raw, b_meta = preprocess._read_counts_python(str(p), '\t', cols, gene_spec)
    assert a_meta == b_meta and len(a_meta) == 200
    assert np.array_equal(np.asarray(a_raw), np.asarray(b_raw, dtype=np.float64))


def test_vectorised_gene_ranking_is_bit_identical():
    rng = np.random.default_rng(8)
    W = rng.normal(size=(500, 37))
    W[5, 3] = np.nan
    W[9] = W[8]                                    # exact tie: falls back to the name
    res = {'g%03d' % i: (None, None, None, None, W[i], None) for i in range(500)}
    fast = jacks_io.getSortedGenes(res, False)
    slow = sorted((np.nanmean(res[g][4]), g) for g in res)
    assert [g for _, g in fast] == [g for _, g in slow]
    assert all(a == b for (a, _), (b, _) in zip(fast, slow))


def test_paper_mode_spec_filters_and_combine():
    """Host helpers of the paper scripts (run_JACKS_single.py:8-31, leaveoneout_reference_test.py:10-18)."""
    from jacks_b200 import paper_modes as P
    sample_spec = {'f.tab': [('CTRL', 'c1'), ('CTRL', 'c2'), ('A', 'a1'), ('B', 'b1'), ('A', 'a2')]}
    ctrl_spec = {'CTRL': 'CTRL', 'A': 'CTRL', 'B': 'CTRL'}
    assert P.filterSampleSpec(sample_spec, 'A', ctrl_spec) == {'f.tab': [('CTRL', 'c1'), ('CTRL', 'c2'), ('A', 'a1'), ('A', 'a2')]}
    assert P.filterCtrlSpec(ctrl_spec, 'B') == {'B': 'CTRL', 'CTRL': 'CTRL'}
    assert P.filterOutSampleSpec(sample_spec, 'A', ctrl_spec) == {'f.tab': [('CTRL', 'c1'), ('CTRL', 'c2'), ('B', 'b1')]}
    singles = [{'g': (None, None, None, None, np.array([0.1 * ts]), np.array([1.0 + ts]))} for ts in range(3)]
    y, tau, x1, x2, w1, w2 = P.combineSingleResults(singles)['g']
    assert y is None and [float(v[0]) for v in w1] == [0.0, 0.1, 0.2] and [float(v[0]) for v in w2] == [1.0, 2.0, 3.0]


# ---- round 2: host logic around the fused screen job and the sharded driver -----------------------------------------
def test_collate_returns_broadcast_control_for_one_control_sample():
    from jacks_b200 import preprocess as P, infer
    rng = np.random.default_rng(0)
    data = rng.normal(size=(7, 4, 2))
    sample_ids = ['A', 'B', 'CTRL', 'D']
    ctrl_spec = {'A': 'CTRL', 'B': 'CTRL', 'CTRL': 'CTRL', 'D': 'CTRL'}
    test, ctrl, idx = P.collateTestControlSamples(data, sample_ids, ctrl_spec)
    assert idx == [0, 1, 3] and test.shape == ctrl.shape == (7, 3, 2)
    assert np.array_equal(ctrl, data[:, [2, 2, 2], :])           # the reference's values (preprocess.py:125) ...
    assert ctrl.strides[1] == 0 and infer._one_control(ctrl)     # ... without the copies: goes to the GPU as [N, 1, 2]
    # matched controls stay materialised
    ctrl_spec = {'A': 'CTRL', 'B': 'D', 'CTRL': 'CTRL', 'D': 'D'}
    test, ctrl, idx = P.collateTestControlSamples(data, sample_ids, ctrl_spec)
    assert idx == [0, 1] and np.array_equal(ctrl, data[:, [2, 3], :]) and not infer._one_control(ctrl)


def test_infer_row_and_plane_conventions():
    from jacks_b200 import infer
    assert infer._rows([0, -1, 3, -5], 5).tolist() == [0, 4, 3, 0]      # numpy fancy indexing wraps negatives (infer.py:21)
    with pytest.raises(IndexError):
        infer._rows([5], 5)
    with pytest.raises(IndexError):
        infer._rows([-6], 5)
    cube = np.zeros((3, 2, 3))
    assert infer._planes(cube).shape == (3, 2, 2)                         # a raw third plane is ignored (infer.py:24)


def test_sharding_helpers():
    from jacks_b200 import dist as jd
    assert jd.line_bounds(400, 8).tolist() == [0, 50, 100, 150, 200, 250, 300, 350, 400]
    assert jd.line_bounds(5, 3).tolist() == [0, 1, 3, 5]
    used, rebased = jd.compact_rows([7, 3, 3, 9, 7], 12)
    assert used.tolist() == [3, 7, 9] and rebased.tolist() == [1, 0, 0, 2, 1]
    # with one rank the sharded job degenerates to the plain calls (no process group needed)
    calls = []

    def em_fn(o, r, t, c, fx1, fx2):
        calls.append((len(o) - 1, t.shape[0]))
        n, L = len(o) - 1, t.shape[1]
        return dict(w1=np.arange(n * L, dtype=np.float64).reshape(n, L), w2=np.ones((n, L)), x1=np.ones(int(o[-1])),
                    x2=np.ones(int(o[-1])), n_iters=np.full(n, 3, dtype=np.int32))

    def kde_fn(g, p, positive):
        return g * 0.0 + p.shape[0]
    offs, rows = np.array([0, 2, 3]), np.array([4, 5, 9])
    poffs, prows = np.array([0, 1, 3]), np.array([9, 9, 4])
    test = np.zeros((12, 6, 2)); ctrl = np.zeros((12, 1, 2))
    res = jd.run_screen_sharded(offs, rows, poffs, prows, test, ctrl, engine=jd.HostEngine(em_fn, kde_fn))
    assert sorted(calls) == [(2, 3), (2, 3)]                # both blocks ran on the 3 rows they reference, not on 12
    assert res['w1'].shape == (2, 6) and res['pvals'].shape == (2, 6) and (res['pvals'] == 2).all()


def test_reference_recipe_imports_the_unmodified_reference():
    import os
    from oracle import make_ref
    if make_ref.make() is None:
        pytest.skip('no /root/reference and no oracle/_ref copy on this box')
    RI, RIO = make_ref.import_reference()
    assert os.path.abspath(RI.__file__).startswith(os.path.join(os.path.abspath(os.path.dirname(make_ref.__file__)), '_ref'))
    data = np.array([[0.1, -0.4, 0.3], [0.2, -0.5, 0.1]])
    err = np.full((2, 3), 0.2)
    y, tau, x1, x2, w1, w2 = RI.inferJACKSGene(data, err.copy(), np.zeros((2, 3)), err.copy(), 5)
    from oracle import jacks_oracle as O
    o = O.em_gene(data, err.copy(), np.zeros((2, 3)), err.copy(), n_iter=5)
    assert np.array_equal(w1, o[4]) and np.array_equal(x1, o[2])          # the restatement is bit-exact against it


def test_cubes_with_a_raw_plane_and_negative_rows_follow_numpy_semantics():
    """The reference reads planes 0 and 1 of [N, L, >=2] cubes (infer.py:57-58) and indexes rows with numpy rules."""
    from jacks_b200 import _native
    cube = np.arange(2 * 3 * 3, dtype=np.float64).reshape(2, 3, 3)
    c2 = _native._cube(cube)
    assert c2.shape == (2, 3, 2) and c2.flags['C_CONTIGUOUS'] and np.array_equal(c2, cube[:, :, :2])
    assert _native._cube(cube[:, :, :2]).shape == (2, 3, 2)
    rows = _native._rows([0, -1, 3, -5], 5)
    assert rows.dtype == np.int32 and rows.tolist() == [0, 4, 3, 0]
    with pytest.raises(IndexError):
        _native._rows([5], 5)
    with pytest.raises(IndexError):
        _native._rows([-6], 5)
    assert _native._rows([], 5).size == 0


def test_two_part_layout_covers_both_indexes():
    """dist.two_part_layout: [control block padded to world*k | my real rows], indexes re-based onto it."""
    from jacks_b200 import dist as jd
    real = np.array([10, 11, 3, 3, 12, 7])
    pseudo_all = np.array([7, 2, 9, 2, 7, 5, 9])
    pseudo_mine = np.array([9, 2, 5])
    P, R, k, ri, pi = jd.two_part_layout(real, pseudo_mine, pseudo_all, 16, 3)
    assert P.tolist() == [2, 5, 7, 9] and k == 2 and R.tolist() == [3, 10, 11, 12]
    cube_rows = np.concatenate([P, np.full(3 * k - len(P), -1), R])
    assert cube_rows[ri].tolist() == real.tolist() and cube_rows[pi].tolist() == pseudo_mine.tolist()
    P, R, k, ri, pi = jd.two_part_layout(real, [], [], 16, 2)
    assert len(P) == 0 and k == 0 and np.unique(real)[ri].tolist() == real.tolist()


def test_native_writer_prints_like_python_percent_e(tmp_path):
    """jacks_table_write formats with std::to_chars(scientific, 6); Python's '%5e' / '%6e' (the reference's writers,
    jacks_io.py:91-156) must come out byte for byte: the whole exponent range, subnormals, halfway cases at the seventh
    digit, exact powers of ten, signed zeros, infinities and NaN."""
    rng = np.random.RandomState(11)
    bits = rng.randint(0, 2 ** 63 - 1, size=60000, dtype=np.int64).view(np.float64)
    bits = bits[np.isfinite(bits)]
    ties = (rng.randint(-10 ** 7, 10 ** 7, size=20000) * 2 + 1) / 2.0 * 10.0 ** rng.randint(-12, 6, size=20000)
    small = 10.0 ** -rng.uniform(0, 320, size=20000) * rng.uniform(0.1, 10, size=20000)
    special = np.array([0.0, -0.0, 1.0, -1.0, 1e-5, 9.9999995e-1, 9.99999949999e7, 1e100, 1e-100, 5e-324, 2.2250738585072014e-308,
                        1.7976931348623157e308, np.inf, -np.inf, np.nan, 0.5, 1.25e-7, 123456.75, 1234567.5, 12345675.0])
    vals = np.concatenate([bits, -bits[:1000], ties, small, special])
    vals = vals[:len(vals) // 4 * 4].reshape(-1, 4)
    labels = ['row%d' % i for i in range(len(vals))]
    for width in (5, 6):
        path = str(tmp_path / ('w%d.txt' % width))
        _native.table_write(path, 'h\n', labels, vals, width)
        got = open(path).read().split('\n')
        assert got[0] == 'h' and got[-1] == ''
        fmt = '%' + str(width) + 'e'
        for i, line in enumerate(got[1:-1]):
            want = labels[i] + '\t' + '\t'.join(fmt % v for v in vals[i])
            assert line == want, (i, line, want)

