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
