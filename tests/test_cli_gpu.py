"""End-to-end drop-in check: run_JACKS semantics on example-small through jacks_b200.jacks_io.runJACKS, compared
with the output FILES of the unmodified reference (tests/golden/small_cli.npz, which also holds the pseudo-gene
index sets the reference drew, so p-values are compared on identical permutations).

Gene / guide ORDER (ranking) must be identical.  Values are printed with 7 significant digits ('%5e'); a value
may differ in its last printed digit when the true number sits on a rounding boundary, so fields are compared
numerically at 2e-6 relative and the fraction of byte-identical lines is reported and required to be > 99 %."""
import io
import os
import sys

import numpy as np
import pytest

from conftest import REPO, load_golden

pytestmark = pytest.mark.gpu


def write_inputs(tmp):
    inp = load_golden('small_inputs.npz')
    counts = inp['counts']
    cfile = os.path.join(tmp, 'counts.tab')
    with io.open(cfile, 'w') as f:
        f.write(u'sgRNA\tGene\t%s\n' % '\t'.join(str(c) for c in inp['columns']))
        for i in range(counts.shape[0]):
            f.write(u'%s\t%s\t%s\n' % (inp['guides'][i], inp['genes'][i], '\t'.join(str(int(v)) for v in counts[i])))
    rfile = os.path.join(tmp, 'repmap.tab')
    with io.open(rfile, 'w') as f:
        f.write(u'Replicate\tSample\n')
        for r, s in zip(inp['rep_names'], inp['rep_samples']):
            f.write(u'%s\t%s\n' % (r, s))
    nfile = os.path.join(tmp, 'neg.txt')
    with io.open(nfile, 'w') as f:
        for g in inp['neg_genes']:
            f.write(u'%s\n' % g)
    return cfile, rfile, nfile


def compare_tables(got, ref, name, n_label_cols):
    got_lines, ref_lines = got.splitlines(), ref.splitlines()
    assert len(got_lines) == len(ref_lines), (name, len(got_lines), len(ref_lines))
    identical = 0
    for a, b in zip(got_lines, ref_lines):
        if a == b:
            identical += 1
            continue
        fa, fb = a.split('\t'), b.split('\t')
        assert fa[:n_label_cols] == fb[:n_label_cols], (name, 'row order / labels differ', a[:60], b[:60])
        assert len(fa) == len(fb), name
        for x, y in zip(fa[n_label_cols:], fb[n_label_cols:]):
            if x == y:
                continue
            xv, yv = float(x), float(y)
            assert abs(xv - yv) <= 2e-6 * max(abs(yv), 1e-300), (name, a[:80], b[:80])
    assert identical >= 0.99 * len(ref_lines), (name, identical, len(ref_lines))
    return identical, len(ref_lines)


def test_cli_example_small_matches_reference_files(tmp_path, monkeypatch):
    from jacks_b200 import jacks_io
    gold = load_golden('small_cli.npz')
    cfile, rfile, nfile = write_inputs(str(tmp_path))
    offs, rows = gold['pseudo_offsets'], gold['pseudo_rows']

    def same_draws(gene_index, ctrl_geneset, n):
        assert n == len(offs) - 1
        return {'JACKS_PSEUDO_GENE_%d' % i: [int(r) for r in rows[offs[i]:offs[i + 1]]] for i in range(n)}
    monkeypatch.setattr(jacks_io, 'createPseudoNonessGenes', same_draws)
    prefix = os.path.join(str(tmp_path), 'out', 'small')
    jacks_io.runJACKS(cfile, rfile, cfile, common_ctrl_sample='CTRL', gene_hdr='Gene', outprefix=prefix,
                      ctrl_genes=nfile, n_pseudo=200)
    report = {}
    for suffix, n_label in [('_gene_JACKS_results', 1), ('_gene_std_JACKS_results', 1), ('_gene_pval_JACKS_results', 1),
                            ('_grna_JACKS_results', 1), ('_logfoldchange_means', 2), ('_logfoldchange_std', 2),
                            ('_pseudo_noness_gene_JACKS_results', 1), ('_pseudo_noness_gene_std_JACKS_results', 1)]:
        with open(prefix + suffix + '.txt', 'rb') as f:
            got = f.read().decode()
        ref = gold['file' + suffix].tobytes().decode()
        report[suffix] = compare_tables(got, ref, suffix, n_label)
    print(report)
    res, lines, grnas = jacks_io.loadJacksFullResultsFromPickle(prefix + jacks_io.PICKLE_FILENAME)
    assert lines == ['HL60', 'MOLM', 'MV411', 'OCIAML2', 'OCIAML3'] and len(res) == 1579
    y, tau, x1, x2, w1, w2 = res['A1BG']
    assert y.shape == tau.shape == (len(grnas['A1BG']), 5) and w1.shape == (5,)
    assert '%5e' % w1[0] == '-2.887833e-01'          # SURVEY.md section 4 spot value


def test_createPseudoNonessGenes_consumes_random_like_the_reference():
    import random
    from conftest import small_screen
    from jacks_b200 import jacks_io
    z = load_golden('pvals.npz')
    gene_index, _, _, _ = small_screen()
    random.seed(int(z['seed']))
    pidx = jacks_io.createPseudoNonessGenes(gene_index, [str(g) for g in z['ctrl_sorted']], 50)
    offs, rows = z['pseudo_offsets'], z['pseudo_rows']
    for i, name in enumerate(pidx):
        assert name == 'JACKS_PSEUDO_GENE_%d' % i and pidx[name] == rows[offs[i]:offs[i + 1]].tolist()


def test_cli_reffile_mode(tmp_path):
    """--reffile: x fixed to the reference efficacies, grna file echoes them, no log-fold-change files."""
    from jacks_b200 import jacks_io
    em = load_golden('small_em.npz')
    pre = load_golden('small_preproc.npz')
    cfile, rfile, nfile = write_inputs(str(tmp_path))
    reffile = os.path.join(str(tmp_path), 'ref.txt')
    with io.open(reffile, 'w') as f:
        f.write(u'sgrna\tX1\tX2\n')
        for g, a, b in zip(pre['meta'][:, 0], em['fx_X1'], em['fx_X2']):
            f.write(u'%s\t%r\t%r\n' % (g, float(a), float(b)))
    prefix = os.path.join(str(tmp_path), 'ref')
    jacks_io.runJACKS(cfile, rfile, cfile, common_ctrl_sample='CTRL', gene_hdr='Gene', outprefix=prefix, reffile=reffile)
    assert not os.path.exists(prefix + '_logfoldchange_means.txt')
    res, lines, grnas = jacks_io.loadJacksFullResultsFromPickle(prefix + jacks_io.PICKLE_FILENAME)
    genes = [str(g) for g in pre['genes']]
    offs = pre['gene_offsets']
    for i in range(0, len(genes), 25):
        g = genes[i]
        assert np.allclose(res[g][4], em['fx_w1'][i], rtol=1e-6, atol=0)
        assert np.array_equal(res[g][2], em['fx_X1'][pre['gene_rows'][offs[i]:offs[i + 1]]])
