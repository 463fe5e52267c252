"""GPU parity for BASELINE.json configs 2 and 3 against fixtures produced by the UNMODIFIED reference
(tests/golden/make_golden_configs.py).

config 3  reference-efficacy mode (infer.py:22-24,77-79,91): inferJACKS with the published Avana efficacies as fixed_x on
          2,048 genes x 4 gRNAs x 64 lines -- w_only=True (API) and fixed_x alone (--reffile): E[w], E[w^2] within 1e-6,
          identical pass counts, x echoed bit for bit.
config 2  the whole command line (jacks_io.runJACKS) on the stand-in for the reference's full example screen: 90,000
          guides, 18,000 genes x 5, 53 replicate columns -> 16 lines + control, 500 control genes, 2,000 pseudo-genes
          (the reference's own draws).  Gene ranking and guide order bit-exact over all rows; values of every 8th row within
          2e-6 of the reference's printed digits; column sums of all rows within 1e-6."""
import os

import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


def test_config3_fixed_x_and_w_only_vs_reference():
    from jacks_b200 import infer, synth
    z = load_golden('config3_fixed_x.npz')
    gene_index, testdata, ctrldata, _ = synth.make_condensed_screen(n_genes=2048, n_lines=64, seed=synth.AVANA_SEED + 3, n_ctrl_genes=64)
    fx = {'X1': z['fx_X1'], 'X2': z['fx_X2']}
    res = infer.inferJACKS(gene_index, testdata, ctrldata, fixed_x=fx, w_only=True)
    glist = list(gene_index)
    assert all(res[g][0] is None and res[g][1] is None and res[g][2] is None and res[g][3] is None for g in glist)   # infer.py:28
    w1 = np.stack([res[g][4] for g in glist])
    w2 = np.stack([res[g][5] for g in glist])
    for got, ref, name in ((w1, z['w1'], 'w1'), (w2, z['w2'], 'w2')):
        err = np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-12))
        assert err < 1e-6, (name, err)
    # pass counts through the array-level call
    from jacks_b200 import _native
    names, offs, rows = _native.as_csr(gene_index)
    r = infer.infer_batched(offs, rows, testdata, ctrldata, fixed_x=fx, want=('w1', 'x1', 'x2', 'tau'))
    assert r['n_iters'].tolist() == z['iters'].tolist()
    # --reffile semantics: x echoed, tau returned
    sub = glist[::16]
    full = infer.inferJACKS({g: gene_index[g] for g in sub}, testdata, ctrldata, fixed_x=fx)
    for i, g in enumerate(sub):
        y, tau, x1, x2, gw1, gw2 = full[g]
        assert np.array_equal(x1, z['full_x1'][i]) and np.array_equal(x1, fx['X1'][gene_index[g]])
        assert np.max(np.abs(gw1 - z['full_w1'][i]) / np.maximum(np.abs(z['full_w1'][i]), 1e-12)) < 1e-6
        assert np.max(np.abs(tau - z['full_tau'][i]) / np.abs(z['full_tau'][i])) < 1e-6


def read_table(path, n_label):
    labels, vals = [], []
    with open(path) as f:
        f.readline()
        for line in f:
            fld = line.rstrip('\n').split('\t')
            labels.append('\t'.join(fld[:n_label]))
            vals.append([float(v) if v != '' else np.nan for v in fld[n_label:]])
    return labels, np.array(vals, dtype=np.float64).reshape(len(labels), -1)


def test_config2_full_example_standin_cli_vs_reference(tmp_path, monkeypatch):
    from jacks_b200 import jacks_io, synth
    z = load_golden('config2_cli.npz')
    guide_ids, gene_ids, cols, samples, counts = synth.make_example_count_screen()
    assert counts.shape == (90000, 53)
    genes = sorted(set(gene_ids))
    cfile, rfile, nfile = synth.write_count_screen(os.path.join(str(tmp_path), 'ex'), guide_ids, gene_ids, cols, samples, counts,
                                                   genes[100:600])
    offs, rows = z['pseudo_offsets'], z['pseudo_rows']

    def same_draws(gene_index, ctrl_geneset, n):
        assert n == len(offs) - 1
        return {'JACKS_PSEUDO_GENE_%d' % i: [int(r) for r in rows[offs[i]:offs[i + 1]]] for i in range(n)}
    monkeypatch.setattr(jacks_io, 'createPseudoNonessGenes', same_draws)
    prefix = os.path.join(str(tmp_path), 'out')
    jacks_io.runJACKS(cfile, rfile, cfile, common_ctrl_sample='CTRL', gene_hdr='Gene', outprefix=prefix, ctrl_genes=nfile, n_pseudo=2000)
    report = {}
    for suffix, n_label in [('_gene_JACKS_results', 1), ('_gene_std_JACKS_results', 1), ('_gene_pval_JACKS_results', 1),
                            ('_grna_JACKS_results', 1), ('_logfoldchange_means', 2), ('_logfoldchange_std', 2)]:
        labels, vals = read_table(prefix + suffix + '.txt', n_label)
        key = suffix.strip('_')
        assert len(labels) == int(z[key + '.n']), key
        if suffix.startswith('_gene'):
            order = np.array([int(g[1:]) for g in labels], dtype=np.int32)
            assert np.array_equal(order, z[key + '.order']), key + ': gene ranking differs'
        elif suffix == '_grna_JACKS_results':
            order = np.array([int(g.split('_')[1]) for g in labels], dtype=np.int32)
            assert np.array_equal(order, z[key + '.order']), 'guide order differs'
        ref = z[key + '.every8']
        got = vals[::8]
        assert got.shape == ref.shape and (np.isnan(got) == np.isnan(ref)).all(), key
        m = ~np.isnan(ref)
        # both sides are 7-digit prints ('%5e'): 2e-6 relative allows the last printed digit to differ at a rounding
        # boundary; p-values additionally inherit the 1e-13 differences of E[w] through the slope of the null density
        err = np.abs(got[m] - ref[m]) / np.maximum(np.abs(ref[m]), 1e-300)
        assert err.max() <= 2e-6, (key, float(err.max()))
        cs = np.nansum(vals, axis=0)
        assert np.max(np.abs(cs - z[key + '.colsum']) / np.maximum(np.abs(z[key + '.colsum']), 1e-9)) < 1e-6, key
        report[key] = (len(labels), float(err.max()), float((err == 0).mean()))
    for suffix in ('_pseudo_noness_gene_JACKS_results', '_pseudo_noness_gene_std_JACKS_results'):
        with open(prefix + suffix + '.txt', 'rb') as f:
            assert f.read() == z[suffix.strip('_') + '.bytes'].tobytes(), suffix
    print(report)
    res, lines, grnas = jacks_io.loadJacksFullResultsFromPickle(prefix + jacks_io.PICKLE_FILENAME)
    assert len(res) == 18000 and len(lines) == 16 and all(len(v) == 5 for v in grnas.values())
