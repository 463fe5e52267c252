#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by RUNNING THE UNMODIFIED REFERENCE.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    PYTHONHASHSEED=0 python tests/golden/make_golden.py

The reference (felicityallen/JACKS, /root/reference/jacks) is imported as-is; the only
shim is ``jacks.infer.SP = numpy`` because infer.py uses scipy's removed top-level numpy
aliases.  Nothing from the repo's own implementation or from
``oracle/`` is used to produce these files, so they pin both.

Outputs (all numpy .npz, float64 unless noted):
  small_inputs.npz    example-small raw counts (int32) + ids + column/sample names
  small_preproc.npz   reference loadDataAndPreprocess output on example-small
  small_em.npz        reference inferJACKS on example-small (joint, apply_w_hp, fixed_x) + iteration counts
  small_cli.npz       reference CLI output files on example-small (bytes) + the pseudo index it drew
  avana_em.npz        reference inferJACKS on a 48-gene sample of the synthetic Avana-shape screen (joint; fixed_x + w_only)
  edge_em.npz         reference inferJACKSGene on edge cases (G=1, L=1, NaNs, repeats, 1-D control, cap)
  sd_model.npz        reference calc_posterior_sd / normalizeLogCounts on small synthetic inputs
  pvals.npz           reference createPseudoNonessGenes (seeded) and computeW1PvalsAndFDRs
"""
import io
import os
import random
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference/jacks'
sys.path.insert(0, REF)
sys.path.insert(1, REPO)

import jacks.infer as RI                                   # noqa: E402  (the reference)
RI.SP = np
import logging                                             # noqa: E402
RI.LOG.setLevel(logging.WARNING)
import jacks.preprocess as RP                              # noqa: E402
import jacks.jacks_io as RIO                               # noqa: E402
assert RI.__file__.startswith(REF), RI.__file__

from jacks_b200 import synth                               # noqa: E402  (input generator only)

SMALL_COUNTS = REF + '/example-small/example_count_data.tab'
SMALL_REPMAP = REF + '/example-small/example_repmap.tab'
NEG_GENES = REF + '/example/NEGv1.txt'


class IterCounter:
    """Counts EM loop passes of the reference by counting lowerBound calls per gene."""

    def __init__(self):
        self.calls = 0
        self._orig = RI.lowerBound

    def __enter__(self):
        def counted(*a):
            self.calls += 1
            return self._orig(*a)
        RI.lowerBound = counted
        return self

    def __exit__(self, *a):
        RI.lowerBound = self._orig


def ref_gene(data, data_err, ctrl, ctrl_err, n_iter=50, **kw):
    with IterCounter() as c:
        out = RI.inferJACKSGene(np.array(data), np.array(data_err), np.array(ctrl), np.array(ctrl_err), n_iter, **kw)
    return out, c.calls - 1


def ref_genes(gene_index, testdata, ctrldata, **kw):
    """Reference inferJACKS gene by gene so that iteration counts can be recorded."""
    res, iters = {}, {}
    for gene in gene_index:
        with IterCounter() as c:
            r = RI.inferJACKS({gene: gene_index[gene]}, testdata.copy(), ctrldata.copy(), **kw)
        res[gene] = r[gene]
        iters[gene] = c.calls - 1
    return res, iters


def pack_results(gene_index, res, iters):
    genes = list(gene_index)
    return dict(
        x1=np.concatenate([res[g][2] for g in genes]), x2=np.concatenate([res[g][3] for g in genes]),
        w1=np.stack([res[g][4] for g in genes]), w2=np.stack([res[g][5] for g in genes]),
        tau=np.concatenate([res[g][1] for g in genes], axis=0),
        y=np.concatenate([res[g][0] for g in genes], axis=0),
        iters=np.array([iters[g] for g in genes], dtype=np.int32))


def save(name, **arrays):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrays)
    print('%-20s %8.1f KB' % (name, os.path.getsize(path) / 1024.0))


def small_inputs():
    import csv
    with io.open(SMALL_COUNTS) as f:
        rdr = csv.reader(f, delimiter='\t')
        hdr = next(rdr)
        rows = [r for r in rdr]
    counts = np.array([[int(v) for v in r[2:]] for r in rows], dtype=np.int32)
    with io.open(SMALL_REPMAP) as f:
        rep = [l.split() for l in f.read().splitlines()[1:] if l.strip()]
    save('small_inputs.npz', counts=counts, guides=np.array([r[0] for r in rows]),
         genes=np.array([r[1] for r in rows]), columns=np.array(hdr[2:]),
         rep_names=np.array([r[0] for r in rep]), rep_samples=np.array([r[1] for r in rep]),
         neg_genes=np.array([l.split()[0] for l in io.open(NEG_GENES) if l.strip()]))


def small_preproc_and_em():
    sample_spec, ctrl_spec, _ = RIO.createSampleSpec(SMALL_COUNTS, SMALL_REPMAP, 'Replicate', 'Sample', 'CTRL')
    gene_spec = RIO.createGeneSpec(SMALL_COUNTS, 'sgRNA', 'Gene', ignore_blank_genes=False)
    data, meta, sample_ids, genes, gene_index = RP.loadDataAndPreprocess(sample_spec, gene_spec, ctrl_spec=ctrl_spec)
    testdata, ctrldata, test_idx = RP.collateTestControlSamples(data, sample_ids, ctrl_spec)
    glist = list(gene_index)
    offsets = np.cumsum([0] + [len(gene_index[g]) for g in glist]).astype(np.int64)
    rows = np.concatenate([gene_index[g] for g in glist]).astype(np.int32)
    save('small_preproc.npz', data=data, meta=meta, sample_ids=np.array(sample_ids), genes=np.array(glist),
         gene_offsets=offsets, gene_rows=rows, test_idx=np.array(test_idx, dtype=np.int32))

    res, iters = ref_genes(gene_index, testdata, ctrldata)
    joint = pack_results(gene_index, res, iters)
    res_hp, iters_hp = ref_genes(gene_index, testdata, ctrldata, apply_w_hp=True)
    hp = pack_results(gene_index, res_hp, iters_hp)
    # fixed-x mode (the --reffile path, jacks_io.py:402-410): reuse the joint x as the reference efficacies
    N = testdata.shape[0]
    fx = {'X1': np.zeros(N), 'X2': np.zeros(N)}
    fx['X1'][rows] = joint['x1']; fx['X2'][rows] = joint['x2']
    res_fx, iters_fx = ref_genes(gene_index, testdata, ctrldata, fixed_x=fx)
    fxr = pack_results(gene_index, res_fx, iters_fx)
    save('small_em.npz',
         x1=joint['x1'], x2=joint['x2'], w1=joint['w1'], w2=joint['w2'], tau=joint['tau'], iters=joint['iters'],
         hp_x1=hp['x1'], hp_w1=hp['w1'], hp_w2=hp['w2'], hp_iters=hp['iters'],
         fx_X1=fx['X1'], fx_X2=fx['X2'], fx_w1=fxr['w1'], fx_w2=fxr['w2'], fx_iters=fxr['iters'])
    return gene_index, testdata, ctrldata, res


def small_cli():
    """The CLI end to end (run_JACKS.py semantics) with pseudo-genes; records the index sets drawn."""
    drawn = {}
    orig = RIO.createPseudoNonessGenes

    def recording(gene_index, ctrl_geneset, n):
        random.seed(0)
        out = orig(gene_index, sorted(ctrl_geneset), n)      # sorted: independent of the hash seed
        drawn.update(out)
        return out
    RIO.createPseudoNonessGenes = recording
    tmp = tempfile.mkdtemp()
    try:
        RIO.runJACKS(SMALL_COUNTS, SMALL_REPMAP, SMALL_COUNTS, common_ctrl_sample='CTRL', gene_hdr='Gene',
                     outprefix=tmp + '/small', ctrl_genes=NEG_GENES, n_pseudo=200)
    finally:
        RIO.createPseudoNonessGenes = orig
    files = {}
    for suffix in ['_gene_JACKS_results.txt', '_gene_std_JACKS_results.txt', '_gene_pval_JACKS_results.txt',
                   '_grna_JACKS_results.txt', '_logfoldchange_means.txt', '_logfoldchange_std.txt',
                   '_pseudo_noness_gene_JACKS_results.txt', '_pseudo_noness_gene_std_JACKS_results.txt']:
        with open(tmp + '/small' + suffix, 'rb') as f:
            files[suffix] = np.frombuffer(f.read(), dtype=np.uint8)
    names = list(drawn)
    offs = np.cumsum([0] + [len(drawn[n]) for n in names]).astype(np.int64)
    prow = np.concatenate([drawn[n] for n in names]).astype(np.int32)
    save('small_cli.npz', pseudo_offsets=offs, pseudo_rows=prow,
         **{'file' + k.replace('.txt', ''): v for k, v in files.items()})


def avana_sample():
    gene_index, testdata, ctrldata, ctrl_genes = synth.make_condensed_screen(n_genes=48, n_ctrl_genes=8)
    res, iters = ref_genes(gene_index, testdata, ctrldata)
    p = pack_results(gene_index, res, iters)
    # w_only + fixed_x with the published Avana efficacies (config 3): first 192 guides of the reffile
    xr = RIO.loadSgrnaReference('/root/reference/reference_grna_efficacies/avana_grna_JACKS_results.txt')
    keys = list(xr)[:testdata.shape[0]]
    fx = {'X1': np.array([float(xr[k]['X1']) for k in keys]), 'X2': np.array([float(xr[k]['X2']) for k in keys])}
    res_w, iters_w = ref_genes(gene_index, testdata, ctrldata, fixed_x=fx, w_only=True)
    glist = list(gene_index)
    save('avana_em.npz', x1=p['x1'], x2=p['x2'], w1=p['w1'], w2=p['w2'], tau_rowsum=p['tau'].sum(axis=1),
         tau_head=p['tau'][:16], iters=p['iters'],
         fx_X1=fx['X1'], fx_X2=fx['X2'], fx_w1=np.stack([res_w[g][4] for g in glist]),
         fx_w2=np.stack([res_w[g][5] for g in glist]), fx_iters=np.array([iters_w[g] for g in glist], dtype=np.int32))


def edge_cases():
    rng = np.random.default_rng(7)
    out = {}

    def make(G, L, w_sd=1.0):
        w = rng.normal(-0.5, w_sd, size=L)
        x = rng.normal(1.0, 0.4, size=G)
        sd_t = rng.uniform(0.1, 0.6, size=(G, L))
        sd_c = rng.uniform(0.05, 0.3, size=(G, L))
        c = rng.normal(0, 1, size=(G, L))
        d = c + np.outer(x, w) + rng.normal(0, 1, size=(G, L)) * np.sqrt(sd_t ** 2 + sd_c ** 2)
        return d, sd_t, c, sd_c

    def record(tag, d, de, c, ce, n_iter=50, **kw):
        (y, tau, x1, x2, w1, w2), it = ref_gene(d, de, c, ce, n_iter, **kw)
        out[tag + '.data'] = d; out[tag + '.data_err'] = de; out[tag + '.ctrl'] = c; out[tag + '.ctrl_err'] = ce
        out[tag + '.y'] = y; out[tag + '.tau'] = tau; out[tag + '.x1'] = x1; out[tag + '.x2'] = x2
        out[tag + '.w1'] = w1; out[tag + '.w2'] = w2; out[tag + '.iters'] = np.int32(it)
        for k, v in kw.items():
            if k == 'fixed_x':
                out[tag + '.fx1'] = v['X1']; out[tag + '.fx2'] = v['X2']
            else:
                out[tag + '.kw_' + k] = np.float64(v)
        out[tag + '.n_iter'] = np.int32(n_iter)

    for tag, (G, L) in {'g1': (1, 7), 'l1': (5, 1), 'g1l1': (1, 1), 'g2': (2, 33), 'g3': (3, 64), 'g7': (7, 31),
                        'g9': (9, 40), 'g18': (18, 5), 'g25': (25, 12), 'l1000': (4, 1000), 'g6l257': (6, 257)}.items():
        record(tag, *make(G, L))
    d, de, c, ce = make(5, 9)
    de2 = de.copy(); de2[:, 2] = np.nan; ce2 = ce.copy(); ce2[:, 4] = np.nan; ce2[1, 2] = np.nan
    record('nan_sd', d, de2, c, ce2)
    d2 = d.copy(); d2[1, 3] = np.nan; d2[4, 0] = np.nan; d2[0, 8] = np.nan
    record('nan_y', d2, de, c, ce)
    d3 = d.copy(); d3[:, 5] = np.nan                             # a whole line missing
    record('nan_line', d3, de, c, ce)
    d4 = d.copy(); d4[2, :] = np.nan                             # a whole guide missing
    record('nan_guide', d4, de, c, ce)
    rep = [0, 2, 2, 4, 0]                                        # pseudo-gene style repeated rows
    record('repeat_rows', d[rep], de[rep], c[rep], ce[rep])
    record('hp', d, de, c, ce, apply_w_hp=True)
    record('fixed', d, de, c, ce, fixed_x={'X1': rng.normal(1, 0.5, 5), 'X2': rng.uniform(0.5, 2.5, 5)})
    record('cap3', d, de, c, ce, n_iter=3)
    record('cap0', d, de, c, ce, n_iter=0)
    record('tol', d, de, c, ce, tol=1e-4)
    record('hyper', d, de, c, ce, mu0_x=0.8, var0_x=2.0, mu0_w=-0.1, var0_w=50.0, tau_prior_strength=0.7)
    dd, dde, cc, cce = make(4, 30, w_sd=0.01)                    # x1m close to zero / slow convergence
    record('flat', dd * 0.01, dde, cc * 0.01, cce)
    # common 1-D control (infer.py:61-65): ctrl [G], ctrl_err [G] with a NaN
    c1 = c[:, 0].copy(); ce1 = ce[:, 0].copy(); ce1[3] = np.nan
    record('ctrl1d', d, de, c1, ce1)
    save('edge_em.npz', **out)


def sd_model():
    rng = np.random.default_rng(11)
    out = {}
    for tag, (N, R) in {'n300r2': (300, 2), 'n4000r3': (4000, 3), 'n9000r2': (9000, 2), 'n100r4': (100, 4)}.items():
        lam = rng.lognormal(np.log(300.0), 0.8, size=N)
        cnt = rng.poisson(lam[:, None] * rng.uniform(0.7, 1.3, size=(1, R)))
        cnt[rng.random(N) < 0.02] = 0                             # ties on the sort key (all-zero rows)
        lc = np.log2(cnt + 32.0)
        out[tag + '.counts'] = cnt.astype(np.int32)
        for nt in ['median', 'zmad', 'mode']:
            out[tag + '.norm_' + nt] = RP.normalizeLogCounts(lc.copy(), normtype=nt)
        norm = out[tag + '.norm_median']
        out[tag + '.sd'] = RP.calc_posterior_sd(norm.copy())
        sub = sorted(rng.choice(N, size=max(70, N // 5), replace=False).tolist())
        out[tag + '.subset'] = np.array(sub, dtype=np.int32)
        out[tag + '.sd_subset'] = RP.calc_posterior_sd(norm.copy(), guideset_indexs=sub)
    x = rng.normal(size=500) ** 2
    out['win.x'] = x
    out['win.w30'] = RP.window_smooth(x.copy(), 30)
    out['win.mono'] = RP.monotonize(out['win.w30'].copy())
    xn = x[:100].copy(); xn[[10, 50, 51]] = np.nan
    out['mono.nan_in'] = xn
    out['mono.nan_out'] = RP.monotonize(xn.copy())
    save('sd_model.npz', **out)


def pvals(gene_index, res):
    """Seeded pseudo index + p-values straight from the reference (negative and positive selection)."""
    neg = sorted(g for g in (l.split()[0] for l in io.open(NEG_GENES) if l.strip()) if g in gene_index)
    random.seed(1234)
    pidx = RIO.createPseudoNonessGenes(gene_index, neg, 50)
    names = list(pidx)
    offs = np.cumsum([0] + [len(pidx[n]) for n in names]).astype(np.int64)
    prow = np.concatenate([pidx[n] for n in names]).astype(np.int32)
    rng = np.random.default_rng(5)
    L = 3
    pseudo_w1 = rng.normal(0.0, 0.2, size=(400, L))
    pseudo_w1[7, 1] = np.nan
    gene_w1 = np.concatenate([rng.normal(-0.8, 0.8, size=(60, L)), np.array([[-30.0, -3.0, 5.0], [0.0, 0.0, 0.0]])])
    results = {}
    for i in range(pseudo_w1.shape[0]):
        results['JACKS_PSEUDO_GENE_%d' % i] = (None, None, None, None, pseudo_w1[i], pseudo_w1[i] ** 2)
    for i in range(gene_w1.shape[0]):
        results['GENE%03d' % i] = (None, None, None, None, gene_w1[i], gene_w1[i] ** 2)
    pset = set(n for n in results if 'JACKS_PSEUDO_GENE' in n)
    lines = ['L%d' % l for l in range(L)]
    pv_neg, _ = RIO.computeW1PvalsAndFDRs(results, lines, noness_genes=pset, pseudo=True, positive=False)
    pv_pos, _ = RIO.computeW1PvalsAndFDRs(results, lines, noness_genes=pset, pseudo=True, positive=True)
    order = list(results)
    # branches the CLI does not reach: leave-one-out kernels for the control genes (pseudo=False) and the local fdr
    noness = set('GENE%03d' % i for i in range(0, 40, 2))
    real = {g: results[g] for g in results if g.startswith('GENE')}
    pv_loo, fdr_loo = RIO.computeW1PvalsAndFDRs(real, lines, noness_genes=noness, pseudo=False, compute_fdr=True)
    pv_loo_pos, _ = RIO.computeW1PvalsAndFDRs(real, lines, noness_genes=noness, pseudo=False, positive=True)
    finite = {g: results[g] for g in results if not np.isnan(results[g][4]).any()}     # scipy's evaluate() rejects NaN points
    _, fdr_pseudo = RIO.computeW1PvalsAndFDRs(finite, lines, noness_genes=pset, pseudo=True, compute_fdr=True)
    fdr_sets = RIO.getFDRGeneSets(pv_neg, 0.2)
    save('pvals.npz', seed=np.int64(1234), ctrl_sorted=np.array(neg), pseudo_offsets=offs, pseudo_rows=prow,
         pseudo_w1=pseudo_w1, gene_w1=gene_w1,
         pv_neg=np.stack([pv_neg[g] for g in order]), pv_pos=np.stack([pv_pos[g] for g in order]),
         order=np.array(order), real_order=np.array(list(real)), loo_noness=np.array(sorted(noness)),
         pv_loo=np.stack([pv_loo[g] for g in real]), pv_loo_pos=np.stack([pv_loo_pos[g] for g in real]),
         fdr_loo=np.stack([fdr_loo[g] for g in real]), fdr_pseudo=np.stack([fdr_pseudo[g] for g in finite]), finite_order=np.array(list(finite)),
         fdr_sets=np.array(['|'.join(sorted(s)) for s in fdr_sets]))


if __name__ == '__main__':
    small_inputs()
    gene_index, testdata, ctrldata, res = small_preproc_and_em()
    small_cli()
    avana_sample()
    edge_cases()
    sd_model()
    pvals(gene_index, res)
