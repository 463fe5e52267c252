#!/usr/bin/env python
"""Golden fixtures for the paper-script batch modes, produced by RUNNING THE UNMODIFIED REFERENCE.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    PYTHONHASHSEED=0 python tests/golden/make_golden_modes.py

  lines_em.npz   per-line ("single JACKS") inference exactly as run_JACKS_single.py:83-85 performs it -- one reference
                 inferJACKS call per line on the [:, [ts], :] slices -- on example-small (all 1,579 genes x 5 lines,
                 joint x and fixed x) and on 16 genes x 400 lines of the synthetic Avana-shape screen, with EM pass
                 counts; plus a few genes with missing values.
  bootstrap.npz  two draws of sample_jacks_screen.py:37-70 on example-small (replicate subsampling through the reference's
                 subsample_and_preprocess, guide subsampling with random.sample, inferJACKSGene per gene), seeded.
"""
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG                                   # noqa: E402  (reference imports + helpers)

RI, RP, RIO = MG.RI, MG.RP, MG.RIO


def ref_lines(gene_index, testdata, ctrldata, **kw):
    """run_JACKS_single.py:83-85 with pass counts: {gene: tuple with a line axis}, {gene: iters [L]}."""
    L = testdata.shape[1]
    cols, its = [], []
    for ts in range(L):
        r, it = MG.ref_genes(gene_index, testdata[:, [ts], :], ctrldata[:, [ts], :], **kw)
        cols.append(r)
        its.append(it)
    genes = list(gene_index)
    return dict(
        w1=np.array([[cols[ts][g][4][0] for ts in range(L)] for g in genes]),
        w2=np.array([[cols[ts][g][5][0] for ts in range(L)] for g in genes]),
        x1=np.concatenate([np.stack([cols[ts][g][2] for ts in range(L)], axis=1) for g in genes]),
        x2=np.concatenate([np.stack([cols[ts][g][3] for ts in range(L)], axis=1) for g in genes]),
        tau=np.concatenate([np.concatenate([cols[ts][g][1] for ts in range(L)], axis=1) for g in genes]),
        y=np.concatenate([np.concatenate([cols[ts][g][0] for ts in range(L)], axis=1) for g in genes]),
        iters=np.array([[its[ts][g] for ts in range(L)] for g in genes], dtype=np.int32))


def lines():
    sample_spec, ctrl_spec, _ = RIO.createSampleSpec(MG.SMALL_COUNTS, MG.SMALL_REPMAP, 'Replicate', 'Sample', 'CTRL')
    gene_spec = RIO.createGeneSpec(MG.SMALL_COUNTS, 'sgRNA', 'Gene', ignore_blank_genes=False)
    data, meta, sample_ids, genes, gene_index = RP.loadDataAndPreprocess(sample_spec, gene_spec, ctrl_spec=ctrl_spec)
    testdata, ctrldata, _ = RP.collateTestControlSamples(data, sample_ids, ctrl_spec)
    small = ref_lines(gene_index, testdata, ctrldata)
    g = np.load(os.path.join(HERE, 'small_em.npz'))
    fx = {'X1': g['fx_X1'], 'X2': g['fx_X2']}
    small_fx = ref_lines(gene_index, testdata, ctrldata, fixed_x=fx)

    gi, test, ctrl, _ = MG.synth.make_condensed_screen(n_genes=18000, n_lines=400)
    names = list(gi)[::1125][:16]
    sub = {n: gi[n] for n in names}
    av = ref_lines(sub, test, ctrl)
    av_rows = np.concatenate([gi[n] for n in names]).astype(np.int32)

    # missing values: NaN means and NaN sds in a copy of the first 40 example-small genes
    rng = np.random.default_rng(7)
    few = {n: gene_index[n] for n in list(gene_index)[:40]}
    t2, c2 = testdata.copy(), ctrldata.copy()
    rows = np.concatenate([few[n] for n in few])
    hit = rng.choice(rows, size=30, replace=False)
    t2[hit[:10], rng.integers(0, t2.shape[1], 10), 0] = np.nan
    t2[hit[10:20], rng.integers(0, t2.shape[1], 10), 1] = np.nan
    c2[hit[20:], rng.integers(0, t2.shape[1], 10), 1] = np.nan
    nan = ref_lines(few, t2, c2)
    MG.save('lines_em.npz',
            **{'small_' + k: v for k, v in small.items() if k != 'y'},
            **{'smallfx_' + k: small_fx[k] for k in ('w1', 'w2', 'iters')},
            av_names=np.array(names), av_rows=av_rows, **{'av_' + k: v for k, v in av.items() if k not in ('y', 'tau')},
            nan_names=np.array(list(few)), nan_test=t2, nan_ctrl=c2, **{'nan_' + k: v for k, v in nan.items()})


def bootstrap():
    """sample_jacks_screen.py:37-70, two draws, every cell line, two replicates each, at most 3 guides per gene."""
    sample_spec, ctrl_spec, _ = RIO.createSampleSpec(MG.SMALL_COUNTS, MG.SMALL_REPMAP, 'Replicate', 'Sample', 'CTRL')
    gene_spec = RIO.createGeneSpec(MG.SMALL_COUNTS, 'sgRNA', 'Gene', ignore_blank_genes=False)
    test_celllines = [s for s in ctrl_spec if ctrl_spec[s] != s]
    test_line = test_celllines[0]
    num_replicates, num_ctrl_reps, num_guides = 2, 2, 3
    random.seed(4242)
    out = {}
    for bs in range(2):
        selected = [test_line] + test_celllines
        screens = [(ctrl_spec[x], num_ctrl_reps) for x in selected] + [(x, num_replicates) for x in selected]
        data, meta, cell_lines, genes, gene_index = RP.subsample_and_preprocess(screens, sample_spec, gene_spec, ctrl_spec=ctrl_spec,
                                                                                ctrl_geneset=set(), normtype='median')
        testdata, ctrldata, _ = RP.collateTestControlSamples(data, cell_lines, ctrl_spec)
        idx = selected.index(test_line)
        w1s, w2s, picks = [], [], []
        for gene in gene_index:
            Ig = gene_index[gene]
            if num_guides > 0 and len(Ig) >= num_guides:
                Ig = random.sample(Ig, num_guides)
            y, tau, x1, x2, w1, w2 = RI.inferJACKSGene(testdata[Ig, :, 0], testdata[Ig, :, 1], ctrldata[Ig, :, 0], ctrldata[Ig, :, 1], 50)
            w1s.append(w1[idx]); w2s.append(w2[idx]); picks.append(np.array(Ig, dtype=np.int32))
        out['w1_%d' % bs] = np.array(w1s)
        out['w2_%d' % bs] = np.array(w2s)
        out['offsets_%d' % bs] = np.cumsum([0] + [len(p) for p in picks]).astype(np.int64)
        out['rows_%d' % bs] = np.concatenate(picks)
        out['data_%d' % bs] = data
        out['cell_lines_%d' % bs] = np.array(cell_lines)
    MG.save('bootstrap.npz', seed=np.int64(4242), test_line=np.array(test_line), genes=np.array(list(gene_index)), **out)


if __name__ == '__main__':
    lines()
    bootstrap()
