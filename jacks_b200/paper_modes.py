"""Batch forms of the reference's paper-script modes (2018_paper_materials/scripts/jacks/), SURVEY 8f-3.

The scripts drive ``jacks.infer`` in Python loops of thousands of small EM problems; here each mode is a handful of
batched GPU calls through the same C-ABI entry points as the CLI path.  Function names follow the scripts.

    run_JACKS_single.py          combineSingleResults, filterSampleSpec, filterCtrlSpec, inferSingleJACKS
                                 (every cell line on its own: one launch set for all (gene, line) pairs)
    sample_jacks_screen.py       sampleJacksScreen  (bootstrap over replicates / cell lines / guides: one batched EM
                                 per draw instead of one inferJACKSGene call per gene)
    leaveoneout_reference_test.py  filterOutSampleSpec, leaveOneOutReference  (x from all other lines, then the
                                 held-out line with fixed x)

No numerical work runs on the host except the scripts' own post-processing (sqrt, normal tail).
"""
import random

import numpy as np

from . import _native
from .infer import inferJACKS, inferJACKSLines, infer_batched
from .preprocess import collateTestControlSamples, subsample_and_preprocess


# ---- run_JACKS_single.py -------------------------------------------------------------------------------------------------

def combineSingleResults(single_jacks_results):
    """run_JACKS_single.py:8-15: stitch per-line result dicts; w1/w2 become lists over the lines."""
    jacks_results = {}
    for gene in single_jacks_results[0].keys():
        y, tau, x1, x2, w1, w2 = single_jacks_results[0][gene]
        w1 = [x[gene][4] for x in single_jacks_results]
        w2 = [x[gene][5] for x in single_jacks_results]
        jacks_results[gene] = y, tau, x1, x2, w1, w2
    return jacks_results


def filterSampleSpec(sample_spec, cell_line, ctrl_spec):
    """run_JACKS_single.py:17-25: keep the replicates of one cell line and of its control."""
    new_sample_spec = {}
    for filename in sample_spec:
        for sample_id, colname in sample_spec[filename]:
            if sample_id == cell_line or sample_id == ctrl_spec[cell_line]:
                new_sample_spec.setdefault(filename, []).append((sample_id, colname))
    return new_sample_spec


def filterCtrlSpec(ctrl_spec, cell_line):
    """run_JACKS_single.py:27-31."""
    return {cell_line: ctrl_spec[cell_line], ctrl_spec[cell_line]: ctrl_spec[cell_line]}


def inferSingleJACKS(gene_index, testdata, ctrldata):
    """run_JACKS_single.py:81-86: ``inferJACKS(gene_index, testdata[:,[ts],:], ctrldata[:,[ts],:], w_only=True)`` for every
    line ts, combined.  Returns {gene: (None, None, None, None, w1 [L], w2 [L])} (arrays where the script builds lists
    of one-element arrays; writeJacksWResults formats both alike)."""
    return inferJACKSLines(gene_index, testdata, ctrldata, w_only=True)


# ---- sample_jacks_screen.py ----------------------------------------------------------------------------------------------

def _norm_sf(z):
    """scipy.stats.norm.sf without the import cost at module load."""
    from scipy.special import ndtr
    return ndtr(-np.asarray(z, dtype=np.float64))


def sampleJacksScreen(sample_spec, gene_spec, ctrl_spec, test_line, num_replicates, num_celllines, num_samples, num_guides,
                      ctrl_geneset=set(), num_ctrl_reps=-1, normtype='median', details=None):
    """sample_jacks_screen.py:37-70: `num_samples` bootstrap draws of a screen -- replicates subsampled per cell line
    (subsample_and_preprocess), optionally `num_celllines` - 1 random companion lines and at most `num_guides` guides
    per gene -- each inferred jointly; collects, for `test_line`, (w1, sqrt(w2 - w1^2), 1 - sf(w1 / std)) per draw.

    The random draws are made in the script's order (cell lines, replicates inside subsample_and_preprocess, then one
    random.sample per gene in gene_index order), so a seeded run selects what the script selects.  `num_ctrl_reps` is
    a name the script uses without defining; here it is an argument (-1: all control replicates).
    Returns {gene: [(wmean, wstd, pval) per draw]}; `details`, if a list, receives per draw a dict with the drawn
    gene index (CSR), cell lines and w1/w2 of the test line."""
    test_celllines = [sample_id for sample_id in ctrl_spec if ctrl_spec[sample_id] != sample_id]
    gene_ws = None
    for bs in range(num_samples):
        if num_celllines < 0:
            selected_celllines = [test_line] + test_celllines
        else:
            selected_celllines = [test_line] + [random.choice(test_celllines) for i in range(num_celllines - 1)]
        selected_screens = [(ctrl_spec[x], num_ctrl_reps) for x in selected_celllines] + \
                           [(x, num_replicates) for x in selected_celllines]
        data, meta, cell_lines, genes, gene_index = subsample_and_preprocess(
            selected_screens, sample_spec, gene_spec, ctrl_spec=ctrl_spec, ctrl_geneset=ctrl_geneset, normtype=normtype)
        testdata, ctrldata, _ = collateTestControlSamples(data, cell_lines, ctrl_spec)
        if testdata.shape != ctrldata.shape:
            raise Exception('Mismatch between test and control shape')
        if bs == 0:
            gene_ws = {gene: [] for gene in gene_index}
        idx = selected_celllines.index(test_line)
        drawn = {}
        for gene in gene_index:
            Ig = gene_index[gene]
            if num_guides > 0 and len(Ig) >= num_guides:
                Ig = random.sample(Ig, num_guides)
            drawn[gene] = Ig
        names, offsets, rows = _native.as_csr(drawn)
        r = infer_batched(offsets, rows, testdata, ctrldata, n_iter=50, want=('w1', 'w2'))
        w1, w2 = r['w1'][:, idx], r['w2'][:, idx]
        with np.errstate(all='ignore'):
            std = np.sqrt(w2 - w1 ** 2.0)
            pval = 1.0 - _norm_sf(w1 / std)
        for i, gene in enumerate(names):
            gene_ws[gene].append((w1[i], std[i], pval[i]))
        if details is not None:
            details.append(dict(offsets=offsets, rows=rows, cell_lines=list(cell_lines), data=data, w1=w1.copy(), w2=w2.copy()))
    return gene_ws


def writeSampledScreen(outfile, gene_ws, num_samples):
    """sample_jacks_screen.py:72-77 (the caller composes the file name)."""
    import io
    with io.open(outfile, 'w') as fout:
        fout.write(u'Gene\t%s\n' % ('\t'.join(['WMean\tWStd\tPval' for i in range(num_samples)])))
        for gene in gene_ws:
            fout.write(u'%s\t%s\n' % (gene, '\t'.join(['%5e\t%5e\t%5e' % (wmean, wstd, pval) for (wmean, wstd, pval) in gene_ws[gene]])))


# ---- leaveoneout_reference_test.py ---------------------------------------------------------------------------------------

def filterOutSampleSpec(sample_spec, cell_line, ctrl_spec):
    """leaveoneout_reference_test.py:10-18: drop the replicates of one cell line."""
    new_sample_spec = {}
    for filename in sample_spec:
        for sample_id, colname in sample_spec[filename]:
            if sample_id == cell_line:
                continue
            new_sample_spec.setdefault(filename, []).append((sample_id, colname))
    return new_sample_spec


def leaveOneOutReference(gene_index, testdata, ctrldata, lines=None):
    """leaveoneout_reference_test.py:59-91 for every (or the given) held-out line at the condensed level: guide
    efficacies x from a joint run over all OTHER lines (:68-69), then the held-out line alone with those x fixed and
    w_only (:87).  Per-column preprocessing makes the script's two loadDataAndPreprocess calls equal to column subsets
    of one cube, so the cube is preprocessed once by the caller.  (The script round-trips x through a '%5e' text file,
    :70 and :80-84; that 7-digit rounding is a file-format artefact and is not reproduced: x is passed on in full
    precision.)  Returns {gene: (None, None, None, None, w1 [n_held_out], w2 [n_held_out])} and the list of lines."""
    testdata, ctrldata = np.asarray(testdata), np.asarray(ctrldata)
    if testdata.shape != ctrldata.shape:
        raise Exception('Expecting matched size between control and test data')
    N, L = testdata.shape[0], testdata.shape[1]
    lines = list(range(L)) if lines is None else [int(l) for l in lines]
    names, offsets, rows = _native.as_csr(gene_index)
    w1 = np.empty((len(names), len(lines)))
    w2 = np.empty((len(names), len(lines)))
    ctx = _native.context(0)
    for k, held in enumerate(lines):
        keep = [l for l in range(L) if l != held]
        if not keep:
            raise Exception('leave-one-out needs at least two lines')
        ref = infer_batched(offsets, rows, np.ascontiguousarray(testdata[:, keep, :]), np.ascontiguousarray(ctrldata[:, keep, :]),
                            want=('x1', 'x2'))
        fx1, fx2 = np.zeros(N), np.zeros(N)
        fx1[rows], fx2[rows] = ref['x1'], ref['x2']
        one = _native.em_lines_host(ctx, offsets, rows, np.ascontiguousarray(testdata[:, [held], :]),
                                    np.ascontiguousarray(ctrldata[:, [held], :]), fixed_x1=fx1, fixed_x2=fx2, want=('w1', 'w2'))
        w1[:, k], w2[:, k] = one['w1'][:, 0], one['w2'][:, 0]
    return {gene: (None, None, None, None, w1[i], w2[i]) for i, gene in enumerate(names)}, lines
