"""Drop-in for the reference module ``jacks.jacks_io`` (jacks/jacks/jacks_io.py): specification readers,
pseudo-gene construction, p-values, result writers, orchestration and the command-line parser, with
the same names, arguments, defaults, file formats and error messages.

Host side by design (as in the reference): file parsing, the ``random`` draws of the pseudo-genes,
gene ranking, Benjamini-Hochberg sets and text formatting -- so that indexing and ranking are
bit-exact.  GPU side: the EM (``inferJACKS``), the preprocessing error model and the KDE p-values / local
fdr (``computeW1PvalsAndFDRs``, every branch).
"""
import argparse
import csv
import io
import os
import random

import numpy as np

from . import _native
from .infer import LOG, inferJACKS, inferJACKSScreen
from .preprocess import loadDataAndPreprocess, collateTestControlSamples

REP_HDR_DEFAULT = "Replicate"
SAMPLE_HDR_DEFAULT = "Sample"
COMMON_CTRL_SAMPLE_DEFAULT = "CONTROL"
SGRNA_HDR_DEFAULT = "sgRNA"
GENE_HDR_DEFAULT = "Gene"
OUTPREFIX_DEFAULT = ""
APPLY_W_HP_DEFAULT = False
PICKLE_FILENAME = '_JACKS_results_full.pickle'
NORM_TYPE_DEFAULT = 'median'
GENE_FILENAME = '_gene_JACKS_results.txt'
GRNA_FILENAME = '_grna_JACKS_results.txt'


def getGeneWs(jacks_results, gene):
    return jacks_results[gene][4]


def _mean_w1(jacks_results):
    """np.nanmean(w1) per gene.  Rows of equal length are averaged in one vectorised call (identical bits: numpy
    reduces each contiguous row with the same pairwise sum as the 1-D call)."""
    genes = [g for g in jacks_results]
    if genes:
        rows = [np.asarray(jacks_results[g][4], dtype=np.float64) for g in genes]
        if all(r.ndim == 1 and r.shape == rows[0].shape for r in rows) and rows[0].size > 0:
            W = np.stack(rows)
            if not np.isnan(W).all(axis=1).any():
                return list(zip(np.nanmean(W, axis=1), genes))
    return [(np.nanmean(jacks_results[gene][4]), gene) for gene in genes]


def getSortedGenes(jacks_results, positive):
    """[(mean w1 over lines, gene)] ranked by effect, strongest depletion first (enrichment first if
    ``positive``); ties break on the gene name (jacks_io.py:26-31)."""
    ordered_genes = _mean_w1(jacks_results)
    ordered_genes.sort(reverse=bool(positive))
    return ordered_genes


def getFDRGeneSets(jacks_w1_pvals, fdr):
    """Per cell line the set of genes passing Benjamini-Hochberg at ``fdr`` (jacks_io.py:52-61)."""
    fdr_gene_sets = []
    a_gene = [x for x in jacks_w1_pvals][0]
    n = len(jacks_w1_pvals)
    for sample_idx in range(len(jacks_w1_pvals[a_gene])):
        pval_genes = sorted((jacks_w1_pvals[gene][sample_idx], gene) for gene in jacks_w1_pvals)
        below_idxs = [i for i, (pval, gene) in enumerate(pval_genes) if pval < fdr * (i + 1) / n]
        k = max(below_idxs) if len(below_idxs) > 0 else -1
        fdr_gene_sets.append(set(gene for (pval, gene) in pval_genes[:k + 1]))
    return fdr_gene_sets


def getLocalFDRGeneSets(jacks_w1_fdrs, test_sample_ids, fdr):
    """Genes whose local fdr is below the threshold, per line.  (The reference's version of this function,
    jacks_io.py:63-67, cannot run: it is called with two arguments and refers to an undefined name.)"""
    return [set(gene for gene in jacks_w1_fdrs if jacks_w1_fdrs[gene][sample_idx] < fdr)
            for sample_idx, _ in enumerate(test_sample_ids)]


def computeW1PvalsAndFDRs(jacks_results, test_sample_ids, noness_genes=set(), pseudo=False, compute_fdr=False,
                          positive=False):
    """P-value of every gene's w1, per cell line, under the Gaussian-KDE null built from the w1 of the
    ``noness_genes`` (jacks_io.py:69-89).  Returns ({gene: pvals[L]}, {gene: fdrs[L]}).

    All KDE sums run on the GPU.  pseudo=True (the command-line path): one shared kernel per line
    (jacks_kde_pvalues_host).  pseudo=False: the control genes themselves are scored under a kernel that leaves
    them out (jacks_kde_eval_host, leave_one_out).  compute_fdr: local fdr = density under the null kernel /
    density under the kernel of all genes (jacks_io.py:76-77,88)."""
    genes = [g for g in jacks_results]
    L = len(test_sample_ids)
    w1_fdrs = {gene: np.zeros(L) for gene in genes}
    if len(genes) == 0:
        return {}, w1_fdrs
    ctx = _native.context(0)
    w1_all = np.stack([np.asarray(jacks_results[g][4], dtype=np.float64) for g in genes])
    null_rows = [i for i, g in enumerate(genes) if g in noness_genes]
    w1_null = w1_all[null_rows]
    pv = _native.kde_pvalues_host(ctx, w1_all, w1_null, positive=positive)
    if not pseudo and null_rows:
        pv[null_rows] = _native.kde_eval_host(ctx, w1_null, w1_null, 'cdf_pos' if positive else 'cdf_neg', leave_one_out=True)
    w1_pvals = {gene: pv[i] for i, gene in enumerate(genes)}
    if compute_fdr:
        dens_null = _native.kde_eval_host(ctx, w1_all, w1_null, 'pdf')
        if not pseudo and null_rows:
            dens_null[null_rows] = _native.kde_eval_host(ctx, w1_null, w1_null, 'pdf', leave_one_out=True)
        dens_all = _native.kde_eval_host(ctx, w1_all, w1_all, 'pdf')
        fdr = dens_null / dens_all
        w1_fdrs = {gene: fdr[i] for i, gene in enumerate(genes)}
    return w1_pvals, w1_fdrs


def writeJacksWResults(outprefix, jacks_results, cell_lines, write_types=[''], ctrl_geneset=set(), fdr=None,
                       fdr_thresh_type='REGULAR', pseudo=False, positive=False, w1_pvals=None):
    """Gene-level result tables, one file per write type ('' = E[w], '_std', '_pval', '_fdr'), genes ranked by
    mean w1, values formatted '%5e' (jacks_io.py:91-133).  w1_pvals: p-values already computed on the GPU by the screen
    job (load_data_and_run); otherwise computeW1PvalsAndFDRs is called here, as in the reference."""
    ordered_genes = getSortedGenes(jacks_results, positive)
    # the files are opened first, as in the reference, so that a failing p-value step leaves the same (empty) files
    fouts = [io.open(outprefix + '_gene%s_JACKS_results.txt' % write_type, 'w') for write_type in write_types]
    if w1_pvals is not None and '_fdr' not in write_types:
        jacks_w1_pvals, jacks_w1_fdrs = w1_pvals, None
    elif '_fdr' in write_types or '_pval' in write_types:
        LOG.info('Computing P-values for %s selection' % ('positive' if positive else 'negative'))
        jacks_w1_pvals, jacks_w1_fdrs = computeW1PvalsAndFDRs(jacks_results, cell_lines, noness_genes=ctrl_geneset,
                                                              pseudo=pseudo, compute_fdr=('_fdr' in write_types),
                                                              positive=positive)
    if fdr is not None:
        if fdr_thresh_type == 'REGULAR':
            fdr_sets = getFDRGeneSets(jacks_w1_pvals, fdr)
        elif fdr_thresh_type == 'LOCAL_FDR':
            fdr_sets = getLocalFDRGeneSets(jacks_w1_fdrs, cell_lines, fdr)
        else:
            raise Exception('Unrecognised FDR threshold type (expecting REGULAR or LOCAL_FDR): ', fdr_thresh_type)

    for fout in fouts:
        fout.close()
    kept, flags = [], []
    if fdr is None:
        # every line of every gene counts as significant (jacks_io.py:118): no per-line flag lists, nothing blanked
        kept = [gene for w1_mean, gene in ordered_genes if len(jacks_results[gene][4]) > 0 and 'JACKS_PSEUDO_GENE' not in gene]
        flags = None
    else:
        for w1_mean, gene in ordered_genes:
            sig_gene_flags = [(gene in x) for x in fdr_sets]
            if sum(sig_gene_flags) == 0:
                continue
            if 'JACKS_PSEUDO_GENE' in gene:
                continue                              # pseudo-genes are never written (jacks_io.py:131)
            kept.append(gene)
            flags.append(sig_gene_flags)
    L = len(cell_lines)
    W1 = np.array([jacks_results[g][4] for g in kept], dtype=np.float64).reshape(len(kept), L)
    header = u'Gene\t%s\n' % ('\t'.join(cell_lines))
    for write_type in write_types:
        blank = None
        if write_type == '_pval':
            vals = np.array([jacks_w1_pvals[g] for g in kept], dtype=np.float64).reshape(len(kept), L)
        elif write_type == '_fdr':
            vals = np.array([jacks_w1_fdrs[g] for g in kept], dtype=np.float64).reshape(len(kept), L)
        elif write_type == '_std':
            W2 = np.array([jacks_results[g][5] for g in kept], dtype=np.float64).reshape(len(kept), L)
            with np.errstate(invalid='ignore'):
                vals = np.sqrt(W2 - W1 ** 2.0)
        elif write_type == '':
            vals = W1
            if flags is not None:
                blank = ~np.array(flags, dtype=bool).reshape(len(kept), L)
        else:
            raise Exception('Unrecognised write type: %s' % write_type)
        _native.table_write(outprefix + '_gene%s_JACKS_results.txt' % write_type, header, kept, vals, 5, blank=blank)


def writeJacksXResults(filename, jacks_results, gene_grnas):
    """Guide efficacies sgrna / X1 / X2, guides grouped by gene, genes ranked by mean w1 (jacks_io.py:135-146)."""
    ordered_genes = sorted(_mean_w1(jacks_results))
    labels, x1s, x2s = [], [], []
    for w1_mean, gene in ordered_genes:
        y, tau, x1, x2, w1, w2 = jacks_results[gene]
        grnas = gene_grnas[gene]
        labels.extend(grnas)
        x1s.append(np.asarray(x1, dtype=np.float64)[:len(grnas)])
        x2s.append(np.asarray(x2, dtype=np.float64)[:len(grnas)])
    vals = np.stack([np.concatenate(x1s), np.concatenate(x2s)], axis=1) if labels else np.zeros((0, 2))
    _native.table_write(filename, u'sgrna\tX1\tX2\n', [str(g) for g in labels], vals, 5)


def writeFoldChanges(filename, testdata, ctrldata, meta, sample_ids, write_std=False, write_raw=False):
    """Per-guide log fold changes (or their standard deviations) against the control (jacks_io.py:148-156)."""
    if write_std:
        vals = np.sqrt(testdata[:, :, 1] ** 2.0 + ctrldata[:, :, 1] ** 2.0)
    elif write_raw:
        vals = np.sqrt(testdata[:, :, 2] ** 2.0 + ctrldata[:, :, 2] ** 2.0)
    else:
        vals = testdata[:, :, 0] - ctrldata[:, :, 0]
    labels = ['%s\t%s' % (meta[i, 0], meta[i, 1]) for i in range(len(meta[:, 0]))]
    _native.table_write(filename, u'gRNA\tgene\t%s\n' % '\t'.join(sample_ids), labels, vals, 6)


def writeLogData(filename, data, meta, sample_ids, write_std=False, write_raw=False):
    """Per-guide condensed log counts (mean, sd or raw plane) of every sample (jacks_io.py:158-166)."""
    fout = io.open(filename, 'w')
    fout.write(u'gRNA\tgene\t%s\n' % '\t'.join(sample_ids))
    plane = 1 if write_std else (2 if write_raw else 0)
    for i in range(len(meta[:, 0])):
        fout.write(u'%s\t%s\t%s\n' % (meta[i, 0], meta[i, 1], '\t'.join(['%6e' % x for x in data[i, :, plane]])))
    fout.close()


def pickleJacksFullResults(filename, jacks_results, cell_lines, gene_grnas):
    """[jacks_results, cell_lines, gene_grnas] -- the layout plot_heatmap.py and the web server load (jacks_io.py:168-173)."""
    import pickle
    with io.open(filename, 'wb') as f:
        pickle.dump([jacks_results, cell_lines, gene_grnas], f)


def loadJacksFullResultsFromPickle(filename):
    import pickle
    with io.open(filename, 'rb') as f:
        jacks_results, cell_lines, gene_grnas = pickle.load(f)
    return jacks_results, cell_lines, gene_grnas


def prepareFile(filename, hdr):
    """Open ``filename`` positioned at the header line containing ``hdr`` (at most 100 lines are skipped) and
    guess the delimiter from the extension (jacks_io.py:182-197)."""
    f = io.open(filename)
    skip_lines, line = 0, f.readline()
    while hdr not in line and skip_lines < 100:
        skip_lines += 1
        line = f.readline()
    f.close()
    if skip_lines >= 100:
        raise Exception('Could not find line with header ' + hdr + ' in ' + filename)
    delim = ',' if (filename.split('.')[-1] == 'csv') else '\t'
    f = io.open(filename)
    for i in range(skip_lines):
        f.readline()
    return f, delim


def createSampleSpec(infile, repfile, rep_hdr, sample_hdr, common_ctrl_sample, ctrl_sample_hdr=None):
    """Replicate map -> ({infile: [(sample_id, replicate column)]}, {sample: control sample}, {sample: n replicates})
    (jacks_io.py:200-226)."""
    f, delim = prepareFile(repfile, rep_hdr)
    rdr = csv.DictReader(f, delimiter=delim)
    sample_spec, sample_num_reps = {infile: []}, {}
    ctrl_per_sample = ctrl_sample_hdr is not None
    if ctrl_per_sample and ctrl_sample_hdr not in rdr.fieldnames:
        raise Exception('Could not find column %s specifying controls in %s' % (ctrl_sample_hdr, repfile))
    ctrl_spec = {}
    for row in rdr:
        sample_id, rep_id = row[sample_hdr], row[rep_hdr]
        sample_num_reps[sample_id] = sample_num_reps.get(sample_id, 0) + 1
        sample_spec[infile].append((sample_id, rep_id))
        if ctrl_per_sample:
            if sample_id in ctrl_spec:
                if ctrl_spec[sample_id] != row[ctrl_sample_hdr]:
                    err_msg = '%s vs %s for %s\n' % (ctrl_spec[sample_id], row[ctrl_sample_hdr], sample_id)
                    raise Exception(err_msg + 'Different controls for replicates of the sample not supported.')
            else:
                ctrl_spec[sample_id] = row[ctrl_sample_hdr]
        else:
            ctrl_spec[sample_id] = common_ctrl_sample
    f.close()
    if not ctrl_per_sample and common_ctrl_sample not in ctrl_spec:
        raise Exception('Could not find control sample %s in %s' % (common_ctrl_sample, repfile))
    return sample_spec, ctrl_spec, sample_num_reps


def createGeneSpec(guidemappingfile, sgrna_hdr, gene_hdr, ignore_blank_genes=True):
    """Guide map -> {guide: gene} (jacks_io.py:229-234)."""
    f, delim = prepareFile(guidemappingfile, sgrna_hdr)
    header_line = f.readline()
    fieldnames = next(csv.reader([header_line], delimiter=delim), None)
    f.close()
    simple = (fieldnames and header_line.count('"') % 2 == 0 and len(set(fieldnames)) == len(fieldnames)
              and sgrna_hdr in fieldnames and gene_hdr in fieldnames)
    if simple:
        # native two-column read (the guide map is usually the count table itself: hundreds of unused columns)
        skip = 0
        with io.open(guidemappingfile) as g:
            line = g.readline()
            while sgrna_hdr not in line and skip < 100:
                skip += 1
                line = g.readline()
        pairs, _, _ = _native.table_read(guidemappingfile, delim, [], skip_lines=skip + 1,
                                         str_cols=(fieldnames.index(sgrna_hdr), fieldnames.index(gene_hdr)))
        # raw field bytes come back: a quoted field anywhere (csv would unescape it) or a short record sends the whole
        # file through csv.DictReader instead
        if all(a is not None and b is not None and '"' not in a and '"' not in b for a, b in pairs):
            return {a: b for a, b in pairs if (not ignore_blank_genes or b != '')}
    f, delim = prepareFile(guidemappingfile, sgrna_hdr)
    rdr = csv.DictReader(f, delimiter=delim)
    gene_spec = {row[sgrna_hdr]: row[gene_hdr] for row in rdr if (not ignore_blank_genes or row[gene_hdr] != '')}
    f.close()
    return gene_spec


def readControlGeneset(ctrl_genes, gene_spec):
    """A gene name, or a file with one gene identifier per line; unknown genes are dropped (jacks_io.py:237-248)."""
    known_genes = set(gene_spec[x] for x in gene_spec)
    if os.path.isfile(ctrl_genes):
        with io.open(ctrl_genes) as f:
            geneset = set(line.split()[0] for line in f if line.split()[0] in known_genes)
        LOG.info('Read %d recognised control genes from %s' % (len(geneset), ctrl_genes))
    else:
        if ctrl_genes not in known_genes:
            raise Exception('Not a file or unrecognised control gene: %s' % ctrl_genes)
        geneset = set([ctrl_genes])
        LOG.info('Using %s as control gene' % (ctrl_genes))
    return geneset


def loadSgrnaReference(filename):
    """Reference efficacies {sgrna: {'sgrna', 'X1', 'X2'}} (strings, as read) (jacks_io.py:250-254)."""
    with io.open(filename) as f:
        x_ref = {row['sgrna']: row for row in csv.DictReader(f, delimiter='\t')}
    return x_ref


def _drawPseudoGenes(gene_index, ctrl_geneset, num_psuedo_noness_genes):
    """The draws of createPseudoNonessGenes: (pseudo_gene_index, csr) with csr = (names, offsets, rows) of the same index,
    or None when the draws were made by the interpreter.

    The reference makes one random.choice over all genes per pseudo-gene (its guide count) and two per guide (a control
    gene, one of its guides) -- 900,000 interpreter-level calls at 1e5 pseudo-genes.  Here the state of the global
    ``random`` generator is handed to jacks_pseudo_draw (csrc/host_io.cpp), which runs the same Mersenne Twister through
    the same rejection sampling, and the advanced state is put back: same sets, same stream position afterwards.
    Gene indexes whose entries are not plain integers keep the interpreter loop."""
    ctrl_geneset_lst = [x for x in ctrl_geneset if x in gene_index]
    n = int(num_psuedo_noness_genes)
    state = random.getstate()
    native = n > 0 and state[0] == 3 and len(state[1]) == 625
    if native:
        try:
            lens = [len(gene_index[g]) for g in gene_index]
            flat = [r for g in gene_index for r in gene_index[g]]
            native = all(isinstance(r, (int, np.integer)) for r in flat) and \
                (len(flat) == 0 or (min(flat) >= -2 ** 31 and max(flat) < 2 ** 31))
        except TypeError:
            native = False
    if not native:
        pseudo_gene_index = {}
        all_geneset_lst = [x for x in gene_index]
        for i in range(n):
            num_guides = len(gene_index[random.choice(all_geneset_lst)])
            pseudo_gene_index['JACKS_PSEUDO_GENE_%d' % i] = [random.choice(gene_index[random.choice(ctrl_geneset_lst)])
                                                             for j in range(num_guides)]
        return pseudo_gene_index, None
    offsets = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(np.asarray(lens, dtype=np.int64), out=offsets[1:])
    pos = {g: i for i, g in enumerate(gene_index)}
    try:
        new_state, poffs, prows = _native.pseudo_draw(state[1], offsets, np.asarray(flat, dtype=np.int32),
                                                      [pos[g] for g in ctrl_geneset_lst], n)
    except _native.JacksError as e:
        if 'empty sequence' in str(e):
            raise IndexError('Cannot choose from an empty sequence')          # what random.choice raises in the reference
        raise
    random.setstate((state[0], tuple(int(v) for v in new_state), state[2]))
    names = ['JACKS_PSEUDO_GENE_%d' % i for i in range(n)]
    lo, rl = poffs.tolist(), prows.tolist()
    pseudo_gene_index = {names[i]: rl[lo[i]:lo[i + 1]] for i in range(n)}
    return pseudo_gene_index, (names, poffs, prows)


def createPseudoNonessGenes(gene_index, ctrl_geneset, num_psuedo_noness_genes):
    """Pseudo-genes: the guide count of a random gene, then that many guides, each the random guide of a random
    control gene, with replacement (jacks_io.py:256-264).  Consumes the global ``random`` stream draw for draw
    like the reference, so the same seed and the same iteration order of ``ctrl_geneset`` give the same sets."""
    return _drawPseudoGenes(gene_index, ctrl_geneset, num_psuedo_noness_genes)[0]


_createPseudoNonessGenes = createPseudoNonessGenes


def getJacksParser():
    """Command-line parser of run_JACKS.py: same positionals, flags, defaults and help (jacks_io.py:266-350)."""
    ap = argparse.ArgumentParser(formatter_class=argparse.ArgumentDefaultsHelpFormatter, description="")
    ap.add_argument("countfile", help="Countfile: assumes sgrna label is always in the first column!")
    rg = ap.add_argument_group("Replicate file arguments")
    rg.add_argument("replicatefile", help="A CSV or tab delimited file mapping replicates to samples")
    rg.add_argument("--rep_hdr", type=str, default=REP_HDR_DEFAULT,
                    help="Column header for column containing the replicate labels")
    rg.add_argument("--sample_hdr", type=str, default=SAMPLE_HDR_DEFAULT,
                    help="Column header for column containing the sample labels")
    rg.add_argument("--common_ctrl_sample", type=str, default=COMMON_CTRL_SAMPLE_DEFAULT,
                    help="Name of the sample to be used as a control for all other samples")
    rg.add_argument("--ctrl_sample_hdr", type=str, default=None,
                    help="Column header for column containing the sample label for the control to be used for each sample")
    gg = ap.add_argument_group("Guidemapping file arguments")
    gg.add_argument("guidemappingfile", help="A CSV or tab delimited file mapping guides to genes")
    gg.add_argument("--sgrna_hdr", type=str, default=SGRNA_HDR_DEFAULT,
                    help="Column header for the column containing the guide labels")
    gg.add_argument("--gene_hdr", type=str, default=GENE_HDR_DEFAULT,
                    help="Column header for the column containing the gene labels")
    gg.add_argument("--ignore_blank_genes", action='store_true', default=False,
                    help="Ignore guides for which the assigned gene is blank")
    ap.add_argument("--outprefix", type=str, default="", help="Output prefix")
    ap.add_argument("--reffile", type=str, default=None, help="Reference file containing pre-trained sgRNA efficacy scores")
    ap.add_argument("--fdr", type=float, default=None,
                    help="False discovery rate threshold at which to accept significant genes (if None output all in order)")
    ap.add_argument("--fdr_thresh_type", type=str, default='REGULAR',
                    help="Method to use to threshold significant genes (REGULAR, LOCAL_FDR)")
    ap.add_argument("--positive", action='store_true', default=False,
                    help="Apply positive selection (default is negative selection)")
    ap.add_argument("--apply_w_hp", action='store_true', default=False,
                    help="Apply hierarchical prior on gene effects (not recommended, use with caution)")
    ap.add_argument("--norm_type", type=str, default=NORM_TYPE_DEFAULT,
                    help="Type of normalisation to use on log count data (median, mode, ctrl_guides - if using ctrl_guides, "
                         "must provide --ctrl_gene_file)")
    ap.add_argument("--ctrl_genes", type=str, default=None,
                    help="(Required if p-value output is wanted) Either, the name of a gene (as used in sgrnamappingfile) "
                         "specifying a set of negative control guides (e.g. these could be intergenic, non-targeting etc) OR "
                         "a text file containing a list (one per line) of genes to use as negative controls. Also used to "
                         "infer variances in case of single replicate data.")
    ap.add_argument("--n_pseudo", type=int, default=2000,
                    help="Number of pseudo genes to create and infer JACKS results for by randomly sampling guides from the "
                         "control genes (specified in --ctrl_genes)")
    ap.add_argument("--count_prior", type=int, default=32,
                    help="Prior count to be added to all read counts prior to log operation")
    return ap


def preprocess(countfile, replicatefile, guidemappingfile, rep_hdr=REP_HDR_DEFAULT, sample_hdr=SAMPLE_HDR_DEFAULT,
               common_ctrl_sample=COMMON_CTRL_SAMPLE_DEFAULT, ctrl_sample_hdr=None, sgrna_hdr=SGRNA_HDR_DEFAULT,
               gene_hdr=GENE_HDR_DEFAULT, ignore_blank_genes=False, outprefix=OUTPREFIX_DEFAULT, reffile=None):
    """Read the sample, gene and (optional) reference-efficacy specifications (jacks_io.py:352-377)."""
    LOG.info('Loading sample specification')
    sample_spec, ctrl_spec, sample_num_reps = createSampleSpec(countfile, replicatefile, rep_hdr, sample_hdr,
                                                               common_ctrl_sample, ctrl_sample_hdr)
    LOG.info('Loading gene mappings')
    gene_spec = createGeneSpec(guidemappingfile, sgrna_hdr, gene_hdr, ignore_blank_genes=ignore_blank_genes)
    x_ref = None
    if reffile:
        LOG.info('Loading sgrna reference values')
        x_ref = loadSgrnaReference(reffile)
        LOG.info('Checking sgrna reference identifiers against gene mappings')
        for guide in gene_spec:
            if guide not in x_ref:
                raise Exception('%s has no sgrna reference in %s' % (guide, reffile))
    return sample_spec, ctrl_spec, gene_spec, x_ref


def load_data_and_run(sample_spec, gene_spec, ctrl_spec, sgrna_reference_file, x_ref, outprefix,
                      apply_w_hp=APPLY_W_HP_DEFAULT, norm_type=NORM_TYPE_DEFAULT, ctrl_genes=None, fdr=None,
                      fdr_thresh_type='REGULAR', n_pseudo=0, count_prior=32, positive=False):
    """Preprocess, infer, (optionally) build the pseudo-gene null and write every output file (jacks_io.py:379-428)."""
    _native.warm_context(0)
    ctrl_geneset = readControlGeneset(ctrl_genes, gene_spec) if ctrl_genes is not None else set()
    if '/' in outprefix and not os.path.exists(os.path.dirname(outprefix)):
        os.makedirs(os.path.dirname(outprefix))
    outfile_x = outprefix + '_grna_JACKS_results.txt'
    outfile_lfc = outprefix + '_logfoldchange_means.txt'
    outfile_lfc_std = outprefix + '_logfoldchange_std.txt'
    outfile_pickle = outprefix + PICKLE_FILENAME

    LOG.info('Loading data and pre-processing')
    data, meta, sample_ids, genes, gene_index = loadDataAndPreprocess(sample_spec, gene_spec, ctrl_spec=ctrl_spec,
                                                                      normtype=norm_type, ctrl_geneset=ctrl_geneset,
                                                                      prior=count_prior)
    gene_grnas = {gene: [x for x in meta[gene_index[gene], 0]] for gene in gene_index}
    testdata, ctrldata, test_sample_idxs = collateTestControlSamples(data, sample_ids, ctrl_spec)
    sample_ids_without_ctrl = [sample_ids[idx] for idx in test_sample_idxs]

    with_pseudo = n_pseudo > 0 and len(ctrl_geneset) > 0
    prefault = None
    if with_pseudo and n_pseudo >= 2:
        # the result arrays of the screen job (0.63 GB at the Avana shape), touched for the first time on a background thread
        # while the fold-change files are written: a one-shot process would take those page faults inside the inference call
        T, n_real, L = sum(len(v) for v in gene_index.values()), len(gene_index), testdata.shape[1]
        f8 = np.float64
        prefault = _native.PrefaultedArrays(dict(x1=((T,), f8), x2=((T,), f8), w1=((n_real, L), f8), w2=((n_real, L), f8),
                                                 y=((T, L), f8), tau=((T, L), f8), pvals=((n_real, L), f8)))
    x_reference = None
    if sgrna_reference_file:
        x_reference = {'X1': np.array([eval(x_ref[x]['X1']) for x in meta[:, 0]]),
                       'X2': np.array([eval(x_ref[x]['X2']) for x in meta[:, 0]])}
    else:
        writeFoldChanges(outfile_lfc, testdata, ctrldata, meta, sample_ids_without_ctrl)
        writeFoldChanges(outfile_lfc_std, testdata, ctrldata, meta, sample_ids_without_ctrl, write_std=True)

    LOG.info('Running JACKS inference')
    if with_pseudo and n_pseudo >= 2:
        # The reference's two inferJACKS calls (jacks_io.py:425, :430) and its p-value pass (:69-89) as ONE job on the GPU:
        # the screen goes up once, the pseudo-gene E[w] feeds the null densities on the device.  The pseudo-genes are
        # drawn first -- the same draws: inference does not touch the random stream.
        if createPseudoNonessGenes is _createPseudoNonessGenes:
            pseudo_gene_index, pseudo_csr = _drawPseudoGenes(gene_index, ctrl_geneset, n_pseudo)
        else:                                       # replaced by the caller (the tests pin the reference's draws this way)
            pseudo_gene_index, pseudo_csr = createPseudoNonessGenes(gene_index, ctrl_geneset, n_pseudo), None
        LOG.info('Running JACKS inference on %d pseudogenes' % n_pseudo)
        jacks_results, w1_pvals, pseudo = inferJACKSScreen(gene_index, pseudo_gene_index, testdata, ctrldata, fixed_x=x_reference,
                                                           apply_w_hp=apply_w_hp, positive=positive, full_pseudo=(fdr is not None),
                                                           pseudo_csr=pseudo_csr, out=prefault.take())
        # no pseudo-gene is ever written (jacks_io.py:131): the reference's '_pseudo_noness' files hold the header only
        writeJacksWResults(outprefix + '_pseudo_noness', {}, sample_ids_without_ctrl, write_types=['', '_std'], positive=positive)
        jacks_pseudo_results = dict(jacks_results)
        if pseudo is not None:               # --fdr: the Benjamini-Hochberg ranks count the pseudo-genes (jacks_io.py:52-61)
            pnames, pw1, pw2, ppv = pseudo
            for i, gene in enumerate(pnames):
                jacks_pseudo_results[gene] = (None, None, None, None, pw1[i], pw2[i])
                w1_pvals[gene] = ppv[i]
    else:
        jacks_results = inferJACKS(gene_index, testdata, ctrldata, apply_w_hp=apply_w_hp, fixed_x=x_reference)
        w1_pvals = None
        if with_pseudo:
            LOG.info('Running JACKS inference on %d pseudogenes' % n_pseudo)
            pseudo_gene_index = createPseudoNonessGenes(gene_index, ctrl_geneset, n_pseudo)
            jacks_pseudo_results = inferJACKS(pseudo_gene_index, testdata, ctrldata, apply_w_hp=apply_w_hp)
            writeJacksWResults(outprefix + '_pseudo_noness', jacks_pseudo_results, sample_ids_without_ctrl,
                               write_types=['', '_std'], positive=positive)
            for gene in jacks_results:
                jacks_pseudo_results[gene] = jacks_results[gene]

    LOG.info('Writing JACKS results')
    if with_pseudo:
        writeJacksWResults(outprefix, jacks_pseudo_results, sample_ids_without_ctrl,
                           ctrl_geneset=set(x for x in jacks_pseudo_results if 'JACKS_PSEUDO_GENE' in x),
                           write_types=['', '_std', '_pval'], fdr=fdr, pseudo=True, fdr_thresh_type=fdr_thresh_type,
                           positive=positive, w1_pvals=w1_pvals)
    else:
        writeJacksWResults(outprefix, jacks_results, sample_ids_without_ctrl, ctrl_geneset=ctrl_geneset,
                           write_types=['', '_std'], positive=positive)
    writeJacksXResults(outfile_x, jacks_results, gene_grnas)
    pickleJacksFullResults(outfile_pickle, jacks_results, sample_ids_without_ctrl, gene_grnas)


def runJACKS(countfile, replicatefile, guidemappingfile, rep_hdr=REP_HDR_DEFAULT, sample_hdr=SAMPLE_HDR_DEFAULT,
             common_ctrl_sample=COMMON_CTRL_SAMPLE_DEFAULT, ctrl_sample_hdr=None, sgrna_hdr=SGRNA_HDR_DEFAULT,
             gene_hdr=GENE_HDR_DEFAULT, apply_w_hp=APPLY_W_HP_DEFAULT, norm_type=NORM_TYPE_DEFAULT, ignore_blank_genes=False,
             ctrl_genes=None, fdr=None, fdr_thresh_type='REGULAR', outprefix=OUTPREFIX_DEFAULT, reffile=None, n_pseudo=0,
             count_prior=32, positive=False):
    """The whole pipeline from three file names (jacks_io.py:430-442)."""
    _native.warm_context(0)        # CUDA initialisation (~1 s) overlaps the parsing of the input files
    sample_spec, ctrl_spec, gene_spec, x_ref = preprocess(countfile, replicatefile, guidemappingfile, rep_hdr, sample_hdr,
                                                          common_ctrl_sample, ctrl_sample_hdr, sgrna_hdr, gene_hdr,
                                                          ignore_blank_genes, outprefix, reffile)
    load_data_and_run(sample_spec, gene_spec, ctrl_spec, reffile, x_ref, outprefix, apply_w_hp=apply_w_hp,
                      norm_type=norm_type, ctrl_genes=ctrl_genes, fdr=fdr, fdr_thresh_type=fdr_thresh_type,
                      n_pseudo=n_pseudo, count_prior=count_prior, positive=positive)


def runJACKSFromArgs():
    """Entry point of run_JACKS.py (jacks_io.py:444-453)."""
    args = getJacksParser().parse_args()
    runJACKS(args.countfile, args.replicatefile, args.guidemappingfile, args.rep_hdr, args.sample_hdr,
             args.common_ctrl_sample, args.ctrl_sample_hdr, args.sgrna_hdr, args.gene_hdr, args.apply_w_hp, args.norm_type,
             args.ignore_blank_genes, args.ctrl_genes, args.fdr, args.fdr_thresh_type, args.outprefix, args.reffile,
             args.n_pseudo, args.count_prior, args.positive)
