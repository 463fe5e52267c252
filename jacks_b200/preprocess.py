"""Drop-in for the reference module ``jacks.preprocess`` (jacks/jacks/preprocess.py).

Same public names, arguments, defaults and return layout.  The numerical work -- log2(count + prior),
per-column centring and the mean-variance error model -- runs on the GPU through the C-ABI
(jacks_normalize_host, jacks_condense_host); the host keeps what the reference does with Python
objects: reading the count table, filtering guides, the string sort of guide ids, the gene index.

"""
import csv
import io
import random

import numpy as np

from . import _native
from .infer import LOG


def _window(n, windowfracN=100.0):
    """Half-width of the smoothing window, preprocess.py:29,35."""
    return min(max(int(n / windowfracN), 30), 800)


def calc_posterior_sd(data, windowfracN=100.0, do_monotonize=True, guideset_indexs=set()):
    """Posterior standard deviation per guide from the mean-variance relationship (preprocess.py:24-58):
    the variance of each guide is replaced by the average variance of the ~2*window guides with the
    nearest means, made non-increasing in the mean.  data: [N, R] normalised log counts of one sample."""
    data = np.asarray(data, dtype=np.float64)
    N = data.shape[0]
    if data.ndim == 1 or data.shape[1] == 1:                     # preprocess.py:32
        return np.zeros(N) * np.nan
    if type(guideset_indexs) == list:
        guideset_indexs = set(guideset_indexs)
    if len(guideset_indexs) not in (0, N):
        # fit on the subset, interpolate to every guide (preprocess.py:44-52)
        mask = np.zeros(N, dtype=np.uint8)
        mask[list(guideset_indexs)] = 1
        return _native.posterior_sd_subset_host(_native.context(0), data, mask, _window(len(guideset_indexs), windowfracN),
                                                monotonize=do_monotonize)
    out = _native.condense_host(_native.context(0), data, [list(range(data.shape[1]))],
                                window=_window(N, windowfracN), monotonize=do_monotonize)
    return np.ascontiguousarray(out[:, 0, 1])


def condense_normalised_counts(counts, Iscreens, sample_ids):
    """Per-sample mean and posterior sd over the replicate columns (preprocess.py:60-67) -> [N, S, 2]."""
    return _native.condense_host(_native.context(0), counts, [list(c) for c in Iscreens])


def compute_extras(meta, sample_ids):
    """Sorted unique gene names and {gene: [row indexes]} in row order (preprocess.py:69-74)."""
    genes = np.unique(sorted(meta[:, 1]))
    gene_guides = {g: [] for g in genes}
    for i, g in enumerate(meta[:, 1]):
        gene_guides[g].append(i)
    return genes, gene_guides


def normalizeLogCounts(logcounts, normtype='median', ctrl_guide_indexes=[]):
    """Per-column normalisation of log counts (preprocess.py:76-96).  Returns the normalised array."""
    LOG.info('Applying %s normalisation' % normtype)
    caller = logcounts
    logcounts = np.asarray(logcounts, dtype=np.float64)
    if normtype in ('median', 'zmad', 'mode'):
        out = _native.normalize_host(_native.context(0), logcounts, normtype)
    elif normtype == 'ctrl_guides':
        if len(ctrl_guide_indexes) == 0:
            raise Exception('No guides specified for ctrl guide normalization')
        mask = np.zeros(logcounts.shape[0], dtype=np.uint8)
        mask[list(ctrl_guide_indexes)] = 1
        out = _native.normalize_host(_native.context(0), logcounts, 'ctrl_guides', row_mask=mask)
    else:
        raise Exception('Unrecognised normalisation type %s' % normtype)
    # the reference normalises IN PLACE for median, mode and ctrl_guides (logcounts -= ..., preprocess.py:81,91,94) and
    # returns the same object; zmad rebinds a new array (preprocess.py:84).  Keep that observable behaviour.
    if normtype != 'zmad' and isinstance(caller, np.ndarray) and caller.dtype == np.float64 and caller.flags.writeable:
        caller[...] = out
        return caller
    return out


def inferMissingVariances(data, meta, sample_ids, ctrl_spec, ctrl_geneset):
    """Fill undefined variances of single-replicate test samples from the sample-vs-control spread of the
    control-gene guides (preprocess.py:98-119)."""
    for s, sample_id in enumerate(sample_ids):
        if np.count_nonzero(~np.isnan(data[:, s, 1])):
            continue                 # the sample has defined variances
        if sample_id in ctrl_spec and ctrl_spec[sample_id] == sample_id:
            continue                 # a control: JACKS uses the data variance for it at model time
        if sample_id in ctrl_spec and len(ctrl_geneset) > 0:
            nan_flags = np.isnan(data[:, s, 1])
            guideset_indexs = [i for i, x in enumerate(meta[:, 1]) if x in ctrl_geneset]
            ctrl_data = data[:, [sample_ids.index(ctrl_spec[sample_id])], 0]
            concat_data = np.concatenate((ctrl_data, data[:, [s], 0]), axis=1)
            data[nan_flags, s, 1] = 2 * calc_posterior_sd(concat_data, guideset_indexs=guideset_indexs)[nan_flags]
        else:
            LOG.warning('Undefined variances in sample %s, set --ctrl_genes input to JACKS to infer variances from '
                        'control genes' % sample_id)
    return data


def collateTestControlSamples(data, sample_ids, ctrl_spec):
    """Split the [N, S, 2] cube into test samples and, per test sample, its control (preprocess.py:121-126)."""
    test_sample_idxs = [i for i, x in enumerate(sample_ids) if ctrl_spec[x] != x]
    LOG.info('Collating %d samples' % len(test_sample_idxs))
    testdata = np.take(data, test_sample_idxs, axis=1)          # the values of data[:, idxs, :], C-contiguous (fancy indexing promises no layout)
    ctrl_idxs = [sample_ids.index(ctrl_spec[sample_ids[idx]]) for idx in test_sample_idxs]
    if len(ctrl_idxs) > 1 and len(set(ctrl_idxs)) == 1:
        # one control sample for every line (--common_ctrl_sample): the same [N, L, 2] values as the reference's copy,
        # as a read-only broadcast view of the one column -- inferJACKS sends such a cube to the GPU as [N, 1, 2]
        ctrldata = np.broadcast_to(data[:, ctrl_idxs[:1], :], testdata.shape)
    else:
        ctrldata = data[:, ctrl_idxs, :]
    return testdata, ctrldata, test_sample_idxs


def _parse_cell(text):
    """The reference eval()s every cell (preprocess.py:147); plain numbers take the fast path."""
    try:
        return float(int(text))
    except ValueError:
        try:
            return float(text)
        except ValueError:
            return float(eval(text))


def _read_counts_python(filename, delim, colnames, gene_spec):
    """The reference's reader (preprocess.py:139-149): csv.DictReader, every cell eval()ed."""
    raw, meta = [], []
    with io.open(filename) as f:
        rdr = csv.DictReader(f, delimiter=delim)
        sgrna_col = rdr.fieldnames[0]                 # the guide label is the first column
        for row in rdr:
            guide = row[sgrna_col]
            if guide not in gene_spec:
                continue                              # no gene mapping: ignored
            raw.append([_parse_cell(row[c]) for c in colnames])
            meta.append([guide, gene_spec[guide]])
    return raw, meta


def _read_counts_native(filename, delim, colnames, gene_spec):
    """Same result through the native table parser (jacks_table_read); returns (None, None) when the file needs the
    csv module after all (a multi-line header, cells that are not plain numbers, duplicate column names)."""
    with io.open(filename) as f:
        header_line = f.readline()
        fieldnames = next(csv.reader([header_line], delimiter=delim), None)
    if not fieldnames or header_line.count('"') % 2 or len(set(fieldnames)) != len(fieldnames):
        return None, None
    try:
        sel = [fieldnames.index(c) for c in colnames]
    except ValueError:
        return None, None                             # let DictReader raise the reference's KeyError
    labels, values, not_numeric = _native.table_read(filename, delim, sel, skip_lines=1)
    if any(g is None or '"' in g for g in labels):
        return None, None                             # quoted guide ids: csv unescapes them, the raw field bytes do not
    keep = np.fromiter((g in gene_spec for g in labels), dtype=bool, count=len(labels))
    if not_numeric[keep].any():
        return None, None
    meta = [[g, gene_spec[g]] for g, k in zip(labels, keep) if k]
    return values[keep], meta


def loadDataAndPreprocess(sample_spec, gene_spec, ctrl_spec={}, ctrl_geneset=set(), normtype='median', prior=32):
    """Load counts from one or more files with the same guides in the same order and condense them
    (preprocess.py:128-173).  sample_spec {filename: [(sample_id, colname)]}.
    Returns data [N, S, 2], meta [N, 2] (guide, gene), sample_ids, genes, gene_guides."""
    if normtype not in ('median', 'zmad', 'ctrl_guides', 'mode'):
        raise Exception('Unrecognised normalisation type %s' % normtype)
    ctx = _native.context(0)
    samples, sample_counts, metas = [], [], []
    single = (len(sample_spec) == 1)          # one count table (the command line): the whole chain in one GPU call
    data = None
    for filename in sample_spec:
        delim = ',' if (filename.split('.')[-1] == 'csv') else '\t'
        samples.extend([sample_id for sample_id, colname in sample_spec[filename]])
        colnames = [colname for sample_id, colname in sample_spec[filename]]
        raw, meta = _read_counts_native(filename, delim, colnames, gene_spec)
        if raw is None:
            raw, meta = _read_counts_python(filename, delim, colnames, gene_spec)
        raw = np.array(raw, dtype=np.float64).reshape(len(meta), len(colnames))
        mask = None
        if normtype == 'ctrl_guides':
            idx = [i for i, m in enumerate(meta) if m[1] in ctrl_geneset]
            if len(idx) == 0:
                raise Exception('No guides specified for ctrl guide normalization')
            mask = np.zeros(len(meta), dtype=np.uint8)
            mask[idx] = 1
        LOG.info('Applying %s normalisation' % normtype)
        meta = np.array(meta)
        order = np.argsort(meta[:, 0])
        if single and len(meta) > 0:
            # log2 + normalisation + row sort + condense on the device: the counts go up once, the cube comes down once
            sample_ids = sorted(set(samples))
            Iscreens = [[col_idx for col_idx, x in enumerate(samples) if x == sample_id] for sample_id in sample_ids]
            data, _ = _native.preprocess_host(ctx, raw, Iscreens, normtype=normtype, prior=prior, row_mask=mask, row_order=order)
            meta = meta[order, :]
            metas.append(meta)
            break
        counts = _native.normalize_host(ctx, raw, normtype, take_log=True, prior=prior, row_mask=mask)
        meta, counts = meta[order, :], counts[order, :]
        metas.append(meta)
        sample_counts.append(counts + 0.0)
        if len(metas) > 1 and not (metas[0] == meta).all():
            raise Exception('Incompatible gRNAs in file %s e.g gRNA naming or order does not match other files' % filename)

    sample_ids = sorted(set(samples))
    if data is None:
        counts = np.concatenate(sample_counts, axis=1)
        Iscreens = [[col_idx for col_idx, x in enumerate(samples) if x == sample_id] for sample_id in sample_ids]
        data = condense_normalised_counts(counts, Iscreens, sample_ids)
    genes, gene_guides = compute_extras(meta, sample_ids)
    data = inferMissingVariances(data, meta, sample_ids, ctrl_spec, ctrl_geneset)
    return data, meta, sample_ids, genes, gene_guides


def subsample_and_preprocess(selected_screens, sample_spec, gene_spec, ctrl_spec={}, ctrl_geneset=set(), normtype='median',
                             norepl=True):
    """Pick a number of replicates for each selected screen and preprocess those (preprocess.py:175-227).
    selected_screens: [(sample_id, n_replicates or -1 for all)]; norepl: sample without replacement."""
    found = {}
    wanted = {sample_id: num_rep for (sample_id, num_rep) in selected_screens}
    for filename in sample_spec:
        for sample_id, colname in sample_spec[filename]:
            if sample_id in wanted:
                found.setdefault(sample_id, []).append((filename, colname))

    chosen_spec, seen = {}, {}
    for sample_id, num_rep in selected_screens:
        if sample_id not in found:
            raise Exception('Could not find data for selected sample %s' % sample_id)
        if num_rep < 0:
            picked = found[sample_id]
        elif norepl:
            picked = random.sample(found[sample_id], num_rep) if num_rep <= len(found[sample_id]) else found[sample_id]
        else:
            picked = [random.choice(found[sample_id]) for i in range(num_rep)]

        ctrl_sample_id = ctrl_spec[sample_id]
        if sample_id in seen:
            seen[sample_id] += 1                      # the same sample again: give the extra copy its own id
            sample_id += ('_%d' % seen[sample_id])
            ctrl_spec[sample_id] = ctrl_sample_id
        elif ctrl_sample_id != sample_id:
            seen[sample_id] = 0
        for filename, colname in picked:
            chosen_spec.setdefault(filename, []).append((sample_id, colname))

    return loadDataAndPreprocess(chosen_spec, gene_spec, ctrl_spec=ctrl_spec, ctrl_geneset=ctrl_geneset, normtype=normtype)
