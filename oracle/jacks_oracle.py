"""CPU oracle for the JACKS hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A numpy restatement of the reference's algorithm (felicityallen/JACKS), written to
reproduce the reference's float64 results bit for bit by issuing the same numpy
primitives in the same association order.  Each function cites the reference
file:line it follows.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module -- as the checker or as the timed
CPU baseline, never on the product path.  ``jacks_b200`` never imports it.

Parity status: PINNED.  ``tests/golden/make_golden.py`` runs the unmodified reference
(imported from /root/reference with the scipy-alias shim of SURVEY.md section 8c) and
stores its outputs under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks
every function here against those files (bit-exact for the EM and the error model,
<=1e-12 relative for KDE p-values against scipy 1.18.1).
"""
import random as _random

import numpy as np

# --------------------------------------------------------------------------------------
# EM  (reference jacks/jacks/infer.py)
# --------------------------------------------------------------------------------------

EM_DEFAULTS = dict(n_iter=50, tol=0.1, mu0_x=1, var0_x=1.0, mu0_w=0.0, var0_w=1e4,
                   tau_prior_strength=0.5)          # infer.py:18, :55


def _rowdot(mat, vec):
    """nansum_l mat[g,l]*vec[l]  -- infer.py:40-42 (matrix . vector branch of nandot)."""
    return np.nansum(np.multiply(mat, np.tile(vec, [mat.shape[0], 1])), axis=1)


def _coldot(vec, mat):
    """nansum_g vec[g]*mat[g,l]  -- infer.py:37-39 (vector . matrix branch of nandot).

    The tiled operand is a transposed view exactly as in the reference, so the product
    array has the same memory order and numpy reduces axis 0 in the same sequence.
    """
    return np.nansum(np.multiply(np.tile(vec, [mat.shape[1], 1]).transpose(), mat), axis=0)


def bound_stat(x1, x2, w1, w2, y, tau):
    """Convergence statistic, infer.py:123-125."""
    xw = np.outer(x1, w1)
    return np.nansum(tau * (y ** 2 + np.outer(x2, w2) - 2 * xw * y))


def step_x(w1, w2, tau, y, mu0_x, var0_x):
    """Posterior update of guide efficacies + median-emphasised renormalisation, infer.py:103-110."""
    ytau = y * tau
    x1 = (mu0_x / var0_x + _rowdot(ytau, w1)) / (_rowdot(tau, w2) + 1.0 / var0_x)
    x2 = x1 ** 2 + 1.0 / (_rowdot(tau, w2) + 1.0 / var0_x)
    wadj = 0.5 / len(x1)
    x1m = x1.mean() + 2 * wadj * np.nanmedian(x1) - wadj * x1.max() - wadj * x1.min()
    return x1 / x1m, x2 / x1m / x1m


def step_w(x1, x2, tau, y, mu0_w, var0_w):
    """Posterior update of per-line gene effects, infer.py:112-116."""
    ytau = y * tau
    w1 = (mu0_w / var0_w + _coldot(x1, ytau)) / (_coldot(x2, tau) + 1.0 / var0_w)
    w2 = w1 ** 2 + 1.0 / (_coldot(x2, tau) + 1.0 / var0_w)
    return w1, w2


def step_tau(x1, x2, w1, w2, y, tau_prior_strength, tau_pr_den):
    """Per-element noise precision, infer.py:118-121 (NaN y propagates to tau)."""
    b_star = y ** 2 - 2 * y * (np.outer(x1, w1)) + np.outer(x2, w2)
    return (tau_prior_strength + 0.5) / (tau_pr_den + 0.5 * b_star)


def _mirror_debug(y, x1, w1):
    # The reference formats its DEBUG messages eagerly (infer.py:87-88,96,109,115), i.e. it
    # evaluates this expression whether or not the message is emitted.  bench.py's CPU
    # baseline switches it on so that the port costs what the reference costs.
    return np.nanmean(abs(y.T - np.outer(w1, x1))).mean()


def em_gene(data, data_err, ctrl, ctrl_err, n_iter=50, tol=0.1, mu0_x=1, var0_x=1.0,
            mu0_w=0.0, var0_w=1e4, tau_prior_strength=0.5, fixed_x=None, apply_w_hp=False,
            info=None, mirror_debug_cost=False):
    """One gene: restatement of inferJACKSGene, infer.py:55-99.

    Unlike the reference this does NOT write into the caller's sd arrays (the reference
    patches NaNs in place, infer.py:58,63,68); the returned values are identical.
    ``info`` (a dict) receives ``n_iters`` (loop passes executed) and ``bound``.
    """
    data = np.asarray(data, dtype=np.float64)
    ctrl = np.asarray(ctrl, dtype=np.float64)
    data_err = np.array(data_err, dtype=np.float64, copy=True)
    ctrl_err = np.array(ctrl_err, dtype=np.float64, copy=True)
    data_err[np.isnan(data_err)] = 2.0                                        # :58
    if ctrl.shape != data.shape:                                              # :61-65 common 1-D control
        gap = np.isnan(ctrl_err)
        ctrl_err[gap] = np.nanmean(data_err, axis=1).reshape(gap.shape)[gap]
        y = (data.T - ctrl).T
        tau_pr_den = tau_prior_strength * 1.0 * ((data_err ** 2).T + ctrl_err ** 2 + 1e-2).T
    else:                                                                     # :66-70 matched control
        gap = np.isnan(ctrl_err)
        ctrl_err[gap] = data_err[gap]
        y = data - ctrl
        tau_pr_den = tau_prior_strength * 1.0 * (data_err ** 2 + ctrl_err ** 2 + 1e-2)
    G, L = y.shape
    if fixed_x is None:                                                       # :73-79
        x1 = mu0_x * np.ones(G)
        x2 = x1 ** 2
    else:
        x1, x2 = fixed_x['X1'], fixed_x['X2']
    w1 = np.nanmedian(y, axis=0)                                              # :81
    tau = tau_prior_strength * 1.0 / tau_pr_den                               # :83
    w2 = w1 ** 2
    bound = bound_stat(x1, x2, w1, w2, y, tau)                                # :86
    if mirror_debug_cost:
        np.nanmean(abs(y)).mean(); _mirror_debug(y, x1, w1); x1.mean(); w1.mean()
    iters = 0
    for _ in range(n_iter):                                                   # :89-98
        last = bound
        if fixed_x is None:
            x1, x2 = step_x(w1, w2, tau, y, mu0_x, var0_x)
            if mirror_debug_cost: _mirror_debug(y, x1, w1)
        if apply_w_hp and len(w1) > 1:
            mu0_w, var0_w = w1.mean(), w1.var() * 3 + 1e-4                    # :92
        w1, w2 = step_w(x1, x2, tau, y, mu0_w, var0_w)
        if mirror_debug_cost: _mirror_debug(y, x1, w1)
        tau = step_tau(x1, x2, w1, w2, y, tau_prior_strength, tau_pr_den)
        bound = bound_stat(x1, x2, w1, w2, y, tau)
        if mirror_debug_cost:
            _mirror_debug(y, x1, w1); np.median((x2 - x1 ** 2) ** 0.5); np.median((w2 - w1 ** 2) ** 0.5)
        iters += 1
        if abs(last - bound) < tol:
            break
    if info is not None:
        info['n_iters'] = iters
        info['bound'] = float(bound)
    return y, tau, x1, x2, w1, w2


def em_genes(gene_index, testdata, ctrldata, fixed_x=None, n_iter=50, apply_w_hp=False,
             w_only=False, iters_out=None, mirror_debug_cost=False):
    """Gene loop: restatement of inferJACKS, infer.py:18-30."""
    out = {}
    for gene in gene_index:
        rows = gene_index[gene]
        fx = None if fixed_x is None else {'X1': fixed_x['X1'][rows], 'X2': fixed_x['X2'][rows]}
        if testdata.shape != ctrldata.shape:                                  # :25-26
            raise Exception('Expecting matched size between control and test data')
        info = {}
        y, tau, x1, x2, w1, w2 = em_gene(testdata[rows, :, 0], testdata[rows, :, 1],
                                         ctrldata[rows, :, 0], ctrldata[rows, :, 1], n_iter,
                                         fixed_x=fx, apply_w_hp=apply_w_hp, info=info,
                                         mirror_debug_cost=mirror_debug_cost)
        if iters_out is not None:
            iters_out[gene] = info['n_iters']
        out[gene] = (None, None, None, None, w1, w2) if w_only else (y, tau, x1, x2, w1, w2)
    return out


def em_lines(gene_index, testdata, ctrldata, fixed_x=None, n_iter=50, iters_out=None):
    """Per-line ("single JACKS") inference: restatement of the loop in
    2018_paper_materials/scripts/jacks/run_JACKS_single.py:83-85 (one inferJACKS call per line on the
    [:, [ts], :] slices) followed by combineSingleResults (:8-15), except that every output keeps its
    line axis: {gene: (y [G,L], tau [G,L], x1 [G,L], x2 [G,L], w1 [L], w2 [L])}.
    iters_out, if given, receives {gene: int array [L]}."""
    L = testdata.shape[1]
    per_line, per_iters = [], []
    for ts in range(L):
        it = {}
        per_line.append(em_genes(gene_index, testdata[:, [ts], :], ctrldata[:, [ts], :], fixed_x=fixed_x,
                                 n_iter=n_iter, iters_out=it))
        per_iters.append(it)
    out = {}
    for gene in gene_index:
        cols = [r[gene] for r in per_line]
        out[gene] = (np.concatenate([c[0] for c in cols], axis=1), np.concatenate([c[1] for c in cols], axis=1),
                     np.stack([c[2] for c in cols], axis=1), np.stack([c[3] for c in cols], axis=1),
                     np.array([c[4][0] for c in cols]), np.array([c[5][0] for c in cols]))
        if iters_out is not None:
            iters_out[gene] = np.array([it[gene] for it in per_iters], dtype=np.int32)
    return out


# --------------------------------------------------------------------------------------
# Preprocessing error model  (reference jacks/jacks/preprocess.py)
# --------------------------------------------------------------------------------------

def log_counts(counts, prior=32):
    """log2(count + prior), preprocess.py:147."""
    return np.log2(np.asarray(counts, dtype=np.float64) + prior)


def normalize_log_counts(logcounts, normtype='median', ctrl_guide_indexes=()):
    """Per-column centring, preprocess.py:77-96.  Returns a new array."""
    lc = np.array(logcounts, dtype=np.float64, copy=True)
    G, L = lc.shape
    if normtype == 'median':
        lc -= np.tile(np.nanmedian(lc, axis=0), (G, 1))
    elif normtype == 'zmad':
        lc -= np.tile(np.nanmedian(lc, axis=0), (G, 1))
        lc = lc / np.tile(1.4826 * np.nanmedian(abs(lc), axis=0), (G, 1))
    elif normtype == 'mode':
        for i in range(L):
            hist, edges = np.histogram(lc[:, i], bins=100)
            smooth = 0.1 * hist[:-4] + 0.2 * hist[1:-3] + 0.4 * hist[2:-2] + 0.2 * hist[3:-1] + 0.1 * hist[4:]
            mids = 0.5 * edges[3:-2] + 0.5 * edges[2:-3]
            lc[:, i] -= mids[np.argmax(smooth)]
    elif normtype == 'ctrl_guides':
        if len(ctrl_guide_indexes) == 0:
            raise Exception('No guides specified for ctrl guide normalization')
        lc -= np.tile(np.nanmedian(lc[list(ctrl_guide_indexes), :], axis=0), (G, 1))
    else:
        raise Exception('Unrecognised normalisation type %s' % normtype)
    return lc


def sliding_mean_shifted(x, window):
    """The reference's sliding mean (preprocess.py:16-22), including its shift of one."""
    N = len(x)
    m = np.zeros(N)
    span = 2 * window + 1
    head = np.mean(x[0:span])
    for k in range(window + 1):
        m[k] = head
    for k in range(window + 1, N - window - 1):
        m[k] = m[k - 1] + (x[k + window + 1] - x[k - window]) / span
    tail = np.mean(x[N - span:])
    for k in range(N - window - 1, N):
        m[k] = tail
    return m


def suffix_max_nanreset(x):
    """Reverse running maximum (preprocess.py:8-13): a NaN entry stays NaN and restarts the run."""
    x = np.array(x, dtype=np.float64, copy=True)
    for j in range(len(x) - 2, -1, -1):
        if not np.isnan(x[j]):
            nxt = x[j + 1]
            if not np.isnan(nxt) and nxt > x[j]:
                x[j] = nxt
    return x


def posterior_sd(data, windowfracN=100.0, do_monotonize=True, guideset_indexs=()):
    """Mean-variance model SD per guide, preprocess.py:25-58."""
    data = np.asarray(data, dtype=np.float64)
    N = data.shape[0]
    if data.ndim == 1 or data.shape[1] == 1:                                  # :32
        return np.zeros(N) * np.nan
    subset = set(guideset_indexs)
    if len(subset) == 0:
        subset = set(range(N))
    window = min(max(int(len(subset) / windowfracN), 30), 800)                # :35
    means = data.mean(axis=1)
    vars_ = np.nanstd(data, axis=1) ** 2                                      # :38
    order = sorted(zip(means.tolist(), vars_.tolist(), range(N)))             # :39-40 tuple sort
    s_mean = np.array([t[0] for t in order])
    s_var = np.array([t[1] for t in order])
    s_idx = np.array([t[2] for t in order], dtype=np.int64)
    in_set = np.array([i in subset for i in s_idx.tolist()], dtype=bool)
    m = sliding_mean_shifted(s_var[in_set], window)
    if do_monotonize:
        m = suffix_max_nanreset(m)
    if len(subset) != N:
        m = np.interp(s_mean, s_mean[in_set], m)                              # :50-52
    sd = np.zeros(N)
    sd[s_idx] = m ** 0.5
    return sd


def condense(counts, Iscreens):
    """Per-sample (mean, sd) cube, preprocess.py:61-67."""
    counts = np.asarray(counts, dtype=np.float64)
    data = np.zeros([counts.shape[0], len(Iscreens), 2])
    for s, cols in enumerate(Iscreens):
        data[:, s, 0] = counts[:, cols].mean(axis=1)
        data[:, s, 1] = posterior_sd(counts[:, cols])
    return data


# --------------------------------------------------------------------------------------
# Pseudo-genes and p-values  (reference jacks/jacks/jacks_io.py)
# --------------------------------------------------------------------------------------

def pseudo_gene_index(gene_index, ctrl_geneset, n_pseudo, rng=_random):
    """Random control-guide sets, jacks_io.py:256-264.  Consumes ``rng`` (default: the global
    ``random`` module) draw for draw like the reference, so the same seed and the same
    iteration order of ``ctrl_geneset`` give the same index sets."""
    ctrl_lst = [g for g in ctrl_geneset if g in gene_index]
    all_lst = [g for g in gene_index]
    out = {}
    for i in range(n_pseudo):
        n_guides = len(gene_index[rng.choice(all_lst)])
        out['JACKS_PSEUDO_GENE_%d' % i] = [rng.choice(gene_index[rng.choice(ctrl_lst)])
                                           for _ in range(n_guides)]
    return out


def _ndtr(z):
    from scipy.special import ndtr
    return ndtr(z)


def kde_bandwidth(sample):
    """Gaussian-KDE kernel standard deviation exactly as scipy.stats.gaussian_kde builds it with
    Scott's rule (scipy 1.18.1 stats/_kde.py: weights = 1/n, neff = 1/sum(weights**2),
    covariance = cov(data, ddof=1, aweights=weights) * neff**(-2/5); integrate_box_1d takes its sqrt)."""
    s = np.asarray(sample, dtype=np.float64)
    wts = np.ones(s.size) / s.size
    neff = 1.0 / np.sum(wts ** 2)
    factor = np.power(neff, -1.0 / 5.0)
    cov = np.atleast_2d(np.cov(s, rowvar=1, bias=False, aweights=wts))
    return np.ravel(np.sqrt(cov * factor ** 2))[0], wts


def kde_pvalues(w1_genes, w1_pseudo, positive=False):
    """P-values of gene w1 against the pseudo-gene w1 KDE, one column per cell line.

    Follows computeW1PvalsAndFDRs with pseudo=True (jacks_io.py:69-89): per line, a
    gaussian_kde over the non-NaN pseudo w1 and ``integrate_box_1d(-100, w)`` (negative
    selection) or ``integrate_box_1d(w, 100)`` (positive), which scipy evaluates as
    vecdot(weights, ndtr((hi - s)/h) - ndtr((lo - s)/h)).
    w1_genes: [n_genes, L]; w1_pseudo: [n_pseudo, L] -> [n_genes, L].
    """
    w1_genes = np.asarray(w1_genes, dtype=np.float64)
    w1_pseudo = np.asarray(w1_pseudo, dtype=np.float64)
    n_genes, L = w1_genes.shape
    out = np.zeros((n_genes, L))
    for l in range(L):
        s = w1_pseudo[:, l]
        s = s[~np.isnan(s)]                                                   # :78
        h, wts = kde_bandwidth(s)
        for g in range(n_genes):
            w = w1_genes[g, l]
            lo, hi = (w, 100.0) if positive else (-100.0, w)
            out[g, l] = np.vecdot(wts, _ndtr((hi - s) / h) - _ndtr((lo - s) / h))
    return out


def fdr_gene_sets(pvals_by_gene, fdr):
    """Benjamini-Hochberg sets per line, jacks_io.py:52-61."""
    sets = []
    first = next(iter(pvals_by_gene))
    for l in range(len(pvals_by_gene[first])):
        ranked = sorted((pvals_by_gene[g][l], g) for g in pvals_by_gene)
        below = [i for i, (p, g) in enumerate(ranked) if p < fdr * (i + 1) / len(ranked)]
        k = max(below) if below else -1
        sets.append(set(g for (p, g) in ranked[:k + 1]))
    return sets


def sorted_genes(results, positive=False):
    """Rank genes by (nanmean(w1), name), jacks_io.py:26-31."""
    ranked = [(np.nanmean(results[g][4]), g) for g in results]
    ranked.sort(reverse=bool(positive))
    return ranked
