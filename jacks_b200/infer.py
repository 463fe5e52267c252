"""Drop-in for the reference module ``jacks.infer`` (jacks/jacks/infer.py).

Same names, arguments, defaults, return layout and error behaviour:

    inferJACKS(gene_index, testdata, ctrldata, fixed_x=None, n_iter=50, apply_w_hp=False, w_only=False)
        -> {gene: (y, tau, x1, x2, w1, w2)}                                   (infer.py:18-30)
    inferJACKSGene(data, data_err, ctrl, ctrl_err, n_iter, tol=0.1, mu0_x=1, var0_x=1.0, mu0_w=0.0,
                   var0_w=1e4, tau_prior_strength=0.5, fixed_x=None, apply_w_hp=False)
        -> (y, tau, x1, x2, w1, w2)                                           (infer.py:55-99)
    LOG                                                                        (infer.py:6-8)

plus the batch form of the paper scripts' per-line loop (run_JACKS_single.py:83-85):

    inferJACKSLines(gene_index, testdata, ctrldata, fixed_x=None, n_iter=50, w_only=False)

All genes of a call go to the GPU in one batch (one gene per warp team, FP64); the host only packs
the gene index into CSR form and slices the flat result arrays into the per-gene tuples.
No numerical work of the EM runs on the host and there is no CPU fallback.
"""
import logging

import numpy as np

from . import _native

logging.basicConfig(level=logging.DEBUG, format='[%(asctime)s] %(name)-5s: %(levelname)-8s %(message)s')
LOG = logging.getLogger("jacks")
LOG.setLevel(logging.DEBUG)

_SHAPE_MSG = 'Expecting matched size between control and test data'      # infer.py:26


def _planes(a):
    """The (mean, sd) planes of a data cube: the reference reads planes 0 and 1 only (infer.py:24), so cubes that carry a
    third, raw plane (writeFoldChanges(write_raw=True), jacks_io.py:153) are accepted."""
    a = np.asarray(a)
    if a.ndim == 3 and a.shape[2] > 2:
        a = a[:, :, :2]
    return a


def _one_control(ctrldata):
    """A control cube whose columns are ONE column repeated without being materialised -- numpy.broadcast_to of an
    [N, 1, 2] block, which is what jacks_b200.preprocess.collateTestControlSamples returns when every line has the same
    control sample -- goes to the GPU as [N, 1, 2]: half the upload, half the reads, identical arithmetic (the matched
    branch of infer.py:66-70 with equal columns)."""
    return ctrldata.ndim == 3 and ctrldata.shape[1] > 1 and ctrldata.strides[1] == 0


def _rows(rows, n_rows):
    """numpy's fancy indexing (infer.py:21) wraps negative row numbers; out-of-range ones raise IndexError."""
    rows = np.asarray(rows, dtype=np.int64)
    if rows.size and (rows.min() < -n_rows or rows.max() >= n_rows):
        raise IndexError('index out of bounds for axis 0 with size %d' % n_rows)
    return np.where(rows < 0, rows + n_rows, rows)


def infer_batched(offsets, rows, testdata, ctrldata, fixed_x=None, n_iter=50, apply_w_hp=False,
                  want=('x1', 'x2', 'w1', 'w2', 'y', 'tau'), device=0, params=None, **hyper):
    """Batched EM on flat arrays (the array-level form of inferJACKS).

    offsets/rows: CSR gene index.  testdata [N,L,2]; ctrldata [N,L,2], or [N,1,2] when one control
    column serves every line.  Returns a dict of flat arrays (x1,x2: [sum G]; w1,w2: [n_genes,L];
    y,tau: [sum G,L]; n_iters: [n_genes]).
    """
    ctx = _native.context(device)
    if params is None:
        params = _native.default_params(n_iter=int(n_iter), apply_w_hp=int(bool(apply_w_hp)), **hyper)
    fx1 = fx2 = None
    if fixed_x is not None:
        fx1, fx2 = np.asarray(fixed_x['X1'], dtype=np.float64), np.asarray(fixed_x['X2'], dtype=np.float64)
    testdata, ctrldata = _planes(testdata), _planes(ctrldata)
    if _one_control(ctrldata):
        ctrldata = ctrldata[:, :1, :]
    return _native.em_batched_host(ctx, offsets, _rows(rows, testdata.shape[0]), testdata, ctrldata, params=params, fixed_x1=fx1,
                                   fixed_x2=fx2, want=tuple(want) + ('n_iters',))


def inferJACKS(gene_index, testdata, ctrldata, fixed_x=None, n_iter=50, apply_w_hp=False, w_only=False):
    """Main function from which to compute JACKS across a set of specified genes (infer.py:10-30).

    gene_index {gene: [row indexes of the gene's gRNAs in testdata]}; testdata, ctrldata
    (num_total_grnas x num_conditions x 2) arrays of (mean, sd).  Returns
    {gene: (y, tau, x1, x2, w1, w2)}; with w_only the first four entries are None.
    """
    results = {}
    if len(gene_index) == 0:
        return results
    testdata, ctrldata = np.asarray(testdata), np.asarray(ctrldata)
    if testdata.shape != ctrldata.shape:                    # infer.py:25-26 (raised on the first gene)
        raise Exception(_SHAPE_MSG)
    names, offsets, rows = _native.as_csr(gene_index)
    want = ('w1', 'w2') if w_only else ('x1', 'x2', 'w1', 'w2', 'y', 'tau')
    r = infer_batched(offsets, rows, testdata, ctrldata, fixed_x=fixed_x, n_iter=n_iter, apply_w_hp=apply_w_hp,
                      want=want)
    w1, w2 = r['w1'], r['w2']
    if w_only:
        for i, gene in enumerate(names):
            results[gene] = (None, None, None, None, w1[i], w2[i])
        return results
    y, tau, x1, x2 = r['y'], r['tau'], r['x1'], r['x2']
    lo = offsets[:-1].tolist()
    hi = offsets[1:].tolist()
    for i, gene in enumerate(names):
        a, b = lo[i], hi[i]
        results[gene] = (y[a:b], tau[a:b], x1[a:b], x2[a:b], w1[i], w2[i])
    return results


def inferJACKSScreen(gene_index, pseudo_gene_index, testdata, ctrldata, fixed_x=None, n_iter=50, apply_w_hp=False,
                     positive=False, device=0, full_pseudo=False, pseudo_csr=None, out=None):
    """The inference job of load_data_and_run in one GPU call (jacks_run_screen_host): what the reference does with

        jacks_results = inferJACKS(gene_index, testdata, ctrldata, apply_w_hp=..., fixed_x=...)        (jacks_io.py:425)
        jacks_pseudo_results = inferJACKS(pseudo_gene_index, testdata, ctrldata, apply_w_hp=...)       (jacks_io.py:430)
        computeW1PvalsAndFDRs(all results, pseudo=True, positive=...)                                  (jacks_io.py:69-89)

    The screen is uploaded once, the pseudo-gene E[w] never leaves the device unless ``full_pseudo`` asks for the
    pseudo-genes' w1, w2 and p-values (only --fdr needs them: the Benjamini-Hochberg ranks count the pseudo-genes,
    jacks_io.py:52-61; no pseudo-gene is ever written to a file, jacks_io.py:131).
    Returns (jacks_results, pvals {gene: [L]}, pseudo) with pseudo = None or (names, w1, w2, pvals)."""
    testdata, ctrldata = np.asarray(testdata), np.asarray(ctrldata)
    if testdata.shape != ctrldata.shape:
        raise Exception(_SHAPE_MSG)
    names, offsets, rows = _native.as_csr(gene_index)
    # pseudo_csr: (names, offsets, rows) of pseudo_gene_index when its maker already holds it (jacks_io._drawPseudoGenes);
    # out: preallocated result arrays {x1, x2, w1, w2, y, tau, pvals} (jacks_io pre-faults them beside its file writes)
    pnames, poffsets, prows = pseudo_csr if pseudo_csr is not None else _native.as_csr(pseudo_gene_index)
    testdata, ctrldata = _planes(testdata), _planes(ctrldata)
    if _one_control(ctrldata):
        ctrldata = ctrldata[:, :1, :]
    N = testdata.shape[0]
    fx1 = fx2 = None
    if fixed_x is not None:
        fx1, fx2 = np.asarray(fixed_x['X1'], dtype=np.float64), np.asarray(fixed_x['X2'], dtype=np.float64)
    params = _native.default_params(n_iter=int(n_iter), apply_w_hp=int(bool(apply_w_hp)))
    r = _native.run_screen_host(_native.context(device), offsets, _rows(rows, N), poffsets, _rows(prows, N), testdata, ctrldata,
                                params=params, fixed_x1=fx1, fixed_x2=fx2, positive=positive,
                                want=('x1', 'x2', 'w1', 'w2', 'y', 'tau', 'pvals'),
                                want_pseudo=('w1', 'w2', 'pvals') if full_pseudo else (), out=out)
    results, pvals = {}, {}
    lo, hi = offsets[:-1].tolist(), offsets[1:].tolist()
    y, tau, x1, x2, w1, w2, pv = r['y'], r['tau'], r['x1'], r['x2'], r['w1'], r['w2'], r['pvals']
    for i, gene in enumerate(names):
        a, b = lo[i], hi[i]
        results[gene] = (y[a:b], tau[a:b], x1[a:b], x2[a:b], w1[i], w2[i])
        pvals[gene] = pv[i]
    pseudo = (pnames, r['pseudo_w1'], r['pseudo_w2'], r['pseudo_pvals']) if full_pseudo else None
    return results, pvals, pseudo


def inferJACKSLines(gene_index, testdata, ctrldata, fixed_x=None, n_iter=50, w_only=False, device=0):
    """Every cell line on its own ("single JACKS"): the result of

        [inferJACKS(gene_index, testdata[:, [ts], :], ctrldata[:, [ts], :], ...) for ts in range(L)]

    (2018_paper_materials/scripts/jacks/run_JACKS_single.py:83-85, combined as in combineSingleResults :8-15) from one
    set of kernel launches -- each (gene, line) pair is an independent one-line EM.  Returns
    {gene: (y [G,L], tau [G,L], x1 [G,L], x2 [G,L], w1 [L], w2 [L])}: column ts is what the ts-th reference call
    returns; with w_only the first four entries are None, as the script's w_only=True calls leave them.
    """
    results = {}
    if len(gene_index) == 0:
        return results
    testdata, ctrldata = np.asarray(testdata), np.asarray(ctrldata)
    if testdata.shape != ctrldata.shape:                    # infer.py:25-26
        raise Exception(_SHAPE_MSG)
    names, offsets, rows = _native.as_csr(gene_index)
    fx1 = fx2 = None
    if fixed_x is not None:
        fx1, fx2 = np.asarray(fixed_x['X1'], dtype=np.float64), np.asarray(fixed_x['X2'], dtype=np.float64)
    want = ('w1', 'w2', 'n_iters') if w_only else ('x1', 'x2', 'w1', 'w2', 'y', 'tau', 'n_iters')
    testdata, ctrldata = _planes(testdata), _planes(ctrldata)
    rows = _rows(rows, testdata.shape[0])
    r = _native.em_lines_host(_native.context(device), offsets, rows, testdata, ctrldata,
                              params=_native.default_params(n_iter=int(n_iter)), fixed_x1=fx1, fixed_x2=fx2, want=want)
    w1, w2 = r['w1'], r['w2']
    lo, hi = offsets[:-1].tolist(), offsets[1:].tolist()
    for i, gene in enumerate(names):
        if w_only:
            results[gene] = (None, None, None, None, w1[i], w2[i])
        else:
            a, b = lo[i], hi[i]
            results[gene] = (r['y'][a:b], r['tau'][a:b], r['x1'][a:b], r['x2'][a:b], w1[i], w2[i])
    return results


def inferJACKSGene(data, data_err, ctrl, ctrl_err, n_iter, tol=0.1, mu0_x=1, var0_x=1.0, mu0_w=0.0, var0_w=1e4,
                   tau_prior_strength=0.5, fixed_x=None, apply_w_hp=False):
    """Run JACKS inference on a single gene (infer.py:47-99).

    data, data_err: G x L guide effects and their standard deviations; ctrl, ctrl_err: the matched
    G x L control, or a length-G common control (infer.py:61-65).  Like the reference, undefined
    standard deviations are filled in place in the caller's arrays (infer.py:58,63,68).
    """
    data = np.asarray(data, dtype=np.float64)
    ctrl = np.asarray(ctrl, dtype=np.float64)
    G, L = data.shape
    # the reference patches the caller's sd arrays; keep that observable side effect
    data_err[np.isnan(data_err)] = 2.0
    common = (ctrl.shape != data.shape)
    test = np.empty((G, L, 2), dtype=np.float64)
    test[:, :, 0] = data
    test[:, :, 1] = data_err
    if common:
        cube = np.empty((G, 1, 2), dtype=np.float64)
        cube[:, 0, 0] = ctrl.reshape(G)
        cube[:, 0, 1] = np.asarray(ctrl_err, dtype=np.float64).reshape(G)
        gap = np.isnan(ctrl_err)
        if gap.any():
            ctrl_err[gap] = np.nanmean(data_err, axis=1).reshape(gap.shape)[gap]
    else:
        cube = np.empty((G, L, 2), dtype=np.float64)
        cube[:, :, 0] = ctrl
        cube[:, :, 1] = ctrl_err
        gap = np.isnan(ctrl_err)
        if gap.any():
            ctrl_err[gap] = data_err[gap]
    params = _native.default_params(n_iter=int(n_iter), apply_w_hp=int(bool(apply_w_hp)), tol=float(tol),
                                    mu0_x=float(mu0_x), var0_x=float(var0_x), mu0_w=float(mu0_w),
                                    var0_w=float(var0_w), tau_prior_strength=float(tau_prior_strength))
    fx1 = fx2 = None
    if fixed_x is not None:
        fx1, fx2 = np.asarray(fixed_x['X1'], dtype=np.float64), np.asarray(fixed_x['X2'], dtype=np.float64)
    ctx = _native.context(0)
    r = _native.em_batched_host(ctx, np.array([0, G], dtype=np.int64), np.arange(G, dtype=np.int32), test, cube,
                                params=params, fixed_x1=fx1, fixed_x2=fx2, ctrl_rowmean_fill=common,
                                want=('x1', 'x2', 'w1', 'w2', 'y', 'tau', 'n_iters'))
    return r['y'], r['tau'], r['x1'], r['x2'], r['w1'][0], r['w2'][0]
