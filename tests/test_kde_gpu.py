"""GPU parity of the KDE p-value kernel (C-ABI jacks_kde_pvalues_host) against the reference's own numbers
(tests/golden/pvals.npz, produced by jacks_io.computeW1PvalsAndFDRs + scipy 1.18.1) and against the oracle
on seeded inputs.  Tolerance 1e-10 relative (summation order and a <=5 ulp normal CDF; far tails amplify
1 ulp of the bandwidth by z^2), NaN patterns identical."""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


def rel_close(a, b, rtol, atol=0.0):
    """Worst relative error over entries that differ by more than atol.  scipy forms positive-selection
    p-values as sum(1 - Phi(z)), accurate to ~1e-17 ABSOLUTE only, so that case is compared with atol=1e-15."""
    nan = np.isnan(a)
    assert (nan == np.isnan(b)).all()
    m = ~nan & (np.abs(a - b) > atol)
    err = np.abs(a[m] - b[m]) / np.maximum(np.abs(b[m]), 1e-300)
    return float(err.max()) if m.any() else 0.0


@pytest.fixture(scope='module')
def nat():
    from jacks_b200 import _native
    return _native


def test_pvalues_vs_reference_golden(nat):
    z = load_golden('pvals.npz')
    all_w1 = np.concatenate([z['pseudo_w1'], z['gene_w1']])
    ctx = nat.context(0)
    for key, positive in (('pv_neg', False), ('pv_pos', True)):
        pv = nat.kde_pvalues_host(ctx, all_w1, z['pseudo_w1'], positive=positive)
        assert rel_close(pv, z[key], 1e-10, atol=1e-15 if positive else 0.0) < 1e-10, key
    # far tails keep full relative accuracy: p ~ 1e-176 in the fixture
    assert np.nanmin(z['pv_neg']) < 1e-150


@pytest.mark.parametrize('n_pseudo,L,n_genes', [(2, 1, 3), (37, 5, 50), (2000, 7, 300), (5000, 3, 1025)])
def test_pvalues_vs_oracle(nat, n_pseudo, L, n_genes):
    from oracle import jacks_oracle as O
    rng = np.random.default_rng(n_pseudo + L)
    pseudo = rng.normal(0.02, 0.25, size=(n_pseudo, L)) * rng.uniform(0.5, 2.0, size=(1, L))
    genes = np.concatenate([rng.normal(-0.5, 0.9, size=(n_genes - 2, L)), np.full((1, L), -40.0), np.full((1, L), 12.0)])
    if n_pseudo > 10:
        pseudo[rng.integers(0, n_pseudo, size=3), rng.integers(0, L, size=3)] = np.nan
        genes[1, 0] = np.nan
    ctx = nat.context(0)
    for positive in (False, True):
        pv = nat.kde_pvalues_host(ctx, genes, pseudo, positive=positive)
        ref = O.kde_pvalues(genes, pseudo, positive=positive)
        assert rel_close(pv, ref, 1e-10, atol=1e-15 if positive else 0.0) < 1e-10


@pytest.mark.parametrize('n_pseudo', [3000, 40000])
def test_heavy_tailed_null_vs_oracle(nat, n_pseudo):
    """Nulls whose range needs more cells than the table holds (a few pseudo-genes hundreds of bandwidths from the rest):
    the table covers a window around the mean, the values outside it are added term by term.  Lines: ordinary, far
    values below, above, on both sides, one value at the very edge of the window, and a gross outlier that inflates the
    bandwidth itself.  Genes sit everywhere: below all null values, between the far values and the bulk, in the bulk,
    above everything.  Against the numpy restatement of gaussian_kde.integrate_box_1d (1e-10), and against the same
    library forced onto its O(n) sums."""
    import os
    from oracle import jacks_oracle as O
    rng = np.random.default_rng(n_pseudo)
    L = 6
    pseudo = rng.normal(0.0, 0.02, size=(n_pseudo, L))
    pseudo[rng.integers(0, n_pseudo, size=4), 1] = np.array([-3.0, -2.5, -2.9, -7.0])           # far below
    pseudo[rng.integers(0, n_pseudo, size=3), 2] = np.array([4.0, 2.0, 2.0])                     # far above (a tie)
    pseudo[rng.integers(0, n_pseudo, size=6), 3] = np.array([-5.0, -1.0, 6.0, 1.5, -1.0, 3.0])   # both sides
    pseudo[7, 4] = -60.0                                                                           # inflates sigma and h
    pseudo[11, 5] = 2.0
    pseudo[13, 5] = np.nan
    grid = np.concatenate([np.linspace(-8.0, 7.0, 61), np.linspace(-0.1, 0.1, 41), [-70.0, -59.9, 25.0]])
    genes = np.repeat(grid[:, None], L, axis=1) + rng.normal(0, 1e-3, size=(len(grid), L))
    ctx = nat.context(0)
    for positive in (False, True):
        pv = nat.kde_pvalues_host(ctx, genes, pseudo, positive=positive)
        ref = O.kde_pvalues(genes, pseudo, positive=positive)
        assert rel_close(pv, ref, 1e-10, atol=1e-15 if positive else 0.0) < 1e-10
        os.environ['JACKS_KDE_BRUTE'] = '1'
        try:
            brute = nat.kde_pvalues_host(ctx, genes, pseudo, positive=positive)
        finally:
            del os.environ['JACKS_KDE_BRUTE']
        assert rel_close(pv, brute, 1e-10, atol=1e-15 if positive else 0.0) < 1e-10


def test_pvalue_properties_full_size(nat):
    """BASELINE size for one line block: 18,000 genes against 100,000 pseudo-genes -- monotone in w, in [0,1],
    negative + positive selection sum to the mass between -100 and 100 (= 1)."""
    rng = np.random.default_rng(9)
    L = 4
    pseudo = rng.normal(0.0, 0.2, size=(100000, L))
    genes = np.sort(rng.normal(-0.3, 0.7, size=(18000, L)), axis=0)
    ctx = nat.context(0)
    neg = nat.kde_pvalues_host(ctx, genes, pseudo, positive=False)
    pos = nat.kde_pvalues_host(ctx, genes, pseudo, positive=True)
    assert (neg >= 0).all() and (neg <= 1).all()
    assert (np.diff(neg, axis=0) >= -1e-15).all()
    assert np.abs(neg + pos - 1.0).max() < 1e-12
    # spot check 8 entries against scipy
    from oracle import jacks_oracle as O
    pick = np.array([0, 17, 4000, 9000, 12000, 15000, 17990, 17999])
    ref = O.kde_pvalues(genes[pick], pseudo, positive=False)
    assert rel_close(neg[pick], ref, 1e-10) < 1e-10


def test_leave_one_out_and_local_fdr_vs_reference(nat):
    """computeW1PvalsAndFDRs with pseudo=False (control genes scored under a kernel that leaves them out) and
    compute_fdr=True (density ratio), against the reference's own numbers."""
    from jacks_b200 import jacks_io
    z = load_golden('pvals.npz')
    lines = ['L0', 'L1', 'L2']
    names = [str(g) for g in z['real_order']]
    gene_w1 = z['gene_w1']
    real = {g: (None, None, None, None, gene_w1[i], gene_w1[i] ** 2) for i, g in enumerate(names)}
    noness = set(str(g) for g in z['loo_noness'])
    pv, fdr = jacks_io.computeW1PvalsAndFDRs(real, lines, noness_genes=noness, pseudo=False, compute_fdr=True)
    got_pv = np.stack([pv[g] for g in names]); got_fdr = np.stack([fdr[g] for g in names])
    assert rel_close(got_pv, z['pv_loo'], 1e-9) < 1e-9
    ok = np.isclose(got_fdr, z['fdr_loo'], rtol=1e-9, atol=0) | (np.isnan(got_fdr) & np.isnan(z['fdr_loo']))
    assert ok.all(), np.argwhere(~ok)[:5]
    pv_pos, _ = jacks_io.computeW1PvalsAndFDRs(real, lines, noness_genes=noness, pseudo=False, positive=True)
    assert rel_close(np.stack([pv_pos[g] for g in names]), z['pv_loo_pos'], 1e-9, atol=1e-15) < 1e-9
    # pseudo=True with the local fdr
    order = [str(g) for g in z['finite_order']]
    all_names = [str(g) for g in z['order']]
    all_w1 = np.concatenate([z['pseudo_w1'], gene_w1])
    results = {g: (None, None, None, None, all_w1[all_names.index(g)], None) for g in order}
    pset = set(g for g in order if 'JACKS_PSEUDO_GENE' in g)
    _, fdr = jacks_io.computeW1PvalsAndFDRs(results, lines, noness_genes=pset, pseudo=True, compute_fdr=True)
    got = np.stack([fdr[g] for g in order])
    ok = np.isclose(got, z['fdr_pseudo'], rtol=1e-9, atol=0) | (np.isnan(got) & np.isnan(z['fdr_pseudo'])) | \
        ((got == 0) & (z['fdr_pseudo'] == 0))
    assert ok.all(), np.argwhere(~ok)[:5]


def test_fast_path_matches_brute_force(nat, monkeypatch):
    """The binned Hermite expansion against the O(n) normal-CDF sums of the same library (JACKS_KDE_BRUTE=1),
    including far tails on both sides, outliers in the null sample and both selection directions."""
    rng = np.random.default_rng(31)
    L = 3
    pseudo = rng.normal(0.0, 0.2, size=(30000, L))
    pseudo[:4, 0] = [-3.0, 2.5, 9.0, -9.0]                       # outliers stretch the cell table
    genes = np.concatenate([rng.normal(-0.3, 0.8, size=(4000, L)), np.linspace(-12, 12, 400)[:, None] * np.ones((1, L)),
                            pseudo[:50] + 1e-9])
    ctx = nat.context(0)
    for positive in (False, True):
        monkeypatch.delenv('JACKS_KDE_BRUTE', raising=False)
        fast = nat.kde_pvalues_host(ctx, genes, pseudo, positive=positive)
        monkeypatch.setenv('JACKS_KDE_BRUTE', '1')
        brute = nat.kde_pvalues_host(ctx, genes, pseudo, positive=positive)
        monkeypatch.delenv('JACKS_KDE_BRUTE', raising=False)
        assert rel_close(fast, brute, 1e-11, atol=1e-300) < 1e-11


def test_pvalues_reproducible_and_block_independent():
    """Bit-reproducible from call to call, and a gene's p-value does not depend on which other genes are evaluated in
    the same call (what lets dist.pvalues_sharded shard by gene with bit-identical results)."""
    from jacks_b200 import _native as nat
    rng = np.random.default_rng(12)
    genes = rng.normal(-0.3, 0.7, size=(700, 7))
    pseudo = rng.normal(0.0, 0.2, size=(30000, 7))
    pseudo[rng.integers(0, 30000, 50), rng.integers(0, 7, 50)] = np.nan
    ctx = nat.context(0)
    a = nat.kde_pvalues_host(ctx, genes, pseudo)
    b = nat.kde_pvalues_host(ctx, genes, pseudo)
    assert np.array_equal(a, b)
    parts = np.concatenate([nat.kde_pvalues_host(ctx, genes[:211], pseudo), nat.kde_pvalues_host(ctx, genes[211:], pseudo)])
    assert np.array_equal(a, parts)
    p = nat.kde_pvalues_host(ctx, genes, pseudo, positive=True)
    assert np.array_equal(p, nat.kde_pvalues_host(ctx, genes, pseudo, positive=True))
