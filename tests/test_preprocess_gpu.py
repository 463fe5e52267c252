"""GPU parity of the preprocessing kernels (C-ABI jacks_lognorm_median_host / jacks_condense_host) against the
reference's own outputs (tests/golden/sd_model.npz, small_preproc.npz -- produced by jacks/jacks/preprocess.py).
Tolerances: log2 is <=1 ulp on both sides but not the same ulp, so normalised log counts agree to 1e-13
absolute; the posterior sd (sort + prefix-sum sliding mean + sqrt) to 1e-9 relative."""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def nat():
    from jacks_b200 import _native
    return _native


def test_lognorm_median_vs_reference(nat):
    z = load_golden('sd_model.npz')
    ctx = nat.context(0)
    for tag in ['n300r2', 'n4000r3', 'n9000r2', 'n100r4']:
        got = nat.lognorm_median_host(ctx, z[tag + '.counts'].astype(np.float64), 32.0)
        ref = z[tag + '.norm_median']
        assert np.abs(got - ref).max() < 1e-13, tag


def test_posterior_sd_vs_reference(nat):
    z = load_golden('sd_model.npz')
    ctx = nat.context(0)
    for tag in ['n300r2', 'n4000r3', 'n9000r2', 'n100r4']:
        norm = z[tag + '.norm_median']
        R = norm.shape[1]
        data = nat.condense_host(ctx, norm, [list(range(R))])
        assert np.array_equal(data[:, 0, 0], norm.mean(axis=1)), tag            # bit-exact means
        ref = z[tag + '.sd']
        err = np.abs(data[:, 0, 1] - ref) / np.maximum(np.abs(ref), 1e-300)
        assert err.max() < 1e-9, (tag, err.max())


def test_preprocess_example_small_vs_reference(nat):
    inp, pre = load_golden('small_inputs.npz'), load_golden('small_preproc.npz')
    ctx = nat.context(0)
    lc = nat.lognorm_median_host(ctx, inp['counts'].astype(np.float64), 32.0)
    order = np.argsort(inp['guides'])
    lc = np.ascontiguousarray(lc[order])
    cols = [str(c) for c in inp['columns']]
    rep2s = dict(zip([str(r) for r in inp['rep_names']], [str(s) for s in inp['rep_samples']]))
    samples = [rep2s[c] for c in cols]
    sample_ids = sorted(set(samples))
    Iscreens = [[i for i, s in enumerate(samples) if s == sid] for sid in sample_ids]
    data = nat.condense_host(ctx, lc, Iscreens)
    ref = pre['data']
    assert np.abs(data[:, :, 0] - ref[:, :, 0]).max() < 1e-13
    err = np.abs(data[:, :, 1] - ref[:, :, 1]) / np.abs(ref[:, :, 1])
    assert err.max() < 1e-9, err.max()


def test_condense_single_replicate_and_mixed(nat):
    rng = np.random.default_rng(3)
    lc = rng.normal(size=(500, 5))
    ctx = nat.context(0)
    data = nat.condense_host(ctx, lc, [[0], [1, 2], [3, 4, 0]])
    from oracle import jacks_oracle as O
    ref = O.condense(lc, [[0], [1, 2], [3, 4, 0]])
    assert np.isnan(data[:, 0, 1]).all() and np.isnan(ref[:, 0, 1]).all()
    assert np.array_equal(data[:, :, 0], ref[:, :, 0])
    err = np.abs(data[:, 1:, 1] - ref[:, 1:, 1]) / np.abs(ref[:, 1:, 1])
    assert err.max() < 1e-9


@pytest.mark.parametrize('N', [31, 45, 61, 62, 63, 2047, 2048, 2049, 70000])
def test_posterior_sd_sizes_vs_oracle(nat, N):
    """Window edge cases (N around 2*window+1, where numpy's slices clip or wrap), the shared-memory / global
    boundary of the sort, and a BASELINE-sized sample (N = 70,000 guides)."""
    from oracle import jacks_oracle as O
    rng = np.random.default_rng(N)
    cnt = rng.poisson(rng.lognormal(np.log(200.0), 1.0, size=(N, 1)) * rng.uniform(0.7, 1.3, size=(1, 2)))
    cnt[rng.random(N) < 0.05] = 0                       # ties on (mean, var): order falls back to the index
    lc = np.log2(cnt + 32.0)
    lc -= np.median(lc, axis=0)
    ctx = nat.context(0)
    data = nat.condense_host(ctx, lc, [[0, 1]])
    ref = O.posterior_sd(lc)
    err = np.abs(data[:, 0, 1] - ref) / np.maximum(np.abs(ref), 1e-300)
    assert err.max() < 1e-9, err.max()


def test_too_few_guides_is_an_error(nat):
    ctx = nat.context(0)
    with pytest.raises(Exception, match='smoothing window'):
        nat.condense_host(ctx, np.zeros((25, 2)), [[0, 1]])


def test_normalisation_types_vs_reference(nat):
    """zmad and mode against the reference's own outputs; ctrl_guides against the oracle (bit-exact inputs:
    the fixtures' counts are integers, so log2 comes from the host table)."""
    from oracle import jacks_oracle as O
    z = load_golden('sd_model.npz')
    ctx = nat.context(0)
    for tag in ['n300r2', 'n4000r3', 'n9000r2', 'n100r4']:
        counts = z[tag + '.counts'].astype(np.float64)
        for nt in ('zmad', 'mode'):
            got = nat.normalize_host(ctx, counts, nt, take_log=True, prior=32.0)
            assert np.abs(got - z[tag + '.norm_' + nt]).max() < 1e-12, (tag, nt)
        lc = np.log2(counts + 32.0)
        got = nat.normalize_host(ctx, lc, 'mode')                       # already-logged input, no table
        assert np.abs(got - z[tag + '.norm_mode']).max() < 1e-12, tag
        rows = list(range(0, counts.shape[0], 7))
        mask = np.zeros(counts.shape[0], dtype=np.uint8); mask[rows] = 1
        got = nat.normalize_host(ctx, counts, 'ctrl_guides', take_log=True, prior=32.0, row_mask=mask)
        assert np.abs(got - O.normalize_log_counts(lc, 'ctrl_guides', rows)).max() < 1e-13, tag


def test_posterior_sd_subset_vs_reference(nat):
    """calc_posterior_sd(guideset_indexs=...) -- the fit inferMissingVariances uses for single-replicate samples."""
    from jacks_b200 import preprocess
    z = load_golden('sd_model.npz')
    for tag in ['n300r2', 'n4000r3', 'n9000r2', 'n100r4']:
        norm = z[tag + '.norm_median']
        got = preprocess.calc_posterior_sd(norm.copy(), guideset_indexs=z[tag + '.subset'].tolist())
        ref = z[tag + '.sd_subset']
        err = np.abs(got - ref) / np.maximum(np.abs(ref), 1e-300)
        assert err.max() < 1e-9, (tag, err.max())
        got = preprocess.calc_posterior_sd(norm.copy())
        err = np.abs(got - z[tag + '.sd']) / np.abs(z[tag + '.sd'])
        assert err.max() < 1e-9, tag


def test_infer_missing_variances_single_replicate(nat):
    """A single-replicate test sample with control genes: sd = 2 * posterior sd of (control mean, sample mean)
    fitted on the control-gene guides (preprocess.py:98-119)."""
    from jacks_b200 import preprocess
    from oracle import jacks_oracle as O
    rng = np.random.default_rng(21)
    N = 1200
    lc = rng.normal(0, 1.2, size=(N, 3))
    data = nat.condense_host(nat.context(0), lc, [[0, 1], [2]])
    assert np.isnan(data[:, 1, 1]).all()
    meta = np.array([['g%04d' % i, 'GENE%03d' % (i // 4)] for i in range(N)])
    ctrl_geneset = set('GENE%03d' % k for k in range(0, 300, 3))
    out = preprocess.inferMissingVariances(data.copy(), meta, ['CTRL', 'S1'], {'CTRL': 'CTRL', 'S1': 'CTRL'}, ctrl_geneset)
    sub = [i for i in range(N) if meta[i, 1] in ctrl_geneset]
    ref = 2 * O.posterior_sd(np.stack([data[:, 0, 0], data[:, 1, 0]], axis=1), guideset_indexs=sub)
    err = np.abs(out[:, 1, 1] - ref) / np.abs(ref)
    assert err.max() < 1e-9, err.max()
    assert np.array_equal(out[:, 0, :], data[:, 0, :])


@pytest.mark.parametrize('case', ['pairs', 'warp_runs', 'cta_runs', 'long_ordered', 'long_unordered', 'poisson300'])
def test_posterior_sd_mean_ties_vs_oracle(nat, case):
    """The sort is ONE stable radix sort by the mean plus a repair of the runs of equal means (by the variance, then the
    index: Python's tuple order, preprocess.py:39-40).  Every repair path: runs a thread sorts itself (<= 16), runs a warp
    or a CTA ranks (<= 256, <= 2048), a longer run that is already in order (nothing to do), a longer run that is not (the
    exact two-key sort takes over), and the bench's Poisson(300) library whose modal count pairs tie ~80 deep."""
    from oracle import jacks_oracle as O
    rng = np.random.default_rng(['pairs', 'warp_runs', 'cta_runs', 'long_ordered', 'long_unordered', 'poisson300'].index(case) + 100)
    N = 6000
    a = rng.normal(0.0, 1.0, size=N)
    d = np.abs(rng.normal(0.0, 0.3, size=N))
    def tie(sel, mean):             # rows `sel` share one mean exactly; their variances stay as drawn
        a[sel] = mean
    if case == 'pairs':
        for k in range(0, 3000, 3):
            tie(slice(k, k + rng.integers(2, 4)), a[k])
    elif case == 'warp_runs':
        for k in range(0, 4000, 400):
            tie(rng.choice(N, size=int(rng.integers(17, 257)), replace=False), float(k) / 1000.0)
    elif case == 'cta_runs':
        tie(rng.choice(N, size=300, replace=False), 0.25)
        tie(rng.choice(N, size=2048, replace=False), -0.5)
    elif case == 'long_ordered':
        sel = np.sort(rng.choice(N, size=3000, replace=False))
        tie(sel, 0.125)
        d[sel] = 0.0                # e.g. zero-count guides: identical mean AND variance, index order stands
    elif case == 'long_unordered':
        tie(rng.choice(N, size=2500, replace=False), 0.125)
    if case == 'poisson300':
        cnt = rng.poisson(300.0, size=(72000, 2)).astype(np.float64)
        lc = np.log2(cnt + 32.0)
        lc -= np.median(lc, axis=0)
    else:
        # two replicates a -/+ d: the mean is a exactly when d is a multiple of a small power of two
        d = np.round(d * 1024.0) / 1024.0
        a = np.round(a * 1024.0) / 1024.0
        lc = np.stack([a - d, a + d], axis=1)
        assert np.array_equal(lc.mean(axis=1), a)
    ctx = nat.context(0)
    data = nat.condense_host(ctx, lc, [[0, 1]])
    ref = O.posterior_sd(lc)
    assert np.array_equal(data[:, 0, 0], lc.mean(axis=1))
    err = np.abs(data[:, 0, 1] - ref) / np.maximum(np.abs(ref), 1e-300)
    assert err.max() < 1e-9, (case, err.max())
