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
