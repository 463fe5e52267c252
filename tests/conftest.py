import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.fixture(scope='session')
def golden():
    return load_golden


def edge_case_tags(npz):
    return sorted(set(k.split('.')[0] for k in npz.files))


def small_screen():
    """example-small at the condensed level, from the reference's own preprocessing output."""
    pre = load_golden('small_preproc.npz')
    data = pre['data']
    sample_ids = [str(s) for s in pre['sample_ids']]
    test_idx = pre['test_idx']
    ctrl_col = sample_ids.index('CTRL')
    testdata = np.ascontiguousarray(data[:, test_idx, :])
    ctrldata = np.ascontiguousarray(data[:, [ctrl_col] * len(test_idx), :])
    genes = [str(g) for g in pre['genes']]
    offs, rows = pre['gene_offsets'], pre['gene_rows']
    gene_index = {g: [int(r) for r in rows[offs[i]:offs[i + 1]]] for i, g in enumerate(genes)}
    return gene_index, testdata, ctrldata, pre
