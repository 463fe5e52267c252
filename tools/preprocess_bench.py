"""Developer probe: device time of the resident preprocessing chain (jacks_normalize_device -> jacks_condense_device) at the
BASELINE shape: 72,000 guides x 802 replicate columns -> 401 samples."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jacks_b200 import _native


def main():
    N = int(os.environ.get('PB_GUIDES', 72000))
    C = int(os.environ.get('PB_COLS', 802))
    rng = np.random.default_rng(2)
    counts = rng.poisson(300.0, size=(N, C)).astype(np.float64)
    ctx = _native.context(0)
    ms = _native.preprocess_device_ms(ctx, counts, [[2 * s, 2 * s + 1] for s in range(C // 2)], prior=32.0)
    print('preprocess chain %d x %d: %.3f ms on the device' % (N, C, ms))


if __name__ == '__main__':
    main()
