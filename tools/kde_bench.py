"""Developer probe: device-resident time of the KDE p-values (null preparation + evaluation) at the BASELINE shape."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jacks_b200 import _native


def main():
    n_genes = int(os.environ.get('KB_GENES', 18000))
    n_pseudo = int(os.environ.get('KB_PSEUDO', 100000))
    L = int(os.environ.get('KB_LINES', 400))
    dev = torch.device('cuda:0')
    ctx = _native.context(0)
    torch.manual_seed(1)
    g = torch.randn((n_genes, L), dtype=torch.float64, device=dev) * 0.7 - 0.3
    pz = torch.randn((n_pseudo, L), dtype=torch.float64, device=dev) * 0.2
    pv = torch.empty_like(g)
    lib = _native.lib()
    stream = torch.cuda.current_stream().cuda_stream

    def kde():
        _native.check(lib.jacks_kde_pvalues_device(ctx.h, g.data_ptr(), n_genes, pz.data_ptr(), n_pseudo, L, 0, pv.data_ptr(),
                                                   C.c_void_p(stream or None)))
    for _ in range(2):
        kde()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        kde()
    e1.record()
    torch.cuda.synchronize()
    print('kde p-values %d genes x %d lines vs %d pseudo-genes: %.3f ms per call' % (n_genes, L, n_pseudo, e0.elapsed_time(e1) / 5))
    print('checksum %.17g  min %.3g max %.3g' % (float(pv.sum()), float(pv.min()), float(pv.max())))


if __name__ == '__main__':
    main()
