"""Developer probe: wall-clock profile of the whole CLI path (run_JACKS semantics) on a synthetic count-level screen.
CP_GENES / CP_LINES / CP_PSEUDO set the shape: 18000 / 400 / 100000 is the full Avana shape (802 count columns)."""
import cProfile
import io
import os
import pstats
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import logging
from jacks_b200 import jacks_io, synth
from jacks_b200.infer import LOG
LOG.setLevel(logging.WARNING)

n_genes = int(os.environ.get('CP_GENES', 18000))
n_lines = int(os.environ.get('CP_LINES', 100))
n_pseudo = int(os.environ.get('CP_PSEUDO', 2000))
tmp = tempfile.mkdtemp()
t0 = time.time()
guides, genes, cols, samples, counts = synth.make_count_screen(n_genes=n_genes, n_lines=n_lines)
cfile, rfile, nfile = synth.write_count_screen(os.path.join(tmp, 'cp'), guides, genes, cols, samples, counts, sorted(set(genes))[5::18])
print('inputs written in %.1f s: %d guides x %d columns' % (time.time() - t0, counts.shape[0], counts.shape[1]), flush=True)
pr = cProfile.Profile()
t0 = time.time()
pr.enable()
jacks_io.runJACKS(cfile, rfile, cfile, common_ctrl_sample='pDNA', gene_hdr='Gene', outprefix=os.path.join(tmp, 'out'),
                  ctrl_genes=nfile, n_pseudo=n_pseudo)
pr.disable()
print('runJACKS wall %.2f s' % (time.time() - t0))
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(28)
print('\n'.join(s.getvalue().splitlines()[:60]))
