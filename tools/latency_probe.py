"""Developer probe: builds and runs tools/probes/fp64_latency_probe.cu (dependent-issue latencies of DFMA / MUFU.RCP64H, DFMA
rate with three distinct register operands)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
out = os.path.join(HERE, 'probes', '_build')
os.makedirs(out, exist_ok=True)
exe = os.path.join(out, 'fp64_latency_probe')
subprocess.check_call(['/usr/local/cuda/bin/nvcc', '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-o', exe,
                       os.path.join(HERE, 'probes', 'fp64_latency_probe.cu')])
sys.exit(subprocess.call([exe]))
