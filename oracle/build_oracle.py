"""gcc build of the C restatement of the EM (oracle/em_oracle.c -> oracle/libem_oracle.so) -- TEST INFRASTRUCTURE.

    python oracle/build_oracle.py [--force]

The reference is pure Python (nothing to compile into oracle/_ref); this builds the repository's own C checker.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'em_oracle.c')
OUT = os.path.join(HERE, 'libem_oracle.so')


def build(force=False):
    if force or not os.path.exists(OUT) or os.path.getmtime(SRC) > os.path.getmtime(OUT):
        subprocess.check_call(['gcc', '-O2', '-fPIC', '-shared', '-fopenmp', '-ffp-contract=off', '-o', OUT, SRC, '-lm'])
    return OUT


if __name__ == '__main__':
    print(build(force='--force' in sys.argv))
