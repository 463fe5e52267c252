#!/usr/bin/env python
"""Command line of the B200-native JACKS engine: same arguments and output files as the reference's
jacks/run_JACKS.py (countfile replicatefile guidemappingfile [--flags], see `python run_JACKS.py -h`)."""
import logging
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from jacks_b200.jacks_io import runJACKSFromArgs   # noqa: E402
from jacks_b200.infer import LOG                    # noqa: E402

LOG.setLevel(logging.INFO)
if __name__ == '__main__':
    runJACKSFromArgs()
