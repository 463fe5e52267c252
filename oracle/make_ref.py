"""Recipe: make the UNMODIFIED reference importable on the GPU box -- TEST / BASELINE INFRASTRUCTURE ONLY.

The reference (felicityallen/JACKS) is pure Python, so there is nothing to compile: this copies the four modules of its
``jacks`` package from where they lie under /root/reference into ``oracle/_ref/jacks/`` -- a directory that is git-ignored
(the reference's sources never enter this repository's history) but travels to the GPU box with the snapshot, like the
built shared libraries.  ``bench.py --impl reference`` and the ``cpu_baseline`` leg import it from there (kind
"reference"); without it they time the numpy restatement (kind "port").  Run in the build container:

    python oracle/make_ref.py

The only shim, applied by the importer and not to the files: ``jacks.infer.SP = numpy`` (infer.py uses scipy's removed
top-level numpy aliases).
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = '/root/reference/jacks/jacks'
DST = os.path.join(HERE, '_ref', 'jacks')
FILES = ('__init__.py', 'infer.py', 'preprocess.py', 'jacks_io.py')


def make(force=False):
    """Returns the directory to put on sys.path, or None when /root/reference is absent and no copy exists."""
    root = os.path.dirname(DST)
    if os.path.isdir(SRC):
        os.makedirs(DST, exist_ok=True)
        for f in FILES:
            s, d = os.path.join(SRC, f), os.path.join(DST, f)
            if force or not os.path.exists(d) or os.path.getmtime(s) > os.path.getmtime(d):
                shutil.copyfile(s, d)
    return root if all(os.path.exists(os.path.join(DST, f)) for f in FILES) else None


def import_reference():
    """(infer, jacks_io) modules of the unmodified reference with the numpy shim applied, or None."""
    root = make()
    if root is None:
        return None
    import logging
    import numpy
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == 'jacks' or k.startswith('jacks.')}
    sys.path.insert(0, root)
    try:
        import jacks.infer as RI
        import jacks.jacks_io as RIO
        assert os.path.abspath(RI.__file__).startswith(os.path.abspath(root)), RI.__file__
    finally:
        sys.path.remove(root)
        ref_mods = {k: sys.modules.pop(k) for k in list(sys.modules) if k == 'jacks' or k.startswith('jacks.')}
        sys.modules.update(saved)
    RI.SP = numpy
    RI.LOG.setLevel(logging.INFO)          # the CLI's level (run_JACKS.py:6): the eager debug arithmetic stays in
    for h in list(logging.getLogger().handlers):
        h.setLevel(logging.WARNING)
    return RI, RIO


if __name__ == '__main__':
    print(make(force='--force' in sys.argv))
