"""ctypes wrapper of oracle/libem_oracle.so (C restatement of the EM) -- TEST INFRASTRUCTURE ONLY.

Same contract as the product's batched EM (CSR gene index in, flat arrays out) so that tests can diff
the two at full size.  Build with ``python oracle/build_oracle.py`` (gcc).  Never imported by jacks_b200.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, 'libem_oracle.so')


class _Params(C.Structure):
    _fields_ = [('n_iter', C.c_int32), ('apply_w_hp', C.c_int32), ('tol', C.c_double), ('mu0_x', C.c_double),
                ('var0_x', C.c_double), ('mu0_w', C.c_double), ('var0_w', C.c_double), ('tau_prior_strength', C.c_double)]


_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH):
            from . import build_oracle
            build_oracle.build()
        _lib = C.CDLL(_PATH)
        vp = C.c_void_p
        _lib.oracle_em_batched.argtypes = [vp, vp, C.c_int64, C.c_int32, C.c_int32, C.c_int32, vp, vp, C.c_int64, vp, vp,
                                           C.POINTER(_Params)] + [vp] * 8 + [C.c_int32]
    return _lib


def em_batched(offsets, rows, test, ctrl, fixed_x1=None, fixed_x2=None, n_iter=50, apply_w_hp=False, tol=0.1,
               mu0_x=1.0, var0_x=1.0, mu0_w=0.0, var0_w=1e4, tau_prior_strength=0.5, ctrl_rowmean_fill=False,
               want_y_tau=True, n_threads=1):
    lib = _load()
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    rows = np.ascontiguousarray(rows, dtype=np.int32)
    test = np.ascontiguousarray(test, dtype=np.float64)
    ctrl = np.ascontiguousarray(ctrl, dtype=np.float64)
    N, L = test.shape[0], test.shape[1]
    n = len(offsets) - 1
    T = int(offsets[-1]) if n > 0 else 0
    r = dict(x1=np.empty(T), x2=np.empty(T), w1=np.empty((n, L)), w2=np.empty((n, L)),
             n_iters=np.empty(n, dtype=np.int32), bound=np.empty(n))
    if want_y_tau:
        r['y'] = np.empty((T, L)); r['tau'] = np.empty((T, L))
    p = _Params(int(n_iter), int(bool(apply_w_hp)), tol, mu0_x, var0_x, mu0_w, var0_w, tau_prior_strength)
    ptr = lambda a: None if a is None else a.ctypes.data
    if fixed_x1 is not None:
        fixed_x1 = np.ascontiguousarray(fixed_x1, dtype=np.float64)
        fixed_x2 = np.ascontiguousarray(fixed_x2, dtype=np.float64)
    rc = lib.oracle_em_batched(ptr(test), ptr(ctrl), N, L, ctrl.shape[1], int(bool(ctrl_rowmean_fill)), ptr(offsets),
                               ptr(rows), n, ptr(fixed_x1), ptr(fixed_x2), C.byref(p), ptr(r['x1']), ptr(r['x2']),
                               ptr(r['w1']), ptr(r['w2']), ptr(r.get('y')), ptr(r.get('tau')), ptr(r['n_iters']),
                               ptr(r['bound']), int(n_threads))
    if rc != 0:
        raise Exception('Expecting matched size between control and test data')
    return r
