"""ctypes binding of libjacks_b200.so (the C-ABI declared in include/jacks_b200.h).

This is the only door between the Python host code and the CUDA kernels.  There is no
CPU fallback: if the shared library is missing, or no CUDA device is visible, the calls
raise -- they never silently compute on the host.
"""
import ctypes as C
import os
import threading

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('JACKS_B200_LIB') or os.path.join(_PKG, 'libjacks_b200.so')   # env: developer A/B builds

JACKS_OK, JACKS_E_INVALID, JACKS_E_CUDA, JACKS_E_NOMEM, JACKS_E_SHAPE = 0, 1, 2, 3, 4


class JacksError(Exception):
    """Raised for any non-zero return code of the C-ABI (the reference raises plain Exception too)."""

    def __init__(self, code, msg):
        Exception.__init__(self, msg)
        self.code = code


class EmParams(C.Structure):
    _fields_ = [('n_iter', C.c_int32), ('apply_w_hp', C.c_int32), ('tol', C.c_double), ('mu0_x', C.c_double),
                ('var0_x', C.c_double), ('mu0_w', C.c_double), ('var0_w', C.c_double),
                ('tau_prior_strength', C.c_double)]


class EmDeviceIO(C.Structure):
    _fields_ = [('test', C.c_void_p), ('ctrl', C.c_void_p), ('ctrl_lines', C.c_int32),
                ('ctrl_rowmean_fill', C.c_int32), ('fixed_x1', C.c_void_p), ('fixed_x2', C.c_void_p),
                ('x1', C.c_void_p), ('x2', C.c_void_p), ('w1', C.c_void_p), ('w2', C.c_void_p),
                ('y', C.c_void_p), ('tau', C.c_void_p), ('n_iters', C.c_void_p), ('bound', C.c_void_p)]


class EmHostIO(C.Structure):
    _fields_ = [('test', C.c_void_p), ('ctrl', C.c_void_p), ('n_rows', C.c_int64), ('n_lines', C.c_int32),
                ('ctrl_lines', C.c_int32), ('ctrl_rowmean_fill', C.c_int32), ('reserved', C.c_int32),
                ('fixed_x1', C.c_void_p), ('fixed_x2', C.c_void_p), ('x1', C.c_void_p), ('x2', C.c_void_p),
                ('w1', C.c_void_p), ('w2', C.c_void_p), ('y', C.c_void_p), ('tau', C.c_void_p),
                ('n_iters', C.c_void_p), ('bound', C.c_void_p)]


class ScreenDeviceIO(C.Structure):
    _fields_ = [('real', EmDeviceIO), ('pvals', C.c_void_p), ('pseudo_w1', C.c_void_p), ('pseudo_w2', C.c_void_p),
                ('pseudo_n_iters', C.c_void_p), ('pseudo_pvals', C.c_void_p), ('positive', C.c_int32), ('reserved', C.c_int32)]


class ScreenHostIO(C.Structure):
    _fields_ = [('real', EmHostIO), ('pvals', C.c_void_p), ('pseudo_w1', C.c_void_p), ('pseudo_w2', C.c_void_p),
                ('pseudo_n_iters', C.c_void_p), ('pseudo_pvals', C.c_void_p), ('positive', C.c_int32), ('reserved', C.c_int32)]


_lib = None
_lock = threading.Lock()


def lib():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise JacksError(JACKS_E_CUDA, 'jacks_b200: %s is missing; run `python -m jacks_b200.build` '
                             '(needs nvcc). There is no CPU fallback.' % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
        L.jacks_last_error.restype = C.c_char_p
        L.jacks_abi_version.restype = C.c_int
        L.jacks_device_count.restype = C.c_int
        L.jacks_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
        L.jacks_ctx_destroy.argtypes = [vp]
        L.jacks_ctx_destroy.restype = None
        L.jacks_ctx_launch_count.argtypes = [vp]
        L.jacks_ctx_launch_count.restype = i64
        L.jacks_ctx_synchronize.argtypes = [vp]
        L.jacks_em_default_params.argtypes = [C.POINTER(EmParams)]
        L.jacks_em_default_params.restype = None
        L.jacks_host_alloc.argtypes = [C.POINTER(vp), C.c_uint64]
        L.jacks_host_free.argtypes = [vp]
        L.jacks_host_free.restype = None
        L.jacks_host_register.argtypes = [vp, C.c_uint64]
        L.jacks_host_unregister.argtypes = [vp]
        L.jacks_copy_to_host_async.argtypes = [vp, vp, C.c_uint64, vp]
        L.jacks_em_plan_create.argtypes = [vp, vp, vp, i64, i64, i32, C.POINTER(vp)]
        L.jacks_em_plan_destroy.argtypes = [vp]
        L.jacks_em_plan_destroy.restype = None
        L.jacks_em_run_device.argtypes = [vp, vp, C.POINTER(EmParams), C.POINTER(EmDeviceIO), vp]
        L.jacks_em_batched_host.argtypes = [vp, C.POINTER(EmParams), vp, vp, i64, C.POINTER(EmHostIO), i32]
        L.jacks_em_lines_device.argtypes = [vp, vp, C.POINTER(EmParams), C.POINTER(EmDeviceIO), vp]
        L.jacks_em_lines_host.argtypes = [vp, C.POINTER(EmParams), vp, vp, i64, C.POINTER(EmHostIO), i32]
        L.jacks_run_screen_device.argtypes = [vp, vp, vp, C.POINTER(EmParams), C.POINTER(ScreenDeviceIO), vp]
        L.jacks_run_screen_host.argtypes = [vp, C.POINTER(EmParams), vp, vp, i64, vp, vp, i64, C.POINTER(ScreenHostIO), i32]
        L.jacks_fp64_peak_flops.argtypes = [vp, C.POINTER(dbl), C.POINTER(dbl)]
        if hasattr(L, 'jacks_kde_pvalues_device'):
            L.jacks_kde_pvalues_device.argtypes = [vp, vp, i64, vp, i64, i32, i32, vp, vp]
            L.jacks_kde_pvalues_host.argtypes = [vp, vp, i64, vp, i64, i32, i32, vp]
            L.jacks_kde_eval_host.argtypes = [vp, vp, i64, vp, i64, i32, i32, i32, vp]
        if hasattr(L, 'jacks_condense_host'):
            L.jacks_condense_host.argtypes = [vp, vp, i64, i32, vp, vp, i32, i32, i32, vp]
            L.jacks_normalize_host.argtypes = [vp, vp, i64, i32, i32, dbl, vp, i64, i32, vp, vp]
            L.jacks_posterior_sd_subset_host.argtypes = [vp, vp, i64, i32, vp, i32, i32, vp]
            L.jacks_normalize_device.argtypes = [vp, vp, i64, i32, i32, dbl, vp, i64, i32, vp, vp, vp]
            L.jacks_condense_device.argtypes = [vp, vp, i64, i32, vp, vp, i32, i32, i32, vp, vp]
            L.jacks_preprocess_host.argtypes = [vp, vp, i64, i32, i32, dbl, vp, i64, i32, vp, vp, vp, vp, i32, i32, i32, vp, vp]
        L.jacks_table_scan.argtypes = [C.c_char_p, i64, C.POINTER(i64), C.POINTER(i64)]
        L.jacks_table_read.argtypes = [C.c_char_p, i64, C.c_char, i64, i32, vp, vp, vp, i32, vp, vp, vp]
        L.jacks_table_write.argtypes = [C.c_char_p, C.c_char_p, i64, i32, vp, vp, vp, vp, i32]
        L.jacks_pseudo_draw.argtypes = [vp, vp, vp, i64, vp, i64, i64, vp, vp, i64, C.POINTER(i64)]
        _lib = L
    return _lib


def check(rc):
    if rc != JACKS_OK:
        msg = lib().jacks_last_error().decode('utf-8', 'replace')
        raise JacksError(rc, msg or ('jacks_b200 error %d' % rc))


def default_params(**over):
    p = EmParams()
    lib().jacks_em_default_params(C.byref(p))
    for k, v in over.items():
        if not hasattr(p, k):
            raise TypeError('unknown EM parameter %s' % k)
        setattr(p, k, v)
    return p


def _ptr(a):
    return None if a is None else a.ctypes.data


class PinnedArray(object):
    """A numpy array backed by pinned (page-locked) host memory from jacks_host_alloc."""

    def __init__(self, shape, dtype=np.float64):
        self.shape = tuple(int(s) for s in np.atleast_1d(shape))
        self.dtype = np.dtype(dtype)
        nbytes = int(np.prod(self.shape, dtype=np.int64)) * self.dtype.itemsize
        p = C.c_void_p()
        check(lib().jacks_host_alloc(C.byref(p), max(nbytes, 1)))
        self._p = p
        buf = (C.c_char * max(nbytes, 1)).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape, dtype=np.int64))).reshape(self.shape)

    def free(self):
        if self._p is not None:
            self.array = None
            lib().jacks_host_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context(object):
    """One per GPU: streams, scratch and resident buffers (JacksCtx)."""

    def __init__(self, device=0):
        h = C.c_void_p()
        check(lib().jacks_ctx_create(int(device), C.byref(h)))
        self.h = h
        self.device = int(device)

    def close(self):
        if self.h is not None:
            lib().jacks_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launch_count(self):
        return int(lib().jacks_ctx_launch_count(self.h))

    def synchronize(self):
        """Wait for asynchronous host calls (em_batched_host(..., sync=False)) on this context."""
        check(lib().jacks_ctx_synchronize(self.h))

    def fp64_peak_flops(self):
        f, s = C.c_double(), C.c_double()
        check(lib().jacks_fp64_peak_flops(self.h, C.byref(f), C.byref(s)))
        return f.value


_ctx = {}
_ctx_lock = threading.Lock()


def context(device=0):
    """Process-wide context of a device (created on first use)."""
    device = int(device)
    with _ctx_lock:
        if device not in _ctx:
            _ctx[device] = Context(device)
        return _ctx[device]


def warm_context(device=0):
    """Start creating the device context (CUDA initialisation, ~1 s) on a background thread, so that it overlaps host
    work such as parsing the count table; the first context() call waits for it.  Errors surface at that call."""
    def run():
        try:
            context(device)
        except Exception:
            pass
    if int(device) not in _ctx:
        threading.Thread(target=run, name='jacks-b200-ctx', daemon=True).start()


class PrefaultedArrays(object):
    """Result arrays of a host call allocated AND touched on a background thread, so that the first-touch page faults of
    hundreds of megabytes (a one-shot command-line process pays them inside its only inference call) overlap host work
    that precedes the call.  shapes: {key: (shape, dtype)}; take() waits for the thread and returns {key: ndarray}.
    (Page-locking the arrays on the same thread as well measured no faster end to end: cudaHostRegister of 0.6 GB stalls
    the allocations of the call that follows.)"""

    def __init__(self, shapes):
        self._out = {}
        self._shapes = dict(shapes)
        self._thread = threading.Thread(target=self._run, name='jacks-b200-prefault', daemon=True)
        self._thread.start()

    def _run(self):
        for k, (shape, dtype) in self._shapes.items():
            a = np.empty(shape, dtype=dtype)
            a.fill(0)                       # numpy releases the GIL for this loop: the faults are taken beside the main thread
            self._out[k] = a

    def take(self):
        self._thread.join()
        return self._out


def as_csr(gene_index):
    """gene_index {gene: [rows]} -> (names, offsets int64[n+1], rows int32[sum G]) in dict order."""
    names = list(gene_index)
    counts = np.fromiter((len(gene_index[g]) for g in names), dtype=np.int64, count=len(names))
    offsets = np.zeros(len(names) + 1, dtype=np.int64)
    np.cumsum(counts, out=offsets[1:])
    if len(names):
        rows = np.fromiter((r for g in names for r in gene_index[g]), dtype=np.int64, count=int(offsets[-1]))
    else:
        rows = np.zeros(0, dtype=np.int64)
    return names, offsets, rows


def _cube(a):
    """float64 C-contiguous [N, L, 2] form of a data cube.  Planes beyond the first two are ignored, as the reference
    reads only [..., 0] and [..., 1] (infer.py:57-58; writeFoldChanges(write_raw=True) implies [N, L, 3] cubes)."""
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 3 and a.shape[2] > 2:
        a = a[:, :, :2]
    return np.ascontiguousarray(a)


def _rows(rows, n_rows):
    """Guide rows as int32 with numpy's index semantics: negative rows count from the end, anything else out of range
    raises IndexError (the reference indexes its cubes with these, infer.py:43-44)."""
    rows = np.asarray(rows)
    if rows.size:
        if rows.min() < -n_rows or rows.max() >= n_rows:
            raise IndexError('guide row index out of bounds for %d rows' % n_rows)
        if rows.min() < 0:
            rows = np.where(rows < 0, rows + n_rows, rows)
    return np.ascontiguousarray(rows, dtype=np.int32)


class EmPlan(object):
    """Device-resident gene index (JacksEmPlan)."""

    def __init__(self, ctx, offsets, rows, n_rows, n_lines):
        self.ctx = ctx
        self.offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        self.rows = _rows(rows, n_rows)
        self.n_genes = len(self.offsets) - 1
        self.total_guides = int(self.offsets[-1]) if self.n_genes >= 0 and len(self.offsets) else 0
        self.n_lines = int(n_lines)
        h = C.c_void_p()
        check(lib().jacks_em_plan_create(ctx.h, _ptr(self.offsets), _ptr(self.rows), self.n_genes, int(n_rows),
                                         int(n_lines), C.byref(h)))
        self.h = h

    def close(self):
        if self.h is not None:
            lib().jacks_em_plan_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run_device(self, params, io, stream=0):
        """Enqueue the batched EM on `stream` (an int cudaStream_t handle); io is an EmDeviceIO of device pointers."""
        check(lib().jacks_em_run_device(self.ctx.h, self.h, C.byref(params), C.byref(io), C.c_void_p(stream or None)))

    def run_lines_device(self, params, io, stream=0):
        """Enqueue the per-line EM (jacks_em_lines_device): outputs carry a line axis, see em_lines_host."""
        check(lib().jacks_em_lines_device(self.ctx.h, self.h, C.byref(params), C.byref(io), C.c_void_p(stream or None)))


def em_batched_host(ctx, offsets, rows, test, ctrl, params=None, fixed_x1=None, fixed_x2=None,
                    want=('x1', 'x2', 'w1', 'w2', 'y', 'tau', 'n_iters', 'bound'), ctrl_rowmean_fill=False,
                    keep_resident=False, n_rows=None, n_lines=None, ctrl_lines=None, out=None, sync=True,
                    used_rows_only=False):
    """jacks_em_batched_host on numpy arrays.  test [N,L,2], ctrl [N,L,2] or [N,1,2] float64 C-contiguous
    (or test=None to reuse the cubes a previous keep_resident call left on the device, with
    n_rows/n_lines/ctrl_lines given).  used_rows_only: upload just the rows the index references
    (JACKS_UPLOAD_USED_ROWS: a pseudo-gene index over a few thousand control guides moves tens of megabytes instead
    of the screen).  Returns a dict of the requested outputs."""
    params = params or default_params()
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    rows = np.asarray(rows)
    if test is not None:
        test, ctrl = _cube(test), _cube(ctrl)
        if test.ndim != 3 or test.shape[2] != 2 or ctrl.ndim != 3 or ctrl.shape[2] != 2 or ctrl.shape[0] != test.shape[0]:
            raise JacksError(JACKS_E_SHAPE, 'Expecting matched size between control and test data')
        n_rows, n_lines, ctrl_lines = test.shape[0], test.shape[1], ctrl.shape[1]
    rows = _rows(rows, n_rows)
    n_genes = len(offsets) - 1
    T = int(offsets[-1]) if n_genes > 0 else 0
    shapes = dict(x1=((T,), np.float64), x2=((T,), np.float64), w1=((n_genes, n_lines), np.float64),
                  w2=((n_genes, n_lines), np.float64), y=((T, n_lines), np.float64), tau=((T, n_lines), np.float64),
                  n_iters=((n_genes,), np.int32), bound=((n_genes,), np.float64))
    res = {}
    for k in set(want) | {'w1'}:
        if out is not None and k in out:
            assert out[k].shape == shapes[k][0] and out[k].dtype == shapes[k][1] and out[k].flags.c_contiguous
            res[k] = out[k]
        else:
            res[k] = np.empty(shapes[k][0], dtype=shapes[k][1])
    io = EmHostIO()
    io.test, io.ctrl = _ptr(test), _ptr(ctrl)
    io.n_rows, io.n_lines, io.ctrl_lines = int(n_rows), int(n_lines), int(ctrl_lines)
    io.ctrl_rowmean_fill = int(bool(ctrl_rowmean_fill))
    if fixed_x1 is not None:
        fixed_x1 = np.ascontiguousarray(fixed_x1, dtype=np.float64)
        fixed_x2 = np.ascontiguousarray(fixed_x2, dtype=np.float64)
        if fixed_x1.shape != (n_rows,) or fixed_x2.shape != (n_rows,):
            raise JacksError(JACKS_E_INVALID, 'fixed_x arrays must have one entry per data row')
        io.fixed_x1, io.fixed_x2 = _ptr(fixed_x1), _ptr(fixed_x2)
    for k in res:
        setattr(io, k, _ptr(res[k]))
    flags = (1 if keep_resident else 0) | (0 if sync else 2) | (4 if used_rows_only else 0)
    check(lib().jacks_em_batched_host(ctx.h, C.byref(params), _ptr(offsets), _ptr(rows), n_genes, C.byref(io), flags))
    if not sync:
        # the C side reads these host arrays until ctx.synchronize(): keep them alive with the result
        res['_keepalive'] = (offsets, rows, test, ctrl, fixed_x1, fixed_x2, params)
    return res


def run_screen_device(ctx, plan_real, plan_pseudo, params, io, stream=0):
    """Enqueue the whole inference job of a screen (jacks_run_screen_device): io is a ScreenDeviceIO of device pointers."""
    check(lib().jacks_run_screen_device(ctx.h, plan_real.h, None if plan_pseudo is None else plan_pseudo.h, C.byref(params),
                                        C.byref(io), C.c_void_p(stream or None)))


def run_screen_host(ctx, offsets, rows, pseudo_offsets, pseudo_rows, test, ctrl, params=None, fixed_x1=None, fixed_x2=None,
                    want=('x1', 'x2', 'w1', 'w2', 'pvals', 'n_iters'), want_pseudo=(), positive=False,
                    ctrl_rowmean_fill=False, out=None, sync=True):
    """jacks_run_screen_host: real-gene EM, pseudo-gene EM and KDE p-values of one screen in one call, the screen
    uploaded once.  test [N,L,2]; ctrl [N,L,2] or [N,1,2] (one control for all lines).  want: real-gene outputs out of
    x1, x2, w1, w2, y, tau, n_iters, bound, pvals; want_pseudo: out of w1, w2, n_iters, pvals (returned as 'pseudo_<k>').
    pseudo_offsets=None runs the real genes alone (no p-values)."""
    params = params or default_params()
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    test, ctrl = _cube(test), _cube(ctrl)
    if test.ndim != 3 or test.shape[2] != 2 or ctrl.ndim != 3 or ctrl.shape[2] != 2 or ctrl.shape[0] != test.shape[0]:
        raise JacksError(JACKS_E_SHAPE, 'Expecting matched size between control and test data')
    n_rows, n_lines = test.shape[0], test.shape[1]

    def index(rows_):
        return _rows(rows_, n_rows)
    rows = index(rows)
    n_genes = len(offsets) - 1
    T = int(offsets[-1]) if n_genes > 0 else 0
    n_pseudo = 0
    if pseudo_offsets is not None:
        pseudo_offsets = np.ascontiguousarray(pseudo_offsets, dtype=np.int64)
        pseudo_rows = index(pseudo_rows)
        n_pseudo = len(pseudo_offsets) - 1
    want = tuple(k for k in want if k != 'pvals' or n_pseudo > 0)       # no pseudo-genes at all: plain inferJACKS
    shapes = dict(x1=((T,), np.float64), x2=((T,), np.float64), w1=((n_genes, n_lines), np.float64),
                  w2=((n_genes, n_lines), np.float64), y=((T, n_lines), np.float64), tau=((T, n_lines), np.float64),
                  n_iters=((n_genes,), np.int32), bound=((n_genes,), np.float64), pvals=((n_genes, n_lines), np.float64),
                  pseudo_w1=((n_pseudo, n_lines), np.float64), pseudo_w2=((n_pseudo, n_lines), np.float64),
                  pseudo_n_iters=((n_pseudo,), np.int32), pseudo_pvals=((n_pseudo, n_lines), np.float64))
    keys = list(set(want) | {'w1'}) + ['pseudo_' + k for k in want_pseudo]
    res = {}
    for k in keys:
        if out is not None and k in out:
            assert out[k].shape == shapes[k][0] and out[k].dtype == shapes[k][1] and out[k].flags.c_contiguous
            res[k] = out[k]
        else:
            res[k] = np.empty(shapes[k][0], dtype=shapes[k][1])
    io = ScreenHostIO()
    io.real.test, io.real.ctrl = _ptr(test), _ptr(ctrl)
    io.real.n_rows, io.real.n_lines, io.real.ctrl_lines = int(n_rows), int(n_lines), int(ctrl.shape[1])
    io.real.ctrl_rowmean_fill = int(bool(ctrl_rowmean_fill))
    if fixed_x1 is not None:
        fixed_x1 = np.ascontiguousarray(fixed_x1, dtype=np.float64)
        fixed_x2 = np.ascontiguousarray(fixed_x2, dtype=np.float64)
        if fixed_x1.shape != (n_rows,) or fixed_x2.shape != (n_rows,):
            raise JacksError(JACKS_E_INVALID, 'fixed_x arrays must have one entry per data row')
        io.real.fixed_x1, io.real.fixed_x2 = _ptr(fixed_x1), _ptr(fixed_x2)
    for k in res:
        if k == 'pvals' or k.startswith('pseudo_'):
            setattr(io, k, _ptr(res[k]))
        else:
            setattr(io.real, k, _ptr(res[k]))
    io.positive = int(bool(positive))
    check(lib().jacks_run_screen_host(ctx.h, C.byref(params), _ptr(offsets), _ptr(rows), n_genes,
                                      None if n_pseudo == 0 else _ptr(pseudo_offsets), None if n_pseudo == 0 else _ptr(pseudo_rows),
                                      n_pseudo, C.byref(io), 0 if sync else 2))
    if not sync:
        res['_keepalive'] = (offsets, rows, pseudo_offsets, pseudo_rows, test, ctrl, fixed_x1, fixed_x2, params)
    return res


def em_lines_host(ctx, offsets, rows, test, ctrl, params=None, fixed_x1=None, fixed_x2=None,
                  want=('x1', 'x2', 'w1', 'w2', 'y', 'tau', 'n_iters', 'bound'), keep_resident=False,
                  n_rows=None, n_lines=None):
    """jacks_em_lines_host: every (gene, line) pair as an independent one-line EM (run_JACKS_single.py:83-85).
    test, ctrl [N,L,2] (matched), or test=None to reuse resident cubes.  Outputs carry a line axis:
    w1, w2, bound, n_iters [n_genes, L]; x1, x2, y, tau [sum G, L]."""
    params = params or default_params()
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    rows = np.asarray(rows)
    if test is not None:
        test, ctrl = _cube(test), _cube(ctrl)
        if test.ndim != 3 or test.shape[2] != 2 or ctrl.shape != test.shape:
            raise JacksError(JACKS_E_SHAPE, 'Expecting matched size between control and test data')
        n_rows, n_lines = test.shape[0], test.shape[1]
    rows = _rows(rows, n_rows)
    n_genes = len(offsets) - 1
    T = int(offsets[-1]) if n_genes > 0 else 0
    per = dict(x1=T, x2=T, w1=n_genes, w2=n_genes, y=T, tau=T, n_iters=n_genes, bound=n_genes)
    res = {k: np.empty((per[k], n_lines), dtype=np.int32 if k == 'n_iters' else np.float64) for k in set(want) | {'w1'}}
    io = EmHostIO()
    io.test, io.ctrl = _ptr(test), _ptr(ctrl)
    io.n_rows, io.n_lines, io.ctrl_lines = int(n_rows), int(n_lines), int(n_lines)
    if fixed_x1 is not None:
        fixed_x1 = np.ascontiguousarray(fixed_x1, dtype=np.float64)
        fixed_x2 = np.ascontiguousarray(fixed_x2, dtype=np.float64)
        if fixed_x1.shape != (n_rows,) or fixed_x2.shape != (n_rows,):
            raise JacksError(JACKS_E_INVALID, 'fixed_x arrays must have one entry per data row')
        io.fixed_x1, io.fixed_x2 = _ptr(fixed_x1), _ptr(fixed_x2)
    for k in res:
        setattr(io, k, _ptr(res[k]))
    check(lib().jacks_em_lines_host(ctx.h, C.byref(params), _ptr(offsets), _ptr(rows), n_genes, C.byref(io),
                                    1 if keep_resident else 0))
    return res


def kde_pvalues_host(ctx, w1_genes, w1_pseudo, positive=False):
    w1_genes = np.ascontiguousarray(w1_genes, dtype=np.float64)
    w1_pseudo = np.ascontiguousarray(w1_pseudo, dtype=np.float64)
    if w1_genes.ndim != 2 or w1_pseudo.ndim != 2 or w1_genes.shape[1] != w1_pseudo.shape[1]:
        raise JacksError(JACKS_E_INVALID, 'w1_genes [n_genes, L] and w1_pseudo [n_pseudo, L] expected')
    out = np.empty(w1_genes.shape, dtype=np.float64)
    check(lib().jacks_kde_pvalues_host(ctx.h, _ptr(w1_genes), w1_genes.shape[0], _ptr(w1_pseudo), w1_pseudo.shape[0],
                                       w1_genes.shape[1], int(bool(positive)), _ptr(out)))
    return out


LOG2_TABLE_MAX = 1 << 24


NORM_TYPES = {'median': 0, 'zmad': 1, 'ctrl_guides': 2, 'mode': 3}


def normalize_host(ctx, values, normtype='median', take_log=False, prior=32.0, row_mask=None, use_table=True):
    """normalizeLogCounts on the GPU, optionally fused with log2(count + prior).  Integer counts take their
    logarithm from a table this host computes with numpy (bit-identical to the reference's np.log2 here)."""
    values = np.ascontiguousarray(values, dtype=np.float64)
    out = np.empty(values.shape, dtype=np.float64)
    table = _log2_table(values, prior) if (take_log and use_table) else None
    mask = None if row_mask is None else np.ascontiguousarray(row_mask, dtype=np.uint8)
    check(lib().jacks_normalize_host(ctx.h, _ptr(values), values.shape[0], values.shape[1], int(bool(take_log)), float(prior),
                                     _ptr(table), 0 if table is None else table.size, NORM_TYPES[normtype], _ptr(mask),
                                     _ptr(out)))
    return out


def _log2_table(values, prior):
    """log2(k + prior) for every integer count k up to the largest in `values`, computed by THIS host's numpy (integer counts
    then take exactly the bits the reference's np.log2 produces here); None when there is nothing to tabulate."""
    if not values.size:
        return None
    vmax = values.max()                      # one pass, no temporary; NaN / inf cells take the slow path
    if not np.isfinite(vmax):
        finite = values[np.isfinite(values)]
        vmax = finite.max() if finite.size else -1.0
    top = int(min(vmax, LOG2_TABLE_MAX - 1)) if vmax >= 0 else -1
    return np.log2(np.arange(top + 1, dtype=np.float64) + prior) if top >= 0 else None


def _sample_csr(Iscreens):
    offs = np.zeros(len(Iscreens) + 1, dtype=np.int32)
    offs[1:] = np.cumsum([len(c) for c in Iscreens])
    cols = np.ascontiguousarray(np.concatenate([np.asarray(c, dtype=np.int32) for c in Iscreens]) if len(Iscreens) else np.zeros(0, np.int32), dtype=np.int32)
    return offs, cols


def preprocess_host(ctx, counts, Iscreens, normtype='median', prior=32.0, take_log=True, row_mask=None, row_order=None, window=0,
                    monotonize=True, want_logcounts=False):
    """jacks_preprocess_host: counts [N, C] -> (data [N, S, 2], logcounts or None) with one upload and one download:
    log2(count + prior), normalizeLogCounts, the row sort by guide id (row_order = argsort of the ids) and
    condense_normalised_counts, chained on the device."""
    counts = np.ascontiguousarray(counts, dtype=np.float64)
    N, Cc = counts.shape
    table = _log2_table(counts, prior) if take_log else None
    offs, cols = _sample_csr(Iscreens)
    mask = None if row_mask is None else np.ascontiguousarray(row_mask, dtype=np.uint8)
    order = None if row_order is None else np.ascontiguousarray(row_order, dtype=np.int64)
    data = np.empty((N, len(Iscreens), 2), dtype=np.float64)
    lc = np.empty((N, Cc), dtype=np.float64) if want_logcounts else None
    check(lib().jacks_preprocess_host(ctx.h, _ptr(counts), N, Cc, int(bool(take_log)), float(prior), _ptr(table),
                                      0 if table is None else table.size, NORM_TYPES[normtype], _ptr(mask), _ptr(order), _ptr(offs),
                                      _ptr(cols), len(Iscreens), int(window), int(bool(monotonize)), _ptr(lc), _ptr(data)))
    return data, lc


def preprocess_device_ms(ctx, counts, Iscreens, prior=32.0, reps=3):
    """Device time (ms, CUDA events) of the resident chain jacks_normalize_device -> jacks_condense_device on counts
    already in HBM -- the figure behind bench.py's aux.preprocess.roofline.  Needs torch for the device buffers."""
    import torch
    counts = np.ascontiguousarray(counts, dtype=np.float64)
    N, Cc = counts.shape
    dev = torch.device('cuda', ctx.device)
    d_counts = torch.from_numpy(counts).to(dev)
    table = _log2_table(counts, prior)
    d_table = torch.from_numpy(table).to(dev)
    d_lc = torch.empty_like(d_counts)
    d_data = torch.empty((N, len(Iscreens), 2), dtype=torch.float64, device=dev)
    offs, cols = _sample_csr(Iscreens)
    stream = torch.cuda.current_stream().cuda_stream

    def chain():
        check(lib().jacks_normalize_device(ctx.h, d_counts.data_ptr(), N, Cc, 1, float(prior), d_table.data_ptr(), table.size,
                                           NORM_TYPES['median'], None, d_lc.data_ptr(), C.c_void_p(stream or None)))
        check(lib().jacks_condense_device(ctx.h, d_lc.data_ptr(), N, Cc, _ptr(offs), _ptr(cols), len(Iscreens), 0, 1,
                                          d_data.data_ptr(), C.c_void_p(stream or None)))
    chain()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        chain()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def lognorm_median_host(ctx, counts, prior, use_table=True):
    return normalize_host(ctx, counts, 'median', take_log=True, prior=prior, use_table=use_table)


def condense_host(ctx, logcounts, Iscreens, window=0, monotonize=True):
    logcounts = np.ascontiguousarray(logcounts, dtype=np.float64)
    offs, cols = _sample_csr(Iscreens)
    data = np.empty((logcounts.shape[0], len(Iscreens), 2), dtype=np.float64)
    check(lib().jacks_condense_host(ctx.h, _ptr(logcounts), logcounts.shape[0], logcounts.shape[1], _ptr(offs),
                                    _ptr(cols), len(Iscreens), int(window), int(bool(monotonize)), _ptr(data)))
    return data


def posterior_sd_subset_host(ctx, data, in_set, window, monotonize=True):
    """calc_posterior_sd fitted on the guides flagged in `in_set` and interpolated to all guides (GPU)."""
    data = np.ascontiguousarray(data, dtype=np.float64)
    mask = np.ascontiguousarray(in_set, dtype=np.uint8)
    out = np.empty(data.shape[0], dtype=np.float64)
    check(lib().jacks_posterior_sd_subset_host(ctx.h, _ptr(data), data.shape[0], data.shape[1], _ptr(mask), int(window),
                                               int(bool(monotonize)), _ptr(out)))
    return out


def kde_eval_host(ctx, w_eval, w_null, kind='cdf_neg', leave_one_out=False):
    """jacks_kde_eval_host: kind in {'cdf_neg', 'cdf_pos', 'pdf'}; w_eval [m, L], w_null [n, L] -> [m, L]."""
    w_eval = np.ascontiguousarray(w_eval, dtype=np.float64)
    w_null = np.ascontiguousarray(w_null, dtype=np.float64)
    out = np.empty(w_eval.shape, dtype=np.float64)
    check(lib().jacks_kde_eval_host(ctx.h, _ptr(w_eval), w_eval.shape[0], _ptr(w_null), w_null.shape[0], w_eval.shape[1],
                                    {'cdf_neg': 0, 'cdf_pos': 1, 'pdf': 2}[kind], int(bool(leave_one_out)), _ptr(out)))
    return out


def table_read(path, delim, sel_cols, skip_lines=1, str_cols=(0,)):
    """Native parse of a delimiter-separated table (jacks_table_read): returns (labels, values, not_numeric) --
    the text of column str_cols[0] of every record (a list of tuples when several text columns are asked for; None
    for a missing field), float64 [n_records, len(sel_cols)] of the selected columns, and a mask of cells that are
    not plain decimal literals (NaN in `values`)."""
    import mmap
    n, off = C.c_int64(), C.c_int64()
    bpath = os.fsencode(path)
    check(lib().jacks_table_scan(bpath, int(skip_lines), C.byref(n), C.byref(off)))
    n = n.value
    sel = np.ascontiguousarray(sel_cols, dtype=np.int32)
    sc = np.ascontiguousarray(str_cols, dtype=np.int32)
    values = np.empty((n, len(sel)), dtype=np.float64)
    bad = np.empty((n, len(sel)), dtype=np.uint8)
    loff = np.empty((n, len(sc)), dtype=np.int64)
    llen = np.empty((n, len(sc)), dtype=np.int32)
    check(lib().jacks_table_read(bpath, off.value, delim.encode()[:1], n, len(sel), _ptr(sel), _ptr(values), _ptr(bad),
                                 len(sc), _ptr(sc), _ptr(loff), _ptr(llen)))
    cols = [[] for _ in range(len(sc))]
    if n:
        with open(path, 'rb') as f:
            mm = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)
            try:
                for k in range(len(sc)):
                    col = cols[k]
                    for o, ln in zip(loff[:, k].tolist(), llen[:, k].tolist()):
                        col.append(None if ln < 0 else mm[o:o + ln].decode())
            finally:
                mm.close()
    labels = cols[0] if len(sc) == 1 else list(zip(*cols))
    return labels, values, bad.astype(bool)


def table_write(path, header, labels, values, width, blank=None):
    """Native formatter (jacks_table_write): one line per row, `label` then tab-separated '%<width>e' values."""
    values = np.ascontiguousarray(values, dtype=np.float64)
    if values.ndim == 1:
        values = values.reshape(len(labels), -1)
    enc = [l.encode() for l in labels]
    offs = np.zeros(len(enc) + 1, dtype=np.int64)
    if enc:
        np.cumsum([len(b) for b in enc], out=offs[1:])
    blob = b''.join(enc)
    mask = None if blank is None else np.ascontiguousarray(blank, dtype=np.uint8)
    check(lib().jacks_table_write(os.fsencode(path), header.encode(), len(enc), values.shape[1] if values.size or len(enc) else 0,
                                  blob, _ptr(offs), _ptr(values), _ptr(mask), int(width)))


def pseudo_draw(mt_state, offsets, rows, ctrl_genes, n_pseudo):
    """jacks_pseudo_draw: the draws of createPseudoNonessGenes on a Mersenne-Twister state.  mt_state: the 625 words of
    random.getstate()[1]; offsets / rows: the gene index (CSR, dict order); ctrl_genes: positions of the control genes
    in it, in list order.  Returns (new_state uint32[625], pseudo_offsets int64[n_pseudo + 1], pseudo_rows int32)."""
    st = np.array(mt_state, dtype=np.uint32)
    if st.shape != (625,):
        raise ValueError('expected the 625 state words of random.getstate()[1]')
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    rows = np.ascontiguousarray(rows, dtype=np.int32)
    ctrl_genes = np.ascontiguousarray(ctrl_genes, dtype=np.int64)
    n_genes = len(offsets) - 1
    n_pseudo = int(n_pseudo)
    poffs = np.zeros(n_pseudo + 1, dtype=np.int64)
    mean_g = (int(offsets[-1]) + n_genes - 1) // n_genes if n_genes > 0 else 1
    cap = n_pseudo * mean_g + (n_pseudo * mean_g) // 4 + 1024
    need = C.c_int64(0)
    while True:
        prows = np.empty(cap, dtype=np.int32)
        rc = lib().jacks_pseudo_draw(st.ctypes.data, _ptr(offsets), _ptr(rows), n_genes, _ptr(ctrl_genes), len(ctrl_genes),
                                     n_pseudo, poffs.ctypes.data, prows.ctypes.data, cap, C.byref(need))
        if rc == JACKS_E_SHAPE and need.value > cap:
            cap = int(need.value)                 # the state was left untouched: draw again into a buffer that holds them
            continue
        check(rc)
        return st, poffs, prows[:int(need.value)]
