"""Gene-sharded data parallelism over the GPUs of one box (one process per GPU, torch.distributed).

Genes are independent (jacks/jacks/infer.py:20-29 carries no state from one gene to the next), so the EM shards by gene
with NO collective inside the iteration.  The job of one screen (jacks/jacks/jacks_io.py:425-441) runs like this:

  1. every rank takes a contiguous block of the real genes and of the pseudo-genes (balanced by guide count) and holds a
     compact cube [control block | rows of my real genes] with a re-based index.  The control block -- the few thousand
     control-guide rows the pseudo-genes draw from, the same on every rank -- goes up in one slice per rank and ONE
     all-gather over NVLink completes it; then every rank uploads the rows of its own real genes: nothing goes up twice;
  2. EM of both blocks on the rank's GPU (the same kernels as the one-GPU call); the pseudo-gene EM starts as soon as the
     control block is complete, beside the rest of the upload and beside the real-gene EM (side stream);
  3. ONE all-to-all over NCCL / NVLink turns "my genes x all lines" into "all genes x my lines" for E[w] of real and
     pseudo-genes (each rank keeps L/W lines: 1/W of what an all-gather of the pseudo-gene E[w] would move);
  4. the KDE null of a line needs all pseudo-genes of that line and nothing else, so the p-values are line-sharded:
     every rank prepares the null densities of its lines and evaluates every real gene on them;
  5. run(): ONE gather (a one-sided all-to-all) hands rank `dst` the per-gene results as device tensors: w1, w2, x1, x2,
     iteration counts and the p-value column blocks.  run_to_host(): no gather -- a second small all-to-all turns the
     p-values back into "my genes x all lines" and every rank copies the rows of ITS genes into one shared, page-locked
     host segment (SharedHostResults) over its own PCIe link.

Nothing is reduced across ranks and every per-line sum runs over the pseudo-genes in their original order, so 1-rank and
N-rank results are bit-identical.  All buffers between steps 2 and 5 stay on the device; under the gloo backend (CPU
tests) the same driver runs on host tensors with an injected engine (CPU restatements of the kernels).
"""
import os

import numpy as np


def shard_bounds(offsets, world):
    """Contiguous gene blocks [b[r], b[r+1]) with near-equal total guide count (EM cost is ~ G*L per pass)."""
    offsets = np.asarray(offsets, dtype=np.int64)
    n = len(offsets) - 1
    total = int(offsets[-1]) if n > 0 else 0
    targets = (np.arange(1, world, dtype=np.float64) * total / world)
    cuts = np.searchsorted(offsets[1:], targets, side='left') + 1 if n > 0 else np.zeros(world - 1, dtype=np.int64)
    b = np.concatenate([[0], np.minimum(cuts, n), [n]]).astype(np.int64)
    return np.maximum.accumulate(b)


def local_index(offsets, rows, lo, hi):
    """CSR index of genes [lo, hi) re-based to start at 0."""
    offsets = np.asarray(offsets, dtype=np.int64)
    return offsets[lo:hi + 1] - offsets[lo], np.asarray(rows)[offsets[lo]:offsets[hi]]


def line_bounds(n_lines, world):
    """Contiguous line blocks [lb[r], lb[r+1]) of the p-value stage."""
    return np.linspace(0, n_lines, world + 1).astype(np.int64)


def compact_rows(index_rows, n_rows):
    """Rows an index references (ascending, unique) and the index re-based onto them."""
    index_rows = np.asarray(index_rows, dtype=np.int64)
    used = np.unique(index_rows)
    remap = np.full(n_rows, -1, dtype=np.int64)
    remap[used] = np.arange(len(used))
    return used, remap[index_rows].astype(np.int32)


def two_part_layout(real_rows, pseudo_rows_mine, pseudo_rows_all, n_rows, world):
    """Row layout of a rank's compact cube: [P | R].  P = every row ANY pseudo-gene references (the control guides; the same
    set on every rank), padded to world * k rows so that each rank uploads only its k-row slice of it and one all-gather
    over NVLink completes the block; R = the rows of this rank's real genes that are not in P.  Returns (P, R, k) and the
    two indexes re-based onto the layout."""
    P = np.unique(np.asarray(pseudo_rows_all, dtype=np.int64))
    k = (len(P) + world - 1) // world
    R = np.setdiff1d(np.unique(np.asarray(real_rows, dtype=np.int64)), P, assume_unique=True)
    pos = np.full(n_rows, -1, dtype=np.int64)
    pos[P] = np.arange(len(P))
    pos[R] = world * k + np.arange(len(R))
    return P, R, k, pos[np.asarray(real_rows, dtype=np.int64)].astype(np.int32), \
        pos[np.asarray(pseudo_rows_mine, dtype=np.int64)].astype(np.int32)


def _world(group):
    import torch.distributed as dist
    if not dist.is_initialized():
        return 1, 0
    return dist.get_world_size(group), dist.get_rank(group)


def _resolve_device(device, group):
    """The CUDA device of this rank: the explicit argument, else torch's current device (what torchrun scripts set from
    LOCAL_RANK) -- one answer for the kernels AND the collectives.  None under gloo."""
    import torch
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_backend(group) != 'nccl':
        return None
    if not torch.cuda.is_available():
        return None
    return int(device) if device is not None else torch.cuda.current_device()


# ----------------------------------------------------------------------------------------------------------------------
# Engines: what one rank computes.  The product engine drives the C-ABI on device tensors; tests inject a CPU engine.
# ----------------------------------------------------------------------------------------------------------------------
class NativeEngine(object):
    """This rank's GPU through the C-ABI device entry points (jacks_em_run_device, jacks_kde_pvalues_device) on torch
    CUDA tensors, everything enqueued on torch's current stream."""

    def __init__(self, device):
        import torch
        from . import _native
        self.torch, self.nat = torch, _native
        self.device = torch.device('cuda', int(device))
        self.ctx = _native.context(int(device))
        self._plans = {}
        self._side = None
        self._up = None
        self._down, self._down_keep = None, []

    def upload(self, array):
        """Host array -> device tensor (through a pinned copy when the source is pageable)."""
        t = self.torch.from_numpy(np.ascontiguousarray(array))
        return t.to(self.device, non_blocking=True)

    def alloc(self, shape):
        return self.torch.empty(tuple(int(x) for x in shape), dtype=self.torch.float64, device=self.device)

    # -- staged upload: copies and the control-row all-gather run on an upload stream; marks are events the EM launches wait
    #    for (the pseudo-gene EM only for the control block, so that it runs while the screen rows are still arriving)
    def begin_upload(self):
        torch = self.torch
        if self._up is None:
            self._up = torch.cuda.Stream(device=self.device)
        self._up.wait_stream(torch.cuda.current_stream(self.device))     # earlier kernels may still read the cubes
        if self._side is not None:
            self._up.wait_stream(self._side)

    def h2d(self, dst, src):
        """dst: contiguous device view; src: host array (numpy, or a torch CPU tensor -- pinned for a truly asynchronous copy)."""
        torch = self.torch
        if dst.numel() == 0:
            return
        t = src if torch.is_tensor(src) else torch.from_numpy(np.ascontiguousarray(src))
        with torch.cuda.stream(self._up):
            dst.copy_(t.view(dst.shape), non_blocking=True)

    def allgather_rows(self, cube, k, world, rank, group):
        """cube[:world*k] <- the k-row slices every rank has uploaded at its own position (in place, NCCL over NVLink)."""
        import torch.distributed as dist
        with self.torch.cuda.stream(self._up):
            dist.all_gather_into_tensor(cube[:world * k], cube[rank * k:(rank + 1) * k], group=group)

    def mark(self):
        ev = self.torch.cuda.Event()
        ev.record(self._up)
        return ev

    def wait(self, ev, side=False):
        if ev is not None:
            (self._side if side else self.torch.cuda.current_stream(self.device)).wait_event(ev)

    def plan(self, key, offsets, rows, n_rows, n_lines):
        """Plans are cached per key: a resident screen is inferred again and again (benchmark, bootstrap draws)."""
        if key is None or key not in self._plans:
            p = self.nat.EmPlan(self.ctx, offsets, rows, n_rows, n_lines)
            if key is None:
                return p
            self._plans[key] = p
        return self._plans[key]

    def fork(self):
        """Open a side stream behind everything enqueued so far: a second EM launched with side=True runs BESIDE the
        kernels enqueued on the current stream afterwards (both are persistent kernels, so its CTAs move in as the first
        kernel's drain -- the first kernel's tail costs nothing).  join() makes the current stream wait for it."""
        torch = self.torch
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
        self._side.wait_stream(torch.cuda.current_stream(self.device))

    def join(self):
        if self._side is not None:
            self.torch.cuda.current_stream(self.device).wait_stream(self._side)

    def em(self, plan, test, ctrl, params, fixed_x1=None, fixed_x2=None, want_x=True, y_tau=False, w1_out=None, side=False):
        """w1_out: an [n, L] device view E[w] is written into (the exchange buffer), instead of a fresh tensor.
        side: launch on the side stream opened by fork(); the caller join()s before the current stream reads the results."""
        torch, nat = self.torch, self.nat
        n, T, L = plan.n_genes, plan.total_guides, plan.n_lines
        f64 = dict(dtype=torch.float64, device=self.device)
        # side: the outputs come from the SIDE stream's pool -- a block the current stream has just freed (a temporary of
        # the gather, say) may still be read by kernels the side stream does not wait for
        with torch.cuda.stream(self._side if side else torch.cuda.current_stream(self.device)):
            out = dict(w1=w1_out if w1_out is not None else torch.empty((n, L), **f64), w2=torch.empty((n, L), **f64),
                       n_iters=torch.empty(n, dtype=torch.int32, device=self.device))
            if want_x:
                out['x1'], out['x2'] = torch.empty(T, **f64), torch.empty(T, **f64)
            if y_tau:
                out['y'], out['tau'] = torch.empty((T, L), **f64), torch.empty((T, L), **f64)
            if n == 0:
                return out
            io = nat.EmDeviceIO()
            io.test, io.ctrl, io.ctrl_lines = test.data_ptr(), ctrl.data_ptr(), int(ctrl.shape[1])
            if fixed_x1 is not None:
                io.fixed_x1, io.fixed_x2 = fixed_x1.data_ptr(), fixed_x2.data_ptr()
            for k, v in out.items():
                setattr(io, k, v.data_ptr())
            plan.run_device(params, io, torch.cuda.current_stream(self.device).cuda_stream)
        out['_keep'] = (test, ctrl, fixed_x1, fixed_x2)
        return out

    def to_host(self, dst, src):
        """Enqueue the copy of device tensor ``src`` into the numpy view ``dst`` (contiguous, page-locked: SharedHostResults)
        on this engine's download stream, behind everything enqueued on the current stream so far."""
        import ctypes as C
        torch = self.torch
        if src.numel() == 0:
            return
        assert dst.flags['C_CONTIGUOUS'] and src.is_contiguous() and dst.nbytes == src.numel() * src.element_size()
        if self._down is None:
            self._down = torch.cuda.Stream(device=self.device)
        self._down.wait_stream(torch.cuda.current_stream(self.device))
        self.nat.check(self.nat.lib().jacks_copy_to_host_async(C.c_void_p(dst.ctypes.data), C.c_void_p(src.data_ptr()),
                                                               dst.nbytes, C.c_void_p(self._down.cuda_stream)))
        self._down_keep.append(src)

    def host_done(self):
        """Wait for the copies enqueued by to_host()."""
        if self._down is not None:
            self._down.synchronize()
        del self._down_keep[:]

    def kde(self, w1_genes, w1_pseudo, positive):
        import ctypes as C
        torch, nat = self.torch, self.nat
        pv = torch.empty_like(w1_genes)
        if w1_genes.shape[0] and w1_genes.shape[1]:
            nat.check(nat.lib().jacks_kde_pvalues_device(self.ctx.h, w1_genes.data_ptr(), w1_genes.shape[0], w1_pseudo.data_ptr(),
                                                         w1_pseudo.shape[0], w1_genes.shape[1], int(bool(positive)), pv.data_ptr(),
                                                         C.c_void_p(torch.cuda.current_stream().cuda_stream or None)))
        return pv


class HostEngine(object):
    """CPU stand-in with the engine interface, built from two callables (the gloo tests pass CPU restatements):
    em_fn(offsets, rows, test, ctrl, fixed_x1, fixed_x2) -> dict(w1, w2, x1, x2, n_iters) of numpy arrays,
    kde_fn(w1_genes, w1_pseudo, positive) -> [m, L]."""

    def __init__(self, em_fn, kde_fn):
        import torch
        self.torch, self.em_fn, self.kde_fn = torch, em_fn, kde_fn
        self.device = torch.device('cpu')

    def upload(self, array):
        return self.torch.from_numpy(np.ascontiguousarray(array))

    def plan(self, key, offsets, rows, n_rows, n_lines):
        class P(object):
            pass
        p = P()
        p.offsets, p.rows = np.asarray(offsets, dtype=np.int64), np.asarray(rows, dtype=np.int32)
        p.n_genes, p.total_guides, p.n_lines = len(p.offsets) - 1, int(p.offsets[-1]) if len(p.offsets) else 0, n_lines
        return p

    def fork(self):
        pass

    def join(self):
        pass

    def alloc(self, shape):
        return self.torch.zeros(tuple(int(x) for x in shape), dtype=self.torch.float64)

    def begin_upload(self):
        pass

    def h2d(self, dst, src):
        if dst.numel():
            t = src if self.torch.is_tensor(src) else self.torch.from_numpy(np.ascontiguousarray(src))
            dst.copy_(t.view(dst.shape))

    def allgather_rows(self, cube, k, world, rank, group):
        import torch.distributed as dist
        parts = [self.torch.empty_like(cube[:k]) for _ in range(world)]
        dist.all_gather(parts, cube[rank * k:(rank + 1) * k].contiguous(), group=group)
        for s_, part in enumerate(parts):
            cube[s_ * k:(s_ + 1) * k] = part

    def mark(self):
        return None

    def wait(self, ev, side=False):
        pass

    def em(self, plan, test, ctrl, params, fixed_x1=None, fixed_x2=None, want_x=True, y_tau=False, w1_out=None, side=False):
        t = self.torch
        L = plan.n_lines
        if plan.n_genes == 0:
            return dict(w1=t.zeros((0, L), dtype=t.float64), w2=t.zeros((0, L), dtype=t.float64), x1=t.zeros(0, dtype=t.float64),
                        x2=t.zeros(0, dtype=t.float64), n_iters=t.zeros(0, dtype=t.int32))
        r = self.em_fn(plan.offsets, plan.rows, test.numpy(), ctrl.numpy(), None if fixed_x1 is None else fixed_x1.numpy(),
                       None if fixed_x2 is None else fixed_x2.numpy())
        out = {k: t.from_numpy(np.ascontiguousarray(r[k])) for k in ('w1', 'w2', 'x1', 'x2', 'n_iters')}
        if w1_out is not None:
            w1_out.copy_(out['w1'])
            out['w1'] = w1_out
        return out

    def to_host(self, dst, src):
        dst[...] = src.numpy().reshape(dst.shape)

    def host_done(self):
        pass

    def kde(self, w1_genes, w1_pseudo, positive):
        if w1_genes.shape[0] == 0 or w1_genes.shape[1] == 0:
            return self.torch.zeros(tuple(w1_genes.shape), dtype=self.torch.float64)
        return self.torch.from_numpy(np.ascontiguousarray(self.kde_fn(w1_genes.numpy(), w1_pseudo.numpy(), positive)))


# ----------------------------------------------------------------------------------------------------------------------
# The sharded screen job
# ----------------------------------------------------------------------------------------------------------------------
class SharedHostResults(object):
    """The result arrays of one screen -- w1, w2, pvals [n, L]; x1, x2 [sum G]; n_iters [n] -- in ONE shared-memory segment
    that every rank process of the node maps and page-locks.  With it ShardedScreen.run_to_host() needs no gather at all:
    each rank copies the rows of its own genes down over its own PCIe link and rank ``owner`` reads the assembled arrays
    in place (the 174 MB of an Avana-shaped screen leave the node's eight GPUs in parallel instead of queueing on one
    link behind an NVLink gather).  Collective over ``group``: construct and close() on every rank."""

    FIELDS = (('w1', 'nL', np.float64), ('w2', 'nL', np.float64), ('pvals', 'nL', np.float64), ('x1', 'T', np.float64),
              ('x2', 'T', np.float64), ('n_iters', 'n', np.int32))

    def __init__(self, n_genes, total_guides, n_lines, group=None, owner=0, register=True):
        import mmap
        import torch.distributed as dist
        world, rank = _world(group)
        self.group, self.owner, self.rank = group, owner, rank
        n, T, L = int(n_genes), int(total_guides), int(n_lines)
        shapes = {'nL': (n, L), 'T': (T,), 'n': (n,)}
        offs, o = {}, 0
        for name, kind, dt in self.FIELDS:
            offs[name] = o
            o += (int(np.prod(shapes[kind], dtype=np.int64)) * np.dtype(dt).itemsize + 255) // 256 * 256
        self.nbytes = max(o, 256)
        path = [None]
        if rank == owner:
            path[0] = '/dev/shm/jacks_b200_%d_%s' % (os.getpid(), os.urandom(6).hex())
            fd = os.open(path[0], os.O_CREAT | os.O_EXCL | os.O_RDWR, 0o600)
            os.ftruncate(fd, self.nbytes)
        if world > 1:
            dist.broadcast_object_list(path, src=owner, group=group)
        if rank != owner:
            fd = os.open(path[0], os.O_RDWR)
        self._map = mmap.mmap(fd, self.nbytes)
        os.close(fd)
        if world > 1:
            dist.barrier(group=group)
        if rank == owner:
            os.unlink(path[0])                   # every rank holds its mapping: nothing is left behind if a rank dies
        base = np.frombuffer(self._map, dtype=np.uint8)
        self.arrays = {}
        for name, kind, dt in self.FIELDS:
            size = int(np.prod(shapes[kind], dtype=np.int64))
            self.arrays[name] = base[offs[name]:offs[name] + size * np.dtype(dt).itemsize].view(dt).reshape(shapes[kind])
        self._registered = None
        if register:
            from . import _native
            import ctypes as C
            self._registered = C.c_void_p(base.ctypes.data)
            _native.check(_native.lib().jacks_host_register(self._registered, self.nbytes))

    def __getitem__(self, key):
        return self.arrays[key]

    def close(self):
        if self._registered is not None:
            from . import _native
            _native.lib().jacks_host_unregister(self._registered)
            self._registered = None
        self.arrays = {}
        try:
            self._map.close()
        except BufferError:
            pass                                 # a caller still holds a view: the mapping goes with it


class ShardedScreen(object):
    """One screen, sharded over the ranks of ``group``: build once (index split, compact upload, plans), run() many times.

    offsets/rows: CSR index of the real genes; pseudo_offsets/pseudo_rows: of the pseudo-genes (or None: no p-values).
    testdata [N,L,2]; ctrldata [N,L,2] or [N,1,2].  fixed_x: {'X1','X2'} by row, real genes only (jacks_io.py:425, :430).
    """

    def __init__(self, offsets, rows, pseudo_offsets, pseudo_rows, testdata, ctrldata, engine=None, group=None, device=None,
                 fixed_x=None, params=None, positive=False, dst=0, shard=True):
        import torch
        self.torch, self.group, self.dst, self.positive = torch, group, dst, bool(positive)
        self.world, self.rank = _world(group) if shard else (1, 0)      # shard=False: the whole screen on this rank (replicas)
        if not shard:
            self.dst = 0
        dev = _resolve_device(device, group)
        if engine is None:
            if dev is None:
                raise RuntimeError('jacks_b200.dist: no CUDA device for this rank and no engine given (there is no CPU fallback)')
            engine = NativeEngine(dev)
        self.engine = engine
        from . import _native
        self.params = params if params is not None else (_native.default_params() if isinstance(engine, NativeEngine) else None)
        offsets = np.asarray(offsets, dtype=np.int64)
        rows = np.asarray(rows)
        N, L = int(testdata.shape[0]), int(testdata.shape[1])
        self.N, self.L = N, L
        self.n_genes = len(offsets) - 1
        self.T = int(offsets[-1]) if self.n_genes > 0 else 0
        self.has_pseudo = pseudo_offsets is not None and len(pseudo_offsets) > 1
        self.n_pseudo = len(pseudo_offsets) - 1 if self.has_pseudo else 0
        W, r = self.world, self.rank
        self.b = shard_bounds(offsets, W)
        self.bt = offsets[self.b]                                       # guide offsets of the gene blocks
        lo, hi = int(self.b[r]), int(self.b[r + 1])
        loffs, lrows = local_index(offsets, rows, lo, hi)
        if self.has_pseudo:
            pseudo_offsets = np.asarray(pseudo_offsets, dtype=np.int64)
            self.bp = shard_bounds(pseudo_offsets, W)
            plo, phi = int(self.bp[r]), int(self.bp[r + 1])
            poffs, prows = local_index(pseudo_offsets, pseudo_rows, plo, phi)
        else:
            self.bp = np.zeros(W + 1, dtype=np.int64)
            poffs, prows = np.zeros(1, dtype=np.int64), np.zeros(0, dtype=np.int64)
        self.lb = line_bounds(L, W)
        # compact cube [P | R]: the control-guide rows of ALL pseudo-genes (each rank uploads 1/W of them, one all-gather
        # over NVLink completes the block) followed by the rows of this rank's real genes -- nothing else goes up
        P, R, k, real_idx, pseudo_idx = two_part_layout(lrows, prows, pseudo_rows if self.has_pseudo else np.zeros(0, np.int64), N, W)
        self.P, self.R, self.k = P, R, k
        self.used = np.concatenate([P, R])                              # rows resident on this rank (cube order, unpadded)
        U = W * k + len(R)
        self.d_test, self.d_ctrl = engine.alloc((U, L, 2)), engine.alloc((U, int(ctrldata.shape[1]), 2))
        self._ev_ctrl = self._ev_all = None
        self.h2d_bytes = 0
        self.upload(self.host_parts(testdata, ctrldata))
        self.d_fx1 = self.d_fx2 = None
        if fixed_x is not None:
            cube_rows = np.concatenate([P, np.zeros(W * k - len(P), dtype=np.int64), R])
            fx1 = np.ascontiguousarray(np.asarray(fixed_x['X1'], dtype=np.float64)[cube_rows])
            fx2 = np.ascontiguousarray(np.asarray(fixed_x['X2'], dtype=np.float64)[cube_rows])
            self.d_fx1, self.d_fx2 = engine.upload(fx1), engine.upload(fx2)
            self.h2d_bytes += fx1.nbytes + fx2.nbytes
        self.plan_real = engine.plan(None, loffs, real_idx, U, L)
        self.plan_pseudo = engine.plan(None, poffs, pseudo_idx, U, L) if self.has_pseudo else None
        self.h2d_bytes += loffs.nbytes + poffs.nbytes + real_idx.nbytes + pseudo_idx.nbytes

    # -- upload ------------------------------------------------------------------------------------------------------
    def host_parts(self, testdata, ctrldata):
        """The four contiguous host blocks this rank uploads: its slice of the control block and its real-gene rows, of
        the test and of the control cube (callers that re-upload per step keep them in pinned memory)."""
        r, k = self.rank, self.k
        mine = self.P[r * k:(r + 1) * k]
        td, cd = np.asarray(testdata), np.asarray(ctrldata)
        return dict(test_p=np.ascontiguousarray(td[mine]), ctrl_p=np.ascontiguousarray(cd[mine]),
                    test_r=np.ascontiguousarray(td[self.R]), ctrl_r=np.ascontiguousarray(cd[self.R]))

    def upload(self, parts):
        """Enqueue the upload of this rank's share of the screen (host_parts) on the engine's upload stream: control block
        first -- my slice, then the all-gather -- so that the pseudo-gene EM of the next run() starts while the real-gene
        rows are still on their way up."""
        eng, W, r, k = self.engine, self.world, self.rank, self.k
        n_mine = len(self.P[r * k:(r + 1) * k])
        eng.begin_upload()
        if k:
            eng.h2d(self.d_test[r * k:r * k + n_mine], parts['test_p'])
            eng.h2d(self.d_ctrl[r * k:r * k + n_mine], parts['ctrl_p'])
            if W > 1:
                eng.allgather_rows(self.d_test, k, W, r, self.group)
                eng.allgather_rows(self.d_ctrl, k, W, r, self.group)
        self._ev_ctrl = eng.mark()
        eng.h2d(self.d_test[W * k:], parts['test_r'])
        eng.h2d(self.d_ctrl[W * k:], parts['ctrl_r'])
        self._ev_all = eng.mark()
        self.h2d_bytes = sum(int(v.numel() * v.element_size()) if self.torch.is_tensor(v) else int(v.nbytes) for v in parts.values())

    # -- collectives -------------------------------------------------------------------------------------------------
    def _all_to_all(self, send, in_splits, out_splits, async_op=False):
        """all_to_all_single with uneven splits.  async_op: returns (work, recv) and does not make the current stream
        wait for the transfer."""
        import torch.distributed as dist
        if self.world == 1:
            return (None, send) if async_op else send
        recv = self.torch.empty(int(sum(out_splits)), dtype=send.dtype, device=send.device)
        work = dist.all_to_all_single(recv, send, output_split_sizes=[int(x) for x in out_splits],
                                      input_split_sizes=[int(x) for x in in_splits], group=self.group, async_op=async_op)
        return (work, recv) if async_op else recv

    def _lines_of_all_genes(self, w1_local, gene_bounds):
        """[my genes, L] on every rank -> [all genes, my lines] on every rank: ONE all-to-all (step 3)."""
        t, lb, r = self.torch, self.lb, self.rank
        n_local = w1_local.shape[0]
        blocks = [w1_local[:, int(lb[s]):int(lb[s + 1])].reshape(-1) for s in range(self.world)]
        send = t.cat(blocks) if self.world > 1 else blocks[0].contiguous()
        in_splits = [n_local * int(lb[s + 1] - lb[s]) for s in range(self.world)]
        my_l = int(lb[r + 1] - lb[r])
        out_splits = [int(gene_bounds[s + 1] - gene_bounds[s]) * my_l for s in range(self.world)]
        recv = self._all_to_all(send, in_splits, out_splits)
        return recv.view(int(gene_bounds[-1]), my_l)

    def run(self, want_pvals=True):
        """Steps 2-5.  Returns on rank ``dst`` a dict of device tensors in the caller's gene order -- w1, w2 [n, L];
        x1, x2 [sum G]; n_iters [n]; pvals [n, L] (if there are pseudo-genes and want_pvals) -- and None elsewhere."""
        t, eng = self.torch, self.engine
        W, r, L = self.world, self.rank, self.L
        do_p = bool(want_pvals and self.has_pseudo and self.n_pseudo >= 2)
        n_lr, n_lp = int(self.b[r + 1] - self.b[r]), int(self.bp[r + 1] - self.bp[r])
        t_l = int(self.bt[r + 1] - self.bt[r])
        pv_block = None
        early = None

        def gather_em(real):
            """Gather of the real genes' EM results (w1 | w2 | x1 | x2 | n_iters as f64) on dst, ASYNCHRONOUS: it travels
            while the pseudo-gene EM and the p-value stage run."""
            send = t.cat([real['w1'].reshape(-1), real['w2'].reshape(-1), real['x1'], real['x2'], real['n_iters'].to(t.float64)])
            sizes = [2 * int(self.b[s + 1] - self.b[s]) * L + 2 * int(self.bt[s + 1] - self.bt[s]) + int(self.b[s + 1] - self.b[s])
                     for s in range(W)]
            return self._all_to_all(send, [send.numel() if s == self.dst else 0 for s in range(W)],
                                    [sizes[s] if r == self.dst else 0 for s in range(W)], async_op=True) + (sizes, send)
        if do_p:
            # one exchange buffer for both: [real | pseudo] E[w] of my genes, written by the kernels themselves
            both = t.empty((n_lr + n_lp, L), dtype=t.float64, device=eng.device)
            eng.fork()
            eng.wait(self._ev_ctrl, side=True)
            eng.wait(self._ev_all)
            real = eng.em(self.plan_real, self.d_test, self.d_ctrl, self.params, self.d_fx1, self.d_fx2, want_x=True, w1_out=both[:n_lr])
            if W > 1:
                early = gather_em(real)
            # beside the real genes, on the side stream: no kernel tail between the two launches
            pseudo = eng.em(self.plan_pseudo, self.d_test, self.d_ctrl, self.params, want_x=False, w1_out=both[n_lr:], side=True)
            eng.join()
            gb = np.concatenate([[0], np.cumsum([(self.b[s + 1] - self.b[s]) + (self.bp[s + 1] - self.bp[s]) for s in range(W)])])
            mine = self._lines_of_all_genes(both, gb)                        # [(real_s, pseudo_s) for s in ranks, my lines]
            parts_r, parts_p, o = [], [], 0
            for s in range(W):
                nr, npz = int(self.b[s + 1] - self.b[s]), int(self.bp[s + 1] - self.bp[s])
                parts_r.append(mine[o:o + nr]); o += nr
                parts_p.append(mine[o:o + npz]); o += npz
            g_block = t.cat(parts_r) if W > 1 else parts_r[0]                # [n_genes, my lines], caller order
            p_block = t.cat(parts_p) if W > 1 else parts_p[0]                # [n_pseudo, my lines], original order
            pv_block = eng.kde(g_block.contiguous(), p_block.contiguous(), self.positive)
        else:
            eng.wait(self._ev_all)
            real = eng.em(self.plan_real, self.d_test, self.d_ctrl, self.params, self.d_fx1, self.d_fx2, want_x=True)
            if W > 1:
                early = gather_em(real)
        if W == 1:
            res = {k: real[k] for k in ('w1', 'w2', 'x1', 'x2', 'n_iters')}
            if do_p:
                res['pvals'] = pv_block
            self.exchange_bytes = 0
            return res
        # late gather: the p-value column blocks (all genes x my lines)
        work, recv_em, em_sizes, em_send = early
        recv_pv = None
        if do_p:
            send = pv_block.reshape(-1)
            pv_sizes = [self.n_genes * int(self.lb[s + 1] - self.lb[s]) for s in range(W)]
            recv_pv = self._all_to_all(send, [send.numel() if s == self.dst else 0 for s in range(W)],
                                       [pv_sizes[s] if r == self.dst else 0 for s in range(W)])
        if work is not None:
            work.wait()
        self.exchange_bytes = (both.numel() * 8 * (W - 1) // W if do_p else 0) + \
            ((em_send.numel() + (pv_block.numel() if do_p else 0)) * 8 if r != self.dst else 0)
        if r != self.dst:
            return None
        out = {k: [] for k in ('w1', 'w2', 'x1', 'x2', 'n_iters')}
        o = 0
        for s in range(W):
            n_s, t_s = int(self.b[s + 1] - self.b[s]), int(self.bt[s + 1] - self.bt[s])
            for key, size, shape in (('w1', n_s * L, (n_s, L)), ('w2', n_s * L, (n_s, L)), ('x1', t_s, (t_s,)), ('x2', t_s, (t_s,)),
                                     ('n_iters', n_s, (n_s,))):
                out[key].append(recv_em[o:o + size].view(shape)); o += size
        res = {k: t.cat(out[k]) for k in ('w1', 'w2', 'x1', 'x2')}
        res['n_iters'] = t.cat(out['n_iters']).to(t.int32)
        if do_p:
            blocks, o = [], 0
            for s in range(W):
                l_s = int(self.lb[s + 1] - self.lb[s])
                blocks.append(recv_pv[o:o + self.n_genes * l_s].view(self.n_genes, l_s)); o += self.n_genes * l_s
            res['pvals'] = t.cat(blocks, dim=1)
        return res

    def run_to_host(self, host, want_pvals=True):
        """Steps 2-5 with the results delivered to ``host`` (SharedHostResults) and NO gather: every rank copies the rows
        of its own genes -- E[w], E[w^2], x, pass counts right behind the real-gene EM, travelling under the pseudo-gene
        EM; p-values after a second all-to-all has turned [all genes, my lines] back into [my genes, all lines] -- and a
        barrier tells rank ``owner`` the arrays are complete.  Returns the arrays (numpy views) on every rank."""
        import torch.distributed as dist
        t, eng = self.torch, self.engine
        W, r, L = self.world, self.rank, self.L
        do_p = bool(want_pvals and self.has_pseudo and self.n_pseudo >= 2)
        lo, hi = int(self.b[r]), int(self.b[r + 1])
        tlo, thi = int(self.bt[r]), int(self.bt[r + 1])
        n_lr, n_lp = hi - lo, int(self.bp[r + 1] - self.bp[r])

        def deliver(real):
            eng.to_host(host['w1'][lo:hi], real['w1'])
            eng.to_host(host['w2'][lo:hi], real['w2'])
            eng.to_host(host['x1'][tlo:thi], real['x1'])
            eng.to_host(host['x2'][tlo:thi], real['x2'])
            eng.to_host(host['n_iters'][lo:hi], real['n_iters'])
        if do_p:
            both = t.empty((n_lr + n_lp, L), dtype=t.float64, device=eng.device)
            eng.fork()
            eng.wait(self._ev_ctrl, side=True)
            eng.wait(self._ev_all)
            real = eng.em(self.plan_real, self.d_test, self.d_ctrl, self.params, self.d_fx1, self.d_fx2, want_x=True, w1_out=both[:n_lr])
            deliver(real)
            pseudo = eng.em(self.plan_pseudo, self.d_test, self.d_ctrl, self.params, want_x=False, w1_out=both[n_lr:], side=True)
            eng.join()
            gb = np.concatenate([[0], np.cumsum([(self.b[s + 1] - self.b[s]) + (self.bp[s + 1] - self.bp[s]) for s in range(W)])])
            mine = self._lines_of_all_genes(both, gb)
            parts_r, parts_p, o = [], [], 0
            for s in range(W):
                nr, npz = int(self.b[s + 1] - self.b[s]), int(self.bp[s + 1] - self.bp[s])
                parts_r.append(mine[o:o + nr]); o += nr
                parts_p.append(mine[o:o + npz]); o += npz
            g_block = t.cat(parts_r) if W > 1 else parts_r[0]
            p_block = t.cat(parts_p) if W > 1 else parts_p[0]
            pv_block = eng.kde(g_block.contiguous(), p_block.contiguous(), self.positive)        # [n_genes, my lines]
            # back: rank s receives the rows of ITS genes from every line block
            my_l = int(self.lb[r + 1] - self.lb[r])
            recv = self._all_to_all(pv_block.reshape(-1), [int(self.b[s + 1] - self.b[s]) * my_l for s in range(W)],
                                    [n_lr * int(self.lb[s + 1] - self.lb[s]) for s in range(W)])
            if W > 1:
                cols, o = [], 0
                for s in range(W):
                    l_s = int(self.lb[s + 1] - self.lb[s])
                    cols.append(recv[o:o + n_lr * l_s].view(n_lr, l_s)); o += n_lr * l_s
                pv_mine = t.cat(cols, dim=1)
            else:
                pv_mine = recv.view(n_lr, L)
            eng.to_host(host['pvals'][lo:hi], pv_mine.contiguous())
            self.exchange_bytes = (both.numel() + pv_block.numel()) * 8 * (W - 1) // W
            del pseudo
        else:
            eng.wait(self._ev_all)
            real = eng.em(self.plan_real, self.d_test, self.d_ctrl, self.params, self.d_fx1, self.d_fx2, want_x=True)
            deliver(real)
            self.exchange_bytes = 0
        eng.host_done()
        if W > 1:
            dist.barrier(group=self.group)
        return host.arrays


def run_screen_sharded(offsets, rows, pseudo_offsets, pseudo_rows, testdata, ctrldata, engine=None, group=None, device=None,
                       fixed_x=None, params=None, positive=False, dst=0, want_pvals=True):
    """One-shot form of ShardedScreen: returns numpy arrays on rank ``dst`` (None elsewhere)."""
    job = ShardedScreen(offsets, rows, pseudo_offsets, pseudo_rows, testdata, ctrldata, engine=engine, group=group, device=device,
                        fixed_x=fixed_x, params=params, positive=positive, dst=dst)
    res = job.run(want_pvals=want_pvals)
    if res is None:
        return None
    return {k: v.cpu().numpy() for k, v in res.items()}


# ----------------------------------------------------------------------------------------------------------------------
# Earlier, array-level drivers (kept: the paper-script batch modes use infer_sharded with per_line=True)
# ----------------------------------------------------------------------------------------------------------------------
def _pack(res, n_local, t_local, L):
    return np.concatenate([res['w1'].reshape(-1), res['w2'].reshape(-1), res['x1'].reshape(-1), res['x2'].reshape(-1),
                           res['n_iters'].astype(np.float64).reshape(-1)]).astype(np.float64), (n_local, t_local, L)


def _unpack(buf, n_local, t_local, L, per_line=False):
    """per_line: the per-line mode's outputs carry a line axis on x1, x2 and n_iters as well."""
    o = 0
    out = {}
    K = L if per_line else 1
    xs, its = ((t_local, L), (n_local, L)) if per_line else ((t_local,), (n_local,))
    for key, size, shape in (('w1', n_local * L, (n_local, L)), ('w2', n_local * L, (n_local, L)), ('x1', t_local * K, xs),
                             ('x2', t_local * K, xs), ('n_iters', n_local * K, its)):
        out[key] = buf[o:o + size].reshape(shape)
        o += size
    out['n_iters'] = out['n_iters'].astype(np.int32)
    return out


def infer_sharded(offsets, rows, testdata, ctrldata, compute=None, group=None, device=None, per_line=False, **em_kwargs):
    """Batched EM of the whole gene index, sharded by gene over the ranks of ``group`` (default: WORLD), host arrays in,
    host arrays out on every rank (one all-gather of the packed results).
    per_line: run the per-line ("single JACKS") mode instead of the joint one -- the batch driver of the paper
    scripts (jacks_b200/paper_modes.py); x1, x2 and n_iters then carry a line axis.

    ``compute(offsets, rows, testdata, ctrldata, **em_kwargs) -> dict(w1, w2, x1, x2, n_iters)`` runs one shard;
    by default the GPU path of this rank.  em_kwargs: n_iter, fixed_x, apply_w_hp and the hyper-parameters of
    jacks_em_default_params."""
    import torch
    import torch.distributed as dist
    world, rank = _world(group)
    offsets = np.asarray(offsets, dtype=np.int64)
    b = shard_bounds(offsets, world)
    lo, hi = int(b[rank]), int(b[rank + 1])
    loffs, lrows = local_index(offsets, rows, lo, hi)
    dev_index = _resolve_device(device, group)
    if compute is None:
        from . import infer, _native
        if dev_index is None:
            raise RuntimeError('jacks_b200.dist: no CUDA device for this rank and no compute given (there is no CPU fallback)')

        def compute(o, r, t, c, **kw):
            if per_line:
                fixed = kw.pop('fixed_x', None)
                kw.pop('apply_w_hp', None)                       # infer.py:92 never applies with one line
                params = _native.default_params(n_iter=int(kw.pop('n_iter', 50)), **kw)
                fx1 = None if fixed is None else np.asarray(fixed['X1'], dtype=np.float64)
                fx2 = None if fixed is None else np.asarray(fixed['X2'], dtype=np.float64)
                return _native.em_lines_host(_native.context(dev_index), o, r, t, c, params=params, fixed_x1=fx1, fixed_x2=fx2,
                                             want=('x1', 'x2', 'w1', 'w2', 'n_iters'))
            return infer.infer_batched(o, r, t, c, want=('x1', 'x2', 'w1', 'w2'), device=dev_index, **kw)
    L = int(np.asarray(testdata).shape[1])
    K = L if per_line else 1
    if hi > lo:
        res = compute(loffs, lrows, testdata, ctrldata, **em_kwargs)
    else:
        xs, its = ((0, L), (0, L)) if per_line else ((0,), (0,))
        res = dict(w1=np.zeros((0, L)), w2=np.zeros((0, L)), x1=np.zeros(xs), x2=np.zeros(xs), n_iters=np.zeros(its, np.int32))
    if world == 1:
        return {k: np.asarray(res[k]) for k in ('w1', 'w2', 'x1', 'x2', 'n_iters')}
    # one collective: every rank contributes its packed block, padded to the largest block
    sizes = [(int(b[r + 1] - b[r]), int(offsets[b[r + 1]] - offsets[b[r]])) for r in range(world)]
    length = lambda n, t: 2 * n * L + 2 * t * K + n * K
    cap = max(length(n, t) for n, t in sizes)
    flat, _ = _pack(res, hi - lo, int(loffs[-1]), L)
    dev = torch.device('cuda', dev_index) if (dev_index is not None and dist.get_backend(group) == 'nccl') else torch.device('cpu')
    send = torch.zeros(cap, dtype=torch.float64, device=dev)
    send[:flat.size] = torch.from_numpy(flat).to(dev)
    recv = torch.empty(world * cap, dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(recv, send, group=group)
    recv = recv.cpu().numpy().reshape(world, cap)
    parts = [_unpack(recv[r], sizes[r][0], sizes[r][1], L, per_line) for r in range(world)]
    return {k: np.concatenate([p[k] for p in parts]) for k in ('w1', 'w2', 'x1', 'x2', 'n_iters')}


def pvalues_sharded(w1_genes, w1_pseudo, positive=False, compute=None, group=None, device=None):
    """KDE p-values of every gene against the pseudo-gene null (computeW1PvalsAndFDRs with pseudo=True,
    jacks_io.py:69-89) from HOST arrays every rank already holds, sharded by gene; one all-gather reassembles [n, L].
    (The screen job above shards this stage by LINE behind an all-to-all instead and never leaves the device.)
    ``compute(w1_block, w1_pseudo, positive) -> [m, L]`` evaluates one block; by default the GPU path of this rank."""
    import torch
    import torch.distributed as dist
    world, rank = _world(group)
    w1_genes = np.ascontiguousarray(w1_genes, dtype=np.float64)
    n, L = w1_genes.shape
    dev_index = _resolve_device(device, group)
    if compute is None:
        from . import _native
        if dev_index is None:
            raise RuntimeError('jacks_b200.dist: no CUDA device for this rank and no compute given (there is no CPU fallback)')

        def compute(block, pseudo, pos):
            return _native.kde_pvalues_host(_native.context(dev_index), block, pseudo, positive=pos)
    b = np.linspace(0, n, world + 1).astype(np.int64)
    lo, hi = int(b[rank]), int(b[rank + 1])
    mine = compute(w1_genes[lo:hi], w1_pseudo, positive) if hi > lo else np.zeros((0, L))
    if world == 1:
        return np.asarray(mine)
    cap = int(np.max(np.diff(b))) * L
    dev = torch.device('cuda', dev_index) if (dev_index is not None and dist.get_backend(group) == 'nccl') else torch.device('cpu')
    send = torch.zeros(cap, dtype=torch.float64, device=dev)
    send[:mine.size] = torch.from_numpy(np.ascontiguousarray(mine).reshape(-1)).to(dev)
    recv = torch.empty(world * cap, dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(recv, send, group=group)
    recv = recv.cpu().numpy().reshape(world, cap)
    return np.concatenate([recv[r, :int(b[r + 1] - b[r]) * L].reshape(-1, L) for r in range(world)])
