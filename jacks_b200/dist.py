"""Gene-sharded data parallelism over the GPUs of one box (one process per GPU, torch.distributed).

Genes are independent (jacks/jacks/infer.py:20-29 carries no state from one gene to the next), so the batched
EM shards by gene with NO collective inside the iteration: every rank runs the kernels on a contiguous block of
the gene index, cost-balanced by guide count, and ONE all-gather of the packed per-gene results (w1, w2, x1, x2,
iteration counts) over NCCL / NVLink hands every rank the full answer.  1-rank and N-rank results are bit-identical
because nothing is reduced across ranks.

The data cubes are replicated (each rank uploads what its genes reference; pseudo-genes reference control rows
anywhere in the cube).  Under the gloo backend (CPU tests) tensors stay on the host.
"""
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
    """Batched EM of the whole gene index, sharded by gene over the ranks of ``group`` (default: WORLD).
    per_line: run the per-line ("single JACKS") mode instead of the joint one -- the batch driver of the paper
    scripts (jacks_b200/paper_modes.py); x1, x2 and n_iters then carry a line axis.

    ``compute(offsets, rows, testdata, ctrldata, **em_kwargs) -> dict(w1, w2, x1, x2, n_iters)`` runs one shard;
    by default the GPU path of this rank (jacks_b200.infer.infer_batched on ``device``).  Returns the gathered
    dict (flat arrays in the caller's gene order) on every rank."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    offsets = np.asarray(offsets, dtype=np.int64)
    b = shard_bounds(offsets, world)
    lo, hi = int(b[rank]), int(b[rank + 1])
    loffs, lrows = local_index(offsets, rows, lo, hi)
    if compute is None:
        from . import infer, _native

        def compute(o, r, t, c, **kw):
            if per_line:
                return _native.em_lines_host(_native.context(device or 0), o, r, t, c, want=('x1', 'x2', 'w1', 'w2', 'n_iters'), **kw)
            return infer.infer_batched(o, r, t, c, want=('x1', 'x2', 'w1', 'w2'), device=device or 0, **kw)
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
    on_gpu = dist.get_backend(group) == 'nccl'
    dev = torch.device('cuda', device if device is not None else torch.cuda.current_device()) if on_gpu else torch.device('cpu')
    send = torch.zeros(cap, dtype=torch.float64, device=dev)
    send[:flat.size] = torch.from_numpy(flat).to(dev)
    recv = torch.empty(world * cap, dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(recv, send, group=group)
    recv = recv.cpu().numpy().reshape(world, cap)
    parts = [_unpack(recv[r], sizes[r][0], sizes[r][1], L, per_line) for r in range(world)]
    return {k: np.concatenate([p[k] for p in parts]) for k in ('w1', 'w2', 'x1', 'x2', 'n_iters')}


def pvalues_sharded(w1_genes, w1_pseudo, positive=False, compute=None, group=None, device=None):
    """KDE p-values of every gene against the pseudo-gene null (computeW1PvalsAndFDRs with pseudo=True,
    jacks_io.py:69-89), sharded by gene over the ranks of ``group``.

    Every rank holds the full ``w1_genes [n, L]`` and ``w1_pseudo [n_pseudo, L]`` (what infer_sharded returns), builds
    the per-line kernel density of the null itself -- it is a few hundred kilobytes of state, cheaper to rebuild than to
    ship -- and evaluates a contiguous block of genes; ONE all-gather reassembles ``[n, L]``.  p-values of a gene do not
    depend on the other genes, so the result is bit-identical to the unsharded call.
    ``compute(w1_block, w1_pseudo, positive) -> [m, L]`` evaluates one block; by default the GPU path of this rank."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    w1_genes = np.ascontiguousarray(w1_genes, dtype=np.float64)
    n, L = w1_genes.shape
    if compute is None:
        from . import _native

        def compute(block, pseudo, pos):
            return _native.kde_pvalues_host(_native.context(device or 0), block, pseudo, positive=pos)
    b = np.linspace(0, n, world + 1).astype(np.int64)
    lo, hi = int(b[rank]), int(b[rank + 1])
    mine = compute(w1_genes[lo:hi], w1_pseudo, positive) if hi > lo else np.zeros((0, L))
    if world == 1:
        return np.asarray(mine)
    cap = int(np.max(np.diff(b))) * L
    on_gpu = dist.get_backend(group) == 'nccl'
    dev = torch.device('cuda', device if device is not None else torch.cuda.current_device()) if on_gpu else torch.device('cpu')
    send = torch.zeros(cap, dtype=torch.float64, device=dev)
    send[:mine.size] = torch.from_numpy(np.ascontiguousarray(mine).reshape(-1)).to(dev)
    recv = torch.empty(world * cap, dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(recv, send, group=group)
    recv = recv.cpu().numpy().reshape(world, cap)
    return np.concatenate([recv[r, :int(b[r + 1] - b[r]) * L].reshape(-1, L) for r in range(world)])
