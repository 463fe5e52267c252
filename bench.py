#!/usr/bin/env python
"""Headline benchmark of jacks_b200: gene-inferences/sec of the batched EM on the Avana shape, ONE screen, 1-8 B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--scaling strong|weak]

The workload is BASELINE.json's configs 4-5: one synthetic Avana-shaped screen -- 18,000 real
genes x 4 gRNAs x 400 cell lines, one control common to all lines -- plus 100,000 pseudo-genes drawn from 1,000 control
genes, i.e. 118,000 gene-inferences per step, and the KDE p-values of the real genes against the pseudo-gene null.

  value     gene-inferences/s of the EM (jacks/jacks/infer.py:18-125; the metric as SURVEY.md 8d defines it) with the
            data cubes resident in HBM: one step = EM of all 118,000 genes, y and tau written for the real genes; with
            N > 1 ranks ("strong", the default) the genes of the ONE screen are sharded over the ranks and the step
            ends with the NCCL gather of the real genes' results on rank 0, inside the timed region.  CUDA events on the
            launching stream, barrier + synchronize on both sides, max over ranks.
  job       the whole north-star job, device resident: EM -> one all-to-all (my genes x all lines -> all genes x my lines)
            -> line-sharded KDE p-values -> gather (jacks_b200/dist.py); ms per step, the collectives' share, bytes over
            NVLink.
  e2e       the same job through the reference-facing entry point on HOST buffers: jacks_run_screen_host (C-ABI) at one
            rank -- pinned cubes, the screen uploaded once, x1, x2, w1, w2, n_iters and the 7.2 M p-values downloaded;
            at N ranks each rank uploads the rows of its shard from pinned memory and rank 0 downloads the results.
            e2e_python: wall time of the drop-in jacks_b200.infer.inferJACKS (pageable numpy arrays, result dict built).
  roofline  the EM kernel against the FP64 FMA-pipe peak measured in this run (MEASURED_PEAKS.json has no FP64 entry):
            algorithmic flops 19*G*L*I + 12*G*L per gene with the iteration counts I the kernel returns,
            over the kernel's own CUDA-event duration; plus the same launch against the measured HBM peak.
  cpu_baseline  the reference's numpy EM on all host cores over a bounded sample of the same genes: the unmodified
            reference when oracle/_ref holds it (kind "reference"), else its bit-exact numpy restatement (kind "port").
  aux       KDE p-values and preprocessing with their own rooflines and CPU baselines; fixed-x / w_only (config 3);
            per-line mode.

--scaling weak: every rank runs its own screen (N independent replicas, no collective).
--impl reference: the reference's EM on the host cores (rank 0 only), same metric and config.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

N_GENES, GUIDES, N_LINES, N_PSEUDO, N_CTRL = 18000, 4, 400, 100000, 1000
METRIC = 'gene-inferences/sec, Avana shape (18k genes x 4 gRNA x 400 lines + 1e5 pseudo-genes)'
UNIT = 'gene-inferences/s'


def build_workload(seed_offset=0, n_genes=N_GENES, n_pseudo=N_PSEUDO, n_lines=N_LINES, guide_counts=None, n_ctrl=N_CTRL):
    from jacks_b200 import synth
    from jacks_b200._native import as_csr
    gi, test, ctrl, ctrl_genes = synth.make_condensed_screen(n_genes=n_genes, guides_per_gene=GUIDES, n_lines=n_lines,
                                                             seed=synth.AVANA_SEED + seed_offset, n_ctrl_genes=n_ctrl,
                                                             guide_counts=guide_counts)
    names, offs, rows = as_csr(gi)
    pos = {g: i for i, g in enumerate(names)}
    ctrl_ids = [pos[g] for g in ctrl_genes]
    poffs, prows = synth.make_pseudo_index_csr(offs, ctrl_ids, n_pseudo, seed=synth.AVANA_SEED + 1 + seed_offset)
    # the control is one sample common to all lines: the product passes it as [N, 1, 2]
    ctrl1 = np.ascontiguousarray(ctrl[:, :1, :])
    return dict(test=test, ctrl=ctrl, ctrl1=ctrl1, offs=offs, rows=rows.astype(np.int32), poffs=poffs, prows=prows,
                all_offs=np.concatenate([offs, offs[-1] + poffs[1:]]),
                all_rows=np.concatenate([rows.astype(np.int32), prows]))


# ------------------------------------------------------------------------------------------------
# CPU baselines -- run BEFORE any CUDA initialisation (they fork)
# ------------------------------------------------------------------------------------------------
_CPU = {}


def _cpu_worker(job):
    lo, hi = job
    w, sel = _CPU['w'], _CPU['sel']
    offs, rows = w['all_offs'], w['all_rows']
    gi = {int(g): rows[offs[g]:offs[g + 1]].tolist() for g in sel[lo:hi]}
    if _CPU['ref'] is not None:
        RI = _CPU['ref'][0]
        calls = [0]
        orig = RI.lowerBound

        def counted(*a):
            calls[0] += 1
            return orig(*a)
        RI.lowerBound = counted
        try:
            RI.inferJACKS(gi, w['test'], w['ctrl'])          # the unmodified reference, jacks/jacks/infer.py:18
        finally:
            RI.lowerBound = orig
        return calls[0] - len(gi)                             # one lowerBound call before the loop of every gene
    from oracle import jacks_oracle as O
    iters = {}
    O.em_genes(gi, w['test'], w['ctrl'], iters_out=iters, mirror_debug_cost=True)
    return sum(iters.values())


def load_reference():
    try:
        from oracle import make_ref
        return make_ref.import_reference()
    except Exception:
        return None


def load_reference_preprocess():
    """jacks.preprocess of the unmodified reference (oracle/_ref), or None."""
    try:
        import sys as _sys
        from oracle import make_ref
        root = make_ref.make()
        if root is None:
            return None
        saved = {k: _sys.modules.pop(k) for k in list(_sys.modules) if k == 'jacks' or k.startswith('jacks.')}
        _sys.path.insert(0, root)
        try:
            import jacks.infer as RI
            import jacks.preprocess as RP
            assert os.path.abspath(RP.__file__).startswith(os.path.abspath(root)), RP.__file__
        finally:
            _sys.path.remove(root)
            for k in list(_sys.modules):
                if k == 'jacks' or k.startswith('jacks.'):
                    _sys.modules.pop(k)
            _sys.modules.update(saved)
        RI.SP = np
        return RP
    except Exception:
        return None


def cpu_reference_rate(w, cores, genes_per_core, ref):
    """Gene-inferences/s of the reference's EM on `cores` processes (fork, arrays shared copy-on-write)."""
    import multiprocessing as mp
    n_all = len(w['all_offs']) - 1
    n_sample = int(min(n_all, cores * genes_per_core))
    sel = np.linspace(0, n_all - 1, n_sample).astype(np.int64)      # evenly over real and pseudo genes
    _CPU['w'], _CPU['sel'], _CPU['ref'] = w, sel, ref
    bounds = np.linspace(0, n_sample, cores + 1).astype(int)
    jobs = [(int(bounds[i]), int(bounds[i + 1])) for i in range(cores) if bounds[i + 1] > bounds[i]]
    ctx = mp.get_context('fork')
    with ctx.Pool(len(jobs)) as pool:
        pool.map(_cpu_worker, [(0, 1)] * len(jobs))                 # warm the workers (imports, page-in)
        t0 = time.perf_counter()
        tot_iters = sum(pool.map(_cpu_worker, jobs))
        dt = time.perf_counter() - t0
    return n_sample / dt, n_sample, dt, tot_iters / float(n_sample)


def cpu_pvalue_rate(n_pseudo, n_genes=200, n_lines=20):
    """p-values/s of the reference's computeW1PvalsAndFDRs restatement (scipy-free numpy port of gaussian_kde.
    integrate_box_1d, oracle/jacks_oracle.py:kde_pvalues) on a (genes x lines) sub-grid at the FULL pseudo-gene count."""
    from oracle import jacks_oracle as O
    rng = np.random.default_rng(5)
    g = rng.normal(-0.3, 0.7, size=(n_genes, n_lines))
    p = rng.normal(0.0, 0.2, size=(n_pseudo, n_lines))
    t0 = time.perf_counter()
    O.kde_pvalues(g, p)
    dt = time.perf_counter() - t0
    return n_genes * n_lines / dt, dt


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler(object):
    Q = 'index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,' \
        'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '250'], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# ------------------------------------------------------------------------------------------------
def workload_config(args, world):
    strong = args.scaling == 'strong'
    return {'workload': 'synthetic Avana-shape joint EM: %d genes x %d gRNAs x %d lines + %d pseudo-genes, ONE screen%s, '
                        'one control common to all lines passed as [N,1,2]'
                        % (N_GENES, GUIDES, N_LINES, N_PSEUDO, (' sharded by gene over %d ranks' % world) if strong and world > 1
                           else (' per rank (replicas)' if world > 1 else '')),
            'genes_per_step': (N_GENES + N_PSEUDO) * (1 if strong else world),
            'n_iter': 50, 'tol': 0.1, 'l2_policy': 'inputs (0.46 GB test cube per screen; 0.08 GB per rank at 8 ranks) and outputs '
            '(0.76 GB of E[w], E[w^2] + 0.46 GB of y, tau per step) exceed the 126 MB L2',
            'parallelism': 'gene-sharded data parallel, %d rank(s); all-to-all + gather over NCCL in the job' % world}


def run_reference_arm(args, rank):
    """--impl reference: the reference's own EM on the host cores (rank 0 only)."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    w = build_workload()
    ref = load_reference()
    kind = 'reference' if ref is not None else 'port'
    rates = []
    gpc = max(40, int(args.ref_genes_per_core))
    for _ in range(args.warmup):
        cpu_reference_rate(w, cores, max(8, gpc // 8), ref)
    n_sample, mean_it = 0, 0.0
    for _ in range(args.steps):
        rate, n_sample, dt, mean_it = cpu_reference_rate(w, cores, gpc, ref)
        rates.append((rate, dt))
    value = float(np.mean([r for r, _ in rates]))
    sample = '%d genes (evenly spaced over the 118,000 real+pseudo genes of the workload) per step, %d processes' % (n_sample, cores)
    note = ('the unmodified reference (jacks/jacks/infer.py, copied by oracle/make_ref.py into the git-ignored oracle/_ref; only shim: '
            'jacks.infer.SP = numpy), matched control cube [N,L,2] as its API requires' if kind == 'reference' else
            'oracle/jacks_oracle.py: numpy restatement of jacks/jacks/infer.py, bit-exact against the unmodified reference '
            '(tests/test_oracle_golden.py); oracle/_ref is absent on this box')
    out = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
           'warmup': args.warmup, 'ms_per_step': 1e3 * float(np.mean([d for _, d in rates])), 'higher_is_better': True,
           'scaling': args.scaling, 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
           'config': workload_config(args, 1),
           'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': kind, 'sample': sample,
                            'mean_iters': mean_it, 'note': note},
           'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(out), flush=True)


def ncu_traffic():
    """dram bytes (read + write) per launch of the EM kernel from the committed ncu capture, or None."""
    for name in ('ncu_traffic_r02.json', 'ncu_traffic_r01.json'):
        try:
            return json.load(open(os.path.join(REPO, 'profiles', name)))['em_fast_kernel_118k_genes_bytes']
        except Exception:
            continue
    return None


def timed(torch, fn, reps, fence=None):
    """ms per call of fn (CUDA events on the current stream, synchronize on both sides)."""
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    (fence or torch.cuda.synchronize)()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    (fence or torch.cuda.synchronize)()
    return e0.elapsed_time(e1) / reps


def lines_measurement(ctx, torch, dev, stream, w, params):
    """Per-line ("single JACKS") mode on the real genes of the workload, device resident: 18,000 genes x 400 lines =
    7.2 M independent one-line EMs per launch set (run_JACKS_single.py:83-85 makes 400 inferJACKS calls for this)."""
    from jacks_b200 import _native
    offs, rows = w['offs'], w['rows']
    N, L = w['test'].shape[0], w['test'].shape[1]
    n = len(offs) - 1
    d_test = torch.from_numpy(w['test']).to(dev)
    d_ctrl = torch.from_numpy(w['ctrl']).to(dev)
    o_w1 = torch.empty((n, L), dtype=torch.float64, device=dev)
    o_w2 = torch.empty((n, L), dtype=torch.float64, device=dev)
    o_it = torch.empty((n, L), dtype=torch.int32, device=dev)
    plan = _native.EmPlan(ctx, offs, rows, N, L)
    io = _native.EmDeviceIO()
    io.test, io.ctrl, io.ctrl_lines = d_test.data_ptr(), d_ctrl.data_ptr(), L
    io.w1, io.w2, io.n_iters = o_w1.data_ptr(), o_w2.data_ptr(), o_it.data_ptr()
    for _ in range(2):
        plan.run_lines_device(params, io, stream)
    ms = timed(torch, lambda: plan.run_lines_device(params, io, stream), 5)
    it = o_it.cpu().numpy()
    plan.close()
    return {'ms': ms, 'value': n * L / (ms * 1e-3), 'unit': 'one-line gene-inferences/s', 'mean_iters': float(it.mean()),
            'shape': '%d genes x %d lines, device resident' % (n, L),
            'hbm_gbs': 2.0 * N * L * 16 / (ms * 1e-3) / 1e9}


def aux_measurements(ctx, torch, dev, stream, w, params, hbm_peak, cpu_pv):
    """The other subsystems of the hot path at the BASELINE shape with their own rooflines (all HBM-bound streaming + sort
    work: algorithmic bytes = inputs read once + outputs written once) and CPU baselines."""
    import ctypes as C
    from jacks_b200 import _native
    out = {}
    L = N_LINES
    lib = _native.lib()
    # ---- KDE p-values: 18,000 genes x L lines against 1e5 pseudo-genes (computeW1PvalsAndFDRs, jacks_io.py:69-89)
    g = torch.randn((N_GENES, L), dtype=torch.float64, device=dev) * 0.7 - 0.3
    pz = torch.randn((N_PSEUDO, L), dtype=torch.float64, device=dev) * 0.2
    pv = torch.empty_like(g)

    def kde():
        _native.check(lib.jacks_kde_pvalues_device(ctx.h, g.data_ptr(), N_GENES, pz.data_ptr(), N_PSEUDO, L, 0, pv.data_ptr(),
                                                   C.c_void_p(stream or None)))
    kde()
    ms = timed(torch, kde, 3)
    kbytes = 8.0 * (N_PSEUDO * L + 2 * N_GENES * L)
    out['kde_pvalues'] = {'value': N_GENES * L / (ms * 1e-3), 'unit': 'p-values/s', 'ms': ms,
                          'shape': '%d genes x %d lines vs %d pseudo-genes, device resident' % (N_GENES, L, N_PSEUDO),
                          'pairs_per_s': N_GENES * L * float(N_PSEUDO) / (ms * 1e-3),
                          'roofline': {'bound': 'hbm', 'achieved': kbytes / (ms * 1e-3) / 1e9, 'peak': hbm_peak, 'unit': 'GB/s',
                                       'frac': kbytes / (ms * 1e-3) / 1e9 / hbm_peak,
                                       'algorithmic_bytes': kbytes, 'note': 'pseudo and gene E[w] read once, p-values written once'},
                          'cpu_baseline': cpu_pv}
    del g, pz, pv
    # ---- preprocessing: 72,000 guides x 802 replicate columns -> 401 samples (preprocess.py:25-96)
    rng = np.random.default_rng(2)
    N, Cc = N_GENES * GUIDES, 2 * (L + 1)
    counts = rng.poisson(300.0, size=(N, Cc)).astype(np.float64)
    _native.normalize_host(ctx, counts[:4096], 'median', take_log=True, prior=32.0)
    t0 = time.perf_counter()
    lc = _native.normalize_host(ctx, counts, 'median', take_log=True, prior=32.0)
    t1 = time.perf_counter()
    data = _native.condense_host(ctx, lc, [[2 * s_, 2 * s_ + 1] for s_ in range(Cc // 2)])
    t2 = time.perf_counter()
    pre = {'normalize_s': t1 - t0, 'condense_s': t2 - t1, 'unit': 's (host entry points, pageable buffers)',
           'shape': '%d guides x %d replicate columns -> %d samples' % (N, Cc, Cc // 2),
           'guide_samples_per_s': N * (Cc // 2) / (t2 - t0)}
    if hasattr(_native, 'preprocess_device_ms'):
        dms = _native.preprocess_device_ms(ctx, counts, [[2 * s_, 2 * s_ + 1] for s_ in range(Cc // 2)], prior=32.0)
        pbytes = 8.0 * N * Cc + 16.0 * N * (Cc // 2)
        pre['device_ms'] = dms
        pre['roofline'] = {'bound': 'hbm', 'achieved': pbytes / (dms * 1e-3) / 1e9, 'peak': hbm_peak, 'unit': 'GB/s',
                           'frac': pbytes / (dms * 1e-3) / 1e9 / hbm_peak, 'algorithmic_bytes': pbytes,
                           'note': 'counts read once, [N,S,2] cube written once; device-resident chain normalise -> condense'}
    # the one-upload chain of the command line (jacks_preprocess_host: counts up once, the [N,S,2] cube down once) on pinned
    # buffers: PCIe floor = 0.46 GB each way around the device time above (the column medians and the per-sample sorts
    # need every row, so neither copy can overlap the kernels)
    pin_in, pin_out = _native.PinnedArray((N, Cc)), _native.PinnedArray((N, Cc // 2, 2))
    pin_in.array[...] = counts
    offs, cols = _native._sample_csr([[2 * s_, 2 * s_ + 1] for s_ in range(Cc // 2)])
    table = _native._log2_table(counts, 32.0)

    def chain_host():
        _native.check(lib.jacks_preprocess_host(ctx.h, _native._ptr(pin_in.array), N, Cc, 1, 32.0, _native._ptr(table), table.size,
                                                _native.NORM_TYPES['median'], None, None, _native._ptr(offs), _native._ptr(cols), Cc // 2, 0, 1,
                                                None, _native._ptr(pin_out.array)))
    chain_host()
    t0 = time.perf_counter()
    for _ in range(3):
        chain_host()
    pre['host_chain_ms'] = (time.perf_counter() - t0) / 3 * 1e3
    pre['host_chain'] = 'jacks_preprocess_host, pinned buffers: %.0f MB up, %.0f MB down per call' % (pin_in.array.nbytes / 1e6, pin_out.array.nbytes / 1e6)
    assert np.array_equal(pin_out.array, data, equal_nan=True)
    # CPU baseline: the unmodified reference's own functions (preprocess.py:77-96 normalizeLogCounts, :61-67
    # condense_normalised_counts -> calc_posterior_sd) on ALL guides of the first few samples, one core (the reference's
    # preprocessing is a serial Python loop over samples), checked against the GPU cube; else the numpy restatement
    n_s = 6
    sub = counts[:, :2 * n_s]
    ref_pre = load_reference_preprocess()
    t0 = time.perf_counter()
    if ref_pre is not None:
        lc_ref = ref_pre.normalizeLogCounts(np.log2(sub + 32.0), normtype='median')
        cube_ref = ref_pre.condense_normalised_counts(lc_ref, [[2 * s_, 2 * s_ + 1] for s_ in range(n_s)], list(range(n_s)))
        kind = 'reference'
    else:
        from oracle import jacks_oracle as O
        lc_ref = O.normalize_log_counts(O.log_counts(sub, 32), 'median')
        cube_ref = O.condense(lc_ref, [[2 * s_, 2 * s_ + 1] for s_ in range(n_s)])
        kind = 'port'
    dt = time.perf_counter() - t0
    err = np.abs(cube_ref[:, :, 1] - data[:, :n_s, 1]) / np.abs(cube_ref[:, :, 1])
    assert np.abs(cube_ref[:, :, 0] - data[:, :n_s, 0]).max() < 1e-12 and err.max() < 1e-9, (err.max(),)
    pre['cpu_baseline'] = {'value': N * n_s / dt, 'unit': 'guide-samples/s', 'cores': 1, 'kind': kind, 'seconds': dt,
                           'sample': 'all %d guides of the first %d samples (%d replicate columns) of the same count matrix: log2, '
                                     'median normalisation, mean + posterior sd per sample; outputs checked against the GPU cube' % (N, n_s, 2 * n_s)}
    if 'device_ms' in pre:
        pre['device_guide_samples_per_s'] = N * (Cc // 2) / (pre['device_ms'] * 1e-3)
    pin_in.free(); pin_out.free()
    out['preprocess'] = pre
    assert np.isfinite(data).all()
    return out


def fixed_x_measurement(ctx, torch, dev, stream, w, params):
    """BASELINE config 3: reference-efficacy mode (--reffile / fixed_x, infer.py:22-24,77-79,91: the X step is skipped) on
    the real genes of the Avana-shaped screen, device resident.  The published Avana efficacies do not travel to the GPU
    box; x is drawn from their distribution (mean 1, sd 0.4) -- the arithmetic does not depend on the values."""
    from jacks_b200 import _native
    offs, rows = w['offs'], w['rows']
    N, L = w['test'].shape[0], w['test'].shape[1]
    n = len(offs) - 1
    rng = np.random.default_rng(11)
    x1 = rng.normal(1.0, 0.4, size=N)
    d_x1, d_x2 = torch.from_numpy(x1).to(dev), torch.from_numpy(x1 * x1 + 0.05).to(dev)
    o_w1 = torch.empty((n, L), dtype=torch.float64, device=dev)
    o_w2 = torch.empty((n, L), dtype=torch.float64, device=dev)
    o_it = torch.empty(n, dtype=torch.int32, device=dev)
    plan = _native.EmPlan(ctx, offs, rows, N, L)
    io = _native.EmDeviceIO()
    d_test, d_ctrl = torch.from_numpy(w['test']).to(dev), torch.from_numpy(w['ctrl1']).to(dev)      # caller row order
    io.test, io.ctrl, io.ctrl_lines = d_test.data_ptr(), d_ctrl.data_ptr(), int(d_ctrl.shape[1])
    io.fixed_x1, io.fixed_x2 = d_x1.data_ptr(), d_x2.data_ptr()
    io.w1, io.w2, io.n_iters = o_w1.data_ptr(), o_w2.data_ptr(), o_it.data_ptr()
    for _ in range(2):
        plan.run_device(params, io, stream)
    ms = timed(torch, lambda: plan.run_device(params, io, stream), 5)
    it = o_it.cpu().numpy().astype(np.float64)
    flops = float(((14.0 * it + 12.0) * GUIDES * L).sum())
    plan.close()
    return {'ms': ms, 'value': n / (ms * 1e-3), 'unit': UNIT, 'mean_iters': float(it.mean()),
            'shape': '%d genes x %d gRNAs x %d lines, x fixed (w_only / --reffile), device resident' % (n, GUIDES, L),
            'algorithmic_tflops': flops / (ms * 1e-3) / 1e12,
            'reference_rate_note': 'the reference runs this mode at ~225 genes/s per core (SURVEY.md 8d)'}


def sharded_parity(torch, dist, jd, _native, local_rank, rank, world):
    """N > 1: the sharded screen job against rank 0's unsharded C-ABI call on a reduced, ragged screen -- bit for bit."""
    counts = np.r_[np.full(3000, 4), [1, 2, 3, 5, 6, 9, 12, 13]].astype(np.int64)
    w = build_workload(seed_offset=7, n_pseudo=5000, n_lines=96, guide_counts=counts, n_ctrl=50)
    w['test'][5, 3, 0] = np.nan
    ok = True
    for ctrl in (w['ctrl'], w['ctrl1']):
        job = jd.ShardedScreen(w['offs'], w['rows'], w['poffs'], w['prows'], w['test'], ctrl, device=local_rank)
        res = job.run()
        torch.cuda.synchronize()
        if rank == 0:
            ref = _native.run_screen_host(_native.context(local_rank), w['offs'], w['rows'], w['poffs'], w['prows'], w['test'], ctrl,
                                          want=('x1', 'x2', 'w1', 'w2', 'n_iters', 'pvals'))
            for k in ('x1', 'x2', 'w1', 'w2', 'n_iters', 'pvals'):
                same = bool(np.array_equal(res[k].cpu().numpy(), ref[k], equal_nan=True))
                if not same:
                    sys.stderr.write('sharded_parity: gathered %s differs from the unsharded call\n' % k)
                ok = ok and same
        # ... and the gather-free delivery into shared host memory (the e2e path)
        host = jd.SharedHostResults(job.n_genes, job.T, job.L)
        arr = job.run_to_host(host)
        if rank == 0:
            for k in ('x1', 'x2', 'w1', 'w2', 'n_iters', 'pvals'):
                same = bool(np.array_equal(arr[k], ref[k], equal_nan=True))
                if not same:
                    bad = np.argwhere(~((arr[k] == ref[k]) | (np.isnan(arr[k]) & np.isnan(ref[k]))))
                    sys.stderr.write('sharded_parity: shared-host %s differs from the unsharded call at %d places, first %s\n'
                                     % (k, len(bad), bad[:3].tolist()))
                ok = ok and same
        del arr
        dist.barrier()
        host.close()
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=torch.device('cuda', local_rank))
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.broadcast(flag, src=0)
    return bool(flag.item())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--scaling', default='strong', choices=['weak', 'strong'])
    ap.add_argument('--ref-genes-per-core', dest='ref_genes_per_core', type=int, default=1200)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-aux', action='store_true')
    ap.add_argument('--parity-only', action='store_true')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup

    rank = int(os.environ.get('RANK', 0))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    if args.impl == 'reference':
        run_reference_arm(args, rank)
        return
    strong = args.scaling == 'strong'

    # ---- workload (host) and the CPU baselines, before CUDA is touched ---------------------------------
    w = build_workload(seed_offset=0 if strong else rank)
    cpu = cpu_pv = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        ref = load_reference()
        rate, n_sample, dt, mean_it = cpu_reference_rate(w, cores, max(40, args.ref_genes_per_core), ref)
        cpu = {'value': rate, 'unit': UNIT, 'cores': cores, 'kind': 'reference' if ref is not None else 'port', 'seconds': dt,
               'mean_iters': mean_it,
               'sample': '%d genes evenly spaced over the 118,000 real+pseudo genes of the workload, one process per core'
                         % n_sample}
        pv_rate, pv_dt = cpu_pvalue_rate(N_PSEUDO)
        cpu_pv = {'value': pv_rate, 'unit': 'p-values/s', 'cores': 1, 'kind': 'port', 'seconds': pv_dt,
                  'sample': '200 genes x 20 lines against 100,000 pseudo-genes (numpy restatement of gaussian_kde.integrate_box_1d, '
                            'one core; the reference makes one scipy call per gene and line)'}

    import torch
    import torch.distributed as dist
    from jacks_b200 import _native, dist as jd
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    ctx = _native.context(local_rank)
    stream = torch.cuda.current_stream().cuda_stream
    params = _native.default_params()
    L = N_LINES
    f64 = dict(dtype=torch.float64, device=dev)

    def fence():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    parity = None
    if world > 1 and strong:
        parity = sharded_parity(torch, dist, jd, _native, local_rank, rank, world)
        if args.parity_only:
            if rank == 0:
                print(json.dumps({'sharded_parity': parity}), flush=True)
            dist.destroy_process_group()
            sys.exit(0 if parity else 3)

    # ---- this rank's share of the screen, resident on the device -------------------------------------------------------
    # strong: gene blocks of ONE screen, each rank holding only the rows its genes reference (dist.ShardedScreen);
    # weak / one rank: the whole screen
    job = jd.ShardedScreen(w['offs'], w['rows'], w['poffs'], w['prows'], w['test'], w['ctrl1'], device=local_rank,
                           shard=(strong and world > 1))
    plan_r, plan_p = job.plan_real, job.plan_pseudo
    n_r, n_p, T_r = plan_r.n_genes, plan_p.n_genes, plan_r.total_guides
    d_test, d_ctrl = job.d_test, job.d_ctrl
    o = dict(w1=torch.empty((n_r, L), **f64), w2=torch.empty((n_r, L), **f64), x1=torch.empty(T_r, **f64), x2=torch.empty(T_r, **f64),
             y=torch.empty((T_r, L), **f64), tau=torch.empty((T_r, L), **f64), it=torch.empty(n_r, dtype=torch.int32, device=dev),
             pw1=torch.empty((n_p, L), **f64), pw2=torch.empty((n_p, L), **f64), pit=torch.empty(n_p, dtype=torch.int32, device=dev))
    io_r, io_p = _native.EmDeviceIO(), _native.EmDeviceIO()
    for io in (io_r, io_p):
        io.test, io.ctrl, io.ctrl_lines = d_test.data_ptr(), d_ctrl.data_ptr(), int(d_ctrl.shape[1])
    io_r.w1, io_r.w2, io_r.x1, io_r.x2, io_r.y, io_r.tau, io_r.n_iters = (o[k].data_ptr() for k in ('w1', 'w2', 'x1', 'x2', 'y', 'tau', 'it'))
    io_p.w1, io_p.w2, io_p.n_iters = o['pw1'].data_ptr(), o['pw2'].data_ptr(), o['pit'].data_ptr()
    fp64_peak = ctx.fp64_peak_flops()
    gather_w = world if (strong and world > 1) else 1
    ev_k0, ev_k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def gather_results():
        """The north star's one collective: rank 0 receives the real genes' w1, w2, x1, x2 (and iteration counts).
        Asynchronous: NCCL waits for the real-gene EM just enqueued and moves the bytes while the pseudo-gene EM runs."""
        send = torch.cat([o['w1'].reshape(-1), o['w2'].reshape(-1), o['x1'], o['x2'], o['it'].to(torch.float64)])
        sizes = [2 * int(job.b[s + 1] - job.b[s]) * L + 2 * int(job.bt[s + 1] - job.bt[s]) + int(job.b[s + 1] - job.b[s]) for s in range(world)]
        recv = torch.empty(sum(sizes) if rank == 0 else 0, **f64)
        work = dist.all_to_all_single(recv, send, output_split_sizes=sizes if rank == 0 else [0] * world,
                                      input_split_sizes=[send.numel()] + [0] * (world - 1), async_op=True)
        return work, recv, send

    side = torch.cuda.Stream(device=dev)
    cur = torch.cuda.current_stream()

    def step():
        ev_k0.record()
        side.wait_stream(cur)
        plan_r.run_device(params, io_r, stream)              # the real genes first: their results travel ...
        pending = gather_results() if gather_w > 1 else None
        if n_p:
            # ... while the pseudo-genes (85 % of the arithmetic) run -- on a second stream, so that their CTAs move in
            # as the real-gene kernel's drain (both kernels are persistent: no tail between the two launches)
            plan_p.run_device(params, io_p, side.cuda_stream)
        cur.wait_stream(side)
        ev_k1.record()
        if pending is not None:
            pending[0].wait()                                # the step ends when rank 0 holds the gathered results

    for _ in range(args.warmup):
        step()
    fence()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    launches0 = ctx.launch_count
    ms = timed(torch, step, args.steps, fence) * args.steps
    launches = ctx.launch_count - launches0
    torch.cuda.synchronize()
    ms_kernels = ev_k0.elapsed_time(ev_k1)               # the EM launches of the last step, on their own
    iters = np.concatenate([o['it'].cpu().numpy(), o['pit'].cpu().numpy()]).astype(np.float64)
    G_of = np.concatenate([np.diff(plan_r.offsets), np.diff(plan_p.offsets)]).astype(np.float64)
    flops_step = float(((19.0 * iters + 12.0) * G_of * L).sum())
    # algorithmic HBM bytes: test rows 16*G*L (+16*G control, one column) read per gene; written: x1, x2
    # (16*G), w1, w2 (16*L) and, for the real genes, y and tau (16*G*L)
    bytes_step = float((16.0 * G_of * L + 32.0 * G_of + 16.0 * L).sum() + 16.0 * float(T_r) * L)
    t = torch.tensor([ms, ms_kernels], **f64)
    f = torch.tensor([flops_step, float(n_r + n_p), bytes_step], **f64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(f, op=dist.ReduceOp.SUM)
    ms_max, ms_k_max = (float(x) for x in t.tolist())
    flops_all, genes_all, bytes_all = (float(x) for x in f.tolist())
    ms_per_step = ms_max / args.steps
    value = genes_all / (ms_per_step * 1e-3)

    # ---- the whole job, device resident: EM -> all-to-all -> line-sharded p-values -> gather ------------------------------
    job_info = None
    if world == 1 or strong:
        for _ in range(2):
            job.run()
        jms = timed(torch, lambda: job.run(), max(3, min(args.steps, 10)), fence)
        # collectives alone (same buffers, no kernels in between): the exchange's share of the step
        coll_ms = nv_bytes = 0.0
        if world > 1:
            both = torch.cat([o['w1'], o['pw1']])
            gb = np.concatenate([[0], np.cumsum([(job.b[s + 1] - job.b[s]) + (job.bp[s + 1] - job.bp[s]) for s in range(world)])])
            pvb = torch.empty((job.n_genes, int(job.lb[rank + 1] - job.lb[rank])), **f64)

            def colls():
                job._lines_of_all_genes(both, gb)
                send = torch.cat([o['w1'].reshape(-1), o['w2'].reshape(-1), o['x1'], o['x2'], o['it'].to(torch.float64), pvb.reshape(-1)])
                sizes = [2 * int(job.b[s + 1] - job.b[s]) * L + 2 * int(job.bt[s + 1] - job.bt[s]) + int(job.b[s + 1] - job.b[s]) +
                         job.n_genes * int(job.lb[s + 1] - job.lb[s]) for s in range(world)]
                recv = torch.empty(sum(sizes) if rank == 0 else 0, **f64)
                dist.all_to_all_single(recv, send, output_split_sizes=sizes if rank == 0 else [0] * world,
                                       input_split_sizes=[send.numel()] + [0] * (world - 1))
            colls()
            coll_ms = timed(torch, colls, 5, fence)
            nv_bytes = float(both.numel() * 8 * (world - 1) / world + (0 if rank == 0 else 8 * (2 * n_r * L + 2 * T_r + n_r + pvb.numel())))
        tj = torch.tensor([jms, coll_ms, nv_bytes], **f64)
        if world > 1:
            tjm = tj.clone()
            dist.all_reduce(tjm, op=dist.ReduceOp.MAX)
            dist.all_reduce(tj, op=dist.ReduceOp.SUM)
            jms, coll_ms, nv_bytes = float(tjm[0]), float(tjm[1]), float(tj[2])
        job_info = {'ms_per_step': jms, 'value': (N_GENES + N_PSEUDO) / (jms * 1e-3), 'unit': UNIT,
                    'pvalues_per_step': N_GENES * L,
                    'stages': 'EM of real + pseudo genes -> all-to-all (E[w]: my genes x all lines -> all genes x my lines) -> '
                              'KDE p-values of the real genes on my lines -> gather of w1, w2, x1, x2, n_iters, p-values on rank 0',
                    'collectives_ms': coll_ms, 'collectives_share': (coll_ms / jms) if jms else None,
                    'nvlink_bytes_per_step': nv_bytes}

    # ---- end to end through the reference-facing entry point, HOST buffers -----------------------------------------------
    e2e_steps = max(2, min(args.steps, 5))
    e2e_python = None
    if world == 1:
        pin_test = _native.PinnedArray(w['test'].shape); pin_test.array[...] = w['test']
        pin_ctrl = _native.PinnedArray(w['ctrl1'].shape); pin_ctrl.array[...] = w['ctrl1']
        Tr = int(w['offs'][-1])
        outs = {k: _native.PinnedArray(s, dt) for k, (s, dt) in dict(
            x1=((Tr,), np.float64), x2=((Tr,), np.float64), w1=((N_GENES, L), np.float64), w2=((N_GENES, L), np.float64),
            n_iters=((N_GENES,), np.int32), pvals=((N_GENES, L), np.float64)).items()}
        n_used = int(np.unique(w['prows']).size)
        h2d = pin_test.array.nbytes + pin_ctrl.array.nbytes + n_used * (L + 1) * 16 + \
            w['offs'].nbytes + w['rows'].nbytes + w['poffs'].nbytes + w['prows'].nbytes
        d2h = sum(v.array.nbytes for v in outs.values())

        def e2e_step():
            _native.run_screen_host(ctx, w['offs'], w['rows'], w['poffs'], w['prows'], pin_test.array, pin_ctrl.array, params=params,
                                    want=tuple(outs), out={k: v.array for k, v in outs.items()})
        api = 'jacks_run_screen_host (C-ABI): pinned host cubes, control as [N,1,2]; real + pseudo EM + 7.2 M p-values; ' \
              'x1, x2, w1, w2, n_iters, pvals downloaded'
    else:
        # N ranks: each uploads the rows of its shard (pinned) and delivers the rows of its own genes into ONE shared,
        # page-locked host segment that rank 0 reads in place (dist.SharedHostResults): no gather, eight PCIe links
        # (control-guide rows: 1/N of them per rank, completed by an all-gather over NVLink, so that the pseudo-gene EM runs
        # while the real-gene rows are still arriving)
        pins = {k: torch.from_numpy(v).pin_memory() for k, v in job.host_parts(w['test'], w['ctrl1']).items()}
        host = jd.SharedHostResults(job.n_genes, job.T, L)
        h2d = sum(p.numel() * 8 for p in pins.values())
        n_mine, t_mine = int(job.b[rank + 1] - job.b[rank]), int(job.bt[rank + 1] - job.bt[rank])
        d2h = 8 * (3 * n_mine * L + 2 * t_mine) + 4 * n_mine

        def e2e_step():
            job.upload(pins)
            job.run_to_host(host)
        api = 'jacks_b200.dist.ShardedScreen.run_to_host: every rank uploads the rows of its gene blocks from pinned memory, EM -> ' \
              'all-to-all -> p-values -> all-to-all back, and copies w1, w2, x1, x2, n_iters, pvals of ITS genes into one shared ' \
              'page-locked host segment (no gather; rank 0 reads the assembled arrays in place)'
    e2e_step()
    fence()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    te = torch.tensor([e2e_s, float(h2d), float(d2h)], **f64)
    if world > 1:
        tem = te.clone()
        dist.all_reduce(tem, op=dist.ReduceOp.MAX)
        dist.all_reduce(te, op=dist.ReduceOp.SUM)
        e2e_s, h2d, d2h = float(tem[0]), float(te[1]), float(te[2])
    e2e_genes = (N_GENES + N_PSEUDO) * (1 if (strong or world == 1) else world)
    e2e_value = e2e_genes / e2e_s
    clk = clocks.stop() if rank == 0 else None      # sampled over the timed regions (device-resident, job, e2e)

    if world == 1:
        # the drop-in Python call on pageable arrays: inferJACKS(gene_index, testdata, ctrldata) -> {gene: 6-tuple}
        from jacks_b200 import infer
        gi = {'G%05d' % i: list(range(int(w['offs'][i]), int(w['offs'][i + 1]))) for i in range(N_GENES)}
        ctrl_b = np.broadcast_to(w['ctrl1'], w['test'].shape)         # what collateTestControlSamples hands over for one control
        infer.inferJACKS(gi, w['test'], ctrl_b)
        t0 = time.perf_counter()
        res = infer.inferJACKS(gi, w['test'], ctrl_b)
        e2e_python = {'seconds': time.perf_counter() - t0, 'call': 'jacks_b200.infer.inferJACKS on the 18,000 real genes, pageable '
                      'numpy arrays, y/tau/x1/x2/w1/w2 per gene in the returned dict', 'genes': len(res)}
        del res, gi

    aux = None
    if rank == 0 and world == 1 and not args.no_aux:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(REPO, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        aux = aux_measurements(ctx, torch, dev, stream, w, params, float(peaks.get('hbm_gbs', 6650.0)), cpu_pv)
        aux['fixed_x'] = fixed_x_measurement(ctx, torch, dev, stream, w, params)
        del o
        aux['em_lines'] = lines_measurement(ctx, torch, dev, stream, w, params)
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(REPO, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
        hbm_src = 'measured (MEASURED_PEAKS.json)' if 'hbm_gbs' in peaks else 'fallback (B200_PROFILING.md)'
        # roofline of the dominant kernel: algorithmic flops of one rank's launch over that launch's own duration
        tf = flops_all / gather_w / (ms_k_max * 1e-3) / 1e12 if (strong or world == 1) else flops_all / world / (ms_k_max * 1e-3) / 1e12
        per_rank_bytes = bytes_all / (gather_w if strong or world == 1 else world)
        out = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
               'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None,
               'dtype': 'f64', 'data': 'synthetic', 'config': workload_config(args, world),
               'clocks': clk, 'gpu_launches': int(launches),
               'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                       'ms_per_step': 1e3 * e2e_s, 'api': api, 'includes': 'EM of 118,000 genes AND the 7.2 M KDE p-values'},
               'e2e_python': e2e_python, 'job': job_info,
               'roofline': {'bound': 'fp64', 'achieved': tf, 'peak': fp64_peak / 1e12, 'unit': 'TFLOP/s',
                            'frac': tf / (fp64_peak / 1e12), 'traffic': ncu_traffic(),
                            'peak_source': 'FP64 FMA-pipe peak measured in this run (jacks_fp64_peak_flops); '
                                           'MEASURED_PEAKS.json carries no FP64 figure',
                            'kernel': 'em_fast_kernel<G=4,W=1,TM=2>', 'kernel_ms': ms_k_max, 'mean_iters': float(iters.mean()),
                            'flops_per_launch': flops_all / (gather_w if strong or world == 1 else world)},
               'roofline_hbm': {'bound': 'hbm', 'achieved': per_rank_bytes / (ms_k_max * 1e-3) / 1e9, 'peak': hbm_peak,
                                'unit': 'GB/s', 'frac': per_rank_bytes / (ms_k_max * 1e-3) / 1e9 / hbm_peak,
                                'peak_source': hbm_src},
               'cpu_baseline': cpu, 'aux': aux}
        if parity is not None:
            out['sharded_parity'] = parity
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if parity is False:
        sys.exit(3)


if __name__ == '__main__':
    main()
