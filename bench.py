#!/usr/bin/env python
"""bench.py -- Curie-depth windows/s (spectrum + Bouligand MAP fit), BASELINE.json config 3:
synthetic 4096 x 4096 grid at 1 km, 200 km windows (201 x 201 nodes, 99 k-bins), 10 km stride
= 390^2 = 152,100 windows.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the unmodified reference on the host cores

One step = one pass of the hot path (gather/taper/2-D DFT/radial bins, then the bounded quasi-Newton fit) over
the whole centroid list.  N > 1 (torchrun, one rank per GPU) is STRONG scaling: the same 152,100 windows at
every N, rank r owns the contiguous block r of the list and holds only the grid rows its windows cover, the
(n,4) results are gathered to rank 0 with one NCCL gather inside the timed step.  The weak-scaling figure (N
copies of the list) is reported beside it (`weak`).

Printed keys follow the driver contract; extra keys: `roofline` (dominant kernel: FP64 flops counted by ncu
against the measured DFMA peak, with the section-8(d) HBM contract figure beside it), `stages` (per-kernel
CUDA-event times), `cpu_baseline` (the reference itself -- baseline/_ref -- on this box's host cores, bounded
sample; the oracle port and the as-shipped-BLAS figure beside it), `e2e` (public API with host buffers: H2D of
the grid rows and centroids, D2H of the results), `mcmc` (BASELINE config 4 at its real chain length on a
bounded window sample per rank, with the reference's own sampler timed on the host), `config5` (the
EMAG2-shaped 21600 x 10800 grid, 100/200/300 km windows, all 1.70 M windows) and `config2` (Tanaka spectrum
sweep), `check` (gathered multi-GPU result == single-GPU result).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

_ENV_ASSHIPPED = dict(os.environ)                        # BLAS threading as the user's shell has it
if not os.environ.get("CB_BENCH_ASSHIPPED"):             # (the as-shipped child keeps the BLAS defaults)
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")  # fair CPU baseline: 1 BLAS thread per worker process
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    os.environ.setdefault("MKL_NUM_THREADS", "1")

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
REF_DIR = os.path.join(ROOT, "baseline", "_ref")

GRID_N = 4096
WINDOW = 200e3
STRIDE = 10e3
SEED = 20261019
METRIC = "curie_depth_windows_per_s"
UNIT = "windows/s"
X0 = (3.0, 1.0, 10.0, 5.0)
MCMC_SCALE = [0.25, 0.1, 1.0, 0.5]


def make_inputs(grid_n):
    from pycurious_b200 import synth
    data = synth.fractal_grid(grid_n, grid_n, seed=SEED)
    return data, synth.grid_extent(grid_n, grid_n)


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


class _HostGrid(object):
    """Just the coordinate attributes create_centroid_list needs (no CUDA)."""

    def __init__(self, data, ext):
        self.xcoords, self.dx = np.linspace(ext[0], ext[1], data.shape[1], retstep=True)
        self.ycoords = np.linspace(ext[2], ext[3], data.shape[0])


def centroid_list(data, ext, window=WINDOW, stride=STRIDE):
    from pycurious_b200.grid import CurieGrid
    return CurieGrid.create_centroid_list(_HostGrid(data, ext), window, stride, stride)


# --------------------------------------------------------------------------- CPU arms
def reference_available():
    return os.path.exists(os.path.join(REF_DIR, "pycurious", "optimise.py"))


def _import_reference():
    """The UNMODIFIED reference package from baseline/_ref (a plain copy of /root/reference/pycurious/*.py made
    by __graft_entry__.build(); its compiled radon extension is the one built from the reference's own sources)."""
    import types
    import warnings
    warnings.simplefilter("ignore")
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    try:
        import pycurious
    except ImportError:                     # radon extension missing: not on this path
        sys.modules["pycurious.radon"] = types.ModuleType("pycurious.radon")
        for m in [m for m in sys.modules if m == "pycurious" or m.startswith("pycurious.") and m != "pycurious.radon"]:
            del sys.modules[m]
        import pycurious
    return pycurious


def ref_windows_per_s(data, ext, xc, yc, nsample, cores):
    """The reference's own fan-out: CurieOptimise(..., max_processors=cores).optimise_routine over an evenly
    spaced sample of the centroid list (fork + Queue(1) feeder, pycurious/optimise.py:386-424)."""
    pyc = _import_reference()
    idx = np.linspace(0, xc.size - 1, nsample).astype(np.int64)
    grid = pyc.CurieOptimise(data, *ext, max_processors=cores)
    grid.optimise_routine(WINDOW, xc[idx[:cores]], yc[idx[:cores]])          # warm: imports, first-call costs
    t0 = time.perf_counter()
    res = grid.optimise_routine(WINDOW, xc[idx], yc[idx])
    dt = time.perf_counter() - t0
    assert len(res) == 4 and len(res[0]) == nsample
    return nsample / dt, dt


def ref_mcmc_steps_per_s(data, ext, xc, yc, nsim, burnin, cores):
    """The reference's metropolis_hastings (optimise.py:511-633), one chain per core through its own
    parallelise_routine, config-4 priors and proposal scales."""
    pyc = _import_reference()
    idx = np.linspace(0, xc.size - 1, cores).astype(np.int64)
    grid = pyc.CurieOptimise(data, *ext, max_processors=cores)
    grid.add_prior(beta=(3.0, 1.0), zt=(1.0, 1.0), dz=(20.0, 10.0), C=(5.0, 5.0))
    t0 = time.perf_counter()
    grid.parallelise_routine(WINDOW, xc[idx], yc[idx], grid.metropolis_hastings, nsim, burnin, MCMC_SCALE)
    dt = time.perf_counter() - t0
    return cores * (nsim + burnin) / dt, dt


_W = {}


def _port_init(data, ext):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import curie_oracle as orc
    _W["orc"] = orc
    _W["spec"] = orc.GridSpec(data, *ext)


def _port_fit(c):
    return _W["orc"].window_fit(_W["spec"], WINDOW, c[0], c[1])


def port_windows_per_s(data, ext, xc, yc, nsample, cores):
    """Oracle port (numpy/scipy restatement of the reference) on `cores` forked workers."""
    import multiprocessing as mp
    idx = np.linspace(0, xc.size - 1, nsample).astype(np.int64)
    work = [(float(xc[i]), float(yc[i])) for i in idx]
    with mp.get_context("fork").Pool(cores, initializer=_port_init, initargs=(data, ext)) as pool:
        pool.map(_port_fit, work[:cores])
        t0 = time.perf_counter()
        pool.map(_port_fit, work, chunksize=max(1, nsample // (cores * 8)))
        dt = time.perf_counter() - t0
    return nsample / dt, dt


def cpu_arm(data, ext, xc, yc, nsample, cores):
    """(value, seconds, kind): the reference itself when baseline/_ref travelled with the tree, else the port."""
    if reference_available():
        v, dt = ref_windows_per_s(data, ext, xc, yc, nsample, cores)
        return v, dt, "reference"
    v, dt = port_windows_per_s(data, ext, xc, yc, nsample, cores)
    return v, dt, "port"


def asshipped_windows_per_s(grid_path, nsample, cores):
    """The same reference run in a child process WITHOUT the BLAS thread caps (as a user's shell has it)."""
    env = dict(_ENV_ASSHIPPED)
    env["CB_BENCH_ASSHIPPED"] = "1"
    try:
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--cpu-arm-child", grid_path,
                              "--cpu-sample", str(nsample)], env=env, capture_output=True, text=True, timeout=300)
        for line in out.stdout.splitlines():
            if line.startswith("{"):
                return json.loads(line)
    except Exception as e:                       # noqa: BLE001
        return {"error": repr(e)}
    return {"error": "no output", "stderr": out.stderr[-300:]}


def run_cpu_arm_child(args):
    from pycurious_b200 import synth
    data = np.load(args.cpu_arm_child)
    ext = synth.grid_extent(*data.shape)
    xc, yc = centroid_list(data, ext)
    cores = host_cores()
    v, dt, kind = cpu_arm(data, ext, xc, yc, args.cpu_sample or cores * 4, cores)
    emit(json.dumps({"value": v, "seconds": dt, "kind": kind, "cores": cores,
                     "blas_env": {k: os.environ.get(k) for k in ("OPENBLAS_NUM_THREADS", "OMP_NUM_THREADS")}}))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    data, ext = make_inputs(args.grid)
    xc, yc = centroid_list(data, ext)
    cores = host_cores()
    nsample = args.cpu_sample or cores * 500           # ~25 ms per window per core -> 10-20 s per step
    vals, kind = [], "port"
    for s in range(args.warmup + args.steps):
        v, dt, kind = cpu_arm(data, ext, xc, yc, nsample, cores)
        if s >= args.warmup:
            vals.append((v, dt))
    value = float(np.mean([v for v, _ in vals]))
    how = ("unmodified reference (baseline/_ref), CurieOptimise(max_processors=%d).optimise_routine" % cores
           if kind == "reference" else "oracle port, fork pool")
    sample = "%d of %d windows per step (evenly spaced), %s, OPENBLAS_NUM_THREADS=1" % (nsample, xc.size, how)
    emit(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean([d for _, d in vals])),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, xc.size, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(args, nwin, world):
    return {"workload": "BASELINE config 3: Bouligand MAP fit, synthetic %dx%d grid @1km, 200 km windows "
                        "(201x201 nodes, 99 bins), 10 km stride" % (args.grid, args.grid),
            "windows_total": int(nwin), "windows_per_gpu": int(-(-nwin // world)), "window_nodes": 201, "nbins": 99,
            "taper": "hanning", "x0": list(X0), "priors": "none",
            "parallelism": "windows sharded in contiguous blocks over %d rank(s), each rank holds the grid rows of "
                           "its block, one NCCL gather" % world,
            "l2_note": "per-step working set (grid rows + row spectra 2.6 GB/N + spectra 240 MB/N) exceeds the 126 MB L2"}


# --------------------------------------------------------------------------- clocks
class ClockSampler(object):
    """nvidia-smi polled beside the run (the recipe's clocks line, every 20 ms, time-stamped).  It is started BEFORE the
    warm-up -- nvidia-smi needs ~0.2 s to produce its first line, more than a whole 8-GPU timed region -- and `mark()`
    / `stop()` bracket the timed region: only samples stamped inside it are summarised.  A region shorter than two
    polls falls back to every sample since the start (the warm-up runs the same kernels) and says so in `window`."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        import datetime
        self.now = datetime.datetime.now
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        self.t0 = self.now()
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "20"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def mark(self):
        """The timed region starts now."""
        self.t0 = self.now()

    def stop(self):
        import datetime
        t1 = self.now()
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        rows = []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f:
            c = [v.strip() for v in line.split(",")]
            if len(c) < 10:
                continue
            try:
                ts = datetime.datetime.strptime(c[0], "%Y/%m/%d %H:%M:%S.%f")
                rows.append((ts, float(c[2]), float(c[3]), float(c[4]),
                             [nm for nm, v in zip(names, c[6:10]) if v.lower().startswith("active")]))
            except ValueError:
                continue
        self.f.close()
        os.unlink(self.f.name)
        inside = [r for r in rows if self.t0 <= r[0] <= t1]
        window = "timed region"
        if len(inside) < 2 and len(rows) > len(inside):
            inside, window = rows, "warm-up + timed region (the timed region is shorter than two 20 ms polls)"
        if inside:
            power = [r[3] for r in inside]
            hot = [r[1] for r in inside if r[3] >= 0.5 * max(power)] or [r[1] for r in inside]
            out.update(sm_mhz=float(np.median(hot)), sm_max_mhz=float(max(r[2] for r in inside)),
                       reasons=sorted({nm for r in inside for nm in r[4]}), samples=len(inside),
                       power_w_max=float(max(power)), window=window)
        return out


# --------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    data, ext = make_inputs(args.grid)
    xc, yc = centroid_list(data, ext)
    nwin = xc.size

    # CPU baselines first (N=1, rank 0 only), before CUDA is initialised in this process: the workers are forked
    cpu, cpu_mcmc = None, None
    if world == 1 and not args.no_cpu_baseline:
        cores = host_cores()
        nsample = args.cpu_sample or cores * 500       # ~25 ms per window per core -> 10-20 s of wall clock
        v, dt, kind = cpu_arm(data, ext, xc, yc, nsample, cores)
        how = ("unmodified reference (baseline/_ref), CurieOptimise(max_processors=%d).optimise_routine" % cores
               if kind == "reference" else "oracle port, fork pool")
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": "%d of %d windows (evenly spaced) in %.1f s, %s, OPENBLAS_NUM_THREADS=1" % (nsample, nwin, dt, how)}
        if kind == "reference":
            pv, pdt = port_windows_per_s(data, ext, xc, yc, cores * 150, cores)
            cpu["port"] = {"value": pv, "sample": "%d windows in %.1f s, oracle port, fork pool" % (cores * 150, pdt)}
            with tempfile.NamedTemporaryFile(suffix=".npy", delete=False) as tf:
                np.save(tf, data)
            try:
                cpu["as_shipped_blas"] = asshipped_windows_per_s(tf.name, cores * 4, cores)     # ~7 windows/s: oversubscribed
            finally:
                os.unlink(tf.name)
            mv, mdt = ref_mcmc_steps_per_s(data, ext, xc, yc, args.cpu_mcmc_steps, args.cpu_mcmc_steps // 100, cores)
            cpu_mcmc = {"value": mv, "unit": "MH steps/s", "cores": cores, "kind": "reference",
                        "sample": "one chain per core, nsim %d + burn-in %d, in %.1f s; reference metropolis_hastings "
                                  "through its own parallelise_routine, config-4 priors" %
                                  (args.cpu_mcmc_steps, args.cpu_mcmc_steps // 100, mdt)}

    import torch
    import torch.distributed as dist
    import pycurious_b200 as pc
    from pycurious_b200 import _lib, _plans, distributed as pdist
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(v):
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    grid = pc.CurieOptimise(data, *ext)
    grid.pin_host()

    # ---- device-timed headline: this rank's block, inputs (its grid rows, origins) resident in HBM ----------
    lo, hi = pdist.shard_range(nwin, rank, world)
    g_dev, iy0, ix0, n = grid._windows_device(WINDOW, xc[lo:hi], yc[lo:hi])
    grid._settle()
    plan = _plans.spectrum_plan(n, grid.dx, _plans.taper_vector(np.hanning, n))
    info = plan.info()

    def step():
        phi, sig = plan.run(g_dev, iy0, ix0, False, 2.0)
        x = grid._fit_device(plan.k, phi, X0)
        if world > 1:
            return pdist.gather_blocks(x, nwin, dst=0)
        return x

    sampler = ClockSampler(local) if rank == 0 else None
    for _ in range(args.warmup):
        step()
    barrier()
    if sampler:
        sampler.mark()
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(args.steps):
        out = step()
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if sampler else None
    shared = getattr(plan, "last_path", "direct") == "shared_rows"

    # ---- end to end through the public API: host grid -> device rows, fits -> host, every step -------------------
    e2e_times = []
    sampler = ClockSampler(local) if rank == 0 else None
    e2e_warm = max(1, min(args.warmup, 3))                 # untimed calls first, like the device-timed loop
    for s in range(e2e_warm + max(3, min(args.steps, 10))):
        barrier()
        t0 = time.perf_counter()
        res = grid.optimise_routine(WINDOW, xc, yc)        # shards over the ranks by itself inside a process group
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        if s >= e2e_warm:
            e2e_times.append(dt)
        elif s == e2e_warm - 1 and sampler:
            sampler.mark()
    rows_up = int(g_dev.shape[0]) * int(g_dev.shape[1]) * 8 + 2 * 4 * (hi - lo) + 8 * plan.k.size
    h2d_t = torch.tensor([float(rows_up)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(h2d_t)
    e2e_clocks = sampler.stop() if sampler else None
    e2e = {"value": nwin / float(np.mean(e2e_times)), "unit": UNIT, "h2d_bytes_per_step": int(h2d_t.item()),
           "d2h_bytes_per_step": int(world * nwin * 4 * 8), "ms_per_step": 1e3 * float(np.mean(e2e_times)),
           "ms_steps": [round(1e3 * t, 2) for t in e2e_times], "clocks": e2e_clocks, "steps": len(e2e_times),
           "api": "CurieOptimise.optimise_routine(200e3, xc, yc) on every rank; grid rows uploaded from the caller's "
                  "(page-locked) array on every call, results returned to every rank"}
    # ---- check: the gathered multi-GPU result equals what ONE GPU computes for the same windows ---------------
    check = None
    if world > 1 and rank == 0:
        # two whole lattice rows on either side of every block boundary, computed by rank 0 alone through the same
        # kernels (shared row spectra forced: 780 windows would otherwise take the scattered-window path, whose
        # spectra differ in the last bits)
        cols = int(np.flatnonzero(xc[1:] == xc[0])[0]) + 1
        cuts = sorted({pdist.shard_range(nwin, r, world)[0] for r in range(1, world)})
        sel = np.unique(np.concatenate([np.arange(max(0, (c // cols - 2) * cols), min(nwin, (c // cols + 2) * cols))
                                        for c in cuts]))
        os.environ["CB_SPEC_SHARE_ROWS"] = "1"
        try:
            solo = grid.optimise_batch(WINDOW, xc[sel], yc[sel], shard=False)
        finally:
            del os.environ["CB_SPEC_SHARE_ROWS"]
        got = out.index_select(0, torch.from_numpy(sel).cuda()).cpu().numpy()
        check = {"windows": int(sel.size), "equal_to_single_gpu": bool(np.array_equal(solo, got)),
                 "max_rel_dev": float(np.max(np.abs(solo - got) / np.maximum(np.abs(solo), 1e-300)))}

    if rank == 0 and world > 1 and check is not None:
        got = np.stack(res, axis=1)
        check["public_api_equal_to_gathered"] = bool(np.array_equal(got, out.cpu().numpy()))


    # ---- per-kernel CUDA-event times (same stream the kernels are launched on), this rank's block ---------------
    def timed(fn, reps):
        fn()
        torch.cuda.synchronize()
        marks = [ev() for _ in range(reps + 1)]
        marks[0].record()
        for i in range(reps):
            out_ = fn()
            marks[i + 1].record()
        torch.cuda.synchronize()
        # median of the back-to-back launches: one slow repetition does not move the per-stage figure
        return float(np.median([marks[i].elapsed_time(marks[i + 1]) for i in range(reps)])), out_

    reps = max(3, min(args.steps, 10))
    t_spec, (phi, sig) = timed(lambda: plan.run(g_dev, iy0, ix0, False, 2.0), reps)
    t_fit, (x_blk, diag) = timed(lambda: grid._fit_device(plan.k, phi, X0, diagnostics=True), reps)
    nfev_sum = float(diag["nfev"].double().sum().item())
    stages = {"windows": int(hi - lo), "spectrum_ms": t_spec, "fit_ms": t_fit,
              "fit_nfev_mean": nfev_sum / max(hi - lo, 1),
              "fit_niter_mean": float(diag["niter"].double().mean().item()),
              "fit_status_counts": [int((diag["status"] == s).sum().item()) for s in range(5)]}

    # ---- weak scaling beside it (N > 1): every rank runs the whole list -------------------------------------------
    weak = None
    if world > 1 and not args.no_weak:
        grid.hold_device()
        gw, iyw, ixw, _ = grid._windows_device(WINDOW, xc, yc)
        grid._settle()

        def wstep():
            p_, _s = plan.run(gw, iyw, ixw, False, 2.0)
            return pdist.gather_blocks(grid._fit_device(plan.k, p_, X0), nwin * world, dst=0)
        wstep()
        barrier()
        a, b = ev(), ev()
        a.record()
        for _ in range(3):
            wstep()
        b.record()
        barrier()
        wms = max_over_ranks(a.elapsed_time(b)) / 3
        weak = {"value": nwin * world / (wms * 1e-3), "unit": UNIT, "ms_per_step": wms, "windows_total": int(nwin * world)}
        del gw, iyw, ixw
        grid.hold_device(False)

    # ---- MCMC samples/s: BASELINE config 4 at its real chain length (nsim 1e5, burn-in 1e3) on a bounded ----------
    # sample of this rank's windows, priors and proposal scales of SURVEY 8d, chains started at the MAP
    # estimate, moments streamed on the device
    mcmc = None
    if args.mcmc_nsim > 0:
        nch = min(hi - lo, args.mcmc_chains)
        sel_t = torch.from_numpy(np.linspace(0, hi - lo - 1, nch).astype(np.int64)).cuda()
        phi_s = phi.index_select(0, sel_t).contiguous()
        x_map = x_blk.index_select(0, sel_t).contiguous()
        cmean = x_blk[:, 3].sum()
        if world > 1:
            dist.all_reduce(cmean)
        grid.add_prior(beta=(3.0, 1.0), zt=(1.0, 1.0), dz=(20.0, 10.0), C=(float(cmean.item()) / nwin, 5.0))
        burn = max(1, args.mcmc_nsim // 100)
        grid._mh_device(plan.k, phi_s, 500, 5, MCMC_SCALE, X0, x_map, SEED, 1, chain_offset=lo)     # warm
        barrier()
        a, b = ev(), ev()
        a.record()
        mom, acc = grid._mh_device(plan.k, phi_s, args.mcmc_nsim, burn, MCMC_SCALE, X0, x_map, SEED, 1, chain_offset=lo)
        b.record()
        barrier()
        mms = max_over_ranks(a.elapsed_time(b))
        grid.reset_priors()
        tot = torch.tensor([float(nch), float(acc.double().sum().item()),
                            float(torch.isfinite(mom).all().item())], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tot)
        steps = tot[0].item() * (args.mcmc_nsim + burn)
        mcmc = {"metric": "mcmc_samples_per_s", "value": steps / (mms * 1e-3), "unit": "MH steps/s (burn-in included)",
                "chains_total": int(tot[0].item()), "chains_per_gpu": int(nch), "nsim": int(args.mcmc_nsim),
                "burnin": int(burn), "ms": mms, "acceptance_mean": tot[1].item() / (tot[0].item() * args.mcmc_nsim),
                "moments_finite_on_all_ranks": bool(tot[2].item() == world),
                "config": "config 4: nsim and burn-in as named, on an evenly spaced sample of each rank's block; priors "
                          "beta=(3,1) zt=(1,1) dz=(20,10) C=(mean MAP C,5), x_scale=%s, start = per-window MAP" % MCMC_SCALE,
                "cpu_baseline": cpu_mcmc}
        if cpu_mcmc:
            mcmc["vs_cpu"] = mcmc["value"] / cpu_mcmc["value"]

    # ---- the other named configurations (driver-visible records) ------------------------------------------------
    config2 = config5 = None
    if not args.no_extra_configs:
        if world == 1:
            config2 = run_config2(torch, pc, grid, data, ext)
        if args.config5 != "off":
            config5 = run_config5(torch, dist, pc, args, rank, world, barrier, max_over_ranks)

    if rank == 0:
        ms_step = ms_total / args.steps
        value = nwin / (ms_step * 1e-3)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        fp64_peak = _lib.fp64_fma_peak()
        nb = plan.nbins
        nblk = hi - lo
        spec_name = "colspec_kernel (+ rowspec_kernel)" if shared else "spectrum_kernel"
        dominant = spec_name if stages["spectrum_ms"] >= stages["fit_ms"] else "fit_kernel"
        counts = {}
        try:        # FP64 instruction counts (ncu, frozen) and per-launch DRAM bytes from the committed captures
            counts = json.load(open(os.path.join(ROOT, "profiles", "fp64_counts.json")))
        except Exception:
            pass
        if dominant == spec_name:
            bytes_per_win = n * n * 8 + 2 * nb * 8            # grid read + Phi, sigma write
            dur = stages["spectrum_ms"] * 1e-3
            flop = counts.get("spectrum_flop_per_window_201") and counts["spectrum_flop_per_window_201"] * nblk
        else:
            bytes_per_win = nb * 8 + 4 * 8                    # Phi read + 4 parameters written
            dur = stages["fit_ms"] * 1e-3
            flop = counts.get("fit_flop_per_point") and counts["fit_flop_per_point"] * nfev_sum * nb
        achieved_gbs = bytes_per_win * nblk / dur / 1e9
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            key = "spectrum_shared_rows" if (shared and dominant == spec_name) else ("spectrum_direct" if dominant == spec_name else "fit")
            if tr.get(key, {}).get("windows") == nblk:
                traffic = tr[key]["dram_bytes"]
        except Exception:
            pass
        fit_flop = counts.get("fit_flop_per_point") and counts["fit_flop_per_point"] * nfev_sum * nb
        spec_flop = counts.get("spectrum_flop_per_window_201") and counts["spectrum_flop_per_window_201"] * nblk
        roofline = {"kernel": dominant, "bound": "fp64",
                    "achieved": (flop / dur / 1e12) if flop else None, "peak": fp64_peak, "unit": "TFLOP/s",
                    "frac": (flop / dur / 1e12 / fp64_peak) if flop else None, "traffic": traffic,
                    "peak_source": "cb_fp64_fma_peak: register-resident DFMA chains measured on this device in this run "
                                   "(MEASURED_PEAKS.json has no FP64 entry)",
                    "flops": "executed FP64 instructions counted by ncu (2 x DFMA + DADD + DMUL, thread level), "
                             "profiles/fp64_counts.json" if flop else "profiles/fp64_counts.json missing",
                    "hbm": {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                            "frac": achieved_gbs / hbm_peak, "peak_source": peak_src,
                            "algorithmic_bytes_per_window": bytes_per_win,
                            "note": "SURVEY 8(d) contract figure; windows overlap ~400x, so neither kernel is HBM bound"},
                    "per_kernel_tflops": {"spectrum": spec_flop and spec_flop / (stages["spectrum_ms"] * 1e-3) / 1e12,
                                          "fit": fit_flop and fit_flop / (stages["fit_ms"] * 1e-3) / 1e12},
                    "fit_flop_per_point": counts.get("fit_flop_per_point"),
                    "spectrum_flop_per_window": counts.get("spectrum_flop_per_window_201")}
        if mcmc and counts.get("mh_flop_per_step_99"):
            mcmc["fp64_tflops_counted"] = counts["mh_flop_per_step_99"] * mcmc["value"] / 1e12
        # kernels of this repo per step: spectrum (row + column pass, or the one direct kernel), the k tables, and
        # the fit's template / begin / first / (rounds - 1) evaluation / rounds step / finish kernels
        import ctypes
        fit_rounds = ctypes.c_int(0)
        _lib.check(_lib.load().cb_fit_schedule(int(hi - lo), ctypes.byref(fit_rounds), None))   # this rank's share
        fit_rounds = fit_rounds.value
        launches_per_step = (2 if shared else 1) + 1 + 3 + (fit_rounds - 1) + fit_rounds + 1
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args, nwin, world),
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches_per_step * args.steps,
                "roofline": roofline, "stages": stages, "weak": weak, "check": check,
                "spectrum_path": "shared_rows" if shared else "direct",
                "plan": {"fft_len": info.fft_len, "bluestein": info.bluestein, "radix": list(info.radix[:info.npass]),
                         "teams": info.teams, "ctas": info.grid_max, "smem_bytes": info.smem_bytes},
                "mcmc": mcmc, "config2": config2, "config5": config5, "cpu_baseline": cpu}
        emit(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_config2(torch, pc, grid, data, ext):
    """BASELINE config 2: Tanaka radial_spectrum(power=0.5) sweep, 128 km windows, 8 km stride (246,016 windows of
    129 x 129), spectrum only and with the batched Tanaka line fits; public batch API, results left on the device."""
    xc, yc = centroid_list(data, ext, 128e3, 8e3)
    out = {"windows": int(xc.size), "window_nodes": 129}
    for name, fn in (("spectrum", lambda: grid.radial_spectrum_batch(128e3, xc, yc, power=0.5, as_numpy=False)),
                     ("spectrum_tanaka", lambda: grid.tanaka_batch(128e3, xc, yc, as_numpy=False))):
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        out[name + "_windows_per_s"] = xc.size / min(ts)
        out[name + "_ms"] = 1e3 * min(ts)
    out["note"] = "end to end per call: grid rows uploaded (134 MB), centroids -> origins on the host, results stay on the device"
    return out


def run_config5(torch, dist, pc, args, rank, world, barrier, max_over_ranks):
    """BASELINE config 5: EMAG2-shaped 21600 x 10800 synthetic grid (1.87 GB), 100/200/300 km windows at 20 km
    stride through curie_depth_maps (spectrum + fit per window; the lattice rows are split over the ranks, every
    rank receives the maps).  `--config5 lattice` uses a 4096^2 grid instead (bounded)."""
    from pycurious_b200 import synth
    full = args.config5 == "full"
    ny, nx = (10800, 21600) if full else (args.grid, args.grid)
    t0 = time.perf_counter()
    data = synth.fractal_grid(ny, nx, seed=SEED + 5, workers=max(1, host_cores() // world), table=1 << 20)
    t_synth = time.perf_counter() - t0
    G = pc.CurieOptimise(data, *synth.grid_extent(ny, nx))
    G.pin_host()
    out = {"grid": [ny, nx], "stride_km": 20, "synth_s": t_synth, "legs": {}}
    total_w, total_s = 0, 0.0
    for wkm in (100, 200, 300):
        G.curie_depth_maps(wkm * 1e3, 20e3, 20e3)                   # warm: plans, row-spectrum scratch
        ts = []
        for _ in range(2):
            barrier()
            t0 = time.perf_counter()
            maps = G.curie_depth_maps(wkm * 1e3, 20e3, 20e3)
            torch.cuda.synchronize()
            ts.append(max_over_ranks(time.perf_counter() - t0))
        nw = maps["beta"].size
        out["legs"]["%dkm" % wkm] = {"windows": int(nw), "ms": 1e3 * min(ts), "windows_per_s": nw / min(ts),
                                     "median_dz": float(np.median(maps["dz"])), "median_beta": float(np.median(maps["beta"])),
                                     "all_finite": bool(np.all(np.isfinite(maps["curie_depth"])))}
        total_w += nw
        total_s += min(ts)
    out["windows_total"] = int(total_w)
    out["windows_per_s"] = total_w / total_s
    out["seconds"] = total_s
    out["note"] = ("end to end per call through the public API: each rank uploads the grid rows of its lattice rows from "
                   "page-locked host memory, maps are all-gathered and returned as numpy arrays on every rank")
    return out


def _claim_stdout():
    """stdout carries exactly ONE JSON line.  Libraries (NCCL's INFO log, for one) write to file descriptor 1
    behind Python's back, so fd 1 is pointed at stderr for the whole run -- NCCL_DEBUG stays as the caller set
    it -- and the JSON line goes to a private duplicate of the original stdout."""
    sys.stdout.flush()
    keep = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(keep, "w")


_OUT = None


def emit(line):
    _OUT.write(line + "\n")
    _OUT.flush()


def main():
    global _OUT
    _OUT = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--grid", type=int, default=GRID_N, help="synthetic grid edge (default: BASELINE's 4096)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="windows in the CPU-baseline sample (0 = auto)")
    ap.add_argument("--cpu-mcmc-steps", type=int, default=40000, help="nsim of the reference's MH chains (one per core)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mcmc-nsim", type=int, default=100000, help="MH samples per chain in the MCMC leg (0 = skip)")
    ap.add_argument("--mcmc-chains", type=int, default=148 * 24 * 4,
                    help="chains (= windows) per GPU in the MCMC leg: four full waves of 24 single-warp CTAs on 148 SMs")
    ap.add_argument("--no-extra-configs", action="store_true", help="skip the config 2 / config 5 legs")
    ap.add_argument("--config5", default="full", choices=["full", "lattice", "off"])
    ap.add_argument("--no-weak", action="store_true", help="skip the weak-scaling leg at N > 1")
    ap.add_argument("--cpu-arm-child", default=None, help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.cpu_arm_child:
        run_cpu_arm_child(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
