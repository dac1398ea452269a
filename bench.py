#!/usr/bin/env python
"""bench.py -- Curie-depth windows/s (spectrum + Bouligand MAP fit), BASELINE.json config 3:
synthetic 4096 x 4096 grid at 1 km, 200 km windows (201 x 201 nodes, 99 k-bins), 10 km stride
= 390^2 = 152,100 windows per GPU.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host cores

One step = one pass of the hot path (gather/taper/2-D DFT/radial bins, then the bounded
quasi-Newton fit) over the whole centroid list.  N > 1 (torchrun, one rank per GPU): weak
scaling -- the global centroid list is N copies of the config-3 list, rank r owns the contiguous
block r, the (n,4) results are gathered to rank 0 with one NCCL gather inside the timed step.

Printed keys follow the driver contract; extra keys: `roofline` (dominant kernel, algorithmic
bytes vs measured HBM peak, plus the FP64 view that actually bounds these kernels), `stages`
(per-kernel CUDA-event times), `cpu_baseline` (the oracle port on this box's host cores, bounded
sample), `e2e` (public API with host buffers: H2D of the grid and centroids, D2H of the results).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")      # fair CPU baseline: 1 BLAS thread per worker process
os.environ.setdefault("OMP_NUM_THREADS", "1")
os.environ.setdefault("MKL_NUM_THREADS", "1")
# stdout carries exactly one JSON line: keep NCCL's banner (NCCL_DEBUG=VERSION/INFO on some boxes) off it
if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "INFO") and not os.environ.get("CB_BENCH_KEEP_NCCL_DEBUG"):
    os.environ["NCCL_DEBUG"] = "WARN"

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GRID_N = 4096
WINDOW = 200e3
STRIDE = 10e3
SEED = 20261019
METRIC = "curie_depth_windows_per_s"
UNIT = "windows/s"


def make_inputs(grid_n):
    from pycurious_b200 import synth
    data = synth.fractal_grid(grid_n, grid_n, seed=SEED)
    return data, synth.grid_extent(grid_n, grid_n)


# --------------------------------------------------------------------------- CPU arm (oracle port)
_W = {}


def _cpu_init(data, ext):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import curie_oracle as orc
    _W["orc"] = orc
    _W["spec"] = orc.GridSpec(data, *ext)


def _cpu_fit(c):
    orc = _W["orc"]
    return orc.window_fit(_W["spec"], WINDOW, c[0], c[1])


def cpu_windows_per_s(data, ext, xc, yc, nsample, cores):
    """Oracle (numpy/scipy restatement of the reference) on `cores` forked workers over an evenly
    spaced sample of the same centroid list; returns windows/s."""
    import multiprocessing as mp
    idx = np.linspace(0, xc.size - 1, nsample).astype(np.int64)
    work = [(float(xc[i]), float(yc[i])) for i in idx]
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_cpu_init, initargs=(data, ext)) as pool:
        pool.map(_cpu_fit, work[:cores])              # warm the workers (imports, first-call costs)
        t0 = time.perf_counter()
        pool.map(_cpu_fit, work, chunksize=max(1, nsample // (cores * 8)))
        dt = time.perf_counter() - t0
    return nsample / dt, dt


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from pycurious_b200.grid import CurieGrid
    data, ext = make_inputs(args.grid)
    xc, yc = CurieGrid.create_centroid_list(_HostGrid(data, ext), WINDOW, STRIDE, STRIDE)
    cores = host_cores()
    nsample = args.cpu_sample or cores * 500           # ~25 ms per window per core -> 10-20 s per step
    vals = []
    for s in range(args.warmup + args.steps):
        v, dt = cpu_windows_per_s(data, ext, xc, yc, nsample, cores)
        if s >= args.warmup:
            vals.append((v, dt))
    value = float(np.mean([v for v, _ in vals]))
    sample = "%d of %d windows per step (evenly spaced), oracle port, fork pool, OPENBLAS_NUM_THREADS=1" % (nsample, xc.size)
    emit(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean([d for _, d in vals])),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, xc.size, 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


class _HostGrid(object):
    """Just the coordinate attributes create_centroid_list needs (no CUDA)."""

    def __init__(self, data, ext):
        self.xcoords, self.dx = np.linspace(ext[0], ext[1], data.shape[1], retstep=True)
        self.ycoords = np.linspace(ext[2], ext[3], data.shape[0])


def workload_config(args, nwin, world):
    return {"workload": "BASELINE config 3: Bouligand MAP fit, synthetic %dx%d grid @1km, 200 km windows "
                        "(201x201 nodes, 99 bins), 10 km stride" % (args.grid, args.grid),
            "windows_per_gpu": int(nwin), "windows_total": int(nwin * world), "window_nodes": 201, "nbins": 99,
            "taper": "hanning", "x0": [3.0, 1.0, 10.0, 5.0], "priors": "none",
            "parallelism": "windows sharded in contiguous blocks, %d rank(s), one NCCL gather" % world,
            "l2_note": "per-step working set (grid 134 MB + per-CTA scratch 72 MB) exceeds the 126 MB L2"}


# --------------------------------------------------------------------------- clocks
class ClockSampler(object):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
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
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f:
            c = [v.strip() for v in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); smax.append(float(c[2])); power.append(float(c[3]))
            except ValueError:
                continue
            for nm, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        self.f.close()
        os.unlink(self.f.name)
        if sm:
            hot = [s for s, p in zip(sm, power) if p >= 0.5 * max(power)] or sm
            out.update(sm_mhz=float(np.median(hot)), sm_max_mhz=float(max(smax)), reasons=sorted(reasons),
                       samples=len(sm), power_w_max=float(max(power)))
        return out


# --------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    data, ext = make_inputs(args.grid)
    hg = _HostGrid(data, ext)
    from pycurious_b200.grid import CurieGrid
    xc, yc = CurieGrid.create_centroid_list(hg, WINDOW, STRIDE, STRIDE)
    nwin = xc.size

    # CPU baseline first (N=1, rank 0 only), before CUDA is initialised in this process: the workers are forked
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = host_cores()
        nsample = args.cpu_sample or cores * 500       # ~25 ms per window per core -> 10-20 s of wall clock
        v, dt = cpu_windows_per_s(data, ext, xc, yc, nsample, cores)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "%d of %d windows (evenly spaced) in %.1f s, oracle port, fork pool, OPENBLAS_NUM_THREADS=1"
                         % (nsample, nwin, dt)}

    import torch
    import torch.distributed as dist
    import pycurious_b200 as pc
    from pycurious_b200 import _lib, _plans, distributed as pdist
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    grid = pc.CurieOptimise(data, *ext)
    grid.pin_host()
    g_dev = grid.device_grid()
    iy0, ix0, n = grid._origins_device(WINDOW, xc, yc)
    plan = _plans.spectrum_plan(n, grid.dx, _plans.taper_vector(np.hanning, n))
    info = plan.info()
    x0 = (3.0, 1.0, 10.0, 5.0)
    gather_buf = None

    def step():
        phi, sig = plan.run(g_dev, iy0, ix0, False, 2.0)
        x = grid._fit_device(plan.k, phi, x0)
        if world > 1:
            return pdist.gather_blocks(x, nwin * world, dst=0)
        return x

    ev = lambda: torch.cuda.Event(enable_timing=True)
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(args.steps):
        out = step()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    clocks = sampler.stop() if sampler else None

    # per-kernel CUDA-event times (same stream the kernels are launched on)
    stages = {}
    if rank == 0:
        # each stage timed on its own over back-to-back launches (events on the launching stream)
        def timed(fn, reps):
            fn()
            torch.cuda.synchronize()
            marks = [ev() for _ in range(reps + 1)]
            marks[0].record()
            for i in range(reps):
                out_ = fn()
                marks[i + 1].record()
            torch.cuda.synchronize()
            # median of the back-to-back launches: one slow repetition (allocator growth, a host hiccup between
            # launches) does not move the per-stage figure
            return float(np.median([marks[i].elapsed_time(marks[i + 1]) for i in range(reps)])), out_
        reps = max(3, args.steps)
        t_spec, (phi, sig) = timed(lambda: plan.run(g_dev, iy0, ix0, False, 2.0), reps)
        t_fit, (x, diag) = timed(lambda: grid._fit_device(plan.k, phi, x0, diagnostics=True), reps)
        ts, tf = [t_spec], [t_fit]
        nfev = diag["nfev"].double()
        stages = {"spectrum_ms": float(np.mean(ts)), "fit_ms": float(np.mean(tf)),
                  "fit_nfev_mean": float(nfev.mean().item()), "fit_niter_mean": float(diag["niter"].double().mean().item()),
                  "fit_status_counts": [int((diag["status"] == s).sum().item()) for s in range(5)]}

    # MCMC samples/s (BASELINE config 4, bounded): one Metropolis-Hastings chain per window on a
    # sample of the centroid list, priors and proposal scale of SURVEY 8d, moments streamed on the device
    mcmc = None
    if rank == 0 and args.mcmc_steps > 0:
        nch = min(nwin, args.mcmc_chains)
        sel = np.linspace(0, nwin - 1, nch).astype(np.int64)
        sel_t = torch.from_numpy(sel).cuda()
        phi_all, _ = plan.run(g_dev, iy0, ix0, False, 2.0)
        phi_s = phi_all.index_select(0, sel_t).contiguous()
        x_map = grid._fit_device(plan.k, phi_s, x0)
        grid.add_prior(beta=(3.0, 1.0), zt=(1.0, 1.0), dz=(20.0, 10.0), C=(float(x_map[:, 3].mean().item()), 5.0))
        burn = max(1, args.mcmc_steps // 100)
        tm = []
        for _ in range(3):
            a, b = ev(), ev()
            a.record()
            mom, acc = grid._mh_device(plan.k, phi_s, args.mcmc_steps, burn, [0.25, 0.1, 1.0, 0.5], x0, x_map,
                                       SEED, 1)
            b.record()
            torch.cuda.synchronize()
            tm.append(a.elapsed_time(b))
        grid.reset_priors()
        steps = nch * (args.mcmc_steps + burn)
        mcmc = {"metric": "mcmc_samples_per_s", "value": steps / (min(tm[1:]) * 1e-3), "unit": "MH steps/s (burn-in included)",
                "chains": int(nch), "nsim": int(args.mcmc_steps), "burnin": int(burn), "ms": float(min(tm[1:])),
                "acceptance_mean": float(acc.double().mean().item() / args.mcmc_steps),
                "nominal_fp64_tflops": steps * nb_points(plan) * 400.0 / (min(tm[1:]) * 1e-3) / 1e12,
                "config": "config 4 sample: priors beta=(3,1) zt=(1,1) dz=(20,10) C=(mean MAP C,5), x_scale=[0.25,0.1,1,0.5], start = per-window MAP"}

    # end to end through the public API: host grid (pinned) -> device, fits -> host, every step
    e2e = None
    if rank == 0 or world > 1:
        lo, hi = 0, nwin
        times = []
        for s in range(1 + max(2, min(args.steps, 5))):
            grid.invalidate_device()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            if world > 1:
                res = pdist.optimise_routine_sharded(grid, WINDOW, np.tile(xc, world), np.tile(yc, world), dst=0)
            else:
                res = grid.optimise_routine(WINDOW, xc, yc)
            torch.cuda.synchronize()
            dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            if s > 0:
                times.append(float(dt.item()))
        e2e = {"value": nwin * world / float(np.mean(times)), "unit": UNIT,
               "h2d_bytes_per_step": int(world * (data.nbytes + 2 * 4 * nwin + 8 * plan.k.size)),
               "d2h_bytes_per_step": int(world * nwin * 4 * 8), "ms_per_step": 1e3 * float(np.mean(times))}

    if rank == 0:
        ms_step = ms_total / args.steps
        value = nwin * world / (ms_step * 1e-3)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        fp64_peak = _lib.fp64_fma_peak()
        nb = plan.nbins
        shared = getattr(plan, "last_path", "direct") == "shared_rows"
        spec_name = "colspec_kernel (+ rowspec_kernel)" if shared else "spectrum_kernel"
        dominant = spec_name if stages["spectrum_ms"] >= stages["fit_ms"] else "fit_kernel"
        if dominant == spec_name:
            bytes_per_win = n * n * 8 + 2 * nb * 8            # grid read + Phi, sigma write
            dur = stages["spectrum_ms"] * 1e-3
        else:
            bytes_per_win = nb * 8 + 4 * 8                    # Phi read + 4 parameters written
            dur = stages["fit_ms"] * 1e-3
        achieved = bytes_per_win * nwin / dur / 1e9
        # FP64 view: nominal 2-D FFT flops (5 n^2 log2 n^2) and forward-model points x 400 flop (SURVEY 8d planning figure)
        fft_flops = 5.0 * n * n * np.log2(float(n * n)) * nwin
        fit_flops = stages["fit_nfev_mean"] * nb * 400.0 * nwin
        traffic = None
        try:        # per-launch DRAM bytes of the dominant kernel from the committed full-size ncu capture
            tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            key = "spectrum_shared_rows" if (shared and dominant == spec_name) else ("spectrum_direct" if dominant == spec_name else "fit")
            if tr.get(key, {}).get("windows") == nwin:
                traffic = tr[key]["dram_bytes"]
        except Exception:
            pass
        roofline = {"kernel": dominant, "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                    "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_window": bytes_per_win,
                    "note": "both kernels are FP64/shared-memory bound, not HBM bound (see fp64)",
                    "fp64": {"peak_tflops_measured": fp64_peak,
                             "spectrum_nominal_tflops": fft_flops / (stages["spectrum_ms"] * 1e-3) / 1e12,
                             "fit_nominal_tflops": fit_flops / (stages["fit_ms"] * 1e-3) / 1e12,
                             "fit_flop_per_point_convention": 400}}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args, nwin, world),
                "clocks": clocks, "e2e": e2e, "gpu_launches": (4 if shared else 3) * args.steps,
                "roofline": roofline, "stages": stages,
                "spectrum_path": "shared_rows" if shared else "direct",
                "plan": {"fft_len": info.fft_len, "bluestein": info.bluestein, "radix": list(info.radix[:info.npass]),
                         "teams": info.teams, "ctas": info.grid_max, "smem_bytes": info.smem_bytes},
                "mcmc": mcmc, "cpu_baseline": cpu}
        emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def nb_points(plan):
    return plan.nbins


def _claim_stdout():
    """stdout carries exactly ONE JSON line.  Libraries (NCCL's version banner, for one) write to file
    descriptor 1 behind Python's back, so fd 1 is pointed at stderr for the whole run and the JSON line goes
    to a private duplicate of the original stdout."""
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
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mcmc-steps", type=int, default=2000, help="MH steps per chain in the MCMC leg (0 = skip)")
    ap.add_argument("--mcmc-chains", type=int, default=148 * 24 * 4,
                    help="chains (= windows) in the MCMC leg: four full waves of 24 single-warp CTAs on 148 SMs")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
