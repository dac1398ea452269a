"""Drop-in `CurieOptimise`: MAP fits and Metropolis-Hastings sampling of the Bouligand et al.
(2009) parameters (beta, zt, dz, C) for windows of a magnetic grid, batched on a B200.

Mirrors the reference's pycurious/optimise.py (v1.1.1): same class, constructor, attributes
(`bounds`, `prior`, `prior_pdf`, `max_processors`), method signatures, return containers and
errors.  Where the reference forks `max_processors` workers
and calls scipy per centroid (optimise.py:386-424), this class launches one kernel for the
whole centroid list.  `max_processors` keeps its meaning of "how many workers": the batch fans out
over min(max_processors, visible GPUs) devices, one host thread per device (no process is ever
forked: CUDA contexts do not survive fork), and under an initialised `torch.distributed` process
group (one process per GPU) the same calls shard over the ranks by themselves
(pycurious_b200/distributed.py).

Differences that are deliberate and documented in DESIGN.md:
  * the optimiser is a bounded quasi-Newton method on the device (csrc/fit_body.cuh), not
    scipy's L-BFGS-B; it converges to the same bounded minimum more tightly
  * Metropolis-Hastings draws from a counter-based Philox stream; `seed=None` takes the seed
    from numpy's global RNG, so `np.random.seed(...)` still makes runs reproducible
"""
from __future__ import annotations

import ctypes as C
import inspect
import threading
import warnings
from multiprocessing import cpu_count

import numpy as np

from . import _lib, _plans
from . import distributed as _pdist
from .grid import CurieGrid

_KEYS = ("beta", "zt", "dz", "C")
_RVS_LOCK = threading.Lock()


class CurieOptimise(CurieGrid):
    """Extends CurieGrid with the inversion routines (reference: optimise.py:50-118).

    Attributes:
        bounds : list of (lower, upper) for beta, zt, dz, C; None = unbounded
        prior, prior_pdf : dicts keyed by 'beta', 'zt', 'dz', 'C'
        max_processors : number of workers (optimise.py:116); here GPUs of this process, capped by the
                         visible device count
        shard : True (default) lets batch calls spread over GPUs / ranks; False keeps them on the
                current device
    """

    MIN_WINDOWS_PER_DEVICE = 2048    # below this a second device costs more (threads, plans, uploads) than it saves

    def __init__(self, grid, xmin, xmax, ymin, ymax, **kwargs):
        super(CurieOptimise, self).__init__(grid, xmin, xmax, ymin, ymax)
        self.reset_priors()
        self.bounds = [(0.0, None), (0.0, None), (0.0, None), (None, None)]
        self.max_processors = kwargs.pop("max_processors", cpu_count())
        self.shard = True

    # ------------------------------------------------------------------ priors (optimise.py:120-194)
    def add_prior(self, **kwargs):
        """add_prior(beta=(p, sigma_p)) or add_prior(beta=scipy.stats.norm(p, sigma_p))."""
        for key, prior in kwargs.items():
            if key not in self.prior:
                raise ValueError("prior must be one of {}".format(self.prior.keys()))
            if type(prior) == tuple:
                from scipy import stats
                pdf = stats.norm(*prior)
            elif type(prior).__name__ == "rv_frozen" or hasattr(prior, "rvs") and hasattr(prior, "args"):
                pdf = prior
            else:
                raise ValueError("Use a distribution from scipy.stats module")
            self.prior_pdf[key] = pdf
            self.prior[key] = list(pdf.args)

    def reset_priors(self):
        self.prior = {"beta": None, "zt": None, "dz": None, "C": None}
        self.prior_pdf = {"beta": None, "zt": None, "dz": None, "C": None}

    def objective_function(self, x, x0, sigma_x0, *args):
        """0.5 * sum((x - x0)^2 / sigma^2)  (optimise.py:181-194); O(1) host helper."""
        return 0.5 * np.sum((x - x0) ** 2 / sigma_x0 ** 2)

    def objective_routine(self, **kwargs):
        """Sum of the prior penalties of the given parameters (optimise.py:159-179)."""
        c = 0.0
        for key, val in kwargs.items():
            if key in self.prior and self.prior[key] is not None:
                c += self.objective_function(val, *self.prior[key])
        return c

    def _prior_arrays(self):
        mu = np.full(4, np.nan)
        sg = np.full(4, np.nan)
        for j, key in enumerate(_KEYS):
            p = self.prior[key]
            if p is not None:
                if len(p) < 2:
                    raise TypeError("prior for {} needs (mean, sigma) positional arguments".format(key))
                mu[j], sg[j] = float(p[0]), float(p[1])     # first two positional args (optimise.py:178)
        return mu, sg

    def _bound_arrays(self):
        lb = np.array([-np.inf if b[0] is None else float(b[0]) for b in self.bounds])
        ub = np.array([np.inf if b[1] is None else float(b[1]) for b in self.bounds])
        if lb.size != 4:
            raise ValueError("bounds must hold four (lower, upper) pairs")
        return lb, ub

    # ------------------------------------------------------------------ objective (optimise.py:196-230)
    def min_func(self, x, kh, Phi, sigma_Phi):
        """Misfit 0.5*sum((bouligand2009(kh,*x) - Phi)^2) + priors; non-finite -> 1e99.
        `sigma_Phi` is accepted and ignored, as in the reference (optimise.py:225)."""
        torch = _lib.require_cuda()
        mu, sg = self._prior_arrays()
        k_t = torch.from_numpy(np.ascontiguousarray(kh, dtype=np.float64)).cuda()
        p_t = torch.from_numpy(np.ascontiguousarray(Phi, dtype=np.float64)).cuda()
        x_t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64).reshape(1, 4)).cuda()
        out = torch.empty(1, dtype=torch.float64, device="cuda")
        _lib.check(_lib.load().cb_misfit(k_t.data_ptr(), k_t.numel(), p_t.data_ptr(), 0, x_t.data_ptr(), 1,
                                         mu.ctypes.data, sg.ctypes.data, out.data_ptr(), _lib.stream_ptr(torch)))
        return float(out.item())

    # ------------------------------------------------------------------ fan-out over devices / ranks
    def _shadow(self):
        """Per-thread stand-in for `self` on another device: shares the data, priors and bounds, owns its
        upload bookkeeping."""
        import copy
        sh = copy.copy(self)
        sh._pending = []
        sh._origin = getattr(self, "_origin", self)
        return sh

    def _to_host(self, t):
        """Device tensor -> numpy array through page-locked memory (asynchronous copy + one stream sync)."""
        import torch
        if not t.is_cuda:
            return t.contiguous().numpy()
        host = torch.empty(tuple(t.shape), dtype=t.dtype, pin_memory=True)
        host.copy_(t, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return host.numpy()

    def _run_blocks(self, n, block, shard=None, as_numpy=True, min_per_worker=None, transpose=False, decide=None):
        """Evaluate `block(obj, lo, hi)` -> device tensor (hi-lo, ...) over range(n) in contiguous blocks and
        return the concatenation in reference order ((C, n) instead of (n, C) with `transpose`).

          * inside torch.distributed (world > 1): this rank's block, then one all-gather: every rank returns
            the full result (the reference's fan-out, optimise.py:386-444, across processes)
          * else over min(max_processors, visible GPUs) devices of this process, one host thread each, every
            device writing its block of one page-locked result array
          * else on the current device.

        `decide()` -> the spectrum path of the WHOLE problem (_plans.share_rows_for_list); when the work is split it
        is taken once and imposed on every block, so that the result does not depend on how many workers there are."""
        import torch
        if shard is None:
            shard = self.shard
        tr = (lambda t: t.t()) if transpose else (lambda t: t)
        if shard and _pdist.is_distributed():
            import torch.distributed as dist
            lo, hi = _pdist.shard_range(n, dist.get_rank(), dist.get_world_size())
            sh = self._shadow()
            sh._share_rows = decide() if decide is not None else None
            full = tr(_pdist.gather_blocks(block(sh, lo, hi), n, dst=None))
            sh._settle()
            out = self._to_host(full) if as_numpy else full
            self._settle()
            return out
        floor = self.MIN_WINDOWS_PER_DEVICE if min_per_worker is None else min_per_worker
        ndev = _pdist.device_workers(self.max_processors, n, floor) if shard else 1
        if ndev > 1:
            box = {}
            lock = threading.Lock()
            path = decide() if decide is not None else None

            def work(lo, hi, dev):
                sh = self._shadow()
                sh._share_rows = path
                t = block(sh, lo, hi)
                with lock:                     # the first block to finish fixes the result's trailing shape
                    if "host" not in box:
                        shape = (n,) + tuple(t.shape[1:])
                        box["host"] = torch.empty(shape[::-1] if transpose else shape, dtype=t.dtype, pin_memory=True)
                dst = box["host"][..., lo:hi] if transpose else box["host"][lo:hi]
                dst.copy_(tr(t))
                sh._settle()

            _pdist.fan_out_devices(ndev, n, work)
            return box["host"].numpy() if as_numpy else box["host"].cuda()
        prev = getattr(self, "_share_rows", None)
        self._share_rows = decide() if decide is not None else None      # the same decision a split run takes
        try:
            t = tr(block(self, 0, n))
        finally:
            self._share_rows = prev
        out = self._to_host(t) if as_numpy else t
        self._settle()
        return out

    def _shared_seed(self, seed):
        """A concrete seed: the caller's, or one drawn from numpy's global RNG (so np.random.seed makes runs
        reproducible) -- drawn once per call and, across ranks, taken from rank 0."""
        if seed is None:
            seed = int(np.random.randint(0, 2 ** 62, dtype=np.int64))
            if self.shard and _pdist.is_distributed():
                import torch.distributed as dist
                box = [seed]
                dist.broadcast_object_list(box, src=0)
                seed = int(box[0])
        return int(seed)

    # ------------------------------------------------------------------ spectra for a centroid batch
    def _is_detrend(self, process_subgrid):
        return (inspect.ismethod(process_subgrid) and self._is_me(process_subgrid.__self__)
                and process_subgrid.__func__ is CurieGrid.remove_trend_linear)

    def _batch_spectra(self, window, xc_list, yc_list, taper, process_subgrid, kwargs):
        """(k, Phi_dev[n, nbins], sigma_dev) on the device for every centroid.  process_subgrid None and
        self.remove_trend_linear run fused in the gather kernel; any other callable is applied on the
        host window by window (its result is then uploaded as a stack of windows)."""
        torch = _lib.require_cuda()
        kwargs = dict(kwargs)
        power = kwargs.pop("power", 2.0)
        lattice = xc_list if isinstance(xc_list, _plans.Lattice) else None
        if process_subgrid is None or self._is_detrend(process_subgrid):
            if lattice is not None:
                g, iy0, ix0, n = self._windows_lattice_device(window, lattice.xs, lattice.ys)
            else:
                g, iy0, ix0, n = self._windows_device(window, xc_list, yc_list)
            plan = _plans.spectrum_plan(n, self.dx, _plans.taper_vector(taper, n, **kwargs))
            phi, sig = plan.run(g, iy0, ix0, process_subgrid is not None, power,
                                share_rows=getattr(self, "_share_rows", None))
            return plan.k, phi, sig
        n_probe = _plans.window_nodes(window, self.dx)[1] * 2 + 1
        filt = self._window_filter(process_subgrid, n_probe)
        if filt is not None:
            return self._batch_spectra_filtered(window, xc_list, yc_list, lattice, filt, taper, power, kwargs)
        if lattice is not None:
            xq, yq = np.meshgrid(lattice.xs, lattice.ys)
            xc_list, yc_list = xq.ravel(), yq.ravel()
        wins = [np.asarray(process_subgrid(self.subgrid(window, xc, yc)), dtype=np.float64)
                for xc, yc in zip(xc_list, yc_list)]
        n = wins[0].shape[0]
        for w in wins:
            if w.shape != (n, n):
                raise ValueError("process_subgrid must return square windows of one size, got {}".format(w.shape))
        stack = torch.from_numpy(np.ascontiguousarray(np.concatenate(wins, axis=0))).cuda()
        iy0 = torch.arange(len(wins), dtype=torch.int32, device="cuda") * n
        ix0 = torch.zeros(len(wins), dtype=torch.int32, device="cuda")
        plan = _plans.spectrum_plan(n, self.dx, _plans.taper_vector(taper, n, **kwargs))
        phi, sig = plan.run(stack, iy0, ix0, False, power)
        return plan.k, phi, sig

    FILTER_CHUNK_BYTES = 6 << 30     # device scratch per slice of the filtered-window path
    SENSITIVITY_ROWS = 1 << 22       # perturbed spectra held at once by curie_depth_maps(uncertainty='sensitivity')

    def _batch_spectra_filtered(self, window, xc_list, yc_list, lattice, filt, taper, power, kwargs):
        """Spectra of reduce_to_pole / upward_continuation windows: windows are filtered on the device in slices
        (cb_fft2_filter_windows -> a dense stack) and each stack goes through the spectrum kernel as a grid of
        its own (one window per n rows)."""
        torch = _lib.require_cuda()
        if lattice is not None:
            g, iy0, ix0, n = self._windows_lattice_device(window, lattice.xs, lattice.ys)
        else:
            g, iy0, ix0, n = self._windows_device(window, xc_list, yc_list)
        plan = _plans.spectrum_plan(n, self.dx, _plans.taper_vector(taper, n, **kwargs))
        nwin = int(iy0.numel())
        phi = torch.empty((nwin, plan.nbins), dtype=torch.float64, device="cuda")
        sig = torch.empty_like(phi)
        step = max(1, int(self.FILTER_CHUNK_BYTES // (n * n * 26)))      # 2 complex half planes + the real stack
        for a in range(0, nwin, step):
            b = min(a + step, nwin)
            stack = self._filter_windows_device(filt, g, iy0[a:b].contiguous(), ix0[a:b].contiguous(), n)
            sy = torch.arange(b - a, dtype=torch.int32, device="cuda") * n
            sx = torch.zeros(b - a, dtype=torch.int32, device="cuda")
            p, s = plan.run(stack.view((b - a) * n, n), sy, sx, False, power, share_rows=False)
            phi[a:b], sig[a:b] = p, s
        return plan.k, phi, sig

    # ------------------------------------------------------------------ MAP fit
    def _fit_device(self, k, phi_t, x0, x0_dev=None, max_iter=0, gtol=0.0, ftol=0.0, diagnostics=False):
        torch = _lib.require_cuda()
        n = phi_t.shape[0]
        if len(k) > _lib.MAX_NK:
            raise ValueError("windows with more than {} radial bins (about {} nodes) are not supported by the fit "
                             "kernel; got {} bins".format(_lib.MAX_NK, 2 * _lib.MAX_NK + 5, len(k)))
        opt = _lib.FitOptions()
        lb, ub = self._bound_arrays()
        mu, sg = self._prior_arrays()
        for j in range(4):
            opt.x0[j], opt.lb[j], opt.ub[j] = float(x0[j]), lb[j], ub[j]
            opt.prior_mu[j], opt.prior_sigma[j] = mu[j], sg[j]
        opt.max_iter, opt.gtol, opt.ftol = int(max_iter), float(gtol), float(ftol)
        k_t = _lib.to_device_async(np.ascontiguousarray(k, dtype=np.float64))
        x_t = torch.empty((n, 4), dtype=torch.float64, device="cuda")
        f_t = torch.empty(n, dtype=torch.float64, device="cuda") if diagnostics else None
        it_t = torch.empty(n, dtype=torch.int32, device="cuda") if diagnostics else None
        st_t = torch.empty(n, dtype=torch.int32, device="cuda") if diagnostics else None
        nf_t = torch.empty(n, dtype=torch.int32, device="cuda") if diagnostics else None
        _lib.check(_lib.load().cb_fit_batch(k_t.data_ptr(), k_t.numel(), phi_t.data_ptr(), n, C.byref(opt),
                                            _lib.ptr(x0_dev), x_t.data_ptr(), _lib.ptr(f_t), _lib.ptr(it_t),
                                            _lib.ptr(st_t), _lib.ptr(nf_t), _lib.stream_ptr(torch)))
        if diagnostics:
            return x_t, dict(f=f_t, niter=it_t, status=st_t, nfev=nf_t)
        return x_t

    def _fit_block(self, window, xc_list, yc_list, x0, taper, process_subgrid, kwargs):
        k, phi, _ = self._batch_spectra(window, xc_list, yc_list, taper, process_subgrid, kwargs)
        return self._fit_device(k, phi, x0)

    def optimise(self, window, xc, yc, beta=3.0, zt=1.0, dz=10.0, C=5.0, taper=np.hanning,
                 process_subgrid=None, **kwargs):
        """Optimal (beta, zt, dz, C) for one centroid; ndarray (4,) like `res.x` (optimise.py:232-301)."""
        x = self._fit_block(window, [xc], [yc], (beta, zt, dz, C), taper, process_subgrid, kwargs)[0].cpu().numpy()
        self._settle()
        return x

    def optimise_routine(self, window, xc_list, yc_list, beta=3.0, zt=1.0, dz=10.0, C=5.0, taper=np.hanning,
                         process_subgrid=None, **kwargs):
        """Fit every centroid of the list; returns [beta(n,), zt(n,), dz(n,), C(n,)] (optimise.py:446-509)."""
        return self.parallelise_routine(window, xc_list, yc_list, self.optimise, beta, zt, dz, C, taper,
                                        process_subgrid, **kwargs)

    def optimise_batch(self, window, xc_list, yc_list, *, beta=3.0, zt=1.0, dz=10.0, C=5.0, taper=np.hanning,
                       process_subgrid=None, diagnostics=False, as_numpy=True, shard=None, _rows_out=False, **kwargs):
        """Batched fit returning an (n, 4) array (and optionally solver diagnostics, current device only)."""
        self._check_lists(xc_list, yc_list)
        x0 = (beta, zt, dz, C)
        if diagnostics:
            k, phi, _ = self._batch_spectra(window, xc_list, yc_list, taper, process_subgrid, kwargs)
            x, diag = self._fit_device(k, phi, x0, diagnostics=True)
            if as_numpy:
                x, diag = x.cpu().numpy(), {kk: v.cpu().numpy() for kk, v in diag.items()}
            self._settle()
            return x, diag
        xc_a, yc_a = np.asarray(xc_list), np.asarray(yc_list)
        return self._run_blocks(len(xc_a), lambda o, lo, hi: o._fit_block(window, xc_a[lo:hi], yc_a[lo:hi], x0, taper,
                                                                          process_subgrid, kwargs),
                                shard=shard, as_numpy=as_numpy, transpose=_rows_out,
                                decide=lambda: _plans.share_rows_for_list(self, window, xc_a, yc_a))

    # ------------------------------------------------------------------ map assembly
    def curie_depth_maps(self, window, spacingX=None, spacingY=None, *, beta=3.0, zt=1.0, dz=10.0, C=5.0,
                         taper=np.hanning, process_subgrid=None, uncertainty=None, nsim=100, burnin=1000,
                         x_scale=None, seed=None, as_numpy=True, shard=None, **kwargs):
        """The map workflow of the reference's notebooks in one call: create_centroid_list(window, spacingX,
        spacingY) -> optimise_routine -> reshape to the centroid lattice -> Curie depth zt + dz, optionally
        with uncertainty maps.  The centroid lists are never formed: the lattice axes go to the device and the
        window origins are expanded there; the (rows, cols) maps are views of the solver's output.  Several
        GPUs / ranks split the lattice by whole rows.

        uncertainty=None           MAP maps only
        uncertainty='sensitivity'  + population std over `nsim` perturbed re-fits per window (optimise.py:635-725)
        uncertainty='mcmc'         + posterior mean / std from one Metropolis-Hastings chain per window
                                     (nsim samples after `burnin`), started at the window's MAP estimate

        Returns a dict: xc (cols,), yc (rows,), beta, zt, dz, C, curie_depth (rows, cols) and, with
        uncertainty, `<name>_std` (plus `<name>_mean` for 'mcmc') for the same five quantities."""
        torch = _lib.require_cuda()
        if uncertainty not in (None, "sensitivity", "mcmc"):
            raise ValueError("uncertainty must be None, 'sensitivity' or 'mcmc'")
        xs, ys = self.centroid_lattice(window, spacingX, spacingY)
        rows, cols = ys.size, xs.size
        x0 = (beta, zt, dz, C)
        if uncertainty is not None:
            seed = self._shared_seed(seed)

        def block(o, rlo, rhi):                         # lattice rows [rlo, rhi) -> (rhi-rlo, cols, K)
            lat = _plans.Lattice(xs, ys[rlo:rhi])
            k, phi, sig = o._batch_spectra(window, lat, lat, taper, process_subgrid, kwargs)
            x = o._fit_device(k, phi, x0)
            parts = [x, (x[:, 1] + x[:, 2]).unsqueeze(1)]
            nloc = x.shape[0]
            if uncertainty == "sensitivity":
                std = torch.empty((nloc, 5), dtype=torch.float64, device="cuda")
                step = max(1, o.SENSITIVITY_ROWS // max(int(nsim), 1))        # bound the perturbed-spectra slab
                for a in range(0, nloc, step):
                    b = min(a + step, nloc)
                    smp = o._sensitivity_device(k, phi[a:b], sig[a:b], nsim, x0, seed, chain_offset=rlo * cols + a)
                    smp = torch.cat([smp, (smp[..., 1] + smp[..., 2]).unsqueeze(-1)], dim=-1)
                    std[a:b] = smp.std(dim=1, unbiased=False)
                parts.append(std)
            elif uncertainty == "mcmc":
                mom, acc = o._mh_device(k, phi, nsim, burnin, x_scale, x0, x, seed, 1, chain_offset=rlo * cols)
                parts += [mom, acc.to(torch.float64).unsqueeze(1)]
            return torch.cat(parts, dim=1).view(rhi - rlo, cols, -1)

        # the per-device floor is in windows: rows * cols
        allc = self._run_blocks(rows, block, shard=shard, as_numpy=as_numpy,
                                min_per_worker=max(1, -(-self.MIN_WINDOWS_PER_DEVICE // max(cols, 1))),
                                decide=lambda: _plans.share_rows_for_lattice(self, window, xs, ys))
        names = ("beta", "zt", "dz", "C", "curie_depth")
        maps = {nm: allc[:, :, j] for j, nm in enumerate(names)}
        if uncertainty == "sensitivity":
            for j, nm in enumerate(names):
                maps[nm + "_std"] = allc[:, :, 5 + j]
        elif uncertainty == "mcmc":
            for j, nm in enumerate(names):
                maps[nm + "_mean"] = allc[:, :, 5 + j]
                maps[nm + "_std"] = allc[:, :, 10 + j]
            acc = allc[:, :, 15].reshape(-1)
            self.last_accepted = acc.astype(np.int64) if as_numpy else acc.to(torch.int64)
        maps["xc"], maps["yc"] = xs, ys
        return maps

    # ------------------------------------------------------------------ fan-out (optimise.py:317-444)
    @staticmethod
    def _check_lists(xc_list, yc_list):
        if len(xc_list) != len(yc_list):
            raise ValueError("xc_list and yc_list must be the same size")

    def parallelise_routine(self, window, xc_list, yc_list, func, *args, **kwargs):
        """Evaluate `func(window, xc, yc, *args, **kwargs)` for every centroid.

        `self.optimise`, `self.metropolis_hastings` and `self.sensitivity` are recognised and run as batched
        device launches over the whole list -- spread over min(max_processors, visible GPUs) devices, or over
        the ranks of an initialised torch.distributed group; any other callable runs as a plain serial loop
        (the reference's `max_processors == 1` branch) -- nothing is ever forked.  Output containers follow the
        reference: 1-D results -> list of 4 arrays (n,); 2-D results -> list[4] of list[n] arrays."""
        n = len(xc_list)
        self._check_lists(xc_list, yc_list)
        nprocs = self.max_processors
        if nprocs < 1:
            raise ValueError("{} processors is invalid, specify a positive integer value".format(nprocs))

        if self._is_own(func, "optimise"):
            b = inspect.signature(CurieOptimise.optimise).bind(self, window, 0.0, 0.0, *args, **kwargs)
            b.apply_defaults()
            a = b.arguments
            x = self.optimise_batch(window, xc_list, yc_list, beta=a["beta"], zt=a["zt"], dz=a["dz"], C=a["C"],
                                    taper=a["taper"], process_subgrid=a["process_subgrid"], _rows_out=True, **a["kwargs"])
            return list(x)                  # transposed on the device: four (n,) rows

        if self._is_own(func, "metropolis_hastings"):
            b = inspect.signature(CurieOptimise.metropolis_hastings).bind(self, window, 0.0, 0.0, *args, **kwargs)
            b.apply_defaults()
            a = b.arguments
            chains = self.metropolis_hastings_batch(
                window, xc_list, yc_list, a["nsim"], a["burnin"], x_scale=a["x_scale"], beta=a["beta"], zt=a["zt"],
                dz=a["dz"], C=a["C"], taper=a["taper"], process_subgrid=a["process_subgrid"], seed=a["seed"],
                reduce="chains", **a["kwargs"])
            return [[chains[i, j] for i in range(n)] for j in range(4)]

        if self._is_own(func, "sensitivity"):
            b = inspect.signature(CurieOptimise.sensitivity).bind(self, window, 0.0, 0.0, *args, **kwargs)
            b.apply_defaults()
            a = b.arguments
            x = self.sensitivity_batch(window, xc_list, yc_list, a["nsim"], beta=a["beta"], zt=a["zt"], dz=a["dz"],
                                       C=a["C"], taper=a["taper"], process_subgrid=a["process_subgrid"],
                                       seed=a["seed"], **a["kwargs"])
            return [[x[i, :, j] for i in range(n)] for j in range(4)]

        res = None
        xOpt = [[] for _ in range(n)]
        for i in range(n):
            res = func(window, xc_list[i], yc_list[i], *args, **kwargs)
            xOpt[i] = res
        ndim = np.array(res).ndim
        if ndim == 1:
            return list(np.vstack(xOpt).T)
        elif ndim > 1:
            out = list(np.hstack(xOpt))
            return [np.split(o, n) for o in out]
        raise ValueError("Cannot determine shape of output")

    def _is_own(self, func, name):
        return (inspect.ismethod(func) and func.__self__ is self
                and func.__func__ is getattr(CurieOptimise, name))

    # ------------------------------------------------------------------ Metropolis-Hastings
    def _mh_device(self, k, phi_t, nsim, burnin, x_scale, x0, x0_dev, seed, reduce_mode, chain_offset=0):
        torch = _lib.require_cuda()
        n = phi_t.shape[0]
        if len(k) > _lib.MAX_NK:
            raise ValueError("windows with more than {} radial bins are not supported by the MCMC kernel; got {}"
                             .format(_lib.MAX_NK, len(k)))
        opt = _lib.MhOptions()
        mu, sg = self._prior_arrays()
        xs = np.ones(4) if x_scale is None else np.asarray(x_scale, dtype=np.float64)
        for j in range(4):
            opt.x0[j], opt.x_scale[j] = float(x0[j]), float(xs[j])
            opt.prior_mu[j], opt.prior_sigma[j] = mu[j], sg[j]
        opt.nsim, opt.burnin = int(nsim), int(burnin)
        if seed is None:
            seed = int(np.random.randint(0, 2 ** 62, dtype=np.int64))
        opt.seed, opt.chain_offset, opt.reduce_mode = int(seed) & (2 ** 64 - 1), int(chain_offset), reduce_mode
        k_t = torch.from_numpy(np.ascontiguousarray(k, dtype=np.float64)).cuda()
        shape = (n, 4, int(nsim)) if reduce_mode == 0 else (n, 10)
        out = torch.empty(shape, dtype=torch.float64, device="cuda")
        acc = torch.empty(n, dtype=torch.int64, device="cuda")
        _lib.check(_lib.load().cb_mh_batch(k_t.data_ptr(), k_t.numel(), phi_t.data_ptr(), n, C.byref(opt),
                                           _lib.ptr(x0_dev), out.data_ptr(), acc.data_ptr(),
                                           _lib.stream_ptr(torch)))
        return out, acc

    def metropolis_hastings(self, window, xc, yc, nsim, burnin, x_scale=None, beta=3.0, zt=1.0, dz=10.0, C=5.0,
                            taper=np.hanning, process_subgrid=None, seed=None, **kwargs):
        """MCMC ensemble for one centroid: [beta(nsim,), zt(nsim,), dz(nsim,), C(nsim,)] (optimise.py:511-633).
        Burn-in is the reference's tempered greedy climb; bounds are not applied, as in the reference."""
        k, phi, _ = self._batch_spectra(window, [xc], [yc], taper, process_subgrid, kwargs)
        out, _ = self._mh_device(k, phi, nsim, burnin, x_scale, (beta, zt, dz, C), None, seed, 0)
        res = list(out[0].cpu().numpy())
        self._settle()
        return res

    def metropolis_hastings_batch(self, window, xc_list, yc_list, nsim, burnin, *, x_scale=None, beta=3.0, zt=1.0,
                                  dz=10.0, C=5.0, x0=None, taper=np.hanning, process_subgrid=None, seed=None,
                                  reduce="moments", chain_offset=0, as_numpy=True, shard=None, **kwargs):
        """One chain per centroid in batched launches.
        reduce='chains'  -> array (n, 4, nsim)
        reduce='moments' -> array (n, 10): mean then std of (beta, zt, dz, C, zt+dz), streamed on the device
        `x0` (n, 4) gives per-window starting points (e.g. the MAP estimates).  The chain id of a window is
        chain_offset + its index in the list, so the result does not depend on how many GPUs share the list."""
        torch = _lib.require_cuda()
        self._check_lists(xc_list, yc_list)
        if reduce not in ("moments", "chains"):
            raise ValueError("reduce must be 'moments' or 'chains'")
        n = len(xc_list)
        x0_h = None
        if x0 is not None:
            if tuple(getattr(x0, "shape", np.shape(x0))) != (n, 4):
                raise ValueError("x0 must have shape (n, 4)")
            x0_h = x0
        seed = self._shared_seed(seed)
        xc_a, yc_a = np.asarray(xc_list), np.asarray(yc_list)
        mode = 0 if reduce == "chains" else 1
        start = (beta, zt, dz, C)

        def block(o, lo, hi):
            k, phi, _ = o._batch_spectra(window, xc_a[lo:hi], yc_a[lo:hi], taper, process_subgrid, kwargs)
            x0_dev = None
            if x0_h is not None:
                part = x0_h[lo:hi]
                x0_dev = (part.to("cuda") if hasattr(part, "data_ptr") else
                          torch.from_numpy(np.ascontiguousarray(part, dtype=np.float64)).cuda()).contiguous()
                assert x0_dev.dtype == torch.float64
            out, acc = o._mh_device(k, phi, nsim, burnin, x_scale, start, x0_dev, seed, mode, chain_offset + lo)
            # the acceptance counts ride along as one more column
            return torch.cat([out.reshape(hi - lo, -1), acc.to(torch.float64).unsqueeze(1)], dim=1)

        res = self._run_blocks(n, block, shard=shard, as_numpy=as_numpy,
                               decide=lambda: _plans.share_rows_for_list(self, window, xc_a, yc_a))
        acc = res[:, -1]
        self.last_accepted = acc.astype(np.int64) if as_numpy else acc.to(torch.int64)
        res = res[:, :-1]
        if mode == 0:
            res = res.reshape(n, 4, int(nsim))
        return np.ascontiguousarray(res) if as_numpy else res.contiguous()

    # ------------------------------------------------------------------ sensitivity (optimise.py:635-725)
    def _sensitivity_device(self, k, phi_t, sig_t, nsim, x0, seed, chain_offset=0):
        """nsim perturbed re-fits for every row of phi_t: returns a cuda tensor (n, nsim, 4).
        Spectra are perturbed on the device (cb_perturb_spectra); the prior means are redrawn on the host
        with the priors' own `rvs` so that any scipy.stats distribution works (optimise.py:709-711)."""
        torch = _lib.require_cuda()
        lib = _lib.load()
        n, nk = int(phi_t.shape[0]), int(phi_t.shape[1])
        nsim = int(nsim)
        if seed is None:
            seed = int(np.random.randint(0, 2 ** 62, dtype=np.int64))
        rows = torch.empty((n * nsim, nk), dtype=torch.float64, device="cuda")
        _lib.check(lib.cb_perturb_spectra(phi_t.data_ptr(), sig_t.data_ptr(), n, nk, nsim, int(seed) & (2 ** 64 - 1),
                                          int(chain_offset), rows.data_ptr(), _lib.stream_ptr(torch)))
        mu_rows = None
        if any(self.prior_pdf[key] is not None for key in _KEYS):
            mu = np.full((n * nsim, 4), np.nan)
            with _RVS_LOCK:                       # scipy's rvs draws from numpy's global RNG
                for j, key in enumerate(_KEYS):
                    pdf = self.prior_pdf[key]
                    if pdf is not None:
                        mu[:, j] = np.asarray(pdf.rvs(size=n * nsim), dtype=np.float64)
            mu_rows = torch.from_numpy(mu).cuda()
        opt = _lib.FitOptions()
        lb, ub = self._bound_arrays()
        pm, ps = self._prior_arrays()
        for j in range(4):
            opt.x0[j], opt.lb[j], opt.ub[j] = float(x0[j]), lb[j], ub[j]
            opt.prior_mu[j], opt.prior_sigma[j] = pm[j], ps[j]
        k_t = torch.from_numpy(np.ascontiguousarray(k, dtype=np.float64)).cuda()
        x_t = torch.empty((n * nsim, 4), dtype=torch.float64, device="cuda")
        _lib.check(lib.cb_fit_batch_rows(k_t.data_ptr(), k_t.numel(), rows.data_ptr(), n * nsim, C.byref(opt), None,
                                         _lib.ptr(mu_rows), x_t.data_ptr(), None, None, None, None,
                                         _lib.stream_ptr(torch)))
        return x_t.view(n, nsim, 4)

    def sensitivity(self, window, xc, yc, nsim, beta=3.0, zt=1.0, dz=10.0, C=5.0, taper=np.hanning,
                    process_subgrid=None, seed=None, **kwargs):
        """Monte-Carlo sensitivity of (beta, zt, dz, C) for one centroid: `nsim` re-fits with the spectrum
        drawn from N(Phi, sigma_Phi) and the prior means redrawn from their distributions
        (optimise.py:635-725).  Returns [beta(nsim,), zt(nsim,), dz(nsim,), C(nsim,)]; all nsim fits run
        in one launch.  `seed=None` takes the seed from numpy's global RNG."""
        k, phi, sig = self._batch_spectra(window, [xc], [yc], taper, process_subgrid, kwargs)
        x = self._sensitivity_device(k, phi, sig, nsim, (beta, zt, dz, C), seed)
        res = list(np.ascontiguousarray(x[0].cpu().numpy().T))
        self._settle()
        return res

    def sensitivity_batch(self, window, xc_list, yc_list, nsim, *, beta=3.0, zt=1.0, dz=10.0, C=5.0,
                          taper=np.hanning, process_subgrid=None, seed=None, reduce="samples", chain_offset=0,
                          as_numpy=True, shard=None, **kwargs):
        """sensitivity for every centroid in batched launches of n*nsim fits.
        reduce='samples' -> (n, nsim, 4);  reduce='moments' -> (n, 8): mean then std (population) of
        (beta, zt, dz, C) over the simulations -- the uncertainty maps of the reference's notebooks."""
        torch = _lib.require_cuda()
        self._check_lists(xc_list, yc_list)
        if reduce not in ("samples", "moments"):
            raise ValueError("reduce must be 'samples' or 'moments'")
        seed = self._shared_seed(seed)
        xc_a, yc_a = np.asarray(xc_list), np.asarray(yc_list)
        start = (beta, zt, dz, C)

        def block(o, lo, hi):
            k, phi, sig = o._batch_spectra(window, xc_a[lo:hi], yc_a[lo:hi], taper, process_subgrid, kwargs)
            x = o._sensitivity_device(k, phi, sig, nsim, start, seed, chain_offset + lo)
            if reduce == "moments":
                return torch.cat([x.mean(dim=1), x.std(dim=1, unbiased=False)], dim=1)
            return x

        # one device holds n * nsim fits: the per-device floor is in fits
        return self._run_blocks(len(xc_a), block, shard=shard, as_numpy=as_numpy,
                                min_per_worker=max(1, self.MIN_WINDOWS_PER_DEVICE // max(int(nsim), 1)),
                                decide=lambda: _plans.share_rows_for_list(self, window, xc_a, yc_a))
