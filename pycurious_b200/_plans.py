"""Host-side planning for the device kernels.

Everything here is O(window) work done once per (window size, grid spacing, taper)
and cached: centroid -> node index rule, taper vectors, the radial-bin membership
CSR and wavenumber axis, Radon ray end points.  The membership and the end
points are computed with the reference's own numpy expressions so that the
integers the kernels consume are bit-identical to what pycurious would use
(pycurious/grid.py:135-139, :242-245, :277-282; src/radon.pyx:30-96).
"""
from __future__ import annotations

import ctypes as C
import warnings
import os
import weakref

import numpy as np

from . import _lib


# ----------------------------------------------------------------------------
# centroid -> window origin  (CurieGrid.subgrid, pycurious/grid.py:130-155)
# ----------------------------------------------------------------------------
def nearest_node(coords, q):
    """Vectorised `np.abs(coords - q).argmin()` for increasing `coords`: the distance profile
    is V-shaped, so the argmin is one of the two bracketing nodes and ties go to the lower
    index, exactly like argmin's first-minimum rule."""
    q = np.asarray(q, dtype=np.float64)
    hi = np.clip(np.searchsorted(coords, q, side="left"), 0, coords.size - 1)
    lo = np.clip(hi - 1, 0, coords.size - 1)
    pick_lo = np.abs(coords[lo] - q) <= np.abs(coords[hi] - q)
    return np.where(pick_lo, lo, hi).astype(np.int64)


def nearest_nodes_of_list(coords, q):
    """nearest_node for one axis of a centroid LIST.  Lists come from create_centroid_list (a raveled meshgrid,
    grid.py:196-197) or are contiguous blocks of one (a rank's share): along x the values repeat with the lattice's
    column count, along y they come in runs.  Either structure is detected with O(n) comparisons and only the distinct
    values are searched (two binary searches per centroid are the largest host cost of a block otherwise);
    anything else takes the plain path.  The result is nearest_node(coords, q) in every case."""
    n = q.size
    if n < 4096:
        return nearest_node(coords, q)
    change = np.flatnonzero(q[1:] != q[:-1])
    if change.size * 16 <= n:                                   # runs of equal values (the slow axis)
        starts = np.concatenate(([0], change + 1))
        return np.repeat(nearest_node(coords, q[starts]), np.diff(np.append(starts, n)))
    rep = np.flatnonzero(q[1:] == q[0])
    if rep.size:                                                # period p (the fast axis)
        p = int(rep[0]) + 1
        if 2 * p <= n and (q[p:] == q[:-p]).all():
            return np.resize(nearest_node(coords, q[:p]), n)
    return nearest_node(coords, q)


def _lattice_shape(xc, yc):
    """(rows, cols) when (xc, yc) is a raveled meshgrid with x fastest (what create_centroid_list returns,
    grid.py:196-197), else None.  O(n) comparisons instead of two searches per centroid."""
    n = xc.size
    if n < 4096 or yc.size != n:
        return None
    rep = np.flatnonzero(xc[1:] == xc[0])
    cols = int(rep[0]) + 1 if rep.size else n
    if n % cols:
        return None
    rows = n // cols
    if not (xc.reshape(rows, cols) == xc[:cols]).all():
        return None
    if not (yc.reshape(rows, cols) == yc[::cols, None]).all():
        return None
    return rows, cols


def window_nodes(window, dx):
    nw = int(round(window / dx))
    return nw, nw // 2


def window_origins(g, window, xc, yc):
    """Top-left nodes (iy0, ix0) and edge n of the windows CurieGrid.subgrid would cut around each
    centroid; raises the reference's ValueErrors for the first offending centroid."""
    xc = np.atleast_1d(np.asarray(xc, dtype=np.float64))
    yc = np.atleast_1d(np.asarray(yc, dtype=np.float64))
    outside = (xc < g.xmin) | (xc > g.xmax) | (yc < g.ymin) | (yc > g.ymax)
    if outside.any():
        i = int(np.flatnonzero(outside)[0])
        raise ValueError("Point {} outside data range".format((xc[i], yc[i])))
    _, half = window_nodes(window, g.dx)
    lat = _lattice_shape(xc, yc)
    if lat is not None:                  # create_centroid_list output: resolve the distinct coordinates only
        rows, cols = lat
        ix = np.tile(nearest_node(g.xcoords, xc[:cols]), rows)
        iy = np.repeat(nearest_node(g.ycoords, yc[::cols]), cols)
    else:                                # blocks of a lattice (a rank's share) keep its structure per axis
        ix = nearest_nodes_of_list(g.xcoords, xc)
        iy = nearest_nodes_of_list(g.ycoords, yc)
    bad = (ix - half < 0) | (ix + half + 1 > g.nx) | (iy - half < 0) | (iy + half + 1 > g.ny)
    if bad.any():
        i = int(np.flatnonzero(bad)[0])
        raise ValueError("Window size {} at centroid {} exceeds the data range".format(window, (xc[i], yc[i])))
    return (iy - half).astype(np.int32), (ix - half).astype(np.int32), 2 * half + 1


def share_rows_for_list(g, window, xc, yc, sample=4096):
    """The spectrum path (shared row spectra or window by window) for a centroid list that is about to be split
    into blocks -- over ranks, devices or scratch slices.  SpectrumPlan.run decides per call from the row reuse
    of the windows it is given, and the two paths agree to ~1e-14 but not bit for bit; a split list must therefore
    take ONE decision, from the same data on every rank: the row reuse of its first `sample` windows."""
    if os.environ.get("CB_SPEC_SHARE_ROWS", "auto") != "auto":
        return None                        # forced for every call by the environment (SpectrumPlan.run reads it)
    xc, yc = np.asarray(xc, dtype=np.float64), np.asarray(yc, dtype=np.float64)
    m = int(min(xc.size, sample))
    if m <= 1:
        return False
    iy0, ix0, nw = window_origins(g, window, xc[:m], yc[:m])
    groups = np.unique(ix0).size
    span = int(iy0.max()) - int(iy0.min()) + nw
    return float(m) * nw / (float(groups) * span) >= SpectrumPlan.SHARE_MIN_REUSE


def share_rows_for_lattice(g, window, xs, ys):
    """share_rows_for_list for the lattice xs x ys (x fastest)."""
    if os.environ.get("CB_SPEC_SHARE_ROWS", "auto") != "auto":
        return None
    if len(xs) * len(ys) <= 1:
        return False
    half = window_nodes(window, g.dx)[1]
    iy = nearest_node(g.ycoords, np.asarray(ys, dtype=np.float64))
    span = int(iy.max()) - int(iy.min()) + 2 * half + 1
    return float(len(ys)) * (2 * half + 1) / span >= SpectrumPlan.SHARE_MIN_REUSE


class Lattice:
    """The centroid lattice xs x ys (x fastest) standing in for create_centroid_list's two raveled lists:
    the batch entry points accept it in place of (xc_list, yc_list) and never expand it on the host."""

    def __init__(self, xs, ys):
        self.xs = np.atleast_1d(np.asarray(xs, dtype=np.float64))
        self.ys = np.atleast_1d(np.asarray(ys, dtype=np.float64))

    @property
    def shape(self):
        return self.ys.size, self.xs.size

    def __len__(self):
        return self.xs.size * self.ys.size


def _first_offender(bad_x, bad_y):
    """(row, col) of the first flagged centroid in the x-fastest order of create_centroid_list."""
    if bad_y[0]:
        return 0, 0
    if bad_x.any():
        return 0, int(np.flatnonzero(bad_x)[0])
    return int(np.flatnonzero(bad_y)[0]), 0


def lattice_origins(g, window, xs, ys):
    """window_origins for the lattice xs x ys (x fastest) without forming the centroid lists: per-axis
    top-left nodes (iyu[rows], ixu[cols]) and the window edge.  The window of centroid (r, c) starts at
    (iyu[r], ixu[c]); validation is O(rows + cols), first offender and messages as in grid.py:131-153."""
    xs = np.atleast_1d(np.asarray(xs, dtype=np.float64))
    ys = np.atleast_1d(np.asarray(ys, dtype=np.float64))
    if xs.size == 0 or ys.size == 0:
        raise ValueError("empty centroid lattice")
    bx, by = (xs < g.xmin) | (xs > g.xmax), (ys < g.ymin) | (ys > g.ymax)
    if bx.any() or by.any():
        r, c = _first_offender(bx, by)
        raise ValueError("Point {} outside data range".format((xs[c], ys[r])))
    _, half = window_nodes(window, g.dx)
    ix, iy = nearest_node(g.xcoords, xs), nearest_node(g.ycoords, ys)
    wx = (ix - half < 0) | (ix + half + 1 > g.nx)
    wy = (iy - half < 0) | (iy + half + 1 > g.ny)
    if wx.any() or wy.any():
        r, c = _first_offender(wx, wy)
        raise ValueError("Window size {} at centroid {} exceeds the data range".format(window, (xs[c], ys[r])))
    return (iy - half).astype(np.int32), (ix - half).astype(np.int32), 2 * half + 1


# ----------------------------------------------------------------------------
# radial bins  (CurieGrid._taper_spectrum / _FFT_spectrum, grid.py:242-245, :277-286)
# ----------------------------------------------------------------------------
_bin_cache = {}


def radial_bins(nr, dx, scale=0.001):
    """(kbins, k, bin_ptr, bin_idx): bin edges, mean wavenumber per bin, and the CSR of member
    indices a*nr+b in the fftshifted plane.  Same numpy expressions as the reference, so membership
    (decided by floating-point rounding of hypot/arange) is bit-identical."""
    key = (int(nr), float(dx), float(scale))
    hit = _bin_cache.get(key)
    if hit is not None:
        return hit
    dk = 2.0 * np.pi / (nr - 1) / (dx * scale)
    kbins = np.arange(dk, dk * nr / 2, dk)
    nbins = kbins.size - 1
    c = int((nr - 1) // 2)
    a, b = np.mgrid[0:nr, 0:nr]
    kk = np.hypot((a - c) * dk, (b - c) * dk)
    flat = kk.ravel()
    k = np.empty(max(nbins, 0))
    ptr = np.zeros(max(nbins, 0) + 1, dtype=np.int32)
    members = []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for i in range(nbins):
            m = np.flatnonzero(np.logical_and(flat >= kbins[i], flat <= kbins[i + 1])).astype(np.int32)
            members.append(m)
            ptr[i + 1] = ptr[i] + m.size
            k[i] = flat[m].mean() if m.size else np.nan
    idx = np.concatenate(members) if members else np.zeros(0, dtype=np.int32)
    out = (kbins, k, ptr, np.ascontiguousarray(idx, dtype=np.int32))
    _bin_cache[key] = out
    return out


# ----------------------------------------------------------------------------
# spectrum plans
# ----------------------------------------------------------------------------
class SpectrumPlan:
    """Owns a cb_spectrum_plan plus its scratch slab on the current device."""

    def __init__(self, nw, dx, taper_vec):
        torch = _lib.require_cuda()
        lib = _lib.load()
        self.nw = int(nw)
        self.kbins, self.k, ptr, idx = radial_bins(nw, dx)
        self.nbins = int(self.k.size)
        if self.nbins < 1:
            raise ValueError("window of {} nodes is too small for a radial spectrum".format(nw))
        self._taper = None if taper_vec is None else np.ascontiguousarray(taper_vec, dtype=np.float64)
        tp = None if self._taper is None else self._taper.ctypes.data
        handle = C.c_void_p()
        _lib.check(lib.cb_spectrum_plan_create(self.nw, self.nbins, tp, tp, ptr.ctypes.data, idx.ctypes.data,
                                               C.byref(handle)))
        self.handle = handle
        self._finalizer = weakref.finalize(self, lib.cb_spectrum_plan_destroy, handle)
        self._ws_bytes = int(lib.cb_spectrum_workspace_bytes(handle))
        self.device = torch.cuda.current_device()
        # scratch (per-CTA slabs, shared row spectra) belongs to the stream it is used on: two run() calls under
        # different torch.cuda.stream contexts must not share it
        self._scratch = {}
        self.last_path = "direct"

    def _stream_scratch(self, torch, stream, rowspec_bytes=0):
        """[workspace, rowspec] of the CUDA stream `stream` (created / grown on demand)."""
        sc = self._scratch.get(stream)
        if sc is None:
            if len(self._scratch) >= 8:                 # a caller cycling through many streams: drop the idle ones
                torch.cuda.synchronize()
                self._scratch.clear()
            sc = [torch.empty(self._ws_bytes, dtype=torch.uint8, device="cuda"), None]
            self._scratch[stream] = sc
        if rowspec_bytes and (sc[1] is None or sc[1].numel() < rowspec_bytes):
            sc[1] = None
            sc[1] = torch.empty(rowspec_bytes, dtype=torch.uint8, device="cuda")
        return sc

    @property
    def workspace(self):
        torch = _lib.require_cuda()
        return self._stream_scratch(torch, _lib.stream_ptr(torch))[0]

    def info(self):
        info = _lib.SpectrumPlanInfo()
        _lib.check(_lib.load().cb_spectrum_plan_get_info(self.handle, C.byref(info)))
        return info

    def folded_bins(self):
        info = self.info()
        ptr = np.empty(self.nbins + 1, dtype=np.int32)
        idx = np.empty(info.nnz_folded, dtype=np.int32)
        _lib.check(_lib.load().cb_spectrum_plan_get_bins(self.handle, ptr.ctypes.data, idx.ctypes.data))
        return ptr, idx

    # windows * nw rows would be transformed without sharing; the shared path transforms groups * rows
    SHARE_MIN_REUSE = 2.0

    def _prepare_groups(self, lib, torch, n, nx, iy0_t, ix0_t, share_rows):
        """Host-side planning of the shared-row path for one centroid list: "direct" when the windows do not
        share enough rows, else a list of slices (column groups that fit the row-spectrum scratch budget), each
        with its device-resident group map, group origins and processing order."""
        if share_rows is False or n <= 1:
            return "direct"
        iy0_h = getattr(iy0_t, "_cb_host", None)
        ix0_h = getattr(ix0_t, "_cb_host", None)
        if iy0_h is None or ix0_h is None:
            iy0_h, ix0_h = iy0_t.cpu().numpy(), ix0_t.cpu().numpy()
        # distinct ix0 values and each window's index into them (counting sort: ix0 < nx)
        present = np.zeros(int(nx) + 1, dtype=bool)
        present[ix0_h] = True
        uniq = np.flatnonzero(present).astype(np.int32)
        inv = (np.cumsum(present) - 1)[ix0_h].astype(np.int32)
        row_lo, row_hi = int(iy0_h.min()), int(iy0_h.max()) + self.nw
        reuse = float(n) * self.nw / (float(uniq.size) * (row_hi - row_lo))
        if not (share_rows or reuse >= self.SHARE_MIN_REUSE):
            return "direct"
        budget = int(float(os.environ.get("CB_SPEC_ROWSPEC_MAX_GB", "24")) * 2 ** 30)
        per_group = int(lib.cb_spectrum_rowspec_bytes(self.handle, 1, row_lo, row_hi))
        gmax = max(1, budget // max(per_group, 1))
        if uniq.size <= gmax:
            slices = [(0, uniq.size, None)]
        else:                                 # too many column groups for one scratch buffer: sort by group, go in slices
            order = np.argsort(inv, kind="stable")
            bounds = np.searchsorted(inv[order], np.arange(0, uniq.size + gmax, gmax))
            slices = [(g0, min(g0 + gmax, uniq.size), order[bounds[i]:bounds[i + 1]])
                      for i, g0 in enumerate(range(0, uniq.size, gmax))]
        prep = []
        for g0, g1, sel in slices:
            c = {"ngroups": g1 - g0, "sel_t": None}
            if sel is None:
                iy_c, gof, iy_hc = iy0_t, inv, iy0_h
            else:
                c["sel_t"] = _lib.to_device_async(sel)
                c["iy_c"], c["ix_c"] = iy0_t.index_select(0, c["sel_t"]), ix0_t.index_select(0, c["sel_t"])
                iy_c, gof, iy_hc = c["iy_c"], inv[sel] - g0, iy0_h[sel]
            c["lo"], c["hi"] = int(iy_hc.min()), int(iy_hc.max()) + self.nw
            c["need"] = int(lib.cb_spectrum_rowspec_bytes(self.handle, g1 - g0, c["lo"], c["hi"]))
            c["gof_t"] = _lib.to_device_async(np.ascontiguousarray(gof, dtype=np.int32))
            c["gix_t"] = _lib.to_device_async(uniq[g0:g1])
            # processing order: by (group, iy0), sorted on the device (plumbing, 0.1 ms for 152,100 keys)
            c["order_t"] = torch.argsort(c["gof_t"].to(torch.int64) * (1 << 20) + iy_c.to(torch.int64)).to(torch.int32)
            prep.append(c)
        return prep

    def run(self, grid_t, iy0_t, ix0_t, detrend, power, share_rows=None):
        """grid_t: 2-D float64 cuda tensor; iy0_t/ix0_t: int32 cuda tensors -> (Phi, sigma) cuda [n, nbins].

        Windows on a lattice (many windows per distinct ix0, overlapping in y) go through
        cb_spectrum_batch_shared_rows, which transforms every needed grid row once per distinct ix0;
        scattered windows go through cb_spectrum_batch.  `share_rows` forces the choice (tests)."""
        torch = _lib.require_cuda()
        lib = _lib.load()
        n = int(iy0_t.numel())
        phi = torch.empty((n, self.nbins), dtype=torch.float64, device="cuda")
        sig = torch.empty((n, self.nbins), dtype=torch.float64, device="cuda")
        if n == 0:
            return phi, sig
        ny, nx = grid_t.shape
        stream = _lib.stream_ptr(torch)
        if share_rows is None:
            share_rows = os.environ.get("CB_SPEC_SHARE_ROWS", "auto")
            share_rows = None if share_rows == "auto" else share_rows not in ("0", "off", "no")
        # The grouping of a centroid list depends on (iy0, ix0) only; it is kept on the origin tensor so that a
        # caller who reuses its origin tensors (a survey processed with several tapers / powers, the timed loop
        # of bench.py) pays the host-side sort and the small uploads once.
        key = (self.nw, int(ny), int(nx), bool(share_rows) if share_rows is not None else None,
               os.environ.get("CB_SPEC_ROWSPEC_MAX_GB", "24"))
        cached = getattr(iy0_t, "_cb_groups", None)
        prep = cached[1] if cached is not None and cached[0] == key and cached[2] is ix0_t else None
        if prep is None:
            prep = self._prepare_groups(lib, torch, n, nx, iy0_t, ix0_t, share_rows)
            try:
                iy0_t._cb_groups = (key, prep, ix0_t)     # (origin tensors are not to be modified in place)
            except AttributeError:
                pass
        if prep is None or prep == "direct":
            ws = self._stream_scratch(torch, stream)[0]
            _lib.check(lib.cb_spectrum_batch(
                self.handle, grid_t.data_ptr(), ny, nx, grid_t.stride(0), iy0_t.data_ptr(), ix0_t.data_ptr(), n,
                1 if detrend else 0, float(power), phi.data_ptr(), sig.data_ptr(),
                ws.data_ptr(), ws.numel(), stream))
            self.last_path = "direct"
            return phi, sig

        for c in prep:
            whole = c["sel_t"] is None
            iy_c = iy0_t if whole else c["iy_c"]
            ix_c = ix0_t if whole else c["ix_c"]
            out_phi = phi if whole else torch.empty((int(iy_c.numel()), self.nbins), dtype=torch.float64, device="cuda")
            out_sig = sig if whole else torch.empty_like(out_phi)
            ws, rowspec = self._stream_scratch(torch, stream, c["need"])
            _lib.check(lib.cb_spectrum_batch_shared_rows(
                self.handle, grid_t.data_ptr(), ny, nx, grid_t.stride(0), iy_c.data_ptr(), ix_c.data_ptr(),
                int(iy_c.numel()), c["gof_t"].data_ptr(), c["gix_t"].data_ptr(), c["ngroups"], c["order_t"].data_ptr(),
                c["lo"], c["hi"], 1 if detrend else 0,
                float(power), out_phi.data_ptr(), out_sig.data_ptr(), rowspec.data_ptr(), rowspec.numel(),
                ws.data_ptr(), ws.numel(), stream))
            if not whole:
                phi.index_copy_(0, c["sel_t"], out_phi)
                sig.index_copy_(0, c["sel_t"], out_sig)
        self.last_path = "shared_rows"
        return phi, sig


_plan_cache = {}


def taper_vector(taper, n, **kwargs):
    """The 1-D taper the reference would build (grid.py:233-237); None stays None."""
    if taper is None:
        return None
    return np.asarray(taper(n, **kwargs), dtype=np.float64)


def spectrum_plan(nw, dx, taper_vec):
    torch = _lib.require_cuda()
    key = (int(nw), float(dx), None if taper_vec is None else taper_vec.tobytes(), torch.cuda.current_device())
    plan = _plan_cache.get(key)
    if plan is None:
        if len(_plan_cache) > 32:
            _plan_cache.clear()
        plan = SpectrumPlan(nw, dx, taper_vec)
        _plan_cache[key] = plan
    return plan


class FftPlan:
    """1-D transform plan of cb_fft2_filter (one per line length and device)."""

    def __init__(self, n):
        _lib.require_cuda()
        lib = _lib.load()
        handle = C.c_void_p()
        _lib.check(lib.cb_fft_plan_create(int(n), C.byref(handle)))
        self.handle = handle
        self.n = int(n)
        self._finalizer = weakref.finalize(self, lib.cb_fft_plan_destroy, handle)


_fft_plan_cache = {}


def fft_plan(n):
    torch = _lib.require_cuda()
    key = (int(n), torch.cuda.current_device())
    plan = _fft_plan_cache.get(key)
    if plan is None:
        if len(_fft_plan_cache) > 32:
            _fft_plan_cache.clear()
        plan = FftPlan(n)
        _fft_plan_cache[key] = plan
    return plan


# ----------------------------------------------------------------------------
# Radon ray end points  (src/radon.pyx:30-96)
# ----------------------------------------------------------------------------
def radon_endpoints(nw, theta):
    """Tx, Rx of shape (nw, 2) for one angle, evaluated with numpy scalars in the reference's
    operation order so the doubles handed to the tracer are the same."""
    nx = ny = nw
    idx = np.arange(nw)
    Tx = np.zeros((nw, 2))
    Rx = np.zeros((nw, 2))
    th = np.float64(theta)
    if th == 0.0:
        Tx[:, 0] = idx
        Rx[:, 0] = idx
        Rx[:, 1] = ny - 1
        return Tx, Rx
    if np.abs(th - np.pi / 2) < 1.e-10:
        Tx[:, 1] = idx
        Rx[:, 0] = nx - 1
        Rx[:, 1] = idx
        return Tx, Rx
    co, si = np.cos(th), np.sin(th)
    xk = nw / 2 * (1. - co) + idx * co
    yk = nw / 2 * (1. - si) + idx * si
    if th < np.pi / 2:
        tn = np.tan(th)
        x0 = np.zeros(nw)
        y0 = yk + (xk - x0) / tn
        over = y0 > ny - 1
        y0 = np.where(over, ny - 1, y0)
        x0 = np.where(over, xk - (y0 - yk) * tn, x0)
        x1 = np.full(nw, nx - 1.0)
        y1 = yk - (x1 - xk) / tn
        under = y1 < 0.0
        y1 = np.where(under, 0.0, y1)
        x1 = np.where(under, xk + (yk - y1) * tn, x1)
    else:
        tq = np.tan(th - np.pi / 2.0)
        x0 = np.full(nw, nx - 1.0)
        y0 = yk + (x0 - xk) * tq
        over = y0 > ny - 1
        y0 = np.where(over, ny - 1, y0)
        x0 = np.where(over, xk + (y0 - yk) / tq, x0)
        x1 = np.zeros(nw)
        y1 = yk - (xk - x1) * tq
        under = y1 < 0.0
        y1 = np.where(under, 0.0, y1)
        x1 = np.where(under, xk - (yk - y1) / tq, x1)
    Tx[:, 0], Tx[:, 1], Rx[:, 0], Rx[:, 1] = x0, y0, x1, y1
    return Tx, Rx


class RadonPlan:
    """Device CSR of the ray/cell intersection lengths for one (nw, theta list)."""

    def __init__(self, nw, theta):
        _lib.require_cuda()
        lib = _lib.load()
        self.nw = int(nw)
        self.theta = np.ascontiguousarray(theta, dtype=np.float64)
        ends = [radon_endpoints(self.nw, t) for t in self.theta]
        self.tx = np.ascontiguousarray(np.stack([e[0] for e in ends]))
        self.rx = np.ascontiguousarray(np.stack([e[1] for e in ends]))
        handle = C.c_void_p()
        _lib.check(lib.cb_radon_plan_create(self.nw, self.theta.size, self.tx.ctypes.data, self.rx.ctypes.data,
                                            C.byref(handle)))
        self.handle = handle
        self._finalizer = weakref.finalize(self, lib.cb_radon_plan_destroy, handle)
        self.nnz = int(lib.cb_radon_plan_nnz(handle))

    def csr(self):
        """(data f64, indices i64, indptr i64 [ntheta, nw+1]) as traced on the device."""
        data = np.empty(self.nnz, dtype=np.float64)
        indices = np.empty(self.nnz, dtype=np.int64)
        indptr = np.empty((self.theta.size, self.nw + 1), dtype=np.int64)
        _lib.check(_lib.load().cb_radon_plan_get_csr(self.handle, data.ctypes.data, indices.ctypes.data,
                                                     indptr.ctypes.data))
        return data, indices, indptr

    def apply(self, grid_t, iy0_t, ix0_t):
        torch = _lib.require_cuda()
        n = int(iy0_t.numel())
        sino = torch.empty((n, self.nw, self.theta.size), dtype=torch.float64, device="cuda")
        ny, nx = grid_t.shape
        _lib.check(_lib.load().cb_radon_apply_batch(self.handle, grid_t.data_ptr(), ny, nx, grid_t.stride(0),
                                                    iy0_t.data_ptr(), ix0_t.data_ptr(), n, sino.data_ptr(),
                                                    _lib.stream_ptr(torch)))
        return sino


_radon_cache = {}


def radon_plan(nw, theta):
    torch = _lib.require_cuda()
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    key = (int(nw), theta.tobytes(), torch.cuda.current_device())
    plan = _radon_cache.get(key)
    if plan is None:
        if len(_radon_cache) > 8:
            _radon_cache.clear()
        plan = RadonPlan(nw, theta)
        _radon_cache[key] = plan
    return plan
