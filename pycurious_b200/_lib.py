"""ctypes binding of libcurie_b200.so (C-ABI declared in include/curie_b200.h).

There is no CPU fallback: if the shared library is missing, or no CUDA device is
present when a compute entry point is called, a RuntimeError is raised.
torch is used only to own device buffers / streams; the ABI sees raw pointers.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CB_LIB_PATH") or os.path.join(_HERE, "libcurie_b200.so")   # override: A/B builds

CB_MAX_RADIX_PASSES = 12
MAX_NK = 512            # radial bins per spectrum the fit / MCMC / misfit kernels hold in a warp slab (csrc/model.cu)


class CurieB200Error(RuntimeError):
    pass


class SpectrumPlanInfo(C.Structure):
    _fields_ = [("nw", C.c_int32), ("nbins", C.c_int32), ("nvh", C.c_int32), ("bluestein", C.c_int32),
                ("fft_len", C.c_int32), ("npass", C.c_int32), ("radix", C.c_int32 * CB_MAX_RADIX_PASSES),
                ("nnz_full", C.c_int32), ("nnz_folded", C.c_int32),
                ("teams", C.c_int32), ("threads", C.c_int32), ("grid_max", C.c_int32),
                ("smem_bytes", C.c_int64), ("workspace_bytes_per_cta", C.c_int64)]


class FitOptions(C.Structure):
    _fields_ = [("x0", C.c_double * 4), ("lb", C.c_double * 4), ("ub", C.c_double * 4),
                ("prior_mu", C.c_double * 4), ("prior_sigma", C.c_double * 4),
                ("max_iter", C.c_int32), ("flags", C.c_int32), ("gtol", C.c_double), ("ftol", C.c_double)]


class MhOptions(C.Structure):
    _fields_ = [("x0", C.c_double * 4), ("x_scale", C.c_double * 4),
                ("prior_mu", C.c_double * 4), ("prior_sigma", C.c_double * 4),
                ("nsim", C.c_int64), ("burnin", C.c_int64), ("seed", C.c_uint64), ("chain_offset", C.c_uint64),
                ("reduce_mode", C.c_int32), ("reserved", C.c_int32)]


_VP, _I, _I64, _D = C.c_void_p, C.c_int, C.c_int64, C.c_double

# name -> (restype, argtypes); every symbol include/curie_b200.h declares
SIGNATURES = {
    "cb_abi_version": (_I, []),
    "cb_last_error": (C.c_char_p, []),
    "cb_device_info": (_I, [C.POINTER(_I), C.POINTER(_I), C.POINTER(_I), C.POINTER(_I64)]),
    "cb_fp64_fma_peak": (_I, [C.POINTER(_D)]),
    "cb_host_register": (_I, [_VP, _I64]),
    "cb_host_unregister": (_I, [_VP]),
    "cb_release_scratch": (_I, []),
    "cb_fit_schedule": (_I, [_I64, C.POINTER(_I), C.POINTER(_I)]),
    "cb_spectrum_plan_create": (_I, [_I, _I, _VP, _VP, _VP, _VP, C.POINTER(_VP)]),
    "cb_spectrum_plan_destroy": (_I, [_VP]),
    "cb_spectrum_plan_get_info": (_I, [_VP, C.POINTER(SpectrumPlanInfo)]),
    "cb_spectrum_plan_get_bins": (_I, [_VP, _VP, _VP]),
    "cb_spectrum_workspace_bytes": (_I64, [_VP]),
    "cb_spectrum_batch": (_I, [_VP, _VP, _I64, _I64, _I64, _VP, _VP, _I64, _I, _D, _VP, _VP, _VP, _I64, _VP]),
    "cb_spectrum_rowspec_bytes": (_I64, [_VP, _I, _I, _I]),
    "cb_spectrum_batch_shared_rows": (_I, [_VP, _VP, _I64, _I64, _I64, _VP, _VP, _I64, _VP, _VP, _I, _VP, _I, _I, _I, _D,
                                           _VP, _VP, _VP, _I64, _VP, _I64, _VP]),
    "cb_detrend_batch": (_I, [_VP, _I64, _I64, _I64, _I, _VP, _VP, _I64, _VP, _VP]),
    "cb_besselk": (_I, [_VP, _VP, _I64, _VP, _VP]),
    "cb_bouligand_forward": (_I, [_VP, _I, _VP, _I64, _VP, _VP]),
    "cb_bouligand_points": (_I, [_VP, _VP, _I64, _VP, _VP]),
    "cb_misfit": (_I, [_VP, _I, _VP, _I64, _VP, _I64, _VP, _VP, _VP, _VP]),
    "cb_fit_batch": (_I, [_VP, _I, _VP, _I64, C.POINTER(FitOptions), _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "cb_fit_batch_rows": (_I, [_VP, _I, _VP, _I64, C.POINTER(FitOptions), _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "cb_perturb_spectra": (_I, [_VP, _VP, _I64, _I, _I64, C.c_uint64, C.c_uint64, _VP, _VP]),
    "cb_fft_plan_create": (_I, [_I, C.POINTER(_VP)]),
    "cb_fft_plan_destroy": (_I, [_VP]),
    "cb_fft2_filter_workspace_bytes": (_I64, [_I64, _I64]),
    "cb_fft2_filter": (_I, [_VP, _VP, _VP, _I64, _I64, _I64, _VP, _VP, _I64, _VP, _I64, _VP]),
    "cb_fft2_filter_windows_workspace_bytes": (_I64, [_I, _I64]),
    "cb_fft2_filter_windows": (_I, [_VP, _VP, _I64, _I64, _I64, _VP, _VP, _I64, _VP, _VP, _VP, _I64, _VP]),
    "cb_tanaka_batch": (_I, [_VP, _I, _VP, _VP, _I64, _D, _D, _D, _D, _VP, _VP]),
    "cb_mh_batch": (_I, [_VP, _I, _VP, _I64, C.POINTER(MhOptions), _VP, _VP, _VP, _VP]),
    "cb_radon_plan_create": (_I, [_I, _I, _VP, _VP, C.POINTER(_VP)]),
    "cb_radon_plan_destroy": (_I, [_VP]),
    "cb_radon_plan_nnz": (_I64, [_VP]),
    "cb_radon_plan_get_csr": (_I, [_VP, _VP, _VP, _VP]),
    "cb_radon_apply_batch": (_I, [_VP, _VP, _I64, _I64, _I64, _VP, _VP, _I64, _VP, _VP]),
    "cb_azimuthal_tail_batch": (_I, [_VP, _I, _I, _I64, _VP, _I, _D, _VP, _VP]),
}

_lib = None


def load():
    """Load the shared library (once) and declare every entry point.  Loading does not need a GPU."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CurieB200Error(
            "pycurious_b200: %s is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C pycurious_b200/csrc`. There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError here means the .so is stale
        fn.restype = res
        fn.argtypes = args
    if lib.cb_abi_version() != 1:
        raise CurieB200Error("libcurie_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().cb_last_error().decode("utf-8", "replace")
        raise CurieB200Error("curie_b200 error %d: %s" % (rc, msg))


def require_cuda():
    """torch + a CUDA device, or a loud failure."""
    import torch
    if not torch.cuda.is_available():
        raise CurieB200Error("pycurious_b200 needs a CUDA device (B200, sm_100a); none is visible and there is "
                             "no CPU fallback.")
    return torch


def ptr(t):
    """Raw address of a torch tensor / numpy array (None -> NULL)."""
    if t is None:
        return None
    if isinstance(t, np.ndarray):
        return t.ctypes.data
    return t.data_ptr()


def stream_ptr(torch):
    return torch.cuda.current_stream().cuda_stream


def fp64_fma_peak():
    """Measured FP64 FMA rate of the current device in TFLOP/s (roofline denominator)."""
    v = _D()
    check(load().cb_fp64_fma_peak(C.byref(v)))
    return v.value


def device_info():
    lib = load()
    a, b, c, d = _I(), _I(), _I(), _I64()
    check(lib.cb_device_info(C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
    return dict(sm_count=a.value, cc=(b.value, c.value), smem_optin=d.value)


def to_device_async(arr):
    """Host array -> device tensor without blocking the host behind the stream's earlier work: the bytes are staged
    in page-locked memory from torch's caching host allocator (which keeps the block until the copy has run) and
    copied asynchronously.  A plain `.cuda()` of a pageable array waits for everything queued before it -- e.g. the
    135 MB grid upload a batch call has just started -- and with it every bit of host-side planning that follows."""
    torch = require_cuda()
    t = torch.from_numpy(np.ascontiguousarray(arr))
    if t.numel() == 0:
        return t.cuda()
    return t.pin_memory().cuda(non_blocking=True)
