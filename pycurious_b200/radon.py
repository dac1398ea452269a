"""`radon2d(data, theta)` with the reference's signature (src/radon.pyx:12-120): sinogram of a
square array for angles theta in [0, pi).  The ray tracer (src/cradon.c) and the sparse product run
on the device (csrc/radon_body.cuh)."""
import numpy as np

from . import _lib, _plans


def radon2d(data, theta):
    theta = np.atleast_1d(np.asarray(theta, dtype=np.float64))
    if np.min(theta) < 0.0 or np.max(theta) >= np.pi:
        raise ValueError('theta should be within [0 pi)')
    data = np.asarray(data)
    nx, ny = data.shape
    if nx != ny:
        raise RuntimeError('data should be a square array')
    torch = _lib.require_cuda()
    plan = _plans.radon_plan(nx, theta)
    g = torch.from_numpy(np.ascontiguousarray(data, dtype=np.float64)).cuda()
    z = torch.zeros(1, dtype=torch.int32, device="cuda")
    return plan.apply(g, z, z)[0].cpu().numpy()


def radon_csr(nw, theta):
    """(data, indices, indptr[ntheta, nw+1]) of the ray/cell lengths traced on the device, with the
    integer types of the reference (int64)."""
    return _plans.radon_plan(nw, np.atleast_1d(theta)).csr()
