"""Synthetic fractal magnetic grids for benchmarks and tests.

The reference's generator (tests/Bouligand_forward.py) is an O(n^3) 3-D construction that cannot
reach 4096^2, so inputs are made by 2-D spectral synthesis: a white Gaussian field is shaped in the
Fourier domain by the Bouligand et al. (2009) radial spectrum for (beta, zt, dz) and transformed
back.  This is input generation, not part of the accelerated path; it is plain numpy/scipy so that
the GPU arm and the CPU reference arm of bench.py are fed the identical array.
"""
import numpy as np


def _log_spectrum(kr, beta, zt, dz):
    from scipy.special import gamma, kv
    x = kr * dz
    nu = 0.5 * (1.0 + beta)
    bracket = 0.5 * np.cosh(x) * gamma(nu) - kv(nu, x) * np.power(0.5 * x, nu)
    return -2.0 * kr * zt - (beta - 1.0) * np.log(kr) - x + np.log(np.sqrt(np.pi) / gamma(1.0 + 0.5 * beta) * bracket)


def fractal_grid(ny, nx, seed=20261019, beta=3.0, zt=1.0, dz=20.0, std=100.0, dx_km=1.0, workers=None, table=0):
    """(ny, nx) float64 anomaly grid with node spacing dx_km whose radial power spectrum follows
    bouligand2009(k, beta, zt, dz, C).  `workers` > 1 runs the two transforms on that many threads
    (scipy.fft; the EMAG2-shaped grid takes minutes on one) -- same field up to the rounding of the
    transform, so goldens are only ever generated with the default.  `table` > 0 evaluates the shaping spectrum
    on that many equispaced wavenumbers and interpolates (the Bessel function over 1e8 Fourier nodes is most of
    the generation time of the large grid); an input-generation shortcut, never used for a grid with goldens."""
    rng = np.random.default_rng(seed)
    if workers and workers > 1:
        import scipy.fft as sfft
        rfft2 = lambda a: sfft.rfft2(a, workers=workers)
        irfft2 = lambda a, s: sfft.irfft2(a, s=s, workers=workers)
    else:
        rfft2, irfft2 = np.fft.rfft2, lambda a, s: np.fft.irfft2(a, s=s)
    F = rfft2(rng.standard_normal((ny, nx)))
    ky = np.fft.fftfreq(ny, dx_km)[:, None]
    kx = np.fft.rfftfreq(nx, dx_km)[None, :]
    kr = 2.0 * np.pi * np.hypot(ky, kx)
    kr[0, 0] = 1.0
    with np.errstate(all="ignore"):
        if table:
            kt = np.linspace(kr[0, 1] * 0.5, kr.max() * 1.0001, int(table))
            amp = np.interp(kr, kt, np.exp(0.5 * _log_spectrum(kt, beta, zt, dz)))
        else:
            amp = np.exp(0.5 * _log_spectrum(kr, beta, zt, dz))
    amp[0, 0] = 0.0
    amp[~np.isfinite(amp)] = 0.0
    F *= amp
    del amp, kr
    out = irfft2(F, (ny, nx))
    return np.ascontiguousarray(out * (std / out.std()), dtype=np.float64)


def grid_extent(ny, nx, dx=1000.0):
    """(xmin, xmax, ymin, ymax) in metres for an (ny, nx) grid with spacing dx."""
    return 0.0, (nx - 1) * dx, 0.0, (ny - 1) * dx
