"""Noll-ordered, Noll-normalised Zernike basis on a square pixel grid.

The reference obtains its Zernike volume from the un-vendored third-party
package ``poppy`` (``poppy.zernike.zernike_basis(nterms, npix, outside=0.0)``,
called at ``Face-DeId/Camera/Utils.py:60-63`` and
``Image_Caption/Camera/Utils.py:75-77``; pinned ``poppy=1.0.3`` only in
``Image_Caption/environment.yml:154``).  poppy is not installable here (no
network) and its source is not under ``/root/reference``, so this module
restates the *published* definition that poppy implements:

* Noll index ``j`` (1-based) -> radial order ``n`` / azimuthal order ``m``
  (Noll 1976, "Zernike polynomials and atmospheric turbulence", Table 1 ordering:
  even ``j`` <-> cosine term, odd ``j`` <-> sine term),
* radial polynomial ``R_n^|m|(rho)``,
* Noll normalisation ``sqrt(n+1)`` (``m == 0``) or ``sqrt(2(n+1))`` (``m != 0``),
* sampling grid ``x = (arange(npix) - (npix-1)/2) / ((npix-1)/2)`` with
  ``meshgrid(x, x)`` ('xy'), unit-disk aperture ``rho <= 1`` and an ``outside``
  fill value.

PARITY UNPINNED against real poppy 1.0.3 (no golden vector exists in the
reference and poppy cannot be imported here).  This only feeds the height-map
projection ``h = sum_j coef_j Z_j`` *above* the CUDA hot path, whose input is
``h`` itself, so kernel parity does not depend on it.  The test-only reference
shim (``oracle/ref_shim.py``) plugs this same function in as ``poppy.zernike``
so that the reference module and ours see an identical basis.
"""
from __future__ import annotations

import math

import numpy as np

__all__ = ["noll_to_nm", "radial_polynomial", "zernike_basis", "zernike_volume"]


def noll_to_nm(j: int) -> tuple[int, int]:
    """Noll index (1-based) -> (n, m) with signed m (m<0: sine term)."""
    if j < 1:
        raise ValueError("Noll indices start at 1")
    n = 0
    rem = j - 1
    while rem > n:
        n += 1
        rem -= n
    # within order n the |m| values ascend in steps of 2, each |m|>0 twice
    abs_m = (n % 2) + 2 * ((rem + ((n + 1) % 2)) // 2)
    sign = 1 if j % 2 == 0 else -1
    return n, sign * abs_m


def radial_polynomial(n: int, m: int, rho: np.ndarray) -> np.ndarray:
    """R_n^m(rho) for m >= 0, (n - m) even."""
    m = abs(m)
    out = np.zeros_like(rho, dtype=np.float64)
    if (n - m) % 2:
        return out
    for k in range((n - m) // 2 + 1):
        c = ((-1.0) ** k * math.factorial(n - k)
             / (math.factorial(k) * math.factorial((n + m) // 2 - k) * math.factorial((n - m) // 2 - k)))
        out += c * rho ** (n - 2 * k)
    return out


def zernike_basis(nterms: int = 15, npix: int = 512, outside: float = float("nan")) -> np.ndarray:
    """(nterms, npix, npix) float64 stack of Noll Zernikes j = 1..nterms."""
    x = (np.arange(npix, dtype=np.float64) - (npix - 1) / 2.0) / ((npix - 1) / 2.0)
    xx, yy = np.meshgrid(x, x)
    rho = np.sqrt(xx ** 2 + yy ** 2)
    theta = np.arctan2(yy, xx)
    inside = rho <= 1.0
    out = np.zeros((nterms, npix, npix), dtype=np.float64)
    for idx in range(nterms):
        n, m = noll_to_nm(idx + 1)
        if m == 0:
            z = math.sqrt(n + 1) * radial_polynomial(n, 0, rho) if n else np.ones_like(rho)
        elif m > 0:
            z = math.sqrt(2.0 * (n + 1)) * radial_polynomial(n, m, rho) * np.cos(m * theta)
        else:
            z = math.sqrt(2.0 * (n + 1)) * radial_polynomial(n, m, rho) * np.sin(-m * theta)
        z = z * inside
        z[~inside] = outside
        out[idx] = z
    return out


def zernike_volume(resolution: int, n_terms: int, scale_factor: float = 1e-6) -> np.ndarray:
    """Basis scaled to metres, as the reference does (x1e-6): float64 (n_terms, res, res)."""
    return zernike_basis(nterms=n_terms, npix=resolution, outside=0.0) * scale_factor
