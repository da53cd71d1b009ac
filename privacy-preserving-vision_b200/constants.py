"""Geometry / dispersion constants and device tables of the Face-DeId camera.

Everything the reference recomputes on every ``get_psf`` call but that does not depend on
the height map is folded into two complex tables, built ONCE at construction with the
reference's fp32 operation order (SURVEY trap T3: the phases reach ~4.7e3 rad, so the fp32
rounding of the *argument* must match; the kernels then never evaluate a large-argument
sin/cos):

* ``A[l]``  = aperture * exp(-i k/(2 f_l) r^2) * exp(+i k/(2 z) r^2) * exp(+i pre_l r^2)
              (``Face-DeId/Camera/Optics.py:95-100`` without the height-map factor)
* ``H[m]``  = exp(-i pi lam_m zi L_len/L_sen |f|^2)   (``Optics.py:103``), stored transposed.

The attribute names mirror the reference ``Camera.__init__`` (``Optics.py:13-55``) because
downstream code may read them (``cam.XY``, ``cam.rho``, ``cam.lamb`` ...).
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import torch

WAVELENGTHS_NM = (640, 550, 440)          # Optics.py:31  (R, G, B)
SENSOR_PITCH = 3.713103e-6                # Optics.py:28
OBJECT_DISTANCE = 0.75                    # Optics.py:36


def glass_minus_air_index(lam_um: torch.Tensor) -> torch.Tensor:
    """|n_glass - n_air| from two Sellmeier-type fits; wavelength in micrometres (Utils.py:33-40)."""
    sq = lam_um ** 2
    glass = torch.sqrt(1 + (0.6961663 * sq / (sq - 0.0684043 ** 2) + 0.4079426 * sq / (sq - 0.1162414 ** 2)
                            + 0.8974794 * sq / (sq - 9.896161 ** 2)))
    isq = lam_um ** -2
    air = 1 + 0.05792105 / (238.0185 - isq) + 0.00167917 / (57.362 - isq)
    return torch.abs(glass - air)


def _unit_phasor(angle: torch.Tensor) -> torch.Tensor:
    return torch.complex(torch.cos(angle), torch.sin(angle))


def _sq_radius_grid(lo: float, hi: float, step: float):
    ax = torch.arange(lo, hi, step)
    g0, g1 = torch.meshgrid(ax, ax, indexing="ij")
    return ax, g0, g1, g0 * g0 + g1 * g1


def build(N: int = 256) -> SimpleNamespace:
    """All constants of ``Camera.__init__`` on the CPU in fp32, plus the kernel tables."""
    c = SimpleNamespace()
    c.N, c.c = N, N // 2
    c.zi, c.z0 = 50e-3, 5.0
    c.f = 1 / (1 / c.zi + 1 / c.z0)
    c.R = c.f * glass_minus_air_index(torch.tensor(550e-9 * 1e6))
    c.radii = 2.0e-3
    c.pi = torch.tensor([math.pi])
    c.L_len = 2 * c.radii * 2
    c.px = SENSOR_PITCH
    c.L_sen = c.px * N
    c.lamb = (torch.tensor(WAVELENGTHS_NM) * 1.e-9).unsqueeze(-1).unsqueeze(-1)
    c.flmb = c.R / glass_minus_air_index(c.lamb * 1e6)
    c.k = 2 * c.pi / c.lamb
    c.z = torch.tensor([OBJECT_DISTANCE])

    c.du = c.L_len / N
    c.u, c.X, c.Y, c.XY = _sq_radius_grid(-1 * c.L_len / 2, c.L_len / 2, c.du)
    c.r = torch.sqrt(c.X ** 2 + c.Y ** 2)
    c.thetha = torch.atan2(c.Y, c.X)
    c.rad = c.r <= c.radii

    fx = torch.arange(-1 / (2 * c.du), 1 / (2 * c.du), 1 / c.L_len)
    c.fx1 = torch.roll(fx, -(N // 2), 0)
    c.FX1, c.FY1 = torch.meshgrid(c.fx1, c.fx1, indexing="ij")
    c.FF = c.FX1 * c.FX1 + c.FY1 * c.FY1

    c.dx2 = c.L_sen / N
    c.x2, c.X2, c.Y2, c.XY2 = _sq_radius_grid(-1 * c.L_sen / 2, c.L_sen / 2, c.dx2)
    c.r2 = torch.sqrt(c.X2 ** 2 + c.Y2 ** 2)
    c.thetha2 = torch.atan2(c.Y2, c.X2)
    c.rho = (c.r2 > c.px * 32) * 1.
    for ax in (c.u, fx, c.x2):
        if ax.numel() != N:
            raise ValueError(f"grid construction produced {ax.numel()} samples for N={N}")

    # ---- kernel tables ------------------------------------------------------------------
    lens = _unit_phasor(-(c.k / (2 * c.flmb)) * c.XY)
    defocus = _unit_phasor((c.k / (2 * c.z[0])) * c.XY)
    pre = _unit_phasor((c.pi / (c.lamb * c.zi * c.L_len) * (c.L_len - c.L_sen)) * c.XY)
    c.table_A = (torch.mul(c.rad, torch.mul(lens, defocus)) * pre).to(torch.complex64).contiguous()
    H = _unit_phasor(-(c.pi * c.lamb * c.zi * c.L_len / c.L_sen) * c.FF).to(torch.complex64)
    c.table_Ht = H.transpose(-1, -2).contiguous()
    c.kappa = [float(v) for v in (c.k * c.flmb).flatten()]
    return c


TENSOR_ATTRS = ("pi", "lamb", "flmb", "k", "u", "X", "Y", "XY", "r", "thetha", "rad", "fx1", "FX1", "FY1",
                "FF", "x2", "X2", "Y2", "XY2", "r2", "thetha2", "rho")
