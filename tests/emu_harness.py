"""TEST INFRASTRUCTURE ONLY - ctypes wrapper around tests/emu/emu.cpp (CPU emulation of the kernel bodies)."""
import ctypes
import subprocess
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parent.parent
SRC = REPO / "tests" / "emu" / "emu.cpp"
OUT = REPO / "tests" / "_build" / "libb200cam_emu.so"
CSRC = REPO / "privacy-preserving-vision_b200" / "csrc"


def build() -> Path:
    deps = [SRC, *CSRC.glob("*.cuh")]
    if OUT.exists() and all(OUT.stat().st_mtime >= d.stat().st_mtime for d in deps):
        return OUT
    OUT.parent.mkdir(parents=True, exist_ok=True)
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", str(OUT), str(SRC)], check=True)
    return OUT


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(str(build()))
    return _lib


def _p(t):
    if t is None:
        return ctypes.c_void_p(0)
    assert t.is_contiguous()
    return ctypes.c_void_p(t.data_ptr())


def fft(x: torch.Tensor, inverse: bool) -> torch.Tensor:
    x = torch.view_as_real(x.to(torch.complex64).contiguous()).contiguous()
    out = torch.empty_like(x)
    rc = lib().emu_fft(ctypes.c_int(x.shape[0]), ctypes.c_int(int(inverse)), _p(x), _p(out))
    assert rc == 0
    return torch.view_as_complex(out)


def psf_fwd(h, tables):
    N = tables.N
    A = torch.view_as_real(tables.table_A).contiguous()
    Ht = torch.view_as_real(tables.table_Ht).contiguous()
    rho = tables.rho.float().contiguous()
    kappa = (ctypes.c_float * 3)(*tables.kappa)
    psf = torch.empty(3, N, N)
    field = torch.empty(3, N, N, 2)
    stats = torch.zeros(4)
    rc = lib().emu_psf_fwd(N, _p(h.contiguous()), _p(A), _p(Ht), _p(rho), kappa, _p(psf), _p(field), _p(stats))
    assert rc == 0
    return psf, field, stats


def psf_bwd(gpsf, gscal, h, tables, psf, field, stats):
    N = tables.N
    A = torch.view_as_real(tables.table_A).contiguous()
    Ht = torch.view_as_real(tables.table_Ht).contiguous()
    rho = tables.rho.float().contiguous()
    kappa = (ctypes.c_float * 3)(*tables.kappa)
    gh = torch.empty(N, N)
    g_rad = gscal[0:1] if gscal is not None else None
    g_cen = gscal[1:2] if gscal is not None else None
    rc = lib().emu_psf_bwd(N, _p(gpsf), _p(g_rad), _p(g_cen), _p(h.contiguous()), _p(A), _p(Ht), _p(rho), kappa, _p(psf),
                           _p(field), _p(stats), _p(gh))
    assert rc == 0
    return gh


def sensor_fwd(img, psf, save_spectrum=False):
    B, _, N, _ = img.shape
    sensor = torch.empty_like(img)
    img_max = torch.empty(B)
    tie_count = torch.zeros(B, dtype=torch.int32)
    tie_pos = torch.zeros(B, 8, dtype=torch.int32)
    otf = torch.empty(3, N // 2 + 1, N, 2)
    spectrum = torch.empty(3 * B, N // 2 + 1, N, 2) if save_spectrum else None
    rc = lib().emu_sensor_fwd(N, B, _p(img.contiguous()), _p(psf.contiguous()), _p(sensor), _p(img_max),
                              _p(tie_count), _p(tie_pos), _p(otf), _p(spectrum))
    assert rc == 0
    if save_spectrum:
        return sensor, img_max, tie_count, tie_pos, otf, spectrum
    return sensor, img_max, tie_count, tie_pos, otf


def sensor_bwd(g, img, sensor, img_max, tie_count, tie_pos, psf, otf, want_img_grad=False, spectrum=None):
    B, _, N, _ = img.shape
    gpsf = torch.empty(3, N, N)
    gimg = torch.empty_like(img) if want_img_grad else None
    rc = lib().emu_sensor_bwd(N, B, _p(g.contiguous()), _p(img.contiguous()), _p(sensor), _p(img_max), _p(tie_count),
                              _p(tie_pos), _p(psf.contiguous()), _p(otf), _p(spectrum), _p(gpsf), _p(gimg))
    assert rc == 0
    return gpsf, gimg


def zernike(coef, Z, gh):
    """(h, gcoef) through the emulated projection bodies."""
    T, NN = Z.shape[0], Z[0].numel()
    h = torch.empty(Z.shape[1:])
    gc = torch.empty(T)
    rc = lib().emu_zernike(T, NN, _p(coef.contiguous()), _p(Z.contiguous()), _p(h), _p(gh.contiguous()), _p(gc))
    assert rc == 0
    return h, gc
