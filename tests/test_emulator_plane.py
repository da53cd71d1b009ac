"""CPU: the plane-resident N=256 sensor kernel bodies (csrc/plane.cuh) through the test-only emulator
(tests/emu/emu_plane.cpp) against torch.fft - index arithmetic, layouts, the packed DC/Nyquist column, the
cross-cluster image max and the backward accumulation, checked before any GPU run."""
import ctypes
import subprocess
from pathlib import Path

import pytest
import torch

from conftest import rel_l2

REPO = Path(__file__).resolve().parent.parent
SRC = REPO / "tests" / "emu" / "emu_plane.cpp"
OUT = REPO / "tests" / "_build" / "libb200cam_emu_plane.so"
CSRC = REPO / "privacy-preserving-vision_b200" / "csrc"
N = 256


def _lib():
    deps = [SRC, *CSRC.glob("*.cuh")]
    if not OUT.exists() or any(OUT.stat().st_mtime < d.stat().st_mtime for d in deps):
        OUT.parent.mkdir(parents=True, exist_ok=True)
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", str(OUT), str(SRC)], check=True)
    return ctypes.CDLL(str(OUT))


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _inputs(B, seed):
    g = torch.Generator().manual_seed(seed)
    img = torch.rand(B, 3, N, N, generator=g)
    psf = torch.rand(3, N, N, generator=g) ** 4
    psf = psf / psf.sum()
    w = torch.rand(B, 3, N, N, generator=g)
    return img, psf, w


def _otf(psf):
    """library layout [3][129][256]: rfft2(roll(psf, -N/2)) / N^2, u-major (Face-DeId/Camera/Optics.py:126, Utils.py:9)"""
    K = torch.fft.rfft2(torch.roll(psf.double(), (-N // 2, -N // 2), (-2, -1))) / (N * N)      # [3][v][u]
    return torch.view_as_real(K.transpose(1, 2).contiguous().to(torch.complex64)).contiguous()


def _rows(lib, x, nctas=5):
    planes = x.shape[0] * 3
    A = torch.zeros(planes, 128, 128, 4)
    assert lib.emu_prow(planes, _p(x.contiguous()), _p(A), nctas) == 0
    return A


def test_row_kernel_layout():
    lib = _lib()
    img, _, _ = _inputs(2, 1)
    A = _rows(lib, img)
    S = 2 * torch.fft.rfft(img.double().reshape(6, N, N), dim=-1)                 # [plane][y][k]
    lo, hi = S[:, :128, :], S[:, 128:, :]                                         # rows P, P+128
    ref = torch.zeros(6, 128, 128, 4, dtype=torch.float64)
    ref[:, 1:, :, 0] = lo[:, :, 1:128].real.transpose(1, 2)
    ref[:, 1:, :, 1] = lo[:, :, 1:128].imag.transpose(1, 2)
    ref[:, 1:, :, 2] = hi[:, :, 1:128].real.transpose(1, 2)
    ref[:, 1:, :, 3] = hi[:, :, 1:128].imag.transpose(1, 2)
    ref[:, 0, :, 0], ref[:, 0, :, 1] = lo[:, :, 0].real, lo[:, :, 128].real
    ref[:, 0, :, 2], ref[:, 0, :, 3] = hi[:, :, 0].real, hi[:, :, 128].real
    assert rel_l2(A.double(), ref) < 1e-6


@pytest.mark.parametrize("B,G3", [(1, 1), (3, 2), (2, 4)])
def test_conv_kernel_matches_fft_convolution(B, G3):
    lib = _lib()
    img, psf, _ = _inputs(B, 10 + B)
    A = _rows(lib, img)
    y = torch.zeros(B, 3, N, N)
    m = torch.zeros(B)
    tc = torch.zeros(B, dtype=torch.int32)
    tp = torch.zeros(B, 8, dtype=torch.int32)
    otf = _otf(psf)
    assert lib.emu_pconv(B, G3, _p(A), _p(otf), _p(y), _p(m), _p(tc), _p(tp), 1, 1) == 0
    K = torch.fft.rfft2(torch.roll(psf.double(), (-N // 2, -N // 2), (-2, -1)))
    X = torch.fft.rfft2(img.double())
    conv = torch.fft.irfft2(X * K, s=(N, N))
    mx = conv.amax(dim=(1, 2, 3))
    assert rel_l2(y.double(), conv / mx[:, None, None, None]) < 1e-6
    assert rel_l2(m.double(), mx) < 1e-6
    assert tc.tolist() == [1] * B
    for b in range(B):
        assert tp[b, 0].item() == int(conv[b].flatten().argmax())
        assert y[b].max().item() == 1.0
    # X^ left in place of A: [plane][u][s] = (X[u][s], X[u][s+128]) with factor 2; column 0 = DC + i Nyquist packed
    Xf = 2 * X.reshape(3 * B, N, 129).transpose(1, 2)                              # [plane][u][v]
    Ah = torch.view_as_complex(A.reshape(3 * B, 128, 128, 2, 2).double().contiguous())   # [plane][u][s][half]
    got = torch.cat([Ah[..., 0], Ah[..., 1]], dim=2)                               # [plane][u][v]
    assert rel_l2(torch.view_as_real(got[:, 1:]), torch.view_as_real(Xf[:, 1:128])) < 1e-6
    assert rel_l2(torch.view_as_real(got[:, 0]), torch.view_as_real(Xf[:, 0] + 1j * Xf[:, 128])) < 1e-6


def test_plain_convolution_mode():
    lib = _lib()
    img, psf, _ = _inputs(2, 3)
    A = _rows(lib, img)
    y = torch.zeros(2, 3, N, N)
    otf = _otf(psf)
    assert lib.emu_pconv(2, 1, _p(A), _p(otf), _p(y), _p(None), _p(torch.zeros(2, dtype=torch.int32)), _p(None), 0, 0) == 0
    K = torch.fft.rfft2(torch.roll(psf.double(), (-N // 2, -N // 2), (-2, -1)))
    conv = torch.fft.irfft2(torch.fft.rfft2(img.double()) * K, s=(N, N))
    assert rel_l2(y.double(), conv) < 1e-6


@pytest.mark.parametrize("B,G3", [(1, 1), (5, 2)])
def test_accumulate_kernel(B, G3):
    lib = _lib()
    img, psf, w = _inputs(B, 20 + B)
    A = _rows(lib, img)
    y = torch.zeros(B, 3, N, N)
    m = torch.zeros(B)
    tc = torch.zeros(B, dtype=torch.int32)
    tp = torch.zeros(B, 8, dtype=torch.int32)
    otf = _otf(psf)
    assert lib.emu_pconv(B, G3, _p(A), _p(otf), _p(y), _p(m), _p(tc), _p(tp), 1, 1) == 0
    partial = torch.zeros(G3, 3, 129, N, 2)
    dotp = torch.zeros(3 * B, 64)
    assert lib.emu_pacc(B, G3, _p(w.contiguous()), _p(A), _p(otf), _p(m), _p(partial), _p(dotp)) == 0
    X = torch.fft.rfft2(img.double())                                               # [B][3][v][u]
    G = torch.fft.rfft2(w.double())
    ref = 4 * (torch.conj(X) * G / m.double()[:, None, None, None]).sum(0).transpose(1, 2)     # [3][u][v]
    got = torch.view_as_complex(partial.double().contiguous()).sum(0)
    assert rel_l2(torch.view_as_real(got), torch.view_as_real(ref)) < 1e-5
    K = torch.fft.rfft2(torch.roll(psf.double(), (-N // 2, -N // 2), (-2, -1)))
    conv = torch.fft.irfft2(X * K, s=(N, N))
    sdot = (w.double() * conv).sum(dim=(1, 2, 3))
    assert rel_l2(dotp.double().reshape(B, 192).sum(1) / 4, sdot) < 1e-5
