"""CPU: run the CUDA kernel bodies through the test-only emulator (tests/emu/emu.cpp) and compare
with the oracle.  This checks every index computation, layout and launch sequence of the generic
pipeline without a GPU; the `-m gpu` tests repeat the comparison on the real device."""
import pytest
import torch

import emu_harness as emu
import b200cam.constants as K
import b200cam.synthetic as synth
from conftest import GOLDEN_CASES, load_golden, rel_l2
from oracle import camera_oracle as co


@pytest.mark.parametrize("N", [64, 128, 256, 512, 1024])
def test_two_pass_fft_all_sizes(N):
    g = torch.Generator().manual_seed(N)
    x = torch.complex(torch.randn(N, generator=g), torch.randn(N, generator=g))
    assert rel_l2(torch.view_as_real(emu.fft(x, False)), torch.view_as_real(torch.fft.fft(x))) < 1e-6
    assert rel_l2(torch.view_as_real(emu.fft(x, True)), torch.view_as_real(torch.fft.ifft(x) * N)) < 1e-6


@pytest.mark.parametrize("name", GOLDEN_CASES[:2])
def test_emulated_pipeline_matches_golden(name):
    gold = load_golden(name)
    N = gold["N"]
    T = K.build(N)
    psf, field, stats = emu.psf_fwd(gold["h"][0], T)
    assert rel_l2(psf, gold["psf"][0]) <= 1e-5
    assert abs(stats[1].item() - gold["loss_rad"].item()) <= 1e-5 * gold["loss_rad"].item()
    assert abs(stats[2].item() - gold["centering_loss"].item()) <= 1e-4 * gold["centering_loss"].item()
    sensor, m, tc, tp, otf, spec = emu.sensor_fwd(gold["img"], psf, save_spectrum=True)
    assert rel_l2(sensor, gold["sensor"]) <= 1e-5
    assert tc.tolist() == [1] * gold["B"]
    gpsf, _ = emu.sensor_bwd(gold["w"], gold["img"], sensor, m, tc, tp, psf, otf, spectrum=spec)   # saved row spectra
    gpsf2, _ = emu.sensor_bwd(gold["w"], gold["img"], sensor, m, tc, tp, psf, otf)                # recomputed
    assert rel_l2(gpsf2, gpsf) <= 1e-6
    gh = emu.psf_bwd(gpsf, torch.tensor([1.0, 1.0]), gold["h"][0], T, psf, field, stats)
    assert rel_l2(gh, gold["grad_h"][0]) <= 1e-4


def test_emulated_img_grad_and_ragged_batch():
    """B=5 is not a multiple of the accumulate chunking; also checks the optional dL/dimg output."""
    N, B = 64, 5
    C = co.build_constants(N)
    img, w, h = synth.images(B, N, 5), synth.upstream_grad(B, N, 6), synth.height_map(N, 7)
    psf, _ = co.psf_from_height(h, C)
    sensor, m, tc, tp, otf = emu.sensor_fwd(img, psf[0])
    gpsf_o, gimg_o = co.sensor_backward(w, img, psf, N, want_img_grad=True)
    gpsf, gimg = emu.sensor_bwd(w, img, sensor, m, tc, tp, psf[0].contiguous(), otf, want_img_grad=True)
    assert rel_l2(gpsf, gpsf_o[0]) <= 1e-5
    assert rel_l2(gimg, gimg_o) <= 1e-5


def test_emulated_tie_splitting():
    """Two exactly equal maxima in different channels: gradient splits evenly (torch amax semantics)."""
    N = 64
    img = torch.zeros(1, 3, N, N)
    img[0, 0, 5, 7] = 1.0
    img[0, 2, 40, 9] = 1.0
    psf = torch.zeros(1, 3, N, N)
    psf[0, :, N // 2, N // 2] = 0.25
    psf[0, :, N // 2 + 1, N // 2] = 0.0625
    w = synth.upstream_grad(1, N, 11)
    sensor, m, tc, tp, otf = emu.sensor_fwd(img, psf[0].contiguous())
    # fp32 FFT rounding may or may not keep the tie exact: force the recorded tie set to the two
    # analytic maxima and give the oracle the same mask, so the tie-splitting arithmetic is what is tested
    tc[0] = 2
    tp[0, 0], tp[0, 1] = 0 * N * N + 5 * N + 7, 2 * N * N + 40 * N + 9
    mask = torch.zeros(1, 3, N, N)
    mask[0, 0, 5, 7] = 1.0
    mask[0, 2, 40, 9] = 1.0
    gpsf_o, gimg_o = co.sensor_backward(w.double(), img.double(), psf.double(), N, want_img_grad=True, tie_mask=mask)
    gpsf, gimg = emu.sensor_bwd(w, img, sensor, m, tc, tp, psf[0].contiguous(), otf, want_img_grad=True)
    assert sorted(tp[0, :2].tolist()) == sorted([0 * N * N + 5 * N + 7, 2 * N * N + 40 * N + 9])
    assert rel_l2(gpsf, gpsf_o[0]) <= 1e-5
    assert rel_l2(gimg, gimg_o) <= 1e-5


@pytest.mark.parametrize("T,N", [(7, 16), (40, 64)])
def test_emulated_zernike_projection(T, N):
    """h = sum_j coef_j Z_j (Optics.py:79-83) and its adjoint, kernel bodies on the CPU emulator."""
    g = torch.Generator().manual_seed(3)
    Z = torch.randn(T, N, N, generator=g)
    c = torch.randn(T, generator=g)
    gh = torch.randn(N, N, generator=g)
    h, gc = emu.zernike(c, Z, gh)
    assert rel_l2(h, (c[:, None, None] * Z).sum(0)) <= 1e-6
    assert rel_l2(gc, (Z * gh).sum((1, 2))) <= 1e-6
