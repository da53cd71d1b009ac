"""GPU (-m gpu): the Image_Caption camera (``b200cam.lens.OpticsZernike``: CUDA convolution kernels behind the
reference's nn.Module interface) against the CPU oracle ``oracle/lens_oracle.py`` and the reference golden vectors.
Tolerances: rel-L2 <= 1e-4 on sensor images and the PSF, <= 1e-3 on gradients (BASELINE.json)."""
import numpy as np
import pytest
import torch

import b200cam.zernike as zern
from b200cam.lens import OpticsZernike
from conftest import GOLDEN_DIR, rel_l2
from oracle import lens_oracle as lo
from test_lens_oracle import CASES, inputs

pytestmark = pytest.mark.gpu


def make(wave, patch, terms, coeffs):
    dev = torch.device("cuda", 0)
    cam = OpticsZernike(input_shape=[1, patch, patch, 3], device=dev, wave_resolution=(wave, wave), patch_size=patch,
                        sample_interval=3e-6, zernike_terms=terms, height_tolerance=None).to(dev)
    with torch.no_grad():
        cam.zernike_coeffs_train.copy_(coeffs[3])
        cam.zernike_coeffs_no_train2.copy_(coeffs[4:])
    return cam


@pytest.mark.parametrize("case", list(CASES))
def test_matches_reference_golden_vectors(case):
    c = CASES[case]
    gold = np.load(GOLDEN_DIR / f"{case}.npz")
    img, w, coeffs = inputs(case)
    cam = make(c["wave"], c["patch"], c["terms"], coeffs)
    sensor, psf, zc, loss = cam(img.cuda())
    (sensor * w.cuda()).sum().backward()
    assert loss is None and zc.shape == (c["terms"], 1, 1)
    assert sensor.shape == img.shape and psf.shape == (1, c["patch"], c["patch"], 3)
    assert rel_l2(sensor, torch.from_numpy(gold["sensor"])) <= 1e-4
    assert rel_l2(psf, torch.from_numpy(gold["psf"])) <= 1e-4
    gd = float(gold["grad_defocus"].reshape(-1)[0])
    assert abs(float(cam.zernike_coeffs_train.grad) - gd) <= 1e-3 * abs(gd)
    assert float(sensor.max()) == 1.0                                          # Lens.py:312


def test_shipped_geometry_patch256():
    """wave_resolution 896, patch 256 (train.py:64-66) -> 1344^2 propagation, 512^2 convolution; 12 Zernike terms
    instead of 350 to keep the test small; `prueba="3"` exercises both masks and the energy loss."""
    wave, patch, terms, B = 896, 256, 12, 3
    g = torch.Generator().manual_seed(5)
    img = torch.rand(B, 3, patch, patch, generator=g)
    img[1, 2, 100, 140] += 3.0
    w = torch.rand(B, 3, patch, patch, generator=g)
    coeffs = torch.zeros(terms, 1, 1)
    coeffs[3], coeffs[4], coeffs[7] = -22.0, 0.5, -0.3
    cam = make(wave, patch, terms, coeffs)
    x = img.cuda().requires_grad_(True)
    sensor, psf, _, loss = cam(x, prueba="3")
    ((sensor * w.cuda()).sum() + loss).backward()

    cfg = lo.LensConfig(wave_res=wave, patch=patch, sample_interval=3e-6)
    vol = torch.tensor(zern.zernike_volume(wave, terms, 1e-6).astype(np.float32))
    cz = coeffs.clone().requires_grad_(True)
    xo = img.clone().requires_grad_(True)
    out = lo.lens_forward(xo, cz, vol, cfg, prueba="3", mask_1=cam.mask_1.cpu(), mask_2=cam.mask_2.cpu())
    ((out["sensor"] * w).sum() + out["loss"]).backward()
    assert rel_l2(sensor, out["sensor"]) <= 1e-4
    assert rel_l2(psf, out["psf"]) <= 1e-4
    assert abs(float(loss) - float(out["loss"])) <= 1e-4 * abs(float(out["loss"]))
    assert abs(float(cam.zernike_coeffs_train.grad) - float(cz.grad[3])) <= 1e-3 * abs(float(cz.grad[3]))
    assert rel_l2(x.grad, xo.grad) <= 1e-3


def test_state_dict_and_signature_match_reference():
    cam = make(128, 64, 10, torch.zeros(10, 1, 1))
    sd = cam.state_dict()
    assert {k: tuple(v.shape) for k, v in sd.items()} == {
        "zernike_coeffs_no_train": (3, 1, 1), "zernike_coeffs_no_train2": (6, 1, 1), "zernike_coeffs_train": (1, 1)}
    assert float(sd["zernike_coeffs_train"]) == 0.0
    fresh = OpticsZernike(input_shape=[1, 64, 64, 3], device=torch.device("cuda", 0), wave_resolution=(128, 128),
                          patch_size=64, zernike_terms=10)
    assert float(fresh.zernike_coeffs_train) == -22.0 and fresh.height_tolerance == 20e-9
    assert [p for p, v in fresh.named_parameters() if v.requires_grad] == ["zernike_coeffs_train"]
    with pytest.raises(NotImplementedError):
        fresh(torch.rand(1, 3, 64, 64).cuda(), psf_lab=True)


@pytest.mark.gpu
@pytest.mark.parametrize("P,B", [(16, 2), (64, 3), (256, 2)])
def test_crop_abs_resize_matches_torch_expression(P, B):
    """The fused epilogue vs the reference's own ops (abs, [129:-128]-style crop, nearest resize; Utils.py:289-295),
    forward and backward.  Bit-exact: no arithmetic besides |.| and sums of at most four gradient terms."""
    from b200cam import functional as F
    from b200cam.lens import CropAbsResize
    dev = torch.device("cuda", 0)
    n = 2 * P
    pad = (n - P) / 2
    pt, pb = int(np.ceil(pad)), int(np.floor(pad))
    g = torch.Generator(device="cpu").manual_seed(9)
    conv = torch.randn(B, 3, n, n, generator=g).to(dev)
    conv[0, 0, pt + 1, pt + 1] = 0.0                      # d|v|/dv at 0
    w = torch.randn(B, 3, P, P, generator=g).to(dev)
    a = conv.clone().requires_grad_(True)
    ref = torch.abs(a)[:, :, pt + 1:n - pb, pt + 1:n - pb]
    idx = torch.clamp(torch.arange(P, device=dev) - 1, min=0)
    ref = ref.index_select(2, idx).index_select(3, idx)
    (ref * w).sum().backward()
    b = conv.clone().requires_grad_(True)
    plan = F.DevicePlan(256, dev, tables=False)
    out = CropAbsResize.apply(b, P, pt + 1, plan)
    (out * w).sum().backward()
    assert torch.equal(out, ref)
    assert torch.allclose(b.grad, a.grad, rtol=0, atol=1e-6)


@pytest.mark.parametrize("wave,patch", [(128, 64), (896, 256), (160, 64)])
@pytest.mark.parametrize("prueba", [None, "3"])
def test_psf_kernels_match_oracle_with_tolerance_noise(wave, patch, prueba):
    """csrc/lens_psf.cu (phase plate + pruned mixed-radix Fresnel propagation + down-sampling + normalisation, masks and
    loss) against oracle/lens_oracle.py::psf_from_height_map, forward and dL/dcoeffs, with the height-tolerance noise live:
    the module draws it with the reference's torch.rand call (Utils.py:401-404), the oracle is given the same field.
    wave 896 -> 1344 = 2^6*3*7, 128 -> 192 = 2^6*3, 160 -> 240 = 2^4*3*5."""
    terms = 10
    if prueba is not None and patch != 256:
        pytest.skip("the reference's disc masks are 256 x 256 (Lens.py:113-129)")
    dev = torch.device("cuda", 0)
    coeffs = torch.zeros(terms, 1, 1)
    coeffs[3], coeffs[4], coeffs[6] = -22.0 * (wave / 896) ** 2, 0.4, -0.2
    cam = OpticsZernike(input_shape=[1, patch, patch, 3], device=dev, wave_resolution=(wave, wave), patch_size=patch,
                        sample_interval=3e-6, zernike_terms=terms, height_tolerance=20e-9).to(dev)
    with torch.no_grad():
        cam.zernike_coeffs_train.copy_(coeffs[3])
        cam.zernike_coeffs_no_train2.copy_(coeffs[4:])
    assert cam._constants(dev)["kernels"]
    flags = 3 if prueba == "3" else 0
    torch.manual_seed(11)
    height = cam.get_Heith_Map().unsqueeze(-1)
    psf, loss = cam._psf_kernels(height, flags)
    g = torch.Generator().manual_seed(3)
    wpsf = torch.rand(1, patch, patch, 3, generator=g, dtype=torch.float64)
    obj = (psf.double() * wpsf.cuda()).sum() * 1e3 + (loss if loss is not None else 0.0)
    obj.backward()
    grad = cam.zernike_coeffs_train.grad.clone()

    torch.manual_seed(11)                                           # the same noise field as the module drew
    noise = ((-20e-9 - 20e-9) * torch.rand([1, wave, wave, 1], dtype=torch.float32, device=dev) + 20e-9).cpu()
    cfg = lo.LensConfig(wave_res=wave, patch=patch, sample_interval=3e-6)
    vol = torch.tensor(zern.zernike_volume(wave, terms, 1e-6).astype(np.float32))
    cz = coeffs.clone().requires_grad_(True)
    hm = torch.sum(cz * vol, dim=0)[None, :, :, None]
    ref = lo.psf_from_height_map(hm, cfg, noise)
    ref_loss = None
    if prueba == "3":
        ref_loss = torch.norm((ref * cam.mask_1.cpu()) - ref)
        ref = ref * cam.mask_2.cpu()
    ((ref.double() * wpsf).sum() * 1e3 + (ref_loss if ref_loss is not None else 0.0)).backward()
    assert psf.dtype == ref.dtype
    assert rel_l2(psf, ref) <= 1e-4
    if ref_loss is not None:
        assert abs(float(loss) - float(ref_loss)) <= 1e-4 * abs(float(ref_loss))
    assert abs(float(grad) - float(cz.grad[3])) <= 1e-3 * abs(float(cz.grad[3]))


@pytest.mark.parametrize("P,B", [(64, 3), (128, 2), (256, 5), (512, 2), (256, 40)])
def test_fused_padded_sensor_matches_oracle(P, B):
    """csrc/lens_conv.cu (zero padding, |.|, crop, nearest resize and batch-global max inside the transform kernels) against
    oracle.lens_oracle.sensor_image + the global max (Utils.py:251-297, Lens.py:312): sensor, dL/dpsf and dL/dimg.  One image
    carries a spike placed so that the maximum lands in crop row 0 - the row the nearest resize duplicates, i.e. an exact tie."""
    from b200cam.lens import LensSensor
    from b200cam import functional as F
    import torch.nn.functional as TF
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(21 + P)
    img = torch.rand(B, 3, P, P, generator=g)
    psf = torch.rand(1, P, P, 3, generator=g) ** 8
    psf = psf / psf.sum(dim=[1, 2], keepdim=True)
    psf[0, P // 2, P // 2, 1] += 0.5                         # a strong centre tap: the spike below maps (almost) onto itself
    img[1, 1, 0, 7] += 40.0                                  # the tap sits one past the centre: conv row pt+1 = crop row 0 -> output rows 0 and 1
    w = torch.rand(B, 3, P, P, generator=g)

    xo, po = img.clone().requires_grad_(True), psf.clone().requires_grad_(True)
    raw = lo.sensor_image(xo, po)
    ref = raw / raw.max()
    (ref * w).sum().backward()

    n = 2 * P
    xg, pg = img.to(dev).requires_grad_(True), psf.to(dev).requires_grad_(True)
    k = TF.pad(pg[0].permute(2, 0, 1), [P // 2 + 1, P // 2 - 1, P // 2 + 1, P // 2 - 1])
    out = LensSensor.apply(xg, k, F.DevicePlan(n, dev, tables=False), None)
    (out * w.to(dev)).sum().backward()
    assert float(out.max()) == 1.0
    assert int((out == 1.0).sum()) >= 2                      # the duplicated row: an exact tie
    assert rel_l2(out, ref) <= 1e-4
    assert rel_l2(pg.grad, po.grad) <= 1e-3
    assert rel_l2(xg.grad, xo.grad) <= 1e-3


def test_constructor_default_geometry_runs():
    """the constructor defaults (wave 736 -> 1104 = 2^4*3*23 propagation, patch 368 -> 736^2 convolution) are not
    powers of two: the PSF runs on the generic-radix kernels, the sensor convolution on the reference's torch expression on the
    GPU; both against the oracle."""
    wave, patch, terms, B = 736, 368, 8, 2
    g = torch.Generator().manual_seed(8)
    img = torch.rand(B, 3, patch, patch, generator=g)
    img[0, 1, 200, 120] += 3.0
    w = torch.rand(B, 3, patch, patch, generator=g)
    coeffs = torch.zeros(terms, 1, 1)
    coeffs[3], coeffs[5] = -15.0, 0.3
    dev = torch.device("cuda", 0)
    cam = OpticsZernike(input_shape=[1, patch, patch, 3], device=dev, zernike_terms=terms, sample_interval=2e-6,
                        height_tolerance=None).to(dev)                  # wave_resolution / patch_size: the defaults
    with torch.no_grad():
        cam.zernike_coeffs_train.copy_(coeffs[3])
        cam.zernike_coeffs_no_train2.copy_(coeffs[4:])
    assert cam._constants(dev)["kernels"]
    sensor, psf, _, _ = cam(img.cuda())
    (sensor * w.cuda()).sum().backward()
    cfg = lo.LensConfig(wave_res=wave, patch=patch, sample_interval=2e-6)
    vol = torch.tensor(zern.zernike_volume(wave, terms, 1e-6).astype(np.float32))
    cz = coeffs.clone().requires_grad_(True)
    out = lo.lens_forward(img, cz, vol, cfg)
    (out["sensor"] * w).sum().backward()
    assert rel_l2(psf, out["psf"]) <= 1e-4
    assert rel_l2(sensor, out["sensor"]) <= 1e-4
    assert abs(float(cam.zernike_coeffs_train.grad) - float(cz.grad[3])) <= 1e-3 * abs(float(cz.grad[3]))


def test_caption_camera_step_is_graph_capturable():
    """Forward + backward of the caption camera captured in a CUDA graph (no host sync, no allocation outside the graph pool
    in any b200cam_lens_* call) replays to the eager result."""
    wave, patch, terms, B = 128, 64, 10, 3
    dev = torch.device("cuda", 0)
    cam = OpticsZernike(input_shape=[1, patch, patch, 3], device=dev, wave_resolution=(wave, wave), patch_size=patch,
                        sample_interval=3e-6, zernike_terms=terms, height_tolerance=None).to(dev)
    with torch.no_grad():
        cam.zernike_coeffs_train.fill_(-0.45)
    g = torch.Generator().manual_seed(12)
    img = torch.rand(B, 3, patch, patch, generator=g).to(dev)
    img[1, 0, 30, 21] += 4.0
    w = torch.rand(B, 3, patch, patch, generator=g).to(dev)
    out = {}

    def step():
        cam.zernike_coeffs_train.grad = None
        sensor, psf, _, _ = cam(img)
        torch.autograd.backward([sensor], [w])
        out["sensor"] = sensor

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step()
        eager_y, eager_g = out["sensor"].detach().clone(), cam.zernike_coeffs_train.grad.clone()
        step()
    torch.cuda.current_stream().wait_stream(side)
    out.clear()
    cam.zernike_coeffs_train.grad = None
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        step()
    y_static = out["sensor"]
    for _ in range(2):
        cam.zernike_coeffs_train.grad.zero_()
        graph.replay()
        torch.cuda.synchronize()
        assert rel_l2(y_static, eager_y) <= 1e-6
        assert abs(float(cam.zernike_coeffs_train.grad) - float(eager_g)) <= 1e-5 * abs(float(eager_g))
