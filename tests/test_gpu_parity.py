"""GPU (-m gpu): parity of the CUDA path, called through the nn.Module / C ABI, against
(1) the committed golden vectors of the unmodified reference and (2) the CPU oracle on seeded inputs.

Tolerances are BASELINE.json's: rel-L2 <= 1e-4 on sensor images (and the PSF), <= 1e-3 on dL/dh.
"""
import ctypes

import pytest
import torch

import b200cam.synthetic as synth
from b200cam import _lib
from b200cam.optics import Camera
from conftest import GOLDEN_CASES, load_golden, rel_l2
from oracle import camera_oracle as co

pytestmark = pytest.mark.gpu
TOL_SENSOR, TOL_GRAD = 1e-4, 1e-3


def make_camera(N, h_cpu, terms=6):
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    cam = Camera(device=dev, N=N, zernike_terms=terms)
    h = h_cpu.to(dev).requires_grad_(True)
    cam.get_Heith_Map = lambda: h
    return cam, h


def oracle_step(img, w, h_cpu, N, g_rad=1.0, g_cen=1.0):
    C = co.build_constants(N)
    h = h_cpu.clone().requires_grad_(True)
    out = co.camera_forward(img, h, C)
    ((out["sensor"] * w).sum() + g_rad * out["loss_rad"] + g_cen * out["centering_loss"]).backward()
    return out, h.grad


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_matches_reference_golden_vectors(name):
    gold = load_golden(name)
    N = gold["N"]
    cam, h = make_camera(N, gold["h"])
    y = cam(gold["img"].cuda())
    ((y * gold["w"].cuda()).sum() + cam.loss_rad + cam.centering_loss).backward()
    assert rel_l2(y, gold["sensor"]) <= TOL_SENSOR
    assert rel_l2(cam.psfs, gold["psf"]) <= TOL_SENSOR
    assert rel_l2(h.grad, gold["grad_h"]) <= TOL_GRAD
    assert abs(cam.loss_rad.item() - gold["loss_rad"].item()) <= 1e-4 * gold["loss_rad"].item()
    assert abs(cam.centering_loss.item() - gold["centering_loss"].item()) <= 1e-3 * gold["centering_loss"].item()
    assert y.shape == gold["sensor"].shape and cam.psfs.shape == (1, 3, N, N)


def test_config1_forward_batch8():
    """BASELINE config 1: forward only, random height map, batch 8 of 256x256 RGB."""
    N, B = 256, 8
    h_cpu, img = synth.height_map(N, 99), synth.images(B, N, 100)
    cam, _ = make_camera(N, h_cpu)
    with torch.no_grad():
        y = cam(img.cuda())
    C = co.build_constants(N)
    out = co.camera_forward(img, h_cpu, C)
    assert rel_l2(y, out["sensor"]) <= TOL_SENSOR
    assert torch.allclose(y.amax((1, 2, 3)).cpu(), torch.ones(B))       # Optics.py:128 property


def test_config2_forward_backward_batch64():
    """BASELINE config 2: forward+backward into the height map, batch 64 of 256x256 RGB."""
    N, B = 256, 64
    h_cpu, img, w = synth.height_map(N), synth.images(B, N), synth.upstream_grad(B, N)
    out, gh = oracle_step(img, w, h_cpu, N)
    assert synth.top2_relative_gap(out["conv"]).min() >= 1e-5           # amax margin (trap T2)
    cam, h = make_camera(N, h_cpu)
    y = cam(img.cuda())
    ((y * w.cuda()).sum() + cam.loss_rad + cam.centering_loss).backward()
    assert rel_l2(y, out["sensor"]) <= TOL_SENSOR
    assert rel_l2(h.grad, gh) <= TOL_GRAD


@pytest.mark.parametrize("N,B", [(64, 3), (128, 5), (512, 2), (1024, 1), (256, 1), (256, 37), (512, 9), (64, 150), (1024, 3)])
def test_other_resolutions(N, B):
    """Other sizes and ragged batches: fewer tiles than persistent CTAs (B = 1), chunk splits that do not divide (B = 37)."""
    h_cpu, img, w = synth.height_map(N, 5), synth.images(B, N, 6), synth.upstream_grad(B, N, 7)
    out, gh = oracle_step(img, w, h_cpu, N, 0.5, 2.0)
    cam, h = make_camera(N, h_cpu)
    y = cam(img.cuda())
    ((y * w.cuda()).sum() + 0.5 * cam.loss_rad + 2.0 * cam.centering_loss).backward()
    assert rel_l2(y, out["sensor"]) <= TOL_SENSOR
    assert rel_l2(cam.psfs, out["psf"]) <= TOL_SENSOR
    assert rel_l2(h.grad, gh) <= TOL_GRAD


def test_zernike_parameter_gradient():
    """End to end through the module's own parameters (Zer_train), default-init lens, T=40."""
    N, B = 256, 4
    dev = torch.device("cuda", 0)
    torch.manual_seed(1)
    cam = Camera(device=dev, N=N, zernike_terms=40)
    img, w = synth.images(B, N, 21), synth.upstream_grad(B, N, 22)
    y = cam(img.cuda())
    ((y * w.cuda()).sum() + cam.loss_rad + cam.centering_loss).backward()
    C = co.build_constants(N)
    zt = cam.Zer_train.detach().cpu().clone().requires_grad_(True)
    h = co.height_map(cam.Zer_no_train.detach().cpu(), zt, cam.zernike_volume.cpu())
    out = co.camera_forward(img, h, C)
    ((out["sensor"] * w).sum() + out["loss_rad"] + out["centering_loss"]).backward()
    assert rel_l2(y, out["sensor"]) <= TOL_SENSOR
    assert rel_l2(cam.Zer_train.grad, zt.grad) <= TOL_GRAD
    assert cam.Zer_no_train.grad is None


def test_image_gradient_optional_output():
    N, B = 128, 3
    h_cpu, w = synth.height_map(N, 5), synth.upstream_grad(B, N, 7)
    img = synth.images(B, N, 6)
    C = co.build_constants(N)
    xo = img.clone().requires_grad_(True)
    out = co.camera_forward(xo, h_cpu, C)
    (out["sensor"] * w).sum().backward()
    cam, _ = make_camera(N, h_cpu)
    xg = img.cuda().requires_grad_(True)
    (cam(xg) * w.cuda()).sum().backward()
    assert rel_l2(xg.grad, xo.grad) <= TOL_GRAD


def test_exact_ties_split_like_torch():
    """Channels 0 and 1 carry identical data and identical PSFs, so their convolutions are bit-identical
    and the per-image maximum is attained exactly twice: the gradient must split evenly (torch amax)."""
    N, B = 128, 2
    img = synth.images(B, N, 31)
    img[:, 1] = img[:, 0]
    img[:, 2] *= 0.25
    C = co.build_constants(N)
    psf, _ = co.psf_from_height(synth.height_map(N, 32), C)
    psf = psf.clone()
    psf[0, 1] = psf[0, 0]
    w = synth.upstream_grad(B, N, 33)
    from b200cam import functional as F
    plan = F.DevicePlan(N, torch.device("cuda", 0))
    p = psf.cuda().requires_grad_(True)
    x = img.cuda().requires_grad_(True)
    y = F.sensor_conv(x, p, plan)
    (y * w.cuda()).sum().backward()
    assert [(y[b] == 1.0).sum().item() for b in range(B)] == [2, 2]
    mask = (y.detach().cpu() == 1.0).float()
    gpsf, gimg = co.sensor_backward(w.double(), img.double(), psf.double(), N, want_img_grad=True, tie_mask=mask)
    assert rel_l2(p.grad, gpsf) <= 1e-5
    assert rel_l2(x.grad, gimg) <= 1e-5


def test_size_independent_properties_full_batch():
    """At the full config-2 size: PSF sums to one, every image peaks at exactly 1, an impulse image
    reproduces the (shifted) PSF, circular shifts commute with the camera, runs are bit-reproducible."""
    N, B = 256, 64
    cam, _ = make_camera(N, synth.height_map(N, 3))
    img = synth.images(B, N, 4).cuda()
    with torch.no_grad():
        y1 = cam(img)
        psf = cam.psfs.clone()
        y2 = cam(img)
        ys = cam(torch.roll(img, (17, -40), (-2, -1)))
        delta = torch.zeros(1, 3, N, N, device="cuda")
        delta[0, :, 10, 20] = 1.0
        yd = cam(delta)
    assert torch.equal(y1, y2)
    assert abs(psf.sum().item() - 1.0) <= 1e-5
    assert torch.equal(y1.amax((1, 2, 3)), torch.ones(B, device="cuda"))
    assert rel_l2(ys, torch.roll(y1, (17, -40), (-2, -1))) <= 1e-5
    expect = torch.roll(psf, (10 - N // 2, 20 - N // 2), (-2, -1))
    assert rel_l2(yd, expect / expect.max()) <= 1e-4


def test_backward_is_deterministic():
    N, B = 256, 16
    h_cpu, img, w = synth.height_map(N), synth.images(B, N).cuda(), synth.upstream_grad(B, N).cuda()
    grads = []
    for _ in range(2):
        cam, h = make_camera(N, h_cpu)
        ((cam(img) * w).sum() + cam.loss_rad).backward()
        grads.append(h.grad.clone())
    assert torch.equal(grads[0], grads[1])


def test_cuda_graph_capture_of_forward_backward():
    N, B = 256, 8
    h_cpu, img, w = synth.height_map(N), synth.images(B, N).cuda(), synth.upstream_grad(B, N).cuda()
    cam, h = make_camera(N, h_cpu)
    one = torch.ones((), device="cuda")

    def step():
        h.grad = None
        torch.autograd.backward([cam(img), cam.loss_rad, cam.centering_loss], [w, one, one])

    def drop_refs():      # the module (like the reference) keeps psfs/losses alive; they pin the old autograd graph
        cam.psfs = cam.loss_rad = cam.centering_loss = cam._pending_centering = None
        h.grad = None

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step()
        eager = h.grad.clone()
    torch.cuda.current_stream().wait_stream(side)
    drop_refs()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        step()
    h.grad.zero_()
    graph.replay()
    torch.cuda.synchronize()
    first = h.grad.clone()
    # eager launches use the cooperative PSF kernel, captured ones the multi-kernel chain: same maths,
    # different partial-sum grouping, so compare to rounding; replays themselves are bit-reproducible
    assert rel_l2(first, eager) <= 1e-5
    h.grad.zero_()
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(h.grad, first)


def test_abi_error_codes_on_device():
    lib = _lib.load_library()
    N, B = 64, 1
    _lib.ensure_init(N, 0)
    t = torch.zeros(B, 3, N, N, device="cuda")
    small = torch.zeros(16, dtype=torch.uint8, device="cuda")
    i32 = torch.zeros(8, dtype=torch.int32, device="cuda")
    otf = torch.zeros(lib.b200cam_otf_bytes(N) // 4, device="cuda")
    p = _lib.ptr
    rc = lib.b200cam_sensor_fwd(p(t), p(t), p(t), p(t), p(i32), p(i32), p(otf), p(None), p(small), small.numel(), B, N,
                                ctypes.c_void_p(0))
    assert rc == -3                                                   # B200CAM_E_WORKSPACE
    with pytest.raises(TypeError):
        cam, _ = make_camera(N, synth.height_map(N))
        cam(t.double())
    with pytest.raises(ValueError):
        cam(torch.zeros(1, 3, 32, 32, device="cuda"))


def test_empty_batch():
    N = 64
    cam, _ = make_camera(N, synth.height_map(N))
    y = cam(torch.zeros(0, 3, N, N, device="cuda"))
    assert y.shape == (0, 3, N, N)


# ---------------------------------------------------------------------------------------------------------
# N=256 has two sensor paths: the generic row/column/row kernels (default, all sizes) and the cluster "plane" kernels of
# plane.cuh (opt-in, B200CAM_PLANE=1, read once per process -> subprocess)
# ---------------------------------------------------------------------------------------------------------
_FORCED_FUSED_SCRIPT = r"""
import sys, torch
sys.path.insert(0, {repo!r}); sys.path.insert(0, {repo!r} + "/tests")
import b200cam.synthetic as synth
from b200cam.optics import Camera
from oracle import camera_oracle as co
N, B = 256, int(sys.argv[1])
dev = torch.device("cuda", 0)
torch.manual_seed(0)
cam = Camera(device=dev, N=N, zernike_terms=6)
h_cpu = synth.height_map(N, 41)
h = h_cpu.to(dev).requires_grad_(True)
cam.get_Heith_Map = lambda: h
img = synth.images(B, N, 42)
img[1] = img[0]                      # identical images must give identical rows
w = synth.upstream_grad(B, N, 43)
x = img.to(dev).requires_grad_(True)
y = cam(x)
((y * w.to(dev)).sum() + cam.loss_rad + cam.centering_loss).backward()
C = co.build_constants(N)
ho = h_cpu.clone().requires_grad_(True)
xo = img.clone().requires_grad_(True)
out = co.camera_forward(xo, ho, C)
((out["sensor"] * w).sum() + out["loss_rad"] + out["centering_loss"]).backward()
rel = lambda a, b: float((a.double().cpu() - b.double()).norm() / b.double().norm())
print("REL", rel(y.detach(), out["sensor"].detach()), rel(h.grad, ho.grad), rel(x.grad, xo.grad))
"""


@pytest.mark.parametrize("B", [3, 52])
def test_plane_kernels_forced_at_256(B):
    """B200CAM_PLANE=1 in a fresh process: the cluster plane kernels (B = 52: several planes per cluster, i.e. the whole
    software pipeline incl. its drain steps) - sensor, dL/dh and the optional dL/dimg against the oracle."""
    import os
    import subprocess
    import sys
    from conftest import REPO
    env = dict(os.environ, B200CAM_PLANE="1")
    res = subprocess.run([sys.executable, "-c", _FORCED_FUSED_SCRIPT.format(repo=str(REPO)), str(B)], env=env,
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    line = [ln for ln in res.stdout.splitlines() if ln.startswith("REL")][-1].split()
    e_sensor, e_grad, e_img = float(line[1]), float(line[2]), float(line[3])
    assert e_sensor <= TOL_SENSOR and e_grad <= TOL_GRAD and e_img <= TOL_GRAD


@pytest.mark.gpu
@pytest.mark.parametrize("T,N", [(12, 64), (300, 256), (37, 128)])
def test_zernike_projection_matches_torch(T, N):
    """h = sum_j coef_j Z_j (Optics.py:79-83) and its adjoint through the C ABI vs the reference's
    torch expression and autograd.  Tolerance: rel-L2 <= 1e-5 (fp32 sums of T / N*N terms in a different order)."""
    from b200cam import functional as F
    dev = torch.device("cuda", 0)
    g = torch.Generator(device="cpu").manual_seed(5)
    Z = (torch.randn(T, N, N, generator=g) * 1e-6).to(dev)
    c_ref = torch.randn(T, 1, 1, generator=g).to(dev).requires_grad_(True)
    c_new = c_ref.detach().clone().requires_grad_(True)
    w = torch.randn(N, N, generator=g).to(dev)
    plan = F.DevicePlan(N, dev)
    h_ref = torch.sum(c_ref * Z, dim=0)
    h_new = F.zernike_project(c_new, Z, plan)
    assert rel_l2(h_new, h_ref) <= 1e-5
    (h_ref * w).sum().backward()
    (h_new * w).sum().backward()
    assert c_new.grad.shape == c_ref.grad.shape
    assert rel_l2(c_new.grad, c_ref.grad) <= 1e-5
    # second call reuses the workspace (arrival counters must have been left at zero)
    assert torch.equal(F.zernike_project(c_new, Z, plan), h_new)


@pytest.mark.gpu
def test_camera_height_map_uses_projection_kernel_and_trains():
    """The module's own get_Heith_Map (learned Zernike coefficients) -> forward -> backward reaches Zer_train."""
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    cam = Camera(device=dev, N=64, zernike_terms=20)
    img = synth.images(2, 64).to(dev)
    y = cam(img)
    (y.sum() + cam.loss_rad + cam.centering_loss).backward()
    assert cam.Zer_train.grad is not None and torch.isfinite(cam.Zer_train.grad).all()
    assert float(cam.Zer_train.grad.abs().max()) > 0
    h_ref = torch.sum(torch.cat((cam.Zer_no_train, cam.Zer_train), 0) * cam.zernike_volume, dim=0)
    assert rel_l2(cam.get_Heith_Map()[0], h_ref) <= 1e-5


@pytest.mark.gpu
def test_uint8_images_are_scaled_on_the_gpu():
    """uint8 input == the reference loader's ToTensor (x/255, Face-DeId/core/data_loader.py:118-124) followed by forward."""
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    cam = Camera(device=dev, N=64, zernike_terms=12)
    u8 = (synth.images(3, 64) * 255).round().to(torch.uint8).to(dev)
    y8 = cam(u8)
    yf = cam(u8.float() / 255.0)
    assert torch.equal(y8, yf)


@pytest.mark.gpu
def test_fused_peer_allreduce_matches_nccl_when_two_gpus_are_visible():
    """Data parallel: the all-reduce of dL/dh fused into the last PSF-backward kernel (NVLink peer memory,
    (value, epoch) words) against the NCCL path, eagerly and through a CUDA graph.  Needs two GPUs: skipped otherwise."""
    import re
    import subprocess
    import sys
    from pathlib import Path
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    repo = Path(__file__).resolve().parent.parent
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29541", str(repo / "tools" / "check_peer_allreduce.py")],
                         capture_output=True, text=True, timeout=300)
    m = re.search(r"PEER_ALLREDUCE world=2 rel_vs_nccl=([0-9.e+-]+) rel_graph=([0-9.e+-]+)", res.stdout + res.stderr)
    assert m, (res.stdout + res.stderr)[-2000:]
    assert float(m.group(1)) <= 1e-5 and float(m.group(2)) <= 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("N,B", [(256, 5), (512, 2)])
def test_opt_in_sensor_noise_and_quantisation(N, B):
    """north_star step 5 / SURVEY trap T6: the opt-in read-out epilogue (include/b200cam.h B200CAM_SENSOR_NOISE / _QUANT).
    Off by default (the reference has neither: Image_Caption/Camera/Lens.py:295-301 is commented out); with it, the
    sensor image equals round(clamp(conv/max + sigma*noise, 0, 1) * L) / L for the SAME torch.randn tensor, and the
    gradient into the height map is the one of the plain read-out (straight through).  N = 256 runs the one-pass inverse-row
    kernel, N = 512 the normalise kernel.  Tolerance: a value within float rounding of a quantisation boundary may fall on
    either side, so at most 0.1 % of the pixels may differ, each by exactly one level."""
    h_cpu, img, w = synth.height_map(N, 3), synth.images(B, N, 4), synth.upstream_grad(B, N, 5)
    out, gh = oracle_step(img, w, h_cpu, N)
    g = torch.Generator(device="cuda").manual_seed(11)
    noise = torch.randn(B, 3, N, N, generator=g, device="cuda")
    sigma, bits = 0.01, 8
    L = float(2 ** bits - 1)
    cam, h = make_camera(N, h_cpu)
    cam.sensor_noise_sigma, cam.sensor_quant_bits = sigma, bits
    y = cam(img.cuda(), noise=noise)
    ((y * w.cuda()).sum() + cam.loss_rad + cam.centering_loss).backward()
    ref = torch.round(torch.clamp(out["sensor"] + sigma * noise.cpu(), 0.0, 1.0) * L) / L
    diff = (y.detach().cpu() - ref).abs()
    assert float(diff.max()) <= 1.0 / L + 1e-6
    assert float((diff > 1e-6).float().mean()) <= 1e-3
    assert torch.all((y.detach() * L - torch.round(y.detach() * L)).abs() < 1e-3)      # on the quantisation grid
    assert rel_l2(h.grad, gh) <= TOL_GRAD                                               # straight-through backward
    # noise only (no quantiser): exact up to fp32 rounding
    cam2, _ = make_camera(N, h_cpu)
    cam2.sensor_noise_sigma = sigma
    with torch.no_grad():
        y2 = cam2(img.cuda(), noise=noise)
    assert rel_l2(y2, out["sensor"] + sigma * noise.cpu()) <= TOL_SENSOR


def test_graphed_callable_matches_eager():
    """Camera.graphed(): forward and backward replayed from CUDA graphs behind an ordinary autograd op - same sensor image,
    regularisers and parameter gradient as the eager module (to rounding: the captured PSF chain is the multi-kernel one).
    Captured BEFORE the module's first eager backward (torch.cuda.make_graphed_callables' rule: an AccumulateGrad node
    created on the default stream would be dragged into the capture)."""
    N, B, T = 128, 5, 10
    torch.manual_seed(3)
    cam = Camera(device=torch.device("cuda", 0), N=N, zernike_terms=T)
    img, w = synth.images(B, N).cuda(), synth.upstream_grad(B, N).cuda()
    step = cam.graphed(img)
    grads = []
    for it in range(2):                                    # the second replay sees the same inputs: identical results
        cam.Zer_train.grad = None
        yg, rad, cen = step(img)
        ((yg * w).sum() + 0.3 * rad + 1.7 * cen).backward()
        grads.append(cam.Zer_train.grad.clone())
    graphed_y = yg.detach().clone()
    img2 = synth.images(B, N, seed=77).cuda()              # new data through the static buffers
    graphed_y2 = step(img2)[0].detach().clone()
    cam.Zer_train.grad = None
    y = cam(img)
    ((y * w).sum() + 0.3 * cam.loss_rad + 1.7 * cam.centering_loss).backward()
    assert torch.equal(grads[0], grads[1])
    assert rel_l2(graphed_y, y) <= 1e-5
    assert rel_l2(grads[0], cam.Zer_train.grad) <= 1e-4
    assert rel_l2(graphed_y2, cam(img2)) <= 1e-5


@pytest.mark.parametrize("switch", ["B200CAM_TIE_SPECTRAL=1", "B200CAM_ONE_PASS=1", "B200CAM_OTF_ROWS=0", "B200CAM_PLANE=1",
                                    "B200CAM_COOP=0", "B200CAM_HOLD_ROWS=0"])
def test_opt_in_switches_keep_parity(switch):
    """Every A/B switch of INTEGRATION.md section 4 selects another kernel path for the same maths: each must stay inside the
    tolerances (the switches are read once per process, hence a subprocess per switch; tests/switch_parity_runner.py)."""
    import os
    import subprocess
    import sys
    from conftest import REPO
    k, v = switch.split("=")
    res = subprocess.run([sys.executable, str(REPO / "tests" / "switch_parity_runner.py")], capture_output=True, text=True,
                         timeout=600, env=dict(os.environ, **{k: v}))
    assert res.returncode == 0 and "ok=True" in res.stdout, (res.stdout + res.stderr)[-2000:]
