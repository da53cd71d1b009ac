"""GPU check of one camera step (N = 256, B = 5 and B = 70) against the CPU oracle, for use under an environment switch
(B200CAM_TIE_SPECTRAL=1, B200CAM_ONE_PASS=1, B200CAM_OTF_ROWS=0, B200CAM_PLANE=1 ...): the switches are read once per
process, so tests/test_gpu_parity.py::test_opt_in_switches_keep_parity runs this file (test infrastructure: it uses the oracle) in a subprocess per switch."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import b200cam.synthetic as synth          # noqa: E402
from b200cam.optics import Camera          # noqa: E402
from oracle import camera_oracle as co     # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    N = 256
    C = co.build_constants(N)
    worst = [0.0, 0.0]
    for B in (5, 70):
        torch.manual_seed(0)
        cam = Camera(device=dev, N=N, zernike_terms=8)
        h_cpu = synth.height_map(N, 3)
        img, w = synth.images(B, N, 40 + B), synth.upstream_grad(B, N, 50 + B)
        h = h_cpu.to(dev).requires_grad_(True)
        cam.get_Heith_Map = lambda: h
        y = cam(img.to(dev))
        ((y * w.to(dev)).sum() + cam.loss_rad + cam.centering_loss).backward()
        cam.check_device_errors()
        ho = h_cpu.clone().requires_grad_(True)
        out = co.camera_forward(img, ho, C)
        ((out["sensor"] * w).sum() + out["loss_rad"] + out["centering_loss"]).backward()
        e_y = float((y.detach().cpu().double() - out["sensor"].double()).norm() / out["sensor"].double().norm())
        e_g = float((h.grad.cpu().double() - ho.grad.double()).norm() / ho.grad.double().norm())
        worst = [max(worst[0], e_y), max(worst[1], e_g)]
    ok = worst[0] <= 1e-4 and worst[1] <= 1e-3
    print(f"SWITCH_PARITY sensor={worst[0]:.2e} grad_h={worst[1]:.2e} ok={ok}", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
