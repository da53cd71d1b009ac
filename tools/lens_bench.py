"""Developer timing (GPU box): Image_Caption camera (OpticsZernike, shipped geometry of train.py:64-66) forward+backward
into the trainable Zernike coefficient, eager, batch B of 256x256 RGB.  usage: python tools/lens_bench.py [B] [reps]"""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from b200cam.lens import OpticsZernike      # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    dev = torch.device("cuda", 0)
    cam = OpticsZernike(input_shape=[None, 256, 256, 3], device=dev, zernike_terms=350, patch_size=256,
                        height_tolerance=2e-8, sensor_distance=0.025, wave_resolution=[896, 896],
                        sample_interval=3e-06, upsample=False).to(dev)      # as train.py:270 does
    img = torch.rand(B, 3, 256, 256, device=dev)
    w = torch.rand(B, 3, 256, 256, device=dev)

    def step():
        cam.zero_grad(set_to_none=True)
        sensor, psf, coeffs, loss = cam(img)
        (sensor * w).sum().backward()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    print(f"OpticsZernike fwd+bwd B={B}: {dt * 1e3:.2f} ms/step, {B / dt:.0f} images/s "
          f"(grad of the trainable coefficient: {float(cam.zernike_coeffs_train.grad.abs().sum()):.3e})")


if __name__ == "__main__":
    main()
