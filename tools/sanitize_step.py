"""Developer tool (GPU box): ONE small camera step (Face-DeId Camera fwd+bwd, then the Image_Caption camera fwd+bwd) for
`compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize_step.py`.
Small sizes: the sanitizer slows kernels by 10-100x."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import b200cam.synthetic as synth          # noqa: E402
from b200cam.optics import Camera          # noqa: E402
from b200cam.lens import OpticsZernike     # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    N, B = int(sys.argv[1]) if len(sys.argv) > 1 else 256, 3
    torch.manual_seed(0)
    cam = Camera(device=dev, N=N, zernike_terms=12)
    h = synth.height_map(N).to(dev).requires_grad_(True)
    cam.get_Heith_Map = lambda: h
    img, w = synth.images(B, N).to(dev), synth.upstream_grad(B, N).to(dev)
    y = cam(img)
    ((y * w).sum() + cam.loss_rad + cam.centering_loss).backward()
    torch.cuda.synchronize()
    print("camera step ok", float(h.grad.abs().sum()))

    lens = OpticsZernike(input_shape=[1, 64, 64, 3], device=dev, wave_resolution=(128, 128), patch_size=64,
                         sample_interval=3e-6, zernike_terms=10, height_tolerance=2e-8).to(dev)
    x = torch.rand(2, 3, 64, 64, device=dev)
    sensor, psf, _, _ = lens(x)
    (sensor * torch.rand_like(sensor)).sum().backward()
    torch.cuda.synchronize()
    print("lens step ok", float(lens.zernike_coeffs_train.grad))


if __name__ == "__main__":
    main()
