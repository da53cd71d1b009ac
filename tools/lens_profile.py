"""Developer tool (GPU box): per-kernel device times (CUPTI through torch.profiler) of one forward+backward of the
Image_Caption camera at the shipped geometry.  usage: python tools/lens_profile.py [B]"""
import sys
from collections import defaultdict
from pathlib import Path

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from b200cam.lens import OpticsZernike      # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    dev = torch.device("cuda", 0)
    cam = OpticsZernike(input_shape=[None, 256, 256, 3], device=dev, zernike_terms=350, patch_size=256,
                        height_tolerance=2e-8, sensor_distance=0.025, wave_resolution=[896, 896],
                        sample_interval=3e-06, upsample=False).to(dev)
    img = torch.rand(B, 3, 256, 256, device=dev)
    w = torch.rand(B, 3, 256, 256, device=dev)

    def step():
        cam.zero_grad(set_to_none=True)
        sensor, psf, coeffs, loss = cam(img)
        torch.autograd.backward([sensor], [w])

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    reps = 5
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(reps):
            step()
        torch.cuda.synchronize()
    tot = defaultdict(float)
    cnt = defaultdict(int)
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            tot[ev.name] += ev.device_time
            cnt[ev.name] += 1
    total = sum(tot.values()) / reps
    print(f"B={B}: sum of kernel times {total:.1f} us per step")
    for name, t in sorted(tot.items(), key=lambda kv: -kv[1])[:40]:
        print(f"{t / reps:9.1f} us  x{cnt[name] / reps:4.1f}  {name[:140]}")


if __name__ == "__main__":
    main()
