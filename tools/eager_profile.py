"""Developer tool (GPU box): host-side cost of one eager (no CUDA graph) camera step - CPU issue time per step and a
cProfile of where it goes.  Measured: 396 us/step of CPU time against 177 us of device time at B=64, N=256, i.e. an eager
loop is launch-bound; capture the step in a CUDA graph (tests/test_gpu_parity.py::test_cuda_graph_capture_of_forward_backward).
usage: python tools/eager_profile.py"""
import cProfile
import io
import pstats
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import b200cam.synthetic as synth
from b200cam.optics import Camera
dev = torch.device('cuda', 0)
N, B = 256, 64
torch.manual_seed(0)
cam = Camera(device=dev, N=N, zernike_terms=12)
h = synth.height_map(N).to(dev).requires_grad_(True)
cam.get_Heith_Map = lambda: h
img = synth.images(B, N).to(dev); w = synth.upstream_grad(B, N).to(dev); one = torch.ones((), device=dev)
def step():
    h.grad = None
    y = cam(img)
    torch.autograd.backward([y, cam.loss_rad, cam.centering_loss], [w, one, one])
for _ in range(20): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(300): step()
t_issue = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print(f"eager: CPU issue {t_issue/300*1e6:.0f} us/step, wall {t_all/300*1e6:.0f} us/step")
pr = cProfile.Profile(); pr.enable()
for _ in range(200): step()
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(18); print(s.getvalue()[:3500])
