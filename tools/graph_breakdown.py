"""Developer timing (GPU box): each C-ABI call of the camera step captured in its own CUDA graph and replayed,
so the numbers are device times without Python / ctypes launch overhead.  Inputs rotate over R sets (> L2).
usage: [B200CAM_FUSED=0|1] python tools/graph_breakdown.py [B] [N] [reps]
"""
import ctypes
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import b200cam.synthetic as synth          # noqa: E402
from b200cam import _lib                   # noqa: E402
from b200cam import functional as F        # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 200
    R = 4
    dev = torch.device("cuda", 0)
    plan = F.DevicePlan(N, dev)
    lib, p = plan.lib, _lib.ptr
    h = synth.height_map(N).to(dev).reshape(N, N).contiguous()
    xs = [synth.images(B, N, 100 + r).to(dev) for r in range(R)]
    gs = [synth.upstream_grad(B, N, 200 + r).to(dev) for r in range(R)]
    psf = torch.empty(3, N, N, device=dev)
    field = torch.empty(3, N, N, 2, device=dev)
    stats = torch.empty(4, device=dev)
    pws = plan.psf_workspace()
    sensor = torch.empty_like(xs[0])
    img_max = torch.empty(B, device=dev)
    tie_count = torch.empty(B, dtype=torch.int32, device=dev)
    tie_pos = torch.empty(B, 8, dtype=torch.int32, device=dev)
    otf = torch.empty(plan.otf_floats, device=dev)
    spectrum = torch.empty(lib.b200cam_spectrum_bytes(N, B) // 4, device=dev)
    ws = plan.sensor_workspace(B)
    gpsf = torch.zeros(3, N, N, device=dev)
    gscal = torch.ones(2, device=dev)
    gh = torch.empty(N, N, device=dev)

    def st():
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def psf_fwd(i):
        _lib.check(lib.b200cam_psf_fwd(p(h), p(plan.A), p(plan.Ht), p(plan.rho), plan.kappa, p(psf), p(field), p(stats),
                                       p(pws), pws.numel(), N, st()))

    def sensor_fwd(i):
        _lib.check(lib.b200cam_sensor_fwd(p(xs[i % R]), p(psf), p(sensor), p(img_max), p(tie_count), p(tie_pos), p(otf),
                                          p(spectrum), p(ws), ws.numel(), B, N, st()))

    def sensor_bwd(i):
        _lib.check(lib.b200cam_sensor_bwd(p(gs[i % R]), p(xs[i % R]), p(sensor), p(img_max), p(tie_count), p(tie_pos),
                                          p(psf), p(otf), p(spectrum), p(gpsf), None, p(ws), ws.numel(), B, N, st()))

    def psf_bwd(i):
        _lib.check(lib.b200cam_psf_bwd(p(gpsf), p(gscal[0:1]), p(gscal[1:2]), p(h), p(plan.A), p(plan.Ht), p(plan.rho), plan.kappa, p(psf),
                                       p(field), p(stats), p(gh), p(pws), pws.numel(), N, st()))

    def step(i):
        psf_fwd(i); sensor_fwd(i); sensor_bwd(i); psf_bwd(i)

    step(0)
    torch.cuda.synchronize()

    def timed(name, fn):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fn(0)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graphs = []
        for r in range(R):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                fn(r)
            graphs.append(g)
        for i in range(8):
            graphs[i % R].replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps):
            graphs[i % R].replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / reps * 1e3
        print(f"{name:12s} {us:8.1f} us")
        return us

    print(f"B={B} N={N} graph replay, {reps} reps")
    # order matters: each call needs the outputs of the previous ones to be valid
    t = [timed("psf_fwd", psf_fwd), timed("sensor_fwd", sensor_fwd), timed("sensor_bwd", sensor_bwd), timed("psf_bwd", psf_bwd)]
    tot = timed("whole step", step)
    print(f"sum of parts {sum(t):.1f} us; step {tot:.1f} us -> {B / tot * 1e6:.0f} images/s")


if __name__ == "__main__":
    main()
