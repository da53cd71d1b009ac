"""Developer check (GPU box): the sensor kernels of libb200cam.so against torch.fft on the same device.

Not a product path and not a parity test (those live in tests/ and use the oracle): this only localises
an error to forward / saved spectrum / backward while a kernel is being written.
usage: python tools/debug_fused.py [B] [N]
"""
import ctypes
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import b200cam.synthetic as synth          # noqa: E402
from b200cam import _lib                   # noqa: E402
from b200cam import functional as F        # noqa: E402


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    dev = torch.device("cuda", 0)
    plan = F.DevicePlan(N, dev)
    lib = plan.lib
    g = torch.Generator().manual_seed(3)
    psf = torch.rand(3, N, N, generator=g)
    psf = (psf / psf.sum()).to(dev)
    x = synth.images(B, N, 11).to(dev)
    w = synth.upstream_grad(B, N, 12).to(dev)

    sensor = torch.empty_like(x)
    img_max = torch.empty(B, device=dev)
    tie_count = torch.empty(B, dtype=torch.int32, device=dev)
    tie_pos = torch.empty(B, 8, dtype=torch.int32, device=dev)
    otf = torch.empty(plan.otf_floats, device=dev)
    spectrum = torch.zeros(lib.b200cam_spectrum_bytes(N, B) // 4, device=dev)
    ws = plan.sensor_workspace(B)
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = _lib.ptr
    _lib.check(lib.b200cam_sensor_fwd(p(x), p(psf), p(sensor), p(img_max), p(tie_count), p(tie_pos), p(otf), p(spectrum),
                                      p(ws), ws.numel(), B, N, stream))
    torch.cuda.synchronize()

    # torch reference on the GPU
    xr = x.clone()
    pr = psf.clone().requires_grad_(True)
    Kf = torch.fft.rfft2(torch.roll(pr, (-N // 2, -N // 2), (-2, -1)))
    conv = torch.fft.irfft2(torch.fft.rfft2(xr) * Kf, s=(N, N))
    m = conv.amax(dim=(1, 2, 3), keepdim=True)
    y = conv / m
    print(f"B={B} N={N}")
    print("sensor rel", rel(sensor, y.detach()), " max rel", rel(img_max, m.flatten().detach()),
          " ties", tie_count.tolist()[:8])
    if N == 256:
        X = torch.view_as_complex(spectrum.view(3 * B, N // 2 + 1, N, 2))           # [plane][u][v]
        Xref = 2 * torch.fft.rfft2(xr).reshape(3 * B, N, N // 2 + 1).transpose(1, 2)
        print("spectrum rel", rel(torch.view_as_real(X), torch.view_as_real(Xref.contiguous())))
        Kt = torch.view_as_complex(otf.view(3, N // 2 + 1, N, 2))
        Kref = (Kf.detach() / (2 * N * N)).transpose(1, 2)
        print("otf rel", rel(torch.view_as_real(Kt), torch.view_as_real(Kref.contiguous())))
        for b in range(min(B, 3)):
            for c in range(3):
                print(f"  plane ({b},{c}) rel {rel(sensor[b, c], y[b, c].detach()):.2e}", end="")
            print()

    (y * w).sum().backward()
    gpsf = torch.zeros(3, N, N, device=dev)
    _lib.check(lib.b200cam_sensor_bwd(p(w), p(x), p(sensor), p(img_max), p(tie_count), p(tie_pos), p(psf), p(otf),
                                      p(spectrum), p(gpsf), None, p(ws), ws.numel(), B, N, stream))
    torch.cuda.synchronize()
    print("grad_psf rel", rel(gpsf, pr.grad))

    # timing (eager launches, CUDA events)
    for name, fn in (("fwd", lambda: lib.b200cam_sensor_fwd(p(x), p(psf), p(sensor), p(img_max), p(tie_count), p(tie_pos),
                                                            p(otf), p(spectrum), p(ws), ws.numel(), B, N, stream)),
                     ("bwd", lambda: lib.b200cam_sensor_bwd(p(w), p(x), p(sensor), p(img_max), p(tie_count), p(tie_pos),
                                                            p(psf), p(otf), p(spectrum), p(gpsf), None, p(ws), ws.numel(),
                                                            B, N, stream))):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            fn()
        e1.record()
        torch.cuda.synchronize()
        print(f"{name}: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us")


if __name__ == "__main__":
    main()
