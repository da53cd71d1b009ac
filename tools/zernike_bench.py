"""Developer timing (GPU box): Zernike projection h = sum_j coef_j Z_j forward+backward, torch expression
(`torch.sum(coef * volume, 0)`, Face-DeId/Camera/Optics.py:79-83) vs b200cam_zernike_fwd/_bwd.
usage: python tools/zernike_bench.py [T] [N]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from b200cam import functional as F        # noqa: E402


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    dev = torch.device("cuda", 0)
    # "basis" as the third argument: the real Noll basis (zero outside the unit disc: the kernels then skip 21 % of the square)
    if len(sys.argv) > 3 and sys.argv[3] == "basis":
        import numpy as np
        from b200cam.zernike import zernike_volume
        Z = torch.tensor(zernike_volume(N, T, 1e-6).astype(np.float32), device=dev)
    else:
        Z = torch.randn(T, N, N, device=dev) * 1e-6
    c = torch.randn(T, 1, 1, device=dev, requires_grad=True)
    w = torch.randn(N, N, device=dev)
    plan = F.DevicePlan(256, dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timed(fn, reps=30):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(reps):
            flush.zero_()                       # basis volume out of L2, as in a real step
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            tot += a.elapsed_time(b)
        return tot / reps * 1e3

    def ref():
        c.grad = None
        (torch.sum(c * Z, dim=0) * w).sum().backward()

    def new():
        c.grad = None
        (F.zernike_project(c, Z, plan) * w).sum().backward()

    t_ref, t_new = timed(ref), timed(new)
    # the two kernels alone (no autograd glue), through the ABI
    from b200cam import _lib
    cc = c.detach().reshape(T).contiguous()
    h = torch.empty(N, N, device=dev)
    gc = torch.empty(T, device=dev)
    ws = plan.zernike_workspace(T, N * N)
    act = plan.zernike_support(Z)
    na = act.numel() if act is not None else 0

    def kernels():
        _lib.check(plan.lib.b200cam_zernike_fwd_ex(_lib.ptr(cc), _lib.ptr(Z), _lib.ptr(h), _lib.ptr(ws), ws.numel(), T, N * N, F._stream(),
                                                   _lib.ptr(act), na))
        _lib.check(plan.lib.b200cam_zernike_bwd_ex(_lib.ptr(w), _lib.ptr(Z), _lib.ptr(gc), T, N * N, F._stream(), _lib.ptr(act), na))

    def kernels_full():
        _lib.check(plan.lib.b200cam_zernike_fwd(_lib.ptr(cc), _lib.ptr(Z), _lib.ptr(h), _lib.ptr(ws), ws.numel(), T, N * N, F._stream()))
        _lib.check(plan.lib.b200cam_zernike_bwd(_lib.ptr(w), _lib.ptr(Z), _lib.ptr(gc), T, N * N, F._stream()))

    def graphed(fn):                 # two ctypes launches cost more host time than the kernels take: replay them from a graph
        fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        return g.replay

    t_k, t_kf = timed(graphed(kernels)), timed(graphed(kernels_full))
    print(f"kernels only: {t_k:.1f} us with the support list ({na} of {N * N // 4} float4 positions), {t_kf:.1f} us over the full square")
    nbytes = 2 * T * N * N * 4
    print(f"T={T} N={N}: torch {t_ref:.1f} us, b200cam {t_new:.1f} us ({nbytes / t_new * 1e-3:.0f} GB/s over the two passes "
          f"incl. autograd glue), speed-up {t_ref / t_new:.2f}x")


if __name__ == "__main__":
    main()
