"""Multi-GPU check (torchrun, 2 ranks): a rank whose peer never joins the fused all-reduce must NOT get a silent partial
sum - dL/dh comes back as NaN and the device error word makes the wrapper raise.
usage: B200CAM_COMM_TIMEOUT_S=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_peer_timeout.py
"""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import b200cam.synthetic as synth          # noqa: E402
from b200cam import parallel               # noqa: E402
from b200cam.optics import Camera          # noqa: E402


def main():
    rank, world, local = parallel.init_from_env("nccl")
    dev = torch.device("cuda", local)
    N, B = 256, 4
    torch.manual_seed(0)
    cam = Camera(device=dev, N=N, zernike_terms=6)
    h = synth.height_map(N, 7).to(dev).requires_grad_(True)
    cam.get_Heith_Map = lambda: h
    cam.data_parallel(average=True, peer_memory=True)
    assert cam._peer_comm is not None
    img, w = synth.images(B, N, 100 + rank).to(dev), synth.upstream_grad(B, N, 200 + rank).to(dev)

    def step():
        h.grad = None
        y = cam(img)
        ((y * w).sum() + cam.loss_rad + cam.centering_loss).backward()

    step()                                   # both ranks: fine
    torch.cuda.synchronize()
    cam.check_device_errors()
    assert torch.isfinite(h.grad).all()
    ok = True
    if rank == 0:
        step()                               # rank 1 does not join: rank 0 must time out, not return garbage
        torch.cuda.synchronize()
        nan = bool(torch.isnan(h.grad).any())
        try:
            cam.check_device_errors()
            raised = False
        except RuntimeError as exc:
            raised = "all-reduce" in str(exc)
        ok = nan and raised
        print(f"PEER_TIMEOUT nan={nan} raised={raised}")
    dist.barrier()
    sys.stdout.flush()
    os._exit(0 if ok else 1)


if __name__ == "__main__":
    main()
