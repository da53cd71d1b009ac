"""Multi-GPU check (run under torchrun, one rank per GPU): the fused peer-memory all-reduce of dL/dh
(b200cam_psf_bwd_allreduce) against the NCCL all-reduce path, over several steps and through a CUDA graph.
usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_peer_allreduce.py
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
    N, B = 256, 8

    def build(peer):
        torch.manual_seed(0)
        cam = Camera(device=dev, N=N, zernike_terms=6)
        h = synth.height_map(N, 7).to(dev).requires_grad_(True)
        cam.get_Heith_Map = lambda: h
        cam.data_parallel(average=True, peer_memory=peer)
        return cam, h

    cam_p, h_p = build(True)
    cam_n, h_n = build(False)
    assert cam_p._peer_comm is not None, "peer-memory all-reduce was not set up"
    worst = 0.0
    for step in range(5):
        img = synth.images(B, N, 100 + 10 * step + rank).to(dev)
        w = synth.upstream_grad(B, N, 200 + 10 * step + rank).to(dev)
        grads = []
        for cam, h in ((cam_p, h_p), (cam_n, h_n)):
            h.grad = None
            y = cam(img)
            ((y * w).sum() + cam.loss_rad + cam.centering_loss).backward()
            grads.append(h.grad.clone())
        rel = float((grads[0] - grads[1]).norm() / grads[1].norm())
        worst = max(worst, rel)
        gathered = [torch.empty_like(grads[0]) for _ in range(world)]
        dist.all_gather(gathered, grads[0])
        assert all(torch.equal(g, gathered[0]) for g in gathered), "ranks disagree bitwise"
    # CUDA graph replay (frozen kernel arguments: epochs / parities must advance inside the buffer)
    img = synth.images(B, N, 900 + rank).to(dev)
    w = synth.upstream_grad(B, N, 901 + rank).to(dev)
    one = torch.ones((), device=dev)

    def step_fn():
        h_p.grad = None
        y = cam_p(img)
        torch.autograd.backward([y, cam_p.loss_rad, cam_p.centering_loss], [w, one, one])

    step_fn()
    ref = h_p.grad.clone()
    cam_p.psfs = None; cam_p.loss_rad = cam_p.centering_loss = cam_p._pending_centering = None; h_p.grad = None
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step_fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    cam_p.psfs = None; cam_p.loss_rad = cam_p.centering_loss = cam_p._pending_centering = None; h_p.grad = None
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        step_fn()
    for _ in range(7):
        g.replay()
    torch.cuda.synchronize()
    rel_graph = float((h_p.grad - ref).norm() / ref.norm())
    # timing: graph replay with the fused all-reduce
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier(); torch.cuda.synchronize()
    e0.record()
    for _ in range(200):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 200 * 1e3
    if rank == 0:
        print(f"PEER_ALLREDUCE world={world} rel_vs_nccl={worst:.3e} rel_graph={rel_graph:.3e} step_us={us:.1f}")
    assert worst <= 1e-5 and rel_graph <= 1e-6
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
