"""CPU, world_size 2, gloo: the data-parallel host logic (batch sharding + dL/dh all-reduce).

Each rank computes the gradient of its shard with the oracle (the CUDA path is exercised by the -m gpu
tests and by `bench.py --gpus N`); after `allreduce_height_grad(average=True)` every rank must hold the
gradient of  mean_over_ranks(sum(sensor*w) on the shard) + loss_rad + centering_loss  computed in one process.
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import b200cam.synthetic as synth
from b200cam import parallel
from conftest import rel_l2
from oracle import camera_oracle as co

N, B, WORLD = 64, 5, 2


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _grad(img, w, h0, scale):
    C = co.build_constants(N)
    h = h0.clone().requires_grad_(True)
    out = co.camera_forward(img, h, C)
    (scale * (out["sensor"] * w).sum() + out["loss_rad"] + out["centering_loss"]).backward()
    return h.grad


def _worker(rank: int, port: int, result_dir: str):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(WORLD),
                      LOCAL_RANK=str(rank))
    torch.set_num_threads(2)
    r, world, _ = parallel.init_from_env("gloo")
    assert (r, world) == (rank, WORLD)
    img, w, h0 = synth.images(B, N, 3), synth.upstream_grad(B, N, 4), synth.height_map(N, 5)
    g = _grad(parallel.shard(img, rank, world), parallel.shard(w, rank, world), h0, 1.0)
    parallel.allreduce_height_grad(g, None, average=True)
    torch.save(g, os.path.join(result_dir, f"grad_{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_cover_the_batch():
    for batch, world in [(64, 8), (5, 2), (7, 4), (3, 8)]:
        spans = [parallel.shard_range(batch, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == batch
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.shard_range(8, 2, 2)


def test_allreduce_is_identity_without_process_group():
    g = torch.ones(4, 4)
    assert parallel.allreduce_height_grad(g) is g and torch.equal(g, torch.ones(4, 4))


@pytest.mark.timeout(180)
def test_two_rank_gradient_equals_single_process(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(port, str(tmp_path)), nprocs=WORLD, join=True)
    g0 = torch.load(tmp_path / "grad_0.pt")
    g1 = torch.load(tmp_path / "grad_1.pt")
    assert torch.equal(g0, g1)                                    # all ranks agree bit for bit
    img, w, h0 = synth.images(B, N, 3), synth.upstream_grad(B, N, 4), synth.height_map(N, 5)
    ref = _grad(img, w, h0, 1.0 / WORLD)                           # mean over ranks of the data term
    assert rel_l2(g0, ref) <= 1e-5


# ---------------------------------------------------------------------------------------------------------------
# Image_Caption camera: batch-global max (Lens.py:312) over two ranks == one process
# ---------------------------------------------------------------------------------------------------------------
def _max_worker(rank: int, port: int, result_dir: str):
    from b200cam.lens import GlobalMaxNormalise
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(WORLD),
                      LOCAL_RANK=str(rank))
    torch.set_num_threads(1)
    parallel.init_from_env("gloo")
    g = torch.Generator().manual_seed(11)
    x_all, w_all = torch.rand(6, 3, 8, 8, generator=g), torch.rand(6, 3, 8, 8, generator=g)
    x = parallel.shard(x_all, rank, WORLD).clone().requires_grad_(True)
    y = GlobalMaxNormalise.apply(x, dist.group.WORLD)
    (y * parallel.shard(w_all, rank, WORLD)).sum().backward()
    torch.save((y.detach(), x.grad), os.path.join(result_dir, f"max_{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_global_max_equals_single_process(tmp_path):
    port = _free_port()
    mp.spawn(_max_worker, args=(port, str(tmp_path)), nprocs=WORLD, join=True)
    g = torch.Generator().manual_seed(11)
    x_all, w_all = torch.rand(6, 3, 8, 8, generator=g), torch.rand(6, 3, 8, 8, generator=g)
    x = x_all.clone().requires_grad_(True)
    (x / x.max() * w_all).sum().backward()                        # the reference expression (Lens.py:312)
    ys, gs = zip(*[torch.load(os.path.join(str(tmp_path), f"max_{r}.pt")) for r in range(WORLD)])
    assert torch.allclose(torch.cat(ys), (x / x.max()).detach(), rtol=1e-6, atol=0)
    assert torch.allclose(torch.cat(gs), x.grad, rtol=1e-5, atol=1e-6)
