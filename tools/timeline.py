"""Developer tool (GPU box): per-kernel TIMELINE of one camera step replayed from a CUDA graph.

ncu serialises kernels and runs them cold, so it cannot show what overlaps what, nor the idle gaps between dependent
kernels.  CUPTI activity records (through torch.profiler) keep the concurrent picture: every kernel's start, duration
and stream inside a graph replay.  Prints one median step as a table (start relative to the step's first kernel) and
the critical-path bookkeeping: busy time (union of kernel intervals), idle gaps, step period.

usage: python tools/timeline.py [B] [N] [steps] [out.json]
"""
import json
import sys
from pathlib import Path

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import b200cam.synthetic as synth          # noqa: E402
from b200cam.optics import Camera          # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 12
    out = sys.argv[4] if len(sys.argv) > 4 else None
    R = 4
    import os
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:               # under torchrun: data-parallel step (fused all-reduce of dL/dh in the last kernel)
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    cam = Camera(device=dev, N=N, zernike_terms=12)
    if world > 1:
        cam.data_parallel(average=True)
    h = synth.height_map(N).to(dev).requires_grad_(True)
    cam.get_Heith_Map = lambda: h
    imgs = [synth.images(B, N, seed=1000 + r).to(dev) for r in range(R)]
    ws = [synth.upstream_grad(B, N, seed=2000 + r).to(dev) for r in range(R)]
    one = torch.ones((), device=dev)

    def step(i):
        h.grad = None
        y = cam(imgs[i % R])
        torch.autograd.backward([y, cam.loss_rad, cam.centering_loss], [ws[i % R], one, one])

    def drop():
        cam.psfs = None
        cam.loss_rad = cam.centering_loss = cam._pending_centering = None
        h.grad = None

    step(0)
    torch.cuda.synchronize()
    drop()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for r in range(R):
            step(r)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graphs = []
    for r in range(R):
        g = torch.cuda.CUDAGraph()
        drop()
        with torch.cuda.graph(g):
            step(r)
        graphs.append(g)
    for i in range(50):
        graphs[i % R].replay()
    torch.cuda.synchronize()

    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for i in range(steps):
            graphs[i % R].replay()
        torch.cuda.synchronize()

    if rank != 0:
        torch.cuda.synchronize()
        os._exit(0)
    evs = []
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None:
            name = e.name
            if name.startswith("Memcpy") or name.startswith("Memset") or "(" in name or "kernel" in name.lower() or "k_" in name:
                evs.append((e.time_range.start, e.time_range.end, name))
    evs.sort()
    if not evs:
        print("no CUDA kernel records (CUPTI unavailable?)")
        return
    # split into steps: a step starts at each k_crows_fwd<.., PupilLoad> / k_psf_fwd_coop (first kernel of the PSF chain)
    # or k_rows_r2c that follows a psf_bwd kernel; simpler: split on gaps between the last kernel of the step
    # (name contains "hgrad" or "psf_bwd") and the next record
    steps_ev, cur = [], []
    for s, t, n in evs:
        cur.append((s, t, n))
        if "hgrad" in n or "psf_bwd_coop" in n:
            steps_ev.append(cur)
            cur = []
    if not steps_ev:
        steps_ev = [evs]
    periods = [steps_ev[i + 1][0][0] - steps_ev[i][0][0] for i in range(len(steps_ev) - 1)]
    periods_sorted = sorted(periods)
    med_period = periods_sorted[len(periods_sorted) // 2] if periods else float("nan")
    k = len(steps_ev) // 2
    one_step = steps_ev[k]
    t0 = one_step[0][0]
    print(f"B={B} N={N}: {len(steps_ev)} steps recorded, median step period {med_period:.1f} us "
          f"({B / med_period * 1e6:.0f} images/s)" if periods else "single step")
    print(f"{'start':>8s} {'dur':>7s} {'end':>8s}  kernel")
    rows = []
    for s, t, n in one_step:
        short = n.replace("b200cam::", "").replace("void ", "")
        short = short[:90]
        print(f"{s - t0:8.1f} {t - s:7.1f} {t - t0:8.1f}  {short}")
        rows.append({"start_us": s - t0, "dur_us": t - s, "name": short})
    # union of intervals = busy time
    busy, end = 0.0, None
    gaps = []
    for s, t, n in one_step:
        if end is None or s > end:
            if end is not None:
                gaps.append((s - end, n))
            busy += t - s
            end = t
        elif t > end:
            busy += t - end
            end = t
    span = one_step[-1][1] - t0
    print(f"span {span:.1f} us, busy (union) {busy:.1f} us, idle inside the step {span - busy:.1f} us over {len(gaps)} gaps; "
          f"sum of kernel durations {sum(t - s for s, t, _ in one_step):.1f} us")
    for g, n in gaps:
        print(f"   gap {g:5.1f} us before {n.replace('b200cam::', '')[:70]}")
    if world > 1:
        sys.stdout.flush()
        os._exit(0)
    if out:
        Path(out).write_text(json.dumps({"B": B, "N": N, "median_period_us": med_period, "span_us": span, "busy_us": busy,
                                         "kernels": rows}, indent=1))


if __name__ == "__main__":
    main()
