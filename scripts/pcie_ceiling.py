"""Ceiling of the host <-> device copies the end-to-end leg of bench.py depends on: every rank (one per GPU, torchrun) moves the
bench's per-step byte counts -- H2D of one step's pictures, D2H of one step's results -- between pinned host memory and its GPU
on two streams, nothing else running; all ranks start together.  Prints per-rank and aggregate GB/s and the frames/s those copies
alone would allow.  Usage: torchrun --nproc-per-node N scripts/pcie_ceiling.py [h2d_MB d2h_MB steps]"""
import os, sys, time, json
import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
h2d_mb = float(sys.argv[1]) if len(sys.argv) > 1 else 199.07      # 64 x 1080p yuv420p pictures
d2h_mb = float(sys.argv[2]) if len(sys.argv) > 2 else 31.6        # 64 x (per-MB decisions + packed levels), BENCH_r01
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 40
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
h_in = torch.empty(int(h2d_mb * 1e6), dtype=torch.uint8).pin_memory(); d_in = torch.empty_like(h_in, device="cuda")
h_out = torch.empty(int(d2h_mb * 1e6), dtype=torch.uint8).pin_memory(); d_out = torch.empty_like(h_out, device="cuda")
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
res = {}
for mode in ("h2d", "d2h", "both"):
    for _ in range(3):
        with torch.cuda.stream(s_in): d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s_out): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1: dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        if mode in ("h2d", "both"):
            with torch.cuda.stream(s_in): d_in.copy_(h_in, non_blocking=True)
        if mode in ("d2h", "both"):
            with torch.cuda.stream(s_out): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); dt = float(t.item())
    nbytes = steps * ((h2d_mb if mode != "d2h" else 0) + (d2h_mb if mode != "h2d" else 0)) * 1e6
    res[mode] = {"seconds": round(dt, 4), "GBps_per_gpu": round(nbytes / dt / 1e9, 2), "GBps_all": round(world * nbytes / dt / 1e9, 2)}
res["frames_per_s_if_copies_only"] = round(world * 64 * steps / res["both"]["seconds"])
if rank == 0:
    print(json.dumps({"n_gpus": world, "h2d_MB_per_step": h2d_mb, "d2h_MB_per_step": d2h_mb, "steps": steps, **res}), flush=True)
if world > 1:
    dist.destroy_process_group()
