#!/usr/bin/env python
"""Copy-only ceiling of the host path bench.py's end-to-end number rides on.

Every rank (one per GPU, same process layout as bench.py under torchrun) moves exactly the bytes one bench step
moves -- CLIPS clips x 441 000 float32 samples host->device and CLIPS x 858 x 1024 result bytes device->host --
between page-locked host buffers and device memory with plain cudaMemcpyAsync on two streams, both directions
in flight at once, no kernel.  The time of the slowest rank gives the frames/s no implementation of the host-buffer
API can exceed on this box; bench.py prints e2e.frac_of_host_ceiling against it.

  python tools/host_path_ceiling.py [--clips 512] [--steps 5] [--bytes-per-sample 4]
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/host_path_ceiling.py
Prints one JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import time

import torch

CLIP_LEN, N_FFT, HOP = 441000, 2048, 512


def measure(dev, clips: int, steps: int, bytes_per_sample: int, barrier=lambda: None) -> dict:
    frames = clips * (1 + (CLIP_LEN - N_FFT) // HOP)
    n_in, n_out = clips * CLIP_LEN * bytes_per_sample, frames * (N_FFT // 2)
    h_in = torch.empty(n_in, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n_out, dtype=torch.uint8).pin_memory()
    h_in.fill_(1)
    d_in = torch.empty(n_in, dtype=torch.uint8, device=dev)
    d_out = torch.zeros(n_out, dtype=torch.uint8, device=dev)
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def one(direction: str, pieces: int = 1) -> float:
        torch.cuda.synchronize(dev)
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            for p in range(pieces):   # pieces > 1: the two directions interleaved in chunks, as a pipelined caller issues them
                if direction in ("h2d", "both"):
                    lo, hi = n_in * p // pieces, n_in * (p + 1) // pieces
                    with torch.cuda.stream(s_in):
                        d_in[lo:hi].copy_(h_in[lo:hi], non_blocking=True)
                if direction in ("d2h", "both"):
                    lo, hi = n_out * p // pieces, n_out * (p + 1) // pieces
                    with torch.cuda.stream(s_out):
                        h_out[lo:hi].copy_(d_out[lo:hi], non_blocking=True)
        torch.cuda.synchronize(dev)
        dt = (time.perf_counter() - t0) / steps
        barrier()
        return dt

    one("both")  # warm-up
    # the ceiling is the fastest way of moving both directions' bytes: whole buffers at once, or interleaved pieces
    schedules = {"whole": one("both"), "32MB_pieces": one("both", max(1, n_in >> 25)), "128MB_pieces": one("both", max(1, n_in >> 27))}
    return {"frames": frames, "h2d_bytes": n_in, "d2h_bytes": n_out, "t_h2d": one("h2d"), "t_d2h": one("d2h"),
            "t_both": min(schedules.values()), "schedules_ms": {k: v * 1e3 for k, v in schedules.items()}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=512)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--bytes-per-sample", type=int, default=4)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    r = measure(dev, args.clips, args.steps, args.bytes_per_sample, (dist.barrier if dist else (lambda: None)))
    t = torch.tensor([r["t_h2d"], r["t_d2h"], r["t_both"]], dtype=torch.float64, device=dev)
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t_h2d, t_d2h, t_both = (float(v) for v in t.tolist())
    if rank == 0:
        print(json.dumps({
            "tool": "host_path_ceiling", "n_gpus": world, "clips_per_gpu": args.clips, "bytes_per_sample": args.bytes_per_sample,
            "h2d_gbs_per_gpu": r["h2d_bytes"] / t_h2d / 1e9, "d2h_gbs_per_gpu": r["d2h_bytes"] / t_d2h / 1e9,
            "both_ms": t_both * 1e3, "both_ms_by_schedule_rank0": r["schedules_ms"], "h2d_gbs_per_gpu_concurrent": r["h2d_bytes"] / t_both / 1e9,
            "d2h_gbs_per_gpu_concurrent": r["d2h_bytes"] / t_both / 1e9,
            "ceiling_frames_per_s": world * r["frames"] / t_both,
            "aggregate_host_gbs": world * (r["h2d_bytes"] + r["d2h_bytes"]) / t_both / 1e9,
        }), flush=True)
    if dist:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
