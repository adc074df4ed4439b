#!/usr/bin/env python
"""BASELINE config 4 at 1/2/4/8 GPUs: 4096 clips x 30 s x 16 kHz, n_fft 400, hop 160, float dB, clip-sharded in
contiguous blocks (clip i -> rank floor(i*G/4096)), device resident, no data-path collective.
STRONG scaling: the 4096 clips are fixed, each rank holds 4096/G of them (per-clip seeded noise, seed = clip index).
    python benchmarks/config4_sharded.py                                     # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \\
        benchmarks/config4_sharded.py
One JSON line from rank 0; time = max over ranks of the CUDA-event time of K steps between barriers."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spectrogram_b200 as sg  # noqa: E402
from bench import barrier, max_over_ranks, shard_plan, sum_over_ranks  # noqa: E402

N_CLIPS, CLIP_LEN, SR = 4096, 480000, 16000.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--clips", type=int, default=N_CLIPS)
    args = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    plan = shard_plan(args.clips, world, rank)
    eng = sg.Engine(local)
    opts = sg.Options(fftSize=400, hop=160, output="db")
    fpc = eng.num_frames(opts, CLIP_LEN)
    x = torch.empty((plan["n_clips"], CLIP_LEN), dtype=torch.float32, device=dev)
    g = torch.Generator(device=dev)
    for i in range(plan["n_clips"]):                       # per-clip seeded noise: the shard content does not depend on G
        g.manual_seed(plan["lo"] + i)
        x[i] = torch.randn(CLIP_LEN, device=dev, generator=g) * 0.1
    out = torch.empty((plan["n_clips"], fpc, 200), dtype=torch.float32, device=dev)
    st = torch.cuda.Stream(device=dev)

    def step():
        eng.spectrogram_device(x.data_ptr(), plan["n_clips"], CLIP_LEN, CLIP_LEN, opts, out.data_ptr(), st.cuda_stream)

    for _ in range(max(3, args.warmup)):
        step()
    torch.cuda.synchronize()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(args.steps):
        step()
    e1.record(st)
    torch.cuda.synchronize()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1) / args.steps, dev)
    frames = sum_over_ranks(plan["n_clips"] * fpc, dev)
    # a checksum of checksums: identical for every G when the shards reproduce the 1-GPU result bit for bit
    fin = torch.where(torch.isfinite(out), out, torch.zeros_like(out))
    chk = sum_over_ranks(float(fin.view(torch.int32).to(torch.int64).sum().item() % (1 << 40)), dev)
    if rank == 0:
        print(json.dumps({"config": "4: 4096x30s n400 hop160 dB, clip-sharded", "n_gpus": world, "scaling": "strong",
                          "clips": args.clips, "frames": int(frames), "ms": ms, "frames_per_s": frames / ms * 1e3,
                          "audio_s_per_s": frames * 160 / SR / ms * 1e3, "kernel": eng.last_kernel,
                          "hbm_frac_per_gpu": frames * 1440 / ms / 1e6 / 6551.4 / world, "checksum": int(chk) % (1 << 40)}), flush=True)
    eng.close()
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
