#!/usr/bin/env python
"""bench.py -- headline benchmark of the frame-producing hot path (BASELINE.json metric:
STFT frames/s & audio-sec/s at n_fft 2048 / hop 512, HBM GB/s fraction).

A "step" is one pass of the fused window -> real FFT -> |X|^2 -> dB -> byte kernel over one batch of
synthetic clips (config 1's shape tiled to a batch): CLIPS_PER_GPU clips x 10 s x
44.1 kHz, n_fft 2048, hop 512, Blackman (the AnalyserNode window), u8 output.

  python bench.py [--gpus N] [--steps K] [--warmup W]          this repo's CUDA path
  python bench.py --impl reference ...                          the CPU arm (oracle port, all host threads)
  torchrun ... bench.py --gpus N ...                            one rank per GPU, clips sharded, no collective
                                                               on the data path (shards are independent)
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_FFT, HOP, SR = 2048, 512, 44100
CLIP_LEN = 441000            # 10 s at 44.1 kHz (BASELINE config 1)
CLIPS_PER_GPU = 512          # 903 MB of float32 input per step: far larger than the 126 MB L2
BYTES_PER_FRAME = 4 * HOP + 1 * (N_FFT // 2)   # algorithmic: each input sample once + one byte per bin = 3072
FALLBACK_HBM_GBS = 6650.0    # B200_PROFILING.md fallback, used only if MEASURED_PEAKS.json is absent


# ------------------------------------------------------------------------------------------------
# multi-rank plumbing (also exercised on CPU/gloo by tests/test_host_logic.py)
# ------------------------------------------------------------------------------------------------
def shard_plan(total_clips: int, world: int, rank: int) -> dict:
    """Contiguous block of clips for this rank: clip i -> rank floor(i*world/total)."""
    lo, hi = total_clips * rank // world, total_clips * (rank + 1) // world
    return {"lo": lo, "hi": hi, "n_clips": hi - lo}


def _dist():
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None


def max_over_ranks(value: float, device="cuda") -> float:
    import torch
    d = _dist()
    if d is None:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    d.all_reduce(t, op=d.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device="cuda") -> float:
    import torch
    d = _dist()
    if d is None:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    d.all_reduce(t, op=d.ReduceOp.SUM)
    return float(t.item())


def barrier():
    d = _dist()
    if d is not None:
        d.barrier()


# ------------------------------------------------------------------------------------------------
# clocks: sampled DURING the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nv = None

    def _run(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.005)

    def start(self):
        if self._nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self) -> dict | None:
        if self._thread is None:
            return None
        self._stop.set()
        self._thread.join()
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (oracle/analyser_ref.c) on the host cores
# ------------------------------------------------------------------------------------------------
def synth_clips_cpu(n_clips: int, first_clip: int = 0):
    """Same family as the GPU batch: linear chirp 20 Hz -> 20 kHz, amplitude 0.5, plus per-clip phase."""
    import numpy as np
    t = np.arange(CLIP_LEN, dtype=np.float64) / SR
    dur = CLIP_LEN / SR
    out = np.empty((n_clips, CLIP_LEN), dtype=np.float32)
    for c in range(n_clips):
        ph = 2 * np.pi * (20.0 * t + 0.5 * (20000.0 - 20.0) / dur * t * t) + 0.37 * (first_clip + c)
        out[c] = (0.5 * np.sin(ph)).astype(np.float32)
    return out


def cpu_arm(n_clips: int, steps: int, warmup: int) -> dict:
    """Times the oracle's C port (all host threads) on a bounded sample of the same workload."""
    from oracle import analyser_oracle as O
    from oracle import cref
    cores = cref.max_threads()
    x = synth_clips_cpu(n_clips)
    cfg = O.Config(n_fft=N_FFT, hop=HOP, window=O.WINDOW_BLACKMAN, output=O.OUT_U8)
    frames = O.num_frames(CLIP_LEN, N_FFT, HOP, O.ALIGN_VALID) * n_clips
    for _ in range(warmup):
        cref.stft_batch(x, cfg, 0)
    t0 = time.perf_counter()
    for _ in range(steps):
        cref.stft_batch(x, cfg, 0)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return {"value": frames / dt, "unit": "frames/s", "cores": cores, "kind": "port",
            "sample": f"{n_clips} clips x 10 s x 44.1 kHz ({frames} frames) per step, {steps} steps, "
                      f"oracle/analyser_ref.c float32, {cores} threads", "ms_per_step": dt * 1e3, "frames": frames}


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # the CPU arm runs once per box
    cores = os.cpu_count() or 1
    n_clips = max(4, min(64, cores))          # bounded: ~0.05 s of single-thread work per clip
    r = cpu_arm(n_clips, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "stft_frames_per_s", "value": r["value"], "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "audio_s_per_s": r["value"] * HOP / SR,
        "config": workload_config(n_clips, note="bounded sample of the same workload on the host cores"),
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(clips_per_gpu: int, note: str = "") -> dict:
    cfg = {
        "workload": "config 1 tiled: batch of 10 s mono 44.1 kHz chirps, n_fft 2048, hop 512, u8 dB bytes",
        "n_fft": N_FFT, "hop": HOP, "window": "blackman (AnalyserNode parity window)", "output": "u8",
        "min_db": -100, "max_db": -30, "smoothing": 0.0, "clips_per_gpu": clips_per_gpu, "clip_len": CLIP_LEN,
        "frames_per_clip": 1 + (CLIP_LEN - N_FFT) // HOP, "sharding": "contiguous clip blocks per rank, no collective",
        "l2": f"inputs larger than L2 ({clips_per_gpu * CLIP_LEN * 4 / 1e6:.0f} MB read per step, no flush needed)",
    }
    if note:
        cfg["note"] = note
    return cfg


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def synth_clips_gpu(torch, n_clips: int, first_clip: int, device):
    t = torch.arange(CLIP_LEN, dtype=torch.float64, device=device) / SR
    dur = CLIP_LEN / SR
    base = 2 * torch.pi * (20.0 * t + 0.5 * (20000.0 - 20.0) / dur * t * t)
    x = torch.empty((n_clips, CLIP_LEN), dtype=torch.float32, device=device)
    for c0 in range(0, n_clips, 32):
        c1 = min(n_clips, c0 + 32)
        ph = 0.37 * torch.arange(first_clip + c0, first_clip + c1, dtype=torch.float64, device=device)
        x[c0:c1] = (0.5 * torch.sin(base[None, :] + ph[:, None])).to(torch.float32)
    return x


def bind_near_gpu(index: int) -> str:
    """Pin this rank's host threads to the cores closest to its GPU (NVML's ideal CPU affinity), so the pinned
    staging buffers it allocates next are first-touched on that NUMA node.  Matters for the end-to-end number at
    4-8 GPUs, where every rank streams ~1.3 GB per step over PCIe."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return f"{len(os.sched_getaffinity(0))} cores"
    except Exception as e:  # no NVML / not permitted: run unbound
        return f"unbound ({type(e).__name__})"


def fp32_bound(clocks: dict, frames_per_s_per_gpu: float) -> dict:
    """The kernel's real ceiling: one frame pair costs 2430 FP32-pipe cycles on one of the 4 x 148
    schedulers (1064 packed FFMA2/FADD2 ops at 2 cycles + 300 scalar), counted from the SASS of
    stft_w32x2p_kernel and confirmed by tools/microbench/pipe_bench.cu (2.05 cycles per packed op)."""
    mhz = float(clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965.0)
    cap = 148 * 4 * mhz * 1e6 / 2430.0 * 2.0
    return {"cycles_per_frame_pair": 2430, "cap_frames_per_s": cap, "frac_of_cap": frames_per_s_per_gpu / cap,
            "cap_as_hbm_frac": cap * BYTES_PER_FRAME / 1e9 / 6551.4}


def run_gpu(args) -> None:
    import numpy as np
    import torch

    import spectrogram_b200 as sg
    from spectrogram_b200 import _lib

    _lib.load()  # fails loudly if the CUDA library has not been built
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (this engine has no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    affinity = bind_near_gpu(local) if world > 1 else "all cores"
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    clips = args.clips_per_gpu
    plan = shard_plan(clips * world, world, rank)
    eng = sg.Engine(local)
    opts = sg.Options(fftSize=N_FFT, hop=HOP, window="blackman", output="u8")
    frames_per_clip = eng.num_frames(opts, CLIP_LEN)
    frames = frames_per_clip * plan["n_clips"]

    x = synth_clips_gpu(torch, plan["n_clips"], plan["lo"], dev)
    out = torch.empty((plan["n_clips"], frames_per_clip, N_FFT // 2), dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    stream = torch.cuda.Stream(device=dev)   # the kernels and the timing events share this stream

    def step():
        eng.spectrogram_device(x.data_ptr(), plan["n_clips"], CLIP_LEN, CLIP_LEN, opts, out.data_ptr(), stream.cuda_stream)

    # ---- device-resident timing: W warm-up steps, K timed steps between barriers + synchronize
    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = eng.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    ev1.record(stream)
    torch.cuda.synchronize()
    barrier()
    ms_local = ev0.elapsed_time(ev1) / args.steps
    launches = eng.launch_count - launches0
    clocks = sampler.stop()
    ms = max_over_ranks(ms_local, dev)
    total_frames = sum_over_ranks(frames, dev)
    value = total_frames / (ms * 1e-3)

    # ---- parity spot check of what was just timed (rank 0, first clip) against the oracle
    parity = None
    if rank == 0:
        from oracle import analyser_oracle as O
        ref = O.spectrogram(x[0].cpu().numpy(), O.Config())[0]
        got = out[0].cpu().numpy()
        d = np.abs(ref.astype(np.int32) - got.astype(np.int32))
        parity = {"max_lsb": int(d.max()), "mismatch_frac": float((d != 0).mean()), "frames_checked": int(ref.shape[0])}
        if d.max() > 1:
            raise SystemExit(f"parity failure in the timed kernel: {parity}")

    # ---- the same kernel with float dB output (getFloatFrequencyData rows, 6144 algorithmic bytes per frame):
    #      secondary figure, rank 0 at N = 1 only; the headline above stays the byte output BASELINE.json names
    float_db = None
    if world == 1:
        opts_db = sg.Options(fftSize=N_FFT, hop=HOP, window="blackman", output="db")
        n_db = min(plan["n_clips"], 512)
        out_db = torch.empty((n_db, frames_per_clip, N_FFT // 2), dtype=torch.float32, device=dev)

        def step_db():
            eng.spectrogram_device(x.data_ptr(), n_db, CLIP_LEN, CLIP_LEN, opts_db, out_db.data_ptr(), stream.cuda_stream)

        for _ in range(3):
            step_db()
        torch.cuda.synchronize()
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d0.record(stream)
        for _ in range(args.steps):
            step_db()
        d1.record(stream)
        torch.cuda.synchronize()
        ms_db = d0.elapsed_time(d1) / args.steps
        fps_db = n_db * frames_per_clip / (ms_db * 1e-3)
        float_db = {"value": fps_db, "unit": "frames/s", "ms_per_step": ms_db, "bytes_per_frame": 4 * HOP + 4 * (N_FFT // 2),
                    "achieved_gbs": fps_db * (4 * HOP + 4 * (N_FFT // 2)) / 1e9, "kernel": eng.last_kernel, "clips": n_db}
        del out_db
        step()                      # leave the engine on the byte kernel (last_kernel, parity check below)
        torch.cuda.synchronize()

    # ---- end to end through the public host API: pinned host buffers, H2D + kernel + D2H every step
    e2e = None
    e2e_clips = min(plan["n_clips"], args.e2e_clips)
    if e2e_clips > 0:
        pin_in = sg.PinnedArray((e2e_clips, CLIP_LEN), np.float32)
        pin_out = sg.PinnedArray((e2e_clips, frames_per_clip, N_FFT // 2), np.uint8)
        pin_in.array[...] = x[:e2e_clips].cpu().numpy()
        for _ in range(2):
            eng.spectrogram(pin_in.array, opts, out=pin_out.array)
        barrier()
        e2e_steps = max(2, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            eng.spectrogram(pin_in.array, opts, out=pin_out.array)   # synchronous: returns when out is filled
        dt_local = (time.perf_counter() - t0) / e2e_steps
        barrier()
        dt = max_over_ranks(dt_local, dev)
        e2e_frames = sum_over_ranks(e2e_clips * frames_per_clip, dev)
        if rank == 0:
            assert np.array_equal(pin_out.array[0], out[0].cpu().numpy()), "host-API result differs from device-API result"
        e2e = {"value": e2e_frames / dt, "unit": "frames/s", "h2d_bytes_per_step": int(world * e2e_clips * CLIP_LEN * 4),
               "d2h_bytes_per_step": int(world * e2e_clips * frames_per_clip * (N_FFT // 2)), "ms_per_step": dt * 1e3,
               "api": "spectrogram_b200.Engine.spectrogram -> sg_stft_batch (pinned host in/out)",
               "host_affinity": affinity}
        # secondary: the same clips as 16-bit PCM through sg_stft_pcm (ingest on the GPU, 2 bytes per sample over PCIe)
        if world == 1:
            from spectrogram_b200 import _lib as L
            import ctypes as C
            pin_s16 = sg.PinnedArray((e2e_clips, CLIP_LEN), np.int16)
            pin_s16.array[...] = np.clip(np.rint(pin_in.array * 32767.0), -32768, 32767).astype(np.int16)
            cfg_c, _keep = opts.to_c()
            info = L.PcmInfo(L.PCM_S16, 1, int(SR), CLIP_LEN, 0)
            lib = L.load()

            def pcm_step():
                L.check(lib.sg_stft_pcm(eng.handle, pin_s16.array.ctypes.data, e2e_clips, C.byref(info), L.PCM_MONO_MIX,
                                        C.byref(cfg_c), pin_out.array.ctypes.data))
            for _ in range(2):
                pcm_step()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                pcm_step()
            dt16 = (time.perf_counter() - t0) / e2e_steps
            e2e["pcm16_input"] = {"value": e2e_clips * frames_per_clip / dt16, "unit": "frames/s", "ms_per_step": dt16 * 1e3,
                                  "h2d_bytes_per_step": int(e2e_clips * CLIP_LEN * 2),
                                  "api": "sg_stft_pcm (16-bit mono PCM, pinned host in/out)"}
            pin_s16.free()
        pin_in.free()
        pin_out.free()

    if rank == 0:
        peak, peak_src = FALLBACK_HBM_GBS, "fallback"
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peak, peak_src = float(json.load(f)["hbm_gbs"]), "measured"
        except Exception:
            pass
        achieved = frames * BYTES_PER_FRAME / (ms_local * 1e-3) / 1e9   # this rank's kernel: bytes per launch / launch time
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
                tr = json.load(f)
                traffic = tr["dram_bytes_per_frame"] * frames
        except Exception:
            pass
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            r = cpu_arm(max(4, min(64, cores)), steps=2, warmup=1)
            cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
        line = {
            "metric": "stft_frames_per_s", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "audio_s_per_s": value * HOP / SR,
            "config": workload_config(clips),
            "kernel": eng.last_kernel,
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "bytes_per_frame": BYTES_PER_FRAME,
                         "frames_per_launch": frames,
                         "fp32_pipe": fp32_bound(clocks, value / max(world, 1)),
                         "note": "co-bound by the FP32 pipe: 2430 FP32-pipe cycles per frame pair per scheduler cap the kernel at 43% of HBM peak (DESIGN.md)"},
            "cpu_baseline": cpu,
            "parity": parity,
        }
        if float_db is not None:
            float_db["frac_of_hbm_peak"] = float_db["achieved_gbs"] / peak
            line["float_db_output"] = float_db
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--clips-per-gpu", type=int, default=CLIPS_PER_GPU)
    ap.add_argument("--e2e-clips", type=int, default=CLIPS_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
