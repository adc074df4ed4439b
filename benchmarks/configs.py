#!/usr/bin/env python
"""Secondary measurements: the other BASELINE.json configs (2-5), device-resident, CUDA-event timed.
bench.py is the contract benchmark (config 1's shape); this script reports where the remaining shapes
stand.  One JSON line per config.  usage: python benchmarks/configs.py [--quick]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spectrogram_b200 as sg  # noqa: E402

PEAK = 6551.4
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def time_device(eng, x, n_clips, clip_len, opts, out, steps=10, warmup=3):
    st = torch.cuda.Stream()
    torch.cuda.synchronize()
    for _ in range(warmup):
        eng.spectrogram_device(x.data_ptr(), n_clips, clip_len, clip_len, opts, out.data_ptr(), st.cuda_stream)
    st.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(steps):
        eng.spectrogram_device(x.data_ptr(), n_clips, clip_len, clip_len, opts, out.data_ptr(), st.cuda_stream)
    e1.record(st)
    st.synchronize()
    return e0.elapsed_time(e1) / steps


def report(name, eng, n_clips, clip_len, opts, sr, elem_bytes, torch_dtype, extra=None):
    frames = eng.num_frames(opts, clip_len)
    bins = opts.fftSize // 2
    g = torch.Generator(device="cuda").manual_seed(1234)
    x = (torch.randn((n_clips, clip_len), device="cuda", generator=g) * 0.1).float()
    out = torch.empty((n_clips, frames, bins), dtype=torch_dtype, device="cuda")
    ms = time_device(eng, x, n_clips, clip_len, opts, out)
    total = n_clips * frames
    bpf = 4 * opts.hop + elem_bytes * bins
    line = {"config": name, "n_fft": opts.fftSize, "hop": opts.hop, "output": opts.output, "tau": opts.smoothingTimeConstant,
            "clips": n_clips, "frames": total, "ms": ms, "frames_per_s": total / ms * 1e3,
            "audio_s_per_s": total * opts.hop / sr / ms * 1e3, "bytes_per_frame": bpf,
            "hbm_frac": total * bpf / ms / 1e6 / PEAK, "kernel": eng.last_kernel}
    if extra:
        line.update(extra)
    print(json.dumps(line), flush=True)
    del x, out
    torch.cuda.empty_cache()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    eng = sg.Engine(0)
    q = args.quick
    # config 2: 1 h mono 16 kHz, n_fft 512, hop 160, float dB
    report("2: 1h 16kHz n512 hop160 dB", eng, 1, 16000 * (600 if q else 3600), sg.Options(fftSize=512, hop=160, window="hann", output="db"), 16000, 4, torch.float32)
    # config 3: sweep on 60 s stereo 48 kHz with tau 0.8 (channels as clips)
    for n in (256, 512, 1024, 2048, 4096, 8192):
        report(f"3: sweep n{n} tau0.8 u8", eng, 2, 48000 * 60, sg.Options(fftSize=n, hop=n // 4, output="u8", smoothingTimeConstant=0.8, align="analyser"), 48000, 1, torch.uint8)
    for n in (256, 512, 1024, 2048, 4096, 8192):
        report(f"3b: sweep n{n} tau0 u8 x64 clips", eng, 64, 48000 * 60, sg.Options(fftSize=n, hop=n // 4, output="u8"), 48000, 1, torch.uint8)
    # config 4: 4096 x 30 s x 16 kHz, n_fft 400, hop 160
    report("4: 4096x30s n400 hop160 dB", eng, 512 if q else 4096, 480000, sg.Options(fftSize=400, hop=160, window="hann", output="db"), 16000, 4, torch.float32)
    # config 5: streaming 256 channels x 48 kHz, n_fft 1024, hop 128, chunk = one render quantum
    opts = sg.Options(fftSize=1024, hop=128, output="u8")
    bank = sg.StreamBank(256, opts, max_chunk=128, engine=eng)
    chunk = sg.PinnedArray((256, 128), np.float32)
    chunk.array[...] = (0.1 * np.random.default_rng(0).standard_normal((256, 128))).astype(np.float32)
    out = sg.PinnedArray((256, 1, 512), np.uint8)
    rgba = sg.PinnedArray((256, 1, 512, 4), np.uint8)
    for _ in range(50):
        bank.push(chunk.array, out=out.array, out_rgba=rgba.array)
    lat = []
    for _ in range(500 if q else 2000):
        t0 = time.perf_counter()
        bank.push(chunk.array, out=out.array, out_rgba=rgba.array)
        lat.append((time.perf_counter() - t0) * 1e6)
    lat = np.sort(np.array(lat))
    print(json.dumps({"config": "5: streaming 256ch n1024 hop128 u8+rgba", "p50_us": float(lat[len(lat) // 2]),
                      "p99_us": float(lat[int(len(lat) * 0.99)]), "chunks_per_s": 1e6 / float(lat.mean()),
                      "realtime_factor": (128 / 48000) / (float(lat.mean()) * 1e-6), "kernel": eng.last_kernel}), flush=True)
    bank.close()
    eng.close()


if __name__ == "__main__":
    main()
