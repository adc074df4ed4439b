#!/usr/bin/env python
"""How much of each parity tolerance (tests/_tol.py) the CUDA kernels actually use.

For every kernel family and the BASELINE shapes it serves, against the float64 oracle on seeded inputs:
  mag_floor   max |d| / frame_peak                      (the absolute floor a float32 FFT has; _tol.MAG_FLOOR bounds it)
  mag_rel60   max |d| / |ref| over bins within 60 dB of their frame's peak   (_tol.MAG_REL bounds it near the peak)
  use@F       max |d| / (1e-4 |ref| + F peak)  for F = 5e-7, 2e-7, 1e-7: the fraction of that tolerance in use
  dB@W        max dB error over bins within W dB of the frame peak, W = 50, 60, 70, 80
  byte        mismatch rate against the oracle's bytes and the largest difference in LSB
Prints a table and writes JSON lines.  usage: python tools/tolerance_report.py [out.jsonl]
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spectrogram_b200 as sg  # noqa: E402
from oracle import analyser_oracle as O  # noqa: E402

WIN = {"blackman": O.WINDOW_BLACKMAN, "hann": O.WINDOW_HANN}


def signals(sr, n, seed):
    rng = np.random.default_rng(seed)
    return {
        "chirp": O.chirp(n, float(sr), 20.0, 0.45 * sr, 0.5),
        "noise": (0.1 * rng.standard_normal(n)).astype(np.float32),
        "band": O.band_noise(n, float(sr), 300.0, 3400.0, 0.1, seed),
        "tone+noise": (0.5 * np.sin(2 * np.pi * 1000.0 * np.arange(n) / sr) + 1e-3 * rng.standard_normal(n)).astype(np.float32),
    }


CASES = [
    # name, n_fft, hop, window, sample rate, kernel variant, tau
    ("2048/512 pair (headline)", 2048, 512, "blackman", 44100, 0, 0.0),
    ("2048/512 pair, TMA staged", 2048, 512, "blackman", 44100, 4, 0.0),
    ("2048/441 pair, unaligned span", 2048, 441, "blackman", 44100, 0, 0.0),
    ("2048/512 one frame per warp", 2048, 512, "blackman", 44100, 2, 0.0),
    ("2048/512 register family", 2048, 512, "blackman", 44100, 3, 0.0),
    ("2048/512 generic smem", 2048, 512, "blackman", 44100, 1, 0.0),
    ("2048/512 fused smoothing 0.8", 2048, 512, "blackman", 44100, 7, 0.8),
    ("2048/160 two-kernel smoothing 0.8", 2048, 160, "blackman", 44100, 0, 0.8),
    ("1024/256 pair L=16", 1024, 256, "blackman", 48000, 0, 0.0),
    ("1024/256 fused smoothing 0.8", 1024, 256, "blackman", 48000, 7, 0.8),
    ("512/128 fused smoothing 0.8", 512, 128, "blackman", 48000, 7, 0.8),
    ("256/64 fused smoothing 0.8", 256, 64, "blackman", 48000, 7, 0.8),
    ("1024/128 pair L=16 (config 5)", 1024, 128, "blackman", 48000, 0, 0.0),
    ("512/160 pair L=8 (config 2)", 512, 160, "hann", 16000, 0, 0.0),
    ("512/100 16x16", 512, 100, "hann", 16000, 0, 0.0),
    ("256/64 pair L=4", 256, 64, "blackman", 48000, 0, 0.0),
    ("256/50 16x8", 256, 50, "blackman", 48000, 0, 0.0),
    ("400/160 r400 (config 4)", 400, 160, "hann", 16000, 0, 0.0),
    ("4096/1024 even/odd", 4096, 1024, "blackman", 48000, 0, 0.0),
    ("8192/2048 register family", 8192, 2048, "blackman", 48000, 0, 0.0),
]


def measure(eng, name, n_fft, hop, window, sr, variant, tau):
    n = max(4 * sr, 40 * n_fft)
    rows = []
    for sig_name, x in signals(sr, n, 1234).items():
        cfg = O.Config(n_fft=n_fft, hop=hop, window=WIN[window], smoothing=tau, output=O.OUT_F32_MAG)
        ref = O.spectrogram(x, cfg)[0]
        eng.set_kernel_variant(variant)
        try:
            kw = dict(fftSize=n_fft, hop=hop, window=window, smoothingTimeConstant=tau)
            mag = eng.spectrogram(x, sg.Options(output="mag", **kw)).astype(np.float64)
            kernel = eng.last_kernel
            db = eng.spectrogram(x, sg.Options(output="db", **kw)).astype(np.float64)
            by = eng.spectrogram(x, sg.Options(output="u8", **kw))
        finally:
            eng.set_kernel_variant(0)
        peak = ref.max(axis=-1, keepdims=True)
        live = peak[:, 0] > 0
        ref, mag, db, by, peak = ref[live], mag[live], db[live], by[live], peak[live]
        d = np.abs(mag - ref)
        with np.errstate(divide="ignore", invalid="ignore"):
            rel_db = 20 * np.log10(ref / peak)
            ref_db = 20 * np.log10(ref)
        row = {"case": name, "kernel": kernel, "signal": sig_name, "frames": int(ref.shape[0]),
               "mag_floor": float((d / peak).max()),
               "mag_rel60": float((d / np.maximum(ref, 1e-300))[rel_db >= -60].max())}
        for f in (5e-7, 2e-7, 1e-7):
            row[f"use@{f:g}"] = float((d / (1e-4 * ref + f * peak)).max())
        err_db = np.abs(db - ref_db)
        for w in (50, 60, 70, 80):
            sel = (ref > 0) & (rel_db >= -w)
            row[f"dB@{w}"] = float(err_db[sel].max())
        ref_by = O.finish(ref, O.Config(n_fft=n_fft, hop=hop, window=WIN[window], smoothing=tau))
        db_ = np.abs(by.astype(np.int32) - ref_by.astype(np.int32))
        row["byte_mismatch"] = float((db_ != 0).mean())
        row["byte_max_lsb"] = int(db_.max())
        rows.append(row)
    return rows


def main():
    eng = sg.Engine(0)
    out = open(sys.argv[1], "w") if len(sys.argv) > 1 else None
    cols = ["mag_floor", "mag_rel60", "use@5e-07", "use@2e-07", "use@1e-07", "dB@50", "dB@60", "dB@70", "dB@80", "byte_mismatch"]
    print(f"{'case':38s} {'kernel':13s} {'signal':10s} " + " ".join(f"{c:>10s}" for c in cols) + " lsb")
    worst = {}
    for case in CASES:
        for row in measure(eng, *case):
            if out:
                out.write(json.dumps(row) + "\n")
            print(f"{row['case']:38s} {row['kernel']:13s} {row['signal']:10s} " + " ".join(f"{row[c]:10.3g}" for c in cols) +
                  f" {row['byte_max_lsb']:3d}", flush=True)
            for c in cols + ["byte_max_lsb"]:
                worst[c] = max(worst.get(c, 0), row[c])
    print("worst over all cases: " + ", ".join(f"{c} {worst[c]:.3g}" for c in cols + ["byte_max_lsb"]))
    if out:
        out.write(json.dumps({"case": "WORST", **worst}) + "\n")
        out.close()
    eng.close()


if __name__ == "__main__":
    main()
