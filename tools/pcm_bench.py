#!/usr/bin/env python
"""PCM ingestion measurements on one B200:
  1. the ingest kernel alone, device resident (CUDA events), as algorithmic GB/s of the measured HBM peak
  2. the whole path from pinned host memory at the headline shape (n_fft 2048 / hop 512 / u8): float32 samples
     through sg_stft_batch against 16-bit samples through sg_stft_pcm (half the host->device bytes)
usage: python tools/pcm_bench.py [--clips 256]"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spectrogram_b200 as sg  # noqa: E402
from spectrogram_b200 import _lib as L  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--clips", type=int, default=256)
ap.add_argument("--steps", type=int, default=5)
args = ap.parse_args()
peak = 6551.4
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbps"]["burst"])
except Exception:
    pass

eng = sg.Engine(0)
lib = L.load()
st = torch.cuda.Stream()
print("-- ingest kernel alone (device resident)")
for fmt, name, sb in ((L.PCM_S16, "s16", 2), (L.PCM_S24, "s24", 3), (L.PCM_F32, "f32", 4)):
    for ch, layout in ((1, L.PCM_MONO_MIX), (2, L.PCM_MONO_MIX), (2, L.PCM_PLANAR), (6, L.PCM_MONO_MIX)):
        frames = (256 << 20) // (ch * sb)
        src = torch.randint(0, 255, (frames * ch * sb,), dtype=torch.uint8, device="cuda")
        planes = ch if layout == L.PCM_PLANAR else 1
        dst = torch.empty((planes, frames), dtype=torch.float32, device="cuda")
        info = L.PcmInfo(fmt, ch, 48000, frames, 0)
        def step():
            L.check(lib.sg_pcm_ingest_device(eng.handle, C.c_void_p(src.data_ptr()), 1, C.byref(info), layout,
                                             C.c_void_p(dst.data_ptr()), frames, C.c_void_p(st.cuda_stream)))
        for _ in range(3):
            step()
        st.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(10):
            step()
        e1.record(st)
        st.synchronize()
        ms = e0.elapsed_time(e1) / 10
        gb = frames * (ch * sb + 4 * planes) / ms / 1e6
        print(f"{name} x{ch} -> {planes} plane(s): {ms:.3f} ms  {gb:.0f} GB/s  ({gb / peak:.1%} of {peak:.0f})", flush=True)
        del src, dst

print("-- whole path from pinned host memory, n_fft 2048 / hop 512 / u8")
clip_len = 441000
opts = sg.Options()
fpc = eng.num_frames(opts, clip_len)
rng = np.random.default_rng(0)
pin_f = sg.PinnedArray((args.clips, clip_len), np.float32)
pin_s = sg.PinnedArray((args.clips, clip_len), np.int16)
pin_o = sg.PinnedArray((args.clips, fpc, 1024), np.uint8)
s16 = (rng.standard_normal((args.clips, clip_len)) * 4000).astype(np.int16)
pin_s.array[...] = s16
pin_f.array[...] = s16.astype(np.float32) / 32768.0
cfg, _keep = opts.to_c()
info = L.PcmInfo(L.PCM_S16, 1, 44100, clip_len, 0)
res = {}
for name, fn in (("f32 sg_stft_batch", lambda: L.check(lib.sg_stft_batch(eng.handle, pin_f.array.ctypes.data, args.clips, clip_len, C.byref(cfg), pin_o.array.ctypes.data))),
                 ("s16 sg_stft_pcm", lambda: L.check(lib.sg_stft_pcm(eng.handle, pin_s.array.ctypes.data, args.clips, C.byref(info), L.PCM_MONO_MIX, C.byref(cfg), pin_o.array.ctypes.data)))):
    fn(); fn()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fn()
    dt = (time.perf_counter() - t0) / args.steps
    res[name] = pin_o.array.copy()
    print(f"{name}: {dt * 1e3:.2f} ms  {args.clips * fpc / dt / 1e6:.2f} M frames/s", flush=True)
print("outputs identical:", bool(np.array_equal(res["f32 sg_stft_batch"], res["s16 sg_stft_pcm"])))
eng.close()
