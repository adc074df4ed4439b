#!/usr/bin/env python
"""sg_stft_batch_multi over 1 .. G GPUs of one box through the Python mirror of the JS `devices` option: page-locked
caller arrays, second call timed, result compared bit for bit with the 1-GPU result.  usage: python tools/multi_api_bench.py"""
import os
import sys
import time

import ctypes as C

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spectrogram_b200 as sg  # noqa: E402
from spectrogram_b200 import _lib as L  # noqa: E402

n = sg.device_count()
clips, clip_len = 512, 441000
opts = sg.Options()
frames = sg.default_engine(0).num_frames(opts, clip_len)
pin_in = sg.PinnedArray((clips, clip_len), np.float32)
pin_out = sg.PinnedArray((clips, frames, 1024), np.uint8)
rng = np.random.default_rng(3)
pin_in.array[...] = (0.2 * rng.standard_normal((clips, clip_len))).astype(np.float32)
one = None
for g in [k for k in (1, 2, 4, 8) if k <= n]:
    engs = [sg.default_engine(d) for d in range(g)]
    best = 1e9
    for rep in range(4):
        t0 = time.perf_counter()
        cfg, _keep = opts.to_c()      # straight through the C ABI: results land in the caller's page-locked array
        handles = (C.c_void_p * g)(*[e.handle for e in engs])
        L.check(L.load().sg_stft_batch_multi(handles, g, pin_in.array.ctypes.data, clips, clip_len, C.byref(cfg), pin_out.array.ctypes.data))
        dt = time.perf_counter() - t0
        if rep:
            best = min(best, dt)
    if one is None:
        one = pin_out.array.copy()
    same = np.array_equal(pin_out.array, one)
    print(f"sg_stft_batch_multi, {g} GPU(s), pinned host arrays: {best * 1e3:.1f} ms  {clips * frames / best / 1e6:.1f} M frames/s  "
          f"({(pin_in.array.nbytes + pin_out.array.nbytes) / best / 1e9:.0f} GB/s over the host path)  identical to 1 GPU: {same}", flush=True)
pin_in.free(); pin_out.free()
