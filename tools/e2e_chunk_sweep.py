#!/usr/bin/env python
"""End-to-end time of the host-buffer entry points against the pipeline's chunk size (SG_CHUNK_MB, read once per process):
float32 clips through sg_stft_batch and the same clips as 16-bit PCM through sg_stft_pcm, pinned host arrays."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spectrogram_b200 as sg  # noqa: E402
from spectrogram_b200 import _lib as L  # noqa: E402

clips, clip_len = 512, 441000
eng = sg.Engine(0)
opts = sg.Options()
frames = eng.num_frames(opts, clip_len)
pin_in = sg.PinnedArray((clips, clip_len), np.float32)
pin_s16 = sg.PinnedArray((clips, clip_len), np.int16)
pin_out = sg.PinnedArray((clips, frames, 1024), np.uint8)
pin_in.array[...] = (0.2 * np.random.default_rng(0).standard_normal((clips, clip_len))).astype(np.float32)
pin_s16.array[...] = np.clip(np.rint(pin_in.array * 32767.0), -32768, 32767).astype(np.int16)
cfg, _k = opts.to_c()
info = L.PcmInfo(L.PCM_S16, 1, 44100, clip_len, 0)
lib = L.load()


def t(fn):
    for _ in range(2):
        fn()
    t0 = time.perf_counter()
    for _ in range(5):
        fn()
    return (time.perf_counter() - t0) / 5


a = t(lambda: eng.spectrogram(pin_in.array, opts, out=pin_out.array))
b = t(lambda: L.check(lib.sg_stft_pcm(eng.handle, pin_s16.array.ctypes.data, clips, C.byref(info), L.PCM_MONO_MIX, C.byref(cfg),
                                      pin_out.array.ctypes.data)))
print(f"chunk {os.environ.get('SG_CHUNK_MB', 'default')} MB: float32 {a * 1e3:.2f} ms ({clips * frames / a / 1e6:.1f} M frames/s), "
      f"16-bit PCM {b * 1e3:.2f} ms ({clips * frames / b / 1e6:.1f} M frames/s)", flush=True)
