#!/usr/bin/env python
"""Latency of the reference's own call pattern (3D/visualizer.js:346-368): push one render quantum, then
getByteFrequencyData into a caller-owned Uint8Array.  usage: python tools/analyser_latency.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spectrogram_b200 as sg  # noqa: E402

eng = sg.Engine(0)
for tau in (0.0, 0.8):
    an = sg.AnalyserNode(eng)
    an.fftSize = 2048
    an.smoothingTimeConstant = tau
    x = (0.1 * np.random.default_rng(0).standard_normal(128)).astype(np.float32)
    buf = np.zeros(an.frequencyBinCount, np.uint8)
    for _ in range(100):
        an.push(x)
        an.getByteFrequencyData(buf)
    lat = []
    for _ in range(3000):
        t0 = time.perf_counter()
        an.push(x)
        an.getByteFrequencyData(buf)
        lat.append((time.perf_counter() - t0) * 1e6)
    lat = np.sort(np.array(lat))
    print(f"tau {tau}: push(128) + getByteFrequencyData p50 {lat[len(lat) // 2]:.1f} us  p99 {lat[int(len(lat) * 0.99)]:.1f} us "
          f"(one display frame at 60 Hz is 16 667 us)  launches/call {eng.launch_count / 3100:.1f}")
    an.close()
eng.close()
