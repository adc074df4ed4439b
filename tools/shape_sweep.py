#!/usr/bin/env python
"""Throughput over a grid of (n_fft, hop, output): finds shapes that fall off the fast paths.
usage: python tools/shape_sweep.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spectrogram_b200 as sg  # noqa: E402

eng = sg.Engine(0)
L = 480000
x = (torch.rand((64, L), device="cuda") - 0.5).float()
st = torch.cuda.Stream()
for n in (256, 400, 512, 1024, 2048, 4096):
    for hop in sorted({n // 8, n // 4, n // 2, n, 160, 441}):
        row = []
        for out, dt, eb in (("u8", torch.uint8, 1), ("db", torch.float32, 4)):
            opts = sg.Options(fftSize=n, hop=hop, output=out)
            fr = eng.num_frames(opts, L)
            o = torch.empty((64, fr, n // 2), dtype=dt, device="cuda")
            for _ in range(2):
                eng.spectrogram_device(x.data_ptr(), 64, L, L, opts, o.data_ptr(), st.cuda_stream)
            st.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(5):
                eng.spectrogram_device(x.data_ptr(), 64, L, L, opts, o.data_ptr(), st.cuda_stream)
            e1.record(st)
            st.synchronize()
            ms = e0.elapsed_time(e1) / 5
            fps = 64 * fr / ms * 1e3
            row.append(f"{out} {fps / 1e6:8.1f} M/s hbm {fps * (4 * hop + eb * n // 2) / 6551.4e9:5.3f} [{eng.last_kernel}]")
            del o
        print(f"n_fft {n:5d} hop {hop:5d} | " + " | ".join(row), flush=True)
eng.close()
