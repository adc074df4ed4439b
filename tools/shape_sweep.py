#!/usr/bin/env python
"""Throughput over a grid of (n_fft, hop, output): finds shapes that fall off the fast paths.  64 clips x 10 s for the
tau = 0 columns (u8, float dB); a third column runs u8 with smoothingTimeConstant 0.8 on 160 clips (enough clips for the
one-pass smoothing kernels to be selected where they exist).
usage: python tools/shape_sweep.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spectrogram_b200 as sg  # noqa: E402

eng = sg.Engine(0)
L = 480000
x = (torch.rand((160, L), device="cuda") - 0.5).float()
st = torch.cuda.Stream()
for n in (256, 400, 512, 1024, 2048, 4096, 8192):
    for hop in sorted({n // 8, n // 4, n // 2, n, 160, 441}):
        row = []
        for out, dt, eb, tau, nc in (("u8", torch.uint8, 1, 0.0, 64), ("db", torch.float32, 4, 0.0, 64), ("u8", torch.uint8, 1, 0.8, 160)):
            opts = sg.Options(fftSize=n, hop=hop, output=out, smoothingTimeConstant=tau)
            fr = eng.num_frames(opts, L)
            o = torch.empty((nc, fr, n // 2), dtype=dt, device="cuda")
            for _ in range(2):
                eng.spectrogram_device(x.data_ptr(), nc, L, L, opts, o.data_ptr(), st.cuda_stream)
            st.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(5):
                eng.spectrogram_device(x.data_ptr(), nc, L, L, opts, o.data_ptr(), st.cuda_stream)
            e1.record(st)
            st.synchronize()
            ms = e0.elapsed_time(e1) / 5
            fps = nc * fr / ms * 1e3
            tag = out if tau == 0 else f"{out} tau {tau}"
            row.append(f"{tag} {fps / 1e6:8.1f} M/s hbm {fps * (4 * hop + eb * n // 2) / 6551.4e9:5.3f} [{eng.last_kernel}]")
            del o
        print(f"n_fft {n:5d} hop {hop:5d} | " + " | ".join(row), flush=True)
eng.close()
