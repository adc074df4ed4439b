#!/usr/bin/env python
"""Developer A/B timing of the n_fft 2048 / hop 512 kernels (device-resident, CUDA events), with a byte
parity check of each variant against the oracle.  usage: python tools/kbench.py [variants...] [--out u8|db]"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spectrogram_b200 as sg  # noqa: E402
from oracle import analyser_oracle as O  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("variants", nargs="*", type=int, default=[0, 4])
ap.add_argument("--out", default="u8")
ap.add_argument("--clips", type=int, default=256)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--nfft", type=int, default=2048)
ap.add_argument("--hop", type=int, default=512)
ap.add_argument("--clip-len", type=int, default=441000)
ap.add_argument("--window", default="blackman")
ap.add_argument("--tau", type=float, default=0.0)
args = ap.parse_args()

eng = sg.Engine(0)
clip_len = args.clip_len
opts = sg.Options(fftSize=args.nfft, hop=args.hop, output=args.out, window=args.window, smoothingTimeConstant=args.tau)
fpc = eng.num_frames(opts, clip_len)
g = torch.Generator(device="cuda").manual_seed(1)
x = (torch.rand((args.clips, clip_len), device="cuda", generator=g) - 0.5).float()
x[0] = torch.from_numpy(O.chirp(clip_len, 44100.0, 20.0, 20000.0, 0.5)).cuda()
dt = torch.uint8 if args.out == "u8" else torch.float32
bins = args.nfft // 2
out = torch.empty((args.clips, fpc, bins), dtype=dt, device="cuda")
WIN = {"blackman": O.WINDOW_BLACKMAN, "hann": O.WINDOW_HANN, "rect": O.WINDOW_RECT}
ref = O.spectrogram(x[0].cpu().numpy(), O.Config(n_fft=args.nfft, hop=args.hop, window=WIN[args.window], smoothing=args.tau,
                                                 output=O.OUT_U8 if args.out == "u8" else O.OUT_F32_DB))[0]
st = torch.cuda.Stream()
bpf = 4 * args.hop + bins * (1 if args.out == "u8" else 4)
for v in args.variants:
    eng.set_kernel_variant(v)
    def step():
        eng.spectrogram_device(x.data_ptr(), args.clips, clip_len, clip_len, opts, out.data_ptr(), st.cuda_stream)
    for _ in range(3):
        step()
    st.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(args.steps):
        step()
    e1.record(st)
    st.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    got = out[0].cpu().numpy()
    if args.out == "u8":
        d = np.abs(got.astype(np.int32) - ref.astype(np.int32))
        par = f"max {d.max()} LSB, mismatch {float((d != 0).mean()):.2e}"
    else:
        # the tolerance of tests/_tol.py: 1e-3 dB on bins within 50 dB of their frame's peak
        near = np.isfinite(ref) & (ref >= ref.max(axis=-1, keepdims=True) - 50.0)
        par = f"max dB err within 50 dB of the frame peak {np.abs(got[near] - ref[near]).max():.2e}"
    fps = args.clips * fpc / ms * 1e3
    print(f"variant {v} [{eng.last_kernel}] {ms:.4f} ms  {fps / 1e6:.1f} Mframes/s  hbm {fps * bpf / 6551.4e9:.3f}  parity: {par}", flush=True)
eng.close()
