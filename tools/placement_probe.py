#!/usr/bin/env python
"""Does the placement of the caller's buffers in device memory move the headline kernel's rate?  The same 512-clip batch at
different offsets of the input and output tensors inside larger allocations (and in fresh allocations)."""
import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spectrogram_b200 as sg
eng = sg.Engine(0)
st = torch.cuda.Stream()
clips, clip_len = 512, 441000
opts = sg.Options()
frames = eng.num_frames(opts, clip_len)
n_in, n_out = clips * clip_len, clips * frames * 1024

def rate(xp, op, reps=20):
    for _ in range(3):
        eng.spectrogram_device(xp, clips, clip_len, clip_len, opts, op, st.cuda_stream)
    st.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps):
        eng.spectrogram_device(xp, clips, clip_len, clip_len, opts, op, st.cuda_stream)
    e1.record(st); st.synchronize()
    return clips * frames / (e0.elapsed_time(e1) / reps) / 1e3

pad = 64 << 20
big_in = torch.rand(n_in + pad // 4, device="cuda") - 0.5
big_out = torch.empty(n_out + pad, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
for off_in in (0, 1 << 12, 1 << 16, 1 << 20, 3 << 20, 17 << 20):
    row = []
    for off_out in (0, 1 << 12, 1 << 16, 1 << 20, 5 << 20, 33 << 20):
        row.append(rate(big_in.data_ptr() + off_in, big_out.data_ptr() + off_out))
    print(f"input +{off_in:>9d} B: " + "  ".join(f"{r:6.1f}" for r in row) + "  M frames/s (output offsets 0, 4K, 64K, 1M, 5M, 33M)", flush=True)
