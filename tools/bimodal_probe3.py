"""Bisect which earlier bench entry makes the later 64-clip warm-up-mode call slow."""
import sys, torch
sys.path.insert(0, '/root/repo')
import spectrogram_b200 as sg
which = sys.argv[1]
eng = sg.Engine(0)
dev = torch.device('cuda', 0)
st = torch.cuda.Stream()
gen = torch.Generator(device=dev).manual_seed(1234)

def timed(n_fft, n_clips, clip_len, tau, reps=20, hop=None, out="u8", align="valid", window="blackman"):
    hop = hop or n_fft // 4
    opts = sg.Options(fftSize=n_fft, hop=hop, output=out, smoothingTimeConstant=tau, align=align, window=window)
    frames = eng.num_frames(opts, clip_len)
    x = (torch.randn((n_clips, clip_len), device=dev, generator=gen) * 0.1).float()
    o = torch.empty((n_clips, frames, n_fft // 2), dtype=torch.uint8 if out == "u8" else torch.float32, device=dev)
    torch.cuda.synchronize()
    for _ in range(3):
        eng.spectrogram_device(x.data_ptr(), n_clips, clip_len, clip_len, opts, o.data_ptr(), st.cuda_stream)
    st.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps):
        eng.spectrogram_device(x.data_ptr(), n_clips, clip_len, clip_len, opts, o.data_ptr(), st.cuda_stream)
    e1.record(st); st.synchronize()
    print(f"n_fft {n_fft} hop {hop} clips {n_clips} tau {tau} {out}: {eng.last_kernel} {e0.elapsed_time(e1) / reps:.3f} ms", flush=True)
    del x, o

if "a" in which: timed(2048, 512, 441000, 0.8)
if "b" in which:
    for n in (256, 512, 1024, 4096, 8192): timed(n, 512, 441000, 0.8, reps=5)
if "c" in which: timed(512, 1, 16000 * 3600, 0.0, hop=160, out="db", window="hann")
if "d" in which:
    for n in (256, 512, 1024, 2048, 4096, 8192): timed(n, 2, 48000 * 60, 0.8, align="analyser")
if "e" in which:
    for n in (256, 512, 1024, 2048, 4096, 8192): timed(n, 64, 48000 * 60, 0.0, reps=5)
if "f" in which: timed(1024, 64, 48000 * 60, 0.8)
timed(2048, 64, 48000 * 60, 0.8)
