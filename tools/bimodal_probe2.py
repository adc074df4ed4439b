"""Why does bench.py's config3b_tau 2048 entry run at 1.78 ms when the same call takes 0.99 ms elsewhere?"""
import sys, torch
sys.path.insert(0, '/root/repo')
import spectrogram_b200 as sg
eng = sg.Engine(0)
dev = torch.device('cuda', 0)
st = torch.cuda.Stream()
gen = torch.Generator(device=dev).manual_seed(1234)

def timed(n_fft, n_clips, clip_len, tau, reps=30):
    opts = sg.Options(fftSize=n_fft, hop=n_fft // 4, output="u8", smoothingTimeConstant=tau)
    frames = eng.num_frames(opts, clip_len)
    x = (torch.randn((n_clips, clip_len), device=dev, generator=gen) * 0.1).float()
    o = torch.empty((n_clips, frames, n_fft // 2), dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    for _ in range(3):
        eng.spectrogram_device(x.data_ptr(), n_clips, clip_len, clip_len, opts, o.data_ptr(), st.cuda_stream)
    st.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps):
        eng.spectrogram_device(x.data_ptr(), n_clips, clip_len, clip_len, opts, o.data_ptr(), st.cuda_stream)
    e1.record(st); st.synchronize()
    print(f"n_fft {n_fft} clips {n_clips} tau {tau}: {eng.last_kernel} {e0.elapsed_time(e1) / reps:.3f} ms", flush=True)
    del x, o

timed(2048, 64, 48000 * 60, 0.8)
for n in (256, 512, 1024, 2048, 4096, 8192):
    timed(n, 64, 48000 * 60, 0.0, reps=5)
timed(2048, 64, 48000 * 60, 0.8)
timed(1024, 64, 48000 * 60, 0.8)
timed(2048, 64, 48000 * 60, 0.8)
timed(2048, 512, 441000, 0.8)
timed(2048, 64, 48000 * 60, 0.8)
import bench
s = bench.ClockSampler(0); s.start()
timed(2048, 64, 48000 * 60, 0.8)
print(s.stop())
r = bench.timed_region(torch, lambda: None, st, 0)
x = (torch.randn((64, 48000 * 60), device=dev, generator=gen) * 0.1).float()
opts = sg.Options(fftSize=2048, hop=512, output="u8", smoothingTimeConstant=0.8)
o = torch.empty((64, eng.num_frames(opts, 48000 * 60), 1024), dtype=torch.uint8, device=dev)
torch.cuda.synchronize()
print(bench.timed_region(torch, lambda: eng.spectrogram_device(x.data_ptr(), 64, 48000 * 60, 48000 * 60, opts, o.data_ptr(), st.cuda_stream), st, 0))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
print(bench.timed_region(torch, lambda: eng.spectrogram_device(x.data_ptr(), 64, 48000 * 60, 48000 * 60, opts, o.data_ptr(), st.cuda_stream), st, 0))
