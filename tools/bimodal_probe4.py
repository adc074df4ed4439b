"""bench.py's own pieces before measure_configs, one at a time, then the 64-clip warm-up-mode call."""
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import spectrogram_b200 as sg
which = sys.argv[1]
eng = sg.Engine(0)
dev = torch.device('cuda', 0)
st = torch.cuda.Stream()
gen = torch.Generator(device=dev).manual_seed(1234)

def timed(n_fft, n_clips, clip_len, tau, reps=20):
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

if "e" in which:     # the end-to-end leg: pinned host arrays through sg_stft_batch
    clips, clip_len = 256, 441000
    opts = sg.Options()
    pin_in = sg.PinnedArray((clips, clip_len), np.float32)
    pin_out = sg.PinnedArray((clips, eng.num_frames(opts, clip_len), 1024), np.uint8)
    pin_in.array[...] = 0.1
    for _ in range(3):
        eng.spectrogram(pin_in.array, opts, out=pin_out.array)
if "h" in which:
    sys.path.insert(0, '/root/repo/tools')
    import host_path_ceiling
    print(host_path_ceiling.measure(dev, 128, 2, 4)["t_both"])
if "c" in which:
    import bench
    x = (0.1 * np.random.default_rng(0).standard_normal((16, 441000))).astype(np.float32)
    print(bench.cpu_baseline_block(x, steps=1)["value"])
if "m" in which:
    torch.cuda.empty_cache()
timed(2048, 64, 48000 * 60, 0.8)
