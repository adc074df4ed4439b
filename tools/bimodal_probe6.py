import sys, torch
sys.path.insert(0, '/root/repo')
import spectrogram_b200 as sg
import bench
which = sys.argv[1]
eng = sg.Engine(0)
dev = torch.device('cuda', 0)
stream = torch.cuda.Stream(dev)
gen = torch.Generator(device=dev).manual_seed(1234)
flush = torch.empty(bench.L2_FLUSH_BYTES, dtype=torch.uint8, device=dev) if "F" in which else None

def entry(n_clips, clip_len, opts, elem_bytes, dtype, flush_l2, sigma=0.1):
    frames = eng.num_frames(opts, clip_len)
    bins = opts.fftSize // 2
    x = (torch.randn((n_clips, clip_len), device=dev, generator=gen) * sigma).float()
    o = torch.empty((n_clips, frames, bins), dtype=dtype, device=dev)
    def fn():
        eng.spectrogram_device(x.data_ptr(), n_clips, clip_len, clip_len, opts, o.data_ptr(), stream.cuda_stream)
    fn()
    r = bench.timed_region(torch, fn, stream, 0, flush_buf=flush if flush_l2 else None)
    print(opts.fftSize, n_clips, opts.smoothingTimeConstant, eng.last_kernel, round(r["ms"], 3), flush=True)
    del x, o
    if "E" in which: torch.cuda.empty_cache()

if "1" in which: entry(512, 441000, sg.Options(fftSize=2048, hop=512, output="u8", smoothingTimeConstant=0.8), 1, torch.uint8, False)
if "2" in which:
    for n in (256, 512, 1024, 4096, 8192): entry(512, 441000, sg.Options(fftSize=n, hop=n // 4, output="u8", smoothingTimeConstant=0.8), 1, torch.uint8, False)
if "3" in which: entry(1, 16000 * 600, sg.Options(fftSize=512, hop=160, window="hann", output="db"), 4, torch.float32, False)
if "4" in which:
    for n in (256, 512, 1024, 2048, 4096, 8192): entry(2, 48000 * 60, sg.Options(fftSize=n, hop=n // 4, output="u8", smoothingTimeConstant=0.8, align="analyser"), 1, torch.uint8, "F" in which)
if "5" in which:
    for n in (256, 512, 1024, 2048, 4096, 8192): entry(64, 48000 * 60, sg.Options(fftSize=n, hop=n // 4, output="u8"), 1, torch.uint8, False)
if "6" in which: entry(64, 48000 * 60, sg.Options(fftSize=1024, hop=256, output="u8", smoothingTimeConstant=0.8), 1, torch.uint8, False)
entry(64, 48000 * 60, sg.Options(fftSize=2048, hop=512, output="u8", smoothingTimeConstant=0.8), 1, torch.uint8, False)
