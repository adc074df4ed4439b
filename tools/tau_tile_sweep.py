import sys, os, torch
sys.path.insert(0, os.getcwd())
import spectrogram_b200 as sg
eng = sg.Engine(0)
st = torch.cuda.Stream()
for n, clips, L in ((2048, 64, 2880000), (1024, 64, 2880000), (2048, 512, 441000), (512, 1, 57600000)):
    hop = n // 4 if n != 512 else 160
    opts = sg.Options(fftSize=n, hop=hop, output="u8", smoothingTimeConstant=0.8)
    fr = eng.num_frames(opts, L)
    x = (torch.randn((clips, L), device="cuda") * 0.1).float()
    o = torch.empty((clips, fr, n // 2), dtype=torch.uint8, device="cuda")
    l0 = eng.launch_count
    for _ in range(2):
        eng.spectrogram_device(x.data_ptr(), clips, L, L, opts, o.data_ptr(), st.cuda_stream)
    st.synchronize()
    nl = (eng.launch_count - l0) // 2
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(5):
        eng.spectrogram_device(x.data_ptr(), clips, L, L, opts, o.data_ptr(), st.cuda_stream)
    e1.record(st); st.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"tile {os.environ.get('SG_TAU_TILE_MB','56')} MB  n_fft {n} clips {clips}: {ms:.3f} ms  {clips*fr/ms/1e3:.1f} M frames/s  launches {nl}", flush=True)
    del x, o
