import time, numpy as np, sys
sys.path.insert(0, ".")
import spectrogram_b200 as sg
eng = sg.Engine(0)
opts = sg.Options(fftSize=1024, hop=128, output="u8")
bank = sg.StreamBank(256, opts, max_chunk=128, engine=eng)
chunk = sg.PinnedArray((256, 128), np.float32)
chunk.array[...] = (0.1 * np.random.default_rng(0).standard_normal((256, 128))).astype(np.float32)
out = sg.PinnedArray((256, 1, 512), np.uint8)
rgba = sg.PinnedArray((256, 1, 512, 4), np.uint8)
for mode in ("u8+rgba", "u8 only"):
    for _ in range(50): bank.push(chunk.array, out=out.array, out_rgba=rgba.array if mode == "u8+rgba" else None)
    lat = []
    for _ in range(3000):
        t0 = time.perf_counter(); bank.push(chunk.array, out=out.array, out_rgba=rgba.array if mode == "u8+rgba" else None); lat.append((time.perf_counter() - t0) * 1e6)
    lat = np.sort(np.array(lat))
    print(mode, "p50 %.1f us p99 %.1f us mean %.1f" % (lat[len(lat)//2], lat[int(len(lat)*0.99)], lat.mean()), "sum", int(out.array.sum()), int(rgba.array.sum()))
bank.close(); eng.close()
