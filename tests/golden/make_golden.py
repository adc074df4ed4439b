"""Generates tests/golden/*.npz: seeded inputs + float64-oracle outputs for the parity tests.

The reference has no golden vectors for this path and cannot run here, so these
fixtures come from the oracle restatement (oracle/analyser_oracle.py), which is itself pinned by the
closed-form tests in tests/test_oracle_kat.py.  Run: python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import analyser_oracle as O  # noqa: E402

CASES = [
    # name, signal, n_fft, hop, window, output, align, tau, clips, clip_len
    dict(name="cfg1_chirp_u8", signal="chirp", n_fft=2048, hop=512, window=0, output=0, align=0, tau=0.0, clips=1, clip_len=44100),
    dict(name="cfg1_chirp_hann_u8", signal="chirp", n_fft=2048, hop=512, window=1, output=0, align=1, tau=0.0, clips=1, clip_len=44100),
    dict(name="cfg1_chirp_mag", signal="chirp", n_fft=2048, hop=512, window=0, output=3, align=0, tau=0.0, clips=1, clip_len=12288),
    dict(name="cfg2_noise_db", signal="band_noise", n_fft=512, hop=160, window=1, output=1, align=0, tau=0.0, clips=1, clip_len=16000),
    dict(name="cfg3_sweep256_tau", signal="chirp_noise", n_fft=256, hop=64, window=0, output=0, align=1, tau=0.8, clips=2, clip_len=9600),
    dict(name="cfg3_sweep8192_tau", signal="chirp_noise", n_fft=8192, hop=2048, window=0, output=1, align=1, tau=0.8, clips=2, clip_len=24576),
    dict(name="cfg4_n400", signal="noise", n_fft=400, hop=160, window=1, output=1, align=0, tau=0.0, clips=3, clip_len=8000),
    dict(name="cfg5_stream_rgba", signal="noise", n_fft=1024, hop=128, window=0, output=2, align=1, tau=0.0, clips=4, clip_len=2048),
]


def build_case(case):
    n, sig = case["clip_len"], case["signal"]
    rows = []
    for c in range(case["clips"]):
        rng = np.random.default_rng(1234 + c)
        if sig == "chirp":
            x = O.chirp(n, 44100.0, 20.0, 20000.0, 0.5)
        elif sig == "band_noise":
            x = O.band_noise(n, 16000.0, 300.0, 3400.0, 0.1, 1234 + c)
        elif sig == "chirp_noise":
            x = O.chirp(n, 48000.0, 50.0, 20000.0, 0.4) + (0.01 * rng.standard_normal(n)).astype(np.float32)
        else:
            x = (0.1 * rng.standard_normal(n)).astype(np.float32)
        rows.append(x.astype(np.float32))
    cfg = O.Config(n_fft=case["n_fft"], hop=case["hop"], window=case["window"], output=case["output"],
                   align=case["align"], smoothing=case["tau"])
    return np.stack(rows), cfg


def main():
    index = {"generator": "tests/golden/make_golden.py", "oracle": "oracle/analyser_oracle.py (float64)", "cases": []}
    for case in CASES:
        x, cfg = build_case(case)
        out = O.spectrogram(x, cfg)
        fn = case["name"] + ".npz"
        np.savez_compressed(os.path.join(HERE, fn), pcm=x, out=out)
        index["cases"].append({**case, "file": fn})
        print(fn, out.shape, out.dtype)
    with open(os.path.join(HERE, "index.json"), "w") as f:
        json.dump(index, f, indent=1)


if __name__ == "__main__":
    main()
