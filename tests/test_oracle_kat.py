"""Pins the oracles (oracle/analyser_oracle.py float64, oracle/analyser_ref.c float32).

The reference repository holds no tests or golden vectors for this path;
these are the closed-form known answers that follow from the Web Audio AnalyserNode algorithm:
Blackman coefficients a0=.42 a1=.5 a2=.08, 1/N scaling, -100/-30 dB byte mapping."""
import json
import os

import numpy as np
import pytest

from oracle import analyser_oracle as O
from oracle import cref

N = 2048
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def both(x, cfg):
    return O.spectrogram(x, cfg), cref.stft_batch(x, cfg)


def test_window_matches_scipy():
    from scipy.signal import get_window
    for n in (32, 400, 2048):
        assert np.allclose(O.make_window(O.WINDOW_BLACKMAN, n), get_window("blackman", n, fftbins=True), atol=1e-15)
        assert np.allclose(O.make_window(O.WINDOW_HANN, n), get_window("hann", n, fftbins=True), atol=1e-15)


def test_kat_dc():
    x = np.full(N, 0.01, np.float32)
    want_db = 20 * np.log10(0.01 * np.array([0.42, 0.25, 0.04]))
    cfg = O.Config(output=O.OUT_F32_DB)
    for got in both(x, cfg):
        assert np.allclose(got[0, 0, :3], want_db, atol=2e-5)
        assert np.all(got[0, 0, 3:] < -150)
    cfg.output = O.OUT_U8
    for got in both(x, cfg):
        assert list(got[0, 0, :5]) == [191, 174, 116, 0, 0]


def test_kat_bin_centred_sine():
    x = (0.1 * np.sin(2 * np.pi * 100 * np.arange(N) / N)).astype(np.float32)
    want = 20 * np.log10(0.1 * np.array([0.02, 0.125, 0.21, 0.125, 0.02]))
    cfg = O.Config(output=O.OUT_F32_DB)
    for got in both(x, cfg):
        assert np.allclose(got[0, 0, 98:103], want, atol=2e-4)
    cfg.output = O.OUT_U8
    for got in both(x, cfg):
        assert list(got[0, 0, 98:103]) == [167, 225, 242, 225, 167]


def test_kat_analyser_aligned_partial_fill():
    x = np.full(4 * 512, 0.01, np.float32)
    cfg = O.Config(output=O.OUT_F32_DB, align=O.ALIGN_ANALYSER)
    want = [-71.8673, -53.5455, -48.0755, -47.5350]
    for got in both(x, cfg):
        assert got.shape == (1, 4, 1024)
        assert np.allclose(got[0, :, 0], want, atol=1e-3)
    cfg.output = O.OUT_U8
    for got in both(x, cfg):
        assert list(got[0, :, 0]) == [102, 169, 189, 191]


def test_kat_smoothing_recurrence():
    tau, frames = 0.8, 12
    x = np.full(N + (frames - 1) * 512, 0.01, np.float32)
    cfg = O.Config(output=O.OUT_F32_DB, smoothing=tau)
    t = np.arange(frames)
    want = 20 * np.log10(0.0042 * (1 - tau ** (t + 1)))
    for got in both(x, cfg):
        assert np.allclose(got[0, :, 0], want, atol=1e-4)


def test_tau_one_freezes_at_zero_and_tau_zero_is_stateless():
    rng = np.random.default_rng(1)
    x = (0.1 * rng.standard_normal(N * 3)).astype(np.float32)
    frozen = O.spectrogram(x, O.Config(output=O.OUT_U8, smoothing=1.0))
    assert frozen.max() == 0
    a = O.spectrogram(x, O.Config(output=O.OUT_F32_MAG))[0]
    b = O.spectrogram(x[512:], O.Config(output=O.OUT_F32_MAG))[0]
    assert np.array_equal(a[1:], b)


def test_frame_counts_config1():
    assert O.num_frames(441000, 2048, 512, O.ALIGN_VALID) == 858
    assert O.num_frames(441000, 2048, 512, O.ALIGN_ANALYSER) == 861
    assert O.num_frames(100, 2048, 512, O.ALIGN_VALID) == 0


@pytest.mark.parametrize("n_fft,hop", [(32, 8), (256, 64), (400, 160), (512, 160), (1024, 128), (2048, 512), (8192, 2048), (32768, 8192)])
def test_f32_oracle_tracks_f64_oracle(n_fft, hop):
    rng = np.random.default_rng(n_fft)
    x = (0.2 * rng.standard_normal(n_fft + 7 * hop)).astype(np.float32)
    for tau in (0.0, 0.8):
        cfg = O.Config(n_fft=n_fft, hop=hop, output=O.OUT_F32_MAG, smoothing=tau)
        a, b = both(x, cfg)
        assert np.abs(a - b).max() <= 1e-4 * np.abs(a).max()
        cfg.output = O.OUT_U8
        a, b = both(x, cfg)
        assert np.abs(a.astype(int) - b.astype(int)).max() <= 1


def test_cross_check_scipy_stft():
    from scipy.signal import stft
    rng = np.random.default_rng(7)
    x = rng.standard_normal(8192)
    w = O.make_window(O.WINDOW_HANN, 512)
    _, _, z = stft(x, window=w, nperseg=512, noverlap=512 - 160, boundary=None, padded=False, return_onesided=True)
    ref = np.abs(z.T[:, :256]) * w.sum() / 512  # scipy scales by 1/sum(w); the analyser by 1/N
    got = O.spectrogram(x, O.Config(n_fft=512, hop=160, window=O.WINDOW_HANN, output=O.OUT_F32_MAG))[0]
    assert np.allclose(got, ref, rtol=1e-10, atol=1e-13)


def test_cross_check_torch_stft():
    import torch
    rng = np.random.default_rng(8)
    x = rng.standard_normal(6000)
    w = O.make_window(O.WINDOW_BLACKMAN, 400)
    z = torch.stft(torch.from_numpy(x), 400, hop_length=160, window=torch.from_numpy(w), center=False,
                   return_complex=True)
    ref = z.abs().numpy().T[:, :200] / 400
    got = O.spectrogram(x, O.Config(n_fft=400, hop=160, output=O.OUT_F32_MAG))[0]
    assert np.allclose(got, ref, rtol=1e-9, atol=1e-12)


def test_cross_check_torchaudio_spectrogram():
    """torchaudio's Spectrogram transform with ITS OWN window function (Blackman, the AnalyserNode window): both the window
    and the framed transform come from a third party here."""
    torchaudio = pytest.importorskip("torchaudio")
    import torch
    rng = np.random.default_rng(10)
    x = rng.standard_normal(3 * 44100 // 10)
    tr = torchaudio.transforms.Spectrogram(n_fft=2048, hop_length=512, power=1.0, center=False,
                                           window_fn=lambda n: torch.blackman_window(n, periodic=True, dtype=torch.float64))
    ref = tr(torch.from_numpy(x)).numpy().T[:, :1024] / 2048
    got = O.spectrogram(x, O.Config(n_fft=2048, hop=512, output=O.OUT_F32_MAG))[0]
    assert got.shape == ref.shape
    assert np.allclose(got, ref, rtol=1e-9, atol=1e-12)


def test_cross_check_transformers_audio_utils():
    """The numpy STFT behind the Hugging Face feature extractors (BASELINE config 4's shape: n_fft 400, hop 160, Hann),
    amplitude and dB; its window function is its own as well."""
    au = pytest.importorskip("transformers.audio_utils")
    rng = np.random.default_rng(11)
    x = rng.standard_normal(16000)
    w = au.window_function(400, "hann", periodic=True)
    ref = au.spectrogram(x, w, frame_length=400, hop_length=160, power=1.0, center=False, dtype=np.float64).T[:, :200] / 400
    got = O.spectrogram(x, O.Config(n_fft=400, hop=160, window=O.WINDOW_HANN, output=O.OUT_F32_MAG))[0]
    assert got.shape == ref.shape
    # (that implementation keeps its frames in single precision whatever dtype it returns: agreement to float32 rounding)
    assert np.allclose(got, ref, rtol=2e-6, atol=1e-8)
    db = O.spectrogram(x, O.Config(n_fft=400, hop=160, window=O.WINDOW_HANN, output=O.OUT_F32_DB))[0]
    near = ref >= ref.max(axis=-1, keepdims=True) * 1e-3
    assert np.abs(db - 20 * np.log10(ref))[near].max() < 1e-4


def test_parseval_rect_window():
    rng = np.random.default_rng(9)
    x = rng.standard_normal(2048)
    m = O.spectrogram(x, O.Config(window=O.WINDOW_RECT, output=O.OUT_F32_MAG))[0, 0]
    nyq = abs(np.fft.rfft(x)[1024]) / 2048
    energy = m[0] ** 2 + 2 * (m[1:] ** 2).sum() + nyq ** 2
    assert np.isclose(energy, (x ** 2).sum() / 2048, rtol=1e-12)


def test_colormap_lut():
    lut = O.colormap_lut()
    assert tuple(lut[255]) == (255, 20, 20, 255)   # hue 0 = red, + 0.08 background
    assert tuple(lut[0]) == (20, 20, 20, 255)      # hueDash == 6 falls through: black + background
    assert np.array_equal(O.colormap_lut_u32(), cref.colormap_lut_u32())


def test_time_domain_byte():
    x = np.array([-2.0, -1.0, -0.5, 0.0, 0.5, 0.999, 1.0, 3.0], np.float32)
    want = [0, 0, 64, 128, 192, 255, 255, 255]
    assert list(O.time_domain_byte(x)) == want
    assert list(cref.time_domain_byte(x)) == want


def test_validation():
    for bad in (31, 48, 16, 65536):
        with pytest.raises(O.IndexSizeError):
            O.validate_analyser_attrs(bad, -100, -30, 0.8)
    with pytest.raises(O.IndexSizeError):
        O.validate_analyser_attrs(2048, -30, -30, 0.8)
    with pytest.raises(O.IndexSizeError):
        O.validate_analyser_attrs(2048, -100, -30, 1.5)
    O.validate_analyser_attrs(32768, -100, -30, 1.0)


def test_analyser_oracle_matches_batch_and_caches_within_a_quantum():
    rng = np.random.default_rng(3)
    x = (0.3 * rng.standard_normal(128 * 40)).astype(np.float32)
    an = O.AnalyserOracle(1024)
    an.smoothingTimeConstant = 0.5
    rows = []
    buf = np.zeros(512, np.uint8)
    for i in range(0, x.size, 128):
        an.push(x[i:i + 128])
        an.getByteFrequencyData(buf)
        rows.append(buf.copy())
        again = np.zeros(512, np.uint8)
        an.getByteFrequencyData(again)       # same render quantum: no state advance
        assert np.array_equal(again, buf)
    ref = O.spectrogram(x, O.Config(n_fft=1024, hop=128, align=O.ALIGN_ANALYSER, smoothing=0.5))[0]
    assert np.array_equal(np.stack(rows), ref)


def test_golden_fixtures_match_oracle():
    """tests/golden/*.npz were generated by tests/golden/make_golden.py from the float64 oracle;
    this guards the oracle against drift (and is what the GPU tests compare with)."""
    with open(os.path.join(GOLDEN, "index.json")) as f:
        index = json.load(f)
    assert index["cases"]
    from golden.make_golden import build_case
    for case in index["cases"]:
        data = np.load(os.path.join(GOLDEN, case["file"]))
        x, cfg = build_case(case)
        assert np.array_equal(x, data["pcm"])
        got = O.spectrogram(x, cfg)
        if got.dtype == np.uint8:
            assert np.array_equal(got, data["out"])
        else:
            assert np.allclose(got, data["out"], rtol=1e-12, atol=0, equal_nan=True)
