"""Parity of the CUDA path (through the C ABI, via the Python host layer) against the oracle.

Every test here needs a B200: run with ``-m gpu``.  The float64 oracle is the truth; tolerances are
tests/_tol.py (1e-4 relative magnitude + float32 floor, 1e-3 dB near the peak, +-1 LSB bytes)."""
import json
import os

import numpy as np
import pytest

import spectrogram_b200 as sg
from oracle import analyser_oracle as O
from _tol import assert_bytes_close, assert_db_close, assert_mag_close

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
WIN = {O.WINDOW_BLACKMAN: "blackman", O.WINDOW_HANN: "hann", O.WINDOW_RECT: "rect"}
OUT = {O.OUT_U8: "u8", O.OUT_F32_DB: "db", O.OUT_RGBA8: "rgba", O.OUT_F32_MAG: "mag"}
ALIGN = {O.ALIGN_VALID: "valid", O.ALIGN_ANALYSER: "analyser"}


def opts_of(cfg: O.Config) -> sg.Options:
    return sg.Options(fftSize=cfg.n_fft, hop=cfg.hop, window=WIN[cfg.window], output=OUT[cfg.output],
                      align=ALIGN[cfg.align], minDecibels=cfg.min_db, maxDecibels=cfg.max_db,
                      smoothingTimeConstant=cfg.smoothing)


def check_all_outputs(engine, x, cfg: O.Config):
    """Runs the four output kinds and compares each with the float64 oracle."""
    base = dict(cfg.__dict__)
    x = np.atleast_2d(x)
    ref_mag = O.spectrogram(x, O.Config(**{**base, "output": O.OUT_F32_MAG}))
    got = engine.spectrogram(x, opts_of(O.Config(**{**base, "output": O.OUT_F32_MAG})))
    assert got.shape == ref_mag.shape and got.dtype == np.float32
    assert_mag_close(got, ref_mag)
    got = engine.spectrogram(x, opts_of(O.Config(**{**base, "output": O.OUT_F32_DB})))
    assert_db_close(got, ref_mag)
    ref_u8 = O.finish(ref_mag, O.Config(**{**base, "output": O.OUT_U8}))
    got_u8 = engine.spectrogram(x, opts_of(O.Config(**{**base, "output": O.OUT_U8})))
    assert got_u8.dtype == np.uint8
    assert_bytes_close(got_u8, ref_u8)
    got_rgba = engine.spectrogram(x, opts_of(O.Config(**{**base, "output": O.OUT_RGBA8})))
    assert got_rgba.shape == ref_u8.shape + (4,)
    assert np.array_equal(got_rgba, O.colormap_lut()[got_u8])   # LUT applied to the engine's own bytes: exact


# ----------------------------------------------------------------------------- golden fixtures
def golden_cases():
    with open(os.path.join(GOLDEN, "index.json")) as f:
        return json.load(f)["cases"]


@pytest.mark.parametrize("case", golden_cases(), ids=lambda c: c["name"])
@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4], ids=["auto", "generic", "w32", "wreg", "x2tma"])
def test_golden(engine, case, variant):
    data = np.load(os.path.join(GOLDEN, case["file"]))
    cfg = O.Config(n_fft=case["n_fft"], hop=case["hop"], window=case["window"], output=case["output"],
                   align=case["align"], smoothing=case["tau"])
    engine.set_kernel_variant(variant)
    try:
        got = engine.spectrogram(data["pcm"], opts_of(cfg))
    finally:
        engine.set_kernel_variant(0)
    ref = data["out"]
    assert got.shape == ref.shape
    if cfg.output == O.OUT_U8:
        assert_bytes_close(got, ref)
    elif cfg.output == O.OUT_RGBA8:
        inv = {tuple(v): i for i, v in enumerate(map(tuple, O.colormap_lut()))}
        ref_b = np.array([inv[tuple(p)] for p in ref.reshape(-1, 4)]).reshape(ref.shape[:-1])
        got_b = np.array([inv[tuple(p)] for p in got.reshape(-1, 4)]).reshape(got.shape[:-1])
        assert_bytes_close(got_b, ref_b)
    else:
        ref_mag = O.spectrogram(data["pcm"], O.Config(**{**cfg.__dict__, "output": O.OUT_F32_MAG}))
        if cfg.output == O.OUT_F32_MAG:
            assert_mag_close(got, ref_mag)
        else:
            assert_db_close(got, ref_mag)


# ----------------------------------------------------------------------------- BASELINE configs
def test_config1_chirp_full(engine):
    """10 s mono 44.1 kHz chirp, n_fft 2048, hop 512: Blackman (parity) and Hann (as named)."""
    x = O.chirp(441000, 44100.0, 20.0, 20000.0, 0.5)
    for window in (O.WINDOW_BLACKMAN, O.WINDOW_HANN):
        for align in (O.ALIGN_VALID, O.ALIGN_ANALYSER):
            cfg = O.Config(window=window, align=align)
            check_all_outputs(engine, x, cfg)
            assert engine.last_kernel in ("warp32x32x2p", "warp32x32x2")   # float dB rows take the TMA-staged kernel
    assert engine.spectrogram(x, sg.Options()).shape == (858, 1024)
    assert engine.spectrogram(x, sg.Options(align="analyser")).shape == (861, 1024)


def test_config1_both_kernels_agree(engine):
    x = O.chirp(100000, 44100.0, 20.0, 20000.0, 0.5)
    a = engine.spectrogram(x, sg.Options(output="mag"))
    engine.set_kernel_variant(1)
    try:
        b = engine.spectrogram(x, sg.Options(output="mag"))
        assert engine.last_kernel == "smem"
    finally:
        engine.set_kernel_variant(0)
    assert_mag_close(a, b.astype(np.float64))


def test_config2_speech_band_noise(engine):
    x = O.band_noise(16000 * 20, 16000.0, 300.0, 3400.0, 0.1, 1234)
    check_all_outputs(engine, x, O.Config(n_fft=512, hop=160, window=O.WINDOW_HANN))


@pytest.mark.parametrize("n_fft", [256, 512, 1024, 2048, 4096, 8192])
def test_config3_fft_sweep_with_smoothing(engine, n_fft):
    rng = np.random.default_rng(1234)
    n = 48000 * 2
    x = np.stack([O.chirp(n, 48000.0, 50.0, 20000.0, 0.4) + (0.01 * rng.standard_normal(n)).astype(np.float32)
                  for _ in range(2)])
    check_all_outputs(engine, x, O.Config(n_fft=n_fft, hop=n_fft // 4, smoothing=0.8, align=O.ALIGN_ANALYSER))


def test_config4_n400_clips(engine):
    rng = np.random.default_rng(0)
    x = (0.1 * rng.standard_normal((16, 16000))).astype(np.float32)
    check_all_outputs(engine, x, O.Config(n_fft=400, hop=160, window=O.WINDOW_HANN))
    assert engine.spectrogram(x, sg.Options(fftSize=400, hop=160)).shape == (16, 98, 200)
    assert engine.last_kernel == "r400"


@pytest.mark.parametrize("n_fft,hop,kernel", [(1024, 256, "p16"), (1024, 128, "p16"), (1024, 512, "p16"), (1024, 160, "p16"), (512, 160, "p8"), (512, 128, "p8"), (512, 64, "p8"), (512, 256, "p8"), (256, 32, "p4"), (512, 100, "w16"), (256, 64, "p4"), (256, 50, "w16"),
                                               (2048, 512, "warp32x32x2p"), (2048, 256, "warp32x32x2p"),
                                               (2048, 1024, "warp32x32x2"), (2048, 768, "warp32x32x2"), (2048, 441, "warp32x32x2"), (2048, 147, "warp32x32x2"),
                                               (2048, 1025, "warp32x32"),
                                               (4096, 1024, "eo4096"), (4096, 512, "eo4096"), (4096, 441, "wreg"), (4096, 444, "eo4096"), (4096, 4096, "eo4096")])
@pytest.mark.parametrize("align,clip_len,n_clips", [("valid", 9000, 5), ("analyser", 4097, 3), ("valid", 2 * 8192 + 2, 1),
                                                    ("analyser", 40000, 2)])
def test_dedicated_kernels_at_clip_edges(engine, n_fft, hop, kernel, align, clip_len, n_clips):
    """Ragged clip lengths (odd frame counts, pairs straddling clips), zero history, unaligned clip starts: the
    dedicated kernels against the oracle and against the register family / generic kernel on the same input."""
    rng = np.random.default_rng(n_fft + hop + clip_len)
    x = (0.2 * rng.standard_normal((n_clips, clip_len))).astype(np.float32)
    al = O.ALIGN_VALID if align == "valid" else O.ALIGN_ANALYSER
    check_all_outputs(engine, x, O.Config(n_fft=n_fft, hop=hop, window=O.WINDOW_BLACKMAN, align=al))
    by = engine.spectrogram(x, sg.Options(fftSize=n_fft, hop=hop, align=align, output="u8"))
    if by.size and (by.shape[1] >= 4 or kernel == "eo4096"):
        assert engine.last_kernel == kernel
    o = sg.Options(fftSize=n_fft, hop=hop, align=align, output="mag")
    a = engine.spectrogram(x, o)
    for variant in (1, 3):
        engine.set_kernel_variant(variant)
        try:
            b = engine.spectrogram(x, o)
        finally:
            engine.set_kernel_variant(0)
        if a.size:
            assert_mag_close(a, b.astype(np.float64))


@pytest.mark.parametrize("hop,align,clip_len,n_clips", [(160, "valid", 16000, 7), (160, "analyser", 4801, 3),
                                                        (100, "valid", 3333, 5), (77, "analyser", 1000, 2),
                                                        (400, "valid", 400, 1), (160, "valid", 399, 2)])
def test_n400_register_kernel_edges(engine, hop, align, clip_len, n_clips):
    """Ragged clip lengths, odd hops (unaligned frame starts), zero history, fewer frames than a warp holds."""
    rng = np.random.default_rng(hop + clip_len)
    x = (0.2 * rng.standard_normal((n_clips, clip_len))).astype(np.float32)
    al = O.ALIGN_VALID if align == "valid" else O.ALIGN_ANALYSER
    check_all_outputs(engine, x, O.Config(n_fft=400, hop=hop, window=O.WINDOW_BLACKMAN, align=al))
    a = engine.spectrogram(x, sg.Options(fftSize=400, hop=hop, align=align, output="mag"))
    engine.set_kernel_variant(1)
    try:
        b = engine.spectrogram(x, sg.Options(fftSize=400, hop=hop, align=align, output="mag"))
        assert engine.last_kernel in ("smem", "none")
    finally:
        engine.set_kernel_variant(0)
    if a.size:
        assert_mag_close(a, b.astype(np.float64))


def test_config5_streaming_equals_batch(engine):
    """256 channels, n_fft 1024, hop 128, one render quantum per push; bytes + colour."""
    rng = np.random.default_rng(5)
    ch, chunks = 256, 12
    x = (0.2 * rng.standard_normal((ch, 128 * chunks))).astype(np.float32)
    opts = sg.Options(fftSize=1024, hop=128, output="u8")
    bank = sg.StreamBank(ch, opts, max_chunk=256, engine=engine)
    rows, rgba_rows = [], []
    i = 0
    for step in (128, 256, 128, 256, 256, 128, 256, 128):
        out, rgba = bank.push(x[:, i:i + step], want_rgba=True)
        rows.append(out)
        rgba_rows.append(rgba)
        i += step
    assert i == x.shape[1] and bank.frames_emitted == chunks
    got = np.concatenate(rows, axis=1)
    got_rgba = np.concatenate(rgba_rows, axis=1)
    # one-frame-per-channel pushes run on the register family; a batch call on the same kernel is bit identical ...
    engine.set_kernel_variant(3)
    try:
        batch = engine.spectrogram(x, sg.Options(fftSize=1024, hop=128, align="analyser"))
    finally:
        engine.set_kernel_variant(0)
    assert np.array_equal(got, batch)            # streaming == batch, bit exact
    # ... and the batch default (the half-warp pair kernel) agrees within one byte level
    batch_default = engine.spectrogram(x, sg.Options(fftSize=1024, hop=128, align="analyser"))
    assert engine.last_kernel == "p16"
    assert_bytes_close(got, batch_default)
    assert np.array_equal(got_rgba, O.colormap_lut()[got])
    ref = O.spectrogram(x, O.Config(n_fft=1024, hop=128, align=O.ALIGN_ANALYSER))
    assert_bytes_close(got, ref)
    bank.reset()
    out = bank.push(x[:, :128])
    assert np.array_equal(out, got[:, :1])
    bank.close()


def test_streaming_with_smoothing_carries_state(engine):
    rng = np.random.default_rng(6)
    x = (0.2 * rng.standard_normal((8, 128 * 10))).astype(np.float32)
    opts = sg.Options(fftSize=512, hop=128, output="db", smoothingTimeConstant=0.8)
    bank = sg.StreamBank(8, opts, max_chunk=128, engine=engine)
    got = np.concatenate([bank.push(x[:, i:i + 128]) for i in range(0, x.shape[1], 128)], axis=1)
    bank.close()
    ref_mag = O.spectrogram(x, O.Config(n_fft=512, hop=128, align=O.ALIGN_ANALYSER, smoothing=0.8, output=O.OUT_F32_MAG))
    assert_db_close(got, ref_mag)


# ----------------------------------------------------------------------------- sizes and edges
@pytest.mark.parametrize("n_fft", [32, 64, 128, 4096, 16384, 32768, 6, 10, 12, 60, 100, 1000, 1200, 3000])
def test_every_legal_size(engine, n_fft):
    rng = np.random.default_rng(n_fft)
    hop = max(1, n_fft // 3)
    x = (0.3 * rng.standard_normal((2, n_fft + 5 * hop + 3))).astype(np.float32)
    check_all_outputs(engine, x, O.Config(n_fft=n_fft, hop=hop))


@pytest.mark.parametrize("hop", [1, 3, 64, 128, 500, 511, 512, 513, 2048, 5000])
def test_ragged_hops_n2048(engine, hop):
    """odd hops give 4-byte-aligned frames (the 8-byte fast loads do not apply), hop > n_fft skips samples"""
    rng = np.random.default_rng(hop)
    x = (0.3 * rng.standard_normal((3, 2048 + 9 * hop + 1))).astype(np.float32)
    for align in (O.ALIGN_VALID, O.ALIGN_ANALYSER):
        check_all_outputs(engine, x, O.Config(hop=hop, align=align))


def test_low_min_decibels_takes_the_exact_log_path(engine):
    """minDecibels below -300 dB: subnormal powers matter, the packed kernel's flush-to-zero lg2 is not used"""
    x = (1e-17 * np.random.default_rng(3).standard_normal((2, 8192))).astype(np.float32)
    cfg = O.Config(min_db=-1000.0, max_db=0.0)
    got = engine.spectrogram(x, opts_of(cfg))
    assert engine.last_kernel == "warp32x32"
    assert_bytes_close(got, O.spectrogram(x, cfg))


def test_empty_and_short_inputs(engine):
    assert engine.spectrogram(np.zeros((3, 100), np.float32), sg.Options()).shape == (3, 0, 1024)
    assert engine.spectrogram(np.zeros((0, 4096), np.float32), sg.Options()).shape == (0, 5, 1024)
    out = engine.spectrogram(np.zeros((2, 700), np.float32), sg.Options(align="analyser"))
    assert out.shape == (2, 1, 1024) and out.max() == 0
    out = engine.spectrogram(np.ones(2048, np.float32), sg.Options())
    assert out.shape == (1, 1024)


def test_silence_full_scale_and_non_finite(engine):
    x = np.zeros((1, 4096), np.float32)
    assert engine.spectrogram(x, sg.Options()).max() == 0
    db = engine.spectrogram(x, sg.Options(output="db"))
    assert np.all(np.isneginf(db))
    loud = np.ones((1, 4096), np.float32) * 30.0
    assert engine.spectrogram(loud, sg.Options())[0, 0, 0] == 255
    bad = (0.1 * np.random.default_rng(0).standard_normal((1, 4096))).astype(np.float32)
    bad[0, 1000] = np.nan
    out = engine.spectrogram(bad, sg.Options(output="mag"))
    ref = O.spectrogram(bad, O.Config(output=O.OUT_F32_MAG))
    assert np.all(np.isfinite(out)) and np.array_equal(out == 0, ref == 0)   # [SPEC] non-finite -> 0
    assert engine.spectrogram(bad, sg.Options()).max() <= 255


def test_db_range_and_custom_window_and_colormap(engine):
    rng = np.random.default_rng(11)
    x = (0.05 * rng.standard_normal((2, 8192))).astype(np.float32)
    cfg = O.Config(min_db=-80.0, max_db=-10.0)
    check_all_outputs(engine, x, cfg)
    w = np.kaiser(2048, 8.0)
    got = engine.spectrogram(x, sg.Options(window=w.astype(np.float32), output="mag"))
    ref = O.spectrogram(x, O.Config(window=O.WINDOW_CUSTOM, custom_window=w.astype(np.float32), output=O.OUT_F32_MAG))
    assert_mag_close(got, ref)
    lut = (np.arange(256, dtype=np.uint32) * 0x01010101)
    got = engine.spectrogram(x, sg.Options(output="rgba", colormap=lut))
    by = engine.spectrogram(x, sg.Options())
    assert np.array_equal(got[..., 0], by) and np.array_equal(got[..., 3], by)


def test_invalid_options_raise_like_web_audio(engine):
    x = np.zeros(8192, np.float32)
    for kw in (dict(fftSize=2047), dict(fftSize=14), dict(hop=0), dict(minDecibels=-30, maxDecibels=-30),
               dict(smoothingTimeConstant=1.01)):
        with pytest.raises(sg.IndexSizeError):
            engine.spectrogram(x, sg.Options(**kw))
    with pytest.raises(TypeError):
        engine.spectrogram(np.zeros((2, 2, 2), np.float32), sg.Options())


# ----------------------------------------------------------------------------- AnalyserNode object
def test_analyser_node_defaults_and_validation(engine):
    an = sg.AnalyserNode(engine)
    assert (an.fftSize, an.frequencyBinCount, an.minDecibels, an.maxDecibels, an.smoothingTimeConstant) == \
        (2048, 1024, -100.0, -30.0, 0.8)
    for bad in (0, 31, 48, 65536, 2048.5):
        with pytest.raises(sg.IndexSizeError):
            an.fftSize = bad
    with pytest.raises(sg.IndexSizeError):
        an.minDecibels = -30
    with pytest.raises(sg.IndexSizeError):
        an.maxDecibels = -100
    with pytest.raises(sg.IndexSizeError):
        an.smoothingTimeConstant = 1.5
    with pytest.raises(TypeError):
        an.getByteFrequencyData(np.zeros(1024, np.float32))
    with pytest.raises(TypeError):
        an.getFloatFrequencyData([0.0] * 1024)
    an.fftSize = 1024                      # player.js:10 mobile branch
    assert an.frequencyBinCount == 512
    an.close()


def test_analyser_node_follows_the_reference_call_pattern(engine):
    """player.js:7-11 configuration, visualizer.js:346-368 polling, against the oracle object."""
    rng = np.random.default_rng(2)
    x = (0.3 * rng.standard_normal(128 * 60)).astype(np.float32)
    for fft, tau in ((2048, 0.0), (2048, 0.75), (1024, 0.1), (32, 0.8), (32768, 0.5)):
        an, ref = sg.AnalyserNode(engine), O.AnalyserOracle()
        an.fftSize = fft
        ref.fftSize = fft
        an.smoothingTimeConstant = tau
        ref.smoothingTimeConstant = tau
        bins = an.frequencyBinCount
        by, by_ref = np.zeros(bins, np.uint8), np.zeros(bins, np.uint8)
        fl, fl_ref = np.zeros(bins, np.float32), np.zeros(bins, np.float64)
        td, td_ref = np.zeros(fft, np.uint8), np.zeros(fft, np.uint8)
        tf, tf_ref = np.zeros(fft, np.float32), np.zeros(fft, np.float64)
        for i in range(0, x.size, 640):   # 5 render quanta between polls
            an.push(x[i:i + 640])
            ref.push(x[i:i + 640])
            an.getByteFrequencyData(by)
            ref.getByteFrequencyData(by_ref)
            assert_bytes_close(by, by_ref, max_mismatch_frac=0.05)
            an.getFloatFrequencyData(fl)   # same quantum: must not advance the smoothing state
            ref.getFloatFrequencyData(fl_ref)
            assert_db_close(fl[None], ref._state[None])
            an.getByteTimeDomainData(td)
            ref.getByteTimeDomainData(td_ref)
            assert np.array_equal(td, td_ref)
            an.getFloatTimeDomainData(tf)
            ref.getFloatTimeDomainData(tf_ref)
            assert np.array_equal(tf, tf_ref.astype(np.float32))
        short = np.full(10, 7, np.uint8)
        an.getByteFrequencyData(short)          # copies min(len, bins) elements
        assert np.array_equal(short, by[:10])
        long_ = np.full(bins + 5, 7, np.uint8)
        an.getByteFrequencyData(long_)
        assert np.array_equal(long_[:bins], by) and np.all(long_[bins:] == 7)   # never resizes, tail untouched
        an.close()


# ----------------------------------------------------------------------------- device pointers, sharding
def test_device_pointer_entry_and_stream(engine):
    import torch
    x = O.chirp(200000, 44100.0, 20.0, 20000.0, 0.5)
    ref = engine.spectrogram(x, sg.Options())
    xd = torch.from_numpy(x).cuda()
    out = torch.empty(ref.shape, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    st = torch.cuda.Stream()
    before = engine.launch_count
    engine.spectrogram_device(xd.data_ptr(), 1, x.size, x.size, sg.Options(), out.data_ptr(), st.cuda_stream)
    st.synchronize()
    assert engine.launch_count == before + 1
    assert np.array_equal(out.cpu().numpy(), ref)


def test_shard_invariance_bit_exact(engine):
    """1-GPU result == concatenation of per-shard results; also the threaded multi-device host path."""
    rng = np.random.default_rng(4)
    x = (0.2 * rng.standard_normal((13, 20000))).astype(np.float32)
    whole = engine.spectrogram(x, sg.Options())
    parts = [engine.spectrogram(x[lo:hi], sg.Options()) for lo, hi in sg.shard_bounds(13, 4)]
    assert np.array_equal(np.concatenate(parts), whole)
    # sg_stft_batch_multi: one host thread per engine inside the library, disjoint slices of one output array
    n_dev = sg.device_count()
    if n_dev > 1:
        assert np.array_equal(sg.spectrogram(x, devices=list(range(n_dev))), whole)
    extra = [sg.Engine(0), sg.Engine(0)]              # three engines (here on one GPU) run their blocks concurrently
    try:
        for opts in (sg.Options(), sg.Options(output="db", smoothingTimeConstant=0.6)):
            one = engine.spectrogram(x, opts)
            assert np.array_equal(sg.spectrogram_multi([engine] + extra, x, opts), one)
            assert np.array_equal(sg.spectrogram_multi(extra, x[:1], opts), one[:1])      # fewer clips than engines
        with pytest.raises(TypeError):
            sg.spectrogram_multi([engine, engine], x, sg.Options())
    finally:
        for e in extra:
            e.close()


def test_pinned_and_pageable_host_buffers_agree(engine):
    rng = np.random.default_rng(12)
    x = (0.2 * rng.standard_normal((40, 300000))).astype(np.float32)   # > one 32 MB chunk: exercises the pipeline
    a = engine.spectrogram(x, sg.Options())
    pin_in = sg.PinnedArray(x.shape, np.float32)
    pin_in.array[...] = x
    pin_out = sg.PinnedArray(a.shape, np.uint8)
    engine.spectrogram(pin_in.array, sg.Options(), out=pin_out.array)
    assert np.array_equal(pin_out.array, a)
    ref = O.spectrogram(x[:2], O.Config())
    assert_bytes_close(a[:2], ref)
    pin_in.free()
    pin_out.free()


def test_long_single_clip_is_chunked_over_frames(engine):
    """config 2 shape (one long stream): frame-range chunks, with and without smoothing carry."""
    rng = np.random.default_rng(13)
    x = (0.1 * rng.standard_normal(16000 * 900)).astype(np.float32)   # 57.6 MB > chunk size
    for tau in (0.0, 0.8):
        got = engine.spectrogram(x, sg.Options(fftSize=512, hop=160, window="hann", output="db", smoothingTimeConstant=tau))
        sel = np.r_[0:50, 40000:40050, got.shape[0] - 50:got.shape[0]]
        if tau == 0.0:
            ref_mag = O.spectrogram(x, O.Config(n_fft=512, hop=160, window=O.WINDOW_HANN, output=O.OUT_F32_MAG))[0]
            assert_db_close(got[sel], ref_mag[sel])
        else:
            # recurrence across chunk borders: compare around the first border with a local oracle run
            per = (32 << 20) // (160 * 4)
            lo = per - 200
            seg = x[lo * 160: (per + 50) * 160 + 512]
            ref_mag = O.spectrogram(seg, O.Config(n_fft=512, hop=160, window=O.WINDOW_HANN, smoothing=tau, output=O.OUT_F32_MAG))[0]
            assert_db_close(got[lo + 150: per + 50], ref_mag[150:250])   # 150 frames of warm-up: 0.8^150 ~ 3e-15


# ----------------------------------------------------------------------------- size-independent properties at full size
def test_properties_at_full_size(engine):
    """BASELINE-sized batch (64 Ki frames at n_fft 2048 / hop 512): linearity, Parseval, determinism."""
    import torch
    frames = 65536
    n = 2048 + (frames - 1) * 512
    g = torch.Generator(device="cuda").manual_seed(1)
    xd = (torch.rand(n, device="cuda", generator=g) - 0.5) * 0.5
    opts = sg.Options(output="mag", window="rect")
    mag = torch.empty((frames, 1024), dtype=torch.float32, device="cuda")
    mag2 = torch.empty_like(mag)
    x2 = xd * 2.0
    torch.cuda.synchronize()
    st_obj = torch.cuda.Stream()
    st = st_obj.cuda_stream
    engine.spectrogram_device(xd.data_ptr(), 1, n, n, opts, mag.data_ptr(), st)
    engine.spectrogram_device(x2.data_ptr(), 1, n, n, opts, mag2.data_ptr(), st)
    torch.cuda.synchronize()
    assert torch.allclose(mag2, 2.0 * mag, rtol=1e-5, atol=0)                 # linearity of |X| (exact power of two)
    # Parseval per frame (rect window): sum_k' |X|^2 = sum x^2 / N, Nyquist bin recomputed separately
    fr = xd.unfold(0, 2048, 512)
    sign = torch.ones(2048, device="cuda")
    sign[1::2] = -1
    nyq = (fr * sign).sum(dim=1) / 2048
    energy = mag[:, 0] ** 2 + 2 * (mag[:, 1:] ** 2).sum(dim=1) + nyq ** 2
    want = (fr ** 2).sum(dim=1) / 2048
    assert torch.allclose(energy, want, rtol=2e-4)
    by1 = torch.empty((frames, 1024), dtype=torch.uint8, device="cuda")
    by2 = torch.empty_like(by1)
    engine.spectrogram_device(xd.data_ptr(), 1, n, n, sg.Options(), by1.data_ptr(), st)
    engine.spectrogram_device(xd.data_ptr(), 1, n, n, sg.Options(), by2.data_ptr(), st)
    torch.cuda.synchronize()
    assert torch.equal(by1, by2)                                              # run-to-run identical
    # byte monotone in amplitude
    engine.spectrogram_device(x2.data_ptr(), 1, n, n, sg.Options(), by2.data_ptr(), st)
    torch.cuda.synchronize()
    assert bool((by2 >= by1).all())
    # spot-check a slice against the oracle
    sl = xd[: 2048 + 63 * 512].cpu().numpy()
    ref = O.spectrogram(sl, O.Config())[0]
    assert_bytes_close(by1[:64].cpu().numpy(), ref)

@pytest.mark.parametrize("seed", range(24))
def test_random_shapes_agree_with_the_generic_kernel(engine, seed):
    """Seeded random (n_fft, hop, clip length, clip count, alignment, output): whatever kernel the dispatcher picks
    must agree with the generic shared-memory kernel on the same input (and the bytes with the oracle)."""
    rng = np.random.default_rng(1000 + seed)
    n_fft = int(rng.choice([256, 400, 512, 1024, 2048, 4096]))
    hop = int(rng.choice([n_fft // 4, n_fft // 8, 160, 128, 100, n_fft // 2, n_fft]))
    n_clips = int(rng.integers(1, 6))
    clip_len = int(rng.integers(n_fft // 2, 12 * n_fft))
    align = str(rng.choice(["valid", "analyser"]))
    out = str(rng.choice(["u8", "db", "mag", "rgba"]))
    x = (0.3 * rng.standard_normal((n_clips, clip_len))).astype(np.float32)
    o = sg.Options(fftSize=n_fft, hop=hop, align=align, output="mag")
    a = engine.spectrogram(x, o)
    engine.set_kernel_variant(1)
    try:
        b = engine.spectrogram(x, o)
    finally:
        engine.set_kernel_variant(0)
    assert a.shape == b.shape
    if a.size:
        assert_mag_close(a, b.astype(np.float64))
    if out != "mag":
        cfg = O.Config(n_fft=n_fft, hop=hop, align=O.ALIGN_VALID if align == "valid" else O.ALIGN_ANALYSER,
                       output={"u8": O.OUT_U8, "db": O.OUT_F32_DB, "rgba": O.OUT_RGBA8}[out])
        got = engine.spectrogram(x, sg.Options(fftSize=n_fft, hop=hop, align=align, output=out))
        ref_mag = O.spectrogram(x, O.Config(**{**cfg.__dict__, "output": O.OUT_F32_MAG}))
        if out == "db":
            assert_db_close(got, ref_mag)
        else:
            ref_u8 = O.finish(ref_mag, O.Config(**{**cfg.__dict__, "output": O.OUT_U8}))
            if out == "u8":
                assert_bytes_close(got, ref_u8)
            else:
                got_u8 = engine.spectrogram(x, sg.Options(fftSize=n_fft, hop=hop, align=align, output="u8"))
                assert_rgba_is_lut_of(got, got_u8)

@pytest.mark.parametrize("n_fft,hop", [(256, 64), (400, 160), (512, 160), (512, 128), (1024, 256), (1024, 128), (2048, 512),
                                       (4096, 1024), (600, 150)])
@pytest.mark.parametrize("bad_value", [np.nan, np.inf, -np.inf])
def test_non_finite_samples_zero_their_frames_only(engine, n_fft, hop, bad_value):
    """[SPEC] "if X^[k] is NaN or infinite, set it to 0": a non-finite sample zeroes exactly the frames that contain it
    (every kernel decides this once per frame) and leaves the neighbours untouched."""
    rng = np.random.default_rng(7)
    n_clips, clip_len = 3, 20 * n_fft
    x = (0.1 * rng.standard_normal((n_clips, clip_len))).astype(np.float32)
    clean = {out: engine.spectrogram(x, sg.Options(fftSize=n_fft, hop=hop, output=out)) for out in ("u8", "mag", "db")}
    pos = 7 * n_fft + 3
    x[1, pos] = bad_value
    frames = clean["u8"].shape[1]
    hit = np.array([t * hop <= pos < t * hop + n_fft for t in range(frames)])
    assert hit.any() and not hit.all()
    for out in ("u8", "mag", "db"):
        got = engine.spectrogram(x, sg.Options(fftSize=n_fft, hop=hop, output=out))
        assert np.array_equal(got[0], clean[out][0]) and np.array_equal(got[2], clean[out][2])
        assert np.array_equal(got[1][~hit], clean[out][1][~hit])
        if out == "db":
            assert np.all(np.isneginf(got[1][hit]))
        else:
            assert not got[1][hit].any()


def test_time_parallel_scan_agrees_with_the_sequential_recurrence(engine):
    """tau > 0: one long clip takes the chunked look-back scan (aggregates carried as float), a large batch takes one
    thread per (clip, bin) walking the frames in order.  Same clip both ways: the smoothed magnitudes agree to float
    rounding of the carried state, the bytes to the LSB."""
    rng = np.random.default_rng(99)
    n_fft, hop, frames = 256, 64, 320
    clip_len = n_fft + (frames - 1) * hop
    one = (0.2 * rng.standard_normal(clip_len)).astype(np.float32)
    batch = np.broadcast_to(one, (2400, clip_len)).copy()      # 2400 x 128 bins > 2048 threads x 148 SMs: sequential kernel
    for out in ("mag", "u8"):
        opts = sg.Options(fftSize=n_fft, hop=hop, output=out, smoothingTimeConstant=0.9)
        a = engine.spectrogram(one, opts)
        b = engine.spectrogram(batch, opts)[17]
        if out == "mag":
            assert np.max(np.abs(a - b) / (np.abs(b) + 1e-30)) < 2e-6
        else:
            d = np.abs(a.astype(np.int32) - b.astype(np.int32))
            assert d.max() <= 1 and (d != 0).mean() < 1e-4


def assert_rgba_is_lut_of(got_rgba, got_u8):
    """RGBA output = the LUT of the byte output.  The two runs are cut into different chunks by the host pipeline (an RGBA
    row is four times a byte row), so with tau > 0 they may take differently rounded smoothing paths (chained segments,
    look-back, two-kernel): a byte on a rounding boundary may then differ by one level between them."""
    lut = O.colormap_lut()
    ok = (got_rgba == lut[got_u8]).all(-1)
    if not ok.all():
        up = (got_rgba == lut[np.clip(got_u8.astype(np.int32) + 1, 0, 255)]).all(-1)
        dn = (got_rgba == lut[np.clip(got_u8.astype(np.int32) - 1, 0, 255)]).all(-1)
        assert (ok | up | dn).all() and (~ok).mean() < 1e-4


# ----------------------------------------------------------------------------- fused smoothing kernel (n_fft 2048, tau > 0)
@pytest.mark.parametrize("n_clips,clip_len,hop,align", [
    (1, 2048 + 511 * 512, 512, O.ALIGN_VALID),        # few clips: aggregate pass + look-back + emit pass, even frames
    (3, 2048 + 300 * 512 + 77, 512, O.ALIGN_ANALYSER),  # zero history at the start, odd frame count
    (2, 2048 + 700 * 256, 256, O.ALIGN_VALID),        # hop 256 instantiation
    (160, 2048 + 120 * 512, 512, O.ALIGN_VALID),      # more clips than half the SMs: chained segments
    (301, 2048 + 97 * 512, 512, O.ALIGN_ANALYSER),    # chained, odd frames per clip, several tasks per CTA
    (150, 2048 + 75 * 1024 + 5, 1024, O.ALIGN_ANALYSER),  # hop n/2 instantiation, chained
    (2, 2048 + 400 * 1024, 1024, O.ALIGN_VALID),      # hop n/2, look-back mode
    (150, 2048 + 130 * 441 + 3, 441, O.ALIGN_ANALYSER),   # any-hop loader: odd hop (10 ms at 44.1 kHz), 4-byte loads
    (3, 2048 + 600 * 735, 735, O.ALIGN_VALID),        # the reference's display cadence as a hop (1/60 s at 44.1 kHz), look-back
    (151, 2048 + 90 * 160, 160, O.ALIGN_VALID),       # even hop, 8-byte loads
    (149, 2048 * 40, 2048, O.ALIGN_VALID),            # frames that do not overlap
])
def test_fused_smoothing_kernel_matches_the_oracle(engine, n_clips, clip_len, hop, align):
    rng = np.random.default_rng(n_clips)
    x = (0.05 * rng.standard_normal((n_clips, clip_len))).astype(np.float32)
    x += O.chirp(clip_len, 44100.0, 100.0, 15000.0, 0.3)[None, :]
    sel = np.unique(np.r_[0, min(1, n_clips - 1), n_clips // 2, n_clips - 1])
    cfg = O.Config(n_fft=2048, hop=hop, smoothing=0.8, align=align, output=O.OUT_F32_MAG)
    ref_mag = O.spectrogram(x[sel], cfg)
    for out in ("mag", "db", "u8", "rgba"):
        opts = sg.Options(fftSize=2048, hop=hop, output=out, smoothingTimeConstant=0.8, align=ALIGN[align])
        engine.set_kernel_variant(7)      # the fused kernel at any clip count (auto: only from ~2/3 of the SMs in clips)
        try:
            got = engine.spectrogram(x, opts)
        finally:
            engine.set_kernel_variant(0)
        assert engine.last_kernel == "warp32x32x2s"
        if out == "mag":
            assert_mag_close(got[sel], ref_mag)
        elif out == "db":
            assert_db_close(got[sel], ref_mag)
        elif out == "u8":
            got_u8 = got
            assert_bytes_close(got[sel], O.finish(ref_mag, O.Config(n_fft=2048, hop=hop, smoothing=0.8, align=align)))
        else:
            assert_rgba_is_lut_of(got, got_u8)


def test_fused_smoothing_chained_segments_are_the_sequential_arithmetic(engine):
    """Chain mode hands the float state from segment to segment: cutting a clip into segments must not change a bit.
    The same clips as a small batch (look-back over float aggregates, or the two-kernel path) agree to rounding.
    Device-resident batches: the host entry cuts 200 clips into chunks of ~38, below the fused kernel's threshold."""
    import torch
    rng = np.random.default_rng(5)
    clip_len = 2048 + 400 * 512
    x = torch.from_numpy((0.2 * rng.standard_normal((200, clip_len))).astype(np.float32)).cuda()
    opts = sg.Options(fftSize=2048, hop=512, output="mag", smoothingTimeConstant=0.75)
    frames = engine.num_frames(opts, clip_len)

    def run(n):
        out = torch.empty((n, frames, 1024), dtype=torch.float32, device="cuda")
        engine.spectrogram_device(x.data_ptr(), n, clip_len, clip_len, opts, out.data_ptr())
        engine.synchronize()
        return out.cpu().numpy()

    a = run(200)                                       # chained segments
    assert engine.last_kernel == "warp32x32x2s"
    b = run(150)                                       # another split of the same clips into tasks
    assert engine.last_kernel == "warp32x32x2s"
    assert np.array_equal(a[:150], b)
    engine.set_kernel_variant(7)
    try:
        c = run(2)                                     # look-back over float aggregates
        assert engine.last_kernel == "warp32x32x2s"
    finally:
        engine.set_kernel_variant(0)
    close = lambda u, v: np.all(np.abs(u - v) <= 2e-5 * np.abs(v) + 1e-7 * v.max(axis=-1, keepdims=True))
    assert close(c, a[:2])
    d = run(2)                                         # auto: two clips take the two-kernel path
    assert engine.last_kernel != "warp32x32x2s"
    assert close(d, a[:2])
    ref = O.spectrogram(x[:1].cpu().numpy(), O.Config(n_fft=2048, hop=512, smoothing=0.75, output=O.OUT_F32_MAG))
    assert_mag_close(a[:1], ref)


def test_fused_smoothing_non_finite_frames_reset_the_state(engine):
    rng = np.random.default_rng(8)
    clip_len = 2048 + 60 * 512
    x = (0.2 * rng.standard_normal((2, clip_len))).astype(np.float32)
    x[0, 2048 + 20 * 512 + 5] = np.nan        # four frames of clip 0 see the NaN
    x[1, 2048 + 30 * 512 + 9] = np.inf
    cfg = O.Config(n_fft=2048, hop=512, smoothing=0.8, output=O.OUT_F32_MAG)
    with np.errstate(invalid="ignore", over="ignore"):
        ref = O.spectrogram(x, cfg)
    engine.set_kernel_variant(7)
    try:
        got = engine.spectrogram(x, sg.Options(fftSize=2048, hop=512, output="mag", smoothingTimeConstant=0.8))
    finally:
        engine.set_kernel_variant(0)
    assert engine.last_kernel == "warp32x32x2s"
    for c in (0, 1):
        bad = np.all(ref[c] == 0, axis=1)
        assert bad.sum() == 4 and np.all(got[c][bad] == 0)
    assert_mag_close(got, ref)


def test_fused_smoothing_kernel_on_frame_ranges_of_a_long_clip(engine):
    """The host entry cuts one long clip into frame ranges and chains them through the per-clip state: the fused kernel
    (forced: one clip is below its automatic threshold) must pick the state up and write rows at the range's offset."""
    rng = np.random.default_rng(21)
    x = (0.1 * rng.standard_normal(44100 * 240)).astype(np.float32)      # 42 MB > one 32 MB chunk
    opts = sg.Options(fftSize=2048, hop=512, output="mag", smoothingTimeConstant=0.8)
    engine.set_kernel_variant(7)
    try:
        got = engine.spectrogram(x, opts)
    finally:
        engine.set_kernel_variant(0)
    assert engine.last_kernel == "warp32x32x2s"
    per = (32 << 20) // (512 * 4)                     # frames per chunk of the host pipeline
    assert got.shape[0] > per + 100
    lo = per - 200
    seg = x[lo * 512: (per + 60) * 512 + 2048]
    ref = O.spectrogram(seg, O.Config(n_fft=2048, hop=512, smoothing=0.8, output=O.OUT_F32_MAG))[0]
    assert_mag_close(got[lo + 150: per + 60], ref[150:260])              # 150 frames of warm-up: 0.8^150 ~ 3e-15
    two = engine.spectrogram(x, opts)                                     # automatic: the two-kernel path
    assert engine.last_kernel != "warp32x32x2s"
    assert np.all(np.abs(two - got) <= 2e-5 * np.abs(got) + 1e-7 * got.max(axis=-1, keepdims=True))


# ----------------------------------------------------------------------------- fused smoothing, n_fft 4096 / 1024 / 512 / 256
PS_KERNEL = {4096: "eo4096s", 1024: "p16s", 512: "p8s", 256: "p4s"}


@pytest.mark.parametrize("n_fft,hop_div,n_clips,frames,extra,align", [
    (1024, 4, 3, 333, 77, O.ALIGN_ANALYSER),     # zero history, odd frame count, partial last step
    (1024, 8, 2, 700, 0, O.ALIGN_VALID),         # hop n_fft / 8
    (1024, 4, 160, 130, 0, O.ALIGN_VALID),       # more tasks than CTAs: several tasks per CTA
    (512, 4, 5, 1001, 13, O.ALIGN_VALID),        # four pairs per warp; a clip cut into chained segments
    (512, 8, 150, 90, 0, O.ALIGN_ANALYSER),
    (256, 4, 2, 2500, 3, O.ALIGN_VALID),         # eight pairs per warp, long chains
    (256, 8, 149, 61, 0, O.ALIGN_ANALYSER),      # clips shorter than one round of the CTA's warps
    (4096, 4, 150, 40, 0, O.ALIGN_VALID),        # one frame per warp (even/odd kernel), several tasks per CTA
    (4096, 4, 3, 200, 37, O.ALIGN_ANALYSER),     # zero history, chained segments of few clips
    (4096, 8, 2, 150, 0, O.ALIGN_VALID),
    (4096, 2, 149, 9, 4, O.ALIGN_VALID),         # clips shorter than one round of the CTA's warps
    (1024, 2, 4, 500, 9, O.ALIGN_ANALYSER),      # hop n_fft / 2
    (256, 2, 151, 77, 0, O.ALIGN_VALID),
    (512, 4, 3, 1, 0, O.ALIGN_VALID),            # a single frame per clip: one lone frame A, every other lane group idle
    (1024, 4, 2, 2, 0, O.ALIGN_VALID),           # exactly one pair
    (256, 4, 2, 5, 7, O.ALIGN_ANALYSER),         # fewer frames than one step holds
])
def test_fused_smoothing_part_warp_kernels_match_the_oracle(engine, n_fft, hop_div, n_clips, frames, extra, align):
    hop = n_fft // hop_div
    clip_len = n_fft + (frames - 1) * hop + extra
    rng = np.random.default_rng(n_fft + n_clips)
    x = (0.05 * rng.standard_normal((n_clips, clip_len))).astype(np.float32)
    x += O.chirp(clip_len, 48000.0, 100.0, 15000.0, 0.3)[None, :]
    sel = np.unique(np.r_[0, min(1, n_clips - 1), n_clips // 2, n_clips - 1])
    cfg = O.Config(n_fft=n_fft, hop=hop, smoothing=0.8, align=align, output=O.OUT_F32_MAG)
    ref_mag = O.spectrogram(x[sel], cfg)
    for out in ("mag", "db", "u8", "rgba"):
        opts = sg.Options(fftSize=n_fft, hop=hop, output=out, smoothingTimeConstant=0.8, align=ALIGN[align])
        engine.set_kernel_variant(7)
        try:
            got = engine.spectrogram(x, opts)
        finally:
            engine.set_kernel_variant(0)
        assert engine.last_kernel == PS_KERNEL[n_fft]
        if out == "mag":
            assert_mag_close(got[sel], ref_mag)
        elif out == "db":
            assert_db_close(got[sel], ref_mag)
        elif out == "u8":
            got_u8 = got
            assert_bytes_close(got[sel], O.finish(ref_mag, O.Config(n_fft=n_fft, hop=hop, smoothing=0.8, align=align)))
        else:
            assert_rgba_is_lut_of(got, got_u8)


@pytest.mark.parametrize("n_fft", [4096, 1024, 512, 256])
def test_fused_smoothing_part_warp_automatic_selection_and_paths_agree(engine, n_fft):
    """From ~2/3 of the SMs in clips the fused kernel is the automatic choice; the two-kernel path (forced by a
    device-resident batch below the threshold) agrees with it to rounding, and both with the oracle."""
    import torch
    rng = np.random.default_rng(n_fft)
    hop = n_fft // 4
    clip_len = n_fft + (499 if n_fft <= 1024 else 149) * hop
    x = torch.from_numpy((0.2 * rng.standard_normal((200, clip_len))).astype(np.float32)).cuda()
    opts = sg.Options(fftSize=n_fft, hop=hop, output="mag", smoothingTimeConstant=0.75)
    frames = engine.num_frames(opts, clip_len)

    def run(n):
        out = torch.empty((n, frames, n_fft // 2), dtype=torch.float32, device="cuda")
        engine.spectrogram_device(x.data_ptr(), n, clip_len, clip_len, opts, out.data_ptr())
        engine.synchronize()
        return out.cpu().numpy()

    a = run(200)
    assert engine.last_kernel == PS_KERNEL[n_fft]
    b = run(3)
    assert engine.last_kernel != PS_KERNEL[n_fft]
    assert np.all(np.abs(b - a[:3]) <= 2e-5 * np.abs(a[:3]) + 1e-7 * a[:3].max(axis=-1, keepdims=True))
    ref = O.spectrogram(x[:2].cpu().numpy(), O.Config(n_fft=n_fft, hop=hop, smoothing=0.75, output=O.OUT_F32_MAG))
    assert_mag_close(a[:2], ref)


@pytest.mark.parametrize("n_fft,window", [(512, "hann"), (1024, "blackman")])
def test_fused_smoothing_at_hop_160(engine, n_fft, window):
    """The 16 kHz speech front-end hop (BASELINE config 2's shape) on the fused smoothing kernels."""
    rng = np.random.default_rng(160 + n_fft)
    n_clips, clip_len = 150, n_fft + 211 * 160 + 31
    x = (0.1 * rng.standard_normal((n_clips, clip_len))).astype(np.float32)
    w = {"hann": O.WINDOW_HANN, "blackman": O.WINDOW_BLACKMAN}[window]
    sel = [0, 77, 149]
    ref = O.spectrogram(x[sel], O.Config(n_fft=n_fft, hop=160, window=w, smoothing=0.6, output=O.OUT_F32_MAG))
    for out in ("mag", "u8"):
        engine.set_kernel_variant(7)
        try:
            got = engine.spectrogram(x, sg.Options(fftSize=n_fft, hop=160, window=window, output=out, smoothingTimeConstant=0.6))
        finally:
            engine.set_kernel_variant(0)
        assert engine.last_kernel == PS_KERNEL[n_fft]
        if out == "mag":
            assert_mag_close(got[sel], ref)
        else:
            assert_bytes_close(got[sel], O.finish(ref, O.Config(n_fft=n_fft, hop=160, window=w, smoothing=0.6)))


@pytest.mark.parametrize("n_fft,hop,n_clips,align", [
    (1024, 441, 150, O.ALIGN_ANALYSER),     # odd hop: 4-byte loads, zero history
    (512, 509, 3, O.ALIGN_VALID),           # hop just under the frame length (longer hops stay on the two-kernel path)
    (256, 100, 152, O.ALIGN_VALID),         # even hop that is no multiple of the lane group's 8 samples: 8-byte loads
    (1024, 200, 4, O.ALIGN_ANALYSER),
])
def test_fused_smoothing_part_warp_kernels_at_any_hop(engine, n_fft, hop, n_clips, align):
    rng = np.random.default_rng(n_fft + hop)
    clip_len = n_fft + 233 * hop + 5
    x = (0.1 * rng.standard_normal((n_clips, clip_len))).astype(np.float32)
    sel = np.unique(np.r_[0, n_clips // 2, n_clips - 1])
    ref = O.spectrogram(x[sel], O.Config(n_fft=n_fft, hop=hop, smoothing=0.7, align=align, output=O.OUT_F32_MAG))
    for out in ("mag", "u8"):
        engine.set_kernel_variant(7)
        try:
            got = engine.spectrogram(x, sg.Options(fftSize=n_fft, hop=hop, output=out, smoothingTimeConstant=0.7, align=ALIGN[align]))
        finally:
            engine.set_kernel_variant(0)
        assert engine.last_kernel == PS_KERNEL[n_fft]
        if out == "mag":
            assert_mag_close(got[sel], ref)
        else:
            assert_bytes_close(got[sel], O.finish(ref, O.Config(n_fft=n_fft, hop=hop, smoothing=0.7, align=align)))


@pytest.mark.parametrize("n_fft,hop,n_clips,frames,extra,align", [
    (8192, 2048, 150, 21, 0, O.ALIGN_VALID),        # two frame slots per CTA, more tasks than CTAs
    (8192, 2048, 3, 150, 77, O.ALIGN_ANALYSER),     # zero history, odd frame count, chained segments of few clips
    (8192, 441, 2, 200, 0, O.ALIGN_VALID),          # odd hop: 4-byte loads
    (8192, 4096, 301, 5, 3, O.ALIGN_VALID),         # clips of a few frames
    (4096, 441, 150, 37, 0, O.ALIGN_ANALYSER),      # n_fft 4096 where the even/odd kernel cannot load: four slots per CTA
    (4096, 735, 2, 301, 9, O.ALIGN_VALID),
])
def test_fused_smoothing_register_family_matches_the_oracle(engine, n_fft, hop, n_clips, frames, extra, align):
    clip_len = n_fft + (frames - 1) * hop + extra
    rng = np.random.default_rng(n_fft + n_clips + hop)
    x = (0.05 * rng.standard_normal((n_clips, clip_len))).astype(np.float32)
    x += O.chirp(clip_len, 48000.0, 100.0, 15000.0, 0.3)[None, :]
    sel = np.unique(np.r_[0, min(1, n_clips - 1), n_clips // 2, n_clips - 1])
    cfg = O.Config(n_fft=n_fft, hop=hop, smoothing=0.8, align=align, output=O.OUT_F32_MAG)
    ref_mag = O.spectrogram(x[sel], cfg)
    for out in ("mag", "db", "u8", "rgba"):
        opts = sg.Options(fftSize=n_fft, hop=hop, output=out, smoothingTimeConstant=0.8, align=ALIGN[align])
        engine.set_kernel_variant(7)
        try:
            got = engine.spectrogram(x, opts)
        finally:
            engine.set_kernel_variant(0)
        assert engine.last_kernel == "wregs"
        if out == "mag":
            assert_mag_close(got[sel], ref_mag)
        elif out == "db":
            assert_db_close(got[sel], ref_mag)
        elif out == "u8":
            got_u8 = got
            assert_bytes_close(got[sel], O.finish(ref_mag, O.Config(n_fft=n_fft, hop=hop, smoothing=0.8, align=align)))
        else:
            assert_rgba_is_lut_of(got, got_u8)


def test_fused_smoothing_register_family_non_finite_frames_reset_the_state(engine):
    rng = np.random.default_rng(81)
    n_fft, hop = 8192, 2048
    clip_len = n_fft + 40 * hop
    x = (0.2 * rng.standard_normal((2, clip_len))).astype(np.float32)
    x[0, n_fft + 10 * hop + 5] = np.nan        # four frames of clip 0 see the NaN
    x[1, n_fft + 23 * hop + 9] = np.inf
    with np.errstate(invalid="ignore", over="ignore"):
        ref = O.spectrogram(x, O.Config(n_fft=n_fft, hop=hop, smoothing=0.8, output=O.OUT_F32_MAG))
    engine.set_kernel_variant(7)
    try:
        got = engine.spectrogram(x, sg.Options(fftSize=n_fft, hop=hop, output="mag", smoothingTimeConstant=0.8))
    finally:
        engine.set_kernel_variant(0)
    assert engine.last_kernel == "wregs"
    for c in (0, 1):
        bad = np.all(ref[c] == 0, axis=1)
        assert bad.sum() == 4 and np.all(got[c][bad] == 0)
    assert_mag_close(got, ref)


@pytest.mark.parametrize("n_fft", [4096, 1024, 256])
def test_fused_smoothing_part_warp_non_finite_frames_reset_the_state(engine, n_fft):
    rng = np.random.default_rng(8)
    hop = n_fft // 4
    clip_len = n_fft + 90 * hop
    x = (0.2 * rng.standard_normal((2, clip_len))).astype(np.float32)
    x[0, n_fft + 20 * hop + 5] = np.nan        # four frames of clip 0 see the NaN
    x[1, n_fft + 33 * hop + 9] = np.inf
    cfg = O.Config(n_fft=n_fft, hop=hop, smoothing=0.8, output=O.OUT_F32_MAG)
    with np.errstate(invalid="ignore", over="ignore"):
        ref = O.spectrogram(x, cfg)
    engine.set_kernel_variant(7)
    try:
        got = engine.spectrogram(x, sg.Options(fftSize=n_fft, hop=hop, output="mag", smoothingTimeConstant=0.8))
    finally:
        engine.set_kernel_variant(0)
    assert engine.last_kernel == PS_KERNEL[n_fft]
    for c in (0, 1):
        bad = np.all(ref[c] == 0, axis=1)
        assert bad.sum() == 4 and np.all(got[c][bad] == 0)
    assert_mag_close(got, ref)


# ----------------------------------------------------------------------------- few clips: independent segments with warm-up
@pytest.mark.parametrize("n_fft,hop,n_clips,frames,tau,kernel", [
    (2048, 512, 8, 4000, 0.6, "warp32x32x2s"),
    (2048, 441, 5, 10000, 0.7, "warp32x32x2s"),      # odd hop
    (1024, 256, 4, 6000, 0.5, "p16s"),
    (256, 64, 2, 40000, 0.8, "p4s"),                 # two channels, the AnalyserNode default
    (4096, 1024, 6, 4000, 0.5, "eo4096s"),
    (8192, 2048, 10, 3000, 0.3, "wregs"),
])
def test_fused_smoothing_independent_segments_match_the_oracle(engine, n_fft, hop, n_clips, frames, tau, kernel):
    """Fewer clips than CTAs: every clip is cut into segments that warm up from a zero state over enough frames for
    tau^warm to vanish in float32 -- rows must be those of the sequential recurrence, also right behind a segment start."""
    import torch
    clip_len = n_fft + (frames - 1) * hop + 4        # (rows of whole 16-byte units: the n_fft 4096 kernel's loader wants them)
    g = torch.Generator(device="cuda").manual_seed(n_fft + n_clips)
    x = (torch.rand((n_clips, clip_len), device="cuda", generator=g) - 0.5).float() * 0.3
    x[:, clip_len // 3: clip_len // 3 + 40 * hop] *= 30.0            # a loud burst: its tail must survive the segment borders
    x[0, clip_len // 2: clip_len // 2 + (50 + n_fft // hop) * hop] = 0.0   # digital silence: rows decay geometrically (kept short
                                                                            # enough to stay clear of float32 underflow)
    torch.cuda.synchronize()                     # torch filled x on its own stream; the engine runs on another
    opts = sg.Options(fftSize=n_fft, hop=hop, output="mag", smoothingTimeConstant=tau)
    assert engine.num_frames(opts, clip_len) == frames
    out = torch.empty((n_clips, frames, n_fft // 2), dtype=torch.float32, device="cuda")
    engine.spectrogram_device(x.data_ptr(), n_clips, clip_len, clip_len, opts, out.data_ptr())
    engine.synchronize()
    assert engine.last_kernel == kernel
    sel = [0, n_clips - 1]
    ref = O.spectrogram(x[sel].cpu().numpy(), O.Config(n_fft=n_fft, hop=hop, smoothing=tau, output=O.OUT_F32_MAG))
    got = out[sel].cpu().numpy()
    assert_mag_close(got, ref)
    # the silent stretch: the recurrence decays geometrically and nothing but float32 rounding may separate the two
    quiet = slice(clip_len // 2 // hop + n_fft // hop + 2, clip_len // 2 // hop + 45)
    assert np.all(np.abs(got[0, quiet] - ref[0, quiet]) <= 2e-5 * ref[0, quiet] + 1e-30)
    out8 = torch.empty((n_clips, frames, n_fft // 2), dtype=torch.uint8, device="cuda")
    engine.spectrogram_device(x.data_ptr(), n_clips, clip_len, clip_len, sg.Options(fftSize=n_fft, hop=hop, output="u8", smoothingTimeConstant=tau),
                              out8.data_ptr())
    engine.synchronize()
    assert_bytes_close(out8[sel].cpu().numpy(), O.finish(ref, O.Config(n_fft=n_fft, hop=hop, smoothing=tau)))


def test_fused_smoothing_independent_segments_fall_back_for_long_memories(engine):
    """tau close to 1: the warm-up would be longer than a segment, the two-kernel path stays."""
    rng = np.random.default_rng(3)
    x = (0.1 * rng.standard_normal((4, 2048 + 1999 * 512))).astype(np.float32)
    got = engine.spectrogram(x, sg.Options(output="mag", smoothingTimeConstant=0.995))
    assert engine.last_kernel != "warp32x32x2s"
    ref = O.spectrogram(x[:1], O.Config(n_fft=2048, hop=512, smoothing=0.995, output=O.OUT_F32_MAG))
    assert_mag_close(got[:1], ref)
