"""PCM ingestion in front of the path: WAVE header walk, sample conversion, speakers
down-mix, and the fused ingest -> frame pipeline.  The reference's side of this is
``context.decodeAudioData`` (util/util.js:9-17) and the AnalyserNode's mono down-mix."""
import io
import struct
import wave

import numpy as np
import pytest

from oracle import analyser_oracle as ao
from oracle import pcm_oracle as po

FMT_NAMES = {po.U8: "u8", po.S16: "s16", po.S24: "s24", po.S32: "s32", po.F32: "f32"}


def random_pcm(fmt, channels, frames, seed):
    """interleaved bytes of `frames` sample frames, full-range values incl. the extremes"""
    rng = np.random.default_rng(seed)
    n = channels * frames
    if fmt == po.U8:
        v = rng.integers(0, 256, n, dtype=np.int64)
        v[:2] = (0, 255)
        return v.astype(np.uint8).tobytes()
    if fmt == po.S16:
        v = rng.integers(-32768, 32768, n, dtype=np.int64)
        v[:2] = (-32768, 32767)
        return v.astype("<i2").tobytes()
    if fmt == po.S24:
        v = rng.integers(-(1 << 23), 1 << 23, n, dtype=np.int64)
        v[:2] = (-(1 << 23), (1 << 23) - 1)
        u = (v & 0xFFFFFF).astype(np.uint32)
        return np.stack([u & 255, (u >> 8) & 255, (u >> 16) & 255], axis=1).astype(np.uint8).tobytes()
    if fmt == po.S32:
        v = rng.integers(-(1 << 31), 1 << 31, n, dtype=np.int64)
        v[:2] = (-(1 << 31), (1 << 31) - 1)
        return v.astype("<i4").tobytes()
    v = rng.standard_normal(n).astype("<f4")
    return v.tobytes()


def wav_file(raw, fmt, channels, rate, extensible=False, junk=False):
    tag = 3 if fmt == po.F32 else 1
    sb = po.SAMPLE_BYTES[fmt]
    body = struct.pack("<HHIIHH", 0xFFFE if extensible else tag, channels, rate, rate * channels * sb, channels * sb, sb * 8)
    if extensible:
        body += struct.pack("<HHIH", 22, sb * 8, 0, tag) + bytes.fromhex("000000001000800000aa00389b71")
    chunks = b""
    if junk:
        chunks += b"LIST" + struct.pack("<I", 5) + b"abcde" + b"\0"      # odd-sized chunk + pad byte
    chunks += b"fmt " + struct.pack("<I", len(body)) + body
    chunks += b"data" + struct.pack("<I", len(raw)) + raw
    return b"RIFF" + struct.pack("<I", 4 + len(chunks)) + b"WAVE" + chunks


# ------------------------------------------------------------------ CPU: oracle known answers ----
def test_oracle_integer_scaling_known_answers():
    assert po.decode_interleaved(bytes([0, 128, 255]), po.U8, 1)[0].tolist() == [-1.0, 0.0, 127 / 128]
    assert po.decode_interleaved(struct.pack("<3h", -32768, 0, 32767), po.S16, 1)[0].tolist() == [-1.0, 0.0, 32767 / 32768]
    s24 = bytes([0, 0, 0x80, 0xFF, 0xFF, 0x7F, 0xFF, 0xFF, 0xFF])
    assert po.decode_interleaved(s24, po.S24, 1)[0].tolist() == [-1.0, (2 ** 23 - 1) / 2 ** 23, -1 / 2 ** 23]
    assert po.decode_interleaved(struct.pack("<2i", -2 ** 31, 2 ** 30), po.S32, 1)[0].tolist() == [-1.0, 0.5]
    assert po.decode_interleaved(struct.pack("<2f", 1.5, -0.25), po.F32, 1)[0].tolist() == [1.5, -0.25]
    # interleaving: frame-major in, channel-major out
    assert po.decode_interleaved(struct.pack("<4h", 1, 2, 3, 4), po.S16, 2).tolist() == [[1 / 32768, 3 / 32768], [2 / 32768, 4 / 32768]]


def test_oracle_speakers_downmix_known_answers():
    one = np.ones((1, 3))
    assert po.downmix_speakers(one).tolist() == [1, 1, 1]
    st = np.array([[1.0, 0.0], [0.0, -1.0]])
    assert po.downmix_speakers(st).tolist() == [0.5, -0.5]
    quad = np.array([[1.0], [2.0], [3.0], [4.0]])
    assert po.downmix_speakers(quad).tolist() == [2.5]
    five1 = np.array([[1.0], [1.0], [0.25], [100.0], [0.5], [-0.5]])          # LFE is dropped
    np.testing.assert_allclose(po.downmix_speakers(five1), [np.sqrt(2.0) + 0.25], rtol=1e-15)
    three = np.array([[7.0], [8.0], [9.0]])                                    # no layout: discrete, channel 0
    assert po.downmix_speakers(three).tolist() == [7.0]


@pytest.mark.parametrize("width,fmt", [(1, po.U8), (2, po.S16), (3, po.S24), (4, po.S32)])
def test_oracle_agrees_with_the_wave_module_and_scipy(width, fmt):
    from scipy.io import wavfile

    raw = random_pcm(fmt, 2, 333, seed=width)
    buf = io.BytesIO()
    with wave.open(buf, "wb") as w:
        w.setnchannels(2); w.setsampwidth(width); w.setframerate(22050); w.writeframes(raw)
    data = buf.getvalue()
    info = po.wav_parse(data)
    assert info == {"format": fmt, "channels": 2, "sample_rate": 22050, "frames": 333, "data_offset": 44}
    rate, arr = wavfile.read(io.BytesIO(data))
    assert rate == 22050
    full = {1: 128.0, 2: 32768.0, 3: 2147483648.0, 4: 2147483648.0}[width]   # scipy left-justifies 24 bit in int32
    ref = (arr.astype(np.float64) - (128.0 if width == 1 else 0.0)) / full
    if width == 4:
        ref = arr.astype(np.float32).astype(np.float64) / full
    got = po.decode_interleaved(data[44:], fmt, 2)
    np.testing.assert_array_equal(got, ref.T)


def test_wav_parse_matches_the_oracle_and_rejects_what_it_cannot_decode():
    import spectrogram_b200 as sg

    raw = random_pcm(po.S24, 6, 41, seed=9)
    for ext in (False, True):
        for junk in (False, True):
            data = wav_file(raw, po.S24, 6, 48000, ext, junk)
            want = po.wav_parse(data)
            got = sg.wav_info(data)
            assert {k: getattr(got, k) for k in want} == want
    f = sg.wav_info(wav_file(random_pcm(po.F32, 1, 8, 0), po.F32, 1, 8000))
    assert (f.format, f.frames) == (po.F32, 8)
    # truncated data chunk: frames follow the bytes that are there
    cut = wav_file(raw, po.S24, 6, 48000)[:-18 * 3 - 5]
    assert sg.wav_info(cut).frames == po.wav_parse(cut)["frames"] == 41 - 4
    for bad in (b"", b"RIFF\0\0\0\0WAVX", wav_file(raw, po.S24, 6, 48000)[:30],
                b"RIFF\0\0\0\0WAVE" + b"fmt " + struct.pack("<IHHIIHH", 16, 85, 2, 44100, 0, 1, 0) + b"data\0\0\0\0"):   # mp3-in-wav
        with pytest.raises(TypeError):
            sg.wav_info(bad)
        with pytest.raises(ValueError):
            po.wav_parse(bad)


def test_abi_rejects_bad_pcm_descriptions():
    import ctypes as C
    from spectrogram_b200 import _lib as L

    lib = L.load()
    assert [lib.sg_pcm_sample_bytes(f) for f in range(-1, 6)] == [0, 1, 2, 3, 4, 4, 0]
    ok = L.PcmInfo(L.PCM_S16, 2, 48000, 10, 0)
    assert lib.sg_pcm_num_planes(C.byref(ok), L.PCM_MONO_MIX) == 1
    assert lib.sg_pcm_num_planes(C.byref(ok), L.PCM_PLANAR) == 2
    for bad in (L.PcmInfo(9, 2, 0, 10, 0), L.PcmInfo(L.PCM_S16, 0, 0, 10, 0), L.PcmInfo(L.PCM_S16, 33, 0, 10, 0),
                L.PcmInfo(L.PCM_S16, 2, 0, -1, 0)):
        assert lib.sg_pcm_num_planes(C.byref(bad), L.PCM_PLANAR) == L.SG_ERR_INVALID_ARG
    assert lib.sg_pcm_num_planes(C.byref(ok), 7) == L.SG_ERR_INVALID_ARG


# ------------------------------------------------------------------ GPU parity -------------------
@pytest.mark.gpu
@pytest.mark.parametrize("fmt", [po.U8, po.S16, po.S24, po.S32, po.F32])
@pytest.mark.parametrize("channels", [1, 2, 3, 4, 6, 32])
def test_planes_are_bit_exact_and_mono_mix_follows_the_speakers_rules(engine, fmt, channels):
    frames = 20011                                  # several tiles for every frame size, ragged tail
    raw = random_pcm(fmt, channels, frames, seed=fmt * 100 + channels)
    want = po.decode_interleaved(raw, fmt, channels)
    buf = engine.decode_pcm(raw, FMT_NAMES[fmt], channels, 48000)
    assert (buf.numberOfChannels, buf.length, buf.sampleRate) == (channels, frames, 48000)
    np.testing.assert_array_equal(buf.planes, want.astype(np.float32))          # conversion is exact in float32
    mono = engine.decode_pcm(raw, FMT_NAMES[fmt], channels, 48000, mix=True)
    assert mono.numberOfChannels == 1
    ref = po.downmix_speakers(want)
    scale = max(1.0, float(np.abs(want).max()))
    np.testing.assert_allclose(mono.getChannelData(0), ref, rtol=0, atol=4e-7 * scale)
    if channels <= 2:                                # x and 0.5(L+R) have one rounding: exact
        np.testing.assert_array_equal(mono.getChannelData(0), ref.astype(np.float32))


@pytest.mark.gpu
@pytest.mark.parametrize("channels,mix", [(1, True), (2, True), (2, False)])
@pytest.mark.parametrize("frames,n_clips", [(20000, 1), (4096, 5), (8, 3)])
def test_sixteen_bit_vector_path_is_bit_exact(engine, channels, mix, frames, n_clips):
    """16-byte friendly 16-bit mono/stereo takes the one-word-per-thread kernel: same results as the oracle"""
    raw = np.frombuffer(random_pcm(po.S16, channels, frames * n_clips, seed=frames + channels), dtype="<i2").copy()
    assert raw.ctypes.data % 16 == 0
    got = engine.decode_pcm(raw, "s16", channels, 48000, n_clips=n_clips, mix=mix).planes.reshape(n_clips, -1, frames)
    for c in range(n_clips):
        want = po.decode_interleaved(raw[c * frames * channels:(c + 1) * frames * channels].tobytes(), po.S16, channels)
        ref = po.downmix_speakers(want)[None] if mix else want
        np.testing.assert_array_equal(got[c], ref.astype(np.float32))


@pytest.mark.gpu
def test_ingest_handles_unaligned_sources_many_clips_and_empty_input(engine):
    raw = random_pcm(po.S24, 2, 3 * 5000, seed=5)   # three clips of 5000 frames, 6-byte frames
    pad = np.frombuffer(b"\x55" * 3 + raw, dtype=np.uint8)[3:]     # source pointer at an odd address
    got = engine.decode_pcm(pad, "s24", 2, 44100, n_clips=3)
    want = po.decode_interleaved(raw, po.S24, 2).astype(np.float32)
    assert got.planes.shape == (3, 2, 5000)
    for c in range(3):
        np.testing.assert_array_equal(got.planes[c], want[:, 5000 * c:5000 * (c + 1)])
    assert engine.decode_pcm(b"", "s16", 2, 44100).length == 0


@pytest.mark.gpu
def test_decode_audio_data_of_a_wave_file(engine):
    raw = random_pcm(po.S16, 2, 44100, seed=3)
    buf = engine.decode_audio_data(wav_file(raw, po.S16, 2, 44100, junk=True))
    assert (buf.numberOfChannels, buf.length, buf.sampleRate, buf.duration) == (2, 44100, 44100, 1.0)
    want = po.decode_interleaved(raw, po.S16, 2)
    np.testing.assert_array_equal(buf.getChannelData(1), want[1].astype(np.float32))
    with pytest.raises(ValueError):
        buf.getChannelData(2)


@pytest.mark.gpu
@pytest.mark.parametrize("fmt,channels,mix", [(po.S16, 2, True), (po.S16, 2, False), (po.S24, 6, True), (po.F32, 1, True), (po.U8, 4, False)])
def test_fused_pcm_pipeline_equals_ingest_then_batch(engine, fmt, channels, mix):
    """sg_stft_pcm == sg_pcm_ingest followed by sg_stft_batch, and the bytes follow the oracle."""
    import spectrogram_b200 as sg

    n_clips, frames = 3, 30000
    raw = random_pcm(fmt, channels, n_clips * frames, seed=17)
    opts = sg.Options(fftSize=1024, hop=256)
    got = engine.spectrogram_pcm(raw, FMT_NAMES[fmt], channels, n_clips=n_clips, mix=mix, opts=opts)
    planes = engine.decode_pcm(raw, FMT_NAMES[fmt], channels, 48000, n_clips=n_clips, mix=mix).planes
    two = engine.spectrogram(planes.reshape(-1, frames), opts)
    np.testing.assert_array_equal(got.reshape(two.shape), two)
    cfg = ao.Config(n_fft=1024, hop=256)
    want = ao.spectrogram(planes.reshape(-1, frames)[0].astype(np.float64), cfg)[0]
    d = np.abs(got.reshape(two.shape)[0].astype(int) - want.astype(int))
    assert d.max() <= 1


@pytest.mark.gpu
def test_fused_pcm_pipeline_on_one_long_stereo_clip(engine):
    """a clip larger than the pipeline's chunk: frame-range chunks, one output run per plane"""
    import spectrogram_b200 as sg

    frames = 12_000_000                              # 48 MB of s16 stereo: two chunks
    rng = np.random.default_rng(2)
    raw = (rng.standard_normal(2 * frames) * 3000).astype("<i2")
    opts = sg.Options(fftSize=512, hop=160, output="db")
    got = engine.spectrogram_pcm(raw, "s16", 2, mix=False, opts=opts)
    planes = engine.decode_pcm(raw, "s16", 2, 16000).planes
    two = engine.spectrogram(planes, opts)
    assert got.shape == (1, 2) + two.shape[1:]
    np.testing.assert_array_equal(got[0], two)


@pytest.mark.gpu
def test_fused_pcm_pipeline_with_smoothing_keeps_one_state_per_plane(engine):
    """tau > 0: every plane of every clip carries its own recurrence state through the pipeline"""
    import spectrogram_b200 as sg

    n_clips, frames = 2, 20000
    raw = random_pcm(po.S16, 2, n_clips * frames, seed=23)
    opts = sg.Options(fftSize=512, hop=128, smoothingTimeConstant=0.8, output="db")
    got = engine.spectrogram_pcm(raw, "s16", 2, n_clips=n_clips, mix=False, opts=opts)
    planes = engine.decode_pcm(raw, "s16", 2, 48000, n_clips=n_clips).planes
    two = engine.spectrogram(planes.reshape(-1, frames), opts)
    np.testing.assert_array_equal(got.reshape(two.shape), two)
