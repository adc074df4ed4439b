"""the 256-row byte ring (3D/visualizer.js:399-416) and the headless sonogram view
(bin/shaders/sonogram-fragment.shader:14-27, sonogram-vertex.shader:19-58).  Oracle known answers run on CPU;
the CUDA path (sg_ring_* through the C ABI) is compared with the oracle under ``-m gpu``."""
import math

import numpy as np
import pytest

from oracle import analyser_oracle as O


def test_ring_wraps_like_the_reference_texture():
    r = O.Ring(bins=4, rows=3)
    rows = np.arange(20, dtype=np.uint8).reshape(5, 4)
    r.append(rows[:2])
    assert r.yoffset == 2 and np.array_equal(r.tex[:2], rows[:2]) and not r.tex[2].any()
    r.append(rows[2:])
    assert r.yoffset == 5 % 3
    assert np.array_equal(r.tex, np.stack([rows[3], rows[4], rows[2]]))   # row t lives at t % rows


def test_view_known_answers():
    # full-scale texture: a = 1 -> hue 0 -> pure red; out = 0.08 + fade * (1, 0, 0)
    img = O.sonogram_view(np.full((256, 1024), 255, np.uint8), 0, 8, 1)
    fade = math.sqrt(math.cos(0.25 * math.pi))
    assert np.all(img[0, :, 0] == int(math.floor(min(0.08 + fade, 1.0) * 255 + 0.5)))
    assert np.all(img[0, :, 1] == 20) and np.all(img[0, :, 2] == 20) and np.all(img[0, :, 3] == 255)
    # silent texture: the background colour everywhere (the shader's HSV ladder yields black at a = 0)
    img = O.sonogram_view(np.zeros((256, 1024), np.uint8), 7, 5, 4)
    assert np.all(img[..., :3] == 20)
    # the edge fade reaches zero at v -> 0: the first row of a tall image is (almost) background
    img = O.sonogram_view(np.full((256, 64), 255, np.uint8), 0, 4, 4096)
    assert img[0, 0, 0] < 30 and img[-1, 0, 0] == 255


def test_view_log_frequency_axis():
    # one bright bin: it lands at u = 1 + log256(s) on the picture's x axis (sonogram-fragment.shader:16)
    bins, width = 1024, 2048
    tex = np.zeros((256, bins), np.uint8)
    k = 100
    tex[:, k] = 255
    img = O.sonogram_view(tex, 0, width, 2)
    px = int(np.argmax(img[1, :, 0].astype(int) + img[1, :, 1] + img[1, :, 2]))
    u_expected = 1.0 + math.log((k + 0.5) / bins) / math.log(256.0)
    assert abs((px + 0.5) / width - u_expected) < 2.0 / width


@pytest.mark.gpu
def test_ring_and_view_match_the_oracle(engine):
    import spectrogram_b200 as sg
    rng = np.random.default_rng(5)
    for bins, rows, n in ((1024, 256, 300), (512, 256, 100), (200, 17, 40), (1, 2, 5)):
        ring = sg.SonogramRing(bins, rows, engine=engine)
        ref = O.Ring(bins, rows)
        frames = rng.integers(0, 256, size=(n, bins), dtype=np.uint8)
        frames[: n // 3] = (np.linspace(0, 255, bins)[None, :] * rng.random((n // 3, 1))).astype(np.uint8)
        for lo, hi in ((0, 1), (1, n // 2), (n // 2, n)):      # single rows and multi-row appends, with wrap
            ring.append(frames[lo:hi])
            ref.append(frames[lo:hi])
            assert ring.yoffset == ref.yoffset
        assert np.array_equal(ring.texture(), ref.tex)
        for w, h in ((640, 256), (33, 7), (1, 1)):
            got = ring.view(w, h)
            want = O.sonogram_view(ref.tex, ref.yoffset, w, h)
            assert got.shape == want.shape
            d = np.abs(got.astype(np.int32) - want.astype(np.int32))
            assert d.max() <= 1, f"{bins}x{rows} {w}x{h}: max {d.max()} LSB"
            assert (d != 0).mean() < 0.02
        assert engine.last_kernel == "sonogram_view"
        ring.reset()
        assert ring.yoffset == 0 and not ring.texture().any()
        ring.close()


@pytest.mark.gpu
def test_ring_fed_by_the_stream_bank(engine):
    """The reference's loop: getByteFrequencyData -> texSubImage2D row -> draw (visualizer.js:346-416), headless."""
    import spectrogram_b200 as sg
    opts = sg.Options(fftSize=2048, hop=512, output="u8")
    x = O.chirp(44100, 44100.0, 20.0, 20000.0, 0.5)
    by = engine.spectrogram(x, opts).reshape(-1, 1024)      # [frames, bins]
    ring = sg.SonogramRing(1024, 256, engine=engine)
    ring.append(by)
    ref = O.Ring(1024, 256)
    ref.append(by)          # byte parity of the frames themselves is test_gpu_parity's job (+-1 LSB)
    got, want = ring.view(512, 256), O.sonogram_view(ref.tex, ref.yoffset, 512, 256)
    assert ring.yoffset == by.shape[0] % 256
    assert np.abs(got.astype(np.int32) - want.astype(np.int32)).max() <= 1
    assert got[..., :3].max() > 100      # the chirp is visible
    ring.close()
