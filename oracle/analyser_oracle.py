"""CPU oracle (float64, numpy) for the frame-producing hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``spectrogram_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs do, and only as the checker or the
timed CPU baseline.

PARITY UNPINNED.  The arithmetic of this path is not in the reference
repository: the reference configures a browser ``AnalyserNode``
(/root/reference/src/javascripts/UI/player.js:7-11) and polls it
(/root/reference/src/javascripts/3D/visualizer.js:346-368).  The implementation
is the host browser's Web Audio engine, which no lock file pins and which cannot
run here (no node, no browser).  The reference holds no tests, fixtures or golden
vectors.  This file therefore restates the published W3C Web Audio API
``AnalyserNode`` algorithm (section "FFT windowing and smoothing over time"),
with Chromium's ``RealtimeAnalyser`` arithmetic as the tie-breaker, and is pinned
only by closed-form known-answer tests (tests/test_oracle_kat.py) and by
cross-checks against scipy / torch.stft.

Steps restated (per ``getByteFrequencyData`` / ``getFloatFrequencyData`` call,
N = fftSize):
  1. time-domain block: the most recent N samples (ring starts zero filled)
  2. Blackman window, alpha = 0.16, periodic form
  3. DFT with 1/N scaling, bins 0..N/2-1 (Nyquist dropped)
  4. smoothing over time: X^[k] = tau*X^-1[k] + (1-tau)*|X[k]|; non-finite -> 0
  5. dB: Y[k] = 20*log10(X^[k])   (0 -> -inf)
  6. byte: clamp(floor(255/(maxDb-minDb) * (Y[k]-minDb)), 0, 255)
  7. attribute validation -> IndexSizeError
Colour map: /root/reference/src/bin/shaders/sonogram-vertex.shader:19-58 and
sonogram-fragment.shader:24-26, background 0.08
(/root/reference/src/javascripts/3D/visualizer.js:69).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

WINDOW_BLACKMAN = 0
WINDOW_HANN = 1
WINDOW_RECT = 2
WINDOW_CUSTOM = 3

OUT_U8 = 0
OUT_F32_DB = 1
OUT_RGBA8 = 2
OUT_F32_MAG = 3

ALIGN_VALID = 0      # frame t starts at t*hop; frames = 1 + (L-N)//hop
ALIGN_ANALYSER = 1   # frame t = the N samples ending at (t+1)*hop, zero history


class IndexSizeError(ValueError):
    """Web Audio ``IndexSizeError`` DOMException (step 7)."""


@dataclass
class Config:
    """Mirror of ``sg_stft_config`` (include/sgcore.h)."""

    n_fft: int = 2048                 # player.js:10
    hop: int = 512
    window: int = WINDOW_BLACKMAN     # AnalyserNode applies Blackman [SPEC step 2]
    output: int = OUT_U8              # visualizer.js:301 Uint8Array
    align: int = ALIGN_VALID
    min_db: float = -100.0            # AnalyserNode default
    max_db: float = -30.0             # AnalyserNode default
    smoothing: float = 0.0            # player.js:11 / visualizer.js:357
    custom_window: np.ndarray | None = field(default=None, repr=False)


# ----------------------------------------------------------------------------
# step 2: windows
# ----------------------------------------------------------------------------
def make_window(kind: int, n: int, custom: np.ndarray | None = None) -> np.ndarray:
    """float64 window table of length n (periodic form, as the spec states)."""
    i = np.arange(n, dtype=np.float64)
    x = i / n
    if kind == WINDOW_BLACKMAN:
        alpha = 0.16
        a0, a1, a2 = 0.5 * (1 - alpha), 0.5, 0.5 * alpha
        return a0 - a1 * np.cos(2 * np.pi * x) + a2 * np.cos(4 * np.pi * x)
    if kind == WINDOW_HANN:
        return 0.5 - 0.5 * np.cos(2 * np.pi * x)
    if kind == WINDOW_RECT:
        return np.ones(n, dtype=np.float64)
    if kind == WINDOW_CUSTOM:
        w = np.asarray(custom, dtype=np.float64)
        if w.shape != (n,):
            raise ValueError("custom window must have n_fft entries")
        return w
    raise ValueError(f"unknown window {kind}")


# ----------------------------------------------------------------------------
# step 7: validation
# ----------------------------------------------------------------------------
def validate_analyser_attrs(fft_size: int, min_db: float, max_db: float, tau: float) -> None:
    if fft_size < 32 or fft_size > 32768 or (fft_size & (fft_size - 1)) != 0:
        raise IndexSizeError(f"fftSize {fft_size} must be a power of two in [32, 32768]")
    if not (min_db < max_db):
        raise IndexSizeError("minDecibels must be < maxDecibels")
    if not (0.0 <= tau <= 1.0):
        raise IndexSizeError("smoothingTimeConstant must be in [0, 1]")


def validate_config(cfg: Config) -> None:
    """Batch path: generalised domain (any even n_fft whose half factors into 2,3,5)."""
    if cfg.n_fft < 4 or cfg.n_fft > 32768 or cfg.n_fft % 2:
        raise IndexSizeError(f"n_fft {cfg.n_fft} must be even and in [4, 32768]")
    if cfg.hop < 1:
        raise IndexSizeError("hop must be >= 1")
    if not (cfg.min_db < cfg.max_db):
        raise IndexSizeError("minDecibels must be < maxDecibels")
    if not (0.0 <= cfg.smoothing <= 1.0):
        raise IndexSizeError("smoothingTimeConstant must be in [0, 1]")


# ----------------------------------------------------------------------------
# step 1: framing
# ----------------------------------------------------------------------------
def num_frames(clip_len: int, n_fft: int, hop: int, align: int) -> int:
    if align == ALIGN_VALID:
        return 0 if clip_len < n_fft else 1 + (clip_len - n_fft) // hop
    return clip_len // hop


def frame_matrix(pcm: np.ndarray, n_fft: int, hop: int, align: int) -> np.ndarray:
    """[frames, n_fft] float64 view/copy of one clip's time-domain blocks."""
    x = np.asarray(pcm, dtype=np.float64)
    f = num_frames(x.shape[0], n_fft, hop, align)
    if f == 0:
        return np.zeros((0, n_fft), dtype=np.float64)
    if align == ALIGN_ANALYSER:
        # the ring starts zero filled: prepend n_fft zeros, frame t ends at (t+1)*hop
        x = np.concatenate([np.zeros(n_fft, dtype=np.float64), x])
        starts = (np.arange(f) + 1) * hop  # start in padded coords = (t+1)*hop - N + N
    else:
        starts = np.arange(f) * hop
    idx = starts[:, None] + np.arange(n_fft)[None, :]
    return x[idx]


# ----------------------------------------------------------------------------
# steps 2-3: window + DFT/N, magnitude
# ----------------------------------------------------------------------------
def magnitudes(frames: np.ndarray, window: np.ndarray, chromium_cast: bool = False) -> np.ndarray:
    """|X[k]|/N for k = 0..N/2-1, float64.  frames: [F, N]."""
    n = frames.shape[-1]
    w = window.astype(np.float32).astype(np.float64) if chromium_cast else window
    xw = frames * w[None, :]
    if chromium_cast:
        xw = xw.astype(np.float32).astype(np.float64)
    spec = np.fft.rfft(xw, axis=-1)[..., : n // 2]
    return np.abs(spec) / n


# ----------------------------------------------------------------------------
# step 4: smoothing over time (per bin first-order recurrence along frames)
# ----------------------------------------------------------------------------
def smooth(mag: np.ndarray, tau: float, state: np.ndarray | None = None) -> tuple[np.ndarray, np.ndarray]:
    """Returns (smoothed [F, bins], final state [bins]).  state=None -> zeros."""
    f, bins = mag.shape
    prev = np.zeros(bins, dtype=np.float64) if state is None else np.asarray(state, np.float64).copy()
    out = np.empty_like(mag)
    for t in range(f):
        with np.errstate(invalid="ignore", over="ignore"):
            cur = tau * prev + (1.0 - tau) * mag[t]
        cur = np.where(np.isfinite(cur), cur, 0.0)
        out[t] = cur
        prev = cur
    return out, prev


# ----------------------------------------------------------------------------
# steps 5-6: dB, byte
# ----------------------------------------------------------------------------
def to_db(mag: np.ndarray) -> np.ndarray:
    with np.errstate(divide="ignore"):
        return 20.0 * np.log10(mag)


def to_byte(db: np.ndarray, min_db: float, max_db: float) -> np.ndarray:
    scale = 255.0 / (max_db - min_db)
    with np.errstate(invalid="ignore"):
        v = scale * (db - min_db)
    v = np.where(np.isnan(v), 0.0, v)           # cannot occur for finite mags; defensive
    v = np.clip(v, 0.0, 255.0)
    return np.floor(v).astype(np.uint8)


# ----------------------------------------------------------------------------
# colour map: sonogram-vertex.shader:19-58, sonogram-fragment.shader:24-26
# ----------------------------------------------------------------------------
def _hsv_to_rgb(hue: float, sat: float, light: float) -> tuple[float, float, float]:
    chroma = light * sat
    hd = hue / 60.0
    x = chroma * (1.0 - abs(math.fmod(hd, 2.0) - 1.0))
    r = g = b = 0.0
    if hd < 1.0:
        r, g = chroma, x
    elif hd < 2.0:
        r, g = x, chroma
    elif hd < 3.0:
        g, b = chroma, x
    elif hd < 4.0:
        g, b = x, chroma
    elif hd < 5.0:
        r, b = x, chroma
    elif hd < 6.0:
        r, b = chroma, x
    # hd == 6.0 (byte 0) falls through every branch -> black, as in the shader
    return r, g, b


def colormap_lut(background: float = 0.08) -> np.ndarray:
    """256-entry RGBA8 table, uint8 [256, 4]: out = clamp(bg + a*HSV(360-360a,1,1)), alpha 1."""
    lut = np.zeros((256, 4), dtype=np.uint8)
    for b in range(256):
        a = b / 255.0
        hue = 360.0 - a * 360.0
        rgb = _hsv_to_rgb(hue, 1.0, 1.0)
        for c in range(3):
            v = min(max(background + a * rgb[c], 0.0), 1.0)
            lut[b, c] = int(math.floor(v * 255.0 + 0.5))
        lut[b, 3] = 255
    return lut


def colormap_lut_u32() -> np.ndarray:
    """Same table packed little-endian R | G<<8 | B<<16 | A<<24."""
    lut = colormap_lut().astype(np.uint32)
    return lut[:, 0] | (lut[:, 1] << 8) | (lut[:, 2] << 16) | (lut[:, 3] << 24)


# ----------------------------------------------------------------------------
# sonogram ring + view
# ----------------------------------------------------------------------------
class Ring:
    """The reference's history texture: ``bins x rows`` bytes, one row per frame written at ``yoffset``, then
    ``yoffset = (yoffset + 1) % rows`` (src/javascripts/3D/visualizer.js:60, 301-329, 399-416)."""

    def __init__(self, bins: int, rows: int = 256):
        self.bins, self.rows = bins, rows
        self.tex = np.zeros((rows, bins), dtype=np.uint8)     # initByteBuffer clears the texture (:317-329)
        self.yoffset = 0

    def append(self, frames) -> None:
        f = np.atleast_2d(np.asarray(frames, dtype=np.uint8))
        for row in f:                                         # texSubImage2D(0, yoffset, bins, 1) then advance
            self.tex[self.yoffset] = row
            self.yoffset = (self.yoffset + 1) % self.rows


def sonogram_view(tex: np.ndarray, yoffset: int, width: int, height: int, background: float = 0.08) -> np.ndarray:
    """Float64 restatement of the sonogram view, uint8 [height, width, 4].  Per pixel (px, py), texCoord
    u = (px+.5)/width, v = (py+.5)/height:
      s = 256^(u-1)                        src/bin/shaders/sonogram-fragment.shader:16, sonogram-vertex.shader:51
      t = v + yoffset/(rows-1)             sonogram-fragment.shader:17, 3D/visualizer.js:460
      a = LINEAR sample of alpha, CLAMP_TO_EDGE in s, REPEAT in t                 3D/visualizer.js:312-315
      rgb = HSV(360 - 360 a, 1, 1)         sonogram-vertex.shader:19-58 (per vertex there, per pixel here)
      fade = sqrt(cos((1-v) pi/2))         sonogram-fragment.shader:24
      out = clamp(background + a*fade*rgb), alpha 1; 8-bit round to nearest       sonogram-fragment.shader:26
    """
    rows, bins = tex.shape
    u = (np.arange(width) + 0.5) / width
    v = (np.arange(height) + 0.5) / height
    s = 256.0 ** (u - 1.0)
    t = v + yoffset / (rows - 1.0)
    t = t - np.floor(t)
    x = s * bins - 0.5
    y = t * rows - 0.5
    x0f, y0f = np.floor(x), np.floor(y)
    fx, fy = x - x0f, y - y0f
    x0 = np.clip(x0f.astype(np.int64), 0, bins - 1)
    x1 = np.clip(x0f.astype(np.int64) + 1, 0, bins - 1)
    y0 = np.mod(y0f.astype(np.int64), rows)
    y1 = np.mod(y0 + 1, rows)
    tx = tex.astype(np.float64)
    top = tx[y0][:, x0] * (1 - fx)[None, :] + tx[y0][:, x1] * fx[None, :]
    bot = tx[y1][:, x0] * (1 - fx)[None, :] + tx[y1][:, x1] * fx[None, :]
    a = (top * (1 - fy)[:, None] + bot * fy[:, None]) / 255.0
    hd = (360.0 - 360.0 * a) / 60.0
    xx = 1.0 - np.abs(np.mod(hd, 2.0) - 1.0)
    r = np.select([hd < 1, hd < 2, hd < 3, hd < 4, hd < 5, hd < 6], [1.0, xx, 0.0, 0.0, xx, 1.0], 0.0)
    g = np.select([hd < 1, hd < 2, hd < 3, hd < 4, hd < 5, hd < 6], [xx, 1.0, 1.0, xx, 0.0, 0.0], 0.0)
    b = np.select([hd < 1, hd < 2, hd < 3, hd < 4, hd < 5, hd < 6], [0.0, 0.0, xx, 1.0, 1.0, xx], 0.0)
    fade = np.sqrt(np.maximum(np.cos((1.0 - v) * 0.5 * math.pi), 0.0))
    k = a * fade[:, None]
    out = np.empty((height, width, 4), dtype=np.uint8)
    for c, ch in enumerate((r, g, b)):
        out[..., c] = np.floor(np.clip(background + k * ch, 0.0, 1.0) * 255.0 + 0.5).astype(np.uint8)
    out[..., 3] = 255
    return out


# ----------------------------------------------------------------------------
# whole path, batched: [clips, clip_len] -> [clips, frames, bins]
# ----------------------------------------------------------------------------
def spectrogram(pcm: np.ndarray, cfg: Config, chromium_cast: bool = False):
    """Float64 truth for ``sg_stft_batch``.  Smoothing state starts at zero per clip."""
    validate_config(cfg)
    x = np.atleast_2d(np.asarray(pcm))
    n_clips, clip_len = x.shape
    bins = cfg.n_fft // 2
    f = num_frames(clip_len, cfg.n_fft, cfg.hop, cfg.align)
    w = make_window(cfg.window, cfg.n_fft, cfg.custom_window)
    mags = np.empty((n_clips, f, bins), dtype=np.float64)
    for c in range(n_clips):
        with np.errstate(invalid="ignore", over="ignore"):
            m = magnitudes(frame_matrix(x[c], cfg.n_fft, cfg.hop, cfg.align), w, chromium_cast)
        if cfg.smoothing > 0.0:
            m, _ = smooth(m, cfg.smoothing)      # [SPEC] step 4: a non-finite X^ is set to 0 (the state restarts)
        else:
            m = np.where(np.isfinite(m), m, 0.0)
        mags[c] = m
    return finish(mags, cfg)


def finish(mags: np.ndarray, cfg: Config):
    if cfg.output == OUT_F32_MAG:
        return mags
    db = to_db(mags)
    if cfg.output == OUT_F32_DB:
        return db
    by = to_byte(db, cfg.min_db, cfg.max_db)
    if cfg.output == OUT_U8:
        return by
    if cfg.output == OUT_RGBA8:
        return colormap_lut()[by]
    raise ValueError("unknown output kind")


# ----------------------------------------------------------------------------
# AnalyserNode-shaped streaming oracle (same surface as visualizer.js uses)
# ----------------------------------------------------------------------------
class AnalyserOracle:
    """push() emulates the audio render thread feeding the ring (multiples of the
    128-frame render quantum in a browser; any length here)."""

    MAX_FFT = 32768

    def __init__(self, fft_size: int = 2048):
        self._fft_size = 2048
        self.minDecibels = -100.0
        self.maxDecibels = -30.0
        self.smoothingTimeConstant = 0.8
        self._ring = np.zeros(self.MAX_FFT, dtype=np.float64)
        self._write = 0
        self._dirty = True
        self._state = np.zeros(1024, dtype=np.float64)
        self.fftSize = fft_size

    @property
    def fftSize(self) -> int:
        return self._fft_size

    @fftSize.setter
    def fftSize(self, n: int) -> None:
        validate_analyser_attrs(n, -1.0, 0.0, 0.0)
        if n != self._fft_size or self._state.shape[0] != n // 2:
            self._fft_size = n
            self._state = np.zeros(n // 2, dtype=np.float64)
            self._dirty = True

    @property
    def frequencyBinCount(self) -> int:
        return self._fft_size // 2

    def push(self, samples) -> None:
        s = np.asarray(samples, dtype=np.float64).ravel()
        if s.size == 0:
            return
        if s.size >= self.MAX_FFT:
            s = s[-self.MAX_FFT:]
        idx = (self._write + np.arange(s.size)) % self.MAX_FFT
        self._ring[idx] = s
        self._write = int((self._write + s.size) % self.MAX_FFT)
        self._dirty = True

    def _block(self) -> np.ndarray:
        n = self._fft_size
        idx = (self._write - n + np.arange(n)) % self.MAX_FFT
        return self._ring[idx]

    def _analyse(self) -> None:
        validate_analyser_attrs(self._fft_size, self.minDecibels, self.maxDecibels,
                                self.smoothingTimeConstant)
        if not self._dirty:
            return  # two calls within one render quantum return the same data
        w = make_window(WINDOW_BLACKMAN, self._fft_size)
        m = magnitudes(self._block()[None, :], w)
        m = np.where(np.isfinite(m), m, 0.0)
        sm, st = smooth(m, float(self.smoothingTimeConstant), self._state)
        self._state = st
        self._dirty = False

    def getFloatFrequencyData(self, dst: np.ndarray) -> None:
        self._analyse()
        n = min(dst.shape[0], self.frequencyBinCount)
        dst[:n] = to_db(self._state[:n])

    def getByteFrequencyData(self, dst: np.ndarray) -> None:
        self._analyse()
        n = min(dst.shape[0], self.frequencyBinCount)
        dst[:n] = to_byte(to_db(self._state[:n]), self.minDecibels, self.maxDecibels)

    def getFloatTimeDomainData(self, dst: np.ndarray) -> None:
        n = min(dst.shape[0], self._fft_size)
        dst[:n] = self._block()[:n]

    def getByteTimeDomainData(self, dst: np.ndarray) -> None:
        n = min(dst.shape[0], self._fft_size)
        dst[:n] = time_domain_byte(self._block()[:n])


def time_domain_byte(x: np.ndarray) -> np.ndarray:
    """b = clamp(floor(128*(1+x)), 0, 255)  [SPEC getByteTimeDomainData]."""
    v = np.clip(128.0 * (1.0 + np.asarray(x, np.float64)), 0.0, 255.0)
    v = np.where(np.isnan(v), 0.0, v)
    return np.floor(v).astype(np.uint8)


# ----------------------------------------------------------------------------
# synthetic inputs of SURVEY.md section 8(d)
# ----------------------------------------------------------------------------
def chirp(n: int, sr: float, f0: float, f1: float, amp: float) -> np.ndarray:
    """Linear chirp f0->f1 over n samples (config 1: 441000, 44100, 20, 20000, 0.5)."""
    t = np.arange(n, dtype=np.float64) / sr
    dur = n / sr
    phase = 2 * np.pi * (f0 * t + 0.5 * (f1 - f0) / dur * t * t)
    return (amp * np.sin(phase)).astype(np.float32)


def band_noise(n: int, sr: float, lo: float, hi: float, sigma: float, seed: int) -> np.ndarray:
    """Band-limited Gaussian noise (config 2), brick-wall filtered in the frequency domain."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal(n)
    spec = np.fft.rfft(x)
    f = np.fft.rfftfreq(n, 1.0 / sr)
    spec[(f < lo) | (f > hi)] = 0.0
    y = np.fft.irfft(spec, n)
    y *= sigma / max(y.std(), 1e-30)
    return y.astype(np.float32)
