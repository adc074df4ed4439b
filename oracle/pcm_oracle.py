"""CPU oracle for PCM ingestion: the step in front of the frame path.

TEST INFRASTRUCTURE ONLY (same rule as analyser_oracle.py: nothing under ``spectrogram_b200/`` imports it).

PARITY UNPINNED.  The reference gives an encoded file to the browser,
``context.decodeAudioData(request.response, buffer => callback(buffer))``
(/root/reference/src/javascripts/util/util.js:9-17), plays the AudioBuffer through a buffer source
(/root/reference/src/javascripts/UI/player.js:110-115,154-170) into the AnalyserNode
(UI/player.js:25), which down-mixes whatever channel count arrives to mono.  Decoder and mixer are the
browser's.  Restated here, for uncompressed PCM only:
  * RIFF/WAVE container walk (Microsoft/IBM "Multimedia Programming Interface and Data Specifications
    1.0" + WAVE_FORMAT_EXTENSIBLE): fmt chunk -> encoding, channels, rate; data chunk -> samples
  * integer -> float: u8 (v-128)/128, s16 v/2^15, s24 v/2^23, s32 v/2^31 (the scaling FFmpeg's sample
    format conversion uses; Chromium decodes through FFmpeg), float32 as is
  * [SPEC] Web Audio "Up-mixing and down-mixing", speakers interpretation, to mono:
      1: x   2: 0.5(L+R)   4: 0.25(L+R+SL+SR)   6: sqrt(1/2)(L+R) + C + 0.5(SL+SR)   else: channel 0
Pinned by known answers and by two independent decoders of the same bytes: the standard library's
``wave`` module and ``scipy.io.wavfile`` (tests/test_pcm_ingest.py).
"""
from __future__ import annotations

import struct

import numpy as np

U8, S16, S24, S32, F32 = 0, 1, 2, 3, 4
SAMPLE_BYTES = {U8: 1, S16: 2, S24: 3, S32: 4, F32: 4}


def decode_interleaved(raw: bytes, fmt: int, channels: int) -> np.ndarray:
    """bytes of interleaved samples -> float64 [channels][frames]."""
    b = np.frombuffer(raw, dtype=np.uint8)
    sb = SAMPLE_BYTES[fmt]
    n = b.size // (sb * channels) * channels
    b = b[:n * sb]
    if fmt == U8:
        x = (b.astype(np.float64) - 128.0) / 128.0
    elif fmt == S16:
        x = b.view("<i2").astype(np.float64) / 32768.0
    elif fmt == S24:
        t = b.reshape(-1, 3).astype(np.int64)
        v = t[:, 0] | (t[:, 1] << 8) | (t[:, 2] << 16)
        v = np.where(v >= 1 << 23, v - (1 << 24), v)
        x = v.astype(np.float64) / 8388608.0
    elif fmt == S32:
        # the int -> float32 rounding is part of the restated conversion
        x = b.view("<i4").astype(np.float32).astype(np.float64) / 2147483648.0
    elif fmt == F32:
        x = b.view("<f4").astype(np.float64)
    else:
        raise TypeError("bad format")
    return x.reshape(-1, channels).T.copy()


def downmix_speakers(planes: np.ndarray) -> np.ndarray:
    """[SPEC] speakers down-mix of [channels][frames] to mono [frames] (float64)."""
    c = planes.shape[0]
    if c == 1:
        return planes[0].copy()
    if c == 2:
        return 0.5 * (planes[0] + planes[1])
    if c == 4:
        return 0.25 * (planes[0] + planes[1] + planes[2] + planes[3])
    if c == 6:
        return np.sqrt(0.5) * (planes[0] + planes[1]) + planes[2] + 0.5 * (planes[4] + planes[5])
    return planes[0].copy()


def wav_parse(data: bytes) -> dict:
    """RIFF/WAVE walk -> {format, channels, sample_rate, frames, data_offset}; ValueError otherwise."""
    if len(data) < 12 or data[:4] != b"RIFF" or data[8:12] != b"WAVE":
        raise ValueError("not a RIFF/WAVE file")
    pos, fmt = 12, None
    while pos + 8 <= len(data):
        cid, size = data[pos:pos + 4], struct.unpack_from("<I", data, pos + 4)[0]
        body = pos + 8
        if cid == b"fmt ":
            if size < 16 or body + 16 > len(data):
                raise ValueError("truncated fmt chunk")
            tag, ch, rate, _byte_rate, align, bits = struct.unpack_from("<HHIIHH", data, body)
            if tag == 0xFFFE:
                if size < 40 or body + 40 > len(data):
                    raise ValueError("truncated extensible fmt chunk")
                tag = struct.unpack_from("<H", data, body + 24)[0]
            fmt = (tag, ch, rate, align, bits)
        elif cid == b"data":
            if fmt is None:
                raise ValueError("data before fmt")
            tag, ch, rate, align, bits = fmt
            code = {(1, 8): U8, (1, 16): S16, (1, 24): S24, (1, 32): S32, (3, 32): F32}.get((tag, bits))
            if code is None:
                raise ValueError("unsupported encoding")
            bpf = ch * SAMPLE_BYTES[code]
            if not 1 <= ch <= 32:
                raise ValueError("channel count")
            if align != bpf:
                raise ValueError("block align")
            avail = len(data) - body
            nbytes = avail if size in (0, 0xFFFFFFFF) or size > avail else size
            return {"format": code, "channels": ch, "sample_rate": rate, "frames": nbytes // bpf, "data_offset": body}
        pos = body + size + (size & 1)
    raise ValueError("no data chunk")
