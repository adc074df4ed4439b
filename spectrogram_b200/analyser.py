"""``AnalyserNode``: the object the reference creates with ``context.createAnalyser()``
(src/javascripts/UI/player.js:7), configures (player.js:10-11, 3D/visualizer.js:351,357,362) and
polls (3D/visualizer.js:352,358,363).  Same property and method names, same argument meaning
(getters write into the caller's typed array, copy min(len, frequencyBinCount) elements, never
resize), same errors (IndexSizeError / TypeError).  ``push()`` replaces the audio render thread.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from .api import Engine, default_engine


class AnalyserNode:
    def __init__(self, engine: Engine | None = None, fftSize: int = 2048):
        self._engine = engine or default_engine(0)
        self._lib = L.load()
        h = C.c_void_p()
        L.check(self._lib.sg_analyser_create(self._engine.handle, C.byref(h)))
        self._h = h
        if fftSize != 2048:
            self.fftSize = fftSize

    # -- attributes ---------------------------------------------------------------------
    @property
    def fftSize(self) -> int:
        return int(self._lib.sg_analyser_get_fft_size(self._h))

    @fftSize.setter
    def fftSize(self, n) -> None:
        if isinstance(n, bool) or not isinstance(n, (int, np.integer)):
            if isinstance(n, float) and n.is_integer():
                n = int(n)
            else:
                raise L.IndexSizeError("fftSize must be a power of two in [32, 32768]")
        L.check(self._lib.sg_analyser_set_fft_size(self._h, int(n)))

    @property
    def frequencyBinCount(self) -> int:
        return int(self._lib.sg_analyser_get_frequency_bin_count(self._h))

    @property
    def minDecibels(self) -> float:
        return float(self._lib.sg_analyser_get_min_decibels(self._h))

    @minDecibels.setter
    def minDecibels(self, v: float) -> None:
        L.check(self._lib.sg_analyser_set_min_decibels(self._h, float(v)))

    @property
    def maxDecibels(self) -> float:
        return float(self._lib.sg_analyser_get_max_decibels(self._h))

    @maxDecibels.setter
    def maxDecibels(self, v: float) -> None:
        L.check(self._lib.sg_analyser_set_max_decibels(self._h, float(v)))

    @property
    def smoothingTimeConstant(self) -> float:
        return float(self._lib.sg_analyser_get_smoothing_time_constant(self._h))

    @smoothingTimeConstant.setter
    def smoothingTimeConstant(self, v: float) -> None:
        L.check(self._lib.sg_analyser_set_smoothing_time_constant(self._h, float(v)))

    # -- audio in -----------------------------------------------------------------------
    def push(self, samples) -> None:
        """Feeds mono float samples to the node (what ``mix.connect(analyser)`` does in the
        browser, UI/player.js:25).  Multi-channel input is down-mixed by the caller."""
        x = np.ascontiguousarray(samples, dtype=np.float32).ravel()
        L.check(self._lib.sg_analyser_push(self._h, x.ctypes.data, x.size))

    def connect(self, *_a, **_k):  # graph plumbing is out of scope; kept so call sites run
        return None

    def disconnect(self, *_a, **_k):
        return None

    # -- getters (write into the caller-owned array) ---------------------------------------
    @staticmethod
    def _dst(array, dtype, what):
        if not isinstance(array, np.ndarray) or array.dtype != dtype or array.ndim != 1 or not array.flags.c_contiguous:
            raise TypeError(f"{what} needs a contiguous 1-D {np.dtype(dtype).name} array")
        return array

    def getByteFrequencyData(self, array) -> None:
        a = self._dst(array, np.uint8, "getByteFrequencyData")
        L.check(self._lib.sg_analyser_get_byte_frequency_data(self._h, a.ctypes.data, a.size))

    def getFloatFrequencyData(self, array) -> None:
        a = self._dst(array, np.float32, "getFloatFrequencyData")
        L.check(self._lib.sg_analyser_get_float_frequency_data(self._h, a.ctypes.data, a.size))

    def getByteTimeDomainData(self, array) -> None:
        a = self._dst(array, np.uint8, "getByteTimeDomainData")
        L.check(self._lib.sg_analyser_get_byte_time_domain_data(self._h, a.ctypes.data, a.size))

    def getFloatTimeDomainData(self, array) -> None:
        a = self._dst(array, np.float32, "getFloatTimeDomainData")
        L.check(self._lib.sg_analyser_get_float_time_domain_data(self._h, a.ctypes.data, a.size))

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.sg_analyser_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
