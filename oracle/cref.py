"""ctypes loader for oracle/liboracle.so (analyser_ref.c).  TEST / BASELINE INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")


class SgoConfig(C.Structure):
    _fields_ = [
        ("n_fft", C.c_int32), ("hop", C.c_int32), ("window", C.c_int32),
        ("output", C.c_int32), ("align", C.c_int32),
        ("min_db", C.c_float), ("max_db", C.c_float), ("smoothing", C.c_float),
        ("custom_window", C.POINTER(C.c_float)), ("colormap", C.POINTER(C.c_uint32)),
    ]


def build() -> str:
    src = os.path.join(_HERE, "analyser_ref.c")
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.sgo_stft_batch.restype = C.c_int
        _lib.sgo_stft_batch.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.POINTER(SgoConfig),
                                        C.c_void_p, C.c_int]
        _lib.sgo_num_frames.restype = C.c_int64
        _lib.sgo_num_frames.argtypes = [C.POINTER(SgoConfig), C.c_int64]
        _lib.sgo_colormap_lut.argtypes = [C.c_void_p]
        _lib.sgo_time_domain_byte.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
        _lib.sgo_max_threads.restype = C.c_int
    return _lib


_OUT_DTYPE = {0: (np.uint8, 1), 1: (np.float32, 1), 2: (np.uint8, 4), 3: (np.float32, 1)}


def stft_batch(pcm: np.ndarray, cfg, n_threads: int = 0) -> np.ndarray:
    """cfg: oracle.analyser_oracle.Config.  Returns [clips, frames, bins(,4)]."""
    x = np.ascontiguousarray(np.atleast_2d(pcm), dtype=np.float32)
    n_clips, clip_len = x.shape
    c = SgoConfig(cfg.n_fft, cfg.hop, cfg.window, cfg.output, cfg.align,
                  cfg.min_db, cfg.max_db, cfg.smoothing, None, None)
    keep = None
    if cfg.custom_window is not None:
        keep = np.ascontiguousarray(cfg.custom_window, dtype=np.float32)
        c.custom_window = keep.ctypes.data_as(C.POINTER(C.c_float))
    frames = lib().sgo_num_frames(C.byref(c), clip_len)
    dt, lanes = _OUT_DTYPE[cfg.output]
    shape = (n_clips, frames, cfg.n_fft // 2) + ((4,) if lanes == 4 else ())
    out = np.empty(shape, dtype=dt)
    rc = lib().sgo_stft_batch(x.ctypes.data, n_clips, clip_len, C.byref(c), out.ctypes.data, n_threads)
    if rc != 0:
        raise ValueError(f"sgo_stft_batch rejected the configuration (rc={rc})")
    return out


def colormap_lut_u32() -> np.ndarray:
    lut = np.empty(256, dtype=np.uint32)
    lib().sgo_colormap_lut(lut.ctypes.data)
    return lut


def time_domain_byte(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty(x.shape, dtype=np.uint8)
    lib().sgo_time_domain_byte(x.ctypes.data, x.size, out.ctypes.data)
    return out


def max_threads() -> int:
    return int(lib().sgo_max_threads())
