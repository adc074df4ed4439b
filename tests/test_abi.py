"""CPU-side checks of the drop-in boundary: libsgcore.so loads, exports every symbol
include/sgcore.h declares, validates like Web Audio, and refuses to run without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "sgcore.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sg_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from spectrogram_b200 import _lib
    lib = _lib.load()
    names = header_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), f"libsgcore.so does not export {n}"
    # and the Python binding declares a signature for each of them
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)


def test_struct_layout_matches_header():
    from spectrogram_b200 import _lib
    assert C.sizeof(_lib.StftConfig) == 48
    assert _lib.StftConfig.custom_window.offset == 32 and _lib.StftConfig.colormap.offset == 40


def test_defaults_are_the_reference_operating_point():
    from spectrogram_b200 import _lib
    lib = _lib.load()
    cfg = _lib.StftConfig()
    assert lib.sg_stft_config_default(C.byref(cfg)) == 0
    # UI/player.js:10-11 -> fftSize 2048, smoothing 0; AnalyserNode default dB range
    assert (cfg.n_fft, cfg.hop, cfg.window, cfg.output, cfg.align) == (2048, 512, 0, 0, 0)
    assert (cfg.min_db, cfg.max_db, cfg.smoothing) == (-100.0, -30.0, 0.0)
    assert lib.sg_stft_num_bins(C.byref(cfg)) == 1024
    assert lib.sg_stft_num_frames(C.byref(cfg), 441000) == 858
    cfg.align = 1
    assert lib.sg_stft_num_frames(C.byref(cfg), 441000) == 861
    assert lib.sg_stft_elem_bytes(C.byref(cfg)) == 1


@pytest.mark.parametrize("field,value", [("n_fft", 2047), ("n_fft", 65536), ("n_fft", 14), ("hop", 0),
                                         ("max_db", -100.0), ("smoothing", 1.5), ("smoothing", -0.1)])
def test_invalid_attributes_are_index_size_errors(field, value):
    from spectrogram_b200 import _lib
    lib = _lib.load()
    cfg = _lib.StftConfig()
    lib.sg_stft_config_default(C.byref(cfg))
    setattr(cfg, field, value)
    assert lib.sg_stft_num_frames(C.byref(cfg), 10000) == -1
    assert len(_lib.last_error()) > 0


def test_generalised_sizes_are_accepted():
    from spectrogram_b200 import _lib
    lib = _lib.load()
    cfg = _lib.StftConfig()
    lib.sg_stft_config_default(C.byref(cfg))
    for n, hop, clip, want in [(400, 160, 480000, 2998), (512, 160, 57600000, 359997), (32768, 8192, 32768, 1), (32, 8, 31, 0)]:
        cfg.n_fft, cfg.hop = n, hop
        assert lib.sg_stft_num_frames(C.byref(cfg), clip) == want


def test_reference_colormap_matches_oracle():
    import spectrogram_b200 as sg
    from oracle import analyser_oracle as O
    assert np.array_equal(sg.colormap_reference(), O.colormap_lut_u32())


def test_no_gpu_means_no_engine():
    """The product has no CPU path: without a device, engine creation fails loudly."""
    import spectrogram_b200 as sg
    from spectrogram_b200 import _lib
    if _lib.load().sg_device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(sg.EngineError):
        sg.Engine(0)
    with pytest.raises(sg.EngineError):
        sg.spectrogram(np.zeros(4096, np.float32))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "spectrogram_b200")
    for dirpath, _d, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".inl", ".h", ".js", ".c")):
                text = open(os.path.join(dirpath, fn), errors="replace").read()
                assert "oracle" not in text.lower(), f"{fn} mentions the oracle"
