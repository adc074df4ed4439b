"""The Node-API addon (spectrogram_b200/js/napi_shim.c) driven by a C stand-in for the Node runtime
(tests/napi_host/fake_napi_host.c): Node itself is not installed in this image, and N-API is a C ABI
that the host process supplies, so a fake host exercises the very object file Node would load."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ADDON = os.path.join(ROOT, "spectrogram_b200", "js", "build", "spectrogram.node")
HOST = os.path.join(ROOT, "tests", "napi_host", "fake_napi_host")


def build():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "spectrogram_b200", "js")], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "napi_host")], stdout=subprocess.DEVNULL)


def test_addon_registers_and_validates_like_web_audio():
    build()
    r = subprocess.run([HOST, ADDON, "cpu"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "FAIL" not in r.stdout and "0 failure(s)" in r.stdout


def test_addon_exports_only_the_module_entry_point():
    build()
    out = subprocess.run(["nm", "-D", "--defined-only", ADDON], capture_output=True, text=True).stdout
    names = {line.split()[-1] for line in out.splitlines() if " T " in line}
    assert "napi_register_module_v1" in names
    assert not any(n.startswith("sg_") for n in names)       # the C ABI stays in libsgcore.so
    undefined = subprocess.run(["nm", "-D", "--undefined-only", ADDON], capture_output=True, text=True).stdout
    assert "sg_stft_batch" in undefined and "napi_create_function" in undefined and "sg_stft_batch_multi" in undefined
    assert "sg_stream_info" in undefined and "sg_ring_info" in undefined       # array-length checks in the shim
    assert "sg_ring_view" in undefined and "sg_ring_append" in undefined
    assert "sg_wav_parse" in undefined and "sg_pcm_ingest" in undefined and "sg_stft_pcm" in undefined


def test_js_facade_keeps_the_analysernode_surface():
    """index.js cannot be executed here; check statically that it exposes the names the reference uses
    (UI/player.js:7-11, 3D/visualizer.js:299-301,351-363)."""
    text = open(os.path.join(ROOT, "spectrogram_b200", "js", "index.js")).read()
    for name in ("get fftSize", "set fftSize", "get frequencyBinCount", "get minDecibels", "set maxDecibels",
                 "set smoothingTimeConstant", "getByteFrequencyData(array)", "getFloatFrequencyData(array)",
                 "getByteTimeDomainData(array)", "getFloatTimeDomainData(array)", "createAnalyser", "connect(", "IndexSizeError",
                 "class SonogramRing", "append(frames)", "view(width, height, out)",
                 "decodeAudioData(", "class AudioBuffer", "getChannelData(", "numberOfChannels", "spectrogramPcm(",
                 "native.stftBatchMulti", "FinalizationRegistry", "closeAll", "OUT_CTOR[o.output]", "wholeNumber("):
        assert name in text, name


@pytest.mark.gpu
def test_addon_matches_the_ctypes_path_on_gpu(tmp_path, engine):
    import spectrogram_b200 as sg
    from oracle import analyser_oracle as O
    build()
    out = tmp_path / "napi_out.bin"
    r = subprocess.run([HOST, ADDON, "gpu", str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    got = np.fromfile(out, dtype=np.uint8).reshape(-1, 1024)
    t = np.arange(44100) / 44100.0
    x = (0.5 * np.sin(2 * np.pi * (20.0 * t + 0.5 * (20000.0 - 20.0) * t * t))).astype(np.float32)
    want = engine.spectrogram(x, sg.Options())
    assert np.array_equal(got, want)
    ref = O.spectrogram(x, O.Config())[0]
    assert np.abs(got.astype(int) - ref.astype(int)).max() <= 1
