import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    # The built artefacts are git-ignored: a fresh checkout builds them once (nvcc cross-compiles without a GPU).
    missing = [p for p in (os.path.join(ROOT, "spectrogram_b200", "libsgcore.so"), os.path.join(ROOT, "oracle", "liboracle.so"))
               if not os.path.exists(p)]
    if missing:
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope="session")
def engine():
    import spectrogram_b200 as sg

    eng = sg.Engine(0)
    yield eng
    eng.close()
