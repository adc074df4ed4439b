"""The SG_DEBUG build of the library (epoch tags beside the shared-memory hand-offs, csrc/common.cuh) over the shape
list of tools/sanitize_run.py: the in-tree stand-in for compute-sanitizer's racecheck, which is closed on this pool."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEBUG_LIB = os.path.join(ROOT, "spectrogram_b200", "libsgcore_debug.so")
pytestmark = pytest.mark.gpu


def run_debug(extra_env=None):
    if not os.path.exists(DEBUG_LIB):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "spectrogram_b200", "csrc"), "debug"], stdout=subprocess.DEVNULL)
    env = dict(os.environ, SG_LIBSGCORE=DEBUG_LIB, **(extra_env or {}))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sanitize_run.py")], capture_output=True, text=True, env=env,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = next(l for l in r.stdout.splitlines() if l.startswith("{"))
    return json.loads(line)


def test_no_hand_off_reads_a_value_of_the_wrong_iteration():
    d = run_debug()
    counts = d["debug_counts"]
    assert counts[15] > 1000, "the checks did not run"           # warp iterations with the tags armed
    assert counts[:15] == [0] * 15, f"tag mismatches per site: {counts[:15]}"
    assert "warp32x32x2p" in d["kernels"] and "warp32x32x2s" in d["kernels"]
    assert {"eo4096s", "p16s", "p8s", "p4s"} <= set(d["kernels"])


def test_the_checker_sees_a_broken_turn_chain():
    """Negative control: with the fused smoothing kernels not waiting for its turn, warps read state slots before the
    pair ahead of them has written them -- the tags must say so (outputs are wrong in this mode by construction)."""
    d = run_debug({"SG_DEBUG_BREAK_CHAIN": "1"})
    assert d["debug_counts"][2] > 0          # n_fft 2048 kernel
    assert d["debug_counts"][3] > 0          # part-warp kernels
    assert d["debug_counts"][4] > 0          # n_fft 4096 kernel
    assert d["debug_counts"][0] == 0 and d["debug_counts"][1] == 0
