"""Host-side logic that needs no GPU: options mapping, sharding, the multi-process bench plumbing."""
import os
import subprocess
import sys

import numpy as np
import pytest

import spectrogram_b200 as sg
from spectrogram_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_and_partition():
    for n, g in [(4096, 8), (4096, 3), (5, 8), (0, 4), (256, 2)]:
        b = sg.shard_bounds(n, g)
        assert b[0][0] == 0 and b[-1][1] == n
        assert all(b[i][1] == b[i + 1][0] for i in range(g - 1))
        sizes = [hi - lo for lo, hi in b]
        assert max(sizes) - min(sizes) <= 1


def test_options_map_to_the_c_struct():
    cfg, _keep = sg.Options(fftSize=512, hop=160, window="hann", output="db", align="analyser",
                            minDecibels=-90, maxDecibels=-10, smoothingTimeConstant=0.8).to_c()
    assert (cfg.n_fft, cfg.hop, cfg.window, cfg.output, cfg.align) == (512, 160, _lib.WINDOW_HANN, _lib.OUT_F32_DB, _lib.ALIGN_ANALYSER)
    assert (cfg.min_db, cfg.max_db) == (-90.0, -10.0) and abs(cfg.smoothing - 0.8) < 1e-7
    w = np.hanning(512).astype(np.float32)
    cfg, keep = sg.Options(fftSize=512, window=w).to_c()
    assert cfg.window == _lib.WINDOW_CUSTOM and keep
    with pytest.raises(TypeError):
        sg.Options(window="kaiser").to_c()
    with pytest.raises(TypeError):
        sg.Options(output="png").to_c()
    with pytest.raises(TypeError):
        sg.Options(fftSize=512, window=np.ones(100, np.float32)).to_c()


def test_status_codes_map_to_web_audio_exceptions():
    with pytest.raises(sg.IndexSizeError):
        _lib.check(_lib.SG_ERR_INDEX_SIZE)
    with pytest.raises(TypeError):
        _lib.check(_lib.SG_ERR_INVALID_ARG)
    with pytest.raises(MemoryError):
        _lib.check(_lib.SG_ERR_OOM)
    with pytest.raises(sg.EngineError):
        _lib.check(_lib.SG_ERR_CUDA)
    assert sg.IndexSizeError.name == "IndexSizeError"


def test_bench_shard_plan_world_size_2_gloo(tmp_path):
    """The N>1 path of bench.py (clip sharding, barrier, max-over-ranks reduction) on CPU/gloo."""
    script = tmp_path / "gloo_ranks.py"
    script.write_text(r"""
import sys
sys.path.insert(0, %r)
import torch, torch.distributed as dist
import bench
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
plan = bench.shard_plan(total_clips=10, world=world, rank=rank)
t = torch.tensor([float(plan["n_clips"])])
dist.all_reduce(t)
assert int(t.item()) == 10, t
ms = bench.max_over_ranks(5.0 + rank, device="cpu")
assert ms == 6.0, ms
frames = bench.sum_over_ranks(100 * (rank + 1), device="cpu")
assert frames == 300, frames
bench.barrier()
dist.destroy_process_group()
print("ok", rank)
""" % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29517", str(script)],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.count("ok") == 2


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--clips-per-gpu", "8"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    import json
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "stft_frames_per_s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["config"]["n_fft"] == 2048
    assert line["config"]["clips_per_gpu"] == 8      # the CPU arm runs the GPU arm's batch, clip for clip
