#!/bin/bash
# round-2 multi-GPU session (run with gpurun --gpus N): sg_stft_batch_multi over N real devices against one device, the
# copy-only host ceiling at 1 .. N ranks, and the bench at N ranks
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
N=${1:-8}
python - > $O/r2_multi_api_n$N.txt 2>&1 <<EOF
import time, numpy as np
import spectrogram_b200 as sg
n = sg.device_count()
print("devices", n)
rng = np.random.default_rng(3)
clips = 256
x = (0.2 * rng.standard_normal((clips, 441000))).astype(np.float32)
one = sg.spectrogram(x, devices=0)
for g in sorted({1, 2, 4, n} & set(range(1, n + 1))):
    devs = list(range(g))
    sg.spectrogram(x[:g * 2], devices=devs)          # engines + plans up
    t0 = time.perf_counter(); y = sg.spectrogram(x, devices=devs); dt = time.perf_counter() - t0
    print(f"sg_stft_batch_multi over {g} GPU(s): {dt*1e3:.1f} ms, {clips*858/dt/1e6:.2f} M frames/s (pageable host arrays), identical to 1 GPU: {np.array_equal(y, one)}")
EOF
cat $O/r2_multi_api_n$N.txt
for g in 1 2 4 8; do
  if [ $g -le $N ]; then
    if [ $g -eq 1 ]; then python tools/host_path_ceiling.py > $O/r2_host_ceiling_g$g.json 2>$O/r2_host_ceiling.err
    else python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29533 tools/host_path_ceiling.py > $O/r2_host_ceiling_g$g.json 2>$O/r2_host_ceiling.err; fi
    tail -1 $O/r2_host_ceiling_g$g.json
  fi
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 20 --warmup 5 > $O/r2_bench_g$N.json 2> $O/r2_bench_g$N.err
tail -c 1500 $O/r2_bench_g$N.json
