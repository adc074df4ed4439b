#!/bin/bash
# tolerance report of every kernel family, then one full capture of the n_fft 1024 fused smoothing kernel
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
python tools/tolerance_report.py $O/r2_tolerance_report.jsonl > $O/r2_tolerance_report.txt 2>&1
tail -3 $O/r2_tolerance_report.txt
K="python tools/kbench.py 0 --nfft 1024 --hop 256 --clips 512 --steps 3 --tau 0.8"
$K > $O/r2_ps_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:stft_pair_s -s 3 -c 1 -f -o $O/r2_prof_ps $K > $O/r2_ps_ncu.log 2>&1
tail -2 $O/r2_ps_plain.log; tail -2 $O/r2_ps_ncu.log
