#!/bin/bash
# full capture of the n_fft 8192 register-family kernel (after the same command ran clean without the profiler)
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
K="python tools/kbench.py 0 --nfft 8192 --hop 2048 --clips 64 --clip-len 2880000 --steps 3"
$K > $O/r2_wreg_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:stft_wreg -s 3 -c 1 -f -o $O/r2_prof_wreg $K > $O/r2_wreg_ncu.log 2>&1
tail -2 $O/r2_wreg_plain.log; tail -2 $O/r2_wreg_ncu.log
