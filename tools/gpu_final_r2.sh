#!/bin/bash
# round-2 closing session: shape sweep, then one full capture of the final n_fft 8192 kernel
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
timeout 600 python tools/shape_sweep.py > $O/r2_shape_sweep.txt 2>&1
tail -4 $O/r2_shape_sweep.txt
K="python tools/kbench.py 0 --nfft 8192 --hop 2048 --clips 64 --clip-len 2880000 --steps 3"
$K > $O/r2_wreg_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:stft_wreg -s 3 -c 1 -f -o $O/r2_prof_wreg $K > $O/r2_wreg_ncu.log 2>&1
tail -1 $O/r2_wreg_plain.log; tail -2 $O/r2_wreg_ncu.log
