#!/bin/bash
# round-2 profiling session: launch list of the bench, full captures of the headline kernel and of the fused smoothing kernel
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --quick-configs --e2e-clips 64"
$B > $O/r2_prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2_launches.csv $B > $O/r2_prof_ncu1.log 2>&1
K="python tools/kbench.py 0 --clips 512 --steps 3"
$K > $O/r2_prof_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:stft_w32x2p -s 3 -c 1 -f -o $O/r2_prof_xp $K > $O/r2_prof_ncu2.log 2>&1
KT="python tools/kbench.py 0 --clips 512 --steps 3 --tau 0.8"
$KT > $O/r2_prof_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:stft_w32x2s -s 3 -c 1 -f -o $O/r2_prof_xs $KT > $O/r2_prof_ncu3.log 2>&1
tail -2 $O/r2_prof_plain2.log $O/r2_prof_plain3.log $O/r2_prof_ncu2.log $O/r2_prof_ncu3.log
