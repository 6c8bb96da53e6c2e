#!/bin/bash
mkdir -p gpurun_out/c4
for t in "c3 full 18944" "c3 reduced 18944" "c5 reduced 378880" "c4 reduced 32768" "c3 full 18944 --mt" "c2 full 262144"; do python scripts/ncu_target.py $t --reps 2 2>&1 | grep -v "^Exception\|^Traceback\|^  File\|^TypeError"; done > gpurun_out/c4/timings.log 2>&1
cat gpurun_out/c4/timings.log
NCU="ncu --set full --clock-control none --import-source on -s 1 -c 1 -f"
$NCU -k regex:quad_sweep -o gpurun_out/c4/c3_full python scripts/ncu_target.py c3 full 18944 --reps 1 > gpurun_out/c4/ncu_c3_full.log 2>&1
$NCU -k regex:quad_sweep -o gpurun_out/c4/c3_red python scripts/ncu_target.py c3 reduced 18944 --reps 1 > gpurun_out/c4/ncu_c3_red.log 2>&1
$NCU -k regex:quad_sweep -o gpurun_out/c4/c3_mt python scripts/ncu_target.py c3 full 18944 --mt --reps 1 > gpurun_out/c4/ncu_c3_mt.log 2>&1
$NCU -k regex:quad_sweep -o gpurun_out/c4/c2_full python scripts/ncu_target.py c2 full 262144 --reps 1 > gpurun_out/c4/ncu_c2_full.log 2>&1
ls -la gpurun_out/c4
