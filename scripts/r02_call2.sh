#!/bin/bash
# round 2, GPU call 2: contract v2 + staged statistics + cluster launches: tests, timings, one full capture of C3
mkdir -p gpurun_out/c2
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/c2/pytest.log 2>&1
tail -5 gpurun_out/c2/pytest.log
for t in "c3 full 18944" "c3 full 18944 --mt" "c3 reduced 18944" "c5 reduced 378880" "c5 reduced 262144" "c4 reduced 32768" "c2 full 262144" \
         "c3 full 2048" "c3 full 2048 --opt 6=1" "c3 full 2048 --opt 6=4" "c3 full 128" "c3 full 128 --opt 6=1" "c3 full 9472" "c3 full 4096" "c1 full 1048576" "c2 full 1024"; do
  python scripts/ncu_target.py $t 2>&1 | grep -v "^Exception\|^Traceback\|^  File\|^TypeError"; done > gpurun_out/c2/timings.log 2>&1
cat gpurun_out/c2/timings.log
NCU="ncu --set full --clock-control none --import-source on -s 1 -c 1 -f"
$NCU -k regex:quad_sweep -o gpurun_out/c2/c3_full python scripts/ncu_target.py c3 full 18944 --reps 1 > gpurun_out/c2/ncu_c3_full.log 2>&1
$NCU -k regex:quad_sweep -o gpurun_out/c2/c3_red python scripts/ncu_target.py c3 reduced 18944 --reps 1 > gpurun_out/c2/ncu_c3_red.log 2>&1
$NCU -k regex:quad_sweep -o gpurun_out/c2/c3_mt python scripts/ncu_target.py c3 full 18944 --mt --reps 1 > gpurun_out/c2/ncu_c3_mt.log 2>&1
cp mc_dagprop_b200/libmcdp_b200.so gpurun_out/c2/libmcdp_b200.so
