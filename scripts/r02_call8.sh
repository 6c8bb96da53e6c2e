#!/bin/bash
mkdir -p gpurun_out/c8
( time python -m pytest tests -m gpu -q -x ) > gpurun_out/c8/pytest.log 2>&1
tail -8 gpurun_out/c8/pytest.log
for rep in 1 2; do for f in scratch_libs/libbase.so scratch_libs/liblit.so; do echo "== $f"; for t in "c3 full 18944" "c3 full 18944 --mt" "c2 full 262144" "c3 reduced 18944"; do MCDP_LIB=$f python scripts/ncu_target.py $t --reps 4 2>&1 | grep -v "^Exception\|^Traceback\|^  File\|^TypeError"; done; done; done > gpurun_out/c8/ab.log 2>&1
cat gpurun_out/c8/ab.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:quad_sweep --csv --log-file gpurun_out/c8/c5_launches.csv python scripts/ncu_target.py c5 reduced 757760 --reps 0 > /dev/null 2>&1
cut -d, -f5,8,9,10,12- gpurun_out/c8/c5_launches.csv | tail -8
