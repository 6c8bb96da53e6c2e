#!/bin/bash
mkdir -p gpurun_out/c12
( time python -m pytest tests/test_gpu_parity.py tests/test_gpu_quad.py tests/test_gpu_statistics.py -q -x ) > gpurun_out/c12/pytest.log 2>&1
tail -5 gpurun_out/c12/pytest.log | cut -c1-300
for t in "c3 full 18944" "c2 full 265216" "c3 full 18944 --mt" "c3 reduced 18944" "c4 reduced 32768"; do python scripts/ncu_target.py $t --reps 4 2>&1 | grep -v "^Exception\|^Traceback\|^  File\|^TypeError"; done
