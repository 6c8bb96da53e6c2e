#!/bin/bash
python -m pytest tests/test_gpu_quad.py tests/test_gpu_parity.py tests/test_gpu_statistics.py -q -x 2>&1 | tail -2
for rep in 1 2; do for f in scratch_libs/libhead.so scratch_libs/libint.so; do echo "== $f"; for t in "c3 full 18944" "c3 full 18944 --mt" "c3 reduced 18944" "c4 reduced 32768" "c5 reduced 378880"; do MCDP_LIB=$f python scripts/ncu_target.py $t --reps 4 2>&1 | grep -v "^Exception\|^Traceback\|^  File\|^TypeError"; done; done; done
