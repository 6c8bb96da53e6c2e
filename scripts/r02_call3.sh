#!/bin/bash
mkdir -p gpurun_out/c3
( time python -m pytest tests -m gpu -q -x ) > gpurun_out/c3/pytest.log 2>&1
tail -40 gpurun_out/c3/pytest.log
for f in scratch_libs/libbase.so scratch_libs/libs0h0.so scratch_libs/libs1h0.so scratch_libs/libs0h1.so scratch_libs/libs1h4.so; do
  echo "== $f"
  for t in "c3 reduced 18944" "c5 reduced 378880" "c4 reduced 32768"; do MCDP_LIB=$f python scripts/ncu_target.py $t --reps 2 2>&1 | grep -v "^Exception\|^Traceback\|^  File\|^TypeError"; done
done > gpurun_out/c3/ab.log 2>&1
cat gpurun_out/c3/ab.log
for t in "c3 full 2048" "c3 full 128" "c3 full 1" "c3 full 4096"; do python scripts/ncu_target.py $t 2>&1 | grep -v "^Exception\|^Traceback\|^  File\|^TypeError"; done
