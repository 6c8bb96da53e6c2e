#!/bin/bash
mkdir -p gpurun_out/c6
( time python -m pytest tests -m gpu -q -x ) > gpurun_out/c6/pytest.log 2>&1
tail -25 gpurun_out/c6/pytest.log
for t in "c3 full 18944" "c3 reduced 18944" "c5 reduced 378880" "c4 reduced 32768" "c3 full 18944 --mt" "c3 full 6656" "c3 full 128" "c3 full 64"; do python scripts/ncu_target.py $t --reps 2 2>&1 | grep -v "^Exception\|^Traceback\|^  File\|^TypeError"; done > gpurun_out/c6/timings.log 2>&1
cat gpurun_out/c6/timings.log
( time python bench.py --steps 3 --warmup 3 ) > gpurun_out/c6/bench.json 2> gpurun_out/c6/bench.err
tail -5 gpurun_out/c6/bench.err; cat gpurun_out/c6/bench.json | cut -c1-6000
