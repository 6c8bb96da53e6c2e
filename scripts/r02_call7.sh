#!/bin/bash
mkdir -p gpurun_out/c7
( time python -m pytest tests -m gpu -q -x ) > gpurun_out/c7/pytest.log 2>&1
tail -8 gpurun_out/c7/pytest.log
for t in "c3 full 18944" "c3 reduced 18944" "c5 reduced 378880" "c5 reduced 757760" "c5 reduced 10000000" "c4 reduced 32768" "c3 full 18944 --mt" "c2 full 262144" "c2 full 265216"; do python scripts/ncu_target.py $t --reps 2 2>&1 | grep -v "^Exception\|^Traceback\|^  File\|^TypeError"; done > gpurun_out/c7/timings.log 2>&1
cat gpurun_out/c7/timings.log
python - <<'PY'
import torch; print(torch.cuda.mem_get_info())
PY
