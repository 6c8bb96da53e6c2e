#!/bin/bash
for rep in 1 2; do for f in scratch_libs/libprev.so scratch_libs/libnew.so scratch_libs/libnew2.so; do echo "== $f"; for t in "c3 full 18944" "c2 full 265216" "c3 reduced 18944"; do MCDP_LIB=$f python scripts/ncu_target.py $t --reps 4 2>&1 | grep -v "^Exception\|^Traceback\|^  File\|^TypeError"; done; done; done
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit,temperature.gpu --format=csv
