#!/bin/bash
for t in "c4 reduced 32768" "c4 reduced 32768 --wpg 10" "c4 reduced 32768 --wpg 10 --gpc 2" "c4 reduced 32768 --wpg 16" "c4 reduced 18944" "c4 reduced 9472 --opt 6=2" "c4 reduced 4352 --opt 6=4" "c4 reduced 4352 --opt 6=1"; do python scripts/ncu_target.py $t --reps 2 2>&1 | grep -v "^Exception\|^Traceback\|^  File\|^TypeError"; done
