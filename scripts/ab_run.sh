#!/bin/bash
# usage: scripts/ab_run.sh "<ncu_target args>" ...   -- times every target with every library under scratch_libs/, twice
for rep in 1 2; do for f in scratch_libs/lib*.so; do echo "== $f"; for t in "$@"; do MCDP_LIB=$f python scripts/ncu_target.py $t --reps 3 2>&1 | grep -v "^Exception\|^Traceback\|^  File\|^TypeError"; done; done; done
