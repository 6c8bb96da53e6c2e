#!/bin/bash
# scratch: time the C3 points with every library under scratch_libs/
for f in scratch_libs/*.so; do echo "== $f"; MCDP_LIB=$f timeout 200 python scripts/ablate_quick.py 2>&1 | grep -v "^c2\|generic"; done
