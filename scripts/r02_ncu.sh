#!/bin/bash
# usage: scripts/r02_ncu.sh <outdir> <name> <kernel mangled substring for by-line> <ncu_target args...>
# capture one launch with --set full, summarise on the box (key metrics + by source line), drop the report
out=$1; name=$2; ksub=$3; shift 3
mkdir -p $out
ncu --set full --clock-control none --import-source on -k regex:quad_sweep -s 1 -c 1 -f -o $out/$name python scripts/ncu_target.py "$@" --reps 1 > $out/ncu_$name.log 2>&1
python scripts/ncu_summary.py $out/$name.ncu-rep $out/$name.json > $out/${name}_summary.txt 2>&1
python scripts/ncu_by_line.py $out/$name.ncu-rep $ksub 70 > $out/${name}_by_line.txt 2>&1
rm -f $out/$name.ncu-rep
