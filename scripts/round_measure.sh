#!/bin/bash
# Round-2 measurement on one B200 (`gpurun -- bash scripts/round_measure.sh`): the bench line and the reference arm as the
# driver runs them, the ncu launch list of the bench command, one --set full capture per workload summarised on the box
# (scripts/r02_ncu.sh: key metrics + attribution to source lines; the reports themselves stay on the box).
out=gpurun_out/r02
mkdir -p $out
( time python bench.py --steps 20 --warmup 5 ) > $out/bench_c3.json 2> $out/bench_c3.err
( time python bench.py --impl reference --steps 2 --warmup 1 ) > $out/bench_c3_reference_arm.json 2> $out/bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu > $out/ncu_launches.log 2>&1
bash scripts/r02_ncu.sh $out c3_full quad_sweep_kernelILi0ELb1ELb1 c3 full 18944
bash scripts/r02_ncu.sh $out c3_red quad_sweep_kernelILi2ELb1ELb1 c3 reduced 18944
bash scripts/r02_ncu.sh $out c3_mt quad_sweep_kernelILi0ELb1ELb1 c3 full 18944 --mt
bash scripts/r02_ncu.sh $out c2_full quad_sweep_kernelILi0ELb1ELb1 c2 full 265216
bash scripts/r02_ncu.sh $out c4_red quad_sweep_kernelILi2ELb1ELb1 c4 reduced 32768
bash scripts/r02_ncu.sh $out c5_red quad_sweep_kernelILi2ELb1ELb0 c5 reduced 37888 --spl 4
for t in "c3 full 18944" "c3 full 18944 --mt" "c2 full 265216" "c3 reduced 18944" "c4 reduced 32768" "c4 reduced 262144" "c5 reduced 378880" "c1 full 1048576" "c3 full 2048" "c3 full 128" "c3 full 64"; do python scripts/ncu_target.py $t --reps 3 2>&1 | grep -v "^Exception\|^Traceback\|^  File\|^TypeError"; done > $out/single_launch_timings.txt
cat $out/single_launch_timings.txt; tail -3 $out/bench_c3.err; head -c 1500 $out/bench_c3.json; echo; head -c 600 $out/bench_c3_reference_arm.json; echo; tail -5 $out/launches_bench.csv | cut -c1-200
