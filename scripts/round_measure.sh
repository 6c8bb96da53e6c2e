#!/bin/bash
# Round measurement on one B200: bench lines, reference arm, ncu launch list, full captures of the sweep kernel.
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_r01_c3.json 2> gpurun_out/err_c3.log
python bench.py --spl 2 --no-cpu > gpurun_out/bench_r01_c3_pair.json 2> gpurun_out/err_c3_pair.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r01_ref.json 2> gpurun_out/err_ref.log
for w in c2 c4 c5 c1; do python bench.py --workload $w --no-cpu > gpurun_out/bench_r01_$w.json 2> gpurun_out/err_$w.log; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r01_c3.csv python bench.py --steps 2 --warmup 3 --no-cpu --e2e-samples 256 > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:quad_sweep -s 3 -c 1 -f -o gpurun_out/prof_r01_c3_quad python bench.py --steps 2 --warmup 3 --no-cpu --e2e-samples 256 > gpurun_out/ncu_full_c3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:quad_sweep -s 3 -c 1 -f -o gpurun_out/prof_r01_c2_quad python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu --e2e-samples 256 > gpurun_out/ncu_full_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:quad_sweep -s 3 -c 1 -f -o gpurun_out/prof_r01_c4_quad python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu --e2e-samples 256 > gpurun_out/ncu_full_c4.log 2>&1
head -c 700 gpurun_out/bench_r01_c3.json; echo; for w in c3_pair c2 c4 c5 c1 ref; do head -c 200 gpurun_out/bench_r01_$w.json; echo; done
