#!/bin/bash
mkdir -p gpurun_out/c9
( time python -m pytest tests -m gpu -q ) > gpurun_out/c9/pytest.log 2>&1
tail -6 gpurun_out/c9/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time python bench.py --steps 4 --warmup 3 ) > gpurun_out/c9/bench.json 2> gpurun_out/c9/bench.err
tail -3 gpurun_out/c9/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c9/bench.json').read().strip().splitlines()[-1])
print('main %.3e frac %.3f avg %.2f ms clocks %s' % (d['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['clocks']))
for s in d['secondary']:
    print(s.get('workload','')[:30], s.get('error') or ('%.3e frac %.3f ms/step %.1f launches %s x %s %s' % (s['value'], s['roofline']['frac'], s['ms_per_step'], s['config']['launches_per_gpu_per_step'], s['config']['samples_per_launch'], s['clocks'])))
print('e2e', d['e2e']['value'], d['e2e']['roofline']['frac'], d['latency'])
PY
# sanitizers over the kernels changed this round (cluster launches, staged statistics, plan sets)
compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_quad.py tests/test_gpu_planset.py -q -x -k "cluster or histogram or reduced or planset or identical" > gpurun_out/c9/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/c9/sanitizer_memcheck.log
compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_quad.py -q -x -k "cluster or histogram" > gpurun_out/c9/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/c9/sanitizer_racecheck.log
