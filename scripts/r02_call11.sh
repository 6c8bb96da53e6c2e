#!/bin/bash
mkdir -p gpurun_out/c11
nvidia-smi -L
( time python -m pytest tests/test_gpu_multi.py tests/test_gpu_planset.py -q -x ) > gpurun_out/c11/pytest.log 2>&1
tail -6 gpurun_out/c11/pytest.log | cut -c1-300
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 ) > gpurun_out/c11/bench2.json 2> gpurun_out/c11/bench2.err
tail -4 gpurun_out/c11/bench2.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/c11/bench2.json').read().strip().splitlines() if l.startswith('{')][-1])
print('main %.3e frac %.3f avg %.2f ms clocks %s' % (d['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['clocks']))
for s in d['secondary']:
    print(s.get('workload','')[:30], s.get('error') or ('%.3e frac %.3f ms/step %.1f launches %s x %s' % (s['value'], s['roofline']['frac'], s['ms_per_step'], s['config']['launches_per_gpu_per_step'], s['config']['samples_per_launch'])), s.get('config',{}).get('collective'))
e=d['e2e']; print('e2e', e['value'], e['api'], e['roofline']); print({k:(v.get('value'),v.get('d2h_gbs')) if isinstance(v,dict) else v for k,v in e.items() if k in ('capi_per_rank','capi_multi_one_process','drop_in_run_many','drop_in_run_many_arrays','drop_in_error')})
PY
