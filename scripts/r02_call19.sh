#!/bin/bash
mkdir -p gpurun_out/c19
nvidia-smi -L | wc -l; free -g | head -2
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 3 --warmup 3 ) > gpurun_out/c19/bench8.json 2> gpurun_out/c19/bench8.err
tail -4 gpurun_out/c19/bench8.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/c19/bench8.json').read().strip().splitlines() if l.startswith('{')][-1])
print('main %.3e frac %.3f avg %.2f ms clocks %s cfg %s' % (d['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['clocks'], {k:d['config'][k] for k in ('samples_per_gpu_per_step','launches_per_gpu_per_step','samples_per_launch')}))
print(d['roofline'].get('tail_launch'), d['roofline'].get('burst'))
for s in d['secondary']:
    print(s.get('workload','')[:30], s.get('error') or ('%.3e frac %.3f ms/step %.1f launches %s x %s' % (s['value'], s['roofline']['frac'], s['ms_per_step'], s['config']['launches_per_gpu_per_step'], s['config']['samples_per_launch'])))
e=d['e2e']; print('e2e', e['value'], e['api'], e['roofline']); print({k:(v.get('value'),v.get('d2h_gbs'),v.get('error')) if isinstance(v,dict) else v for k,v in e.items() if k in ('capi_per_rank','capi_multi_one_process','drop_in_run_many','drop_in_run_many_arrays','drop_in_error')})
PY
python -m pytest tests/test_gpu_planset.py tests/test_gpu_multi.py -q -x 2>&1 | tail -2
