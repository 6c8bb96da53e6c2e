#!/bin/bash
# round 2, GPU call 1: phase-0 validation + baseline captures that round 1 never took
mkdir -p gpurun_out/c1
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/c1/pytest.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c1/smoke.log 2>&1
python - > gpurun_out/c1/d2h.log 2>&1 <<'PY'
import torch, time
for gb in (0.25, 1, 4):
    n = int(gb * (1 << 30))
    d = torch.empty(n, dtype=torch.uint8, device="cuda"); h = torch.empty(n, dtype=torch.uint8).pin_memory()
    for _ in range(2): h.copy_(d, non_blocking=True); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5): h.copy_(d, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    print(f"pinned D2H {gb} GiB: {n/dt/1e9:.1f} GB/s")
    t0 = time.perf_counter()
    for _ in range(5): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    print(f"pinned H2D {gb} GiB: {n/dt/1e9:.1f} GB/s")
import os; print("cpus", os.cpu_count()); os.system("free -g | head -2; lscpu | grep -E 'Model name|Socket|NUMA' ")
PY
for t in "c3 full 18944 --mt" "c3 reduced 18944" "c3 full 18944" "c5 reduced 378880 --spl 4" "c5 reduced 262144" "c4 reduced 32768" "c2 full 262144" "c3 full 2048" "c3 full 2048 --spl 4" "c3 full 128"; do python scripts/ncu_target.py $t; done > gpurun_out/c1/timings.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -s 1 -c 1 -f"
$NCU -k regex:quad_sweep -o gpurun_out/c1/c3_mt python scripts/ncu_target.py c3 full 18944 --mt --reps 1 > gpurun_out/c1/ncu_c3_mt.log 2>&1
$NCU -k regex:quad_sweep -o gpurun_out/c1/c3_red python scripts/ncu_target.py c3 reduced 18944 --reps 1 > gpurun_out/c1/ncu_c3_red.log 2>&1
$NCU -k regex:quad_sweep -o gpurun_out/c1/c5_red python scripts/ncu_target.py c5 reduced 37888 --spl 4 --reps 1 > gpurun_out/c1/ncu_c5_red.log 2>&1
tail -3 gpurun_out/c1/pytest.log; cat gpurun_out/c1/smoke.log gpurun_out/c1/d2h.log gpurun_out/c1/timings.log
