#!/bin/bash
mkdir -p gpurun_out/c10
( time python -m pytest tests/test_gpu_analytic.py tests/test_gpu_reference_verbatim.py tests/test_gpu_planset.py -q -x ) > gpurun_out/c10/pytest.log 2>&1
tail -40 gpurun_out/c10/pytest.log | cut -c1-300
