#!/bin/bash
# round 2, session w (1 GPU): capacities grow ahead of the counts (no overflow re-run by fluctuation); emit / overflow tests, short bench
mkdir -p gpurun_out
T=r02w
timeout 900 python -X faulthandler -m pytest tests/test_gpu_direct_pipeline.py tests/test_gpu_scenarios.py tests/test_gpu_random_parity.py tests/test_gpu_kat.py -x -q --capture=sys > gpurun_out/${T}_tests.log 2>&1
tail -2 gpurun_out/${T}_tests.log | cut -c1-300
GSS_HOST_PROF=1 timeout 600 python bench.py --steps 30 --warmup 3 --no-streamed --no-latency --no-ref-gpu --no-dense --no-cpu > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${T}_bench.json").read().strip().splitlines()[-1])
for k in ["value","ms_per_step","e2e","gpu_launches","host_during_timed_region","remeasured","device_step_complete","phases_us_per_step","e2e_host_us_per_step","kernel_us"]:
    if k in d: print(k, json.dumps(d.get(k))[:1500])
PY
