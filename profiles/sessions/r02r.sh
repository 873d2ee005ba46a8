#!/bin/bash
# round 2, session r (1 GPU): per-solver k_emit_scan (parity tests of the emit path), level-1 variant sweep incl. the lean
# variants, bench line
mkdir -p gpurun_out
T=r02r
timeout 900 python -X faulthandler -m pytest tests/test_gpu_direct_pipeline.py tests/test_gpu_scenarios.py tests/test_gpu_random_parity.py tests/test_gpu_kat.py tests/test_gpu_peer.py -x -q --capture=sys > gpurun_out/${T}_tests.log 2>&1
tail -3 gpurun_out/${T}_tests.log | cut -c1-300
GSS_HOST_PROF=1 timeout 900 python bench.py --filter-sweep --no-streamed --no-latency --no-ref-gpu > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
awk '/reset/{n++} n==1' gpurun_out/${T}_bench.err | grep -E "launch|bump|hand|prepare|collect" | head -30
python - <<PY
import json
T="${T}"
d=json.loads(open(f"gpurun_out/{T}_bench.json").read().strip().splitlines()[-1])
for k in ["value","ms_per_step","e2e","gpu_launches","host_during_timed_region","remeasured","device_step_complete","phases_us_per_step","e2e_host_us_per_step","kernel_us","parity_sample","filter_variants","exact_variants"]:
    if k in d: print(k, json.dumps(d.get(k))[:2500])
PY
