#!/bin/bash
# round 2, session x (1 GPU): k_emit_fused (sort + write in one launch, start-order tickets) against the two-kernel form
mkdir -p gpurun_out
T=r02x
timeout 900 python -X faulthandler -m pytest tests/test_gpu_direct_pipeline.py tests/test_gpu_scenarios.py tests/test_gpu_random_parity.py tests/test_gpu_kat.py tests/test_gpu_reduce_device.py -x -q --capture=sys > gpurun_out/${T}_tests.log 2>&1
tail -2 gpurun_out/${T}_tests.log | cut -c1-300
B="python bench.py --steps 20 --warmup 3 --no-streamed --no-latency --no-ref-gpu --no-dense --no-cpu"
timeout 600 $B > gpurun_out/${T}_bench_fused.json 2> gpurun_out/${T}_bench_fused.err
GSS_EMIT_SPLIT=1 timeout 600 $B > gpurun_out/${T}_bench_split.json 2> gpurun_out/${T}_bench_split.err
timeout 300 python bench.py --clauses 1000000 --vars 50000 --solvers 1 --slots 32 --no-cpu --no-ref-gpu --no-streamed --no-latency --no-dense > gpurun_out/${T}_bench_config2.json 2>/dev/null
python - <<PY
import json
for f in ("bench_fused","bench_split","bench_config2"):
    d=json.loads(open("gpurun_out/${T}_"+f+".json").read().strip().splitlines()[-1])
    print("==",f)
    for k in ["value","ms_per_step","gpu_launches","host_during_timed_region","device_step_complete","phases_us_per_step","kernel_us"]:
        if k in d: print(k, json.dumps(d.get(k))[:600])
    print("e2e", d["e2e"]["ms_per_step"])
PY
