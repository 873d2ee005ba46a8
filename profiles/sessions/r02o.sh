#!/bin/bash
# round 2, session o (1 GPU): the round's evidence run after the container was re-created — full GPU tier, the default
# bench line (both arms), launch list + DRAM traffic, one `ncu --set full` capture of the check kernels, latency harness
mkdir -p gpurun_out
T=r02o
timeout 1500 python -X faulthandler -m pytest tests -m gpu -x -q --capture=sys > gpurun_out/${T}_tests.log 2>&1
tail -3 gpurun_out/${T}_tests.log | cut -c1-300
timeout 600 python bench.py --impl reference > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err
GSS_HOST_PROF=1 timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
grep "host prof" gpurun_out/${T}_bench.err | tail -40
timeout 300 python bench.py --clauses 1000000 --vars 50000 --solvers 1 --slots 32 --no-cpu --no-ref-gpu --no-streamed --no-latency > gpurun_out/${T}_bench_config2.json 2>/dev/null
timeout 100 tests/latency/latency_harness 16 200000 1000000 300 -1 999 > gpurun_out/${T}_latency16.jsonl 2>&1
timeout 700 python profiles/capture_traffic.py ${T} > gpurun_out/${T}_traffic.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_filter|k_exact|k_check_dense_sliced|k_emit|k_apply_direct' \
   --launch-skip 30 -c 8 -o gpurun_out/${T}_check -f python bench.py --steps 4 --warmup 3 --no-cpu --no-ref-gpu --no-streamed --no-latency \
   --prod-iters 2 --dense-iters 1 > gpurun_out/${T}_ncu.log 2>&1
tail -3 gpurun_out/${T}_ncu.log
python - <<PY
import json
T="${T}"
for f in ("bench","bench_reference","bench_config2"):
    try:
        d=json.loads(open(f"gpurun_out/{T}_{f}.json").read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "FAILED", e); continue
    print("=====", f)
    for k in ["value","ms_per_step","e2e","gpu_launches","clocks","roofline","roofline_k_filter","device_step_complete","phases_us_per_step","e2e_host_us_per_step","kernel_us","cpu_baseline","reference_gpu","streamed_db","import_latency","parity_sample"]:
        if k in d: print(k, json.dumps(d.get(k))[:1800])
print(open(f"gpurun_out/{T}_latency16.jsonl").read()[-1500:])
try:
    t=json.load(open(f"gpurun_out/{T}_traffic.json")); print(json.dumps(t)[:2500])
except Exception as e: print("traffic", e, open(f"gpurun_out/{T}_traffic.log").read()[-800:])
PY
