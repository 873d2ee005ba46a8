#!/bin/bash
# round 2, session n (1 GPU): full GPU tier; in-place host mirrors, persistent spare arenas, run header fused into
# k_apply_direct; streamed pass with host profile; latency harness with probes that get satisfied; config 4 (3-SAT-like instance)
mkdir -p gpurun_out
timeout 1200 python -X faulthandler -m pytest tests -m gpu -x -q --capture=sys > gpurun_out/r02n_tests.log 2>&1
tail -3 gpurun_out/r02n_tests.log | cut -c1-300
GSS_HOST_PROF=1 timeout 900 python bench.py --steps 20 --warmup 5 --no-ref-gpu > gpurun_out/r02n_bench.json 2> gpurun_out/r02n_bench.err
grep "host prof" gpurun_out/r02n_bench.err | head -120
timeout 100 tests/latency/latency_harness 16 200000 1000000 300 -1 999 > gpurun_out/r02n_latency16.jsonl 2>&1
timeout 500 python profiles/bench_config4_glucose.py --seconds 110 > gpurun_out/r02n_config4_glucose.json 2> gpurun_out/r02n_config4.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02n_bench.json").read().strip().splitlines()[-1])
for k in ["value","ms_per_step","e2e","phases_us_per_step","e2e_host_us_per_step","kernel_us","streamed_db","import_latency","roofline"]:
    print(k, json.dumps(d.get(k))[:3000])
try:
    f=json.loads(open("gpurun_out/r02n_bench_dense_flat.json").read().strip().splitlines()[-1])
    print("dense flat", json.dumps(f.get("roofline"))[:600])
except Exception as e: print("flat", e)
print(open("gpurun_out/r02n_latency16.jsonl").read())
c=json.load(open("gpurun_out/r02n_config4_glucose.json"))
for k,v in c.items():
    print(k, json.dumps(v)[:1500])
try:
    t=json.load(open("gpurun_out/r02n_traffic.json")); print(json.dumps(t)[:2500])
except Exception as e: print("traffic", e, open("gpurun_out/r02n_traffic.log").read()[-800:])
PY
