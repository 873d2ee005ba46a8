#!/bin/bash
# round 2, session m (1 GPU): dense L2-sliced kernel (parity tests + timing), streamed pass with host profile,
# latency harness (quiet / saturated), config 4 with the real solver (no preprocessing), ncu traffic capture
mkdir -p gpurun_out
timeout 600 python -X faulthandler -m pytest tests/test_gpu_random_parity.py tests/test_gpu_kat.py tests/test_gpu_reduce_device.py -x -q --capture=sys > gpurun_out/r02m_tests.log 2>&1
tail -3 gpurun_out/r02m_tests.log | cut -c1-300
GSS_HOST_PROF=1 timeout 900 python bench.py --steps 20 --warmup 5 --no-ref-gpu > gpurun_out/r02m_bench.json 2> gpurun_out/r02m_bench.err
grep "host prof" gpurun_out/r02m_bench.err | head -120
GSS_DENSE_FLAT=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-ref-gpu --no-cpu --no-streamed --no-latency > gpurun_out/r02m_bench_dense_flat.json 2>/dev/null
timeout 100 tests/latency/latency_harness 16 200000 1000000 300 -1 999 > gpurun_out/r02m_latency16.jsonl 2>&1
timeout 400 python profiles/bench_config4_glucose.py --seconds 75 > gpurun_out/r02m_config4_glucose.json 2> gpurun_out/r02m_config4.err
timeout 700 python profiles/capture_traffic.py r02m > gpurun_out/r02m_traffic.log 2>&1
python - <<PY
import json
d=json.loads(open("gpurun_out/r02m_bench.json").read().strip().splitlines()[-1])
for k in ["value","ms_per_step","e2e","phases_us_per_step","e2e_host_us_per_step","kernel_us","streamed_db","import_latency","roofline"]:
    print(k, json.dumps(d.get(k))[:3000])
try:
    f=json.loads(open("gpurun_out/r02m_bench_dense_flat.json").read().strip().splitlines()[-1])
    print("dense flat", json.dumps(f.get("roofline"))[:600])
except Exception as e: print("flat", e)
print(open("gpurun_out/r02m_latency16.jsonl").read())
c=json.load(open("gpurun_out/r02m_config4_glucose.json"))
for k,v in c.items():
    print(k, json.dumps(v)[:1500])
try:
    t=json.load(open("gpurun_out/r02m_traffic.json")); print(json.dumps(t)[:2500])
except Exception as e: print("traffic", e, open("gpurun_out/r02m_traffic.log").read()[-800:])
PY
