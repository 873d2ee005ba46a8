#!/bin/bash
# round 2, session l (1 GPU): full GPU tier, full bench line (streamed-database pass, host profile), config 4 with the real solver
mkdir -p gpurun_out
timeout 1200 python -X faulthandler -m pytest tests -m gpu -x -q --capture=sys > gpurun_out/r02l_tests.log 2>&1
tail -4 gpurun_out/r02l_tests.log | cut -c1-300
GSS_HOST_PROF=1 timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02l_bench.json 2> gpurun_out/r02l_bench.err
grep "host prof" gpurun_out/r02l_bench.err | head -40
timeout 400 python profiles/bench_config4_glucose.py --seconds 40 > gpurun_out/r02l_config4_glucose.json 2> gpurun_out/r02l_config4.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02l_bench.json").read().strip().splitlines()[-1])
for k in ["value","ms_per_step","e2e","phases_us_per_step","e2e_host_us_per_step","kernel_us","parity_sample","streamed_db","import_latency","roofline"]:
    print(k, json.dumps(d.get(k))[:1500])
c=json.load(open("gpurun_out/r02l_config4_glucose.json"))
for k,v in c.items():
    print(k, json.dumps(v)[:1800])
PY
