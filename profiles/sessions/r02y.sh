#!/bin/bash
# round 2, session y (1 GPU): final evidence run in the driver's order (GPU tier, smoke, reference arm, default bench) plus
# config 2, latency harness at 16 solvers, launch list + DRAM traffic of the bench command, ncu --set full of the check kernels
# and of the dense sliced kernel, config 4 (the real 32-thread glucose portfolio on the 2 M-variable instance)
mkdir -p gpurun_out
T=r02y
timeout 1500 python -X faulthandler -m pytest tests -m gpu -x -q --capture=sys > gpurun_out/${T}_tests.log 2>&1
tail -3 gpurun_out/${T}_tests.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --impl reference > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err
GSS_HOST_PROF=1 timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
timeout 300 python bench.py --clauses 1000000 --vars 50000 --solvers 1 --slots 32 --no-cpu --no-ref-gpu --no-streamed --no-latency > gpurun_out/${T}_bench_config2.json 2>/dev/null
timeout 100 tests/latency/latency_harness 16 200000 1000000 300 -1 999 > gpurun_out/${T}_latency16.jsonl 2>&1
timeout 600 python profiles/capture_traffic.py ${T} > gpurun_out/${T}_traffic.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_filter|k_exact|k_emit|k_apply_direct' \
   --launch-skip 24 -c 6 -o gpurun_out/${T}_check -f python bench.py --steps 3 --warmup 3 --no-cpu --no-ref-gpu --no-streamed --no-latency \
   --no-dense --prod-iters 1 > gpurun_out/${T}_ncu.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_check_dense_sliced' -c 2 -o gpurun_out/${T}_dense -f \
   python bench.py --steps 1 --warmup 3 --no-cpu --no-ref-gpu --no-streamed --no-latency --prod-iters 1 --dense-iters 1 > gpurun_out/${T}_ncu2.log 2>&1
tail -1 gpurun_out/${T}_ncu.log; tail -1 gpurun_out/${T}_ncu2.log
timeout 400 python profiles/bench_config4_glucose.py --seconds 45 > gpurun_out/${T}_config4_glucose.json 2> gpurun_out/${T}_config4.err
python - <<PY
import json
T="${T}"
for f in ("bench","bench_reference","bench_config2"):
    try:
        d=json.loads(open(f"gpurun_out/{T}_{f}.json").read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "FAILED", e); continue
    print("=====", f)
    for k in ["value","ms_per_step","e2e","gpu_launches","clocks","host_during_timed_region","remeasured","device_step_complete","phases_us_per_step","e2e_host_us_per_step","kernel_us","cpu_baseline","parity_sample"]:
        if k in d: print(k, json.dumps(d.get(k))[:700])
    for k in ["roofline","roofline_k_filter","reference_gpu","streamed_db","import_latency"]:
        if k in d: print(k, json.dumps(d.get(k))[:400])
print(open(f"gpurun_out/{T}_latency16.jsonl").read()[-700:])
try:
    c=json.load(open(f"gpurun_out/{T}_config4_glucose.json"))
    for k,v in c.items(): print(k, json.dumps(v)[:900])
except Exception as e: print("config4", e, open(f"gpurun_out/{T}_config4.err").read()[-500:])
PY
