#!/bin/bash
# round 2, session q (1 GPU): the driver's order -- GPU tier, smoke, reference arm, default bench -- with the host watch
# (context switches / steal) on the bench line; k_emit_scan, 16-byte PCIe loads in k_apply_direct, idle runs without apply;
# ncu --set full of the dense sliced kernel and the emit kernels
mkdir -p gpurun_out
T=r02q
timeout 1500 python -X faulthandler -m pytest tests -m gpu -x -q --capture=sys > gpurun_out/${T}_tests.log 2>&1
tail -3 gpurun_out/${T}_tests.log | cut -c1-300
ps aux --sort=-%cpu | head -8 | cut -c1-200; cat /proc/loadavg
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --impl reference > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err
cat /proc/loadavg
GSS_HOST_PROF=1 timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
awk '/reset/{n++} n==1' gpurun_out/${T}_bench.err | head -30
timeout 300 python bench.py --clauses 1000000 --vars 50000 --solvers 1 --slots 32 --no-cpu --no-ref-gpu --no-streamed --no-latency > gpurun_out/${T}_bench_config2.json 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_check_dense_sliced|k_emit|k_slice' \
   --launch-skip 12 -c 8 -o gpurun_out/${T}_dense -f python bench.py --steps 2 --warmup 3 --no-cpu --no-ref-gpu --no-streamed --no-latency \
   --prod-iters 2 --dense-iters 1 > gpurun_out/${T}_ncu.log 2>&1
tail -2 gpurun_out/${T}_ncu.log
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
        if k in d: print(k, json.dumps(d.get(k))[:900])
    for k in ["roofline","roofline_k_filter","reference_gpu","streamed_db","import_latency"]:
        if k in d: print(k, json.dumps(d.get(k))[:500])
PY
