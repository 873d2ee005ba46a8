#!/bin/bash
# round 2, session f: TMA-staged k_filter vs the register variants, staged k_exact appends, vectorised k_emit_write,
# import-latency harness, ncu launch list
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_direct_pipeline.py tests/test_gpu_random_parity.py tests/test_gpu_kat.py -q -x 2>&1 | tail -5 > gpurun_out/r02f_tests.log; tail -5 gpurun_out/r02f_tests.log
GSS_FILTER_VARIANT=10 timeout 300 python -m pytest tests/test_gpu_random_parity.py tests/test_gpu_kat.py tests/test_gpu_config3.py -q -x 2>&1 | tail -5 > gpurun_out/r02f_tests_tma.log; tail -5 gpurun_out/r02f_tests_tma.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-ref-gpu --no-latency --filter-sweep > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; tail -3 gpurun_out/r02f_bench.err
GSS_FILTER_VARIANT=10 timeout 300 python bench.py --steps 20 --warmup 5 --no-ref-gpu --no-latency --no-dense --no-cpu > gpurun_out/r02f_bench_tma.json 2>> gpurun_out/r02f_bench.err
timeout 100 tests/latency/latency_harness 64 200000 1000000 300 -1 > gpurun_out/r02f_latency.jsonl 2>&1
timeout 100 tests/latency/latency_harness 16 200000 1000000 300 -1 >> gpurun_out/r02f_latency.jsonl 2>&1
cat gpurun_out/r02f_latency.jsonl; nproc
timeout 300 python profiles/capture_traffic.py r02f > gpurun_out/r02f_traffic.log 2>&1
python - <<EOF
import json
for f in ["r02f_bench","r02f_bench_tma"]:
    d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
    print("==",f)
    for k in ["value","ms_per_step","e2e","phases_us_per_step","e2e_host_us_per_step","kernel_us","parity_sample","filter_variants"]:
        print(k, json.dumps(d.get(k))[:1800])
t=json.load(open("gpurun_out/r02f_traffic.json"))
print({k:round(v,1) for k,v in t["us_per_launch_under_ncu"].items()})
EOF
