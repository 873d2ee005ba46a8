#!/bin/bash
# round 2, session i (2 GPUs): full GPU tier (merged hand-over order), 2-GPU bench with rank 0's host phases
mkdir -p gpurun_out
timeout 1200 python -X faulthandler -m pytest tests -m gpu -x -q > gpurun_out/r02i_tests.log 2>&1
grep -n "passed\|failed\|Fatal\|FAILED\|Error" gpurun_out/r02i_tests.log | tail -15
grep -n -B5 -A40 "Fatal Python" gpurun_out/r02i_tests.log | head -90
GSS_PEER_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02i_bench_n2.json 2> gpurun_out/r02i_bench_n2.err
grep "peer trace" gpurun_out/r02i_bench_n2.err | sed -n '20,30p'
python - <<PY
import json
for f in ["r02i_bench_n2"]:
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "no line", e); continue
    print("==",f)
    for k in ["value","ms_per_step","scaling","e2e","phases_us_per_step","e2e_host_us_per_step_rank0","hits_per_step","parity_sample","strong_scaling"]:
        print(k, json.dumps(d.get(k))[:1800])
PY
