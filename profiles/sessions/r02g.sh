#!/bin/bash
# round 2, session g (2 GPUs): full GPU tier on the working tree, multi-process exchange with results through
# shared host memory (peer.cu), 2-GPU bench (weak + strong), in-process 2-device bench
mkdir -p gpurun_out
nvidia-smi -L; df -h /dev/shm | tail -1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02g_tests.log; cat gpurun_out/r02g_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02g_bench_n2.json 2> gpurun_out/r02g_bench_n2.err; tail -5 gpurun_out/r02g_bench_n2.err
timeout 300 python bench.py --devices 2 --steps 20 --warmup 5 --no-ref-gpu --no-latency --no-dense --no-cpu > gpurun_out/r02g_bench_dev2.json 2> gpurun_out/r02g_bench_dev2.err; tail -3 gpurun_out/r02g_bench_dev2.err
python - <<PY
import json
for f in ["r02g_bench_n2","r02g_bench_dev2"]:
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "no line", e); continue
    print("==",f)
    for k in ["value","ms_per_step","scaling","e2e","phases_us_per_step","hits_per_step","parity_sample","strong_scaling"]:
        print(k, json.dumps(d.get(k))[:1500])
PY
