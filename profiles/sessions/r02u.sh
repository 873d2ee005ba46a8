#!/bin/bash
# round 2, session u (8 GPUs): record buckets scaled to the device's own share; tests on distinct devices; N = 8 and N = 4 lines
mkdir -p gpurun_out
T=r02u
timeout 600 python -X faulthandler -m pytest tests/test_gpu_multi_device.py tests/test_gpu_peer.py -x -q --capture=sys > gpurun_out/${T}_tests.log 2>&1
tail -2 gpurun_out/${T}_tests.log | cut -c1-300
for N in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${T}_bench_n$N.json 2> gpurun_out/${T}_bench_n$N.err
done
python - <<PY
import json
T="${T}"
for f in ("bench_n8","bench_n4"):
    try:
        d=json.loads(open(f"gpurun_out/{T}_{f}.json").read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "FAILED", e); continue
    print("=====", f)
    for k in ["value","ms_per_step","scaling","e2e","gpu_launches","hits_per_step","phases_us_per_step","e2e_host_us_per_step_rank0","parity_sample","strong_scaling"]:
        if k in d: print(k, json.dumps(d.get(k))[:1700])
PY
