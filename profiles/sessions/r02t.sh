#!/bin/bash
# round 2, session t (8 GPUs): the driver's N = 8 launch (weak line + strong sub-object, parity sample in both)
mkdir -p gpurun_out
T=r02t
nvidia-smi -L | wc -l; free -g | head -2
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/${T}_bench_n8.json 2> gpurun_out/${T}_bench_n8.err
tail -3 gpurun_out/${T}_bench_n8.err | cut -c1-300
python - <<PY
import json
T="${T}"
for f in ("bench_n8",):
    try:
        d=json.loads(open(f"gpurun_out/{T}_{f}.json").read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "FAILED", e); continue
    print("=====", f)
    for k in ["value","ms_per_step","scaling","e2e","gpu_launches","hits_per_step","phases_us_per_step","e2e_host_us_per_step_rank0","parity_sample","strong_scaling","clocks"]:
        if k in d: print(k, json.dumps(d.get(k))[:2000])
PY
