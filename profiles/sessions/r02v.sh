#!/bin/bash
# round 2, session v (2 GPUs): workers keep their record keys / masks unless asked (every rank bumps its own hits);
# peer tests with the export switched per step, N = 2 line
mkdir -p gpurun_out
T=r02v
timeout 600 python -X faulthandler -m pytest tests/test_gpu_peer.py tests/test_gpu_multi_device.py -x -q --capture=sys > gpurun_out/${T}_tests.log 2>&1
tail -3 gpurun_out/${T}_tests.log | cut -c1-400
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/${T}_bench_n2.json 2> gpurun_out/${T}_bench_n2.err
tail -2 gpurun_out/${T}_bench_n2.err | cut -c1-300
python - <<PY
import json
T="${T}"
d=json.loads(open(f"gpurun_out/{T}_bench_n2.json").read().strip().splitlines()[-1])
for k in ["value","ms_per_step","scaling","e2e","gpu_launches","hits_per_step","phases_us_per_step","e2e_host_us_per_step_rank0","parity_sample","strong_scaling"]:
    if k in d: print(k, json.dumps(d.get(k))[:1300])
PY
