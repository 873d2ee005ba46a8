#!/bin/bash
# round 2, session s (2 GPUs): multi-device / peer tests on distinct devices, bench at N = 2 both ways
# (one process per GPU under torchrun = the driver's launch; one process, two devices = GPUSHARE_DEVICES)
mkdir -p gpurun_out
T=r02s
nvidia-smi -L
timeout 900 python -X faulthandler -m pytest tests/test_gpu_multi_device.py tests/test_gpu_peer.py tests/test_gpu_sharded.py -x -q --capture=sys > gpurun_out/${T}_tests.log 2>&1
tail -3 gpurun_out/${T}_tests.log | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/${T}_bench_n2.json 2> gpurun_out/${T}_bench_n2.err
tail -2 gpurun_out/${T}_bench_n2.err | cut -c1-300
timeout 600 python bench.py --devices 2 --steps 10 --warmup 3 --no-streamed --no-latency > gpurun_out/${T}_bench_devices2.json 2> gpurun_out/${T}_bench_devices2.err
timeout 600 python bench.py --devices 2 --scaling strong --steps 10 --warmup 3 --no-streamed --no-latency > gpurun_out/${T}_bench_devices2_strong.json 2>> gpurun_out/${T}_bench_devices2.err
python - <<PY
import json
T="${T}"
for f in ("bench_n2","bench_devices2","bench_devices2_strong"):
    try:
        d=json.loads(open(f"gpurun_out/{T}_{f}.json").read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "FAILED", e); continue
    print("=====", f)
    for k in ["value","ms_per_step","scaling","e2e","gpu_launches","hits_per_step","phases_us_per_step","e2e_host_us_per_step","e2e_host_us_per_step_rank0","parity_sample","strong_scaling","weak_scaling"]:
        if k in d: print(k, json.dumps(d.get(k))[:1500])
PY
