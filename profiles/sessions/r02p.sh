#!/bin/bash
# round 2, session p (1 GPU): why is the host side of a config-3 step milliseconds long?  (host profile with max /
# context switches / page faults per scope; with and without the nvidia-smi clock sampler; eager results off)
mkdir -p gpurun_out
T=r02p
nproc; cat /proc/loadavg; cat /sys/kernel/mm/transparent_hugepage/enabled; cat /proc/sys/kernel/numa_balancing
B="python bench.py --steps 10 --warmup 3 --no-cpu --no-dense --no-streamed --no-latency --no-ref-gpu"
GSS_HOST_PROF=1 timeout 300 $B > gpurun_out/${T}_a.json 2> gpurun_out/${T}_a.err
GSS_HOST_PROF=1 GSS_NO_CLOCK_SAMPLER=1 timeout 300 $B > gpurun_out/${T}_b.json 2> gpurun_out/${T}_b.err
GSS_HOST_PROF=1 GSS_NO_CLOCK_SAMPLER=1 GPUSHARE_EAGER_RESULTS=0 timeout 300 $B > gpurun_out/${T}_c.json 2> gpurun_out/${T}_c.err
for x in a b c; do
  echo "=== $x"
  python - <<PY
import json
d=json.loads(open("gpurun_out/${T}_$x.json").read().strip().splitlines()[-1])
print(d["e2e"]["ms_per_step"], d["e2e"]["ms_every_step"]); print(d["e2e_host_us_per_step"]); print(d["phases_us_per_step"])
PY
  awk '/reset/{n++} n==1' gpurun_out/${T}_$x.err | head -30
done
