#!/bin/bash
# round 2, session z (1 GPU): config 4 alone -- the reference's real 32-thread GPU portfolio solver on the 2 M-variable /
# 8 M-clause instance, linked against this library and against the reference's own, 100 s each
mkdir -p gpurun_out
timeout 500 python profiles/bench_config4_glucose.py --seconds 100 > gpurun_out/r02z_config4_glucose.json 2> gpurun_out/r02z_config4.err
python - <<PY
import json
c=json.load(open("gpurun_out/r02z_config4_glucose.json"))
for k,v in c.items(): print(k, json.dumps(v)[:1400])
PY
