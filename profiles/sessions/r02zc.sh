#!/bin/bash
# round 2, session zc (1 GPU): latency harness with 100 unmeasured warm-up probes, both regimes at 64 solvers, 16 solvers quiet
mkdir -p gpurun_out
timeout 50 tests/latency/latency_harness 64 200000 1000000 300 -1 999 2>&1 | tail -1 > gpurun_out/r02zd_latency64_quiet.jsonl
timeout 50 tests/latency/latency_harness 64 200000 1000000 300 -1 985 2>&1 | tail -1 > gpurun_out/r02zd_latency64_saturated.jsonl
cut -c1-760 gpurun_out/r02zd_latency64_quiet.jsonl gpurun_out/r02zd_latency64_saturated.jsonl
