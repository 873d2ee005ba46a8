#!/bin/bash
# round 2, session ze (1 GPU): latency harness without the undefined-variable leak of its trail churn
mkdir -p gpurun_out
timeout 22 tests/latency/latency_harness 64 200000 1000000 300 -1 999 2>&1 | tail -1 > gpurun_out/r02ze_latency64_quiet.jsonl
timeout 22 tests/latency/latency_harness 64 200000 1000000 300 -1 985 2>&1 | tail -1 > gpurun_out/r02ze_latency64_saturated.jsonl
cut -c1-760 gpurun_out/r02ze_latency64_quiet.jsonl gpurun_out/r02ze_latency64_saturated.jsonl
