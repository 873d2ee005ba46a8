#!/bin/bash
# round 2, session za (1 GPU): the saturated regime of the latency harness (64 solvers) died in session y: stderr, fused vs split emit
mkdir -p gpurun_out
echo "== fused"; timeout 100 tests/latency/latency_harness 64 200000 1000000 300 -1 985 2>&1 | tail -5 | cut -c1-600
echo "== split"; GSS_EMIT_SPLIT=1 timeout 100 tests/latency/latency_harness 64 200000 1000000 300 -1 985 2>&1 | tail -3 | cut -c1-600
echo "== quiet fused"; timeout 100 tests/latency/latency_harness 64 200000 1000000 300 -1 999 2>&1 | tail -2 | cut -c1-600
