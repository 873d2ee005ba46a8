#!/bin/bash
# round 2, session zb (1 GPU): directory searches past 128 distinct lengths (regression test), latency harness: binary probes
mkdir -p gpurun_out
timeout 120 python -X faulthandler -m pytest tests/test_gpu_scenarios.py tests/test_gpu_kat.py -x -q --capture=sys 2>&1 | tail -3 | cut -c1-400
timeout 70 tests/latency/latency_harness 64 200000 1000000 300 -1 999 > gpurun_out/r02zb_latency64_quiet.jsonl 2>&1; tail -1 gpurun_out/r02zb_latency64_quiet.jsonl | cut -c1-700
