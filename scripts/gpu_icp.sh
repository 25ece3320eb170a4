#!/bin/bash
# ICP consumer: GPU parity tests + the 1 M two-cloud benchmark.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_icp.py tests/test_gpu_icp_downstream.py -x -q -m gpu 2>&1 | tail -15
timeout 600 python scripts/gpu_icp_bench.py 1000000 > gpurun_out/icp_bench.json 2> gpurun_out/icp_bench.err; tail -3 gpurun_out/icp_bench.err; cat gpurun_out/icp_bench.json
