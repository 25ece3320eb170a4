#!/bin/bash
# full GPU suite + smoke + both bench arms (round-2 bench.py)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt
lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/gpu.txt
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 1500 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_ref.json | cut -c1-600
bash scripts/gpu_prof.sh 3m
timeout 900 bash scripts/gpu_sanitize.sh 2>&1 | tee gpurun_out/sanitize.log | tail -30
timeout 600 python bench.py --workload 10m --steps 5 --warmup 3 --skip-e2e --no-cpu-baseline --no-extra 2>&1 | tail -1 | tee gpurun_out/bench10m_n1.json | cut -c1-200
timeout 600 python bench.py --workload 1m --steps 10 --warmup 3 --no-cpu-baseline --no-extra 2>/dev/null | tail -1 | tee gpurun_out/bench1m_n1.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('1M x 5: value %.3e (%.3f ms)  e2e %.3e (%.2f ms)' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']))"
