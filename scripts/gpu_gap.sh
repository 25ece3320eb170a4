#!/bin/bash
# host timeline of a build and GPU idle between levels (HEM_GAP_DEBUG), then the plain bench
mkdir -p gpurun_out
HEM_GAP_DEBUG=1 timeout 600 python bench.py --skip-e2e --no-cpu-baseline --no-extra --steps 10 --warmup 3 2> gpurun_out/gap.err | cut -c1-120
grep "GPU idle" gpurun_out/gap.err | tail -7
grep "host timeline" gpurun_out/gap.err | tail -4
timeout 600 python bench.py --skip-e2e --no-cpu-baseline --no-extra --steps 20 --warmup 3 2>/dev/null | cut -c1-120
