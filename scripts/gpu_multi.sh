#!/bin/bash
# multi-GPU visit (gpurun --gpus N): bit-identity test, then the driver's bench command at N ranks
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/multi_test_n$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
python - <<PY || tail -5 gpurun_out/bench_n$N.err
import json
d = json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1])
print('N=$N 3M: %.3e G/s  %.3f ms/step  e2e %.3e (%.2f ms)' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']))
print('  levels', [(l['shards'], round(l['ms'], 3)) for l in d['levels']])
print('  kernels', {k: (round(v['frac'], 4), round(v['ms_per_step'], 3)) for k, v in d['kernels'].items() if 'ms_per_step' in v})
print('  checksums', d['checksums'][:3], '1M:', d['workloads'].get('1m', {}).get('ms_per_step'), d['workloads'].get('1m', {}).get('checksums', [None])[0])
PY
if [ "$2" = "10m" ]; then
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 5 --warmup 3 --workload 10m --skip-e2e --no-cpu-baseline --no-extra > gpurun_out/bench10m_n$N.json 2> gpurun_out/bench10m_n$N.err
tail -c 300 gpurun_out/bench10m_n$N.json || tail -5 gpurun_out/bench10m_n$N.err
fi
