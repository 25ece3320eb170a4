#!/bin/bash
N=${1:-2}
nvidia-smi -L | wc -l
if [ "$N" = "2" ]; then timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -2; fi
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
python -c "import sys,json; d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1]); print('N=$N 1M', '%.3e'%d['value'], d['ms_per_step'], '%.3e'%d['e2e']['value'], [round(l['ms'],3) for l in d['levels']], d['kernels'])" || tail -5 gpurun_out/bench_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 10 --warmup 3 --splats 3000000 --levels 3 > gpurun_out/bench3m_n$N.json 2> gpurun_out/bench3m_n$N.err
python -c "import sys,json; d=json.loads(open('gpurun_out/bench3m_n$N.json').read().strip().splitlines()[-1]); print('N=$N 3M', '%.3e'%d['value'], d['ms_per_step'], [round(l['ms'],3) for l in d['levels']], d['kernels'])" || tail -5 gpurun_out/bench3m_n$N.err
if [ "$2" = "10m" ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 5 --warmup 3 --splats 10000000 --levels 4 --skip-e2e --no-cpu-baseline > gpurun_out/bench10m_n$N.json 2> gpurun_out/bench10m_n$N.err
python -c "import sys,json; d=json.loads(open('gpurun_out/bench10m_n$N.json').read().strip().splitlines()[-1]); print('N=$N 10M', '%.3e'%d['value'], d['ms_per_step'], [round(l['ms'],3) for l in d['levels']], d['kernels'])" || tail -5 gpurun_out/bench10m_n$N.err
fi
