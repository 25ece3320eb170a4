#!/bin/bash
N=${1:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -3 gpurun_out/bench_n$N.err
python -c "import sys,json; d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1]); print('N=$N 1M', d['value'], d['ms_per_step'], d['e2e']['value'], [ (l['ms']) for l in d['levels']])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 5 --warmup 3 --splats 3000000 --levels 3 > gpurun_out/bench3m_n$N.json 2> gpurun_out/bench3m_n$N.err
python -c "import sys,json; d=json.loads(open('gpurun_out/bench3m_n$N.json').read().strip().splitlines()[-1]); print('N=$N 3M', d['value'], d['ms_per_step'], [ (l['ms']) for l in d['levels']])"
