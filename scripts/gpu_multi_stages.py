"""Per-stage times of the sharded levels under torchrun (rank 0 prints)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from gaussiansplattingregistration_b200.synth import synth
from gaussiansplattingregistration_b200.hem import HemContext
rank, lr, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3_000_000
c = synth(n, seed=1, density=5000, aspect=(2, 1, 0.5))
dev = [torch.from_numpy(a).cuda() for a in c.arrays()]
ctx = HemContext(lr, 7); ctx.set_stream(torch.cuda.current_stream().cuda_stream); ctx.set_option("timing", 1)
ctx.init_distributed()
for k, v in [a.split("=") for a in sys.argv[2:]]:
    ctx.set_option(k, float(v))
for r in range(4):
    ctx.build(*dev, cluster_level=3)
torch.cuda.synchronize(); dist.barrier()
if rank == 0:
    for l in range(3):
        s = ctx.step_stats(l)
        print(f"world {world} level {l}: n {s['n_in']} shards {s['shards']} | tot {s['ms_total']:.3f} prep {s['ms_prep']:.3f} grid {s['ms_grid']:.3f} gate {s['ms_gate']:.3f} (k {s['ms_k_gate']:.3f}) sumlw {s['ms_sumlw']:.3f} mom {s['ms_moments']:.3f} (k {s['ms_k_moments']:.3f}) comp {s['ms_compact']:.3f} host {s['ms_host_enqueue']:.3f}")
dist.barrier(); dist.destroy_process_group()
