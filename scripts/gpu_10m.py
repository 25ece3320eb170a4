import sys, time; sys.path.insert(0,'.')
import numpy as np, torch
from gaussiansplattingregistration_b200.synth import synth
from gaussiansplattingregistration_b200.hem import HemContext
n = 10_000_000
t=time.time(); c = synth(n, seed=1, density=5000.0, aspect=(2,1,0.5)); print("synth %.1f s"%(time.time()-t))
dev = [torch.from_numpy(a).cuda() for a in c.arrays()]
ctx = HemContext(0, 7); ctx.set_option("timing", 1)
for r in range(2):
    torch.cuda.synchronize(); t=time.perf_counter(); ctx.build(*dev, cluster_level=4); print("build %.2f ms"%((time.perf_counter()-t)*1e3))
tot_n=tot_ms=0
for l in range(4):
    s = ctx.step_stats(l); tot_n+=s['n_in']; tot_ms+=s['ms_total']
    print(f"level {l}: n_in {s['n_in']} np {s['n_parents']} P {s['n_pairs']} n_out {s['n_out']} ms {s['ms_total']:.3f} gate {s['ms_k_gate']:.3f} mom {s['ms_k_moments']:.3f}")
print("pyramid %.3e Gaussians/s; mem allocated by torch %.1f GB, device used %.1f GB"%(tot_n/tot_ms*1e3, torch.cuda.memory_allocated()/1e9, (torch.cuda.mem_get_info()[1]-torch.cuda.mem_get_info()[0])/1e9))
lv = ctx.level_torch(4, with_internal=True)
print("level 4 mass", float(lv["weight"].double().sum()), "min weight", float(lv["weight"].min()), "finite", bool(torch.isfinite(lv["xyz"]).all()))
