import sys, time; sys.path.insert(0,'.')
import torch
from gaussiansplattingregistration_b200.synth import synth
from gaussiansplattingregistration_b200.hem import HemContext
c = synth(1_000_000, seed=1, density=5000)
dev = [torch.from_numpy(a).cuda() for a in c.arrays()]
ctx = HemContext(0, 7); ctx.set_option("timing", 1)
for r in range(3): ctx.build(*dev, cluster_level=5)
for l in range(5):
    s = ctx.step_stats(l)
    print(f"level {l}: n {s['n_in']} gpu ms {s['ms_total']:.3f} host enqueue ms {s['ms_host_enqueue']:.3f}")
ctx2 = HemContext(0, 7)    # without timing events
for r in range(3): ctx2.build(*dev, cluster_level=5)
print("no-timing host enqueue:", [round(ctx2.step_stats(l)['ms_host_enqueue'],3) for l in range(5)])
