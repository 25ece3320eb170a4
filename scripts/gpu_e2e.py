import sys, time; sys.path.insert(0,'.')
import torch, numpy as np
from gaussiansplattingregistration_b200.synth import synth
from gaussiansplattingregistration_b200.hem import HemContext
c = synth(1_000_000, seed=1, density=5000)
host = [torch.from_numpy(a).pin_memory() for a in c.arrays()]
hn = [h.numpy() for h in host]
# raw PCIe rates
d = torch.empty_like(host[4], device="cuda")
for _ in range(2): d.copy_(host[4], non_blocking=True)
torch.cuda.synchronize(); t=time.perf_counter()
for _ in range(5): d.copy_(host[4], non_blocking=True)
torch.cuda.synchronize(); dt=(time.perf_counter()-t)/5
print("H2D pinned %.1f GB/s (%.2f ms for 180 MB)"%(host[4].numel()*4/dt/1e9, dt*1e3))
torch.cuda.synchronize(); t=time.perf_counter()
for _ in range(5): host[4].copy_(d, non_blocking=True)
torch.cuda.synchronize(); dt=(time.perf_counter()-t)/5
print("D2H pinned %.1f GB/s"%(host[4].numel()*4/dt/1e9))
ctx = HemContext(0, 7); ctx.set_option("timing", 1)
ctx.build(*hn, cluster_level=5)
sizes=[ctx.level_size(l) for l in range(6)]
outs=[]
for l in range(1,6):
    cap=sizes[l]+1024
    b=[torch.empty((cap,3)).pin_memory().numpy(), torch.empty((cap,3)).pin_memory().numpy(), torch.empty(cap).pin_memory().numpy(), torch.empty((cap,6)).pin_memory().numpy(), torch.empty((cap,45)).pin_memory().numpy()]
    outs.append(b); ctx.bind_output(l,*b)
for r in range(4):
    torch.cuda.synchronize(); t=time.perf_counter()
    ctx.build(*hn, cluster_level=5)
    dt=time.perf_counter()-t
    print("e2e build %.2f ms"%(dt*1e3), [round(ctx.step_stats(l)['ms_total'],3) for l in range(5)], "mom L0", round(ctx.step_stats(0)['ms_moments'],3), "prep L0", round(ctx.step_stats(0)['ms_prep'],3))
ctx.clear_outputs()
for r in range(3):
    torch.cuda.synchronize(); t=time.perf_counter()
    ctx.build(*hn, cluster_level=5)
    t1=time.perf_counter()
    for l in range(1,6): ctx.level_into(l, *outs[l-1])
    t2=time.perf_counter()
    print("no-bind: build %.2f ms, readback %.2f ms"%((t1-t)*1e3,(t2-t1)*1e3))
dev=[torch.from_numpy(a).cuda() for a in c.arrays()]
for l in range(1,6): ctx.bind_output(l,*outs[l-1])
for r in range(3):
    torch.cuda.synchronize(); t=time.perf_counter()
    ctx.build(*dev, cluster_level=5)
    print("device-in + bound out: %.2f ms"%((time.perf_counter()-t)*1e3), [round(ctx.step_stats(l)['ms_total'],3) for l in range(5)])
for l in (1,2):
    s=ctx.step_stats(l); print("bound-out level",l,{k:round(s[k],3) for k in ("ms_total","ms_prep","ms_grid","ms_gate","ms_sumlw","ms_moments","ms_compact","ms_k_gate","ms_k_moments")})
