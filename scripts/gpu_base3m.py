"""Per-level stage timings of the 3M scene-scale full pyramid (BASELINE configs[2]) and the 1M x 5 pyramid."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gaussiansplattingregistration_b200.synth import synth
from gaussiansplattingregistration_b200.hem import HemContext

def timing(n, dens, lv, reps=4, aspect=(1, 1, 1)):
    c = synth(n, seed=1, density=dens, aspect=aspect)
    dev = [torch.from_numpy(a).cuda() for a in c.arrays()]
    ctx = HemContext(0, 7)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    ctx.set_option("timing", 1)
    for r in range(reps):
        torch.cuda.synchronize(); t = time.time()
        ctx.build(*dev, cluster_level=lv)
        torch.cuda.synchronize(); dt = time.time() - t
        print(f"== timing n={n} dens={dens} rep {r}: wall {dt*1e3:.2f} ms", flush=True)
    tot_in = 0; tot_ms = 0
    for l in range(lv):
        st = ctx.step_stats(l)
        tot_in += st["n_in"]; tot_ms += st["ms_total"]
        print(f"   step {l}: n_in {st['n_in']} np {st['n_parents']} P {st['n_pairs']} orph {st['n_orphans']} n_out {st['n_out']} cells {st['grid_cells']} h {st['cell_size']:.3f} maxR {st['max_radius']:.3f} | "
              f"ms tot {st['ms_total']:.3f} prep {st['ms_prep']:.3f} grid {st['ms_grid']:.3f} gate {st['ms_gate']:.3f} (k {st['ms_k_gate']:.3f}) sumlw {st['ms_sumlw']:.3f} mom {st['ms_moments']:.3f} (k {st['ms_k_moments']:.3f}) comp {st['ms_compact']:.3f} | {st['n_in']/st['ms_total']*1e3:.3e} G/s")
    print(f"   pyramid: {tot_in/tot_ms*1e3:.3e} Gaussians/s, sum ms {tot_ms:.3f}")
    ctx.close()

if __name__ == "__main__":
    timing(3_000_000, 5000, 8, aspect=(2, 1, 0.5))
    timing(1_000_000, 5000, 5)
