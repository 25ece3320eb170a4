"""Developer check run on the GPU box: CUDA path vs oracle on a few configs + stage timings."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaussiansplattingregistration_b200.synth import synth
from gaussiansplattingregistration_b200.hem import HemContext
from oracle import porthem, refhem

def relerr(a, b):
    a = a.astype(np.float64); b = b.astype(np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))

def check(n, dens, lv, seed=7, use_ref=False):
    c = synth(n, seed=1, density=dens)
    ctx = HemContext(0, seed)
    ctx.set_option("keep_pairs", 1)
    ctx.set_option("timing", 1)
    t = time.time(); ctx.build(*c.arrays(), cluster_level=lv); tg = time.time() - t
    t = time.time(); p = porthem.run(*c.arrays(), levels=lv, seed=seed); tp = time.time() - t
    print(f"== n={n} dens={dens} levels={lv}: gpu {tg*1e3:.1f} ms (cold), port {tp:.2f} s")
    print("   sizes gpu", [ctx.level_size(l) for l in range(lv + 1)], "port", [p.size(l) for l in range(lv + 1)])
    ok = True
    for l in range(lv):
        gp, pc = ctx.step_pairs(l), p.children(l)
        st = ctx.step_stats(l)
        same_par = np.array_equal(gp["parent_idx"], pc["parent_idx"])
        same_off = np.array_equal(gp["off"], pc["off"])
        same_ch = False
        if same_par and same_off:
            # sort gpu children per parent
            ch = gp["child"].copy(); wl = gp["wl"].copy()
            for s in range(len(gp["parent_idx"])):
                a, b = gp["off"][s], gp["off"][s + 1]
                o = np.argsort(ch[a:b], kind="stable"); ch[a:b] = ch[a:b][o]; wl[a:b] = wl[a:b][o]
            same_ch = np.array_equal(ch, pc["child"])
            wlerr = relerr(wl, pc["wl"]) if same_ch else -1
            slerr = relerr(gp["sumlw"], pc["sumlw"]) if same_ch else -1
            raderr = float(np.max(np.abs(gp["radius"].view(np.int32).astype(np.int64) - pc["radius"].view(np.int32)))) if len(pc["radius"]) else 0
        else:
            wlerr = slerr = raderr = -1
        print(f"   step {l}: parents eq {same_par} off eq {same_off} children eq {same_ch} wl {wlerr:.1e} sumlw {slerr:.1e} radius ulp {raderr} | "
              f"np {st['n_parents']} P {st['n_pairs']} orph {st['n_orphans']} drop {st['n_dropped']} cells {st['grid_cells']} h {st['cell_size']:.3f} "
              f"ms tot {st['ms_total']:.3f} prep {st['ms_prep']:.3f} grid {st['ms_grid']:.3f} gate {st['ms_gate']:.3f} sumlw {st['ms_sumlw']:.3f} mom {st['ms_moments']:.3f} comp {st['ms_compact']:.3f}")
        ok &= same_par and same_off and same_ch
        if not (same_par and same_off and same_ch): break
    for l in range(lv + 1):
        if ctx.level_size(l) != p.size(l): print("   size mismatch level", l); ok = False; break
        a, b = ctx.level(l, with_internal=True), p.level(l)
        errs = {k: relerr(a[k], b[k]) for k in ["xyz", "colors", "opacity", "cov6", "sh", "weight"]}
        fe = np.array_equal(a["is_parent"], b["is_parent"])
        print(f"   level {l}: flags eq {fe}", {k: f"{v:.1e}" for k, v in errs.items()})
        ok &= fe and max(errs.values()) < 1e-4
    print("   RESULT", "OK" if ok else "MISMATCH")
    return ok

def timing(n, dens, lv, reps=3, aspect=(1, 1, 1)):
    c = synth(n, seed=1, density=dens, aspect=aspect)
    import torch
    dev = [torch.from_numpy(a).cuda() for a in c.arrays()]
    ctx = HemContext(0, 7)
    ctx.set_option("timing", 1)
    for r in range(reps):
        torch.cuda.synchronize(); t = time.time()
        ctx.build(*dev, cluster_level=lv)
        torch.cuda.synchronize(); dt = time.time() - t
        print(f"== timing n={n} dens={dens} rep {r}: wall {dt*1e3:.2f} ms")
    tot_in = 0; tot_ms = 0
    for l in range(lv):
        st = ctx.step_stats(l)
        tot_in += st["n_in"]; tot_ms += st["ms_total"]
        print(f"   step {l}: n_in {st['n_in']} np {st['n_parents']} P {st['n_pairs']} orph {st['n_orphans']} n_out {st['n_out']} cells {st['grid_cells']} h {st['cell_size']:.3f} maxR {st['max_radius']:.3f} | "
              f"ms tot {st['ms_total']:.3f} prep {st['ms_prep']:.3f} grid {st['ms_grid']:.3f} gate {st['ms_gate']:.3f} sumlw {st['ms_sumlw']:.3f} mom {st['ms_moments']:.3f} comp {st['ms_compact']:.3f} | {st['n_in']/st['ms_total']*1e3:.3e} G/s")
    print(f"   pyramid: {tot_in/tot_ms*1e3:.3e} Gaussians/s (sum n_in / sum ms)")

if __name__ == "__main__":
    ok = True
    ok &= check(3000, 20000, 3)
    ok &= check(20000, 5000, 4)
    ok &= check(100000, 5000, 3)
    ok &= check(50000, 1000, 3)
    timing(1_000_000, 5000, 5)
    timing(3_000_000, 5000, 3, reps=2, aspect=(2, 1, 0.5))
    print("ALL", "OK" if ok else "MISMATCH")
