#!/usr/bin/env python
"""bench.py — HEM pyramid throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload 3m|1m|10m]

A "step" is one full pass of the hot path: build the whole HEM pyramid of one synthetic SH-degree-3 splat
cloud with the inputs already resident in HBM.  The headline workload is BASELINE.json configs[2], the
configuration the metric is quoted on: the 3M-Gaussian scene-scale cloud (box 2:1:0.5, 5000 splats per
unit^3), full pyramid = 8 levels (the level sizes stop falling below ~60k).  The 1M x 5
pyramid (configs[1]) is measured in the same run and reported under `workloads`.
`value` = (sum over levels of the level's input Gaussians) * steps / timed seconds, i.e. pyramid
throughput SUM(N_l)/SUM(t_l) in Gaussians/s; under torchrun ONE cloud is sharded over the ranks (strong
scaling), every rank ends with the same pyramid and `checksums` proves it across N.
`e2e` is the same metric through the public API with pinned HOST buffers: H2D of the cloud and D2H of
every returned level inside the timed region.  `--impl reference` times the reference's own CPU
implementation (oracle/_ref, the unmodified cpp_ext; else the C port) on the box's host cores on a
bounded SAMPLE of the workload and says so in `config` (`n` is the size it ran).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "HEM Gaussians reduced/sec per level (pyramid SUM N_l / SUM t_l)"
UNIT = "Gaussians/s"
ROW_BYTES = 236            # 59 floats per splat at SH degree 3

WORKLOADS = {
    # name: (n, levels, aspect, BASELINE.json config)
    "3m": (3_000_000, 8, (2.0, 1.0, 0.5), "configs[2]: HEM full pyramid, synthetic 3M-Gaussian scene-scale splat"),
    "1m": (1_000_000, 5, (1.0, 1.0, 1.0), "configs[1]: HEM 5-level pyramid, synthetic 1M-Gaussian splat"),
    "10m": (10_000_000, 8, (2.0, 1.0, 0.5), "configs[3]: HEM full pyramid, synthetic 10M-Gaussian splat"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="3m", choices=sorted(WORKLOADS))
    ap.add_argument("--splats", dest="n", type=int, default=0, help="override the workload's size")
    ap.add_argument("--levels", type=int, default=0, help="override the workload's level count")
    ap.add_argument("--density", type=float, default=5000.0)
    ap.add_argument("--cpu-sample", type=int, default=200_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary workload and the list-marshalling baseline")
    ap.add_argument("--skip-e2e", action="store_true", help="profiling runs only")
    a = ap.parse_args()
    n, lv, aspect, cfg = WORKLOADS[a.workload]
    a.n = a.n or n
    a.levels = a.levels or lv
    a.aspect = aspect
    a.cfgname = cfg
    return a


def workload_config(a, world, n=None, sample=False):
    n = a.n if n is None else n
    c = {"workload": f"HEM {a.levels}-level pyramid, synthetic {n}-Gaussian SH-deg-3 splat, box {a.aspect[0]:g}:{a.aspect[1]:g}:{a.aspect[2]:g}, "
                     f"density {a.density:g}/unit^3, geometric+colour distance (BASELINE {a.cfgname})",
         "n": n, "levels": a.levels, "density": a.density, "aspect": list(a.aspect), "sh_degree": 3,
         "params": {"hem_reduction": 3.0, "distance_delta": 3.0, "color_delta": 2.5, "decay_rate": 1.0},
         "parallelism": "single GPU" if world == 1 else f"{world} ranks: one cloud, parents sharded by spatial slab (grid-cell order), per-level NCCL exchange; every rank ends with the full pyramid",
         "l2": "no flush: the level-0 working set (geometry 64 B + SH 180 B per splat + sort/pair buffers) exceeds the 126 MB L2"}
    if sample:
        c["sample_of_n"] = a.n
        c["same_config"] = False
        c["note"] = (f"bounded sample: the CPU reference ran {n} Gaussians of the {a.n}-Gaussian workload (same generator, density, "
                     "parameters and level count); it gets slower per Gaussian as N grows (BASELINE.md §2), so the sample flatters it")
    return c


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region.  NVML is polled from a thread
    every ~2 ms (the timed region of the default run is tens of milliseconds; `nvidia-smi -lms`
    needs longer than that to print its first line and is only the fallback)."""
    REASONS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20}

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.sm, self.reasons, self.max_sm = [], set(), None
        self._stop = threading.Event()
        self.thread = None
        self.nvml = None
        self.proc = None
        self.rows = []

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.idx]) if vis and vis.split(",")[self.idx].isdigit() else self.idx
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                 "clocks_event_reasons.sw_power_cap")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while not self._stop.is_set():
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)))
                r = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for name, bit in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        self._stop.set()
        if self.nvml is not None:
            self.thread.join(timeout=1)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_sm,
                    "samples": len(self.sm), "reasons": sorted(self.reasons), "source": "nvml, polled during the timed region"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["no clock source available"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx = [], []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    self.reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(self.reasons), "source": "nvidia-smi -lms 20"}


# --------------------------------------------------------------------------- CPU reference leg
def cpu_reference_run(cloud, levels, nthreads):
    """One run of the reference CPU implementation; returns (sum N_l, sum t_l seconds, kind, sizes)."""
    from oracle import refhem
    if refhem.available():
        r = refhem.run(*cloud.arrays(), levels=levels, seed=7, nthreads=nthreads)
        kind = "reference"
    else:
        from oracle import porthem
        r = porthem.run(*cloud.arrays(), levels=levels, seed=7, nthreads=nthreads)
        kind = "port"
    sizes = [r.size(l) for l in range(levels + 1)]
    secs = sum(r.seconds(l) for l in range(levels))
    tot = sum(sizes[:levels])
    r.close()
    return tot, secs, kind, sizes


def run_reference_arm(a, rank, world):
    if rank != 0:
        return
    from gaussiansplattingregistration_b200.synth import synth
    cores = os.cpu_count() or 1
    n_s = min(a.n, a.cpu_sample if (a.steps + a.warmup) <= 20 else max(50_000, a.cpu_sample // 2))
    cloud = synth(n_s, seed=1, density=a.density, aspect=a.aspect)
    kind = "reference"
    for _ in range(a.warmup):
        cpu_reference_run(cloud, a.levels, cores)
    tot = secs = 0.0
    t0 = time.perf_counter()
    for _ in range(a.steps):
        t, s, kind, sizes = cpu_reference_run(cloud, a.levels, cores)
        tot += t
        secs += s
    wall = time.perf_counter() - t0
    value = tot / secs
    sample = f"{n_s}-Gaussian sample of the {a.n}-Gaussian workload (same generator/density/params/levels), per-level createClusterLevel time"
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
           "warmup": a.warmup, "ms_per_step": secs / a.steps * 1e3, "higher_is_better": True,
           "scaling": "strong",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(a, world, n=n_s, sample=n_s != a.n),
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample, "n": n_s},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "wall_s": wall, "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def gui_path_baseline(n_s, density, levels=1):
    """what the reference's GUI worker pays around the algorithm — Python lists in
    (mixturelevel.cpp:14-28), CreateMixture, Python lists out (mixturelevel.cpp:30-70) — through the reference's own
    pybind module when oracle/_ref holds it, beside this repository's `mixture_bind` stand-in on the same lists."""
    from gaussiansplattingregistration_b200.synth import synth
    c = synth(n_s, seed=1, density=density)
    lists = [x.tolist() for x in c.arrays()]
    xyz, col, op, cov, sh = lists
    out = {"n": n_s, "levels": levels}
    try:
        sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
        import mixture_bind as ref_mb            # the unmodified reference module (oracle/Makefile target refpy)
        t0 = time.perf_counter()
        lvl = ref_mb.MixtureLevel.CreateMixtureLevel(xyz, col, op, cov, sh)
        t1 = time.perf_counter()
        res = ref_mb.MixtureCreator.CreateMixture(levels, 3.0, 3.0, 2.5, 1.0, lvl)
        t2 = time.perf_counter()
        for r in res:
            ref_mb.MixtureLevel.CreatePythonLists(r)
        t3 = time.perf_counter()
        out["reference"] = {"lists_in_s": t1 - t0, "create_mixture_s": t2 - t1, "lists_out_s": t3 - t2, "total_s": t3 - t0}
    except Exception as e:       # module not built on this box
        out["reference"] = {"unavailable": str(e)[:120]}
    finally:
        sys.path.pop(0)
    from gaussiansplattingregistration_b200 import mixture_bind as mb
    for _ in range(2):           # second pass = warm context
        t0 = time.perf_counter()
        lvl = mb.MixtureLevel.CreateMixtureLevel(xyz, col, op, cov, sh)
        t1 = time.perf_counter()
        res = mb.MixtureCreator.CreateMixture(levels, 3.0, 3.0, 2.5, 1.0, lvl)
        t2 = time.perf_counter()
        for r in res:
            mb.MixtureLevel.CreatePythonLists(r)
        t3 = time.perf_counter()
    out["b200"] = {"lists_in_s": t1 - t0, "create_mixture_s": t2 - t1, "lists_out_s": t3 - t2, "total_s": t3 - t0}
    return out


# --------------------------------------------------------------------------- own arm
def level_checksum(ctx, level):
    """64-bit order-sensitive checksum of a level's rows (geometry, weights, flags, SH): equal checksums at
    N = 1, 2, 4, 8 mean bit-identical pyramids in the same row order.  Computed outside the timed region."""
    import torch
    d = ctx.level_torch(level, with_internal=True)
    acc = torch.zeros((), dtype=torch.int64, device="cuda")
    for i, k in enumerate(("xyz", "colors", "opacity", "cov6", "sh", "weight")):
        t = d[k].contiguous().view(torch.int32).to(torch.int64).reshape(d[k].shape[0], -1)
        if t.numel() == 0:
            continue
        w = torch.arange(1, 2 * t.shape[0], 2, dtype=torch.int64, device="cuda")[:, None] * (i + 3)
        c = torch.arange(1, t.shape[1] + 1, dtype=torch.int64, device="cuda")[None, :] * 0x9E3779B1
        acc = acc + ((t + c) * w).sum()
    acc = acc + (d["is_parent"].to(torch.int64) * torch.arange(1, d["is_parent"].shape[0] + 1, dtype=torch.int64, device="cuda")).sum()
    return f"{int(acc.item()) & 0xFFFFFFFFFFFFFFFF:016x}"


def measure(ctx, dev, levels, steps, warmup, barrier, dist, world):
    """K timed builds of one cloud resident in HBM.  Returns whole-job ms (max over ranks), the per-level statistics and
    the algorithmic bytes / CUDA-event times of the two hot kernels summed over this rank's launches."""
    import torch
    for _ in range(max(warmup, 3)):
        ctx.build(*dev, cluster_level=levels)
    stats = [ctx.step_stats(l) for l in range(levels)]
    l0 = ctx.info("launches")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gate_ms = gate_bytes = mom_ms = mom_bytes = 0.0
    lvl_ms = [0.0] * levels
    barrier()
    e0.record()
    tables = []
    for _ in range(steps):
        ctx.build(*dev, cluster_level=levels)
        tables.append(ctx.step_stats_table(levels))     # the event times the library recorded for this build (one small array)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    from gaussiansplattingregistration_b200.hem import STAT_NAMES
    col = {k: i for i, k in enumerate(STAT_NAMES)}
    for tab in tables:
        for l in range(levels):
            s = {k: float(tab[l, i]) for k, i in col.items()}
            lvl_ms[l] += s["ms_total"]
            gate_ms += s["ms_k_gate"]
            mom_ms += s["ms_k_moments"]
            # algorithmic bytes of THIS RANK's launch (S3 = 60N + 60Np + 12P + 4N ; S4 = 12P + 236 N_claimed
            # + 236 Np); a step sharded over G ranks gives each rank 1/G of the candidates and of the claimed children
            G = max(1, s["shards"])
            gate_bytes += 64 * s["n_in"] / G + 60 * s["n_parents_local"] + 12 * s["n_pairs_local"]
            mom_bytes += 12 * s["n_pairs_local"] + ROW_BYTES * (s["n_in"] - s["n_orphans"]) / G + ROW_BYTES * s["n_parents_local"]
    launches = ctx.info("launches") - l0
    t = torch.tensor([ms, gate_ms, mom_ms], dtype=torch.float64, device="cuda")
    b = torch.tensor([gate_bytes, mom_bytes], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)       # a sharded kernel is as slow as its slowest rank
        dist.all_reduce(b, op=dist.ReduceOp.SUM)       # bytes moved by all ranks
    ms, gate_ms, mom_ms = (float(x) for x in t.tolist())
    gate_bytes, mom_bytes = (float(x) for x in b.tolist())
    return dict(ms=ms, stats=stats, lvl_ms=lvl_ms, gate_ms=gate_ms, gate_bytes=gate_bytes, mom_ms=mom_ms, mom_bytes=mom_bytes,
                launches=int(launches), sum_n=sum(s["n_in"] for s in stats))


def run_b200_arm(a, rank, local_rank, world):
    import torch
    from gaussiansplattingregistration_b200.hem import HemContext
    from gaussiansplattingregistration_b200.synth import synth

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    cloud = synth(a.n, seed=1, density=a.density, aspect=a.aspect)       # the same cloud on every rank (strong scaling)
    dev = [torch.from_numpy(x).cuda() for x in cloud.arrays()]
    ctx = HemContext(local_rank, seed=7)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    ctx.set_option("timing", 2)        # level totals + the two hot kernels (what this script reports)
    if world > 1:
        ctx.init_distributed()

    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    M = measure(ctx, dev, a.levels, a.steps, a.warmup, barrier, dist, world)
    clocks = sampler.stop() if rank == 0 else None
    ms, stats, lvl_ms, sum_n = M["ms"], M["stats"], M["lvl_ms"], M["sum_n"]
    value = sum_n * a.steps / (ms / 1e3)                # one cloud, sharded over the ranks
    checksums = [level_checksum(ctx, l) for l in range(1, a.levels + 1)] if rank == 0 else None
    sizes = [ctx.level_size(l) for l in range(a.levels + 1)]

    # ---- end to end through the public API with pinned host buffers ----------------------------
    if a.skip_e2e:
        if rank == 0:
            print(json.dumps({'profile_only': True, 'value': value, 'ms_per_step': ms / a.steps}))
        return
    host_in = [torch.from_numpy(x).pin_memory() for x in cloud.arrays()]
    host_np = [h.numpy() for h in host_in]
    F = cloud.sh_floats
    # one pinned destination per returned level (what CreateMixture hands back), bound once: each level
    # is copied out on the library's copy stream while the next one is being built
    # Under torchrun the pyramid is returned to ONE caller: rank 0 binds the destinations, the other ranks are
    # workers (each uploads 1/N of the host arrays over its own PCIe link, the slices are all-gathered over NVLink).
    outb = []
    for l in range(1, a.levels + 1 if rank == 0 else 1):
        cap = int(sizes[l] * 1.02) + 1024
        bufs = [torch.empty((cap, 3), dtype=torch.float32).pin_memory(), torch.empty((cap, 3), dtype=torch.float32).pin_memory(),
                torch.empty(cap, dtype=torch.float32).pin_memory(), torch.empty((cap, 6), dtype=torch.float32).pin_memory(),
                torch.empty((cap, F), dtype=torch.float32).pin_memory()]
        outb.append(bufs)
        ctx.bind_output(l, *[t_.numpy() for t_ in bufs])

    def e2e_step():
        ctx.build(*host_np, cluster_level=a.levels)        # H2D of the cloud + D2H of every level inside

    for _ in range(2):
        e2e_step()
    k_e2e = max(3, min(a.steps, 5))
    h0, d0 = ctx.info("bytes_h2d"), ctx.info("bytes_d2h")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(k_e2e):
        e2e_step()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    e2e_value = sum_n * k_e2e / (e2e_ms / 1e3)
    io = torch.tensor([ctx.info("bytes_h2d") - h0, ctx.info("bytes_d2h") - d0], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(io, op=dist.ReduceOp.SUM)          # bytes over all ranks' PCIe links
    h2d, d2h = float(io[0].item()) / k_e2e, float(io[1].item()) / k_e2e
    # the bound host buffers now hold the last pyramid: spot-check them against a plain readback
    if rank == 0:
        chk = ctx.level(1)
        assert np.array_equal(outb[0][0].numpy()[:sizes[1]], chk["xyz"]) and np.array_equal(outb[0][4].numpy()[:sizes[1]], chk["sh"])
    ctx.clear_outputs()
    del outb, host_in, host_np

    # ---- the other BASELINE workload that fits one GPU, device-resident value only ------------------------
    extra = {}
    if not a.no_extra:
        other = "1m" if a.workload != "1m" else "3m"
        n2, lv2, asp2, cfg2 = WORKLOADS[other]
        del dev
        torch.cuda.empty_cache()
        c2 = synth(n2, seed=1, density=a.density, aspect=asp2)
        dev2 = [torch.from_numpy(x).cuda() for x in c2.arrays()]
        k2 = max(3, min(a.steps, 10))
        M2 = measure(ctx, dev2, lv2, k2, 3, barrier, dist, world)
        if rank == 0:
            extra[other] = {"config": cfg2, "n": n2, "levels": lv2, "value": M2["sum_n"] * k2 / (M2["ms"] / 1e3), "unit": UNIT,
                            "ms_per_step": M2["ms"] / k2, "steps": k2, "gpu_launches": M2["launches"],
                            "level_ms": [x / k2 for x in M2["lvl_ms"]], "level_sizes": [s_["n_in"] for s_ in M2["stats"]],
                            "k_gate_ms_per_step": M2["gate_ms"] / k2, "k_moments_ms_per_step": M2["mom_ms"] / k2,
                            "checksums": [level_checksum(ctx, l) for l in range(1, lv2 + 1)]}
        del dev2

    if rank != 0:
        return
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    # DRAM traffic of the dominant kernel: the committed `ncu --set full` capture of this round's code (level-0 launch of
    # this very workload on one GPU); bench.py itself never runs under a profiler
    traffic, traffic_note = None, None
    prof = os.path.join(ROOT, "profiles", f"r02_ncu_full_hot_kernels_{a.workload}.csv")
    if os.path.exists(prof) and a.n == WORKLOADS[a.workload][0]:
        import csv
        rows = {r[0]: r for r in csv.reader(open(prof)) if r}
        try:
            unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            traffic = sum(float(rows[m][2]) * unit[rows[m][1]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
            s0 = stats[0]
            traffic_note = (f"ncu --set full, level-0 k_gate launch on one GPU ({os.path.basename(prof)}); algorithmic bytes "
                            f"of that launch = {(64 * s0['n_in'] + 60 * s0['n_parents'] + 12 * s0['n_pairs']) / 1e6:.1f} MB")
        except (KeyError, ValueError, IndexError):
            traffic = None
    # aggregate over ranks: bytes of all ranks / slowest rank's kernel time, against `world` x the per-GPU peak
    gate_gbs = M["gate_bytes"] / (M["gate_ms"] * 1e-3) / 1e9 if M["gate_ms"] > 0 else 0.0
    mom_gbs = M["mom_bytes"] / (M["mom_ms"] * 1e-3) / 1e9 if M["mom_ms"] > 0 else 0.0
    peak_all = peak * world
    # whole pyramid by SURVEY.md §8d's B_level
    blevel = sum(ROW_BYTES * s_["n_in"] + ROW_BYTES * s_["n_out"] + 64 * s_["n_in"] + 12 * s_["n_pairs"] + 9 * s_["n_in"] for s_ in stats)
    lvl = [{"level": l, "n_in": stats[l]["n_in"], "n_parents": stats[l]["n_parents"], "n_pairs": stats[l]["n_pairs"],
            "n_out": stats[l]["n_out"], "shards": stats[l]["shards"], "ms": lvl_ms[l] / a.steps,
            "gaussians_per_s": stats[l]["n_in"] / (lvl_ms[l] / a.steps * 1e-3) if lvl_ms[l] > 0 else None}
           for l in range(a.levels)]
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
           "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic", "config": workload_config(a, world), "clocks": clocks,
           "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                   "ms_per_step": e2e_ms / k_e2e, "steps": k_e2e},
           "gpu_launches": M["launches"],
           "roofline": {"kernel": "hem::k_gate (gates + likelihood)", "bound": "hbm", "achieved": gate_gbs, "peak": peak_all,
                        "unit": "GB/s", "frac": gate_gbs / peak_all, "traffic": traffic, "traffic_note": traffic_note, "peak_source": peak_src + (f" x {world} GPUs" if world > 1 else ""),
                        "bytes_per_launch": M["gate_bytes"] / (a.steps * a.levels * world), "ms_per_launch": M["gate_ms"] / (a.steps * a.levels),
                        "note": "ALU/L2-bound by construction (SURVEY.md §8d): ~220 distance tests and ~63 in-radius candidates per parent against 60 B of compulsory traffic per child"},
           "kernels": {"k_gate": {"gbs": gate_gbs, "frac": gate_gbs / peak_all, "ms_per_step": M["gate_ms"] / a.steps},
                       "k_moments": {"gbs": mom_gbs, "frac": mom_gbs / peak_all, "ms_per_step": M["mom_ms"] / a.steps},
                       "pyramid_B_level": {"gbs": blevel * a.steps / (ms * 1e-3) / 1e9, "frac": blevel * a.steps / (ms * 1e-3) / 1e9 / peak_all,
                                           "bytes_per_step": blevel}},
           "levels": lvl, "level_sizes": sizes, "checksums": checksums, "workloads": extra}
    if world == 1 and not a.no_cpu_baseline:
        from gaussiansplattingregistration_b200.synth import synth as _s
        cores = os.cpu_count() or 1
        n_s = min(a.n, a.cpu_sample)
        sample_cloud = _s(n_s, seed=1, density=a.density, aspect=a.aspect)
        tot, secs, kind, _ = cpu_reference_run(sample_cloud, a.levels, cores)
        out["cpu_baseline"] = {"value": tot / secs, "unit": UNIT, "cores": cores, "kind": kind, "n": n_s,
                               "sample": f"{n_s}-Gaussian sample of the {a.n}-Gaussian workload (same generator/density/params/levels), "
                                         f"one run, per-level createClusterLevel time {secs:.2f} s; not the same size as the GPU arm"}
        if not a.no_extra:
            try:
                out["gui_path"] = gui_path_baseline(100_000, a.density)
            except Exception as e:      # never let the optional baseline take the bench line down
                out["gui_path"] = {"error": str(e)[:200]}
    print(json.dumps(out), flush=True)


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if a.impl == "reference":
        run_reference_arm(a, rank, world)
        return
    try:
        run_b200_arm(a, rank, local_rank, world)
    finally:
        if world > 1:
            import torch.distributed as dist
            if dist.is_initialized():
                dist.destroy_process_group()


if __name__ == "__main__":
    main()
