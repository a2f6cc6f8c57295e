#!/usr/bin/env python
"""bench.py -- particle-timesteps/s of the DSMC particle pipeline on N B200s (one process per GPU).

Default workload = BASELINE.json configs[2] (C3), the config the metric is quoted on: 1-D Couette Ar VHS, equal weight.  One "step" is
the reference's per-timestep pipeline in the order of simulations/1D/couette_benchmarking.jl:58-85:
    ntc_equal_weight! (all cells) -> convect_particles! (diffuse walls) -> [slab exchange, N > 1] -> sort_particles! ->
    compute_props_sorted!
on a synthetic population sampled on the device, 1.25e8 particles per GPU (1e9 on 8 GPUs; weak scaling, slab partition).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config c3|c4|c2|c5]
                    [--scaling same-dx|published-dx|same-L] [--particles-per-gpu P] [--band w] [--no-others]

--scaling (C3 only) says HOW the vs-SPARTA case is scaled up to 1.25e8 particles per GPU:
    same-dx       the small benchmark's cell (dx = 1e-5 m, ppc = 1000, BENCHMARKS.md:10-35), more cells, longer domain (default)
    published-dx  the large benchmark's cell (dx = 2.5e-7 m, ppc = 250, BENCHMARKS.md:93-99: L = 5e-4 m over 2000 cells), more cells:
                  sigma_v dt = 2.6 cells, the sort's band is 15 cells wide and the rare outliers go through its hybrid path
    same-L        L = 5e-4 m fixed over ALL ranks, ppc = 250, nx = particles / 250 ("same L, more cells", 4e6 cells for
                  1e9 particles): a particle crosses hundreds of cells per step, every sort is a full re-sort (general path), 5e5
                  particles cross each slab face per step (full exchange)
--config selects another BASELINE.json config as THE workload of the line (same JSON contract): c4 = 1-D Couette variable weight +
octree merging (configs[3], slab-partitioned like c3), c2 = 0-D BKW variable weight + octree merging (configs[1], replicas),
c5 = 0-D Fokker-Planck ensemble (configs[4], replicas).  The default c3 line also carries a short run of each of them (and of the
other two scalings) under "other_configs", so that one driver run times every config; --no-others skips that.

Prints ONE JSON line (see the contract in DESIGN.md "Measurement").
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "merzbild.jl_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

AR = 66.3e-27
K_B = 1.380649e-23
NDENS, DT, T_WALL, V_WALL = 5e22, 2.59e-9, 300.0, 500.0
L_REF = 5e-4          # the reference's Couette domain (BENCHMARKS.md:93-99, couette_benchmarking.jl:95)
BYTES_SORT = 116      # sort -- stable scatter (key 4 + record 56 read, record 56 write)
BYTES_STEP = 220      # full Couette step (convect + sort + collide + props)
BYTES_FP = 56         # fp_linear (w, v read 32 B, v written 24 B)
METRIC = "particle-timesteps/s, 1D Couette Ar VHS"

SCALINGS = {
    "same-dx": {"dx": 1e-5, "ppc": 1000, "band": 2, "xmode": 0,
                "name": "couette_ar_vhs_equal_weight (BENCHMARKS.md vs-SPARTA case: ppc=1000, dx=1e-5 m, dt=2.59e-9 s, n=5e22), scaling same-dx (more cells of "
                        "the same size, longer domain)"},
    "published-dx": {"dx": 2.5e-7, "ppc": 250, "band": 15, "xmode": 0,
                     "name": "couette_ar_vhs_equal_weight (BENCHMARKS.md:93-99 large case: ppc=250, dx=2.5e-7 m = 5e-4 m / 2000, dt=2.59e-9 s, n=5e22), scaling "
                             "published-dx (more cells of the published size: sigma_v dt = 2.6 cells)"},
    "same-L": {"dx": None, "ppc": 250, "band": 0, "xmode": 1,
               "name": "couette_ar_vhs_equal_weight (BENCHMARKS.md:93-99 large case: L=5e-4 m, ppc=250, dt=2.59e-9 s, n=5e22), scaling same-L (SURVEY.md 8(d): "
                       "same L over all ranks, more cells: dx = L / nx)"},
}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_cpu_port(nx, ppc, steps, warm, threads, L=None):
    """The C++ restatement of the reference's multithreaded Couette loop (oracle/couette_cpu.cpp), timed on the host cores."""
    from oracle import oracle

    oracle.build()
    exe = os.path.join(ROOT, "oracle", "_build", "couette_cpu")
    cmd = [exe, str(nx), str(ppc), str(steps), str(warm), str(threads)] + (["1234", repr(L)] if L is not None else [])
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, check=True).stdout
    except (subprocess.CalledProcessError, OSError):
        oracle.build(force=True)  # e.g. built on a different CPU
        out = subprocess.run(cmd, capture_output=True, text=True, check=True).stdout
    return json.loads(out.strip().splitlines()[-1])


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for ts, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                clk, mx = float(f[1]), float(f[2])
            except ValueError:
                continue
            smax = mx
            if t0 - 0.05 <= ts <= t1 + 0.15:
                sm.append(clk)
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        if not sm:
            sm = [float(x[1].split(",")[1]) for x in self.rows[-3:] if len(x[1].split(",")) > 2] or [0.0]
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line of the contract goes to the real stdout; everything libraries print (NCCL's version banner, torchrun
    notices) was redirected to stderr at start-up."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def hbm_peak():
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    if "hbm_gbs" in peaks:
        return float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback 6650 GB/s (B200_PROFILING.md)"


def committed_traffic(kernel, particles):
    """DRAM traffic of a kernel from the committed ncu --set full capture (profiles/traffic.json); only quoted for the particle count it
    was captured at."""
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[kernel]
        if abs(tr["particles"] - particles) <= 0.01 * particles:
            return tr["dram_bytes_read"] + tr["dram_bytes_write"]
    except (OSError, ValueError, KeyError):
        pass
    return None


def step_traffic_frac(workload_key, particles, ms_per_step, peak):
    """Measured DRAM bytes of ALL kernels of one step (ncu --set full, profiles/traffic.json "step_<workload>") / step time / peak: the
    bandwidth the step really draws, next to the byte-model number step_frac."""
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["step_" + workload_key]
        if abs(tr["particles"] - particles) <= 0.01 * particles:
            b = tr["dram_bytes_read"] + tr["dram_bytes_write"]
            return {"dram_bytes_per_step": b, "achieved_GBps": b / (ms_per_step * 1e-3) / 1e9, "frac": b / (ms_per_step * 1e-3) / 1e9 / peak,
                    "source": "profiles/traffic.json (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum over the step's kernels)"}
    except (OSError, ValueError, KeyError):
        pass
    return None


class Env:
    """ranks, device, barrier / reductions over ranks (torch.distributed is plumbing only)"""

    def __init__(self):
        import torch

        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: libmerzbild_b200 has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        self.numa = bind_to_gpu_numa_node(self.local_rank)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist

            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.dist = dist

    def barrier(self, ctx=None):
        if ctx is not None:
            ctx.sync()
        if self.dist:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, x, op):
        if not self.dist:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=getattr(self.dist.ReduceOp, op))
        return float(t.item())

    def max(self, x):
        return self.reduce(x, "MAX")

    def sum(self, x):
        return self.reduce(x, "SUM")

    def new_context(self, mb, seed=1234):
        ctx = mb.Context(self.local_rank, seed + self.rank)  # rank r uses seed + r, as couette_multithreaded.jl:17 does per chunk
        if self.world > 1:
            uid = [mb.comm_unique_id() if self.rank == 0 else None]
            self.dist.broadcast_object_list(uid, src=0)
            mb.comm_init(ctx, uid[0], self.rank, self.world)
        return ctx

    def pinned(self, n, k=7):
        torch = self.torch
        try:
            pin = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(k)]
        except RuntimeError:  # the box refuses to pin that much: pageable host buffers (the e2e leg is then slower, still valid)
            pin = [torch.empty(n, dtype=torch.float64) for _ in range(k)]
        return pin, [p.numpy() for p in pin]

    def close(self):
        if self.dist:
            self.dist.destroy_process_group()


def bind_to_gpu_numa_node(local_rank):
    """Pin this rank (and therefore the pinned host buffers it allocates next: first touch) to the CPUs of the NUMA node its GPU hangs
    off, so that the e2e leg's host->device copies of 8 ranks do not all cross one socket's memory controller."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read().strip())
        if node < 0:
            return None
        cpus = []
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        allowed = set(os.sched_getaffinity(0))
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return {"node": node, "cpus": len(cpus)}
    except Exception:  # no NVML / sysfs / permission: leave the affinity alone
        return None
    return None


def roofline_block(kernel, bytes_per_particle, particles, ms, peak, peak_src, traffic=None):
    achieved = bytes_per_particle * particles / (ms * 1e-3) / 1e9 if ms and ms > 0 else None
    return {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
            "traffic": traffic, "algorithmic_bytes": bytes_per_particle * particles, "peak_source": peak_src,
            "algorithmic_bytes_per_particle": bytes_per_particle}


# ------------------------------------------------------------------------------------------------------------------- C3
def c3_shape(scaling, particles_per_gpu, world):
    S = SCALINGS[scaling]
    ppc = S["ppc"]
    nx_local = max(int(round(particles_per_gpu / ppc)), 1)
    nx_global = nx_local * world
    dx = S["dx"] if S["dx"] is not None else L_REF / nx_global
    return S, ppc, nx_local, nx_global, dx


def run_c3(env, args, scaling, steps, warmup, e2e_steps, band=None):
    import numpy as np

    import merzbild_b200 as mb

    rank, world = env.rank, env.world
    S, ppc, nx_local, nx_global, dx = c3_shape(scaling, args.particles_per_gpu, world)
    band = S["band"] if band is None else band
    ctx = env.new_context(mb)
    ctx.set_band_halfwidth(band)
    mb.exchange_set_mode(ctx, S["xmode"])
    G = mb.Grid1DUniform(nx_global * dx, nx_global)
    slab = G.slab(rank, world)
    nx = slab.n_cells
    n = nx * ppc
    Fnum = dx * NDENS / ppc
    cap = int(n * (1.05 if scaling != "same-L" else 1.15)) + 4096
    pv = mb.ParticleVector(cap, ctx)
    pia = mb.ParticleIndexerArray(nx, 1, ctx)
    it = mb.make_interaction(AR, AR, 4.11e-10, 0.81, 273.0)
    cf = mb.CollisionFactors(nx, mb.estimate_sigma_g_w_max(it, AR, AR, T_WALL, T_WALL, Fnum), ctx)
    walls = mb.MaxwellWalls1D(T_WALL, T_WALL, -V_WALL, V_WALL, 1.0, 1.0)
    props = mb.PhysProps(nx, 1, ctx=ctx)
    # initial condition sampled on the device (sample_particles_equal_weight!(rng, grid, ..., ppc, T, Fnum), grid_uniform1D.jl:117-152)
    mb.sample_particles_equal_weight(mb.PhiloxRng(0, 0), slab, pv, pia, 1, AR, ppc, T_WALL, Fnum)
    ctx.sync()
    tstep = [0]
    dt = DT * getattr(args, "dt_mult", 1.0)

    def step():
        tstep[0] += 1
        r = mb.PhiloxRng(tstep[0], 0)
        mb.ntc_equal_weight(r, cf, None, it, pv, pia, (1, nx), 1, dt, slab.dx)
        mb.convect_particles(r, slab, walls, pv, pia, 1, AR, dt)
        if world > 1:
            mb.exchange_slab(ctx, slab, pv, pia, 1)
        mb.sort_particles(None, slab, pv, pia, 1)
        mb.compute_props_sorted([pv], pia, [AR], props)

    host = pin = None
    if e2e_steps > 0:  # the e2e leg uploads the initial state from pinned host memory every step
        pin, host = env.pinned(n)
        pv.download_soa(1, n, host)
        indexer = pia.indexer.copy()
        n_total = np.array([n], dtype=np.int64)
        contiguous = np.array([1], dtype=np.uint8)
    for _ in range(warmup):
        step()
    env.barrier(ctx)
    # ---- device-resident timing (value)
    ctx.prof_enable(True)
    l0 = ctx.kernel_launches
    env.barrier(ctx)
    w0 = time.time()
    ctx.timer_start()
    for _ in range(steps):
        step()
    ms = ctx.timer_stop()
    env.barrier(ctx)
    w1 = time.time()
    launches = ctx.kernel_launches - l0
    sections = ctx.prof_read()
    ctx.prof_enable(False)
    sort_path = ctx.sort_last_path
    extras = ctx.sort_last_extras
    n_now = int(pia.n_total[0])
    ms_max = env.max(ms)
    particles_all = env.sum(float(n_now))
    res = {"value": particles_all * steps / (ms_max * 1e-3), "ms_per_step": ms_max / steps, "particles": int(particles_all), "cells": nx * world, "ppc": ppc,
           "launches": launches, "sections": {k: v[0] / steps for k, v in sections.items()}, "sort_path": "band" if sort_path == 1 else "general",
           "band_halfwidth": band, "sort_extras_last_step": extras, "n_rank": n_now, "wall": (w0, w1), "dx": dx,
           "sigma_v_dt_over_dx": math.sqrt(K_B * T_WALL / AR) * DT / dx,
           "workload": "%s, %.3g particles/GPU, slab partition" % (S["name"], n), "scaling_rule": scaling}
    sc_ms, sc_n = sections.get("sort.scatter", (0.0, 0))
    if sort_path == 1 and sc_n:
        if ctx.sort_last_pass_b == 1:
            res["dominant"] = ("k_band_tile (sort_particles! pass B: a CTA per tile of cells, TMA bulk loads / stores, permutation in shared memory)", BYTES_SORT,
                               sc_ms / sc_n)
        else:
            res["dominant"] = ("k_band_scatter (sort_particles! pass B: stable scatter by source cell)", BYTES_SORT, sc_ms / sc_n)
    else:
        g_ms, g_n = sections.get("sort.general", (0.0, 0))
        res["dominant"] = ("general sort path (classify, scan, index scatter, per-cell index sort, gather by cell)", BYTES_SORT, g_ms / max(g_n, 1))
    # ---- end to end through the C ABI with HOST buffers: H2D of the particle state + pia, the step, D2H of the props
    if e2e_steps > 0:
        props_host = {k: np.empty(s) for k, s in (("np", (1, nx)), ("n", (1, nx)), ("v", (1, nx, 3)), ("T", (1, nx)))}

        def upload():
            pv.upload_soa(1, n, host)
            pia.upload(indexer, n_total, contiguous)

        upload()
        step()
        env.barrier(ctx)
        ctx.timer_start()
        for _ in range(e2e_steps):
            upload()
            step()
            mb._ck(mb.lib().mb_props_download(props.h, None, mb._p(props_host["np"]), mb._p(props_host["n"]), mb._p(props_host["v"]),
                                              mb._p(props_host["T"]), None))
        ms_e2e = env.max(ctx.timer_stop())
        env.barrier(ctx)
        res["e2e"] = {"value": particles_all * e2e_steps / (ms_e2e * 1e-3), "unit": "particle-timesteps/s", "h2d_bytes_per_step": 56 * n + indexer.nbytes + 8,
                      "d2h_bytes_per_step": 48 * nx, "steps": e2e_steps, "ms_per_step": ms_e2e / e2e_steps}
        res["mean_T_K"] = float(props_host["T"].mean())
        res["wall"] = (w0, time.time())
    else:
        res["mean_T_K"] = float(props.download()["T"].mean())
    for o in (pv, pia):
        o.close()
    ctx.close()
    del host, pin
    return res


# ------------------------------------------------------------------------------------------------------------------- C4
def run_c4(env, args, steps, warmup, e2e_steps):
    """BASELINE.json configs[3]: 1-D Couette, variable weight, octree merging (couette_multithreaded_varweight_octree.jl:205-206,
    couette_varweight_octree.jl:86-135): 500 particles sampled per cell, merged to 100 at t = 0; per step ntc! (splits) ->
    merge_octree_N2_based! where n_local > 130 -> convect_particles! -> [slab exchange: reads the non-contiguous layout through the
    squash map] -> sort_particles! (squashes first) -> compute_props_sorted!.  Slab-partitioned like C3; ~1.1e8 live particles/GPU."""
    import numpy as np

    import merzbild_b200 as mb

    rank, world = env.rank, env.world
    dx, ppc_s, thr, tgt = 1e-5, 500, 130, 100
    nx_local = max(int(round(args.particles_per_gpu / 109.0)), 64)  # 1.15e6 cells / GPU -> 1.25e8 live particles after the merge (1e9 on 8 GPUs)
    nx_global = nx_local * world
    ctx = env.new_context(mb)
    mb.exchange_set_mode(ctx, 0)  # edge exchange: sigma_v dt = 0.065 cells, the leavers sit in the w = 2 cells next to a slab face
    G = mb.Grid1DUniform(nx_global * dx, nx_global, wall_offset=1e-6)  # L ~ 10 m: the default offset dx * 1e-12 is below ulp(L)
    slab = G.slab(rank, world)
    nx = slab.n_cells
    Fnum = dx * NDENS / ppc_s
    pv, pia = mb.ParticleVector(int(nx * ppc_s * 1.01) + 1024, ctx), mb.ParticleIndexerArray(nx, 1, ctx)
    oc = mb.OctreeN2Merge(mb.OctreeN2Merge.OctreeBinMidSplit, mb.OctreeN2Merge.OctreeInitBinMinMaxVel, max_Nbins=6000)
    it = mb.make_interaction(AR, AR, 4.11e-10, 0.81, 273.0)
    mb.sample_particles_equal_weight(mb.PhiloxRng(0), slab, pv, pia, 1, AR, float(NDENS), T_WALL, Fnum)
    n_sampled = int(pia.n_total[0])
    ctx.sync()
    ctx.timer_start()
    mb.merge_octree_N2_based(mb.PhiloxRng(0), oc, pv, pia, (1, nx), 1, tgt, slab, threshold=thr)
    ms_merge0 = ctx.timer_stop()
    mb.squash_pia(pv, pia, 1)
    walls = mb.MaxwellWalls1D(T_WALL, T_WALL, -V_WALL, V_WALL, 1.0, 1.0)
    cf = mb.CollisionFactors(nx, mb.estimate_sigma_g_w_max(it, AR, AR, T_WALL, T_WALL, dx * NDENS / tgt), ctx)
    props = mb.PhysProps(nx, 1, ctx=ctx)
    tstep = [0]

    def step():
        tstep[0] += 1
        r = mb.PhiloxRng(tstep[0], 0)
        mb.ntc(r, cf, None, it, pv, pia, (1, nx), 1, DT, slab.dx)
        mb.merge_octree_N2_based(r, oc, pv, pia, (1, nx), 1, tgt, slab, threshold=thr)
        mb.convect_particles(r, slab, walls, pv, pia, 1, AR, DT)
        if world > 1:
            mb.exchange_slab(ctx, slab, pv, pia, 1)
        mb.sort_particles(None, slab, pv, pia, 1)
        mb.compute_props_sorted([pv], pia, [AR], props)

    def sum_w():
        return env.sum(float(props.download()["n"].sum()))

    for _ in range(max(warmup, 3)):
        step()
    env.barrier(ctx)
    w_before = sum_w()
    ctx.prof_enable(True)
    l0 = ctx.kernel_launches
    env.barrier(ctx)
    w0 = time.time()
    ctx.timer_start()
    for _ in range(steps):
        step()
    ms = ctx.timer_stop()
    env.barrier(ctx)
    w1 = time.time()
    launches = ctx.kernel_launches - l0
    sections = ctx.prof_read()
    ctx.prof_enable(False)
    w_after = sum_w()
    n_now = int(pia.n_total[0])
    ms_max = env.max(ms)
    particles_all = env.sum(float(n_now))
    g_ms, g_n = sections.get("sort.general", (0.0, 0))
    res = {"value": particles_all * steps / (ms_max * 1e-3), "ms_per_step": ms_max / steps, "particles": int(particles_all), "cells": nx * world,
           "ppc": particles_all / (nx * world), "launches": launches, "sections": {k: v[0] / steps for k, v in sections.items()}, "n_rank": n_now,
           "wall": (w0, w1), "dominant": ("general sort path with the squash folded in (sort_particles! after merge_octree_N2_based!)", BYTES_SORT, g_ms / max(g_n, 1)),
           "workload": "couette_ar_vhs_variable_weight_octree (couette_multithreaded_varweight_octree.jl:205-206: 500 sampled per cell merged to 100, threshold 130, "
                       "OctreeBinMidSplit, MinMaxVel, max_Nbins 6000; dx=1e-5 m, dt=2.59e-9 s), %d cells/GPU, slab partition" % nx,
           "initial_merge": {"particles_sampled_per_gpu": n_sampled, "ms": ms_merge0},
           "weight_conservation": {"sum_w_before": w_before, "sum_w_after": w_after, "rel_drift": abs(w_after - w_before) / w_before,
                                   "bound_4eps_per_step": 4 * 2.220446049250313e-16 * steps,
                                   "ok": abs(w_after - w_before) / w_before <= 4 * 2.220446049250313e-16 * steps}}
    res["mean_T_K"] = float(props.download()["T"].mean())
    if e2e_steps > 0:
        res["e2e"] = generic_e2e(env, mb, ctx, pv, pia, props, step, n_now, nx, e2e_steps, particles_all)
    for o in (pv, pia):
        o.close()
    ctx.close()
    return res


def generic_e2e(env, mb, ctx, pv, pia, props, step, n, nx, e2e_steps, particles_all, props_keys=("np", "n", "v", "T")):
    """the step through the C ABI with HOST buffers: the particle state + pia of the current step are uploaded from pinned memory,
    the step runs, the props come back -- every step"""
    import numpy as np

    pin, host = env.pinned(n)
    pv.download_soa(1, n, host)
    indexer, n_total, contiguous = pia.download()
    props_host = {k: np.empty(s) for k, s in (("np", (1, nx)), ("n", (1, nx)), ("v", (1, nx, 3)), ("T", (1, nx)))}

    def upload():
        pv.upload_soa(1, n, host)
        pia.upload(indexer, n_total, contiguous)

    upload()
    step()
    env.barrier(ctx)
    ctx.timer_start()
    for _ in range(e2e_steps):
        upload()
        step()
        mb._ck(mb.lib().mb_props_download(props.h, None, mb._p(props_host["np"]), mb._p(props_host["n"]), mb._p(props_host["v"]),
                                          mb._p(props_host["T"]), None))
    ms_e2e = env.max(ctx.timer_stop())
    env.barrier(ctx)
    return {"value": particles_all * e2e_steps / (ms_e2e * 1e-3), "unit": "particle-timesteps/s", "h2d_bytes_per_step": 56 * n + indexer.nbytes + 8,
            "d2h_bytes_per_step": 48 * nx, "steps": e2e_steps, "ms_per_step": ms_e2e / e2e_steps}


# ------------------------------------------------------------------------------------------------------------------- C2
def run_c2(env, args, steps, warmup, e2e_steps):
    """BASELINE.json configs[1]: 0-D BKW variable-weight relaxation with octree N:2 merging (simulations/0D/BKW/bkw_varweight_octree.jl:90-104,
    test/test_bkw_varweight_octree.jl:43-47): an ensemble of independent cells (replicas; every rank its own), each the nv = 40 velocity
    grid sample (~33.5k particles) merged to 8000 at t = 0; per step ntc! -> merge_octree_N2_based! where n_local > 10000 -> re-sort by
    cell id (folds the squash_pia!) -> compute_props_with_total_moments!.  1e8 LIVE particles per GPU: 12500 cells x 8000."""
    import numpy as np

    import merzbild_b200 as mb

    ctx = mb.Context(env.local_rank, 1234 + env.rank)
    itm = mb.make_interaction(AR, AR, 4.11e-10, 1.0, 273.0)  # data/pseudo_maxwell.toml
    T0, n_dens, nv, tgt, thr = 273.0, 1e23, 40, 8000, 10000
    probe_pv, probe_pia = mb.ParticleVector(nv ** 3, ctx), mb.ParticleIndexerArray(1, 1, ctx)
    n_s = mb.sample_on_grid(mb.PhiloxRng(0), "bkw", probe_pv, probe_pia, 1, 1, nv, AR, T0, n_dens)
    probe_pv.close()
    probe_pia.close()
    ncell = max(int(round(args.particles_per_gpu * 0.8 / tgt)), 1)  # 1.25e8 * 0.8 = 1e8 live
    n0 = ncell * n_s
    pv, pia = mb.ParticleVector(n0, ctx), mb.ParticleIndexerArray(ncell, 1, ctx)
    oc = mb.OctreeN2Merge(mb.OctreeN2Merge.OctreeBinMidSplit, mb.OctreeN2Merge.OctreeInitBinMinMaxVel, max_Nbins=6000)
    mb.sample_on_grid(mb.PhiloxRng(0), "bkw", pv, pia, (1, ncell), 1, nv, AR, T0, n_dens)
    ctx.sync()
    ctx.timer_start()
    mb.merge_octree_N2_based(mb.PhiloxRng(0), oc, pv, pia, (1, ncell), 1, tgt, threshold=thr)
    ms_merge0 = ctx.timer_stop()
    mb.sort_particles(None, pv, pia, 1)
    props = mb.PhysProps(ncell, 1, (4, 6, 8, 10), Tref=T0, ctx=ctx)
    cf = mb.CollisionFactors(ncell, mb.estimate_sigma_g_w_max(itm, AR, AR, T0, T0, n_dens / tgt), ctx)
    tref = 1.0 / (n_dens * math.pi * 4.11e-10 ** 2) / math.sqrt(2 * K_B * T0 / AR)
    dt = 0.025 * tref
    tstep = [0]
    census = []

    def step(count=False):
        tstep[0] += 1
        r = mb.PhiloxRng(tstep[0])
        mb.ntc(r, cf, None, itm, pv, pia, (1, ncell), 1, dt, 1.0)
        if count:  # which cells merge this step (one small download: the merge's algorithmic bytes are 56 (N + N_target) per merging cell)
            nl = pia.indexer[0, :, 0]
            m = nl > thr
            census.append((int(m.sum()), int(nl[m].sum())))
        mb.merge_octree_N2_based(r, oc, pv, pia, (1, ncell), 1, tgt, threshold=thr)
        mb.sort_particles(None, pv, pia, 1)
        mb.compute_props_with_total_moments([pv], pia, [AR], props)

    for _ in range(max(warmup, 3)):
        step()
    env.barrier(ctx)
    ctx.prof_enable(True)
    l0 = ctx.kernel_launches
    env.barrier(ctx)
    w0 = time.time()
    ctx.timer_start()
    for _ in range(steps):
        step(count=True)
    ms = ctx.timer_stop()
    env.barrier(ctx)
    w1 = time.time()
    launches = ctx.kernel_launches - l0
    sections = ctx.prof_read()
    ctx.prof_enable(False)
    n_now = int(pia.n_total[0])
    ms_max = env.max(ms)
    particles_all = env.sum(float(n_now))
    d = props.download()
    sec = {k: v[0] / steps for k, v in sections.items()}
    merged_cells = sum(c for c, _ in census)
    merge_bytes = sum(56.0 * (np_in + tgt * c) for c, np_in in census)
    if sec.get("merge", 0.0) >= sec.get("sort.general", 0.0) and merged_cells > 0:
        dom = ("k_merge (merge_octree_N2_based!, CTA per merging cell)", merge_bytes / max(n_now, 1) / steps, sec["merge"])
    else:
        dom = ("general sort path by cell id with the squash folded in", BYTES_SORT, sec.get("sort.general", 0.0))
    res = {"value": particles_all * steps / (ms_max * 1e-3), "ms_per_step": ms_max / steps, "particles": int(particles_all), "cells": ncell * env.world,
           "ppc": n_now / ncell, "launches": launches, "sections": sec, "n_rank": n_now, "wall": (w0, w1), "dominant": dom,
           "workload": "bkw_0d_variable_weight_octree (test_bkw_varweight_octree.jl:43-47: nv=40 grid sample %d/cell merged to 8000, threshold 10000, pseudo-Maxwell "
                       "molecules, dt = 0.025 t_ref), %d independent cells per GPU = %.3g live particles (replicas only: 0-D cases do not shard)" % (n_s, ncell, n_now),
           "initial_merge": {"particles_sampled_per_gpu": n0, "cells": ncell, "ms": ms_merge0},
           "merges_in_timed_steps": {"cells": merged_cells, "per_step": [c for c, _ in census],
                                     "note": "the census (one download of the cell populations per step, inside the timed region) names the merging cells"},
           "mean_T_K": float(d["T"].mean()), "mean_M4": float(d["moments"][0, :, 0].mean())}
    if e2e_steps > 0:
        res["e2e"] = generic_e2e(env, mb, ctx, pv, pia, props, step, n_now, ncell, e2e_steps, particles_all)
    for o in (pv, pia):
        o.close()
    ctx.close()
    return res


# ------------------------------------------------------------------------------------------------------------------- C5
def run_c5(env, args, steps, warmup, e2e_steps):
    """BASELINE.json configs[4]: 0-D Fokker-Planck ensemble (test/test_collision_fp.jl:7, test_1D_couette_fp.jl:3-11): 1e6 independent cells
    x 100 particles per GPU, Ar, T = 300 K; per step fp_linear! on every cell (replicas only)."""
    import merzbild_b200 as mb

    ctx = mb.Context(env.local_rank, 1234 + env.rank)
    ppc, V = 100, 1e-5
    nc = max(int(round(args.particles_per_gpu * 0.8 / ppc)), 1)  # 1e8 particles
    n = nc * ppc
    it = mb.make_interaction(AR, AR, 4.11e-10, 0.81, 273.0)
    grid = mb.Grid1DUniform(nc * 1e-5, nc)
    pv, pia = mb.ParticleVector(n, ctx), mb.ParticleIndexerArray(nc, 1, ctx)
    mb.sample_particles_equal_weight(mb.PhiloxRng(0), grid, pv, pia, 1, AR, ppc, T_WALL, V * NDENS / ppc)
    props = mb.PhysProps(nc, 1, ctx=ctx)
    tstep = [0]

    def step():
        tstep[0] += 1
        mb.fp_linear(mb.PhiloxRng(tstep[0]), None, it, AR, pv, pia, (1, nc), 1, DT, V)

    for _ in range(max(warmup, 3)):
        step()
    env.barrier(ctx)
    l0 = ctx.kernel_launches
    env.barrier(ctx)
    w0 = time.time()
    ctx.timer_start()
    for _ in range(steps):
        step()
    ms = ctx.timer_stop()
    env.barrier(ctx)
    w1 = time.time()
    launches = ctx.kernel_launches - l0
    ms_max = env.max(ms)
    particles_all = env.sum(float(n))
    mb.compute_props_sorted([pv], pia, [AR], props)
    res = {"value": particles_all * steps / (ms_max * 1e-3), "ms_per_step": ms_max / steps, "particles": int(particles_all), "cells": nc * env.world, "ppc": ppc,
           "launches": launches, "sections": {"fp": ms_max / steps}, "n_rank": n, "wall": (w0, w1),
           "dominant": ("k_fp_linear_reg (fp_linear!, warp per cell, the cell in registers)", BYTES_FP, ms_max / steps),
           "workload": "fokker_planck_0d_ensemble (test_collision_fp.jl / test_1D_couette_fp.jl shape: Ar, 300 K, dt=2.59e-9 s), %d independent cells x %d "
                       "particles per GPU (replicas only)" % (nc, ppc),
           "mean_T_K": float(props.download()["T"].mean())}
    if e2e_steps > 0:
        def step_props():
            step()
            mb.compute_props_sorted([pv], pia, [AR], props)
        res["e2e"] = generic_e2e(env, mb, ctx, pv, pia, props, step_props, n, nc, e2e_steps, particles_all)
    for o in (pv, pia):
        o.close()
    ctx.close()
    return res


# ------------------------------------------------------------------------------------------------------------------- CPU legs
def cpu_c3(scaling, steps, warmup, threads, calibrate=True, world=1, particles_per_gpu=1.25e8):
    S, ppc, nx_local, nx_global, dx = c3_shape(scaling, particles_per_gpu, world)
    nx_cpu = 8000 if ppc == 1000 else 32000  # 8e6 particles: a bounded sample of the same workload (ppc, dx, dt, physics identical)
    L = nx_cpu * dx
    if calibrate:
        cal = run_cpu_port(nx_cpu, ppc, 5, 2, threads, L)  # calibration: size the sample to ~15 s of CPU work
        steps = int(min(max(15.0 * cal["particle_steps_per_s"] / (nx_cpu * ppc), 20), 2000))
        warmup = 3
    r = run_cpu_port(nx_cpu, ppc, steps, warmup, threads, L)
    sample = ("C++ restatement of the reference's multithreaded Couette loop (couette_multithreaded.jl:97-173; Julia is not installed; the same operator code "
              "replays the reference's golden runs to round-off, tests/test_oracle_reference_bitlevel.py): %d cells x %d ppc = %.1e particles at dx = %.3g m "
              "(scaling %s), %d steps, %d OpenMP threads, %.1f s; collide+convect+sort %.2fs, exchange %.2fs, resort+props %.2fs" %
              (nx_cpu, ppc, nx_cpu * ppc, dx, scaling, steps, threads, r["seconds"], r["collide_convect_sort_s"], r["exchange_s"], r["resort_props_s"]))
    return r["particle_steps_per_s"], r["seconds"], steps, nx_cpu * ppc, sample, threads


def cpu_ops(config, budget_s=12.0):
    """C4 / C2 / C5 on the host: the oracle's operators (the C++ restatement, one thread) driven through oracle.py on a bounded sample."""
    import numpy as np

    from oracle import oracle

    oracle.lib()
    rng = np.random.default_rng(1)
    if config == "c5":
        nc, ppc = 2000, 100
        n = nc * ppc
        rows = np.zeros((n, 7))
        rows[:, 0] = 1e-5 * NDENS / ppc
        rows[:, 1:4] = rng.normal(0, math.sqrt(K_B * T_WALL / AR), (n, 3))
        opv, opia = oracle.OPV(n), oracle.OPIA(nc, 1)
        opv.particles[:n] = rows
        opv.nbuffer = 0
        for c in range(nc):
            opia.indexer[0, c] = (ppc, c * ppc + 1, (c + 1) * ppc, ppc, 0, -1, 0)
        opia.n_total[0] = n
        oit = oracle.interaction("Ar", "Ar")
        t0, k = time.time(), 0
        while time.time() - t0 < budget_s:
            k += 1
            oracle.fp_linear(oracle.Rng.philox(1234, k), oit, AR, opv, opia, 1, nc, 1, DT, 1e-5)
        el = time.time() - t0
        return n * k / el, el, k, n, "oracle fp_linear! (collision_fp.jl:24-125 restated), %d cells x %d, %d steps, 1 thread, %.1f s" % (nc, ppc, k, el), 1
    if config == "c4":
        nx, ppc_s, thr, tgt, dx = 400, 500, 130, 100, 1e-5
        L = nx * dx
        opv, opia = oracle.OPV(int(nx * ppc_s * 1.05)), oracle.OPIA(nx, 1)
        oracle.sample_equal_weight_cells(oracle.Rng.philox(1234, 0), opv, opia, 1, nx, 1, -1, AR, T_WALL, dx * NDENS / ppc_s, grid=(L, nx), ndens=NDENS)
        oc = oracle.Octree(max_Nbins=6000)
        oracle.merge_octree_N2(oracle.Rng.philox(1234, 0), oc, opv, opia, 1, nx, 1, tgt, threshold=thr, grid=(L, nx), squash_after_each=True)
        oit = oracle.interaction("Ar", "Ar")
        ocf = oracle.CF(nx, oracle.estimate_sigma_g_w_max(oit, AR, AR, T_WALL, T_WALL, dx * NDENS / tgt))
        t0, k, ps = time.time(), 0, 0
        while time.time() - t0 < budget_s:
            k += 1
            r = oracle.Rng.philox(1234, k)
            oracle.ntc(r, ocf, oit, opv, opia, 1, nx, 1, DT, dx)
            oracle.merge_octree_N2(r, oc, opv, opia, 1, nx, 1, tgt, threshold=thr, grid=(L, nx), squash_after_each=True)
            oracle.convect_particles(r, (L, nx), (T_WALL, T_WALL, -V_WALL, V_WALL, 1.0, 1.0), opv, opia, 1, [AR], DT)
            oracle.sort_particles(opv, opia, 1, grid=(L, nx))
            oracle.compute_props_sorted([opv], opia, [AR])
            ps += int(opia.n_total[0])
        el = time.time() - t0
        return ps / el, el, k, int(opia.n_total[0]), ("oracle variable-weight Couette loop (couette_varweight_octree.jl:86-135 restated: ntc!, per-cell octree merge + "
                                                      "squash_pia!, convect, sort, props), %d cells, ~%d particles, %d steps, 1 thread, %.1f s" %
                                                      (nx, int(opia.n_total[0]), k, el)), 1
    # c2
    ncell, nv, tgt, thr, T0, n_dens = 2, 40, 8000, 10000, 273.0, 1e23
    opv, opia = oracle.OPV(ncell * 40000), oracle.OPIA(ncell, 1)
    oracle.sample_on_grid_cells(oracle.Rng.philox(1234, 0), "bkw", opv, opia, 1, ncell, 1, nv, AR, T0, n_dens)
    oc = oracle.Octree(max_Nbins=6000)
    oracle.merge_octree_N2(oracle.Rng.philox(1234, 0), oc, opv, opia, 1, ncell, 1, tgt, threshold=thr, squash_after_each=True)
    oit = oracle.make_interaction(AR, AR, 4.11e-10, 1.0, 273.0)  # data/pseudo_maxwell.toml
    ocf = oracle.CF(ncell, oracle.estimate_sigma_g_w_max(oit, AR, AR, T0, T0, n_dens / tgt))
    tref = 1.0 / (n_dens * math.pi * 4.11e-10 ** 2) / math.sqrt(2 * K_B * T0 / AR)
    t0, k, ps = time.time(), 0, 0
    while time.time() - t0 < budget_s:
        k += 1
        r = oracle.Rng.philox(1234, k)
        oracle.ntc(r, ocf, oit, opv, opia, 1, ncell, 1, 0.025 * tref, 1.0)
        oracle.merge_octree_N2(r, oc, opv, opia, 1, ncell, 1, tgt, threshold=thr, squash_after_each=True)
        oracle.compute_props([opv], opia, [AR], (4, 6, 8, 10), T0, with_moments=True)
        ps += int(opia.n_total[0])
    el = time.time() - t0
    return ps / el, el, k, int(opia.n_total[0]), ("oracle 0-D BKW variable-weight loop (bkw_varweight_octree.jl:90-99 restated: ntc!, octree merge above 10000, squash_pia!, "
                                                  "props with total moments), %d cells of 8000-10000, %d steps, 1 thread, %.1f s" % (ncell, k, el)), 1


def cpu_leg(config, scaling, args, calibrate):
    if config == "c3":
        return cpu_c3(scaling, args.steps, args.warmup, host_threads(), calibrate, max(args.gpus, 1), args.particles_per_gpu)
    return cpu_ops(config)


def reference_arm(args, rank):
    """The reference's CPU implementation of the path on the host cores: the oracle port (Julia is not installed anywhere), on a bounded
    sample of the arm's workload.  The line's config states what RAN."""
    if rank != 0:
        return
    v, seconds, steps, n_part, sample, threads = cpu_leg(args.config, args.scaling, args, calibrate=False)
    world = max(args.gpus, 1)
    if args.config == "c3":
        S, ppc, nx_local, nx_global, dx = c3_shape(args.scaling, args.particles_per_gpu, world)
        workload = "%s; the b200 arm runs %.3g particles/GPU" % (S["name"], nx_local * ppc)
    else:
        workload = {"c4": "couette_ar_vhs_variable_weight_octree", "c2": "bkw_0d_variable_weight_octree", "c5": "fokker_planck_0d_ensemble"}[args.config]
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "particle-timesteps/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 * seconds / max(steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload, "bench_config": args.config, "particles": n_part,
                   "note": "this arm ran the bounded sample described under cpu_baseline.sample (%.3g particles on the host), not the b200 arm's full size: a "
                           "larger working set only slows a CPU down" % n_part},
        "cpu_baseline": {"value": v, "unit": "particle-timesteps/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "particle-timesteps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def summary(res, peak, peak_src):
    """compact record of a secondary workload (other_configs)"""
    k, bpp, ms = res["dominant"]
    out = {"workload": res["workload"], "value": res["value"], "unit": "particle-timesteps/s", "ms_per_step": res["ms_per_step"], "particles": res["particles"],
           "cells": res["cells"], "gpu_launches": res["launches"], "sections_ms_per_step": res["sections"], "mean_T_K": res.get("mean_T_K"),
           "roofline": roofline_block(k, bpp, res["n_rank"], ms, peak, peak_src,
                                      committed_traffic(k.split(" ")[0], res["n_rank"]) if k.startswith("k_band_") else None)}
    for key in ("sort_path", "band_halfwidth", "sort_extras_last_step", "sigma_v_dt_over_dx", "initial_merge", "weight_conservation", "merges_in_timed_steps",
                "mean_M4", "scaling_rule"):
        if key in res:
            out[key] = res[key]
    return out


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c3", choices=["c3", "c4", "c2", "c5"])
    ap.add_argument("--scaling", default="same-dx", choices=sorted(SCALINGS))
    ap.add_argument("--particles-per-gpu", type=float, default=1.25e8)
    ap.add_argument("--e2e-steps", type=int, default=3, help="0 (experiments only): skip the end-to-end leg, the line then carries e2e = null")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-others", action="store_true", help="do not append the short runs of the other configs / scalings to the default line")
    ap.add_argument("--dt-mult", type=float, default=1.0, help="experiment knob: multiplies the time step (0: nothing moves, the sort is a pure segmented copy)")
    ap.add_argument("--band", type=int, default=None, help="band half-width of the sort fast path (0: general path only); default: the scaling's")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        reference_arm(args, int(os.environ.get("RANK", "0")))
        return

    env = Env()
    rank, world = env.rank, env.world
    peak, peak_src = hbm_peak()
    sampler = ClockSampler(env.local_rank)
    sampler.start()
    time.sleep(0.3)
    if args.config == "c3":
        res = run_c3(env, args, args.scaling, args.steps, args.warmup, args.e2e_steps, args.band)
    else:
        res = {"c4": run_c4, "c2": run_c2, "c5": run_c5}[args.config](env, args, args.steps, args.warmup, max(args.e2e_steps, 1))
    clocks = sampler.stop(*res["wall"])

    others = {}
    if args.config == "c3" and args.scaling == "same-dx" and not args.no_others and args.band is None:
        k = max(min(args.steps, 10), 3)
        for sc in ("published-dx", "same-L"):
            others["c3/" + sc] = summary(run_c3(env, args, sc, k, 3, 0), peak, peak_src)
        others["c4"] = summary(run_c4(env, args, k, 3, 0), peak, peak_src)
        others["c2"] = summary(run_c2(env, args, max(args.steps, 20), 3, 0), peak, peak_src)
        others["c5"] = summary(run_c5(env, args, k, 3, 0), peak, peak_src)

    if rank != 0:
        env.close()
        return

    kname, bpp, kms = res["dominant"]
    traffic = committed_traffic(kname.split(" ")[0], res["n_rank"]) if kname.startswith("k_band_") else None
    roofline = roofline_block(kname, bpp, res["n_rank"], kms, peak, peak_src, traffic)
    roofline["sections_ms_per_step"] = res["sections"]
    if args.config == "c3":
        step_gbs = BYTES_STEP * res["particles"] / (res["ms_per_step"] * 1e-3) / 1e9 / world
        roofline["step_achieved_GBps"] = step_gbs
        roofline["step_frac"] = step_gbs / peak  # byte MODEL of SURVEY.md 8(d) (220 B per particle-step) / time / peak ...
        roofline["step_dram"] = step_traffic_frac(args.scaling, res["n_rank"], res["ms_per_step"], peak)  # ... and what the kernels really moved
    config = {"workload": res["workload"], "bench_config": args.config, "particles": res["particles"], "cells": res["cells"], "ppc": res["ppc"],
              "l2": "working set %.1f GB per GPU >> 126 MB L2 (no flush needed)" % (56 * res["n_rank"] / 1e9), "mean_T_K": res.get("mean_T_K"),
              "numa_binding": env.numa}
    for key in ("sort_path", "band_halfwidth", "sort_extras_last_step", "sigma_v_dt_over_dx", "dx", "initial_merge", "weight_conservation",
                "merges_in_timed_steps", "mean_M4", "scaling_rule"):
        if key in res:
            config[key] = res[key]
    if args.config == "c3":
        config["step"] = "ntc_equal_weight+convect+%ssort+props_sorted" % ("exchange+" if world > 1 else "")
    line = {
        "metric": METRIC if args.config in ("c3", "c4") else "particle-timesteps/s, " + {"c2": "0D BKW variable weight + octree merging", "c5": "0D Fokker-Planck ensemble"}[args.config],
        "value": res["value"], "unit": "particle-timesteps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": config, "roofline": roofline, "e2e": res.get("e2e"), "gpu_launches": res["launches"], "clocks": clocks,
    }
    if others:
        line["other_configs"] = others
    if world == 1 and not args.no_cpu_baseline:
        v, seconds, steps, n_part, sample, threads = cpu_leg(args.config, args.scaling, args, calibrate=True)
        line["cpu_baseline"] = {"value": v, "unit": "particle-timesteps/s", "cores": threads, "kind": "port", "sample": sample}
    emit(line)
    env.close()


if __name__ == "__main__":
    main()
