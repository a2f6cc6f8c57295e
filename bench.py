#!/usr/bin/env python
"""bench.py -- particle-timesteps/s of the 1-D Couette Ar VHS step (BASELINE.json configs[2]) on N B200s.

One "step" = the reference's per-timestep pipeline in the order of simulations/1D/couette_benchmarking.jl:58-85:
    ntc_equal_weight! (all cells) -> convect_particles! (diffuse walls) -> [slab exchange, N > 1] -> sort_particles! ->
    compute_props_sorted!
on a synthetic population of the vs-SPARTA case's shape (Ar, vhs.toml, T_wall 300 K, v_wall +-500 m/s, n = 5e22 m^-3,
dt = 2.59e-9 s, ppc = 1000, dx = 1e-5 m), scaled to 1.25e8 particles per GPU (1e9 on 8 GPUs; weak scaling, slab partition).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--particles-per-gpu P]

Prints ONE JSON line (see the contract in DESIGN.md "Measurement").
"""
import argparse
import json
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
DX, PPC, NDENS, DT, T_WALL, V_WALL = 1e-5, 1000, 5e22, 2.59e-9, 300.0, 500.0
BYTES_SCATTER = 116  # sort -- stable scatter (key 4 + record 56 read, record 56 write)
BYTES_STEP = 220     # full Couette step (convect + sort + collide + props)


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_cpu_port(nx, ppc, steps, warm, threads):
    """The C++ restatement of the reference's multithreaded Couette loop (oracle/couette_cpu.cpp), timed on the host cores."""
    from oracle import oracle

    oracle.build()
    exe = os.path.join(ROOT, "oracle", "_build", "couette_cpu")
    try:
        out = subprocess.run([exe, str(nx), str(ppc), str(steps), str(warm), str(threads)], capture_output=True, text=True, check=True).stdout
    except (subprocess.CalledProcessError, OSError):
        oracle.build(force=True)  # e.g. built on a different CPU
        out = subprocess.run([exe, str(nx), str(ppc), str(steps), str(warm), str(threads)], capture_output=True, text=True, check=True).stdout
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


def workload_name(particles_per_gpu):
    n = max(int(round(particles_per_gpu / PPC)), 1) * PPC
    return ("couette_ar_vhs_equal_weight (BENCHMARKS.md vs-SPARTA case: ppc=1000, dx=1e-5 m, dt=2.59e-9 s, n=5e22) scaled to "
            "%.3g particles/GPU, slab partition" % n)


def reference_arm(args, rank):
    if rank != 0:
        return
    threads = host_threads()
    nx = 16000  # 1.6e7 particles: a bounded sample of the same workload (ppc, dx, dt, physics identical)
    r = run_cpu_port(nx, PPC, args.steps, args.warmup, threads)
    v = r["particle_steps_per_s"]
    line = {
        "impl": "reference", "metric": "particle-timesteps/s, 1D Couette Ar VHS", "value": v, "unit": "particle-timesteps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * r["seconds"] / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.particles_per_gpu), "particles": int(round(args.particles_per_gpu / PPC)) * PPC * max(args.gpus, 1),
                   "cells": int(round(args.particles_per_gpu / PPC)) * max(args.gpus, 1), "ppc": PPC,
                   "step": "ntc_equal_weight+convect+exchange(chunks)+sort+props_sorted",
                   "sample": "each step runs on a bounded sample of this workload: %d cells x %d ppc = %.1e particles, all host threads" % (nx, PPC, nx * PPC)},
        "cpu_baseline": {"value": v, "unit": "particle-timesteps/s", "cores": threads, "kind": "port",
                         "sample": "C++ restatement of the reference's multithreaded Couette loop (Julia is not installed; the same operator code replays the "
                                   "reference's golden runs to round-off, tests/test_oracle_reference_bitlevel.py): %d cells x %d ppc, %d steps; "
                                   "collide+convect+sort %.2fs, exchange %.2fs, resort+props %.2fs" %
                                   (nx, PPC, args.steps, r["collide_convect_sort_s"], r["exchange_s"], r["resort_props_s"])},
        "e2e": {"value": v, "unit": "particle-timesteps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


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
    ap.add_argument("--particles-per-gpu", type=float, default=1.25e8)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--band", type=int, default=2, help="band half-width of the sort fast path (0: general path only)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        reference_arm(args, rank)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import merzbild_b200 as mb

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libmerzbild_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = mb.Context(local_rank, 1234 + rank)  # rank r uses seed + r, as couette_multithreaded.jl:17 does per chunk
    ctx.set_band_halfwidth(args.band)

    # ---- workload: slab `rank` of a global grid of world * nx_local cells
    nx_local = max(int(round(args.particles_per_gpu / PPC)), 1)
    nx_global = nx_local * world
    G = mb.Grid1DUniform(nx_global * DX, nx_global)
    slab = G.slab(rank, world)
    nx = slab.n_cells
    n = nx * PPC
    Fnum = DX * NDENS / PPC
    cap = int(n * 1.05) + 4096
    indexer = np.zeros((1, nx, 7), dtype=np.int64)
    n_total = np.array([n], dtype=np.int64)
    contiguous = np.array([1], dtype=np.uint8)
    try:
        pin = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(7)]
    except RuntimeError:  # the box refuses to pin 7 GB per rank: pageable host buffers (the e2e leg is then slower, still valid)
        pin = [torch.empty(n, dtype=torch.float64) for _ in range(7)]
    host = [p.numpy() for p in pin]

    pv = mb.ParticleVector(cap, ctx)
    pia = mb.ParticleIndexerArray(nx, 1, ctx)
    it = mb.make_interaction(AR, AR, 4.11e-10, 0.81, 273.0)
    sgwm0 = mb.estimate_sigma_g_w_max(it, AR, AR, T_WALL, T_WALL, Fnum)
    cf = mb.CollisionFactors(nx, sgwm0, ctx)
    walls = mb.MaxwellWalls1D(T_WALL, T_WALL, -V_WALL, V_WALL, 1.0, 1.0)
    props = mb.PhysProps(nx, 1, ctx=ctx)
    props_host = {k: np.empty(s) for k, s in (("np", (1, nx)), ("n", (1, nx)), ("v", (1, nx, 3)), ("T", (1, nx)))}
    if world > 1:
        uid = [mb.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        mb.comm_init(ctx, uid[0], rank, world)

    # initial condition sampled on the device (sample_particles_equal_weight!(rng, grid, ..., ppc, T, Fnum), grid_uniform1D.jl:117-152),
    # then copied once to pinned host memory: the end-to-end leg uploads it from there every step
    mb.sample_particles_equal_weight(mb.PhiloxRng(0, 0), slab, pv, pia, 1, AR, PPC, T_WALL, Fnum)
    ctx.sync()
    pv.download_soa(1, n, host)
    indexer[:] = pia.indexer

    def upload():
        pv.upload_soa(1, n, host)
        pia.upload(indexer, n_total, contiguous)

    tstep = [0]

    def step():
        tstep[0] += 1
        r = mb.PhiloxRng(tstep[0], 0)
        mb.ntc_equal_weight(r, cf, None, it, pv, pia, (1, nx), 1, DT, slab.dx)
        mb.convect_particles(r, slab, walls, pv, pia, 1, AR, DT)
        if world > 1:
            mb.exchange_slab(ctx, slab, pv, pia, 1)
        mb.sort_particles(None, slab, pv, pia, 1)
        mb.compute_props_sorted([pv], pia, [AR], props)

    def barrier():
        ctx.sync()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    upload()
    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)

    # ---- device-resident timing (value)
    ctx.prof_enable(True)
    l0 = ctx.kernel_launches
    barrier()
    w0 = time.time()
    ctx.timer_start()
    for _ in range(args.steps):
        step()
    ms = ctx.timer_stop()
    barrier()
    w1 = time.time()
    launches = ctx.kernel_launches - l0
    sections = ctx.prof_read()
    ctx.prof_enable(False)
    sort_path = ctx.sort_last_path
    n_now = int(pia.n_total[0])
    ms_max = max_over_ranks(ms)
    particles_all = sum_over_ranks(float(n_now))
    value = particles_all * args.steps / (ms_max * 1e-3)

    # ---- end to end through the C ABI with HOST buffers: H2D of the particle state + pia, the step, D2H of the props
    e2e_steps = max(args.e2e_steps, 1)
    upload()
    step()
    barrier()
    ctx.timer_start()
    for _ in range(e2e_steps):
        upload()
        step()
        mb._ck(mb.lib().mb_props_download(props.h, None, mb._p(props_host["np"]), mb._p(props_host["n"]), mb._p(props_host["v"]),
                                          mb._p(props_host["T"]), None))
    ms_e2e = max_over_ranks(ctx.timer_stop())
    barrier()
    w2 = time.time()
    clocks = sampler.stop(w0, w2)
    e2e_value = particles_all * e2e_steps / (ms_e2e * 1e-3)
    h2d = 56 * n + indexer.nbytes + 8
    d2h = 48 * nx
    T_mean = float(props_host["T"].mean())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    sc_ms, sc_n = sections.get("sort.scatter", (0.0, 0))
    if sort_path != 1 or sc_n == 0:
        sc_ms, sc_n = sections.get("sort.general", (0.0, 0))
    achieved = BYTES_SCATTER * n_now / (sc_ms / max(sc_n, 1) * 1e-3) / 1e9 if sc_ms > 0 else None
    # DRAM traffic of the dominant kernel from the committed ncu --set full capture (profiles/traffic.json); only quoted for the
    # particle count it was captured at
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["k_band_scatter"]
        if sort_path == 1 and abs(tr["particles"] - n_now) <= 0.01 * n_now:
            traffic = tr["dram_bytes_read"] + tr["dram_bytes_write"]
    except (OSError, ValueError, KeyError):
        pass
    roofline = {"bound": "hbm", "kernel": "k_band_scatter (sort_particles! pass B: stable scatter by source cell)" if sort_path == 1 else "general sort path",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                "algorithmic_bytes": BYTES_SCATTER * n_now,
                "peak_source": peak_src, "algorithmic_bytes_per_particle": BYTES_SCATTER,
                "step_achieved_GBps": BYTES_STEP * particles_all * args.steps / (ms_max * 1e-3) / 1e9 / world,
                "step_frac": BYTES_STEP * particles_all * args.steps / (ms_max * 1e-3) / 1e9 / world / peak,
                "sections_ms_per_step": {k: v[0] / args.steps for k, v in sections.items()}}
    line = {
        "metric": "particle-timesteps/s, 1D Couette Ar VHS", "value": value, "unit": "particle-timesteps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload_name(args.particles_per_gpu), "particles": int(particles_all), "cells": nx * world, "ppc": PPC,
                   "step": "ntc_equal_weight+convect+%ssort+props_sorted" % ("exchange+" if world > 1 else ""), "sort_path": "band" if sort_path == 1 else "general",
                   "l2": "working set %.1f GB per GPU >> 126 MB L2 (no flush needed)" % (56 * n / 1e9), "mean_T_K": T_mean},
        "roofline": roofline,
        "e2e": {"value": e2e_value, "unit": "particle-timesteps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                "ms_per_step": ms_e2e / e2e_steps},
        "gpu_launches": launches,
        "clocks": clocks,
    }
    if world == 1 and not args.no_cpu_baseline:
        threads = host_threads()
        nx_cpu = 8000
        cal = run_cpu_port(nx_cpu, PPC, 5, 2, threads)  # calibration: size the sample to ~15 s of CPU work
        k_cpu = int(min(max(15.0 * cal["particle_steps_per_s"] / (nx_cpu * PPC), 20), 2000))
        r = run_cpu_port(nx_cpu, PPC, k_cpu, 3, threads)
        line["cpu_baseline"] = {"value": r["particle_steps_per_s"], "unit": "particle-timesteps/s", "cores": threads, "kind": "port",
                                "sample": "C++ restatement of the reference's multithreaded Couette loop (couette_multithreaded.jl:97-173; the same operator code "
                                          "replays the reference's golden runs to round-off): %d cells x %d ppc "
                                          "(%.0e particles), %d steps, %d OpenMP threads, %.1f s" % (nx_cpu, PPC, nx_cpu * PPC, k_cpu, threads, r["seconds"])}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
