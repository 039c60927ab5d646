#!/usr/bin/env python
"""Benchmark of the Particle Life physics step (Physics.update) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C3]

Prints ONE JSON line (rank 0).  A "step" is one Physics.update() over the whole
particle set: cell-list rebuild + 3x3 force pass + friction/integrate/wrap.

  value     particle-steps/s with the state resident in HBM, CUDA events on the
            launching stream, max over ranks
  e2e       the same metric through the C ABI with host buffers: every step
            re-sends settings + matrix from host memory and pulls the full fp32
            render snapshot (xy, vxy as fp32, type as u8 = 17 B/particle) into pinned host memory
  roofline  dominant kernel (force/integrate) against measured HBM peak; the
            path is FP32-issue bound at the benchmark density, so `fp32` carries
            the binding fraction
  cpu_baseline  the CPU oracle (C restatement of the reference algorithm, no JVM
            exists here) on the host cores, bounded sample of the same workload

--impl reference times that CPU restatement with all host threads instead.
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
sys.path.insert(0, os.path.join(ROOT, "particle-life-app_b200"))

METRIC = "particle_steps_per_sec"
UNIT = "particle-steps/s"
DT = 0.02
ALGO_BYTES_STEP = 84       # fp32 SoA three-pass design, bytes per particle-step
ALGO_BYTES_FORCE = 36      # pass C: read 20 + write 16
FLOP_PER_PAIR = 17         # wrap on, default accelerator


def workload(name):
    from plife import synth
    c = dict(synth.CONFIGS[name])
    c["name"] = name
    return c


def scaled_workload(name, world):
    """Weak scaling: per-GPU particle count fixed, density fixed (rmax shrinks with sqrt(world)).
    world == 8 on C3 reproduces BASELINE config 4's 128M particles (nx = 2828 instead of 2800)."""
    import numpy as np
    c = workload(name)
    if world > 1:
        nx1 = int(np.floor(1.0 / c["rmax"]))
        nx = int(round(nx1 * world ** 0.5))
        rmax = 1.0 / nx
        while int(np.floor(1.0 / rmax)) != nx:
            rmax = float(np.nextafter(rmax, 0.0))
        c["rmax"] = rmax
        c["n_per_gpu"] = c["n"]
        c["n"] = c["n"] * world
    else:
        c["n_per_gpu"] = c["n"]
    return c


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_run(cfg, steps, warmup, threads, sample_n):
    """The oracle's threaded variant (B/Physics.java:116-134 phase structure) on a bounded sample:
    same density, matrix size, settings; fewer particles."""
    import numpy as np
    import oracle
    from plife import synth
    scale = (sample_n / cfg["n"]) ** 0.5
    rmax = cfg["rmax"] / scale  # same particles per cell
    pos, vel, types = synth.uniform_state(sample_n, cfg["m"], cfg["seed"])
    M = synth.random_matrix(cfg["m"], cfg["seed"])
    o = oracle.Oracle(rmax=rmax, matrix=M, wrap=cfg["wrap"], dt=DT, threads=threads)
    o.set_particles(pos, vel, types)
    for _ in range(warmup):
        o.update()
    t0 = time.perf_counter()
    for _ in range(steps):
        o.update()
    dt = time.perf_counter() - t0
    evals = o.pair_stats()[0]
    return dict(value=sample_n * steps / dt, ms_per_step=dt / steps * 1e3, pair_evals_per_s=evals * steps / dt,
                sample=f"{sample_n} particles, rmax={rmax:.6g} (same {sample_n / int(1 / rmax) ** 2:.1f} particles/cell as the workload), "
                       f"m={cfg['m']}, {steps} steps after {warmup} warm-up, {threads} threads",
                nx=int(1 / rmax))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cfg = scaled_workload(args.workload, max(1, args.gpus))
    threads = os.cpu_count() or 1
    sample_n = min(cfg["n"], 1_000_000)
    r = cpu_reference_run(cfg, max(1, args.steps), max(0, args.warmup), threads, sample_n)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{cfg['name']}: {cfg['n']} particles, {cfg['m']} types, rmax={cfg['rmax']}, wrap={cfg['wrap']} "
                               f"(CPU arm runs a bounded same-density sample)", "sample": r["sample"]},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": threads, "kind": "port", "sample": r["sample"],
                         "note": "C restatement of the reference algorithm (no JVM available); an optimistic proxy for the Java path"},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "pair_evals_per_sec": r["pair_evals_per_s"],
    }
    print(json.dumps(line), flush=True)
    return 0


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # one process per GPU: run on the CPUs next to this GPU, so that the pinned snapshot buffers (first touch) and the
        # launch thread live on its NUMA node; best effort, the numbers are valid without it
        try:
            import pynvml
            pynvml.nvmlInit()
            pr = torch.cuda.get_device_properties(local_rank)
            bus = f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"  # CUDA order may differ from NVML's
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode()))
        except Exception:  # noqa: BLE001
            pass
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import plife
    from plife import _native as N

    cfg = scaled_workload(args.workload, world)
    n, m = cfg["n"], cfg["m"]  # n = GLOBAL particle count
    precision = plife.F64 if args.precision == "f64" else plife.F32

    stream = torch.cuda.Stream()
    if world == 1:
        p = plife.NativePhysics(device=local_rank, precision=precision, stream=stream.cuda_stream)
        p.set_settings(cfg["rmax"], 0.85, 1.0, cfg["wrap"])
        p.random_matrix(m, cfg["seed"])
        p.init_uniform(n, cfg["seed"])
        slab = None

        def run_steps(k):
            p.step(DT, k)
    else:
        from plife.slab import DistExchange, SlabPhysics, grid_rows
        if precision != plife.F32:
            raise SystemExit("bench.py: the slab path is fp32")
        nx = grid_rows(cfg["rmax"])
        rho = n / nx ** 2
        share = n // world
        def make_slab(mode):
            with torch.cuda.stream(stream):
                sl = SlabPhysics(rank, world, cfg["rmax"], device=local_rank, capacity=share + share // 8 + 65536,
                                 halo_cap=int(nx * rho * 1.5) + 4096, mig_cap=max(65536, share // 64), wrap=cfg["wrap"],
                                 stream=stream.cuda_stream, exchange=mode)
                if mode == "peer":
                    sl.connect_dist()
            return sl

        mode = args.exchange
        try:
            slab = make_slab(mode)
        except RuntimeError as e:  # connect_dist raises on EVERY rank if CUDA IPC is unavailable on any: NCCL instead
            if mode != "peer":
                raise
            if rank == 0:
                print(f"bench.py: {e}; using NCCL send/recv", file=sys.stderr)
            mode = args.exchange = "nccl"
            slab = make_slab(mode)
        p = slab.native
        p.random_matrix(m, cfg["seed"])
        p.init_uniform(n, cfg["seed"])  # every rank scans the global stream and keeps its own rows
        ex = DistExchange(rank, world)

        def run_steps(k):
            slab.step(DT, ex, k)
    matrix = p.get_matrix()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident-state throughput ----
    with torch.cuda.stream(stream):
        run_steps(max(3, args.warmup))
        barrier()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        run_steps(args.steps)
        e1.record(stream)
        barrier()
        clocks = sampler.stop() if rank == 0 else None
        ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = n * args.steps / (ms * 1e-3)
    stats = p.step_stats()
    n_local = p.count

    # ---- per-kernel device time of the same steps (events between kernels) ----
    with torch.cuda.stream(stream):
        # slabs: every rank runs the same profiled steps (the exchange needs all of them); rank 0's times are reported,
        # and the "gather" bucket there also holds the halo pack / push / wait / unpack kernels
        p.set_profiling(True)
        run_steps(args.steps)
        kt = p.kernel_times()
        p.set_profiling(False)
    force_ms = kt["force"][0] / max(1, kt["force"][1])
    per_kernel = {k: v[0] / args.steps for k, v in kt.items()}

    # ---- end to end through the C ABI with host buffers ----
    ncap = n if world == 1 else n_local + n_local // 8 + 65536
    pin_pos = torch.empty((ncap, 2), dtype=torch.float32).pin_memory()
    pin_vel = torch.empty((ncap, 2), dtype=torch.float32).pin_memory()
    pin_typ = torch.empty((ncap,), dtype=torch.uint8).pin_memory()  # one byte per type: the app allows at most 256 types
    h2d = (matrix.nbytes + 32) * world
    d2h = n * 17  # xy + vxy as fp32, type as u8
    e2e_steps = max(3, min(args.steps, 10))

    pins = [(pin_pos, pin_vel, pin_typ),
            (torch.empty_like(pin_pos).pin_memory(), torch.empty_like(pin_vel).pin_memory(), torch.empty_like(pin_typ).pin_memory())]
    e2e_k = [0]

    def e2e_step():
        # settings + matrix go host -> device every step; the snapshot of step k copies out (pinned host memory,
        # second stream) while step k+1 computes; every snapshot is awaited before its buffer is reused
        p.set_settings(cfg["rmax"], 0.85, 1.0, cfg["wrap"])
        p.set_matrix(matrix)
        run_steps(1)
        p.snapshot_wait()  # the previous snapshot is complete (its buffer is the renderer's now)
        a, b, c = pins[e2e_k[0] & 1]
        e2e_k[0] += 1
        p.snapshot_async(a.data_ptr(), b.data_ptr(), c.data_ptr(), types_u8=True)

    with torch.cuda.stream(stream):
        for _ in range(3):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        p.snapshot_wait()
        barrier()
        e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = n * e2e_steps / float(t.item())

    secondary = None
    if world == 1 and args.workload == "C3" and precision == plife.F32:
        # BASELINE config 2 (1M particles, 8 types): the metric names both sizes; this one fits in L2
        c2 = workload("C2")
        q = plife.NativePhysics(device=local_rank, precision=precision, stream=stream.cuda_stream)
        q.set_settings(c2["rmax"], 0.85, 1.0, c2["wrap"])
        q.random_matrix(c2["m"], c2["seed"])
        q.init_uniform(c2["n"], c2["seed"])
        with torch.cuda.stream(stream):
            q.step(DT, 20)
            torch.cuda.synchronize()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(stream)
            q.step(DT, 200)
            a1.record(stream)
            torch.cuda.synchronize()
        ms2 = a0.elapsed_time(a1) / 200
        secondary = {"workload": f"C2: {c2['n']} particles, {c2['m']} types, rmax={c2['rmax']} (16 particles/cell), state fits in L2 (cache-resident)",
                     "value": c2["n"] / (ms2 * 1e-3), "unit": UNIT, "ms_per_step": ms2, "pair_evals_per_sec": q.step_stats()["pair_evals"] / (ms2 * 1e-3)}
        # the same state after it has evolved into clusters (what the app actually runs): most cells empty, some with
        # thousands of particles; candidates per particle grow by an order of magnitude, so particle-steps/s drops while
        # pair-evals/s does not
        with torch.cuda.stream(stream):
            q.step(DT, 2000)
            torch.cuda.synchronize()
            a0.record(stream)
            q.step(DT, 50)
            a1.record(stream)
            torch.cuda.synchronize()
        ms3 = a0.elapsed_time(a1) / 50
        pe = q.step_stats()["pair_evals"]
        secondary["evolved"] = {"workload": "the C2 state 2220 steps later (clustered)", "value": c2["n"] / (ms3 * 1e-3), "unit": UNIT,
                                "ms_per_step": ms3, "pair_evals_per_particle": pe / c2["n"], "pair_evals_per_sec": pe / (ms3 * 1e-3)}
        q.close()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    hbm_peak, peak_src = peaks()
    force_gbs = ALGO_BYTES_FORCE * n_local / (force_ms * 1e-3) / 1e9 if force_ms > 0 else None  # one launch = this rank's particles
    step_gbs = ALGO_BYTES_STEP * n / (ms / args.steps * 1e-3) / 1e9
    pair_rate = stats["pair_evals"] * world * args.steps / (ms * 1e-3)
    fp32_nominal_tf = 148 * 128 * 2 * (clocks["sm_max_mhz"] or 1965.0) * 1e6 / 1e12
    import ctypes
    meas = ctypes.c_double(0.0)
    fp32_peak_tf = meas.value if N.lib().plife_measure_fp32_peak(local_rank, ctypes.byref(meas)) == 0 and meas.value > 0 else fp32_nominal_tf
    fp32_peak_src = "measured FFMA loop on this GPU (plife_measure_fp32_peak)" if meas.value > 0 else "nominal"
    fp32_force_tf = FLOP_PER_PAIR * stats["pair_evals"] / (force_ms * 1e-3) / 1e12 if force_ms > 0 else None
    fp32_step_tf = FLOP_PER_PAIR * stats["pair_evals"] * world / (ms / args.steps * 1e-3) / 1e12

    cpu = None
    if world == 1:
        threads = os.cpu_count() or 1
        sample_n = min(n, 1_000_000)
        r = cpu_reference_run(cfg, 2, 1, threads, sample_n)
        # grow the sample toward ~10 s of CPU work
        extra = int(min(20, max(0, 10.0 / (r["ms_per_step"] * 1e-3) - 2)))
        if extra >= 2:
            r = cpu_reference_run(cfg, extra, 1, threads, sample_n)
        cpu = {"value": r["value"], "unit": UNIT, "cores": threads, "kind": "port", "sample": r["sample"],
               "pair_evals_per_sec": r["pair_evals_per_s"],
               "note": "C restatement of the reference algorithm (no JVM available); an optimistic proxy for the Java path"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if precision == plife.F32 else "f64", "data": "synthetic",
        "config": {"workload": f"{cfg['name']}: {n} particles ({cfg['n_per_gpu']} per GPU), {m} types, rmax={cfg['rmax']:.9g} (nx={stats['nx']}, "
                               f"{n / stats['nx'] ** 2:.1f} particles/cell), wrap={cfg['wrap']}, default accelerator, "
                               f"per-step cell-list rebuild, uniform-random state",
                   "l2": "per-GPU state (2 x 24 B x N) exceeds the 126 MB L2; no flush needed" if cfg["n_per_gpu"] * 48 > 126e6 else "state fits in L2: cache-resident",
                   "parallelism": "1 GPU" if world == 1 else f"{world} slabs over grid rows, halo exchange + particle migration every step via " + ("kernel pushes into CUDA-IPC peer memory over NVLink" if args.exchange == "peer" else "NCCL send/recv")},
        "pair_evals_per_sec": pair_rate,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "what": "per step: set_settings + set_matrix from host, plife_step, full snapshot (xy, vxy as fp32, type as u8: 17 B/particle) into pinned host memory" + " (copy of step k overlaps step k+1; every snapshot awaited)"},
        # per step: 3 scan + scatter + gather + force (binning is fused into the force pass); slabs add the halo pack
        # (writes into the neighbours' memory and signals), the halo wait + unpack, 2 migration pushes (with signal) and the
        # wait-and-collect of phase FINISH; arrival appends are not counted
        "gpu_launches": (6 if world == 1 else (11 if args.exchange == "peer" else 9)) * args.steps * world,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "force_kernel (3x3 force + friction + integrate + wrap)",
                     "achieved": force_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": force_gbs / hbm_peak if force_gbs else None, "traffic": 54.5 * cfg["n_per_gpu"] / 1e9,
                     "traffic_note": "GB per launch: dram__bytes_read+write of force_kernel_staged from profiles/r1_force_kernel.md (54.5 B/particle)",
                     "peak_source": peak_src, "algorithmic_bytes_per_particle": ALGO_BYTES_FORCE,
                     "binding": "fp32 issue (9*rho-1 = 143 pair evaluations per particle at 16 particles/cell); HBM fraction is low by construction, see fp32"},
        "roofline_step": {"bound": "hbm", "achieved": step_gbs, "peak": hbm_peak * world, "unit": "GB/s", "frac": step_gbs / (hbm_peak * world),
                          "algorithmic_bytes_per_particle_step": ALGO_BYTES_STEP},
        "fp32": {"achieved": fp32_force_tf, "peak": fp32_peak_tf, "unit": "TFLOP/s", "frac": fp32_force_tf / fp32_peak_tf if fp32_force_tf else None,
                 "whole_step_achieved": fp32_step_tf, "whole_step_frac": fp32_step_tf / (fp32_peak_tf * world),
                 "flop_per_pair_eval": FLOP_PER_PAIR, "peak_source": fp32_peak_src, "nominal_peak": fp32_nominal_tf},
        "kernel_ms_per_step": per_kernel,
        "secondary": secondary,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C3")
    ap.add_argument("--precision", default="f32", choices=["f32", "f64"])
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"], help="multi-GPU message transport")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
