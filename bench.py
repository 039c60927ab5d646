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
  secondary the other BASELINE configs on one GPU: C2 (1M, L2-resident), C3-lo (the HBM-bound
            regime, with its own roofline), C3 in fp64 (bit-exact mode), C5 (clamped borders)
  parity_check  multi-GPU runs: particle count conserved, pair evaluations equal to the value the
            cell histogram implies, and a 1M-particle slab run bit-equal to the single-GPU run
  cpu_baseline  the CPU oracle (C restatement of the reference algorithm, no JVM
            exists here) on the host cores, on the SAME workload, with all host threads and with
            the reference's default of 12 (B/Physics.java:37)

--impl reference times that CPU restatement on the same workload instead.
Every rank reports its own failure on stderr (rank, traceback, plife_last_error) before the
process group is torn down.
"""
from __future__ import annotations

import argparse
import faulthandler
import json
import os
import subprocess
import sys
import threading
import time
import traceback

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "particle-life-app_b200"))

METRIC = "particle_steps_per_sec"
UNIT = "particle-steps/s"
DT = 0.02
ALGO_BYTES_STEP = 84       # fp32 SoA three-pass design, bytes per particle-step
ALGO_BYTES_FORCE = 36      # pass C: read 20 + write 16
FLOP_PER_PAIR = 17         # wrap on, default accelerator
REF_THREADS = 12           # B/Physics.java:37 preferredNumberOfThreads


def log(msg):
    print(f"[bench rank {os.environ.get('RANK', '0')}] {msg}", file=sys.stderr, flush=True)


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


def workload_text(cfg, nx=None):
    import numpy as np
    nx = nx or int(np.floor(1.0 / cfg["rmax"]))
    return (f"{cfg['name']}: {cfg['n']} particles ({cfg['n_per_gpu']} per GPU), {cfg['m']} types, rmax={cfg['rmax']:.9g} "
            f"(nx={nx}, {cfg['n'] / nx ** 2:.1f} particles/cell), wrap={cfg['wrap']}, default accelerator, "
            f"per-step cell-list rebuild, uniform-random state")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic():
    """dram__bytes per particle of the dominant kernels, from the ncu --set full captures of this round
    (profiles/r2_traffic.json, written by hand from the capture named there)."""
    p = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return {}


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


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle's threaded variant (B/Physics.java:116-134 phase structure: serial makeContainers,
# two parallel passes with ceil(N/T) contiguous chunks as in B/LoadDistributor.java:42-56)
# ---------------------------------------------------------------------------------------------

class CpuArm:
    def __init__(self, cfg):
        import numpy as np
        import oracle
        from plife import synth
        self.cfg = cfg
        n = cfg["n"]
        need = n * 112 * 1.6  # two AoS buffers of 56 B + numpy staging
        avail = None
        try:
            with open("/proc/meminfo") as f:
                for ln in f:
                    if ln.startswith("MemAvailable:"):
                        avail = int(ln.split()[1]) * 1024
        except OSError:
            pass
        self.sample_n = n
        self.rmax = cfg["rmax"]
        if avail is not None and need > 0.5 * avail:  # does not fit this host: same-density sample instead (say so)
            self.sample_n = cfg["n_per_gpu"]
            self.rmax = cfg["rmax"] * (n / self.sample_n) ** 0.5
        pos, vel, types = synth.uniform_state(self.sample_n, cfg["m"], cfg["seed"])
        # the GPU arm's fp32 handle holds fp32-rounded positions; the oracle gets the same values
        pos = pos.astype(np.float32).astype(np.float64)
        M = synth.random_matrix(cfg["m"], cfg["seed"])
        self.o = oracle.Oracle(rmax=self.rmax, matrix=M, wrap=cfg["wrap"], dt=DT, threads=1)
        self.o.set_particles(pos, vel, types)
        del pos, vel, types
        self.nx = int(1 / self.rmax)

    def run(self, steps, warmup, threads, probed=1):
        """`probed` = steps already run on this state by the caller (the timing probe): they count as warm-up."""
        for _ in range(warmup):
            self.o.update(threads=threads)
        t0 = time.perf_counter()
        for _ in range(steps):
            self.o.update(threads=threads)
        dt = time.perf_counter() - t0
        evals = self.o.pair_stats()[0]
        full = self.sample_n == self.cfg["n"]
        sample = (f"{'the whole workload: ' if full else 'same-density sample (host memory): '}{self.sample_n} particles, rmax={self.rmax:.9g} "
                  f"({self.sample_n / self.nx ** 2:.1f} particles/cell), m={self.cfg['m']}, {steps} steps after {warmup + probed} warm-up, {threads} threads")
        return dict(value=self.sample_n * steps / dt, ms_per_step=dt / steps * 1e3, pair_evals_per_s=evals * steps / dt,
                    sample=sample, steps=steps, warmup=warmup, full=full)

    def bounded(self, threads, budget_s, max_steps, warmup=1):
        """1 probe step decides how many timed steps fit the budget (at least 1)."""
        t0 = time.perf_counter()
        self.o.update(threads=threads)
        probe = time.perf_counter() - t0
        steps = int(max(1, min(max_steps, budget_s / max(probe, 1e-6) - warmup)))
        return self.run(steps, max(0, warmup - 1), threads)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    world = max(1, args.gpus)
    cfg = scaled_workload(args.workload, world)
    threads = os.cpu_count() or 1
    arm = CpuArm(cfg)
    # the real workload: K steps after W warm-ups when that fits ~150 s of CPU time, otherwise fewer (stated)
    t0 = time.perf_counter()
    arm.o.update(threads=threads)
    probe = time.perf_counter() - t0
    want = max(1, args.steps) + max(0, args.warmup)
    if probe * want <= 150.0:
        steps, warm = max(1, args.steps), max(0, args.warmup - 1)
    else:
        warm = 0
        steps = int(max(1, min(args.steps, 150.0 / probe - 1)))
    r = arm.run(steps, warm, threads)
    note = None if steps == args.steps else f"timed {steps} of the requested {args.steps} steps (one CPU step of this workload takes {probe:.1f} s)"
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warm + 1, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_text(cfg), "sample": r["sample"], "same_workload_as_gpu_arm": r["full"], "note": note},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": threads, "kind": "port", "sample": r["sample"],
                         "note": "C restatement of the reference algorithm (no JVM available); an optimistic proxy for the Java path"},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "pair_evals_per_sec": r["pair_evals_per_s"],
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------

def timed_steps(torch, stream, fn, steps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    fn(steps)
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


def single_gpu_secondary(torch, plife, stream, device, name, precision, steps, label, hbm_peak, traffic, accel=None):
    c = workload(name)
    q = plife.NativePhysics(device=device, precision=precision, stream=stream.cuda_stream)
    try:
        q.set_settings(c["rmax"], 0.85, 1.0, c["wrap"])
        q.random_matrix(c["m"], c["seed"])
        if accel is not None:
            q.set_accelerator(accel)
        q.init_uniform(c["n"], c["seed"])
        with torch.cuda.stream(stream):
            q.step(DT, 5)
            torch.cuda.synchronize()
            ms = timed_steps(torch, stream, lambda k: q.step(DT, k), steps) / steps
            st = q.step_stats()
            q.set_profiling(True)
            q.step(DT, steps)
            kt = q.kernel_times()
            q.set_profiling(False)
        out = {"workload": label, "value": c["n"] / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps,
               "dtype": "f64" if precision == plife.F64 else "f32", "nx": st["nx"],
               "pair_evals_per_sec": st["pair_evals"] / (ms * 1e-3),
               "kernel_ms_per_step": {k: v[0] / steps for k, v in kt.items()}}
        bytes_step = ALGO_BYTES_STEP if precision == plife.F32 else 156
        gbs = bytes_step * c["n"] / (ms * 1e-3) / 1e9
        out["roofline_step"] = {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                                "algorithmic_bytes_per_particle_step": bytes_step,
                                "traffic": (traffic[name + "_step_bytes_per_particle"] * c["n"] / 1e9) if (precision == plife.F32 and accel is None and name + "_step_bytes_per_particle" in traffic) else None,
                                "traffic_note": "GB per step: dram__bytes_read + dram__bytes_write summed over the step's kernels (ncu --set full, profiles/r2_traffic.json)"}
        return q, out
    except Exception:
        q.close()
        raise


def slab_parity_check(torch, dist, plife, stream, rank, world, local_rank, exchange):
    """A 1M-particle state stepped by `world` slabs and by one GPU: every rank compares the particles it ends up
    owning with the single-GPU result, bit for bit (positions, velocities, types, order by id)."""
    import numpy as np
    from plife import synth
    from plife.slab import DistExchange, SlabPhysics, owner_of_position
    c = workload("C2")
    n, m, rmax, steps = c["n"], c["m"], c["rmax"], 12
    pos, vel, types = synth.uniform_state(n, m, c["seed"])
    pos = pos.astype(np.float32).astype(np.float64)
    M = synth.random_matrix(m, c["seed"])
    # same fine-bin count on both sides: the fp32 summation order follows the internal cell list
    single = plife.NativePhysics(device=local_rank, precision=plife.F32, stream=stream.cuda_stream, bins=8)
    single.set_settings(rmax, 0.85, 1.0, True)
    single.set_matrix(M)
    single.upload(pos, vel, types)
    with torch.cuda.stream(stream):
        single.step(DT, steps)
        ref = single.download()
    single.close()
    with torch.cuda.stream(stream):
        sl = SlabPhysics(rank, world, rmax, device=local_rank, capacity=n // world + n // 8 + 65536, halo_cap=16384, mig_cap=16384,
                         wrap=True, stream=stream.cuda_stream, exchange=exchange, bins=8)
        if exchange == "peer":
            sl.connect_dist()
        sl.native.set_matrix(M)
        own = owner_of_position(pos[:, 1], rmax, world) == rank
        sl.native.upload(pos[own], vel[own], types[own], np.arange(n, dtype=np.uint32)[own])
        dist.barrier()
        sl.step(DT, DistExchange(rank, world), steps)
        got = sl.native.download()
    lo, hi, nx = sl.rows()
    # the single-GPU particles whose final row lies in this rank's slab, in array order
    cy = np.minimum((ref.position[:, 1] / rmax).astype(np.int64), nx - 1)
    mine = (cy >= lo) & (cy < hi)
    # (the slab's array is residents-then-arrivals until its next cell-list build: compare as sets keyed by id)
    ig, ir = np.argsort(got.id), np.nonzero(mine)[0][np.argsort(ref.id[mine])]
    ok = (len(got.id) == int(mine.sum()) and np.array_equal(got.id[ig], ref.id[ir]) and np.array_equal(got.position[ig], ref.position[ir])
          and np.array_equal(got.velocity[ig], ref.velocity[ir]) and np.array_equal(got.type[ig], ref.type[ir]))
    # particles whose owner now differs from their owner at upload: the check is only worth something if some migrated
    moved = int((own[ref.id] != mine).sum())
    t = torch.tensor([1 if ok else 0, moved], device="cuda", dtype=torch.int64)
    dist.all_reduce(t[:1], op=dist.ReduceOp.MIN)
    dist.barrier()
    sl.native.close()
    return bool(int(t[0].item())), moved


def histogram_pair_evals(torch, dist, containers, nx, nly, first, rank, world, wrap, pos_xy, rmax, row_lo, device="cuda"):
    """Candidate pairs (i, j != i in the 3x3 cells of i) of this rank's OWNED rows as the cell histograms imply them:
    sum_c occ(c) * sum_{3x3} occ - n_owned.  The rows next to the slab come from their OWNERS' histograms (all-gather),
    not from this rank's ghost rows, so the number also checks that the halo exchange delivered the right rows.
    Particles sitting exactly on x == 1.0 or y == 1.0 (Range.wrap can return it) scan the cells around the
    UN-clamped coordinate (B/Physics.java:404-405), e.g. columns nx-1, 0, 1 instead of nx-2, nx-1, 0: corrected one by one."""
    import numpy as np
    ny = nx
    ends = containers.astype(np.int64).reshape(nly, nx)[1:nly - 1]  # owned rows (local rows 1 .. nly-2)
    occ = np.diff(np.concatenate([[first], ends.reshape(-1)])).reshape(nly - 2, nx)
    edge = torch.tensor(np.stack([occ[0], occ[1], occ[-1]]), device=device, dtype=torch.int64)  # first, second, last owned row
    edges = [torch.zeros_like(edge) for _ in range(world)]
    dist.all_gather(edges, edge)
    zero = np.zeros(nx, np.int64)
    has_dn, has_up = (wrap or rank > 0), (wrap or rank < world - 1)
    below = edges[(rank - 1) % world][2].cpu().numpy() if has_dn else zero
    above = edges[(rank + 1) % world][0].cpu().numpy() if has_up else zero
    above2 = edges[(rank + 1) % world][1].cpu().numpy() if has_up else zero
    full = np.concatenate([below[None], occ, above[None]])
    row3 = full + np.roll(full, 1, axis=1) + np.roll(full, -1, axis=1)
    nine = row3[:-2] + row3[1:-1] + row3[2:]
    total = int((occ * nine).sum() - occ.sum())
    # corrections for targets with an un-clamped coordinate == n (x or y exactly 1.0)
    x, y = pos_xy[:, 0].astype(np.float64), pos_xy[:, 1].astype(np.float64)
    cx0, cy0 = (x / rmax).astype(np.int64), (y / rmax).astype(np.int64)
    special = np.nonzero((cx0 >= nx) | (cy0 >= ny))[0]
    rows_of = {row_lo - 1 + k: full[k] for k in range(full.shape[0])}
    rows_of[row_lo + occ.shape[0] + 1] = above2

    def W(c, n):
        return c + n if c < 0 else (c - n if c >= n else c)
    exact = True
    for i in special:
        cols = [W(W(int(cx0[i]) + d, nx), nx) for d in (-1, 0, 1)]
        # (y: a slab scans the rows ny-2, ny-1, 0 around a particle on y == 1.0 - global row 1 lives two slabs away and holds
        # nothing within rmax of it - so its candidate count is the clamped row's; plife_internal.h: scan_row)
        cyi = min(int(cy0[i]), ny - 1)
        rws = [W(W(cyi + d, ny), ny) for d in (-1, 0, 1)]
        got = 0
        for r in rws:
            key = r if r in rows_of else (r + ny if (r + ny) in rows_of else r - ny)
            if key not in rows_of:
                exact = False
                continue
            got += int(sum(rows_of[key][c] for c in cols))
        ccx, ccy = min(int(cx0[i]), nx - 1), min(int(cy0[i]), ny - 1)
        total += got - int(nine[ccy - row_lo, ccx])
    return total, bool((occ >= 0).all()) and exact, len(special)


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
    extra_barriers = os.environ.get("PLIFE_BENCH_NO_BARRIERS", "0") != "1"  # off = the round-1 phase structure (for the repro)
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
    hbm_peak, peak_src = peaks()
    traffic = measured_traffic()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def phase(name):
        """every rank enters every stepping phase together: a rank that is late by more than the device-side spin
        limit would otherwise look like a dead neighbour"""
        if extra_barriers:
            barrier()
        if args.verbose:
            log(f"phase {name}")

    stream = torch.cuda.Stream()
    parity = None
    if world > 1 and not args.no_parity:
        ok, moved = slab_parity_check(torch, dist, plife, stream, rank, world, local_rank, args.exchange)
        parity = {"slab_vs_single_gpu_1M_bit_equal": ok, "what": f"C2 state (1M particles), 12 steps, {world} slabs vs one GPU: ids, positions, velocities, types bit-equal per rank; {moved} particles moved in or out of rank 0's slab"}
        if not ok:
            raise RuntimeError("parity check failed: the slab run differs from the single-GPU run")

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
        halo_cap = int(nx * rho * 1.5) + 4096

        def make_slab(mode):
            with torch.cuda.stream(stream):
                sl = SlabPhysics(rank, world, cfg["rmax"], device=local_rank, capacity=share + share // 8 + 65536,
                                 halo_cap=halo_cap, mig_cap=max(65536, share // 64), wrap=cfg["wrap"],
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

    # ---- resident-state throughput ----
    phase("warmup")
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

    # ---- correctness words of the timed run itself (multi-GPU) ----
    if world > 1:
        # the positions now, then one more step with dt = 0: it builds the cell list of exactly these positions, and its
        # cell offsets and pair count are what is checked below.  (Not "a step that moves nothing": a particle sitting on
        # y == 1.0 is wrapped to 0.0 by updatePosition even with dt = 0, B/Range.java:46-57, and changes slab.)
        pos32 = np.empty((p.count, 2), np.float32)
        p.download_f32(pos32, None, None)
        with torch.cuda.stream(stream):
            slab.step(0.0, ex, 1)
        stats = p.step_stats()
        n_local = p.count
        lo, hi, nxg = slab.rows()
        cont = p.containers_local(nxg * (hi - lo + 2))
        tot = torch.tensor([n_local, stats["pair_evals"]], device="cuda", dtype=torch.int64)
        dist.all_reduce(tot)
        conserved = int(tot[0].item()) == n
        expect, sane, n_special = histogram_pair_evals(torch, dist, cont, nxg, hi - lo + 2, halo_cap, rank, world, cfg["wrap"], pos32, cfg["rmax"], lo)
        del pos32
        if args.verbose or expect != stats["pair_evals"]:
            c2 = cont.astype(np.int64).reshape(hi - lo + 2, nxg)
            g_above = np.diff(np.concatenate([[c2[-2, -1]], c2[-1]]))  # the ghost rows as this rank holds them
            g_below = np.diff(c2[0])
            log(f"pair evaluations {stats['pair_evals']}, implied by the owners' cell histograms {expect} (rows [{lo},{hi}), n={n_local}, "
                f"{n_special} particles exactly on 1.0, sane={sane}); ghost rows held: above {int(g_above.sum())} particles "
                f"(min {int(g_above.min())}), below {int(g_below.sum())}+first cell, end of owned block {int(c2[-2, -1])} = first + n = {halo_cap + n_local}")
        flag = torch.tensor([1 if (sane and expect == stats["pair_evals"]) else 0], device="cuda", dtype=torch.int64)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        hist_ok = bool(int(flag.item()))
        parity = dict(parity or {})
        parity.update({"particles_conserved": conserved, "particles_total": int(tot[0].item()),
                       "pair_evals_total": int(tot[1].item()), "pair_evals_match_cell_histogram": hist_ok})
        if not conserved or hist_ok is False:
            raise RuntimeError(f"parity check failed after the timed steps: {parity}")

    # ---- per-kernel device time of the same steps (events between kernels) ----
    phase("profile")
    with torch.cuda.stream(stream):
        # slabs: every rank runs the same profiled steps (the exchange needs all of them); rank 0's times are reported,
        # and the "gather" bucket there also holds the halo pack / push / wait / unpack kernels
        p.set_profiling(True)
        run_steps(args.steps)
        kt = p.kernel_times()
        p.set_profiling(False)
    force_ms = kt["force"][0] / max(1, kt["force"][1])
    per_kernel = {k: v[0] / args.steps for k, v in kt.items()}
    per_rank = None
    if world > 1:
        # the ranks step in lockstep (every step waits for both neighbours' migrants), so the slowest GPU sets the pace:
        # report every rank's own kernel time next to the step time
        mine = torch.tensor([per_kernel["force"], sum(per_kernel.values())], device="cuda", dtype=torch.float64)
        allk = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allk, mine)
        per_rank = {"force_ms": [round(float(x[0]), 4) for x in allk], "kernels_ms": [round(float(x[1]), 4) for x in allk],
                    "note": "per rank: force bucket (both force launches, halo wait and unpack in between) and the sum of all buckets; ms_per_step minus the largest sum is what the end-of-step migration exchange and the rank-to-rank skew cost"}

    # ---- end to end through the C ABI with host buffers ----
    ncap = n if world == 1 else n_local + n_local // 8 + 65536
    pins = [tuple(torch.empty(shape, dtype=dt_, pin_memory=True) for shape, dt_ in
                  (((ncap, 2), torch.float32), ((ncap, 2), torch.float32), ((ncap,), torch.uint8)))  # u8: the app allows at most 256 types
            for _ in range(2)]
    h2d = (matrix.nbytes + 32) * world
    d2h = n * 17  # xy + vxy as fp32, type as u8
    e2e_steps = max(3, min(args.steps, 20))
    e2e_k = [0]
    e2e_pending = [0]

    def e2e_step():
        # settings + matrix go host -> device every step; the snapshot of step k copies out (pinned host memory,
        # second stream) while step k+1 computes; every snapshot is awaited before its buffer is reused
        p.set_settings(cfg["rmax"], 0.85, 1.0, cfg["wrap"])
        p.set_matrix(matrix)
        run_steps(1)
        a, b, c = pins[e2e_k[0] & 1]  # (the snapshot that last used these buffers was awaited one step ago)
        e2e_k[0] += 1
        p.snapshot_async(a.data_ptr(), b.data_ptr(), c.data_ptr(), types_u8=True)
        e2e_pending[0] += 1
        if e2e_pending[0] > 1:
            p.snapshot_wait()  # the PREVIOUS snapshot is complete (its buffers are the renderer's now); this one is in flight
            e2e_pending[0] -= 1

    def e2e_drain():
        while e2e_pending[0] > 0:
            p.snapshot_wait()
            e2e_pending[0] -= 1

    phase("e2e")
    with torch.cuda.stream(stream):
        for _ in range(3):
            e2e_step()
        e2e_drain()
        barrier()
        t0 = time.perf_counter()
        marks = []
        for _ in range(e2e_steps):
            e2e_step()
            marks.append(time.perf_counter() - t0)
        e2e_drain()
        barrier()
        e2e_s = time.perf_counter() - t0
        if args.verbose:
            log("e2e iteration ends (ms): " + " ".join(f"{m * 1e3:.2f}" for m in marks) + f" | drained {e2e_s * 1e3:.2f}")
    t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = n * e2e_steps / float(t.item())

    # what the host path can take: this rank's pinned D2H alone (others idle), then all ranks at once
    d2h_probe = None
    if world > 1 and not args.no_probe:
        src = torch.empty(n_local * 17, dtype=torch.uint8, device="cuda")
        dst = torch.empty(n_local * 17, dtype=torch.uint8, pin_memory=True)

        def copy_ms(reps=3):
            torch.cuda.synchronize()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for _ in range(reps):
                dst.copy_(src, non_blocking=True)
            a1.record()
            torch.cuda.synchronize()
            return a0.elapsed_time(a1) / reps
        copy_ms(1)
        solo = torch.zeros(world, device="cuda", dtype=torch.float64)
        for r in range(world):  # one rank at a time
            dist.barrier()
            if r == rank:
                solo[r] = src.numel() / (copy_ms() * 1e-3) / 1e9
        dist.all_reduce(solo)
        dist.barrier()
        allr = torch.tensor([src.numel() / (copy_ms() * 1e-3) / 1e9], device="cuda", dtype=torch.float64)
        alls = [torch.zeros_like(allr) for _ in range(world)]
        dist.all_gather(alls, allr)
        d2h_probe = {"pinned_d2h_GBps_one_rank_at_a_time": [round(float(x), 2) for x in solo.tolist()],
                     "pinned_d2h_GBps_all_ranks_together": [round(float(x.item()), 2) for x in alls],
                     "note": "bare cudaMemcpyAsync of one snapshot's bytes (17 B/particle) device -> pinned host, no physics"}
        del src, dst

    secondary = None
    if world == 1 and args.workload == "C3" and precision == plife.F32 and not args.no_secondary:
        secondary = {}
        # BASELINE config 2 (1M particles, 8 types): the metric names both sizes; this one fits in L2
        q, s2 = single_gpu_secondary(torch, plife, stream, local_rank, "C2", plife.F32, 200,
                                     "C2: 1000000 particles, 8 types, rmax=0.004 (16 particles/cell), state fits in L2 (cache-resident)", hbm_peak, traffic)
        # the same state after it has evolved into clusters (what the app actually runs): most cells empty, some with
        # thousands of particles; candidates per particle grow by an order of magnitude, so particle-steps/s drops while
        # pair-evals/s does not
        with torch.cuda.stream(stream):
            q.step(DT, 2000 - 410)
            torch.cuda.synchronize()
            ms3 = timed_steps(torch, stream, lambda k: q.step(DT, k), 50) / 50
        pe = q.step_stats()["pair_evals"]
        s2["evolved"] = {"workload": "the C2 state 2000 steps later (clustered)", "value": 1_000_000 / (ms3 * 1e-3), "unit": UNIT,
                         "ms_per_step": ms3, "pair_evals_per_particle": pe / 1_000_000, "pair_evals_per_sec": pe / (ms3 * 1e-3)}
        q.close()
        secondary["C2"] = s2
        # the HBM-bound regime: same 16M particles on a 4096^2 grid (0.95 particles/cell, 7.6 candidates per particle)
        q, s = single_gpu_secondary(torch, plife, stream, local_rank, "C3lo", plife.F32, 30,
                                    "C3-lo: 16000000 particles, 16 types, rmax=2^-12 (nx=4096, 0.95 particles/cell): HBM-bound regime", hbm_peak, traffic)
        q.close()
        secondary["C3lo"] = s
        # BASELINE config 3's fp64 mode: bit-exact with the CPU oracle
        q, s = single_gpu_secondary(torch, plife, stream, local_rank, "C3", plife.F64, 10,
                                    "C3 in fp64 (bit-exact mode): 16000000 particles, 16 types, rmax=0.001", hbm_peak, traffic)
        q.close()
        secondary["C3_f64"] = s
        # BASELINE config 5: clamped (non-wrapping) borders, and one builder-defined accelerator
        q, s = single_gpu_secondary(torch, plife, stream, local_rank, "C5", plife.F32, 50,
                                    "C5: 4000000 particles, 8 types, rmax=0.002, wrap=False (clamped borders), default accelerator", hbm_peak, traffic)
        q.close()
        secondary["C5_clamp"] = s
        c5w = dict(workload("C5"))
        q, s = single_gpu_secondary(torch, plife, stream, local_rank, "C5", plife.F32, 50,
                                    "C5: 4000000 particles, 8 types, rmax=0.002, wrap=False, accelerator kind 3 (rotator, builder-defined)", hbm_peak, traffic,
                                    accel=plife.ACC_ROTATOR_90)
        q.close()
        secondary["C5_rotator"] = s

        # BASELINE config 1 (the app's own default scale; the reference's CPU-runnable case): 10 000 particles at rmax 0.04
        # and at the snapshot default rmax 0.02, replayed as a CUDA graph
        for nm, label in (("C1", "C1: 10000 particles, 6 types, rmax=0.04 (nx=25, 16 particles/cell): launch-latency regime, CUDA-graph replay"),
                          ("C1d", "C1 at the reference's default rmax=0.02 (nx=50, 4 particles/cell)")):
            q, s = single_gpu_secondary(torch, plife, stream, local_rank, nm, plife.F32, 100, label, hbm_peak, traffic)  # (100 steps: still the uniform state)
            s["graph_steps"] = q.step_stats().get("graph_steps")
            q.close()
            secondary[nm] = s
        if not args.no_cpu:
            # the CPU arm at C1 (both rmax) and C2 as well, all host threads and the reference's default 12
            threads = os.cpu_count() or 1
            for nm, steps in (("C1", 100), ("C1d", 100), ("C2", 10)):
                c = workload(nm)
                c["n_per_gpu"] = c["n"]
                arm = CpuArm(c)
                r = arm.run(steps, 3, threads, probed=0)
                entry = {"value": r["value"], "unit": UNIT, "cores": threads, "kind": "port", "sample": r["sample"], "ms_per_step": r["ms_per_step"]}
                if threads != REF_THREADS:
                    r12 = arm.run(steps, 1, REF_THREADS, probed=0)
                    entry["reference_default_threads"] = {"value": r12["value"], "unit": UNIT, "cores": REF_THREADS, "ms_per_step": r12["ms_per_step"]}
                secondary[nm]["cpu_baseline"] = entry
                del arm

    if rank != 0:
        return 0

    force_gbs = ALGO_BYTES_FORCE * n_local / (force_ms * 1e-3) / 1e9 if force_ms > 0 else None  # one launch = this rank's particles
    step_gbs = ALGO_BYTES_STEP * n / (ms / args.steps * 1e-3) / 1e9
    pair_total = parity["pair_evals_total"] if (parity and "pair_evals_total" in parity) else stats["pair_evals"] * world
    pair_rate = pair_total * args.steps / (ms * 1e-3)
    fp32_nominal_tf = 148 * 128 * 2 * (clocks["sm_max_mhz"] or 1965.0) * 1e6 / 1e12
    import ctypes
    meas = ctypes.c_double(0.0)
    fp32_peak_tf = meas.value if N.lib().plife_measure_fp32_peak(local_rank, ctypes.byref(meas)) == 0 and meas.value > 0 else fp32_nominal_tf
    fp32_peak_src = "measured FFMA loop on this GPU (plife_measure_fp32_peak)" if meas.value > 0 else "nominal"
    fp32_force_tf = FLOP_PER_PAIR * stats["pair_evals"] / (force_ms * 1e-3) / 1e12 if force_ms > 0 else None
    fp32_step_tf = FLOP_PER_PAIR * pair_total / (ms / args.steps * 1e-3) / 1e12

    cpu = None
    if world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        arm = CpuArm(cfg)
        r = arm.bounded(threads, 12.0, 6)
        cpu = {"value": r["value"], "unit": UNIT, "cores": threads, "kind": "port", "sample": r["sample"],
               "pair_evals_per_sec": r["pair_evals_per_s"], "ms_per_step": r["ms_per_step"],
               "note": "C restatement of the reference algorithm (no JVM available); an optimistic proxy for the Java path"}
        if threads != REF_THREADS:
            r12 = arm.bounded(REF_THREADS, 12.0, 4)
            cpu["reference_default_threads"] = {"value": r12["value"], "unit": UNIT, "cores": REF_THREADS, "sample": r12["sample"],
                                                "ms_per_step": r12["ms_per_step"],
                                                "note": "T = 12 is the reference's default preferredNumberOfThreads (B/Physics.java:37)"}

    tr_force = traffic.get("C3_force_bytes_per_particle")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if precision == plife.F32 else "f64", "data": "synthetic",
        "config": {"workload": workload_text(cfg, stats["nx"]),
                   "l2": "per-GPU state (2 x 24 B x N) exceeds the 126 MB L2; no flush needed" if cfg["n_per_gpu"] * 48 > 126e6 else "state fits in L2: cache-resident",
                   "parallelism": "1 GPU" if world == 1 else f"{world} slabs over grid rows, halo exchange + particle migration every step via " + ("kernel pushes into CUDA-IPC peer memory over NVLink" if args.exchange == "peer" else "NCCL send/recv")},
        "pair_evals_per_sec": pair_rate,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                "what": "per step: set_settings + set_matrix from host, plife_step, full snapshot (xy, vxy as fp32, type as u8: 17 B/particle) into pinned host memory" + " (the copy of snapshot k overlaps step k+1 and the taking of snapshot k+1; every snapshot awaited)",
                "host_path_probe": d2h_probe},
        "gpu_launches": launches_per_step(world, args.exchange) * args.steps * world,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "force_kernel (3x3 force + friction + integrate + wrap)",
                     "achieved": force_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": force_gbs / hbm_peak if force_gbs else None,
                     "traffic": tr_force * cfg["n_per_gpu"] / 1e9 if tr_force else None,
                     "traffic_note": traffic.get("C3_force_note"),
                     "peak_source": peak_src, "algorithmic_bytes_per_particle": ALGO_BYTES_FORCE,
                     "binding": "fp32 issue (candidate pair evaluations per particle at 16 particles/cell); HBM fraction is low by construction, see fp32 and secondary.C3lo for the HBM-bound regime"},
        "roofline_step": {"bound": "hbm", "achieved": step_gbs, "peak": hbm_peak * world, "unit": "GB/s", "frac": step_gbs / (hbm_peak * world),
                          "algorithmic_bytes_per_particle_step": ALGO_BYTES_STEP,
                          "traffic": traffic.get("C3_step_bytes_per_particle", 0) * n / 1e9 or None,
                          "traffic_note": "GB per step over all GPUs: dram bytes of the step's kernels (ncu, profiles/r2_traffic.json)"},
        "fp32": {"achieved": fp32_force_tf, "peak": fp32_peak_tf, "unit": "TFLOP/s", "frac": fp32_force_tf / fp32_peak_tf if fp32_force_tf else None,
                 "whole_step_achieved": fp32_step_tf, "whole_step_frac": fp32_step_tf / (fp32_peak_tf * world),
                 "flop_per_pair_eval": FLOP_PER_PAIR, "peak_source": fp32_peak_src, "nominal_peak": fp32_nominal_tf,
                 "pair_evals_counted": "the reference's candidate pairs (3x3 cells, B/Physics.java:423-439); the kernel itself evaluates fewer (finer internal binning)"},
        "kernel_ms_per_step": per_kernel,
        "kernel_ms_per_step_by_rank": per_rank,
        "parity_check": parity,
        "secondary": secondary,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    return 0


def launches_per_step(world, exchange):
    # per step: scan (one launch, decoupled look-back) + scatter + gather + force (binning is fused into the force pass); slabs
    # add the halo pack (written into the neighbours' memory, then a flag), the halo wait + unpack, a second force launch
    # (interior rows / edge rows) and ONE kernel for phase FINISH (push migrants + wait + new counts + append arrivals)
    return 4 if world == 1 else 8


def main():
    faulthandler.enable(all_threads=True)  # a native abort (SIGABRT / SIGSEGV) still leaves a Python traceback on stderr
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C3")
    ap.add_argument("--precision", default="f32", choices=["f32", "f64"])
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"], help="multi-GPU message transport")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-probe", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    args = ap.parse_args()
    rc = 1
    try:
        rc = run_reference(args) if args.impl == "reference" else run_ours(args)
    except BaseException as e:  # noqa: BLE001 - report, then fail
        if isinstance(e, SystemExit) and not e.code:
            rc = 0
        else:
            log(f"FAILED: {type(e).__name__}: {e}")
            traceback.print_exc(file=sys.stderr)
            sys.stderr.flush()
            rc = 1
    finally:
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                if rc == 0:
                    dist.destroy_process_group()
                else:
                    # peers may be blocked in a collective or spinning on this rank's flag: leave without the collective
                    # teardown, torchrun ends them; os._exit avoids the c10d destructor's abort hiding the real error
                    sys.stdout.flush()
                    os._exit(rc)
        except Exception:  # noqa: BLE001
            pass
    return rc


if __name__ == "__main__":
    sys.exit(main())
