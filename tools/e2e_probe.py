"""Timing probe (not a test): the end-to-end loop of bench.py (step + full snapshot to pinned host memory) in the two orders."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "particle-life-app_b200"))
import numpy as np, torch
import plife
from plife import synth

c = synth.CONFIGS["C3"]; n = c["n"]
ts = torch.cuda.Stream() if len(sys.argv) > 1 and sys.argv[1] == 'torch' else None
p = plife.NativePhysics(stream=ts.cuda_stream) if ts else plife.NativePhysics()
p.set_settings(c["rmax"], 0.85, 1.0, c["wrap"]); p.random_matrix(c["m"], c["seed"]); p.init_uniform(n, c["seed"])
M = synth.random_matrix(c["m"], c["seed"])
pins = [tuple(torch.empty(s, dtype=d, pin_memory=True) for s, d in (((n, 2), torch.float32), ((n, 2), torch.float32), ((n,), torch.uint8))) for _ in range(2)]
p.step(0.02, 5); p.sync()

def loop(order, steps, settings=True):
    k = 0
    t0 = time.perf_counter()
    marks = []
    for _ in range(steps):
        if settings:
            p.set_settings(c["rmax"], 0.85, 1.0, c["wrap"]); p.set_matrix(M)
        p.step(0.02, 1)
        a, b, cc = pins[k & 1]; k += 1
        if order == "old":
            p.snapshot_wait()
            p.snapshot_async(a.data_ptr(), b.data_ptr(), cc.data_ptr(), types_u8=True)
        else:
            p.snapshot_async(a.data_ptr(), b.data_ptr(), cc.data_ptr(), types_u8=True)
            if k > 1: p.snapshot_wait()
        marks.append(time.perf_counter() - t0)
    p.snapshot_wait(); p.snapshot_wait()
    dt = (time.perf_counter() - t0) / steps
    return dt, np.diff(marks)

for order, st in (("old", True), ("new", True)):
    loop(order, 3, st)
    dt, d = loop(order, 20, st)
    print(f"{order} settings={st}: {dt*1e3:.3f} ms/step = {n*17/dt/1e9:.1f} GB/s; iteration gaps ms: {np.round(d[:8]*1e3,2)}", flush=True)
