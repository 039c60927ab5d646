"""Timing probe (not a test): C5 (4M particles, clamped borders) with every accelerator kind."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "particle-life-app_b200"))
import plife
from plife import synth

c = synth.CONFIGS["C5"]
for kind in range(6):
    p = plife.NativePhysics()
    p.set_settings(c["rmax"], 0.85, 1.0, c["wrap"])
    p.random_matrix(c["m"], c["seed"])
    p.set_accelerator(kind, [0.3, 0.0, 0.0, 0.0])
    p.init_uniform(c["n"], c["seed"])
    p.step(0.02, 5); p.sync()
    t = time.perf_counter(); p.step(0.02, 20); p.sync(); dt = (time.perf_counter() - t) / 20
    print(f"C5 accelerator kind {kind}: {dt * 1e3:.3f} ms/step", flush=True)
    p.close()
