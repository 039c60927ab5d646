"""Timing of the builder-defined accelerator kinds (literal force visitors): python tools/quick_time_accel.py [config]"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "particle-life-app_b200"))
import plife
from plife import synth

name = sys.argv[1] if len(sys.argv) > 1 else "C5"
c = synth.CONFIGS[name]
for kind, params in [(0, (0.3,)), (1, (0.3,)), (2, (0.3,)), (3, ()), (4, ()), (5, ())]:
    p = plife.NativePhysics()
    p.set_settings(c["rmax"], 0.85, 1.0, True)
    p.random_matrix(c["m"], c["seed"])
    p.set_accelerator(kind, params)
    p.init_uniform(c["n"], c["seed"])
    p.step(0.02, 5); p.sync()
    t = time.perf_counter(); p.step(0.02, 20); p.sync(); dt = (time.perf_counter() - t) / 20
    print(f"{name} kind {kind}: {dt*1e3:.3f} ms/step  {c['n']/dt:.3e} particle-steps/s", flush=True)
    p.close()
