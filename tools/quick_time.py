"""Ad-hoc timing probe (not a test): python tests/quick_time.py [config ...]"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # repo root (this file lives in tools/)
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "particle-life-app_b200"))
import numpy as np
import plife
from plife import synth

def run(name, steps=20, precision=plife.F32, flags=0):
    c = synth.CONFIGS[name]
    p = plife.NativePhysics(precision=precision, flags=flags)
    p.set_settings(c["rmax"], 0.85, 1.0, c["wrap"])
    p.random_matrix(c["m"], c["seed"])
    p.init_uniform(c["n"], c["seed"])
    p.step(0.02, 5); p.sync()
    t = time.perf_counter(); p.step(0.02, steps); p.sync(); dt = (time.perf_counter() - t) / steps
    st = p.step_stats()
    p.set_profiling(True); p.step(0.02, steps); kt = p.kernel_times(); p.set_profiling(False)
    print(f"{name} prec={precision} flags={flags}: {dt*1e3:.3f} ms/step  {c['n']/dt:.3e} particle-steps/s  "
          f"{st['pair_evals']/dt:.3e} pair-evals/s  nx={st['nx']}", flush=True)
    print("   per-kernel ms/step:", {k: round(v[0] / steps, 4) for k, v in kt.items()}, flush=True)

if __name__ == "__main__":
    names = sys.argv[1:] or ["C1", "C2", "C3", "C3lo", "C5"]
    for nme in names:
        run(nme)
    if not sys.argv[1:]:
        run("C3", precision=plife.F64, steps=5)
        run("C3", flags=plife.FLAG_UNSTABLE_SORT)
        run("C3", flags=4)
