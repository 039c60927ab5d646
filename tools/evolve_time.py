"""How the step time moves as a uniform state evolves into clusters (not a test, not a bench value):
python tools/evolve_time.py [config] [total_steps] [chunk] [flags]"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "particle-life-app_b200"))
import numpy as np
import plife
from plife import synth


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "C2"
    total = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
    chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 250
    flags = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    c = synth.CONFIGS[name]
    p = plife.NativePhysics(flags=flags)
    p.set_settings(c["rmax"], 0.85, 1.0, c["wrap"])
    p.random_matrix(c["m"], c["seed"])
    p.init_uniform(c["n"], c["seed"])
    p.step(0.02, 3); p.sync()
    done = 3
    print(f"{name}: n={c['n']} m={c['m']} rmax={c['rmax']} flags={flags}", flush=True)
    while done < total:
        t = time.perf_counter(); p.step(0.02, chunk); p.sync(); dt = (time.perf_counter() - t) / chunk
        done += chunk
        st = p.step_stats()
        ends = np.asarray(p.containers())
        counts = np.diff(np.r_[0, ends])
        p.set_profiling(True); p.step(0.02, 4); kt = p.kernel_times(); p.set_profiling(False)
        done += 4
        print(f"  step {done:6d}: {dt*1e3:8.3f} ms/step  {c['n']/dt:.3e} p-steps/s  pair_evals/particle={st['pair_evals']/c['n']:8.1f} "
              f"max cell={counts.max():5d}  empty cells={np.mean(counts == 0):.2f}  "
              f"kernels={ {k: round(v[0] / 4, 3) for k, v in kt.items() if v[0] > 0} }", flush=True)


if __name__ == "__main__":
    main()
