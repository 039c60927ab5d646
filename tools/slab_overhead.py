"""What slab mode costs on ONE GPU with no neighbours (world = 1): the local extra kernels (halo pack / unpack, header
reset) and the per-step host synchronisation, against the plain step.  python tools/slab_overhead.py [config] [steps]"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "particle-life-app_b200"))
import plife
from plife import synth
from plife.slab import SlabPhysics, grid_rows


def timed(fn, sync, steps):
    fn(5); sync()
    t = time.perf_counter(); fn(steps); sync()
    return (time.perf_counter() - t) / steps * 1e3


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "C3"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    c = synth.CONFIGS[name]
    p = plife.NativePhysics()
    p.set_settings(c["rmax"], 0.85, 1.0, c["wrap"])
    p.random_matrix(c["m"], c["seed"])
    p.init_uniform(c["n"], c["seed"])
    plain = timed(lambda k: p.step(0.02, k), p.sync, steps)
    p.close()
    nx = grid_rows(c["rmax"])
    rho = c["n"] / nx ** 2
    s = SlabPhysics(0, 1, c["rmax"], capacity=c["n"] + c["n"] // 8, halo_cap=int(nx * rho * 1.5) + 4096, mig_cap=65536, wrap=c["wrap"])
    s.native.random_matrix(c["m"], c["seed"])
    s.native.init_uniform(c["n"], c["seed"])
    slab = timed(lambda k: s.step(0.02, None, k), s.native.sync, steps)
    print(f"{name}: plain {plain:.3f} ms/step, slab(world=1) {slab:.3f} ms/step, overhead {1e3 * (slab - plain):.0f} us", flush=True)


if __name__ == "__main__":
    main()
