"""Driver for compute-sanitizer over the dense-cluster code paths (chunked staging, hit lists, list rebase, rank loop):
compute-sanitizer --tool memcheck python tools/sanitize_dense.py"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "particle-life-app_b200"))
import numpy as np
import plife

rng = np.random.default_rng(5)


def blob(n, centre, sigma, n_bg):
    pos = np.concatenate([np.asarray(centre) + rng.normal(0, sigma, (n, 2)), rng.random((n_bg, 2))])
    return np.clip(pos, 0, 0.999999)


def run(precision, pos, m, rmax, wrap, accel=(0, (0.3,)), flags=0, steps=2):
    p = plife.NativePhysics(precision=precision, flags=flags)
    p.set_settings(rmax, 0.85, 1.0, wrap)
    p.set_matrix(rng.random((m, m)) * 2 - 1)
    p.set_accelerator(*accel)
    p.upload(pos, None, rng.integers(0, m, len(pos)).astype(np.int32))
    p.step(0.02, steps)
    out = p.download()
    assert np.isfinite(out.position).all() and np.isfinite(out.velocity).all()
    p.close()


# several chunks per row, CTAs ending inside a chunk, seam blob, trailing partial CTA
pos = np.concatenate([blob(6000, (0.5, 0.5), 0.03, 0), blob(2000, (0.98, 0.02), 0.01, 3001)])
for wrap in (True, False):
    run(plife.F32, pos, 4, 0.02, wrap)
    run(plife.F32, pos, 4, 0.02, wrap, accel=(3, ()))
    run(plife.F32, pos, 4, 0.02, wrap, flags=plife.FLAG_FORCE_V1)
    run(plife.F64, pos, 4, 0.02, wrap)
# one cell holding more candidates than the hit list's 14-bit offsets: rebase + many flushes (fp64), many chunks (fp32)
pos = blob(17000, (0.525, 0.475), 0.006, 500)
run(plife.F64, pos, 3, 0.05, True, steps=1)
run(plife.F32, pos, 3, 0.05, True, steps=1)
print("dense paths ok")
