"""Compares what the REAL reference produced (DumpVectors.java -> tests/golden/java/NAME.out.txt) with the committed
fixtures (outputs of the C oracle).  Bit equality is expected: the oracle restates the reference's operation order,
including JOML's div = multiply by reciprocal and non-fused length / mulAdd.  A difference of an ulp would point at
`-Djoml.useMathFma` or at Math.pow (B/Physics.java:401, within 1 ulp by contract); the report says which."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    bad = 0
    for f in sorted(os.listdir(os.path.join(HERE, "java"))):
        if not f.endswith(".out.txt"):
            continue
        name = f[:-8]
        z = np.load(os.path.join(HERE, name + ".npz"))
        n, steps = len(z["typ0"]), int(z["steps"])
        rows = [ln.split() for ln in open(os.path.join(HERE, "java", f)) if ln.strip()]
        assert len(rows) == n * steps, f"{name}: {len(rows)} rows, expected {n * steps}"
        raw = np.array([[int(c, 16) for c in r[:4]] for r in rows], dtype=np.uint64).view(np.float64).reshape(steps, n, 4)
        typ = np.array([int(r[4]) for r in rows]).reshape(steps, n)
        for s in range(steps):
            pos, vel = z[f"pos{s + 1}"], z[f"vel{s + 1}"]
            order_ok = np.array_equal(typ[s], z[f"typ{s + 1}"])
            dp, dv = np.abs(raw[s, :, :2] - pos).max(), np.abs(raw[s, :, 2:] - vel).max()
            exact = order_ok and np.array_equal(raw[s, :, :2], pos) and np.array_equal(raw[s, :, 2:], vel)
            print(f"{name} step {s + 1}: {'bit-exact' if exact else f'DIFFERS (order ok: {order_ok}, max |dpos| {dp:.3e}, max |dvel| {dv:.3e})'}")
            bad += 0 if exact else 1
    if bad == 0:
        print("oracle pinned: every step of every fixture is bit-identical to the reference's Physics.update()")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
