"""Writes the INPUTS of the golden fixtures (accelerator kind 0 only: the one the reference ships) as text files that
particle-life-app_b200/java/com/particle_life/backend/DumpVectors.java reads.  Doubles travel as raw IEEE bits in hex."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def hx(v):
    return format(int(np.float64(v).view(np.uint64)), "x")


def main():
    os.makedirs(os.path.join(HERE, "java"), exist_ok=True)
    for f in sorted(os.listdir(HERE)):
        if not f.endswith(".npz"):
            continue
        z = np.load(os.path.join(HERE, f))
        if int(z["accel_kind"]) != 0:
            continue
        pos, vel, typ, M = z["pos0"], z["vel0"], z["typ0"], z["matrix"]
        with open(os.path.join(HERE, "java", f[:-4] + ".in.txt"), "w") as o:
            o.write(f"{len(typ)} {M.shape[0]} {int(z['steps'])} {int(bool(z['wrap']))} {hx(z['rmax'])} {hx(0.85)} {hx(1.0)} {hx(0.02)}\n")
            o.write(" ".join(hx(v) for v in M.reshape(-1)) + "\n")
            for i in range(len(typ)):
                o.write(f"{hx(pos[i, 0])} {hx(pos[i, 1])} {hx(vel[i, 0])} {hx(vel[i, 1])} {int(typ[i])}\n")
        print("wrote", f[:-4] + ".in.txt")


if __name__ == "__main__":
    main()
