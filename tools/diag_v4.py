"""Diagnostic (not a test): the catalog-setter test body, printing the per-generator error against the oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "particle-life-app_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import plife, oracle
from plife import setters as S

def oracle_step(pos, vel, types, matrix, ids, rmax, dt):
    o = oracle.Oracle(matrix=matrix, rmax=rmax, dt=dt, threads=os.cpu_count() or 1)
    o.set_particles(pos, vel, types, ids)
    o.update()
    return o

p = plife.Physics(position_setter=S.POSITION_SETTERS["color battle"], matrix_generator=S.MATRIX_GENERATORS["snakes"],
                  type_setter=S.TYPE_SETTERS["more of first"], particle_count=6000, seed=11)
p.settings.rmax = 0.05
for name in S.POSITION_SETTERS:
    p.position_setter = S.POSITION_SETTERS[name]; p.set_positions()
for name in S.TYPE_SETTERS:
    p.type_setter = S.TYPE_SETTERS[name]; p.set_types()
p.set_type_count_equal()
p.set_type_count([10, 20, 30, 40, 50, 7000])
p.set_type_count([500, 500, 500, 500, 500, 500])
for gname, gen in S.MATRIX_GENERATORS.items():
    p.matrix_generator = gen; p.generate_matrix()
    b = p.particles
    o = oracle_step(b.position, b.velocity, b.type, p.settings.matrix, b.id, 0.05, 0.02)
    p.update()
    a = p.particles
    _, ovel, _, oid = o.get_particles()
    d = a.velocity - ovel
    err = np.linalg.norm(d) / max(np.linalg.norm(ovel), 1e-300)
    worst = int(np.argmax(np.abs(d).sum(axis=1)))
    cnt = np.bincount((np.floor(b.position[:, 0] / 0.05).astype(int).clip(0, 19) + 20 * np.floor(b.position[:, 1] / 0.05).astype(int).clip(0, 19)), minlength=400)
    print(f"{gname:24s} rel_l2={err:.3e} order_ok={np.array_equal(a.id, oid)} worst={worst} d={d[worst]} ovel={ovel[worst]} pos_before={b.position[np.where(b.id == a.id[worst])[0][0]]} maxcell={cnt.max()} stats={p.native.step_stats() if hasattr(p, 'native') else ''}", flush=True)
