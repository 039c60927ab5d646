"""Known-answer tests that pin the CPU oracle (none exist in the reference: SURVEY.md 4, Appendix C).

Every expected value is derived by hand from the formulas at A/Main.java:275-280 (accelerator),
B/Range.java (wrap / wrapConnection / clamp) and B/Physics.java:362-375,401,437,447.
"""
import numpy as np
import pytest

import oracle
from oracle import bruteforce


def test_force_law_breakpoints():
    f = oracle.lib().oracle_particle_life_force
    assert f(0.7, 0.15, 0.3) == pytest.approx(-0.5, abs=1e-15)   # repulsion is independent of a
    assert f(-0.7, 0.15, 0.3) == pytest.approx(-0.5, abs=1e-15)
    assert f(0.9, 0.3, 0.3) == pytest.approx(0.0, abs=1e-15)
    assert f(1.0, 0.65, 0.3) == pytest.approx(1.0, abs=1e-15)    # peak of the triangular lobe
    assert f(0.5, 0.475, 0.3) == pytest.approx(0.25, abs=1e-15)
    assert f(-1.0, 0.825, 0.3) == pytest.approx(-0.5, abs=1e-15)
    assert f(0.8, 1.0, 0.3) == pytest.approx(0.0, abs=1e-15)
    assert f(0.8, 1e-12, 0.3) == pytest.approx(-1.0, abs=1e-11)


def test_range_wrap_clamp_examples():
    L = oracle.lib()
    assert L.oracle_range_wrap(-0.3) == pytest.approx(0.7)        # B/Range.java:24-26 doc example
    assert L.oracle_range_wrap(2.0) == 0.0
    assert L.oracle_range_wrap(0.4) == 0.4
    assert L.oracle_range_wrap(-1e-17) == 1.0                     # E1: can return exactly 1.0
    assert L.oracle_range_wrap(-2.0) == 0.0
    assert L.oracle_range_wrap(-5.3) == pytest.approx(0.7)
    assert L.oracle_range_clamp(-1.3) == 0.0                      # B/Range.java:14-17 doc example
    assert L.oracle_range_clamp(2.0) == 1.0
    assert L.oracle_range_clamp(0.4) == 0.4
    assert L.oracle_range_wrap_connection(-0.98) == pytest.approx(0.02)
    assert L.oracle_range_wrap_connection(0.5) == -0.5            # [-0.5, 0.5)
    assert L.oracle_range_wrap_connection(-0.5) == -0.5
    assert L.oracle_range_wrap_connection(0.49) == 0.49


def test_container_index():
    import ctypes as C
    nx = C.c_int32()
    L = oracle.lib()
    assert L.oracle_container_index(0.04, 0.999999, 0.0, C.byref(nx)) == 24 and nx.value == 25
    assert L.oracle_container_index(0.04, 1.0, 0.0, None) == 24         # == nx clamp
    assert L.oracle_container_index(0.04, 1.0, 1.0, None) == 24 + 24 * 25
    assert L.oracle_container_index(0.065, 0.98, 0.0, C.byref(nx)) == 14 and nx.value == 15  # fat last cell
    assert L.oracle_container_index(0.02, 0.5, 0.25, C.byref(nx)) == 25 + 12 * 50 and nx.value == 50
    # GUI float-widened rmax (A/utils/ImGuiUtils.java:83-85): 0.04f -> nx stays 25
    assert L.oracle_container_index(float(np.float32(0.04)), 0.5, 0.5, C.byref(nx)) >= 0 and nx.value == 25


def _two(pos, wrap=True, rmax=0.04, matrix=None, **kw):
    o = oracle.Oracle(rmax=rmax, wrap=wrap, matrix=np.ones((1, 1)) if matrix is None else matrix, **kw)
    o.set_particles(np.array(pos, np.float64), None, np.zeros(len(pos), np.int32))
    o.update()
    p, v, t, i = o.get_particles()
    order = np.argsort(i)
    return p[order], v[order], o


def test_two_particle_step():
    p, v, _ = _two([[0.5, 0.5], [0.5 + 0.65 * 0.04, 0.5]])
    assert 0.85 ** 1.2 == pytest.approx(0.8228159672521091, abs=1e-16)
    assert v[0, 0] == pytest.approx(8.0e-4, abs=1e-15) and v[1, 0] == pytest.approx(-8.0e-4, abs=1e-15)
    assert v[0, 1] == 0 and v[1, 1] == 0
    assert p[0, 0] == pytest.approx(0.500016, abs=1e-15) and p[1, 0] == pytest.approx(0.525984, abs=1e-15)


def test_seam_pair():
    k = 0.5714285714285714 * 0.04 * 0.02
    _, v, _ = _two([[0.99, 0.5], [0.01, 0.5]], wrap=True)
    assert v[0, 0] == pytest.approx(k, abs=1e-16) and v[1, 0] == pytest.approx(-k, abs=1e-16)
    _, v, _ = _two([[0.99, 0.5], [0.01, 0.5]], wrap=False)
    assert np.all(v == 0)


def test_x_equal_one_scans_cells_24_0_1():
    # x == 1.0 sits in container 24 but its force pass uses cx0 = 25 -> cells 24, 0, 1 (Appendix C)
    p, v, o = _two([[1.0, 0.5], [0.02, 0.5], [0.97, 0.5]], wrap=True)
    assert o.pair_stats()[1] == 4  # (0,1),(1,0),(0,2),(2,0) in range; (1,2) is 0.05 apart
    assert v[0, 0] != 0


def test_duplicate_visit_quirk_nx2():
    # rmax = 0.5 -> nx = 2: offsets -1 and +1 alias the same cell, pairs are counted twice per
    # aliased axis (E4, reference bug reproduced on purpose)
    pos = [[0.45, 0.25], [0.55, 0.25]]
    _, v, o = _two(pos, rmax=0.5)
    evals, hits = o.pair_stats()
    assert (evals, hits) == (4, 4)  # each particle meets the other twice: its cell is reached via ox=-1 and ox=+1
    _, vb, _ = bruteforce.step(np.array(pos), np.zeros((2, 2)), np.zeros(2, int), np.ones((1, 1)), rmax=0.5)
    ratio = v[0, 0] / vb[0, 0]
    assert ratio == pytest.approx(2.0, rel=1e-12)


def test_coincident_particles_exert_no_force():
    _, v, o = _two([[0.3, 0.3], [0.3, 0.3]])
    assert np.all(v == 0) and o.pair_stats() == (2, 0)


def test_clamp_keeps_velocity():
    o = oracle.Oracle(rmax=0.04, wrap=False, matrix=np.zeros((1, 1)), friction=1.0)
    o.set_particles(np.array([[0.999, 0.5]]), np.array([[1.0, 0.0]]), np.zeros(1, np.int32))
    o.update()
    p, v, _, _ = o.get_particles()
    assert p[0, 0] == 1.0 and v[0, 0] == 1.0  # E8: no reflection, no zeroing


def test_friction_zero_and_dt_zero():
    o = oracle.Oracle(rmax=0.04, matrix=np.zeros((1, 1)), friction=0.0, dt=0.0)
    o.set_particles(np.array([[0.5, 0.5]]), np.array([[1.0, 2.0]]), np.zeros(1, np.int32))
    o.update()
    _, v, _, _ = o.get_particles()
    assert v.tolist() == [[1.0, 2.0]]  # pow(0, 0) == 1 (E9)
    o.update(dt=0.02)
    assert o.get_particles()[1].tolist() == [[0.0, 0.0]]


def test_stable_sort_and_end_offsets():
    rng = np.random.default_rng(3)
    pos = rng.random((500, 2))
    o = oracle.Oracle(rmax=0.1, matrix=np.zeros((1, 1)), dt=0.0)
    o.set_particles(pos, None, np.zeros(500, np.int32))
    o.update()
    p, _, _, ids = o.get_particles()
    cells = (p[:, 0] / 0.1).astype(int) + 10 * (p[:, 1] / 0.1).astype(int)
    assert np.all(np.diff(cells) >= 0)
    for c in np.unique(cells):
        assert np.all(np.diff(ids[cells == c].astype(np.int64)) > 0)  # stable: previous order kept
    ends = o.containers()
    assert ends[-1] == 500 and np.array_equal(ends, np.cumsum(np.bincount(cells, minlength=100)))


@pytest.mark.parametrize("wrap", [True, False])
@pytest.mark.parametrize("rmax", [0.04, 0.065, 0.11, 0.3])
@pytest.mark.parametrize("kind", [0, 1, 3, 4, 5])
def test_cell_list_equals_bruteforce(wrap, rmax, kind):
    """The 3x3 cell scan visits exactly the minimum-image neighbour set for nx >= 3."""
    from plife import synth
    n, m = 1200, 5
    pos, vel, types = synth.uniform_state(n, m, 4242)
    vel = (synth.uniform01(7, np.arange(2 * n, dtype=np.uint64)).reshape(n, 2) - 0.5) * 0.1
    M = synth.random_matrix(m, 4242)
    o = oracle.Oracle(rmax=rmax, wrap=wrap, matrix=M, diag=True, accel_kind=kind)
    o.set_particles(pos, vel, types)
    o.update()
    p, v, t, ids = o.get_particles()
    cnt, _, _ = o.neighbor_diag()
    bp, bv, bc = bruteforce.step(pos, vel, types, M, rmax=rmax, wrap=wrap, accel_kind=kind)
    assert np.array_equal(cnt, bc[ids])
    assert np.abs(v - bv[ids]).max() <= 1e-12 * max(1.0, np.abs(bv).max())
    d = np.abs(p - bp[ids])
    assert np.minimum(d, 1 - d).max() <= 1e-14


def test_threaded_equals_serial():
    from plife import synth
    pos, vel, types = synth.uniform_state(20000, 6, 11)
    M = synth.random_matrix(6, 11)
    outs = []
    for T in (1, 3, 12):
        o = oracle.Oracle(rmax=0.02, matrix=M, threads=T)
        o.set_particles(pos, vel, types)
        o.update(); o.update()
        outs.append(o.get_particles())
    for a in outs[1:]:
        for x, y in zip(a, outs[0]):
            assert np.array_equal(x, y)


def test_symmetric_matrix_conserves_momentum():
    from plife import synth
    pos, vel, types = synth.uniform_state(3000, 4, 5)
    M = synth.random_matrix(4, 5)
    M = 0.5 * (M + M.T)
    o = oracle.Oracle(rmax=0.05, matrix=M, friction=1.0)
    o.set_particles(pos, vel, types)
    o.update()
    v = o.get_particles()[1]
    assert np.abs(v.sum(axis=0)).max() <= 1e-15 * 3000
