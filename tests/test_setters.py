"""Catalogs of position setters, type setters and matrix generators and the type-count
editor (A/ExtendedPhysics.java).  The reference draws from Math.random(), so the checks are on supports, moments
and the deterministic structure, not on streams."""
import numpy as np
import pytest

from plife import setters as S

N = 200_000


def rng():
    return np.random.default_rng(7)


def test_catalog_names_match_the_reference_lists():
    assert list(S.POSITION_SETTERS) == ["centered", "uniform", "uniform circle", "centered circle", "ring", "rainbow ring",
                                        "color battle", "color wheel", "line", "spiral", "rainbow spiral"]
    assert list(S.TYPE_SETTERS) == ["random", "randomize 10%", "slices", "onion", "rotate", "flip", "more of first", "kill still"]
    assert list(S.MATRIX_GENERATORS) == ["random", "symmetry", "chains", "chains 2", "chains 3", "snakes", "zero"]


@pytest.mark.parametrize("name", list(S.POSITION_SETTERS))
def test_position_setters_shape_and_empty(name):
    s = S.POSITION_SETTERS[name]
    t = rng().integers(0, 5, 1000)
    p = s.set(t, 5, rng())
    assert p.shape == (1000, 2) and p.dtype == np.float64 and np.isfinite(p).all()
    assert s.set(np.zeros(0, np.int32), 5, rng()).shape == (0, 2)


def _centred(p):
    """back to the setters' [-1, 1]^2 working frame"""
    return (p - 0.5) * 2.0


def test_position_setter_distributions():
    t = rng().integers(0, 6, N)
    g = lambda name: _centred(S.POSITION_SETTERS[name].set(t, 6, rng()))
    c = g("centered")
    assert abs(c.mean()) < 5e-3 and abs(c.std() - 0.3) < 5e-3
    u = S.POSITION_SETTERS["uniform"].set(t, 6, rng())
    assert u.min() >= 0 and u.max() < 1 and abs(u.mean() - 0.5) < 5e-3
    r = np.hypot(*g("uniform circle").T)
    assert r.max() <= 0.5 + 1e-12 and abs((r < 0.25).mean() - 0.25) < 5e-3      # area-uniform: P(r < R/2) = 1/4
    r = np.hypot(*g("centered circle").T)
    assert r.max() <= 0.5 + 1e-12 and abs((r < 0.25).mean() - 0.5) < 5e-3       # radius-uniform: P(r < R/2) = 1/2
    r = np.hypot(*g("ring").T)
    assert abs(r.mean() - 0.7) < 1e-3 and abs(r.std() - 0.02) < 1e-3
    ln = g("line")
    assert np.abs(ln[:, 0]).max() <= 1 and np.abs(ln[:, 1]).max() <= 0.15 + 1e-7
    sp = np.hypot(*g("spiral").T)
    assert sp.max() < 0.9 + 0.1 * 0.1 * 6 and abs(np.median(sp) - 0.45) < 0.01   # radius ~ 0.9 f, f uniform


def test_type_dependent_position_setters():
    m = 6
    t = rng().integers(0, m, N)
    for name, centre_r, blob in (("color battle", 0.5, 0.1), ("color wheel", 0.3, None)):
        p = _centred(S.POSITION_SETTERS[name].set(t, m, rng()))
        for k in range(m):
            c = p[t == k].mean(axis=0)
            want = centre_r * np.array([np.cos(k / m * S.TAU), np.sin(k / m * S.TAU)])
            assert np.abs(c - want).max() < 5e-3
            if blob:
                assert np.hypot(*(p[t == k] - want).T).max() <= blob + 1e-7
    p = _centred(S.POSITION_SETTERS["rainbow ring"].set(t, m, rng()))
    ang = np.arctan2(p[:, 1], p[:, 0])
    for k in range(m):
        d = np.angle(np.exp(1j * (ang[t == k] - k / m * S.TAU)))
        assert abs(d.mean()) < 0.01 and abs(d.std() - 0.3 / m * S.TAU) < 0.01
    p = _centred(S.POSITION_SETTERS["rainbow spiral"].set(t, m, rng()))
    r = np.hypot(*p.T)
    means = [r[t == k].mean() for k in range(m)]
    assert all(np.diff(means) > 0)                                              # ordered outwards by type
    assert abs(means[2] - 0.9 * 3 / 8) < 0.01


def test_map_type_clamps_like_the_reference():
    assert S.map_type([-0.2, 0.0, 0.49, 0.5, 0.999, 1.0, 7.0], 2).tolist() == [0, 0, 0, 1, 1, 1, 1]


def test_type_setters():
    m = 5
    r = rng()
    pos = r.random((N, 2))
    vel = r.normal(size=(N, 2)) * 0.01
    t = r.integers(0, m, N).astype(np.int32)
    T = S.TYPE_SETTERS
    out = T["random"].get_type(pos, vel, t, m, rng())
    assert out.dtype == np.int32 and out.min() == 0 and out.max() == m - 1
    assert np.abs(np.bincount(out) / N - 1 / m).max() < 5e-3
    out = T["randomize 10%"].get_type(pos, vel, t, m, rng())
    assert abs((out != t).mean() - 0.1 * (1 - 1 / m)) < 3e-3
    assert np.array_equal(T["slices"].get_type(pos, vel, t, m, rng()), np.floor(pos[:, 0] * m).astype(np.int32))
    onion = T["onion"].get_type(pos, vel, t, m, rng())
    rad = np.hypot(pos[:, 0] - 0.5, pos[:, 1] - 0.5) * 2
    assert np.array_equal(onion, np.minimum(np.floor(rad * m), m - 1).astype(np.int32))
    assert (onion[rad >= 1] == m - 1).all()
    assert np.array_equal(T["rotate"].get_type(pos, vel, t, m, rng()), (t + 1) % m)
    assert np.array_equal(T["flip"].get_type(pos, vel, t, m, rng()), m - 1 - t)
    first = np.bincount(T["more of first"].get_type(pos, vel, t, m, rng()), minlength=m) / N
    assert all(np.diff(first) < 0)
    # P(UV < 1/m) = (1 + ln m) / m
    assert abs(first[0] - (1 + np.log(m)) / m) < 5e-3
    still = np.hypot(*vel.T) < 0.01
    out = T["kill still"].get_type(pos, vel, t, m, rng())
    assert (out[still] == m - 1).all() and np.array_equal(out[~still], t[~still])
    # 3-component inputs (the reference's Vector3d) are accepted
    out3 = T["onion"].get_type(np.c_[pos, np.zeros(N)], np.c_[vel, np.zeros(N)], t, m, rng())
    assert np.array_equal(out3, onion)


def test_matrix_generators():
    G = S.MATRIX_GENERATORS
    m = G["random"].make_matrix(7, rng())
    assert m.shape == (7, 7) and m.min() >= -1 and m.max() < 1 and not np.allclose(m, m.T)
    base = 2.0 * rng().random((7, 7)) - 1.0
    s = G["symmetry"].make_matrix(7, rng())
    assert np.array_equal(s, s.T) and np.array_equal(np.tril(s), np.tril(base))     # upper takes lower, (A/MatrixGeneratorProvider.java:20-24)
    assert G["chains"].make_matrix(5, rng()).tolist() == [[1, 1, -1, -1, 1], [1, 1, 1, -1, -1], [-1, 1, 1, 1, -1], [-1, -1, 1, 1, 1], [1, -1, -1, 1, 1]]
    assert G["chains 2"].make_matrix(4, rng()).tolist() == [[1, .2, -1, .2], [.2, 1, .2, -1], [-1, .2, 1, .2], [.2, -1, .2, 1]]
    assert G["chains 3"].make_matrix(4, rng()).tolist() == [[1, .2, 0, .2], [.2, 1, .2, 0], [0, .2, 1, .2], [.2, 0, .2, 1]]
    assert G["snakes"].make_matrix(3, rng()).tolist() == [[1, .2, 0], [0, 1, .2], [.2, 0, 1]]
    assert not G["zero"].make_matrix(6, rng()).any()
    # degenerate sizes: every branch of the reference's if-chains
    assert G["chains"].make_matrix(1, rng()).tolist() == [[1]] and G["chains"].make_matrix(2, rng()).tolist() == [[1, 1], [1, 1]]
    assert G["chains 2"].make_matrix(2, rng()).tolist() == [[1, .2], [.2, 1]] and G["chains 3"].make_matrix(1, rng()).tolist() == [[1]]
    assert G["snakes"].make_matrix(1, rng()).tolist() == [[.2]] and G["snakes"].make_matrix(2, rng()).tolist() == [[1, .2], [.2, 1]]
    for g in G.values():
        assert g.make_matrix(0, rng()).shape == (0, 0)


# -- type counts --------------------------------------------------------------

def _serial_set_type_count(types, want, order):
    """The reference's sweep (A/ExtendedPhysics.java:55-96), run literally on a given shuffle: returns the multiset
    of kept original indices and the final type histogram."""
    idx = list(order)
    have = [0] * len(want)
    i, j = 0, len(idx) - 1
    while i < j:
        t = types[idx[i]]
        if have[t] < want[t]:
            have[t] += 1
            i += 1
        else:
            idx[i], idx[j] = idx[j], idx[i]
            j -= 1
    return idx[:i], have


def test_rank_within_type():
    t = np.array([2, 0, 2, 2, 1, 0])
    assert S._rank_within_type(t).tolist() == [0, 0, 1, 2, 0, 1]
    assert S._rank_within_type(np.zeros(0, np.int64)).tolist() == []


def test_equal_type_count():
    assert S.equal_type_count(10, 3).tolist() == [4, 4, 2] and S.equal_type_count(12, 4).tolist() == [3, 3, 3, 3]
    assert S.equal_type_count(10, 1) is None
    assert S.equal_type_count(2, 4).tolist() == [1, 1, 1, -1]     # the reference's formula, negative remainder and all


@pytest.mark.parametrize("want", [[40, 40, 40, 40], [10, 0, 5, 3], [100, 100, 50, 1], [0, 0, 0, 0], [0, 0, 0, 500]])
def test_plan_type_count_resizing(want):
    r = rng()
    types = r.integers(0, 4, 120)
    src, new_types, fresh = S.plan_type_count(types, want, r)
    n_new = sum(want)
    assert len(src) == len(new_types) == len(fresh) == n_new
    assert np.bincount(new_types, minlength=4).tolist() == want
    old = src[src >= 0]
    assert len(np.unique(old)) == len(old) and len(old) == min(n_new, 120)        # nothing duplicated, as much reused as fits
    kept = ~fresh
    assert np.array_equal(new_types[kept], types[src[kept]])                      # kept particles keep their type
    assert (src[fresh & (src >= 0)] >= 0).all() and fresh[src < 0].all()          # new particles always get a position
    # as many kept per type as the literal sweep keeps on some shuffle (it depends on the shuffle only through
    # the type of the one particle the sweep never examines)
    counts = np.bincount(types[src[kept]], minlength=4)
    full = np.minimum(np.bincount(types, minlength=4), want)
    assert (counts <= full).all() and full.sum() - counts.sum() <= 1


def test_plan_type_count_matches_the_serial_sweep_in_distribution():
    types = np.array([0, 0, 0, 1, 1, 2])
    want = [2, 1, 2, 0]                       # 5 != 6: the resizing branch
    trials = 4000
    r = rng()
    a = np.zeros(6)
    b = np.zeros(6)
    for _ in range(trials):
        src, _, fresh = S.plan_type_count(types, want, r)
        a[src[~fresh]] += 1
        kept, _ = _serial_set_type_count(types, want, r.permutation(6))
        b[kept] += 1
    assert np.abs(a - b).max() / trials < 0.04                                   # per-particle keep probability agrees


def test_plan_type_count_same_total_retypes_in_place():
    r = rng()
    types = r.integers(0, 3, 90)
    have = np.bincount(types, minlength=3)
    want = have + np.array([7, -3, -4])
    src, new_types, fresh = S.plan_type_count(types, want, r)
    assert sorted(src.tolist()) == list(range(90)) and not fresh.any()
    assert np.bincount(new_types, minlength=3).tolist() == want.tolist()
    changed = new_types != types[src]
    assert changed.sum() == 7 and set(types[src][changed]) == {1, 2} and set(new_types[changed]) == {0}


def test_plan_type_count_rejects_bad_input():
    with pytest.raises(ValueError):
        S.plan_type_count(np.array([0, 3]), [1, 1], rng())
    with pytest.raises(ValueError):
        S.plan_type_count(np.array([0, 1]), [3, -1], rng())
