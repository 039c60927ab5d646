"""Property tests (hypothesis) of the host-side logic: type-count planning, Java number formatting, clipboard matrices,
slab ownership.  CPU only, bounded example counts."""
import numpy as np
from hypothesis import given, settings, strategies as st

from plife import io as pio
from plife import setters as S
from plife import slab


@settings(max_examples=150, deadline=None)
@given(st.integers(1, 6).flatmap(lambda m: st.tuples(st.just(m), st.lists(st.integers(0, m - 1), min_size=0, max_size=60),
                                                      st.lists(st.integers(0, 25), min_size=m, max_size=m))), st.integers(0, 2 ** 31))
def test_plan_type_count_always_meets_the_request(case, seed):
    m, types, want = case
    types = np.array(types, np.int64)
    src, new_types, fresh = S.plan_type_count(types, want, np.random.default_rng(seed))
    assert np.bincount(new_types, minlength=m).tolist() == want
    old = src[src >= 0]
    assert len(set(old.tolist())) == len(old) and len(old) == min(sum(want), len(types))
    if sum(want) == len(types):                                           # retyped in place, no position re-drawn
        assert not fresh.any() and sorted(src.tolist()) == list(range(len(types)))
        assert (new_types != types[src]).sum() == np.maximum(np.bincount(types, minlength=m) - want, 0).sum()
    else:
        assert fresh[src < 0].all()
        assert np.array_equal(new_types[~fresh], types[src[~fresh]])      # whoever is not re-placed keeps its type
        kept = np.bincount(types[src[~fresh]], minlength=m)
        full = np.minimum(np.bincount(types, minlength=m), want)
        assert (kept <= full).all() and full.sum() - kept.sum() <= 1        # the sweep's one unexamined particle


@settings(max_examples=300, deadline=None)
@given(st.floats(allow_nan=False, allow_infinity=False, width=64))
def test_java_double_round_trips_and_has_java_shape(x):
    s = pio.java_double(x)
    assert float(s.replace("E", "e")) == x
    mant = s.lstrip("-").split("E")[0]
    assert "." in mant and "e" not in s and "+" not in s
    a = abs(x)
    assert ("E" in s) == (a != 0 and not (1e-3 <= a < 1e7))
    if "E" in s:
        assert len(mant.split(".")[0]) == 1 and mant[0] != "0"


@settings(max_examples=100, deadline=None)
@given(st.integers(1, 6).flatmap(lambda m: st.lists(st.floats(-1, 1, width=32), min_size=m * m, max_size=m * m)))
def test_clipboard_matrix_round_trip(values):
    m = int(round(len(values) ** 0.5))
    a = np.array(values, np.float64).reshape(m, m)
    back = pio.parse_matrix(pio.matrix_to_string(a))
    assert back.shape == (m, m) and np.abs(back - a).max() <= 5.1e-7          # %f keeps six decimals
    rounded = pio.parse_matrix(pio.matrix_to_string(a, rounded=True))
    assert np.abs(rounded - a).max() <= 0.05 + 1e-6


@settings(max_examples=200, deadline=None)
@given(st.integers(1, 8), st.integers(8, 4000))
def test_slab_rows_partition_the_grid(world, ny):
    rows = [slab.slab_rows(r, world, ny) for r in range(world)]
    assert rows[0][0] == 0 and rows[-1][1] == ny
    assert all(rows[r][1] == rows[r + 1][0] for r in range(world - 1)) and all(hi > lo for lo, hi in rows)
    cy = np.arange(ny)
    own = slab.owner_of_row(cy, world, ny)
    for r, (lo, hi) in enumerate(rows):
        assert (own[lo:hi] == r).all()
