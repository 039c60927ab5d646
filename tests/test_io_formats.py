"""Save-file formats of the reference: TSV / TOML / zip readers and writers (host side, CPU)."""
import io
import zipfile
from types import SimpleNamespace

import numpy as np
import pytest

from plife import io as pio


def test_java_double_to_string():
    cases = {1.0: "1.0", 0.001: "0.001", 1e-4: "1.0E-4", 1e7: "1.0E7", 9999999.0: "9999999.0", 123456789.0: "1.23456789E8",
             0.1 + 0.2: "0.30000000000000004", 100.0: "100.0", -0.85: "-0.85", 0.02: "0.02", 0.0: "0.0",
             1.5e-7: "1.5E-7", 0.5: "0.5", 1234.5678: "1234.5678", 2.5e10: "2.5E10", -3e-12: "-3.0E-12"}
    for x, s in cases.items():
        assert pio.java_double(x) == s
        assert float(s.replace("E", "e")) == x
    rng = np.random.default_rng(1)
    for x in np.concatenate([rng.random(200), rng.normal(size=200) * 1e-5, rng.normal(size=200) * 1e9]):
        assert float(pio.java_double(x)) == x  # shortest digits still round-trip


def test_particles_tsv_round_trip_and_java_written_file():
    rng = np.random.default_rng(2)
    pos, vel, typ = rng.random((50, 2)), rng.normal(size=(50, 2)) * 0.01, rng.integers(0, 6, 50).astype(np.int32)
    buf = io.BytesIO()
    pio.save_particles(buf, pos, vel, typ)
    text = buf.getvalue().decode()
    assert text.splitlines()[0] == "x\ty\tvx\tvy\tcolor" and len(text.splitlines()) == 51
    p2, v2, t2 = pio.load_particles(io.BytesIO(buf.getvalue()))
    assert np.array_equal(p2, pos) and np.array_equal(v2, vel) and np.array_equal(t2, typ)
    # what the Java app writes (Double.toString forms)
    java = "x\ty\tvx\tvy\tcolor\n0.5\t0.25\t1.0E-4\t-0.0\t3\n0.999\t1.0E-5\t0.0\t2.5E-3\t0\n"
    p, v, t = pio.load_particles(io.StringIO(java))
    assert p.tolist() == [[0.5, 0.25], [0.999, 1e-5]] and v.tolist() == [[1e-4, -0.0], [0.0, 2.5e-3]] and t.tolist() == [3, 0]


def test_matrix_tsv_and_physics_toml():
    M = np.array([[1.0, -0.5, 0.25], [0.0, 1e-4, -1.0], [0.3, 0.7, -0.123456789]])
    buf = io.BytesIO()
    pio.save_matrix(buf, M)
    assert buf.getvalue().decode().splitlines()[1] == "0.0\t1.0E-4\t-1.0"
    assert np.array_equal(pio.load_matrix(io.BytesIO(buf.getvalue())), M)
    s = SimpleNamespace(wrap=False, rmax=0.04, friction=0.85, force=2.0)
    buf = io.BytesIO()
    pio.save_physics_toml(buf, s)
    assert 'boundaries = "clamped"' in buf.getvalue().decode() and "radius = 0.04" in buf.getvalue().decode()
    t = pio.load_physics_toml(io.BytesIO(buf.getvalue()), SimpleNamespace(wrap=True, rmax=0.02, friction=0.1, force=1.0))
    assert (t.wrap, t.rmax, t.friction, t.force) == (False, 0.04, 0.85, 2.0)
    with pytest.raises(IOError):
        pio.load_physics_toml(io.StringIO("radius = 0.1\nradiuss = 3\n"), SimpleNamespace())


class FakePhysics:
    """The slice of plife.Physics that load_state / save_state touch."""

    def __init__(self):
        self.settings = SimpleNamespace(wrap=True, rmax=0.02, friction=0.85, force=1.0, matrix=np.zeros((2, 2)))
        self.state = None
        self.ensured = 0

    def set_particles(self, pos, vel, types):
        self.state = (np.asarray(pos), np.asarray(vel), np.asarray(types))

    def ensure_position(self, pos):
        return np.clip(pos, 0, 1)

    def ensure_types(self):
        self.ensured += 1

    @property
    def particles(self):
        return SimpleNamespace(position=self.state[0], velocity=self.state[1], type=self.state[2])


def test_zip_save_load(tmp_path):
    a = FakePhysics()
    rng = np.random.default_rng(3)
    a.set_particles(rng.random((20, 2)), rng.normal(size=(20, 2)), rng.integers(0, 4, 20))
    a.settings.matrix = rng.random((4, 4)) * 2 - 1
    a.settings.rmax, a.settings.wrap = 0.05, False
    path = tmp_path / "save.zip"
    pio.save_state(path, a)
    with zipfile.ZipFile(path) as z:
        assert sorted(z.namelist()) == ["matrix.tsv", "particles.tsv", "physics.toml"]
    b = FakePhysics()
    pio.load_state(path, b)
    assert np.array_equal(b.state[0], a.state[0]) and np.array_equal(b.state[2], a.state[2])
    assert np.array_equal(b.settings.matrix, a.settings.matrix) and b.ensured == 1
    assert (b.settings.rmax, b.settings.wrap) == (0.05, False)


def test_clipboard_matrix_text():
    """A/MatrixParser.java:28-74: parse through float, tolerate junk, floor(sqrt(count)) size; %f and %4.1f writers."""
    m = pio.parse_matrix("0.1 0.2 -0.3\n-0.1 0.4 0.1\n1.0 -1.0 0.0")
    assert m.shape == (3, 3) and m[0, 0] == float(np.float32(0.1)) and m[0, 0] != 0.1 and m[2].tolist() == [1.0, -1.0, 0.0]
    assert pio.parse_matrix("") is None and pio.parse_matrix("a b\tc") is None
    assert pio.parse_matrix("1 2 3 4 5 6 7 8").tolist() == [[1, 2], [3, 4]]          # surplus dropped
    assert pio.parse_matrix("x 1\t\t2\n\n3 y 4").tolist() == [[1, 2], [3, 4]]         # empty / junk tokens skipped
    # Float.parseFloat accepts suffixes, hex floats, NaN/Infinity; rejects what only Python accepts
    got = pio.parse_matrix("1f 2d .5 1e1 0x1p1 inf 1_0 nan -Infinity")
    assert got.shape == (2, 2) and got.tolist() == [[1.0, 2.0], [0.5, 10.0]]
    got = pio.parse_matrix("0x1p1 -Infinity 3. +4")
    assert got.tolist() == [[2.0, float("-inf")], [3.0, 4.0]]
    assert pio.matrix_to_string([[0.1, -1 / 3], [1e-7, 1.0]]) == "0.100000\t-0.333333\n0.000000\t1.000000\n"
    # HALF_UP on the shortest decimal digits (JDK Formatter), width 4
    assert pio.matrix_to_string([[0.25, -0.35], [1, -0.04]], rounded=True) == " 0.3\t-0.4\n 1.0\t-0.0\n"
    rt = pio.parse_matrix(pio.matrix_to_string(m))
    assert np.abs(rt - m).max() <= 5e-7
