"""On-disk formats of the reference's saves, host side.

A save is a zip with `particles.tsv`, `physics.toml`, `matrix.tsv` (and an `img.png` that is ignored):
A/Main.java:1258-1348.  These readers/writers let the native backend load real saved states (clustered,
non-uniform) and write saves the Java app can open.

  particles.tsv  header `x\\ty\\tvx\\tvy\\tcolor`, one particle per line (A/io/ParticlesIO.java:8-46);
                 z and the third velocity component are dropped on save and zero on load
  matrix.tsv     one matrix row per line, tab separated (A/io/MatrixIO.java:10-42)
  physics.toml   boundaries = "periodic" | "clamped", radius, friction, force
                 (A/PhysicsSettingsToml.java:9-34); dt is NOT saved

  clipboard      matrix text copied / pasted by the GUI (A/MatrixParser.java:28-74): parsed through **float**,
                 written with `%f` (6 decimals) or `%4.1f`

Numbers are written like Java's Double.toString (shortest digits that round-trip; decimal notation for
1e-3 <= |x| < 1e7, otherwise `d.dddE[-]n`), so files are textually what the app would write with JDK >= 19.
"""
from __future__ import annotations

import io
import re
import zipfile
from decimal import ROUND_HALF_UP, Decimal
from typing import Optional, Tuple

import numpy as np

PARTICLES_HEADER = "x\ty\tvx\tvy\tcolor"


def java_double(x: float) -> str:
    """Double.toString(x)."""
    x = float(x)
    if x != x:
        return "NaN"
    if x in (float("inf"), float("-inf")):
        return "Infinity" if x > 0 else "-Infinity"
    if x == 0.0:
        return "-0.0" if str(x).startswith("-") else "0.0"
    sign = "-" if x < 0 else ""
    a = abs(x)
    # shortest round-trip digits via repr
    r = repr(a)
    if "e" in r or "E" in r:
        m, e = r.lower().split("e")
        digits = m.replace(".", "")
        point = len(m.split(".")[0])
        exp10 = int(e) + point - 1  # exponent of the first digit
    else:
        ip, _, fp = r.partition(".")
        if ip.strip("0") == "":
            stripped = fp.lstrip("0")
            exp10 = -(len(fp) - len(stripped)) - 1
            digits = stripped
        else:
            digits = ip + fp
            exp10 = len(ip) - 1
    digits = digits.rstrip("0") or "0"
    if 1e-3 <= a < 1e7:
        if exp10 >= 0:
            ip = digits[: exp10 + 1].ljust(exp10 + 1, "0")
            fp = digits[exp10 + 1:] or "0"
        else:
            ip = "0"
            fp = "0" * (-exp10 - 1) + digits
        return f"{sign}{ip}.{fp}"
    return f"{sign}{digits[0]}.{digits[1:] or '0'}E{exp10}"


# -- particles.tsv ------------------------------------------------------------

def save_particles(stream, position, velocity, types):
    """A/io/ParticlesIO.java:30-46"""
    position = np.asarray(position, np.float64).reshape(-1, 2)
    velocity = np.asarray(velocity, np.float64).reshape(-1, 2)
    types = np.asarray(types).reshape(-1)
    lines = [PARTICLES_HEADER]
    for (x, y), (vx, vy), t in zip(position, velocity, types):
        lines.append(f"{java_double(x)}\t{java_double(y)}\t{java_double(vx)}\t{java_double(vy)}\t{int(t)}")
    stream.write(("\n".join(lines) + "\n").encode())


def load_particles(stream) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """A/io/ParticlesIO.java:10-28 (the header line is skipped, whatever it says)."""
    text = stream.read()
    if isinstance(text, bytes):
        text = text.decode()
    rows = [ln.split("\t") for ln in text.splitlines()[1:] if ln.strip()]
    n = len(rows)
    pos = np.empty((n, 2), np.float64)
    vel = np.empty((n, 2), np.float64)
    types = np.empty(n, np.int32)
    for i, p in enumerate(rows):
        pos[i] = (float(p[0]), float(p[1]))
        vel[i] = (float(p[2]), float(p[3]))
        types[i] = int(p[4])
    return pos, vel, types


# -- matrix.tsv ---------------------------------------------------------------

def save_matrix(stream, matrix):
    """A/io/MatrixIO.java:27-42"""
    m = np.asarray(matrix, np.float64)
    stream.write(("".join("\t".join(java_double(v) for v in row) + "\n" for row in m)).encode())


def load_matrix(stream) -> np.ndarray:
    """A/io/MatrixIO.java:12-25 (square: the row count decides the size)."""
    text = stream.read()
    if isinstance(text, bytes):
        text = text.decode()
    rows = [[float(v) for v in ln.split("\t")] for ln in text.splitlines() if ln.strip()]
    n = len(rows)
    return np.array([r[:n] for r in rows], np.float64).reshape(n, n)


# -- clipboard matrix text ----------------------------------------------------

_JAVA_FLOAT = re.compile(r"[+-]?(NaN|Infinity|(\d+\.?\d*|\.\d+)([eE][+-]?\d+)?[fFdD]?"
                         r"|0[xX]([0-9a-fA-F]+\.?[0-9a-fA-F]*|\.[0-9a-fA-F]+)[pP][+-]?\d+[fFdD]?)")


def _java_parse_float(token: str):
    """Float.parseFloat: None where Java throws NumberFormatException (Python's float() is laxer: '1_0', 'inf')."""
    t = token.strip()
    if not _JAVA_FLOAT.fullmatch(t):
        return None
    low = t.lower().lstrip("+-")
    if low.startswith("0x"):
        v = float.fromhex(t[:-1] if t[-1] in "fFdD" else t)   # after the p-exponent a trailing f/d is a suffix
    elif low == "nan":
        v = float("nan")
    elif low == "infinity":
        v = float("-inf") if t.startswith("-") else float("inf")
    else:
        v = float(t.rstrip("fFdD"))
    return float(np.float32(v))


def parse_matrix(text: str) -> Optional[np.ndarray]:
    """MatrixParser.parseMatrix (A/MatrixParser.java:28-52): split on single whitespace characters, keep what
    parses as a float, size = floor(sqrt(count)), surplus numbers dropped; None when nothing parses.  Values pass
    through float, so a pasted 0.1 becomes 0.10000000149011612."""
    numbers = [v for v in (_java_parse_float(p) for p in re.split(r"\s", text)) if v is not None]
    size = int(np.sqrt(len(numbers)))
    if size < 1:
        return None
    return np.array(numbers[: size * size], np.float64).reshape(size, size)


def _java_format_f(x: float, decimals: int, width: int = 0) -> str:
    """String.format(Locale.US, "%<width>.<decimals>f"): the JDK rounds HALF_UP on the decimal digits of
    Double.toString, not on the binary value (0.25 -> "0.3", 0.35 -> "0.4")."""
    x = float(x)
    if x != x:
        out = "NaN"
    elif x in (float("inf"), float("-inf")):
        out = "Infinity" if x > 0 else "-Infinity"
    else:
        d = Decimal(repr(x)).quantize(Decimal(1).scaleb(-decimals), rounding=ROUND_HALF_UP)
        out = f"{d:.{decimals}f}"
        if d == 0 and str(x).startswith("-"):
            out = "-" + out.lstrip("-")
    return out.rjust(width)


def matrix_to_string(matrix, rounded: bool = False) -> str:
    """MatrixParser.matrixToString / matrixToStringRoundAndFormat (A/MatrixParser.java:54-74)."""
    m = np.asarray(matrix, np.float64)
    enc = (lambda v: _java_format_f(v, 1, 4)) if rounded else (lambda v: _java_format_f(v, 6))
    return "".join("\t".join(enc(v) for v in row) + "\n" for row in m)


# -- physics.toml -------------------------------------------------------------

def save_physics_toml(stream, settings):
    """A/PhysicsSettingsToml.java:19-25 (toml4j writes a flat key = value table; key order is not significant)."""
    s = settings
    stream.write((f'boundaries = "{"periodic" if s.wrap else "clamped"}"\n'
                  f"radius = {java_double(s.rmax)}\nfriction = {java_double(s.friction)}\nforce = {java_double(s.force)}\n").encode())


def load_physics_toml(stream, settings):
    """A/PhysicsSettingsToml.java:27-33; unknown keys raise, like the reference's TomlFile (toml_util/TomlFile.java:52-106)."""
    text = stream.read()
    if isinstance(text, bytes):
        text = text.decode()
    allowed = {"boundaries", "radius", "friction", "force"}
    for ln in text.splitlines():
        ln = ln.split("#", 1)[0].strip()
        if not ln:
            continue
        key, _, val = ln.partition("=")
        key, val = key.strip(), val.strip()
        if key not in allowed:
            raise IOError(f"Unknown key '{key}' in physics.toml")
        if key == "boundaries":
            settings.wrap = val.strip('"\'') == "periodic"
        elif key == "radius":
            settings.rmax = float(val)
        elif key == "friction":
            settings.friction = float(val)
        else:
            settings.force = float(val)
    return settings


# -- zip container ------------------------------------------------------------

def save_state(path, physics):
    """Main.saveState (A/Main.java:1258-1298) for a `plife.Physics`; particles are written in the current
    (cell-sorted) order, exactly what the reference's array holds after a step."""
    p = physics.particles
    with zipfile.ZipFile(path, "w", zipfile.ZIP_DEFLATED) as z:
        for name, writer in (("particles.tsv", lambda b: save_particles(b, p.position, p.velocity, p.type)),
                             ("physics.toml", lambda b: save_physics_toml(b, physics.settings)),
                             ("matrix.tsv", lambda b: save_matrix(b, physics.settings.matrix))):
            buf = io.BytesIO()
            writer(buf)
            z.writestr(name, buf.getvalue())


def load_state(path, physics):
    """Main.loadState (A/Main.java:1314-1348): entries are applied in file order; a missing entry keeps the current
    state; after a matrix load `ensureTypes()` runs (B/Physics.java:266-272)."""
    with zipfile.ZipFile(path) as z:
        for info in z.infolist():
            with z.open(info) as f:
                if info.filename == "particles.tsv":
                    pos, vel, types = load_particles(f)
                    m = int(types.max()) + 1 if len(types) else 1
                    if m > physics.settings.matrix.shape[0]:
                        # the reference would crash on the next update(); grow the matrix with zeros until matrix.tsv arrives
                        grown = np.zeros((m, m))
                        k = physics.settings.matrix.shape[0]
                        grown[:k, :k] = physics.settings.matrix
                        physics.settings.matrix = grown
                    physics.set_particles(physics.ensure_position(pos), vel, types)
                elif info.filename == "physics.toml":
                    load_physics_toml(f, physics.settings)
                elif info.filename == "matrix.tsv":
                    physics.settings.matrix = load_matrix(f)
                    physics.ensure_types()
                elif info.filename == "img.png":
                    pass
                else:
                    print("Unknown file in ZIP:", info.filename)
    return physics
