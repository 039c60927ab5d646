"""examples/plife_headless.c: the C ABI used from plain C (no Python, no torch, no CUDA headers)."""
import os
import re
import subprocess

import numpy as np
import pytest

import plife

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "examples", "plife_headless")


def _fnv1a(chunks):
    h = 0xCBF29CE484222325
    for c in chunks:
        for b in c.tobytes():
            h = ((h ^ b) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h


def test_c_client_builds_and_fails_loudly_without_a_gpu(native_lib):
    assert os.path.exists(BIN), "build() compiles the C client next to the library"
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the failure path is covered on the CPU box")
    r = subprocess.run([BIN, "1000", "3", "0.1", "2"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 2 and "plife_create failed" in r.stderr and r.stdout == ""   # no CPU fallback


@pytest.mark.gpu
@pytest.mark.parametrize("prec", [32, 64])
def test_c_client_matches_the_python_host(native_lib, prec):
    n, m, rmax, steps, seed = 6000, 5, 0.05, 7, 0x5EED0042
    r = subprocess.run([BIN, str(n), str(m), str(rmax), str(steps), str(prec), hex(seed)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    got = dict(kv.split("=") for kv in r.stdout.split())
    p = plife.NativePhysics(precision=plife.F64 if prec == 64 else plife.F32)
    p.set_settings(rmax, 0.85, 1.0, True)
    p.random_matrix(m, seed)
    p.init_uniform(n, seed)
    p.step(0.02, 1)
    p.step(0.02, steps - 1)
    q = p.download()
    st = p.step_stats()
    want = _fnv1a([np.ascontiguousarray(q.position, np.float64), np.ascontiguousarray(q.velocity, np.float64),
                   np.ascontiguousarray(q.type, np.int32), np.ascontiguousarray(q.id, np.uint32)])
    assert int(got["n"]) == n and int(got["nx"]) == st["nx"] and int(got["pair_evals"]) == st["pair_evals"]
    assert int(got["checksum"], 16) == want
    assert re.fullmatch(r"\d+\.\d+", got["ms_per_step"])
