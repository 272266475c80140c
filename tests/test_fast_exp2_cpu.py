"""Accuracy of the FMA-pipe 2^x (time_r1_b200/csrc/fast_exp2.cuh, groundwork for moving the attention softmax off the
MUFU) against numpy's exp2 in float64, through a test-only host build of the same header."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("fe_host") / "fast_exp2_host.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-I", os.path.join(ROOT, "time_r1_b200", "csrc"),
                    os.path.join(ROOT, "tests", "native", "fast_exp2_host.cpp"), "-o", so], check=True)
    return ctypes.CDLL(so)


def run(host, x, deg):
    x = np.ascontiguousarray(x, np.float32)
    out = np.empty_like(x)
    host.fe_exp2_batch(x.ctypes.data_as(ctypes.c_void_p), len(x), deg, out.ctypes.data_as(ctypes.c_void_p))
    return out


@pytest.mark.parametrize("deg,bound", [(3, 8e-5), (4, 3e-6), (5, 3e-7)])
def test_relative_error_over_the_softmax_range(host, deg, bound):
    rng = np.random.default_rng(deg)
    x = np.concatenate([np.linspace(-125.9, 8.0, 400001), rng.uniform(-30, 0, 200000), np.arange(-125, 9) + 0.5, np.arange(-125, 9) - 0.5,
                        np.arange(-125, 9).astype(np.float64)]).astype(np.float32)
    got = run(host, x, deg).astype(np.float64)
    want = np.exp2(x.astype(np.float64))
    assert np.abs(got / want - 1).max() < bound
    assert (got > 0).all()
    # monotone enough for an online softmax: a larger score never gets a visibly smaller weight
    xs = np.sort(x)
    g = run(host, xs, deg).astype(np.float64)
    assert (g[1:] >= g[:-1] * (1 - 4 * bound)).all()


def test_edges(host):
    for deg in (3, 4, 5):
        out = run(host, np.array([-np.inf, -1e30, -127.0, -126.0, 0.0, 1.0, -1.0, 8.0], np.float32), deg)
        assert list(out[:4]) == [0.0, 0.0, 0.0, 0.0]                      # masked / flushed scores are exact zeros
        assert np.allclose(out[4:], [1.0, 2.0, 0.5, 256.0], rtol=1e-4)
