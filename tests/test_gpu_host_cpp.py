"""The C++ host mirror (mfem-ad_b200/host/madb.hpp: reference class names above the C ABI) gives the same
numbers as the ctypes path: runs examples/ex_assemble (built by __graft_entry__.build())."""
import os
import re
import subprocess

import numpy as np
import pytest

import spec as S
from mfem_ad_b200 import meshgen as G
from oracle import oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _nums(line):
    return [float(v) for v in re.findall(r"[-+]?\d+\.?\d*(?:[eE][-+]?\d+)?", line)]


def _kv(line):
    return {k: float(v) for k, v in re.findall(r"(\w+)=([-+0-9.eE]+)", line)}


def test_cpp_examples_match(ctx):
    exe = os.path.join(ROOT, "examples", "ex_assemble")
    assert os.path.exists(exe), "run python __graft_entry__.py first"
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.splitlines()
    # ex0 known answers (ex0.cpp:36-61)
    assert abs(_nums(out[0])[1] - 0.30321372968699545) <= 1e-14
    assert np.max(np.abs(np.array(_nums(out[1])[1:]) - [2.3855167309591354, 1.3032137296869954, 3.0])) <= 1e-14
    assert np.max(np.abs(np.array(_nums(out[2])[1:]) - [-1.3032137296869954, 2.3855167309591354, 1.3032137296869954, -6.0])) <= 1e-14
    # ex2 twice with eps halved
    mesh = G.cartesian_mesh((16, 12))
    s = G.h1_space(mesh, 2, mode=O.GRAD)
    x = np.sin(0.37 * np.arange(s["ndofs"])) * 0.3
    for k, eps in enumerate((0.5, 0.25)):
        of = O.OracleForm(mesh, [s], S.minsurf(2, eps).oracle())
        kv = _kv(out[3 + k])
        y, v = of.mult(x), of.grad(x)[2]
        assert abs(kv["eps"] - eps) < 1e-12
        assert abs(kv["energy"] - of.energy(x)) <= 1e-12 * abs(of.energy(x))
        assert abs(kv["y2"] - y @ y) <= 1e-12 * (y @ y)
        assert abs(kv["K2"] - v @ v) <= 1e-12 * (v @ v)
        assert int(kv["nnz"]) == v.size
    # ex4 block
    order = 2
    mesh = G.cartesian_mesh((6, 5))
    h1 = G.h1_space(mesh, order + 1, mode=O.VALUE | O.GRAD)
    l2 = G.l2_space(mesh, order - 1, mode=O.VALUE)
    n = h1["ndofs"] + l2["ndofs"]
    x = np.cos(0.11 * np.arange(n)) * 0.4
    psik = np.sin(0.23 * np.arange(l2["ndofs"]))
    of = O.OracleForm(mesh, [h1, l2], S.pg(S.obstacle(2), S.fermidirac(0.0, 0.5), 0.8).oracle(), quad_order=9,
                      params=[dict(type=O.PRM_GF, size=1, data=psik, space=l2)])
    kv = _kv(out[5])
    y, v = of.mult(x), of.grad(x)[2]
    assert abs(kv["alpha"] - 0.8) < 1e-12
    assert abs(kv["y2"] - y @ y) <= 1e-12 * (y @ y)
    assert abs(kv["K2"] - v @ v) <= 1e-12 * (v @ v)
    # ex1-type: load vector of 2 pi^2 sin sin (sum = int f = 8 up to the rule) and the Poisson solve on the device:
    # the Q2 solution on a 12 x 12 mesh is within 1e-3 of sin(pi x) sin(pi y) at the nodes
    kv = _kv(out[6])
    assert abs(kv["bsum"] - 8.0) <= 1e-3 and 0 < kv["iters"] < 2000 and kv["relres"] <= 1e-11 and kv["maxerr"] <= 1e-3
