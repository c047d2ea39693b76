"""Generates the committed fixtures of tests/golden/ (run from the repo root: python tests/golden/make_golden.py).

ex0_known_answers.json -- the reference's OWN known answers: the closed forms its driver checks the AD results
    against (ex0.cpp:20 f = sin(x0) e^x1 + x2^3 with the Jacobian / Hessian of ex0.cpp:36-61, and
    ex0.cpp:23-35 F = (sin(x0 x1), cos(x0 x1 x2)) with ex0.cpp:62-98), evaluated here in plain numpy at the
    driver's point (ex0.cpp:102) and a few more.  Independent of the oracle and of the CUDA code.
assembly_*.npz -- small whole-mesh assemblies produced by the CPU oracle (oracle/oracle.cpp).  The reference cannot be
    built in this image (MFEM, MPI, hypre, SuiteSparse are absent: DESIGN.md section 2), so these are NOT outputs of
    the reference itself: they freeze the oracle (regression fixture for both sides), they do not pin it.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def ex0_scalar(x):
    s, c, e = np.sin(x[0]), np.cos(x[0]), np.exp(x[1])
    v = s * e + x[2] ** 3
    g = [c * e, s * e, 3 * x[2] ** 2]
    H = [[-s * e, c * e, 0.0], [c * e, s * e, 0.0], [0.0, 0.0, 6 * x[2]]]
    return float(v), [float(t) for t in g], [[float(t) for t in r] for r in H]


def ex0_vector(x):
    a, b, c = x
    p, q = a * b, a * b * c
    F = [np.sin(p), np.cos(q)]
    J = [[b * np.cos(p), a * np.cos(p), 0.0], [-b * c * np.sin(q), -a * c * np.sin(q), -a * b * np.sin(q)]]
    H0 = [[-b * b * np.sin(p), np.cos(p) - a * b * np.sin(p), 0.0],
          [np.cos(p) - a * b * np.sin(p), -a * a * np.sin(p), 0.0], [0.0, 0.0, 0.0]]
    dq = [b * c, a * c, a * b]          # gradient of q
    d2q = [[0.0, c, b], [c, 0.0, a], [b, a, 0.0]]
    H1 = [[-np.cos(q) * dq[i] * dq[j] - np.sin(q) * d2q[i][j] for j in range(3)] for i in range(3)]
    tofl = lambda M: [[float(t) for t in r] for r in M]
    return [float(t) for t in F], tofl(J), [tofl(H0), tofl(H1)]


def main():
    pts = [[0.5, 1.0, -1.0], [0.0, 0.0, 0.0], [-1.3, 0.25, 2.0], [2.0, -0.7, 0.4]]
    out = {"source": "closed forms of ex0.cpp:36-98 evaluated in numpy (tests/golden/make_golden.py)", "points": []}
    for x in pts:
        v, g, H = ex0_scalar(np.array(x))
        F, J, HH = ex0_vector(np.array(x))
        out["points"].append({"x": x, "f": v, "grad": g, "hess": H, "F": F, "jac": J, "Hess": HH})
    with open(os.path.join(HERE, "ex0_known_answers.json"), "w") as fh:
        json.dump(out, fh, indent=1)

    from mfem_ad_b200 import meshgen as G
    from oracle import oracle as O
    import spec as S
    cases = {
        "assembly_minsurf_q2": dict(n=(6, 5), perturb=0.15, order=2, fs=S.minsurf(2, 0.5), seed=11),
        "assembly_diffusion_q1": dict(n=(7, 4), perturb=0.2, order=1, fs=S.diffusion(2), seed=12),
    }
    for name, c in cases.items():
        mesh = G.cartesian_mesh(c["n"], perturb=c["perturb"])
        sp = G.h1_space(mesh, c["order"], mode=O.GRAD)
        of = O.OracleForm(mesh, [sp], c["fs"].oracle())
        xc = G.dof_coords(mesh, sp)
        x = np.sin(np.pi * xc[:, 0]) * np.sin(np.pi * xc[:, 1]) + 0.1 * np.random.default_rng(c["seed"]).uniform(-1, 1, sp["ndofs"])
        rp, ci, vals = of.grad(x)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), n=np.array(c["n"]), perturb=c["perturb"], order=c["order"],
                            x=x, y=of.mult(x), energy=of.energy(x), rowptr=rp, colidx=ci, vals=vals)
        print(name, "ndofs", sp["ndofs"], "nnz", vals.size)


if __name__ == "__main__":
    main()
