"""The minimax sign-polynomial table compiled into csrc/dense_proj.cu must be the one
scripts/sign_poly_table.py generates, and must have the two properties the kernel relies on
(p <= 1 on [0, 1]; p(x) >= x below the design interval).  CPU only."""
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))


def _compiled_table():
    src = open(os.path.join(ROOT, "cuadmm_b200", "csrc", "dense_proj.cu")).read()
    body = src[src.index("kSignPoly[kSignTable][3] = {"):]
    body = body[:body.index("};")]
    rows = re.findall(r"\{\s*([-0-9.e+]+),\s*([-0-9.e+]+),\s*([-0-9.e+]+)\s*\}", body)
    return np.array(rows, dtype=np.float64)


def test_compiled_table_matches_generator():
    import sign_poly_table as spt
    tab = np.array([t[:3] for t in spt.table()])
    ctab = _compiled_table()
    assert ctab.shape == tab.shape == (7, 3)
    assert np.allclose(ctab, tab, rtol=1e-9, atol=0)


def test_table_properties_and_composite_convergence():
    import sign_poly_table as spt
    ctab = _compiled_table()
    lows = [t[3] for t in spt.table()]
    xs = np.concatenate([np.linspace(0, 1, 400001), np.geomspace(1e-12, 1, 400001)])
    for (a, b, c), l in zip(ctab, lows):
        p = a * xs + b * xs ** 3 + c * xs ** 5
        assert p.max() <= 1 + 1e-9
        assert (p - xs)[xs <= l].min() >= -1e-15
    # the composite maps [1e-4, 1] into [0.94, 1]; three Newton-Schulz-5 steps finish the job
    y = np.geomspace(1e-4, 1, 200001)
    for a, b, c in ctab:
        y = a * y + b * y ** 3 + c * y ** 5
    assert y.min() > 0.94 and y.max() <= 1 + 1e-9
    for _ in range(3):
        y = (15 * y - 10 * y ** 3 + 3 * y ** 5) / 8
    assert np.abs(y - 1).max() < 1e-15


def test_numpy_emulation_of_the_iteration_projects_correctly():
    # the schedule of dense_part_project (scale from ||A0^2||_F, table, NS5, residual freeze) in numpy
    ctab = _compiled_table()
    rng = np.random.default_rng(0)
    n = 120
    G = rng.standard_normal((n, n)); A = (G + G.T) / 2
    w, V = np.linalg.eigh(A); ref = (V * np.maximum(w, 0)) @ V.T
    X = A / np.linalg.norm(A); I = np.eye(n); res = []
    for k in range(40):
        if k >= 2 and (res[k - 1] < 1e-10 or (k - 1 >= 8 and abs(res[k - 2] - res[k - 1]) <= 1e-13 * res[k - 1])):
            break
        X2 = X @ X
        sc = 1 / np.sqrt(np.linalg.norm(X2)) if k == 0 else 1.0
        res.append(np.linalg.norm(X2 - I) ** 2)
        a, b, c = ctab[k] if k < len(ctab) else (15 / 8, -10 / 8, 3 / 8)
        W = c * sc ** 4 * (X2 @ X2) + b * sc ** 2 * X2
        X = sc * (X @ W) + a * sc * X
        X = (X + X.T) / 2
    P = 0.5 * (X @ A + A); P = (P + P.T) / 2
    assert k <= 14
    assert np.linalg.norm(P - ref) <= 1e-12 * np.linalg.norm(ref)
