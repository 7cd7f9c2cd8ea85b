"""Generates the coefficient table kSignPoly of cuadmm_b200/csrc/dense_proj.cu.

Greedy minimax composition for the matrix sign function: given that the spectrum of X lies in
[-1, -l] u [l, 1], the odd degree-5 polynomial p(x) = a x + b x^3 + c x^5 with  p <= 1 on [l, 1]  that
maximises  min_{[l,1]} p  (a linear programme on a grid) maps [l, 1] into [l', 1] with the largest possible
l'; iterate l <- l' from l_0 = 1e-4 until l >= 0.9, then the solver continues with the cubically convergent
Newton-Schulz polynomial (15 x - 10 x^3 + 3 x^5) / 8.  The script also checks the two properties the
kernel relies on: p <= 1 (+1e-9) on [0, 1], and p(x) >= x on [0, l] (eigenvalues below the assumed range are
never pushed back).  CPU only (scipy.optimize.linprog).
"""
import numpy as np
from scipy.optimize import linprog


def step(l):
    xs = np.unique(np.concatenate([np.geomspace(l, 1.0, 4000), np.linspace(l, 1.0, 4000)]))
    A = np.stack([xs, xs ** 3, xs ** 5], 1)
    n = len(xs)
    if l > 1e-3:      # variables a, b, c, r: maximise r subject to r * l <= p(x) <= 1
        Aub = np.block([[A, np.zeros((n, 1))], [-A, l * np.ones((n, 1))]])
        bub = np.concatenate([np.ones(n), np.zeros(n)])
        a, b, c, _ = linprog([0, 0, 0, -1], A_ub=Aub, b_ub=bub, bounds=[(None, None)] * 3 + [(0, None)], method="highs").x
    else:             # tiny l: min p is p(l) ~ a l; maximise the slope a subject to a * l <= p(x) <= 1
        A2 = A.copy()
        A2[:, 0] -= l
        Aub = np.block([[A], [-A2]])
        bub = np.concatenate([np.ones(n), np.zeros(n)])
        a, b, c = linprog([-1, 0, 0], A_ub=Aub, b_ub=bub, bounds=[(None, None)] * 3, method="highs").x
    xx = np.unique(np.concatenate([np.geomspace(l, 1.0, 100000), np.linspace(l, 1.0, 100000)]))
    p = a * xx + b * xx ** 3 + c * xx ** 5
    sc = 1.0 / max(p.max(), 1.0)      # the grid LP may overshoot 1 between grid points: shrink
    return a * sc, b * sc, c * sc, float((p * sc).min())


def table(l0=1e-4, lend=0.9):
    l, tab = l0, []
    while l < lend:
        a, b, c, lo = step(l)
        tab.append((a, b, c, l, lo))
        l = lo
    return tab


if __name__ == "__main__":
    tab = table()
    xs = np.concatenate([np.linspace(0, 1, 2000001), np.geomspace(1e-12, 1, 1000001)])
    for a, b, c, l, lo in tab:
        p = a * xs + b * xs ** 3 + c * xs ** 5
        assert p.max() <= 1 + 1e-9 and (p - xs)[xs <= l].min() >= -1e-15
        print("    {%.17g, %.17g, %.17g},   // [%.3e, 1] -> [%.6f, 1]" % (a, b, c, l, lo))
