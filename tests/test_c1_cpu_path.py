"""BASELINE.json configs[0]: PlanarHand_N=1_MOMENT with the PSD projection on the reference's LAPACK dsyevd CPU path
(one thread-pool run per iteration) — the CPU oracle (oracle_np.ADMMOracle + oracle/cpu_baseline.cpp) on the
regenerated fixture must follow the reference's committed log.  CPU only; the full 800 / 878-iteration runs were
checked once with the same script (scripts/make_bundled_fixtures.py docstring) and are asserted on the GPU path in
tests/test_gpu_bundled.py."""
import os

import numpy as np

import oracle_np as onp
from util_problems import GOLD, load_fixture, parse_log


def test_planarhand_first_iterations_follow_the_reference_log():
    P = load_fixture("planarhand_n1")
    assert P["vec_len"] == 55179 and P["con_num"] == 66008 and len(P["vals"]) == 156635
    blk = np.ascontiguousarray(P["blk"], np.int32)
    o = onp.ADMMOracle(P["vec_len"], P["con_num"], P["col_ptrs"], P["row_ids"], P["vals"], P["b_idx"], P["b_val"],
                       P["C_idx"], P["C_val"], blk, project=lambda v: onp.project_svec_cpp(blk, v, 4))
    # iteration-0 line of the log: the residuals of the zero start
    rows = parse_log(os.path.join(GOLD, "planarhand_n1_sgs.log"))
    assert abs(o.errRp - rows[0]["errRp"]) <= 6e-3 * rows[0]["errRp"] and abs(o.errRd - rows[0]["errRd"]) <= 6e-3 * rows[0]["errRd"]
    o.solve(50, 1e-3, 0, 50, 100, 11000, 1.05)
    r = [x for x in rows if x["it"] == 50][0]
    assert abs(o.hist["errRp"][49] - r["errRp"]) <= 6e-3 * r["errRp"]
    assert abs(o.hist["errRd"][49] - r["errRd"]) <= 6e-3 * r["errRd"]
    assert abs(o.hist["pobj"][49] - r["pobj"]) <= 6e-4 * abs(r["pobj"])
    assert abs(o.hist["dobj"][49] - r["dobj"]) <= 6e-4 * abs(r["dobj"])
    assert abs(o.hist["sig"][49] - r["sig"]) <= 0.06 * r["sig"]
