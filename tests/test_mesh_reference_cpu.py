"""CPU tier: tests/mesh_reference.py (SuperLU on oracle-stamped matrices + restated trapezoidal right-hand sides) is held to the
oracle's own drivers on a small RLC mesh, so that the GPU tier may use it at BASELINE config-5 size where the oracle's
natural-order LU cannot run in seconds."""
import numpy as np

import mesh_reference as mr
import netgen
from helpers import assert_close


def test_trap_rows_match_the_oracle(oracle):
    text = netgen.rlc_mesh(6, 7, transient=True)
    oc = oracle.Circuit(text)
    tab = mr.ElementTable(oc)
    h, steps = 2e-7, 60
    for e in range(oc.num_elements):                  # a Monte-Carlo variant, not the nominal circuit
        if oc.element_kind(e) in (mr.KIND_R, mr.KIND_C, mr.KIND_L):
            oc.set_element_value(e, oc.element_value(e) * netgen.mc_factor(3, e, oc.num_elements))
    rows = mr.trap_rows(oc, tab, h, steps)
    _, ref = oc.solve_tran(h, steps * h, integrator=1)
    assert_close(rows, ref, rtol=1e-9, what="mesh_reference.trap_rows vs oracle trapezoidal driver")


def test_ac_and_dc_systems_match_the_oracle(oracle):
    import scipy.sparse.linalg as spl
    text = netgen.rlc_mesh(6, 7, transient=False)
    oc = oracle.Circuit(text)
    tab = mr.ElementTable(oc)
    val = tab.values(oc)
    freqs = np.array([1e3, 1e6, 1e8])
    _, ref = oc.solve_ac(freqs)
    for f, r in zip(freqs, ref):
        x = spl.splu(mr.csc(oc, "ac", float(f))).solve(mr.ac_rhs(oc, tab, val))
        assert_close(x, r, rtol=1e-9, what=f"AC system at {f} Hz")
    oc = oracle.Circuit(netgen.rlc_mesh(6, 7, transient=True))
    tab = mr.ElementTable(oc)
    x = spl.splu(mr.csc(oc, "dc", 0.0)).solve(mr.dc_rhs(oc, tab, tab.values(oc)))
    assert_close(x, oc.solve_op(), rtol=1e-9, what="DC system")
