"""CPU tier: level-1 NMOS companion stamps (EXTENSION, SURVEY §8 f4 — the reference's NMOS stamps nothing, stampable.rs:842-844).

No upstream oracle exists for this element, so the chain of evidence is: (1) the oracle restatement's extension
(oracle/krets_oracle.c: nmos_gm / nmos_gds / nmos_id follow nmosfet.rs:51-102 line by line) is pinned to an INDEPENDENT solve of the
circuit's KCL equations written here from documentation/main.typ:375-424; (2) the lane program (numpy emulation of the interpreter)
and the generated specialised kernel (host build) are held to that oracle at 1e-9 (OP, nonlinear DC sweep lanes, transient, AC);
(3) with the flag off, everything reproduces the reference's no-op (the inverter golden, covered by the existing tests)."""
import numpy as np
import pytest
from scipy.optimize import brentq

from helpers import assert_close
from jit_host import JitHost
from vm_emulator import Emulator, input_layout

import krets_b200 as kb
from krets_b200 import api

# common-source stage with source degeneration; the .model line exercises every level-1 parameter (models/nmosfet.rs:48-61)
AMP = """VDD vdd 0 DC 5
VG g 0 DC 1.5 AC 1
RD vdd d 10k
MN1 d g s 0 NM
RS s 0 1k
.model NM NMOS (VTO=0.7 KP=1e-4 W=20u L=2u LAMBDA=0.05)
"""
AMP_TRAN = AMP.replace("VG g 0 DC 1.5 AC 1", "VG g 0 SIN(1.5 0.2 1k 0 0 0)") + "CL d 0 10n\n"
VTH, BETA, LAM = 0.7, 1e-4 * (20e-6 / 2e-6), 0.05


def i_d(vgs, vds):   # documentation/main.typ:375-424
    if vgs <= VTH or vds < 0:
        return 0.0
    vov = vgs - VTH
    return BETA * (vov * vds - vds * vds / 2) if vds <= vov else BETA / 2 * vov * vov * (1 + LAM * vds)


def independent_op(vg, vdd=5.0, rd=10e3, rs=1e3):
    """KCL: the same current flows through RD, the channel and RS -> one scalar equation in i."""
    f = lambda i: i_d(vg - i * rs, (vdd - i * rd) - i * rs) - i
    i = brentq(f, 0.0, vdd / (rd + rs), xtol=1e-18, rtol=1e-15)
    return i, vdd - i * rd, i * rs


@pytest.fixture()
def level1(oracle):
    oracle.set_device_models(1)
    yield
    oracle.set_device_models(0)


def circuits(oracle, text):
    kc = kb.parse_circuit_description(text)
    assert kc.set_device_models(kb.DEVICES_LEVEL1) == 1
    return kc, oracle.Circuit(text)


def test_oracle_extension_against_independent_kcl_solve(oracle, level1):
    oc = oracle.Circuit(AMP)
    tight = oracle.SolverConfig(relative_tolerance=1e-13, voltage_absolute_tolerance=1e-15, current_absolute_tolerance=1e-18)
    x = oc.solve_op(tight)
    names = oc.names
    i, vd, vs = independent_op(1.5)
    assert vd - vs > 1.5 - vs - VTH > 0       # saturation at this bias
    assert abs(x[names.index("V(d)")] - vd) < 1e-10 and abs(x[names.index("V(s)")] - vs) < 1e-10
    assert abs(x[names.index("I(VDD)")] + i) < 1e-13
    # a sweep through cut-off, saturation and the linear region, warm-started from the previous point as dc.rs:58 does
    sw = oc.solve_dc("VG", 0.0, 5.0, 0.25, cfg=tight, warm_start=True)
    regions = set()
    for k, row in enumerate(sw):
        vg = 0.25 * k
        i, vd, vs = independent_op(vg)
        regions.add("off" if vg - vs <= VTH else ("lin" if vd - vs <= vg - vs - VTH else "sat"))
        assert abs(row[names.index("V(d)")] - vd) < 1e-9 and abs(row[names.index("V(s)")] - vs) < 1e-9, vg
    assert regions == {"off", "lin", "sat"}
    # flag off: the reference's behaviour — no channel current at all
    oracle.set_device_models(0)
    x0 = oc.solve_op()
    assert abs(x0[names.index("V(d)")] - 5.0) < 1e-12 and abs(x0[names.index("V(s)")]) < 1e-12


def test_flag_off_reproduces_the_reference_noop(oracle):
    kc = kb.parse_circuit_description(AMP)             # no set_device_models: reference behaviour
    plan = kb.Plan(kc, api.MODE_DC)
    X = Emulator(plan, 1).run(np.zeros((0, 1)))
    assert_close(X[0][:, 0], oracle.Circuit(AMP).solve_op(), what="NMOS no-op")
    with pytest.raises(kb.KretsError) as e:
        kb.Plan(kc, api.MODE_AC)                       # stampable.rs:914-921 todo!()
    assert e.value.kind == "Unsupported"


@pytest.mark.parametrize("runner", ["emulator", "specialised"])
def test_lane_program_op_and_sweep_lanes(oracle, level1, runner):
    kc, oc = circuits(oracle, AMP)
    Run = Emulator if runner == "emulator" else JitHost
    plan = kb.Plan(kc, api.MODE_DC)
    assert plan.info.nonlinear and plan.info.executor == 0
    r = Run(plan, 1)
    X = r.run(np.zeros((0, 1)))
    ref, it = oc.solve_op(return_iterations=True)
    assert_close(X[0][:, 0], ref, what="common-source OP")
    assert r.iters[0] == it and r.status[0] == 0
    # DC sweep of the gate source, one lane per point, every lane from the empty guess (as the diode sweep: test_gpu_parity.py)
    sweep = kb.Plan(kc, api.MODE_DC, sweep_element=kc.element_index("VG"))
    vals = np.arange(10) * 0.25          # cut-off, saturation and the first linear-region points
    r = Run(sweep, len(vals))
    X = r.run(vals[None, :])
    cold = oc.solve_dc("VG", 0.0, 2.25, 0.25, warm_start=False)
    assert cold.shape[0] == len(vals)
    assert_close(X[0].T, cold, what="gate sweep vs cold-start oracle")
    assert not r.status.any()
    # the reference's sweep order (dc.rs:58,112): the same points as the rows of one lane, warm-started from the previous point
    from vm_emulator import run_sequential_sweep
    Xs, rs, _ = run_sequential_sweep(Run, sweep, 0.0, 0.25, len(vals))
    assert_close(Xs[:, : sweep.n, 0], oc.solve_dc("VG", 0.0, 2.25, 0.25, warm_start=True), what="gate sweep vs the sequential warm-started oracle")
    assert not rs.status.any()
    # deep in the linear region the reference's un-limited Newton iteration cycles from an empty guess (v_ds < 0 falls into the
    # all-zero branch of nmosfet.rs:51-102): the oracle reports MaximumIterationsExceeded, the lane its NOT_CONVERGED bit
    cfg = dict(reltol=1e-3, iabstol=1e-12, vabstol=1e-6, max_iter=40)
    r = Run(sweep, 1, cfg=cfg)
    r.run(np.array([[3.0]]))
    assert r.status[0] & 2
    with pytest.raises(oracle.OracleError) as e:
        oc.solve_dc("VG", 3.0, 3.0, 0.25, cfg=oracle.SolverConfig(maximum_iterations=40), warm_start=False)
    assert e.value.kind == "MaximumIterationsExceeded"


@pytest.mark.parametrize("runner", ["emulator", "specialised"])
def test_lane_program_transient_and_ac(oracle, level1, runner):
    Run = Emulator if runner == "emulator" else JitHost
    kc, oc = circuits(oracle, AMP_TRAN)
    h, stop = 10e-6, 2e-3
    pdc = kb.Plan(kc, api.MODE_DC)
    ptr = kb.Plan(kc, api.MODE_TRAN, time_step=h)
    t, ref = oc.solve_tran(h, stop)
    rows = len(t)
    X = Run(pdc, 1, rows=rows).run(np.zeros((0, 1)), row_begin=0, row_end=1)
    r = Run(ptr, 1, rows=rows, h=h)
    X = r.run(np.zeros((0, 1)), X=X, row_begin=1, guess_empty_at_iter0=False, pseudo_time=True)
    assert_close(X[:, :, 0], ref, what="common-source transient")
    assert not r.status.any()
    d = oc.names.index("V(d)")
    assert ref[:, d].max() - ref[:, d].min() > 0.5          # the stage amplifies: the drain swings
    # AC: small-signal g_m / g_ds at the operating point
    kc, oc = circuits(oracle, AMP)
    freqs = np.array([10.0, 1e3, 1e5])
    op = oc.solve_op()
    pac = kb.Plan(kc, api.MODE_AC, nominal_frequency=1e3)
    Xa = Run(pac, len(freqs)).run(freqs[None, :], Xprev0=np.repeat(op[:, None], len(freqs), axis=1))
    _, refa = oc.solve_ac(freqs)
    assert_close(Xa[0].T, refa, what="common-source AC")
    gain = refa[0, d]
    i, vd, vs = independent_op(1.5)
    vov = 1.5 - vs - VTH
    gm, gds = BETA * vov * (1 + LAM * (vd - vs)), BETA / 2 * LAM * vov * vov
    expect = -gm * 10e3 / (1 + gm * 1e3 + gds * (10e3 + 1e3))       # degenerated common-source gain with finite r_o
    assert abs(gain.real - expect) < 2e-3 * abs(expect) and abs(gain.imag) < 1e-9
