"""CPU tier: randomly generated small circuits through the plan compiler, the lane program run on the CPU emulator, against the oracle
(OP, AC, backward-Euler transient; linear and diode circuits).  The generator is deterministic (numpy Generator, fixed seeds).

This is how the static pivot plan was found to need a numeric check: a floating voltage source whose terminals are also joined through
companion conductances gives a pivot sequence that cancels exactly for any element values (compiler.cpp: static_pivots_ok); the
reference never sees it because faer pivots numerically at every factorisation (op.rs:50)."""
import numpy as np
import pytest

import krets_b200 as kb
from helpers import rel_err
from krets_b200 import _lib, api
from vm_emulator import Emulator

FLOATING_SOURCE = "R2 2 0 1031.41\nR3 3 2 200.105\nC1 4 3 1.01013e-07\nV2 4 2 -1.689\n"


def random_circuit(rng, diode):
    """nodes 1..nn on a random resistor tree to ground (every node has a DC path), one grounded source, up to five more elements between
    random node pairs; diodes come with a series resistor (a diode straight across a source has no meaningful solution)."""
    nn = int(rng.integers(2, 7))
    lines = [f"R{k} {k} {int(rng.integers(0, k))} {10 ** rng.uniform(1, 4):.6g}" for k in range(1, nn + 1)]
    lines.append(f"V1 1 0 {rng.uniform(0.5, 5):.4g} AC 1")
    cnt = {"R": nn, "C": 0, "L": 0, "I": 0, "V": 1, "D": 0}
    for _ in range(int(rng.integers(1, 6))):
        a, b = rng.choice(nn + 1, size=2, replace=False)
        kind = rng.choice(["R", "C", "L", "I", "V"] + (["D"] if diode else []))
        if kind == "V" and (a == 1 or b == 1 or cnt["V"] >= 2):
            kind = "R"
        cnt[kind] += 1
        nm = f"{kind}{cnt[kind]}"
        if kind == "R":
            lines.append(f"{nm} {a} {b} {10 ** rng.uniform(1, 4):.6g}")
        elif kind == "C":
            lines.append(f"{nm} {a} {b} {10 ** rng.uniform(-9, -6):.6g}")
        elif kind == "L":
            lines.append(f"{nm} {a} {b} {10 ** rng.uniform(-6, -3):.6g}")
        elif kind == "I":
            lines.append(f"{nm} {a} {b} {rng.uniform(-1e-3, 1e-3):.6g}")
        elif kind == "V":
            lines.append(f"{nm} {a} {b} {rng.uniform(-2, 2):.4g}")
        else:
            lines.append(f"{nm} {a} x{cnt['D']} DMOD")
            lines.append(f"RD{cnt['D']} x{cnt['D']} {b} {10 ** rng.uniform(2, 4):.6g}")
    if cnt["D"]:
        lines.append(".model DMOD D (is=1e-12)")
    return "\n".join(lines) + "\n"


def _oracle(fn):
    try:
        return fn(), None
    except Exception as e:      # singular at this operating point, not convergent ...: the CUDA side must fail as well
        return None, str(e)


def _emulated(fn):
    try:
        X, status = fn()
        return (None, f"lane status {status}") if np.any(status) else (X, None)
    except kb.KretsError as e:
        return None, str(e)


def run_transient(kc, h, nsteps):
    pdc = kb.Plan(kc, api.MODE_DC, force_lane_vm=True)
    ptr = kb.Plan(kc, api.MODE_TRAN, time_step=h, force_lane_vm=True)
    X = Emulator(pdc, 1, rows=nsteps + 1).run(np.zeros((0, 1)), row_begin=0, row_end=1)
    em = Emulator(ptr, 1, rows=nsteps + 1, h=h)
    X = em.run(np.zeros((0, 1)), X=X, row_begin=1, guess_empty_at_iter0=False, pseudo_time=True)
    return X[:, : kc.n, 0], em.status, ptr


def check_circuit(oracle, text, diode):
    oc, kc = oracle.Circuit(text), kb.parse_circuit_description(text)
    assert kc.names == oc.names
    problems = []

    def compare(what, ref, oerr, got, kerr, tol):
        if (ref is None) != (got is None):
            problems.append(f"{what}: oracle {'failed: ' + oerr if ref is None else 'solved'}, lane program {'failed: ' + kerr if got is None else 'solved'}")
        elif ref is not None and not rel_err(got, ref) < tol:
            problems.append(f"{what}: max rel err {rel_err(got, ref):.3g}")

    ref, oerr = _oracle(oc.solve_op)

    def op():
        em = Emulator(kb.Plan(kc, api.MODE_DC, force_lane_vm=True), 1)
        return em.run(np.zeros((0, 1)))[0][:, 0], em.status
    got, kerr = _emulated(op)
    compare("OP", ref, oerr, got, kerr, 1e-8)
    if ref is None:
        return problems
    if not diode:
        freqs = np.array([1.0, 1e3, 1e6])
        refa, oerr = _oracle(lambda: oc.solve_ac(freqs)[1])

        def ac():
            em = Emulator(kb.Plan(kc, api.MODE_AC, nominal_frequency=1e3, force_lane_vm=True), 3)
            return em.run(freqs[None, :])[0].T, em.status
        got, kerr = _emulated(ac)
        compare("AC", refa, oerr, got, kerr, 1e-8)
    reft, oerr = _oracle(lambda: oc.solve_tran(1e-6, 20e-6)[1])
    got, kerr = _emulated(lambda: run_transient(kc, 1e-6, 20)[:2])
    compare("transient", reft, oerr, got, kerr, 1e-6 if diode else 1e-7)
    return problems


@pytest.mark.parametrize("diode, seeds", [(False, range(0, 300)), (True, range(300, 600))])
def test_random_circuits_lane_program_vs_oracle(oracle, diode, seeds):
    bad = {}
    for seed in seeds:
        text = random_circuit(np.random.default_rng(seed), diode)
        problems = check_circuit(oracle, text, diode)
        if problems:
            bad[seed] = (problems, text)
    assert not bad, bad


def test_found_by_the_fuzz_floating_source_with_companion_path(oracle, monkeypatch):
    """The transversal + minimum-degree pivot sequence of this transient plan cancels exactly (I(V2)'s pivot: the KCL rows of the source's
    terminals sum to an equation without the current).  The numeric check at the nominal values re-pivots it; without the check every lane
    reports a zero pivot where the reference solves."""
    kc, oc = kb.parse_circuit_description(FLOATING_SOURCE), oracle.Circuit(FLOATING_SOURCE)
    _, ref = oc.solve_tran(1e-6, 20e-6)
    X, status, plan = run_transient(kc, 1e-6, 20)
    assert _lib.lib().krb_plan_repivoted(plan._h) == 1 and not status.any()
    assert np.max(np.abs(X - ref)) < 1e-12
    assert _lib.lib().krb_plan_repivoted(kb.Plan(kc, api.MODE_DC, force_lane_vm=True)._h) == 0        # the DC plan (no capacitor) never needed it
    monkeypatch.setenv("KRB_NO_PIVOT_CHECK", "1")
    _, status, plan = run_transient(kc, 1e-6, 20)
    assert _lib.lib().krb_plan_repivoted(plan._h) == 0 and (status[0] & 1)          # KRB_LANE_ZERO_PIVOT


def test_found_by_the_fuzz_nonlinear_variant(oracle):
    """the same sub-circuit inside a diode circuit (lane VM with Newton-Raphson: no pivoting fallback exists for nonlinear plans)"""
    text = ("R1 1 0 2069.22\nR2 2 0 1031.41\nR3 3 2 200.105\nR4 4 2 9813.09\nV1 1 0 4.765\nC1 4 3 1.01013e-07\nI1 2 4 0.00031127\nV2 4 2 -1.689\n"
            "D1 1 x1 DMOD\nRD1 x1 0 110.868\n.model DMOD D (is=1e-12)\n")
    assert check_circuit(oracle, text, True) == []


# ---- ladder-like circuits through the default executor selection (chain executor when the plan is banded) ------------------------
def random_line(rng, mixed):
    """a line of 4..40 nodes: series R between neighbours (mixed: also L, and now and then a floating voltage source — those take the plan
    out of the chain executor's band class), shunt R / C / nothing to ground, a source at one end, a load at the other"""
    nn = int(rng.integers(4, 41))
    lines = [f"V1 1 0 {rng.uniform(0.5, 5):.4g} AC 1"]
    nv = 1
    for k in range(1, nn):
        kind = rng.choice(["R", "R", "R", "L", "V"]) if mixed else "R"
        if kind == "R":
            lines.append(f"RS{k} {k} {k + 1} {10 ** rng.uniform(0, 4):.6g}")
        elif kind == "L":
            lines.append(f"LS{k} {k} {k + 1} {10 ** rng.uniform(-7, -3):.6g}")
        else:
            nv += 1
            lines.append(f"V{nv} {k} {k + 1} {rng.uniform(-1, 1):.4g}")
        sh = rng.choice(["R", "C", "-"])
        if sh == "R":
            lines.append(f"RP{k} {k + 1} 0 {10 ** rng.uniform(1, 5):.6g}")
        elif sh == "C":
            lines.append(f"CP{k} {k + 1} 0 {10 ** rng.uniform(-12, -7):.6g}")
    lines.append(f"RL {nn} 0 {10 ** rng.uniform(1, 4):.6g}")
    return "\n".join(lines) + "\n"


def _solve_default(plan, IN, cplx):
    from vm_emulator import run_chain
    if plan.info.executor == 1:
        X, status = run_chain(plan, IN, cplx)
        return X.T, status
    assert plan.info.executor == 0
    em = Emulator(plan, IN.shape[1])
    return em.run(IN)[0].T, em.status


def test_random_lines_default_executor_vs_oracle(oracle):
    bad, chain_plans = {}, 0
    for seed in range(1000, 1150):
        text = random_line(np.random.default_rng(seed), mixed=seed >= 1100)
        oc, kc = oracle.Circuit(text), kb.parse_circuit_description(text)
        vals = np.array([-1.0, 0.5, 2.0])
        ref, oerr = _oracle(lambda: np.array([oc.solve_dc("V1", v, v, 1.0)[0] for v in vals]))
        try:
            plan = kb.Plan(kc, api.MODE_DC, sweep_element=kc.element_index("V1"))
            chain_plans += plan.info.executor == 1
            got, status = _solve_default(plan, vals[None, :], False)
            kerr = f"lane status {status}" if status.any() else None
        except kb.KretsError as e:
            got, kerr = None, str(e)
        if (ref is None) != (got is None or kerr is not None):
            bad[seed] = ("DC", oerr, kerr, text)
            continue
        if ref is not None and not rel_err(got, ref) < 1e-8:
            bad[seed] = ("DC", rel_err(got, ref), text)
            continue
        freqs = np.array([10.0, 1e4, 1e7])
        refa, oerr = _oracle(lambda: oc.solve_ac(freqs)[1])
        try:
            plan = kb.Plan(kc, api.MODE_AC, nominal_frequency=1e4)
            chain_plans += plan.info.executor == 1
            got, status = _solve_default(plan, freqs[None, :], True)
            kerr = f"lane status {status}" if status.any() else None
        except kb.KretsError as e:
            got, kerr = None, str(e)
        if (refa is None) != (got is None or kerr is not None):
            bad[seed] = ("AC", oerr, kerr, text)
        elif refa is not None and not rel_err(got, refa) < 1e-8:
            bad[seed] = ("AC", rel_err(got, refa), text)
    assert not bad, bad
    assert chain_plans >= 50          # the generator really exercises the chain executor's tables


# ---- the front executor's tables (fronts.cpp) on random circuits, through the numpy multifrontal emulator ------------------------
def test_random_circuits_front_tables_vs_superlu(oracle):
    """Random circuits and lines analysed with KRB_EXEC_STAGED: supernodes, relative indices, assembly lists and level order are run by
    tests/front_emulator.py (threshold pivoting inside the fully-summed blocks, as the kernels do) on the oracle-stamped matrix and
    compared with SuperLU.  Matrices that are singular or worse conditioned than 1e13 are skipped."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl

    import front_emulator as fe
    from test_front_cpu import _check_tables
    done, bad = 0, {}
    for seed in range(2000, 2120):
        rng = np.random.default_rng(seed)
        text = random_circuit(rng, False) if seed % 2 else random_line(rng, True)
        oc, kc = oracle.Circuit(text), kb.parse_circuit_description(text)
        n = kc.n
        for mode, kmode, par in (("dc", api.MODE_DC, 0.0), ("ac", api.MODE_AC, 2.5e4), ("tran", api.MODE_TRAN, 1e-6)):
            try:
                ocp, ori, vals = oc.stamp_matrix(mode, par)
                A = sp.csc_matrix((vals, ori, ocp), shape=(n, n))
                b = np.random.default_rng(7).standard_normal(n) + (1j * np.random.default_rng(8).standard_normal(n) if mode == "ac" else 0.0)
                with np.errstate(all="ignore"):
                    xr = spl.splu(A).solve(b)
                    if not np.all(np.isfinite(xr)) or np.linalg.cond(A.toarray()) > 1e13:
                        continue
            except Exception:      # singular for SuperLU / the oracle: nothing to compare
                continue
            plan = kb.Plan(kc, kmode, executor=api.EXEC_STAGED, time_step=1e-6 if mode == "tran" else 0.0, nominal_frequency=2.5e4)
            ft = plan.front_tables()
            _check_tables(ft, n)
            ro, co = plan.pivot_order()
            x = fe.factor_and_solve(ft, ro, co, vals, b, dtype=np.complex128 if mode == "ac" else np.float64)
            done += 1
            err = np.max(np.abs(x - xr)) / np.max(np.abs(xr))
            if not err <= 1e-8:
                bad[(seed, mode)] = (err, text)
    assert not bad, bad
    assert done >= 200


# ---- netlist front end: random formatting, value suffixes and source forms, against the oracle's restatement of the parser -----------
_SUFFIXES = ["", "k", "K", "meg", "MEG", "Meg", "m", "u", "n", "p", "f", "g", "G", "t", "T", "mil", "e3", "e-3", "E6", "kohm", "uF", "nH", "V", "A", "Hz", "x"]


def _value(rng):
    m = rng.choice([f"{rng.uniform(0.1, 999):.4g}", f"{int(rng.integers(1, 1000))}", f"{rng.uniform(0.1, 9):.3f}", f".{int(rng.integers(1, 99))}", f"{int(rng.integers(1, 99))}."])
    return m + (rng.choice(_SUFFIXES[:12]) if rng.random() < 0.9 else rng.choice(_SUFFIXES))


def random_netlist_text(rng):
    nn = int(rng.integers(2, 6))
    names = [rng.choice(["n", "N", "out", "in", "a", "x_"]) + str(k) for k in range(nn)]
    nodes = ["0"] + names
    lines = []
    if rng.random() < 0.3:
        lines.append(rng.choice(["* title line", "% comment", "", "   "]))
    for k, nm in enumerate(names):
        lines.append(f"R{k} {nm} {nodes[int(rng.integers(0, k + 1))]} {_value(rng)}" + (rng.choice(["", " extra", " ; c", " $ c"]) if rng.random() < 0.2 else ""))
    for j in range(int(rng.integers(1, 6))):
        a, b = rng.choice(len(nodes), size=2, replace=False)
        kind = rng.choice(["R", "C", "L", "V", "I", "r", "c", "v"])
        ident = f"{kind}x{j}"
        if kind in "Vv":
            form = rng.choice(["dc", "plain", "ac", "sin", "pulse"])
            v = {"dc": f"DC {_value(rng)}", "plain": _value(rng), "ac": f"DC {_value(rng)} AC {_value(rng)}",
                 "sin": f"SIN({_value(rng)} {_value(rng)} {_value(rng)} 0 0 0)", "pulse": f"PULSE(0 {_value(rng)} 1n 1n 1n 1u 2u)"}[form]
            lines.append(f"{ident} {nodes[a]} {nodes[b]} {v}")
        else:
            lines.append(f"{ident} {nodes[a]} {nodes[b]} {_value(rng)}" + (" G2" if kind in "Cc" and rng.random() < 0.3 else ""))
        if rng.random() < 0.15:
            lines.append(rng.choice(["* c", "%c", "", ".control\nop\n.endc"]))
    return "\n".join(lines) + "\n"


def test_random_netlists_front_end_vs_oracle(oracle):
    """accept / reject decisions, the MNA index map (parser.rs:105-125: bit-exact) and every parsed element value (utils.rs:32-88 suffix
    rules, voltage_source.rs:194 implicit DC) agree between the C++ host front end and the oracle's restatement of the parser"""
    bad, parsed = {}, 0
    for seed in range(3000, 3300):
        text = random_netlist_text(np.random.default_rng(seed))
        oc, oerr = _oracle(lambda: oracle.Circuit(text))
        try:
            kc, kerr = kb.parse_circuit_description(text), None
        except kb.KretsError as e:
            kc, kerr = None, str(e)
        if (oc is None) != (kc is None):
            bad[seed] = ("accept/reject", oerr, kerr, text)
            continue
        if oc is None:
            continue
        parsed += 1
        if oc.names != kc.names:
            bad[seed] = ("index map", oc.names, kc.names, text)
            continue
        for e in range(kc.num_elements):
            a, b = oc.element_value(e), kc.element_value(e)
            if not (a == b or (a != a and b != b)):
                bad[seed] = ("value", kc.identifiers[e], a, b, text)
                break
    assert not bad, bad
    assert parsed >= 100


# ---- the block-partitioned chain kernel's tables (the headline kernel) on random NON-uniform lines ---------------------------------
def test_random_long_lines_block_chain_tables_vs_oracle(oracle):
    """Lines of 515..900 nodes with series resistors over four decades and random R / C shunts: the numpy restatements of
    krb_chain_block_kernel (Wang partition + cyclic reduction) and of the sequential chain kernel against the oracle, DC sweep and AC."""
    from vm_emulator import chain_block_info, run_chain, run_chain_block
    bad, blocks = {}, 0
    for seed in range(4000, 4030):
        rng = np.random.default_rng(seed)
        nn = int(rng.integers(515, 900))
        lines = [f"V1 1 0 {rng.uniform(0.5, 5):.4g} AC 1"]
        for k in range(1, nn):
            lines.append(f"RS{k} {k} {k + 1} {10 ** rng.uniform(0, 4):.6g}")
            sh = rng.choice(["R", "C", "-"])
            if sh == "R":
                lines.append(f"RP{k} {k + 1} 0 {10 ** rng.uniform(1, 5):.6g}")
            elif sh == "C":
                lines.append(f"CP{k} {k + 1} 0 {10 ** rng.uniform(-12, -7):.6g}")
        lines.append(f"RL {nn} 0 {10 ** rng.uniform(1, 4):.6g}")
        text = "\n".join(lines) + "\n"
        oc, kc = oracle.Circuit(text), kb.parse_circuit_description(text)
        for mode in ("dc", "ac"):
            if mode == "dc":
                plan, IN, cplx = kb.Plan(kc, api.MODE_DC, sweep_element=kc.element_index("V1")), np.array([[-1.0, 0.5, 2.0]]), False
                ref = np.array([oc.solve_dc("V1", v, v, 1.0)[0] for v in IN[0]])
            else:
                plan, IN, cplx = kb.Plan(kc, api.MODE_AC, nominal_frequency=1e4), np.array([[10.0, 1e4, 1e7]]), True
                ref = oc.solve_ac(IN[0])[1]
            assert plan.info.executor == 1
            Xs, st = run_chain(plan, IN, cplx)
            if st.any() or not rel_err(Xs.T, ref) < 1e-8:
                bad[(seed, mode, "sequential")] = (rel_err(Xs.T, ref), st)
            valid = chain_block_info(plan)[0]
            if valid:
                blocks += 1
                X, stb = run_chain_block(plan, IN, cplx)
                if stb.any() or not rel_err(X.T, ref) < 1e-8:
                    bad[(seed, mode, "block")] = (rel_err(X.T, ref), stb)
    assert not bad, bad
    assert blocks >= 50


# ---- the circuit-specialised kernel's generated statements (vm_jit.cpp, host build) against the interpreter emulator ----------------
def test_random_circuits_specialised_kernel_vs_emulator(oracle):
    """vm_jit.cpp writes the lane program out as straight-line code; tests/jit_host.py compiles the same text for the host.  On random
    circuits (every second one with diodes) the OP + 10 backward-Euler steps and the AC sweep must agree with the interpreter emulator:
    same failed / solved lanes, same Newton-Raphson iteration counts, values to 1e-8."""
    from jit_host import JitHost
    checked = 0
    for seed in range(5000, 5024):
        diode = seed % 2 == 1
        text = random_circuit(np.random.default_rng(seed), diode)
        kc = kb.parse_circuit_description(text)
        try:
            pdc = kb.Plan(kc, api.MODE_DC, force_lane_vm=True)
            ptr = kb.Plan(kc, api.MODE_TRAN, time_step=1e-6, force_lane_vm=True)
        except kb.KretsError:
            continue
        outs = []
        for driver in (Emulator, JitHost):
            d0 = driver(pdc, 1, rows=11)
            X = d0.run(np.zeros((0, 1)), row_begin=0, row_end=1)
            d = driver(ptr, 1, rows=11, h=1e-6)
            X = d.run(np.zeros((0, 1)), X=X, row_begin=1, guess_empty_at_iter0=False, pseudo_time=True)
            outs.append((X[:, : kc.n, 0].copy(), bool(d0.status[0] or d.status[0]), int(d0.iters[0]), int(d.iters[0])))
        (xa, fa, i0a, ia), (xb, fb, i0b, ib) = outs
        assert fa == fb and (i0a, ia) == (i0b, ib), (seed, outs, text)
        if not fa:
            assert rel_err(xb, xa) < 1e-8, (seed, rel_err(xb, xa), text)
            checked += 1
        if not diode:
            try:
                pa = kb.Plan(kc, api.MODE_AC, nominal_frequency=1e3, force_lane_vm=True)
            except kb.KretsError:      # AC with a current source: structurally singular (SURVEY B-3)
                continue
            f = np.array([[1.0, 1e3, 1e6]])
            assert rel_err(JitHost(pa, 3).run(f)[0].T, Emulator(pa, 3).run(f)[0].T) < 1e-8, (seed, text)
    assert checked >= 12


# ---- Monte-Carlo lanes (R / C / L / V / I values per lane) x the three integrators ----------------------------------------------------
def _mc_transient(oracle, text, varied, vals, integ, h=1e-6, nsteps=12):
    """OP + nsteps steps of every lane on the emulator and on the oracle (element values set per lane); returns the list of problems"""
    from vm_emulator import input_layout
    kc = kb.parse_circuit_description(text)
    lanes = vals.shape[1]
    pdc = kb.Plan(kc, api.MODE_DC, varied=varied, force_lane_vm=True)
    ptr = kb.Plan(kc, api.MODE_TRAN, varied=varied, time_step=h, integrator=integ, force_lane_vm=True)
    assert input_layout(pdc)[2] == input_layout(ptr)[2]               # one input buffer serves both plans (runtime.cu)
    IN = np.zeros((max(ptr.info.n_inputs, 1), lanes))
    for k, pl in enumerate(input_layout(ptr)[2]):
        IN[pl] = vals[k]
    rows = nsteps + 1
    e0 = Emulator(pdc, lanes, rows=rows)
    X = e0.run(IN, row_begin=0, row_end=1)
    et = Emulator(ptr, lanes, rows=rows, h=h)
    if et.width > kc.n:                                               # trapezoidal: capacitor-current state columns, i_C(0) = 0
        X2 = np.zeros((rows, et.width, lanes))
        X2[:, : kc.n] = X[:, : kc.n]
        X = X2
    X = et.run(IN, X=X, row_begin=1, guess_empty_at_iter0=False, pseudo_time=True)
    problems = []
    for l in range(lanes):
        oc = oracle.Circuit(text)
        for k, e in enumerate(varied):
            oc.set_element_value(e, vals[k, l])
        ref, _ = _oracle(lambda: oc.solve_tran(h, nsteps * h, integrator=integ)[1])
        failed = bool(e0.status[l] or et.status[l])
        if (ref is None) != failed:
            problems.append(f"lane {l}: oracle {'failed' if ref is None else 'solved'}, lane status {e0.status[l]} / {et.status[l]}")
        elif ref is not None and not rel_err(X[:, : kc.n, l], ref) < 1e-7:
            problems.append(f"lane {l}: max rel err {rel_err(X[:, : kc.n, l], ref):.3g}")
    return problems


def test_random_circuits_monte_carlo_lanes_and_integrators(oracle):
    bad, done = {}, 0
    for seed in range(6000, 6150):
        rng = np.random.default_rng(seed)
        text = random_circuit(rng, False)
        kc = kb.parse_circuit_description(text)
        cand = [e for e in range(kc.num_elements) if kc.element(e).kind in (api.KIND_R, api.KIND_C, api.KIND_L, api.KIND_V, api.KIND_I)]
        varied = sorted(rng.choice(cand, size=min(len(cand), int(rng.integers(1, 5))), replace=False).tolist())
        integ = int(rng.choice([api.INTEGRATOR_BE_KRETS, api.INTEGRATOR_TRAP, api.INTEGRATOR_BE]))
        vals = np.array([[kc.element_value(e) * (1.0 if l == 0 else 0.95 + 0.1 * rng.random()) for l in range(3)] for e in varied])
        try:
            problems = _mc_transient(oracle, text, varied, vals, integ)
        except kb.KretsError:          # structurally singular circuit: rejected at analysis time
            continue
        done += 1
        if problems:
            bad[seed] = (problems, integ, varied, text)
    assert not bad, bad
    assert done >= 100


def test_found_by_the_fuzz_varied_dc_source_in_a_transient_plan(oracle):
    """A voltage source without a waveform whose DC value varies per lane: its transient value is that DC value (voltage_source.rs:244-252),
    so every lane must step with ITS value — the lane-invariant per-row source table used to hold the nominal value for all of them."""
    text = "R1 1 0 2220.02\nR2 2 1 792.65\nV1 1 0 2.975\nL1 2 1 0.000121718\nR3 1 0 550.966\n"
    for integ in (api.INTEGRATOR_BE_KRETS, api.INTEGRATOR_TRAP):
        assert _mc_transient(oracle, text, [2], np.array([[2.975, 3.1, 1.0]]), integ) == []


def test_random_monte_carlo_dc_sweep_and_ac_all_lane_executors(oracle):
    """DC-sweep lanes and AC lanes with per-lane element values, on whichever lane executor the plan selects (sequential chain tables or
    the lane program; half of the plans forced onto the lane VM), every lane against the oracle with that lane's values."""
    from vm_emulator import input_layout, run_chain
    bad, done = {}, 0

    def solve(plan, IN, cplx):
        if plan.info.executor == 1:
            X, st = run_chain(plan, IN, cplx)
            return X.T, st
        em = Emulator(plan, IN.shape[1])
        return em.run(IN)[0].T, em.status

    for seed in range(7000, 7150):
        rng = np.random.default_rng(seed)
        diode = seed % 3 == 2
        text = random_line(rng, False) if seed % 3 == 0 else random_circuit(rng, diode)
        kc = kb.parse_circuit_description(text)
        cand = [e for e in range(kc.num_elements)
                if kc.element(e).kind in (api.KIND_R, api.KIND_C, api.KIND_L, api.KIND_V, api.KIND_I) and kc.identifiers[e] != "V1"]
        varied = sorted(rng.choice(cand, size=min(len(cand), int(rng.integers(1, 5))), replace=False).tolist())
        vals = np.array([[kc.element_value(e) * (1.0 if l == 0 else 0.95 + 0.1 * rng.random()) for l in range(3)] for e in varied])
        force = bool(rng.integers(0, 2))

        def lane_oracle(l):
            oc = oracle.Circuit(text)
            for k, e in enumerate(varied):
                oc.set_element_value(e, vals[k, l])
            return oc

        sweep = np.array([0.3, 1.7, -0.8])
        cases = [("DC", api.MODE_DC, dict(sweep_element=kc.element_index("V1")), sweep, lambda oc, l: oc.solve_dc("V1", sweep[l], sweep[l], 1.0, warm_start=False)[0])]
        if not diode:
            f = np.array([10.0, 1e4, 1e7])
            cases.append(("AC", api.MODE_AC, dict(nominal_frequency=1e4), f, lambda oc, l: oc.solve_ac([f[l]])[1][0]))
        for what, mode, kw, param, ref_of in cases:
            try:
                plan = kb.Plan(kc, mode, varied=varied, force_lane_vm=force, **kw)
            except kb.KretsError:
                continue
            si, fi, iov = input_layout(plan)
            IN = np.zeros((plan.info.n_inputs, 3))
            IN[si if mode == api.MODE_DC else fi] = param
            for k, pl in enumerate(iov):
                IN[pl] = vals[k]
            X, st = solve(plan, IN, mode == api.MODE_AC)
            for l in range(3):
                ref, _ = _oracle(lambda: ref_of(lane_oracle(l), l))
                if ref is None and what == "AC" and _oracle(lane_oracle(l).solve_op)[0] is None:
                    continue      # ac.rs:25: the reference's AC analysis fails in its operating point (a V / L loop); the plan-level AC solve has no OP
                if (ref is None) != bool(st[l]):
                    bad[(seed, what, l)] = ("solved / failed mismatch", st[l], text)
                elif ref is not None:
                    done += 1
                    if not rel_err(X[l], ref) < (1e-6 if diode else 1e-8):
                        bad[(seed, what, l)] = (rel_err(X[l], ref), plan.info.executor, text)
    assert not bad, bad
    assert done >= 400


def test_random_diode_circuits_sequential_dc_sweep_vs_the_reference_order(oracle):
    """krb_solve_dc_sweep_sequential's launch schedule (vm_emulator.run_sequential_sweep) on random diode circuits: 13 sweep points in
    order, each warm-started from its predecessor, against the oracle's restatement of dc.rs:47-113 (warm_start = True)."""
    from vm_emulator import run_sequential_sweep
    bad, done = {}, 0
    for seed in range(8000, 8100):
        text = random_circuit(np.random.default_rng(seed), True)
        if "DMOD" not in text:
            continue
        kc, oc = kb.parse_circuit_description(text), oracle.Circuit(text)
        start, step, num = -1.0, 0.25, 13
        ref, _ = _oracle(lambda: oc.solve_dc("V1", start, start + (num - 1) * step, step, warm_start=True))
        try:
            plan = kb.Plan(kc, api.MODE_DC, sweep_element=kc.element_index("V1"))
            X, em, _ = run_sequential_sweep(Emulator, plan, start, step, num)
            failed = bool(em.status.any())
        except kb.KretsError:
            failed = True
        if (ref is None) != failed:
            bad[seed] = ("solved / failed mismatch", text)
        elif ref is not None:
            done += 1
            if not rel_err(X[:, : kc.n, 0], ref) < 1e-7:
                bad[seed] = (rel_err(X[:, : kc.n, 0], ref), text)
    assert not bad, bad
    assert done >= 20


# ---- transient with SIN / PULSE sources (per-row source table, the reference's source lag) -------------------------------------------
# Known limit: a source that jumps by volts within one time step across a diode drives the reference's un-limited
# Newton-Raphson (diode.rs:65-70: the clamp is a no-op) through conductances of 1e20 S and more; there the Schur complements cancel
# completely, the answer of ANY elimination order is rounding noise, and a static order can land on an exactly zero pivot (NaN, the lane
# is flagged) where a per-factorisation pivot search happens to step around it.  Seeds of this generator that do so:
WAVEFORM_KNOWN_FLAGGED = {9027, 9139}        # (9139 lies beyond the range the test runs by default)


def test_random_circuits_transient_with_waveforms(oracle):
    import re
    bad, done = {}, 0
    for seed in range(9000, 9060):
        rng = np.random.default_rng(seed - 9000)
        diode = seed % 2 == 1
        text = random_circuit(rng, diode)
        if rng.random() < 0.5:
            w = (f"SIN({rng.uniform(-1, 1):.3g} {rng.uniform(0.1, 3):.3g} {10 ** rng.uniform(3, 5.5):.4g} {rng.choice([0, 2e-6, 5e-6])} "
                 f"{rng.choice([0, 1e4])} {rng.choice([0, 30, 90])})")
        else:
            w = (f"PULSE({rng.uniform(-1, 1):.3g} {rng.uniform(0.5, 3):.3g} {rng.choice([0, 1e-6, 3e-6])} {rng.choice([1e-6, 2e-6, 1e-9])} "
                 f"{rng.choice([1e-6, 3e-6])} {rng.choice([2e-6, 5e-6])} {rng.choice([8e-6, 2e-5])})")
        text = re.sub(r"^V1 1 0 .*$", f"V1 1 0 {w}", text, flags=re.M)
        oc, kc = oracle.Circuit(text), kb.parse_circuit_description(text)
        ref, _ = _oracle(lambda: oc.solve_tran(1e-6, 25e-6)[1])
        try:
            X, st, _ = run_transient(kc, 1e-6, 25)
            failed = bool(st.any())
        except kb.KretsError:
            failed = True
        if (ref is None) != failed:
            bad[seed] = ("solved / failed mismatch", text)
        elif ref is not None:
            done += 1
            if not rel_err(X, ref) < (1e-6 if diode else 1e-7):
                bad[seed] = (rel_err(X, ref), text)
    assert set(bad) <= WAVEFORM_KNOWN_FLAGGED, {k: v for k, v in bad.items() if k not in WAVEFORM_KNOWN_FLAGGED}
    assert done >= 45
