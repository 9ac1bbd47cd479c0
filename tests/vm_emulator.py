"""numpy emulator of the krets_b200 lane program (vm_ops.h) — TEST INFRASTRUCTURE ONLY.

Lets the CPU-only test tier check the host plan compiler (stamps -> pivot plan -> symbolic LU -> scratch allocation)
against the oracle without a GPU: every op is lane-invariant, so numpy executes each op on the whole lane vector.
It mirrors krb_lane_vm in krets_b200/csrc/vm_kernel.cuh; the product never imports it.
"""
import ctypes as C

import numpy as np

from krets_b200 import _lib

(END, LDB, LDBS, LDI, LDG, STG, LDX, XOUT, ZERO, MOV, NEG, ADD, SUB, MUL, FNMA, RCP, INV, DIVR, MULJ, DIOG, DIOI, LDK, NOP, WAIT, MULG, MOSGM, MOSGDS, MOSID) = range(28)
ZERO_PIVOT, NOT_CONVERGED, NONFINITE, INACCURATE = 1, 2, 4, 8
GROWTH_LIMIT = 2.0 ** 24   # KRB_GROWTH_LIMIT_LOG2


def _plan_tables(plan, rows, h):
    L = _lib.lib()
    lag = 1 if int(L.krb_plan_integrator(plan._h)) == 0 else 0     # only the reference integrator lags the source
    npool = int(L.krb_plan_pool_size(plan._h))
    pool = np.zeros(max(npool, 1))
    L.krb_plan_eval_pool(plan._h, pool.ctypes.data_as(C.POINTER(C.c_double)))
    nw = int(L.krb_plan_wave_count(plan._h))
    bstep = np.zeros((max(rows, 1), 2, max(nw, 1)))
    for r in range(1, rows):
        for k in range(nw):
            v = C.c_double()
            L.krb_plan_wave_value(plan._h, k, (r - lag) * h, C.byref(v))
            bstep[r, 0, k] = v.value
            L.krb_plan_wave_value(plan._h, k, r * h, C.byref(v))
            bstep[r, 1, k] = v.value
    return pool, bstep


def input_layout(plan):
    L = _lib.lib()
    a, b = C.c_int32(), C.c_int32()
    arr = (C.c_int32 * max(len(plan.varied), 1))()
    L.krb_plan_input_layout(plan._h, C.byref(a), C.byref(b), arr)
    return a.value, b.value, list(arr)[: len(plan.varied)]


class Emulator:
    def __init__(self, plan, lanes, *, rows=1, h=0.0, cfg=None, nonlinear=None, cplx=None):
        self.plan, self.L, self.rows, self.h = plan, lanes, rows, h
        self.n = plan.n
        self.width = plan.n + max(int(_lib.lib().krb_plan_state_columns(plan._h)), 0)
        self.cplx = (plan.mode == 1) if cplx is None else cplx
        self.dtype = np.complex128 if self.cplx else np.float64
        self.pool, self.bstep = _plan_tables(plan, rows, h)
        self.nonlinear = bool(plan.info.nonlinear) if nonlinear is None else nonlinear
        self.cfg = cfg or dict(reltol=1e-3, iabstol=1e-12, vabstol=1e-6, max_iter=300)
        self.split = bool(_lib.lib().krb_plan_has_step_sections(plan._h))
        self.progs = [plan.program(s) for s in range(3)]
        self.scr, self.G = {}, {}
        self.status = np.zeros(lanes, dtype=np.int32)
        self.iters = np.zeros(lanes, dtype=np.int32)
        self.stats = dict(ldg=0, stg=0, ops=0)
        self.pending = set()      # scratch slots with an asynchronous load in flight (must not be touched before WAIT)

    def run_prog(self, prog, IN, X, row, it, guess_empty, guess, prev, active):
        scr, G = self.scr, self.G
        notconv = np.zeros(self.L, dtype=bool)
        with np.errstate(all="ignore"):
            pcount = len(prog)
            pci = 0
            while pci < pcount:
                w0, imm = int(prog[pci][0]), int(prog[pci][1])
                pci += 1
                opc, d, a, b = w0 & 0xFF, (w0 >> 8) & 0xFF, (w0 >> 16) & 0xFF, (w0 >> 24) & 0xFF
                self.stats["ops"] += 1
                if opc in (ADD, SUB, MUL, MULG, FNMA, DIVR, MOSGM, MOSGDS, MOSID):
                    touched = {d, a, b}
                elif opc in (MOV, NEG, RCP, INV, MULJ, DIOG, DIOI):
                    touched = {d, a}
                elif opc in (STG, XOUT):
                    touched = {a}
                elif opc in (LDB, LDBS, LDI, LDG, LDX, ZERO, LDK):
                    touched = {d}
                else:
                    touched = set()
                assert not (touched & self.pending), f"op {opc} touches a slot with an async load in flight (missing WAIT)"
                if opc in (LDBS, LDI, LDG) or (opc == LDX and not ((b & 1) and guess_empty)):
                    self.pending.add(d)
                if opc == END:
                    assert not self.pending, "program ended with async loads in flight"
                    break
                elif opc == LDK:
                    assert (pci - 1) % 32 != 31, "LDK must not sit at the end of a 32-op chunk"
                    bits = np.array([int(prog[pci][0]) | (int(prog[pci][1]) << 32)], dtype=np.uint64)
                    pci += 1
                    scr[d] = np.full(self.L, bits.view(np.float64)[0], dtype=self.dtype)
                elif opc == NOP:
                    pass
                elif opc == WAIT:
                    self.pending.clear()
                elif opc == LDB:
                    scr[d] = np.full(self.L, self.pool[imm], dtype=self.dtype)
                elif opc == LDBS:
                    scr[d] = np.full(self.L, self.bstep[row, 1 if it > 0 else 0, imm], dtype=self.dtype)
                elif opc == LDI:
                    scr[d] = IN[imm].astype(self.dtype)
                elif opc == LDG:
                    scr[d] = G[imm].copy(); self.stats["ldg"] += 1
                elif opc == STG:
                    G[imm] = scr[a].copy(); self.stats["stg"] += 1
                elif opc == LDX:
                    if b & 1:
                        v = np.full(self.L, 0.5 if (b & 2) else 0.0) if guess_empty else guess[imm]
                    else:
                        v = prev[imm]
                    scr[d] = np.asarray(v).real.astype(self.dtype)
                elif opc == XOUT:
                    v = scr[a]
                    self.status[~np.isfinite(v)] |= NONFINITE
                    if (not self.cplx) and self.nonlinear and not guess_empty and not (b & 4):
                        ov = guess[imm]
                        diff = np.abs(v - ov)
                        scale = np.maximum(np.abs(v), np.abs(ov))
                        atol = self.cfg["iabstol"] if (b & 1) else self.cfg["vabstol"]
                        notconv |= ~(diff <= self.cfg["reltol"] * scale + atol)
                    X[row][imm] = np.where(active, v, X[row][imm])
                elif opc == ZERO:
                    scr[d] = np.zeros(self.L, dtype=self.dtype)
                elif opc == MOV:
                    scr[d] = scr[a].copy()
                elif opc == NEG:
                    scr[d] = -scr[a]
                elif opc == ADD:
                    scr[d] = scr[a] + scr[b]
                elif opc == SUB:
                    scr[d] = scr[a] - scr[b]
                elif opc == MUL:
                    scr[d] = scr[a] * scr[b]
                elif opc == MULG:                    # u^ = u / pivot: growth-tracked
                    scr[d] = scr[a] * scr[b]
                    v = scr[d]
                    mag = np.maximum(np.abs(v.real), np.abs(v.imag)) if np.iscomplexobj(v) else np.abs(v)
                    if not self.nonlinear:           # NR convergence, not pivot growth, guards nonlinear circuits (vm_kernel.cuh)
                        self.status[np.isfinite(mag) & (mag >= GROWTH_LIMIT)] |= INACCURATE
                elif opc == FNMA:
                    scr[d] = scr[d] - scr[a] * scr[b]
                elif opc == RCP:
                    bad = ~(np.abs(scr[a]) > 0) | ~np.isfinite(scr[a])
                    self.status[bad] |= ZERO_PIVOT
                    scr[d] = 1.0 / scr[a]
                elif opc == INV:
                    scr[d] = (1.0 / scr[a].real).astype(self.dtype)
                elif opc == DIVR:
                    scr[d] = (scr[a].real / scr[b].real).astype(self.dtype)
                elif opc == MULJ:
                    scr[d] = (1j * scr[a].real).astype(self.dtype)
                elif opc in (DIOG, DIOI):
                    vd = scr[a].real
                    isnvt, nvt, is_, vcrit = self.pool[imm: imm + 4]
                    vl = np.minimum(np.maximum(vd, -vcrit), vcrit)
                    ex = np.exp(vl / nvt)
                    g = isnvt * ex
                    scr[d] = (g if opc == DIOG else (is_ * (ex - 1.0) - g * vd)).astype(self.dtype)
                elif opc in (MOSGM, MOSGDS, MOSID):     # level-1 NMOS, nmosfet.rs:51-102 (extension)
                    vgs, vds = scr[a].real, scr[b].real
                    vth, beta, lam = self.pool[imm: imm + 3]
                    vov = vgs - vth
                    on = vgs > vth
                    lin = on & (vds >= 0.0) & (vds <= vov)
                    sat = on & ~lin & (vds >= vov) & (vds >= 0.0)
                    if opc == MOSGM:
                        v = np.where(lin, beta * vds, np.where(sat, (beta * vov) * (1.0 + lam * vds), 0.0))
                    elif opc == MOSGDS:
                        v = np.where(lin, beta * (vov - vds), np.where(sat, ((beta / 2.0) * lam) * (vov * vov), 0.0))
                    else:
                        v = np.where(lin, beta * (vov * vds - (vds * vds) / 2.0), np.where(sat, ((beta / 2.0) * (vov * vov)) * (1.0 + lam * vds), 0.0))
                    scr[d] = v.astype(self.dtype)
                else:
                    raise ValueError(f"bad opcode {opc}")
        return notconv

    def run(self, IN, X=None, row_begin=0, row_end=None, guess_empty_at_iter0=True, pseudo_time=False, Xprev0=None):
        """IN: [n_inputs][L]; X: preallocated [rows][n][L] (transient continuation) or None."""
        row_end = self.rows if row_end is None else row_end
        if X is None:
            X = np.zeros((self.rows, self.width, self.L), dtype=self.dtype)
        max_iter = self.cfg["max_iter"] if self.nonlinear and not self.cplx else 1
        for row in range(row_begin, row_end):
            prev = Xprev0 if (row == row_begin and Xprev0 is not None) else X[max(row - 1, 0)]
            active = np.ones(self.L, dtype=bool)
            passes = [1, 2] if (self.split and row == row_begin) else ([2] if self.split else [0])
            for sec in passes:
                solve_pass = sec != 1
                for it in range(max_iter):
                    guess_empty = guess_empty_at_iter0 and it == 0
                    guess = prev if (it == 0 and not guess_empty_at_iter0) else X[row].copy()
                    notconv = self.run_prog(self.progs[sec], IN, X, row, it, guess_empty, guess, prev, active)
                    if not solve_pass or not self.nonlinear or self.cplx:
                        break
                    pseudo_ok = True
                    if pseudo_time and it == 0:
                        tn, tp = row * self.h, (row - 1) * self.h
                        pseudo_ok = abs(tn - tp) <= self.cfg["reltol"] * max(abs(tn), abs(tp)) + self.cfg["vabstol"]
                    self.iters[active] = it + 1
                    conv = (not guess_empty) & (~notconv) & pseudo_ok
                    active = active & ~conv
                    if it == max_iter - 1:
                        self.status[active] |= NOT_CONVERGED
                    if not active.any():
                        break
        return X


def run_sequential_sweep(driver, plan, start, step, num, lanes=1, IN_extra=None):
    """The launch schedule of krb_solve_dc_sweep_sequential (runtime.cu run_lanes, `sweep_rows`) on a CPU driver (Emulator or
    jit_host.JitHost): one launch per sweep point, the sweep element's input plane filled with start + r*step (dc.rs:48), row 0 from the
    empty guess (dc.rs:41), row r warm-started from row r-1 (dc.rs:58,112), the `step` pseudo-unknown as |r h - (r-1) h| with h = 1."""
    i_sweep, _, _ = input_layout(plan)
    IN = np.zeros((i_sweep + 1, lanes)) if IN_extra is None else np.array(IN_extra, dtype=np.float64)
    em = driver(plan, lanes, rows=num, h=1.0)
    X = None
    its = []
    for r in range(num):
        IN[i_sweep, :] = start + (float(r) * step)
        X = em.run(IN, X=X, row_begin=r, row_end=r + 1, guess_empty_at_iter0=(r == 0), pseudo_time=True)
        its.append(em.iters.copy())
    return X, em, np.array(its)


# ---- chain executor term stream (chain_kernel.cuh) --------------------------------------------------------------
def run_chain(plan, IN, cplx):
    """numpy restatement of krb_chain_kernel (row-info words + 16-byte record stream): IN [n_inputs][L] -> X [n][L]."""
    L_ = _lib.lib()
    nw = int(L_.krb_plan_chain_size(plan._h))
    assert nw > 0, "plan has no chain tables"
    n = plan.n
    rowinfo = np.zeros(n + 1 + 1100, dtype=np.uint32)
    words = np.zeros(nw, dtype=np.uint64)
    unk = np.zeros(n, dtype=np.int32)
    L_.krb_plan_get_chain(plan._h, rowinfo.ctypes.data_as(C.POINTER(C.c_uint32)), words.ctypes.data_as(C.POINTER(C.c_uint64)),
                          unk.ctypes.data_as(C.POINTER(C.c_int32)))
    dwords = words.view(np.float64)
    lanes = IN.shape[1]
    sweep, freq, _ = input_layout(plan)
    w = (2.0 * np.pi) * IN[freq] if (cplx and freq >= 0) else np.zeros(lanes)
    dt = np.complex128 if cplx else np.float64
    U1 = np.zeros((n, lanes), dtype=dt)
    U2 = np.zeros((n, lanes), dtype=dt)
    Y = np.zeros((n, lanes), dtype=dt)
    pu1 = np.zeros(lanes, dtype=dt)
    pu2 = np.zeros(lanes, dtype=dt)
    py = np.zeros(lanes, dtype=dt)
    status = np.zeros(lanes, dtype=np.int32)
    has_u2 = np.zeros(n, dtype=bool)
    fix_units = 4 if cplx else 2
    i = 0          # position in 16-byte units
    n_fixed = 0
    with np.errstate(all="ignore"):
        for k in range(n):
            info = int(rowinfo[k])
            e = [np.zeros(lanes, dtype=dt) for _ in range(5)]
            has_u2[k] = bool(info & 2)
            if info & 1:
                d = dwords[2 * i: 2 * (i + fix_units)]
                n_fixed += 1
                if cplx:
                    e[0] = d[0] + 1j * (w * d[4]); e[1] = d[1] + 1j * (w * d[5]); e[2] = d[2] + 1j * (w * d[6])
                    e[4] = np.full(lanes, d[3], dtype=dt)
                else:
                    e[0] = np.full(lanes, d[0]); e[1] = np.full(lanes, d[1]); e[2] = np.full(lanes, d[2]); e[4] = np.full(lanes, d[3])
                i += fix_units
            else:
                nt = info >> 8
                for t in range(nt):
                    mw = int(words[2 * (i + t)])
                    meta, plane, K = mw & 0xFFFFFFFF, mw >> 32, float(dwords[2 * (i + t) + 1])
                    kind = (meta >> 4) & 0xF
                    if kind == 0:
                        v = np.full(lanes, K, dtype=dt)
                    elif kind == 1:
                        v = IN[plane].astype(dt)
                    elif kind == 2:
                        v = (1.0 / IN[plane]).astype(dt)
                    elif kind == 3:
                        v = (IN[plane] / K).astype(dt)
                    elif kind == 4:
                        v = 1j * (w * K)
                    elif kind == 5:
                        v = 1j * (w * IN[plane])
                    else:
                        raise ValueError(kind)
                    if meta & (1 << 8):
                        v = -v
                    e[meta & 7] = e[meta & 7] + v
                i += nt
            wk = e[1] - e[0] * pu1
            s1 = e[2] - e[0] * pu2
            yk = e[4] - e[0] * py
            inv = 1.0 / wk
            status[~np.isfinite(inv)] |= ZERO_PIVOT
            pu1, pu2, py = s1 * inv, e[3] * inv, yk * inv
            assert has_u2[k] or not np.any(pu2 != 0), "u2 plane flag missing"
            U1[k], U2[k], Y[k] = pu1, pu2, py
        X = np.zeros((n, lanes), dtype=dt)
        x1 = np.zeros(lanes, dtype=dt)
        x2 = np.zeros(lanes, dtype=dt)
        for k in range(n - 1, -1, -1):
            x = Y[k] - U1[k] * x1
            if has_u2[k]:
                x = x - U2[k] * x2
            X[unk[k]] = x
            x2, x1 = x1, x
    assert 2 * i == nw - 2200, "record stream must end with the prefetch padding"
    run_chain.last_fixed_rows = n_fixed
    return X, status


# ---- block-partitioned chain kernel (chain_block_kernel.cuh) ------------------------------------------------------
def chain_block_info(plan):
    """(valid, T, M, n_dict, nominal_diff, why_not) of the plan's block-partitioned chain tables."""
    L_ = _lib.lib()
    L_.krb_plan_chain_block_info.restype = C.c_int32
    L_.krb_plan_chain_block_info.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_double)]
    info = (C.c_int32 * 4)()
    nd = C.c_double(0.0)
    rc = L_.krb_plan_chain_block_info(plan._h, info, C.byref(nd))
    assert rc == 0, "plan has no chain tables"
    why = "" if info[0] else L_.krb_last_error_string().decode()
    return bool(info[0]), int(info[1]), int(info[2]), int(info[3]), float(nd.value), why


def run_chain_block(plan, IN, cplx):
    """numpy restatement of krb_chain_block_kernel: Wang partition (T blocks of M rows), one cyclic-reduction level + parallel
    cyclic reduction on the odd separators, interiors by back substitution.  IN [n_inputs][L] -> X [n][L] (unknown order), status [L]."""
    L_ = _lib.lib()
    valid, T, M, n_dict, _, why = chain_block_info(plan)
    assert valid, f"block tables not valid: {why}"
    L_.krb_plan_get_chain_block.restype = C.c_int32
    L_.krb_plan_get_chain_block.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]
    n = plan.n
    fw = 8 if cplx else 4
    rows = np.zeros(2 * T * M, dtype=np.uint32)
    dict_w = np.zeros(max(n_dict * fw, 1), dtype=np.uint64)
    assert L_.krb_plan_get_chain_block(plan._h, rows.ctypes.data_as(C.POINTER(C.c_uint32)), dict_w.ctypes.data_as(C.POINTER(C.c_uint64))) == 0
    rows = rows.reshape(M, T, 2)
    dict_d = dict_w.view(np.float64)
    nw = int(L_.krb_plan_chain_size(plan._h))
    words = np.zeros(nw, dtype=np.uint64)
    unk = np.zeros(n, dtype=np.int32)
    L_.krb_plan_get_chain(plan._h, None, words.ctypes.data_as(C.POINTER(C.c_uint64)), unk.ctypes.data_as(C.POINTER(C.c_int32)))
    dwords = words.view(np.float64)
    lanes = IN.shape[1]
    _, freq, _ = input_layout(plan)
    w = (2.0 * np.pi) * IN[freq] if (cplx and freq >= 0) else np.zeros(lanes)
    dt = np.complex128 if cplx else np.float64
    E = np.zeros((5, T, M, lanes), dtype=dt)          # sub, diag, sup1, sup2, rhs of row t*M + j
    has_u2 = np.zeros((T, M), dtype=bool)
    slots = []
    for t in range(T):
        for j in range(M):
            w0, w1 = int(rows[j, t, 0]), int(rows[j, t, 1])
            if w0 & 1:
                d = dict_d[(w0 >> 1) * fw: (w0 >> 1) * fw + fw]
                if cplx:
                    E[0, t, j] = d[0] + 1j * (w * d[4]); E[1, t, j] = d[1] + 1j * (w * d[5]); E[2, t, j] = d[2] + 1j * (w * d[6]); E[4, t, j] = d[3]
                else:
                    E[0, t, j] = d[0]; E[1, t, j] = d[1]; E[2, t, j] = d[2]; E[4, t, j] = d[3]
            else:
                i, nt = w0 >> 1, (w1 >> 2) & 0x3FFF           # word 1: bit 1 HAS_U2, bits 2..15 #terms, bits 16.. slot
                has_u2[t, j] = bool(w1 & 2)
                slots.append((t * M + j, w1 >> 16))
                assert t * M + j < n
                for q in range(nt):
                    mw = int(words[2 * (i + q)])
                    meta, plane, K = mw & 0xFFFFFFFF, mw >> 32, float(dwords[2 * (i + q) + 1])
                    kind = (meta >> 4) & 0xF
                    if kind == 0: v = np.full(lanes, K, dtype=dt)
                    elif kind == 1: v = IN[plane].astype(dt)
                    elif kind == 2: v = (1.0 / IN[plane]).astype(dt)
                    elif kind == 3: v = (IN[plane] / K).astype(dt)
                    elif kind == 4: v = 1j * (w * K)
                    elif kind == 5: v = 1j * (w * IN[plane])
                    else: raise ValueError(kind)
                    if meta & (1 << 8):
                        v = -v
                    E[meta & 7, t, j] = E[meta & 7, t, j] + v
    if len(slots) < 65536:                             # slot = rank among the GENERIC rows (shared-dictionary mode)
        assert [sl for _, sl in sorted(slots)] == list(range(len(slots)))
    for t in range(T):                                 # layout rules the kernel relies on
        js = np.nonzero(has_u2[t])[0]
        assert len(js) <= 1 and all(j <= M - 3 for j in js), "u2 row next to a block boundary / two per block"
    assert not np.any(E[3][~has_u2] != 0)
    sub, diag, sup, sup2, rhs = E
    status = np.zeros(lanes, dtype=np.int32)
    with np.errstate(all="ignore"):
        u = np.zeros((T, M - 1, lanes), dtype=dt); u2 = np.zeros_like(u); y = np.zeros_like(u); l = np.zeros_like(u)
        for j in range(M - 1):
            if j == 0:
                wk, s1, yk, lk = diag[:, 0], sup[:, 0], rhs[:, 0], sub[:, 0]
            else:
                wk = diag[:, j] - sub[:, j] * u[:, j - 1]
                s1 = sup[:, j] - sub[:, j] * u2[:, j - 1]
                yk = rhs[:, j] - sub[:, j] * y[:, j - 1]
                lk = -sub[:, j] * l[:, j - 1]
            inv = 1.0 / wk
            status[~np.isfinite(inv).all(axis=0)] |= ZERO_PIVOT
            u[:, j], u2[:, j], y[:, j], l[:, j] = s1 * inv, sup2[:, j] * inv, yk * inv, lk * inv
        A = -sub[:, M - 1] * l[:, M - 2]
        Bp = diag[:, M - 1] - sub[:, M - 1] * u[:, M - 2]
        Dp = rhs[:, M - 1] - sub[:, M - 1] * y[:, M - 2]
        p = np.zeros((T, M + 1, lanes), dtype=dt); r = np.zeros_like(p); q = np.zeros_like(p)
        r[:, M - 1] = -1.0
        for j in range(M - 2, -1, -1):
            p[:, j] = y[:, j] - u[:, j] * p[:, j + 1] - u2[:, j] * p[:, j + 2]
            r[:, j] = -u[:, j] * r[:, j + 1] - u2[:, j] * r[:, j + 2]
            q[:, j] = l[:, j] - u[:, j] * q[:, j + 1] - u2[:, j] * q[:, j + 2]
        # the kernel accumulates the first row's relation x_0 = P - R s_t - Q s_{t-1} top-down (x_0 = accP - accQ s_{t-1}
        # + fA x_j + fB x_{j+1}, row j substituted); the bottom-up recurrences above are the independent check
        accP = np.zeros((T, lanes), dtype=dt); accQ = np.zeros_like(accP); fA = np.ones_like(accP); fB = np.zeros_like(accP)
        for j in range(M - 1):
            accP = accP + fA * y[:, j]
            accQ = accQ + fA * l[:, j]
            fA, fB = fB - fA * u[:, j], -fA * u2[:, j]
        assert not np.any(fB != 0)
        scale = max(1e-300, float(np.abs(p[:, 0]).max()), float(np.abs(q[:, 0]).max()), float(np.abs(r[:, 0]).max()))
        assert np.abs(accP - p[:, 0]).max() <= 1e-9 * scale and np.abs(accQ - q[:, 0]).max() <= 1e-9 * scale and np.abs(-fA - r[:, 0]).max() <= 1e-9 * scale
        def nxt(v):
            o = np.zeros_like(v); o[:-1] = v[1:]; return o
        ss = sup[:, M - 1]
        Bq = Bp - ss * nxt(accQ); Cq = -ss * nxt(-fA); Dq = Dp - ss * nxt(accP)
        inv = 1.0 / Bq
        status[~np.isfinite(inv).all(axis=0)] |= ZERO_PIVOT
        a, c, d = A * inv, Cq * inv, Dq * inv
        def sh(v, k):
            o = np.zeros_like(v)
            if k > 0: o[k:] = v[:-k]
            else: o[:k] = v[-k:]
            return o
        def reduce(a, c, d, am, cm, dm, ap, cp, dp):
            inv = 1.0 / (1.0 - a * cm - c * ap)
            status[~np.isfinite(inv).all(axis=0)] |= ZERO_PIVOT
            return (-a * am) * inv, (-c * cp) * inv, (d - a * dm - c * dp) * inv
        # one cyclic-reduction level: the odd rows 2k+1 absorb rows 2k and 2k+2 ...
        ae, ce, de = a[0::2], c[0::2], d[0::2]
        a1, c1, d1 = reduce(a[1::2], c[1::2], d[1::2], ae, ce, de, sh(ae, -1), sh(ce, -1), sh(de, -1))
        # ... parallel cyclic reduction on the T/2 odd rows ...
        s = 1
        while s < T // 2:
            a1, c1, d1 = reduce(a1, c1, d1, sh(a1, s), sh(c1, s), sh(d1, s), sh(a1, -s), sh(c1, -s), sh(d1, -s))
            s *= 2
        # ... even rows by substitution
        sep = np.zeros_like(d)
        sep[1::2] = d1
        sep[0::2] = de - ae * sh(d1, 1) - ce * d1
        d = sep
        sprev = np.zeros_like(d); sprev[1:] = d[:-1]
        # interiors: back substitution x_j = (y^_j - l^_j s_{t-1}) - u^_j x_{j+1} - u2^_j x_{j+2}, x_{M-1} = s_t
        x = np.zeros((T, M + 1, lanes), dtype=dt)
        x[:, M - 1] = d
        x[:, M] = 0.0
        for j in range(M - 2, -1, -1):
            x2 = x[:, j + 2] if j + 2 <= M - 1 else 0.0
            x[:, j] = (y[:, j] - l[:, j] * sprev) - u[:, j] * x[:, j + 1] - u2[:, j] * x2
        x = x[:, :M]
    xr = x.reshape(T * M, lanes)
    assert not np.any(xr[n:] != 0), "padding rows must stay zero"
    X = np.zeros((n, lanes), dtype=dt)
    X[unk] = xr[:n]
    status[~np.isfinite(d).all(axis=0)] |= NONFINITE
    return X, status
