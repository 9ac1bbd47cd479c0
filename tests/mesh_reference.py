"""Independent CPU reference for RLC meshes at sizes the oracle's natural-order LU cannot factor in seconds — TEST INFRASTRUCTURE ONLY.

The matrices are the ORACLE's stamps (oracle.Circuit.stamp_matrix: stampable.rs restated); the factorisation is SuperLU (scipy) with
its own ordering and partial pivoting; the trapezoidal right-hand sides are restated here from the element table and are themselves
held to the oracle's trapezoidal driver on a small mesh by tests/test_mesh_reference_cpu.py.  The product never imports this module.
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spl

KIND_R, KIND_C, KIND_L, KIND_V, KIND_I = 0, 1, 2, 3, 4


class ElementTable:
    def __init__(self, oc):
        ne = oc.num_elements
        self.kind = np.array([oc.element_kind(e) for e in range(ne)])
        idx = np.array([oc.element_indices(e) for e in range(ne)])
        self.ip, self.im, self.ic = idx[:, 0], idx[:, 1], idx[:, 2]

    def values(self, oc):
        return np.array([oc.element_value(e) for e in range(len(self.kind))])


def csc(oc, mode, w):
    colptr, rowidx, v = oc.stamp_matrix(mode, w)
    return sp.csc_matrix((v, rowidx, colptr), shape=(oc.n, oc.n))


def _dv(x, ip, im):
    """v(plus) - v(minus) with ground = 0"""
    return np.where(ip >= 0, x[np.maximum(ip, 0)], 0.0) - np.where(im >= 0, x[np.maximum(im, 0)], 0.0)


def _add(b, idx, v):
    m = idx >= 0
    np.add.at(b, idx[m], v[m])


def dc_rhs(oc, tab, val):
    b = np.zeros(oc.n)
    src = (tab.kind == KIND_V) | (tab.kind == KIND_I)
    b[tab.ic[src]] = val[src]                         # stampable.rs:1016-1018, :566-569
    return b


def ac_rhs(oc, tab, val, ac_amp_of_v=1.0):
    b = np.zeros(oc.n, dtype=complex)
    b[tab.ic[tab.kind == KIND_V]] = ac_amp_of_v       # stampable.rs:1030-1032
    b[tab.ic[tab.kind == KIND_I]] = val[tab.kind == KIND_I]
    return b


def trap_rows(oc, tab, h, steps):
    """rows 0..steps of the trapezoidal transient (KRB_INTEGRATOR_TRAP): row 0 = operating point, companions g = 2C/h with the
    capacitor current carried as state (i_C(0) = 0), inductor row v - (2L/h) i = -(2L/h) i_prev - v_prev, sources at t_n."""
    val = tab.values(oc)
    x = spl.splu(csc(oc, "dc", 0.0)).solve(dc_rhs(oc, tab, val))
    lu = spl.splu(csc(oc, "tran", h / 2))
    C, Lm, V, I = (tab.kind == k for k in (KIND_C, KIND_L, KIND_V, KIND_I))
    gC, gL = val[C] / (h / 2), val[Lm] / (h / 2)
    iC = np.zeros(int(C.sum()))
    v_idx = np.nonzero(V)[0]
    rows = [x]
    for r in range(1, steps + 1):
        b = np.zeros(oc.n)
        for e in v_idx:
            b[tab.ic[e]] = oc.transient_value_at(int(e), r * h)
        b[tab.ic[I]] = val[I]
        dvC = _dv(x, tab.ip[C], tab.im[C])
        ieq = gC * dvC + iC
        _add(b, tab.ip[C], ieq)
        _add(b, tab.im[C], -ieq)
        b[tab.ic[Lm]] = -gL * x[tab.ic[Lm]] - _dv(x, tab.ip[Lm], tab.im[Lm])
        xn = lu.solve(b)
        iC = gC * (_dv(xn, tab.ip[C], tab.im[C]) - dvC) - iC
        x = xn
        rows.append(x)
    return np.array(rows)
