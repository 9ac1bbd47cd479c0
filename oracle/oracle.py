"""ctypes binding of the CPU oracle (oracle/libkrets_oracle.so).

ORACLE — TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; the krets_b200 package must never import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libkrets_oracle.so")

ERR_NAMES = {
    0: "Ok", 1: "ParseError", 2: "EmptyNetlist", 3: "ElementNotFound", 4: "InvalidElementFormat",
    5: "MaximumIterationsExceeded", 6: "DecompositionFailed", 7: "Unsupported", 8: "Buffer", 9: "MissingModel",
}


class OracleError(RuntimeError):
    def __init__(self, code, msg=""):
        super().__init__(f"{ERR_NAMES.get(code, code)}: {msg}")
        self.code = code
        self.kind = ERR_NAMES.get(code, str(code))


class _Config(C.Structure):
    _fields_ = [("relative_tolerance", C.c_double), ("current_absolute_tolerance", C.c_double),
                ("voltage_absolute_tolerance", C.c_double), ("maximum_iterations", C.c_int)]


def build(force: bool = False) -> str:
    """Compile the oracle with the committed recipe (oracle/Makefile).  Building the checker is not using it."""
    src_newer = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB_PATH)
        for f in ("krets_oracle.c", "krets_oracle.h", "ko_lu_template.h"))
    if force or src_newer:
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        L.ko_parse.restype = C.c_void_p
        L.ko_parse.argtypes = [C.c_char_p, ip, C.c_char_p, C.c_int]
        L.ko_free.argtypes = [C.c_void_p]
        L.ko_num_unknowns.argtypes = [C.c_void_p]
        L.ko_num_elements.argtypes = [C.c_void_p]
        L.ko_unknown_name.restype = C.c_char_p
        L.ko_unknown_name.argtypes = [C.c_void_p, C.c_int]
        L.ko_index_of.argtypes = [C.c_void_p, C.c_char_p]
        L.ko_element_index.argtypes = [C.c_void_p, C.c_char_p]
        L.ko_element_ident.restype = C.c_char_p
        L.ko_element_ident.argtypes = [C.c_void_p, C.c_int]
        L.ko_element_kind.argtypes = [C.c_void_p, C.c_int]
        L.ko_element_value.restype = C.c_double
        L.ko_element_value.argtypes = [C.c_void_p, C.c_int]
        L.ko_set_element_value.argtypes = [C.c_void_p, C.c_int, C.c_double]
        L.ko_element_indices.argtypes = [C.c_void_p, C.c_int, ip, ip, ip]
        L.ko_is_nonlinear.argtypes = [C.c_void_p]
        L.ko_set_device_models.argtypes = [C.c_int]
        L.ko_set_device_models.restype = None
        L.ko_stamp_matrix.argtypes = [C.c_void_p, C.c_int, C.c_double, dp, ip, ip, dp, C.c_int]
        L.ko_generate_frequencies.argtypes = [C.c_int, C.c_uint, C.c_double, C.c_double, dp, C.c_int]
        L.ko_transient_value_at.restype = C.c_double
        L.ko_transient_value_at.argtypes = [C.c_void_p, C.c_int, C.c_double]
        L.ko_solve_op.argtypes = [C.c_void_p, C.POINTER(_Config), dp, ip]
        L.ko_dc_num_steps.argtypes = [C.c_double, C.c_double, C.c_double]
        L.ko_solve_dc_best_effort.restype = C.c_int
        L.ko_solve_dc_best_effort.argtypes = [C.c_void_p, C.c_char_p, C.c_double, C.c_double, C.c_int, C.POINTER(C.c_double), C.c_int]
        L.ko_solve_dc.argtypes = [C.c_void_p, C.POINTER(_Config), C.c_char_p, C.c_double, C.c_double, C.c_double,
                                  dp, C.c_int, ip, C.c_int, C.c_int]
        L.ko_solve_ac.argtypes = [C.c_void_p, C.POINTER(_Config), dp, C.c_int, dp, ip, C.c_int]
        L.ko_tran_num_steps.argtypes = [C.c_double, C.c_double]
        L.ko_solve_tran.argtypes = [C.c_void_p, C.POINTER(_Config), C.c_double, C.c_double, dp, dp, C.c_int, ip]
        L.ko_solve_tran_ex.argtypes = [C.c_void_p, C.POINTER(_Config), C.c_double, C.c_double, C.c_int, dp, dp, C.c_int, ip]
        L.ko_default_config.argtypes = [C.POINTER(_Config)]
        L.ko_version.restype = C.c_char_p
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def set_device_models(models: int):
    """EXTENSION (no upstream oracle): 1 = NMOS elements stamp the level-1 companion model of nmosfet.rs:51-102; 0 (default) = the
    reference's behaviour (an NMOS stamps nothing, stampable.rs:842-844).  Process-wide switch of the restatement."""
    lib().ko_set_device_models(int(models))


class SolverConfig:
    """crates/krets-solver/src/config.rs:4-37"""

    def __init__(self, relative_tolerance=0.001, current_absolute_tolerance=1e-12,
                 voltage_absolute_tolerance=1e-6, maximum_iterations=300):
        self.c = _Config(relative_tolerance, current_absolute_tolerance, voltage_absolute_tolerance, maximum_iterations)


def generate_frequencies(kind: str, points: int, fstart: float, fstop: float) -> np.ndarray:
    """crates/krets-parser/src/analyses.rs:80-158; kind in {'Linear','Decade','Octave'}"""
    k = {"Linear": 0, "Decade": 1, "Octave": 2}[kind]
    cap = 64
    if kind == "Linear":
        cap = points + 8
    elif fstart > 0 and fstop >= fstart:
        import math
        cap = int(round((math.log10(fstop / fstart) if k == 1 else math.log2(fstop / fstart)) * points)) + 16
    out = np.empty(max(cap, 16), dtype=np.float64)
    n = lib().ko_generate_frequencies(k, points, fstart, fstop, _dp(out), out.size)
    if n < 0:
        raise OracleError(8, "frequency buffer")
    return out[:n].copy()


class Circuit:
    """crates/krets-parser/src/circuit.rs:7-20 as produced by parser.rs:27-154"""

    def __init__(self, text: str):
        err = C.c_int(0)
        msg = C.create_string_buffer(512)
        self._h = lib().ko_parse(text.encode(), C.byref(err), msg, 512)
        if not self._h:
            raise OracleError(err.value, msg.value.decode(errors="replace"))
        self.n = lib().ko_num_unknowns(self._h)
        self.num_elements = lib().ko_num_elements(self._h)
        self.names = [lib().ko_unknown_name(self._h, i).decode() for i in range(self.n)]
        self.index_map = {nm: i for i, nm in enumerate(self.names)}

    @classmethod
    def from_file(cls, path):
        with open(path, "r") as f:
            return cls(f.read())

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and _lib is not None:
            _lib.ko_free(h)
            self._h = None

    # --- element table -------------------------------------------------------------------------
    def element_index(self, ident: str) -> int:
        return lib().ko_element_index(self._h, ident.encode())

    def element_ident(self, e: int) -> str:
        return lib().ko_element_ident(self._h, e).decode()

    def element_kind(self, e: int) -> int:
        return lib().ko_element_kind(self._h, e)

    def element_value(self, e: int) -> float:
        return lib().ko_element_value(self._h, e)

    def set_element_value(self, e: int, v: float):
        lib().ko_set_element_value(self._h, e, float(v))

    def element_indices(self, e: int):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        lib().ko_element_indices(self._h, e, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    @property
    def nonlinear(self) -> bool:
        return bool(lib().ko_is_nonlinear(self._h))

    # --- stamped matrix (pattern checks) ---------------------------------------------------------
    def stamp_matrix(self, mode: str, w: float = 0.0, guess=None):
        """returns (colptr, rowidx, vals) of the summed CSC matrix; vals complex for mode 'ac'"""
        m = {"dc": 0, "ac": 1, "tran": 2}[mode]
        cap = 8 * self.num_elements + 16
        colptr = np.zeros(self.n + 1, dtype=np.int32)
        rowidx = np.zeros(cap, dtype=np.int32)
        vals = np.zeros(cap * (2 if m == 1 else 1), dtype=np.float64)
        g = None if guess is None else np.ascontiguousarray(guess, dtype=np.float64)
        nnz = lib().ko_stamp_matrix(self._h, m, float(w), _dp(g) if g is not None else None, _ip(colptr), _ip(rowidx),
                                    _dp(vals), cap)
        if nnz < 0:
            raise OracleError(7 if nnz <= -1000000 else 8, "stamp_matrix")
        if m == 1:
            v = vals[: 2 * nnz].view(np.complex128).copy()
        else:
            v = vals[:nnz].copy()
        return colptr, rowidx[:nnz].copy(), v

    # --- analyses ---------------------------------------------------------------------------------
    def solve_op(self, cfg: SolverConfig | None = None, return_iterations=False):
        cfg = cfg or SolverConfig()
        x = np.zeros(max(self.n, 1), dtype=np.float64)
        it = C.c_int(0)
        rc = lib().ko_solve_op(self._h, C.byref(cfg.c), _dp(x), C.byref(it))
        if rc:
            raise OracleError(rc, "op")
        x = x[: self.n]
        return (x, it.value) if return_iterations else x

    def solve_dc(self, element: str, start: float, stop: float, step: float, cfg=None, warm_start=True, nthreads=1):
        cfg = cfg or SolverConfig()
        num = lib().ko_dc_num_steps(start, stop, step)
        out = np.zeros((num, max(self.n, 1)), dtype=np.float64)
        npts = C.c_int(0)
        rc = lib().ko_solve_dc(self._h, C.byref(cfg.c), element.encode(), start, stop, step, _dp(out), num, C.byref(npts),
                               1 if warm_start else 0, nthreads)
        if rc:
            raise OracleError(rc, "dc")
        return out[:, : self.n]

    def solve_dc_best_effort(self, element: str, start: float, stop: float, step: float, nthreads=1):
        """SURVEY 8(d) cpu_best_effort: one-time analysis, per point re-stamp + numeric LU + solve (linear circuits)."""
        num = lib().ko_dc_num_steps(start, stop, step)
        out = np.zeros((num, max(self.n, 1)), dtype=np.float64)
        rc = lib().ko_solve_dc_best_effort(self._h, element.encode(), start, step, num, _dp(out), nthreads)
        if rc:
            raise OracleError(rc, "dc best effort")
        return out[:, : self.n]

    def solve_ac(self, freqs, cfg=None, nthreads=1):
        cfg = cfg or SolverConfig()
        f = np.ascontiguousarray(freqs, dtype=np.float64)
        out = np.zeros((max(f.size, 1), max(self.n, 1)), dtype=np.complex128)
        nout = C.c_int(0)
        rc = lib().ko_solve_ac(self._h, C.byref(cfg.c), _dp(f), f.size, out.ctypes.data_as(C.POINTER(C.c_double)),
                               C.byref(nout), nthreads)
        if rc:
            raise OracleError(rc, "ac")
        return f[f > 0.0].copy(), out[: nout.value, : self.n]

    def solve_tran(self, h: float, stop: float, cfg=None, integrator: int = 0):
        """integrator 0: reference backward Euler (transient.rs); 1: trapezoidal, 2: BE without source lag (non-reference)"""
        cfg = cfg or SolverConfig()
        ns = lib().ko_tran_num_steps(h, stop)
        out = np.zeros((ns + 1, max(self.n, 1)), dtype=np.float64)
        t = np.zeros(ns + 1, dtype=np.float64)
        nrows = C.c_int(0)
        rc = lib().ko_solve_tran_ex(self._h, C.byref(cfg.c), h, stop, integrator, _dp(out), _dp(t), ns + 1, C.byref(nrows))
        if rc:
            raise OracleError(rc, "tran")
        return t, out[:, : self.n]

    def transient_value_at(self, e: int, time: float) -> float:
        return lib().ko_transient_value_at(self._h, e, time)
