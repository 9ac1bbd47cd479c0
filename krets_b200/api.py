"""Python veneer over the krets_b200 C ABI.

Two layers, both thin:
  * `Circuit`, `Solver`, `Analysis`, `AnalysisResult` mirror the krets public API (parser.rs:27,156; solver.rs:17-52;
    lib.rs:13-77) through the C++ host mirror (krh_* in include/krets_b200_host.h);
  * `Plan` exposes the array-level drop-in boundary (krb_* in include/krets_b200.h) for lane batches: numpy arrays are
    passed as host memory, torch CUDA tensors as device memory on torch's current stream.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import KrbConfig, KrbIO, KrbOverrides, KrbPlanInfo, KrbPlanOpts

MODE_DC, MODE_AC, MODE_TRAN = 0, 1, 2
MEM_HOST, MEM_DEVICE = 0, 1
LAYOUT_POINT_MAJOR, LAYOUT_LANE_INTERLEAVED = 0, 1
EXEC_FUSED, EXEC_STAGED, EXEC_LANE_VM = 0, 1, 2   # krb_plan_opts.exec (STAGED = the supernodal front executor)
LANE_ZERO_PIVOT, LANE_NOT_CONVERGED, LANE_NONFINITE, LANE_INACCURATE = 1, 2, 4, 8
KIND_R, KIND_C, KIND_L, KIND_V, KIND_I, KIND_D, KIND_NMOS = range(7)   # krb_element.kind (elements.rs:45-55 order)

_ERR_NAMES = {
    1: "InvalidArgument", 3: "ElementNotFound", 4: "InvalidElementFormat", 5: "MaximumIterationsExceeded",
    6: "DecompositionFailed", 7: "Unsupported", 20: "NoDevice", 21: "Cuda", 22: "OutOfMemory", 23: "Internal",
    30: "ParseError", 31: "EmptyNetlist", 32: "UnsupportedElement", 33: "MissingModel", 34: "Io",
}


class KretsError(RuntimeError):
    """krets_solver::Error / krets_parser::Error carried across the C ABI."""

    def __init__(self, code: int, message: str):
        self.code = code
        self.kind = _ERR_NAMES.get(code, str(code))
        super().__init__(f"{self.kind}: {message}")


def _check_krb(rc: int):
    if rc:
        raise KretsError(rc, _lib.lib().krb_last_error_string().decode(errors="replace"))


def _check_krh(rc: int):
    if rc:
        raise KretsError(rc, _lib.lib().krh_last_error_string().decode(errors="replace"))


def version() -> str:
    return _lib.lib().krb_version().decode()


def device_count() -> int:
    return int(_lib.lib().krb_device_count())


def kernel_launch_count() -> int:
    return int(_lib.lib().krb_kernel_launch_count())


def last_kernel_ms() -> float:
    return float(_lib.lib().krb_last_kernel_ms())


class SolverConfig:
    """krets_solver::config::SolverConfig (config.rs:4-37)."""

    def __init__(self, relative_tolerance=0.001, current_absolute_tolerance=1e-12, voltage_absolute_tolerance=1e-6,
                 maximum_iterations=300):
        self.c = KrbConfig(relative_tolerance, current_absolute_tolerance, voltage_absolute_tolerance, maximum_iterations)

    @staticmethod
    def default():
        return SolverConfig()


# ---- krets_parser::analyses ---------------------------------------------------------------------------------
@dataclass
class DcAnalysis:            # analyses.rs:44-54
    element: str
    start: float
    stop: float
    step_size: float
    mode: int = 0            # extension: 0 = the reference (nonlinear sweeps sequential, warm-started dc.rs:58), 1 = nonlinear points as parallel cold-start lanes


@dataclass
class AcSweep:               # analyses.rs:56-65
    variation: str           # "Linear" | "Decade" | "Octave"
    points: int

    @staticmethod
    def Linear(total_points):
        return AcSweep("Linear", total_points)

    @staticmethod
    def Decade(points_per_decade):
        return AcSweep("Decade", points_per_decade)

    @staticmethod
    def Octave(points_per_octave):
        return AcSweep("Octave", points_per_octave)


@dataclass
class AcAnalysis:            # analyses.rs:68-76
    sweep: AcSweep
    fstart: float
    fstop: float

    def generate_frequencies(self):
        return generate_frequencies(self.sweep.variation, self.sweep.points, self.fstart, self.fstop)


INTEGRATOR_BE_KRETS, INTEGRATOR_TRAP, INTEGRATOR_BE = 0, 1, 2
DC_REFERENCE, DC_PARALLEL_COLD_START = 0, 1        # DcAnalysis.mode (krets_b200_host.h KRH_DC_*)


@dataclass
class TransientAnalysis:     # analyses.rs:161-165
    time_step: float
    stop_time: float
    integrator: int = 0      # extension: 0 = the reference's backward Euler (incl. its source lag), 1 = trapezoidal, 2 = BE at t_n


class Analysis:              # analyses.rs:27-41
    def __init__(self, kind, params=None):
        self.kind, self.params = kind, params

    @staticmethod
    def Op():
        return Analysis("op")

    @staticmethod
    def Dc(p: DcAnalysis):
        return Analysis("dc", p)

    @staticmethod
    def Ac(p: AcAnalysis):
        return Analysis("ac", p)

    @staticmethod
    def Transient(p: TransientAnalysis):
        return Analysis("transient", p)


_SWEEP_KIND = {"Linear": 0, "Decade": 1, "Octave": 2}


def generate_frequencies(variation: str, points: int, fstart: float, fstop: float) -> np.ndarray:
    """AcAnalysis::generate_frequencies (analyses.rs:80-158), computed by the C++ host mirror."""
    L = _lib.lib()
    k = _SWEEP_KIND[variation]
    n = L.krh_generate_frequencies(k, points, fstart, fstop, None, 0)
    out = np.empty(max(int(n), 1), dtype=np.float64)
    n2 = L.krh_generate_frequencies(k, points, fstart, fstop, out.ctypes.data_as(C.POINTER(C.c_double)), out.size)
    assert n2 == n
    return out[:n]


# ---- krets_parser::circuit::Circuit ---------------------------------------------------------------------------
DEVICES_REFERENCE, DEVICES_LEVEL1 = 0, 1


class Circuit:
    def __init__(self, handle):
        L = _lib.lib()
        self._h = handle
        self.n = L.krh_circuit_num_unknowns(handle)
        self.num_elements = L.krh_circuit_num_elements(handle)
        self.names = [L.krh_circuit_unknown_name(handle, i).decode() for i in range(self.n)]
        self.index_map = {nm: i for i, nm in enumerate(self.names)}
        self.identifiers = [L.krh_circuit_element_identifier(handle, e).decode() for e in range(self.num_elements)]

    def __del__(self):
        if getattr(self, "_h", None) and _lib._lib is not None:
            _lib._lib.krh_circuit_free(self._h)
            self._h = None

    def element_index(self, identifier: str) -> int:
        try:
            return self.identifiers.index(identifier)
        except ValueError:
            return -1

    def element(self, e: int):
        return _lib.lib().krh_circuit_elements(self._h)[e]

    def element_value(self, e: int) -> float:
        return float(self.element(e).value)

    def set_device_models(self, models: int) -> int:
        """EXTENSION (krets_b200_host.h): DEVICES_LEVEL1 makes every NMOS stamp the level-1 companion model (nmosfet.rs:51-102);
        DEVICES_REFERENCE (default) reproduces the reference, whose NMOS stamps nothing.  Call before building Solver / Plan objects."""
        k = int(_lib.lib().krh_circuit_set_device_models(self._h, models))
        if k < 0:
            raise ValueError("set_device_models: unknown model set")
        return k


def parse_circuit_description(text: str) -> Circuit:
    h = C.c_void_p()
    _check_krh(_lib.lib().krh_parse_circuit_description(text.encode(), C.byref(h)))
    return Circuit(h)


def parse_circuit_description_file(path) -> Circuit:
    h = C.c_void_p()
    _check_krh(_lib.lib().krh_parse_circuit_description_file(str(path).encode(), C.byref(h)))
    return Circuit(h)


# ---- krets_solver::AnalysisResult ------------------------------------------------------------------------------
class AnalysisResult:
    """lib.rs:13-77.  `.array` is [rows][n] (complex for AC) in index-map order, `.axis` the step/frequency/time column."""

    def __init__(self, kind: str, names, array: np.ndarray, axis):
        self.kind, self.names, self.array, self.axis = kind, names, array, axis

    def _rows(self, axis_key, cast=float):
        out = []
        for r in range(self.array.shape[0]):
            d = {nm: (complex(v) if self.kind == "ac" else float(v)) for nm, v in zip(self.names, self.array[r])}
            if axis_key:
                d[axis_key] = complex(self.axis[r], 0.0) if self.kind == "ac" else float(self.axis[r])
            out.append(d)
        return out

    def into_op(self):
        if self.kind != "op":
            raise RuntimeError("Called `into_op()` on a non-Op result")
        return self._rows(None)[0]

    def into_dc(self):
        if self.kind != "dc":
            raise RuntimeError("Called `into_dc()` on a non-Dc result")
        return self._rows("step")

    def into_ac(self):
        if self.kind != "ac":
            raise RuntimeError("Called `into_ac()` on a non-Ac result")
        return self._rows("frequency")

    def into_transient(self):
        if self.kind != "transient":
            raise RuntimeError("Called `into_transient()` on a non-Transient result")
        return self._rows("time")


class Solver:
    """krets_solver::solver::Solver (solver.rs:17-52) over the CUDA backend."""

    def __init__(self, circuit: Circuit, config: SolverConfig | None = None):
        self.circuit = circuit
        self.config = config or SolverConfig()
        h = C.c_void_p()
        _check_krh(_lib.lib().krh_solver_new(circuit._h, C.byref(self.config.c), C.byref(h)))
        self._h = h

    def __del__(self):
        if getattr(self, "_h", None) and _lib._lib is not None:
            _lib._lib.krh_solver_free(self._h)
            self._h = None

    def solve(self, analysis: Analysis) -> AnalysisResult:
        L = _lib.lib()
        r = C.c_void_p()
        p = analysis.params
        if analysis.kind == "op":
            _check_krh(L.krh_solver_solve_op(self._h, C.byref(r)))
        elif analysis.kind == "dc":
            _check_krh(L.krh_solver_solve_dc_with(self._h, p.element.encode(), p.start, p.stop, p.step_size, p.mode, C.byref(r)))
        elif analysis.kind == "ac":
            _check_krh(L.krh_solver_solve_ac(self._h, _SWEEP_KIND[p.sweep.variation], p.sweep.points, p.fstart, p.fstop, C.byref(r)))
        elif analysis.kind == "transient":
            _check_krh(L.krh_solver_solve_transient_with(self._h, p.time_step, p.stop_time, p.integrator, C.byref(r)))
        else:
            raise ValueError(analysis.kind)
        try:
            rows, n = int(L.krh_result_rows(r)), int(L.krh_result_num_unknowns(r))
            width = n * (2 if analysis.kind == "ac" else 1)
            data = np.ctypeslib.as_array(L.krh_result_data(r), shape=(rows * max(width, 1),)).copy() if rows * width else np.zeros(0)
            arr = data.view(np.complex128).reshape(rows, n) if analysis.kind == "ac" else data.reshape(rows, n)
            ax_ptr = L.krh_result_axis(r)
            axis = np.ctypeslib.as_array(ax_ptr, shape=(rows,)).copy() if (ax_ptr and rows) else None
        finally:
            L.krh_result_free(r)
        return AnalysisResult(analysis.kind, self.circuit.names, arr, axis)


# ---- array-level boundary (krb_*) -------------------------------------------------------------------------------
def _is_device(x) -> bool:
    return hasattr(x, "data_ptr") and hasattr(x, "is_cuda") and bool(x.is_cuda)


def _ptr(x):
    if x is None:
        return None
    if _is_device(x):
        return C.c_void_p(x.data_ptr())
    return C.c_void_p(x.ctypes.data)


def _f64_in(x, shape, name):
    """A lane-sized INPUT of the C ABI as packed float64: host arrays are converted (np.ascontiguousarray), device tensors must
    already be contiguous float64 (no silent device copies); the shape is checked either way."""
    if x is None:
        return None
    if _is_device(x):
        import torch
        if x.dtype != torch.float64 or not x.is_contiguous():
            raise ValueError(f"{name}: CUDA tensors must be contiguous float64 (got {x.dtype}, contiguous={x.is_contiguous()})")
    else:
        x = np.ascontiguousarray(x, dtype=np.float64)
    if shape is not None and tuple(x.shape) != tuple(shape):
        raise ValueError(f"{name}: expected shape {tuple(shape)}, got {tuple(x.shape)}")
    return x


def _check_out(x, shape, complex_, name="out"):
    """A caller-provided OUTPUT buffer is written in place: it must be contiguous, of the right dtype and shape."""
    if _is_device(x):
        import torch
        ok = x.dtype == (torch.complex128 if complex_ else torch.float64) and x.is_contiguous()
    else:
        ok = isinstance(x, np.ndarray) and x.dtype == (np.complex128 if complex_ else np.float64) and x.flags["C_CONTIGUOUS"]
    if not ok or tuple(x.shape) != tuple(shape):
        raise ValueError(f"{name}: expected a contiguous {'complex128' if complex_ else 'float64'} array of shape {tuple(shape)}")
    return x


class Plan:
    """One-time symbolic analysis of a circuit for one stamp mode (krb_analyze) + lane-batch solves (krb_solve_*)."""

    def __init__(self, circuit: Circuit, mode: int, *, sweep_element: int = -1, varied=(), time_step: float = 0.0,
                 nominal_frequency: float = 0.0, scratch_slots: int = 0, threads_per_block: int = 0, force_lane_vm: bool = False,
                 integrator: int = 0, executor: int | None = None, refine_steps: int = 0):
        L = _lib.lib()
        self.circuit, self.mode = circuit, mode
        self.varied = list(varied)
        kc = C.c_void_p()
        _check_krb(L.krb_circuit_create(L.krh_circuit_elements(circuit._h), circuit.num_elements, circuit.n,
                                        L.krh_circuit_is_current(circuit._h), C.byref(kc)))
        self._kc = kc
        o = KrbPlanOpts()
        L.krb_default_plan_opts(C.byref(o))
        o.sweep_element = sweep_element
        o.time_step = time_step
        o.nominal_frequency = nominal_frequency
        o.scratch_slots = scratch_slots
        o.threads_per_block = threads_per_block
        o.exec = executor if executor is not None else (EXEC_LANE_VM if force_lane_vm else EXEC_FUSED)
        o.integrator = integrator
        o.refine_steps = refine_steps     # front executor: k > 0 = k refinement steps + backward-error check (krets_b200.h)
        self._varied_arr = (C.c_int32 * max(len(self.varied), 1))(*self.varied)
        o.n_varied = len(self.varied)
        o.varied_elements = C.cast(self._varied_arr, C.POINTER(C.c_int32))
        h = C.c_void_p()
        try:
            _check_krb(L.krb_analyze(kc, mode, C.byref(o), C.byref(h)))
        except Exception:
            L.krb_circuit_destroy(kc)
            self._kc = None
            raise
        self._h = h
        info = KrbPlanInfo()
        _check_krb(L.krb_plan_get_info(h, C.byref(info)))
        self.info = info
        self.n = info.n

    def __del__(self):
        if _lib._lib is None:
            return
        if getattr(self, "_h", None):
            _lib._lib.krb_plan_destroy(self._h)
            self._h = None
        if getattr(self, "_kc", None):
            _lib._lib.krb_circuit_destroy(self._kc)
            self._kc = None

    # -- introspection --
    def front_tables(self) -> dict:
        """Tables of the front executor (debug introspection; tests run them through a numpy multifrontal emulator)."""
        L = _lib.lib()
        names = ["piv0", "idx_ptr", "idx", "parent", "child_ptr", "child", "rel", "level_ptr", "level_fronts", "asm_ptr", "asm_ent", "asm_pos"]
        out = {}
        for which, name in enumerate(names):
            size = L.krb_plan_front_array(self._h, which, None)
            if size < 0:
                raise KretsError(1, "plan does not use the front executor")
            a = np.zeros(max(size, 1), dtype=np.int32)
            L.krb_plan_front_array(self._h, which, a.ctypes.data_as(C.POINTER(C.c_int32)))
            out[name] = a[:size]
        info = (C.c_int64 * 8)()
        _check_krb(L.krb_plan_front_info(self._h, info))
        for k, name in enumerate(["n_fronts", "n_levels", "max_m", "max_p", "lu_size", "cb_size", "cbv_size", "flops"]):
            out[name] = int(info[k])
        return out

    def pattern(self):
        colptr = np.zeros(self.n + 1, dtype=np.int32)
        rowidx = np.zeros(max(self.info.nnz, 1), dtype=np.int32)
        _check_krb(_lib.lib().krb_plan_get_pattern(self._h, colptr.ctypes.data_as(C.POINTER(C.c_int32)),
                                                   rowidx.ctypes.data_as(C.POINTER(C.c_int32))))
        return colptr, rowidx[: self.info.nnz]

    def pivot_order(self):
        r = np.zeros(max(self.n, 1), dtype=np.int32)
        c = np.zeros(max(self.n, 1), dtype=np.int32)
        _check_krb(_lib.lib().krb_plan_get_pivot_order(self._h, r.ctypes.data_as(C.POINTER(C.c_int32)),
                                                       c.ctypes.data_as(C.POINTER(C.c_int32))))
        return r[: self.n], c[: self.n]

    def program(self, section: int = 0) -> np.ndarray:
        n = int(_lib.lib().krb_plan_program_size(self._h, section))
        ops = np.zeros((max(n, 0), 2), dtype=np.uint32)
        if n > 0:
            _check_krb(_lib.lib().krb_plan_get_program(self._h, section, ops.ctypes.data_as(C.POINTER(C.c_uint32))))
        return ops

    def jit_source(self, host_mode: bool = False) -> str | None:
        """CUDA text of the plan's circuit-specialised lane kernel (krb_plan_jit_source), None when the plan is not eligible."""
        L = _lib.lib()
        need = int(L.krb_plan_jit_source(self._h, 1 if host_mode else 0, None, 0))
        if need <= 0:
            return None
        buf = C.create_string_buffer(need)
        L.krb_plan_jit_source(self._h, 1 if host_mode else 0, buf, need)
        return buf.value.decode()

    @property
    def specialised(self) -> bool:
        """True once the plan's lane program runs as an NVRTC-compiled kernel on this device (decided at its first solve)."""
        return int(_lib.lib().krb_plan_jit_active(self._h)) == 1

    # -- solves --
    def _io(self, arrays, layout, select=None):
        io = KrbIO()
        _lib.lib().krb_default_io(C.byref(io))
        if select is not None:        # extension: deliver only these unknowns (index-map indices), krets_b200.h krb_io.select
            self._sel = np.ascontiguousarray(select, dtype=np.int32)
            io.select = self._sel.ctypes.data_as(C.POINTER(C.c_int32))
            io.n_select = int(self._sel.size)
        dev = [_is_device(a) for a in arrays if a is not None]
        if any(dev) and not all(dev):
            raise ValueError("all lane-sized arrays of one call must be on the same side (host numpy or CUDA tensors)")
        io.memory = MEM_DEVICE if any(dev) else MEM_HOST
        io.layout = layout
        if io.memory == MEM_DEVICE:
            import torch
            io.stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        return io

    def _ov(self, values):
        if not self.varied:
            return None, None
        ov = KrbOverrides()
        ov.n_varied = len(self.varied)
        ov.values = _ptr(values)
        return ov, values

    def solve_op(self, lanes: int = 1, varied_values=None, out=None, cfg: SolverConfig | None = None, layout=LAYOUT_POINT_MAJOR,
                 return_iterations=False, select=None):
        cfg = cfg or SolverConfig()
        varied_values = _f64_in(varied_values, (len(self.varied), lanes) if self.varied else None, "varied_values")
        cols = self.n if select is None else len(select)
        out = _check_out(out, _shape(layout, lanes, 1, cols), False) if out is not None else _like_out(varied_values, _shape(layout, lanes, 1, cols), False)
        status = _like_int(out, lanes)
        iters = _like_int(out, lanes) if return_iterations else None
        ov, _keep = self._ov(varied_values)
        io = self._io([out, varied_values], layout, select)
        _check_krb(_lib.lib().krb_solve_op(self._h, C.byref(cfg.c), lanes, C.byref(ov) if ov else None, _ptr(out), _ptr(iters),
                                           _ptr(status), C.byref(io)))
        return (out, status, iters) if return_iterations else (out, status)

    def solve_dc_sweep(self, sweep_values, varied_values=None, out=None, cfg: SolverConfig | None = None, layout=LAYOUT_POINT_MAJOR, select=None):
        cfg = cfg or SolverConfig()
        sweep_values = _f64_in(sweep_values, None, "sweep_values")
        if sweep_values.ndim != 1:
            raise ValueError("sweep_values: expected a 1-D array of per-lane source values")
        lanes = int(sweep_values.shape[0])
        varied_values = _f64_in(varied_values, (len(self.varied), lanes) if self.varied else None, "varied_values")
        cols = self.n if select is None else len(select)
        out = _check_out(out, _shape(layout, lanes, 1, cols), False) if out is not None else _like_out(sweep_values, _shape(layout, lanes, 1, cols), False)
        status = _like_int(out, lanes)
        ov, _keep = self._ov(varied_values)
        io = self._io([sweep_values, out, varied_values], layout, select)
        _check_krb(_lib.lib().krb_solve_dc_sweep(self._h, C.byref(cfg.c), _ptr(sweep_values), lanes, C.byref(ov) if ov else None,
                                                 _ptr(out), _ptr(status), C.byref(io)))
        return out, status

    def solve_dc_sweep_sequential(self, start: float, step_size: float, n_points: int, lanes: int = 1, varied_values=None, out=None,
                                  cfg: SolverConfig | None = None, layout=LAYOUT_POINT_MAJOR, return_iterations=False, select=None):
        """dc.rs:47-113 as written (nonlinear circuits): the sweep points are the rows of a lane, solved in order, Newton-Raphson
        warm-started from the previous point (dc.rs:58,112); lanes are Monte-Carlo variants.  Returns x[lane][row][unknown]."""
        cfg = cfg or SolverConfig()
        varied_values = _f64_in(varied_values, (len(self.varied), lanes) if self.varied else None, "varied_values")
        cols = self.n if select is None else len(select)
        shape = (lanes, n_points, cols) if layout == LAYOUT_POINT_MAJOR else (n_points, cols, lanes)
        if out is None:
            out = _like_out(varied_values, shape, False) if varied_values is not None else np.zeros(shape, dtype=np.float64)
        else:
            _check_out(out, shape, False)
        status = _like_int(out, lanes)
        iters = _like_int(out, lanes) if return_iterations else None
        ov, _keep = self._ov(varied_values)
        io = self._io([out, varied_values], layout, select)
        _check_krb(_lib.lib().krb_solve_dc_sweep_sequential(self._h, C.byref(cfg.c), float(start), float(step_size), int(n_points), lanes,
                                                            C.byref(ov) if ov else None, _ptr(out), _ptr(iters), _ptr(status), C.byref(io)))
        return (out, status, iters) if return_iterations else (out, status)

    def solve_ac(self, frequencies, varied_values=None, op_solution=None, out=None, cfg: SolverConfig | None = None,
                 layout=LAYOUT_POINT_MAJOR, select=None):
        cfg = cfg or SolverConfig()
        frequencies = _f64_in(frequencies, None, "frequencies")
        if frequencies.ndim != 1:
            raise ValueError("frequencies: expected a 1-D array")
        lanes = int(frequencies.shape[0])
        varied_values = _f64_in(varied_values, (len(self.varied), lanes) if self.varied else None, "varied_values")
        cols = self.n if select is None else len(select)
        out = _check_out(out, _shape(layout, lanes, 1, cols), True) if out is not None else _like_out(frequencies, _shape(layout, lanes, 1, cols), True)
        status = _like_int(out, lanes)
        ov, _keep = self._ov(varied_values)
        lanes_op = 0
        if op_solution is not None:
            op_solution = _f64_in(op_solution, None, "op_solution")
            if op_solution.ndim == 1 and op_solution.shape[0] == self.n:
                lanes_op = 1
            elif op_solution.ndim == 2 and tuple(op_solution.shape) == (lanes, self.n):
                lanes_op = lanes
            else:
                raise ValueError(f"op_solution: expected shape ({self.n},) or ({lanes}, {self.n})")
        io = self._io([frequencies, out, varied_values, op_solution], layout, select)
        _check_krb(_lib.lib().krb_solve_ac(self._h, C.byref(cfg.c), _ptr(frequencies), lanes, C.byref(ov) if ov else None,
                                           _ptr(op_solution), lanes_op, _ptr(out), _ptr(status), C.byref(io)))
        return out, status

    def solve_transient(self, plan_dc: "Plan", n_steps: int, lanes: int = 1, varied_values=None, out=None,
                        cfg: SolverConfig | None = None, layout=LAYOUT_POINT_MAJOR, select=None):
        """self must be a MODE_TRAN plan; plan_dc the MODE_DC plan of the same circuit (row 0 = OP, transient.rs:21)."""
        cfg = cfg or SolverConfig()
        varied_values = _f64_in(varied_values, (len(self.varied), lanes) if self.varied else None, "varied_values")
        cols = self.n if select is None else len(select)
        shape = (lanes, n_steps + 1, cols) if layout == LAYOUT_POINT_MAJOR else (n_steps + 1, cols, lanes)
        if out is None:
            out = _like_out(varied_values, shape, False) if varied_values is not None else np.zeros(shape, dtype=np.float64)
        else:
            _check_out(out, shape, False)
        status = _like_int(out, lanes)
        ov, _keep = self._ov(varied_values)
        io = self._io([out, varied_values], layout, select)
        _check_krb(_lib.lib().krb_solve_transient(plan_dc._h, self._h, C.byref(cfg.c), n_steps, lanes, C.byref(ov) if ov else None,
                                                  _ptr(out), _ptr(status), C.byref(io)))
        return out, status


def _shape(layout, lanes, rows, n):
    """result shape of a one-row analysis: x[lane][unknown] point-major, x[unknown][lane] lane-interleaved"""
    return (lanes, n) if layout == LAYOUT_POINT_MAJOR else (n, lanes)


def _like_out(ref, shape, complex_):
    if _is_device(ref):
        import torch
        return torch.empty(shape, dtype=torch.complex128 if complex_ else torch.float64, device=ref.device)
    return np.zeros(shape, dtype=np.complex128 if complex_ else np.float64)


def _like_int(ref, n):
    if _is_device(ref):
        import torch
        return torch.zeros(n, dtype=torch.int32, device=ref.device)
    return np.zeros(n, dtype=np.int32)
