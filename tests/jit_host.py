"""Host build of a circuit-specialised lane kernel — TEST INFRASTRUCTURE ONLY.

`krb_plan_jit_source(plan, host_mode=1)` returns the CUDA text NVRTC would compile on the GPU box (krets_b200/csrc/vm_jit.cpp)
prefixed for a host build; g++ compiles it behind tests/jit_host_shim.h and this module runs it lane by lane on numpy buffers
laid out like the device's (lane-interleaved planes).  Same driver interface as vm_emulator.Emulator, so the CPU test tier can hold
the generated statements to the oracle without a GPU.  The product never imports this.
"""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

from krets_b200 import _lib
from vm_emulator import _plan_tables

HERE = os.path.dirname(os.path.abspath(__file__))


class VmParams(C.Structure):   # krb::VmParams (vm_kernel.cuh) == the struct in vm_jit_prelude.inc
    _fields_ = [("sec0", C.c_void_p), ("sec1", C.c_void_p), ("sec2", C.c_void_p), ("B", C.c_void_p), ("Bstep", C.c_void_p),
                ("n_wave", C.c_int), ("IN", C.c_void_p), ("G", C.c_void_p), ("X", C.c_void_p), ("Xprev0", C.c_void_p),
                ("n", C.c_int), ("row_width", C.c_int), ("L", C.c_longlong), ("row_begin", C.c_int), ("row_end", C.c_int),
                ("nonlinear", C.c_int), ("max_iter", C.c_int), ("guess_empty_at_iter0", C.c_int), ("pseudo_time", C.c_int),
                ("h", C.c_double), ("reltol", C.c_double), ("iabstol", C.c_double), ("vabstol", C.c_double),
                ("status", C.c_void_p), ("iters", C.c_void_p)]


def jit_source(plan, host_mode=False):
    return plan.jit_source(host_mode)


class JitHost:
    def __init__(self, plan, lanes, *, rows=1, h=0.0, cfg=None):
        src = jit_source(plan, host_mode=True)
        assert src is not None, "plan is not eligible for a specialised lane kernel"
        self._dir = tempfile.TemporaryDirectory(prefix="krb_jit_")
        cpp, so = os.path.join(self._dir.name, "k.cpp"), os.path.join(self._dir.name, "k.so")
        with open(cpp, "w") as f:
            f.write(src)
        # -ffp-contract=off: like the device build, nothing may be fused across lane-program operations
        subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-I", HERE, cpp, "-o", so], check=True)
        self._so = C.CDLL(so)
        self._fn = self._so.krb_jit_vm
        self._fn.restype = None
        self._fn.argtypes = [VmParams, C.c_longlong]
        self.plan, self.L, self.rows, self.h = plan, lanes, rows, h
        self.n = plan.n
        self.width = plan.n + max(int(_lib.lib().krb_plan_state_columns(plan._h)), 0)
        self.cplx = plan.mode == 1
        self.dtype = np.complex128 if self.cplx else np.float64
        self.pool, self.bstep = _plan_tables(plan, rows, h)
        self.n_wave = int(_lib.lib().krb_plan_wave_count(plan._h))
        self.nonlinear = bool(plan.info.nonlinear)
        self.cfg = cfg or dict(reltol=1e-3, iabstol=1e-12, vabstol=1e-6, max_iter=300)
        self.split = bool(_lib.lib().krb_plan_has_step_sections(plan._h))
        self.G = np.zeros((max(int(plan.info.spill_planes), 1), lanes), dtype=self.dtype)
        self.status = np.zeros(lanes, dtype=np.int32)
        self.iters = np.zeros(lanes, dtype=np.int32)

    def run(self, IN, X=None, row_begin=0, row_end=None, guess_empty_at_iter0=True, pseudo_time=False, Xprev0=None):
        row_end = self.rows if row_end is None else row_end
        if X is None:
            X = np.zeros((self.rows, self.width, self.L), dtype=self.dtype)
        assert X.flags.c_contiguous and X.dtype == self.dtype and X.shape == (self.rows, self.width, self.L)
        IN = np.ascontiguousarray(IN, dtype=np.float64).reshape(-1, self.L) if IN.size else np.zeros((1, self.L))
        bstep = np.ascontiguousarray(self.bstep[:, :, : max(self.n_wave, 1)])
        pool = np.ascontiguousarray(self.pool)
        xp0 = None if Xprev0 is None else np.ascontiguousarray(np.asarray(Xprev0).real, dtype=np.float64)
        nl = self.nonlinear and not self.cplx
        p = VmParams()
        one = 1   # non-null marker: the kernel only tests the section pointers
        p.sec0, p.sec1, p.sec2 = (None, one, one) if self.split else (one, None, None)
        p.B, p.Bstep, p.n_wave = pool.ctypes.data, bstep.ctypes.data, self.n_wave
        p.IN, p.G, p.X = IN.ctypes.data, self.G.ctypes.data, X.ctypes.data
        p.Xprev0 = None if xp0 is None else xp0.ctypes.data
        p.n, p.row_width, p.L = self.n, self.width, self.L
        p.row_begin, p.row_end = row_begin, row_end
        p.nonlinear, p.max_iter = (1 if nl else 0), (self.cfg["max_iter"] if nl else 1)
        p.guess_empty_at_iter0, p.pseudo_time = int(guess_empty_at_iter0), int(pseudo_time)
        p.h, p.reltol, p.iabstol, p.vabstol = self.h, self.cfg["reltol"], self.cfg["iabstol"], self.cfg["vabstol"]
        p.status, p.iters = self.status.ctypes.data, self.iters.ctypes.data
        for lane in range(self.L):
            self._fn(p, lane)
        return X
