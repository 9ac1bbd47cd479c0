"""Shared helpers for the parity tests."""
import numpy as np


def rel_err(a, b):
    """max over unknowns of |a-b| / max(|b|, 1e-6*||b||_inf): relative error per unknown, except that unknowns more than six
    orders below the solution's scale (mathematically zero node voltages / branch currents that come out as rounding
    noise in either solver) are measured against that floor, i.e. with absolute tolerance rtol*1e-6*||b||_inf."""
    a = np.asarray(a)
    b = np.asarray(b)
    top = float(np.max(np.abs(b))) if b.size else 0.0
    scale = np.maximum(np.abs(b), 1e-6 * top if top > 0 else 1e-300)
    return float(np.max(np.abs(a - b) / scale)) if b.size else 0.0


def assert_close(a, b, rtol=1e-9, what=""):
    """north-star tolerance: max relative error 1e-9 in f64 (complex compared through |a-b|/|b|, which bounds both
    the magnitude and the phase error)."""
    e = rel_err(a, b)
    assert e <= rtol, f"{what}: max relative error {e:.3e} > {rtol:.1e}"
