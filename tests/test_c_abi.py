"""A plain-C client of include/krets_b200.h (tests/c_abi/abi_check.c) and the struct-layout check of the Rust shim.

The Rust crate under integration/krets-solver-b200 cannot be compiled in this image (no cargo/rustc): its `#[repr(C)]` structs
in src/ffi.rs are parsed here, laid out by the C rules (natural alignment) and compared field by field with offsetof/sizeof
printed by the C program built against the real header."""
import json
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c_abi", "abi_check.c")
RUST_TYPES = {"i32": (4, 4), "u32": (4, 4), "i64": (8, 8), "u64": (8, 8), "f64": (8, 8), "u8": (1, 1)}


@pytest.fixture(scope="module")
def abi_check(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("c_abi") / "abi_check")
    libdir = os.path.join(ROOT, "krets_b200")
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-O1", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe,
                           "-L", libdir, "-lkrets_b200", f"-Wl,-rpath,{libdir}", "-lm"])
    return exe


def rust_layout(struct_body):
    """C layout of a #[repr(C)] struct from its `pub name: type,` lines"""
    off, align_max, fields = 0, 1, []
    for name, ty in re.findall(r"pub\s+(\w+)\s*:\s*([^,\n]+),", struct_body):
        ty = ty.strip()
        m = re.fullmatch(r"\[(\w+);\s*(\d+)\]", ty)
        if m:
            size, align = RUST_TYPES[m.group(1)][0] * int(m.group(2)), RUST_TYPES[m.group(1)][1]
        elif ty.startswith("*"):
            size, align = 8, 8
        else:
            size, align = RUST_TYPES[ty]
        off = (off + align - 1) // align * align
        fields.append([name, off])
        off += size
        align_max = max(align_max, align)
    return {"size": (off + align_max - 1) // align_max * align_max, "fields": fields}


def test_rust_shim_struct_layouts_match_the_header(abi_check):
    c_layout = json.loads(subprocess.check_output([abi_check, "--layout"]))
    with open(os.path.join(ROOT, "integration", "krets-solver-b200", "src", "ffi.rs")) as f:
        ffi = f.read()
    for name, want in c_layout.items():
        m = re.search(r"#\[repr\(C\)\]\s*(?:#\[[^\]]*\]\s*)*pub struct " + name + r"\s*\{(.*?)\n\}", ffi, re.S)
        assert m, f"ffi.rs has no #[repr(C)] struct {name}"
        got = rust_layout(m.group(1))
        rename = {"is_": "is"}
        got["fields"] = [[rename.get(n, n), o] for n, o in got["fields"]]
        assert got == want, f"{name}: ffi.rs layout {got} != header layout {want}"


def test_rust_shim_declares_every_entry_point_it_calls():
    """every krb_* function ffi.rs binds exists in the header with the same number of parameters"""
    with open(os.path.join(ROOT, "integration", "krets-solver-b200", "src", "ffi.rs")) as f:
        ffi = f.read()
    with open(os.path.join(ROOT, "include", "krets_b200.h")) as f:
        hdr = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    for name, params in re.findall(r"pub fn (krb_\w+)\s*\((.*?)\)", ffi, re.S):
        m = re.search(r"\b" + name + r"\s*\(([^;]*?)\)\s*;", hdr, re.S)
        assert m, f"{name} is bound in ffi.rs but not declared in krets_b200.h"
        n_rust = len([p for p in params.split(",") if p.strip()])
        c_params = m.group(1).strip()
        n_c = 0 if c_params in ("", "void") else len(c_params.split(","))
        assert n_rust == n_c, f"{name}: {n_rust} parameters in ffi.rs, {n_c} in the header"


def test_c_client_host_side(abi_check):
    """analysis and introspection from plain C; without a GPU the numeric call must fail with KRB_ERR_NO_DEVICE"""
    out = subprocess.check_output([abi_check, "--solve"], text=True)
    assert "abi_check:" in out


@pytest.mark.gpu
def test_c_client_solves_on_the_gpu(abi_check):
    out = subprocess.check_output([abi_check, "--solve"], text=True)
    assert "solved 3 sweep points on the GPU" in out


# ---- sequential DC sweep (dc.rs:58,112 warm start) from plain C ---------------------------------------------------------------------
@pytest.fixture(scope="module")
def seq_sweep(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("c_abi_seq") / "seq_sweep")
    libdir = os.path.join(ROOT, "krets_b200")
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-O1", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c_abi", "seq_sweep.c"), "-o", exe, "-L", libdir, "-lkrets_b200", f"-Wl,-rpath,{libdir}", "-lm"])
    return exe


def _netlist_file(golden, key, tmp_path):
    path = str(tmp_path / (key + ".cir"))
    with open(path, "w") as f:
        f.write(golden["netlists"][key])
    return path


def test_c_client_sequential_sweep_host_side(seq_sweep, golden, tmp_path):
    import krets_b200 as kb
    if kb.device_count() > 0:
        pytest.skip("a CUDA device is present")
    out = json.loads(subprocess.check_output([seq_sweep, _netlist_file(golden, "diode_iv_curve", tmp_path), "V1", "0", "2", "0.125"]))
    assert out == {"no_device": True}


def check_sequential_sweep_output(out, oracle, golden):
    """shared with tools/seq_sweep_check.py: the C client's JSON for circuits/diode_iv_curve against the oracle"""
    import numpy as np
    from helpers import assert_close
    oc = oracle.Circuit(golden["netlists"]["diode_iv_curve"])
    warm = oc.solve_dc("V1", 0.0, 2.0, 0.125, warm_start=True)
    cold = oc.solve_dc("V1", 0.0, 2.0, 0.125, warm_start=False)
    assert out["rows"] == 17 and out["n"] == warm.shape[1]
    assert_close(np.array(out["reference"]), warm, rtol=1e-9, what="Solver DC sweep (sequential) vs the reference's warm start")
    assert_close(np.array(out["parallel"]), cold, rtol=1e-8, what="parallel cold-start extension vs the oracle run the same way")
    if "lanes" in out:
        lanes = np.array(out["lanes"]).reshape(-1, 17, out["n"])
        e = oc.element_index("R1")
        for l, r in enumerate(out["_values"]):
            oc.set_element_value(e, r)
            assert_close(lanes[l], oc.solve_dc("V1", 0.0, 2.0, 0.125, warm_start=True), rtol=1e-9, what=f"Monte-Carlo lane {l} (R1 = {r})")
        assert not any(out["status"]) and all(i >= 2 for i in out["iterations"])


@pytest.mark.gpu
def test_c_client_sequential_sweep_on_the_gpu(seq_sweep, oracle, golden, tmp_path):
    values = [1000.0, 900.0, 1500.0]
    out = json.loads(subprocess.check_output([seq_sweep, _netlist_file(golden, "diode_iv_curve", tmp_path), "V1", "0", "2", "0.125", "R1"] +
                                             [repr(v) for v in values]))
    out["_values"] = values
    check_sequential_sweep_output(out, oracle, golden)
