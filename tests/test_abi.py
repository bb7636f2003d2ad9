"""The drop-in boundary: libfrkb200.so loads and exports exactly the symbols include/frkb200.h declares; the ctypes
binding lists every one of them; without a GPU the library refuses to create a context (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "frkb200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(frk_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from frk_b200._lib import LIB_PATH
    out = subprocess.run(["nm", "-D", "--defined-only", LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\sT\s+(frk_[a-z0-9_]+)", out))
    decl = _declared()
    assert len(decl) >= 50
    missing = [s for s in decl if s not in exported]
    assert not missing, f"declared in include/frkb200.h but not exported: {missing}"
    extra = sorted(exported - set(decl))
    assert not extra, f"exported but undeclared: {extra}"


def test_ctypes_binding_covers_the_header():
    from frk_b200._lib import SIGNATURES, lib
    assert sorted(SIGNATURES) == _declared()
    assert lib.frk_abi_version() == 1


def test_header_is_plain_c():
    """the boundary is a C ABI: the header must compile as C (no C++/torch types in the signatures)."""
    r = subprocess.run(["gcc", "-std=c99", "-fsyntax-only", "-x", "c", HEADER], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_every_entry_point_cites_the_reference():
    src = open(HEADER).read()
    assert len(re.findall(r"R/[A-Za-z]+\.R:\d+", src)) >= 40


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from frk_b200._lib import FRKError, call, last_error, vp
    ctx = vp()
    with pytest.raises(FRKError, match="no CUDA device|no CPU fallback"):
        call("frk_ctx_create", 0, vp(None), C.byref(ctx))
    import frk_b200
    with pytest.raises(FRKError, match="no CPU fallback"):
        frk_b200.get_device()


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under frk_b200/ may import or execute it."""
    for dp, _, fs in os.walk(os.path.join(ROOT, "frk_b200")):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.replace("frk_oracle", "oracle") or "import" not in "".join(
                    ln for ln in txt.splitlines() if "oracle" in ln), f"{f} references the oracle"


def test_r_call_shim_type_checks_against_the_header():
    """r/src/frkb200_R.c (the .Call shim of INTEGRATION.md) cannot be built here (no R), but it must at least type-check
    against include/frkb200.h: compile it with -fsyntax-only against declaration-only stand-ins for R's headers."""
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "tests", "r_stub"),
                        "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "r", "src", "frkb200_R.c")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_overlay_calls_match_the_registered_routines():
    """r/R/overlay.R cannot run here (no R): check statically that every `.Call("<routine>", ...)` it makes is registered in the shim's
    R_CallMethodDef table with the same number of arguments, that every registered routine has a definition with that many SEXP
    parameters, and that the overlay replaces the predict path as a whole (no stop() on the default predict())."""
    import re
    shim = open(os.path.join(ROOT, "r", "src", "frkb200_R.c")).read()
    over = open(os.path.join(ROOT, "r", "R", "overlay.R")).read()
    reg = {m.group(1): int(m.group(2)) for m in re.finditer(r'\{"(frkb200_\w+)",\s*\(DL_FUNC\)&\w+,\s*(\d+)\}', shim)}
    assert len(reg) >= 15
    for name, nargs in reg.items():
        m = re.search(r"SEXP\s+" + name + r"\s*\(([^)]*)\)", shim)
        assert m, f"{name} registered but not defined"
        params = [p for p in m.group(1).split(",") if p.strip() and p.strip() != "void"]
        assert len(params) == nargs, f"{name}: {len(params)} parameters, registered with {nargs}"

    def call_args(src, start):           # number of top-level arguments of the call whose '(' is at `start`
        depth, n, i, any_tok = 0, 0, start, False
        in_str = None
        while i < len(src):
            ch = src[i]
            if in_str:
                if ch == "\\":
                    i += 1
                elif ch == in_str:
                    in_str = None
            elif ch in "\"'":
                in_str = ch; any_tok = True
            elif ch == "#":                # R comment: skip to the end of the line
                while i < len(src) and src[i] != "\n":
                    i += 1
            elif ch in "([{":
                depth += 1
            elif ch in ")]}":
                depth -= 1
                if depth == 0:
                    return n + (1 if any_tok else 0)
            elif ch == "," and depth == 1:
                n += 1
            elif not ch.isspace():
                any_tok = True
            i += 1
        raise AssertionError("unbalanced call")
    calls = list(re.finditer(r'\.Call\("(frkb200_\w+)"', over))
    assert calls
    for m in calls:
        name = m.group(1)
        assert name in reg, f"overlay calls {name}, which the shim does not register"
        nargs = call_args(over, m.start() + len(".Call")) - 1          # minus the routine name
        assert nargs == reg[name], f"overlay passes {nargs} arguments to {name}, registered with {reg[name]}"
    assert 'assignInNamespace(".SRE.predict_EM"' in over and "frkb200_model_predict" in over
    body = over[over.index(".frkb200_batch_compute_var <- function"):over.index("frkb200_enable <- function")]
    assert "stop(" not in body, "the default predict() (fs_in_process = TRUE) must not hit a stop()"
