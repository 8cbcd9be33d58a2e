"""The C ABI consumed from C: tests/abi_smoke.c is compiled with gcc -std=c99 against include/slamklt.h and
linked with libslamklt.so.  Its compile-time checks pin the struct layouts the ctypes mirror and the Julia shim assume."""
import ctypes as C
import os
import subprocess

import pytest

import slamklt

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "slam.jl_b200", "csrc")


def _build(tmp_path):
    exe = str(tmp_path / "abi_smoke")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "abi_smoke.c"),
                           "-o", exe, "-L" + CSRC, "-lslamklt", "-Wl,-rpath," + CSRC, "-lm"])
    return exe


def test_header_is_c99_and_layouts_match(tmp_path):
    exe = _build(tmp_path)
    out = subprocess.run([exe, "--layout-only"], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stderr
    # the ctypes mirror agrees with the C compiler
    assert C.sizeof(slamklt.LKParams) == 40 and C.sizeof(slamklt.DetectParams) == 40
    assert C.sizeof(slamklt.CameraC) == 208 and C.sizeof(slamklt.MatchingParams) == 56 and C.sizeof(slamklt.Stats) == 40


@pytest.mark.gpu
def test_c_program_tracks_and_detects(tmp_path):
    exe = _build(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, (out.stdout, out.stderr)
    assert "abi_smoke: ok" in out.stdout


def _split_top(s):
    """split at top-level commas (parentheses and braces nest)"""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "({[":
            depth += 1
        elif ch in ")}]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def _balanced(s, i):
    """s[i] == '(' -> index just past its matching ')'"""
    depth = 0
    for j in range(i, len(s)):
        depth += s[j] == "("
        depth -= s[j] == ")"
        if depth == 0:
            return j + 1
    raise ValueError("unbalanced")


def test_julia_shim_ccalls_match_the_header():
    """The Julia shim cannot be compiled here (no Julia in the image), so its `ccall`s are checked statically against
    include/slamklt.h: every symbol exists in the header and in the library, the argument-type tuple has as many entries as the C
    prototype has parameters, as many values follow it, and pointer / integer / double positions agree."""
    import re
    hdr = open(os.path.join(ROOT, "include", "slamklt.h")).read()
    hdr = re.sub(r"/\*.*?\*/", " ", hdr, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(?:int|const char\*)\s+(slamklt_\w+)\s*\(", hdr):
        end = _balanced(hdr, m.end() - 1)
        params = _split_top(hdr[m.end():end - 1])
        protos[m.group(1)] = [] if params == ["void"] else params

    def c_kind(p):
        if "*" in p:
            return "ptr"
        t = p.split()[0] if not p.startswith("const") else p.split()[1]
        return {"double": "f64", "size_t": "size"}.get(t, "int")

    def jl_kind(t):
        if t.startswith(("Ptr{", "Ref{")) or t == "Cstring":
            return "ptr"
        return {"Cdouble": "f64", "Float64": "f64", "Csize_t": "size"}.get(t, "int")

    lib = slamklt.lib()
    n_calls = 0
    for fn in ("SlamKLT.jl", "SlamKLTOptional.jl", "dump_golden.jl"):
        src = open(os.path.join(ROOT, "julia", fn)).read()
        for m in re.finditer(r"ccall\(\s*\(\s*:(slamklt_\w+)\s*,\s*libslamklt\s*\)\s*,", src):
            name = m.group(1)
            end = _balanced(src, m.start() + len("ccall"))
            parts = _split_top(src[m.end():end - 1])          # return type, type tuple, values...
            assert name in protos, (fn, name, "not declared in slamklt.h")
            assert hasattr(lib, name), (fn, name, "not exported by libslamklt.so")
            rettype, types, values = parts[0], parts[1], parts[2:]
            assert types.startswith("(") and types.endswith(")"), (fn, name, types)
            jl_types = _split_top(types[1:-1])
            assert len(jl_types) == len(protos[name]), (fn, name, len(jl_types), len(protos[name]))
            assert len(values) == len(jl_types), (fn, name, "values", len(values), "types", len(jl_types))
            assert [jl_kind(t) for t in jl_types] == [c_kind(p) for p in protos[name]], (fn, name, jl_types, protos[name])
            assert rettype in ("Cint", "Cstring"), (fn, name, rettype)
            n_calls += 1
    assert n_calls >= 20


def test_julia_shim_struct_layouts():
    """The isbits structs the shim passes by reference have the sizes the C compiler gives the header's structs (40 / 40 / 208 / 56:
    checked against the compiled abi_smoke program in test_header_is_c99_and_layouts_match); all fields are 4- or 8-byte scalars in
    an order that needs no padding surprises: 4-byte fields come in pairs before every 8-byte field."""
    import re
    size = {"Int32": 4, "Float64": 8, "Int64": 8, "NTuple{16, Float64}": 128}
    structs = {}
    for fn in ("SlamKLT.jl", "SlamKLTOptional.jl"):
        src = open(os.path.join(ROOT, "julia", fn)).read()
        for m in re.finditer(r"^struct (SlamKlt\w+)\n(.*?)^end", src, flags=re.S | re.M):
            fields = re.findall(r"\w+::([\w{}, ]+?)(?:;|\n|$)", m.group(2))
            off = 0
            for t in fields:
                t = t.strip()
                sz = size[t] if t in size else structs[t]
                al = min(sz, 8)
                assert off % al == 0, (m.group(1), t, off)      # no implicit padding anywhere
                off += sz
            structs[m.group(1)] = off
    assert structs == {"SlamKltLKParams": C.sizeof(slamklt.LKParams), "SlamKltDetectParams": C.sizeof(slamklt.DetectParams),
                       "SlamKltCamera": C.sizeof(slamklt.CameraC), "SlamKltMatchingParams": C.sizeof(slamklt.MatchingParams)}, structs
