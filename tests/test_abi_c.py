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
