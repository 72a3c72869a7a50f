"""include/uhdr_b200_jpegr.hpp (C++ mirror of ultrahdr::JpegR over the C ABI): compiles and links
against libuhdr_b200.so everywhere; on a GPU the program runs and checks the mirror against the C API."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "jpegr_mirror_test.cpp")
LIBDIR = os.path.join(ROOT, "libultrahdr_b200")


def _build(tmp_path):
    if not os.path.exists(os.path.join(LIBDIR, "libuhdr_b200.so")):
        sys.path.insert(0, ROOT)
        import __graft_entry__ as G
        G.build()
    exe = str(tmp_path / "jpegr_mirror_test")
    cudalib = "/usr/local/cuda/lib64"
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe,
           "-L", LIBDIR, "-luhdr_b200", "-Wl,-rpath," + LIBDIR, "-Wl,-rpath," + cudalib, "-L", cudalib, "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_mirror_header_compiles_and_links(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "linked" in r.stdout, (r.stdout, r.stderr)


@pytest.mark.gpu
def test_mirror_matches_c_api(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe, "run"], capture_output=True, text=True)
    assert r.returncode == 0 and "ok" in r.stdout, (r.stdout, r.stderr)
