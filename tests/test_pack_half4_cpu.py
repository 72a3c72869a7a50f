"""pack_half4 (apply_fast.cu): reference floatToHalf == hardware RN conversion of `bits | 1` on
[0, 10000/203].  tools/check_pack_half4.c proves it by enumerating all 1.1e9 binary32 values of the
interval (about a second with OpenMP), so the CPU suite simply runs the whole proof."""
import os
import shutil
import subprocess

import pytest

import uhdr_testlib as T


def test_pack_half4_equivalence(tmp_path):
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    if "f16c" not in open("/proc/cpuinfo").read():
        pytest.skip("host CPU without F16C")
    exe = str(tmp_path / "check_pack_half4")
    r = subprocess.run(["gcc", "-O2", "-mf16c", "-fopenmp", os.path.join(T.ROOT, "tools", "check_pack_half4.c"), "-o", exe],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    stride = "1"  # the whole sweep takes about a second
    r = subprocess.run([exe, stride], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and ": 0 mismatches" in r.stdout, r.stdout
