"""SURVEY 8(f)3 asks for an allocation-free host container layer.  tests/cpp/alloc_probe.c replaces malloc for the
whole process, drives the C API on warmed handles and counts the heap calls that come out of libuhdr_b200.so: none in
steady state -- API-4 assembly + probe (host only, runs here) and API-1 encode / decode on the device (GPU box)."""
import io
import os
import subprocess

import numpy as np
import pytest

import uhdr_testlib as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def probe(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("alloc") / "alloc_probe")
    so = T.GPU_SO
    cmd = ["gcc", "-O1", "-g", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "alloc_probe.c"), "-o", exe,
           "-L", os.path.dirname(so), "-l:" + os.path.basename(so), "-Wl,-rpath," + os.path.dirname(so), "-ldl", "-rdynamic"]
    subprocess.run(cmd, check=True, capture_output=True)
    return exe


def _run(exe, args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([exe] + args, capture_output=True, text=True, timeout=300, env=e)
    return r.returncode, r.stdout, r.stderr


def _counts(out):
    res = {}
    for line in out.splitlines():
        if "ours=" in line:
            name = line[:line.index("ours=")].strip()
            res[name] = {k: int(v) for k, v in (f.split("=") for f in line[line.index("ours="):].split())}
    return res


def test_api4_assembly_and_probe_do_not_touch_the_heap(probe, tmp_path):
    PIL = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(1)
    a = (rng.random((128, 192, 3)) * 255).astype(np.uint8)
    base, gm = str(tmp_path / "base.jpg"), str(tmp_path / "gm.jpg")
    PIL.fromarray(a).save(base, quality=90)
    PIL.fromarray(a[::2, ::2, 0]).save(gm, quality=90)
    rc, out, err = _run(probe, ["api4", base, gm])
    assert rc == 0, (out, err)
    c = _counts(out)
    assert c and all(v["ours"] == 0 for v in c.values()), out
    # self check: the same run with the warm-up iteration counted does see the library's first-use allocations
    rc, out, err = _run(probe, ["api4", base, gm], {"ALLOC_PROBE_COUNT_WARMUP": "1"})
    assert rc == 1 and sum(v["ours"] for v in _counts(out).values()) > 0, (out, err[-2000:])


@pytest.mark.gpu
@pytest.mark.parametrize("w,h", [(1280, 720), (1920, 1080)])
def test_encode_decode_steady_state_does_not_touch_the_heap(probe, w, h):
    rc, out, err = _run(probe, ["gpu", str(w), str(h)])
    assert rc == 0, (out, err[-4000:])
    c = _counts(out)
    assert len(c) == 3 and all(v["ours"] == 0 for v in c.values()), out
