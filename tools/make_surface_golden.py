"""Builds tests/cpp/jpegr_surface_test.cpp against the REFERENCE's own headers and objects
(oracle/_ref/obj_turbo, made by oracle/Makefile) and stores what it prints in
tests/golden/jpegr_surface_ref.txt.  Run in the build container (needs /root/reference):
    python tools/make_surface_golden.py
tests/test_cpp_surface.py compiles the same source, unmodified, against include/ + libuhdr_b200.so on the
GPU box and compares the lines."""
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
objs = [o for o in glob.glob(os.path.join(ROOT, "oracle", "_ref", "obj_turbo", "*.o")) if "ref_capi" not in o]
jpeg = glob.glob(os.path.join(ROOT, "oracle", "_ref", "libjpeg-*.so.62*"))[0]
exe = "/tmp/jpegr_surface_ref"
cmd = ["g++", "-std=c++17", "-O1", "-w", "-I" + os.path.join(ROOT, "oracle", "ref_turbo"), "-I" + REF, "-I" + REF + "/lib/include",
       "-I" + REF + "/third_party/image_io/includes", os.path.join(ROOT, "tests", "cpp", "jpegr_surface_test.cpp")] + objs + \
      [jpeg, "-Wl,-rpath," + os.path.dirname(jpeg), "-lpthread", "-o", exe]
subprocess.check_call(cmd)
out = subprocess.check_output([exe, REF + "/tests/data/raw_p010_image.p010", REF + "/tests/data/raw_yuv420_image.yuv420"], text=True)
sys.stdout.write(out)
assert "surface test done" in out and "FAILED" not in out
with open(os.path.join(ROOT, "tests", "golden", "jpegr_surface_ref.txt"), "w") as f:
    f.write(out)
