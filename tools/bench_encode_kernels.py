"""Per-kernel CUDA-event times of one 4K API-1 encode (resident inputs, one handle): quick A/B tool.
  python tools/bench_encode_kernels.py [iterations]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import uhdr_testlib as T  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
gpu = T.Gpu()
lib = gpu.lib
T.UhdrApi(lib)
slots = []
for i in range(4):
    p, y = bench.make_frame(3840, 2160, i)
    hdr, sdr, keep = bench.frame_descs(p, y, 3840, 2160)
    s = bench.EncoderSlot(lib)
    s.set_inputs(hdr, sdr)
    slots.append((s, keep, p, y))
for s, *_ in slots:
    s.encode()
lib.uhdr_b200_set_kernel_timing(1)
for _ in range(3):
    for s, *_ in slots:
        s.rearm()
        s.encode()
bench.kernel_report(lib)
for _ in range(n):
    for s, *_ in slots:
        s.rearm()
        s.encode()
kt = bench.kernel_report(lib)
tot = 0.0
for k, v in sorted(kt.items()):
    per_frame = v[1] / (n * len(slots))
    tot += per_frame
    print("%-18s launches/frame %.1f  avg %.4f ms  per frame %.4f ms" % (k, v[0] / (n * len(slots)), v[1] / v[0], per_frame))
print("sum per frame %.4f ms" % tot)
