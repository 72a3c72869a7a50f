"""Where does the end-to-end encode lose PCIe bandwidth?  Same threads / handles / pinned frames as bench.py's e2e arm:
  (a) uploads only: reset + uhdr_enc_set_raw_image x2 per frame, no encode
  (b) uploads + encode (the bench's e2e step)
for several slot counts.  Prints GB/s of host->device traffic."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import torch  # noqa: E402

so = os.path.join(ROOT, "libultrahdr_b200", "libuhdr_b200.so")
api, lib = bench.load_api(so)
F = 32
frames, pins = [], []
for i in range(F):
    p, y = bench.make_frame(bench.W4K, bench.H4K, i)
    tp, ty = torch.from_numpy(p).pin_memory(), torch.from_numpy(y).pin_memory()
    pins.append((tp, ty))
    frames.append((tp.numpy(), ty.numpy()))
descs = [bench.frame_descs(p, y, bench.W4K, bench.H4K) for (p, y) in frames]
nbytes = sum(p.nbytes + y.nbytes for (p, y) in frames)
for slots in (2, 4, 8, 12):
    sl = [bench.EncoderSlot(lib) for _ in range(slots)]
    for mode in ("upload", "upload+encode"):
        def step():
            def work(s):
                for i in range(s, F, slots):
                    sl[s].reset()
                    sl[s].set_inputs(descs[i][0], descs[i][1])
                    if mode != "upload":
                        sl[s].encode()
            bench.run_threads(slots, work)
        step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 4
        for _ in range(n):
            step()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n
        print("slots %2d  %-14s %6.2f ms/step  %5.1f GB/s h2d" % (slots, mode, dt * 1e3, nbytes / dt / 1e9))
