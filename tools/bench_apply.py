"""applyGainMap @8K kernel timing only (quick iteration on the GPU box):  python tools/bench_apply.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import uhdr_testlib as T  # noqa: E402

gpu = T.Gpu()
pk, _kind = bench.peaks()
hbm = pk["hbm_gbs"]
print(json.dumps(bench.apply_8k(gpu.lib, hbm, iters=12), indent=1))
