"""Summarise ncu reports for profiles/.
  python tools/ncu_summary.py OUT.json REPORT.ncu-rep [REPORT2.ncu-rep ...]
Reads each report with `ncu -i … --page raw --csv` and keeps, per kernel launch, the figures the
roofline discussion in DESIGN.md uses.  Also (re)writes profiles/traffic.json: dram bytes
(read + write) per launch, keyed by the kernel names bench.py reports.
"""
import csv
import io
import json
import os
import subprocess
import sys

KEEP = {
    "gpu__time_duration.sum": "time_us",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "smsp__inst_executed.sum": "warp_insts",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "pipe_fp64_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "pipe_alu_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "pipe_fma_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "pipe_xu_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "pipe_lsu_pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_sb",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_sb",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "stall_math_throttle",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait",
}
# ncu kernel-name substring -> bench.py kernel name
BENCH_NAMES = (("k_gainmap_fast<0", "gainmap_pass1"), ("k_gainmap_fast<1", "gainmap_onepass"),
               ("k_gainmap_affine", "gainmap_affine"), ("k_affine_fast", "gainmap_affine"), ("k_fdct8", "fdct_quant"), ("k_huff_encode", "huff_encode"),
               ("k_apply_lin1", "apply_gainmap"), ("k_tonemap", "tonemap"), ("k_yuv420_fast", "yuv_convert"), ("k_yuv_convert", "yuv_convert"))


def to_bytes(v, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    return float(v) * mult


def summarise(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        e = {"kernel": r[hdr.index("Kernel Name")].replace("void ", "").replace("<unnamed>::", "").replace("unnamed>::", "")}
        for i, h in enumerate(hdr):
            if h in KEEP and r[i] != "":
                v = float(r[i].replace(",", ""))
                if KEEP[h].startswith("dram_r") or KEEP[h].startswith("dram_w"):
                    v = to_bytes(v, units[i])
                elif KEEP[h] == "time_us":
                    v = v * {"ns": 1e-3, "us": 1, "ms": 1e3, "usecond": 1, "nsecond": 1e-3, "msecond": 1e3}.get(units[i], 1)
                e[KEEP[h]] = round(v, 3)
        out.append(e)
    return out


def main():
    dst, reps = sys.argv[1], sys.argv[2:]
    doc = {"how": "ncu --set full --clock-control none --import-source on (cold-cache, serialised replays; "
                  "use for shares, pipe mix and dram bytes, not absolute times)", "reports": {}}
    traffic_acc = {}
    for rep in reps:
        s = summarise(rep)
        doc["reports"][os.path.basename(rep)] = s
        for e in s:
            for sub, name in BENCH_NAMES:
                if sub in e["kernel"] and "dram_read" in e:
                    traffic_acc.setdefault(name, []).append(e["dram_read"] + e["dram_write"])
    json.dump(doc, open(dst, "w"), indent=1)
    tpath = os.path.join(os.path.dirname(dst), "traffic.json")
    traffic = {}
    if os.path.exists(tpath):
        traffic = json.load(open(tpath))
    for k, v in traffic_acc.items():
        traffic[k] = int(sum(v) / len(v))
    json.dump(traffic, open(tpath, "w"), indent=1, sort_keys=True)
    print(json.dumps(traffic))


if __name__ == "__main__":
    main()
