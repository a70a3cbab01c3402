#!/usr/bin/env python
"""Summarises an `ncu --set full` capture of k_tile_score into profiles/<name>.json: the limiter figures bench.py
attaches to its `roofline` object (issue-active %, shared-memory wavefront %, DRAM fraction, DRAM bytes per launch,
stall mix).  The summary records the SHA-256 of the kernel's source file so that bench.py can tell whether the profile
still describes the shipped code.  Usage: python benchmarks/ncu_limiter.py gpurun_out/prof.ncu-rep profiles/r2_tile_limiter.json
"""
import csv
import hashlib
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "document_retrieval_b200", "csrc", "br_tile.cu")


def source_sha():
    return hashlib.sha256(open(SRC, "rb").read()).hexdigest()


def main(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    h, v = rows[0], rows[-1]

    def g(name, default=None):
        try:
            return float(v[h.index(name)].replace(",", ""))
        except (ValueError, IndexError):
            return default

    unit = {k: rows[1][i] for i, k in enumerate(h)}

    def to_bytes(name):
        x, u = g(name), unit.get(name, "")
        return None if x is None else x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)

    stalls = {k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]: g(k)
              for k in h if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")}
    top = dict(sorted(((k, round(x, 3)) for k, x in stalls.items() if x), key=lambda kv: -kv[1])[:6])
    grid = [int(float(v[h.index(k)])) for k in ("launch__grid_size",)][0]
    res = {
        "kernel": v[h.index("Kernel Name")], "report": os.path.basename(rep), "source": os.path.relpath(SRC, ROOT),
        "source_sha256": source_sha(), "grid_ctas": grid,
        "duration_ms": g("gpu__time_duration.sum") * {"ms": 1, "us": 1e-3, "s": 1e3}.get(unit.get("gpu__time_duration.sum", "ms"), 1),
        "inst_executed": g("smsp__inst_executed.sum"),
        "issue_active_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "smem_wavefront_pct": g("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
        "dram_frac": (g("dram__bytes_read.sum.pct_of_peak_sustained_elapsed", 0) + g("dram__bytes_write.sum.pct_of_peak_sustained_elapsed", 0)) / 100.0,
        "dram_bytes": (to_bytes("dram__bytes_read.sum") or 0) + (to_bytes("dram__bytes_write.sum") or 0),
        "l2_hit_pct": g("lts__t_sector_hit_rate.pct"),
        "warps_active_pct": g("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "registers_per_thread": g("launch__registers_per_thread"),
        "stall_cycles_per_issue": top,
        "note": "one launch of the tile kernel captured with ncu --set full --clock-control none; per-launch figures",
    }
    with open(out, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
