#!/usr/bin/env python
"""Summarise one kernel of an `ncu --set full` report as JSON (profiles/<name>.json), the file
bench.py attaches to its roofline objects (load_ncu).  Runs here (no GPU needed):

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep KERNEL_SUBSTRING profiles/r2_extend_ncu.json \
        --command "python bench.py --no-extra ..." [--index N]

Values are per launch (the N-th captured launch whose name contains the substring; default the
slowest one).  dram_bytes = dram__bytes_read.sum + dram__bytes_write.sum."""
import argparse
import csv
import io
import json
import subprocess
import sys

FIELDS = {
    "duration_us": ("gpu__time_duration.sum", 1.0),
    "dram_bytes_read": ("dram__bytes_read.sum", 1.0),
    "dram_bytes_write": ("dram__bytes_write.sum", 1.0),
    "registers_per_thread": ("launch__registers_per_thread", 1.0),
    "warp_instructions": ("smsp__inst_executed.sum", 1.0),
    "sm_throughput_pct_of_elapsed": ("sm__throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
    "issue_slots_pct_while_active": ("smsp__issue_active.avg.pct_of_peak_sustained_active", 1.0),
    "warps_active_pct": ("sm__warps_active.avg.pct_of_peak_sustained_active", 1.0),
    "fp64_pipe_pct_while_active": ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", 1.0),
    "tensor_pipe_pct_of_elapsed": ("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", 1.0),
    "hmma_pipe_pct_while_active": ("sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", 1.0),
    "dram_throughput_pct": ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
    "l2_hit_pct": ("lts__t_sector_hit_rate.pct", 1.0),
    "threads_per_warp_instruction": ("smsp__thread_inst_executed_per_inst_executed.ratio", 1.0),
}
STALLS = ["long_scoreboard", "wait", "short_scoreboard", "barrier", "math_pipe_throttle", "mio_throttle",
          "branch_resolving", "no_instruction", "lg_throttle", "not_selected", "dispatch_stall", "membar", "tex_throttle"]
UNIT_SCALE = {"nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6,
              "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("kernel")
    ap.add_argument("out")
    ap.add_argument("--command", default="")
    ap.add_argument("--index", type=int, default=-1)
    a = ap.parse_args()
    txt = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    cand = [r for r in rows[2:] if a.kernel in r[col["Kernel Name"]]]
    if not cand:
        sys.exit(f"no launch of a kernel named *{a.kernel}* in {a.report}")

    def val(r, name):
        if name not in col or r[col[name]] in ("", "n/a"):
            return None
        v = float(r[col[name]].replace(",", ""))
        return v * UNIT_SCALE.get(units[col[name]], 1.0)
    r = cand[a.index] if a.index >= 0 else max(cand, key=lambda x: val(x, "gpu__time_duration.sum") or 0.0)
    out = {"kernel": r[col["Kernel Name"]], "report": a.report, "command": a.command, "launches_in_report": len(cand),
           "grid": r[col["Grid Size"]] if "Grid Size" in col else None, "block": r[col["Block Size"]] if "Block Size" in col else None}
    for k, (m, s) in FIELDS.items():
        v = val(r, m)
        out[k] = None if v is None else v * s
    if out["dram_bytes_read"] is not None and out["dram_bytes_write"] is not None:
        out["dram_bytes"] = out["dram_bytes_read"] + out["dram_bytes_write"]
    st = {}
    for s in STALLS:
        v = val(r, f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio")
        if v is not None:
            st[s] = round(v, 3)
    out["stall_warps_per_issue"] = dict(sorted(st.items(), key=lambda kv: -kv[1])[:6])
    out["note"] = "one launch under ncu --set full --clock-control none (cold caches, serialised): shares and pipe utilisation, not a bench time"
    json.dump(out, open(a.out, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
