"""Summarise an `ncu --set full` report for profiles/: per captured launch the duration, DRAM bytes, pipe utilisation, issue
activity and the top stall reasons (read with `ncu -i REP --page raw --csv`).

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep "command the capture was taken with" > profiles/ncu_xxx.json
"""
import csv
import io
import json
import re
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pipe_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "fma_pipe_pct",
    "sm__inst_issued.avg.pct_of_peak_sustained_active": "issue_slots_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "launch__registers_per_thread": "registers",
    "launch__occupancy_limit_registers": "occ_limit_regs",
    "launch__occupancy_limit_shared_mem": "occ_limit_smem",
    "launch__occupancy_limit_warps": "occ_limit_warps",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "smsp__inst_executed.sum": "warp_instructions",
}
SCALE = {"nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main() -> None:
    rep, cmd = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, body = rows[0], rows[1], rows[2:]
    col = {n: k for k, n in enumerate(head)}
    stall_cols = [n for n in head if n.startswith("smsp__average_warps_issue_stalled_") and n.endswith("_per_issue_active.ratio")]
    launches = []
    for r in body:
        name = r[col["Kernel Name"]]
        m = re.search(r"(svp_\w+)(<[^>]*>)?", name)
        d = {"kernel": (m.group(1) + (m.group(2) or "")) if m else name[:60], "block": r[col["Block Size"]], "grid": r[col["Grid Size"]]}
        for metric, key in WANT.items():
            if metric not in col or r[col[metric]] in ("", "n/a"):
                continue
            v = float(r[col[metric]].replace(",", ""))
            u = units[col[metric]]
            if key == "duration":
                d["duration_ms"] = round(v * SCALE.get(u, 1.0), 4)
            elif key.startswith("dram_") and key != "dram_pct":
                d[key + "_gbyte"] = round(v * SCALE.get(u, 1.0) / 1e9, 6)
            else:
                d[key] = round(v, 2)
        stalls = []
        for n in stall_cols:
            if r[col[n]] not in ("", "n/a"):
                stalls.append((n[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], float(r[col[n]].replace(",", ""))))
        d["top_stalls_warps_per_issue"] = {k: round(v, 2) for k, v in sorted(stalls, key=lambda kv: -kv[1])[:5]}
        launches.append(d)
    out = {"source": cmd or rep, "report": rep, "launches": launches,
           "dram_write_gbyte_total": round(sum(x.get("dram_write_gbyte", 0) for x in launches), 6),
           "dram_read_gbyte_total": round(sum(x.get("dram_read_gbyte", 0) for x in launches), 6)}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
