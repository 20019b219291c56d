#!/usr/bin/env python
"""Reads one kernel launch out of an `ncu --set full` report and (a) prints the counters DESIGN.md / profiles/ quote,
(b) records the launch's DRAM traffic next to its algorithmic bytes in profiles/r02_ncu_traffic.json, which bench.py
uses to fill `roofline.traffic`.

    python tools/ncu_traffic.py report.ncu-rep --key bkt_Cosine_1000000x768_mc8192 --alg-bytes 261625046688 --nq 10000
"""
import argparse
import csv
import io
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size",
        "sm__inst_executed_pipe_tensor.sum", "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]


def read_raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        if len(r) == len(hdr):
            out.append({h: (r[i], units[i]) for i, h in enumerate(hdr)})
    return out


def num(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--kernel", default="search_kernel")
    ap.add_argument("--key")
    ap.add_argument("--alg-bytes", type=float)
    ap.add_argument("--nq", type=int)
    ap.add_argument("--lookups", type=float, help="quantized runs: table look-ups of the launch (sum D_q x M)")
    a = ap.parse_args()
    launches = [l for l in read_raw(a.report) if a.kernel in l.get("Kernel Name", ("", ""))[0]]
    if not launches:
        raise SystemExit("no launch of %s in %s" % (a.kernel, a.report))
    l = launches[0]
    print("kernel:", l["Kernel Name"][0][:140])
    vals = {}
    for k in WANT:
        if k in l:
            v, u = l[k]
            vals[k] = (num(v), u)
            print("  %-80s %s %s" % (k, v, u))

    def scaled(name):  # ncu prints byte counters in a convenient unit
        v, u = vals[name]
        mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1.0)
        return v * mult

    dram = scaled("dram__bytes_read.sum") + scaled("dram__bytes_write.sum")
    print("  DRAM bytes per launch: %.3f GB" % (dram / 1e9))
    if a.alg_bytes:
        print("  traffic / algorithmic: %.4f" % (dram / a.alg_bytes))
    if a.key and a.alg_bytes:
        path = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
        rec = json.load(open(path)) if os.path.exists(path) else {}
        rec[a.key] = {"dram_bytes": dram, "algorithmic_bytes": a.alg_bytes, "nq": a.nq,
                      "source": "ncu --set full --clock-control none, one launch; report %s" % os.path.basename(a.report),
                      "dram_read_bytes": scaled("dram__bytes_read.sum"), "dram_write_bytes": scaled("dram__bytes_write.sum")}
        if a.lookups and "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum" in vals:
            rec[a.key]["l1_global_ld_sectors"] = vals["l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"][0]
            rec[a.key]["l1_global_ld_sectors_per_lookup"] = vals["l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"][0] / a.lookups
        with open(path, "w") as f:
            json.dump(rec, f, indent=1, sort_keys=True)
        print("  recorded under %s in %s" % (a.key, path))


if __name__ == "__main__":
    main()
