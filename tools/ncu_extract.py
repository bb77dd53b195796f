#!/usr/bin/env python3
"""Summarise an `ncu --set full ... --page raw --csv` export: one row per captured launch with the counters DESIGN.md / bench.py quote
(duration, DRAM bytes read + written, registers, SM / FMA-pipe / DRAM throughput, the top warp-stall reasons), and -- with --traffic-json --
the per-launch DRAM traffic of the dominant kernel in the form bench.py's `roofline.traffic` reads (profiles/r2_ncu_traffic.json).

    python tools/ncu_extract.py gpurun_out/r2_tree_raw.csv --curve bls12381 --log2n 20 --csv-out profiles/r2_ncu_full_tree_kernels_2p20_bls.csv \
        --traffic-json profiles/r2_ncu_traffic.json
"""
import argparse, csv, json, os, re
ap = argparse.ArgumentParser(); ap.add_argument("raw"); ap.add_argument("--curve", default="bls12381"); ap.add_argument("--log2n", type=int, default=20)
ap.add_argument("--csv-out", default=""); ap.add_argument("--traffic-json", default="")
a = ap.parse_args()
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}
rows = list(csv.reader(open(a.raw))); hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
def val(r, name):
    i = col.get(name)
    if i is None or r[i] in ("", "n/a"): return None
    return float(r[i].replace(",", "")) * UNIT.get(units[i], 1)
WANT = [("ms", "gpu__time_duration.sum"), ("dram_bytes_read", "dram__bytes_read.sum"), ("dram_bytes_write", "dram__bytes_write.sum"),
        ("registers", "launch__registers_per_thread"), ("sm_throughput_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("fma_pipe_pct", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"), ("dram_throughput_pct", "dram__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("achieved_occupancy_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
        ("stall_long_scoreboard", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
        ("stall_math_pipe", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"),
        ("stall_wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
        ("stall_not_selected", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"),
        ("active_threads_per_inst", "smsp__thread_inst_executed_per_inst_executed.ratio")]
out = []
for r in data:
    name = r[col["Kernel Name"]]
    m = re.match(r"(?:void )?(?:b200::)?(\w+)(?:<([^>]*(?:<[^>]*>)?[^>]*)>)?", name)
    short = m.group(1) + ("<" + m.group(2) + ">" if m.group(2) else "")
    rec = {"kernel": short, "grid": r[col["Grid Size"]], "block": r[col["Block Size"]]}
    for k, n in WANT: rec[k] = val(r, n)
    out.append(rec)
if a.csv_out:
    with open(a.csv_out, "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=list(out[0].keys())); w.writeheader()
        for rec in out: w.writerow(rec)
for rec in out: print(json.dumps(rec))
if a.traffic_json:
    d = json.load(open(a.traffic_json)) if os.path.exists(a.traffic_json) else {}
    for rec in out:          # first launch of each (kernel, FIRST) = the largest round of that form
        is_first = re.search(r",\s*(?:1|true)\s*>$", rec["kernel"]) is not None
        key = "%s<FIRST=%d>|%s|2^%d" % (rec["kernel"].split("<")[0], 1 if is_first else 0, a.curve, a.log2n)
        if key not in d or d[key].get("source") != os.path.basename(a.raw) or d[key]["ms"] < rec["ms"]:
            if rec["dram_bytes_read"] is not None:
                d[key] = {"dram_bytes_read": rec["dram_bytes_read"], "dram_bytes_write": rec["dram_bytes_write"], "ms": rec["ms"], "registers": rec["registers"],
                          "sm_throughput_pct": rec["sm_throughput_pct"], "source": os.path.basename(a.raw)}
    json.dump(d, open(a.traffic_json, "w"), indent=1, sort_keys=True)
