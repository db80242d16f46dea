#!/usr/bin/env python
"""Summarise ncu output for profiles/: `ncu_summary.py raw <rep> <kernel-regex> <out.json>` pulls the headline
counters of every captured launch from a --set full report; `ncu_summary.py launches <csv> <out.json>` groups a
`--metrics gpu__time_duration.sum` launch list by kernel; `ncu_summary.py sass <rep> <out.json> <units>` counts
executed warp instructions per opcode (divided by `units`, e.g. the triplets per launch)."""
import collections
import csv
import io
import json
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.sum", "lts__t_bytes.sum"]


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def raw(rep, regex, dst):
    rows = ncu_csv(rep, "raw")
    hdr, units, data = rows[0], rows[1], rows[2:]
    ik = hdr.index("Kernel Name")
    launches = []
    for d in data:
        if not re.search(regex, d[ik]):
            continue
        rec = {"kernel": d[ik].split("(")[0]}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                try:
                    rec[k] = {"value": float(d[i].replace(",", "")), "unit": units[i]}
                except ValueError:
                    pass
        stalls = {}
        for i, h in enumerate(hdr):
            if "issue_stalled_" in h and h.endswith("per_issue_active.ratio"):
                try:
                    v = float(d[i].replace(",", ""))
                except ValueError:
                    continue
                if v >= 0.2:
                    stalls[h.split("issue_stalled_")[-1].replace("_per_issue_active.ratio", "")] = round(v, 3)
        rec["warp_stalls_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1]))
        launches.append(rec)
    json.dump({"report": rep, "launches": launches}, open(dst, "w"), indent=1)
    print(f"{len(launches)} launch(es) -> {dst}")


def launches(path, dst):
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(list)
    for r in rows[1:]:
        if len(r) > iv:
            try:
                agg[r[ik].split("(")[0]].append(float(r[iv].replace(",", "")) / 1e3)
            except ValueError:
                pass
    out = [{"kernel": k, "launches": len(v), "avg_us": sum(v) / len(v), "total_us": sum(v)}
           for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))]
    json.dump({"source": path, "kernels": out}, open(dst, "w"), indent=1)
    for o in out[:8]:
        print(f"{o['kernel'][:60]:60s} n={o['launches']:4d} avg={o['avg_us']:9.1f} us")


def sass(rep, dst, units):
    rows = ncu_csv(rep, "source", ["--print-source", "sass"])
    hdr = rows[1]
    isrc, ins, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    ops, samples = collections.Counter(), collections.Counter()
    for r in rows[2:]:
        if len(r) < 10 or r[0] == "Kernel Name":
            break
        toks = r[isrc].split()
        op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
        ops[op] += int(r[ins])
        samples[op] += int(r[isamp])
    tot, ts = sum(ops.values()), max(1, sum(samples.values()))
    out = {"report": rep, "kernel": rows[0][1].split("(")[0], "units_per_launch": units,
           "warp_instructions_per_unit": tot / units,
           "by_opcode": {k: {"per_unit": round(v / units, 3), "stall_samples_pct": round(100 * samples[k] / ts, 2)}
                         for k, v in ops.most_common(24)}}
    json.dump(out, open(dst, "w"), indent=1)
    print(f"{tot / units:.1f} warp instructions per unit -> {dst}")


if __name__ == "__main__":
    if sys.argv[1] == "raw":
        raw(sys.argv[2], sys.argv[3], sys.argv[4])
    elif sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        sass(sys.argv[2], sys.argv[3], float(sys.argv[4]))
