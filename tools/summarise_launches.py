#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
--csv` launch list (tools/profile_eval.py --evals 2) per kernel for the LAST evaluation, and
write profiles/r02_ncu_traffic.json for config 2 (keyed by the library's source digest, which
bench.py checks before quoting `roofline.traffic`).

    python tools/summarise_launches.py gpurun_out/r02_config2_launches.csv [--write-traffic STAMP]
"""
import collections
import csv
import json
import re
import sys


def load(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 8]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    per = collections.OrderedDict()
    for r in rows[1:]:
        key = int(r[ix["ID"]])
        rec = per.setdefault(key, {"name": r[ix["Kernel Name"]]})
        try:
            rec[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
            rec["unit:" + r[ix["Metric Name"]]] = r[ix["Metric Unit"]]
        except ValueError:
            pass
    return list(per.values())


def to_bytes(v, unit):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def to_us(v, unit):
    return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)


def main():
    path = sys.argv[1]
    launches = load(path)
    half = len(launches) // 2
    last = launches[half:]           # second evaluation
    agg = collections.OrderedDict()
    for l in last:
        m = re.search(r"([A-Za-z_][A-Za-z0-9_]*(?:<[^>()]*>)?)\(", l["name"].replace("(int)", "")
                      .replace("(bool)", ""))
        name = m.group(1) if m else l["name"][:60]
        a = agg.setdefault(name, dict(n=0, us=0.0, rd=0.0, wr=0.0, dmma=0.0))
        a["n"] += 1
        us = to_us(l.get("gpu__time_duration.sum", 0.0), l.get("unit:gpu__time_duration.sum", "ns"))
        a["us"] += us
        a["rd"] += to_bytes(l.get("dram__bytes_read.sum", 0.0), l.get("unit:dram__bytes_read.sum"))
        a["wr"] += to_bytes(l.get("dram__bytes_write.sum", 0.0), l.get("unit:dram__bytes_write.sum"))
        a["dmma"] += us * l.get("sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", 0.0)
    total = sum(a["us"] for a in agg.values())
    print("| kernel | launches | time (us) | share | DRAM read (GB) | DRAM written (GB) | GB/s | DMMA pipe busy |")
    print("|---|---|---|---|---|---|---|---|")
    for name, a in agg.items():
        gbs = (a["rd"] + a["wr"]) / (a["us"] * 1e-6) / 1e9 if a["us"] else 0.0
        print("| `%s` | %d | %.1f | %.1f %% | %.3f | %.3f | %.0f | %s |" % (
            name[:70], a["n"], a["us"], 100 * a["us"] / total, a["rd"] / 1e9, a["wr"] / 1e9, gbs,
            ("%.0f %%" % (a["dmma"] / a["us"])) if a["dmma"] else "-"))
    print("| total | %d | %.1f | | %.3f | %.3f | | |" % (
        sum(a["n"] for a in agg.values()), total, sum(a["rd"] for a in agg.values()) / 1e9,
        sum(a["wr"] for a in agg.values()) / 1e9))
    if "--write-traffic" in sys.argv:
        stamp = sys.argv[sys.argv.index("--write-traffic") + 1]
        pre = [a for n, a in agg.items() if n.startswith("bwd4_")]
        post = [a for n, a in agg.items() if n.startswith("fwd4") or n.startswith("cherry")]
        rec = {"lib_stamp": stamp, "source": path,
               "preorder_bytes_per_step": sum(a["rd"] + a["wr"] for a in pre),
               "preorder_launches": sum(a["n"] for a in pre),
               "preorder_us_under_ncu": sum(a["us"] for a in pre),
               "postorder_bytes_per_step": sum(a["rd"] + a["wr"] for a in post),
               "postorder_us_under_ncu": sum(a["us"] for a in post),
               "all_kernels_bytes_per_step": sum(a["rd"] + a["wr"] for a in agg.values()),
               "all_kernels_us_under_ncu": total}
        json.dump(rec, open("profiles/r02_ncu_traffic.json", "w"), indent=1)
        print(json.dumps(rec))


if __name__ == "__main__":
    main()
