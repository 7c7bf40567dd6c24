#!/usr/bin/env python
"""Stall samples and executed instructions of one profiled kernel per barrier-delimited code segment.

  ncu -i X.ncu-rep --page source --csv --print-source sass > sass.csv
  tools/ncu_segments.py sass.csv [min_share_percent]

The SASS listing is cut at every BAR.SYNC; a segment is one phase of the kernel between two barriers (the
static counts of LDS / STS / LDG / STG / MUFU / SHFL instructions identify it)."""
import csv
import re
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
    hdr = next(r for r in rows if "Address" in r and "Instructions Executed" in r)
    col = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[rows.index(hdr) + 1:] if len(r) == len(hdr)]
    tot = sum(int(r[col["# Samples"]]) for r in data)
    toti = sum(int(r[col["Instructions Executed"]]) for r in data)
    kinds = ("LDS", "STS", "LDG", "STG", "MUFU", "SHFL", "LDL", "STL")

    def new(i):
        d = {"start": i, "n": 0, "s": 0, "i": 0, "ssb": 0, "lsb": 0, "bar": 0, "wf": 0, "wfx": 0, "exec": 0, "name": ""}
        d.update({k: 0 for k in kinds})
        return d
    segs, cur = [], new(0)
    for idx, r in enumerate(data):
        src = r[col["Source"]]
        cur["n"] += 1
        cur["s"] += int(r[col["# Samples"]])
        cur["i"] += int(r[col["Instructions Executed"]])
        cur["ssb"] += int(r[col["stall_short_sb"]] or 0)
        cur["lsb"] += int(r[col["stall_long_sb"]] or 0)
        cur["bar"] += int(r[col["stall_barrier"]] or 0)
        cur["wf"] += int(r[col["L1 Wavefronts Shared"]] or 0)
        cur["wfx"] += int(r[col["L1 Wavefronts Shared Excessive"]] or 0)
        for k in kinds:
            if re.search(r"\b" + k, src):
                cur[k] += 1
        if "BAR.SYNC" in src:
            cur["exec"] = int(r[col["Instructions Executed"]])
            cur["name"] = src.strip().split(";")[0][-12:]
            segs.append(cur)
            cur = new(idx + 1)
    segs.append(cur)
    totwf = sum(s["wf"] for s in segs)
    print("samples %d, warp instructions %d, shared wavefronts %d" % (tot, toti, totwf))
    print("start    n | samples  inst  | short_sb long_sb barrier | wavefronts (excess) | lds sts ldg stg mufu shfl ldl stl | barrier executions")
    for s in segs:
        if 100.0 * s["s"] / tot < thr:
            continue
        print("%5d %4d | %5.1f%% %5.1f%% | %5.1f %5.1f %5.1f | %5.1f%% (%4.1f%%) | %3d %3d %3d %3d %3d %3d %3d %3d | %d %s"
              % (s["start"], s["n"], 100.0 * s["s"] / tot, 100.0 * s["i"] / toti, 100.0 * s["ssb"] / tot, 100.0 * s["lsb"] / tot,
                 100.0 * s["bar"] / tot, 100.0 * s["wf"] / max(totwf, 1), 100.0 * s["wfx"] / max(totwf, 1),
                 s["LDS"], s["STS"], s["LDG"], s["STG"], s["MUFU"], s["SHFL"], s["LDL"], s["STL"], s["exec"], s["name"]))


if __name__ == "__main__":
    main()
