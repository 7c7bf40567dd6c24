#!/usr/bin/env python
"""Executed warp instructions and stall samples per CUDA source line of one profiled kernel.

  ncu -i X.ncu-rep --page source --csv > sass.csv          (SASS rows: address, counters)
  nvdisasm -g -c <cubin> > k.sass                           (same SASS with //## File/line notes)
  tools/ncu_lines.py sass.csv k.sass <mangled-name-substring> [top]

Rows of the two listings are matched by instruction offset inside the function.  Per line: share of the
executed warp instructions, share of the stall samples, and the three largest stall reasons of the line."""
import collections
import csv
import re
import sys


def main():
    sass_csv, disasm, needle = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    rows = list(csv.reader(open(sass_csv)))
    hdr = next(r for r in rows if "Address" in r and "Instructions Executed" in r)
    ia, ie, ism = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    stall_cols = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    data = [r for r in rows[rows.index(hdr) + 1:] if len(r) == len(hdr)]
    base = int(data[0][ia], 16)
    line_of, cur, inside = {}, None, False
    for ln in open(disasm):
        if ln.startswith(".text."):
            inside = needle in ln
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
        if m:
            line_of[int(m.group(1), 16)] = cur
    inst, samp = collections.Counter(), collections.Counter()
    reasons = collections.defaultdict(collections.Counter)
    total_reasons = collections.Counter()
    for r in data:
        key = line_of.get(int(r[ia], 16) - base, ("?", 0))
        inst[key] += int(r[ie])
        samp[key] += int(r[ism])
        for i, name in stall_cols:
            v = int(r[i] or 0)
            reasons[key][name] += v
            total_reasons[name] += v
    tot_i, tot_s = sum(inst.values()), sum(samp.values())
    print("total warp instructions %d, samples %d" % (tot_i, tot_s))
    print("stall samples by reason: " + ", ".join("%s %.1f%%" % (k, 100.0 * v / max(tot_s, 1)) for k, v in total_reasons.most_common(8)))
    order = samp if "--by-samples" in sys.argv else inst
    for key, _ in order.most_common(top):
        rs = ", ".join("%s %.1f" % (k, 100.0 * v / max(tot_s, 1)) for k, v in reasons[key].most_common(3) if v)
        print("%5.1f%% inst %5.1f%% samples  %s:%d   [%s]" % (100.0 * inst[key] / tot_i, 100.0 * samp[key] / max(tot_s, 1), key[0], key[1], rs))


if __name__ == "__main__":
    main()
