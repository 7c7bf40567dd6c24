#!/usr/bin/env python
"""Executed warp instructions and stall samples per CUDA source line of one profiled kernel.

  ncu -i X.ncu-rep --page source --csv > sass.csv          (SASS rows: address, counters)
  nvdisasm -g -c <cubin> > k.sass                           (same SASS with //## File/line notes)
  tools/ncu_lines.py sass.csv k.sass <mangled-name-substring> [top]

Rows of the two listings are matched by instruction offset inside the function."""
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
    data = [r for r in rows[rows.index(hdr) + 1:] if len(r) == len(hdr)]
    base = int(data[0][ia], 16)
    by_off = {int(r[ia], 16) - base: (int(r[ie]), int(r[ism]), r[hdr.index("Source")]) for r in data}
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
    for off, (n, s, _) in by_off.items():
        key = line_of.get(off, ("?", 0))
        inst[key] += n
        samp[key] += s
    tot_i, tot_s = sum(inst.values()), sum(samp.values())
    print("total warp instructions %d, samples %d" % (tot_i, tot_s))
    for key, n in inst.most_common(top):
        print("%5.1f%% inst %5.1f%% samples  %s:%d" % (100.0 * n / tot_i, 100.0 * samp[key] / max(tot_s, 1), key[0], key[1]))


if __name__ == "__main__":
    main()
