#!/usr/bin/env python
"""Join an `ncu --page source --csv` export (SASS level) with `nvdisasm -g -c` line info of the same kernel and
print executed instructions, stall samples and L2 sectors per SOURCE line.  Reads files only.
usage: python tools/ncu_lines.py source.csv kernel.sass [units]   (units: divide instruction counts, e.g. warp-steps)"""
import collections
import csv
import re
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    units = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
    hdr = rows[1]
    ci = {h: i for i, h in enumerate(hdr)}
    body = rows[2:]
    amap, loc = {}, None
    for line in open(sys.argv[2]):
        m = re.search(r'//## File "(.*)", line (\d+)', line)
        if m:
            loc = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*);", line)
        if m:
            amap[int(m.group(1), 16)] = (loc, m.group(2))
    base = int(body[0][ci["Address"]], 16)
    inst, samp, lsb, sect, noinst = (collections.Counter() for _ in range(5))
    for r in body:
        e = amap.get(int(r[ci["Address"]], 16) - base)
        if e is None:
            continue
        k = e[0]
        inst[k] += int(r[ci["Instructions Executed"]] or 0)
        samp[k] += int(r[ci["# Samples"]] or 0)
        lsb[k] += int(r[ci["stall_long_sb"]] or 0)
        noinst[k] += int(r[ci["stall_no_inst"]] or 0)
        sect[k] += int(r[ci["L2 Theoretical Sectors Global"]] or 0)
    ti, ts, tl, tsec = sum(inst.values()), sum(samp.values()), sum(lsb.values()), sum(sect.values())
    print(f"total instructions {ti} ({ti / units:.1f} per unit), samples {ts}, long_sb samples {tl}, no_inst {sum(noinst.values())}, L2 sectors {tsec}")
    print("  inst/unit   inst%  samples%  long_sb%  sectors%  line")
    for k, v in samp.most_common(40):
        print(f"{inst[k] / units:10.1f} {100 * inst[k] / ti:6.1f} {100 * v / ts:8.1f} {100 * lsb[k] / max(tl, 1):8.1f} {100 * sect[k] / max(tsec, 1):8.1f}   {k[0]}:{k[1]}")


if __name__ == "__main__":
    main()
