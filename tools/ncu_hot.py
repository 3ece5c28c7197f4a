#!/usr/bin/env python
"""Hottest SASS instructions (warp-stall samples) from `ncu --page source --csv` output."""
import csv
import sys


def main(path, n=30):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    si = idx['# Samples']
    data = [r for r in rows[2:] if len(r) > si and r[si].isdigit()]
    tot = sum(int(r[si]) for r in data) or 1
    print("total samples", tot, "instructions", len(data))
    acc = 0
    for k, r in enumerate(data):
        r.append(k)
    for r in sorted(data, key=lambda r: -int(r[si]))[:int(n)]:
        print("%5d %7s %5.1f%%  %s" % (r[-1], r[si], 100 * int(r[si]) / tot, r[idx['Source']][:120]))


if __name__ == '__main__':
    main(*sys.argv[1:])
