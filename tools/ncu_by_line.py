#!/usr/bin/env python
"""Attribute an ncu SASS source page to CUDA source lines.

    ncu -i rep.ncu-rep --page source --csv --print-source sass > sass.csv
    cuobjdump -xelf all lib.so ; nvdisasm -g -c kernel.cubin > kernel.dis
    python tools/ncu_by_line.py sass.csv kernel.dis <kernel-name-substring> [top]

Sums "Instructions Executed", thread instructions and stall samples per (file, line) -- outermost inlining
site inside the kernel's own file is reported as `via`."""
import collections
import csv
import re
import sys


def main():
    sass_csv, dis, kname = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    # address -> (file, line, via) from nvdisasm -g
    loc = {}
    in_k = False
    cur = None
    for ln in open(dis):
        if ln.startswith("//---") and ".text." in ln:
            in_k = kname in ln
            continue
        if not in_k:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            via = re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3))
            cur = (m.group(1).split("/")[-1], int(m.group(2)), tuple((f.split("/")[-1], int(l)) for f, l in via))
            continue
        m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*);", ln)
        if m and cur:
            loc[int(m.group(1), 16)] = (cur, m.group(2).strip())
    rows = list(csv.reader(open(sass_csv)))
    hdr = rows[1]
    ia, ie, it, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index(
        "Thread Instructions Executed"), hdr.index("# Samples")
    base = None
    agg = collections.defaultdict(lambda: [0, 0, 0])
    tot = [0, 0, 0]
    for r in rows[2:]:
        if len(r) <= isamp or not r[ia]:
            continue
        addr = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
        if base is None:
            base = addr
        key = loc.get(addr - base)
        ex, th, sm = int(r[ie] or 0), int(r[it] or 0), int(r[isamp] or 0)
        tot[0] += ex; tot[1] += th; tot[2] += sm
        if key is None:
            k = ("?", 0, ())
        else:
            k = key[0]
        # report at the outermost site in the main file when inlined
        site = k[2][-1] if k[2] else (k[0], k[1])
        agg[(site, (k[0], k[1]))][0] += ex
        agg[(site, (k[0], k[1]))][1] += th
        agg[(site, (k[0], k[1]))][2] += sm
    print(f"total: inst {tot[0]:.3e}  thread-inst {tot[1]:.3e}  samples {tot[2]}")
    by_site = collections.defaultdict(lambda: [0, 0, 0])
    for (site, inner), v in agg.items():
        for i in range(3):
            by_site[site][i] += v[i]
    print("--- by outermost site")
    for site, v in sorted(by_site.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{site[0]}:{site[1]:<5d} inst {100*v[0]/tot[0]:5.1f}%  samples {100*v[2]/max(tot[2],1):5.1f}%  avg-threads {v[1]/max(v[0],1):5.1f}")
    print("--- by innermost line")
    by_in = collections.defaultdict(lambda: [0, 0, 0])
    for (site, inner), v in agg.items():
        for i in range(3):
            by_in[inner][i] += v[i]
    for inner, v in sorted(by_in.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{inner[0]}:{inner[1]:<5d} inst {100*v[0]/tot[0]:5.1f}%  samples {100*v[2]/max(tot[2],1):5.1f}%  avg-threads {v[1]/max(v[0],1):5.1f}")


if __name__ == "__main__":
    main()
