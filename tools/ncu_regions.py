#!/usr/bin/env python
"""Instruction share of consecutive SASS blocks of one kernel (which part of a long kernel body costs what):

    ncu -i rep.ncu-rep --page source --csv --print-source sass > sass.csv
    cuobjdump -xelf all lib.so ; nvdisasm -g -c file.cubin > kernel.dis
    python tools/ncu_regions.py sass.csv kernel.dis <mangled-kernel-name-substring> [instructions per block = 48]

Per block: start address, % of executed warp instructions, average active threads, dominant source lines, dominant
opcodes.  (Lines of inlined libdevice code -- sincos, atan2 -- carry unrelated line numbers of the main file.)"""
import csv, re, sys, collections
sass_csv, dis, kname = sys.argv[1:4]
loc = {}; in_k=False; cur=None
for ln in open(dis):
    if ln.startswith("//---") and ".text." in ln:
        in_k = kname in ln; continue
    if not in_k: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur=(m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*);", ln)
    if m and cur: loc[int(m.group(1),16)] = (cur, m.group(2).strip())
rows = list(csv.reader(open(sass_csv))); hdr=rows[1]
ia, ie, it = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
isrc = hdr.index("Source")
base=None; data=[]
for r in rows[2:]:
    if len(r)<=it or not r[ia]: continue
    addr=int(r[ia],16) if r[ia].startswith("0x") else int(r[ia])
    if base is None: base=addr
    data.append((addr-base, int(r[ie] or 0), int(r[it] or 0), r[isrc]))
tot=sum(d[1] for d in data)
# chunk in blocks of N instructions
N=int(sys.argv[4]) if len(sys.argv)>4 else 48
for i in range(0,len(data),N):
    blk=data[i:i+N]
    ex=sum(d[1] for d in blk); th=sum(d[2] for d in blk)
    if ex/tot < 0.004: continue
    lines=collections.Counter()
    for d in blk:
        k=loc.get(d[0]); 
        if k: lines[f"{k[0][0].replace('nm_','').replace('.cuh','').replace('.cu','')}:{k[0][1]}"]+=d[1]
    ops=collections.Counter(d[3].split()[0] if not d[3].startswith('@') else d[3].split()[1] for d in blk)
    print(f"{blk[0][0]:6x} {100*ex/tot:5.2f}% thr {th/max(ex,1):4.1f}  {dict(lines.most_common(3))}  {dict(ops.most_common(4))}")
