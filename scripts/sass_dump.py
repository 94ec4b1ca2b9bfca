#!/usr/bin/env python
"""List the SASS lines of an `ncu --page source --csv` export with their execution counts (hot lines only)."""
import csv, sys
path = sys.argv[1]; thresh = float(sys.argv[2]) if len(sys.argv) > 2 else 0.002
rows=list(csv.reader(open(path)))
hi=next(i for i,r in enumerate(rows) if "Source" in r and "# Samples" in r)
hdr=rows[hi]; ia,isamp,iex,ith=hdr.index("Source"),hdr.index("# Samples"),hdr.index("Instructions Executed"),hdr.index("Avg. Threads Executed")
data=[r for r in rows[hi+1:] if len(r)>ith and r[isamp].isdigit()]
seen=set();u=[]
for r in data:
    if r[0] in seen: break
    seen.add(r[0]);u.append(r)
tot=sum(int(r[iex]) for r in u); ts=sum(int(r[isamp]) for r in u)
print("total warp-instructions", tot, "samples", ts)
for i,r in enumerate(u):
    if int(r[iex])>tot*thresh: print(i, r[ia].strip()[:70].ljust(70), r[iex].rjust(10), r[ith].rjust(3), r[isamp].rjust(6))
