"""Executed-instruction histogram by SASS opcode from `ncu --page source --csv --print-source sass`."""
import csv, collections, re, sys
rows=list(csv.reader(open(sys.argv[1])))
nwarps=float(sys.argv[2]) if len(sys.argv)>2 else 1.0
hi=next(i for i,r in enumerate(rows) if r and r[0]=='Address')
h=rows[hi]; si=h.index('Source'); ii=h.index('Instructions Executed')
cnt=collections.Counter(); tot=0
for r in rows[hi+1:]:
    try: n=int(r[ii])
    except (ValueError, IndexError): continue
    s=re.sub(r'^@!?U?P\d+\s+','',r[si].strip())
    op=s.split()[0].split('.')[0] if s else '?'
    cnt[op]+=n; tot+=n
print("total", tot)
for op,n in cnt.most_common(45): print(f"{op:12s} {100*n/tot:5.1f}%  {n/nwarps:8.1f}/warp")
