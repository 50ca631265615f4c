"""Per-source-line / per-region instruction and stall-sample shares from `ncu --page source --csv --print-source cuda,sass`."""
import csv, collections, sys
rows=list(csv.reader(open(sys.argv[1])))
top=int(sys.argv[2]) if len(sys.argv)>2 else 40
out=[]; cur=None
for r in rows:
    if r and r[0]=='File Path': cur=r[1].split('/')[-1]; continue
    if len(r)>8 and r[0] not in ('','Line No'):
        try: out.append((cur,int(r[0]),r[1].strip(),int(r[6]),int(r[7]),int(r[8])))
        except ValueError: pass
tot=sum(o[4] for o in out); ts=sum(o[3] for o in out); tt=sum(o[5] for o in out)
print("warp inst", tot, "thread inst", tt, "threads/inst", round(tt/tot,2))
agg=collections.defaultdict(lambda:[0,0,'',0])
for f,l,s,smp,ins,tins in out:
    a=agg[(f,l)]; a[0]+=ins; a[1]+=smp; a[2]=s; a[3]+=tins
print("--- by file")
byf=collections.defaultdict(lambda:[0,0])
for (f,l),a in agg.items(): byf[f][0]+=a[0]; byf[f][1]+=a[1]
for f,a in sorted(byf.items(), key=lambda kv:-kv[1][0]): print(f"{f:28s} inst={100*a[0]/tot:5.1f}% samp={100*a[1]/ts:5.1f}%")
print("--- top lines by stall samples")
for (f,l),a in sorted(agg.items(), key=lambda kv:-kv[1][1])[:top//2]:
    print(f"{f[:16]:16s}:{l:4d} inst={100*a[0]/tot:5.1f}% thr/inst={a[3]/max(a[0],1):5.1f} samp={100*a[1]/ts:5.1f}%  {a[2][:100]}")
print("--- top lines by instructions")
for (f,l),a in sorted(agg.items(), key=lambda kv:-kv[1][0])[:top]:
    print(f"{f[:16]:16s}:{l:4d} inst={100*a[0]/tot:5.1f}% thr/inst={a[3]/max(a[0],1):5.1f} samp={100*a[1]/ts:5.1f}%  {a[2][:100]}")
