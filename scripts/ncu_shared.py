"""Shared-memory wavefronts (actual / ideal / excessive = bank conflicts) per source line, from
`ncu --page source --csv --print-source cuda,sass`.  Columns are taken from the right because source text with commas
shifts the left part of a row."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = next(r for r in rows if r and r[0] == 'Line No')
n = len(hdr)
iw, ii, ie = hdr.index('L1 Wavefronts Shared') - n, hdr.index('L1 Wavefronts Shared Ideal') - n, hdr.index('L1 Wavefronts Shared Excessive') - n
agg = collections.defaultdict(lambda: [0, 0, 0, ''])
cur = None
for r in rows:
    if r and r[0] == 'File Path':
        cur = r[1].split('/')[-1]
        continue
    if len(r) >= n and r[0] not in ('', 'Line No'):
        try:
            w, i, e = int(r[iw]), int(r[ii]), int(r[ie])
        except ValueError:
            continue
        a = agg[(cur, int(r[0]))]
        a[0] += w; a[1] += i; a[2] += e; a[3] = ','.join(r[1:len(r) - n + 2]).strip()
tw = sum(a[0] for a in agg.values()); ti = sum(a[1] for a in agg.values()); te = sum(a[2] for a in agg.values())
print(f"shared wavefronts {tw}  ideal {ti}  excessive {te} ({100 * te / max(tw, 1):.1f}%)")
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    if a[0] == 0: break
    print(f"{f[:18]:18s}:{l:4d} wavefronts={100 * a[0] / tw:5.1f}%  ideal={100 * a[1] / tw:5.1f}%  excess={100 * a[2] / tw:5.1f}%  {a[3][:90]}")
