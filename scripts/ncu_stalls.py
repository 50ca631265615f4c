"""Per source line: stall samples by reason (from `ncu --page source --csv --print-source cuda,sass`).  usage: ncu_stalls.py <csv> [top]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 12
hi = next(i for i, r in enumerate(rows) if 'Source' in r and any('Instructions Executed' in c for c in r))
h = rows[hi]
reasons = ['stall_wait', 'stall_long_sb', 'stall_short_sb', 'stall_no_inst', 'stall_barrier', 'stall_math', 'stall_branch_resolving',
           'stall_lg', 'stall_mio', 'stall_not_selected', 'stall_selected', 'stall_dispatch']
idx = {r: h.index(r) for r in reasons if r in h}
isrc = h.index('Source')
tot = collections.Counter()
by_line = collections.defaultdict(collections.Counter)
cur = None
for r in rows[hi + 1:]:
    if len(r) < len(h):
        continue
    src = r[isrc]
    # source rows carry the CUDA line; sass rows follow.  The csv interleaves them: a CUDA-line row has its own totals
    for k, i in idx.items():
        try:
            v = int(r[i])
        except ValueError:
            v = 0
        by_line[src][k] += v
        tot[k] += v
grand = sum(tot.values())
print("total samples", grand, {k: f"{100 * v / grand:.1f}%" for k, v in tot.most_common()})
for k in ['stall_long_sb', 'stall_wait', 'stall_no_inst', 'stall_short_sb', 'stall_barrier', 'stall_branch_resolving', 'stall_lg']:
    print("---", k)
    lines = sorted(by_line.items(), key=lambda kv: -kv[1][k])[:top]
    for src, c in lines:
        if c[k]:
            print(f"  {100 * c[k] / grand:5.2f}%  {src.strip()[:130]}")
