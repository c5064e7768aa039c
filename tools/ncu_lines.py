"""tools/ncu_lines.py <csv from `ncu --page source --csv --print-source cuda,sass`> [top]
Aggregates warp-stall samples and executed instructions per CUDA source line."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
def num(x):
    try: return int(x)
    except ValueError: return 0
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
agg = collections.defaultdict(lambda: [0, 0, 0.0]); text = {}
stall = collections.defaultdict(collections.Counter)
cur = None; idx = {}
KEYS = ['stall_long_sb', 'stall_wait', 'stall_math', 'stall_short_sb', 'stall_branch_resolving', 'stall_no_inst',
        'stall_not_selected', 'stall_selected', 'stall_lg', 'stall_mio', 'stall_dispatch', 'stall_barrier', 'stall_sleep', 'stall_membar']
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No':
        idx = {}
        for i, h in enumerate(r): idx.setdefault(h, i)
        continue
    try: ln = int(r[0])
    except ValueError: continue
    text[(cur, ln)] = r[1]
    s = num(r[idx['# Samples']]); ie = num(r[idx['Instructions Executed']])
    te = num(r[idx['Thread Instructions Executed']])
    a = agg[(cur, ln)]; a[0] += s; a[1] += ie; a[2] += te
    for k in KEYS:
        if k in idx: stall[(cur, ln)][k] += num(r[idx[k]])
tot = sum(v[0] for v in agg.values()); toti = sum(v[1] for v in agg.values())
print("total samples", tot, "warp instructions", toti)
allst = collections.Counter()
for c in stall.values(): allst.update(c)
print("stalls:", [(k[6:], round(100 * v / tot, 1)) for k, v in allst.most_common(8)])
for (f, ln), v in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    st = [(k[6:], round(100 * c / tot, 1)) for k, c in stall[(f, ln)].most_common(3)]
    thr = v[2] / v[1] if v[1] else 0
    print(f"{f}:{ln:4d} {100*v[0]/tot:5.1f}% inst {100*v[1]/toti:5.1f}% thr {thr:4.1f} {st} | {text[(f, ln)].strip()[:90]}")
