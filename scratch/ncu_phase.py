"""Aggregate an ncu SASS source page by the outermost source line of the kernel body.
usage: ncu_phase.py <annotated nvdisasm -gi of the function> <ncu --page source --csv> <kernel file basename>"""
import csv, re, sys, collections
sass, src, kfile = sys.argv[1], sys.argv[2], sys.argv[3]
off2line = {}; off2inner = {}
cur = []
for ln in open(sass):
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur.append((m.group(1).split('/')[-1], int(m.group(2))))
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
    if m:
        off = int(m.group(1), 16)
        if cur: last = cur
        outer = [l for f, l in last if f == kfile]
        off2line[off] = outer[-1] if outer else -1
        off2inner[off] = last[0]
        cur = []
rows = list(csv.reader(open(src)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[hi]
col = {n: i for i, n in enumerate(hdr)}
base = int(rows[hi + 1][0], 16)
stalls = [h for h in hdr if h.startswith('stall_') and '(Not Issued)' not in h]
agg = collections.defaultdict(lambda: collections.Counter())
inner = collections.defaultdict(lambda: collections.Counter())
for r in rows[hi + 1:]:
    off = int(r[0], 16) - base
    line = off2line.get(off, -2)
    a = agg[line]
    a['samples'] += int(r[col['# Samples']]); a['inst'] += int(r[col['Instructions Executed']])
    op = r[1].split()[0] if not r[1].strip().startswith('@') else r[1].split()[1]
    if op.startswith(('DFMA', 'DADD', 'DMUL')): a['fp64'] += int(r[col['Instructions Executed']])
    if op.startswith(('LDS', 'STS', 'LDGSTS')): a['lsu'] += int(r[col['Instructions Executed']]); a['wavefronts'] += int(r[col['L1 Wavefronts Shared']] or 0)
    for s in stalls: a[s] += int(r[col[s]] or 0)
    inner[(line, off2inner.get(off, ('?', 0)))]['samples'] += int(r[col['# Samples']])
tot = sum(a['samples'] for a in agg.values()); toti = sum(a['inst'] for a in agg.values())
print(f"total samples {tot}, warp instructions {toti}")
print(f"{'line':>5} {'samp%':>6} {'inst%':>6} {'fp64%':>6} {'lsu%':>5} {'wavef%':>6}  top stalls")
totw = sum(a['wavefronts'] for a in agg.values())
for line, a in sorted(agg.items()):
    if a['samples'] < tot * 0.003: continue
    top = sorted(((a[s], s[6:]) for s in stalls), reverse=True)[:5]
    print(f"{line:>5} {100*a['samples']/tot:6.2f} {100*a['inst']/toti:6.2f} {100*a['fp64']/max(1,a['inst']):6.1f} {100*a['lsu']/max(1,a['inst']):5.1f} {100*a['wavefronts']/max(1,totw):6.1f}  " +
          ' '.join(f"{n}:{100*v/max(1,a['samples']):.0f}" for v, n in top))
if len(sys.argv) > 4:
    print("\ninner lines with most samples:")
    for (line, inn), c in sorted(inner.items(), key=lambda kv: -kv[1]['samples'])[:40]:
        print(f"  outer {line:>4} inner {inn[0]}:{inn[1]:<4} {100*c['samples']/tot:5.2f}%")
