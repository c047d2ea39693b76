"""Summarise an `ncu --page source --csv` dump: stall samples per region of the SASS listing.
usage: ncu_src_summary.py src.csv [nbuckets]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
body = rows[2:]
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 40
ci = {h: i for i, h in enumerate(hdr)}
samp = [int(r[ci['# Samples']] or 0) for r in body]
tot = sum(samp)
stall_cols = [h for h in hdr if h.startswith('stall_')]
n = len(body)
print("instructions", n, "samples", tot)
step = (n + nb - 1) // nb
for b in range(0, n, step):
    s = sum(samp[b:b + step])
    if s < 0.01 * tot:
        continue
    # dominant stall reasons
    acc = {}
    for r in body[b:b + step]:
        for h in stall_cols:
            v = r[ci[h]]
            if v and v != '0':
                acc[h] = acc.get(h, 0) + int(v)
    top = sorted(acc.items(), key=lambda kv: -kv[1])[:3]
    ops = {}
    for r in body[b:b + step]:
        op = r[ci['Source']].split()[0] if r[ci['Source']].split() else ''
        if op.startswith('@'):
            op = r[ci['Source']].split()[1]
        ops[op.split('.')[0]] = ops.get(op.split('.')[0], 0) + 1
    topo = sorted(ops.items(), key=lambda kv: -kv[1])[:4]
    print("%5d-%5d  %5.1f%%  %s  | %s" % (b, min(n, b + step), 100.0 * s / tot, top, topo))
# top individual instructions
order = sorted(range(n), key=lambda i: -samp[i])[:25]
print("top instructions:")
for i in sorted(order):
    print("%5d %6d  %s" % (i, samp[i], body[i][ci['Source']].strip()[:90]))
