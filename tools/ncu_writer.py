"""Stall samples of the writer warps of k_patch_ws from an `ncu --page source --csv` dump (everything after
USETMAXREG.DEALLOC): per region (split at barriers / mbarrier waits) and the top instructions.
usage: ncu_writer.py src.csv [min_samples]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, body = rows[1], rows[2:]
ci = {h: i for i, h in enumerate(hdr)}
mins = int(sys.argv[2]) if len(sys.argv) > 2 else 40
samp = [int(r[ci['# Samples']] or 0) for r in body]
start = [i for i, r in enumerate(body) if 'USETMAXREG.DEALLOC' in r[ci['Source']]][0]
tot = sum(samp)
print("all samples", tot, "writer samples", sum(samp[start:]), "compute samples", sum(samp[:start]))
stall = [h for h in hdr if h.startswith('stall_') and 'Not' not in h]
acc = {}
for i in range(start, len(body)):
    for h in stall:
        v = body[i][ci[h]]
        if v and v != '0':
            acc[h] = acc.get(h, 0) + int(v)
print("writer stall totals:", sorted(acc.items(), key=lambda kv: -kv[1])[:8])
# regions
reg0, racc = start, 0
for i in range(start, len(body)):
    src = body[i][ci['Source']]
    racc += samp[i]
    if 'BAR.SYNC' in src or 'SYNCS.PHASECHK' in src or i == len(body) - 1:
        if racc > 0.01 * tot:
            print("  region %d-%d: %d samples (%.1f%% of writer)" % (reg0, i, racc, 100.0 * racc / max(1, sum(samp[start:]))))
        reg0, racc = i + 1, 0
for i in range(start, len(body)):
    if samp[i] >= mins:
        r = body[i]
        st = {h[6:]: r[ci[h]] for h in stall if r[ci[h]] not in ('', '0')}
        st = dict(sorted(st.items(), key=lambda kv: -int(kv[1]))[:3])
        print(i, samp[i], r[ci['Source']].strip()[:60], st, r[ci['L1 Wavefronts Shared']], r[ci['L1 Wavefronts Shared Ideal']])
