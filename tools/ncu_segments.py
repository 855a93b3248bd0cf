#!/usr/bin/env python3
"""Warp-sample totals of a K1 capture by code segment (contiguous SASS lines of one execution class).
  python tools/ncu_segments.py rep [min_samples]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; thr = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out))); hdr, data = rows[1], rows[2:]
ix = {k: i for i, k in enumerate(hdr)}
S = [int(r[ix['# Samples']] or 0) for r in data]; E = [int(r[ix['Instructions Executed']] or 0) for r in data]
src = [r[ix['Source']] for r in data]
stalls = [k for k in hdr if k.startswith('stall_') and 'Not Issued' not in k]
batch = max(E[i] for i in range(len(E)) if 'VOTE.ANY' in src[i])
tiles_b = max(E[i] for i in range(len(E)) if 'BAR.ARV' in src[i] or 'BAR.SYNC' in src[i])
def cls(e):
    if e == 0: return 'zero'
    if abs(e - batch) <= batch * 0.002: return 'batch'
    if e > batch * 1.5: return 'query'
    if e >= batch * 0.02 and e < batch * 0.98: return 'dup/part'
    return 'tile'
seg = []
for i, (s, e) in enumerate(zip(S, E)):
    c = cls(e)
    if seg and seg[-1][0] == c: seg[-1][2] = i; seg[-1][3] += s
    else: seg.append([c, i, i, s])
print("total samples", sum(S), "batches", batch)
for c, a, b, s in seg:
    if s >= thr:
        agg = {}
        for i in range(a, b + 1):
            for t in stalls:
                agg[t] = agg.get(t, 0) + int(data[i][ix[t]] or 0)
        top = sorted(agg.items(), key=lambda x: -x[1])[:3]
        print("%-9s %5d-%5d %8d  exec~%10d  %s  | %s" % (c, a, b, s, E[a], ' '.join('%s:%d' % (k[6:], v) for k, v in top), src[a][:40]))
