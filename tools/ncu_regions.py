#!/usr/bin/env python3
"""Per-region warp-sample totals of a K1 capture: one row per build level (the loop around each
MATCH.ANY), barrier waits, gram fill, query loop.  python tools/ncu_regions.py rep [--dump LEVEL]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, data = rows[1], rows[2:]
ix = {k: i for i, k in enumerate(hdr)}
S = [int(r[ix['# Samples']] or 0) for r in data]
E = [int(r[ix['Instructions Executed']] or 0) for r in data]
src = [r[ix['Source']] for r in data]
stalls = [k for k in hdr if k.startswith('stall_') and 'Not Issued' not in k]
tot = sum(S)
print("total samples", tot, "lines", len(data))
m_idx = [i for i, s in enumerate(src) if 'MATCH' in s]
batches = max(E[i] for i in range(len(E)) if 'VOTE.ANY' in src[i]) if any('VOTE.ANY' in s for s in src) else 0
print("batches per level", batches)
used = set()
for n, m in enumerate(m_idx[:11]):
    lo = m
    while lo > 0 and 'BAR.' not in src[lo - 1] and E[lo - 1] <= batches * 1.01 and E[lo - 1] > 0 and (lo - 1) not in used and m - lo < 70: lo -= 1
    hi = m
    while hi < len(E) - 1 and 'BAR.' not in src[hi + 1] and E[hi + 1] <= batches * 1.01 and hi - m < 45 and not (E[hi + 1] < batches * 0.5 and E[hi+1] < E[m] * 0.5): hi += 1
    used.update(range(lo, hi + 1))
    main = sum(1 for i in range(lo, hi + 1) if E[i] >= batches * 0.9)
    print("level %2d lines %4d-%4d main-path instr %3d dup-path exec %9d samples %7d (%.1f%%)" % (n + 2, lo, hi, main, E[m], sum(S[lo:hi + 1]), 100.0 * sum(S[lo:hi + 1]) / tot))
    if len(sys.argv) > 3 and int(sys.argv[3]) == n + 2:
        for i in range(lo, hi + 1):
            st = sorted(((int(data[i][ix[s]] or 0), s) for s in stalls), reverse=True)[:2]
            print(i, str(S[i]).rjust(6), str(E[i]).rjust(10), src[i][:72].ljust(72), ' '.join('%s:%d' % (s[6:], n_) for n_, s in st if n_))
rest = [(S[i], i) for i in range(len(S)) if i not in used]
print("outside level loops:", sum(s for s, _ in rest))
for s, i in sorted(rest, reverse=True)[:14]:
    st = max(((int(data[i][ix[t]] or 0), t) for t in stalls))
    print("  %7d %10d  %-70s %s" % (s, E[i], src[i][:70], st[1]))
