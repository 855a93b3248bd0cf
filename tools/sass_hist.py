#!/usr/bin/env python3
"""Opcode histogram per kernel of a built library (cuobjdump -sass), so that what the sm_100a code
uses -- SYNCS (mbarrier), ATOMS (shared-memory atomics), UBLKCP (cp.async.bulk / TMA 1-D), LDGSTS
(cp.async), UTMALDG (tensor TMA), HMMA / UTCMMA (tensor cores) -- is on file beside the ncu text.
  python tools/sass_hist.py lzs-compression_b200/liblzs.so [variants/bulk.so ...] > profiles/r2_sass_opcodes.txt"""
import collections, re, subprocess, sys

MARK = ["SYNCS", "ATOMS", "UBLKCP", "LDGSTS", "UTMALDG", "UTMASTG", "HMMA", "UTCMMA", "UTCHMMA", "MATCH", "REDUX", "SHFL", "VOTE", "NANOSLEEP", "BAR", "MEMBAR", "LDS", "STS", "LDG", "STG"]


def main():
    for lib in sys.argv[1:]:
        out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
        arch = re.findall(r"arch = (sm_\w+)", out)
        print("# %s  (code objects: %s)" % (lib, ", ".join(sorted(set(arch)))))
        kernels = re.split(r"\n\s*Function : ", out)[1:]
        for k in kernels:
            name = k.split("\n", 1)[0].strip()
            demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
            ops = collections.Counter()
            for m in re.finditer(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)", k):
                ops[m.group(1).split(".")[0]] += 1
            total = sum(ops.values())
            if total == 0:
                continue
            short = demangled.replace("(anonymous namespace)::", "")
            short = re.sub(r"\(.*", "", short)
            print("\n## %s   (%d SASS instructions)" % (short, total))
            print("   marked: " + "  ".join("%s %d" % (m, ops[m]) for m in MARK if ops.get(m)))
            print("   top:    " + "  ".join("%s %d" % kv for kv in ops.most_common(14)))
        print()


if __name__ == "__main__":
    main()
