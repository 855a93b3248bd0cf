#!/usr/bin/env python3
"""compute-sanitizer --tool racecheck workload: only the piece kernels (shared-memory hazards of the copy pass)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lzs-compression_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, lzs_b200 as B, helpers
o = helpers.oracle()
data = [helpers.corpus(kind, 1, 60000, first_index=3).tobytes() for kind in (helpers.CORPUS_TEXT, helpers.CORPUS_BINARY, helpers.CORPUS_MIXED)]
data += [b"\0" * 30000 + helpers.corpus(helpers.CORPUS_TEXT, 1, 5000, first_index=1).tobytes() + b"ab" * 5000]
comp = [o.compress(d) for d in data]
B.set_decode_piece_bytes(256)
got = B.decompress_streams(comp, [len(d) for d in data])
assert got == data
B.set_decode_piece_bytes(2048)
big = [helpers.corpus(kind, 1, 600000, first_index=5).tobytes() for kind in (helpers.CORPUS_TEXT, helpers.CORPUS_MIXED)]
assert B.decompress_streams([o.compress(d) for d in big], [len(d) for d in big]) == big      # pointer doubling (out_span >= 1 MiB)
print("racecheck workload ok")
