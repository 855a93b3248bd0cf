#!/usr/bin/env python3
"""Regenerate tests/golden/*.bin.  Run in the build container only.

Needs /root/reference (read-only) and oracle/_ref/liblzs_ref.so (make -C oracle);
neither exists on the GPU box, which is why the outputs are committed.

  golden1_compressed.bin / golden1_plain.bin
      the reference's own golden vector, parsed out of
      c/src/test/test-lzs-decompression.c:34-96 (324-byte stream, 507-byte text;
      all nested '#if 1' blocks active, the '#if 0' block skipped)
  uncompressible.bin
      the 506-byte digram-free sequence of c/src/test/test-lzs.c:44-66
  ref_cases.npz
      inputs from tests/helpers.edge_case_inputs() plus small seeded corpora, each
      with the unmodified reference's lzs_compress output and, for a few damaged
      streams, its lzs_decompress output
"""
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import helpers  # noqa: E402

REF_TESTS = "/root/reference/c/src/test"


def active_lines(text):
    """Tiny preprocessor: honour '#if 0' / '#if 1' / '#endif' nesting."""
    out, stack = [], []
    for line in text.splitlines():
        s = line.strip()
        if s.startswith("#if"):
            stack.append(s.split()[1] != "0")
            continue
        if s.startswith("#endif"):
            stack.pop()
            continue
        if all(stack):
            out.append(line)
    return "\n".join(out)


def c_string_literals(block):
    return "".join(re.findall(r'"((?:[^"\\]|\\.)*)"', block))


def main():
    src = open(os.path.join(REF_TESTS, "test-lzs-decompression.c")).read()
    m = re.search(r"compressed_data_1\[\]\s*=\s*\{(.*?)\};", src, re.S)
    comp = bytes(int(x, 16) for x in re.findall(r"0x([0-9A-Fa-f]{2})", active_lines(m.group(1))))
    m = re.search(r"decompressed_data_1\[\]\s*=(.*?);", src, re.S)
    plain = c_string_literals(m.group(1)).encode("ascii")
    assert len(comp) == 324 and len(plain) == 507, (len(comp), len(plain))
    open(os.path.join(HERE, "golden1_compressed.bin"), "wb").write(comp)
    open(os.path.join(HERE, "golden1_plain.bin"), "wb").write(plain)

    src = open(os.path.join(REF_TESTS, "test-lzs.c")).read()
    m = re.search(r"uncompressible_sequence\[\]\s*=(.*?);", src, re.S)
    seq = c_string_literals(re.sub(r"//.*", "", m.group(1))).encode("ascii")
    assert len(seq) == 506, len(seq)
    open(os.path.join(HERE, "uncompressible.bin"), "wb").write(seq)

    ref = helpers.reference()
    assert ref is not None, "build oracle/_ref first: make -C oracle"
    assert ref.compress(plain) == comp, "reference lzs_compress must reproduce the golden stream"

    arrays = {}
    cases = dict(helpers.edge_case_inputs())
    for kind, name in ((helpers.CORPUS_TEXT, "text"), (helpers.CORPUS_BINARY, "binary"),
                       (helpers.CORPUS_RANDOM, "random"), (helpers.CORPUS_PACKET, "packet")):
        cases["corpus_%s_5000" % name] = helpers.corpus(kind, 1, 5000, seed=0x5EED0000 + 9).tobytes()
    for name, data in cases.items():
        arrays["in__" + name] = np.frombuffer(data, dtype=np.uint8)
        arrays["out__" + name] = np.frombuffer(ref.compress(data), dtype=np.uint8)

    # damaged / unusual streams for the decoder
    base = ref.compress(cases["corpus_text_5000"])
    damaged = {
        "truncated_3": base[:-3],
        "truncated_half": base[: len(base) // 2],
        "trailing_garbage": base + b"\xAA\x55\xFF",
        "two_streams": base + base,
        "offset_before_start": bytes([0xC0 | 0x0A, 0x80 | 0x40, 0x00, 0xC0, 0x00]),
        "long_offset_zero": bytes([0x80, 0x00, 0x20, 0xD8, 0x00, 0x00]),
        "only_marker": b"\xC0\x00",
        "empty": b"",
        "all_ones": b"\xFF" * 40,
        "all_zero": b"\x00" * 40,
    }
    rng = np.random.default_rng(77)
    for k in range(8):
        damaged["noise_%d" % k] = rng.integers(0, 256, 200, dtype=np.uint8).tobytes()
    for name, stream in damaged.items():
        for cap in (100, 8000):
            arrays["dec_in__%s__%d" % (name, cap)] = np.frombuffer(stream, dtype=np.uint8)
            arrays["dec_out__%s__%d" % (name, cap)] = np.frombuffer(ref.decompress(stream, cap), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "ref_cases.npz"), **arrays)
    print("wrote golden fixtures:", len(cases), "compress cases,", len(damaged) * 2, "decode cases")


if __name__ == "__main__":
    main()
