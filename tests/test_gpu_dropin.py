"""Drop-in proof on the GPU box: the reference's OWN test programs and file CLIs, compiled from
its unmodified sources against include/lzs.h and linked to the B200 liblzs.so (oracle/Makefile
target `dropin`, binaries in the git-ignored oracle/_ref/), must pass / interoperate with the
reference's own binaries.  Covers c/src/test/test-lzs.c, test-lzs-decompression.c and the
flow of c/src/test/test-lzs.sh:24-27."""
import os
import subprocess

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu
REF = os.path.join(helpers.ORACLE_DIR, "_ref")


def _bin(name):
    path = os.path.join(REF, name)
    if not os.path.exists(path):
        pytest.skip("%s not built (make -C oracle dropin needs the reference tree)" % name)
    return path


@pytest.mark.parametrize("prog", ["b200-test-lzs", "b200-test-lzs-decompression"])
def test_reference_unity_tests_pass_against_b200_library(prog):
    r = subprocess.run([_bin(prog)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert " 0 Failures" in r.stdout and "OK" in r.stdout, r.stdout[-500:]


def test_reference_clis_interoperate(tmp_path):
    """lzs-compress / lzs-decompress (reference sources) linked to the B200 library produce the
    same file as the reference's own build and decode each other's output."""
    data = helpers.corpus(helpers.CORPUS_MIXED, 3, 20000, seed=0x5EED0000 + 6).tobytes() + b"tail"
    src = tmp_path / "in.bin"
    src.write_bytes(data)
    out = {}
    for who in ("b200", "ref"):
        c = tmp_path / (who + ".lzs")
        subprocess.run([_bin(who + "-lzs-compress"), str(src), str(c)], check=True, timeout=600)
        out[who] = c.read_bytes()
    assert out["b200"] == out["ref"]                       # byte-identical compressed files
    for comp, dec in (("b200", "ref"), ("ref", "b200"), ("b200", "b200")):
        d = tmp_path / ("%s_by_%s.out" % (comp, dec))
        subprocess.run([_bin(dec + "-lzs-decompress"), str(tmp_path / (comp + ".lzs")), str(d)], check=True, timeout=600)
        assert d.read_bytes() == data, (comp, dec)


def test_concatenated_chunk_streams_decode_with_reference_cli(tmp_path):
    """a file made of back-to-back per-chunk streams (what the batch
    compressor emits) is decoded by the reference's own lzs-decompress, marker by marker."""
    from gpu_common import binding
    B = binding()
    chunks = [helpers.corpus(helpers.CORPUS_MIXED, 1, 8192, first_index=i).tobytes() for i in range(5)]
    streams = B.compress_streams(chunks)
    f = tmp_path / "cat.lzs"
    f.write_bytes(b"".join(streams))
    d = tmp_path / "cat.out"
    subprocess.run([_bin("ref-lzs-decompress"), str(f), str(d)], check=True, timeout=600)
    assert d.read_bytes() == b"".join(chunks)


# ---- the file CLI over the batch ABI (lzs-compression_b200/utils) ----

def _cli():
    path = os.path.join(helpers.ROOT, "lzs-compression_b200", "bin", "lzs-b200")
    if not os.path.exists(path):
        pytest.skip("lzs-b200 not built (make -C lzs-compression_b200)")
    return path


def _run(*args):
    r = subprocess.run(list(args), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (args, r.stdout[-1000:], r.stderr[-1000:])


def test_file_cli_chunked_output_is_decoded_by_the_reference_cli(tmp_path):
    """`lzs-b200 c` writes back-to-back independent chunk streams: the reference's own
    lzs-decompress restores the file, and so does `lzs-b200 d`, with the index (one batch) and
    without it (one resumable stream, like the reference)."""
    data = helpers.corpus(helpers.CORPUS_MIXED, 1, 300000, seed=0x5EED0000 + 8).tobytes() + b"odd tail"
    src = tmp_path / "in.bin"
    src.write_bytes(data)
    comp, idx = tmp_path / "in.lzs", tmp_path / "in.lzsx"
    _run(_cli(), "c", "-b", "16384", "-x", str(idx), str(src), str(comp))
    # the file is exactly the concatenation of what lzs_compress gives for every chunk
    o = helpers.oracle()
    want = b"".join(o.compress(data[i:i + 16384]) for i in range(0, len(data), 16384))
    assert comp.read_bytes() == want
    for name, cmd in (("ref", [_bin("ref-lzs-decompress"), str(comp)]),
                      ("b200_indexed", [_cli(), "d", "-x", str(idx), str(comp)]),
                      ("b200_stream", [_cli(), "d", str(comp)])):
        out = tmp_path / (name + ".out")
        _run(*cmd, str(out))
        assert out.read_bytes() == data, name


def test_file_cli_single_chunk_equals_reference_compressor(tmp_path):
    """With a chunk at least as large as the file, `lzs-b200 c` writes the very stream the
    reference's lzs-compress writes; `lzs-b200 d` decodes the reference's file; empty file too."""
    for tag, data in (("text", helpers.corpus(helpers.CORPUS_TEXT, 1, 50000, seed=0x5EED0000 + 9).tobytes()),
                      ("empty", b"")):
        src = tmp_path / (tag + ".bin")
        src.write_bytes(data)
        ours, theirs = tmp_path / (tag + ".b200.lzs"), tmp_path / (tag + ".ref.lzs")
        _run(_cli(), "c", "-b", "1048576", str(src), str(ours))
        _run(_bin("ref-lzs-compress"), str(src), str(theirs))
        assert ours.read_bytes() == theirs.read_bytes(), tag
        back = tmp_path / (tag + ".back")
        _run(_cli(), "d", str(theirs), str(back))
        assert back.read_bytes() == data, tag


def test_file_cli_single_stream_of_a_large_file_equals_reference_compressor(tmp_path):
    """`lzs-b200 c -s` on a 12 MiB file: ONE stream, compressed in parallel inside the library (cut into
    pieces, csrc/k23_pieces.cuh), byte for byte the file the reference's own lzs-compress writes -- and the
    reference's lzs-decompress restores the input from it."""
    data = b"".join(helpers.corpus(kind, 1, 3 << 20, seed=0x5EED0000 + 30 + kind).tobytes()
                    for kind in (helpers.CORPUS_TEXT, helpers.CORPUS_BINARY, helpers.CORPUS_MIXED)) + bytes(3 << 20) + b"tail"
    src = tmp_path / "big.bin"
    src.write_bytes(data)
    ours, theirs, back = tmp_path / "big.b200.lzs", tmp_path / "big.ref.lzs", tmp_path / "big.back"
    _run(_cli(), "c", "-s", str(src), str(ours))
    _run(_bin("ref-lzs-compress"), str(src), str(theirs))
    assert ours.read_bytes() == theirs.read_bytes()
    _run(_bin("ref-lzs-decompress"), str(ours), str(back))
    assert back.read_bytes() == data


def test_file_cli_decodes_long_streams_of_an_index_less_file_in_parallel(tmp_path):
    """`lzs-b200 d` on the reference's own files: ONE long stream (what lzs-compress writes), and two long
    streams followed by short ones laid end to end.  Long streams are decoded by one batch-class call each
    (csrc/k4_pieces.cuh), found by the end marker's position; the rest goes the reference CLI's way."""
    a = b"".join(helpers.corpus(kind, 1, 2 << 20, seed=0x5EED0000 + 40 + kind).tobytes()
                 for kind in (helpers.CORPUS_TEXT, helpers.CORPUS_BINARY, helpers.CORPUS_MIXED)) + bytes(1 << 20)
    b = helpers.corpus(helpers.CORPUS_MIXED, 1, 5 << 20, seed=0x5EED0000 + 44).tobytes()
    small = [helpers.corpus(helpers.CORPUS_TEXT, 1, 3000 + i, seed=0x5EED0000 + 45 + i).tobytes() for i in range(3)]
    for tag, parts in (("one", [a]), ("several", [a, b] + small)):
        comp = tmp_path / (tag + ".lzs")
        blob = b""
        for i, part in enumerate(parts):
            src, one = tmp_path / ("%s_%d.bin" % (tag, i)), tmp_path / ("%s_%d.lzs" % (tag, i))
            src.write_bytes(part)
            _run(_bin("ref-lzs-compress"), str(src), str(one))
            blob += one.read_bytes()
        comp.write_bytes(blob)
        back = tmp_path / (tag + ".back")
        _run(_cli(), "d", str(comp), str(back))
        assert back.read_bytes() == b"".join(parts), tag
