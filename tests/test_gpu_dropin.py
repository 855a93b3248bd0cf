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
