"""Pin the CPU restatement (oracle/lzs_oracle.c) before anything trusts it.

Sources of truth, strongest first:
  * the reference's golden vector and known-answer size laws
    (c/src/test/test-lzs-decompression.c:34-96, c/src/test/test-lzs.c:93-167);
  * committed outputs of the unmodified reference (tests/golden/ref_cases.npz);
  * the unmodified reference itself (oracle/_ref), when the prebuilt .so is present.
"""
import os

import numpy as np
import pytest

import helpers
from helpers import GOLDEN_DIR, compressed_max


def _golden(name):
    return open(os.path.join(GOLDEN_DIR, name), "rb").read()


def _ref_cases():
    return np.load(os.path.join(GOLDEN_DIR, "ref_cases.npz"))


def test_golden_vector_decode_and_encode():
    comp, plain = _golden("golden1_compressed.bin"), _golden("golden1_plain.bin")
    o = helpers.oracle()
    assert o.decompress(comp, len(plain) + 520) == plain
    assert o.compress(plain) == comp


def length_bits(rep):
    """c/src/test/test-lzs.c:73-87."""
    if rep == 0:
        return 0
    if rep == 1:
        return 9
    if rep <= 4:
        return 2
    if rep <= 7:
        return 4
    return ((rep + 22) // 15) * 4


def test_uncompressible_size_law():
    """c/src/test/test-lzs.c:93-119: all literals + end marker, and round trip."""
    seq = _golden("uncompressible.bin")
    o = helpers.oracle()
    for n in range(len(seq) + 1):
        c = o.compress(seq[:n], 1000)
        assert len(c) == (n * 9 + 9 + 7) // 8
        assert o.decompress(c, 1000) == seq[:n]


def test_repeated_byte_size_law():
    """c/src/test/test-lzs.c:121-167."""
    o = helpers.oracle()
    for n in range(0, 1001):
        data = b"X" * n
        if n == 0:
            bits = 0
        elif n == 1:
            bits = 9
        elif n == 2:
            bits = 18
        else:
            bits = 9 + 2 + 7 + length_bits(n - 1)
        c = o.compress(data, 1000)
        assert len(c) == (bits + 9 + 7) // 8, n
        assert o.decompress(c, 1000) == data


def test_known_tiny_streams():
    """SURVEY.md Appendix A probes: empty -> C0 00, 'a' -> 30 E0 00."""
    o = helpers.oracle()
    assert o.compress(b"") == bytes([0xC0, 0x00])
    assert o.compress(b"a") == bytes([0x30, 0xE0, 0x00])


def test_committed_reference_outputs_compress():
    z = _ref_cases()
    o = helpers.oracle()
    names = [k[4:] for k in z.files if k.startswith("in__")]
    assert len(names) >= 20
    for name in names:
        data = z["in__" + name].tobytes()
        want = z["out__" + name].tobytes()
        assert o.compress(data) == want, name
        assert o.decompress(want, len(data) + 10) == data, name
        assert len(want) <= compressed_max(len(data))


def test_committed_reference_outputs_decode_damaged():
    z = _ref_cases()
    o = helpers.oracle()
    keys = [k for k in z.files if k.startswith("dec_in__")]
    assert len(keys) >= 30
    for k in keys:
        _, name, cap = k.split("__")
        want = z["dec_out__%s__%s" % (name, cap)].tobytes()
        assert o.decompress(z[k].tobytes(), int(cap)) == want, (name, cap)


def test_truncated_output_is_prefix():
    """lzs-compression.c:306-309: too-small outSize returns a plain prefix."""
    o = helpers.oracle()
    data = helpers.corpus(helpers.CORPUS_TEXT, 1, 3000).tobytes()
    full = o.compress(data)
    for cap in (0, 1, 2, 10, len(full) - 1):
        assert o.compress(data, cap) == full[:cap]


@pytest.mark.parametrize("kind", [helpers.CORPUS_TEXT, helpers.CORPUS_BINARY, helpers.CORPUS_RANDOM,
                                  helpers.CORPUS_PACKET])
def test_against_live_reference(kind):
    ref = helpers.reference()
    if ref is None:
        pytest.skip("oracle/_ref/liblzs_ref.so not built")
    o = helpers.oracle()
    for n in (1, 2, 3, 15, 16, 100, 1500, 2047, 2048, 2049, 4096, 20000):
        data = helpers.corpus(kind, 1, n, seed=0x5EED0000 + n).tobytes()
        want = ref.compress(data)
        assert o.compress(data) == want, (kind, n)
        assert ref.decompress(want, n + 8) == data
        assert o.decompress(want, n + 8) == data
        assert o.decompress(want, n // 2) == ref.decompress(want, n // 2)


def test_fuzz_against_live_reference():
    ref = helpers.reference()
    if ref is None:
        pytest.skip("oracle/_ref/liblzs_ref.so not built")
    o = helpers.oracle()
    rng = np.random.default_rng(2024)
    for it in range(120):
        n = int(rng.integers(0, 6000))
        alpha = int(rng.choice([2, 3, 20, 256]))
        buf = bytearray(rng.integers(0, alpha, n, dtype=np.uint8).tobytes())
        for _ in range(int(rng.integers(0, 6))):          # plant copies, some near the window edge
            if n < 2200:
                break
            d = int(rng.choice([1, 2, 7, 128, 2040, 2047, 2048, 2060]))
            p = int(rng.integers(d, n - 1))
            ln = int(rng.integers(2, 60))
            buf[p:p + ln] = buf[p - d:p - d + ln]
        data = bytes(buf[:n])
        want = ref.compress(data)
        assert o.compress(data) == want, it
        assert o.decompress(want, n + 4) == data
    for it in range(200):                                   # random bit strings through both decoders
        s = rng.integers(0, 256, int(rng.integers(0, 300)), dtype=np.uint8).tobytes()
        cap = int(rng.integers(0, 5000))
        assert o.decompress(s, cap) == ref.decompress(s, cap), it


def test_all_matches_table_agrees_with_compress():
    """The per-position table the GPU match finder is checked against is the same rule."""
    o = helpers.oracle()
    data = helpers.corpus(helpers.CORPUS_TEXT, 1, 4000).tobytes()
    ln, off = helpers.oracle_all_matches(data)
    assert ln.max() <= 12 and off.max() <= 2047
    assert (ln[off == 0] == 0).all() and (ln[ln > 0] >= 2).all()
    i = 0
    while i < len(data):
        if ln[i] >= 2:
            o_, l_ = int(off[i]), int(ln[i])
            assert data[i:i + l_] == bytes(data[i - o_ + k % o_] if k >= o_ else data[i - o_ + k] for k in range(l_))
            i += l_
        else:
            i += 1
    assert o.decompress(o.compress(data), len(data)) == data
