"""Long streams cut into pieces (csrc/k23_pieces.cuh): the kernels' logic on the CPU SIMT emulator
against the oracle.  Whatever the piece size and wherever the pieces' borders fall -- inside
literal runs, inside short matches, inside matches thousands of bytes long -- the stitched stream
must be the one stream the reference produces (c/src/liblzs/lzs-compression.c:249-467)."""
import os

import numpy as np
import pytest

import emu
import helpers


def _check(data, piece, **kw):
    o = helpers.oracle()
    got, stats = emu.compress_pieces(data, piece, **kw)
    assert got is not None, "piece table overflow"
    for i, (g, d) in enumerate(zip(got, data)):
        assert g == o.compress(d), "stream %d (%d bytes), piece %d" % (i, len(d), piece)
    return stats


@pytest.mark.parametrize("piece", [64, 100, 777])
def test_pieces_edge_cases_and_corpora(piece):
    cases = helpers.edge_case_inputs()
    data = [cases[k] for k in cases if len(cases[k]) <= 6000]
    data += [helpers.corpus(kind, 1, 2500, first_index=3).tobytes()
             for kind in (helpers.CORPUS_MIXED, helpers.CORPUS_PACKET)]
    stats = _check(data, piece)
    assert stats[0] >= sum(max(1, -(-len(d) // piece)) for d in data)


def test_pieces_long_matches_across_many_pieces():
    """Runs and periodic data: one match covers many pieces (those hold no token at all), the pieces
    that look further than four pieces ahead are left open and the sweep measures the match itself."""
    rng = np.random.default_rng(11)
    noise = lambda n: rng.integers(0, 256, n, dtype=np.uint8).tobytes()
    data = [
        noise(300) + b"\0" * 9000 + noise(200),                     # a run through ~70 pieces of 128
        noise(100) + (b"abcdefg" * 700) + noise(50) + (b"xy" * 900),
        b"\0" * 4096,                                                # the run ends with the stream
        noise(64) + b"q" * 127 + noise(1) + b"q" * 129 + noise(7),  # matches ending at / next to piece borders
        (noise(40) + b"\0" * 88) * 30,                               # period 128 = the piece size
    ]
    stats = _check(data, 128)
    assert stats[1] > 0 and stats[2] > 0            # some pieces were left open ...
    assert stats[3] < stats[0]                      # ... and some hold no token


def test_pieces_text_records_random_4k():
    for kind in (helpers.CORPUS_TEXT, helpers.CORPUS_BINARY, helpers.CORPUS_RANDOM):
        data = [helpers.corpus(kind, 1, 12000, first_index=5).tobytes()]
        stats = _check(data, 4096)
        assert stats[0] == 3


@pytest.mark.parametrize("lead,out_lead", [(1, 0), (3, 5)])
def test_pieces_unaligned_and_truncated(lead, out_lead):
    """Streams at odd addresses, output slots at odd addresses, capacities smaller than the stream:
    the result is the prefix that fits (lzs-compression.c:306-309), nothing behind it is written."""
    o = helpers.oracle()
    data = [helpers.corpus(helpers.CORPUS_MIXED, 1, 3000 + 7 * i, first_index=i).tobytes() for i in range(4)]
    full = [o.compress(d) for d in data]
    caps = [len(full[0]) - 1, 0, len(full[2]) // 2, len(full[3]) + 9]
    got, _ = emu.compress_pieces(data, 200, caps=caps, align=4, lead=lead, out_lead=out_lead)
    for g, f, c in zip(got, full, caps):
        assert g == f[:c]


def test_pieces_table_too_small_produces_nothing():
    data = [helpers.corpus(helpers.CORPUS_MIXED, 1, 3000, first_index=1).tobytes()]
    got, _ = emu.compress_pieces(data, 100, cap_entries=10)
    assert got is None


# ---------------------------------------------------------------- the decoder for long streams (csrc/k4_pieces.cuh)

JUMP = [False]          # the same checks with the copies resolved by pointer doubling (the host calls' way for a handful of streams)


def _decode_check(streams, caps, piece, **kw):
    kw.setdefault("jump", JUMP[0])
    """Outputs and stop reasons must be the oracle's (= the reference's, tests/test_oracle.py) whether a
    stream is finished by the piece passes or handed to k4_decode as dirty."""
    o = helpers.oracle()
    got, stats, status = emu.decode_pieces(streams, caps, piece, with_status=True, **kw)
    for i, (s, c, g) in enumerate(zip(streams, caps, got)):
        want, why = o.decompress_status(s, c)
        assert g == want, "stream %d, piece %d" % (i, piece)
        assert int(status[i]) == why, "status of stream %d, piece %d: %d, want %d" % (i, piece, int(status[i]), why)
    return stats


@pytest.mark.parametrize("piece", [16, 50, 333])
def test_decode_pieces_clean_streams(piece):
    o = helpers.oracle()
    cases = helpers.edge_case_inputs()
    data = [cases[k] for k in cases if len(cases[k]) <= 6000]
    data += [helpers.corpus(kind, 1, 4000, first_index=4).tobytes()
             for kind in (helpers.CORPUS_TEXT, helpers.CORPUS_BINARY, helpers.CORPUS_RANDOM, helpers.CORPUS_PACKET)]
    comp = [o.compress(d) for d in data]
    for slack in (0, 1, 100):                   # a full output stops the decoder before the end marker: another status
        stats = _decode_check(comp, [len(d) + slack for d in data], piece)
        assert stats[1] == 0 and stats[3] == 0  # nothing left to k4_decode


def test_decode_pieces_long_tokens_and_unaligned():
    """Matches whose continuation nibbles run through many pieces (runs), streams and outputs at odd
    addresses."""
    o = helpers.oracle()
    rng = np.random.default_rng(3)
    noise = lambda n: rng.integers(0, 256, n, dtype=np.uint8).tobytes()
    data = [noise(50) + b"\0" * 60000 + noise(30), b"ab" * 20000, noise(20) + b"z" * 2100 + noise(5) + b"y" * 2046 + b"x" * 2048]
    comp = [o.compress(d) for d in data]
    for lead, out_lead in ((0, 0), (1, 3), (3, 5)):
        stats = _decode_check(comp, [len(d) + 3 for d in data], 16, align=4, lead=lead, out_lead=out_lead)
        assert stats[1] == 0


def test_decode_pieces_damaged_streams_and_short_outputs():
    """Everything that is not a clean stream is k4_decode's: the committed outputs of the unmodified
    reference on damaged streams and short capacities, random bit strings, a table that is too small."""
    z = np.load(os.path.join(helpers.GOLDEN_DIR, "ref_cases.npz"))
    keys = [k for k in z.files if k.startswith("dec_in__")]
    streams = [z[k].tobytes() for k in keys]
    caps = [int(k.split("__")[2]) for k in keys]
    want = [z["dec_out__" + k[len("dec_in__"):]].tobytes() for k in keys]
    got, stats = emu.decode_pieces(streams, caps, 24)
    assert got == want
    assert stats[1] > 0
    rng = np.random.default_rng(5)
    streams = [rng.integers(0, 256, int(rng.integers(0, 400)), dtype=np.uint8).tobytes() for _ in range(60)]
    caps = [int(rng.integers(0, 3000)) for _ in streams]
    _decode_check(streams, caps, 32)
    o = helpers.oracle()
    data = [helpers.corpus(helpers.CORPUS_MIXED, 1, 3000, first_index=i).tobytes() for i in range(3)]
    stats = _decode_check([o.compress(d) for d in data], [len(d) for d in data], 64, cap_entries=5)
    assert stats[3] == 1 and stats[1] == 3


# ---------------------------------------------------------------- structured fuzz, both directions

def _structured(rng, n):
    """Runs, short periods, copies of earlier bytes, small alphabets, noise, text: what makes piece borders
    fall into every kind of token."""
    out = bytearray()
    while len(out) < n:
        r = rng.integers(0, 6)
        if r == 0:
            out += rng.integers(0, 256, int(rng.integers(1, 60)), dtype=np.uint8).tobytes()
        elif r == 1:
            out += bytes([int(rng.integers(0, 256))]) * int(rng.integers(1, 4000))
        elif r == 2:
            out += rng.integers(0, 4, int(rng.integers(1, 9)), dtype=np.uint8).tobytes() * int(rng.integers(1, 300))
        elif r == 3 and len(out) > 10:
            a = int(rng.integers(0, len(out)))
            out += out[a:a + int(rng.integers(2, 40))]
        elif r == 4:
            out += rng.integers(0, 3, int(rng.integers(1, 100)), dtype=np.uint8).tobytes()
        else:
            out += b"the quick brown fox "[:int(rng.integers(1, 20))]
    return bytes(out[:n])


def test_pieces_structured_fuzz_compressor():
    rng = np.random.default_rng(77)
    for _ in range(8):
        data = [_structured(rng, int(rng.integers(1, 9000))) for _ in range(4)]
        _check(data, int(rng.integers(64, 400)), grid=2)


def test_pieces_structured_fuzz_decoder():
    """Clean, truncated, damaged and padded streams with exact, generous and short capacities, at odd
    addresses: bytes and stop reasons are the oracle's for every piece size."""
    o = helpers.oracle()
    rng = np.random.default_rng(78)
    for _ in range(40):
        data = [_structured(rng, int(rng.integers(1, 20000))) for _ in range(6)]
        streams, caps = [], []
        for d in data:
            c = o.compress(d)
            k = int(rng.integers(0, 6))
            if k == 0:
                streams.append(c); caps.append(len(d))
            elif k == 1:
                streams.append(c); caps.append(len(d) + int(rng.integers(1, 50)))
            elif k == 2:
                streams.append(c[:int(rng.integers(0, len(c) + 1))]); caps.append(len(d) + 5)
            elif k == 3:
                streams.append(c); caps.append(int(rng.integers(0, len(d) + 1)))
            elif k == 4:
                b = bytearray(c)
                for _ in range(int(rng.integers(1, 4))):
                    b[int(rng.integers(0, len(b)))] ^= 1 << int(rng.integers(0, 8))
                streams.append(bytes(b)); caps.append(len(d) + int(rng.integers(0, 3000)))
            else:
                streams.append(c + rng.integers(0, 256, 7, dtype=np.uint8).tobytes()); caps.append(len(d) + 9)
        _decode_check(streams, caps, int(rng.integers(16, 300)), lead=int(rng.integers(0, 4)), align=4,
                      out_lead=int(rng.integers(0, 8)))


def test_decode_pieces_pointer_doubling_instead_of_replay():
    """k4j_*: every byte finds the literal it is a copy of by pointer doubling.  The decoder's tests again,
    with that in place of the replay."""
    JUMP[0] = True
    try:
        test_decode_pieces_clean_streams(50)
        test_decode_pieces_long_tokens_and_unaligned()
        test_decode_pieces_damaged_streams_and_short_outputs()
        test_pieces_structured_fuzz_decoder()
    finally:
        JUMP[0] = False


def test_decode_pieces_report_where_the_end_marker_is():
    """in_used: bytes of a stream up to and including its end marker (what walks a file of several streams laid
    end to end); unknown (all ones) for streams that are not clean."""
    o = helpers.oracle()
    data = [helpers.corpus(helpers.CORPUS_MIXED, 1, 5000 + 13 * i, first_index=i).tobytes() for i in range(4)]
    comp = [o.compress(d) for d in data]
    streams = [comp[0], comp[1] + b"\x12\x34\x56" * 40, comp[2] + comp[3], comp[3][:len(comp[3]) // 2]]
    caps = [len(data[0]), len(data[1]) + 7, len(data[2]) + 100, len(data[3])]
    used = np.zeros(4, dtype=np.uint32)
    got, _ = emu.decode_pieces(streams, caps, 40, in_used=used, lead=1, align=4)
    assert got[:3] == data[:3]
    assert list(used[:3]) == [len(comp[0]), len(comp[1]), len(comp[2])]
    assert used[3] == 0xFFFFFFFF
