"""Long streams cut into pieces (csrc/k23_pieces.cuh): the kernels' logic on the CPU SIMT emulator
against the oracle.  Whatever the piece size and wherever the pieces' borders fall -- inside
literal runs, inside short matches, inside matches thousands of bytes long -- the stitched stream
must be the one stream the reference produces (c/src/liblzs/lzs-compression.c:249-467)."""
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
