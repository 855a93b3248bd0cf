"""Kernel LOGIC checks without a GPU: the product's .cuh kernel sources are compiled
against tests/simt (a CPU model of warps, collectives and shared memory) and compared
with the oracle.  This is a development aid for a container with no GPU -- the parity
tests proper are tests/test_gpu_*.py (-m gpu), which run the real sm_100a kernels."""
import os

import numpy as np
import pytest

import emu
import helpers


def _want_matches(data):
    ln, off = helpers.oracle_all_matches(data)
    return (ln.astype(np.uint16) << 11) | off


@pytest.fixture(scope="module")
def cases():
    return helpers.edge_case_inputs()


@pytest.mark.parametrize("lanes", [4, 8, 16, 32])
def test_decode_edge_cases(cases, lanes):
    o = helpers.oracle()
    names = list(cases)
    streams = [o.compress(cases[k]) for k in names]
    got = emu.decode(streams, [len(cases[k]) + 7 for k in names], lanes=lanes)
    for k, g in zip(names, got):
        assert g == cases[k], k


def test_decode_golden_vector():
    comp = open(os.path.join(helpers.GOLDEN_DIR, "golden1_compressed.bin"), "rb").read()
    plain = open(os.path.join(helpers.GOLDEN_DIR, "golden1_plain.bin"), "rb").read()
    assert emu.decode([comp], [len(plain) + 520]) == [plain]


@pytest.mark.parametrize("lead,out_lead", [(0, 0), (1, 0), (3, 5), (2, 16)])
def test_decode_damaged_and_unaligned(lead, out_lead):
    """Malformed streams and capacity limits must give exactly the reference's bytes
    (committed outputs of the unmodified reference, tests/golden/ref_cases.npz)."""
    z = np.load(os.path.join(helpers.GOLDEN_DIR, "ref_cases.npz"))
    keys = [k for k in z.files if k.startswith("dec_in__")]
    streams = [z[k].tobytes() for k in keys]
    caps = [int(k.split("__")[2]) for k in keys]
    want = [z["dec_out__" + k[len("dec_in__"):]].tobytes() for k in keys]
    got = emu.decode(streams, caps, lanes=8, align=4, lead=lead, out_lead=out_lead)
    for k, g, w in zip(keys, got, want):
        assert g == w, k


def test_decode_random_bitstrings_match_oracle():
    rng = np.random.default_rng(5)
    o = helpers.oracle()
    streams = [rng.integers(0, 256, int(rng.integers(0, 400)), dtype=np.uint8).tobytes() for _ in range(60)]
    caps = [int(rng.integers(0, 3000)) for _ in streams]
    got = emu.decode(streams, caps, lanes=8, grid=3)
    for s, c, g in zip(streams, caps, got):
        assert g == o.decompress(s, c)


@pytest.mark.parametrize("lanes", [4, 8, 32])
def test_decode_status_words(cases, lanes):
    """one status byte per stream saying why decoding stopped (end marker /
    output full with input left / input ended first), against the oracle's statement of the
    reference loop's exits -- on valid streams, capacity limits, truncations and random bits."""
    o = helpers.oracle()
    rng = np.random.default_rng(17)
    streams, caps = [], []
    for name, d in cases.items():
        c = o.compress(d)
        streams += [c, c, c[:len(c) // 2], c[:max(0, len(c) - 1)], c + b"\x55\xAA"]
        caps += [len(d) + 8, max(0, len(d) - 3), len(d) + 8, len(d), len(d)]
    streams += [rng.integers(0, 256, int(rng.integers(0, 300)), dtype=np.uint8).tobytes() for _ in range(40)]
    caps += [int(rng.integers(0, 2000)) for _ in range(40)]
    got, status = emu.decode(streams, caps, lanes=lanes, grid=3, with_status=True)
    seen = set()
    for s, c, g, st in zip(streams, caps, got, status):
        want, why = o.decompress_status(s, c)
        assert g == want and st == why, (len(s), c, st, why)
        seen.add(why)
    assert seen == {0x01, 0x04, 0x08}


def test_match_finder_edge_cases(cases):
    names = list(cases)
    got, _ = emu.match([cases[k] for k in names])
    for k, m in zip(names, got):
        assert (m == _want_matches(cases[k])).all(), k


@pytest.mark.parametrize("kind", [helpers.CORPUS_TEXT, helpers.CORPUS_BINARY, helpers.CORPUS_RANDOM,
                                  helpers.CORPUS_PACKET])
def test_compress_corpora(kind):
    o = helpers.oracle()
    sizes = [1500, 4096, 9000]
    data = [helpers.corpus(kind, 1, n, seed=0x5EED0000 + n).tobytes() for n in sizes]
    got = emu.compress(data, lead=kind & 3, align=4)
    for d, g in zip(data, got):
        assert g == o.compress(d)


def test_compress_edge_cases_and_truncation(cases):
    o = helpers.oracle()
    names = list(cases)
    data = [cases[k] for k in names]
    full = [o.compress(d) for d in data]
    assert emu.compress(data) == full
    caps = [max(0, len(f) - 1 - (i % 7)) for i, f in enumerate(full)]
    got = emu.compress(data, caps=caps, out_lead=1)
    for f, c, g in zip(full, caps, got):
        assert g == f[:c]


def test_compress_streams_share_tables_safely():
    """Many short streams through ONE thread block: stale table entries from earlier
    streams must never leak into later ones (epoch rule in k1_match.cuh)."""
    o = helpers.oracle()
    rng = np.random.default_rng(11)
    base = helpers.corpus(helpers.CORPUS_TEXT, 1, 600).tobytes()
    data = []
    for i in range(40):
        n = int(rng.integers(0, 600))
        data.append(base[:n] if i % 2 else helpers.corpus(helpers.CORPUS_PACKET, 1, n, first_index=i).tobytes())
    assert emu.compress(data, grid=1) == [o.compress(d) for d in data]


def test_match_finder_exact_for_any_exchange_order(cases):
    """K1's fast insert relies on sm_100a serving the lanes of a shared-memory exchange in ascending
    lane order.  With the emulator serving them in a scrambled order the fast launch must notice,
    and the safe launch that follows must produce the exact records anyway."""
    data = [helpers.corpus(helpers.CORPUS_BINARY, 1, 3000, first_index=2).tobytes(),
            helpers.corpus(helpers.CORPUS_TEXT, 1, 2500, first_index=3).tobytes(), b"\0" * 700, cases["run_a_5000"]
            if "run_a_5000" in cases else b"a" * 5000]
    want = [_want_matches(d) for d in data]
    got, _ = emu.match(data)
    assert emu.last_match_disorder == 0, "in-order exchanges must not trigger the safe launch"
    for w, g in zip(want, got):
        assert (w == g).all()
    emu.scramble_exchanges(True)
    try:
        got, _ = emu.match(data)
        assert emu.last_match_disorder == 1, "the scrambled order went unnoticed"
    finally:
        emu.scramble_exchanges(False)
    for w, g in zip(want, got):
        assert (w == g).all()


def test_compress_64k_chunk_crosses_16bit_positions():
    o = helpers.oracle()
    d = helpers.corpus(helpers.CORPUS_MIXED, 1, 70000, first_index=1).tobytes()
    e = helpers.corpus(helpers.CORPUS_MIXED, 1, 3000, first_index=0).tobytes()
    assert emu.compress([d, e]) == [o.compress(d), o.compress(e)]
