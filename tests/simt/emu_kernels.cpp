/*
 * emu_kernels.cpp -- builds the product's kernel sources against the CPU SIMT
 * emulator (simt.h) and exposes them to the non-GPU tests through ctypes.
 * TEST INFRASTRUCTURE: this is how kernel logic is checked in a container that
 * has no GPU; the shipped library never contains or calls any of this.
 */
#define LZS_SIMT_EMU 1
#include <vector>
#include <cstdio>
#include <cstdlib>
#include "../../lzs-compression_b200/csrc/k1_match.cuh"
#include "../../lzs-compression_b200/csrc/k23_parse_pack.cuh"
#include "../../lzs-compression_b200/csrc/k23_pieces.cuh"
#include "../../lzs-compression_b200/csrc/k4_decode.cuh"
#include "../../lzs-compression_b200/csrc/k4_pieces.cuh"

template <int G>
static void run_decode(const uint8_t *in, const uint64_t *in_off, const uint32_t *in_len, uint8_t *out,
                       const uint64_t *out_off, const uint32_t *out_cap, uint32_t *out_len, uint32_t n,
                       unsigned grid, uint8_t *status, const uint32_t *hist = nullptr)
{
    uint32_t counter = 0;
    simt::launch(dim3(grid), dim3(lzs::kDecThreads), lzs::k4_smem_bytes<G>(), [&] {
        lzs::k4_decode<G>(in, in_off, in_len, out, out_off, out_cap, out_len, n, &counter, status, nullptr, hist);
    });
}

extern "C" int emu_decode(const uint8_t *in, const uint64_t *in_off, const uint32_t *in_len, uint8_t *out,
                          const uint64_t *out_off, const uint32_t *out_cap, uint32_t *out_len, uint32_t n,
                          int lanes_per_stream, unsigned grid, uint8_t *status)
{
    switch (lanes_per_stream) {
        case 4:  run_decode<4>(in, in_off, in_len, out, out_off, out_cap, out_len, n, grid, status); break;
        case 8:  run_decode<8>(in, in_off, in_len, out, out_off, out_cap, out_len, n, grid, status); break;
        case 16: run_decode<16>(in, in_off, in_len, out, out_off, out_cap, out_len, n, grid, status); break;
        case 32: run_decode<32>(in, in_off, in_len, out, out_off, out_cap, out_len, n, grid, status); break;
        default: return -1;
    }
    return 0;
}

static const uint32_t *g_emu_hist = nullptr;   /* kept-history lengths for the next emu_match / emu_decode (test knob) */
extern "C" void emu_set_hist(const uint32_t *hist) { g_emu_hist = hist; }
static const uint32_t *g_emu_seg = nullptr;    /* packet length per stream for the next emu_match (flows as one stream) */
extern "C" void emu_set_seg(const uint32_t *seg) { g_emu_seg = seg; }

extern "C" int emu_decode_hist(const uint8_t *in, const uint64_t *in_off, const uint32_t *in_len, uint8_t *out,
                               const uint64_t *out_off, const uint32_t *out_cap, uint32_t *out_len, uint32_t n,
                               unsigned grid, const uint32_t *hist)
{
    run_decode<8>(in, in_off, in_len, out, out_off, out_cap, out_len, n, grid, nullptr, hist);
    return 0;
}

extern "C" int emu_match(const uint8_t *in, const uint64_t *in_off, const uint32_t *in_len,
                         uint16_t *matches, uint32_t n, unsigned grid)
{
    /* as the host does: the fast launch, then the safe launch (which returns at once unless the
     * fast one recorded an exchange order it does not handle).  Returns that record. */
    uint32_t ctl[4] = {0, 0, 0, 0};
    simt::launch(dim3(grid), dim3(lzs::kK1Threads), lzs::kK1SmemBytes, [&] {
        lzs::k1_match<false>(in, in_off, in_len, matches, n, ctl, g_emu_hist, g_emu_seg);
    });
    simt::launch(dim3(grid), dim3(lzs::kK1Threads), lzs::kK1SmemBytes, [&] {
        lzs::k1_match<true>(in, in_off, in_len, matches, n, ctl, g_emu_hist, g_emu_seg);
    });
    return static_cast<int>(ctl[2]);
}

/* Test knob: serve the lanes of an atomic exchange in a scrambled order (real hardware serves
 * them in ascending lane order; the product must be exact either way). */
extern "C" void emu_scramble_exchanges(int on) { simt::g_scramble_exchanges = on; }

extern "C" int emu_parse_pack(const uint8_t *in, const uint64_t *in_off, const uint32_t *in_len,
                              const uint16_t *matches, uint8_t *out, const uint64_t *out_off,
                              const uint32_t *out_cap, uint32_t *out_len, uint32_t n)
{
    unsigned grid = (n + lzs::kK2Warps - 1) / lzs::kK2Warps;
    simt::launch(dim3(grid), dim3(lzs::kK2Threads), 0, [&] {
        lzs::k23_parse_pack(in, in_off, in_len, matches, out, out_off, out_cap, out_len, n);
    });
    return 0;
}

/* The compressor for long streams (k23_pieces.cuh), the launches of compress_pieces() in
 * lzs_b200.cu one after the other.  `cap` entries of piece table; returns the table's overflow flag.
 * stats (optional, 4 words): pieces, pieces left open by spec, by fix, pieces in use. */
extern "C" int emu_compress_pieces(const uint8_t *in, const uint64_t *in_off, const uint32_t *in_len, uint64_t in_span,
                                   uint8_t *out, const uint64_t *out_off, const uint32_t *out_cap, uint32_t *out_len,
                                   uint32_t n, uint32_t piece, uint32_t cap, unsigned grid, uint32_t *stats)
{
    std::vector<uint64_t> mem((lzs::piece_table_bytes(cap) + 7) / 8, 0);
    std::vector<uint16_t> matches(in_span + 64, 0xFFFF);
    const lzs::PieceTable t = lzs::piece_table_at(mem.data(), cap);
    simt::launch(dim3(1), dim3(lzs::kPlanThreads), 0, [&] { lzs::k23p_plan_count(in_len, n, piece, t); });
    simt::launch(dim3(n), dim3(128), 0, [&] { lzs::k23p_plan_fill(in_off, in_len, out_len, n, piece, t); });
    uint32_t ctl[4] = {0, 0, 0, 0};
    simt::launch(dim3(grid), dim3(lzs::kK1Threads), lzs::kK1SmemBytes, [&] {
        lzs::k1_match<false>(in, t.off, t.len, matches.data(), cap, ctl, t.hist, nullptr, t.look);
    });
    simt::launch(dim3(grid), dim3(lzs::kK1Threads), lzs::kK1SmemBytes, [&] {
        lzs::k1_match<true>(in, t.off, t.len, matches.data(), cap, ctl, t.hist, nullptr, t.look);
    });
    const unsigned pgrid = (cap + lzs::kPieceWarps - 1) / lzs::kPieceWarps;
    const unsigned sgrid = (n + lzs::kPieceWarps - 1) / lzs::kPieceWarps;
    simt::launch(dim3(pgrid), dim3(lzs::kPieceThreads), 0, [&] { lzs::k23p_spec(in, in_off, in_len, matches.data(), piece, t); });
    simt::launch(dim3(pgrid), dim3(lzs::kPieceThreads), 0, [&] { lzs::k23p_fix(in, in_off, in_len, matches.data(), piece, t); });
    simt::launch(dim3(sgrid), dim3(lzs::kPieceThreads), 0, [&] {
        lzs::k23p_sweep(in, in_off, in_len, matches.data(), out_cap, out_len, n, piece, t);
    });
    simt::launch(dim3(pgrid), dim3(lzs::kPieceThreads), 0, [&] {
        lzs::k23p_pack(in, in_off, in_len, matches.data(), out, out_off, out_cap, t);
    });
    if (stats) {
        stats[0] = t.count[0];
        stats[1] = stats[2] = stats[3] = 0;
        for (uint32_t i = 0; i < t.count[0]; i++) {
            if (t.flags[i] & lzs::kPieceSpecOpen) stats[1]++;
            if (t.flags[i] & lzs::kPieceFixOpen) stats[2]++;
            if (t.entry[i] < t.p0[i] + t.len[i]) stats[3]++;
        }
    }
    return static_cast<int>(t.count[1]);
}

/* The decoder for long streams (k4_pieces.cuh), the launches of decompress_pieces() in lzs_b200.cu one
 * after the other, the k4_decode launch for the dirty streams included.  stats (4 words): pieces,
 * dirty streams, pieces fix left open, table overflow. */
extern "C" int emu_decode_pieces(const uint8_t *in, const uint64_t *in_off, const uint32_t *in_len, uint8_t *out,
                                 const uint64_t *out_off, const uint32_t *out_cap, uint32_t *out_len, uint32_t n,
                                 uint32_t piece, uint32_t cap, uint8_t *status, uint32_t *stats, uint32_t jump_span, uint32_t *in_used)
{
    std::vector<uint32_t> mem(lzs::dpiece_table_bytes(cap, piece) / 4 + 16, 0xCDCDCDCDu);
    std::vector<uint32_t> S(jump_span + 1, 0xCDCDCDCDu);
    const lzs::DPieceTable t = lzs::dpiece_table_at(mem.data(), cap, piece);
    const unsigned pgrid = (cap + 127) / 128, sgrid = (n + 3) / 4;
    simt::launch(dim3(1), dim3(1024), 0, [&] { lzs::k4p_plan(in_len, n, piece, t); });
    simt::launch(dim3(pgrid), dim3(128), 0, [&] { lzs::k4p_spec(in, in_off, in_len, n, piece, t); });
    simt::launch(dim3(pgrid), dim3(128), 0, [&] { lzs::k4p_fix(in, in_off, in_len, n, piece, 0u, t); });
    for (int rep = 0; rep < 3; rep++)
        simt::launch(dim3(pgrid), dim3(128), 0, [&] { lzs::k4p_fix(in, in_off, in_len, n, piece, 1u, t); });
    simt::launch(dim3(sgrid), dim3(128), 0, [&] { lzs::k4p_sweep(in, in_off, in_len, out_cap, out_len, status, n, piece, t, in_used); });
    simt::launch(dim3(pgrid), dim3(128), 0, [&] { lzs::k4p_emit(in, in_off, in_len, out, out_off, n, piece, t); });
    if (jump_span) {
        /* as decompress_pieces() does for a handful of streams: pointer doubling instead of the replay */
        uint32_t *flags = t.count + 8;
        simt::launch(dim3(3), dim3(256), 0, [&] { lzs::k4j_init(S.data(), jump_span, flags); });
        simt::launch(dim3((cap + 3) / 4), dim3(128), 0, [&] { lzs::k4j_fill(out_off, 0, n, S.data(), t); });
        for (int r = 0; r < lzs::kJumpRounds; r++)
            simt::launch(dim3(3), dim3(256), 0, [&] { lzs::k4j_jump(S.data(), jump_span, flags, static_cast<uint32_t>(r)); });
        simt::launch(dim3(3), dim3(256), 0, [&] { lzs::k4j_gather(out, 0, S.data(), jump_span); });
    } else {
        simt::launch(dim3(n), dim3(lzs::kDCopyThreads), 0, [&] { lzs::k4p_copy(out, out_off, out_len, n, t); });
    }
    simt::launch(dim3((n + 127) / 128), dim3(128), 0, [&] { lzs::k4p_dirty_list(n, t); });
    simt::launch(dim3(2), dim3(lzs::kDecThreads), lzs::k4_smem_bytes<8>(), [&] {
        lzs::k4_decode<8>(in, in_off, in_len, out, out_off, out_cap, out_len, n, &t.count[3], status, t.dirty_list, nullptr,
                          &t.count[2]);
    });
    if (stats) {
        stats[0] = t.count[0];
        stats[1] = t.count[2];
        stats[2] = 0;
        for (uint32_t i = 0; i < t.count[0]; i++) stats[2] += t.fix_status[i] == lzs::kDStOpen;
        stats[3] = t.count[1];
        /* batches of 32 pieces the sweep can take at once (its chain condition), of all batches */
        if (getenv("EMU_DBG_SWEEP")) {
            unsigned all = 0, fast = 0;
            for (uint32_t s2 = 0; s2 < n; s2++) {
                for (uint32_t k0 = t.first[s2]; k0 < t.first[s2 + 1]; k0 += 32) {
                    bool ok = k0 != t.first[s2];
                    for (uint32_t k = k0; k < k0 + 32 && k < t.first[s2 + 1] && ok; k++)
                        ok = t.fix_status[k] == lzs::kDStOk && t.fix_entry[k] == t.fix_exit[k - 1];
                    all++;
                    fast += ok;
                }
            }
            fprintf(stderr, "sweep batches %u chained %u\n", all, fast);
        }
    }
    return 0;
}

#include "../../lzs-compression_b200/csrc/incremental.cuh"

/* One incremental call on the emulator.  `state` is the caller's private state block
 * (host memory stands in for device memory here). */
extern "C" int emu_inc_call(int decompress, void *state, const uint8_t *in, uint32_t in_len, uint8_t *out,
                            uint32_t out_cap, int add_end_marker, uint32_t *in_used, uint32_t *out_used,
                            uint32_t *status)
{
    lzs::IncJob job;
    job.state = state;
    job.in = in;
    job.out = out;
    job.in_len = in_len;
    job.out_cap = out_cap;
    job.in_used = job.out_used = job.status = 0;
    job.add_end_marker = add_end_marker ? 1u : 0u;
    simt::launch(dim3(1), dim3(128), 0, [&] {
        if (decompress) lzs::kinc_decompress(&job, 1);
        else            lzs::kinc_compress(&job, 1);
    });
    *in_used = job.in_used;
    *out_used = job.out_used;
    *status = job.status;
    return 0;
}

extern "C" void emu_inc_init(int decompress, void *state)
{
    if (decompress) {
        lzs::IncDecompressState *s = static_cast<lzs::IncDecompressState *>(state);
        memset(s, 0, offsetof(lzs::IncDecompressState, ring));
        s->state = lzs::kDTokenType;
    } else {
        memset(state, 0, offsetof(lzs::IncCompressState, ring));
    }
}
