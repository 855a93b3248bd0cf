/*
 * k23_pieces.cuh -- K2+K3 for LONG streams: the greedy parse and the bit packer of
 * k23_parse_pack.cuh (reference: c/src/liblzs/lzs-compression.c:301-466) made parallel INSIDE a
 * stream, so that a batch of a few large streams -- or one lzs_compress call on one large buffer --
 * uses the whole GPU instead of one warp per stream.
 *
 * A stream is cut into pieces of P bytes.  K1 takes the pieces as streams of their own (2047 bytes of
 * history in front, 11 bytes of look-ahead behind: k1_match.cuh `hist_len` / `look_len`) and leaves
 * the same per-position records as for the whole stream.  The parse is the orbit of
 * next(i) = i + advance(i) from position 0, and which positions of a piece are on it depends on where
 * the last token of the piece before ends.  But orbits from different starts run into each other
 * after a few tokens (any literal, or any position both reach, joins them), so:
 *
 *   plan    pieces per stream, exclusive scan, piece table                       (2 small kernels)
 *   spec    one warp per piece parses it from its first position -- a guess -- and
 *           records the bits it emits and the first token start behind the piece (its exit)
 *   fix     one warp per piece takes the exit of the piece before as its entry and walks both
 *           orbits token by token until they meet: bits(entry) = bits(guess) - bits the guess spent
 *           before the meeting point + bits the true orbit spent; exit as guessed
 *   sweep   one warp per stream goes through its pieces in order: true entry e of a piece is the
 *           exit of the piece before; if it is what `fix` assumed (almost always) its numbers
 *           stand, otherwise the piece is parsed again from e.  A running sum gives every piece
 *           the bit offset of its first token, and the stream its length.
 *   pack    one warp per piece parses from its true entry and writes its tokens at its bit offset.
 *           A byte shared by two pieces is written by the first of them, which takes the missing
 *           low bits from the head of the next token; the second starts at its first whole byte.
 *
 * Long matches (K1 length 12 = "12 or more") need their bytes compared to know where they end.  spec
 * and fix stop comparing 4 pieces behind their own (a run of zeros through a whole file would
 * otherwise be compared once per piece) and mark the piece unresolved; the sweep then measures such a
 * match once.  pack never compares beyond its piece: a match that reaches the piece's end is its
 * last token and ends at the piece's exit, which the sweep has stored.
 *
 * Results are byte for byte those of k23_parse_pack on the whole stream (tests/test_pieces.py,
 * tests/test_gpu_parity.py); extra HBM traffic: the 2 B/position records are read twice.
 */
#ifndef LZS_B200_K23_PIECES_CUH
#define LZS_B200_K23_PIECES_CUH

#include "k23_parse_pack.cuh"

namespace lzs {

struct PieceTable {
    uint64_t *off;        /* piece's first byte, offset in `in` (K1: in_off)                      */
    uint64_t *bitoff;     /* sweep: bit offset of the piece's first token in the stream's output    */
    uint32_t *len;        /* own bytes (K1: in_len); 0 for unused entries and empty streams         */
    uint32_t *hist;       /* bytes of the stream in front of the piece that offsets may reach       */
    uint32_t *look;       /* bytes of the stream behind the piece that its matches may extend into  */
    uint32_t *sid;        /* stream                                                                 */
    uint32_t *p0;         /* first position of the piece inside its stream                          */
    uint32_t *spec_exit, *spec_bits;   /* parse from p0                                             */
    uint32_t *fix_exit, *fix_bits;     /* parse from the exit of the piece before                   */
    uint32_t *flags;      /* kPieceSpecOpen / kPieceFixOpen                                         */
    uint32_t *entry, *exit;            /* sweep: the true ones                                      */
    uint32_t *run_end;    /* spec: 0, or the piece starts inside a long match (K1 length 12 at p0) whose bytes
                             agree with those `offset` before them from p0 up to here                */
    uint32_t *first;      /* [n_streams + 1] first piece of every stream                            */
    uint32_t *count;      /* [0] pieces in use, [1] != 0: the table was too small, nothing is produced */
    uint32_t  cap;        /* entries                                                                */
};
constexpr uint32_t kPieceSpecOpen = 1u, kPieceFixOpen = 2u;
constexpr uint32_t kPieceLookPieces = 4;      /* spec / fix compare at most this many pieces ahead  */
constexpr uint32_t kPieceWalkMax = 2048;      /* fix gives up after this many tokens (sweep parses the piece instead) */
constexpr size_t   kPieceEntryBytes = 2 * 8 + 14 * 4;

__host__ __device__ inline size_t piece_table_bytes(uint32_t cap) { return static_cast<size_t>(cap) * kPieceEntryBytes + 256; }

/* carve the table out of `base` (256-byte aligned); the 64-bit arrays come first */
__host__ __device__ inline PieceTable piece_table_at(void *base, uint32_t cap)
{
    PieceTable t;
    uint64_t  *q = static_cast<uint64_t *>(base);
    t.count = reinterpret_cast<uint32_t *>(q);
    q += 32;
    t.off = q;          q += cap;
    t.bitoff = q;       q += cap;
    uint32_t *w = reinterpret_cast<uint32_t *>(q);
    t.len = w;          w += cap;
    t.hist = w;         w += cap;
    t.look = w;         w += cap;
    t.sid = w;          w += cap;
    t.p0 = w;           w += cap;
    t.spec_exit = w;    w += cap;
    t.spec_bits = w;    w += cap;
    t.fix_exit = w;     w += cap;
    t.fix_bits = w;     w += cap;
    t.flags = w;        w += cap;
    t.entry = w;        w += cap;
    t.exit = w;         w += cap;
    t.run_end = w;      w += cap;
    t.first = w;
    t.cap = cap;
    return t;
}

/* ---------------------------------------------------------------- plan */

constexpr int kPlanThreads = 1024;

/* One block: pieces per stream (at least one, so that an empty stream still gets its end marker),
 * exclusive scan into first[]. */
__global__ void __launch_bounds__(kPlanThreads)
k23p_plan_count(const uint32_t *__restrict__ in_len, uint32_t n_streams, uint32_t piece, PieceTable t)
{
    __shared__ uint32_t s_warp[kPlanThreads / 32];
    __shared__ uint32_t s_base;
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (uint32_t s0 = 0; s0 < n_streams; s0 += kPlanThreads) {
        const uint32_t s = s0 + threadIdx.x;
        const uint32_t n = s < n_streams ? in_len[s] : 0u;
        const uint32_t np = s < n_streams ? (n ? (n - 1u) / piece + 1u : 1u) : 0u;
        uint32_t       incl = np;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(LZS_FULL_MASK, incl, static_cast<unsigned>(d));
            if (lane >= static_cast<uint32_t>(d)) incl += v;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        uint32_t before = s_base;
        for (uint32_t w = 0; w < warp; w++) before += s_warp[w];
        if (s < n_streams) t.first[s] = before + incl - np;
        __syncthreads();
        if (threadIdx.x == kPlanThreads - 1) s_base = before + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        t.first[n_streams] = s_base;
        t.count[0] = s_base <= t.cap ? s_base : 0u;
        t.count[1] = s_base <= t.cap ? 0u : 1u;
    }
}

/* One block per stream fills its pieces; with an overflowing table every stream gets length 0. */
__global__ void k23p_plan_fill(const uint64_t *__restrict__ in_off, const uint32_t *__restrict__ in_len,
                               uint32_t *__restrict__ out_len, uint32_t n_streams, uint32_t piece, PieceTable t)
{
    const uint32_t s = blockIdx.x;
    if (s >= n_streams) return;
    if (t.count[1]) {
        if (threadIdx.x == 0) out_len[s] = 0;
        return;
    }
    const uint32_t n = in_len[s];
    const uint32_t first = t.first[s], np = t.first[s + 1] - first;
    for (uint32_t k = threadIdx.x; k < np; k += blockDim.x) {
        const uint32_t idx = first + k;
        const uint32_t p0 = k * piece;
        const uint32_t len = umin32(piece, n - p0);
        t.off[idx] = in_off[s] + p0;
        t.len[idx] = len;
        t.hist[idx] = umin32(p0, kWindow);
        t.look[idx] = umin32(kSearchMax - 1u, n - p0 - len);
        t.sid[idx] = s;
        t.p0[idx] = p0;
    }
}

/* ---------------------------------------------------------------- tokens */

/* Bit pattern of the token at a position: value and width; for a long match (K1 length 12) the
 * header and the first 1111 only -- its nibbles follow once its length is known. */
__device__ __forceinline__ void k23_token(uint32_t len, uint32_t off, uint32_t byte, uint32_t &val, uint32_t &nb)
{
    if (len < kMinLen) {
        val = byte;
        nb = 9u;
        return;
    }
    if (off <= kShortOffMax) { val = 0x180u | off;  nb = 9u; }
    else                     { val = 0x1000u | off; nb = 13u; }
    if (len <= 4u) {
        val = (val << 2) | (len - 2u);
        nb += 2u;
    } else if (len < kMaxShortLen) {
        val = (val << 4) | (0xCu + len - 5u);
        nb += 4u;
    } else if (len < kSearchMax) {
        val = (val << 8) | 0xF0u | (len - kMaxShortLen);
        nb += 8u;
    } else {
        val = (val << 4) | 0xFu;
        nb += 4u;
    }
}

/* Whole warp: length of the match at p with offset loff, known to be >= 12, compared as far as
 * `limit` (<= n).  Most long matches end within the first 32 bytes; the ones that do not are
 * compared 256 bytes per round, the loads of a round in flight together (a run through a whole
 * file is measured by ONE warp, in the sweep: ~1 GB/s). */
__device__ __forceinline__ uint32_t k23_long_length(const uint8_t *src, uint32_t p, uint32_t loff, uint32_t limit,
                                                    uint32_t L = kSearchMax)
{
    const uint32_t lane = lane_id();
    const uint8_t *from = src - loff;
    {
        const uint32_t idx = p + L + lane;
        const uint32_t ball = __ballot_sync(LZS_FULL_MASK, (idx < limit) && (src[idx] == from[idx]));
        if (ball != LZS_FULL_MASK) return L + static_cast<uint32_t>(__ffs(static_cast<int>(~ball)) - 1);
        L += 32u;
    }
    for (;;) {
        uint32_t ball[8];
#pragma unroll
        for (int t = 0; t < 8; t++) {
            const uint32_t idx = p + L + lane + 32u * static_cast<uint32_t>(t);
            ball[t] = __ballot_sync(LZS_FULL_MASK, (idx < limit) && (src[idx] == from[idx]));
        }
#pragma unroll
        for (int t = 0; t < 8; t++) {
            if (ball[t] != LZS_FULL_MASK) return L + static_cast<uint32_t>(__ffs(static_cast<int>(~ball[t])) - 1);
            L += 32u;
        }
    }
}

/* The same without a limit, for the sweep: a match that runs through many pieces (zeros, a period) is
 * not compared byte by byte by this one warp.  Every piece that starts inside such a match has
 * measured -- in spec, in parallel -- how far the bytes from ITS start agree at ITS offset
 * (run_end); where that offset is ours, the agreement is the same fact, and the length jumps from
 * piece to piece.  `first` is the stream's first entry in the table. */
__device__ __forceinline__ uint32_t k23_long_length_hops(const uint8_t *src, const match_t *m, uint32_t n, uint32_t p,
                                                         uint32_t loff, uint32_t piece, const uint32_t *run_end,
                                                         uint32_t first)
{
    uint32_t q = p + kSearchMax;                       /* bytes in [p, q) are known to agree */
    while (q < n) {
        const uint32_t j = q / piece;
        const uint32_t mv = m[j * piece];
        if ((mv >> kMatchOffBits) >= kSearchMax && (mv & ((1u << kMatchOffBits) - 1u)) == loff) {
            const uint32_t re = run_end[first + j];
            if (re > q) {
                q = re;
                continue;
            }
        }
        const uint64_t far = static_cast<uint64_t>(j + 1u) * piece;
        const uint32_t bound = far < n ? static_cast<uint32_t>(far) : n;
        const uint32_t q2 = p + k23_long_length(src, p, loff, bound, q - p);
        if (q2 < bound) return q2 - p;
        q = bound;
    }
    return umin32(q, n) - p;
}

/* bits of the nibbles that follow the first 1111 of a match of length L >= 8 */
__device__ __forceinline__ uint32_t k23_nibble_bits(uint32_t L)
{
    return 4u * ((L - kMaxShortLen) / kMaxExtLen) + 4u;
}

/* ---------------------------------------------------------------- output stage of a piece */

struct PieceStage {
    uint32_t *buf;        /* 64 words of shared memory, zero where no bit was put yet            */
    uint8_t  *dst;
    uint64_t  lo, hi;     /* bytes [lo, hi) of dst are this piece's to write                       */
    uint32_t  cur;        /* bits pending in buf (counted from the start of word wdone)            */
    uint32_t  wdone;      /* index in dst of the word buf[0] stands for                            */
    bool      aligned;
};

__device__ __forceinline__ void pstage_store_word(const PieceStage &s, uint32_t widx, uint32_t w)
{
    const uint64_t bi = static_cast<uint64_t>(widx) * 4u;
    if (s.aligned && bi >= s.lo && bi + 4u <= s.hi) {
        *reinterpret_cast<uint32_t *>(s.dst + bi) = bswap32(w);
    } else {
        for (uint32_t b = 0; b < 4u; b++)
            if (bi + b >= s.lo && bi + b < s.hi) s.dst[bi + b] = static_cast<uint8_t>(w >> (24u - 8u * b));
    }
}

__device__ __forceinline__ void pstage_flush_if_full(PieceStage &s)
{
    if (s.cur >= 1024u) {
        const uint32_t lane = lane_id();
        const uint32_t w = s.buf[lane];
        const uint32_t hi = s.buf[lane + 32];
        pstage_store_word(s, s.wdone + lane, w);
        __syncwarp();
        s.buf[lane] = hi;
        s.buf[lane + 32] = 0;
        __syncwarp();
        s.wdone += 32;
        s.cur -= 1024u;
    }
}

__device__ __forceinline__ void pstage_emit_uniform(PieceStage &s, uint32_t val, uint32_t nb)
{
    if (lane_id() == 0) stage_put(s.buf, s.cur, val, nb);
    s.cur += nb;
    __syncwarp();
    pstage_flush_if_full(s);
}

/* ---------------------------------------------------------------- the parse of one piece */

/* Whole warp: tokens from `pos` (a token start) while they start before p1.  Counts their bits and,
 * with kPack, writes them to the stage.  A long match is compared up to `limit`; if it is still
 * running there and the stream is not over, its end is `known_exit` (kPack: the sweep stored it) or
 * unknown (counting: returns false, exit and bits are then meaningless). */
struct PieceHops {          /* sweep only: lets a long match jump over the pieces it runs through */
    const uint32_t *run_end;
    uint32_t        first, piece;
};

template <bool kPack>
__device__ __forceinline__ bool k23_piece(const uint8_t *src, const match_t *m, uint32_t n, uint32_t pos, uint32_t p1,
                                          uint32_t limit, uint32_t known_exit, PieceStage &s, uint32_t &bits_out,
                                          uint32_t &exit_out, uint32_t *first_run_end = nullptr,
                                          const PieceHops *hops = nullptr)
{
    const uint32_t lane = lane_id();
    const uint32_t pos0 = pos;
    uint32_t       bits = 0;
    if (first_run_end) *first_run_end = 0;
    while (pos < p1) {
        const uint32_t i = pos + lane;
        const bool     valid = i < p1;
        const uint32_t mv = valid ? m[i] : 0u;
        const uint32_t byte = (kPack && valid) ? src[i] : 0u;
        const uint32_t len = mv >> kMatchOffBits;
        const uint32_t off = mv & ((1u << kMatchOffBits) - 1u);
        const bool     is_long = len >= kSearchMax;
        const uint32_t nxt = lane + (len >= kMinLen ? len : 1u);

        uint32_t j = (is_long || !valid) ? 32u : umin32(nxt, 32u);
        uint32_t reach = 1u;
        if (__any_sync(LZS_FULL_MASK, len >= kMinLen)) {
#pragma unroll
            for (int r = 0; r < 5; r++) {
                const uint32_t c = (((reach >> lane) & 1u) && j < 32u) ? (1u << j) : 0u;
                reach |= __reduce_or_sync(LZS_FULL_MASK, c);
                const uint32_t jj = __shfl_sync(LZS_FULL_MASK, j, static_cast<int>(j & 31u));
                j = (j < 32u) ? jj : 32u;
            }
        } else {
            reach = 0xFFFFFFFFu;            /* literals only (incompressible data): every position starts a token */
        }
        const uint32_t nvalid = umin32(32u, p1 - pos);
        if (nvalid < 32u) reach &= (1u << nvalid) - 1u;
        const bool tok = (reach >> lane) & 1u;

        uint32_t val = 0, nb = 0;
        if (tok) k23_token(len, off, byte, val, nb);
        uint32_t total;
        if (kPack) {
            uint32_t incl = nb;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t v = __shfl_up_sync(LZS_FULL_MASK, incl, static_cast<unsigned>(d));
                if (lane >= static_cast<uint32_t>(d)) incl += v;
            }
            total = __shfl_sync(LZS_FULL_MASK, incl, 31);
            if (tok) stage_put(s.buf, s.cur + incl - nb, val, nb);
            s.cur += total;
            __syncwarp();
        } else {
            total = __reduce_add_sync(LZS_FULL_MASK, nb);      /* counting needs no per-token offsets */
        }
        bits += total;

        const int      last = 31 - __clz(static_cast<int>(reach));
        const uint32_t last_long = __shfl_sync(LZS_FULL_MASK, is_long ? 1u : 0u, last);
        uint32_t       next_pos = pos + __shfl_sync(LZS_FULL_MASK, nxt, last);
        if (kPack) pstage_flush_if_full(s);

        if (last_long) {
            const uint32_t p = pos + static_cast<uint32_t>(last);
            const uint32_t loff = __shfl_sync(LZS_FULL_MASK, off, last);
            uint32_t       L = hops ? k23_long_length_hops(src, m, n, p, loff, hops->piece, hops->run_end, hops->first)
                                    : k23_long_length(src, p, loff, limit);
            if (first_run_end && p == pos0) *first_run_end = p + L;
            if (p + L >= limit && limit < n) {               /* still running where the comparing stops */
                if (!kPack) return false;
                L = known_exit - p;
            }
            bits += k23_nibble_bits(L);
            if (kPack) {
                const uint32_t e = L - kMaxShortLen;
                uint32_t       q = e / kMaxExtLen;
                const uint32_t r = e - q * kMaxExtLen;
                while (q >= 8u) {
                    pstage_emit_uniform(s, 0xFFFFFFFFu, 32u);
                    q -= 8u;
                }
                pstage_emit_uniform(s, (((1u << (4u * q)) - 1u) << 4) | r, 4u * q + 4u);
            }
            next_pos = p + L;
        }
        pos = next_pos;
    }
    bits_out = bits;
    exit_out = pos;
    return true;
}

/* Whole warp, every lane the same: the token at position i -- its bits and where the next one
 * starts; false if a long match is still running at `limit` < n. */
__device__ __forceinline__ bool k23_step(const uint8_t *src, const match_t *m, uint32_t n, uint32_t i, uint32_t limit,
                                         uint32_t &bits, uint32_t &next)
{
    const uint32_t mv = m[i];
    const uint32_t len = mv >> kMatchOffBits;
    const uint32_t off = mv & ((1u << kMatchOffBits) - 1u);
    uint32_t       val, nb;
    k23_token(len, off, 0u, val, nb);
    if (len < kSearchMax) {
        bits = nb;
        next = i + (len >= kMinLen ? len : 1u);
        return true;
    }
    const uint32_t L = k23_long_length(src, i, off, limit);
    if (i + L >= limit && limit < n) return false;
    bits = nb + k23_nibble_bits(L);
    next = i + L;
    return true;
}

/* ---------------------------------------------------------------- spec, fix, sweep, pack */

constexpr int kPieceThreads = 128;
constexpr int kPieceWarps = kPieceThreads / 32;

__device__ __forceinline__ uint32_t piece_limit(uint32_t p1, uint32_t n, uint32_t piece)
{
    const uint64_t far = static_cast<uint64_t>(p1) + static_cast<uint64_t>(kPieceLookPieces) * piece;
    return far < n ? static_cast<uint32_t>(far) : n;
}

__global__ void __launch_bounds__(kPieceThreads)
k23p_spec(const uint8_t *__restrict__ in, const uint64_t *__restrict__ in_off, const uint32_t *__restrict__ in_len,
          const match_t *__restrict__ matches, uint32_t piece, PieceTable t)
{
    const uint32_t idx = blockIdx.x * kPieceWarps + (threadIdx.x >> 5);
    if (idx >= t.count[0]) return;
    const uint32_t sid = t.sid[idx];
    const uint32_t n = in_len[sid];
    const uint32_t p0 = t.p0[idx], p1 = p0 + t.len[idx];
    PieceStage     none = {};
    uint32_t       bits = 0, exit = 0, run_end = 0;
    const bool     ok = k23_piece<false>(in + in_off[sid], matches + in_off[sid], n, p0, p1, piece_limit(p1, n, piece), 0u,
                                         none, bits, exit, &run_end);
    if (lane_id() == 0) {
        t.spec_bits[idx] = bits;
        t.spec_exit[idx] = exit;
        t.flags[idx] = ok ? 0u : kPieceSpecOpen;
        t.run_end[idx] = run_end;
    }
}

__global__ void __launch_bounds__(kPieceThreads)
k23p_fix(const uint8_t *__restrict__ in, const uint64_t *__restrict__ in_off, const uint32_t *__restrict__ in_len,
         const match_t *__restrict__ matches, uint32_t piece, PieceTable t)
{
    const uint32_t idx = blockIdx.x * kPieceWarps + (threadIdx.x >> 5);
    if (idx >= t.count[0]) return;
    const uint32_t sid = t.sid[idx];
    const uint32_t p0 = t.p0[idx], p1 = p0 + t.len[idx];
    const uint32_t flags = t.flags[idx];
    uint32_t       fbits = t.spec_bits[idx], fexit = t.spec_exit[idx];
    bool           open = (flags & kPieceSpecOpen) != 0u;
    if (p0 != 0u) {
        const bool     prev_open = (t.flags[idx - 1] & kPieceSpecOpen) != 0u;
        const uint32_t a = t.spec_exit[idx - 1];               /* the entry assumed here */
        if (prev_open) {
            open = true;                                        /* no assumption to work with */
        } else if (a >= p1) {
            fbits = 0;                                          /* a match runs over the whole piece */
            fexit = a;
            open = false;
        } else if (a != p0) {
            const uint8_t *src = in + in_off[sid];
            const match_t *m = matches + in_off[sid];
            const uint32_t n = in_len[sid];
            const uint32_t limit = piece_limit(p1, n, piece);
            uint32_t       A = p0, B = a, bitsA = 0, bitsB = 0;
            bool           met = false, lost = false;
            for (uint32_t steps = 0; B < p1; steps++) {
                if (A == B) {
                    met = true;
                    break;
                }
                if (steps >= kPieceWalkMax) {
                    lost = true;
                    break;
                }
                uint32_t b = 0, nx = 0;
                if (A < B) {
                    if (!k23_step(src, m, n, A, limit, b, nx)) { lost = true; break; }
                    bitsA += b;
                    A = nx;
                } else {
                    if (!k23_step(src, m, n, B, limit, b, nx)) { lost = true; break; }
                    bitsB += b;
                    B = nx;
                }
            }
            if (lost) {
                open = true;
            } else if (met) {
                if (open) {
                    /* the guess itself is unresolved, and so is everything that joins it */
                } else {
                    fbits = fbits - bitsA + bitsB;
                }
            } else {                                            /* B left the piece on its own */
                fbits = bitsB;
                fexit = B;
                open = false;
            }
        }
    }
    if (lane_id() == 0) {
        t.fix_bits[idx] = fbits;
        t.fix_exit[idx] = fexit;
        if (open) t.flags[idx] = flags | kPieceFixOpen;
    }
}

/* One warp per stream: true entries, bit offsets, the stream's length. */
__global__ void __launch_bounds__(kPieceThreads)
k23p_sweep(const uint8_t *__restrict__ in, const uint64_t *__restrict__ in_off, const uint32_t *__restrict__ in_len,
           const match_t *__restrict__ matches, const uint32_t *__restrict__ out_cap, uint32_t *__restrict__ out_len,
           uint32_t n_streams, uint32_t piece, PieceTable t)
{
    const uint32_t sid = blockIdx.x * kPieceWarps + (threadIdx.x >> 5);
    if (sid >= n_streams || t.count[1]) return;
    const uint32_t lane = lane_id();
    const uint32_t n = in_len[sid];
    const uint8_t *src = in + in_off[sid];
    const match_t *m = matches + in_off[sid];
    const uint32_t first = t.first[sid], np = t.first[sid + 1] - first;
    uint32_t       e = 0;                                       /* true entry of the piece at hand */
    uint64_t       bit = 0;
    for (uint32_t k0 = 0; k0 < np; k0 += 32u) {
        const uint32_t cnt = umin32(32u, np - k0);
        const uint32_t idx = first + k0 + lane;
        const bool     have = lane < cnt;
        const uint32_t r_p0 = have ? t.p0[idx] : 0u, r_len = have ? t.len[idx] : 0u;
        const uint32_t r_flags = have ? t.flags[idx] : 0u;
        const uint32_t r_sx = have ? t.spec_exit[idx] : 0u, r_sb = have ? t.spec_bits[idx] : 0u;
        const uint32_t r_fx = have ? t.fix_exit[idx] : 0u, r_fb = have ? t.fix_bits[idx] : 0u;
        uint32_t       r_a = (have && k0 + lane != 0u) ? t.spec_exit[idx - 1] : 0u;
        if (have && k0 + lane != 0u && (t.flags[idx - 1] & kPieceSpecOpen)) r_a = 0xFFFFFFFFu;   /* no assumption was made */
        uint32_t my_entry = 0, my_exit = 0;
        uint64_t my_bit = 0;
        for (uint32_t k = 0; k < cnt; k++) {
            const uint32_t p0 = __shfl_sync(LZS_FULL_MASK, r_p0, static_cast<int>(k));
            const uint32_t p1 = p0 + __shfl_sync(LZS_FULL_MASK, r_len, static_cast<int>(k));
            const uint32_t flags = __shfl_sync(LZS_FULL_MASK, r_flags, static_cast<int>(k));
            const uint32_t a = __shfl_sync(LZS_FULL_MASK, r_a, static_cast<int>(k));
            uint32_t       bits = 0, x = e;
            if (e >= p1 && p1 != n) {
                /* a match from an earlier piece runs over this one: no token starts here */
            } else if (e >= p1) {
                /* the last piece, and the last token ended with the stream (or the stream is empty) */
            } else if (e == a && !(flags & kPieceFixOpen)) {
                bits = __shfl_sync(LZS_FULL_MASK, r_fb, static_cast<int>(k));
                x = __shfl_sync(LZS_FULL_MASK, r_fx, static_cast<int>(k));
            } else if (e == p0 && !(flags & kPieceSpecOpen)) {
                bits = __shfl_sync(LZS_FULL_MASK, r_sb, static_cast<int>(k));
                x = __shfl_sync(LZS_FULL_MASK, r_sx, static_cast<int>(k));
            } else {
                PieceStage     none = {};
                const PieceHops hops = {t.run_end, first, piece};
                k23_piece<false>(src, m, n, e, p1, n, 0u, none, bits, x, nullptr, &hops);
            }
            if (lane == k) {
                my_entry = e;
                my_bit = bit;
                my_exit = x;
            }
            bit += bits;
            e = x;
        }
        if (have) {
            t.entry[idx] = my_entry;
            t.bitoff[idx] = my_bit;
            t.exit[idx] = my_exit;
        }
    }
    if (lane == 0) {
        const uint64_t bytes = (bit + 9u + 7u) >> 3;            /* end marker, padded to a byte */
        const uint32_t cap = out_cap[sid];
        out_len[sid] = bytes < cap ? static_cast<uint32_t>(bytes) : cap;
    }
}

__global__ void __launch_bounds__(kPieceThreads)
k23p_pack(const uint8_t *__restrict__ in, const uint64_t *__restrict__ in_off, const uint32_t *__restrict__ in_len,
          const match_t *__restrict__ matches, uint8_t *__restrict__ out, const uint64_t *__restrict__ out_off,
          const uint32_t *__restrict__ out_cap, PieceTable t)
{
    __shared__ uint32_t s_buf[kPieceWarps][64];
    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    const uint32_t idx = blockIdx.x * kPieceWarps + warp;
    if (idx >= t.count[0]) return;
    const uint32_t sid = t.sid[idx];
    const uint32_t n = in_len[sid];
    const uint32_t p1 = t.p0[idx] + t.len[idx];
    const uint32_t e = t.entry[idx], x = t.exit[idx];
    /* nothing starts in this piece -- unless it is the one that has to close an empty stream */
    if (e >= p1 && !(n == 0u)) return;
    const uint8_t *src = in + in_off[sid];
    const match_t *m = matches + in_off[sid];
    const uint64_t bit0 = t.bitoff[idx];

    PieceStage s;
    s.buf = s_buf[warp];
    s.dst = out + out_off[sid];
    s.lo = (bit0 + 7u) >> 3;
    s.hi = out_cap[sid];
    s.cur = static_cast<uint32_t>(bit0 & 31u);
    s.wdone = static_cast<uint32_t>(bit0 >> 5);
    s.aligned = (reinterpret_cast<uintptr_t>(s.dst) & 3u) == 0;
    s.buf[lane] = 0;
    s.buf[lane + 32] = 0;
    __syncwarp();

    uint32_t bits = 0, exit = 0;
    k23_piece<true>(src, m, n, e, p1, p1, x, s, bits, exit);

    if (exit >= n) {
        pstage_emit_uniform(s, 0x180u, 9u);                     /* end marker; the zero padding is in the stage already */
        s.cur = (s.cur + 7u) & ~7u;
    } else {
        /* the byte this piece ends in is completed with the first bits of the next token */
        const uint32_t k = (8u - (s.cur & 7u)) & 7u;
        if (k) {
            const uint32_t mv = m[exit];
            uint32_t       val, nb;
            k23_token(mv >> kMatchOffBits, mv & ((1u << kMatchOffBits) - 1u), src[exit], val, nb);
            pstage_emit_uniform(s, val >> (nb - k), k);
        }
    }
    const uint32_t rest = s.cur >> 3;
    const uint64_t base = static_cast<uint64_t>(s.wdone) * 4u;
    for (uint32_t k = lane; k < rest; k += 32u) {
        const uint32_t b = (s.buf[k >> 2] >> (24u - 8u * (k & 3u))) & 0xFFu;
        if (base + k >= s.lo && base + k < s.hi) s.dst[base + k] = static_cast<uint8_t>(b);
    }
}

}  // namespace lzs

#endif /* LZS_B200_K23_PIECES_CUH */
