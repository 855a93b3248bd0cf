/*
 * k4_pieces.cuh -- the decoder for LONG streams: where the tokens of an LZS stream start is found in
 * parallel INSIDE the stream, so that a batch of a few large streams does not decode at the rate of
 * one serial parser per stream (k4_decode.cuh: ~9 MB/s per stream, whatever the GPU has idle).
 *
 * Reference: c/src/liblzs/lzs-decompression.c:156-412.  What is parallel here and what is not:
 *
 *   A token's bit length is a function of the bit position it starts at (literal 9; match 9 or 13
 *   bits of offset, 2 or 4 of length, then 4-bit continuations while they are 1111), so the token
 *   starts are the orbit of next(b) = b + bits(b) from the stream's first bit -- the same structure
 *   as the compressor's parse (k23_pieces.cuh), over bit positions.  The compressed stream is cut
 *   into pieces of C bytes:
 *     spec   one THREAD per piece parses from the piece's first bit -- a guess -- and records the
 *            output bytes its tokens make and the first token start behind the piece (exit);
 *     fix    one thread per piece takes the exit of the piece before as its entry and walks both
 *            orbits until they meet (a wrong guess is noise, but noise re-synchronises);
 *     sweep  one warp per stream: true entries in order, a running sum of output bytes gives every
 *            piece its place in the output, the stream its length;
 *     emit   one thread per piece parses from its true entry: literals go straight to their place
 *            in the output, matches become 4-byte records (literals before it, length, offset);
 *     copy   one warp per stream replays the records in order.  THIS is the serial part that is
 *            left -- a match may copy bytes the match before it produced -- but it is a load, a
 *            store and a barrier per match instead of a bit parser's dependent chain.
 *
 * Only CLEAN streams are finished here: every token complete, no offset of zero other than the end
 * marker, no offset that reaches before the start of the output, the first end marker reached, and
 * the output fits.  Anything else sets the stream's dirty flag, and k4_decode -- which restates the
 * reference's behaviour on malformed streams and short capacities rule by rule -- decodes the dirty
 * streams afterwards from scratch (a launch that ends at once when there are none).
 */
#ifndef LZS_B200_K4_PIECES_CUH
#define LZS_B200_K4_PIECES_CUH

#include "k4_decode.cuh"

namespace lzs {

constexpr uint32_t kDPieceDead = 0xFFFFFFFFu;           /* no token of the stream starts in this piece */
constexpr uint32_t kDTokLiteral = 0, kDTokMatch = 1, kDTokEnd = 2, kDTokBad = 3;
constexpr uint32_t kDStOk = 0, kDStEnd = 1, kDStBad = 2, kDStOpen = 3;
constexpr uint32_t kDOutMax = 0xF0000000u;              /* output positions are 32-bit                 */
constexpr uint32_t kDWalkMax = 1u << 16;                /* fix gives up after this many tokens         */
constexpr uint32_t kDRecLitsMax = 1023, kDRecLenMax = 2047;
constexpr uint32_t kDGuesses = 1;                       /* start bits spec tries until an orbit survives its piece (more than one: measured slower, the
                                                           retries of a few threads hold their warps) */

/* record words a piece of `piece` compressed bytes can make: one per match (>= 11 bits of input), two for
 * a match longer than 2047 bytes (>= 560 bits), one per 1023 literals, and the one that carries the
 * literals behind the last match */
__host__ __device__ inline uint32_t dpiece_record_stride(uint32_t piece) { return (piece * 8u) / 9u + 4u; }
/* lengths of the matches longer than 2047 bytes that can start in a piece (each spends >= 560 bits) */
__host__ __device__ inline uint32_t dpiece_long_stride(uint32_t piece) { return (piece * 8u) / 560u + 2u; }

struct DPieceTable {
    uint32_t *spec_exit, *spec_out, *spec_status;
    uint32_t *fix_exit, *fix_out, *fix_status, *fix_entry;   /* parse from fix_entry, the exit of the piece before */
    uint32_t *entry, *out_at, *nrec;      /* sweep: true entry bit (or dead) and output position; emit: records made */
    uint32_t *first;                      /* [n_streams + 1]                                               */
    uint32_t *dirty;                      /* [n_streams] stream goes to k4_decode                          */
    uint32_t *dirty_list;                 /* [n_streams] compacted                                         */
    uint32_t *count;                      /* [0] pieces, [1] table overflow, [2] dirty streams, [3] work counter of the fallback */
    uint32_t *records;                    /* cap * stride                                                  */
    uint32_t *longs;                      /* cap * lstride: lengths of the piece's long matches, in order  */
    uint32_t  cap, stride, lstride;
};
constexpr size_t kDPieceEntryWords = 10 + 3;            /* per piece, plus first / dirty / dirty_list per stream (<= cap) */

__host__ __device__ inline size_t dpiece_table_bytes(uint32_t cap, uint32_t piece)
{
    return 256 + static_cast<size_t>(cap) * 4u * (kDPieceEntryWords + dpiece_record_stride(piece) + dpiece_long_stride(piece));
}

__host__ __device__ inline DPieceTable dpiece_table_at(void *base, uint32_t cap, uint32_t piece)
{
    DPieceTable t;
    uint32_t   *w = static_cast<uint32_t *>(base);
    t.count = w;        w += 64;
    t.spec_exit = w;    w += cap;
    t.spec_out = w;     w += cap;
    t.spec_status = w;  w += cap;
    t.fix_exit = w;     w += cap;
    t.fix_out = w;      w += cap;
    t.fix_status = w;   w += cap;
    t.fix_entry = w;    w += cap;
    t.entry = w;        w += cap;
    t.out_at = w;       w += cap;
    t.nrec = w;         w += cap;
    t.first = w;        w += cap;
    t.dirty = w;        w += cap;
    t.dirty_list = w;   w += cap;
    t.records = w;
    t.stride = dpiece_record_stride(piece);
    t.lstride = dpiece_long_stride(piece);
    t.longs = w + static_cast<size_t>(cap) * t.stride;
    t.cap = cap;
    return t;
}

/* ---------------------------------------------------------------- bits of one stream */

struct BitSrc {
    const uint32_t *wbase;        /* aligned word that holds the stream's first byte   */
    uint32_t        nwords, tail; /* words that hold stream bytes; bytes of the last   */
    uint32_t        first, end;   /* first bit of the stream, bit behind its last      */
};

__device__ __forceinline__ BitSrc bitsrc_open(const uint8_t *p, uint32_t len)
{
    BitSrc          s;
    const uint32_t  nin = umin32(len, 0x1FFFFF00u);                 /* 32-bit bit positions, as k4_decode */
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t  lead = static_cast<uint32_t>(a & 3u);
    s.wbase = reinterpret_cast<const uint32_t *>(a & ~static_cast<uintptr_t>(3));
    s.nwords = (lead + nin + 3u) >> 2;
    s.tail = (lead + nin) & 3u;
    s.first = 8u * lead;
    s.end = s.first + 8u * nin;
    return s;
}

__device__ __forceinline__ uint32_t bitsrc_word(const BitSrc &s, uint32_t w)
{
    if (w >= s.nwords) return 0u;
    uint32_t v = bswap32(__ldg(s.wbase + w));
    if (w == s.nwords - 1u && s.tail) v &= 0xFFFFFFFFu << (8u * (4u - s.tail));
    return v;
}

/* 32 bits of the stream from bit b (zeros behind its end).  A parser moves forward, so the reader
 * keeps the word at hand and the two behind it: one load per 32 bits of stream, issued two words
 * before its bits are looked at. */
struct BitCache {
    uint32_t widx, w0, w1, w2;
};
__device__ __forceinline__ BitCache bitcache_none()
{
    BitCache c;
    c.widx = 0xFFFFFFF0u;
    c.w0 = c.w1 = c.w2 = 0;
    return c;
}
__device__ __forceinline__ uint32_t bitsrc_peek(const BitSrc &s, BitCache &c, uint32_t b)
{
    const uint32_t w = b >> 5;
    if (w != c.widx) {
        if (w == c.widx + 1u) {
            c.w0 = c.w1;
            c.w1 = c.w2;
            c.w2 = bitsrc_word(s, w + 2u);
        } else {
            c.w0 = bitsrc_word(s, w);
            c.w1 = bitsrc_word(s, w + 1u);
            c.w2 = bitsrc_word(s, w + 2u);
        }
        c.widx = w;
    }
    return __funnelshift_l(c.w1, c.w0, b & 31u);
}

struct DTok {
    uint32_t kind, used, out, off, byte;
};

/* The token that starts at bit b (lzs-decompression.c:178-408 for one token, continuations included).
 * Bad: not complete inside the stream, a long offset of zero, or longer than 32-bit positions hold --
 * whatever k4_decode has a rule for and a clean stream does not contain. */
__device__ __forceinline__ DTok dtoken(const BitSrc &s, BitCache &c, uint32_t b)
{
    DTok t;
    t.kind = kDTokBad;
    t.used = t.out = t.off = t.byte = 0;
    if (b >= s.end) return t;
    const uint32_t left = s.end - b;
    if (left < 9u) return t;
    const uint32_t top = bitsrc_peek(s, c, b);
    if ((top >> 31) == 0u) {
        t.kind = kDTokLiteral;
        t.used = 9u;
        t.out = 1u;
        t.byte = (top >> 23) & 0xFFu;
        return t;
    }
    const uint32_t is_short = (top >> 30) & 1u;
    const uint32_t hdr = is_short ? 9u : 13u;
    if (left < hdr) return t;
    const uint32_t o = is_short ? ((top >> 23) & 0x7Fu) : ((top >> 19) & 0x7FFu);
    if (o == 0u) {
        if (is_short) {
            t.kind = kDTokEnd;
            t.used = 9u;
        }
        return t;
    }
    const uint32_t code = (top << hdr) >> 28;
    uint32_t       w, len;
    if (code < 12u) { len = (code >> 2) + 2u; w = 2u; }
    else            { len = code - 7u;        w = 4u; }
    if (left < hdr + w) return t;
    uint32_t used = hdr + w;
    if (len == kMaxShortLen) {
        for (;;) {
            if (s.end - (b + used) < 4u) return t;
            const uint32_t next = bitsrc_peek(s, c, b + used);
            if (next == 0xFFFFFFFFu && s.end - (b + used) >= 36u) {      /* eight continuations at once */
                used += 32u;
                len += 8u * kMaxExtLen;
                if (len > kDOutMax) return t;
                continue;
            }
            const uint32_t nib = next >> 28;
            used += 4u;
            len += nib;
            if (nib != kMaxExtLen) break;
        }
    }
    t.kind = kDTokMatch;
    t.used = used;
    t.out = len;
    t.off = o;
    return t;
}

/* ---------------------------------------------------------------- plan */

__global__ void __launch_bounds__(1024)
k4p_plan(const uint32_t *__restrict__ in_len, uint32_t n_streams, uint32_t piece, DPieceTable t)
{
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_base;
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (uint32_t s0 = 0; s0 < n_streams; s0 += 1024u) {
        const uint32_t s = s0 + threadIdx.x;
        const uint32_t n = s < n_streams ? umin32(in_len[s], 0x1FFFFF00u) : 0u;
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
        if (s < n_streams) {
            t.first[s] = before + incl - np;
            t.dirty[s] = 0;
        }
        __syncthreads();
        if (threadIdx.x == 1023) s_base = before + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        t.first[n_streams] = s_base;
        t.count[0] = s_base <= t.cap ? s_base : 0u;
        t.count[1] = s_base <= t.cap ? 0u : 1u;
        t.count[2] = 0;
        t.count[3] = 0;
    }
}

/* piece index -> (stream, piece of the stream): binary search in first[] */
__device__ __forceinline__ uint32_t dpiece_stream(const uint32_t *first, uint32_t n_streams, uint32_t idx)
{
    uint32_t lo = 0, hi = n_streams;                    /* first[lo] <= idx < first[hi] */
    while (hi - lo > 1u) {
        const uint32_t mid = (lo + hi) >> 1;
        if (first[mid] <= idx) lo = mid; else hi = mid;
    }
    return lo;
}

/* ---------------------------------------------------------------- spec, fix */

__global__ void __launch_bounds__(128)
k4p_spec(const uint8_t *__restrict__ in, const uint64_t *__restrict__ in_off, const uint32_t *__restrict__ in_len,
         uint32_t n_streams, uint32_t piece, DPieceTable t)
{
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= t.count[0]) return;
    const uint32_t sid = dpiece_stream(t.first, n_streams, idx);
    const BitSrc   s = bitsrc_open(in + in_off[sid], in_len[sid]);
    const uint32_t k = idx - t.first[sid];
    const uint32_t p0 = s.first + 8u * k * piece;
    const uint32_t pend = umin32(p0 + 8u * piece, s.end);
    /* A guess that is noise often runs into an end marker or an impossible token before it has joined
     * the true orbit; the next bit is then tried (a guess that reaches the end of the piece has very
     * likely joined).  The first piece of a stream has nothing to guess. */
    uint32_t b = p0, out = 0, status = kDStOk, shift = 0;
    for (uint32_t g = 0; g < kDGuesses; g++) {
        uint32_t gb = p0 + g, gout = 0, gstatus = kDStOk;
        BitCache bc = bitcache_none();
        while (gb < pend) {
            const DTok tk = dtoken(s, bc, gb);
            if (tk.kind == kDTokEnd) { gstatus = kDStEnd; break; }
            if (tk.kind == kDTokBad || gout + tk.out > kDOutMax) { gstatus = kDStBad; break; }
            gout += tk.out;
            gb += tk.used;
        }
        if (g == 0 || gstatus == kDStOk) {
            b = gb; out = gout; status = gstatus; shift = g;
        }
        if (gstatus == kDStOk || k == 0u || p0 + g + 1u >= pend) break;
    }
    status |= shift << 8;                                /* where the recorded guess started */
    t.spec_exit[idx] = b;
    t.spec_out[idx] = out;
    t.spec_status[idx] = status;
}

/* pass 0: the entry assumed is the exit of the guess of the piece before.  Where that guess was noise
 * that ended early (it ran into an end-marker pattern before it joined the true orbit) there is
 * nothing to assume yet, and where it never joined, the assumption is wrong; pass 1 takes the exit
 * pass 0 found for that piece instead. */
__global__ void __launch_bounds__(128)
k4p_fix(const uint8_t *__restrict__ in, const uint64_t *__restrict__ in_off, const uint32_t *__restrict__ in_len,
        uint32_t n_streams, uint32_t piece, uint32_t pass, DPieceTable t)
{
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= t.count[0]) return;
    const uint32_t sid = dpiece_stream(t.first, n_streams, idx);
    const uint32_t k = idx - t.first[sid];
    const uint32_t sraw = t.spec_status[idx];
    uint32_t       fexit = t.spec_exit[idx], fout = t.spec_out[idx], fstatus = sraw & 0xFFu, fentry = kDPieceDead;
    /* pass 1 repairs what pass 0 could not know: pieces it left open, and pieces whose assumed entry is
     * not the exit pass 0 found for the piece before (that piece's guess never joined its true orbit).
     * Whatever is still wrong afterwards costs the sweep a serial parse of the piece. */
    if (pass != 0u) {
        if (k == 0u) return;
        const bool open = t.fix_status[idx] == kDStOpen;
        const bool stale = t.fix_status[idx - 1] == kDStOk && t.fix_exit[idx - 1] != t.fix_entry[idx];
        if (!open && !stale) return;
    }
    if (k != 0u) {
        const BitSrc   s = bitsrc_open(in + in_off[sid], in_len[sid]);
        const uint32_t pstart = s.first + 8u * k * piece;
        const uint32_t pend = umin32(pstart + 8u * piece, s.end);
        const uint32_t p0 = pstart + (sraw >> 8);        /* where this piece's guess started */
        const bool     have = pass == 0u ? (t.spec_status[idx - 1] & 0xFFu) == kDStOk : t.fix_status[idx - 1] == kDStOk;
        const uint32_t a = pass == 0u ? t.spec_exit[idx - 1] : t.fix_exit[idx - 1];
        if (!have) {
            fstatus = kDStOpen;                          /* no entry to assume: pass 1, or the sweep, decides */
        } else if (a >= pend) {
            fentry = a;
            fexit = a;                                   /* a token of the piece before runs over this one */
            fout = 0;
            fstatus = kDStOk;
        } else {
            fentry = a;
            if (a != p0) {
                uint32_t A = p0, B = a, outA = 0, outB = 0;
                bool     a_alive = true, met = false, lost = false;
                uint32_t b_status = kDStOk;
                BitCache ca = bitcache_none(), cb = bitcache_none();
                for (uint32_t steps = 0; B < pend; steps++) {
                    if (a_alive && A == B) { met = true; break; }
                    if (steps >= kDWalkMax) { lost = true; break; }
                    if (a_alive && A < B) {
                        const DTok tk = dtoken(s, ca, A);
                        if (tk.kind == kDTokEnd || tk.kind == kDTokBad) a_alive = false;   /* the guess was noise and ended */
                        else { outA += tk.out; A += tk.used; }
                    } else {
                        const DTok tk = dtoken(s, cb, B);
                        if (tk.kind == kDTokEnd) { b_status = kDStEnd; break; }
                        if (tk.kind == kDTokBad || outB + tk.out > kDOutMax) { b_status = kDStBad; break; }
                        outB += tk.out;
                        B += tk.used;
                    }
                }
                if (lost) {
                    fstatus = kDStOpen;
                } else if (met) {
                    fout = fout - outA + outB;           /* exit and status as guessed */
                    if (fstatus == kDStBad || fout > kDOutMax) fstatus = kDStBad;
                } else {                                 /* B went through the piece on its own */
                    fexit = B;
                    fout = outB;
                    fstatus = b_status;
                }
            }
        }
    }
    t.fix_exit[idx] = fexit;
    t.fix_out[idx] = fout;
    t.fix_entry[idx] = fentry;
    t.fix_status[idx] = fstatus;
}

/* ---------------------------------------------------------------- sweep */

/* One warp per stream.  A clean stream gets its length and status here; a stream that is not clean
 * gets its dirty flag (and whatever the later passes do to its output slot is overwritten by k4_decode). */
__global__ void __launch_bounds__(128)
k4p_sweep(const uint8_t *__restrict__ in, const uint64_t *__restrict__ in_off, const uint32_t *__restrict__ in_len,
          const uint32_t *__restrict__ out_cap, uint32_t *__restrict__ out_len, uint8_t *__restrict__ status,
          uint32_t n_streams, uint32_t piece, DPieceTable t, uint32_t *__restrict__ in_used = nullptr)
{
    const uint32_t sid = blockIdx.x * 4u + (threadIdx.x >> 5);
    if (sid >= n_streams) return;
    const uint32_t lane = lane_id();
    if (t.count[1]) {                                    /* no table: everything to k4_decode */
        if (lane == 0) t.dirty[sid] = 1;
        if (lane == 0 && in_used != nullptr) in_used[sid] = 0xFFFFFFFFu;
        return;
    }
    const BitSrc   s = bitsrc_open(in + in_off[sid], in_len[sid]);
    const uint32_t first = t.first[sid], np = t.first[sid + 1] - first;
    const uint32_t cap = out_cap[sid];
    uint32_t       e = s.first;                          /* true entry of the piece at hand */
    uint64_t       pos = 0;
    bool           ended = false, bad = false;
    for (uint32_t k0 = 0; k0 < np; k0 += 32u) {
        const uint32_t cnt = umin32(32u, np - k0);
        const uint32_t idx = first + k0 + lane;
        const bool     have = lane < cnt;
        const uint32_t r_sx = have ? t.spec_exit[idx] : 0u, r_so = have ? t.spec_out[idx] : 0u;
        const uint32_t r_ss = have ? t.spec_status[idx] : 0u;  /* status | start shift << 8 */
        const uint32_t r_fx = have ? t.fix_exit[idx] : 0u, r_fo = have ? t.fix_out[idx] : 0u;
        const uint32_t r_fs = have ? t.fix_status[idx] : 0u;
        const uint32_t r_a = have ? t.fix_entry[idx] : kDPieceDead;   /* the entry fix assumed, if it assumed one */
        uint32_t my_entry = kDPieceDead, my_at = 0;
        /* The usual case, 32 pieces at once: every piece's assumed entry is the exit fix found for the
         * piece before it, so all 32 results stand and a scan places them.  (The first and the last
         * pieces of a stream, and whatever fix left open, take the loop below.) */
        {
            uint32_t prev_exit = __shfl_up_sync(LZS_FULL_MASK, r_fx, 1);
            if (lane == 0) prev_exit = e;
            const bool chain = !have || (r_fs == kDStOk && r_a == prev_exit);
            if (!ended && !bad && __all_sync(LZS_FULL_MASK, chain)) {
                const uint32_t pend = umin32(s.first + 8u * (k0 + lane + 1u) * piece, s.end);
                uint64_t       incl = have ? r_fo : 0u;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint64_t u = __shfl_up_sync(LZS_FULL_MASK, incl, static_cast<unsigned>(d));
                    if (lane >= static_cast<uint32_t>(d)) incl += u;
                }
                const uint64_t at = pos + incl - (have ? r_fo : 0u);
                pos += __shfl_sync(LZS_FULL_MASK, incl, 31);
                e = __shfl_sync(LZS_FULL_MASK, r_fx, static_cast<int>(cnt - 1u));
                if (pos > cap) bad = true;
                if (have) {
                    t.entry[idx] = (bad || prev_exit >= pend) ? kDPieceDead : prev_exit;
                    t.out_at[idx] = static_cast<uint32_t>(at);
                }
                continue;
            }
        }
        for (uint32_t k = 0; k < cnt; k++) {
            const uint32_t p0 = s.first + 8u * (k0 + k) * piece;
            const uint32_t pend = umin32(p0 + 8u * piece, s.end);
            uint32_t       out = 0, x = e, st = kDStOk;
            bool           live = false;
            if (ended || bad || (e >= pend && !(np == 1u && s.end == s.first))) {
                /* the stream is over, or a token of an earlier piece runs over this one */
            } else if (e >= pend) {
                st = kDStBad;                            /* an empty stream has no end marker */
            } else {
                live = true;
                const uint32_t a = __shfl_sync(LZS_FULL_MASK, r_a, static_cast<int>(k));
                const uint32_t fs = __shfl_sync(LZS_FULL_MASK, r_fs, static_cast<int>(k));
                const uint32_t ss = __shfl_sync(LZS_FULL_MASK, r_ss, static_cast<int>(k));
                if (k0 + k != 0u && e == a && fs != kDStOpen) {
                    out = __shfl_sync(LZS_FULL_MASK, r_fo, static_cast<int>(k));
                    x = __shfl_sync(LZS_FULL_MASK, r_fx, static_cast<int>(k));
                    st = fs;
                } else if (e == p0 + (ss >> 8)) {
                    out = __shfl_sync(LZS_FULL_MASK, r_so, static_cast<int>(k));
                    x = __shfl_sync(LZS_FULL_MASK, r_sx, static_cast<int>(k));
                    st = ss & 0xFFu;
                } else {                                 /* parse the piece again from e (every lane the same) */
                    uint32_t b = e;
                    BitCache bc = bitcache_none();
                    while (b < pend) {
                        const DTok tk = dtoken(s, bc, b);
                        if (tk.kind == kDTokEnd) { st = kDStEnd; break; }
                        if (tk.kind == kDTokBad || out + tk.out > kDOutMax) { st = kDStBad; break; }
                        out += tk.out;
                        b += tk.used;
                    }
                    x = b;
                }
            }
            if (lane == k && live) {
                my_entry = e;
                my_at = static_cast<uint32_t>(pos);
            }
            pos += out;
            e = x;
            if (st == kDStEnd) ended = true;
            if (st == kDStBad || pos > cap) bad = true;
        }
        if (have) {
            t.entry[idx] = bad ? kDPieceDead : my_entry;
            t.out_at[idx] = my_at;
        }
    }
    /* (entries stored before `bad` was known are harmless: a dirty stream is decoded again) */
    if (lane == 0) {
        /* bytes of the stream up to and including its end marker (e stands on it), for callers that
         * walk a file of several streams; unknown for a stream that is not clean */
        if (in_used != nullptr) in_used[sid] = (bad || !ended) ? 0xFFFFFFFFu : (e + 9u - s.first + 7u) >> 3;
        if (bad || !ended) {
            t.dirty[sid] = 1;
        } else {
            out_len[sid] = static_cast<uint32_t>(pos);
            /* the decoder stops on a full output before it reads the end marker (lzs-decompression.c:200-203) */
            if (status != nullptr) status[sid] = static_cast<uint8_t>(pos >= cap ? kDecNoSpace : kDecEndMarker);
        }
    }
}

/* ---------------------------------------------------------------- emit */

/* record: literals in front (10 bits) | length (11 bits) | offset (11 bits).  Length 0 and offset 0:
 * literals only.  Length 0 with an offset: a match longer than 2047 bytes, its length is the next
 * entry of the piece's `longs`. */
__device__ __forceinline__ uint32_t drecord(uint32_t lits, uint32_t len, uint32_t off) { return (lits << 22) | (len << 11) | off; }

__global__ void __launch_bounds__(128)
k4p_emit(const uint8_t *__restrict__ in, const uint64_t *__restrict__ in_off, const uint32_t *__restrict__ in_len,
         uint8_t *__restrict__ out, const uint64_t *__restrict__ out_off, uint32_t n_streams, uint32_t piece, DPieceTable t)
{
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= t.count[0]) return;
    const uint32_t sid = dpiece_stream(t.first, n_streams, idx);
    uint32_t       nrec = 0;
    const uint32_t e = t.entry[idx];
    if (e != kDPieceDead && !t.dirty[sid]) {
        const BitSrc   s = bitsrc_open(in + in_off[sid], in_len[sid]);
        const uint32_t k = idx - t.first[sid];
        const uint32_t pend = umin32(s.first + 8u * (k + 1u) * piece, s.end);
        uint8_t       *dst = out + out_off[sid];
        uint32_t      *rec = t.records + static_cast<size_t>(idx) * t.stride;
        uint32_t      *longs = t.longs + static_cast<size_t>(idx) * t.lstride;
        uint32_t       b = e, pos = t.out_at[idx], lits = 0, nlong = 0;
        BitCache       bc = bitcache_none();
        while (b < pend) {
            const DTok tk = dtoken(s, bc, b);
            if (tk.kind == kDTokLiteral) {
                dst[pos++] = static_cast<uint8_t>(tk.byte);
                if (++lits == kDRecLitsMax) {
                    rec[nrec++] = drecord(lits, 0u, 0u);
                    lits = 0;
                }
            } else if (tk.kind == kDTokMatch) {
                if (tk.out <= kDRecLenMax) {
                    rec[nrec++] = drecord(lits, tk.out, tk.off);
                } else {
                    rec[nrec++] = drecord(lits, 0u, tk.off);
                    longs[nlong++] = tk.out;
                }
                lits = 0;
                pos += tk.out;
            } else {
                break;                                   /* the end marker (the sweep saw nothing bad) */
            }
            b += tk.used;
        }
        if (lits) rec[nrec++] = drecord(lits, 0u, 0u);
    }
    t.nrec[idx] = nrec;
}

/* ---------------------------------------------------------------- copy */

/* One block per stream replays the match records in order -- the serial step that is left.  The
 * block keeps the last 4 KiB of the output in shared memory: bytes enter it from the output slot
 * (where emit has put the literals), matches are copied inside it, finished bytes go back to the slot.
 *   Byte k of a match is byte (k mod offset) behind pos - offset, all of which exist before the match
 * starts, so a match depends on earlier matches only through its source bytes.  128 records are
 * taken at a time, one per thread; a scan of literals + lengths gives every match its place, and a
 * thread copies its match as soon as everything in front of its source range is final -- most
 * sources lie before the 128 records altogether, so a few rounds do what 128 dependent steps
 * would.  Matches longer than 16 bytes are copied by the 32 lanes of their thread's warp; a record
 * that does not fit the window (2 KiB ahead) is copied in the slot itself by the whole block. */
constexpr uint32_t kDRing = 4096, kDRingAhead = 2048;
#ifndef LZS_K4P_LANE_LEN
#define LZS_K4P_LANE_LEN 16
#endif
constexpr uint32_t kDLaneLen = LZS_K4P_LANE_LEN;        /* matches up to this long are copied by one thread (4 / 8 / 16 / 32: no
                                                           difference beyond the noise between boxes) */
#ifndef LZS_K4P_COPY_THREADS
#define LZS_K4P_COPY_THREADS 128
#endif
constexpr int      kDCopyThreads = LZS_K4P_COPY_THREADS;   /* records replayed together, one per thread (measured: 64 / 128 / 256
                                                              threads -> 24.1 / 23.8 / 29.2 ms per GiB in 1 MiB chunks, 2.6 / 1.9 / 1.6 s for
                                                              4 streams of 256 MiB; loading the next records a batch ahead: no change) */
constexpr uint32_t kDCopyWarps = kDCopyThreads / 32;

/* bytes [flushed, upto) leave the ring, bytes [loaded, ...) enter it as far as the window allows; whole block */
__device__ __noinline__ void dring_refill(uint8_t *ring, uint8_t *dst, uint32_t total, uint32_t upto, uint32_t &flushed,
                                          uint32_t &loaded)
{
    constexpr uint32_t kMask = kDRing - 1u, kT = kDCopyThreads;
    for (uint32_t p = flushed + threadIdx.x; p < upto; p += kT) dst[p] = ring[p & kMask];
    flushed = upto;
    __syncthreads();
    const uint32_t want = umin32(total, upto + kDRingAhead);
    uint32_t       p = loaded + threadIdx.x;
    for (; p + 7u * kT < want; p += 8u * kT) {           /* eight loads in flight per thread */
        uint8_t b[8];
#pragma unroll
        for (uint32_t j = 0; j < 8u; j++) b[j] = dst[p + kT * j];
#pragma unroll
        for (uint32_t j = 0; j < 8u; j++) ring[(p + kT * j) & kMask] = b[j];
    }
    for (; p < want; p += kT) ring[p & kMask] = dst[p];
    if (want > loaded) loaded = want;
    __syncthreads();
}

__global__ void __launch_bounds__(kDCopyThreads)
k4p_copy(uint8_t *out, const uint64_t *__restrict__ out_off, const uint32_t *__restrict__ out_len, uint32_t n_streams,
         DPieceTable t)
{
    __shared__ uint8_t  ring[kDRing];
    __shared__ uint32_t s_start[kDCopyThreads];          /* where every thread's match starts */
    __shared__ uint32_t s_mend[kDCopyThreads];           /* and ends */
    __shared__ uint32_t s_w[2][kDCopyWarps];             /* per-warp words of the block-wide steps */
    __shared__ uint32_t s_any;
    const uint32_t sid = blockIdx.x;
    if (sid >= n_streams || t.count[1] || t.dirty[sid]) return;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    uint8_t       *dst = out + out_off[sid];
    const uint32_t total = out_len[sid];                 /* the sweep's */
    const uint32_t first = t.first[sid], np = t.first[sid + 1] - first;
    constexpr uint32_t kMask = kDRing - 1u, kT = kDCopyThreads;
    uint32_t       pos = 0, flushed = 0, loaded = 0;     /* output bytes done / written back / present in the ring; the same in every thread */
    bool           before_start = false;

    /* sum over the warps before mine, and over all, of one word per warp */
    auto across_warps = [&](uint32_t mine_total, uint32_t &before, uint32_t &all, int buf) {
        if (lane == 0) s_w[buf][warp] = mine_total;
        __syncthreads();
        before = 0;
        all = 0;
#pragma unroll
        for (uint32_t w = 0; w < kDCopyWarps; w++) {
            const uint32_t x = s_w[buf][w];
            if (w < warp) before += x;
            all += x;
        }
    };

    dring_refill(ring, dst, total, 0u, flushed, loaded);
    for (uint32_t k = 0; k < np && !before_start; k++) {
        const uint32_t idx = first + k;
        const uint32_t nrec = t.nrec[idx];
        if (nrec == 0u) continue;
        const uint32_t *rec = t.records + static_cast<size_t>(idx) * t.stride;
        const uint32_t *longs = t.longs + static_cast<size_t>(idx) * t.lstride;
        uint32_t        nlong = 0, r0 = 0;
        while (r0 < nrec) {
            const bool     have = r0 + tid < nrec;
            const uint32_t v = have ? rec[r0 + tid] : 0u;
            const uint32_t off = v & 0x7FFu, lits = v >> 22;
            uint32_t       len = (v >> 11) & 0x7FFu;
            const bool     is_long = have && len == 0u && off != 0u;
            {
                const uint32_t lm = __ballot_sync(LZS_FULL_MASK, is_long);
                uint32_t       before, all;
                across_warps(static_cast<uint32_t>(__popc(lm)), before, all, 0);
                if (is_long) len = longs[nlong + before + static_cast<uint32_t>(__popc(lm & ((1u << lane) - 1u)))];
            }
            /* where every record's match starts: a scan of literals + lengths over the block */
            uint32_t incl = have ? lits + len : 0u;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t u = __shfl_up_sync(LZS_FULL_MASK, incl, static_cast<unsigned>(d));
                if (lane >= static_cast<uint32_t>(d)) incl += u;
            }
            {
                uint32_t before, all;
                across_warps(__shfl_sync(LZS_FULL_MASK, incl, 31), before, all, 1);
                incl += before;
            }
            /* records are taken while they fit the window; the first one always is */
            const bool     fits = have && (incl <= kDRingAhead || tid == 0u);
            uint32_t       nb;
            {
                const uint32_t fm = __ballot_sync(LZS_FULL_MASK, fits);
                /* fits is a prefix property (incl grows): count the leading ones over the block */
                uint32_t before, all;
                across_warps(static_cast<uint32_t>(__popc(fm)), before, all, 0);
                nb = all;
            }
            const bool     valid = tid < nb;
            {
                /* long lengths consumed by the records taken */
                const uint32_t lm = __ballot_sync(LZS_FULL_MASK, is_long && valid);
                uint32_t       before, all;
                across_warps(static_cast<uint32_t>(__popc(lm)), before, all, 1);
                nlong += all;
            }
            const uint32_t mstart = pos + incl - len;    /* first byte of my match */
            if (tid == nb - 1u) s_any = pos + incl;      /* where the batch ends */
            /* an offset that reaches before the output: the reference has a rule, k4_decode knows it */
            if (__syncthreads_or(valid && len != 0u && off > mstart)) {
                before_start = true;
                break;
            }
            const uint32_t batch_end = s_any;
            s_start[tid] = valid ? mstart : 0xFFFFFFFFu;
            s_mend[tid] = valid ? mstart + len : 0xFFFFFFFFu;
            if (batch_end > loaded) dring_refill(ring, dst, total, pos, flushed, loaded);
            else __syncthreads();
            if (batch_end > loaded) {
                /* one record that does not fit the window: in the slot itself, by the whole block
                 * (everything before it is in the slot after the refill) */
                const uint32_t m0 = s_start[0];
                const uint32_t b_len = batch_end - m0;
                const uint32_t b_off = rec[r0] & 0x7FFu;
                const uint8_t *src = dst + m0 - b_off;
                for (uint32_t i = tid; i < b_len; i += kT) dst[m0 + i] = src[i < b_off ? i : i % b_off];
                __syncthreads();
                flushed = batch_end;                     /* the window starts again behind the match */
                loaded = batch_end > kWindow ? batch_end - kWindow : 0u;
                for (uint32_t p = loaded + tid; p < batch_end; p += kT) ring[p & kMask] = dst[p];
                loaded = batch_end;
                __syncthreads();
                dring_refill(ring, dst, total, batch_end, flushed, loaded);
            } else {
                /* The matches my source bytes come from: records [qa, qb) of this batch (matches lie in
                 * the order of their records, so both ends are found by bisection); bytes before the
                 * batch, and literals, are final already. */
                const uint32_t src_lo = mstart - off, src_end = src_lo + umin32(len, off);
                bool           done = !valid || len == 0u;
                uint32_t       qa = 0, qb = 0;
                if (!done && src_end > pos) {
                    uint32_t lo = 0, hi = tid;               /* first record whose match ends behind src_lo */
                    while (lo < hi) {
                        const uint32_t mid = (lo + hi) >> 1;
                        if (s_mend[mid] > src_lo) hi = mid; else lo = mid + 1u;
                    }
                    qa = lo;
                    lo = qa;
                    hi = tid;                                /* first record whose match starts at or behind src_end */
                    while (lo < hi) {
                        const uint32_t mid = (lo + hi) >> 1;
                        if (s_start[mid] >= src_end) hi = mid; else lo = mid + 1u;
                    }
                    qb = lo;
                }
                uint32_t dep[kDCopyWarps];               /* the same as bits of the warps' done words */
#pragma unroll
                for (uint32_t w = 0; w < kDCopyWarps; w++) {
                    const uint32_t a = qa > w * 32u ? qa - w * 32u : 0u;
                    const uint32_t b = qb < w * 32u + 32u ? (qb > w * 32u ? qb - w * 32u : 0u) : 32u;
                    dep[w] = a < b ? (b == 32u ? 0xFFFFFFFFu : (1u << b) - 1u) & ~((1u << a) - 1u) : 0u;
                }
                /* one barrier per round: the done words alternate between two buffers, and the ring
                 * bytes a round writes are read after the next round's barrier at the earliest */
                for (int buf = 0;; buf ^= 1) {
                    const uint32_t dm = __ballot_sync(LZS_FULL_MASK, done);
                    if (lane == 0) s_w[buf][warp] = dm;
                    __syncthreads();
                    uint32_t undone = 0, waits = 0;
#pragma unroll
                    for (uint32_t w = 0; w < kDCopyWarps; w++) {
                        const uint32_t u = ~s_w[buf][w];
                        undone |= u;
                        waits |= u & dep[w];
                    }
                    if (undone == 0u) break;
                    const bool blocked = waits != 0u;
                    const bool     ready = !done && !blocked;
                    const uint32_t from0 = src_lo;
                    if (ready && len <= kDLaneLen) {
                        /* all source bytes lie before the match, so reads and writes of different
                         * bytes never meet: four at a time, the offset's period kept by a counter */
                        uint32_t f = 0;
                        for (uint32_t i0 = 0; i0 < len; i0 += 4u) {
                            uint8_t b[4];
#pragma unroll
                            for (uint32_t j = 0; j < 4u; j++) {
                                /* no byte behind the source range is touched: another thread may be writing it */
                                b[j] = i0 + j < len ? ring[(from0 + f) & kMask] : static_cast<uint8_t>(0);
                                f = f + 1u == off ? 0u : f + 1u;
                            }
#pragma unroll
                            for (uint32_t j = 0; j < 4u; j++)
                                if (i0 + j < len) ring[(mstart + i0 + j) & kMask] = b[j];
                        }
                    }
                    /* the longer ones of this round: the 32 lanes of the warp, one match after the other */
                    uint32_t bm = __ballot_sync(LZS_FULL_MASK, ready && len > kDLaneLen);
                    while (bm) {
                        const int      l = __ffs(static_cast<int>(bm)) - 1;
                        const uint32_t b_pos = __shfl_sync(LZS_FULL_MASK, mstart, l);
                        const uint32_t b_len = __shfl_sync(LZS_FULL_MASK, len, l);
                        const uint32_t b_off = __shfl_sync(LZS_FULL_MASK, off, l);
                        const uint32_t b_from = b_pos - b_off;
                        if (b_off >= b_len) {
                            for (uint32_t i = lane; i < b_len; i += 32u) ring[(b_pos + i) & kMask] = ring[(b_from + i) & kMask];
                        } else {
                            for (uint32_t i = lane; i < b_len; i += 32u) ring[(b_pos + i) & kMask] = ring[(b_from + i % b_off) & kMask];
                        }
                        bm &= bm - 1u;
                    }
                    done = done || ready;
                }
                __syncthreads();                         /* nobody still reads this batch's words when the next one writes them */
            }
            pos = batch_end;
            r0 += nb;
        }
    }
    __syncthreads();
    if (!before_start) {
        for (uint32_t p = flushed + tid; p < pos; p += kT) dst[p] = ring[p & kMask];
    } else if (tid == 0) {
        t.dirty[sid] = 1;
    }
}

/* ---------------------------------------------------------------- few streams: pointer jumping instead of the replay
 *
 * The replay above runs at the latency of one thread block per stream (~140 MB/s): fine for hundreds of
 * streams, hopeless for ONE (a single lzs_decompress call on a large buffer).  For that case the copies
 * are not replayed at all.  Every output byte gets a pointer to the byte it is a copy of -- itself for
 * a literal; pos - offset + (k mod offset) for byte k of a match, always an EARLIER byte -- and the
 * pointers are doubled, S[p] = S[S[p]], until every one of them points at a literal (log2 of the longest
 * chain of copies rounds; in place, since a pointer only ever moves towards its literal).  Then every
 * match byte is fetched from its literal.  O(n log n) work and 4 bytes of scratch per output byte,
 * but all of it parallel over the bytes of the stream.  Positions are offsets from `base` in `out`.
 */
constexpr int kJumpRounds = 32;                          /* chains are shorter than 2^32 */

__global__ void k4j_init(uint32_t *__restrict__ S, uint32_t span, uint32_t *__restrict__ flags)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= static_cast<uint32_t>(kJumpRounds)) flags[i] = i == 0u ? 1u : 0u;
    for (uint32_t p = i; p < span; p += gridDim.x * blockDim.x) S[p] = p;
}

/* One warp per piece: the pointers of its matches (32 records at a time, a scan gives their places;
 * one lane per match, the long ones by the whole warp). */
__global__ void __launch_bounds__(128)
k4j_fill(const uint64_t *__restrict__ out_off, uint64_t base, uint32_t n_streams, uint32_t *__restrict__ S, DPieceTable t)
{
    const uint32_t idx = blockIdx.x * 4u + (threadIdx.x >> 5);
    if (idx >= t.count[0]) return;
    const uint32_t nrec = t.nrec[idx];
    if (nrec == 0u) return;
    const uint32_t lane = lane_id();
    const uint32_t sid = dpiece_stream(t.first, n_streams, idx);
    if (t.dirty[sid]) return;
    const uint32_t o = static_cast<uint32_t>(out_off[sid] - base);
    const uint32_t *rec = t.records + static_cast<size_t>(idx) * t.stride;
    const uint32_t *longs = t.longs + static_cast<size_t>(idx) * t.lstride;
    uint32_t        pos = t.out_at[idx], nlong = 0;
    bool            before_start = false;
    for (uint32_t r0 = 0; r0 < nrec; r0 += 32u) {
        const bool     valid = r0 + lane < nrec;
        const uint32_t v = valid ? rec[r0 + lane] : 0u;
        const uint32_t off = v & 0x7FFu, lits = v >> 22;
        uint32_t       len = (v >> 11) & 0x7FFu;
        {
            const bool     is_long = valid && len == 0u && off != 0u;
            const uint32_t lm = __ballot_sync(LZS_FULL_MASK, is_long);
            if (is_long) len = longs[nlong + static_cast<uint32_t>(__popc(lm & ((1u << lane) - 1u)))];
            nlong += static_cast<uint32_t>(__popc(lm));
        }
        uint32_t incl = lits + len;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t u = __shfl_up_sync(LZS_FULL_MASK, incl, static_cast<unsigned>(d));
            if (lane >= static_cast<uint32_t>(d)) incl += u;
        }
        const uint32_t mstart = pos + incl - len;
        pos += __shfl_sync(LZS_FULL_MASK, incl, 31);
        if (__any_sync(LZS_FULL_MASK, valid && len != 0u && off > mstart)) {
            before_start = true;                         /* k4_decode knows the rule */
            break;
        }
        const uint32_t from0 = o + mstart - off, to0 = o + mstart;
        if (len <= 32u) {
            uint32_t f = 0;
            for (uint32_t i = 0; i < len; i++) {
                S[to0 + i] = from0 + f;
                f = f + 1u == off ? 0u : f + 1u;
            }
        }
        uint32_t bm = __ballot_sync(LZS_FULL_MASK, len > 32u);
        while (bm) {
            const int      l = __ffs(static_cast<int>(bm)) - 1;
            const uint32_t b_to = __shfl_sync(LZS_FULL_MASK, to0, l), b_from = __shfl_sync(LZS_FULL_MASK, from0, l);
            const uint32_t b_len = __shfl_sync(LZS_FULL_MASK, len, l), b_off = __shfl_sync(LZS_FULL_MASK, off, l);
            if (b_off >= b_len) {
                for (uint32_t i = lane; i < b_len; i += 32u) S[b_to + i] = b_from + i;
            } else {
                for (uint32_t i = lane; i < b_len; i += 32u) S[b_to + i] = b_from + i % b_off;
            }
            bm &= bm - 1u;
        }
    }
    if (before_start && lane == 0) t.dirty[sid] = 1;
}

/* One round of doubling; returns at once when the round before changed nothing. */
__global__ void k4j_jump(uint32_t *S, uint32_t span, uint32_t *flags, uint32_t round)
{
    if (flags[round] == 0u) return;
    bool changed = false;
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < span; p += gridDim.x * blockDim.x) {
        const uint32_t s = S[p];
        if (s != p) {
            const uint32_t s2 = S[s];
            if (s2 != s) {
                S[p] = s2;
                changed = true;
            }
        }
    }
    if (changed) flags[round + 1u] = 1u;
}

__global__ void k4j_gather(uint8_t *out, uint64_t base, const uint32_t *__restrict__ S, uint32_t span)
{
    uint8_t *o = out + base;
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < span; p += gridDim.x * blockDim.x) {
        const uint32_t s = S[p];
        if (s != p) o[p] = o[s];
    }
}

/* ---------------------------------------------------------------- dirty streams -> k4_decode */

__global__ void k4p_dirty_list(uint32_t n_streams, DPieceTable t)
{
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n_streams && t.dirty[s]) t.dirty_list[atomicAdd(&t.count[2], 1u)] = s;
}

}  // namespace lzs

#endif /* LZS_B200_K4_PIECES_CUH */
