/*
 * lzs_oracle.c -- CPU restatement of the LZS codec, used ONLY as a checker.
 *
 * TEST INFRASTRUCTURE.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may call it.  The shipped library (lzs-compression_b200/) never links or
 * loads this file and has no CPU fallback.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement against
 *   - the reference's golden vector (c/src/test/test-lzs-decompression.c:34-96,
 *     extracted into tests/golden/ by tests/golden/make_golden.py),
 *   - the two known-answer size laws of c/src/test/test-lzs.c:93-167,
 *   - byte-for-byte against the unmodified reference compiled into oracle/_ref/
 *     (oracle/Makefile) on seeded corpora, when that build is present, and
 *   - committed outputs of that same reference build (tests/golden/ref_*.bin).
 *
 * The compressor is written as the *specification* of the reference's match
 * rule rather than as its hash-chain loop.  For position i of an n-byte buffer:
 *     H = min(i, 2047)                          c/src/liblzs/lzs-compression.c:447
 *     M = min(n - i, 12)                        lzs-compression.c:62, :325
 *     over offsets o = 1..H, len(o) = common prefix of in[i..] and in[i-o..],
 *     capped at M; keep the longest, ties go to the smallest o
 *                                               lzs-compression.c:334-361 (strict '>'
 *                                               and nearest-first order), identical to
 *                                               the brute-force loop in
 *                                               lzs-compression-simple.c:264-278
 *     len < 2  -> literal  (0 + 8 bits)         lzs-compression.c:365-375
 *     else     -> 1, then 1+7-bit or 0+11-bit offset, then the length code of
 *                 min(len, 8); if that is 8 the match continues at the same offset
 *                 in 4-bit steps of up to 15 bytes until a step is short
 *                                               lzs-compression.c:381-431, tables :100-124
 *     after the last byte: end marker 110000000, zero padded to a byte
 *                                               lzs-compression.c:449-466
 * The hash (lzs-compression.c:139-142) is not normative: it only prunes
 * candidates that cannot reach length 2, so the brute-force rule gives the same
 * bytes.
 */
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#define LZS_WINDOW        2047u   /* c/src/liblzs/lzs.h:60                    */
#define LZS_SEARCH_MAX    12u     /* lzs-compression.c:62                     */
#define LZS_SHORT_OFF_MAX 127u    /* lzs-common.h:43                          */
#define LZS_MIN_LEN       2u      /* lzs-common.h:51                          */
#define LZS_MAX_SHORT_LEN 8u      /* lzs-common.h:52                          */
#define LZS_MAX_EXT_LEN   15u     /* lzs-common.h:53                          */

/* ------------------------------------------------------------------ bit sink */

typedef struct {
    uint8_t *out;
    size_t   cap;
    size_t   count;      /* bytes that fitted                                  */
    uint64_t acc;        /* pending bits, right aligned                        */
    unsigned nacc;       /* number of pending bits (< 8 between calls)         */
    int      full;       /* set once a byte did not fit                        */
} bitsink_t;

/* Append nbits (<= 32) MSB-first; flush whole bytes.  Mirrors the truncation
 * rule of lzs-compression.c:304-313: once the buffer is full the function
 * returns what fitted, a plain prefix of the full stream. */
static void sink_put(bitsink_t *s, uint32_t value, unsigned nbits)
{
    s->acc = (s->acc << nbits) | (uint64_t)value;
    s->nacc += nbits;
    while (s->nacc >= 8u) {
        uint8_t b = (uint8_t)(s->acc >> (s->nacc - 8u));
        s->nacc -= 8u;
        if (s->count < s->cap) {
            s->out[s->count++] = b;
        } else {
            s->full = 1;
        }
    }
    s->acc &= ((uint64_t)1 << s->nacc) - 1u;
}

/* ------------------------------------------------------------ match finding */

/* Capped common-prefix length; lzs-compression.c:178-191. */
static unsigned prefix_len(const uint8_t *a, const uint8_t *b, unsigned cap)
{
    unsigned l = 0;
    while (l < cap && a[l] == b[l]) {
        l++;
    }
    return l;
}

/*
 * Best (length, offset) at position i by the normative rule above.
 * Returns the capped length (0 or 1 mean "literal"); *off is set when >= 2.
 */
unsigned lzs_oracle_best_match(const uint8_t *in, size_t n, size_t i, unsigned *off)
{
    size_t   H = i < LZS_WINDOW ? i : LZS_WINDOW;
    size_t   rem = n - i;
    unsigned M = rem < LZS_SEARCH_MAX ? (unsigned)rem : LZS_SEARCH_MAX;
    unsigned best = 0, best_off = 0;
    size_t   o;

    if (M < LZS_MIN_LEN) {
        *off = 0;
        return 0;
    }
    for (o = 1; o <= H; o++) {
        unsigned l;
        /* A candidate only replaces `best` when strictly longer (ties keep the
         * nearer one), so byte number `best` must match; best < M here. */
        if (best > 0 && in[i - o + best] != in[i + best]) continue;
        l = prefix_len(in + i, in + i - o, M);
        if (l > best) {
            best = l;
            best_off = (unsigned)o;
            if (l >= M) break;                  /* lzs-compression.c:341-344 */
        }
    }
    *off = best_off;
    return best;
}

/* Fill len12[i] (0..12) and off[i] for every position: the per-position table
 * the GPU match finder (kernel K1) must reproduce. */
void lzs_oracle_all_matches(const uint8_t *in, size_t n, uint8_t *len12, uint16_t *off)
{
    size_t i;
    for (i = 0; i < n; i++) {
        unsigned o = 0;
        unsigned l = lzs_oracle_best_match(in, n, i, &o);
        if (l < LZS_MIN_LEN) { l = 0; o = 0; }
        len12[i] = (uint8_t)l;
        off[i] = (uint16_t)o;
    }
}

/* ----------------------------------------------------------------- compress */

/* Length code for 2..8; lzs-compression.c:91-124. */
static void put_short_length(bitsink_t *s, unsigned len)
{
    if (len <= 4u) {
        sink_put(s, len - 2u, 2u);
    } else {
        sink_put(s, 0xCu + (len - 5u), 4u);
    }
}

size_t lzs_oracle_compress(uint8_t *out, size_t out_cap, const uint8_t *in, size_t n)
{
    bitsink_t s;
    size_t    i = 0;

    memset(&s, 0, sizeof s);
    s.out = out;
    s.cap = out_cap;

    while (i < n) {
        unsigned off = 0;
        unsigned len = lzs_oracle_best_match(in, n, i, &off);

        if (len < LZS_MIN_LEN) {
            sink_put(&s, in[i], 9u);            /* '0' + byte, :370-373       */
            i += 1;
            continue;
        }
        if (off <= LZS_SHORT_OFF_MAX) {
            sink_put(&s, 0x180u | off, 9u);     /* '1','1', 7-bit  :381-393   */
        } else {
            sink_put(&s, 0x1000u | off, 13u);   /* '1','0', 11-bit :396-402   */
        }
        if (len > LZS_MAX_SHORT_LEN) {
            len = LZS_MAX_SHORT_LEN;            /* :404                       */
        }
        put_short_length(&s, len);
        i += len;
        if (len == LZS_MAX_SHORT_LEN) {
            /* extended state, :417-431: 4-bit steps until one is short */
            for (;;) {
                size_t   rem = n - i;
                unsigned cap = rem < LZS_MAX_EXT_LEN ? (unsigned)rem : LZS_MAX_EXT_LEN;
                unsigned step = prefix_len(in + i, in + i - off, cap);
                sink_put(&s, step, 4u);
                i += step;
                if (step != LZS_MAX_EXT_LEN) break;
            }
        }
    }
    /* end marker + pad: 9 marker bits followed by 7 zero bits, then only whole
     * bytes leave the queue (:452-465) */
    sink_put(&s, 0x180u, 9u);
    sink_put(&s, 0u, 7u);
    return s.count;
}

/* --------------------------------------------------------------- decompress */

/*
 * Restatement of lzs_decompress, c/src/liblzs/lzs-decompression.c:156-412.
 * The reference keeps a 32-bit queue that it tops up while <= 24 bits are held
 * (:181-187); a token never needs more than 17 bits, so "not enough bits in the
 * queue" (:220,:238,:248,:272,:332,:373) only ever happens when the input
 * itself is exhausted.  That lets the rule be stated on the whole bit string:
 * stop as soon as the next field does not fit in the bits that remain; bits
 * past the end read as zero for the 4-bit length-table peek (:325-327).
 */
typedef struct {
    const uint8_t *in;
    size_t         nbits;   /* total bits in the stream                       */
    size_t         pos;     /* bits consumed                                  */
} bitsrc_t;

static uint32_t src_peek(const bitsrc_t *b, unsigned nbits)
{
    uint32_t v = 0;
    unsigned k;
    for (k = 0; k < nbits; k++) {
        size_t p = b->pos + k;
        unsigned bit = 0;
        if (p < b->nbits) {
            bit = (b->in[p >> 3] >> (7u - (p & 7u))) & 1u;
        }
        v = (v << 1) | bit;
    }
    return v;
}

static size_t src_left(const bitsrc_t *b) { return b->nbits - b->pos; }

/* The same loop also reports which of its exits was taken, in the values of
 * LzsDecompressStatus_t (lzs.h:170-178): 0x04 end marker (:255-261), 0x08 output full with input
 * left (:200-203, :361-364), 0x01 input ended first (every "goto finish" on missing bits, and
 * :189-192).  When the output fills up exactly as the input ends, the input check comes first,
 * as at the top of the reference's loop. */
size_t lzs_oracle_decompress_status(uint8_t *out, size_t out_cap, const uint8_t *in, size_t in_len, int *why)
{
    bitsrc_t b;
    size_t   n = 0;
    unsigned offset = 0;
    int      extended = 0;

    b.in = in;
    b.nbits = in_len * 8u;
    b.pos = 0;

    for (;;) {
        unsigned length, k;

        *why = 0x01;                            /* unless said otherwise below */
        if (src_left(&b) == 0) break;           /* :189-192                   */
        if (n >= out_cap) { *why = 0x08; break; }   /* :200-203               */

        if (!extended) {
            unsigned type = src_peek(&b, 1);
            b.pos += 1;
            if (type == 0) {                    /* literal, :217-233          */
                if (src_left(&b) < 8u) break;
                out[n++] = (uint8_t)src_peek(&b, 8);
                b.pos += 8;
                continue;
            }
            if (src_left(&b) < 1u) break;       /* :238-241                   */
            type = src_peek(&b, 1);
            b.pos += 1;
            if (type) {                         /* short offset, :245-268     */
                if (src_left(&b) < 7u) break;
                offset = src_peek(&b, 7);
                b.pos += 7;
                if (offset == 0) { *why = 0x04; break; }   /* end marker, :255-261 */
            } else {                            /* long offset, :269-279      */
                if (src_left(&b) < 11u) break;
                offset = src_peek(&b, 11);
                b.pos += 11;
                if (offset == 0) continue;      /* :280: no length field read */
            }
            {                                   /* length code, :323-343      */
                unsigned code = src_peek(&b, 4);
                unsigned width;
                if (code < 0xCu) { length = (code >> 2) + 2u; width = 2u; }
                else             { length = code - 0xCu + 5u; width = 4u; }
                if (src_left(&b) < width) break;
                b.pos += width;
                if (length == LZS_MAX_SHORT_LEN) extended = 1;
            }
        } else {                                /* :370-406                   */
            if (src_left(&b) < 4u) break;
            length = src_peek(&b, 4);
            b.pos += 4;
            if (length != LZS_MAX_EXT_LEN) extended = 0;
        }
        /* copy, one byte at a time so overlaps repeat; offsets that reach
         * before the start of the output give zero bytes (:346-365) */
        for (k = 0; k < length; k++) {
            out[n] = (n >= offset) ? out[n - offset] : 0;
            n++;
            if (n >= out_cap) {                 /* :361-364                   */
                *why = src_left(&b) == 0 ? 0x01 : 0x08;
                return n;
            }
        }
    }
    return n;
}

size_t lzs_oracle_decompress(uint8_t *out, size_t out_cap, const uint8_t *in, size_t in_len)
{
    int why;
    return lzs_oracle_decompress_status(out, out_cap, in, in_len, &why);
}

/* Upper bound used by callers for the output buffer; lzs.h:77. */
size_t lzs_oracle_compressed_max(size_t n) { return n + (n + 7u) / 8u + 3u; }
