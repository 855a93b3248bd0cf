/*
 * corpus.h -- deterministic synthetic corpora for tests and bench.py
 * (SURVEY.md section 8d).  Integer arithmetic only, so the host (gcc) and device
 * (nvcc) builds of lzs_corpus_fill() produce identical bytes and any shard can be
 * generated in place on the GPU that will process it.
 *
 * Every stream (64 KiB chunk, packet, ...) is generated independently from
 * (seed, stream index), so stream boundaries never change the bytes.
 *
 *   TEXT    words drawn log-uniformly (Zipf-like, p(rank) ~ 1/rank) from a
 *           4096-word vocabulary of lowercase words of 2..12 letters, joined by
 *           " " (mostly), ", " or ".\n"
 *   BINARY  32-byte records: u32 LE running id, u32 LE small-range field,
 *           8 zero bytes, 16 bytes over a 4-symbol alphabet
 *   RANDOM  uniform bytes (incompressible)
 *   MIXED   kind = stream index mod 3 (TEXT, BINARY, RANDOM)
 *   PACKET  IPComp/PPP-like: a text header of 64..319 bytes, then a payload
 *           whose kind is stream index mod 3
 */
#ifndef LZS_B200_CORPUS_H
#define LZS_B200_CORPUS_H

#include <stdint.h>

#ifdef __CUDACC__
#define LZS_HD __host__ __device__
#else
#define LZS_HD
#endif

enum {
    LZS_CORPUS_TEXT   = 0,
    LZS_CORPUS_BINARY = 1,
    LZS_CORPUS_RANDOM = 2,
    LZS_CORPUS_MIXED  = 3,
    LZS_CORPUS_PACKET = 4
};

static inline LZS_HD uint64_t lzs_splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

static inline LZS_HD uint64_t lzs_xorshift64s(uint64_t *s)
{
    uint64_t x = *s;
    x ^= x >> 12;
    x ^= x << 25;
    x ^= x >> 27;
    *s = x;
    return x * 0x2545F4914F6CDD1Dull;
}

/* one vocabulary word: length and letters are pure functions of the rank */
static inline LZS_HD uint32_t lzs_corpus_word(uint32_t rank, uint8_t *dst, uint32_t room)
{
    uint64_t h = lzs_splitmix64(0xC0FFEEull + rank);
    uint32_t wl = 2u + (uint32_t)(h % 11u);
    uint32_t k;
    h = lzs_splitmix64(h);
    for (k = 0; k < wl && k < room; k++) {
        dst[k] = (uint8_t)('a' + (uint32_t)(h % 26u));
        h /= 26u;
        if (k == 9u) h = lzs_splitmix64(h + rank);
    }
    return k;
}

static inline LZS_HD uint32_t lzs_corpus_text(uint8_t *dst, uint32_t len, uint64_t *rng)
{
    uint32_t p = 0;
    while (p < len) {
        uint64_t r = lzs_xorshift64s(rng) >> 16;
        uint32_t oct = (uint32_t)(r % 12u);                    /* octave 0..11   */
        uint32_t rank = ((1u << oct) - 1u) + ((uint32_t)(r >> 8) & ((1u << oct) - 1u));
        uint32_t sep = (uint32_t)(r >> 40) & 15u;
        p += lzs_corpus_word(rank, dst + p, len - p);
        if (p < len) {
            if (sep == 0u) {
                dst[p++] = '.';
                if (p < len) dst[p++] = '\n';
            } else if (sep <= 2u) {
                dst[p++] = ',';
                if (p < len) dst[p++] = ' ';
            } else {
                dst[p++] = ' ';
            }
        }
    }
    return p;
}

static inline LZS_HD void lzs_corpus_binary(uint8_t *dst, uint32_t len, uint64_t *rng, uint32_t id0)
{
    uint32_t p = 0, id = id0;
    while (p < len) {
        uint8_t  rec[32];
        uint64_t r = lzs_xorshift64s(rng);
        uint64_t a = lzs_xorshift64s(rng);
        uint32_t f = (uint32_t)(r >> 20) % 1000u;
        uint32_t k;
        rec[0] = (uint8_t)id; rec[1] = (uint8_t)(id >> 8);
        rec[2] = (uint8_t)(id >> 16); rec[3] = (uint8_t)(id >> 24);
        rec[4] = (uint8_t)f; rec[5] = (uint8_t)(f >> 8); rec[6] = 0; rec[7] = 0;
        for (k = 8; k < 16; k++) rec[k] = 0;
        for (k = 16; k < 32; k++) {
            rec[k] = (uint8_t)(0x40u + 0x11u * (uint32_t)(a & 3u));
            a >>= 2;
        }
        for (k = 0; k < 32u && p < len; k++) dst[p++] = rec[k];
        id++;
    }
}

static inline LZS_HD void lzs_corpus_random(uint8_t *dst, uint32_t len, uint64_t *rng)
{
    uint32_t p = 0;
    while (p < len) {
        uint64_t r = lzs_xorshift64s(rng);
        uint32_t k;
        for (k = 0; k < 8u && p < len; k++) {
            dst[p++] = (uint8_t)r;
            r >>= 8;
        }
    }
}

/* Fill one stream of `len` bytes.  `index` is the global stream index. */
static inline LZS_HD void lzs_corpus_fill(uint8_t *dst, uint32_t len, uint64_t seed,
                                          uint64_t index, int kind)
{
    uint64_t rng = lzs_splitmix64(seed ^ (index * 0xD1B54A32D192ED03ull));
    if (rng == 0) rng = 0x5EED5EED5EEDull;
    if (kind == LZS_CORPUS_MIXED) kind = (int)(index % 3u);
    if (kind == LZS_CORPUS_PACKET) {
        uint32_t hdr = 64u + (uint32_t)(lzs_xorshift64s(&rng) & 255u);
        if (hdr > len) hdr = len;
        lzs_corpus_text(dst, hdr, &rng);
        dst += hdr;
        len -= hdr;
        kind = (int)(index % 3u);
    }
    if (kind == LZS_CORPUS_TEXT) {
        lzs_corpus_text(dst, len, &rng);
    } else if (kind == LZS_CORPUS_BINARY) {
        lzs_corpus_binary(dst, len, &rng, (uint32_t)(index * 2048u));
    } else {
        lzs_corpus_random(dst, len, &rng);
    }
}

#endif /* LZS_B200_CORPUS_H */
