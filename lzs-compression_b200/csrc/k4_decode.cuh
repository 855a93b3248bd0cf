/*
 * k4_decode.cuh -- K4, the LZS stream decoder for sm_100a.
 *
 * Replaces the serial loop of lzs_decompress, c/src/liblzs/lzs-decompression.c:156-412,
 * with identical results on well-formed AND malformed streams:
 *   - stops at the first end marker (:255-261), when the remaining bits cannot
 *     hold the next field (:220,:238,:248,:272,:332,:373), or when the output
 *     capacity is reached (:200-203, :361-364);
 *   - a long offset of 0 consumes 13 bits and nothing else (:280);
 *   - offsets that reach before the start of the output produce zero bytes (:346-357).
 *
 * Mapping: G lanes of a warp decode one stream (G = 4/8/16/32), so one warp
 * instruction serves 32/G streams.  The token parse is inherently serial per
 * stream, so the lanes of a group run it redundantly and split the back-reference
 * copy.  All 32/G groups of a warp advance in lock step under warp-uniform control
 * flow; one step takes a run of up to 7 literals (lane g looks at the token that
 * would start 9*g bits after the cursor; a ballot gives the length of the run) and
 * then at most one match or continuation token.  The bit stream is read through a
 * 4-word register cache (96 bits visible from the cursor).  Every collective uses
 * the full mask (a
 * variable member mask makes nvcc emit MATCH.ANY, ~50-390 cycles on sm_100a), and
 * a group whose stream ended simply idles until it has fetched the next stream
 * index from a global counter (which also balances incompressible vs. text chunks).
 *
 * History lives in a 2 KiB shared-memory ring per stream (the window is 2047
 * bytes); output leaves the ring in 16*G-byte blocks with 16-byte vector stores.
 * Overlapping copies (offset < length) need no serialisation: byte k of a match
 * equals history byte (k mod offset), which was written by an earlier token.
 *
 * HBM traffic per stream: c compressed bytes read + n bytes written (the
 * algorithmic minimum).
 */
#ifndef LZS_B200_K4_DECODE_CUH
#define LZS_B200_K4_DECODE_CUH

#include "lzs_common.cuh"

namespace lzs {

constexpr int      kDecThreads = 128;
constexpr uint32_t kDecRing = 2048;

template <int G>
constexpr size_t k4_smem_bytes() { return static_cast<size_t>(kDecThreads / G) * kDecRing; }

/* One compressed stream seen as aligned big-endian 32-bit words. */
struct DecInput {
    const uint32_t *wbase;
    uint32_t        nwords;
    uint32_t        tail;     /* valid bytes in the last word, 0 = all four */
    __device__ __forceinline__ uint32_t fetch(uint32_t w) const
    {
        if (w >= nwords) return 0u;
        uint32_t v = bswap32(__ldg(wbase + w));
        if (w == nwords - 1u && tail) v &= 0xFFFFFFFFu << (8u * (4u - tail));
        return v;                                       /* bits past the end read as zero */
    }
};

template <int G>
__global__ void __launch_bounds__(kDecThreads)
k4_decode(const uint8_t *__restrict__ in, const uint64_t *__restrict__ in_off,
          const uint32_t *__restrict__ in_len, uint8_t *__restrict__ out,
          const uint64_t *__restrict__ out_off, const uint32_t *__restrict__ out_cap,
          uint32_t *__restrict__ out_len, uint32_t n_streams, uint32_t *__restrict__ next_stream)
{
    LZS_DYN_SMEM(uint8_t, smem);
    const uint32_t lane = lane_id();
    const uint32_t gl = lane % G;
    const uint32_t gshift = lane - gl;                  /* first lane of my group            */
    uint8_t *ring = smem + static_cast<size_t>(threadIdx.x / G) * kDecRing;
    constexpr int      kPass = (static_cast<int>(kMaxExtLen) + G - 1) / G;
    constexpr uint32_t kMaxLit = G < 7 ? G : 7;         /* literals per step: 9 * 7 <= 64 bits */
    constexpr uint32_t kGroupBits = (G == 32) ? 0xFFFFFFFFu : ((1u << (G & 31)) - 1u);

    /* per-stream state; identical in all lanes of a group */
    bool     active = false, exhausted = false, ext = false, vec_ok = false;
    uint32_t sid = 0, cap = 0, pos = 0, flushed = 0, off = 1;
    uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0, wbase = 0; /* stream words [wbase, wbase+4)     */
    uint64_t cur = 0, end = 0;                          /* bit cursor / end, from word 0      */
    uint8_t *dst = nullptr;
    DecInput src{nullptr, 0, 0};

    for (;;) {
        /* ---- idle groups fetch the next stream ---- */
        const bool want = !active && !exhausted;
        if (__any_sync(LZS_FULL_MASK, want)) {
            uint32_t s = 0;
            if (want && gl == 0) s = atomicAdd(next_stream, 1u);
            s = __shfl_sync(LZS_FULL_MASK, s, 0, G);
            if (want) {
                if (s >= n_streams) {
                    exhausted = true;
                } else {
                    sid = s;
                    const uint8_t  *p = in + in_off[sid];
                    const uint32_t  nin = in_len[sid];
                    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
                    const uint32_t  lead = static_cast<uint32_t>(a & 3u);
                    src.wbase = reinterpret_cast<const uint32_t *>(a & ~static_cast<uintptr_t>(3));
                    src.nwords = (lead + nin + 3u) >> 2;
                    src.tail = (lead + nin) & 3u;
                    dst = out + out_off[sid];
                    cap = out_cap[sid];
                    vec_ok = (reinterpret_cast<uintptr_t>(dst) & 15u) == 0;
                    cur = 8u * lead;                    /* bytes before the stream in word 0 */
                    end = cur + static_cast<uint64_t>(nin) * 8u;
                    wbase = 0;
                    w0 = src.fetch(0); w1 = src.fetch(1); w2 = src.fetch(2); w3 = src.fetch(3);
                    pos = 0;
                    flushed = 0;
                    off = 1;
                    ext = false;
                    active = true;
                }
            }
        }
        if (__all_sync(LZS_FULL_MASK, !active)) break;

        /* ---- 96 bits of the stream starting at the cursor ---- */
        {
            const uint32_t wi = static_cast<uint32_t>(cur >> 5);
            while (active && wbase < wi) {
                w0 = w1; w1 = w2; w2 = w3;
                w3 = src.fetch(wbase + 4u);
                wbase++;
            }
        }
        const uint32_t sh = static_cast<uint32_t>(cur) & 31u;
        const uint32_t b0 = __funnelshift_l(w1, w0, sh);
        const uint32_t b1 = __funnelshift_l(w2, w1, sh);
        const uint32_t b2 = __funnelshift_l(w3, w2, sh);
        const uint64_t left = end - cur;
        uint32_t       avail = left > 0xFFFFu ? 0xFFFFu : static_cast<uint32_t>(left);

        /* ---- phase A: a run of literals, one per lane (token gl starts at bit 9*gl) ---- */
        uint32_t nlit = 0;
        {
            const uint32_t at = 9u * gl;                /* < 64 for gl < 7                    */
            const uint64_t b01 = (static_cast<uint64_t>(b0) << 32) | b1;
            const uint32_t field = static_cast<uint32_t>((b01 << (at & 63u)) >> 55);   /* 9 bits */
            const bool     is_lit = active && !ext && gl < kMaxLit && (field >> 8) == 0u &&
                                avail >= at + 9u && pos + gl < cap;
            const uint32_t mine = (__ballot_sync(LZS_FULL_MASK, is_lit) >> gshift) & kGroupBits;
            nlit = static_cast<uint32_t>(__ffs(static_cast<int>(~mine))) - 1u;          /* leading literals */
            if (gl < nlit) ring[(pos + gl) & (kDecRing - 1u)] = static_cast<uint8_t>(field);
        }
        pos += nlit;
        avail -= 9u * nlit;
        uint32_t used = 9u * nlit;                      /* bits consumed this step            */
        __syncwarp();

        /* ---- phase B: at most one match / continuation token ---- */
        bool     done = false;
        uint32_t L = 0;
        if (active) {
            const uint32_t top = used < 32u ? __funnelshift_l(b1, b0, used) : __funnelshift_l(b2, b1, used - 32u);
            if (avail == 0u || pos >= cap) {
                done = true;
            } else if (ext) {                           /* 4-bit continuation */
                if (avail < 4u) {
                    done = true;
                } else {
                    used += 4u;
                    L = top >> 28;
                    ext = (L == kMaxExtLen);
                }
            } else if ((top >> 31) == 0u) {             /* a literal phase A could not take   */
                if (avail < 9u) done = true;            /* (8th in a row: next step)          */
            } else {
                const uint32_t is_short = (top >> 30) & 1u;
                const uint32_t hdr = is_short ? 9u : 13u;
                const uint32_t o = is_short ? ((top >> 23) & 0x7Fu) : ((top >> 19) & 0x7FFu);
                const uint32_t code = (top << hdr) >> 28;
                uint32_t       w, len;
                if (code < 12u) { len = (code >> 2) + 2u; w = 2u; }
                else            { len = code - 7u;        w = 4u; }
                if (avail < hdr) {
                    done = true;
                } else if (o == 0u) {
                    if (is_short) done = true;          /* end marker */
                    else used += 13u;                   /* long offset 0: no length field */
                } else if (avail < hdr + w) {
                    done = true;
                } else {
                    used += hdr + w;
                    L = len;
                    off = o;
                    ext = (len == kMaxShortLen);
                }
            }
            L = umin32(L, cap - pos);
            cur += used;
        }

        /* copy: read everything, then write (the ring is one byte larger than the window,
         * so byte k+1 lands on the slot byte k reads at offset 2047) */
        uint8_t v[kPass];
#pragma unroll
        for (int t = 0; t < kPass; t++) {
            const uint32_t k = gl + static_cast<uint32_t>(t) * G;
            v[t] = 0;
            if (k < L) {
                uint32_t kk = k;
                if (kk >= off) kk %= off;               /* overlap: periodic extension */
                const int32_t s = static_cast<int32_t>(pos + kk) - static_cast<int32_t>(off);
                if (s >= 0) v[t] = ring[static_cast<uint32_t>(s) & (kDecRing - 1u)];
            }
        }
        __syncwarp();
#pragma unroll
        for (int t = 0; t < kPass; t++) {
            const uint32_t k = gl + static_cast<uint32_t>(t) * G;
            if (k < L) ring[(pos + k) & (kDecRing - 1u)] = v[t];
        }
        pos += L;
        __syncwarp();

        if (active && pos - flushed >= 16u * G) {
            const uint32_t p = flushed + 16u * gl;
            if (vec_ok) {
                const uint4 q = *reinterpret_cast<const uint4 *>(ring + (p & (kDecRing - 1u)));
                *reinterpret_cast<uint4 *>(dst + p) = q;
            } else {
                for (uint32_t b = 0; b < 16u; b++) dst[p + b] = ring[(p + b) & (kDecRing - 1u)];
            }
            flushed += 16u * G;
        }
        if (active && done) {
            for (uint32_t k = flushed + gl; k < pos; k += G) dst[k] = ring[k & (kDecRing - 1u)];
            if (gl == 0) out_len[sid] = pos;
            active = false;
        }
    }
}

}  // namespace lzs

#endif /* LZS_B200_K4_DECODE_CUH */
