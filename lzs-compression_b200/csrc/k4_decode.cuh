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
 * stream, so the lanes of a group run it redundantly (no divergence inside the
 * group) and split the back-reference copy.  History lives in a 2 KiB shared-
 * memory ring per stream (the window is 2047 bytes); output leaves the ring in
 * 16*G-byte blocks with 16-byte vector stores.  Groups are persistent and pull
 * stream indices from a global counter, which balances streams whose token
 * counts differ (incompressible vs. text).
 *
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
    const uint32_t gmask = (G == 32) ? LZS_FULL_MASK : (((1u << (G & 31)) - 1u) << (lane - gl));
    uint8_t *ring = smem + static_cast<size_t>(threadIdx.x / G) * kDecRing;

    for (;;) {
        uint32_t sid = 0;
        if (gl == 0) sid = atomicAdd(next_stream, 1u);
        sid = __shfl_sync(gmask, sid, 0, G);
        if (sid >= n_streams) break;

        const uint8_t *src = in + in_off[sid];
        const uint32_t nin = in_len[sid];
        uint8_t       *dst = out + out_off[sid];
        const uint32_t cap = out_cap[sid];
        const bool     vec_ok = (reinterpret_cast<uintptr_t>(dst) & 15u) == 0;

        /* aligned 32-bit view of the stream; `skip` leading bits belong to the
         * bytes before src inside the first aligned word */
        const uintptr_t a = reinterpret_cast<uintptr_t>(src);
        const uint32_t *wbase = reinterpret_cast<const uint32_t *>(a & ~static_cast<uintptr_t>(3));
        const uint32_t  lead = static_cast<uint32_t>(a & 3u);
        const uint32_t  nbytes = lead + nin;
        const uint32_t  nwords = (nbytes + 3u) >> 2;
        const uint32_t  tail = nbytes & 3u;            /* valid bytes in the last word, 0 = all */

        uint32_t wi = 0;
        auto fetch = [&](uint32_t w) -> uint32_t {
            if (w >= nwords) return 0u;
            uint32_t v = bswap32(__ldg(wbase + w));
            if (w == nwords - 1u && tail) v &= 0xFFFFFFFFu << (8u * (4u - tail));
            return v;                                   /* bits past the end read as zero */
        };
        uint32_t nextw = fetch(wi++);
        uint64_t win = 0;                               /* next bit is bit 63 */
        int      nb = 0;                                /* bits held in win   */
        uint64_t avail = static_cast<uint64_t>(nin) * 8u;   /* stream bits not yet consumed */

        win = static_cast<uint64_t>(nextw) << 32;
        nb = 32;
        nextw = fetch(wi++);
        win <<= 8u * lead;
        nb -= static_cast<int>(8u * lead);

        uint32_t pos = 0, flushed = 0, off = 0;
        bool     ext = false;

        for (;;) {
            if (avail == 0 || pos >= cap) break;
            if (nb <= 32) {
                win |= static_cast<uint64_t>(nextw) << (32 - nb);
                nb += 32;
                nextw = fetch(wi++);
            }
            const uint32_t top = static_cast<uint32_t>(win >> 32);
            uint32_t need, L, lit_byte = 0;
            bool     lit = false;
            if (!ext) {
                if ((top >> 31) == 0u) {                /* literal: 0 + 8 bits */
                    if (avail < 9u) break;
                    need = 9u;
                    lit_byte = (top >> 23) & 0xFFu;
                    L = 1u;
                    lit = true;
                } else {
                    const uint32_t is_short = (top >> 30) & 1u;
                    const uint32_t hdr = is_short ? 9u : 13u;
                    const uint32_t o = is_short ? ((top >> 23) & 0x7Fu) : ((top >> 19) & 0x7FFu);
                    if (avail < hdr) break;
                    if (o == 0u) {
                        if (is_short) break;            /* end marker */
                        win <<= 13;                     /* long offset 0: no length field */
                        nb -= 13;
                        avail -= 13u;
                        continue;
                    }
                    const uint32_t code = (top << hdr) >> 28;
                    uint32_t       w;
                    if (code < 12u) { L = (code >> 2) + 2u; w = 2u; }
                    else            { L = code - 7u;        w = 4u; }
                    need = hdr + w;
                    if (avail < need) break;
                    off = o;
                    ext = (L == kMaxShortLen);
                }
            } else {                                    /* 4-bit continuation */
                if (avail < 4u) break;
                L = top >> 28;
                need = 4u;
                ext = (L == kMaxExtLen);
            }
            win <<= need;
            nb -= static_cast<int>(need);
            avail -= need;

            L = umin32(L, cap - pos);
            if (lit) {
                if (gl == 0) ring[pos & (kDecRing - 1u)] = static_cast<uint8_t>(lit_byte);
            } else {
                /* The ring is one byte larger than the window, so the byte written for
                 * k+1 lands on the slot that k reads at offset 2047: read everything
                 * first, then write. */
                constexpr int kPass = (static_cast<int>(kMaxExtLen) + G - 1) / G;
                uint8_t       v[kPass];
#pragma unroll
                for (int t = 0; t < kPass; t++) {
                    const uint32_t k = gl + static_cast<uint32_t>(t) * G;
                    v[t] = 0;
                    if (k < L) {
                        uint32_t kk = k;
                        if (kk >= off) kk %= off;       /* overlap: periodic extension */
                        const int32_t s = static_cast<int32_t>(pos + kk) - static_cast<int32_t>(off);
                        if (s >= 0) v[t] = ring[static_cast<uint32_t>(s) & (kDecRing - 1u)];
                    }
                }
                __syncwarp(gmask);
#pragma unroll
                for (int t = 0; t < kPass; t++) {
                    const uint32_t k = gl + static_cast<uint32_t>(t) * G;
                    if (k < L) ring[(pos + k) & (kDecRing - 1u)] = v[t];
                }
            }
            pos += L;
            __syncwarp(gmask);

            if (pos - flushed >= 16u * G) {
                const uint32_t p = flushed + 16u * gl;
                if (vec_ok) {
                    const uint4 v = *reinterpret_cast<const uint4 *>(ring + (p & (kDecRing - 1u)));
                    *reinterpret_cast<uint4 *>(dst + p) = v;
                } else {
                    for (uint32_t b = 0; b < 16u; b++) dst[p + b] = ring[(p + b) & (kDecRing - 1u)];
                }
                flushed += 16u * G;
            }
        }
        for (uint32_t k = flushed + gl; k < pos; k += G) dst[k] = ring[k & (kDecRing - 1u)];
        if (gl == 0) out_len[sid] = pos;
        __syncwarp(gmask);
    }
}

}  // namespace lzs

#endif /* LZS_B200_K4_DECODE_CUH */
