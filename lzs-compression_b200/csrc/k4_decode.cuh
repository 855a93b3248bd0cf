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
 * copy.  All 32/G groups of a warp advance in lock step, one token per iteration,
 * under warp-uniform control flow: every collective uses the full mask (a
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
    uint8_t *ring = smem + static_cast<size_t>(threadIdx.x / G) * kDecRing;
    constexpr int kPass = (static_cast<int>(kMaxExtLen) + G - 1) / G;

    /* per-stream state; identical in all lanes of a group */
    bool     active = false, exhausted = false, ext = false, vec_ok = false;
    uint32_t sid = 0, cap = 0, pos = 0, flushed = 0, off = 1, wi = 0, nextw = 0;
    int      nb = 0;
    uint64_t win = 0, avail = 0;
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
                    avail = static_cast<uint64_t>(nin) * 8u;
                    /* the `lead` bytes before the stream inside its first word are skipped */
                    win = static_cast<uint64_t>(src.fetch(0)) << (32u + 8u * lead);
                    nb = 32 - static_cast<int>(8u * lead);
                    nextw = src.fetch(1);
                    wi = 2;
                    pos = 0;
                    flushed = 0;
                    off = 1;
                    ext = false;
                    active = true;
                }
            }
        }
        if (__all_sync(LZS_FULL_MASK, !active)) break;

        /* ---- one token per active group ---- */
        bool     done = false, lit = false;
        uint32_t L = 0, lit_byte = 0;
        if (active) {
            if (avail == 0 || pos >= cap) {
                done = true;
            } else {
                if (nb <= 32) {
                    win |= static_cast<uint64_t>(nextw) << (32 - nb);
                    nb += 32;
                    nextw = src.fetch(wi++);
                }
                const uint32_t top = static_cast<uint32_t>(win >> 32);
                uint32_t       need = 0;
                if (!ext) {
                    if ((top >> 31) == 0u) {            /* literal: 0 + 8 bits */
                        need = 9u;
                        lit_byte = (top >> 23) & 0xFFu;
                        L = 1u;
                        lit = true;
                        if (avail < need) done = true;
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
                            if (is_short) done = true;  /* end marker */
                            else need = 13u;            /* long offset 0: no length field */
                        } else if (avail < hdr + w) {
                            done = true;
                        } else {
                            need = hdr + w;
                            L = len;
                            off = o;
                            ext = (len == kMaxShortLen);
                        }
                    }
                } else {                                /* 4-bit continuation */
                    if (avail < 4u) {
                        done = true;
                    } else {
                        need = 4u;
                        L = top >> 28;
                        ext = (L == kMaxExtLen);
                    }
                }
                if (done) {
                    L = 0;
                } else {
                    win <<= need;
                    nb -= static_cast<int>(need);
                    avail -= need;
                    L = umin32(L, cap - pos);
                }
            }
        }

        /* ---- copy: read everything, then write (the ring is one byte larger than the
         * window, so byte k+1 lands on the slot byte k reads at offset 2047) ---- */
        uint8_t v[kPass];
#pragma unroll
        for (int t = 0; t < kPass; t++) {
            const uint32_t k = gl + static_cast<uint32_t>(t) * G;
            v[t] = static_cast<uint8_t>(lit_byte);
            if (!lit && k < L) {
                uint32_t kk = k;
                if (kk >= off) kk %= off;               /* overlap: periodic extension */
                const int32_t s = static_cast<int32_t>(pos + kk) - static_cast<int32_t>(off);
                v[t] = (s >= 0) ? ring[static_cast<uint32_t>(s) & (kDecRing - 1u)] : 0;
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
