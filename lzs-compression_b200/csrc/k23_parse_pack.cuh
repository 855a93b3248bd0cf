/*
 * k23_parse_pack.cuh -- K2 (greedy parse) and K3 (bit packer) for sm_100a, one
 * warp per stream, launched as one kernel because the parse already holds every
 * token in registers when the packer needs it.
 *
 * K2 restates the reference's token loop, c/src/liblzs/lzs-compression.c:301-447:
 * starting at position 0, a position with no match (K1 length < 2) emits a
 * literal and advances by 1; otherwise a match token is emitted with initial
 * length min(len, 8) (:404) and, when that is 8, the match continues at the SAME
 * offset (:417-431) for as long as the bytes agree -- total length
 * L = min(uncapped common prefix, n - i) -- and the parse advances by L.
 *   The chain "next(i) = i + advance(i)" is resolved 32 positions at a time:
 *   every lane knows its own advance from K1's table, reachability from the
 *   entry lane is closed by pointer doubling (5 rounds of shuffle +
 *   __reduce_or_sync), and a long match ends the group so that its true length
 *   can be measured by the whole warp (32 bytes per ballot).
 *
 * K3 restates the bit layout (:365-409, :423-425, :449-466): literal 0+8 bits;
 * match 1, then 1+7-bit or 0+11-bit offset, then 00/01/10 for 2/3/4,
 * 1100/1101/1110 for 5/6/7, 1111 + nibbles for >= 8 where the nibbles are
 * floor((L-8)/15) times 1111 followed by (L-8) mod 15; end marker 110000000 and
 * zero padding to a byte.  A warp scan of token bit lengths gives every token
 * its bit offset; tokens are OR-ed MSB-first into a 64-word shared staging ring
 * and leave as byte-swapped 32-bit words, 128 bytes per flush.  If the output
 * capacity is too small the result is the prefix that fits (:306-309).
 *
 * HBM traffic per stream: n bytes + 2n bytes of K1 records read, c bytes written.
 */
#ifndef LZS_B200_K23_PARSE_PACK_CUH
#define LZS_B200_K23_PARSE_PACK_CUH

#include "lzs_common.cuh"

namespace lzs {

constexpr int kK2Threads = 128;
constexpr int kK2Warps = kK2Threads / 32;

struct BitStage {
    uint32_t *buf;        /* 64 words of shared memory, zero where no bit was put yet */
    uint8_t  *dst;
    uint32_t  cap;        /* output capacity in bytes                                */
    uint32_t  cur;        /* bits pending in buf                                     */
    uint32_t  wdone;      /* 32-bit words already written to dst                     */
    bool      aligned;
};

/* OR `nb` (<= 32) bits of `val` into the stage at bit offset `b`, MSB first. */
__device__ __forceinline__ void stage_put(uint32_t *buf, uint32_t b, uint32_t val, uint32_t nb)
{
    const uint32_t wi = b >> 5, sh = b & 31u;
    const uint64_t v = static_cast<uint64_t>(val) << (64u - nb - sh);
    atomicOr(&buf[wi], static_cast<uint32_t>(v >> 32));
    const uint32_t lo = static_cast<uint32_t>(v);
    if (lo) atomicOr(&buf[wi + 1], lo);
}

__device__ __forceinline__ void stage_store_word(const BitStage &s, uint32_t widx, uint32_t w)
{
    const uint32_t bi = widx * 4u;
    if (s.aligned && bi + 4u <= s.cap) {
        *reinterpret_cast<uint32_t *>(s.dst + bi) = bswap32(w);
    } else {
        for (uint32_t b = 0; b < 4u; b++)
            if (bi + b < s.cap) s.dst[bi + b] = static_cast<uint8_t>(w >> (24u - 8u * b));
    }
}

/* Whole warp: if 32 or more words are pending, write the first 32 (128 bytes). */
__device__ __forceinline__ void stage_flush_if_full(BitStage &s)
{
    if (s.cur >= 1024u) {
        const uint32_t lane = lane_id();
        const uint32_t w = s.buf[lane];
        const uint32_t hi = s.buf[lane + 32];
        stage_store_word(s, s.wdone + lane, w);
        __syncwarp();
        s.buf[lane] = hi;
        s.buf[lane + 32] = 0;
        __syncwarp();
        s.wdone += 32;
        s.cur -= 1024u;
    }
}

/* Whole warp emits one field that every lane agrees on. */
__device__ __forceinline__ void stage_emit_uniform(BitStage &s, uint32_t val, uint32_t nb)
{
    if (lane_id() == 0) stage_put(s.buf, s.cur, val, nb);
    s.cur += nb;
    __syncwarp();
    stage_flush_if_full(s);
}

__global__ void __launch_bounds__(kK2Threads)
k23_parse_pack(const uint8_t *__restrict__ in, const uint64_t *__restrict__ in_off,
               const uint32_t *__restrict__ in_len, const match_t *__restrict__ matches,
               uint8_t *__restrict__ out, const uint64_t *__restrict__ out_off,
               const uint32_t *__restrict__ out_cap, uint32_t *__restrict__ out_len,
               uint32_t n_streams)
{
    __shared__ uint32_t s_buf[kK2Warps][64];
    const uint32_t lane = lane_id();
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t sid = blockIdx.x * kK2Warps + warp;
    if (sid >= n_streams) return;

    const uint32_t  n = in_len[sid];
    const uint8_t  *src = in + in_off[sid];
    const match_t  *m = matches + in_off[sid];

    BitStage s;
    s.buf = s_buf[warp];
    s.dst = out + out_off[sid];
    s.cap = out_cap[sid];
    s.cur = 0;
    s.wdone = 0;
    s.aligned = (reinterpret_cast<uintptr_t>(s.dst) & 3u) == 0;
    s.buf[lane] = 0;
    s.buf[lane + 32] = 0;
    __syncwarp();

    uint32_t pos = 0;
    while (pos < n) {
        const uint32_t i = pos + lane;
        const bool     valid = i < n;
        const uint32_t mv = valid ? m[i] : 0u;
        const uint32_t byte = valid ? src[i] : 0u;
        const uint32_t len = mv >> kMatchOffBits;
        const uint32_t off = mv & ((1u << kMatchOffBits) - 1u);
        /* K1 caps lengths at 12: 8..11 are exact (one 4-bit continuation, known here);
         * only 12 means "12 or more" and needs the bytes compared further */
        const bool     is_long = len >= kSearchMax;
        const uint32_t nxt = lane + (len >= kMinLen ? len : 1u);

        /* K2: token starts reachable from lane 0 inside this group */
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
        const uint32_t nvalid = umin32(32u, n - pos);
        if (nvalid < 32u) reach &= (1u << nvalid) - 1u;
        const bool tok = (reach >> lane) & 1u;

        /* K3: this lane's token */
        uint32_t val = 0, nb = 0;
        if (tok) {
            if (len < kMinLen) {
                val = byte;                                   /* 0 + 8 bits */
                nb = 9u;
            } else {
                if (off <= kShortOffMax) { val = 0x180u | off;  nb = 9u; }
                else                     { val = 0x1000u | off; nb = 13u; }
                if (len <= 4u) {
                    val = (val << 2) | (len - 2u);
                    nb += 2u;
                } else if (len < kMaxShortLen) {
                    val = (val << 4) | (0xCu + len - 5u);
                    nb += 4u;
                } else if (!is_long) {                        /* 8..11: 1111 + (len - 8) */
                    val = (val << 8) | 0xF0u | (len - kMaxShortLen);
                    nb += 8u;
                } else {                                      /* >= 12: 1111, nibbles follow below */
                    val = (val << 4) | 0xFu;
                    nb += 4u;
                }
            }
        }
        uint32_t incl = nb;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(LZS_FULL_MASK, incl, static_cast<unsigned>(d));
            if (lane >= static_cast<uint32_t>(d)) incl += t;
        }
        const uint32_t total = __shfl_sync(LZS_FULL_MASK, incl, 31);
        if (tok) stage_put(s.buf, s.cur + incl - nb, val, nb);
        s.cur += total;
        __syncwarp();

        const int      last = 31 - __clz(static_cast<int>(reach));
        const uint32_t last_long = __shfl_sync(LZS_FULL_MASK, is_long ? 1u : 0u, last);
        uint32_t       next_pos = pos + __shfl_sync(LZS_FULL_MASK, nxt, last);
        stage_flush_if_full(s);

        if (last_long) {
            /* true length of the long match that ended the group */
            const uint32_t p = pos + static_cast<uint32_t>(last);
            const uint32_t loff = __shfl_sync(LZS_FULL_MASK, off, last);
            uint32_t       L = kSearchMax;                    /* the first 12 bytes are known to match */
            const uint8_t *from = src - loff;                 /* may point into the flow's kept history, in front of the packet */
            for (;;) {
                const uint32_t idx = p + L + lane;
                const bool     same = (idx < n) && (src[idx] == from[idx]);
                const uint32_t ball = __ballot_sync(LZS_FULL_MASK, same);
                if (ball == LZS_FULL_MASK) {
                    L += 32u;
                } else {
                    L += static_cast<uint32_t>(__ffs(static_cast<int>(~ball)) - 1);
                    break;
                }
            }
            const uint32_t e = L - kMaxShortLen;
            uint32_t       q = e / kMaxExtLen;
            const uint32_t r = e - q * kMaxExtLen;
            while (q >= 8u) {
                stage_emit_uniform(s, 0xFFFFFFFFu, 32u);
                q -= 8u;
            }
            stage_emit_uniform(s, (((1u << (4u * q)) - 1u) << 4) | r, 4u * q + 4u);
            next_pos = p + L;
        }
        pos = next_pos;
    }

    stage_emit_uniform(s, 0x180u, 9u);                        /* end marker */
    const uint32_t rest = (s.cur + 7u) >> 3;                  /* pad to a byte */
    const uint32_t base = s.wdone * 4u;
    for (uint32_t k = lane; k < rest; k += 32u) {
        const uint32_t b = (s.buf[k >> 2] >> (24u - 8u * (k & 3u))) & 0xFFu;
        if (base + k < s.cap) s.dst[base + k] = static_cast<uint8_t>(b);
    }
    if (lane == 0) out_len[sid] = umin32(base + rest, s.cap);
}

}  // namespace lzs

#endif /* LZS_B200_K23_PARSE_PACK_CUH */
