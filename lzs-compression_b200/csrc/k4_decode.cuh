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
 * then at most one match or continuation token, and a step does kDecRounds of those (fewer
 * near the end of a stream) before it pays the fixed costs (votes, refill and flush checks) again.  The
 * compressed stream is staged in shared memory (64 words per stream, refilled 32
 * words at a time well ahead of the cursor), so a bit field at any position is two
 * shared-memory loads and a funnel shift.  Every collective uses
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
constexpr uint32_t kDecRing = 2048;                 /* history bytes per stream             */
constexpr uint32_t kDecInWords = 64;                /* staged input words per stream        */
constexpr uint32_t kDecStreamSmem = kDecRing + 4 * kDecInWords;
#ifndef LZS_K4_ROUNDS
#define LZS_K4_ROUNDS 8
#endif
constexpr int      kDecRounds = LZS_K4_ROUNDS;      /* (literal run + one token) per step   */
constexpr uint32_t kDecAhead = 24;                  /* words kept staged ahead of the cursor */
static_assert(kDecRounds * 80 + 64 <= 32 * static_cast<int>(kDecAhead), "a step may not outrun the staged input");

/* k mod m for 0 <= k < 16 and 1 <= m <= k (a match token carries at most 15 bytes): the
 * quotient is floor((k + 0.5) / m), at least 1/30 away from the next integer, so an
 * approximate reciprocal is exact here -- four predicated instructions instead of the ~18
 * and a branch of a general 32-bit remainder (LZS_K4_GENERIC_MOD restores the latter). */
__device__ __forceinline__ uint32_t small_mod(uint32_t k, uint32_t m)
{
#if defined(LZS_K4_GENERIC_MOD)
    return k % m;
#elif defined(LZS_SIMT_EMU)
    const float q = (static_cast<float>(k) + 0.5f) * (1.0f / static_cast<float>(m));
    return k - static_cast<uint32_t>(q) * m;
#else
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(static_cast<float>(m)));   /* one MUFU.RCP, no range fix-up */
    return k - __float2uint_rz((static_cast<float>(k) + 0.5f) * r) * m;
#endif
}

template <bool B>
struct CarefulTag { static constexpr bool value = B; };

template <int G>
constexpr size_t k4_smem_bytes() { return static_cast<size_t>(kDecThreads / G) * kDecStreamSmem; }

/* per-stream stop reasons, the values of LzsDecompressStatus_t (reference lzs.h:170-178) */
constexpr uint32_t kDecStarved = 0x01u, kDecEndMarker = 0x04u, kDecNoSpace = 0x08u;

template <int G>
__global__ void __launch_bounds__(kDecThreads)
k4_decode(const uint8_t *__restrict__ in, const uint64_t *__restrict__ in_off,
          const uint32_t *__restrict__ in_len, uint8_t *__restrict__ out,
          const uint64_t *__restrict__ out_off, const uint32_t *__restrict__ out_cap,
          uint32_t *__restrict__ out_len, uint32_t n_streams, uint32_t *__restrict__ next_stream,
          uint8_t *__restrict__ status, const uint32_t *__restrict__ order, const uint32_t *__restrict__ hist_len,
          const uint32_t *__restrict__ n_dev = nullptr)
{
    /* a list of streams made on the device (k4_pieces.cuh: the streams its passes left to this kernel):
     * `order` holds the list, *n_dev its length, n_streams only bounds the grid */
    if (n_dev != nullptr) n_streams = *n_dev;
    LZS_DYN_SMEM(uint8_t, smem);
    const uint32_t lane = lane_id();
    const uint32_t gl = lane % G;
    const uint32_t gshift = lane - gl;                  /* first lane of my group            */
    /* shared memory is addressed by offset from the block's base (cheap LDS/STS forms) */
    const uint32_t ring0 = (threadIdx.x / G) * kDecStreamSmem;
    uint32_t      *words = reinterpret_cast<uint32_t *>(smem);
    const uint32_t in0 = (ring0 + kDecRing) / 4;        /* word index of my input buffer     */
    constexpr int      kPass = (static_cast<int>(kMaxExtLen) + G - 1) / G;
    constexpr uint32_t kMaxLit = G < 7 ? G : 7;         /* literals per run: 9 * 7 <= 64 bits */
    constexpr uint32_t kGroupBits = (G == 32) ? 0xFFFFFFFFu : ((1u << (G & 31)) - 1u);
    constexpr uint32_t kFillPerLane = 32 / G;           /* one refill = 32 words per stream   */

    /* per-stream state; identical in all lanes of a group */
    bool     active = false, exhausted = false, ext = false, vec_ok = false;
    uint32_t sid = 0, cap = 0, pos = 0, flushed = 0, off = 1;
    int32_t  hist = 0;                                  /* kept history of the flow: bytes before dst that offsets may reach */
    uint32_t cur = 0, end = 0;                          /* bit cursor / end, from word 0      */
    uint32_t fill_hi = 0, nwords = 0, tail = 0;         /* words staged so far; stream extent */
    const uint32_t *wbase = nullptr;
    uint8_t *dst = nullptr;

    /* big-endian word w of the stream, zero past the end */
    auto fetch = [&](uint32_t w) -> uint32_t {
        if (w >= nwords) return 0u;
        uint32_t v = bswap32(__ldg(wbase + w));
        if (w == nwords - 1u && tail) v &= 0xFFFFFFFFu << (8u * (4u - tail));
        return v;
    };
    /* stage words [fill_hi, fill_hi + 32) of my stream (whole group) */
    auto refill = [&]() {
#pragma unroll
        for (uint32_t t = 0; t < kFillPerLane; t++) {
            const uint32_t w = fill_hi + gl * kFillPerLane + t;
            words[in0 + (w & (kDecInWords - 1u))] = fetch(w);
        }
        fill_hi += 32u;
    };
    /* 32 bits of the stream starting at bit position p */
    auto bits32 = [&](uint32_t p) -> uint32_t {
        const uint32_t w = p >> 5;
        return __funnelshift_l(words[in0 + ((w + 1u) & (kDecInWords - 1u))], words[in0 + (w & (kDecInWords - 1u))], p & 31u);
    };

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
                    sid = order ? order[s] : s;             /* streams in the order the launcher chose */
                    const uint8_t  *p = in + in_off[sid];
                    const uint32_t  nin = umin32(in_len[sid], 0x1FFFFF00u);   /* 32-bit bit positions */
                    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
                    const uint32_t  lead = static_cast<uint32_t>(a & 3u);
                    wbase = reinterpret_cast<const uint32_t *>(a & ~static_cast<uintptr_t>(3));
                    nwords = (lead + nin + 3u) >> 2;
                    tail = (lead + nin) & 3u;
                    dst = out + out_off[sid];
                    cap = out_cap[sid];
                    vec_ok = (reinterpret_cast<uintptr_t>(dst) & 15u) == 0;
                    cur = 8u * lead;                    /* bytes before the stream in word 0 */
                    end = cur + nin * 8u;
                    fill_hi = 0;
                    refill();
                    refill();
                    pos = 0;
                    flushed = 0;
                    off = 1;
                    ext = false;
                    active = true;
                    /* Flows with kept history: the decoder of the reference keeps its history across end
                     * markers (lzs-decompression.c:564-576), so a packet's offsets may reach into the
                     * flow's earlier packets -- the hist bytes in front of this packet's output, already
                     * decoded by an earlier launch.  They are loaded into the ring behind position 0. */
                    hist = hist_len != nullptr ? static_cast<int32_t>(umin32(hist_len[sid], kWindow)) : 0;
                    for (int32_t k = static_cast<int32_t>(gl); k < hist; k += G)
                        smem[ring0 + (static_cast<uint32_t>(k - hist) & (kDecRing - 1u))] = dst[k - hist];
                }
            }
            __syncwarp();
        }
        if (__all_sync(LZS_FULL_MASK, !active)) break;

        /* ---- keep kDecAhead words staged ahead of the cursor ---- */
        {
            const bool low = active && fill_hi < (cur >> 5) + kDecAhead;
            if (__any_sync(LZS_FULL_MASK, low)) {
                if (low) refill();
                __syncwarp();
            }
        }

        /* Far from the end of the input and of the output capacity nothing can run out
         * within one step, so the bound checks of the reference's loop are only compiled
         * into the "careful" variant used for the last steps of a stream. */
        bool       done = false;
        /* How many rounds can every stream of the warp run without any bound check?  A round
         * takes at most 80 bits and looks 64 + 32 bits ahead, and writes at most 22 bytes; the
         * limits below are a little wider. */
        uint32_t safe = static_cast<uint32_t>(kDecRounds);
        if (active) {
            const uint32_t bits = end - cur, room = cap - pos;
            const uint32_t by_in = bits >= 64u ? (bits - 64u) / 96u : 0u;
            const uint32_t by_out = room >= 8u ? (room - 8u) / 24u : 0u;
            safe = umin32(safe, umin32(by_in, by_out));
        }
        safe = __reduce_min_sync(LZS_FULL_MASK, safe);
        auto one_round = [&](auto careful_tag) {
            constexpr bool kCareful = decltype(careful_tag)::value;
            /* ---- a run of literals, one per lane (token gl starts 9*gl bits after the cursor) ---- */
            uint32_t nlit;
            {
                const uint32_t at = 9u * gl;
                const uint32_t field = bits32(cur + at) >> 23;                     /* 9 bits */
                bool           is_lit = active && !done && !ext && gl < kMaxLit && (field >> 8) == 0u;
                if (kCareful) is_lit = is_lit && (end - cur >= at + 9u) && (pos + gl < cap);
                const uint32_t mine = (__ballot_sync(LZS_FULL_MASK, is_lit) >> gshift) & kGroupBits;
                nlit = static_cast<uint32_t>(__ffs(static_cast<int>(~mine))) - 1u;  /* leading literals */
                if (gl < nlit) smem[ring0 + ((pos + gl) & (kDecRing - 1u))] = static_cast<uint8_t>(field);
            }
            pos += nlit;
            cur += 9u * nlit;
            __syncwarp();

            /* ---- at most one match / continuation token ---- */
            uint32_t L = 0;
            if (active && !done) {
                const uint32_t top = bits32(cur);
                const uint32_t left = kCareful ? end - cur : 0xFFFFu;
                uint32_t       used = 0;
                if (kCareful && (left == 0u || pos >= cap)) {
                    done = true;
                } else if (ext) {                       /* 4-bit continuation */
                    if (left < 4u) {
                        done = true;
                    } else {
                        used = 4u;
                        L = top >> 28;
                        ext = (L == kMaxExtLen);
                    }
                } else if ((top >> 31) == 0u) {         /* a literal the run could not take   */
                    if (left < 9u) done = true;         /* (8th in a row: next round)         */
                } else {
                    const uint32_t is_short = (top >> 30) & 1u;
                    const uint32_t hdr = is_short ? 9u : 13u;
                    const uint32_t o = is_short ? ((top >> 23) & 0x7Fu) : ((top >> 19) & 0x7FFu);
                    const uint32_t code = (top << hdr) >> 28;
                    uint32_t       w, len;
                    if (code < 12u) { len = (code >> 2) + 2u; w = 2u; }
                    else            { len = code - 7u;        w = 4u; }
                    if (left < hdr) {
                        done = true;
                    } else if (o == 0u) {
                        if (is_short) done = true;      /* end marker */
                        else used = 13u;                /* long offset 0: no length field */
                    } else if (left < hdr + w) {
                        done = true;
                    } else {
                        used = hdr + w;
                        L = len;
                        off = o;
                        ext = (len == kMaxShortLen);
                    }
                }
                if (kCareful) L = umin32(L, cap - pos);
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
                    if (kk >= off) kk = small_mod(kk, off);   /* overlap: periodic extension */
                    const int32_t s = static_cast<int32_t>(pos + kk) - static_cast<int32_t>(off);
                    if (s >= -hist) v[t] = smem[ring0 + (static_cast<uint32_t>(s) & (kDecRing - 1u))];
                }
            }
            __syncwarp();
#pragma unroll
            for (int t = 0; t < kPass; t++) {
                const uint32_t k = gl + static_cast<uint32_t>(t) * G;
                if (k < L) smem[ring0 + ((pos + k) & (kDecRing - 1u))] = v[t];
            }
            pos += L;
            __syncwarp();
        };
        /* Full steps while every stream of the warp is far from its end; as many rounds as are
         * still safe (one copy of the round in a loop) when one of them is closer -- with
         * 1500-byte packets that is most of the time -- and the careful variant, one round per
         * step, only for the last tokens of a stream. */
        if (safe >= static_cast<uint32_t>(kDecRounds)) {
#pragma unroll
            for (int round = 0; round < kDecRounds; round++) one_round(CarefulTag<false>{});
        } else if (safe != 0u) {
#pragma unroll 1
            for (uint32_t round = 0; round < safe; round++) one_round(CarefulTag<false>{});
        } else {
            one_round(CarefulTag<true>{});
        }

        /* flush once per step (a step adds at most kDecRounds * 22 bytes, far less than the ring) */
        while (active && pos - flushed >= 16u * G) {
            const uint32_t p = flushed + 16u * gl;
            if (vec_ok) {
                const uint4 q = *reinterpret_cast<const uint4 *>(smem + ring0 + (p & (kDecRing - 1u)));
                *reinterpret_cast<uint4 *>(dst + p) = q;
            } else {
                for (uint32_t b = 0; b < 16u; b++) dst[p + b] = smem[ring0 + ((p + b) & (kDecRing - 1u))];
            }
            flushed += 16u * G;
        }
        if (active && done) {
            for (uint32_t k = flushed + gl; k < pos; k += G) dst[k] = smem[ring0 + (k & (kDecRing - 1u))];
            if (gl == 0) {
                out_len[sid] = pos;
                if (status != nullptr) {
                    /* why the stream stopped, from where it stopped (a stop consumes nothing):
                     * capacity reached with input left / nothing left / an end marker / a token
                     * the remaining bits do not complete (include/lzs_b200.h) */
                    const uint32_t left = end - cur;
                    uint32_t       why = kDecStarved;
                    if (left != 0u && pos >= cap) why = kDecNoSpace;
                    else if (left >= 9u && !ext && (bits32(cur) >> 23) == 0x180u) why = kDecEndMarker;
                    status[sid] = static_cast<uint8_t>(why);
                }
            }
            active = false;
        }
    }
}

/* ---- launch order of the streams ----
 * The groups of a warp decode in lock step, so a warp costs what its slowest stream costs, and
 * the launch ends when the last stream does.  Streams of similar density (compressed bytes per
 * byte of capacity: mostly literals / mixed / mostly long matches) take similar numbers of
 * rounds, so the launcher hands them out bucket by bucket -- warps get streams of one kind --
 * in the order of rising density: the streams made of literals, which decode fastest, come last
 * and fill the tail of the launch.  A counting sort into 16 buckets: counters, then slots. */
constexpr uint32_t kDecBuckets = 16;
__device__ __forceinline__ uint32_t k4_bucket(uint32_t in_len, uint32_t cap)
{
    if (cap == 0u) return kDecBuckets - 1u;
    const uint64_t b = (static_cast<uint64_t>(in_len) * 12u) / cap;        /* 12 per unit of density: literals-only is 1.125 */
    return b < kDecBuckets - 1u ? static_cast<uint32_t>(b) : kDecBuckets - 1u;
}
__global__ void k4_order_count(const uint32_t *__restrict__ in_len, const uint32_t *__restrict__ out_cap, uint32_t n,
                               uint32_t *__restrict__ counts)
{
    __shared__ uint32_t s_c[kDecBuckets];
    if (threadIdx.x < kDecBuckets) s_c[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n) atomicAdd(&s_c[k4_bucket(in_len[s], out_cap[s])], 1u);
    __syncthreads();
    if (threadIdx.x < kDecBuckets && s_c[threadIdx.x]) atomicAdd(&counts[threadIdx.x], s_c[threadIdx.x]);
}
/* counts[0..16) = bucket sizes in, running cursors out (counts[16..32) is used for them) */
__global__ void k4_order_scatter(const uint32_t *__restrict__ in_len, const uint32_t *__restrict__ out_cap, uint32_t n,
                                 uint32_t *__restrict__ counts, uint32_t *__restrict__ order)
{
    __shared__ uint32_t s_base[kDecBuckets], s_c[kDecBuckets], s_at[kDecBuckets];
    if (threadIdx.x < kDecBuckets) s_c[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t       b = 0, rank = 0;
    if (s < n) {
        b = k4_bucket(in_len[s], out_cap[s]);
        rank = atomicAdd(&s_c[b], 1u);
    }
    __syncthreads();
    if (threadIdx.x < kDecBuckets) {
        uint32_t start = 0;
        for (uint32_t k = 0; k < threadIdx.x; k++) start += counts[k];
        s_base[threadIdx.x] = start;
        s_at[threadIdx.x] = s_c[threadIdx.x] ? atomicAdd(&counts[kDecBuckets + threadIdx.x], s_c[threadIdx.x]) : 0u;
    }
    __syncthreads();
    if (s < n) order[s_base[b] + s_at[b] + rank] = s;
}

}  // namespace lzs

#endif /* LZS_B200_K4_DECODE_CUH */
