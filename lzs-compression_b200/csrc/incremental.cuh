/*
 * incremental.cuh -- resumable (incremental) LZS compressor and decompressor as
 * device code: one warp advances one caller-owned stream state by one call.
 *
 * These restate, call for call, the observable behaviour of
 *   lzs_compress_incremental        c/src/liblzs/lzs-compression.c:553-823
 *   lzs_simple_compress_incremental c/src/liblzs/lzs-compression-simple.c:435-647
 *   lzs_decompress_incremental      c/src/liblzs/lzs-decompression.c:459-743
 * i.e. how much input is taken, how much output is produced, and the status flags
 * after every call, for any slicing of input and output -- so the concatenated
 * output is byte-identical to the single-call functions.
 *
 * A stream carried across calls is inherently serial, so the unit of parallelism is
 * the stream: the batch entry points run one warp per state block.  Inside a warp
 * the control flow is uniform; the match search -- which the reference does with a
 * hash chain (compression.c) or a linear scan (compression-simple.c), both
 * equivalent to "longest capped prefix over offsets 1..historyLen, nearest wins" --
 * is spread over the 32 lanes (offsets lane+1, lane+33, ...) and reduced with
 * __reduce_max_sync on (length, -offset).
 *
 * The state lives in the private part of the caller's parameter block (include/lzs.h);
 * the layouts below are this library's own and fit the reference's sizes.
 */
#ifndef LZS_B200_INCREMENTAL_CUH
#define LZS_B200_INCREMENTAL_CUH

#include "lzs_common.cuh"

namespace lzs {

constexpr uint32_t kIncCRing = 2062;     /* history + look-ahead, lzs.h:64          */
constexpr uint32_t kIncDRing = 2047;     /* lzs.h:66                                */
constexpr uint32_t kLookAhead = 15;      /* LZS_MAX_LOOK_AHEAD_LEN, lzs.h:57        */

/* status bits, identical for both directions (lzs.h:90-99, :170-178) */
enum : uint32_t {
    kStStarved = 0x01, kStFinished = 0x02, kStEndMarker = 0x04, kStNoSpace = 0x08, kStError = 0x10
};

struct IncCompressState {                /* 2078 bytes <= 2079 available            */
    uint32_t queue;                      /* bit queue, right aligned                */
    uint16_t latest;                     /* ring index of the next byte to encode   */
    uint16_t la_idx;                     /* ring index where the next input byte goes */
    uint16_t hist_len;
    uint16_t offset;                     /* offset of the match being extended      */
    uint8_t  la_len;                     /* bytes buffered ahead of `latest`        */
    uint8_t  qlen;                       /* bits in the queue                       */
    uint8_t  extended;                   /* 0 = normal, 1 = inside a long match     */
    uint8_t  reserved;
    uint8_t  ring[kIncCRing];
};

enum : uint8_t {                         /* decoder states, lzs-decompression.c:81-94 */
    kDCopy = 0, kDTokenType, kDLiteral, kDOffsetType, kDOffsetShort, kDOffsetLong, kDLength,
    kDCopyExt, kDExtLength
};

struct IncDecompressState {              /* 2063 bytes <= 2063 available            */
    uint32_t queue;                      /* bit queue, left aligned (next bit = 31) */
    uint16_t read_idx;
    uint16_t latest;
    uint16_t hist_len;
    uint16_t offset;
    uint8_t  qlen;
    uint8_t  length;
    uint8_t  state;
    uint8_t  reserved;
    uint8_t  ring[kIncDRing];
};

struct IncJob {
    void          *state;                /* device copy of the state block          */
    const uint8_t *in;
    uint8_t       *out;
    uint32_t       in_len, out_cap;
    uint32_t       in_used, out_used, status;
    uint32_t       add_end_marker;
};

__device__ __forceinline__ uint32_t ring_add(uint32_t idx, uint32_t inc, uint32_t size)
{
    idx += inc;
    return idx >= size ? idx - size : idx;
}

/* ------------------------------------------------------------------ compress */

/* Copy `bytes` bytes between a 4-byte aligned global address and shared memory (whole warp). */
__device__ __forceinline__ void inc_copy_ring(uint8_t *dst, const uint8_t *src, uint32_t bytes)
{
    const uint32_t lane = lane_id();
    const uint32_t words = bytes >> 2;
    if ((reinterpret_cast<uintptr_t>(src) & 3u) == 0u && (reinterpret_cast<uintptr_t>(dst) & 3u) == 0u) {
        for (uint32_t w = lane; w < words; w += 32u)
            reinterpret_cast<uint32_t *>(dst)[w] = reinterpret_cast<const uint32_t *>(src)[w];
        for (uint32_t b = 4u * words + lane; b < bytes; b += 32u) dst[b] = src[b];
    } else {
        for (uint32_t b = lane; b < bytes; b += 32u) dst[b] = src[b];
    }
    __syncwarp();
}

/* One call of lzs_compress_incremental; whole warp, uniform control flow.  `ring` is the warp's
 * shared-memory copy of the state's ring for the duration of the call (the search reads it
 * thousands of times per token): loaded on entry, written back on exit. */
__device__ inline void inc_compress_call(IncCompressState *S, IncJob *J, uint8_t *ring)
{
    const uint32_t lane = lane_id();
    const bool     finish = J->add_end_marker != 0;
    const uint8_t *in = J->in;
    uint8_t       *out = J->out;
    uint32_t       in_left = J->in_len, out_left = J->out_cap, in_pos = 0, out_pos = 0;
    uint32_t       status = 0;
    uint32_t       queue = S->queue, qlen = S->qlen, latest = S->latest, la_idx = S->la_idx;
    uint32_t       hist_len = S->hist_len, moff = S->offset, la_len = S->la_len, ext = S->extended;
    /* nothing offered, nothing asked for, nothing to flush: the reference's loop (:574-610) stops
     * at once with "finished | starved" -- no need to bring the ring in for that */
    if (in_left == 0 && !finish && qlen < 8u) {
        __syncwarp();
        if (lane == 0) { J->in_used = 0; J->out_used = 0; J->status = kStFinished | kStStarved; }
        return;
    }
    inc_copy_ring(ring, S->ring, kIncCRing);

    for (;;) {
        uint32_t length = 0;
        while (qlen >= 8u) {                             /* :574-588 */
            if (out_left == 0) { status |= kStNoSpace; break; }
            if (lane == 0) out[out_pos] = static_cast<uint8_t>(queue >> (qlen - 8u));
            out_pos++; out_left--; qlen -= 8u;
        }
        if (status != 0) break;                          /* :597-601 */
        if (in_left == 0) {                              /* :603-610 */
            status |= kStFinished | kStStarved;
            if (!finish) break;
        }
        {                                                /* top up the look-ahead, :613-635 */
            const uint32_t t = umin32(kLookAhead - la_len, in_left);
            if (lane < t) ring[ring_add(la_idx, lane, kIncCRing)] = in[in_pos + lane];
            la_idx = ring_add(la_idx, t, kIncCRing);
            la_len += t; in_left -= t; in_pos += t;
            __syncwarp();
        }
        if (!ext) {
            if (la_len < (finish ? 1u : kSearchMax)) {   /* :641-647 */
                status |= kStStarved;
            } else {
                const uint32_t M = umin32(la_len, kSearchMax);
                uint32_t       key = 0;
                if (M >= kMinLen) {
                    /* the window already in the ring, offsets lane+1, lane+33, ...: an offset whose
                     * first byte differs (nearly all of them) costs one load and one compare */
                    const uint32_t first = ring[latest];
                    for (uint32_t o = lane + 1u; o <= hist_len; o += 32u) {
                        const uint32_t from = ring_add(latest, kIncCRing - o, kIncCRing);
                        if (ring[from] != first) continue;
                        uint32_t l = 1;
                        while (l < M && ring[ring_add(latest, l, kIncCRing)] == ring[ring_add(from, l, kIncCRing)]) l++;
                        const uint32_t cand = (l << 12) | (4095u - o);
                        if (l >= kMinLen && cand > key) key = cand;
                    }
                }
                key = __reduce_max_sync(LZS_FULL_MASK, key);
                const uint32_t best = key >> 12, boff = 4095u - (key & 4095u);
                if (best < kMinLen) {                    /* literal, :695-706 */
                    queue = (queue << 9) | ring[latest];
                    qlen += 9u;
                    length = 1u;
                } else {                                 /* match, :707-747 */
                    if (boff <= kShortOffMax) { queue = (queue << 9) | 0x180u | boff;   qlen += 9u; }
                    else                      { queue = (queue << 13) | 0x1000u | boff; qlen += 13u; }
                    length = umin32(best, kMaxShortLen);
                    if (length <= 4u) { queue = (queue << 2) | (length - 2u);        qlen += 2u; }
                    else              { queue = (queue << 4) | (0xCu + length - 5u); qlen += 4u; }
                    if (length == kMaxShortLen) { moff = boff; ext = 1u; }
                }
            }
        } else {
            if (!finish && la_len < kMaxExtLen) {        /* :750-758 */
                status |= kStStarved;
            } else {                                     /* :760-773 */
                const uint32_t M = umin32(la_len, kMaxExtLen);
                const uint32_t from = ring_add(latest, kIncCRing - moff, kIncCRing);
                const bool     same = lane < M &&
                                  ring[ring_add(latest, lane, kIncCRing)] == ring[ring_add(from, lane, kIncCRing)];
                const uint32_t ball = __ballot_sync(LZS_FULL_MASK, same);
                length = static_cast<uint32_t>(__ffs(static_cast<int>(~ball)) - 1);
                queue = (queue << 4) | length;
                qlen += 4u;
                if (length != kMaxExtLen) ext = 0u;
            }
        }
        la_len -= length;                                /* :777-793 */
        latest = ring_add(latest, length, kIncCRing);
        hist_len = umin32(hist_len + length, kWindow);
    }

    if (finish && in_left == 0 && !ext && la_len == 0 && qlen < 8u && out_left >= (qlen + 16u) / 8u) {
        queue = (queue << 16) | (3u << 14);              /* :796-820 */
        qlen += 16u;
        while (qlen >= 8u) {
            if (lane == 0) out[out_pos] = static_cast<uint8_t>(queue >> (qlen - 8u));
            out_pos++; out_left--; qlen -= 8u;
        }
        qlen = 0;
        status |= kStEndMarker;
    }

    __syncwarp();
    if (in_pos != 0) inc_copy_ring(S->ring, ring, kIncCRing);   /* the ring only changes where input is appended */
    if (lane == 0) {
        S->queue = queue; S->qlen = static_cast<uint8_t>(qlen); S->latest = static_cast<uint16_t>(latest);
        S->la_idx = static_cast<uint16_t>(la_idx); S->hist_len = static_cast<uint16_t>(hist_len);
        S->offset = static_cast<uint16_t>(moff); S->la_len = static_cast<uint8_t>(la_len);
        S->extended = static_cast<uint8_t>(ext);
        J->in_used = in_pos; J->out_used = out_pos; J->status = status;
    }
}

/* ---------------------------------------------------------------- decompress */

/* One call of lzs_decompress_incremental.  The bit parse is a dependent chain, so
 * lane 0's view is authoritative; all lanes execute the same (uniform) path and the
 * copy states move up to 32 bytes per step when the offset allows it. */
__device__ inline void inc_decompress_call(IncDecompressState *S, IncJob *J)
{
    const uint32_t lane = lane_id();
    const uint8_t *in = J->in;
    uint8_t       *out = J->out;
    uint32_t       in_left = J->in_len, out_left = J->out_cap, in_pos = 0, out_pos = 0;
    uint32_t       status = 0;
    uint32_t       queue = S->queue, qlen = S->qlen, read_idx = S->read_idx, latest = S->latest;
    uint32_t       hist_len = S->hist_len, off = S->offset, length = S->length, state = S->state;
    uint8_t       *ring = S->ring;

    for (;;) {
        while (in_left > 0 && qlen <= 24u) {             /* :472-478 */
            queue |= static_cast<uint32_t>(in[in_pos]) << (24u - qlen);
            qlen += 8u; in_pos++; in_left--;
        }
        if (qlen == 0) status |= kStFinished | kStStarved;      /* :480-483 */
        uint32_t need;                                   /* StateBitMinimumWidth, :124-135 */
        switch (state) {
            case kDTokenType: case kDOffsetType: need = 1u; break;
            case kDLiteral:     need = 8u; break;
            case kDOffsetShort: need = 7u; break;
            case kDOffsetLong:  need = 11u; break;
            case kDExtLength:   need = 4u; break;
            default:            need = 0u; break;
        }
        if (qlen < need) status |= kStStarved;           /* :491-495 */
        if (status != 0) break;                          /* :498-502 */

        switch (state) {
            case kDTokenType:                            /* :507-519 */
                state = (queue >> 31) ? kDOffsetType : kDLiteral;
                queue <<= 1; qlen -= 1u;
                break;
            case kDLiteral:                              /* :521-548 */
                if (out_left == 0) {
                    status |= kStNoSpace;
                } else {
                    const uint32_t b = queue >> 24;
                    queue <<= 8; qlen -= 8u;
                    if (lane == 0) { out[out_pos] = static_cast<uint8_t>(b); ring[latest] = static_cast<uint8_t>(b); }
                    out_pos++; out_left--;
                    latest = ring_add(latest, 1u, kIncDRing);
                    hist_len = umin32(hist_len + 1u, kWindow);
                    state = kDTokenType;
                    __syncwarp();
                }
                break;
            case kDOffsetType:                           /* :550-557 */
                state = (queue >> 31) ? kDOffsetShort : kDOffsetLong;
                queue <<= 1; qlen -= 1u;
                break;
            case kDOffsetShort: {                        /* :559-583 */
                const uint32_t o = queue >> 25;
                queue <<= 7; qlen -= 7u;
                if (o == 0) {                            /* end marker: byte align, keep history */
                    const uint32_t pad = qlen & 7u;
                    queue <<= pad; qlen -= pad;
                    status |= kStEndMarker;
                    state = kDTokenType;
                } else {
                    off = o;
                    state = kDLength;
                }
                break;
            }
            case kDOffsetLong:                           /* :585-593 */
                off = queue >> 21;
                queue <<= 11; qlen -= 11u;
                state = kDLength;
                break;
            case kDLength: {                             /* :595-659 */
                const uint32_t code = queue >> 28;
                uint32_t       len, w;
                if (code < 12u) { len = (code >> 2) + 2u; w = 2u; }
                else            { len = code - 7u;        w = 4u; }
                length = len;                            /* the reference stores it before the check */
                if (qlen < w) {
                    status |= kStStarved;
                } else {
                    queue <<= w; qlen -= w;
                    state = (len == kMaxShortLen) ? kDCopyExt : kDCopy;
                    read_idx = latest < off ? latest + kIncDRing - off : latest - off;
                }
                break;
            }
            case kDCopy:
            case kDCopyExt:                              /* :661-711, byte at a time like the reference */
                for (;;) {
                    if (length == 0) { state = state + 1u; break; }
                    if (out_left == 0) { status |= kStNoSpace; break; }
                    if (lane == 0) {
                        const uint8_t b = (off <= hist_len) ? ring[read_idx] : 0;
                        out[out_pos] = b;
                        ring[latest] = b;
                    }
                    __syncwarp();
                    read_idx = ring_add(read_idx, 1u, kIncDRing);
                    out_pos++; out_left--; length--;
                    latest = ring_add(latest, 1u, kIncDRing);
                    hist_len = umin32(hist_len + 1u, kWindow);
                }
                break;
            case kDExtLength:                            /* :713-730 */
                length = queue >> 28;
                queue <<= 4; qlen -= 4u;
                state = (length == kMaxExtLen) ? kDCopyExt : kDCopy;
                break;
            default:
                state = kDTokenType;
                status |= kStError;
                break;
        }
    }

    __syncwarp();
    if (lane == 0) {
        S->queue = queue; S->qlen = static_cast<uint8_t>(qlen); S->read_idx = static_cast<uint16_t>(read_idx);
        S->latest = static_cast<uint16_t>(latest); S->hist_len = static_cast<uint16_t>(hist_len);
        S->offset = static_cast<uint16_t>(off); S->length = static_cast<uint8_t>(length);
        S->state = static_cast<uint8_t>(state);
        J->in_used = in_pos; J->out_used = out_pos; J->status = status;
    }
}

/* One warp per job. */
__global__ void __launch_bounds__(128)
kinc_compress(IncJob *jobs, uint32_t n)
{
    __shared__ __align__(16) uint8_t s_ring[4][(kIncCRing + 15) / 16 * 16];
    const uint32_t j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (j < n) inc_compress_call(static_cast<IncCompressState *>(jobs[j].state), &jobs[j], s_ring[threadIdx.x >> 5]);
}

/* Fresh states for n streams, `stride` bytes apart (device memory). */
__global__ void kinc_init(uint8_t *states, size_t stride, uint32_t n, int decompress)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    uint32_t *w = reinterpret_cast<uint32_t *>(states + static_cast<size_t>(j) * stride);
    w[0] = w[1] = w[2] = w[3] = 0;                       /* both headers are 16 bytes; the rings need no clearing */
    if (decompress) reinterpret_cast<IncDecompressState *>(w)->state = kDTokenType;
}

__global__ void __launch_bounds__(128)
kinc_decompress(IncJob *jobs, uint32_t n)
{
    const uint32_t j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (j < n) inc_decompress_call(static_cast<IncDecompressState *>(jobs[j].state), &jobs[j]);
}

}  // namespace lzs

#endif /* LZS_B200_INCREMENTAL_CUH */
