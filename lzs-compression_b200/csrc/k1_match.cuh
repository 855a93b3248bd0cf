/*
 * k1_match.cuh -- K1, the all-positions LZS match finder for sm_100a.
 *
 * For EVERY position i of every stream it produces the (length, offset) the
 * reference's search would choose there (c/src/liblzs/lzs-compression.c:322-363,
 * equivalently the brute-force loop of lzs-compression-simple.c:264-278):
 *     M = min(n - i, 12), window H = min(i, 2047);
 *     longest common prefix capped at M over offsets 1..H, ties -> smallest offset;
 *     shorter than 2 -> "no match".
 *
 * The reference walks one hash chain of 2-byte prefixes and compares every
 * candidate, which is data dependent and pathological on runs.  This kernel
 * restates the rule so that its cost does not depend on the data:
 *
 *   For k = 2..12 let P_k(i) be the nearest earlier position (inside the window)
 *   whose next k bytes equal the k bytes at i.  "Some offset reaches length >= k"
 *   is monotone in k, so   best length = max { k <= M : P_k(i) exists }
 *   and the winning offset is  i - P_best(i)  (nearest candidate of that length,
 *   exactly the reference's strict '>' / nearest-first tie-break).
 *
 *   P_k is "previous equal element" over the sequence of k-grams.  One warp per
 *   level k keeps a 4096-slot last-occurrence table in shared memory and inserts
 *   positions in order, 32 at a time: __match_any_sync finds predecessors inside
 *   the batch, the table gives the predecessor from earlier batches.  Each
 *   position stores the distance to its predecessor in the same SLOT; slots are
 *   hashes, so a query verifies bytes and, on a foreign entry, follows the
 *   distance chain (every in-window position of the slot is on it, nearest
 *   first).  Table and chain garbage (stale entries, aliasing) can only produce
 *   candidates that fail the byte check, never a wrong answer -- the same
 *   argument that lets the reference run on uninitialised tables
 *   (lzs-compression.c:253-254).
 *
 *   A query probes level 2 first (incompressible data stops there), then
 *   bisects the remaining levels; a verified candidate of length l found at
 *   level k is also the nearest candidate at level l, so the search jumps.
 *
 * Layout: one persistent CTA per SM (192 KiB of shared memory: 11 head tables,
 * 11 link rings, a ring of 4-byte grams), streams pulled from a global counter,
 * each stream walked in 1024-position tiles.  Output: one uint16 per input byte,
 * (len << 11) | offset, consumed by the parse kernel.
 *
 * HBM traffic per input byte: 1 B read + 2 B written (intermediate).
 */
#ifndef LZS_B200_K1_MATCH_CUH
#define LZS_B200_K1_MATCH_CUH

#include "lzs_common.cuh"

namespace lzs {

constexpr int      kK1Levels = 11;          /* k = 2 .. 12 */
constexpr uint32_t kK1Slots = 4096;
constexpr uint32_t kK1Ring = 4096;          /* >= 2047 + tile + 8 */
constexpr uint32_t kK1Tile = 1024;
constexpr int      kK1Threads = 512;
constexpr size_t   kK1SmemBytes =
    static_cast<size_t>(kK1Levels) * kK1Slots * 2 + static_cast<size_t>(kK1Levels) * kK1Ring * 2 + kK1Ring * 4;

/* mask selecting the low `bytes` bytes of a little-endian word, bytes in 1..4 */
__device__ __forceinline__ constexpr uint32_t low_bytes_mask(int bytes)
{
    return bytes >= 4 ? 0xFFFFFFFFu : ((1u << (8 * bytes)) - 1u);
}

/* table slot of the K-gram whose bytes are the first K bytes of (w0, w1, w2) */
template <int K>
__device__ __forceinline__ uint32_t gram_slot(uint32_t w0, uint32_t w1, uint32_t w2)
{
    uint32_t h;
    if (K <= 4) {
        h = (w0 & low_bytes_mask(K)) * 0x9E3779B1u;
    } else if (K <= 8) {
        h = w0 * 0x9E3779B1u;
        h = (h ^ (w1 & low_bytes_mask(K - 4))) * 0x85EBCA77u;
    } else {
        h = w0 * 0x9E3779B1u;
        h = (h ^ w1) * 0x85EBCA77u;
        h = (h ^ (w2 & low_bytes_mask(K - 8))) * 0xC2B2AE3Du;
    }
    h ^= h >> 15;
    h *= 0x27D4EB2Fu;
    return h >> 20;                          /* 12 bits */
}

/* common prefix length (0..12) of two 12-byte strings given as LE words */
__device__ __forceinline__ uint32_t lcp12(uint32_t a0, uint32_t a1, uint32_t a2,
                                          uint32_t b0, uint32_t b1, uint32_t b2)
{
    const uint32_t x0 = a0 ^ b0, x1 = a1 ^ b1, x2 = a2 ^ b2;
    if (x0) return static_cast<uint32_t>(__ffs(static_cast<int>(x0)) - 1) >> 3;
    if (x1) return 4u + (static_cast<uint32_t>(__ffs(static_cast<int>(x1)) - 1) >> 3);
    if (x2) return 8u + (static_cast<uint32_t>(__ffs(static_cast<int>(x2)) - 1) >> 3);
    return 12u;
}

/* Insert the positions of one tile into level K's table, in order, and record
 * for each the distance to the previous position of the same slot (0 = none in
 * the window).  Executed by one whole warp. */
template <int K>
__device__ __forceinline__ void k1_build_level(uint16_t *heads, uint16_t *links, const uint32_t *W,
                                               uint32_t t0, uint32_t tile_n, uint32_t epoch)
{
    const uint32_t lane = lane_id();
    uint16_t      *hd = heads + (K - 2) * kK1Slots;
    uint16_t      *lk = links + (K - 2) * kK1Ring;
    for (uint32_t b = 0; b < tile_n; b += 32) {
        const uint32_t i = t0 + b + lane;
        const bool     act = (b + lane) < tile_n;
        const uint32_t w0 = W[i & (kK1Ring - 1)];
        const uint32_t w1 = W[(i + 4) & (kK1Ring - 1)];
        const uint32_t w2 = W[(i + 8) & (kK1Ring - 1)];
        const uint32_t slot = act ? gram_slot<K>(w0, w1, w2) : (0x10000u | lane);
        const uint32_t grp = __match_any_sync(LZS_FULL_MASK, slot);
        const uint32_t lower = grp & ((1u << lane) - 1u);
        const uint32_t pos16 = (epoch + i) & 0xFFFFu;
        if (act) {
            uint32_t dist;
            if (lower) dist = lane - (31u - static_cast<uint32_t>(__clz(static_cast<int>(lower))));
            else       dist = (pos16 - hd[slot]) & 0xFFFFu;
            if (dist > umin32(kWindow, i)) dist = 0;
            lk[i & (kK1Ring - 1)] = static_cast<uint16_t>(dist);
        }
        __syncwarp();                        /* all table reads before any insert */
        if (act && (grp >> lane) == 1u) hd[slot] = static_cast<uint16_t>(pos16);
        __syncwarp();
    }
}

/* Walk level k's chain from position i: first in-window candidate whose common
 * prefix (capped at M) reaches k.  Returns that prefix length, or 0. */
__device__ __forceinline__ uint32_t k1_probe(const uint16_t *links, const uint32_t *W, uint32_t k,
                                             uint32_t i, uint32_t maxd, uint32_t M, uint32_t w0,
                                             uint32_t w1, uint32_t w2, uint32_t &dist_out)
{
    const uint16_t *lk = links + (k - 2) * kK1Ring;
    uint32_t        tot = 0;
    uint32_t        d = lk[i & (kK1Ring - 1)];
    while (d != 0) {
        tot += d;
        if (tot > maxd) break;
        const uint32_t j = i - tot;
        const uint32_t l = umin32(lcp12(w0, w1, w2, W[j & (kK1Ring - 1)], W[(j + 4) & (kK1Ring - 1)],
                                        W[(j + 8) & (kK1Ring - 1)]), M);
        if (l >= k) {
            dist_out = tot;
            return l;
        }
        d = lk[j & (kK1Ring - 1)];
    }
    return 0;
}

__device__ __forceinline__ uint32_t k1_query(const uint16_t *links, const uint32_t *W, uint32_t i, uint32_t n)
{
    const uint32_t M = umin32(kSearchMax, n - i);
    const uint32_t maxd = umin32(kWindow, i);
    if (M < kMinLen || maxd == 0) return 0;
    const uint32_t w0 = W[i & (kK1Ring - 1)];
    const uint32_t w1 = W[(i + 4) & (kK1Ring - 1)];
    const uint32_t w2 = W[(i + 8) & (kK1Ring - 1)];
    uint32_t bd = 0;
    uint32_t best = k1_probe(links, W, 2, i, maxd, M, w0, w1, w2, bd);
    if (best == 0) return 0;
    uint32_t lo = best + 1, hi = M;
    while (lo <= hi) {
        const uint32_t mid = (lo + hi) >> 1;
        uint32_t       dd = 0;
        const uint32_t l = k1_probe(links, W, mid, i, maxd, M, w0, w1, w2, dd);
        if (l) { best = l; bd = dd; lo = l + 1; }
        else   { hi = mid - 1; }
    }
    return (best << kMatchOffBits) | bd;
}

__global__ void __launch_bounds__(kK1Threads, 1)
k1_match(const uint8_t *__restrict__ in, const uint64_t *__restrict__ in_off,
         const uint32_t *__restrict__ in_len, match_t *__restrict__ matches, uint32_t n_streams,
         uint32_t *__restrict__ next_stream)
{
    LZS_DYN_SMEM(uint8_t, smem);
    uint16_t *heads = reinterpret_cast<uint16_t *>(smem);
    uint16_t *links = heads + kK1Levels * kK1Slots;
    uint32_t *W = reinterpret_cast<uint32_t *>(links + kK1Levels * kK1Ring);
    __shared__ uint32_t s_sid;

    const uint32_t tid = threadIdx.x;
    const uint32_t warp = tid >> 5;

    for (uint32_t x = tid; x < kK1Levels * (kK1Slots + kK1Ring); x += kK1Threads) heads[x] = 0;
    uint32_t epoch = 1;                      /* running 16-bit position base across streams */

    for (;;) {
        __syncthreads();
        if (tid == 0) s_sid = atomicAdd(next_stream, 1u);
        __syncthreads();
        const uint32_t sid = s_sid;
        if (sid >= n_streams) break;

        const uint32_t n = in_len[sid];
        const uint8_t *src = in + in_off[sid];
        const uint8_t *end = src + n;
        match_t       *mout = matches + in_off[sid];

        for (uint32_t t0 = 0; t0 < n; t0 += kK1Tile) {
            const uint32_t tile_n = umin32(kK1Tile, n - t0);
            /* 4-byte grams for the new positions (+8 look-ahead for 12-byte compares) */
            const uint32_t p_lo = (t0 == 0) ? 0u : t0 + 8u;
            const uint32_t p_hi = t0 + kK1Tile + 8u;
            for (uint32_t p = p_lo + tid; p < p_hi; p += kK1Threads)
                W[p & (kK1Ring - 1)] = (p < n) ? load4_unaligned(src + p, end) : 0u;
            __syncthreads();

            switch (warp) {
                case 0:  k1_build_level<2>(heads, links, W, t0, tile_n, epoch); break;
                case 1:  k1_build_level<3>(heads, links, W, t0, tile_n, epoch); break;
                case 2:  k1_build_level<4>(heads, links, W, t0, tile_n, epoch); break;
                case 3:  k1_build_level<5>(heads, links, W, t0, tile_n, epoch); break;
                case 4:  k1_build_level<6>(heads, links, W, t0, tile_n, epoch); break;
                case 5:  k1_build_level<7>(heads, links, W, t0, tile_n, epoch); break;
                case 6:  k1_build_level<8>(heads, links, W, t0, tile_n, epoch); break;
                case 7:  k1_build_level<9>(heads, links, W, t0, tile_n, epoch); break;
                case 8:  k1_build_level<10>(heads, links, W, t0, tile_n, epoch); break;
                case 9:  k1_build_level<11>(heads, links, W, t0, tile_n, epoch); break;
                case 10: k1_build_level<12>(heads, links, W, t0, tile_n, epoch); break;
                default: break;
            }
            __syncthreads();

            for (uint32_t r = tid; r < tile_n; r += kK1Threads) {
                const uint32_t i = t0 + r;
                mout[i] = static_cast<match_t>(k1_query(links, W, i, n));
            }
            __syncthreads();
        }
        epoch = (epoch + n) & 0xFFFFu;
    }
}

}  // namespace lzs

#endif /* LZS_B200_K1_MATCH_CUH */
