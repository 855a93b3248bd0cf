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
 *   level k keeps a 2048-slot last-occurrence table (32-bit heads) in shared memory
 *   and inserts positions in order, 32 at a time: one atomic exchange per lane puts
 *   the position into its slot and returns the predecessor (see k1_build_level for
 *   why that is exact, and for the safe launch that backs the assumption it makes).
 *   Each position stores the distance to its predecessor in the same SLOT; slots
 *   are hashes, so a query verifies bytes and, on a foreign entry, follows the
 *   distance chain (every in-window position of the slot is on it, nearest
 *   first).  Table and chain garbage (stale entries, aliasing) can only
 *   produce candidates that fail the byte check, never a wrong answer -- the same
 *   argument that lets the reference run on uninitialised tables
 *   (lzs-compression.c:253-254).
 *
 *   A query probes level 2 first (incompressible data stops there) and then moves
 *   upwards; a verified candidate of length l found at level k is also the nearest
 *   candidate at level l, so the next level tried is l + 1, and the first level
 *   without a candidate ends the search.
 *
 * Layout: one persistent CTA per SM (216 KiB of shared memory: 11 head tables,
 * 11 link rings, the run table, a ring of 4-byte grams), 29 warps in four roles:
 *   - 1 loader warp pulls streams from a global counter, cuts them into 448-position
 *     tiles and fills the gram ring (aligned word loads, a tile ahead; shuffles and
 *     funnel shifts make the grams), publishing each tile by a flag;
 *   - 11 level warps + 1 run-table warp build a tile each at their own pace;
 *   - 16 query warps take the tile's positions in chunks of 32 from a counter.
 * Tiles are handed from the build warps to the query warps and back through
 * mbarriers (arrive / wait-on-phase), four tiles deep, so no warp ever has to meet
 * the other warps of its own group: a level warp only waits for "the query warps
 * have left the tile that used this stage before", a query warp only for "all
 * twelve build warps have finished this tile".  Positions are numbered
 * continuously across the streams a CTA processes ("virtual positions", every
 * stream starting on a multiple of 32), so the rings need no clearing between
 * streams; a candidate is valid only if its distance does not exceed the position
 * inside the current stream.
 *
 * What was tried and what it measured: profiles/k1_experiments.md.
 *
 * Output: one uint16 per input byte, (len << 11) | offset, consumed by K2.
 * HBM traffic per input byte: 1 B read + 2 B written (intermediate).
 */
#ifndef LZS_B200_K1_MATCH_CUH
#define LZS_B200_K1_MATCH_CUH

#include "lzs_common.cuh"

namespace lzs {

constexpr int      kK1Levels = 11;          /* k = 2 .. 12 */
constexpr uint32_t kK1Slots = 2048;          /* 32-bit heads (exchanged atomically) */
constexpr uint32_t kK1LinkRing = 4096;      /* >= 2047 + 2 * (tile + gap)          */
constexpr uint32_t kK1WRing = 8192;
constexpr uint32_t kK1WMirror = 64;         /* the first grams again behind the ring: index, +4, +8 need one wrap */
#ifndef LZS_K1_TILE
#define LZS_K1_TILE 448
#endif
#ifndef LZS_K1_FAKE_BUILD_WARPS
#define LZS_K1_FAKE_BUILD_WARPS 99          /* timing experiments only: build warps from this index on do nothing (wrong results) */
#endif
#ifndef LZS_K1_FAKE_LOADER_GRAMS
#define LZS_K1_FAKE_LOADER_GRAMS 4          /* timing experiments only: < 4 leaves grams unwritten (wrong results) */
#endif
#ifndef LZS_K1_LAG
#define LZS_K1_LAG 0                        /* 1: a batch's links are made one batch later (two batches of exchanges in flight) */
#endif
#ifndef LZS_K1_EAGER
#define LZS_K1_EAGER 0                      /* 1: a query step loads the candidate's bytes together with its link entry */
#endif
#ifndef LZS_K1_BUILD_UNROLL
#define LZS_K1_BUILD_UNROLL 2
#endif
#ifndef LZS_K1_DEPTH
#define LZS_K1_DEPTH 4
#endif
constexpr uint32_t kK1Tile = LZS_K1_TILE;   /* a multiple of 32; 448 = 14 batches  */
constexpr int      kK1BuildUnroll = LZS_K1_BUILD_UNROLL;
constexpr uint32_t kK1Depth = LZS_K1_DEPTH; /* tiles the build warps may be ahead of the query group (power of two) */
/* Virtual positions between streams: 12 zero grams behind the last byte, then up to the next
 * multiple of 32 (every stream starts on a batch boundary, so a batch never straddles a ring
 * wrap and the positions a last batch inserts past the end of its stream belong to no stream). */
constexpr uint32_t kK1StreamGap = 16 + 31;
#ifndef LZS_K1_LPW
#define LZS_K1_LPW 1                        /* levels per build warp */
#endif
constexpr int      kK1Lpw = LZS_K1_LPW;
#ifndef LZS_K1_GROUPED
#define LZS_K1_GROUPED 1                    /* 1: the group code (compile-time levels, exchanges in flight while the next batch is hashed) also for one level per warp */
#endif
constexpr bool     kK1Grouped = LZS_K1_GROUPED != 0;
/* one warp per level + the run-table warp, or a group of levels per warp (the last one also builds the run table) */
constexpr int      kK1BuildWarps = kK1Lpw == 1 ? kK1Levels + 1 : (kK1Levels + kK1Lpw - 1) / kK1Lpw;
#ifndef LZS_K1_QW
#define LZS_K1_QW 18
#endif
constexpr int      kK1QueryWarps = LZS_K1_QW;
constexpr int      kK1Threads = 32 * (kK1BuildWarps + 1 + kK1QueryWarps);   /* + the loader warp */
constexpr unsigned kK1BuildThreads = 32 * kK1BuildWarps;
constexpr unsigned kK1QueryThreads = 32 * kK1QueryWarps;
constexpr size_t   kK1SmemBytes = static_cast<size_t>(kK1Levels) * kK1Slots * 4 +
                                static_cast<size_t>(kK1Levels + 1) * kK1LinkRing * 2 +
                                (kK1WRing + kK1WMirror) * 4;
/* run table entry: (forward run length capped at 12) << 12 | distance back to the run start */
constexpr uint32_t kRunBackMask = 0xFFFu;
static_assert((kK1Depth & (kK1Depth - 1)) == 0 && kK1Depth >= 2 && kK1Depth <= 4, "pipeline depth");
static_assert(kWindow + kK1Depth * (kK1Tile + kK1StreamGap) < kK1LinkRing, "link ring too small for the pipeline");
constexpr uint32_t kK1Ahead = 40;           /* grams filled beyond the tile being built */
static_assert(kWindow + (kK1Depth + 2) * (kK1Tile + kK1StreamGap) + 16 + kK1Ahead < kK1WRing, "gram ring: window + the tiles in the pipeline + those the loader may be ahead");

constexpr uint32_t kK1EndOfWork = 0xFFFFFFFFu;

#if defined(LZS_K1_TIMELINE) && !defined(LZS_SIMT_EMU)
/* Debug builds only (tools/k1_timeline.py): per tile of CTA 0, when each role started and finished it
 * (SM clock).  [tile][0] loader may start (ring space free), [1] loader published, [2] first build warp
 * starts, [3] last build warp starts, [4] first build warp done, [5] last build warp done, [6] first
 * query warp enters, [7] last query warp leaves. */
constexpr uint32_t kTlTiles = 8192;
__device__ unsigned long long g_k1_timeline[kTlTiles][8];
__device__ unsigned long long g_k1_warp_busy[64][2];      /* per warp of CTA 0: cycles spent building / waiting, summed over tiles */
__device__ __forceinline__ void tl_set(uint32_t g, int k)
{
    if (blockIdx.x == 0 && g < kTlTiles && lane_id() == 0) g_k1_timeline[g][k] = clock64();
}
__device__ __forceinline__ void tl_min(uint32_t g, int k)
{
    if (blockIdx.x == 0 && g < kTlTiles && lane_id() == 0) atomicMin(&g_k1_timeline[g][k], static_cast<unsigned long long>(clock64()));
}
__device__ __forceinline__ void tl_max(uint32_t g, int k)
{
    if (blockIdx.x == 0 && g < kTlTiles && lane_id() == 0) atomicMax(&g_k1_timeline[g][k], static_cast<unsigned long long>(clock64()));
}
#else
#define tl_set(g, k) ((void)0)
#define tl_min(g, k) ((void)0)
#define tl_max(g, k) ((void)0)
#endif

struct K1Tile {
    uint32_t sid, t0, tile_n, n, v0;
    uint32_t skip;      /* the first `skip` positions are kept history of the flow (built into the tables, not matched) */
    uint32_t seg;       /* 0, or the stream is a flow of packets of this many bytes: tokens end with the packet */
    uint32_t look;      /* the last `look` positions belong to the NEXT piece of a long stream: built and compared against, not matched */
};

/* Hash of the k-gram that starts the 12 bytes (w0, w1, w2); (m0, m1, m2) mask the bytes that
 * belong to the gram.  Bits 31..21 are the table slot, bits 19..15 a 5-bit tag kept beside every
 * chain link so that a query can reject most foreign entries of its slot without touching their
 * bytes.  The three products are independent (one dependent multiply less than a chained hash).
 * The masks are run-time values on purpose: all eleven level warps execute the SAME loop code,
 * which is what keeps it resident in the instruction cache. */
__device__ __forceinline__ uint32_t gram_hash(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t m0, uint32_t m1,
                                              uint32_t m2)
{
    uint32_t h = ((w0 & m0) * 0x9E3779B1u) ^ ((w1 & m1) * 0x85EBCA77u) ^ ((w2 & m2) * 0xC2B2AE3Du);
    h ^= h >> 15;
    h *= 0x27D4EB2Fu;
    return h;
}
constexpr uint32_t kSlotShift = 21;          /* slot = h >> 21 (11 bits)                    */
constexpr uint32_t kTagShift = 15;           /* tag  = (h >> 15) & 31                       */
constexpr uint32_t kLinkDistMask = 0x7FFu;   /* link entry: (tag << 11) | distance          */

/* common prefix length (0..12) of two 12-byte strings given as LE words; branch-free, so the
 * lanes of a warp stay together whatever the data */
__device__ __forceinline__ uint32_t lcp12(uint32_t a0, uint32_t a1, uint32_t a2,
                                          uint32_t b0, uint32_t b1, uint32_t b2)
{
    const uint32_t x0 = a0 ^ b0, x1 = a1 ^ b1, x2 = a2 ^ b2;
    /* first differing word, and 4 x its index; the sentinel bit makes "all equal" come out as 12 */
    const uint32_t x = x0 ? x0 : (x1 ? x1 : (x2 ? x2 : 1u));
    const uint32_t base = x0 ? 0u : (x1 ? 4u : 8u + ((x2 == 0u) ? 4u : 0u));
    return base + (static_cast<uint32_t>(__ffs(static_cast<int>(x)) - 1) >> 3);
}

/* Exchange on a shared-memory word: returns the previous value.  Called by all 32 lanes together. */
__device__ __forceinline__ uint32_t smem_exch(uint32_t *p, uint32_t v)
{
#ifdef LZS_SIMT_EMU
    return simt_warp_exch(p, v);
#else
    uint32_t o;
    asm volatile("atom.shared.exch.b32 %0, [%1], %2;"
                 : "=r"(o)
                 : "r"(static_cast<uint32_t>(__cvta_generic_to_shared(p))), "r"(v)
                 : "memory");
    return o;
#endif
}

/* Exact repair of one batch's exchanges, for the case that the lanes sharing a slot were not
 * served in ascending lane order (sm_100a serves them in ascending order --
 * tools/micro/atoms_exch.cu -- so this is insurance, exercised by the emulator tests).
 * Whatever the order, exactly one lane of every group received the pre-batch head.  Returns the
 * position each lane should have received and leaves the group's highest lane in the head. */
__device__ __noinline__ uint32_t k1_relink(uint32_t *slot, uint32_t key, uint32_t old, uint32_t vb, uint32_t pos)
{
    const uint32_t lane = lane_id();
    const uint32_t grp = __match_any_sync(LZS_FULL_MASK, key);
    const uint32_t lower = grp & ((1u << lane) - 1u);
    const uint32_t outside = __ballot_sync(LZS_FULL_MASK, (old - vb) >= 32u);
    const uint32_t pre = __shfl_sync(LZS_FULL_MASK, old, __ffs(static_cast<int>(grp & outside)) - 1);
    if ((grp >> lane) == 1u) *slot = pos;
    __syncwarp();
    return lower ? (vb | (31u - static_cast<uint32_t>(__clz(static_cast<int>(lower))))) : pre;
}

/* Insert the positions of one tile into level K's table, in order, and record for each the
 * distance to the previous position of the same slot (0 = none within 2047).  Executed by one
 * whole warp; vt = virtual position of the tile start, a multiple of 32.  One atomic exchange
 * per lane puts the position into the slot's head and returns its predecessor: lanes of one
 * batch that share a slot are served in ascending lane (= position) order, so each receives the
 * lane before it and the highest one stays in the head.  That order is what sm_100a does
 * (tools/micro/atoms_exch.cu), not something PTX promises, so the fast kernel only RECORDS whether
 * a lane ever received a higher lane of its own batch (the returned flag; nothing in the loop
 * waits for it) and the safe kernel (kSafe: every batch repaired with k1_relink, exact for any
 * service order) re-does the whole batch of streams if that was ever seen.  The last batch of a stream runs all 32
 * lanes: the positions past the end sit in the gap before the next stream, where no query ever
 * looks (a candidate is valid only up to the query's own position inside its stream). */
template <bool kSafe>
__device__ __forceinline__ uint32_t k1_build_level(uint32_t *hd, uint16_t *lk, const uint32_t *W, uint32_t vt,
                                                   uint32_t tile_n, uint32_t m0, uint32_t m1, uint32_t m2)
{
    const uint32_t  lane = lane_id();
    const uint32_t *Wl = W + lane;
    lk += lane;
    uint32_t x = vt & (kK1WRing - 1);
    uint32_t hnext = gram_hash(Wl[x], Wl[x + 4], Wl[x + 8], m0, m1, m2);
    uint32_t disorder = 0;
#pragma unroll kK1BuildUnroll   /* one copy serves all eleven levels */
    for (uint32_t b = 0; b < tile_n; b += 32) {
        const uint32_t vb = vt + b;                       /* warp-uniform, a multiple of 32 */
        const uint32_t pos = vb | lane;
        const uint32_t h = hnext;
        uint32_t      *slot = hd + (h >> kSlotShift);
        x = (vb + 32) & (kK1WRing - 1);
        hnext = gram_hash(Wl[x], Wl[x + 4], Wl[x + 8], m0, m1, m2);   /* next batch: independent of the table */

        uint32_t       old = smem_exch(slot, pos);
        __syncwarp();                                     /* batch after batch, also formally */
        if (kSafe)
            old = k1_relink(slot, h >> kSlotShift, old, vb, pos);
        else
            disorder |= (pos - old) >> 31;                /* received a LATER position: not the order assumed */
        const uint32_t dist = pos - old;
        uint32_t       e = (h >> (kTagShift - 11)) & 0xF800u;   /* tag << 11 */
        if (dist <= kWindow) e |= dist;                   /* further than the window: no link */
        lk[vb & (kK1LinkRing - 1)] = static_cast<uint16_t>(e);
    }
    return disorder;
}

/* Hashes of the k-grams that start the 12 bytes (w0, w1, w2), k = 2..12.  Bits 31..21 are the
 * table slot, bits 20..16 a 5-bit tag kept beside every chain link so that a query can reject
 * most foreign entries of its slot without touching their bytes.  A build warp needs the hashes
 * of a few consecutive levels of the same position: the first is made from the masked words, each
 * further one by mixing one more byte into the previous (one LOP3, one IMAD).  Nobody else ever
 * computes these hashes (a query compares the tags stored beside the links), so the levels need
 * not agree on a formula. */
constexpr uint32_t kHashC1 = 0x9E3779B1u, kHashC2 = 0x85EBCA77u;
__device__ __forceinline__ uint32_t gram_mask(int bytes) { return bytes >= 4 ? 0xFFFFFFFFu : ((1u << (8 * bytes)) - 1u); }
__device__ __forceinline__ uint32_t k1_hash_start(int k, uint32_t w0, uint32_t w1, uint32_t w2)
{
    if (k <= 4) return (w0 & gram_mask(k)) * kHashC1;
    if (k <= 8) return ((w0 * kHashC1) ^ (w1 & gram_mask(k - 4))) * kHashC2;
    return ((((w0 * kHashC1) ^ w1) * kHashC2) ^ (w2 & gram_mask(k - 8))) * kHashC1;
}
/* hash of the k-gram from the hash of the (k-1)-gram: byte k-1 comes in where it sits in its word */
__device__ __forceinline__ uint32_t k1_hash_roll(int k, uint32_t h, uint32_t w0, uint32_t w1, uint32_t w2)
{
    const int      b = k - 1;
    const uint32_t w = b < 4 ? w0 : (b < 8 ? w1 : w2);
    return (h ^ (w & (0xFFu << (8 * (b & 3))))) * kHashC1;
}

/* Insert the positions of one tile into the tables of levels K0 .. K0+NL-1, in order, and record
 * for each position and level the distance to the previous position of the same slot (0 = none
 * within 2047).  Executed by one whole warp; vt = virtual position of the tile start, a multiple
 * of 32.  One atomic exchange per lane and level puts the position into the slot's head and
 * returns its predecessor: lanes of one batch that share a slot are served in ascending lane
 * (= position) order, so each receives the lane before it and the highest one stays in the head.
 * That order is what sm_100a does (tools/micro/atoms_exch.cu), not something PTX promises, so the
 * fast kernel only RECORDS whether a lane ever received a higher lane of its own batch (the
 * returned flag; nothing in the loop waits for it) and the safe kernel (kSafe: every batch
 * repaired with k1_relink, exact for any service order) re-does the whole batch of streams if that
 * was ever seen.  The last batch of a stream runs all 32 lanes: the positions past the end sit in
 * the gap before the next stream, where no query ever looks (a candidate is valid only up to the
 * query's own position inside its stream).  The levels of a group are independent of each other,
 * so their exchanges are in flight together. */
template <int K0, int NL, bool kSafe>
__device__ __forceinline__ uint32_t k1_build_group(uint32_t *heads, uint16_t *links, const uint32_t *W, uint32_t vt,
                                                   uint32_t tile_n)
{
    const uint32_t  lane = lane_id();
    const uint32_t *Wl = W + lane;
    uint32_t       *hd = heads + (K0 - 2) * kK1Slots;
    uint16_t       *lk0 = links + (K0 - 2) * kK1LinkRing + lane;
    uint32_t        disorder = 0;
    /* Software pipeline: the exchanges of a batch are issued, then the grams of the NEXT batch are
     * loaded and hashed while the exchanges are under way, and only then are their results turned
     * into links -- a warp with several levels is bound by what it issues, not by the round trip
     * of an exchange (exchanges of one warp reach the table in program order). */
    uint32_t h[NL];
    {
        const uint32_t x = vt & (kK1WRing - 1);
        const uint32_t w0 = Wl[x], w1 = Wl[x + 4], w2 = Wl[x + 8];
        h[0] = k1_hash_start(K0, w0, w1, w2);
#pragma unroll
        for (int l = 1; l < NL; l++) h[l] = k1_hash_roll(K0 + l, h[l - 1], w0, w1, w2);
    }
#if LZS_K1_LAG
    if (!kSafe) {
        /* One batch further: the results of a batch's exchanges are turned into links only after the
         * NEXT batch's exchanges have been issued (and the batch after that hashed), so two batches
         * of exchanges are in flight and nothing in the loop waits for a round trip. */
        uint32_t  p_old[NL], p_tag[NL], p_pos = 0;
        uint16_t *p_lk = lk0;
        bool      pending = false;
#pragma unroll
        for (int l = 0; l < NL; l++) p_old[l] = p_tag[l] = 0;
#pragma unroll 2
        for (uint32_t b = 0; b < tile_n; b += 32) {
            const uint32_t vb = vt + b;
            const uint32_t pos = vb | lane;
            uint32_t       old[NL], tagbits[NL];
#pragma unroll
            for (int l = 0; l < NL; l++) {
                old[l] = smem_exch(hd + l * kK1Slots + (h[l] >> kSlotShift), pos);
                tagbits[l] = (h[l] >> 5) & 0xF800u;
            }
            {
                const uint32_t x = (vb + 32u) & (kK1WRing - 1);
                const uint32_t w0 = Wl[x], w1 = Wl[x + 4], w2 = Wl[x + 8];
                h[0] = k1_hash_start(K0, w0, w1, w2);
#pragma unroll
                for (int l = 1; l < NL; l++) h[l] = k1_hash_roll(K0 + l, h[l - 1], w0, w1, w2);
            }
            if (pending) {
#pragma unroll
                for (int l = 0; l < NL; l++) {
                    const uint32_t dist = p_pos - p_old[l];
                    disorder |= dist;
                    uint32_t e = p_tag[l];
                    if (dist <= kWindow) e |= dist;
                    p_lk[l * kK1LinkRing] = static_cast<uint16_t>(e);
                }
            }
#pragma unroll
            for (int l = 0; l < NL; l++) { p_old[l] = old[l]; p_tag[l] = tagbits[l]; }
            p_pos = pos;
            p_lk = lk0 + (vb & (kK1LinkRing - 1));
            pending = true;
            __syncwarp();
        }
        if (pending) {
#pragma unroll
            for (int l = 0; l < NL; l++) {
                const uint32_t dist = p_pos - p_old[l];
                disorder |= dist;
                uint32_t e = p_tag[l];
                if (dist <= kWindow) e |= dist;
                p_lk[l * kK1LinkRing] = static_cast<uint16_t>(e);
            }
        }
        return disorder >> 31;
    }
#endif
#pragma unroll 2
    for (uint32_t b = 0; b < tile_n; b += 32) {
        const uint32_t vb = vt + b;                       /* warp-uniform, a multiple of 32 */
        const uint32_t pos = vb | lane;
        uint16_t      *lk = lk0 + (vb & (kK1LinkRing - 1));
        uint32_t       old[NL], tagbits[NL];
#pragma unroll
        for (int l = 0; l < NL; l++) {
            uint32_t *slot = hd + l * kK1Slots + (h[l] >> kSlotShift);
            old[l] = smem_exch(slot, pos);
            if (kSafe) {
                __syncwarp();
                old[l] = k1_relink(slot, h[l] >> kSlotShift, old[l], vb, pos);
            }
            tagbits[l] = (h[l] >> 5) & 0xF800u;           /* tag << 11 */
        }
        {
            /* the batch after this one (past the tile's end: grams that are there anyway) */
            const uint32_t x = (vb + 32u) & (kK1WRing - 1);
            const uint32_t w0 = Wl[x], w1 = Wl[x + 4], w2 = Wl[x + 8];
            h[0] = k1_hash_start(K0, w0, w1, w2);
#pragma unroll
            for (int l = 1; l < NL; l++) h[l] = k1_hash_roll(K0 + l, h[l - 1], w0, w1, w2);
        }
#pragma unroll
        for (int l = 0; l < NL; l++) {
            const uint32_t dist = pos - old[l];
            if (!kSafe) disorder |= dist;                 /* sign bit: received a LATER position, not the order assumed */
            uint32_t e = tagbits[l];
            if (dist <= kWindow) e |= dist;               /* further than the window: no link */
            lk[l * kK1LinkRing] = static_cast<uint16_t>(e);
        }
        __syncwarp();                                     /* batch after batch, also formally */
    }
    return disorder >> 31;
}

/* Run table of one tile (one whole warp).  For every position p it records how far back
 * the run of identical bytes containing p starts (0 = p starts a run, capped at 4095)
 * and how many identical bytes follow from p (capped at 12).  Positions p-1 and p have the
 * same k-gram whenever the forward run at p-1 covers k+1 bytes, so inside a run of one
 * byte value every level's chain visits the run members one by one; the table lets a
 * query hop over all of them at once (see k1_query). */
__device__ __forceinline__ void k1_build_runs(uint16_t *runs, const uint32_t *W, uint32_t v0, uint32_t t0,
                                              uint32_t tile_n)
{
    const uint32_t  lane = lane_id();
    const uint32_t  vt = v0 + t0;                         /* a multiple of 32 */
    const uint32_t *Wl = W + lane;
    uint16_t       *rl = runs + lane;
    /* carried from batch to batch in a register: how far back the run of the batch's last byte starts.
     * The byte in front of every position comes from the gram ring (the gram one position back), so
     * the only thing a batch waits for from the batch before it is that one number. */
    uint32_t carry = (t0 == 0) ? 0u : (runs[(vt - 1u) & (kK1LinkRing - 1)] & kRunBackMask);
#pragma unroll 2
    for (uint32_t b = 0; b < tile_n; b += 32) {
        const uint32_t vb = vt + b;
        const uint32_t x = vb & (kK1WRing - 1);
        const uint32_t w0 = Wl[x], w1 = Wl[x + 4], w2 = Wl[x + 8];
        const uint32_t wb = W[(vb + lane - 1u) & (kK1WRing - 1)];              /* the gram that starts one byte earlier */
        const uint32_t byte = w0 & 0xFFu;
        const uint32_t rep = byte * 0x01010101u;
        const uint32_t fwd = lcp12(w0, w1, w2, rep, rep, rep);                 /* 1..12 */
        /* a stream's first position starts a run whatever lies in front of it in the ring */
        const bool     start = (t0 + b + lane == 0u) || ((wb & 0xFFu) != byte);
        const uint32_t starts = __ballot_sync(LZS_FULL_MASK, start);           /* positions that begin a run */
        const uint32_t below = starts & ((2u << lane) - 1u);                   /* starts at or below my lane */
        const uint32_t back = below ? lane - (31u - static_cast<uint32_t>(__clz(static_cast<int>(below))))
                                    : umin32(carry + lane + 1u, kRunBackMask);  /* the run began in an earlier batch */
        rl[vb & (kK1LinkRing - 1)] = static_cast<uint16_t>((fwd << 12) | back);
        carry = __shfl_sync(LZS_FULL_MASK, back, 31);
    }
}

/* One build warp's share of a tile: its group of levels, and for the last group the run table. */
template <bool kSafe>
__device__ __forceinline__ uint32_t k1_build_tile(uint32_t warp, uint32_t *heads, uint16_t *links, uint16_t *runs,
                                                  const uint32_t *W, const K1Tile &d)
{
    const uint32_t vt = d.v0 + d.t0;
    uint32_t       dis = 0;
#if LZS_K1_LPW == 3
    switch (warp) {
        case 0: dis = k1_build_group<2, 3, kSafe>(heads, links, W, vt, d.tile_n); break;
        case 1: dis = k1_build_group<5, 3, kSafe>(heads, links, W, vt, d.tile_n); break;
        case 2: dis = k1_build_group<8, 3, kSafe>(heads, links, W, vt, d.tile_n); break;
        default:
            dis = k1_build_group<11, 2, kSafe>(heads, links, W, vt, d.tile_n);
            k1_build_runs(runs, W, d.v0, d.t0, d.tile_n);
            break;
    }
#elif LZS_K1_LPW == 2
    switch (warp) {
        case 0: dis = k1_build_group<2, 2, kSafe>(heads, links, W, vt, d.tile_n); break;
        case 1: dis = k1_build_group<4, 2, kSafe>(heads, links, W, vt, d.tile_n); break;
        case 2: dis = k1_build_group<6, 2, kSafe>(heads, links, W, vt, d.tile_n); break;
        case 3: dis = k1_build_group<8, 2, kSafe>(heads, links, W, vt, d.tile_n); break;
        case 4: dis = k1_build_group<10, 2, kSafe>(heads, links, W, vt, d.tile_n); break;
        default:
            dis = k1_build_group<12, 1, kSafe>(heads, links, W, vt, d.tile_n);
            k1_build_runs(runs, W, d.v0, d.t0, d.tile_n);
            break;
    }
#elif LZS_K1_LPW == 1
    switch (warp) {
        case 0: dis = k1_build_group<2, 1, kSafe>(heads, links, W, vt, d.tile_n); break;
        case 1: dis = k1_build_group<3, 1, kSafe>(heads, links, W, vt, d.tile_n); break;
        case 2: dis = k1_build_group<4, 1, kSafe>(heads, links, W, vt, d.tile_n); break;
        case 3: dis = k1_build_group<5, 1, kSafe>(heads, links, W, vt, d.tile_n); break;
        case 4: dis = k1_build_group<6, 1, kSafe>(heads, links, W, vt, d.tile_n); break;
        case 5: dis = k1_build_group<7, 1, kSafe>(heads, links, W, vt, d.tile_n); break;
        case 6: dis = k1_build_group<8, 1, kSafe>(heads, links, W, vt, d.tile_n); break;
        case 7: dis = k1_build_group<9, 1, kSafe>(heads, links, W, vt, d.tile_n); break;
        case 8: dis = k1_build_group<10, 1, kSafe>(heads, links, W, vt, d.tile_n); break;
        case 9: dis = k1_build_group<11, 1, kSafe>(heads, links, W, vt, d.tile_n); break;
        case 10: dis = k1_build_group<12, 1, kSafe>(heads, links, W, vt, d.tile_n); break;
        default: k1_build_runs(runs, W, d.v0, d.t0, d.tile_n); break;
    }
#elif LZS_K1_LPW == 4
    switch (warp) {
        case 0: dis = k1_build_group<2, 4, kSafe>(heads, links, W, vt, d.tile_n); break;
        case 1: dis = k1_build_group<6, 4, kSafe>(heads, links, W, vt, d.tile_n); break;
        default:
            dis = k1_build_group<10, 3, kSafe>(heads, links, W, vt, d.tile_n);
            k1_build_runs(runs, W, d.v0, d.t0, d.tile_n);
            break;
    }
#else
#error "LZS_K1_LPW must be 1, 2, 3 or 4"
#endif
    return dis;
}


/* One query = a single flat loop of chain steps (no nested loops, so lanes of a warp
 * stay together).  Levels are tried upwards: a verified candidate of length l at level
 * k answers every level up to l, so the next level tried is l + 1, and the first level
 * whose chain ends without a verified candidate ends the search ("some candidate
 * reaches k" is monotone in k).  Every chain link carries the 5-bit tag of the entry
 * it belongs to, so a foreign entry costs one shared-memory load. */
#if defined(LZS_SIMT_EMU) && defined(LZS_K1_STATS)
extern "C" unsigned long long g_k1_stats[8];   /* queries, steps, foreign, verified-fail, levels, run skips, max steps */
extern "C" unsigned char g_k1_walk[1 << 22];    /* chain steps of the query at virtual position v (mod 4 Mi) */
extern "C" unsigned char g_k1_foreign[1 << 22]; /* ... of which landed on a foreign entry (tag mismatch)     */
#define LZS_STAT(i, v) (g_k1_stats[i] += (v))
#define LZS_STAT_WALK(v, steps) (g_k1_walk[(v) & ((1u << 22) - 1u)] = static_cast<unsigned char>((steps) > 255u ? 255u : (steps)))
#define LZS_STAT_FOREIGN(v) (g_k1_foreign[(v) & ((1u << 22) - 1u)] += g_k1_foreign[(v) & ((1u << 22) - 1u)] < 255u ? 1u : 0u)
#else
#define LZS_STAT_FOREIGN(v) ((void)0)
#define LZS_STAT(i, v) ((void)0)
#define LZS_STAT_WALK(v, steps) ((void)0)
#endif

__device__ __forceinline__ uint32_t k1_query(const uint16_t *links, const uint16_t *runs, const uint32_t *W,
                                             uint32_t v0, uint32_t i, uint32_t n)
{
    const uint32_t M = umin32(kSearchMax, n - i);
    const uint32_t maxd = umin32(kWindow, i);
    if (M < kMinLen || maxd == 0) return 0;
    const uint32_t v = v0 + i;
    const uint32_t *wv = W + (v & (kK1WRing - 1));           /* the mirror covers +4 and +8 */
    const uint32_t w0 = wv[0], w1 = wv[4], w2 = wv[8];
    uint32_t best = 0, bd = 0;
    uint32_t k = kMinLen;
    const uint16_t *lk = links;                              /* level k's ring              */
    uint32_t e = lk[v & (kK1LinkRing - 1)];                  /* own entry: tag + first link */
    uint32_t tag = e >> 11;
    uint32_t d = e & kLinkDistMask;
    uint32_t tot = 0;
    uint32_t walked = 0;                                     /* statistics builds only */
    LZS_STAT(0, 1); LZS_STAT(4, 1);
    for (;;) {
        tot += d;
        if (d == 0 || tot > maxd) break;                     /* level k has no candidate: done */
        LZS_STAT(1, 1);
        walked++;
        const uint32_t j = v - tot;
        e = lk[j & (kK1LinkRing - 1)];
#if LZS_K1_EAGER
        /* the candidate's twelve bytes are requested together with its link entry: one
         * shared-memory round trip per step instead of two when the entry is one of ours */
        const uint32_t *wje = W + (j & (kK1WRing - 1));
        const uint32_t  y0 = wje[0], y1 = wje[4], y2 = wje[8];
#ifndef LZS_SIMT_EMU
        asm volatile("" ::"r"(e), "r"(y0), "r"(y1), "r"(y2));
#endif
#endif
        d = e & kLinkDistMask;
        if ((e >> 11) != tag) {                              /* foreign entry of this slot  */
            LZS_STAT(2, 1);
            LZS_STAT_FOREIGN(v);
            if (d == 1u) {
                /* its predecessor is the adjacent position: if j sits inside a run of one
                 * byte value that covers k bytes from j, every run member before j has j's
                 * gram (not ours) and is the next entry of this chain -- skip to the run start */
                const uint32_t r = runs[j & (kK1LinkRing - 1)];
                const uint32_t back = r & kRunBackMask;
                if ((r >> 12) >= k && back != 0u) {
                    tot += back;
                    LZS_STAT(5, 1);
                    d = lk[(j - back) & (kK1LinkRing - 1)] & kLinkDistMask;
                }
            }
            continue;
        }
#if LZS_K1_EAGER
        const uint32_t l = umin32(lcp12(w0, w1, w2, y0, y1, y2), M);
#else
        const uint32_t *wj = W + (j & (kK1WRing - 1));
        const uint32_t l = umin32(lcp12(w0, w1, w2, wj[0], wj[4], wj[8]), M);
#endif
        if (l < k) { LZS_STAT(3, 1); continue; }
        LZS_STAT(4, 1);
        best = l;                                            /* nearest candidate of length l */
        bd = tot;
        if (l >= M) break;
        k = l + 1;                                           /* next level to try           */
        lk = links + (k - 2) * kK1LinkRing;
        e = lk[v & (kK1LinkRing - 1)];
        tag = e >> 11;
        d = e & kLinkDistMask;
        tot = 0;
    }
    LZS_STAT_WALK(v, walked);
    (void)walked;
    return (best << kMatchOffBits) | bd;
}

/* The loader's view of one tile: the aligned words that cover the tile's new grams.  Lane l holds
 * words l, l + 32, ... of the range that starts at the aligned word containing byte p_lo of the
 * stream (one more 32-word row than there are 128-position groups, because a gram reaches into
 * the following word).  Addresses beyond the stream's last word are clamped and read as zero. */
constexpr int kK1LoadGroups = static_cast<int>((kK1Tile + kK1Ahead + 127) / 128);
struct K1Words {
    uint32_t w[kK1LoadGroups + 1];
};
__device__ __forceinline__ void k1_load_words(K1Words &r, const uint8_t *src, uint32_t p_lo, uintptr_t wlast)
{
    const uintptr_t base = (reinterpret_cast<uintptr_t>(src) + p_lo) & ~static_cast<uintptr_t>(3);
#pragma unroll
    for (int k = 0; k <= kK1LoadGroups; k++) {
        const uintptr_t a = base + 4u * (static_cast<uint32_t>(k) * 32u + lane_id());
        const uint32_t  v = __ldg(reinterpret_cast<const uint32_t *>(a <= wlast ? a : wlast));
        r.w[k] = (a <= wlast) ? v : 0u;
    }
}
/* grams p_lo .. p_hi-1 of the stream from the words loaded above: lane l makes the four grams
 * 4l .. 4l+3 of every 128-position group out of words l, l+1, l+2 */
__device__ __forceinline__ void k1_store_grams(const K1Words &r, uint32_t *W, const uint8_t *src, uint32_t v0,
                                               uint32_t p_lo, uint32_t p_hi, uint32_t n)
{
    const uint32_t lane = lane_id();
    const uint32_t m = static_cast<uint32_t>((reinterpret_cast<uintptr_t>(src) + p_lo) & 3u);   /* warp-uniform */
#pragma unroll
    for (int k = 0; k < kK1LoadGroups; k++) {
        const uint32_t x0 = r.w[k];
        const uint32_t d1 = __shfl_down_sync(LZS_FULL_MASK, r.w[k], 1), n1 = __shfl_sync(LZS_FULL_MASK, r.w[k + 1], 0);
        const uint32_t d2 = __shfl_down_sync(LZS_FULL_MASK, r.w[k], 2),
                       n2 = __shfl_sync(LZS_FULL_MASK, r.w[k + 1], (lane + 2u) & 31u);
        const uint32_t x1 = lane == 31u ? n1 : d1;
        const uint32_t x2 = lane >= 30u ? n2 : d2;
        const uint32_t q0 = p_lo + static_cast<uint32_t>(k) * 128u + 4u * lane;
#pragma unroll
        for (uint32_t j = 0; j < LZS_K1_FAKE_LOADER_GRAMS; j++) {
            const uint32_t o = m + j;                     /* byte offset from word l: 0..6 */
            const uint32_t g = __funnelshift_r(o < 4u ? x0 : x1, o < 4u ? x1 : x2, (o & 3u) * 8u);
            const uint32_t q = q0 + j;
            if (q < p_hi) {
                const uint32_t x = (v0 + q) & (kK1WRing - 1);
                const uint32_t w = (q < n) ? g : 0u;
                W[x] = w;
                if (x < kK1WMirror) W[kK1WRing + x] = w;
            }
        }
    }
}

#ifndef LZS_K1_EAGER
#define LZS_K1_EAGER 0                      /* 1: a query step loads the candidate's bytes together with its link entry */
#endif
#ifndef LZS_K1_BULK
#define LZS_K1_BULK 0                       /* 1: the loader stages the input with cp.async.bulk (TMA, 1-D) instead of word loads */
#endif
#if defined(LZS_SIMT_EMU)
#undef LZS_K1_BULK
#define LZS_K1_BULK 0                       /* the emulator has no asynchronous copy engine */
#endif
constexpr uint32_t kK1StageBytes = 576;     /* a tile's new bytes (448 + 40 + 3) plus alignment on both sides, rounded up */

#if LZS_K1_BULK
/* One tile's input bytes, requested from the copy engine: [a0, a0 + bytes) is the 16-byte aligned
 * range that covers the bytes q_lo .. hi_b-1 of the stream (bytes == 0: nothing to fetch). */
struct K1Stage {
    uintptr_t a0;
    uint32_t  bytes;
};
__device__ __forceinline__ K1Stage k1_stage_range(const uint8_t *src, uint32_t q_lo, uint32_t q_hi, uint32_t n)
{
    K1Stage        s;
    const uint32_t hi_b = umin32(q_hi + 3u, n);                     /* grams reach 3 bytes on; nothing past the stream */
    const uintptr_t lo = reinterpret_cast<uintptr_t>(src) + q_lo, hi = reinterpret_cast<uintptr_t>(src) + hi_b;
    s.a0 = lo & ~static_cast<uintptr_t>(15);
    s.bytes = hi_b > q_lo ? static_cast<uint32_t>(((hi + 15u) & ~static_cast<uintptr_t>(15)) - s.a0) : 0u;
    return s;
}
/* lane 0: arm the barrier with the byte count and start the copy (global -> shared, completes on the barrier) */
__device__ __forceinline__ void k1_stage_issue(uint8_t *stage, uint64_t *bar, const K1Stage &r)
{
    const uint32_t b = static_cast<uint32_t>(__cvta_generic_to_shared(bar));
    const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(stage));
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    /* earlier reads of the buffer, then the engine's writes */
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(b), "r"(r.bytes) : "memory");
    if (r.bytes)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(d), "l"(r.a0), "r"(r.bytes), "r"(b) : "memory");
}
/* grams p_lo .. p_hi-1 of the stream from the staged bytes: lane l makes the four grams 4l .. 4l+3
 * of every 128-position group out of three staged words */
__device__ __forceinline__ void k1_store_grams_staged(const uint8_t *stage, const K1Stage &r, uint32_t *W,
                                                      const uint8_t *src, uint32_t v0, uint32_t p_lo, uint32_t p_hi,
                                                      uint32_t n)
{
    const uint32_t  lane = lane_id();
    const uint32_t  lead = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(src) + p_lo - r.a0);   /* 0..15 */
    const uint32_t  m = lead & 3u;                                                                 /* warp-uniform */
    const uint32_t *sw = reinterpret_cast<const uint32_t *>(stage) + (lead >> 2);
    const uint32_t  words = r.bytes >> 2;
#pragma unroll
    for (int k = 0; k < kK1LoadGroups; k++) {
        const uint32_t wi = static_cast<uint32_t>(k) * 32u + lane;
        const uint32_t q0 = p_lo + static_cast<uint32_t>(k) * 128u + 4u * lane;
        if (q0 >= p_hi) continue;
        const uint32_t base = (lead >> 2) + wi;
        const uint32_t x0 = base < words ? sw[wi] : 0u;
        const uint32_t x1 = base + 1u < words ? sw[wi + 1u] : 0u;
        const uint32_t x2 = base + 2u < words ? sw[wi + 2u] : 0u;
#pragma unroll
        for (uint32_t j = 0; j < 4; j++) {
            const uint32_t o = m + j;                     /* byte offset from word wi: 0..6 */
            const uint32_t g = __funnelshift_r(o < 4u ? x0 : x1, o < 4u ? x1 : x2, (o & 3u) * 8u);
            const uint32_t q = q0 + j;
            if (q < p_hi) {
                const uint32_t x = (v0 + q) & (kK1WRing - 1);
                const uint32_t w = (q < n) ? g : 0u;
                W[x] = w;
                if (x < kK1WMirror) W[kK1WRing + x] = w;
            }
        }
    }
}
#endif

/* Queries of one tile, handed out in chunks of 32 positions (one per warp pass) from a counter in
 * shared memory, so that a warp that drew cheap positions takes more of them. */
__device__ __forceinline__ void k1_query_chunks(const K1Tile &d, uint32_t *next_chunk, const uint16_t *links,
                                                const uint16_t *runs, const uint32_t *W, match_t *mout)
{
    const uint32_t lane = lane_id();
    for (;;) {
        uint32_t c = 0;
        if (lane == 0) c = atomicAdd(next_chunk, 1u);
        c = __shfl_sync(LZS_FULL_MASK, c, 0);
        const uint32_t r = c * 32u + lane;
        if (c * 32u >= d.tile_n) break;
        if (r < d.tile_n) {
            const uint32_t i = d.t0 + r;
            /* a flow of equal packets as ONE stream: the look-ahead ends with the packet position i lies in
             * (what the reference does when a packet is flushed with add_end_marker), the history does not */
            const uint32_t n_eff = d.seg ? umin32(d.n, (i / d.seg + 1u) * d.seg) : d.n;
            if (i >= d.skip && i < d.n - d.look) mout[i - d.skip] = static_cast<match_t>(k1_query(links, runs, W, d.v0, i, n_eff));
        }
        __syncwarp();
    }
}

/* ctl[0]: stream counter of the fast launch, ctl[1]: of the safe launch, ctl[2]: set by the fast
 * launch when an exchange order was observed that the fast insert does not handle.  The safe
 * launch follows the fast one on the same stream and returns at once unless ctl[2] is set. */
template <bool kSafe>
__global__ void __launch_bounds__(kK1Threads, 1)
k1_match(const uint8_t *__restrict__ in, const uint64_t *__restrict__ in_off,
         const uint32_t *__restrict__ in_len, match_t *__restrict__ matches, uint32_t n_streams,
         uint32_t *__restrict__ ctl, const uint32_t *__restrict__ hist_len, const uint32_t *__restrict__ seg_len,
         const uint32_t *__restrict__ look_len = nullptr)
{
    if (kSafe && *reinterpret_cast<volatile uint32_t *>(ctl + 2) == 0u) return;
    uint32_t *next_stream = ctl + (kSafe ? 1 : 0);
    LZS_DYN_SMEM(uint8_t, smem);
    uint32_t *heads = reinterpret_cast<uint32_t *>(smem);
    uint16_t *links = reinterpret_cast<uint16_t *>(heads + kK1Levels * kK1Slots);
    uint16_t *runs = links + kK1Levels * kK1LinkRing;
    uint32_t *W = reinterpret_cast<uint32_t *>(runs + kK1LinkRing);
    __shared__ K1Tile   s_desc[8];           /* tile g is described in s_desc[g & 7]             */
    __shared__ uint64_t s_full[kK1Depth];    /* tile g built: one arrival per build thread, phase g / depth */
    __shared__ uint64_t s_empty[kK1Depth];   /* tile g queried: one arrival per query thread              */
    __shared__ uint32_t s_filled;            /* tiles whose grams and descriptor are in place     */
    __shared__ uint32_t s_qdone[kK1Depth];   /* query-warp completions per stage: tile q is done at 16 (q / depth + 1) in [q % depth] */
    __shared__ uint32_t s_qnext[16];         /* next chunk of tile g to query, in s_qnext[g & 15]  */
#if LZS_K1_BULK
    __shared__ __align__(128) uint8_t s_stage[2][kK1StageBytes];   /* input bytes of the tile being filled and the next */
    __shared__ uint64_t              s_stage_bar[2];
#endif

    const uint32_t tid = threadIdx.x;
    const uint32_t warp = tid >> 5;
    const uint32_t lane = tid & 31u;

    for (uint32_t x = tid; x < kK1Levels * kK1Slots; x += kK1Threads) heads[x] = 0;
    for (uint32_t x = tid; x < (kK1Levels + 1) * kK1LinkRing; x += kK1Threads) links[x] = 0;
    if (tid == 0) s_filled = 0;
    if (tid < kK1Depth) s_qdone[tid] = 0;
    if (tid < 16) s_qnext[tid] = 0;
    if (tid < kK1Depth) {
        mbar_init(&s_full[tid], kK1BuildThreads);
        mbar_init(&s_empty[tid], kK1QueryThreads);
    }
#if LZS_K1_BULK
    if (tid < 2) mbar_init(&s_stage_bar[tid], 1);
    if (tid == 0) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#endif
    __syncthreads();

    if (warp == static_cast<uint32_t>(kK1BuildWarps)) {
        /* ================= loader: streams -> tiles, 4-byte grams into the ring =================
         * Runs ahead of the build warps (up to depth + 1 tiles ahead of the slowest query warp,
         * which is what the gram ring holds), so the global-load latency is off everybody's
         * critical path and the build warps never have to meet each other. */
        uint32_t g = 0;                      /* tiles described so far                  */
        uint32_t vnext = 4096;               /* virtual position of the next stream     */
#if LZS_K1_BULK
        uint32_t stage_n = 0;                /* staging requests consumed so far: buffer and barrier phase */
#endif
        for (;;) {
            uint32_t sid = 0;
            if (lane == 0) sid = atomicAdd(next_stream, 1u);
            sid = __shfl_sync(LZS_FULL_MASK, sid, 0);
            if (sid >= n_streams) break;
            /* Flows with kept history (RFC 1974 style; the reference does not reset its history at an
             * end marker, lzs-compression.c:796-820): the hist bytes that precede the packet in memory
             * are the flow's earlier packets.  They go through the tables like any other position --
             * the stream the loader and the build warps see starts `hist` bytes early -- and only the
             * packet's own positions are matched (candidates may reach back into the history, the
             * look-ahead ends with the packet). */
            const uint32_t own = in_len[sid];
            const uint32_t hist = (hist_len != nullptr && own != 0u) ? umin32(hist_len[sid], kWindow) : 0u;
            /* Pieces of a long stream (k23_pieces.cuh): the piece is followed in memory by the rest of its
             * stream, and its last positions may match up to 11 bytes into it -- the `look` bytes behind the
             * piece are part of what the tables and the compares see, and their own positions are left
             * to the next piece. */
            const uint32_t look = (look_len != nullptr && own != 0u) ? umin32(look_len[sid], kSearchMax - 1u) : 0u;
            const uint32_t n = own + hist + look;
            const uint8_t *src = in + in_off[sid] - hist;
            /* last aligned word that holds a byte of the stream (n > 0 inside the tile loop) */
            const uintptr_t wlast = (reinterpret_cast<uintptr_t>(src) + (n ? n - 1u : 0u)) & ~static_cast<uintptr_t>(3);
            const uint32_t v0 = vnext;
            vnext = (v0 + n + 16u + 31u) & ~31u;
            if (n == 0) continue;            /* nothing to match, and no word of it may be touched */
            /* grams are kept kK1Ahead positions beyond the tile: 8 for the 12-byte compares plus the
             * 32 positions whose hashes the build warps prefetch in their last batch.  The words of
             * the NEXT tile are requested before this tile's grams are written, so the DRAM latency
             * is paid once per stream, not once per tile. */
#if LZS_K1_BULK
            /* the copy engine brings the tile's bytes into one of two staging buffers (a tile ahead,
             * like the word loads of the other variant); the loader turns them into grams */
            (void)wlast;
            K1Stage cur = k1_stage_range(src, 0u, umin32(kK1Tile + kK1Ahead, n + 12u), n);
            if (lane == 0) k1_stage_issue(s_stage[stage_n & 1u], &s_stage_bar[stage_n & 1u], cur);
#else
            K1Words cur;
            k1_load_words(cur, src, 0u, wlast);
#endif
            for (uint32_t t0 = 0; t0 < n; t0 += kK1Tile, g++) {
                const uint32_t p_lo = (t0 == 0) ? 0u : t0 + kK1Ahead;
                const uint32_t p_hi = umin32(t0 + kK1Tile + kK1Ahead, n + 12u);
#if LZS_K1_BULK
                const uint32_t t1 = t0 + kK1Tile;
                K1Stage        nxt = {0, 0};
                if (t1 < n) {
                    nxt = k1_stage_range(src, t1 + kK1Ahead, umin32(t1 + kK1Tile + kK1Ahead, n + 12u), n);
                    __syncwarp();
                    if (lane == 0) k1_stage_issue(s_stage[(stage_n + 1u) & 1u], &s_stage_bar[(stage_n + 1u) & 1u], nxt);
                }
#else
                K1Words nxt;
                k1_load_words(nxt, src, t0 + kK1Tile + kK1Ahead, wlast);
#endif
                if (g > kK1Depth) {          /* every query warp has left tile g - depth - 1 */
                    const uint32_t q = g - kK1Depth - 1u;
                    while (*reinterpret_cast<volatile uint32_t *>(&s_qdone[q & (kK1Depth - 1u)]) <
                           (q / kK1Depth + 1u) * kK1QueryWarps)
                        spin_pause();
                    __threadfence_block();
                }
                tl_set(g, 0);
#if LZS_K1_BULK
                mbar_wait(&s_stage_bar[stage_n & 1u], (stage_n >> 1) & 1u);
                k1_store_grams_staged(s_stage[stage_n & 1u], cur, W, src, v0, p_lo, p_hi, n);
                stage_n++;
                cur = nxt;
#else
                k1_store_grams(cur, W, src, v0, p_lo, p_hi, n);
                cur = nxt;
#endif
                if (lane == 0) {
                    K1Tile d;
                    d.sid = sid; d.t0 = t0; d.tile_n = umin32(kK1Tile, n - t0); d.n = n; d.v0 = v0; d.skip = hist;
                    d.seg = seg_len != nullptr ? seg_len[sid] : 0u;
                    d.look = look;
                    s_desc[g & 7u] = d;
                    s_qnext[g & 15u] = 0;    /* nobody can still be on tile g - 16 */
                }
                __threadfence_block();
                __syncwarp();
                if (lane == 0) *reinterpret_cast<volatile uint32_t *>(&s_filled) = g + 1u;
                tl_set(g, 1);
            }
        }
        if (g > kK1Depth) {
            const uint32_t q = g - kK1Depth - 1u;
            while (*reinterpret_cast<volatile uint32_t *>(&s_qdone[q & (kK1Depth - 1u)]) <
                   (q / kK1Depth + 1u) * kK1QueryWarps)
                spin_pause();
        }
        if (lane == 0) s_desc[g & 7u].sid = kK1EndOfWork;
        __threadfence_block();
        __syncwarp();
        if (lane == 0) *reinterpret_cast<volatile uint32_t *>(&s_filled) = g + 1u;
    } else if (warp < static_cast<uint32_t>(kK1BuildWarps)) {
        /* ================= build warps: one level each (the twelfth: the run table) =================
         * Independent of each other: a warp waits for the loader (flag) and for the query warps to
         * have left the tile that used this stage before (mbarrier phase), builds, and arrives on
         * the stage's "built" mbarrier. */
        /* byte masks of this warp's gram length k = warp + 2 */
        const uint32_t k = warp + 2u;
        const uint32_t m0 = k >= 4u ? 0xFFFFFFFFu : ((1u << (8u * k)) - 1u);
        const uint32_t m1 = k >= 8u ? 0xFFFFFFFFu : (k <= 4u ? 0u : ((1u << (8u * (k - 4u))) - 1u));
        const uint32_t m2 = k >= 12u ? 0xFFFFFFFFu : (k <= 8u ? 0u : ((1u << (8u * (k - 8u))) - 1u));
        uint32_t disorder = 0;
        for (uint32_t g = 0;; g++) {
            const uint32_t buf = g & (kK1Depth - 1u);
            while (*reinterpret_cast<volatile uint32_t *>(&s_filled) <= g) spin_pause();
            __threadfence_block();
            const K1Tile d = s_desc[g & 7u];
            /* the query warps must have left the tile that used this stage before (tile g - depth) */
            if (g >= kK1Depth) mbar_wait(&s_empty[buf], (g / kK1Depth - 1u) & 1u);
            tl_min(g, 2);
            tl_max(g, 3);
#if defined(LZS_K1_TIMELINE) && !defined(LZS_SIMT_EMU)
            const long long tb0 = clock64();
#endif
            if (d.sid != kK1EndOfWork && warp < static_cast<uint32_t>(LZS_K1_FAKE_BUILD_WARPS)) {
                const uint32_t vt = d.v0 + d.t0;
                if (kK1Lpw > 1 || kK1Grouped)
                    disorder |= k1_build_tile<kSafe>(warp, heads, links, runs, W, d);
                else if (warp < static_cast<uint32_t>(kK1Levels))
                    disorder |= k1_build_level<kSafe>(heads + warp * kK1Slots, links + warp * kK1LinkRing, W, vt,
                                                      d.tile_n, m0, m1, m2);
                else
                    k1_build_runs(runs, W, d.v0, d.t0, d.tile_n);
            }
            tl_min(g, 4);
            tl_max(g, 5);
#if defined(LZS_K1_TIMELINE) && !defined(LZS_SIMT_EMU)
            if (blockIdx.x == 0 && lane == 0) g_k1_warp_busy[warp][0] += static_cast<unsigned long long>(clock64() - tb0);
#endif
            mbar_arrive(&s_full[buf]);
            if (d.sid == kK1EndOfWork) break;
        }
        if (!kSafe && __any_sync(LZS_FULL_MASK, disorder != 0u) && lane == 0) atomicOr(ctl + 2, 1u);
    } else {
        /* ================= query warps: one query per position ================= */
        for (uint32_t g = 0;; g++) {
            const uint32_t buf = g & (kK1Depth - 1u);
            mbar_wait(&s_full[buf], (g / kK1Depth) & 1u);
            const K1Tile d = s_desc[g & 7u];
            if (d.sid == kK1EndOfWork) break;
            tl_min(g, 6);
            k1_query_chunks(d, &s_qnext[g & 15u], links, runs, W, matches + in_off[d.sid]);
            tl_max(g, 7);
            mbar_arrive(&s_empty[buf]);
            __syncwarp();
            if (lane == 0) atomicAdd(&s_qdone[buf], 1u);  /* the loader may reuse the ring behind us */
        }
    }
}

}  // namespace lzs

#endif /* LZS_B200_K1_MATCH_CUH */
