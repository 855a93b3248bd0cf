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
 *   P_k is "previous equal element" over the sequence of k-grams.  Every level k
 *   keeps a 2048-slot last-occurrence table (32-bit heads) in shared memory;
 *   positions are inserted in order, 32 at a time: one atomic exchange per lane puts
 *   the position into its slot and returns the predecessor (see k1_build_group for
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
 * Layout: one persistent CTA per SM (224 KiB of shared memory: 11 head tables,
 * 11 link rings, the run table, a ring of 4-byte grams, a ring of results), warps in
 * four roles that never meet each other -- every hand-off is a counter in shared
 * memory that only ever grows:
 *   - 1 loader warp pulls streams from a global counter, cuts them into 448-position
 *     tiles and fills the gram ring (aligned word loads, a tile ahead; shuffles and
 *     funnel shifts make the grams);
 *   - a few build warps, each owning a GROUP of consecutive levels: one set of gram
 *     loads per 32 positions, the hash rolled from level to level (two instructions),
 *     one exchange and one link store per level.  The last group also builds the
 *     run table;
 *   - query warps whose LANES are independent walkers: a lane that has finished its
 *     position takes the next unclaimed position of the oldest built tile at once,
 *     so a warp iteration (one chain step for every lane) is always full, whatever
 *     the spread of walk lengths (round 1 gave a warp 32 positions at a time and
 *     waited for the longest of the 32 walks: 7 of 32 lanes active on average);
 *   - 1 writer warp that copies the results of finished tiles from the result ring
 *     to global memory with 16-byte stores and releases the tile's ring space.
 * Positions are numbered continuously across the streams a CTA processes
 * ("virtual positions", every stream starting on a multiple of 32), so the rings
 * need no clearing between streams; a candidate is valid only if its distance does
 * not exceed the position inside the current stream.
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
constexpr uint32_t kK1LinkRing = 4096;      /* >= 2047 + depth * (tile + gap)      */
constexpr uint32_t kK1WRing = 8192;
constexpr uint32_t kK1WMirror = 64;         /* the first grams again behind the ring: index, +4, +8 need one wrap */
#ifndef LZS_K1_TILE
#define LZS_K1_TILE 448
#endif
#ifndef LZS_K1_DEPTH
#define LZS_K1_DEPTH 4
#endif
#ifndef LZS_K1_LPW
#define LZS_K1_LPW 3                        /* levels per build warp */
#endif
#ifndef LZS_K1_QW
#define LZS_K1_QW 20
#endif
constexpr uint32_t kK1Tile = LZS_K1_TILE;   /* a multiple of 32; 448 = 14 batches  */
constexpr uint32_t kK1Depth = LZS_K1_DEPTH; /* tiles the build warps may be ahead of the writer (power of two) */
constexpr int      kK1Lpw = LZS_K1_LPW;
/* Virtual positions between streams: 12 zero grams behind the last byte, then up to the next
 * multiple of 32 (every stream starts on a batch boundary, so a batch never straddles a ring
 * wrap and the positions a last batch inserts past the end of its stream belong to no stream). */
constexpr uint32_t kK1StreamGap = 16 + 31;
constexpr int      kK1BuildWarps = (kK1Levels + kK1Lpw - 1) / kK1Lpw;
constexpr int      kK1QueryWarps = LZS_K1_QW;
constexpr int      kK1LoaderWarp = kK1BuildWarps;
constexpr int      kK1WriterWarp = kK1BuildWarps + 1;
constexpr int      kK1FirstQueryWarp = kK1BuildWarps + 2;
constexpr int      kK1Threads = 32 * (kK1BuildWarps + 2 + kK1QueryWarps);
constexpr size_t   kK1SmemBytes = static_cast<size_t>(kK1Levels) * kK1Slots * 4 +        /* heads       */
                                static_cast<size_t>(kK1Levels + 2) * kK1LinkRing * 2 +  /* links, runs, results */
                                (kK1WRing + kK1WMirror) * 4;                            /* grams       */
/* run table entry: (forward run length capped at 12) << 12 | distance back to the run start */
constexpr uint32_t kRunBackMask = 0xFFFu;
static_assert((kK1Depth & (kK1Depth - 1)) == 0 && kK1Depth >= 2 && kK1Depth <= 4, "pipeline depth");
static_assert(kWindow + kK1Depth * (kK1Tile + kK1StreamGap) < kK1LinkRing, "link ring too small for the pipeline");
constexpr uint32_t kK1Ahead = 40;           /* grams filled beyond the tile being built */
static_assert(kWindow + (kK1Depth + 2) * (kK1Tile + kK1StreamGap) + 16 + kK1Ahead < kK1WRing, "gram ring: window + the tiles in the pipeline + those the loader may be ahead");
static_assert(kK1Tile % 32 == 0, "a tile is whole batches");

constexpr uint32_t kK1EndOfWork = 0xFFFFFFFFu;
constexpr uint32_t kK1ClaimSpan = 512;      /* claim counts per tile: a power of two >= the tile, a multiple of 32 */
static_assert(kK1ClaimSpan >= kK1Tile && (kK1ClaimSpan & (kK1ClaimSpan - 1)) == 0, "claim span");

struct K1Tile {
    uint32_t sid, t0, tile_n, n, v0, g;      /* g: the tile's index (descriptors live in a ring of 8) */
};

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
constexpr uint32_t kSlotShift = 21;          /* slot = h >> 21 (11 bits)                    */
constexpr uint32_t kLinkDistMask = 0x7FFu;   /* link entry: (tag << 11) | distance, tag = bits 20..16 of h */

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

/* A counter in shared memory that only grows: the hand-offs between the roles. */
__device__ __forceinline__ uint32_t flag_read(const uint32_t *p) { return *reinterpret_cast<const volatile uint32_t *>(p); }
__device__ __forceinline__ void     flag_wait_ge(const uint32_t *p, uint32_t want)
{
    while (static_cast<int32_t>(flag_read(p) - want) < 0) spin_pause();
    __threadfence_block();
}

/* Exact repair of one batch's exchanges, for the case that the lanes sharing a slot were not
 * served in ascending lane order (sm_100a serves them in ascending order --
 * tools/micro/atoms_exch.cu -- so this is insurance, exercised by the emulator tests and by a GPU
 * test that forces the safe launch).  Whatever the order, exactly one lane of every group received
 * the pre-batch head.  Returns the position each lane should have received and leaves the group's
 * highest lane in the head. */
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
#pragma unroll 2
    for (uint32_t b = 0; b < tile_n; b += 32) {
        const uint32_t vb = vt + b;                       /* warp-uniform, a multiple of 32 */
        const uint32_t pos = vb | lane;
        const uint32_t x = vb & (kK1WRing - 1);
        const uint32_t w0 = Wl[x], w1 = Wl[x + 4], w2 = Wl[x + 8];
        uint16_t      *lk = lk0 + (vb & (kK1LinkRing - 1));
        uint32_t       h = k1_hash_start(K0, w0, w1, w2);
#pragma unroll
        for (int l = 0; l < NL; l++) {
            if (l) h = k1_hash_roll(K0 + l, h, w0, w1, w2);
            uint32_t *slot = hd + l * kK1Slots + (h >> kSlotShift);
            uint32_t  old = smem_exch(slot, pos);
            if (kSafe) {
                __syncwarp();
                old = k1_relink(slot, h >> kSlotShift, old, vb, pos);
            }
            const uint32_t dist = pos - old;
            if (!kSafe) disorder |= dist;                 /* sign bit: received a LATER position, not the order assumed */
            uint32_t e = (h >> 5) & 0xF800u;              /* tag << 11 */
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
 * query hop over all of them at once (see k1_query_warp). */
__device__ __forceinline__ void k1_build_runs(uint16_t *runs, const uint32_t *W, uint32_t v0, uint32_t t0,
                                              uint32_t tile_n)
{
    const uint32_t  lane = lane_id();
    const uint32_t  vt = v0 + t0;                         /* a multiple of 32 */
    const uint32_t *Wl = W + lane;
    uint16_t       *rl = runs + lane;
    /* carried from batch to batch in registers: the byte before the batch and how far back its run starts */
    uint32_t prev_byte = (t0 == 0) ? 0x100u : (W[(vt - 1u) & (kK1WRing - 1)] & 0xFFu);
    uint32_t carry = (t0 == 0) ? 0u : (runs[(vt - 1u) & (kK1LinkRing - 1)] & kRunBackMask);
#pragma unroll 1
    for (uint32_t b = 0; b < tile_n; b += 32) {
        const uint32_t vb = vt + b;
        const uint32_t x = vb & (kK1WRing - 1);
        const uint32_t w0 = Wl[x], w1 = Wl[x + 4], w2 = Wl[x + 8];
        const uint32_t byte = w0 & 0xFFu;
        const uint32_t rep = byte * 0x01010101u;
        const uint32_t fwd = lcp12(w0, w1, w2, rep, rep, rep);                 /* 1..12 */
        uint32_t       before = __shfl_up_sync(LZS_FULL_MASK, byte, 1);
        if (lane == 0) before = prev_byte;
        const uint32_t starts = __ballot_sync(LZS_FULL_MASK, before != byte);  /* positions that begin a run */
        const uint32_t below = starts & ((2u << lane) - 1u);                   /* starts at or below my lane */
        const uint32_t back = below ? lane - (31u - static_cast<uint32_t>(__clz(static_cast<int>(below))))
                                    : umin32(carry + lane + 1u, kRunBackMask);  /* the run began in an earlier batch */
        rl[vb & (kK1LinkRing - 1)] = static_cast<uint16_t>((fwd << 12) | back);
        carry = __shfl_sync(LZS_FULL_MASK, back, 31);
        prev_byte = __shfl_sync(LZS_FULL_MASK, byte, 31);
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
#error "LZS_K1_LPW must be 2, 3 or 4"
#endif
    return dis;
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
        for (uint32_t j = 0; j < 4; j++) {
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

/* Shared control block of a CTA: every member only ever grows (or is written by one role). */
struct K1Ctl {
    K1Tile   desc[8];                 /* tile g is described in desc[g & 7]                              */
    uint32_t filled;                  /* loader: tiles whose grams and descriptor are in place            */
    uint32_t built[kK1Depth];         /* build warps: completions per stage; tile g is built at
                                         kK1BuildWarps * (g / depth + 1) in built[g % depth]            */
    uint32_t qdone[kK1Depth];         /* query lanes: finished positions per stage (cumulative)           */
    uint32_t written;                 /* writer: tiles whose results are in global memory                 */
    uint32_t qcursor;                 /* next unclaimed position, kK1ClaimSpan counts per tile            */
};

#if defined(LZS_SIMT_EMU) && defined(LZS_K1_DEBUG)
static K1Ctl *g_k1_ctl_dbg = nullptr;
static unsigned char g_k1_seen[4][70000];
#define LZS_K1_DBG_HAND(sid, i, g, c_lo, c_hi, r)                                                              \
    do {                                                                                                      \
        if (g_k1_seen[(sid) & 3][(i)])                                                                      \
            fprintf(stderr, "DUP sid %u i %u tile %u claim [%u,%u) r %u warp %u first: tile %u\n", (sid), (i), (g), (c_lo), (c_hi), (r), threadIdx.x >> 5, g_k1_seen[(sid) & 3][(i)] - 1); \
        g_k1_seen[(sid) & 3][(i)] = (g) + 1; \
    } while (0)
#else
#define LZS_K1_DBG_HAND(sid, i, g, c_lo, c_hi, r) ((void)0)
#endif

/* ---------------------------------------------------------------- query warps ----
 * Every lane is a walker with its own position.  One iteration of the warp is one chain step
 * for every lane: load the link entry `tot` positions back at the lane's level and the twelve
 * bytes there, compare.  A foreign entry (tag mismatch) or a candidate that is too short just
 * moves on along the chain; a verified candidate of length l answers every level up to l, so the
 * lane restarts at level l + 1 (its own entry there is the first link); the first level whose
 * chain ends without a verified candidate ends the search ("some candidate reaches k" is monotone
 * in k).  Lanes without a position take the next unclaimed position of the oldest tile that has
 * unclaimed positions (claimed 32 at a time by the warp, handed to lanes as they fall idle), so
 * the lanes stay busy across tiles and streams.  A warp never blocks while one of its lanes still
 * holds a position: the tile that lane belongs to may be the one the build warps are waiting for. */
__device__ __forceinline__ void k1_query_warp(K1Ctl *ctl, const uint16_t *links, const uint16_t *runs,
                                              const uint32_t *W, uint16_t *res)
{
    const uint32_t lane = lane_id();
    const uint32_t lt = (1u << lane) - 1u;
    /* warp-uniform claim state: positions are claimed 32 at a time from ONE counter that runs
     * through all tiles (kK1ClaimSpan counts per tile, the tail of a shorter tile is void), so a
     * claim can never land on a recycled per-tile counter */
    uint32_t g = 0;                          /* tile of the current claim                      */
    uint32_t c_lo = 0, c_hi = 0;             /* unhanded positions [c_lo, c_hi) of the claim   */
    bool     claimed = false, have = false, eow = false;
    K1Tile   d = {0, 0, 0, 0, 0, 0};
    /* per-lane walker */
    bool            active = false;
    uint32_t        v = 0, k = 0, tot = 0, tag = 0, best = 0, bd = 0, M = 0, maxd = 0, buf = 0;
    uint32_t        w0 = 0, w1 = 0, w2 = 0;
    const uint16_t *lk = links;

    for (;;) {
        bool fin = false;
        /* ---- hand positions to idle lanes ---- */
        const uint32_t idle = __ballot_sync(LZS_FULL_MASK, !active);
        if (idle != 0u && !eow) {
            /* a claim whose tile is not built yet is kept; the warp never blocks on it while a
             * lane of this warp is walking (that lane's tile may be the one everybody waits for) */
            while (c_lo >= c_hi) {
                if (!claimed) {
                    uint32_t q = 0;
                    if (lane == 0) q = atomicAdd(&ctl->qcursor, 32u);
                    q = __shfl_sync(LZS_FULL_MASK, q, 0);
                    const uint32_t gq = q / kK1ClaimSpan;
                    have = have && gq == g;
                    g = gq;
                    c_lo = c_hi = q % kK1ClaimSpan;                      /* nothing to hand out until the tile is known */
                    claimed = true;
                }
                if (!have) {
                    const uint32_t want = static_cast<uint32_t>(kK1BuildWarps) * (g / kK1Depth + 1u);
                    /* a vote, so that the 32 lanes take the same branch even if they read the counter at
                     * different moments */
                    if (!__all_sync(LZS_FULL_MASK, static_cast<int32_t>(flag_read(&ctl->built[g & (kK1Depth - 1u)]) - want) >= 0)) {
                        if (idle != LZS_FULL_MASK) break;                /* walkers pending: do a step instead */
                        spin_pause();
                        continue;
                    }
                    __threadfence_block();
                    d = ctl->desc[g & 7u];
                    have = true;
                }
                claimed = false;
                /* The descriptor ring has moved on: every position of tile g was finished long ago,
                 * so this claim lies in the void tail of a short tile. */
                if (d.g != g) {
                    have = false;
                    continue;
                }
                if (d.sid == kK1EndOfWork) {
                    eow = true;
                    break;
                }
                c_hi = umin32(c_lo + 32u, d.tile_n);                     /* void when the tile is shorter */
            }
            if (c_lo < c_hi && !eow) {
                const uint32_t r = c_lo + static_cast<uint32_t>(__popc(idle & lt));
                if (!active && r < c_hi) {
                    const uint32_t i = d.t0 + r;
                    LZS_K1_DBG_HAND(d.sid, i, g, c_lo, c_hi, r);
                    v = d.v0 + i;
                    M = umin32(kSearchMax, d.n - i);
                    maxd = umin32(kWindow, i);
                    buf = g & (kK1Depth - 1u);
                    const uint32_t *wv = W + (v & (kK1WRing - 1));       /* the mirror covers +4 and +8 */
                    w0 = wv[0]; w1 = wv[4]; w2 = wv[8];
                    k = kMinLen;
                    lk = links;
                    const uint32_t e = lk[v & (kK1LinkRing - 1)];       /* own entry: tag + first link */
                    tag = e >> 11;
                    tot = e & kLinkDistMask;
                    best = 0;
                    bd = 0;
                    active = true;
                    fin = (M < kMinLen) || tot == 0u || tot > maxd;     /* nothing to look for / level 2 is empty */
                }
                c_lo = umin32(c_hi, c_lo + static_cast<uint32_t>(__popc(idle)));
            }
        }
        if (__all_sync(LZS_FULL_MASK, !active)) {
            if (eow) break;
            continue;
        }

        /* ---- one chain step ---- */
        if (active && !fin) {
            const uint32_t  j = v - tot;
            const uint32_t  e = lk[j & (kK1LinkRing - 1)];
            const uint32_t *wj = W + (j & (kK1WRing - 1));
            const uint32_t  x0 = wj[0], x1 = wj[4], x2 = wj[8];
            uint32_t        dn = e & kLinkDistMask;
            if ((e >> 11) == tag) {
                const uint32_t l = umin32(lcp12(w0, w1, w2, x0, x1, x2), M);
                if (l >= k) {
                    best = l;                                            /* nearest candidate of length l */
                    bd = tot;
                    if (l >= M) {
                        fin = true;
                    } else {
                        k = l + 1u;                                      /* next level to try */
                        lk = links + (k - 2u) * kK1LinkRing;
                        const uint32_t e2 = lk[v & (kK1LinkRing - 1)];
                        tag = e2 >> 11;
                        dn = e2 & kLinkDistMask;
                        tot = 0;
                    }
                }
            } else if (dn == 1u) {
                /* a foreign entry whose predecessor is the adjacent position: if j sits inside a run
                 * of one byte value that covers k bytes from j, every run member before j has j's
                 * gram (not ours) and is the next entry of this chain -- skip to the run start */
                const uint32_t r = runs[j & (kK1LinkRing - 1)];
                const uint32_t back = r & kRunBackMask;
                if ((r >> 12) >= k && back != 0u) {
                    tot += back;
                    dn = lk[(j - back) & (kK1LinkRing - 1)] & kLinkDistMask;
                }
            }
            if (!fin) {
                tot += dn;
                fin = (dn == 0u) || tot > maxd;                          /* level k has no (further) candidate */
            }
        }

        /* ---- retire finished lanes: result into the ring, positions counted per stage ---- */
        uint32_t rem = __ballot_sync(LZS_FULL_MASK, fin);
        if (rem != 0u) {
            if (fin) {
                res[v & (kK1LinkRing - 1)] = static_cast<uint16_t>((best << kMatchOffBits) | bd);
                active = false;
            }
            __syncwarp();
            while (rem != 0u) {
                const int      leader = __ffs(static_cast<int>(rem)) - 1;
                const uint32_t b0 = __shfl_sync(LZS_FULL_MASK, buf, leader);
                const uint32_t same = __ballot_sync(LZS_FULL_MASK, fin && buf == b0);
                if (lane == static_cast<uint32_t>(leader)) {
                    __threadfence_block();
                    atomicAdd(&ctl->qdone[b0], static_cast<uint32_t>(__popc(same)));
                }
                rem &= ~same;
            }
        }
    }
}

/* ctl[0]: stream counter of the fast launch, ctl[1]: of the safe launch, ctl[2]: set by the fast
 * launch when an exchange order was observed that the fast insert does not handle.  The safe
 * launch follows the fast one on the same stream and returns at once unless ctl[2] is set. */
template <bool kSafe>
__global__ void __launch_bounds__(kK1Threads, 1)
k1_match(const uint8_t *__restrict__ in, const uint64_t *__restrict__ in_off,
         const uint32_t *__restrict__ in_len, match_t *__restrict__ matches, uint32_t n_streams,
         uint32_t *__restrict__ ctl)
{
    if (kSafe && *reinterpret_cast<volatile uint32_t *>(ctl + 2) == 0u) return;
    uint32_t *next_stream = ctl + (kSafe ? 1 : 0);
    LZS_DYN_SMEM(uint8_t, smem);
    uint32_t *heads = reinterpret_cast<uint32_t *>(smem);
    uint16_t *links = reinterpret_cast<uint16_t *>(heads + kK1Levels * kK1Slots);
    uint16_t *runs = links + kK1Levels * kK1LinkRing;
    uint16_t *res = runs + kK1LinkRing;
    uint32_t *W = reinterpret_cast<uint32_t *>(res + kK1LinkRing);
    __shared__ K1Ctl s_ctl;
#if defined(LZS_SIMT_EMU) && defined(LZS_K1_DEBUG)
    g_k1_ctl_dbg = &s_ctl;
#endif

    const uint32_t tid = threadIdx.x;
    const uint32_t warp = tid >> 5;
    const uint32_t lane = tid & 31u;

    for (uint32_t x = tid; x < kK1Levels * kK1Slots; x += kK1Threads) heads[x] = 0;
    for (uint32_t x = tid; x < (kK1Levels + 2) * kK1LinkRing; x += kK1Threads) links[x] = 0;
    for (uint32_t x = tid; x < sizeof(K1Ctl) / 4; x += kK1Threads) reinterpret_cast<uint32_t *>(&s_ctl)[x] = 0;
    __syncthreads();

    if (warp == static_cast<uint32_t>(kK1LoaderWarp)) {
        /* ================= loader: streams -> tiles, 4-byte grams into the ring =================
         * Runs ahead of the build warps (up to depth + 1 tiles ahead of the writer, which is what
         * the gram ring holds), so the global-load latency is off everybody's critical path. */
        uint32_t g = 0;                      /* tiles described so far                  */
        uint32_t vnext = 4096;               /* virtual position of the next stream     */
        for (;;) {
            uint32_t sid = 0;
            if (lane == 0) sid = atomicAdd(next_stream, 1u);
            sid = __shfl_sync(LZS_FULL_MASK, sid, 0);
            if (sid >= n_streams) break;
            const uint32_t n = in_len[sid];
            const uint8_t *src = in + in_off[sid];
            /* last aligned word that holds a byte of the stream (n > 0 inside the tile loop) */
            const uintptr_t wlast = (reinterpret_cast<uintptr_t>(src) + (n ? n - 1u : 0u)) & ~static_cast<uintptr_t>(3);
            const uint32_t v0 = vnext;
            vnext = (v0 + n + 16u + 31u) & ~31u;
            if (n == 0) continue;            /* nothing to match, and no word of it may be touched */
            /* grams are kept kK1Ahead positions beyond the tile: 8 for the 12-byte compares plus
             * slack.  The words of the NEXT tile are requested before this tile's grams are
             * written, so the DRAM latency is paid once per stream, not once per tile. */
            K1Words cur;
            k1_load_words(cur, src, 0u, wlast);
            for (uint32_t t0 = 0; t0 < n; t0 += kK1Tile, g++) {
                const uint32_t p_lo = (t0 == 0) ? 0u : t0 + kK1Ahead;
                const uint32_t p_hi = umin32(t0 + kK1Tile + kK1Ahead, n + 12u);
                K1Words nxt;
                k1_load_words(nxt, src, t0 + kK1Tile + kK1Ahead, wlast);
                if (g > kK1Depth) flag_wait_ge(&s_ctl.written, g - kK1Depth);   /* tile g - depth - 1 has left the rings */
                k1_store_grams(cur, W, src, v0, p_lo, p_hi, n);
                cur = nxt;
                if (lane == 0) {
                    K1Tile d;
                    d.sid = sid; d.t0 = t0; d.tile_n = umin32(kK1Tile, n - t0); d.n = n; d.v0 = v0; d.g = g;
                    s_ctl.desc[g & 7u] = d;
                }
                __threadfence_block();
                __syncwarp();
                if (lane == 0) *reinterpret_cast<volatile uint32_t *>(&s_ctl.filled) = g + 1u;
            }
        }
        if (g > kK1Depth) flag_wait_ge(&s_ctl.written, g - kK1Depth);
        if (lane == 0) {
            s_ctl.desc[g & 7u].sid = kK1EndOfWork;
            s_ctl.desc[g & 7u].g = g;
        }
        __threadfence_block();
        __syncwarp();
        if (lane == 0) *reinterpret_cast<volatile uint32_t *>(&s_ctl.filled) = g + 1u;
    } else if (warp < static_cast<uint32_t>(kK1BuildWarps)) {
        /* ================= build warps: a group of levels each =================
         * Independent of each other: a warp waits for the loader (filled) and for the writer to
         * have released the tile that used this stage before (written), builds, and counts itself
         * in the stage's `built`. */
        uint32_t disorder = 0;
        for (uint32_t g = 0;; g++) {
            flag_wait_ge(&s_ctl.filled, g + 1u);
            const K1Tile d = s_ctl.desc[g & 7u];
            if (g >= kK1Depth) flag_wait_ge(&s_ctl.written, g - kK1Depth + 1u);
            if (d.sid != kK1EndOfWork) disorder |= k1_build_tile<kSafe>(warp, heads, links, runs, W, d);
            __threadfence_block();
            __syncwarp();
            if (lane == 0) atomicAdd(&s_ctl.built[g & (kK1Depth - 1u)], 1u);
            if (d.sid == kK1EndOfWork) break;
        }
        if (!kSafe && __any_sync(LZS_FULL_MASK, disorder != 0u) && lane == 0) atomicOr(ctl + 2, 1u);
    } else if (warp == static_cast<uint32_t>(kK1WriterWarp)) {
        /* ================= writer: results of finished tiles -> global memory =================
         * Tile g is finished when every one of its positions has been counted in its stage. */
        uint32_t expect[kK1Depth];
#pragma unroll
        for (uint32_t s = 0; s < kK1Depth; s++) expect[s] = 0;
        for (uint32_t g = 0;; g++) {
            flag_wait_ge(&s_ctl.filled, g + 1u);
            const K1Tile d = s_ctl.desc[g & 7u];
            if (d.sid == kK1EndOfWork) break;
            uint32_t want = 0;
#pragma unroll
            for (uint32_t s = 0; s < kK1Depth; s++)
                if (s == (g & (kK1Depth - 1u))) want = (expect[s] += d.tile_n);
            flag_wait_ge(&s_ctl.qdone[g & (kK1Depth - 1u)], want);
            match_t       *dst = matches + in_off[d.sid] + d.t0;
            const uint32_t vt = d.v0 + d.t0;
            const uint32_t vec = ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0u) ? (d.tile_n & ~7u) : 0u;
            for (uint32_t p = 8u * lane; p < vec; p += 256u)
                *reinterpret_cast<uint4 *>(dst + p) = *reinterpret_cast<const uint4 *>(res + ((vt + p) & (kK1LinkRing - 1)));
            for (uint32_t p = vec + lane; p < d.tile_n; p += 32u) dst[p] = res[(vt + p) & (kK1LinkRing - 1)];
            __threadfence_block();
            __syncwarp();
            if (lane == 0) *reinterpret_cast<volatile uint32_t *>(&s_ctl.written) = g + 1u;
        }
    } else {
        k1_query_warp(&s_ctl, links, runs, W, res);
    }
}

}  // namespace lzs

#endif /* LZS_B200_K1_MATCH_CUH */
