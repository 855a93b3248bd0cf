/*
 * lzs_b200.cu -- the C ABI (include/lzs_b200.h and the single-call part of
 * include/lzs.h) over the sm_100a kernels K1..K4.
 *
 * Host code here is launch plumbing only: argument checks, scratch carving,
 * stream-ordered launches, host<->device copies for the host-pointer entry points.
 * All codec work happens in the kernels; there is no CPU implementation behind any
 * entry point, and a missing device is reported as an error, never papered over.
 */
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/lzs.h"
#include "../../include/lzs_b200.h"
#include "corpus.h"
#include "k1_match.cuh"
#include "k23_parse_pack.cuh"
#include "k23_pieces.cuh"
#include "k4_decode.cuh"
#include "k4_pieces.cuh"

namespace {

thread_local char          g_err[512] = "";
std::atomic<uint64_t>      g_launches{0};
std::atomic<int>           g_decode_lanes{0};
std::atomic<int>           g_force_safe_match{0};
std::atomic<int>           g_zero_copy_out{-1};   /* -1: read LZS_B200_ZEROCOPY on first use */
std::atomic<int64_t>       g_piece_bytes{-1};     /* -1: read LZS_B200_PIECE on first use; 0: long streams are never cut */
std::atomic<int64_t>       g_dpiece_bytes{-1};    /* the same for the decoder (LZS_B200_DPIECE), compressed bytes per piece */

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

#define CUDA_TRY(expr)                                                                         \
    do {                                                                                       \
        cudaError_t e_ = (expr);                                                               \
        if (e_ != cudaSuccess)                                                                 \
            return fail(e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver           \
                            ? LZS_B200_ENODEVICE : LZS_B200_ECUDA,                             \
                        "%s failed: %s", #expr, cudaGetErrorString(e_));                       \
    } while (0)

struct DeviceInfo {
    int  sms = 0;
    bool ok = false;
    int  dec_blocks[4] = {0, 0, 0, 0};   /* resident K4 blocks per SM for G = 4, 8, 16, 32 */
};

/* per-device one-time setup: opt in to large dynamic shared memory, query occupancy */
int device_info(DeviceInfo **out)
{
    static DeviceInfo info[64];
    static std::mutex mu;
    int               dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return fail(LZS_B200_EINVAL, "device ordinal %d out of range", dev);
    std::lock_guard<std::mutex> lock(mu);
    DeviceInfo &d = info[dev];
    if (!d.ok) {
        CUDA_TRY(cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev));
        CUDA_TRY(cudaFuncSetAttribute(lzs::k1_match<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(lzs::kK1SmemBytes)));
        CUDA_TRY(cudaFuncSetAttribute(lzs::k1_match<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(lzs::kK1SmemBytes)));
        CUDA_TRY(cudaFuncSetAttribute(lzs::k4_decode<4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(lzs::k4_smem_bytes<4>())));
        CUDA_TRY(cudaFuncSetAttribute(lzs::k4_decode<8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(lzs::k4_smem_bytes<8>())));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d.dec_blocks[0], lzs::k4_decode<4>,
                                                               lzs::kDecThreads, lzs::k4_smem_bytes<4>()));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d.dec_blocks[1], lzs::k4_decode<8>,
                                                               lzs::kDecThreads, lzs::k4_smem_bytes<8>()));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d.dec_blocks[2], lzs::k4_decode<16>,
                                                               lzs::kDecThreads, lzs::k4_smem_bytes<16>()));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d.dec_blocks[3], lzs::k4_decode<32>,
                                                               lzs::kDecThreads, lzs::k4_smem_bytes<32>()));
        d.ok = true;
    }
    *out = &d;
    return LZS_B200_OK;
}

int decode_lanes()
{
    int g = g_decode_lanes.load();
    if (g == 0) {
        const char *e = getenv("LZS_B200_DECODE_LANES");
        g = e ? atoi(e) : 8;
        if (g != 4 && g != 8 && g != 16 && g != 32) g = 8;
        g_decode_lanes.store(g);
    }
    return g;
}

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

/* LZS_B200_ORDER=0 turns the decoder's launch order off (streams in index order), for measurements */
bool order_streams()
{
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("LZS_B200_ORDER");
        v = e ? atoi(e) : 1;
    }
    return v != 0;
}

constexpr size_t kCounterBytes = 256;    /* work counters live at the start of scratch */

/* Long streams are cut into pieces of this many bytes for the match finder and the parse/pack
 * kernels (k23_pieces.cuh) when the streams of a batch average two pieces or more -- fewer streams
 * than that do not fill the GPU with one warp each.  LZS_B200_PIECE sets it (0: never cut). */
uint32_t piece_bytes()
{
    int64_t v = g_piece_bytes.load();
    if (v < 0) {
        const char *e = getenv("LZS_B200_PIECE");
        v = e ? atoll(e) : 65536;
        if (v < 0 || v > (1ll << 28)) v = 65536;
        if (v != 0 && v < 64) v = 64;
        g_piece_bytes.store(v);
    }
    return static_cast<uint32_t>(v);
}
/* entries of the piece table for a batch over in_span bytes: in_span / P pieces, one more per
 * stream, and a batch is only cut when it has at most in_span / 2P streams */
uint32_t piece_table_entries(uint64_t in_span, uint32_t piece)
{
    const uint64_t e = in_span / piece + in_span / (2ull * piece) + 16u;
    return e > 0x7FFFFFFFull ? 0x7FFFFFFFu : static_cast<uint32_t>(e);
}
/* The decoder's pieces (k4_pieces.cuh) are pieces of the COMPRESSED stream.  LZS_B200_DPIECE sets their
 * size (0: long streams are decoded by one group of lanes each, as short ones are). */
uint32_t dpiece_bytes()
{
    int64_t v = g_dpiece_bytes.load();
    if (v < 0) {
        const char *e = getenv("LZS_B200_DPIECE");
        v = e ? atoll(e) : 2048;                         /* measured: 2048 / 4096 / 8192 -> 25.3 / 27.6 / 35.9 ms per GiB in 1 MiB chunks */
        if (v < 0 || v > (1ll << 24)) v = 2048;
        if (v != 0 && v < 16) v = 16;
        g_dpiece_bytes.store(v);
    }
    return static_cast<uint32_t>(v);
}
constexpr uint32_t kCutStreamsMaxDecode = 4096;
/* ... when the output is at least this long (LZS_B200_JUMP_MIN) */
uint64_t jump_bytes_min()
{
    static long long v = -1;
    if (v < 0) {
        const char *e = getenv("LZS_B200_JUMP_MIN");
        v = e ? atoll(e) : 16384;    /* measured, one lzs_decompress call: 32 KiB .. 512 KiB take 1.2 .. 2.4 ms this way, 1.3 .. 5.3 ms
                                        with the replay, 2.2 .. 34 ms with one group of lanes (tools/small_calls_probe.py) */
        if (v < 4096) v = 4096;
    }
    return static_cast<uint64_t>(v);
}
/* repair passes of k4p_fix after the first (LZS_B200_FIX_REPAIRS): a piece whose guess never joined its true
 * orbit gives the NEXT piece a wrong entry, and runs of such pieces are repaired one piece per pass */
int fix_repairs()
{
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("LZS_B200_FIX_REPAIRS");
        v = e ? atoi(e) : 3;         /* measured on 4 streams of 256 MiB: 1 / 2 / 4 passes leave 631 / 46 / 0 of 361 283 pieces
                                        with a wrong entry, each of which costs the sweep a serial parse (34 ms / 3 / 1) */
        if (v < 1) v = 1;
    }
    return v;
}
/* at most this many long streams are decoded by pointer doubling (host calls; LZS_B200_JUMP_STREAMS, 0 = never) */
uint32_t jump_streams_max()
{
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("LZS_B200_JUMP_STREAMS");
        v = e ? atoi(e) : 128;       /* measured: doubling beats the replay below ~150 equal streams (1 GiB: 45 ms whatever their number,
                                        against 7.6 s / streams for the replay) */
        if (v < 0) v = 0;
    }
    return static_cast<uint32_t>(v);
}
size_t order_bytes(uint32_t n_streams) { return align_up(static_cast<size_t>(n_streams) * sizeof(uint32_t), 256); }

size_t matches_bytes(uint64_t in_span) { return align_up(static_cast<size_t>(in_span) * sizeof(lzs::match_t) + 64, 256); }
/* Cut the streams of this batch?  Decided from what the host knows without looking at the lengths:
 * few streams (one warp per stream fills the parse kernel from ~2000 streams on: measured, 1 GiB in
 * 512 KiB chunks is where cut and uncut meet, profiles/r2_pieces_bench.jsonl) that average two
 * pieces or more, and scratch with room for the piece table.  A batch of many short streams with a
 * long one among them is not cut.  Returns the table's entries, 0 = do not cut. */
constexpr uint32_t kCutStreamsMax = 2048;
uint32_t cut_into_pieces(uint32_t n_streams, uint64_t in_span, size_t scratch_bytes, uint32_t *piece_out)
{
    const uint32_t piece = piece_bytes();
    if (piece == 0 || n_streams > kCutStreamsMax || static_cast<uint64_t>(n_streams) * 2u * piece > in_span) return 0;
    const uint32_t cap = piece_table_entries(in_span, piece);
    if (scratch_bytes < kCounterBytes + matches_bytes(in_span) + lzs::piece_table_bytes(cap)) return 0;
    *piece_out = piece;
    return cap;
}

__global__ void corpus_fill_kernel(uint8_t *dst, uint64_t stride, uint32_t stream_len, uint64_t first_index,
                                   uint64_t n, uint64_t seed, int kind)
{
    const uint64_t s = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (s < n) lzs_corpus_fill(dst + s * stride, stream_len, seed, first_index + s, kind);
}

/* Copy n streams from their slots to packed positions; both offsets are multiples of 16. */
__global__ void gather_streams_kernel(const uint8_t *__restrict__ src, const uint64_t *__restrict__ src_off,
                                      const uint32_t *__restrict__ len, uint8_t *__restrict__ dst,
                                      const uint64_t *__restrict__ dst_off, uint32_t n)
{
    const uint32_t s = blockIdx.x;
    if (s >= n) return;
    const uint4   *from = reinterpret_cast<const uint4 *>(src + src_off[s]);
    uint4         *to = reinterpret_cast<uint4 *>(dst + dst_off[s]);
    const uint32_t vecs = (len[s] + 15u) >> 4;
    for (uint32_t i = threadIdx.x; i < vecs; i += blockDim.x) to[i] = from[i];
}

/* The pack kernel fused with the all-gather of the packed streams: every stream is read once from
 * this GPU's slots and written into the gathered buffer of EVERY rank of the box -- the buffers are
 * one symmetric allocation, dst[r] the address of rank r's copy as THIS GPU sees it over NVLink
 * (peer memory) -- at this rank's place in it (base) plus the stream's packed offset.  No staging
 * copy and no separate collective: the NVLink transfer is the store. */
__global__ void scatter_streams_peers_kernel(const uint8_t *__restrict__ src, const uint64_t *__restrict__ src_off,
                                             const uint32_t *__restrict__ len, uint8_t *const *__restrict__ dst,
                                             uint32_t n_dst, uint64_t base, const uint64_t *__restrict__ dst_off, uint32_t n)
{
    const uint32_t s = blockIdx.x;
    if (s >= n) return;
    const uint4   *from = reinterpret_cast<const uint4 *>(src + src_off[s]);
    const uint64_t at = base + dst_off[s];
    const uint32_t vecs = (len[s] + 15u) >> 4;
    /* every block starts with a different rank's copy, so that the GPUs of the box do not all store
     * into the same GPU at the same moment */
    const uint32_t d0 = (s + static_cast<uint32_t>(base >> 8)) % n_dst;
    for (uint32_t i = threadIdx.x; i < vecs; i += blockDim.x) {
        const uint4 v = from[i];
        for (uint32_t k = 0; k < n_dst; k++) {
            const uint32_t d = d0 + k < n_dst ? d0 + k : d0 + k - n_dst;
            reinterpret_cast<uint4 *>(dst[d] + at)[i] = v;
        }
    }
}

/* The same through the NVSwitch multicast address of the symmetric buffer: ONE store per 16 bytes,
 * replicated to all ranks inside the switch (multimem.st), so the sending GPU's NVLink carries every
 * byte once instead of once per peer. */
__global__ void scatter_streams_multicast_kernel(const uint8_t *__restrict__ src, const uint64_t *__restrict__ src_off,
                                                 const uint32_t *__restrict__ len, uint8_t *__restrict__ mc, uint64_t base,
                                                 const uint64_t *__restrict__ dst_off, uint32_t n)
{
    const uint32_t s = blockIdx.x;
    if (s >= n) return;
    const uint4   *from = reinterpret_cast<const uint4 *>(src + src_off[s]);
    uint8_t       *to = mc + base + dst_off[s];
    const uint32_t vecs = (len[s] + 15u) >> 4;
    for (uint32_t i = threadIdx.x; i < vecs; i += blockDim.x) {
        const uint4 v = from[i];
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(to + 16ull * i),
                     "f"(__uint_as_float(v.x)), "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)),
                     "f"(__uint_as_float(v.w))
                     : "memory");
    }
}

}  // namespace

extern "C" {

const char *lzs_b200_last_error(void) { return g_err; }

int lzs_b200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

uint64_t lzs_b200_kernel_launches(void) { return g_launches.load(); }

int lzs_b200_set_decode_lanes(int lanes)
{
    if (lanes != 4 && lanes != 8 && lanes != 16 && lanes != 32)
        return fail(LZS_B200_EINVAL, "decode lanes must be 4, 8, 16 or 32 (got %d)", lanes);
    g_decode_lanes.store(lanes);
    return LZS_B200_OK;
}

#ifdef LZS_K1_TIMELINE
/* debug builds only: reset / fetch the per-tile timeline of K1's CTA 0 (tools/k1_timeline.py) */
int lzs_b200_debug_timeline(unsigned long long *dst, int reset)
{
    static std::vector<unsigned long long> init;
    const size_t n = static_cast<size_t>(lzs::kTlTiles) * 8;
    if (reset) {
        init.assign(n, 0ull);
        for (size_t i = 0; i < n; i++)
            if ((i & 7) == 2 || (i & 7) == 4 || (i & 7) == 6) init[i] = ~0ull;      /* the atomicMin slots */
        return cudaMemcpyToSymbol(lzs::g_k1_timeline, init.data(), n * 8) == cudaSuccess ? 0 : -2;
    }
    return cudaMemcpyFromSymbol(dst, lzs::g_k1_timeline, n * 8) == cudaSuccess ? 0 : -2;
}
int lzs_b200_debug_warp_busy(unsigned long long *dst, int reset)
{
    static unsigned long long zero[128];
    if (reset) return cudaMemcpyToSymbol(lzs::g_k1_warp_busy, zero, sizeof zero) == cudaSuccess ? 0 : -2;
    return cudaMemcpyFromSymbol(dst, lzs::g_k1_warp_busy, 128 * 8) == cudaSuccess ? 0 : -2;
}
#endif

int lzs_b200_set_zero_copy_output(int on)
{
    g_zero_copy_out.store(on ? 1 : 0);
    return LZS_B200_OK;
}

int lzs_b200_set_force_safe_match(int on)
{
    g_force_safe_match.store(on ? 1 : 0);
    return LZS_B200_OK;
}

size_t lzs_b200_compress_scratch_bytes(uint64_t in_span)
{
    const uint32_t piece = piece_bytes();
    const size_t   table = piece ? align_up(lzs::piece_table_bytes(piece_table_entries(in_span, piece)), 256) : 0;
    return kCounterBytes + matches_bytes(in_span) + table;
}

int lzs_b200_set_piece_bytes(uint32_t bytes)
{
    if (bytes != 0 && (bytes < 64 || bytes > (1u << 28))) return fail(LZS_B200_EINVAL, "piece size must be 0 or 64 .. 2^28");
    g_piece_bytes.store(bytes);
    return LZS_B200_OK;
}

size_t lzs_b200_decompress_scratch_bytes(void) { return kCounterBytes; }

/* with room for the launch order of n streams (see k4_decode.cuh) */
size_t lzs_b200_decompress_scratch_bytes_for(uint32_t n_streams)
{
    return kCounterBytes + order_bytes(n_streams);
}

/* with room for the piece table of a batch of few long streams (k4_pieces.cuh): in_span compressed bytes */
size_t lzs_b200_decompress_scratch_bytes_long(uint64_t in_span, uint32_t n_streams)
{
    const uint32_t piece = dpiece_bytes();
    if (piece == 0 || n_streams > kCutStreamsMaxDecode) return lzs_b200_decompress_scratch_bytes_for(n_streams);
    const uint64_t cap = in_span / piece + 2ull * n_streams + 16u;
    return kCounterBytes + order_bytes(n_streams) + align_up(lzs::dpiece_table_bytes(static_cast<uint32_t>(cap > 0x7FFFFFFFull ? 0x7FFFFFFFull : cap), piece), 256);
}

/* ... and for the pointers of lzs_b200_decompress_long_batch_device: 4 bytes per byte of out_span more */
size_t lzs_b200_decompress_scratch_bytes_jump(uint64_t in_span, uint64_t out_span, uint32_t n_streams)
{
    return lzs_b200_decompress_scratch_bytes_long(in_span, n_streams) + align_up(static_cast<size_t>(out_span) * 4u, 256) + 4096;
}

int lzs_b200_set_decode_piece_bytes(uint32_t bytes)
{
    if (bytes != 0 && (bytes < 16 || bytes > (1u << 24))) return fail(LZS_B200_EINVAL, "piece size must be 0 or 16 .. 2^24");
    g_dpiece_bytes.store(bytes);
    return LZS_B200_OK;
}

}  // extern "C"

namespace {
int match_batch(const uint8_t *in, const uint64_t *in_off, const uint32_t *in_len, const uint32_t *hist_len,
                uint16_t *matches, uint32_t n_streams, uint32_t *counter, void *stream, const uint32_t *seg_len = nullptr,
                const uint32_t *look_len = nullptr)
{
    if (n_streams == 0) return LZS_B200_OK;
    if (!in || !in_off || !in_len || !matches || !counter) return fail(LZS_B200_EINVAL, "null pointer");
    DeviceInfo *d = nullptr;
    int         rc = device_info(&d);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CUDA_TRY(cudaMemsetAsync(counter, 0, 4 * sizeof(uint32_t), st));
    /* test knob: pretend the fast launch saw an exchange order it does not handle, so that the
     * exact-for-any-order launch really runs (and overwrites every record) on this hardware */
    if (g_force_safe_match.load()) CUDA_TRY(cudaMemsetAsync(counter + 2, 1, 1, st));
    const unsigned grid = n_streams < static_cast<uint32_t>(d->sms) ? n_streams : static_cast<unsigned>(d->sms);
    lzs::k1_match<false><<<grid, lzs::kK1Threads, lzs::kK1SmemBytes, st>>>(in, in_off, in_len, matches, n_streams,
                                                                          counter, hist_len, seg_len, look_len);
    /* the exact-for-any-hardware variant: returns at once unless the fast launch saw an exchange
     * order it does not handle (never on sm_100a); on the stream, so nothing waits on the host */
    lzs::k1_match<true><<<grid, lzs::kK1Threads, lzs::kK1SmemBytes, st>>>(in, in_off, in_len, matches, n_streams,
                                                                         counter, hist_len, seg_len, look_len);
    g_launches += 2;
    CUDA_TRY(cudaGetLastError());
    return LZS_B200_OK;
}

/* The compressor for few long streams: plan, K1 over the pieces, then the four passes of
 * k23_pieces.cuh.  All on the stream; nothing comes back to the host. */
int compress_pieces(const uint8_t *in, const uint64_t *in_off, const uint32_t *in_len, uint8_t *out,
                    const uint64_t *out_off, const uint32_t *out_cap, uint32_t *out_len, uint32_t n_streams,
                    uint32_t piece, uint32_t cap, uint32_t *counter, uint16_t *matches, void *table_mem, void *stream)
{
    if (!in || !in_off || !in_len || !out || !out_off || !out_cap || !out_len) return fail(LZS_B200_EINVAL, "null pointer");
    cudaStream_t          st = static_cast<cudaStream_t>(stream);
    const lzs::PieceTable t = lzs::piece_table_at(table_mem, cap);
    CUDA_TRY(cudaMemsetAsync(t.count, 0, 256, st));
    CUDA_TRY(cudaMemsetAsync(t.len, 0, static_cast<size_t>(cap) * sizeof(uint32_t), st));
    lzs::k23p_plan_count<<<1, lzs::kPlanThreads, 0, st>>>(in_len, n_streams, piece, t);
    lzs::k23p_plan_fill<<<n_streams, 128, 0, st>>>(in_off, in_len, out_len, n_streams, piece, t);
    g_launches += 2;
    CUDA_TRY(cudaGetLastError());
    int rc = match_batch(in, t.off, t.len, t.hist, matches, cap, counter, stream, nullptr, t.look);
    if (rc) return rc;
    const unsigned pgrid = (cap + lzs::kPieceWarps - 1) / lzs::kPieceWarps;
    const unsigned sgrid = (n_streams + lzs::kPieceWarps - 1) / lzs::kPieceWarps;
    lzs::k23p_spec<<<pgrid, lzs::kPieceThreads, 0, st>>>(in, in_off, in_len, matches, piece, t);
    lzs::k23p_fix<<<pgrid, lzs::kPieceThreads, 0, st>>>(in, in_off, in_len, matches, piece, t);
    lzs::k23p_sweep<<<sgrid, lzs::kPieceThreads, 0, st>>>(in, in_off, in_len, matches, out_cap, out_len, n_streams, piece, t);
    lzs::k23p_pack<<<pgrid, lzs::kPieceThreads, 0, st>>>(in, in_off, in_len, matches, out, out_off, out_cap, t);
    g_launches += 4;
    CUDA_TRY(cudaGetLastError());
    return LZS_B200_OK;
}

/* k4_decode over n_streams streams, or over the list `order` whose length *n_dev the device knows
 * (grid_streams bounds it) */
int launch_k4(const uint8_t *in, const uint64_t *in_off, const uint32_t *in_len, uint8_t *out, const uint64_t *out_off,
              const uint32_t *out_cap, uint32_t *out_len, uint32_t n_streams, uint32_t *counter, uint8_t *status,
              const uint32_t *order, const uint32_t *hist_len, const uint32_t *n_dev, uint32_t grid_streams, cudaStream_t st,
              DeviceInfo *d)
{
    const int lanes = decode_lanes();
    const int idx = lanes == 4 ? 0 : lanes == 8 ? 1 : lanes == 16 ? 2 : 3;
    const unsigned per_block = lzs::kDecThreads / lanes;
    unsigned       want = (grid_streams + per_block - 1) / per_block;
    unsigned       resident = static_cast<unsigned>(d->sms) * static_cast<unsigned>(d->dec_blocks[idx] > 0 ? d->dec_blocks[idx] : 1);
    const unsigned grid = want < resident ? want : resident;
    switch (lanes) {
        case 4:
            lzs::k4_decode<4><<<grid, lzs::kDecThreads, lzs::k4_smem_bytes<4>(), st>>>(
                in, in_off, in_len, out, out_off, out_cap, out_len, n_streams, counter, status, order, hist_len, n_dev);
            break;
        case 16:
            lzs::k4_decode<16><<<grid, lzs::kDecThreads, lzs::k4_smem_bytes<16>(), st>>>(
                in, in_off, in_len, out, out_off, out_cap, out_len, n_streams, counter, status, order, hist_len, n_dev);
            break;
        case 32:
            lzs::k4_decode<32><<<grid, lzs::kDecThreads, lzs::k4_smem_bytes<32>(), st>>>(
                in, in_off, in_len, out, out_off, out_cap, out_len, n_streams, counter, status, order, hist_len, n_dev);
            break;
        default:
            lzs::k4_decode<8><<<grid, lzs::kDecThreads, lzs::k4_smem_bytes<8>(), st>>>(
                in, in_off, in_len, out, out_off, out_cap, out_len, n_streams, counter, status, order, hist_len, n_dev);
            break;
    }
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return LZS_B200_OK;
}

/* The decoder for few long streams: the passes of k4_pieces.cuh, then k4_decode for the streams they
 * left (malformed, short of output): a launch that ends at once when there are none. */
int decompress_pieces(const uint8_t *in, const uint64_t *in_off, const uint32_t *in_len, uint8_t *out,
                      const uint64_t *out_off, const uint32_t *out_cap, uint32_t *out_len, uint8_t *status,
                      uint32_t n_streams, uint32_t piece, uint32_t cap, void *table_mem, cudaStream_t st, DeviceInfo *d,
                      uint32_t *jump_S = nullptr, uint32_t jump_span = 0, uint64_t jump_base = 0, uint32_t *in_used = nullptr)
{
    const lzs::DPieceTable t = lzs::dpiece_table_at(table_mem, cap, piece);
    const unsigned pgrid = (cap + 127u) / 128u, sgrid = (n_streams + 3u) / 4u;
    lzs::k4p_plan<<<1, 1024, 0, st>>>(in_len, n_streams, piece, t);
    lzs::k4p_spec<<<pgrid, 128, 0, st>>>(in, in_off, in_len, n_streams, piece, t);
    lzs::k4p_fix<<<pgrid, 128, 0, st>>>(in, in_off, in_len, n_streams, piece, 0u, t);
    for (int rep = 0; rep < fix_repairs(); rep++) lzs::k4p_fix<<<pgrid, 128, 0, st>>>(in, in_off, in_len, n_streams, piece, 1u, t);
    lzs::k4p_sweep<<<sgrid, 128, 0, st>>>(in, in_off, in_len, out_cap, out_len, status, n_streams, piece, t, in_used);
    lzs::k4p_emit<<<pgrid, 128, 0, st>>>(in, in_off, in_len, out, out_off, n_streams, piece, t);
    g_launches += 5 + fix_repairs();
    if (jump_S != nullptr) {
        /* a handful of streams: every byte finds the literal it is a copy of by pointer doubling
         * (k4_pieces.cuh) -- parallel over the bytes of a stream, where the replay is one block per stream */
        uint32_t      *flags = t.count + 8;
        const unsigned jgrid = static_cast<unsigned>(d->sms) * 8u;
        lzs::k4j_init<<<jgrid, 256, 0, st>>>(jump_S, jump_span, flags);
        lzs::k4j_fill<<<(cap + 3u) / 4u, 128, 0, st>>>(out_off, jump_base, n_streams, jump_S, t);
        for (int r = 0; r < lzs::kJumpRounds; r++) lzs::k4j_jump<<<jgrid, 256, 0, st>>>(jump_S, jump_span, flags, static_cast<uint32_t>(r));
        lzs::k4j_gather<<<jgrid, 256, 0, st>>>(out, jump_base, jump_S, jump_span);
        g_launches += 3 + lzs::kJumpRounds;
    } else {
        lzs::k4p_copy<<<n_streams, lzs::kDCopyThreads, 0, st>>>(out, out_off, out_len, n_streams, t);
        g_launches += 1;
    }
    lzs::k4p_dirty_list<<<(n_streams + 127u) / 128u, 128, 0, st>>>(n_streams, t);
    g_launches += 1;
    CUDA_TRY(cudaGetLastError());
    return launch_k4(in, in_off, in_len, out, out_off, out_cap, out_len, n_streams, t.count + 3, status, t.dirty_list, nullptr,
                     t.count + 2, n_streams, st, d);
}
}  // namespace

namespace {
int decompress_impl(const uint8_t *in, const uint64_t *in_off, const uint32_t *in_len, uint8_t *out,
                    const uint64_t *out_off, const uint32_t *out_cap, const uint32_t *hist_len, uint32_t *out_len,
                    uint8_t *status, uint32_t n_streams, void *scratch, size_t scratch_bytes, void *stream,
                    uint32_t jump_span, uint32_t *in_used = nullptr);
}

extern "C" {

int lzs_b200_match_batch_device(const uint8_t *in, const uint64_t *in_off, const uint32_t *in_len,
                                uint16_t *matches, uint32_t n_streams, uint32_t *counter, void *stream)
{
    return match_batch(in, in_off, in_len, nullptr, matches, n_streams, counter, stream);
}

/* The same for a TABLE of flows whose packets all have seg_len[f] bytes (the last one may be shorter): the
 * match finder takes every flow as ONE stream (flow f = in + flow_off[f], flow_len[f] bytes), so a flow's
 * bytes go through the tables once, and ends every position's look-ahead with its packet; the parse/pack
 * kernel then takes the packets one by one (pkt_off / pkt_len: n_packets entries, every packet inside a
 * flow).  Same bytes as lzs_b200_compress_flows_batch_device with hist_len = bytes of the flow before the
 * packet, about 1.6 times faster on 1500-byte packets. */
int lzs_b200_compress_flow_table_device(const uint8_t *in, const uint64_t *flow_off, const uint32_t *flow_len,
                                        const uint32_t *seg_len, uint32_t n_flows, const uint64_t *pkt_off,
                                        const uint32_t *pkt_len, uint32_t n_packets, uint64_t in_span, uint8_t *out,
                                        const uint64_t *out_off, const uint32_t *out_cap, uint32_t *out_len,
                                        void *scratch, size_t scratch_bytes, void *stream)
{
    if (n_flows == 0 || n_packets == 0) return LZS_B200_OK;
    if (!seg_len) return fail(LZS_B200_EINVAL, "null pointer");
    if (!scratch || scratch_bytes < kCounterBytes + matches_bytes(in_span))
        return fail(LZS_B200_EINVAL, "scratch too small: need %zu bytes, got %zu",
                    lzs_b200_compress_scratch_bytes(in_span), scratch_bytes);
    uint32_t *counter = static_cast<uint32_t *>(scratch);
    uint16_t *matches = reinterpret_cast<uint16_t *>(static_cast<uint8_t *>(scratch) + kCounterBytes);
    int rc = match_batch(in, flow_off, flow_len, nullptr, matches, n_flows, counter, stream, seg_len);
    if (rc) return rc;
    return lzs_b200_parse_pack_batch_device(in, pkt_off, pkt_len, matches, out, out_off, out_cap, out_len, n_packets,
                                            stream);
}

/* Packets of flows with kept history, the bulk path: stream s is ONE packet,
 * compressed as lzs_compress_incremental(add_end_marker = true) does on a state whose history holds
 * the hist_len[s] (<= 2047) bytes that PRECEDE in + in_off[s] in memory (the flow's earlier packets,
 * kept contiguous by the caller).  With hist_len all zero this is lzs_b200_compress_batch_device. */
int lzs_b200_compress_flows_batch_device(const uint8_t *in, const uint64_t *in_off, const uint32_t *in_len,
                                         const uint32_t *hist_len, uint64_t in_span, uint8_t *out,
                                         const uint64_t *out_off, const uint32_t *out_cap, uint32_t *out_len,
                                         uint32_t n_streams, void *scratch, size_t scratch_bytes, void *stream)
{
    if (n_streams == 0) return LZS_B200_OK;
    if (!scratch || scratch_bytes < kCounterBytes + matches_bytes(in_span))
        return fail(LZS_B200_EINVAL, "scratch too small: need %zu bytes, got %zu",
                    lzs_b200_compress_scratch_bytes(in_span), scratch_bytes);
    uint32_t *counter = static_cast<uint32_t *>(scratch);
    uint16_t *matches = reinterpret_cast<uint16_t *>(static_cast<uint8_t *>(scratch) + kCounterBytes);
    int rc = match_batch(in, in_off, in_len, hist_len, matches, n_streams, counter, stream);
    if (rc) return rc;
    return lzs_b200_parse_pack_batch_device(in, in_off, in_len, matches, out, out_off, out_cap, out_len,
                                            n_streams, stream);
}

int lzs_b200_parse_pack_batch_device(const uint8_t *in, const uint64_t *in_off, const uint32_t *in_len,
                                     const uint16_t *matches, uint8_t *out, const uint64_t *out_off,
                                     const uint32_t *out_cap, uint32_t *out_len, uint32_t n_streams,
                                     void *stream)
{
    if (n_streams == 0) return LZS_B200_OK;
    if (!in || !in_off || !in_len || !matches || !out || !out_off || !out_cap || !out_len)
        return fail(LZS_B200_EINVAL, "null pointer");
    cudaStream_t   st = static_cast<cudaStream_t>(stream);
    const unsigned grid = (n_streams + lzs::kK2Warps - 1) / lzs::kK2Warps;
    lzs::k23_parse_pack<<<grid, lzs::kK2Threads, 0, st>>>(in, in_off, in_len, matches, out, out_off, out_cap,
                                                          out_len, n_streams);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return LZS_B200_OK;
}

int lzs_b200_compress_batch_device(const uint8_t *in, const uint64_t *in_off, const uint32_t *in_len,
                                   uint64_t in_span, uint8_t *out, const uint64_t *out_off,
                                   const uint32_t *out_cap, uint32_t *out_len, uint32_t n_streams,
                                   void *scratch, size_t scratch_bytes, void *stream)
{
    if (n_streams == 0) return LZS_B200_OK;
    const size_t plain = kCounterBytes + matches_bytes(in_span);
    if (!scratch || scratch_bytes < plain)
        return fail(LZS_B200_EINVAL, "scratch too small: need %zu bytes, got %zu",
                    lzs_b200_compress_scratch_bytes(in_span), scratch_bytes);
    uint32_t *counter = static_cast<uint32_t *>(scratch);
    uint16_t *matches = reinterpret_cast<uint16_t *>(static_cast<uint8_t *>(scratch) + kCounterBytes);
    /* few long streams: cut into pieces (k23_pieces.cuh) */
    uint32_t       piece = 0;
    const uint32_t cap = cut_into_pieces(n_streams, in_span, scratch_bytes, &piece);
    if (cap)
        return compress_pieces(in, in_off, in_len, out, out_off, out_cap, out_len, n_streams, piece, cap, counter,
                               matches, static_cast<uint8_t *>(scratch) + plain, stream);
    int rc = lzs_b200_match_batch_device(in, in_off, in_len, matches, n_streams, counter, stream);
    if (rc) return rc;
    return lzs_b200_parse_pack_batch_device(in, in_off, in_len, matches, out, out_off, out_cap, out_len,
                                            n_streams, stream);
}

int lzs_b200_decompress_batch_device(const uint8_t *in, const uint64_t *in_off, const uint32_t *in_len,
                                     uint8_t *out, const uint64_t *out_off, const uint32_t *out_cap,
                                     uint32_t *out_len, uint32_t n_streams, void *scratch,
                                     size_t scratch_bytes, void *stream)
{
    return lzs_b200_decompress_status_batch_device(in, in_off, in_len, out, out_off, out_cap, out_len, nullptr,
                                                   n_streams, scratch, scratch_bytes, stream);
}

int lzs_b200_decompress_status_batch_device(const uint8_t *in, const uint64_t *in_off, const uint32_t *in_len,
                                            uint8_t *out, const uint64_t *out_off, const uint32_t *out_cap,
                                            uint32_t *out_len, uint8_t *status, uint32_t n_streams,
                                            void *scratch, size_t scratch_bytes, void *stream)
{
    return lzs_b200_decompress_flows_batch_device(in, in_off, in_len, out, out_off, out_cap, nullptr, out_len, status,
                                                  n_streams, scratch, scratch_bytes, stream);
}

/* Decoder side of the same: packet s is decoded as lzs_decompress_incremental does on a state whose
 * history holds the hist_len[s] (<= 2047) bytes that precede out + out_off[s] -- the flow's earlier
 * packets, ALREADY decoded by an earlier launch (a flow is serial: one launch per packet index,
 * all flows together).  hist_len == NULL: plain independent streams. */
int lzs_b200_decompress_flows_batch_device(const uint8_t *in, const uint64_t *in_off, const uint32_t *in_len,
                                           uint8_t *out, const uint64_t *out_off, const uint32_t *out_cap,
                                           const uint32_t *hist_len, uint32_t *out_len, uint8_t *status,
                                           uint32_t n_streams, void *scratch, size_t scratch_bytes, void *stream)
{
    return decompress_impl(in, in_off, in_len, out, out_off, out_cap, hist_len, out_len, status, n_streams, scratch,
                           scratch_bytes, stream, 0);
}

/* A handful of long streams: the copies resolved by pointer doubling (k4_pieces.cuh) -- parallel over the
 * bytes of a stream.  out_span = bytes of `out` covered by the slots (< 2 GiB); scratch of
 * lzs_b200_decompress_scratch_bytes_jump(in_span, out_span, n_streams). */
int lzs_b200_decompress_long_batch_device(const uint8_t *in, const uint64_t *in_off, const uint32_t *in_len,
                                          uint8_t *out, const uint64_t *out_off, const uint32_t *out_cap, uint64_t out_span,
                                          uint32_t *out_len, uint8_t *status, uint32_t n_streams, void *scratch,
                                          size_t scratch_bytes, void *stream)
{
    if (out_span == 0 || out_span > (1ull << 31)) return fail(LZS_B200_EINVAL, "out_span must be 1 .. 2^31");
    return decompress_impl(in, in_off, in_len, out, out_off, out_cap, nullptr, out_len, status, n_streams, scratch,
                           scratch_bytes, stream, static_cast<uint32_t>(out_span));
}

}  // extern "C"

namespace {
/* jump_span != 0 (host path, a handful of long streams): the last 4 * jump_span bytes of scratch hold one
 * pointer per byte of out[0 .. jump_span), and the copies are resolved by pointer doubling instead of
 * being replayed (k4_pieces.cuh) */
int decompress_impl(const uint8_t *in, const uint64_t *in_off, const uint32_t *in_len, uint8_t *out,
                    const uint64_t *out_off, const uint32_t *out_cap, const uint32_t *hist_len, uint32_t *out_len,
                    uint8_t *status, uint32_t n_streams, void *scratch, size_t scratch_bytes, void *stream,
                    uint32_t jump_span, uint32_t *in_used)
{
    if (n_streams == 0) return LZS_B200_OK;
    if (!in || !in_off || !in_len || !out || !out_off || !out_cap || !out_len)
        return fail(LZS_B200_EINVAL, "null pointer");
    if (!scratch || scratch_bytes < kCounterBytes) return fail(LZS_B200_EINVAL, "scratch too small");
    DeviceInfo *d = nullptr;
    int         rc = device_info(&d);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    uint32_t    *counter = static_cast<uint32_t *>(scratch);
    CUDA_TRY(cudaMemsetAsync(counter, 0, kCounterBytes, st));
    /* few long streams, and scratch with room for a piece table (lzs_b200_decompress_scratch_bytes_long):
     * token starts found in parallel inside the streams, k4_pieces.cuh */
    const uint32_t dpiece = dpiece_bytes();
    const size_t   fixed = kCounterBytes + order_bytes(n_streams);
    uint32_t      *jump_S = nullptr;
    if (jump_span != 0) {
        const size_t sbytes = align_up(static_cast<size_t>(jump_span) * sizeof(uint32_t), 256);
        if (scratch_bytes > fixed + 4096 + sbytes) {
            scratch_bytes -= sbytes;
            jump_S = reinterpret_cast<uint32_t *>(static_cast<uint8_t *>(scratch) + scratch_bytes);
        }
    }
    if (dpiece != 0 && hist_len == nullptr && n_streams <= kCutStreamsMaxDecode && scratch_bytes > fixed + 4096) {
        /* as many table entries as the scratch holds (dpiece_table_bytes is linear in them) */
        const size_t   per_piece = lzs::dpiece_table_bytes(1, dpiece) - lzs::dpiece_table_bytes(0, dpiece);
        const uint64_t fit = (scratch_bytes - fixed - lzs::dpiece_table_bytes(0, dpiece)) / per_piece;
        const uint32_t cap = fit > 0x00FFFFFFull ? 0x00FFFFFFu : static_cast<uint32_t>(fit);
        if (cap >= 2u * n_streams + 16u && fixed + lzs::dpiece_table_bytes(cap, dpiece) <= scratch_bytes)
            return decompress_pieces(in, in_off, in_len, out, out_off, out_cap, out_len, status, n_streams, dpiece, cap,
                                     static_cast<uint8_t *>(scratch) + fixed, st, d, jump_S, jump_span, 0, in_used);
    }
    /* launch order: streams of similar density together, the fastest kind last (k4_decode.cuh);
     * needs scratch for one index per stream, otherwise the streams go in index order */
    const uint32_t *order = nullptr;
    if (n_streams >= 1024u && scratch_bytes >= lzs_b200_decompress_scratch_bytes_for(n_streams) && order_streams()) {
        uint32_t *ord = reinterpret_cast<uint32_t *>(static_cast<uint8_t *>(scratch) + kCounterBytes);
        uint32_t *counts = counter + 8;                  /* 32 words inside the counter block */
        const unsigned blocks = (n_streams + 255u) / 256u;
        lzs::k4_order_count<<<blocks, 256, 0, st>>>(in_len, out_cap, n_streams, counts);
        lzs::k4_order_scatter<<<blocks, 256, 0, st>>>(in_len, out_cap, n_streams, counts, ord);
        g_launches += 2;
        order = ord;
    }
    rc = launch_k4(in, in_off, in_len, out, out_off, out_cap, out_len, n_streams, counter, status, order, hist_len, nullptr,
                   n_streams, st, d);
    return rc;
}
}  // namespace

extern "C" {

int lzs_b200_corpus_fill_device(uint8_t *dst, uint64_t stride, uint32_t stream_len, uint64_t first_index,
                                uint64_t n, uint64_t seed, int kind, void *stream)
{
    if (n == 0) return LZS_B200_OK;
    if (!dst) return fail(LZS_B200_EINVAL, "null pointer");
    const unsigned threads = 64;
    const unsigned grid = static_cast<unsigned>((n + threads - 1) / threads);
    corpus_fill_kernel<<<grid, threads, 0, static_cast<cudaStream_t>(stream)>>>(dst, stride, stream_len,
                                                                              first_index, n, seed, kind);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return LZS_B200_OK;
}

int lzs_b200_pack_streams_device(const uint8_t *src, const uint64_t *src_off, const uint32_t *len, uint8_t *dst,
                                 const uint64_t *dst_off, uint32_t n_streams, void *stream)
{
    if (n_streams == 0) return LZS_B200_OK;
    if (!src || !src_off || !len || !dst || !dst_off) return fail(LZS_B200_EINVAL, "null pointer");
    gather_streams_kernel<<<n_streams, 128, 0, static_cast<cudaStream_t>(stream)>>>(src, src_off, len, dst, dst_off,
                                                                                    n_streams);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return LZS_B200_OK;
}

int lzs_b200_pack_streams_peers_device(const uint8_t *src, const uint64_t *src_off, const uint32_t *len,
                                       uint8_t *const *dst_ptrs, uint32_t n_dst, uint64_t base, const uint64_t *dst_off,
                                       uint32_t n_streams, void *stream)
{
    if (n_streams == 0 || n_dst == 0) return LZS_B200_OK;
    if (!src || !src_off || !len || !dst_ptrs || !dst_off) return fail(LZS_B200_EINVAL, "null pointer");
    scatter_streams_peers_kernel<<<n_streams, 128, 0, static_cast<cudaStream_t>(stream)>>>(src, src_off, len, dst_ptrs, n_dst,
                                                                                          base, dst_off, n_streams);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return LZS_B200_OK;
}

int lzs_b200_pack_streams_multicast_device(const uint8_t *src, const uint64_t *src_off, const uint32_t *len, uint8_t *mc_ptr,
                                           uint64_t base, const uint64_t *dst_off, uint32_t n_streams, void *stream)
{
    if (n_streams == 0) return LZS_B200_OK;
    if (!src || !src_off || !len || !mc_ptr || !dst_off) return fail(LZS_B200_EINVAL, "null pointer");
    scatter_streams_multicast_kernel<<<n_streams, 128, 0, static_cast<cudaStream_t>(stream)>>>(src, src_off, len, mc_ptr, base,
                                                                                              dst_off, n_streams);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return LZS_B200_OK;
}

uint32_t lzs_b200_chunk_count(uint64_t total, uint32_t chunk)
{
    if (chunk == 0) return 0;
    return static_cast<uint32_t>((total + chunk - 1) / chunk);
}

void lzs_b200_chunk_layout(uint64_t total, uint32_t chunk, uint64_t out_stride, uint64_t *in_off,
                           uint32_t *in_len, uint64_t *out_off, uint32_t *out_cap)
{
    const uint32_t n = lzs_b200_chunk_count(total, chunk);
    for (uint32_t s = 0; s < n; s++) {
        const uint64_t o = static_cast<uint64_t>(s) * chunk;
        const uint64_t l = total - o < chunk ? total - o : chunk;
        if (in_off) in_off[s] = o;
        if (in_len) in_len[s] = static_cast<uint32_t>(l);
        if (out_off) out_off[s] = static_cast<uint64_t>(s) * out_stride;
        if (out_cap) out_cap[s] = static_cast<uint32_t>(out_stride);
    }
}

}  // extern "C"

/* ------------------------------------------------------------------ host batches */

namespace {

/* Grow-only device arena for the host-pointer entry points (one per device). */
/* Result lengths of a slice go to the host by plain stores into mapped pinned memory.  A
 * cudaMemcpyAsync for them would sit in the download engine's queue in ISSUE order, ahead of the
 * bulk downloads of earlier slices that the host only issues once it has seen their lengths, and
 * hold those back until the later slice's kernels are done (measured on B200, 1 GiB: first
 * download finished at 51 ms instead of 8 ms). */
__global__ void publish_lengths_kernel(uint32_t *__restrict__ host_mapped, const uint32_t *__restrict__ dev, uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) host_mapped[i] = dev[i];
}

struct HostPath {
    std::mutex     mu;
    std::once_flag once;                    /* streams are created once per device        */
    int            init_rc = LZS_B200_OK;
    cudaStream_t stream = nullptr;          /* copies of small arrays, simple path        */
    cudaStream_t work[8] = {};              /* slice k: upload + kernels on work[k % 8]   */
    cudaStream_t down = nullptr;            /* downloads of finished slices               */
    cudaStream_t up = nullptr;              /* uploads, in slice order, ahead of the kernels */
    void        *buf[11] = {};
    size_t       cap[11] = {};
    uint32_t    *pinned_len = nullptr;      /* pinned, device-mapped: per-stream result lengths */
    uint32_t    *pinned_len_dev = nullptr;  /* the same memory as the device addresses it       */
    size_t       pinned_cap = 0;
    uint8_t     *bounce = nullptr;          /* pinned staging for outputs whose slots have gaps */
    size_t       bounce_cap = 0;

    int reserve_bounce(size_t bytes)
    {
        if (bounce_cap >= bytes) return LZS_B200_OK;
        if (bounce) cudaFreeHost(bounce);
        bounce = nullptr;
        bounce_cap = 0;
        if (cudaHostAlloc(reinterpret_cast<void **>(&bounce), bytes, cudaHostAllocDefault) != cudaSuccess) {
            cudaGetLastError();
            return fail(LZS_B200_ENOMEM, "cudaHostAlloc(%zu) failed", bytes);
        }
        bounce_cap = bytes;
        return LZS_B200_OK;
    }

    /* give the device and pinned memory back (the streams stay) */
    void release()
    {
        for (int i = 0; i < 11; i++) {
            if (buf[i]) cudaFree(buf[i]);
            buf[i] = nullptr;
            cap[i] = 0;
        }
        if (pinned_len) cudaFreeHost(pinned_len);
        pinned_len = pinned_len_dev = nullptr;
        pinned_cap = 0;
        if (bounce) cudaFreeHost(bounce);
        bounce = nullptr;
        bounce_cap = 0;
    }

    int reserve_pinned(size_t count)
    {
        if (pinned_cap >= count) return LZS_B200_OK;
        if (pinned_len) cudaFreeHost(pinned_len);
        pinned_len = nullptr;
        pinned_len_dev = nullptr;
        pinned_cap = 0;
        if (cudaHostAlloc(reinterpret_cast<void **>(&pinned_len), count * sizeof(uint32_t), cudaHostAllocMapped) !=
                cudaSuccess ||
            cudaHostGetDevicePointer(reinterpret_cast<void **>(&pinned_len_dev), pinned_len, 0) != cudaSuccess) {
            cudaGetLastError();
            return fail(LZS_B200_ENOMEM, "cudaHostAlloc(%zu, mapped) failed", count * sizeof(uint32_t));
        }
        pinned_cap = count;
        return LZS_B200_OK;
    }

    int reserve(int slot, size_t bytes)
    {
        bytes = align_up(bytes ? bytes : 1, 256);
        if (cap[slot] >= bytes) return LZS_B200_OK;
        if (buf[slot]) cudaFree(buf[slot]);
        buf[slot] = nullptr;
        cap[slot] = 0;
        cudaError_t e = cudaMalloc(&buf[slot], bytes);
        if (e != cudaSuccess) cudaGetLastError();        /* not sticky; a caller may go on with a smaller request */
        if (e != cudaSuccess)
            return fail(e == cudaErrorNoDevice ? LZS_B200_ENODEVICE : LZS_B200_ENOMEM,
                        "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
        cap[slot] = bytes;
        return LZS_B200_OK;
    }
};

HostPath *host_path_slot(int *rc_out)
{
    static HostPath paths[64];
    int             dev = 0;
    cudaError_t     e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        *rc_out = fail(e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? LZS_B200_ENODEVICE : LZS_B200_ECUDA,
                       "cudaGetDevice failed: %s", cudaGetErrorString(e));
        return nullptr;
    }
    if (dev < 0 || dev >= 64) {
        *rc_out = fail(LZS_B200_EINVAL, "device ordinal %d out of range", dev);
        return nullptr;
    }
    *rc_out = LZS_B200_OK;
    return &paths[dev];
}

int host_path_init(HostPath &p)
{
    /* Priorities fall from work[0] to work[7]: the decoder's slices run concurrently (it
     * takes several to fill the GPU), and the earlier slice should win the SMs so that its
     * download can start while the later ones are still being decoded.  `down` also runs
     * the small gather kernel of the packed compressor: highest priority as well. */
    int prio_lo = 0, prio_hi = 0;           /* numerically lower = more urgent */
    CUDA_TRY(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    for (int i = 0; i < 8; i++) {
        const int pr = prio_hi + i < prio_lo ? prio_hi + i : prio_lo;
        CUDA_TRY(cudaStreamCreateWithPriority(&p.work[i], cudaStreamNonBlocking, pr));
    }
    CUDA_TRY(cudaStreamCreateWithPriority(&p.down, cudaStreamNonBlocking, prio_hi));
    CUDA_TRY(cudaStreamCreateWithFlags(&p.up, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&p.stream, cudaStreamNonBlocking));
    return LZS_B200_OK;
}

int host_path(HostPath **out)
{
    int       rc = LZS_B200_OK;
    HostPath *p = host_path_slot(&rc);
    if (!p) return rc;
    std::call_once(p->once, [p] { p->init_rc = host_path_init(*p); });
    if (p->init_rc) return fail(p->init_rc, "creating the CUDA streams of the host path failed");
    *out = p;
    return LZS_B200_OK;
}

/* Everything the sliced pipelines create per call; the destructor runs on EVERY exit, so an
 * error in the middle of a pipeline leaves no copy in flight on the caller's buffers, no kernel
 * on the arena the next call will reuse, and no event behind. */
struct PipelineScope {
    HostPath                &p;
    std::vector<cudaEvent_t> events;
    explicit PipelineScope(HostPath &hp) : p(hp) {}
    int make(cudaEvent_t *e, unsigned flags)
    {
        CUDA_TRY(cudaEventCreateWithFlags(e, flags));
        events.push_back(*e);
        return LZS_B200_OK;
    }
    cudaError_t drain()
    {
        cudaError_t first = cudaSuccess;
        auto note = [&](cudaError_t e) { if (e != cudaSuccess && first == cudaSuccess) first = e; };
        for (auto &w : p.work) note(cudaStreamSynchronize(w));
        note(cudaStreamSynchronize(p.down));
        note(cudaStreamSynchronize(p.up));
        note(cudaStreamSynchronize(p.stream));
        return first;
    }
    ~PipelineScope()
    {
        drain();
        for (auto &e : events) cudaEventDestroy(e);
        cudaGetLastError();
    }
};

/* If [p, p + bytes) lies in ONE pinned, device-mapped host allocation (cudaHostAlloc / cudaHostRegister
 * under unified addressing), the address the device uses for it; else nullptr.  Off unless
 * LZS_B200_ZEROCOPY=1 / lzs_b200_set_zero_copy_output(1): measured on B200 (1 GiB, PCIe gen 5) the decode
 * call takes 31 ms writing straight into pinned memory against 29 ms with a staging buffer and
 * copy-engine downloads -- the SMs' posted writes do not reach the copy engine's rate. */
uint8_t *mapped_host_range(const uint8_t *p, uint64_t bytes)
{
    int on = g_zero_copy_out.load();
    if (on < 0) {
        const char *e = getenv("LZS_B200_ZEROCOPY");
        on = e ? atoi(e) : 0;
        g_zero_copy_out.store(on);
    }
    if (!on || !p || bytes == 0) return nullptr;
    cudaPointerAttributes a0, a1;
    if (cudaPointerGetAttributes(&a0, p) != cudaSuccess || cudaPointerGetAttributes(&a1, p + bytes - 1) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    if (a0.type != cudaMemoryTypeHost || a1.type != cudaMemoryTypeHost || !a0.devicePointer || !a1.devicePointer) return nullptr;
    if (static_cast<uint8_t *>(a1.devicePointer) - static_cast<uint8_t *>(a0.devicePointer) != static_cast<ptrdiff_t>(bytes - 1))
        return nullptr;                                  /* two allocations that happen to be neighbours */
    return static_cast<uint8_t *>(a0.devicePointer);
}

/* The caller's description of a batch must stay inside the spans it states: the device arena
 * is sized from the spans, and a slot that reaches beyond them would be written outside it. */
int check_layout(const uint64_t *in_off, const uint32_t *in_len, uint64_t in_span, const uint64_t *out_off,
                 const uint32_t *out_cap, uint64_t out_span, uint32_t n)
{
    for (uint32_t s = 0; s < n; s++) {
        if (in_off[s] > in_span || in_len[s] > in_span - in_off[s])
            return fail(LZS_B200_EINVAL, "stream %u: input [%llu, +%u) reaches beyond in_span %llu", s,
                        static_cast<unsigned long long>(in_off[s]), in_len[s], static_cast<unsigned long long>(in_span));
        if (out_off && (out_off[s] > out_span || out_cap[s] > out_span - out_off[s]))
            return fail(LZS_B200_EINVAL, "stream %u: output slot [%llu, +%u) reaches beyond out_span %llu", s,
                        static_cast<unsigned long long>(out_off[s]), out_cap[s], static_cast<unsigned long long>(out_span));
    }
    return LZS_B200_OK;
}

/* Bring the bytes streams [a, b) produced back to the caller: only [out_off, out_off + out_len)
 * of a stream, never a byte outside the slots (the caller may keep framing between them).
 * Slots that touch are fetched in one copy up to the last produced byte (what lies between two
 * produced streams is then slot space, which the contract leaves unspecified); a slice whose
 * slots have gaps everywhere goes through a pinned staging buffer and host copies instead of
 * thousands of small device copies. */
int download_streams(HostPath &p, uint8_t *out, const uint8_t *d_out, const uint64_t *out_off, const uint32_t *out_cap,
                     const uint32_t *out_len, uint32_t a, uint32_t b, cudaStream_t st, bool *used_bounce)
{
    struct Run { uint64_t lo, hi; };
    std::vector<Run> runs;
    uint64_t         slot_end = 0;
    for (uint32_t s = a; s < b; s++) {
        if (out_len[s] == 0) continue;
        const uint64_t lo = out_off[s], hi = out_off[s] + out_len[s];
        if (!runs.empty() && lo == slot_end) runs.back().hi = hi;     /* touches the previous slot */
        else runs.push_back({lo, hi});
        slot_end = out_off[s] + out_cap[s];
    }
    if (runs.size() <= 64) {
        for (const Run &r : runs)
            CUDA_TRY(cudaMemcpyAsync(out + r.lo, d_out + r.lo, r.hi - r.lo, cudaMemcpyDeviceToHost, st));
        return LZS_B200_OK;
    }
    const uint64_t lo = runs.front().lo, hi = runs.back().hi;
    int            rc = p.reserve_bounce(hi - lo);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(p.bounce, d_out + lo, hi - lo, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    for (const Run &r : runs) memcpy(out + r.lo, p.bounce + (r.lo - lo), r.hi - r.lo);
    if (used_bounce) *used_bounce = true;
    return LZS_B200_OK;
}

enum { S_IN, S_OUT, S_INOFF, S_INLEN, S_OUTOFF, S_OUTCAP, S_OUTLEN, S_SCRATCH, S_COUNTERS, S_PACKED, S_PACKOFF };

/* Uncompressed bytes per pipeline slice.  Compressor: a quarter of the batch, between 32 and
 * 256 MiB (its kernels run slice after slice, and every slice pays the tail of the match finder
 * and the latency of the parse kernel's slowest stream; measured on B200, 1 GiB: 128 / 192 / 256 /
 * 384 MiB slices -> 66.0 / 62.1 / 60.9 / 64.2 ms).  Decompressor: 64 MiB -- its slices run
 * concurrently on eight streams, so small slices cost nothing on the GPU and let the first
 * download start early (1 GiB: 16 / 32 / 48 / 64 / 96 / 128 / 256 MiB -> 59.5 / 35.5 / 29.2 / 28.8 /
 * 29.4 / 30.5 / 32.0 ms; below 64 MiB slices start to queue behind each other on the eight
 * streams).  LZS_B200_SLICE_MIB overrides both, for tuning. */
uint64_t slice_bytes(uint64_t total, bool decompress)
{
    static long env_mib = -1;
    if (env_mib < 0) {
        const char *e = getenv("LZS_B200_SLICE_MIB");
        env_mib = e ? atol(e) : 0;
    }
    if (env_mib > 0) return static_cast<uint64_t>(env_mib) << 20;
    if (decompress) return 64ull << 20;
    uint64_t v = total / 4;
    if (v < (32ull << 20)) v = 32ull << 20;
    if (v > (256ull << 20)) v = 256ull << 20;
    return v;
}

/* The compressor's parse/pack kernel (one warp per stream: a slice of it lasts as long as its
 * slowest stream, ~3 ms for 64 KiB streams, however few there are) runs on a second, less urgent
 * stream behind the match finder's, so that it shares the SMs with the next slice's match finder
 * instead of holding it up.  LZS_B200_K23_STREAM=0 puts both back on one stream. */
int parse_side_streams()
{
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("LZS_B200_K23_STREAM");
        v = e ? atoi(e) : 1;
        if (v < 0) v = 0;
        if (v > 7) v = 7;
    }
    return v;
}

/* First stream of every pipeline slice (slices are sized by UNcompressed bytes, `weight[s]`).
 * With `short_ends` (the compressor, which is kernel bound) the first and the last slice are a
 * quarter of the others: the first kernel starts after a short upload and the last download,
 * which has to wait for the last kernel, is short (measured on B200, 1 GiB: 67 -> 64.5 ms).
 * The decompressor is bound by the download and measured slower that way (31.8 -> 35.7 ms), so it
 * keeps equal slices.  LZS_B200_SLICE_RAMP=0 turns the short ends off. */
std::vector<uint32_t> plan_slices(const uint32_t *weight, uint32_t n, uint64_t total, bool short_ends)
{
    static int ramp = -1;
    if (ramp < 0) {
        const char *e = getenv("LZS_B200_SLICE_RAMP");
        ramp = e ? atoi(e) : 1;
    }
    const uint64_t big = slice_bytes(total, !short_ends);   /* short ends: the compressor */
    const uint64_t small = big / 4;
    std::vector<uint64_t> plan;
    uint64_t              left = total;
    if (ramp && short_ends && total >= 2 * big) {
        plan.push_back(small);
        left -= small;
        while (left > small) {
            const uint64_t v = left - small < big + big / 2 ? left - small : big;
            plan.push_back(v);
            left -= v;
        }
    }
    std::vector<uint32_t> first;
    size_t                k = 0;
    uint64_t              acc = 0, want = 0;
    for (uint32_t s2 = 0; s2 < n; s2++) {
        if (acc >= want) {
            first.push_back(s2);
            acc = 0;
            want = k < plan.size() ? plan[k++] : big;
        }
        acc += weight[s2];
    }
    return first;
}

int run_host_batch(bool decompress, const uint8_t *in, const uint64_t *in_off, const uint32_t *in_len,
                   uint64_t in_span, uint8_t *out, const uint64_t *out_off, const uint32_t *out_cap,
                   uint32_t *out_len, uint64_t out_span, uint32_t n, uint32_t *in_used = nullptr)
{
    if (n == 0) return LZS_B200_OK;
    if (!in_off || !in_len || !out_off || !out_cap || !out_len || (!in && in_span) || (!out && out_span))
        return fail(LZS_B200_EINVAL, "null pointer");
    int rc = check_layout(in_off, in_len, in_span, out_off, out_cap, out_span, n);
    if (rc) return rc;
    if (lzs_b200_device_count() <= 0) return fail(LZS_B200_ENODEVICE, "no CUDA device: the LZS codec has no CPU path");
    HostPath *hp = nullptr;
    if ((rc = host_path(&hp))) return rc;
    std::lock_guard<std::mutex> lock(hp->mu);
    HostPath     &p = *hp;
    PipelineScope scope(p);                 /* declared after the lock: drains before the arena is released */
    cudaStream_t  st = p.stream;
    /* few long streams to decode: one call with room for the piece table (k4_pieces.cuh) instead of the
     * slices, whose streams would each be decoded by one group of lanes */
    const bool    long_decode = decompress && dpiece_bytes() != 0 && n <= kCutStreamsMaxDecode &&
                             static_cast<uint64_t>(n) * 2u * dpiece_bytes() <= in_span;
    /* a handful of them: pointer doubling instead of the replay (4 bytes of scratch per byte of output) */
    /* (one stream cannot make more than 30 bytes per byte: 15 per continuation nibble) */
    const uint64_t jump_need = n == 1 && out_span > 30ull * in_span + 64u ? 30ull * in_span + 64u : out_span;
    uint32_t jump_span = long_decode && n <= jump_streams_max() && jump_need >= jump_bytes_min() && jump_need <= (1ull << 31)
                             ? static_cast<uint32_t>(jump_need) : 0u;
    auto scratch_for = [&](uint32_t span) {
        return decompress ? (long_decode ? lzs_b200_decompress_scratch_bytes_long(in_span, n) +
                                               (span ? align_up(static_cast<size_t>(span) * 4u, 256) + 4096 : 0)
                                         : lzs_b200_decompress_scratch_bytes_for(n))
                          : lzs_b200_compress_scratch_bytes(in_span);
    };
    size_t scratch = scratch_for(jump_span);
    /* A decoder writes every output byte exactly once and reads none back (its history is in shared
     * memory), so when the caller's output buffer is pinned and device mapped it CAN decode straight
     * into it over PCIe (option, off by default: see mapped_host_range). */
    /* (not for the piece decoder: its copy pass reads the output back) */
    uint8_t *const direct_out = decompress && !long_decode ? mapped_host_range(out, out_span) : nullptr;
    if ((rc = p.reserve(S_IN, in_span + 64))) return rc;
    if (!direct_out && (rc = p.reserve(S_OUT, out_span + 64))) return rc;
    if ((rc = p.reserve(S_INOFF, n * sizeof(uint64_t)))) return rc;
    if ((rc = p.reserve(S_INLEN, n * sizeof(uint32_t)))) return rc;
    if ((rc = p.reserve(S_OUTOFF, n * sizeof(uint64_t)))) return rc;
    if ((rc = p.reserve(S_OUTCAP, n * sizeof(uint32_t)))) return rc;
    if ((rc = p.reserve(S_OUTLEN, n * sizeof(uint32_t)))) return rc;
    if ((rc = p.reserve(S_SCRATCH, scratch))) {
        if (jump_span == 0) return rc;
        jump_span = 0;                      /* no room for a pointer per output byte: the replay instead */
        scratch = scratch_for(0);
        if ((rc = p.reserve(S_SCRATCH, scratch))) return rc;
    }

    uint8_t  *d_in = static_cast<uint8_t *>(p.buf[S_IN]);
    uint8_t  *d_out = direct_out ? direct_out : static_cast<uint8_t *>(p.buf[S_OUT]);
    uint64_t *d_inoff = static_cast<uint64_t *>(p.buf[S_INOFF]);
    uint32_t *d_inlen = static_cast<uint32_t *>(p.buf[S_INLEN]);
    uint64_t *d_outoff = static_cast<uint64_t *>(p.buf[S_OUTOFF]);
    uint32_t *d_outcap = static_cast<uint32_t *>(p.buf[S_OUTCAP]);
    uint32_t *d_outlen = static_cast<uint32_t *>(p.buf[S_OUTLEN]);
    CUDA_TRY(cudaMemcpyAsync(d_inoff, in_off, n * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_inlen, in_len, n * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_outoff, out_off, n * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_outcap, out_cap, n * sizeof(uint32_t), cudaMemcpyHostToDevice, st));

    /* Streams laid out in increasing order (the usual chunked file / packet table) are
     * processed in slices so that upload, kernels and download of different slices
     * overlap: slice k uploads and computes on work[k % 8] (the decoder needs several
     * slices resident at once to fill the GPU), finished slices are downloaded on `down`, and
     * only the bytes a slice really produced come back. */
    bool ordered = n > 8 && !long_decode && in_used == nullptr;
    for (uint32_t s2 = 1; s2 < n && ordered; s2++)
        ordered = in_off[s2] >= in_off[s2 - 1] + in_len[s2 - 1] && out_off[s2] >= out_off[s2 - 1] + out_cap[s2 - 1];
    if (ordered) {
        std::vector<uint32_t> first =
            plan_slices(decompress ? out_cap : in_len, n, decompress ? out_span : in_span, !decompress);
        const uint32_t nslice = static_cast<uint32_t>(first.size());
        first.push_back(n);
        if ((rc = p.reserve(S_COUNTERS, static_cast<size_t>(nslice) * 512))) return rc;
        if ((rc = p.reserve_pinned(n))) return rc;          /* D2H into pageable memory would block the host */
        uint8_t *d_cnt = static_cast<uint8_t *>(p.buf[S_COUNTERS]);
        std::vector<cudaEvent_t> ev(nslice + 1), up_ev(nslice);
        for (auto &e : ev) if ((rc = scope.make(&e, cudaEventDisableTiming))) return rc;
        for (auto &e : up_ev) if ((rc = scope.make(&e, cudaEventDisableTiming))) return rc;
        CUDA_TRY(cudaEventRecord(ev[nslice], st));
        for (auto &w : p.work) CUDA_TRY(cudaStreamWaitEvent(w, ev[nslice], 0));
        CUDA_TRY(cudaStreamWaitEvent(p.up, ev[nslice], 0));
        uint16_t *matches = reinterpret_cast<uint16_t *>(static_cast<uint8_t *>(p.buf[S_SCRATCH]) + kCounterBytes);
        uint32_t       piece = 0;
        const uint32_t piece_cap = decompress ? 0u : cut_into_pieces(n, in_span, p.cap[S_SCRATCH], &piece);
        const bool     pieces = piece_cap != 0;
        /* LZS_B200_TRACE=1 prints when every stage of every slice finished (ms from the start) */
        const bool               trace = getenv("LZS_B200_TRACE") != nullptr;
        std::vector<cudaEvent_t> tr(trace ? nslice * 4 + 1 : 0);
        for (auto &e : tr) if ((rc = scope.make(&e, cudaEventDefault))) return rc;
        if (trace) cudaEventRecord(tr[nslice * 4], st);
        for (uint32_t k = 0; k < nslice; k++) {
            const uint32_t a = first[k], b = first[k + 1], cnt = b - a;
            /* The decoder runs slice k on work[k % 8]: it needs several slices resident at once to
             * fill the GPU.  The compressor's match finder fills the GPU with any slice, so its
             * kernels run in slice order on ONE stream (uploads ahead of them on `up`): slices
             * then finish one after the other and their downloads overlap the later slices'
             * kernels, instead of all slices finishing together at the end. */
            cudaStream_t   ws = decompress ? p.work[k % 8u] : p.work[0];
            const uint64_t lo = in_off[a], hi = in_off[b - 1] + in_len[b - 1];
            /* uploads in slice order on their own stream (issued on the work streams, the copy
             * engine takes them in any order and an early slice can arrive after a later one) */
            if (hi > lo) CUDA_TRY(cudaMemcpyAsync(d_in + lo, in + lo, hi - lo, cudaMemcpyHostToDevice, p.up));
            if (trace) cudaEventRecord(tr[k * 4 + 0], p.up);
            CUDA_TRY(cudaEventRecord(up_ev[k], p.up));
            CUDA_TRY(cudaStreamWaitEvent(ws, up_ev[k], 0));
            if (decompress) {
                rc = lzs_b200_decompress_batch_device(d_in, d_inoff + a, d_inlen + a, d_out, d_outoff + a, d_outcap + a,
                                                      d_outlen + a, cnt, d_cnt + k * 512, 256, ws);
            } else if (pieces) {
                /* few long streams: cut into pieces; the slices share one piece table, which is safe
                 * because the compressor's slices follow each other on one stream */
                rc = compress_pieces(d_in, d_inoff + a, d_inlen + a, d_out, d_outoff + a, d_outcap + a, d_outlen + a, cnt,
                                     piece, piece_cap, reinterpret_cast<uint32_t *>(d_cnt + k * 512), matches,
                                     static_cast<uint8_t *>(p.buf[S_SCRATCH]) + kCounterBytes + matches_bytes(in_span), ws);
            } else {
                rc = lzs_b200_match_batch_device(d_in, d_inoff + a, d_inlen + a, matches, cnt,
                                                 reinterpret_cast<uint32_t *>(d_cnt + k * 512), ws);
                if (!rc && parse_side_streams() > 0) {
                    CUDA_TRY(cudaEventRecord(up_ev[k], ws));            /* reused: the upload has been waited for */
                    ws = p.work[1 + k % static_cast<uint32_t>(parse_side_streams())];
                    CUDA_TRY(cudaStreamWaitEvent(ws, up_ev[k], 0));
                }
                if (!rc)
                    rc = lzs_b200_parse_pack_batch_device(d_in, d_inoff + a, d_inlen + a, matches, d_out, d_outoff + a,
                                                          d_outcap + a, d_outlen + a, cnt, ws);
            }
            if (rc) return rc;
            if (trace) cudaEventRecord(tr[k * 4 + 1], ws);
            publish_lengths_kernel<<<(cnt + 255u) / 256u, 256, 0, ws>>>(p.pinned_len_dev + a, d_outlen + a, cnt);
            CUDA_TRY(cudaGetLastError());
            CUDA_TRY(cudaEventRecord(ev[k], ws));
        }
        for (uint32_t k = 0; k < nslice; k++) {
            const uint32_t a = first[k], b = first[k + 1];
            CUDA_TRY(cudaEventSynchronize(ev[k]));
            memcpy(out_len + a, p.pinned_len + a, (b - a) * sizeof(uint32_t));
            if (trace) cudaEventRecord(tr[k * 4 + 2], p.down);
            if (!direct_out && (rc = download_streams(p, out, d_out, out_off, out_cap, out_len, a, b, p.down, nullptr))) return rc;
            if (trace) cudaEventRecord(tr[k * 4 + 3], p.down);
        }
        CUDA_TRY(scope.drain());
        if (trace) {
            for (uint32_t k = 0; k < nslice; k++) {
                float t[4] = {0, 0, 0, 0};
                for (int j = 0; j < 4; j++) cudaEventElapsedTime(&t[j], tr[nslice * 4], tr[k * 4 + j]);
                fprintf(stderr, "lzs (b200) trace: %s slice %u, %u streams: upload done %.2f  kernels done %.2f | download %.2f-%.2f\n",
                        decompress ? "decompress" : "compress", k, first[k + 1] - first[k], t[0], t[1], t[2], t[3]);
            }
        }
        return LZS_B200_OK;
    }

    if (in_span) CUDA_TRY(cudaMemcpyAsync(d_in, in, in_span, cudaMemcpyHostToDevice, st));
    uint32_t *d_used = nullptr;
    if (decompress && in_used != nullptr) {             /* bytes every stream took up to its end marker (unknown: all ones) */
        if ((rc = p.reserve(S_COUNTERS, static_cast<size_t>(n) * sizeof(uint32_t)))) return rc;
        d_used = static_cast<uint32_t *>(p.buf[S_COUNTERS]);
        CUDA_TRY(cudaMemsetAsync(d_used, 0xFF, static_cast<size_t>(n) * sizeof(uint32_t), st));
    }
    if (decompress)
        rc = decompress_impl(d_in, d_inoff, d_inlen, d_out, d_outoff, d_outcap, nullptr, d_outlen, nullptr, n,
                             p.buf[S_SCRATCH], scratch, st, jump_span, d_used);
    else
        rc = lzs_b200_compress_batch_device(d_in, d_inoff, d_inlen, in_span, d_out, d_outoff, d_outcap, d_outlen, n,
                                            p.buf[S_SCRATCH], p.cap[S_SCRATCH], st);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(out_len, d_outlen, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    if (d_used) CUDA_TRY(cudaMemcpyAsync(in_used, d_used, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    /* only what was produced comes back, and nothing outside the slots is touched */
    if (!direct_out && (rc = download_streams(p, out, d_out, out_off, out_cap, out_len, 0, n, st, nullptr))) return rc;
    CUDA_TRY(cudaStreamSynchronize(st));
    return LZS_B200_OK;
}

/* Compress n ordered streams and return them PACKED: out_off[s] (written here, multiples of
 * 16) and out_len[s]; the bytes between streams are padding.  Same sliced pipeline as
 * run_host_batch, plus a gather kernel per slice, so only compressed bytes cross PCIe. */
int run_compress_packed(const uint8_t *in, const uint64_t *in_off, const uint32_t *in_len, uint64_t in_span,
                        uint8_t *out, uint64_t out_capacity, uint64_t *out_off, uint32_t *out_len, uint32_t n,
                        uint64_t *out_used)
{
    if (out_used) *out_used = 0;
    if (n == 0) return LZS_B200_OK;
    if (!in_off || !in_len || !out_off || !out_len || (!in && in_span) || !out)
        return fail(LZS_B200_EINVAL, "null pointer");
    for (uint32_t s2 = 1; s2 < n; s2++)
        if (in_off[s2] < in_off[s2 - 1] + in_len[s2 - 1])
            return fail(LZS_B200_EINVAL, "packed compression needs streams in increasing, non-overlapping order");
    int rc = check_layout(in_off, in_len, in_span, nullptr, nullptr, 0, n);
    if (rc) return rc;
    if (lzs_b200_device_count() <= 0) return fail(LZS_B200_ENODEVICE, "no CUDA device: the LZS codec has no CPU path");
    HostPath *hp = nullptr;
    if ((rc = host_path(&hp))) return rc;
    std::lock_guard<std::mutex> lock(hp->mu);
    HostPath     &p = *hp;
    PipelineScope scope(p);
    cudaStream_t  st = p.stream;

    /* internal slots: worst-case size of every stream, 16-byte aligned */
    std::vector<uint64_t> slot_off(n);
    std::vector<uint32_t> slot_cap(n);
    uint64_t              slots = 0;
    for (uint32_t s2 = 0; s2 < n; s2++) {
        slot_off[s2] = slots;
        slot_cap[s2] = static_cast<uint32_t>(align_up(LZS_COMPRESSED_MAX(static_cast<size_t>(in_len[s2])), 16));
        slots += slot_cap[s2];
    }
    std::vector<uint32_t> first = plan_slices(in_len, n, in_span, true);
    const uint32_t nslice = static_cast<uint32_t>(first.size());
    first.push_back(n);

    if ((rc = p.reserve(S_IN, in_span + 64))) return rc;
    if ((rc = p.reserve(S_OUT, slots + 64))) return rc;
    if ((rc = p.reserve(S_PACKED, out_capacity + 64))) return rc;
    if ((rc = p.reserve(S_INOFF, n * sizeof(uint64_t)))) return rc;
    if ((rc = p.reserve(S_INLEN, n * sizeof(uint32_t)))) return rc;
    if ((rc = p.reserve(S_OUTOFF, n * sizeof(uint64_t)))) return rc;
    if ((rc = p.reserve(S_OUTCAP, n * sizeof(uint32_t)))) return rc;
    if ((rc = p.reserve(S_OUTLEN, n * sizeof(uint32_t)))) return rc;
    if ((rc = p.reserve(S_PACKOFF, n * sizeof(uint64_t)))) return rc;
    if ((rc = p.reserve(S_SCRATCH, lzs_b200_compress_scratch_bytes(in_span)))) return rc;
    if ((rc = p.reserve(S_COUNTERS, static_cast<size_t>(nslice) * 512))) return rc;
    if ((rc = p.reserve_pinned(n))) return rc;
    uint8_t  *d_in = static_cast<uint8_t *>(p.buf[S_IN]);
    uint8_t  *d_slots = static_cast<uint8_t *>(p.buf[S_OUT]);
    uint8_t  *d_packed = static_cast<uint8_t *>(p.buf[S_PACKED]);
    uint8_t  *d_cnt = static_cast<uint8_t *>(p.buf[S_COUNTERS]);
    uint64_t *d_inoff = static_cast<uint64_t *>(p.buf[S_INOFF]);
    uint32_t *d_inlen = static_cast<uint32_t *>(p.buf[S_INLEN]);
    uint64_t *d_slotoff = static_cast<uint64_t *>(p.buf[S_OUTOFF]);
    uint32_t *d_slotcap = static_cast<uint32_t *>(p.buf[S_OUTCAP]);
    uint32_t *d_outlen = static_cast<uint32_t *>(p.buf[S_OUTLEN]);
    uint64_t *d_packoff = static_cast<uint64_t *>(p.buf[S_PACKOFF]);
    uint16_t *matches = reinterpret_cast<uint16_t *>(static_cast<uint8_t *>(p.buf[S_SCRATCH]) + kCounterBytes);
    uint32_t       piece = 0;
    const uint32_t piece_cap = cut_into_pieces(n, in_span, p.cap[S_SCRATCH], &piece);

    CUDA_TRY(cudaMemcpyAsync(d_inoff, in_off, n * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_inlen, in_len, n * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_slotoff, slot_off.data(), n * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_slotcap, slot_cap.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    std::vector<cudaEvent_t> ev(nslice + 1), up_ev(nslice);
    for (auto &e : ev) if ((rc = scope.make(&e, cudaEventDisableTiming))) return rc;
    for (auto &e : up_ev) if ((rc = scope.make(&e, cudaEventDisableTiming))) return rc;
    CUDA_TRY(cudaEventRecord(ev[nslice], st));
    for (auto &w : p.work) CUDA_TRY(cudaStreamWaitEvent(w, ev[nslice], 0));
    CUDA_TRY(cudaStreamWaitEvent(p.up, ev[nslice], 0));
    /* LZS_B200_TRACE=1 prints when every stage of every slice finished (ms from the start) */
    const bool               trace = getenv("LZS_B200_TRACE") != nullptr;
    std::vector<cudaEvent_t> tr(trace ? nslice * 6 + 1 : 0);
    for (auto &e : tr) if ((rc = scope.make(&e, cudaEventDefault))) return rc;
    auto mark = [&](uint32_t k, int j, cudaStream_t s2) { if (trace) cudaEventRecord(tr[k * 6 + j], s2); };
    if (trace) cudaEventRecord(tr[nslice * 6], st);
    for (uint32_t k = 0; k < nslice && !rc; k++) {
        const uint32_t a = first[k], b = first[k + 1], cnt = b - a;
        cudaStream_t   ws = p.work[0];             /* kernels in slice order on one stream: see run_host_batch */
        const uint64_t lo = in_off[a], hi = in_off[b - 1] + in_len[b - 1];
        mark(k, 0, p.up);
        if (hi > lo) CUDA_TRY(cudaMemcpyAsync(d_in + lo, in + lo, hi - lo, cudaMemcpyHostToDevice, p.up));
        mark(k, 1, p.up);
        CUDA_TRY(cudaEventRecord(up_ev[k], p.up));
        CUDA_TRY(cudaStreamWaitEvent(ws, up_ev[k], 0));
        if (piece_cap) {                           /* few long streams: cut into pieces, as in run_host_batch */
            rc = compress_pieces(d_in, d_inoff + a, d_inlen + a, d_slots, d_slotoff + a, d_slotcap + a, d_outlen + a, cnt,
                                 piece, piece_cap, reinterpret_cast<uint32_t *>(d_cnt + k * 512), matches,
                                 static_cast<uint8_t *>(p.buf[S_SCRATCH]) + kCounterBytes + matches_bytes(in_span), ws);
            mark(k, 2, ws);
        } else {
            rc = lzs_b200_match_batch_device(d_in, d_inoff + a, d_inlen + a, matches, cnt,
                                             reinterpret_cast<uint32_t *>(d_cnt + k * 512), ws);
            mark(k, 2, ws);
        }
        if (!rc && !piece_cap && parse_side_streams() > 0) {
            CUDA_TRY(cudaEventRecord(up_ev[k], ws));                    /* reused: the upload has been waited for */
            ws = p.work[1 + k % static_cast<uint32_t>(parse_side_streams())];
            CUDA_TRY(cudaStreamWaitEvent(ws, up_ev[k], 0));
        }
        if (!rc && !piece_cap)
            rc = lzs_b200_parse_pack_batch_device(d_in, d_inoff + a, d_inlen + a, matches, d_slots, d_slotoff + a,
                                                  d_slotcap + a, d_outlen + a, cnt, ws);
        if (rc) break;
        mark(k, 3, ws);
        publish_lengths_kernel<<<(cnt + 255u) / 256u, 256, 0, ws>>>(p.pinned_len_dev + a, d_outlen + a, cnt);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaEventRecord(ev[k], ws));
    }
    uint64_t cursor = 0;
    for (uint32_t k = 0; k < nslice && !rc; k++) {
        const uint32_t a = first[k], b = first[k + 1], cnt = b - a;
        CUDA_TRY(cudaEventSynchronize(ev[k]));
        memcpy(out_len + a, p.pinned_len + a, cnt * sizeof(uint32_t));
        const uint64_t begin = cursor;
        for (uint32_t s2 = a; s2 < b; s2++) {
            out_off[s2] = cursor;
            cursor += align_up(out_len[s2], 16);
        }
        if (cursor > out_capacity) {
            rc = fail(LZS_B200_EINVAL, "packed output needs more than the %llu bytes offered",
                      static_cast<unsigned long long>(out_capacity));
            break;
        }
        CUDA_TRY(cudaMemcpyAsync(d_packoff + a, out_off + a, cnt * sizeof(uint64_t), cudaMemcpyHostToDevice, p.down));
        mark(k, 4, p.down);
        gather_streams_kernel<<<cnt, 128, 0, p.down>>>(d_slots, d_slotoff + a, d_outlen + a, d_packed, d_packoff + a, cnt);
        g_launches++;
        CUDA_TRY(cudaGetLastError());
        if (cursor > begin)
            CUDA_TRY(cudaMemcpyAsync(out + begin, d_packed + begin, cursor - begin, cudaMemcpyDeviceToHost, p.down));
        mark(k, 5, p.down);
    }
    if (rc) return rc;
    CUDA_TRY(scope.drain());
    if (trace) {
        for (uint32_t k = 0; k < nslice; k++) {
            float t[6];
            for (int j = 0; j < 6; j++) cudaEventElapsedTime(&t[j], tr[nslice * 6], tr[k * 6 + j]);
            fprintf(stderr, "lzs (b200) trace: slice %u, %u streams: upload %.2f-%.2f  K1 -%.2f  K2K3 -%.2f | offsets up + gather from %.2f, download done %.2f\n",
                    k, first[k + 1] - first[k], t[0], t[1], t[2], t[3], t[4], t[5]);
        }
    }
    if (out_used) *out_used = cursor;
    return LZS_B200_OK;
}

/* single stream through the host path; loud on failure, 0 bytes as the reference's
 * only "error" value */
size_t single_call(bool decompress, uint8_t *out, size_t out_cap, const uint8_t *in, size_t in_len)
{
    if (in_len > 0xFFFFFF00ull || out_cap > 0xFFFFFF00ull) {
        /* the batch ABI uses 32-bit stream lengths; clamp the capacity, refuse huge inputs */
        if (in_len > 0xFFFFFF00ull) {
            fprintf(stderr, "lzs (b200): single-call input of %zu bytes exceeds the 4 GiB stream limit\n", in_len);
            return 0;
        }
        out_cap = 0xFFFFFF00ull;
    }
    const uint64_t in_off = 0, out_off = 0;
    const uint32_t ilen = static_cast<uint32_t>(in_len), ocap = static_cast<uint32_t>(out_cap);
    uint32_t       olen = 0;
    uint8_t        dummy = 0;
    int rc = run_host_batch(decompress, in ? in : &dummy, &in_off, &ilen, in_len, out ? out : &dummy, &out_off,
                            &ocap, &olen, out_cap, 1);
    if (rc) {
        fprintf(stderr, "lzs (b200): %s failed (%d): %s\n", decompress ? "lzs_decompress" : "lzs_compress", rc,
                lzs_b200_last_error());
        return 0;
    }
    return olen;
}

}  // namespace

void lzs_b200_release_incremental_arena();

extern "C" {

/* The host-pointer entry points (and the drop-in lzs.h calls on top of them) keep grow-only
 * device and pinned buffers per device between calls; this gives them back. */
int lzs_b200_release(void)
{
    if (lzs_b200_device_count() <= 0) return LZS_B200_OK;
    int       rc = LZS_B200_OK;
    HostPath *p = host_path_slot(&rc);
    if (!p) return rc;
    {
        std::lock_guard<std::mutex> lock(p->mu);
        p->release();
    }
    lzs_b200_release_incremental_arena();
    return LZS_B200_OK;
}

int lzs_b200_compress_batch_host(const uint8_t *in, const uint64_t *in_off, const uint32_t *in_len,
                                 uint64_t in_span, uint8_t *out, const uint64_t *out_off,
                                 const uint32_t *out_cap, uint32_t *out_len, uint64_t out_span,
                                 uint32_t n_streams)
{
    return run_host_batch(false, in, in_off, in_len, in_span, out, out_off, out_cap, out_len, out_span, n_streams);
}

int lzs_b200_compress_packed_host(const uint8_t *in, const uint64_t *in_off, const uint32_t *in_len,
                                  uint64_t in_span, uint8_t *out, uint64_t out_capacity, uint64_t *out_off,
                                  uint32_t *out_len, uint32_t n_streams, uint64_t *out_used)
{
    return run_compress_packed(in, in_off, in_len, in_span, out, out_capacity, out_off, out_len, n_streams, out_used);
}

/* lzs_b200_decompress_batch_host that also says how many bytes of every stream were read up to and including its
 * end marker -- what a caller needs to walk a buffer of several streams laid end to end without an index.
 * 0xFFFFFFFF where that is not known (the stream is short, not clean, or its output did not fit). */
int lzs_b200_decompress_used_batch_host(const uint8_t *in, const uint64_t *in_off, const uint32_t *in_len,
                                        uint64_t in_span, uint8_t *out, const uint64_t *out_off,
                                        const uint32_t *out_cap, uint32_t *out_len, uint32_t *in_used, uint64_t out_span,
                                        uint32_t n_streams)
{
    if (!in_used) return fail(LZS_B200_EINVAL, "null pointer");
    for (uint32_t s = 0; s < n_streams; s++) in_used[s] = 0xFFFFFFFFu;
    return run_host_batch(true, in, in_off, in_len, in_span, out, out_off, out_cap, out_len, out_span, n_streams, in_used);
}

int lzs_b200_decompress_batch_host(const uint8_t *in, const uint64_t *in_off, const uint32_t *in_len,
                                   uint64_t in_span, uint8_t *out, const uint64_t *out_off,
                                   const uint32_t *out_cap, uint32_t *out_len, uint64_t out_span,
                                   uint32_t n_streams)
{
    return run_host_batch(true, in, in_off, in_len, in_span, out, out_off, out_cap, out_len, out_span, n_streams);
}

/* ---- reference single-call entry points (c/src/liblzs/lzs.h:218, :224, :229) ---- */

size_t lzs_compress(uint8_t *a_pOutData, size_t a_outBufferSize, const uint8_t *a_pInData, size_t a_inLen)
{
    return single_call(false, a_pOutData, a_outBufferSize, a_pInData, a_inLen);
}

/* The reference's "simple" variant differs only in how it searches (brute force,
 * lzs-compression-simple.c:264-278); its output is byte-identical, so it is the same
 * engine here. */
size_t lzs_simple_compress(uint8_t *a_pOutData, size_t a_outBufferSize, const uint8_t *a_pInData, size_t a_inLen)
{
    return single_call(false, a_pOutData, a_outBufferSize, a_pInData, a_inLen);
}

size_t lzs_decompress(uint8_t *a_pOutData, size_t a_outBufferSize, const uint8_t *a_pInData, size_t a_inLen)
{
    return single_call(true, a_pOutData, a_outBufferSize, a_pInData, a_inLen);
}

}  // extern "C"
