/*
 * lzs_common.cuh -- LZS format constants and small device helpers shared by the
 * sm_100a kernels.  Constants restate c/src/liblzs/lzs-common.h:38-53,
 * lzs.h:57-81 and lzs-compression.c:62 of the reference.
 *
 * The same source compiles under the CPU emulator used by the non-GPU tests
 * (tests/simt/simt.h, -DLZS_SIMT_EMU); the product build is nvcc only.
 */
#ifndef LZS_B200_COMMON_CUH
#define LZS_B200_COMMON_CUH

#include <stdint.h>

#ifdef LZS_SIMT_EMU
#include "simt.h"
#define LZS_DYN_SMEM(type, name) type *name = reinterpret_cast<type *>(simt::dyn_smem())
#define LZS_SPIN_HINT() simt_yield()
#else
#include <cuda_runtime.h>
#define LZS_DYN_SMEM(type, name) extern __shared__ __align__(16) type name[]
#define LZS_SPIN_HINT() __nanosleep(20)
#endif

#define LZS_FULL_MASK 0xFFFFFFFFu

namespace lzs {

constexpr uint32_t kWindow = 2047;        /* LZS_MAX_HISTORY_SIZE, lzs.h:60            */
constexpr uint32_t kSearchMax = 12;       /* LZS_SEARCH_MATCH_MAX, lzs-compression.c:62 */
constexpr uint32_t kMinLen = 2;           /* MIN_LENGTH, lzs-common.h:51               */
constexpr uint32_t kMaxShortLen = 8;      /* MAX_SHORT_LENGTH, lzs-common.h:52         */
constexpr uint32_t kMaxExtLen = 15;       /* MAX_EXTENDED_LENGTH, lzs-common.h:53      */
constexpr uint32_t kShortOffMax = 127;    /* SHORT_OFFSET_MAX, lzs-common.h:43         */

/* Per-position match record written by K1 and read by K2: (len << 11) | offset,
 * len in 0 or 2..12 (capped at LZS_SEARCH_MATCH_MAX), offset in 1..2047. */
typedef uint16_t match_t;
constexpr uint32_t kMatchOffBits = 11;

__device__ __forceinline__ uint32_t umin32(uint32_t a, uint32_t b) { return a < b ? a : b; }
__device__ __forceinline__ uint32_t umax32(uint32_t a, uint32_t b) { return a > b ? a : b; }

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

/* Byte swap: the stream is MSB-first, the GPU is little endian. */
__device__ __forceinline__ uint32_t bswap32(uint32_t v) { return __byte_perm(v, 0, 0x0123); }

/* Four bytes starting at byte address `p` (any alignment), built from aligned
 * 32-bit loads only.  Words that lie entirely at or beyond `end` are not touched
 * and read as zero, so nothing outside the buffer's own aligned words is read. */
__device__ __forceinline__ uint32_t load4_unaligned(const uint8_t *p, const uint8_t *end)
{
    uintptr_t       a = reinterpret_cast<uintptr_t>(p);
    const uint32_t *w = reinterpret_cast<const uint32_t *>(a & ~static_cast<uintptr_t>(3));
    uint32_t        sh = static_cast<uint32_t>(a & 3u) * 8u;
    uint32_t        lo = 0, hi = 0;
    if (reinterpret_cast<const uint8_t *>(w) < end) lo = __ldg(w);
    if (sh != 0 && reinterpret_cast<const uint8_t *>(w + 1) < end) hi = __ldg(w + 1);
    return __funnelshift_r(lo, hi, sh);
}

/* Orders this thread's earlier shared-memory accesses before its later ones as seen by the
 * other threads of the block (release before publishing a counter, acquire after reading one). */
__device__ __forceinline__ void cta_fence()
{
#ifndef LZS_SIMT_EMU
    asm volatile("fence.acq_rel.cta;" ::: "memory");
#endif
}

/* Body of a spin-wait on a shared-memory flag. */
__device__ __forceinline__ void spin_pause()
{
#ifdef LZS_SIMT_EMU
    simt_yield();
#else
    __nanosleep(20);
#endif
}

/* The same with a back-off: a warp that polls steals issue slots from the warps it is waiting
 * for, so the pause doubles with every unsuccessful poll (`ns` is the caller's, reset per wait). */
#ifndef LZS_SPIN_NS_MIN
#define LZS_SPIN_NS_MIN 32
#endif
#ifndef LZS_SPIN_NS_MAX
#define LZS_SPIN_NS_MAX 64
#endif
__device__ __forceinline__ void spin_backoff(uint32_t &ns)
{
#ifdef LZS_SIMT_EMU
    (void)ns;
    simt_yield();
#else
    __nanosleep(ns);
    if (ns < LZS_SPIN_NS_MAX) ns *= 2u;
#endif
}

/* Shared-memory arrival barriers (PTX mbarrier): `count` arrivals complete a phase; any number
 * of threads may wait for a phase without being counted, so producers and consumers of a
 * hand-off never have to meet among themselves.  arrive has release, wait has acquire semantics
 * at CTA scope.  Phases are used strictly in turn (a phase cannot complete twice before every
 * waiter of the previous one has seen it -- the protocol of the caller guarantees that), so a
 * waiter only needs the parity of the phase it waits for. */
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
#ifdef LZS_SIMT_EMU
    *bar = (static_cast<uint64_t>(count) << 32) | count;      /* [63] phase, [62:32] count, [31:0] pending */
#else
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(bar))),
                 "r"(count)
                 : "memory");
#endif
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
#ifdef LZS_SIMT_EMU
    uint64_t v = *bar;
    uint32_t pending = static_cast<uint32_t>(v) - 1u;
    if (pending == 0u) {
        v ^= 1ull << 63;
        pending = static_cast<uint32_t>(v >> 32) & 0x7FFFFFFFu;
    }
    *bar = (v & 0xFFFFFFFF00000000ull) | pending;
#else
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(
                     static_cast<uint32_t>(__cvta_generic_to_shared(bar)))
                 : "memory");
#endif
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
#ifdef LZS_SIMT_EMU
    while (static_cast<uint32_t>(*reinterpret_cast<volatile uint64_t *>(bar) >> 63) == (parity & 1u)) simt_yield();
#else
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "LZS_MBAR_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra LZS_MBAR_DONE;\n\t"
        "bra LZS_MBAR_WAIT;\n\t"
        "LZS_MBAR_DONE:\n\t}" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(bar))),
        "r"(parity & 1u)
        : "memory");
#endif
}

}  // namespace lzs

#endif /* LZS_B200_COMMON_CUH */
