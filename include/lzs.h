/*
 * lzs.h -- drop-in C interface of the B200 LZS codec.
 *
 * Binary and source compatible with the public interface of cmcqueen/lzs-compression
 * (reference header c/src/liblzs/lzs.h): the same ten entry points (lzs.h:218-232),
 * the same inline lzs_compress_init (lzs.h:239-242), the same status flags
 * (lzs.h:90-99, :170-178), the same size macros (lzs.h:57-81) and parameter structs
 * with the same public members at the same offsets and the same total sizes
 * (14432 / 2112 / 2096 bytes on LP64; reference lzs.h:101-134, :136-167, :180-211).
 * The reference documents everything after `status` as private; here that space is
 * an opaque scratch area owned by the library.
 *
 * Every function takes HOST pointers and is synchronous, like the reference.  The
 * work is done on the GPU (CUDA kernels for sm_100a); there is no CPU code path.
 * If no usable CUDA device is present the calls write a diagnostic to stderr and
 * return 0 bytes (incremental calls also set LZS_*_STATUS_ERROR).  For throughput
 * use the batch interface in lzs_b200.h, which keeps data resident on the device.
 */
#ifndef LZS_B200_LZS_H
#define LZS_B200_LZS_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- format limits (reference lzs.h:57-68) ---- */
#define LZS_MAX_LOOK_AHEAD_LEN      15u
#define LZS_MAX_HISTORY_SIZE        ((1u << 11u) - 1u)                  /* 2047 */
#define LZS_COMPRESS_HISTORY_SIZE   (LZS_MAX_HISTORY_SIZE + LZS_MAX_LOOK_AHEAD_LEN)
#define LZS_DECOMPRESS_HISTORY_SIZE LZS_MAX_HISTORY_SIZE
#define INPUT_HASH_SIZE             (1u << 12u)

/* ---- buffer bounds (reference lzs.h:75-81) ---- */
#define LZS_COMPRESSED_MAX(X)       ((X) + ((X) + 7u) / 8u + 3u)       /* 9/8 n + marker */
#define LZS_DECOMPRESSED_MAX(X)     ((X) * 16u)

typedef uint16_t lzs_input_hash_t;

typedef enum {
    LZS_C_STATUS_NONE                   = 0x00,
    LZS_C_STATUS_INPUT_STARVED          = 0x01,   /* all available input was read          */
    LZS_C_STATUS_INPUT_FINISHED         = 0x02,   /* all available input was read          */
    LZS_C_STATUS_END_MARKER             = 0x04,   /* the output now contains an end marker */
    LZS_C_STATUS_NO_OUTPUT_BUFFER_SPACE = 0x08,   /* output buffer is full                 */
    LZS_C_STATUS_ERROR                  = 0x10
} LzsCompressStatus_t;

typedef enum {
    LZS_D_STATUS_NONE                   = 0x00,
    LZS_D_STATUS_INPUT_STARVED          = 0x01,   /* input read; bits may remain queued    */
    LZS_D_STATUS_INPUT_FINISHED         = 0x02,   /* input read and fully consumed         */
    LZS_D_STATUS_END_MARKER             = 0x04,   /* an end marker was decoded             */
    LZS_D_STATUS_NO_OUTPUT_BUFFER_SPACE = 0x08,   /* output buffer is full                 */
    LZS_D_STATUS_ERROR                  = 0x10
} LzsDecompressStatus_t;

/*
 * Public members of every parameter block.  Set them before each incremental
 * call; the call advances the pointers, decreases the lengths by what it
 * consumed / produced and overwrites `status`.
 */
#define LZS_B200_PUBLIC_MEMBERS                                                              \
    const uint8_t *inPtr;      /* in: next input byte;     out: first unread input byte   */ \
    uint8_t       *outPtr;     /* in: output buffer;       out: one past last byte written */\
    size_t         inLength;   /* in: input bytes offered; out: input bytes left unread   */ \
    size_t         outLength;  /* in: output space;        out: output space left         */ \
    uint8_t        status      /* LzsCompressStatus_t / LzsDecompressStatus_t flags       */

#define LZS_B200_PUBLIC_BYTES (2u * sizeof(void *) + 2u * sizeof(size_t) + 1u)

typedef struct {
    LZS_B200_PUBLIC_MEMBERS;
    uint8_t lzs_private_[14432u - LZS_B200_PUBLIC_BYTES];   /* do not touch */
} LzsCompressParameters_t;

typedef struct {
    LZS_B200_PUBLIC_MEMBERS;
    uint8_t lzs_private_[2112u - LZS_B200_PUBLIC_BYTES];    /* do not touch */
} LzsSimpleCompressParameters_t;

typedef struct {
    LZS_B200_PUBLIC_MEMBERS;
    uint8_t lzs_private_[2096u - LZS_B200_PUBLIC_BYTES];    /* do not touch */
} LzsDecompressParameters_t;

/* ---- single call (reference lzs.h:218, :224, :229) ----
 * Return the number of bytes written.  A too-small output buffer yields the prefix
 * that fits (compress) or the bytes decoded so far (decompress); no error code. */
size_t lzs_compress(uint8_t *a_pOutData, size_t a_outBufferSize, const uint8_t *a_pInData, size_t a_inLen);
size_t lzs_simple_compress(uint8_t *a_pOutData, size_t a_outBufferSize, const uint8_t *a_pInData, size_t a_inLen);
size_t lzs_decompress(uint8_t *a_pOutData, size_t a_outBufferSize, const uint8_t *a_pInData, size_t a_inLen);

/* ---- incremental (reference lzs.h:220-222, :226-227, :231-232) ---- */
void   lzs_compress_init_quick(LzsCompressParameters_t *pParams);
void   lzs_compress_init_full(LzsCompressParameters_t *pParams);
size_t lzs_compress_incremental(LzsCompressParameters_t *pParams, bool add_end_marker);

void   lzs_simple_compress_init(LzsSimpleCompressParameters_t *pParams);
size_t lzs_simple_compress_incremental(LzsSimpleCompressParameters_t *pParams, bool add_end_marker);

void   lzs_decompress_init(LzsDecompressParameters_t *pParams);
size_t lzs_decompress_incremental(LzsDecompressParameters_t *pParams);

static inline void lzs_compress_init(LzsCompressParameters_t *pParams)
{
    lzs_compress_init_full(pParams);
}

#ifdef __cplusplus
}
#endif

#endif /* LZS_B200_LZS_H */
