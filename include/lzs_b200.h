/*
 * lzs_b200.h -- batch C ABI of the B200 LZS codec (extension of lzs.h).
 *
 * The reference library has one call per stream on host memory
 * (c/src/liblzs/lzs.h:218 lzs_compress, :229 lzs_decompress) and leaves batching to
 * the caller (c/src/utils/lzs-compress.c:82-134 loops over one file).  A GPU needs
 * thousands of independent streams per launch, so this header adds the batch form
 * of exactly those two calls: stream s is
 *      in  + in_off[s],  in_len[s]  bytes   ->   out + out_off[s], at most out_cap[s] bytes
 * and out_len[s] receives what lzs_compress / lzs_decompress would have returned
 * for that stream alone (including the truncated-prefix behaviour when out_cap[s]
 * is too small).  Streams are independent LZS streams, each ended by its own end
 * marker, byte-identical to the reference's output for the same boundaries.
 *
 * Plain C, plain pointers and sizes: bind it from cgo / JNI / ctypes / FFI as is.
 * All functions return 0 on success and a negative LZS_B200_E* code on failure;
 * lzs_b200_last_error() describes the last failure of the calling thread.
 * There is no CPU fallback: without a CUDA device every call fails with
 * LZS_B200_ENODEVICE.
 */
#ifndef LZS_B200_BATCH_H
#define LZS_B200_BATCH_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LZS_B200_OK         0
#define LZS_B200_ENODEVICE (-1)   /* no usable CUDA device / driver                     */
#define LZS_B200_ECUDA     (-2)   /* a CUDA runtime call or kernel failed               */
#define LZS_B200_EINVAL    (-3)   /* bad argument (null pointer, scratch too small ...) */
#define LZS_B200_ENOMEM    (-4)   /* device or host allocation failed                   */

/* Offsets handed to the *_device calls should be multiples of this for the
 * vectorised paths (other alignments work, byte by byte). */
#define LZS_B200_ALIGN 16u

const char *lzs_b200_last_error(void);
int         lzs_b200_device_count(void);

/* ------------------------------------------------------------------------------
 * Device-resident batches.  Every pointer below is a DEVICE pointer on the current
 * CUDA device; `stream` is a cudaStream_t (NULL = default stream).  Calls enqueue
 * work and return without synchronising.
 *
 * in_span = number of bytes of `in` covered by the batch, i.e. max(in_off+in_len);
 * the compressor keeps one 16-bit match record per covered input byte in scratch.
 * ---------------------------------------------------------------------------- */
size_t lzs_b200_compress_scratch_bytes(uint64_t in_span);
size_t lzs_b200_decompress_scratch_bytes(void);                     /* the minimum                         */
size_t lzs_b200_decompress_scratch_bytes_for(uint32_t n_streams);   /* with room for the decoder's launch order
                                                                       (streams of similar density share a warp;
                                                                       about 15 % faster on mixed batches)   */

/* batch form of lzs_compress (reference lzs.h:218) */
int lzs_b200_compress_batch_device(const uint8_t *in, const uint64_t *in_off, const uint32_t *in_len,
                                   uint64_t in_span, uint8_t *out, const uint64_t *out_off,
                                   const uint32_t *out_cap, uint32_t *out_len, uint32_t n_streams,
                                   void *scratch, size_t scratch_bytes, void *stream);

/* batch form of lzs_decompress (reference lzs.h:229); a compressed stream may be at most
 * 512 MiB - 256 B long (the decoder keeps 32-bit bit positions) */
int lzs_b200_decompress_batch_device(const uint8_t *in, const uint64_t *in_off, const uint32_t *in_len,
                                     uint8_t *out, const uint64_t *out_off, const uint32_t *out_cap,
                                     uint32_t *out_len, uint32_t n_streams, void *scratch,
                                     size_t scratch_bytes, void *stream);

/* The same with one status byte per stream saying why its decoding stopped, in the values of
 * LzsDecompressStatus_t (reference lzs.h:170-178; SURVEY.md section 8f-4, for untrusted packets):
 *   LZS_D_STATUS_END_MARKER (0x04)             an end marker was reached (the normal case);
 *   LZS_D_STATUS_NO_OUTPUT_BUFFER_SPACE (0x08) out_cap bytes were written and input was left;
 *   LZS_D_STATUS_INPUT_STARVED (0x01)          the input ended first: nothing left, or the
 *                                              remaining bits do not complete a token.
 * The bytes and lengths are those of lzs_b200_decompress_batch_device; `status` may be NULL. */
int lzs_b200_decompress_status_batch_device(const uint8_t *in, const uint64_t *in_off, const uint32_t *in_len,
                                            uint8_t *out, const uint64_t *out_off, const uint32_t *out_cap,
                                            uint32_t *out_len, uint8_t *status, uint32_t n_streams,
                                            void *scratch, size_t scratch_bytes, void *stream);

/* ------------------------------------------------------------------------------
 * Packets of flows with KEPT HISTORY, the bulk path (RFC 1974 style: the
 * reference resets neither its compressor's nor its decoder's history at an end marker,
 * lzs-compression.c:796-820, lzs-decompression.c:564-576).  Stream s is ONE packet; the caller keeps
 * a flow's packets contiguous in memory, and hist_len[s] (<= 2047, more is clamped) says how many
 * bytes in front of the packet are the flow's earlier packets.  The packet is compressed exactly as
 * init-once + lzs_compress_incremental(add_end_marker = true)-until-END_MARKER per packet does it:
 * matches may reach back into the history, tokens end with the packet, every packet ends with its
 * own end marker.  Compressing is parallel over ALL packets of ALL flows (one call); decoding is
 * serial inside a flow -- packet k needs packets < k of its flow decoded, in front of its own output
 * -- so the decoder is called once per packet index with all flows.  hist_len == NULL or all zero:
 * the plain batch calls.
 * ---------------------------------------------------------------------------- */
int lzs_b200_compress_flows_batch_device(const uint8_t *in, const uint64_t *in_off, const uint32_t *in_len,
                                         const uint32_t *hist_len, uint64_t in_span, uint8_t *out,
                                         const uint64_t *out_off, const uint32_t *out_cap, uint32_t *out_len,
                                         uint32_t n_streams, void *scratch, size_t scratch_bytes, void *stream);
int lzs_b200_decompress_flows_batch_device(const uint8_t *in, const uint64_t *in_off, const uint32_t *in_len,
                                           uint8_t *out, const uint64_t *out_off, const uint32_t *out_cap,
                                           const uint32_t *hist_len, uint32_t *out_len, uint8_t *status,
                                           uint32_t n_streams, void *scratch, size_t scratch_bytes, void *stream);
/* A table of flows whose packets have seg_len[f] bytes each (the last of a flow may be shorter): the match
 * finder takes flow f (in + flow_off[f], flow_len[f] bytes) as ONE stream and ends every position's
 * look-ahead with its packet, so a flow's bytes go through the tables once; the parse/pack kernel takes the
 * n_packets packets (pkt_off / pkt_len, each inside a flow) one by one.  Same bytes as the call above. */
int lzs_b200_compress_flow_table_device(const uint8_t *in, const uint64_t *flow_off, const uint32_t *flow_len,
                                        const uint32_t *seg_len, uint32_t n_flows, const uint64_t *pkt_off,
                                        const uint32_t *pkt_len, uint32_t n_packets, uint64_t in_span, uint8_t *out,
                                        const uint64_t *out_off, const uint32_t *out_cap, uint32_t *out_len,
                                        void *scratch, size_t scratch_bytes, void *stream);

/* Copies n streams from their slots (src + src_off[s], len[s] bytes) to packed positions
 * (dst + dst_off[s]); both offsets must be multiples of 16 and every slot readable up to the
 * next multiple of 16 of its length.  What the packed host compressor and the multi-GPU gather
 * of variable-size outputs use to put streams back to back. */
int lzs_b200_pack_streams_device(const uint8_t *src, const uint64_t *src_off, const uint32_t *len, uint8_t *dst,
                                 const uint64_t *dst_off, uint32_t n_streams, void *stream);

/* The pack kernel FUSED with the all-gather of the packed streams over NVLink peer memory (multi-GPU
 * gather of variable-size outputs).  dst_ptrs: device array of n_dst addresses,
 * one per rank of the box, of the SAME symmetric buffer as this GPU sees each rank's copy (peer
 * mappings, e.g. torch.distributed._symmetric_memory: buffer_ptrs_dev); every stream is read once and
 * stored at base + dst_off[s] in every copy.  The multicast form stores once per 16 bytes to the
 * buffer's NVSwitch multicast address (multimem.st) and lets the switch replicate.  The caller
 * separates the calls from the readers with a cross-rank barrier. */
int lzs_b200_pack_streams_peers_device(const uint8_t *src, const uint64_t *src_off, const uint32_t *len,
                                       uint8_t *const *dst_ptrs, uint32_t n_dst, uint64_t base, const uint64_t *dst_off,
                                       uint32_t n_streams, void *stream);
int lzs_b200_pack_streams_multicast_device(const uint8_t *src, const uint64_t *src_off, const uint32_t *len, uint8_t *mc_ptr,
                                           uint64_t base, const uint64_t *dst_off, uint32_t n_streams, void *stream);

/* Individual stages of the compressor, for tests and profiling:
 * K1 writes one record per input byte, (len << 11) | offset with len 0 or 2..12;
 * K2+K3 turn records + input into streams.  `counter` is device scratch of at least 16 bytes
 * (work counters and a status word, cleared by the call). */
int lzs_b200_match_batch_device(const uint8_t *in, const uint64_t *in_off, const uint32_t *in_len,
                                uint16_t *matches, uint32_t n_streams, uint32_t *counter, void *stream);
int lzs_b200_parse_pack_batch_device(const uint8_t *in, const uint64_t *in_off, const uint32_t *in_len,
                                     const uint16_t *matches, uint8_t *out, const uint64_t *out_off,
                                     const uint32_t *out_cap, uint32_t *out_len, uint32_t n_streams,
                                     void *stream);

/* ------------------------------------------------------------------------------
 * Host-resident batches: same meaning, HOST pointers, synchronous.  Input is
 * copied to the device, processed and copied back inside the call (pinned host
 * memory makes the copies faster but is not required).  in_span / out_span are the
 * number of bytes of `in` / `out` covered by the batch; a stream or slot that reaches
 * beyond them is refused (LZS_B200_EINVAL), nothing is processed.
 * What the call writes to `out`: the out_len[s] bytes of every stream.  The rest of a
 * slot, [out_off[s] + out_len[s], out_off[s] + out_cap[s]), is unspecified afterwards;
 * bytes outside every slot (framing the caller keeps between slots) are never touched.
 * The calls keep grow-only device buffers between calls; lzs_b200_release() frees them.
 * ---------------------------------------------------------------------------- */
int lzs_b200_compress_batch_host(const uint8_t *in, const uint64_t *in_off, const uint32_t *in_len,
                                 uint64_t in_span, uint8_t *out, const uint64_t *out_off,
                                 const uint32_t *out_cap, uint32_t *out_len, uint64_t out_span,
                                 uint32_t n_streams);
int lzs_b200_decompress_batch_host(const uint8_t *in, const uint64_t *in_off, const uint32_t *in_len,
                                   uint64_t in_span, uint8_t *out, const uint64_t *out_off,
                                   const uint32_t *out_cap, uint32_t *out_len, uint64_t out_span,
                                   uint32_t n_streams);

/* The same, and in_used[s] = bytes of stream s read up to and including its end marker: what a caller
 * needs to walk a buffer of several streams laid end to end without an index (the reference's file
 * format, c/src/utils/lzs-decompress.c:75-118).  0xFFFFFFFF where that is not known (short streams, streams
 * that are not clean, outputs that did not fit): decode those the slow way. */
int lzs_b200_decompress_used_batch_host(const uint8_t *in, const uint64_t *in_off, const uint32_t *in_len,
                                        uint64_t in_span, uint8_t *out, const uint64_t *out_off,
                                        const uint32_t *out_cap, uint32_t *out_len, uint32_t *in_used,
                                        uint64_t out_span, uint32_t n_streams);

/* Packed variant of lzs_b200_compress_batch_host for streams given in increasing order (chunks of
 * a file, a packet table): the streams are written back to back, each starting at a multiple of
 * 16, and out_off[s] / out_len[s] are OUTPUTS.  Only compressed bytes travel back over PCIe, and
 * the concatenation (minus the padding) is what the reference's lzs-decompress reads marker by
 * marker.  out_capacity: LZS_COMPRESSED_MAX of every stream rounded up to 16 always suffices.
 * *out_used (may be NULL) receives the bytes of `out` covered. */
int lzs_b200_compress_packed_host(const uint8_t *in, const uint64_t *in_off, const uint32_t *in_len,
                                  uint64_t in_span, uint8_t *out, uint64_t out_capacity, uint64_t *out_off,
                                  uint32_t *out_len, uint32_t n_streams, uint64_t *out_used);

/* ------------------------------------------------------------------------------
 * Batch form of the incremental calls of lzs.h (reference lzs.h:222, :232): advance
 * n independent caller-owned streams by ONE lzs_compress_incremental /
 * lzs_decompress_incremental call each, in a single launch (one warp per stream).
 * Each block is used exactly as for the single call: set inPtr/inLength/outPtr/
 * outLength before, read them and `status` after.  produced[s] (may be NULL) is what
 * the single call would have returned.  For whole packets (init, then one call with
 * add_end_marker until LZS_C_STATUS_END_MARKER) lzs_b200_compress_batch_* gives the
 * same bytes and is much faster.
 * ---------------------------------------------------------------------------- */
#ifdef LZS_B200_LZS_H
int lzs_b200_compress_incremental_batch(LzsCompressParameters_t **params, uint32_t n, int add_end_marker,
                                        size_t *produced);
int lzs_b200_decompress_incremental_batch(LzsDecompressParameters_t **params, uint32_t n, size_t *produced);
#endif

/* ------------------------------------------------------------------------------
 * The same with everything resident on the device (packet tables of flows: BASELINE configs[2]
 * "via the incremental API"): n state blocks in device memory,
 * lzs_b200_incremental_state_bytes() each (a multiple of 16; blocks 16-byte aligned, `stride`
 * bytes apart), set up by lzs_b200_incremental_init_device (= lzs_compress_init /
 * lzs_decompress_init, lzs.h:220-232), and a DEVICE array of job records, one per stream:
 * `state`, `in`, `in_len`, `out`, `out_cap` and `add_end_marker` are the caller's inputs (the
 * inPtr/inLength/outPtr/outLength of the reference's parameter block and the flag of
 * lzs_compress_incremental); the call fills `in_used`, `out_used` (= the reference's return
 * value) and `status` (LzsCompressStatus_t / LzsDecompressStatus_t bits).  One launch, one warp
 * per stream, nothing crosses PCIe; the caller advances its pointers by in_used / out_used and
 * calls again until LZS_C_STATUS_END_MARKER, exactly as c/src/utils/lzs-compress.c:91-134 does.
 * The history survives end markers, so a flow's packets share it (RFC 1974 style).
 * ---------------------------------------------------------------------------- */
typedef struct {
    void          *state;
    const uint8_t *in;
    uint8_t       *out;
    uint32_t       in_len, out_cap;
    uint32_t       in_used, out_used, status;
    uint32_t       add_end_marker;
} lzs_b200_inc_job_t;

size_t lzs_b200_incremental_state_bytes(int decompress);
int    lzs_b200_incremental_init_device(void *states, size_t stride, uint32_t n_streams, int decompress, void *stream);
int    lzs_b200_compress_incremental_batch_device(lzs_b200_inc_job_t *jobs, uint32_t n_streams, void *stream);
int    lzs_b200_decompress_incremental_batch_device(lzs_b200_inc_job_t *jobs, uint32_t n_streams, void *stream);

/* Uniform chunking helpers (host arrays): stream s covers [s*chunk, min((s+1)*chunk, total))
 * and its output slot starts at s*out_stride. */
uint32_t lzs_b200_chunk_count(uint64_t total, uint32_t chunk);
void     lzs_b200_chunk_layout(uint64_t total, uint32_t chunk, uint64_t out_stride, uint64_t *in_off,
                               uint32_t *in_len, uint64_t *out_off, uint32_t *out_cap);

/* Synthetic corpus generated in place on the device:
 * n streams of stream_len bytes at dst + s*stride, stream index first_index + s. */
int lzs_b200_corpus_fill_device(uint8_t *dst, uint64_t stride, uint32_t stream_len, uint64_t first_index,
                                uint64_t n, uint64_t seed, int kind, void *stream);

/* Frees the device and pinned buffers the host-pointer entry points (and the lzs.h drop-in
 * calls, which sit on them) keep between calls on the current device.  Safe at any time
 * between calls; the next call allocates again. */
int lzs_b200_release(void);

/* Tuning knobs (also read once from the environment: LZS_B200_DECODE_LANES). */
int lzs_b200_set_decode_lanes(int lanes_per_stream);   /* 4, 8, 16 or 32 */
/* Option (also LZS_B200_ZEROCOPY=1): lzs_b200_decompress_batch_host writes straight into an output
 * buffer that is pinned and device mapped instead of staging it in device memory and copying.
 * Off by default: measured slightly slower on B200 / PCIe gen 5 (31 vs 29 ms per GiB). */
int lzs_b200_set_zero_copy_output(int on);
/* Test knob: run the match finder's exact-for-any-exchange-order launch after every fast launch
 * (normally it returns at once: sm_100a serves the exchanges in the order the fast launch
 * assumes).  Same records, several times slower. */
int lzs_b200_set_force_safe_match(int on);

/* Long streams.  The parse of an LZS stream is serial (lzs-compression.c:301-447: where a token
 * starts depends on where the one before ended), and one warp per stream leaves the GPU empty when a
 * batch has few streams -- one lzs_compress call on a large buffer has one.  A batch of at most 2048
 * streams that average two pieces or more (in_span >= 2 * piece * n_streams) is therefore cut: every
 * stream into pieces of `piece` bytes, matches found and pieces parsed in parallel, the pieces'
 * tokens stitched into the one stream the reference produces, byte for byte (csrc/k23_pieces.cuh).
 * Default 65536 (environment: LZS_B200_PIECE); 0 turns the cutting off, 64 .. 2^28 sets the piece
 * size.  lzs_b200_compress_scratch_bytes() includes the piece table for the setting in force when it
 * is called; with less scratch than that the streams are not cut.  Streams of a batch that is cut
 * must not overlap in `in`.  Applies to lzs_b200_compress_batch_device, the host batch calls and
 * lzs_compress / lzs_simple_compress; the decoder has no counterpart. */
int lzs_b200_set_piece_bytes(uint32_t bytes);

/* The decoder's counterpart (csrc/k4_pieces.cuh).  Where the tokens of a stream start is found in
 * parallel inside the stream (pieces of `bytes` COMPRESSED bytes, default 2048, environment
 * LZS_B200_DPIECE, 0 = off), literals go straight to their place and the matches are replayed in
 * order by one warp per stream -- the one serial step that is left, a copy per match instead of a
 * bit parser.  Used by lzs_b200_decompress_batch_device / _status_batch_device for batches of at most
 * 4096 streams whose scratch has room for the piece table, i.e. was sized with
 * lzs_b200_decompress_scratch_bytes_long(in_span, n_streams) (in_span = compressed bytes covered by the
 * batch); with less scratch every stream is decoded by one group of lanes, as before.  The host batch
 * call and lzs_decompress do this themselves.  Malformed streams and outputs that are too small
 * come out exactly as before: the piece passes hand such streams to the serial decoder. */
size_t lzs_b200_decompress_scratch_bytes_long(uint64_t in_span, uint32_t n_streams);
int    lzs_b200_set_decode_piece_bytes(uint32_t bytes);

/* A HANDFUL of long streams -- one lzs_decompress call on a large buffer is the case -- cannot keep the
 * GPU busy even with one thread block replaying each stream's matches (~140 MB/s per stream).  For
 * them the copies are not replayed at all: every output byte gets a pointer to the byte it is a copy
 * of, the pointers are doubled until each points at a literal, and every byte is fetched from its
 * literal (csrc/k4_pieces.cuh: k4j_*).  O(n log n) work and 4 bytes of scratch per byte of out_span,
 * but parallel over the bytes of a stream: one stream of 256 MiB decodes in 80 ms instead of 1.9 s.
 * out_span = bytes of `out` the slots cover, at most 2^31.  Same bytes, lengths and stop reasons as
 * lzs_b200_decompress_status_batch_device (status may be NULL).  lzs_decompress and the host batch call
 * use it for up to 128 streams (environment: LZS_B200_JUMP_STREAMS, 0 = never). */
size_t lzs_b200_decompress_scratch_bytes_jump(uint64_t in_span, uint64_t out_span, uint32_t n_streams);
int    lzs_b200_decompress_long_batch_device(const uint8_t *in, const uint64_t *in_off, const uint32_t *in_len,
                                             uint8_t *out, const uint64_t *out_off, const uint32_t *out_cap,
                                             uint64_t out_span, uint32_t *out_len, uint8_t *status,
                                             uint32_t n_streams, void *scratch, size_t scratch_bytes, void *stream);

/* Number of kernels launched by this library in the calling process so far. */
uint64_t lzs_b200_kernel_launches(void);

#ifdef __cplusplus
}
#endif

#endif /* LZS_B200_BATCH_H */
