/*
 * corpus_host.c -- host build of the synthetic corpus generator (tests and the
 * CPU-baseline leg of bench.py).  TEST INFRASTRUCTURE; the generator itself lives
 * in lzs-compression_b200/csrc/corpus.h so the device build makes the same bytes.
 */
#include <stddef.h>
#include <stdint.h>
#include "../lzs-compression_b200/csrc/corpus.h"

/* Fill n streams of stream_len bytes laid out back to back with `stride`. */
void lzs_corpus_fill_host(uint8_t *dst, uint64_t stride, uint32_t stream_len, uint64_t first_index,
                          uint64_t n, uint64_t seed, int kind)
{
    uint64_t s;
    for (s = 0; s < n; s++) {
        lzs_corpus_fill(dst + s * stride, stream_len, seed, first_index + s, kind);
    }
}
