/*
 * chunk_driver.c -- run a CPU LZS codec over a batch of independent streams,
 * optionally on several host threads, and time it.
 *
 * TEST / BASELINE INFRASTRUCTURE (see oracle/lzs_oracle.c header).  Built twice
 * by oracle/Makefile:
 *   - with -DLZS_DRIVER_REFERENCE and the unmodified reference sources from
 *     /root/reference/c/src/liblzs  ->  oracle/_ref/liblzs_ref.so
 *     (CPU baseline kind "reference"; also exports the reference's own symbols)
 *   - with oracle/lzs_oracle.c       ->  oracle/liblzs_oracle.so
 *     (CPU baseline kind "port")
 *
 * Work split follows BASELINE.md section 4: identical stream boundaries to the
 * GPU run, one lzs_compress / lzs_decompress call per stream, streams handed to
 * threads in static contiguous ranges, CLOCK_MONOTONIC around the whole batch.
 * The reference reads in[inLen] while hashing (lzs-compression.c:437), so the
 * caller must leave one readable byte after the last stream.
 */
#define _POSIX_C_SOURCE 200809L
#include <pthread.h>
#include <stddef.h>
#include <stdint.h>
#include <time.h>

#ifdef LZS_DRIVER_REFERENCE
#include "lzs.h"            /* the reference's header, found via -I           */
#define CODEC_COMPRESS   lzs_compress
#define CODEC_DECOMPRESS lzs_decompress
#else
size_t lzs_oracle_compress(uint8_t *, size_t, const uint8_t *, size_t);
size_t lzs_oracle_decompress(uint8_t *, size_t, const uint8_t *, size_t);
#define CODEC_COMPRESS   lzs_oracle_compress
#define CODEC_DECOMPRESS lzs_oracle_decompress
#endif

typedef struct {
    const uint8_t  *in;
    const uint64_t *in_off;
    const uint32_t *in_len;
    uint8_t        *out;
    const uint64_t *out_off;
    const uint32_t *out_cap;
    uint32_t       *out_len;
    uint32_t        first, last;     /* [first, last) */
    int             decompress;
} job_t;

static void *run_job(void *arg)
{
    job_t   *j = (job_t *)arg;
    uint32_t s;
    for (s = j->first; s < j->last; s++) {
        const uint8_t *src = j->in + j->in_off[s];
        uint8_t       *dst = j->out + j->out_off[s];
        size_t         r;
        if (j->decompress) r = CODEC_DECOMPRESS(dst, j->out_cap[s], src, j->in_len[s]);
        else               r = CODEC_COMPRESS(dst, j->out_cap[s], src, j->in_len[s]);
        j->out_len[s] = (uint32_t)r;
    }
    return NULL;
}

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

#define MAX_THREADS 256

static double run_batch(int decompress,
                        const uint8_t *in, const uint64_t *in_off, const uint32_t *in_len,
                        uint8_t *out, const uint64_t *out_off, const uint32_t *out_cap,
                        uint32_t *out_len, uint32_t n, int threads)
{
    pthread_t tid[MAX_THREADS];
    job_t     job[MAX_THREADS];
    double    t0, t1;
    int       t;

    if (threads < 1) threads = 1;
    if (threads > MAX_THREADS) threads = MAX_THREADS;
    if ((uint32_t)threads > n && n > 0) threads = (int)n;

    for (t = 0; t < threads; t++) {
        job[t].in = in; job[t].in_off = in_off; job[t].in_len = in_len;
        job[t].out = out; job[t].out_off = out_off; job[t].out_cap = out_cap;
        job[t].out_len = out_len;
        job[t].first = (uint32_t)(((uint64_t)n * (uint64_t)t) / (uint64_t)threads);
        job[t].last = (uint32_t)(((uint64_t)n * (uint64_t)(t + 1)) / (uint64_t)threads);
        job[t].decompress = decompress;
    }
    t0 = now_s();
    if (threads == 1) {
        run_job(&job[0]);
    } else {
        for (t = 0; t < threads; t++) pthread_create(&tid[t], NULL, run_job, &job[t]);
        for (t = 0; t < threads; t++) pthread_join(tid[t], NULL);
    }
    t1 = now_s();
    return t1 - t0;
}

/* Both return elapsed wall seconds for the whole batch. */
double lzsdrv_compress_streams(const uint8_t *in, const uint64_t *in_off, const uint32_t *in_len,
                               uint8_t *out, const uint64_t *out_off, const uint32_t *out_cap,
                               uint32_t *out_len, uint32_t n, int threads)
{
    return run_batch(0, in, in_off, in_len, out, out_off, out_cap, out_len, n, threads);
}

double lzsdrv_decompress_streams(const uint8_t *in, const uint64_t *in_off, const uint32_t *in_len,
                                 uint8_t *out, const uint64_t *out_off, const uint32_t *out_cap,
                                 uint32_t *out_len, uint32_t n, int threads)
{
    return run_batch(1, in, in_off, in_len, out, out_off, out_cap, out_len, n, threads);
}

/* 1 when this build wraps the unmodified reference, 0 for the restatement. */
int lzsdrv_is_reference(void)
{
#ifdef LZS_DRIVER_REFERENCE
    return 1;
#else
    return 0;
#endif
}
