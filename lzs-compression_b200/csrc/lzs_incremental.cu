/*
 * lzs_incremental.cu -- the incremental half of include/lzs.h
 * (reference prototypes c/src/liblzs/lzs.h:220-232) plus its batch form.
 *
 * The codec state lives in the caller's parameter block, as in the reference.  A
 * call copies the state, the offered input and nothing else to the device, runs one
 * warp per stream (incremental.cuh) and copies back the state, the produced bytes
 * and the bookkeeping (bytes consumed / produced, status flags).  There is no CPU
 * implementation: without a device the call reports LZS_*_STATUS_ERROR and moves
 * nothing.
 */
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/lzs.h"
#include "../../include/lzs_b200.h"
#include "incremental.cuh"

static_assert(sizeof(LzsCompressParameters_t) == 14432, "reference layout: lzs.h:101-134");
static_assert(sizeof(LzsSimpleCompressParameters_t) == 2112, "reference layout: lzs.h:136-167");
static_assert(sizeof(LzsDecompressParameters_t) == 2096, "reference layout: lzs.h:180-211");
/* bytes of state that travel: header + ring (the structs' tail padding does not) */
constexpr size_t kIncCBytes = offsetof(lzs::IncCompressState, ring) + lzs::kIncCRing;
constexpr size_t kIncDBytes = offsetof(lzs::IncDecompressState, ring) + lzs::kIncDRing;
static_assert(kIncCBytes <= sizeof(((LzsSimpleCompressParameters_t *)0)->lzs_private_),
              "compress state must fit the smaller parameter block");
static_assert(kIncDBytes <= sizeof(((LzsDecompressParameters_t *)0)->lzs_private_), "decompress state must fit");

namespace {

constexpr size_t kPieceMax = 1u << 30;          /* per device pass; larger slices are looped */

struct Arena {
    std::mutex     mu;
    std::once_flag once;
    cudaStream_t stream = nullptr;
    void        *buf[4] = {};
    size_t       cap[4] = {};
    bool reserve(int slot, size_t bytes)
    {
        bytes = (bytes + 255) / 256 * 256 + 256;
        if (cap[slot] >= bytes) return true;
        if (buf[slot]) cudaFree(buf[slot]);
        buf[slot] = nullptr;
        cap[slot] = 0;
        if (cudaMalloc(&buf[slot], bytes) != cudaSuccess) { cudaGetLastError(); return false; }
        cap[slot] = bytes;
        return true;
    }
};

Arena *arena()
{
    static Arena a[64];
    int          dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { cudaGetLastError(); return nullptr; }
    Arena &x = a[dev];
    std::call_once(x.once, [&x] {
        if (cudaStreamCreateWithFlags(&x.stream, cudaStreamNonBlocking) != cudaSuccess) {
            cudaGetLastError();
            x.stream = nullptr;
        }
    });
    return x.stream ? &x : nullptr;
}

enum { A_STATE, A_IN, A_OUT, A_JOBS };

}  // namespace

/* lzs_b200_release(): the incremental calls' share (lzs_b200.cu calls this) */
void lzs_b200_release_incremental_arena()
{
    Arena *ap = arena();
    if (!ap) return;
    std::lock_guard<std::mutex> lock(ap->mu);
    for (int i = 0; i < 4; i++) {
        if (ap->buf[i]) cudaFree(ap->buf[i]);
        ap->buf[i] = nullptr;
        ap->cap[i] = 0;
    }
}

namespace {

struct View {                 /* the public members of any parameter block (lzs.h) */
    const uint8_t **in_ptr;
    uint8_t       **out_ptr;
    size_t         *in_len;
    size_t         *out_len;
    uint8_t        *status;
    uint8_t        *priv;
};

template <typename P>
View view_of(P *p)
{
    View v = {&p->inPtr, &p->outPtr, &p->inLength, &p->outLength, &p->status, p->lzs_private_};
    return v;
}

bool cuda_ok(cudaError_t e, const char *what)
{
    if (e == cudaSuccess) return true;
    fprintf(stderr, "lzs (b200): %s failed: %s\n", what, cudaGetErrorString(e));
    cudaGetLastError();
    return false;
}

/* Advance n streams by one incremental call each.  Returns bytes produced per stream
 * through `produced`; false on a device failure (every status then carries ERROR). */
bool run_calls(bool decompress, View *views, uint32_t n, bool add_end_marker, size_t *produced)
{
    const size_t state_bytes = decompress ? kIncDBytes : kIncCBytes;
    const size_t state_stride = (state_bytes + 31) / 16 * 16;
    for (uint32_t s = 0; s < n; s++) produced[s] = 0;
    if (n == 0) return true;
    Arena *ap = lzs_b200_device_count() > 0 ? arena() : nullptr;
    if (!ap) {
        fprintf(stderr, "lzs (b200): no CUDA device: the incremental LZS codec has no CPU path\n");
        for (uint32_t s = 0; s < n; s++) *views[s].status = decompress ? LZS_D_STATUS_ERROR : LZS_C_STATUS_ERROR;
        return false;
    }
    std::lock_guard<std::mutex> lock(ap->mu);
    Arena &a = *ap;

    std::vector<lzs::IncJob> jobs(n);
    std::vector<bool>        active(n, true);
    for (uint32_t s = 0; s < n; s++) *views[s].status = 0;
    bool again = true, ok = true;
    while (again && ok) {
        again = false;
        /* lay out this pass: one piece of input and output space per active stream */
        size_t in_total = 0, out_total = 0;
        std::vector<size_t> in_at(n), out_at(n), in_piece(n), out_piece(n);
        uint32_t            n_act = 0;
        for (uint32_t s = 0; s < n; s++) {
            if (!active[s]) continue;
            in_piece[s] = *views[s].in_len < kPieceMax ? *views[s].in_len : kPieceMax;
            out_piece[s] = *views[s].out_len < kPieceMax ? *views[s].out_len : kPieceMax;
            in_at[s] = in_total;   in_total += (in_piece[s] + 15) / 16 * 16;
            out_at[s] = out_total; out_total += (out_piece[s] + 15) / 16 * 16;
            n_act++;
        }
        if (!a.reserve(A_STATE, n * state_stride) || !a.reserve(A_IN, in_total) || !a.reserve(A_OUT, out_total) ||
            !a.reserve(A_JOBS, n * sizeof(lzs::IncJob))) {
            fprintf(stderr, "lzs (b200): device allocation failed\n");
            ok = false;
            break;
        }
        uint8_t *d_state = static_cast<uint8_t *>(a.buf[A_STATE]);
        uint8_t *d_in = static_cast<uint8_t *>(a.buf[A_IN]);
        uint8_t *d_out = static_cast<uint8_t *>(a.buf[A_OUT]);
        uint32_t j = 0;
        std::vector<uint32_t> who(n_act);
        for (uint32_t s = 0; s < n && ok; s++) {
            if (!active[s]) continue;
            const bool last_piece = in_piece[s] == *views[s].in_len;
            ok = ok && cuda_ok(cudaMemcpyAsync(d_state + j * state_stride, views[s].priv, state_bytes,
                                               cudaMemcpyHostToDevice, a.stream), "state upload");
            if (in_piece[s])
                ok = ok && cuda_ok(cudaMemcpyAsync(d_in + in_at[s], *views[s].in_ptr, in_piece[s],
                                                   cudaMemcpyHostToDevice, a.stream), "input upload");
            lzs::IncJob &job = jobs[j];
            job.state = d_state + j * state_stride;
            job.in = d_in + in_at[s];
            job.out = d_out + out_at[s];
            job.in_len = static_cast<uint32_t>(in_piece[s]);
            job.out_cap = static_cast<uint32_t>(out_piece[s]);
            job.in_used = job.out_used = job.status = 0;
            job.add_end_marker = (add_end_marker && last_piece) ? 1u : 0u;
            who[j++] = s;
        }
        if (!ok) break;
        ok = cuda_ok(cudaMemcpyAsync(a.buf[A_JOBS], jobs.data(), n_act * sizeof(lzs::IncJob), cudaMemcpyHostToDevice,
                                     a.stream), "job upload");
        if (!ok) break;
        const unsigned grid = (n_act * 32 + 127) / 128;
        if (decompress) lzs::kinc_decompress<<<grid, 128, 0, a.stream>>>(static_cast<lzs::IncJob *>(a.buf[A_JOBS]), n_act);
        else            lzs::kinc_compress<<<grid, 128, 0, a.stream>>>(static_cast<lzs::IncJob *>(a.buf[A_JOBS]), n_act);
        ok = cuda_ok(cudaGetLastError(), "incremental kernel launch") &&
             cuda_ok(cudaMemcpyAsync(jobs.data(), a.buf[A_JOBS], n_act * sizeof(lzs::IncJob), cudaMemcpyDeviceToHost,
                                     a.stream), "job download") &&
             cuda_ok(cudaStreamSynchronize(a.stream), "incremental kernel");
        if (!ok) break;
        for (j = 0; j < n_act && ok; j++) {
            const uint32_t     s = who[j];
            const lzs::IncJob &job = jobs[j];
            ok = ok && cuda_ok(cudaMemcpyAsync(views[s].priv, d_state + j * state_stride, state_bytes,
                                               cudaMemcpyDeviceToHost, a.stream), "state download");
            if (job.out_used)
                ok = ok && cuda_ok(cudaMemcpyAsync(*views[s].out_ptr, job.out, job.out_used, cudaMemcpyDeviceToHost,
                                                   a.stream), "output download");
            const bool last_piece = in_piece[s] == *views[s].in_len;
            const bool out_limited = out_piece[s] < *views[s].out_len;
            *views[s].in_ptr += job.in_used;
            *views[s].in_len -= job.in_used;
            *views[s].out_ptr += job.out_used;
            *views[s].out_len -= job.out_used;
            *views[s].status = static_cast<uint8_t>(job.status);
            produced[s] += job.out_used;
            /* oversize slices: keep going while the stop was only our piece limit */
            const bool stopped_on_piece_input = !last_piece && job.in_used == in_piece[s] && !(job.status & lzs::kStNoSpace);
            const bool stopped_on_piece_output = out_limited && (job.status & lzs::kStNoSpace);
            active[s] = stopped_on_piece_input || stopped_on_piece_output;
            again = again || active[s];
        }
        ok = ok && cuda_ok(cudaStreamSynchronize(a.stream), "result download");
    }
    if (!ok)
        for (uint32_t s = 0; s < n; s++) *views[s].status |= decompress ? LZS_D_STATUS_ERROR : LZS_C_STATUS_ERROR;
    return ok;
}

template <typename P>
size_t one_call(bool decompress, P *p, bool add_end_marker)
{
    View   v = view_of(p);
    size_t produced = 0;
    run_calls(decompress, &v, 1, add_end_marker, &produced);
    return produced;
}

void reset_compress_state(uint8_t *priv)
{
    lzs::IncCompressState hdr;
    memset(&hdr, 0, offsetof(lzs::IncCompressState, ring));
    memcpy(priv, &hdr, offsetof(lzs::IncCompressState, ring));      /* the ring itself needs no clearing */
}

}  // namespace

extern "C" {

/* reference c/src/liblzs/lzs-compression.c:479-516.  The reference's two initialisers
 * differ only in whether its hash tables are cleared; this implementation keeps no
 * hash tables in the state, so both do the same thing. */
void lzs_compress_init_quick(LzsCompressParameters_t *pParams)
{
    pParams->status = LZS_C_STATUS_NONE;
    reset_compress_state(pParams->lzs_private_);
}
void lzs_compress_init_full(LzsCompressParameters_t *pParams) { lzs_compress_init_quick(pParams); }

size_t lzs_compress_incremental(LzsCompressParameters_t *pParams, bool add_end_marker)
{
    return one_call(false, pParams, add_end_marker);
}

/* reference c/src/liblzs/lzs-compression-simple.c:382, :435 -- same bytes as above */
void lzs_simple_compress_init(LzsSimpleCompressParameters_t *pParams)
{
    pParams->status = LZS_C_STATUS_NONE;
    reset_compress_state(pParams->lzs_private_);
}
size_t lzs_simple_compress_incremental(LzsSimpleCompressParameters_t *pParams, bool add_end_marker)
{
    return one_call(false, pParams, add_end_marker);
}

/* reference c/src/liblzs/lzs-decompression.c:420-428, :459 */
void lzs_decompress_init(LzsDecompressParameters_t *pParams)
{
    lzs::IncDecompressState hdr;
    memset(&hdr, 0, offsetof(lzs::IncDecompressState, ring));
    hdr.state = lzs::kDTokenType;
    memcpy(pParams->lzs_private_, &hdr, offsetof(lzs::IncDecompressState, ring));
    pParams->status = LZS_D_STATUS_NONE;
}
size_t lzs_decompress_incremental(LzsDecompressParameters_t *pParams)
{
    return one_call(true, pParams, false);
}

/* ---- device-resident batches: states, job table, input and output all in device memory ---- */

static_assert(sizeof(lzs_b200_inc_job_t) == sizeof(lzs::IncJob) && offsetof(lzs_b200_inc_job_t, status) == offsetof(lzs::IncJob, status) &&
                  offsetof(lzs_b200_inc_job_t, add_end_marker) == offsetof(lzs::IncJob, add_end_marker),
              "the public job record is the kernels' job record");
static_assert(offsetof(lzs::IncCompressState, ring) == 16 && offsetof(lzs::IncDecompressState, ring) == 16, "16-byte headers");

size_t lzs_b200_incremental_state_bytes(int decompress)
{
    return ((decompress ? kIncDBytes : kIncCBytes) + 15) / 16 * 16;
}

int lzs_b200_incremental_init_device(void *states, size_t stride, uint32_t n_streams, int decompress, void *stream)
{
    if (n_streams == 0) return LZS_B200_OK;
    if (!states || stride < lzs_b200_incremental_state_bytes(decompress) || (stride & 15u) ||
        (reinterpret_cast<uintptr_t>(states) & 15u))
        return LZS_B200_EINVAL;
    lzs::kinc_init<<<(n_streams + 255u) / 256u, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<uint8_t *>(states), stride, n_streams, decompress);
    return cudaGetLastError() == cudaSuccess ? LZS_B200_OK : LZS_B200_ECUDA;
}

int lzs_b200_compress_incremental_batch_device(lzs_b200_inc_job_t *jobs, uint32_t n_streams, void *stream)
{
    if (n_streams == 0) return LZS_B200_OK;
    if (!jobs) return LZS_B200_EINVAL;
    const unsigned grid = (n_streams + 3u) / 4u;
    lzs::kinc_compress<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<lzs::IncJob *>(jobs), n_streams);
    return cudaGetLastError() == cudaSuccess ? LZS_B200_OK : LZS_B200_ECUDA;
}

int lzs_b200_decompress_incremental_batch_device(lzs_b200_inc_job_t *jobs, uint32_t n_streams, void *stream)
{
    if (n_streams == 0) return LZS_B200_OK;
    if (!jobs) return LZS_B200_EINVAL;
    const unsigned grid = (n_streams + 3u) / 4u;
    lzs::kinc_decompress<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<lzs::IncJob *>(jobs), n_streams);
    return cudaGetLastError() == cudaSuccess ? LZS_B200_OK : LZS_B200_ECUDA;
}

/* Batch forms: advance n independent streams by one incremental call each, one warp
 * per stream in a single launch.  produced[s] (optional) receives each call's return
 * value. */
int lzs_b200_compress_incremental_batch(LzsCompressParameters_t **params, uint32_t n, int add_end_marker,
                                        size_t *produced)
{
    std::vector<View>   views(n);
    std::vector<size_t> out(n ? n : 1);
    for (uint32_t s = 0; s < n; s++) views[s] = view_of(params[s]);
    const bool ok = run_calls(false, views.data(), n, add_end_marker != 0, out.data());
    if (produced) memcpy(produced, out.data(), n * sizeof(size_t));
    return ok ? LZS_B200_OK : LZS_B200_ECUDA;
}

int lzs_b200_decompress_incremental_batch(LzsDecompressParameters_t **params, uint32_t n, size_t *produced)
{
    std::vector<View>   views(n);
    std::vector<size_t> out(n ? n : 1);
    for (uint32_t s = 0; s < n; s++) views[s] = view_of(params[s]);
    const bool ok = run_calls(true, views.data(), n, false, out.data());
    if (produced) memcpy(produced, out.data(), n * sizeof(size_t));
    return ok ? LZS_B200_OK : LZS_B200_ECUDA;
}

}  // extern "C"
