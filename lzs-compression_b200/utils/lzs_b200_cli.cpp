/*
 * lzs-b200 -- file compressor / decompressor over the batch ABI.
 *
 *   lzs-b200 c [-b chunk_bytes | -s] [-x index_file] infile outfile
 *   lzs-b200 d [-x index_file] infile outfile
 *
 * The reference's file format is a raw LZS stream with no header (c/src/utils/lzs-compress.c:82-134
 * writes what lzs_compress_incremental produces, lzs-decompress.c:75-118 feeds the file to
 * lzs_decompress_incremental).  Its decoder keeps going after an end marker
 * (lzs-decompression.c:564-576), so a file made of back-to-back INDEPENDENT streams, one per chunk,
 * is a valid input for the reference's lzs-decompress -- and that is what `c` writes, because
 * independent chunks are what both the compressor and the decoder of the GPU take in parallel.
 * With `-s` (or a chunk size of at least the file size) the output is the single stream the
 * reference's lzs-compress writes, byte for byte; the library compresses it in parallel all the
 * same (cut into pieces inside, csrc/k23_pieces.cuh), and `d` decodes long streams of an index-less file in
 * parallel inside each stream (csrc/k4_pieces.cuh), walking from end marker to end marker.
 *
 * Finding the stream starts again needs a scan of the whole bit stream, so `c -x` also writes a
 * small index (one (uncompressed, compressed) length pair per chunk) and `d -x` uses it to decode
 * all chunks in one batch.  Without an index `d` decodes the file as the reference does, as one
 * resumable stream through lzs_decompress_incremental (on the GPU as well, one call per buffer:
 * correct for any LZS file, but serial).
 *
 * No CPU codec in here: every byte goes through liblzs.so (B200).
 */
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "lzs.h"
#include "lzs_b200.h"

namespace {

const char kIndexMagic[8] = {'L', 'Z', 'S', 'X', '1', 0, 0, 0};

bool read_file(const char *path, std::vector<uint8_t> &data)
{
    FILE *f = fopen(path, "rb");
    if (!f) { perror(path); return false; }
    uint8_t buf[1 << 16];
    size_t  got;
    while ((got = fread(buf, 1, sizeof buf, f)) > 0) data.insert(data.end(), buf, buf + got);
    const bool ok = !ferror(f);
    fclose(f);
    if (!ok) perror(path);
    return ok;
}

bool write_file(const char *path, const uint8_t *data, size_t n)
{
    FILE *f = fopen(path, "wb");
    if (!f) { perror(path); return false; }
    const bool ok = fwrite(data, 1, n, f) == n;
    if (fclose(f) != 0 || !ok) { perror(path); return false; }
    return true;
}

int usage()
{
    fprintf(stderr, "usage: lzs-b200 c [-b chunk_bytes | -s] [-x index_file] infile outfile\n"
                    "       lzs-b200 d [-x index_file] infile outfile\n");
    return 2;
}

int compress_file(const char *in_path, const char *out_path, const char *index_path, uint64_t chunk)
{
    std::vector<uint8_t> in;
    if (!read_file(in_path, in)) return 1;
    /* an empty file is one empty stream (the reference writes the end marker alone) */
    const uint64_t total = in.size();
    const uint32_t n = static_cast<uint32_t>(total == 0 ? 1 : (total + chunk - 1) / chunk);
    std::vector<uint64_t> in_off(n), out_off(n);
    std::vector<uint32_t> in_len(n), out_len(n);
    uint64_t              cap = 0;
    for (uint32_t s = 0; s < n; s++) {
        in_off[s] = static_cast<uint64_t>(s) * chunk;
        in_len[s] = static_cast<uint32_t>(total - in_off[s] < chunk ? total - in_off[s] : chunk);
        cap += (LZS_COMPRESSED_MAX(static_cast<size_t>(in_len[s])) + 15u) & ~static_cast<size_t>(15);
    }
    in.resize(in.size() + 16);                      /* slack for the library's aligned loads */
    std::vector<uint8_t> packed(cap + 64);
    uint64_t             used = 0;
    int rc = lzs_b200_compress_packed_host(in.data(), in_off.data(), in_len.data(), total, packed.data(), cap,
                                           out_off.data(), out_len.data(), n, &used);
    if (rc != LZS_B200_OK) {
        fprintf(stderr, "lzs-b200: %s\n", lzs_b200_last_error());
        return 1;
    }
    /* the batch call aligns every stream to 16 bytes; the file has them back to back */
    std::vector<uint8_t> file;
    file.reserve(used);
    for (uint32_t s = 0; s < n; s++) file.insert(file.end(), packed.begin() + out_off[s], packed.begin() + out_off[s] + out_len[s]);
    if (!write_file(out_path, file.data(), file.size())) return 1;
    if (index_path) {
        std::vector<uint8_t> idx(sizeof kIndexMagic + 8 + static_cast<size_t>(n) * 8);
        memcpy(idx.data(), kIndexMagic, sizeof kIndexMagic);
        const uint64_t count = n;
        memcpy(idx.data() + 8, &count, 8);
        for (uint32_t s = 0; s < n; s++) {
            memcpy(idx.data() + 16 + static_cast<size_t>(s) * 8, &in_len[s], 4);
            memcpy(idx.data() + 20 + static_cast<size_t>(s) * 8, &out_len[s], 4);
        }
        if (!write_file(index_path, idx.data(), idx.size())) return 1;
    }
    return 0;
}

int decompress_with_index(const std::vector<uint8_t> &in, const std::vector<uint8_t> &idx, const char *out_path)
{
    if (idx.size() < 16 || memcmp(idx.data(), kIndexMagic, sizeof kIndexMagic) != 0) {
        fprintf(stderr, "lzs-b200: not an index file\n");
        return 1;
    }
    uint64_t count = 0;
    memcpy(&count, idx.data() + 8, 8);
    if (count > 0xFFFFFFFFull || idx.size() != 16 + count * 8) {
        fprintf(stderr, "lzs-b200: damaged index file\n");
        return 1;
    }
    const uint32_t        n = static_cast<uint32_t>(count);
    std::vector<uint64_t> in_off(n), out_off(n);
    std::vector<uint32_t> in_len(n), out_cap(n), out_len(n);
    uint64_t              ci = 0, co = 0;
    for (uint32_t s = 0; s < n; s++) {
        uint32_t u, c;
        memcpy(&u, idx.data() + 16 + static_cast<size_t>(s) * 8, 4);
        memcpy(&c, idx.data() + 20 + static_cast<size_t>(s) * 8, 4);
        in_off[s] = ci; in_len[s] = c; ci += c;
        out_off[s] = co; out_cap[s] = u; co += (static_cast<uint64_t>(u) + 15u) & ~15ull;   /* aligned slots */
    }
    if (ci != in.size()) {
        fprintf(stderr, "lzs-b200: index describes %llu compressed bytes, file has %zu\n",
                static_cast<unsigned long long>(ci), in.size());
        return 1;
    }
    std::vector<uint8_t> src(in);
    src.resize(src.size() + 16);
    std::vector<uint8_t> slots(co + 64);
    int rc = lzs_b200_decompress_batch_host(src.data(), in_off.data(), in_len.data(), ci, slots.data(), out_off.data(),
                                            out_cap.data(), out_len.data(), co, n);
    if (rc != LZS_B200_OK) {
        fprintf(stderr, "lzs-b200: %s\n", lzs_b200_last_error());
        return 1;
    }
    std::vector<uint8_t> out;
    for (uint32_t s = 0; s < n; s++) {
        if (out_len[s] != out_cap[s])
            fprintf(stderr, "lzs-b200: chunk %u decoded to %u bytes, index says %u\n", s, out_len[s], out_cap[s]);
        out.insert(out.end(), slots.begin() + out_off[s], slots.begin() + out_off[s] + out_len[s]);
    }
    return write_file(out_path, out.data(), out.size()) ? 0 : 1;
}

/* the reference CLI's loop (c/src/utils/lzs-decompress.c:75-118) over the same incremental call */
/* A file without an index whose streams are LONG (the reference's own lzs-compress writes ONE): every
 * stream is decoded by one batch-class call (token starts found in parallel, copies resolved by pointer
 * doubling: csrc/k4_pieces.cuh), which also says where the stream's end marker is, so the next stream
 * can be found.  Returns the bytes of `in` that were decoded this way (0: none); whatever is left --
 * short streams, a damaged tail -- goes through the incremental calls like the reference's CLI. */
size_t decompress_long_streams(const std::vector<uint8_t> &in, std::vector<uint8_t> &out)
{
    const size_t kWorth = 1u << 20;                  /* below this a stream is not worth a call of its own */
    size_t       pos = 0;
    while (in.size() - pos >= kWorth) {
        const uint64_t in_off = 0, out_off = 0;
        const uint64_t rest = in.size() - pos;
        const uint32_t in_len = static_cast<uint32_t>(rest < 0x1FFFFF00ull ? rest : 0x1FFFFF00ull);
        /* one stream cannot make more than 30 bytes per byte; 1 GiB per call at most */
        const uint64_t want = 30ull * in_len + 64u;
        const uint32_t cap = static_cast<uint32_t>(want < (1ull << 30) ? want : (1ull << 30));
        std::vector<uint8_t> buf(static_cast<size_t>(cap) + 64u);
        std::vector<uint8_t> src(in.begin() + pos, in.begin() + pos + in_len);
        src.resize(src.size() + 16);
        uint32_t out_len = 0, used = 0xFFFFFFFFu;
        if (lzs_b200_decompress_used_batch_host(src.data(), &in_off, &in_len, in_len, buf.data(), &out_off, &cap, &out_len,
                                                &used, cap, 1) != LZS_B200_OK)
            break;
        if (used == 0xFFFFFFFFu || used < kWorth) break;   /* not clean, did not fit, or short streams: the slow way from here */
        out.insert(out.end(), buf.begin(), buf.begin() + out_len);
        pos += used;
    }
    return pos;
}

int decompress_stream(const std::vector<uint8_t> &in, const char *out_path)
{
    LzsDecompressParameters_t p;
    lzs_decompress_init(&p);
    std::vector<uint8_t> out;
    const size_t         done = decompress_long_streams(in, out);
    std::vector<uint8_t> buf(1u << 20);
    p.inPtr = in.data() + done;
    p.inLength = in.size() - done;
    for (;;) {
        if (p.inLength == 0 && (p.status & LZS_D_STATUS_INPUT_STARVED) != 0) break;    /* as the reference CLI */
        p.outPtr = buf.data();
        p.outLength = buf.size();
        const size_t got = lzs_decompress_incremental(&p);
        out.insert(out.end(), buf.begin(), buf.begin() + got);
        if (p.status & LZS_D_STATUS_ERROR) {
            fprintf(stderr, "lzs-b200: decoder error (no CUDA device?)\n");
            return 1;
        }
    }
    return write_file(out_path, out.data(), out.size()) ? 0 : 1;
}

}  // namespace

int main(int argc, char **argv)
{
    if (argc < 4) return usage();
    const bool  compress = strcmp(argv[1], "c") == 0;
    if (!compress && strcmp(argv[1], "d") != 0) return usage();
    uint64_t    chunk = 65536;
    const char *index_path = nullptr;
    int         a = 2;
    while (a < argc && argv[a][0] == '-' && argv[a][1] != 0) {
        if (strcmp(argv[a], "-b") == 0 && a + 1 < argc) {
            chunk = strtoull(argv[a + 1], nullptr, 0);
            a += 2;
        } else if (strcmp(argv[a], "-s") == 0) {        /* one stream: the reference's own file */
            chunk = 0xFFFFFF00ull;
            a += 1;
        } else if (strcmp(argv[a], "-x") == 0 && a + 1 < argc) {
            index_path = argv[a + 1];
            a += 2;
        } else {
            return usage();
        }
    }
    if (argc - a != 2 || chunk == 0 || chunk > 0xFFFFFF00ull) return usage();
    if (compress) return compress_file(argv[a], argv[a + 1], index_path, chunk);
    std::vector<uint8_t> in;
    if (!read_file(argv[a], in)) return 1;
    if (index_path) {
        std::vector<uint8_t> idx;
        if (!read_file(index_path, idx)) return 1;
        return decompress_with_index(in, idx, argv[a + 1]);
    }
    return decompress_stream(in, argv[a + 1]);
}
