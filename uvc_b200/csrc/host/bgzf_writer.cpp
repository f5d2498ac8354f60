// bgzf_writer.cpp - see bgzf_writer.h.
#include "bgzf_writer.h"

#include <zlib.h>

#include <atomic>
#include <thread>
#include <vector>

#include <stdint.h>
#include <string.h>

namespace {

const size_t kBlock = 0xff00;

// one BGZF member: gzip header with the BC extra field, raw deflate stream, CRC32 and input size
int compress_block(std::string & dst, const char *src, size_t n, int level) {
    uint8_t buf[0x10000 + 64];
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) { return -1; }
    zs.next_in = (Bytef*)src; zs.avail_in = (uInt)n;
    zs.next_out = buf + 18; zs.avail_out = 0x10000 - 18 - 8;
    const int rc = deflate(&zs, Z_FINISH);
    size_t clen = zs.total_out;
    deflateEnd(&zs);
    if (rc != Z_STREAM_END) {
        // incompressible input: store it (a stored deflate block always fits: 0xff00 + 5 bytes)
        memset(&zs, 0, sizeof(zs));
        if (deflateInit2(&zs, 0, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) { return -1; }
        zs.next_in = (Bytef*)src; zs.avail_in = (uInt)n;
        zs.next_out = buf + 18; zs.avail_out = 0x10000 - 18 - 8;
        const int rc2 = deflate(&zs, Z_FINISH);
        clen = zs.total_out;
        deflateEnd(&zs);
        if (rc2 != Z_STREAM_END) { return -1; }
    }
    const uint8_t head[16] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 'B', 'C', 2, 0};
    memcpy(buf, head, 16);
    const size_t bsize = 18 + clen + 8;
    buf[16] = (uint8_t)((bsize - 1) & 0xff); buf[17] = (uint8_t)((bsize - 1) >> 8);
    const uint32_t crc = (uint32_t)crc32(crc32(0L, NULL, 0), (const Bytef*)src, (uInt)n);
    uint8_t *t = buf + 18 + clen;
    t[0] = crc & 0xff; t[1] = (crc >> 8) & 0xff; t[2] = (crc >> 16) & 0xff; t[3] = (crc >> 24) & 0xff;
    t[4] = n & 0xff; t[5] = (n >> 8) & 0xff; t[6] = (n >> 16) & 0xff; t[7] = (n >> 24) & 0xff;
    dst.assign((const char*)buf, bsize);
    return 0;
}

} // namespace

int uvchost_bgzf_compress(std::string & out, const char *text, size_t len, int level, int n_threads) {
    if (0 == len) { return 0; }
    const size_t n_blocks = (len + kBlock - 1) / kBlock;
    std::vector<std::string> parts(n_blocks);
    std::atomic<size_t> next(0);
    std::atomic<int> failed(0);
    auto work = [&]() {
        for (;;) {
            const size_t i = next.fetch_add(1);
            if (i >= n_blocks) { break; }
            const size_t off = i * kBlock;
            const size_t n = (len - off < kBlock ? len - off : kBlock);
            if (compress_block(parts[i], text + off, n, level) != 0) { failed.store(1); }
        }
    };
    if (n_threads <= 1 || n_blocks < 4) { work(); }
    else {
        std::vector<std::thread> pool;
        const size_t nt = ((size_t)n_threads < n_blocks ? (size_t)n_threads : n_blocks);
        for (size_t k = 0; k < nt; k++) { pool.emplace_back(work); }
        for (auto & th : pool) { th.join(); }
    }
    if (failed.load()) { return -1; }
    size_t total = 0;
    for (const auto & p : parts) { total += p.size(); }
    out.reserve(out.size() + total);
    for (const auto & p : parts) { out += p; }
    return 0;
}

void uvchost_bgzf_eof(std::string & out) {
    static const uint8_t eof[28] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43, 0x02, 0, 0x1b, 0, 0x03, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    out.append((const char*)eof, sizeof(eof));
}
