// bam_reader.cpp - BGZF/BAM/BAI and FASTA/FAI decoding into SoA buffers (see bam_reader.h).
// Formats follow the SAM/BAM specification (sections 4.1 BGZF, 4.2 BAM, 5.2 BAI) and the samtools faidx five-column index.
#include "bam_reader.h"
#include "inflate_fast.h"

#include <zlib.h>

#include <algorithm>
#include <map>
#include <string>
#include <vector>

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>

namespace {

const int kMaxBlock = 0x10000;

// Sequential read-ahead for whole-file scans (the region tiler): compressed BGZF members are read in file order in groups of kGroup and
// inflated on a small pool of threads (members are independent deflate streams); the consumer then takes them in order.
struct SeqInflater {
    static const int kGroup = 96;
    struct Slot { std::vector<uint8_t> c, u; int clen = 0, ulen = 0, bsize = 0; int64_t addr = 0; bool ok = true; };
    FILE *fp = NULL;
    int64_t next_addr = 0;
    std::vector<Slot> slots;
    int n_filled = 0, n_taken = 0;
    bool eof = false, failed = false;
    std::vector<std::thread> pool;
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    int generation = 0, pending = 0;
    std::atomic<int> next_job{0};
    bool stop = false;

    void start(const char *path, int64_t addr, int n_threads) {
        fp = fopen(path, "rb");
        next_addr = addr;
        slots.resize(kGroup);
        for (auto & s : slots) { s.c.resize(kMaxBlock + 64); s.u.resize(kMaxBlock); }
        for (int t = 0; t < n_threads; t++) { pool.emplace_back([this]() { worker(); }); }
    }
    static void inflate_slot(Slot & s) {
        const int64_t n = uvc_inflate_member(s.c.data(), (size_t)s.clen, s.u.data(), (size_t)kMaxBlock);    // (s.c holds the footer and 56 bytes of slack after the stream)
        s.ok = (n >= 0);
        s.ulen = (int)(n >= 0 ? n : 0);
    }
    void worker() {
        int seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_work.wait(lk, [&]() { return stop || generation != seen; });
                if (stop) { return; }
                seen = generation;
            }
            for (;;) {
                const int j = next_job.fetch_add(1);
                if (j >= n_filled) { break; }
                inflate_slot(slots[j]);
                std::lock_guard<std::mutex> lk(mu);
                if (--pending == 0) { cv_done.notify_all(); }
            }
        }
    }
    // reads and inflates the next group; returns false at end of file / on error
    bool fill() {
        n_filled = 0; n_taken = 0;
        if (eof || failed || NULL == fp) { return false; }
        if (fseeko(fp, next_addr, SEEK_SET) != 0) { failed = true; return false; }
        while (n_filled < kGroup) {
            Slot & s = slots[n_filled];
            uint8_t h[18];
            const size_t n = fread(h, 1, 18, fp);
            if (0 == n) { eof = true; break; }
            if (n != 18 || h[0] != 31 || h[1] != 139 || !(h[3] & 4)) { failed = true; break; }
            const int xlen = h[10] | (h[11] << 8);
            uint8_t extra[512];
            memcpy(extra, h + 12, 6);
            if (xlen > 512 || (xlen > 6 && fread(extra + 6, 1, xlen - 6, fp) != (size_t)(xlen - 6))) { failed = true; break; }
            int bsize = -1;
            for (int off = 0; off + 4 <= xlen;) {
                const int slen = extra[off + 2] | (extra[off + 3] << 8);
                if (extra[off] == 'B' && extra[off + 1] == 'C' && slen == 2) { bsize = (extra[off + 4] | (extra[off + 5] << 8)) + 1; }
                off += 4 + slen;
            }
            if (bsize < 0) { failed = true; break; }
            s.clen = bsize - 12 - xlen - 8;
            if (s.clen < 0) { failed = true; break; }
            if (fread(s.c.data(), 1, s.clen + 8, fp) != (size_t)(s.clen + 8)) { failed = true; break; }
            s.addr = next_addr; s.bsize = bsize;
            next_addr += bsize;
            n_filled++;
        }
        if (0 == n_filled) { return false; }
        if (pool.empty()) { for (int j = 0; j < n_filled; j++) { inflate_slot(slots[j]); } }
        else {
            {
                std::lock_guard<std::mutex> lk(mu);
                pending = n_filled; next_job.store(0); generation++;
            }
            cv_work.notify_all();
            std::unique_lock<std::mutex> lk(mu);
            cv_done.wait(lk, [&]() { return 0 == pending; });
        }
        for (int j = 0; j < n_filled; j++) { if (!slots[j].ok) { failed = true; return false; } }
        return true;
    }
    // next block in file order; NULL at end of file or on error
    Slot *take() {
        if (n_taken >= n_filled && !fill()) { return NULL; }
        return &slots[n_taken++];
    }
    ~SeqInflater() {
        { std::lock_guard<std::mutex> lk(mu); stop = true; }
        cv_work.notify_all();
        for (auto & th : pool) { th.join(); }
        if (fp) { fclose(fp); }
    }
};

struct BgzfIn {
    FILE *fp = NULL;
    std::vector<uint8_t> cbuf, ubuf;
    int64_t block_addr = 0;   // compressed offset of the block in ubuf
    int block_clen = 0;       // 0 = nothing loaded
    int ulen = 0, uoff = 0;
    z_stream zs;
    bool zs_init = false;
    SeqInflater *seq = NULL;    // when set, blocks come from the parallel sequential read-ahead

    // 0 ok, 1 end of file, -1 error
    int load(int64_t caddr) {
        if (seq) {
            SeqInflater::Slot *s = seq->take();
            if (NULL == s) { if (seq->failed) { return -1; } block_addr = caddr; block_clen = 0; ulen = uoff = 0; return 1; }
            ubuf.swap(s->u);
            if ((int)s->u.size() < kMaxBlock) { s->u.resize(kMaxBlock); }
            ulen = s->ulen; uoff = 0; block_addr = s->addr; block_clen = s->bsize;
            return 0;
        }
        uint8_t h[18];
        if (fseeko(fp, caddr, SEEK_SET) != 0) { return -1; }
        const size_t n = fread(h, 1, 18, fp);
        if (0 == n) { block_addr = caddr; block_clen = 0; ulen = uoff = 0; return 1; }
        if (n != 18 || h[0] != 31 || h[1] != 139 || !(h[3] & 4)) { return -1; }
        const int xlen = h[10] | (h[11] << 8);
        int bsize = -1;
        cbuf.resize(kMaxBlock + 64);
        memcpy(cbuf.data(), h + 12, 6);
        if (xlen > 6 && fread(cbuf.data() + 6, 1, xlen - 6, fp) != (size_t)(xlen - 6)) { return -1; }
        for (int off = 0; off + 4 <= xlen;) {
            const int slen = cbuf[off + 2] | (cbuf[off + 3] << 8);
            if (cbuf[off] == 'B' && cbuf[off + 1] == 'C' && slen == 2) { bsize = (cbuf[off + 4] | (cbuf[off + 5] << 8)) + 1; }
            off += 4 + slen;
        }
        if (bsize < 0) { return -1; }
        const int clen = bsize - 12 - xlen - 8;
        if (clen < 0) { return -1; }
        if (fread(cbuf.data(), 1, clen + 8, fp) != (size_t)(clen + 8)) { return -1; }
        ubuf.resize(kMaxBlock);
        const int64_t n_out = uvc_inflate_member(cbuf.data(), (size_t)clen, ubuf.data(), (size_t)kMaxBlock);     // (cbuf holds the footer and 56 bytes of slack after the stream)
        if (n_out < 0) { return -1; }
        ulen = (int)n_out; uoff = 0; block_addr = caddr; block_clen = bsize;
        return 0;
    }
    int seek(int64_t voff) {
        stop_seq();
        const int64_t caddr = voff >> 16;
        if (!(block_clen > 0 && block_addr == caddr)) { if (load(caddr) < 0) { return -1; } }
        uoff = (int)(voff & 0xffff);
        return 0;
    }
    // returns bytes read (less than n only at end of file), -1 on error
    int64_t read(void *dst, int64_t n) {
        uint8_t *out = (uint8_t*)dst;
        int64_t done = 0;
        while (done < n) {
            if (uoff >= ulen) {
                const int r = load(block_addr + block_clen);
                if (r < 0) { return -1; }
                if (r == 1) { break; }
                if (0 == ulen) { continue; }
            }
            int64_t k = ulen - uoff;
            if (k > n - done) { k = n - done; }
            memcpy(out + done, ubuf.data() + uoff, (size_t)k);
            uoff += (int)k; done += k;
        }
        return done;
    }
    void stop_seq() { if (seq) { delete seq; seq = NULL; block_clen = 0; ulen = uoff = 0; } }
    ~BgzfIn() { stop_seq(); if (zs_init) { inflateEnd(&zs); } if (fp) { fclose(fp); } }
};

inline uint32_t le32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

} // namespace

struct uvchost_bam {
    int index_status = 0;                      // 1: .bai loaded, 0: no index file, -1: index file present but invalid (bad magic, truncated, wrong number of references)
    BgzfIn in;
    std::string path;
    uvchost_readbuf *span = NULL;              // reusable span buffer of uvchost_bam_fetch_tiles
    std::vector<std::string> names;
    std::vector<int64_t> lens;
    std::vector<std::vector<uint64_t>> lidx;   // linear index per reference
    int64_t first_record_voff = 0;
    std::vector<uint8_t> rec;
};

struct uvchost_readbuf {
    std::vector<int32_t> endpos_scratch;   // only filled for the span buffer of uvchost_bam_fetch_tiles
    std::vector<int32_t> pos, mpos, isize, mtid, l_qseq, n_cigar, nm;
    std::vector<uint16_t> flag;
    std::vector<uint8_t> mapq;
    std::vector<uint64_t> seq_off, qual_off, cigar_off, qname_off;
    std::vector<uint8_t> seq, qual;
    std::vector<uint32_t> cigar;
    std::vector<char> qname;
};

struct uvchost_fasta {
    FILE *fp = NULL;
    struct Entry { int64_t len, offset; int32_t linebases, linewidth; };
    std::map<std::string, Entry> entries;
};

extern "C" {

uvchost_bam *uvchost_bam_open(const char *path) {
    uvchost_bam *b = new uvchost_bam();
    b->path = path;
    b->in.fp = fopen(path, "rb");
    if (NULL == b->in.fp) { delete b; return NULL; }
    uint8_t w[8];
    if (b->in.seek(0) != 0 || b->in.read(w, 8) != 8 || memcmp(w, "BAM\1", 4) != 0) { delete b; return NULL; }
    const int32_t l_text = (int32_t)le32(w + 4);
    std::vector<char> text((size_t)l_text);
    if (b->in.read(text.data(), l_text) != l_text || b->in.read(w, 4) != 4) { delete b; return NULL; }
    const int32_t n_ref = (int32_t)le32(w);
    for (int32_t i = 0; i < n_ref; i++) {
        if (b->in.read(w, 4) != 4) { delete b; return NULL; }
        const int32_t l_name = (int32_t)le32(w);
        std::vector<char> nm((size_t)l_name);
        if (b->in.read(nm.data(), l_name) != l_name || b->in.read(w, 4) != 4) { delete b; return NULL; }
        b->names.push_back(std::string(nm.data()));
        b->lens.push_back((int32_t)le32(w));
    }
    b->first_record_voff = (b->in.uoff >= b->in.ulen ? ((b->in.block_addr + b->in.block_clen) << 16) : ((b->in.block_addr << 16) | b->in.uoff));
    // index: <path>.bai or <path minus .bam>.bai
    std::string idx = std::string(path) + ".bai";
    FILE *f = fopen(idx.c_str(), "rb");
    if (NULL == f) {
        std::string alt(path);
        if (alt.size() > 4 && alt.substr(alt.size() - 4) == ".bam") { alt = alt.substr(0, alt.size() - 4) + ".bai"; f = fopen(alt.c_str(), "rb"); }
    }
    b->index_status = 0;
    if (f) {
        fseek(f, 0, SEEK_END);
        const long sz = ftell(f);
        fseek(f, 0, SEEK_SET);
        std::vector<uint8_t> d((size_t)(sz > 0 ? sz : 0));
        // every offset is checked against the file size: a truncated or foreign file must not pass as "no reads anywhere"
        bool ok = (sz >= 8 && fread(d.data(), 1, (size_t)sz, f) == (size_t)sz && memcmp(d.data(), "BAI\1", 4) == 0);
        size_t o = 4;
        uint32_t nr = 0;
        if (ok) { nr = le32(&d[o]); o += 4; ok = (nr == (uint32_t)n_ref); }
        if (ok) { b->lidx.resize(nr); }
        for (uint32_t r = 0; ok && r < nr; r++) {
            if (o + 4 > d.size()) { ok = false; break; }
            const uint32_t n_bin = le32(&d[o]); o += 4;
            for (uint32_t bi = 0; ok && bi < n_bin; bi++) {
                if (o + 8 > d.size()) { ok = false; break; }
                const uint32_t n_chunk = le32(&d[o + 4]);
                if ((d.size() - (o + 8)) / 16 < (size_t)n_chunk) { ok = false; break; }
                o += 8 + (size_t)n_chunk * 16;
            }
            if (!ok || o + 4 > d.size()) { ok = false; break; }
            const uint32_t n_intv = le32(&d[o]); o += 4;
            if ((d.size() - o) / 8 < (size_t)n_intv) { ok = false; break; }
            b->lidx[r].resize(n_intv);
            for (uint32_t i = 0; i < n_intv; i++) { b->lidx[r][i] = (uint64_t)le32(&d[o]) | ((uint64_t)le32(&d[o + 4]) << 32); o += 8; }
        }
        if (!ok) { b->lidx.clear(); }
        b->index_status = (ok ? 1 : -1);
        fclose(f);
    }
    return b;
}

int uvchost_bam_index_status(const uvchost_bam *b) { return (b ? b->index_status : 0); }
void uvchost_bam_close(uvchost_bam *b) { if (b) { delete b->span; } delete b; }
int32_t uvchost_bam_n_targets(const uvchost_bam *b) { return (int32_t)b->names.size(); }
const char *uvchost_bam_target_name(const uvchost_bam *b, int32_t tid) { return b->names[tid].c_str(); }
int64_t uvchost_bam_target_len(const uvchost_bam *b, int32_t tid) { return b->lens[tid]; }

uvchost_readbuf *uvchost_readbuf_new(void) {
    uvchost_readbuf *rb = new uvchost_readbuf();
    uvchost_readbuf_clear(rb);
    return rb;
}
void uvchost_readbuf_free(uvchost_readbuf *rb) { delete rb; }
void uvchost_readbuf_clear(uvchost_readbuf *rb) {
    rb->pos.clear(); rb->mpos.clear(); rb->isize.clear(); rb->mtid.clear(); rb->l_qseq.clear(); rb->n_cigar.clear(); rb->nm.clear();
    rb->flag.clear(); rb->mapq.clear();
    rb->seq_off.assign(1, 0); rb->qual_off.assign(1, 0); rb->cigar_off.assign(1, 0); rb->qname_off.assign(1, 0);
    rb->seq.clear(); rb->qual.clear(); rb->cigar.clear(); rb->qname.clear();
}
int64_t uvchost_readbuf_size(const uvchost_readbuf *rb) { return (int64_t)rb->pos.size(); }
void uvchost_readbuf_view(const uvchost_readbuf *rb, uvcgpu_reads_soa *out) {
    out->n_reads = (int64_t)rb->pos.size();
    out->pos = rb->pos.data(); out->mpos = rb->mpos.data(); out->isize = rb->isize.data(); out->mtid = rb->mtid.data();
    out->l_qseq = rb->l_qseq.data(); out->n_cigar = rb->n_cigar.data(); out->nm = rb->nm.data();
    out->flag = rb->flag.data(); out->mapq = rb->mapq.data();
    out->seq_off = rb->seq_off.data(); out->qual_off = rb->qual_off.data(); out->cigar_off = rb->cigar_off.data(); out->qname_off = rb->qname_off.data();
    out->seq = rb->seq.data(); out->qual = rb->qual.data(); out->cigar = rb->cigar.data(); out->qname = rb->qname.data();
}

} // extern "C"

namespace {

struct Core { int32_t tid, pos, l_qname, mapq, n_cigar, flag, l_qseq, mtid, mpos, isize, endpos; };

// reads the next record into b->rec; returns 1 ok, 0 end of file, -1 error
int next_record(uvchost_bam *b, Core & c) {
    uint8_t w[4];
    const int64_t n = b->in.read(w, 4);
    if (0 == n) { return 0; }
    if (n != 4) { return -1; }
    const int32_t bs = (int32_t)le32(w);
    if (bs < 32) { return -1; }
    b->rec.resize((size_t)bs);
    if (b->in.read(b->rec.data(), bs) != bs) { return -1; }
    const uint8_t *x = b->rec.data();
    {   // the variable-length parts must fit the record (a corrupt file must not be read out of bounds)
        const int64_t l_qname = x[8], n_cigar = (int64_t)(x[12] | (x[13] << 8)), l_qseq = (int64_t)(int32_t)le32(x + 16);
        if (l_qseq < 0 || 32 + l_qname + 4 * n_cigar + (l_qseq + 1) / 2 + l_qseq > (int64_t)bs || l_qname < 1) { return -1; }
    }
    c.tid = (int32_t)le32(x); c.pos = (int32_t)le32(x + 4);
    c.l_qname = x[8]; c.mapq = x[9];
    c.n_cigar = x[12] | (x[13] << 8); c.flag = x[14] | (x[15] << 8);
    c.l_qseq = (int32_t)le32(x + 16); c.mtid = (int32_t)le32(x + 20); c.mpos = (int32_t)le32(x + 24); c.isize = (int32_t)le32(x + 28);
    int64_t rlen = 0;
    if (!(c.flag & 4)) {
        const uint8_t *cg = x + 32 + c.l_qname;
        for (int k = 0; k < c.n_cigar; k++) {
            const uint32_t v = le32(cg + 4 * k);
            const int op = v & 0xf;
            if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) { rlen += v >> 4; }
        }
    }
    if (0 == rlen) { rlen = 1; }
    c.endpos = (int32_t)(c.pos + rlen);
    return 1;
}

void append_record(uvchost_readbuf *rb, const uvchost_bam *b, const Core & c) {
    const uint8_t *x = b->rec.data();
    const size_t total = b->rec.size();
    rb->pos.push_back(c.pos); rb->mpos.push_back(c.mpos); rb->isize.push_back(c.isize); rb->mtid.push_back(c.mtid);
    rb->l_qseq.push_back(c.l_qseq); rb->n_cigar.push_back(c.n_cigar); rb->flag.push_back((uint16_t)c.flag); rb->mapq.push_back((uint8_t)c.mapq);
    const uint8_t *qn = x + 32;
    const uint8_t *cg = qn + c.l_qname;
    const uint8_t *sq = cg + 4 * (size_t)c.n_cigar;
    const uint8_t *ql = sq + (size_t)((c.l_qseq + 1) / 2);
    const uint8_t *aux = ql + c.l_qseq;
    rb->qname.insert(rb->qname.end(), (const char*)qn, (const char*)qn + c.l_qname);
    for (int k = 0; k < c.n_cigar; k++) { rb->cigar.push_back(le32(cg + 4 * k)); }
    rb->seq.insert(rb->seq.end(), sq, ql);
    rb->qual.insert(rb->qual.end(), ql, aux);
    rb->qname_off.push_back(rb->qname.size()); rb->cigar_off.push_back(rb->cigar.size());
    rb->seq_off.push_back(rb->seq.size()); rb->qual_off.push_back(rb->qual.size());
    // NM aux tag (bam_aux_get + bam_aux2i); -1 if absent
    int32_t nm = -1;
    const uint8_t *end = x + total;
    const uint8_t *s = aux;
    while (end - s >= 3) {
        const bool hit = (s[0] == 'N' && s[1] == 'M');
        const uint8_t type = s[2];
        const uint8_t *val = s + 3;
        int sz = 0;
        if (type == 'A' || type == 'c' || type == 'C') { sz = 1; } else if (type == 's' || type == 'S') { sz = 2; }
        else if (type == 'i' || type == 'I' || type == 'f') { sz = 4; } else if (type == 'd') { sz = 8; }
        if (sz) {
            if (hit) {
                if (type == 'c') { nm = (int8_t)val[0]; } else if (type == 'C') { nm = val[0]; }
                else if (type == 's') { nm = (int16_t)(val[0] | (val[1] << 8)); } else if (type == 'S') { nm = (uint16_t)(val[0] | (val[1] << 8)); }
                else if (type == 'i' || type == 'I') { nm = (int32_t)le32(val); } else { nm = 0; }
                break;
            }
            s = val + sz;
        } else if (type == 'Z' || type == 'H') {
            if (hit) { nm = 0; break; }
            s = val;
            while (s < end && *s) { s++; }
            s++;
        } else if (type == 'B') {
            if (end - val < 5) { break; }
            const uint8_t st = val[0];
            const int esz = ((st == 'c' || st == 'C') ? 1 : ((st == 's' || st == 'S') ? 2 : 4));
            s = val + 5 + (size_t)esz * le32(val + 1);
        } else {
            break;
        }
    }
    rb->nm.push_back(nm);
}

} // namespace

extern "C" {

int64_t uvchost_bam_fetch(uvchost_bam *b, int32_t tid, int64_t beg, int64_t end, uvchost_readbuf *rb) {
    if (b->index_status != 1) { return -1; }
    if (beg < 0) { beg = 0; }
    if (tid < 0 || (size_t)tid >= b->lidx.size() || beg >= end) { return 0; }
    const std::vector<uint64_t> & l = b->lidx[tid];
    size_t w = (size_t)(beg >> 14);
    while (w < l.size() && 0 == l[w]) { w++; }
    if (w >= l.size()) { return 0; }
    if (b->in.seek((int64_t)l[w]) != 0) { return -1; }
    int64_t n = 0;
    Core c;
    for (;;) {
        const int r = next_record(b, c);
        if (r < 0) { return -1; }
        if (0 == r) { break; }
        if (c.tid != tid || c.pos >= end) {
            if (c.tid >= 0 && c.tid < tid) { continue; }
            break;
        }
        if (c.endpos > beg) { append_record(rb, b, c); n++; }
    }
    return n;
}

// appends record j of src to dst
static void copy_record(uvchost_readbuf *dst, const uvchost_readbuf *src, size_t j) {
    dst->pos.push_back(src->pos[j]); dst->mpos.push_back(src->mpos[j]); dst->isize.push_back(src->isize[j]); dst->mtid.push_back(src->mtid[j]);
    dst->l_qseq.push_back(src->l_qseq[j]); dst->n_cigar.push_back(src->n_cigar[j]); dst->nm.push_back(src->nm[j]);
    dst->flag.push_back(src->flag[j]); dst->mapq.push_back(src->mapq[j]);
    dst->qname.insert(dst->qname.end(), src->qname.begin() + src->qname_off[j], src->qname.begin() + src->qname_off[j + 1]);
    dst->cigar.insert(dst->cigar.end(), src->cigar.begin() + src->cigar_off[j], src->cigar.begin() + src->cigar_off[j + 1]);
    dst->seq.insert(dst->seq.end(), src->seq.begin() + src->seq_off[j], src->seq.begin() + src->seq_off[j + 1]);
    dst->qual.insert(dst->qual.end(), src->qual.begin() + src->qual_off[j], src->qual.begin() + src->qual_off[j + 1]);
    dst->qname_off.push_back(dst->qname.size()); dst->cigar_off.push_back(dst->cigar.size());
    dst->seq_off.push_back(dst->seq.size()); dst->qual_off.push_back(dst->qual.size());
}

int64_t uvchost_bam_fetch_tiles(uvchost_bam *b, int32_t tid, int32_t n, const int64_t *begs, const int64_t *ends, uvchost_readbuf *rb,
        int64_t *read_begin, int64_t *read_end) {
    if (n <= 0) { return 0; }
    if (b->index_status != 1) { return -1; }
    int64_t span_beg = begs[0], span_end = ends[0], sum_len = 0;
    for (int32_t k = 0; k < n; k++) {
        if (begs[k] < span_beg) { span_beg = begs[k]; }
        if (ends[k] > span_end) { span_end = ends[k]; }
        sum_len += (ends[k] - begs[k]) + 8192;   // a separate query also parses, on average, half an index window before its start
    }
    bool ascending = true;
    for (int32_t k = 0; k + 1 < n; k++) { if (begs[k] > begs[k + 1]) { ascending = false; } }
    int64_t total = 0;
    if (1 == n || !ascending || span_end - span_beg > sum_len) {   // sparse (or unsorted) windows: one index query each
        for (int32_t k = 0; k < n; k++) {
            read_begin[k] = uvchost_readbuf_size(rb);
            const int64_t r = uvchost_bam_fetch(b, tid, begs[k], ends[k], rb);
            if (r < 0) { return r; }
            read_end[k] = uvchost_readbuf_size(rb);
            total += r;
        }
        return total;
    }
    if (NULL == b->span) { b->span = uvchost_readbuf_new(); }
    uvchost_readbuf *sp = b->span;
    uvchost_readbuf_clear(sp);
    sp->endpos_scratch.clear();
    {   // one pass over the span, keeping every record's end position
        int64_t beg = (span_beg < 0 ? 0 : span_beg);
        if (uvchost_bam_seek_region(b, tid, beg) == 0) {
            Core c;
            for (;;) {
                const int r = next_record(b, c);
                if (r < 0) { return -1; }
                if (0 == r) { break; }
                if (c.tid != tid || c.pos >= span_end) {
                    if (c.tid >= 0 && c.tid < tid) { continue; }
                    break;
                }
                if (c.endpos > beg) { append_record(sp, b, c); sp->endpos_scratch.push_back(c.endpos); }
            }
        }
    }
    const size_t m = sp->pos.size();
    size_t lo = 0;
    for (int32_t k = 0; k < n; k++) {
        const int64_t beg = (begs[k] < 0 ? 0 : begs[k]), end = ends[k];
        read_begin[k] = uvchost_readbuf_size(rb);
        if (beg < end) {
            while (lo < m && sp->endpos_scratch[lo] <= beg) { lo++; }   // ascending windows: such a record overlaps no later window either
            for (size_t j = lo; j < m && sp->pos[j] < end; j++) {
                if (sp->endpos_scratch[j] > beg) { copy_record(rb, sp, j); total++; }
            }
        }
        read_end[k] = uvchost_readbuf_size(rb);
    }
    return total;
}

int64_t uvchost_bam_fetch_span(uvchost_bam *b, int32_t tid, int32_t n, const int64_t *begs, const int64_t *ends, uvchost_readbuf *rb,
        int64_t *read_begin, int64_t *read_end) {
    if (n <= 0) { return 0; }
    if (b->index_status != 1) { return -1; }
    for (int32_t k = 0; k + 1 < n; k++) { if (begs[k] > begs[k + 1] || ends[k] > ends[k + 1]) { return -2; } }
    const int64_t span_beg = (begs[0] < 0 ? 0 : begs[0]), span_end = ends[n - 1];
    const int64_t base = uvchost_readbuf_size(rb);
    {   // room for the records the index says the span holds: growing the arrays by doubling copies every byte about once more and faults its pages in
        const int64_t est = uvchost_bam_estimate_reads(b, tid, span_beg, span_end);
        if (est > 4096 && est < ((int64_t)1 << 28)) {
            const size_t m = (size_t)base + (size_t)est + (size_t)est / 8;
            if (m > rb->pos.capacity()) {
                rb->pos.reserve(m); rb->mpos.reserve(m); rb->isize.reserve(m); rb->mtid.reserve(m); rb->l_qseq.reserve(m); rb->n_cigar.reserve(m); rb->nm.reserve(m);
                rb->flag.reserve(m); rb->mapq.reserve(m);
                rb->seq_off.reserve(m + 1); rb->qual_off.reserve(m + 1); rb->cigar_off.reserve(m + 1); rb->qname_off.reserve(m + 1);
                // (bytes per record from what the buffer already holds, else typical short-read sizes)
                const size_t have = (size_t)base;
                const size_t seq_b = (have > 1000 ? rb->seq.size() / have + 1 : 80), qual_b = (have > 1000 ? rb->qual.size() / have + 1 : 160);
                const size_t cig_w = (have > 1000 ? rb->cigar.size() / have + 1 : 3), name_b = (have > 1000 ? rb->qname.size() / have + 1 : 48);
                rb->seq.reserve(rb->seq.size() + ((size_t)est + (size_t)est / 8) * seq_b); rb->qual.reserve(rb->qual.size() + ((size_t)est + (size_t)est / 8) * qual_b);
                rb->cigar.reserve(rb->cigar.size() + ((size_t)est + (size_t)est / 8) * cig_w); rb->qname.reserve(rb->qname.size() + ((size_t)est + (size_t)est / 8) * name_b);
            }
        }
    }
    std::vector<int32_t> & endpos = rb->endpos_scratch;
    endpos.clear();
    int64_t total = 0;
    if (uvchost_bam_seek_region(b, tid, span_beg) == 0) {
        Core c;
        int32_t k_open = 0;      // first window that can still need a record (ascending begs)
        for (;;) {
            const int r = next_record(b, c);
            if (r < 0) { return -1; }
            if (0 == r) { break; }
            if (c.tid != tid || c.pos >= span_end) {
                if (c.tid >= 0 && c.tid < tid) { continue; }
                break;
            }
            // keep the record if it overlaps some window: ends are ascending, so windows whose end <= pos are closed for good
            while (k_open < n && ends[k_open] <= c.pos) { k_open++; }
            bool wanted = false;
            for (int32_t k = k_open; k < n && begs[k] < c.endpos; k++) { if (c.pos < ends[k]) { wanted = true; break; } }
            if (wanted) { append_record(rb, b, c); endpos.push_back(c.endpos); total++; }
        }
    }
    // slices: window k = [first kept record with endpos > beg, first kept record with pos >= end)
    const size_t m = endpos.size();
    size_t lo = 0, hi = 0;
    for (int32_t k = 0; k < n; k++) {
        const int64_t beg = (begs[k] < 0 ? 0 : begs[k]), end = ends[k];
        while (lo < m && endpos[lo] <= beg) { lo++; }
        if (hi < lo) { hi = lo; }
        while (hi < m && rb->pos[(size_t)base + hi] < end) { hi++; }
        read_begin[k] = base + (int64_t)lo; read_end[k] = base + (int64_t)(hi > lo ? hi : lo);
    }
    return total;
}

int uvchost_bam_scan(uvchost_bam *b, uvchost_scan_cb cb, void *user) {
    if (b->in.seek(b->first_record_voff) != 0) { return -1; }
    Core c;
    for (;;) {
        const int r = next_record(b, c);
        if (r < 0) { return -1; }
        if (0 == r) { break; }
        if (cb(c.tid, c.pos, c.endpos, (uint16_t)c.flag, c.isize, c.l_qseq, (uint8_t)c.mapq, user)) { break; }
    }
    return 0;
}

int uvchost_bam_rewind(uvchost_bam *b) { return b->in.seek(b->first_record_voff); }

int uvchost_bam_rewind_parallel(uvchost_bam *b, int n_threads) {
    if (b->in.seek(b->first_record_voff) != 0) { return -1; }
    if (n_threads <= 1) { return 0; }
    // the block that holds the first record is already inflated: the read-ahead continues with the block after it
    SeqInflater *q = new SeqInflater();
    q->start(b->path.c_str(), b->in.block_addr + b->in.block_clen, n_threads);
    if (NULL == q->fp) { delete q; return 0; }
    b->in.seq = q;
    return 0;
}

int uvchost_bam_next_core(uvchost_bam *b, uvchost_core *out) {
    Core c;
    const int r = next_record(b, c);
    if (r <= 0) { return r; }
    out->tid = c.tid; out->pos = c.pos; out->endpos = c.endpos; out->isize = c.isize; out->l_qseq = c.l_qseq; out->flag = (uint16_t)c.flag; out->mapq = (uint8_t)c.mapq;
    return 1;
}

int uvchost_bam_seek_region(uvchost_bam *b, int32_t tid, int64_t beg) {
    if (b->index_status != 1) { return -1; }
    if (beg < 0) { beg = 0; }
    if (tid < 0 || (size_t)tid >= b->lidx.size()) { return 1; }
    const std::vector<uint64_t> & l = b->lidx[tid];
    size_t w = (size_t)(beg >> 14);
    while (w < l.size() && 0 == l[w]) { w++; }
    if (w >= l.size()) { return 1; }
    return (b->in.seek((int64_t)l[w]) != 0 ? -1 : 0);
}

int64_t uvchost_bam_estimate_reads(const uvchost_bam *b, int32_t tid, int64_t beg, int64_t end) {
    if (b->index_status != 1 || tid < 0 || (size_t)tid >= b->lidx.size() || beg >= end) { return 0; }
    const std::vector<uint64_t> & l = b->lidx[tid];
    if (l.empty()) { return 0; }
    if (beg < 0) { beg = 0; }
    size_t w0 = (size_t)(beg >> 14), w1 = (size_t)((end - 1) >> 14) + 1;
    if (w0 >= l.size()) { return 0; }
    // compressed bytes between the first record of window w0 and the first record of the window after w1 (or of the last window)
    size_t a = w0, z = (w1 < l.size() ? w1 : l.size() - 1);
    while (a < l.size() && 0 == l[a]) { a++; }
    while (z > a && 0 == l[z]) { z--; }
    if (a >= l.size() || z <= a) {
        // a single window: fall back on the file's average density over the contig's windows
        a = 0; z = l.size() - 1;
        while (a < l.size() && 0 == l[a]) { a++; }
        while (z > a && 0 == l[z]) { z--; }
        if (a >= l.size() || z <= a) { return 0; }
    }
    const double bytes = (double)((l[z] >> 16) - (l[a] >> 16));
    const double positions = (double)(z - a) * 16384.0;
    const double density = bytes / positions;                     // compressed bytes per reference position
    return (int64_t)(density * (double)(end - beg + 300) / 50.0) + 1;   // ~50 compressed bytes per record: errs on the high side
}

int64_t uvchost_bam_count(uvchost_bam *b, int32_t tid, int64_t beg, int64_t end) {
    if (uvchost_bam_seek_region(b, tid, beg) != 0) { return 0; }
    int64_t n = 0;
    Core c;
    for (;;) {
        const int r = next_record(b, c);
        if (r <= 0) { break; }
        if (c.tid != tid || c.pos >= end) {
            if (c.tid >= 0 && c.tid < tid) { continue; }
            break;
        }
        if (c.endpos > beg) { n++; }
    }
    return n;
}

int uvchost_bam_infer(uvchost_bam *b, int64_t max_records, uvchost_infer_stats *out) {
    memset(out, 0, sizeof(*out));
    if (b->in.seek(b->first_record_voff) != 0) { return -1; }
    std::vector<int32_t> qlens;
    qlens.push_back(150);
    Core c;
    while ((out->count_pe + out->count_se) < max_records) {
        const int r = next_record(b, c);
        if (r < 0) { return -1; }
        if (0 == r) { break; }
        if (c.mapq > out->max_mapq) { out->max_mapq = c.mapq; }
        if (c.flag & 0x1) { out->count_pe++; } else { out->count_se++; }
        qlens.push_back(c.l_qseq);
        const uint8_t *ql = b->rec.data() + 32 + c.l_qname + 4 * (size_t)c.n_cigar + (size_t)((c.l_qseq + 1) / 2);
        for (int32_t i = 0; i < c.l_qseq; i++) {
            if (ql[i] < 30) { out->q30_n_fail_bases++; } else { out->q30_n_pass_bases++; }
            if (ql[i] < 20) { out->q20_n_fail_bases++; }
        }
    }
    std::sort(qlens.begin(), qlens.end());
    out->median_qlen = qlens[qlens.size() / 2];
    out->max_qlen = qlens.back();
    return 0;
}

uvchost_fasta *uvchost_fasta_open(const char *path) {
    const std::string fai = std::string(path) + ".fai";
    FILE *fi = fopen(fai.c_str(), "r");
    if (NULL == fi) { return NULL; }
    uvchost_fasta *f = new uvchost_fasta();
    f->fp = fopen(path, "rb");
    if (NULL == f->fp) { fclose(fi); delete f; return NULL; }
    char line[4096];
    while (fgets(line, sizeof(line), fi)) {
        char name[2048];
        long long len, off; int lb, lw;
        if (sscanf(line, "%2047s\t%lld\t%lld\t%d\t%d", name, &len, &off, &lb, &lw) == 5) {
            uvchost_fasta::Entry e; e.len = len; e.offset = off; e.linebases = lb; e.linewidth = lw;
            f->entries[name] = e;
        }
    }
    fclose(fi);
    return f;
}

void uvchost_fasta_close(uvchost_fasta *f) { if (f) { if (f->fp) { fclose(f->fp); } delete f; } }

char *uvchost_fasta_fetch_contig(uvchost_fasta *f, const char *name, int64_t *len) {
    auto it = f->entries.find(name);
    if (it == f->entries.end()) { *len = 0; return NULL; }
    const uvchost_fasta::Entry & e = it->second;
    char *out = (char*)malloc((size_t)e.len + 1);
    const int64_t nlines = (e.len + e.linebases - 1) / e.linebases;
    const int64_t nbytes = e.len + nlines * (e.linewidth - e.linebases);
    std::vector<char> raw((size_t)nbytes + 16);
    fseeko(f->fp, e.offset, SEEK_SET);
    const size_t got = fread(raw.data(), 1, (size_t)nbytes, f->fp);
    int64_t k = 0;
    for (size_t i = 0; i < got && k < e.len; i++) { if ((unsigned char)raw[i] > ' ') { out[k++] = raw[i]; } }
    out[k] = 0;
    *len = k;
    return out;
}

} // extern "C"
