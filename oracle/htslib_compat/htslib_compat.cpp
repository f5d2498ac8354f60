/* htslib-compat shim: TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Implements, over zlib, the 30 htslib-1.11 symbols that the unmodified reference
 * (genetronhealth/uvc 0.15.1) needs for a tumor-only run, plus loud-failing stubs
 * for the 17 VCF/BCF symbols used only by --tumor-vcf.
 * htslib itself is an un-vendored dependency of the reference (Makefile:16-17,
 * install-dependencies.sh:13-25, pinned version 1.11) and there is no network, so
 * this file restates the published formats (SAM/BAM spec section 4 for BGZF+BAM,
 * section 5 for the BAI index; faidx .fai five-column format). None of the
 * path's arithmetic lives here - htslib only decodes bytes.
 *
 * Query semantics reproduced from htslib: sam_itr_next yields records with
 * tid == query tid, pos < end and bam_endpos > beg; bam_endpos treats a zero
 * reference length as 1.
 */
#include "htslib/bgzf.h"
#include "htslib/faidx.h"
#include "htslib/hts.h"
#include "htslib/sam.h"
#include "htslib/synced_bcf_reader.h"
#include "htslib/vcf.h"

#include <zlib.h>

#include <map>
#include <string>
#include <vector>

#include <assert.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

extern "C" {

const unsigned char seq_nt16_table[256] = {
    15,15,15,15, 15,15,15,15, 15,15,15,15, 15,15,15,15,
    15,15,15,15, 15,15,15,15, 15,15,15,15, 15,15,15,15,
    15,15,15,15, 15,15,15,15, 15,15,15,15, 15,15,15,15,
     1, 2, 4, 8, 15,15,15,15, 15,15,15,15, 15, 0 /*=*/,15,15,
    15, 1,14, 2, 13,15,15, 4, 11,15,15,12, 15, 3,15,15,
    15,15, 5, 6,  8,15, 7, 9, 15,10,15,15, 15,15,15,15,
    15, 1,14, 2, 13,15,15, 4, 11,15,15,12, 15, 3,15,15,
    15,15, 5, 6,  8,15, 7, 9, 15,10,15,15, 15,15,15,15,
    15,15,15,15, 15,15,15,15, 15,15,15,15, 15,15,15,15,
    15,15,15,15, 15,15,15,15, 15,15,15,15, 15,15,15,15,
    15,15,15,15, 15,15,15,15, 15,15,15,15, 15,15,15,15,
    15,15,15,15, 15,15,15,15, 15,15,15,15, 15,15,15,15,
    15,15,15,15, 15,15,15,15, 15,15,15,15, 15,15,15,15,
    15,15,15,15, 15,15,15,15, 15,15,15,15, 15,15,15,15,
    15,15,15,15, 15,15,15,15, 15,15,15,15, 15,15,15,15,
    15,15,15,15, 15,15,15,15, 15,15,15,15, 15,15,15,15
};
const char seq_nt16_str[] = "=ACMGRSVTWYHKDBN";
const int seq_nt16_int[] = { 4, 0, 1, 4, 2, 4, 4, 4, 3, 4, 4, 4, 4, 4, 4, 4 };

/* ------------------------------------------------------------------ BGZF */

static void die(const char *msg) {
    fprintf(stderr, "[htslib-compat] fatal: %s\n", msg);
    abort();
}

BGZF *bgzf_open(const char *path, const char *mode) {
    BGZF *fp = (BGZF*)calloc(1, sizeof(BGZF));
    fp->is_write = (strchr(mode, 'w') != NULL);
    fp->fp = fopen(path, fp->is_write ? "wb" : "rb");
    if (NULL == fp->fp) { free(fp); return NULL; }
    fp->compress_level = -1;
    if (fp->is_write) {
        fp->wbuf = (uint8_t*)malloc(BGZF_MAX_BLOCK_SIZE);
    } else {
        fp->ublock = (uint8_t*)malloc(BGZF_MAX_BLOCK_SIZE);
        fp->block_address = 0;
    }
    return fp;
}

int bgzf_compress(void *_dst, size_t *dlen, const void *src, size_t slen, int level) {
    uint8_t *dst = (uint8_t*)_dst;
    static const uint8_t header[18] = {31,139,8,4, 0,0,0,0, 0,255, 6,0, 'B','C',2,0, 0,0};
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    zs.next_in = (Bytef*)src;
    zs.avail_in = slen;
    zs.next_out = dst + 18;
    zs.avail_out = BGZF_MAX_BLOCK_SIZE - 18 - 8;
    if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) { return -1; }
    if (deflate(&zs, Z_FINISH) != Z_STREAM_END) { deflateEnd(&zs); return -1; }
    deflateEnd(&zs);
    size_t total = zs.total_out + 18 + 8;
    memcpy(dst, header, 18);
    dst[16] = (uint8_t)((total - 1) & 0xff);
    dst[17] = (uint8_t)((total - 1) >> 8);
    uint32_t crc = crc32(crc32(0L, NULL, 0L), (const Bytef*)src, slen);
    for (int i = 0; i < 4; i++) { dst[total - 8 + i] = (uint8_t)(crc >> (8 * i)); }
    for (int i = 0; i < 4; i++) { dst[total - 4 + i] = (uint8_t)(((uint32_t)slen) >> (8 * i)); }
    *dlen = total;
    return 0;
}

int bgzf_flush(BGZF *fp) {
    if (!fp->is_write) { return 0; }
    while (fp->wbuf_len > 0) {
        uint8_t cbuf[BGZF_MAX_BLOCK_SIZE];
        size_t clen = 0;
        int n = (fp->wbuf_len > BGZF_BLOCK_SIZE ? BGZF_BLOCK_SIZE : fp->wbuf_len);
        if (bgzf_compress(cbuf, &clen, fp->wbuf, n, fp->compress_level) != 0) { return -1; }
        if (fwrite(cbuf, 1, clen, fp->fp) != clen) { return -1; }
        memmove(fp->wbuf, fp->wbuf + n, fp->wbuf_len - n);
        fp->wbuf_len -= n;
    }
    return (fflush(fp->fp) == 0 ? 0 : -1);
}

ssize_t bgzf_write(BGZF *fp, const void *data, size_t length) {
    const uint8_t *in = (const uint8_t*)data;
    size_t remaining = length;
    while (remaining > 0) {
        size_t n = BGZF_BLOCK_SIZE - fp->wbuf_len;
        if (n > remaining) { n = remaining; }
        memcpy(fp->wbuf + fp->wbuf_len, in, n);
        fp->wbuf_len += n;
        in += n;
        remaining -= n;
        if (fp->wbuf_len == BGZF_BLOCK_SIZE) {
            if (bgzf_flush(fp) != 0) { return -1; }
        }
    }
    return length;
}

ssize_t bgzf_raw_write(BGZF *fp, const void *data, size_t length) {
    size_t ret = fwrite(data, 1, length, fp->fp);
    return (ssize_t)ret;
}

int bgzf_close(BGZF *fp) {
    if (NULL == fp) { return 0; }
    int ret = 0;
    if (fp->is_write) {
        if (bgzf_flush(fp) != 0) { ret = -1; }
        uint8_t cbuf[BGZF_MAX_BLOCK_SIZE];
        size_t clen = 0;
        bgzf_compress(cbuf, &clen, NULL, 0, fp->compress_level); /* EOF marker block */
        if (fwrite(cbuf, 1, clen, fp->fp) != clen) { ret = -1; }
        free(fp->wbuf);
    } else {
        free(fp->ublock);
    }
    if (fclose(fp->fp) != 0) { ret = -1; }
    free(fp);
    return ret;
}

/* Load the block starting at compressed offset caddr. Returns 0 ok, 1 EOF, -1 error. */
static int bgzf_load_block(BGZF *fp, int64_t caddr) {
    uint8_t hdr[18];
    if (fseeko(fp->fp, caddr, SEEK_SET) != 0) { return -1; }
    size_t n = fread(hdr, 1, 18, fp->fp);
    if (0 == n) { fp->block_address = caddr; fp->block_clen = 0; fp->ublock_len = 0; fp->ublock_off = 0; return 1; }
    if (n != 18 || hdr[0] != 31 || hdr[1] != 139 || hdr[2] != 8 || !(hdr[3] & 4)) { return -1; }
    int xlen = hdr[10] | (hdr[11] << 8);
    int bsize = -1;
    std::vector<uint8_t> extra(xlen);
    memcpy(extra.data(), hdr + 12, (xlen < 6 ? xlen : 6));
    if (xlen > 6) {
        if (fread(extra.data() + 6, 1, xlen - 6, fp->fp) != (size_t)(xlen - 6)) { return -1; }
    }
    for (int off = 0; off + 4 <= xlen; ) {
        int slen = extra[off + 2] | (extra[off + 3] << 8);
        if (extra[off] == 'B' && extra[off + 1] == 'C' && slen == 2) { bsize = (extra[off + 4] | (extra[off + 5] << 8)) + 1; }
        off += 4 + slen;
    }
    if (bsize < 0) { return -1; }
    int cdata_len = bsize - 12 - xlen - 8;
    std::vector<uint8_t> cdata(cdata_len + 8);
    if (fread(cdata.data(), 1, cdata_len + 8, fp->fp) != (size_t)(cdata_len + 8)) { return -1; }
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    zs.next_in = cdata.data();
    zs.avail_in = cdata_len;
    zs.next_out = fp->ublock;
    zs.avail_out = BGZF_MAX_BLOCK_SIZE;
    if (inflateInit2(&zs, -15) != Z_OK) { return -1; }
    int zret = inflate(&zs, Z_FINISH);
    inflateEnd(&zs);
    if (zret != Z_STREAM_END) { return -1; }
    fp->ublock_len = zs.total_out;
    fp->ublock_off = 0;
    fp->block_address = caddr;
    fp->block_clen = bsize;
    return 0;
}

ssize_t bgzf_read(BGZF *fp, void *data, size_t length) {
    uint8_t *out = (uint8_t*)data;
    size_t done = 0;
    while (done < length) {
        if (fp->ublock_off >= fp->ublock_len) {
            int r = bgzf_load_block(fp, fp->block_address + fp->block_clen);
            if (r < 0) { return -1; }
            if (r == 1) { break; }
            if (0 == fp->ublock_len) { continue; } /* empty (EOF-marker) block; try the next one */
        }
        size_t n = fp->ublock_len - fp->ublock_off;
        if (n > length - done) { n = length - done; }
        memcpy(out + done, fp->ublock + fp->ublock_off, n);
        fp->ublock_off += n;
        done += n;
    }
    return done;
}

int64_t bgzf_tell_compat(BGZF *fp) {
    if (fp->ublock_off >= fp->ublock_len && fp->block_clen > 0) {
        return ((fp->block_address + fp->block_clen) << 16);
    }
    return (fp->block_address << 16) | (fp->ublock_off & 0xffff);
}

int bgzf_seek_compat(BGZF *fp, int64_t voffset) {
    int64_t caddr = voffset >> 16;
    int uoff = voffset & 0xffff;
    if (!(fp->block_clen > 0 && fp->block_address == caddr)) {
        int r = bgzf_load_block(fp, caddr);
        if (r < 0) { return -1; }
    }
    fp->ublock_off = uoff;
    return 0;
}

/* ------------------------------------------------------------------ hts/sam files */

htsFile *hts_open(const char *fn, const char *mode) {
    if (strchr(mode, 'w')) { die("hts_open for writing is not supported by the shim"); }
    BGZF *bg = bgzf_open(fn, "r");
    if (NULL == bg) { return NULL; }
    htsFile *fp = (htsFile*)calloc(1, sizeof(htsFile));
    fp->bgzf = bg;
    fp->fn = strdup(fn);
    return fp;
}

int hts_close(htsFile *fp) {
    if (NULL == fp) { return 0; }
    int ret = bgzf_close(fp->bgzf);
    free(fp->fn);
    free(fp);
    return ret;
}

samFile *sam_open(const char *fn, const char *mode) { return hts_open(fn, mode); }
int sam_close(samFile *fp) { return hts_close(fp); }

static int read_i32(BGZF *fp, int32_t *v) {
    uint8_t b[4];
    if (bgzf_read(fp, b, 4) != 4) { return -1; }
    *v = (int32_t)((uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24));
    return 0;
}

sam_hdr_t *sam_hdr_read(samFile *fp) {
    BGZF *bg = fp->bgzf;
    if (bgzf_seek_compat(bg, 0) != 0) { return NULL; }
    char magic[4];
    if (bgzf_read(bg, magic, 4) != 4 || memcmp(magic, "BAM\1", 4) != 0) { return NULL; }
    sam_hdr_t *h = (sam_hdr_t*)calloc(1, sizeof(sam_hdr_t));
    int32_t l_text = 0;
    if (read_i32(bg, &l_text) != 0) { free(h); return NULL; }
    h->l_text = l_text;
    h->text = (char*)calloc(l_text + 1, 1);
    if (bgzf_read(bg, h->text, l_text) != l_text) { return NULL; }
    if (read_i32(bg, &h->n_targets) != 0) { return NULL; }
    h->target_name = (char**)calloc(h->n_targets > 0 ? h->n_targets : 1, sizeof(char*));
    h->target_len = (uint32_t*)calloc(h->n_targets > 0 ? h->n_targets : 1, sizeof(uint32_t));
    for (int32_t i = 0; i < h->n_targets; i++) {
        int32_t l_name = 0, l_ref = 0;
        if (read_i32(bg, &l_name) != 0) { return NULL; }
        h->target_name[i] = (char*)calloc(l_name + 1, 1);
        if (bgzf_read(bg, h->target_name[i], l_name) != l_name) { return NULL; }
        if (read_i32(bg, &l_ref) != 0) { return NULL; }
        h->target_len[i] = (uint32_t)l_ref;
    }
    return h;
}

void bam_hdr_destroy(sam_hdr_t *h) {
    if (NULL == h) { return; }
    for (int32_t i = 0; i < h->n_targets; i++) { free(h->target_name[i]); }
    free(h->target_name);
    free(h->target_len);
    free(h->text);
    free(h);
}

bam1_t *bam_init1(void) { return (bam1_t*)calloc(1, sizeof(bam1_t)); }

void bam_destroy1(bam1_t *b) {
    if (NULL == b) { return; }
    free(b->data);
    free(b);
}

bam1_t *bam_dup1(const bam1_t *bsrc) {
    if (NULL == bsrc) { return NULL; }
    bam1_t *b = bam_init1();
    *b = *bsrc;
    b->m_data = (bsrc->l_data > 0 ? bsrc->l_data : 1);
    b->data = (uint8_t*)malloc(b->m_data);
    memcpy(b->data, bsrc->data, bsrc->l_data);
    return b;
}

static hts_pos_t cigar2rlen(uint32_t n_cigar, const uint32_t *cigar) {
    hts_pos_t l = 0;
    for (uint32_t k = 0; k < n_cigar; k++) {
        if (bam_cigar_type(bam_cigar_op(cigar[k])) & 2) { l += bam_cigar_oplen(cigar[k]); }
    }
    return l;
}

hts_pos_t bam_endpos(const bam1_t *b) {
    hts_pos_t rlen = ((b->core.flag & BAM_FUNMAP) ? 0 : cigar2rlen(b->core.n_cigar, bam_get_cigar(b)));
    if (0 == rlen) { rlen = 1; }
    return b->core.pos + rlen;
}

/* Reads one BAM record at the current position. >=0 ok, -1 EOF, < -1 error. */
static int bam_read1_compat(BGZF *bg, bam1_t *b) {
    int32_t block_len = 0;
    uint8_t lb[4];
    ssize_t n = bgzf_read(bg, lb, 4);
    if (0 == n) { return -1; }
    if (4 != n) { return -4; }
    block_len = (int32_t)((uint32_t)lb[0] | ((uint32_t)lb[1] << 8) | ((uint32_t)lb[2] << 16) | ((uint32_t)lb[3] << 24));
    if (block_len < 32) { return -4; }
    uint8_t x[32];
    if (bgzf_read(bg, x, 32) != 32) { return -3; }
    auto u32 = [&](int off) { return (uint32_t)x[off] | ((uint32_t)x[off+1] << 8) | ((uint32_t)x[off+2] << 16) | ((uint32_t)x[off+3] << 24); };
    bam1_core_t *c = &b->core;
    c->tid = (int32_t)u32(0);
    c->pos = (int32_t)u32(4);
    uint32_t bin_mq_nl = u32(8);
    c->bin = bin_mq_nl >> 16;
    c->qual = (bin_mq_nl >> 8) & 0xff;
    c->l_qname = bin_mq_nl & 0xff;
    c->l_extranul = 0;
    uint32_t flag_nc = u32(12);
    c->flag = flag_nc >> 16;
    c->n_cigar = flag_nc & 0xffff;
    c->l_qseq = (int32_t)u32(16);
    c->mtid = (int32_t)u32(20);
    c->mpos = (int32_t)u32(24);
    c->isize = (int32_t)u32(28);
    int l_data = block_len - 32;
    if ((uint32_t)l_data > b->m_data) {
        b->m_data = l_data + 64;
        b->data = (uint8_t*)realloc(b->data, b->m_data);
    }
    b->l_data = l_data;
    if (bgzf_read(bg, b->data, l_data) != l_data) { return -4; }
    return 4 + block_len;
}

int sam_read1(samFile *fp, sam_hdr_t *h, bam1_t *b) {
    (void)h;
    int r = bam_read1_compat(fp->bgzf, b);
    return (r >= 0 ? r : r);
}

static int aux_type2size(uint8_t type) {
    switch (type) {
        case 'A': case 'c': case 'C': return 1;
        case 's': case 'S': return 2;
        case 'i': case 'I': case 'f': return 4;
        case 'd': return 8;
        default: return 0;
    }
}

uint8_t *bam_aux_get(const bam1_t *b, const char tag[2]) {
    uint8_t *s = bam_get_aux(b);
    uint8_t *end = b->data + b->l_data;
    while (s != NULL && end - s >= 3) {
        const bool hit = (s[0] == (uint8_t)tag[0] && s[1] == (uint8_t)tag[1]);
        uint8_t *val = s + 2;
        if (hit) { return val; }
        uint8_t type = *val;
        s = val + 1;
        if (type == 'Z' || type == 'H') {
            while (s < end && *s) { s++; }
            s++;
        } else if (type == 'B') {
            if (end - s < 5) { return NULL; }
            int sz = aux_type2size(*s);
            uint32_t cnt = (uint32_t)s[1] | ((uint32_t)s[2] << 8) | ((uint32_t)s[3] << 16) | ((uint32_t)s[4] << 24);
            s += 5 + (size_t)sz * cnt;
        } else {
            int sz = aux_type2size(type);
            if (0 == sz) { return NULL; }
            s += sz;
        }
    }
    return NULL;
}

int64_t bam_aux2i(const uint8_t *s) {
    uint8_t type = *s++;
    switch (type) {
        case 'c': return (int8_t)s[0];
        case 'C': return s[0];
        case 's': return (int16_t)((uint16_t)s[0] | ((uint16_t)s[1] << 8));
        case 'S': return (uint16_t)((uint16_t)s[0] | ((uint16_t)s[1] << 8));
        case 'i': return (int32_t)((uint32_t)s[0] | ((uint32_t)s[1] << 8) | ((uint32_t)s[2] << 16) | ((uint32_t)s[3] << 24));
        case 'I': return (uint32_t)((uint32_t)s[0] | ((uint32_t)s[1] << 8) | ((uint32_t)s[2] << 16) | ((uint32_t)s[3] << 24));
        default: return 0;
    }
}

/* ------------------------------------------------------------------ BAI index */

struct hts_idx_t {
    /* per reference: linear index (16 kbp windows) of virtual file offsets, plus the smallest chunk offset of any bin */
    std::vector<std::vector<uint64_t>> lidx;
    std::vector<uint64_t> first_off;
    std::vector<bool> has_data;
};

hts_idx_t *sam_index_load2(samFile *fp, const char *fn, const char *fnidx) {
    (void)fp;
    std::string idxfn = (fnidx != NULL ? std::string(fnidx) : (std::string(fn) + ".bai"));
    FILE *f = fopen(idxfn.c_str(), "rb");
    if (NULL == f && NULL == fnidx) {
        std::string alt(fn);
        if (alt.size() > 4 && alt.substr(alt.size() - 4) == ".bam") {
            alt = alt.substr(0, alt.size() - 4) + ".bai";
            f = fopen(alt.c_str(), "rb");
        }
    }
    if (NULL == f) { return NULL; }
    auto rd32 = [&](uint32_t &v) { uint8_t b[4]; if (fread(b, 1, 4, f) != 4) { return false; } v = (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24); return true; };
    auto rd64 = [&](uint64_t &v) { uint32_t lo, hi; if (!rd32(lo) || !rd32(hi)) { return false; } v = (uint64_t)lo | ((uint64_t)hi << 32); return true; };
    char magic[4];
    if (fread(magic, 1, 4, f) != 4 || memcmp(magic, "BAI\1", 4) != 0) { fclose(f); return NULL; }
    uint32_t n_ref = 0;
    if (!rd32(n_ref)) { fclose(f); return NULL; }
    hts_idx_t *idx = new hts_idx_t();
    idx->lidx.resize(n_ref);
    idx->first_off.assign(n_ref, UINT64_MAX);
    idx->has_data.assign(n_ref, false);
    for (uint32_t r = 0; r < n_ref; r++) {
        uint32_t n_bin = 0;
        if (!rd32(n_bin)) { delete idx; fclose(f); return NULL; }
        for (uint32_t bi = 0; bi < n_bin; bi++) {
            uint32_t bin = 0, n_chunk = 0;
            if (!rd32(bin) || !rd32(n_chunk)) { delete idx; fclose(f); return NULL; }
            for (uint32_t ci = 0; ci < n_chunk; ci++) {
                uint64_t cb = 0, ce = 0;
                if (!rd64(cb) || !rd64(ce)) { delete idx; fclose(f); return NULL; }
                if (bin != 37450) { /* 37450 is the metadata pseudo-bin */
                    idx->has_data[r] = true;
                    if (cb < idx->first_off[r]) { idx->first_off[r] = cb; }
                }
            }
        }
        uint32_t n_intv = 0;
        if (!rd32(n_intv)) { delete idx; fclose(f); return NULL; }
        idx->lidx[r].resize(n_intv);
        for (uint32_t i = 0; i < n_intv; i++) {
            if (!rd64(idx->lidx[r][i])) { delete idx; fclose(f); return NULL; }
        }
    }
    fclose(f);
    return idx;
}

hts_idx_t *sam_index_load(samFile *fp, const char *fn) { return sam_index_load2(fp, fn, NULL); }

void hts_idx_destroy(hts_idx_t *idx) { delete idx; }

hts_itr_t *sam_itr_queryi(const hts_idx_t *idx, int tid, hts_pos_t beg, hts_pos_t end) {
    if (NULL == idx) { return NULL; }
    hts_itr_t *it = (hts_itr_t*)calloc(1, sizeof(hts_itr_t));
    if (beg < 0) { beg = 0; }
    it->tid = tid;
    it->beg = beg;
    it->end = end;
    it->start_voffset = -1;
    if (tid >= 0 && (size_t)tid < idx->lidx.size() && idx->has_data[tid] && beg < end) {
        const std::vector<uint64_t> &l = idx->lidx[tid];
        /* The linear index gives the smallest offset of a record overlapping the 16 kbp window;
         * empty windows (offset 0) are skipped forward to the next filled one, because every
         * record overlapping [beg, end) either overlaps the window of beg or starts after it. */
        uint64_t off = 0;
        size_t w = (size_t)(beg >> 14);
        if (w < l.size()) {
            while (w < l.size() && 0 == l[w]) { w++; }
            if (w < l.size()) { off = l[w]; }
        }
        if (0 != off) {
            it->start_voffset = (int64_t)off;
        } else {
            it->finished = 1; /* no record at or after this window */
        }
    } else {
        it->finished = 1;
    }
    return it;
}

hts_itr_t *sam_itr_querys(const hts_idx_t *idx, sam_hdr_t *hdr, const char *region) {
    /* "name", "name:beg-end" or "name:pos" (1-based, inclusive); only the first region of a comma list is used by htslib too */
    std::string reg(region);
    size_t comma = reg.find(',');
    if (comma != std::string::npos) { reg = reg.substr(0, comma); }
    std::string name = reg;
    hts_pos_t beg = 0, end = HTS_POS_MAX;
    size_t colon = reg.rfind(':');
    if (colon != std::string::npos) {
        name = reg.substr(0, colon);
        std::string rest = reg.substr(colon + 1);
        size_t dash = rest.find('-');
        if (dash != std::string::npos) {
            beg = atoll(rest.substr(0, dash).c_str()) - 1;
            end = atoll(rest.substr(dash + 1).c_str());
        } else {
            beg = atoll(rest.c_str()) - 1;
        }
    }
    int tid = -1;
    for (int32_t i = 0; i < hdr->n_targets; i++) {
        if (name == hdr->target_name[i]) { tid = i; }
    }
    if (tid < 0) { return NULL; }
    if (end == HTS_POS_MAX) { end = hdr->target_len[tid]; }
    return sam_itr_queryi(idx, tid, beg, end);
}

void hts_itr_destroy(hts_itr_t *iter) { free(iter); }

int sam_itr_next(samFile *fp, hts_itr_t *itr, bam1_t *b) {
    if (NULL == itr || itr->finished) { return -1; }
    BGZF *bg = fp->bgzf;
    if (!itr->positioned) {
        if (bgzf_seek_compat(bg, itr->start_voffset) != 0) { itr->finished = 1; return -2; }
        itr->positioned = 1;
    }
    for (;;) {
        int r = bam_read1_compat(bg, b);
        if (r < 0) { itr->finished = 1; return (r == -1 ? -1 : r); }
        if (b->core.tid != itr->tid || b->core.pos >= itr->end) {
            if (b->core.tid >= 0 && b->core.tid < itr->tid) { continue; }
            itr->finished = 1;
            return -1;
        }
        if (bam_endpos(b) > itr->beg) { return r; }
    }
}

/* ------------------------------------------------------------------ faidx */

struct fai_entry { std::string name; int64_t len, offset; int32_t linebases, linewidth; };
struct faidx_t {
    FILE *fp;
    std::vector<fai_entry> entries;
    std::map<std::string, size_t> name2idx;
};

faidx_t *fai_load(const char *fn) {
    std::string faifn = std::string(fn) + ".fai";
    FILE *fi = fopen(faifn.c_str(), "r");
    if (NULL == fi) { return NULL; }
    FILE *fa = fopen(fn, "rb");
    if (NULL == fa) { fclose(fi); return NULL; }
    faidx_t *fai = new faidx_t();
    fai->fp = fa;
    char line[4096];
    while (fgets(line, sizeof(line), fi)) {
        char name[2048];
        long long len, offset;
        int lb, lw;
        if (sscanf(line, "%2047s\t%lld\t%lld\t%d\t%d", name, &len, &offset, &lb, &lw) == 5) {
            fai_entry e;
            e.name = name; e.len = len; e.offset = offset; e.linebases = lb; e.linewidth = lw;
            fai->name2idx[e.name] = fai->entries.size();
            fai->entries.push_back(e);
        }
    }
    fclose(fi);
    return fai;
}

void fai_destroy(faidx_t *fai) {
    if (NULL == fai) { return; }
    fclose(fai->fp);
    delete fai;
}

const char *faidx_iseq(const faidx_t *fai, int i) { return fai->entries.at(i).name.c_str(); }

char *faidx_fetch_seq(const faidx_t *fai, const char *c_name, int p_beg_i, int p_end_i, int *len) {
    auto it = fai->name2idx.find(c_name);
    if (it == fai->name2idx.end()) { *len = -2; return NULL; }
    const fai_entry &e = fai->entries[it->second];
    int64_t beg = p_beg_i, end = (int64_t)p_end_i + 1; /* p_end_i is inclusive */
    if (beg < 0) { beg = 0; }
    if (end > e.len) { end = e.len; }
    if (beg >= end) { *len = 0; char *s = (char*)calloc(1, 1); return s; }
    char *seq = (char*)malloc(end - beg + 1);
    int64_t l = 0;
    int64_t fileoff = e.offset + (beg / e.linebases) * e.linewidth + (beg % e.linebases);
    fseeko(fai->fp, fileoff, SEEK_SET);
    while (l < end - beg) {
        int ch = fgetc(fai->fp);
        if (EOF == ch) { break; }
        if (ch > ' ' ) { seq[l++] = (char)ch; } /* isgraph */
    }
    seq[l] = '\0';
    *len = (int)l;
    return seq;
}

/* ------------------------------------------------------------------ VCF/BCF stubs (--tumor-vcf only) */

#define UNSUPPORTED(name) die(name " (tumor-normal VCF input) is not supported by the htslib-compat shim")

bcf_hdr_t *bcf_hdr_read(htsFile *fp) { (void)fp; UNSUPPORTED("bcf_hdr_read"); return NULL; }
void bcf_hdr_destroy(bcf_hdr_t *h) { (void)h; }
int bcf_unpack(bcf1_t *b, int which) { (void)b; (void)which; UNSUPPORTED("bcf_unpack"); return -1; }
bcf1_t *bcf_dup(bcf1_t *src) { (void)src; UNSUPPORTED("bcf_dup"); return NULL; }
void bcf_destroy(bcf1_t *v) { (void)v; }
int bcf_get_format_values(const bcf_hdr_t *hdr, bcf1_t *line, const char *tag, void **dst, int *ndst, int type) {
    (void)hdr; (void)line; (void)tag; (void)dst; (void)ndst; (void)type; UNSUPPORTED("bcf_get_format_values"); return -1;
}
int vcf_format(const bcf_hdr_t *h, const bcf1_t *v, kstring_t *s) { (void)h; (void)v; (void)s; UNSUPPORTED("vcf_format"); return -1; }
bcf_srs_t *bcf_sr_init(void) { UNSUPPORTED("bcf_sr_init"); return NULL; }
void bcf_sr_destroy(bcf_srs_t *readers) { (void)readers; }
int bcf_sr_set_opt(bcf_srs_t *readers, bcf_sr_opt_t opt, ...) { (void)readers; (void)opt; UNSUPPORTED("bcf_sr_set_opt"); return -1; }
int bcf_sr_set_regions(bcf_srs_t *readers, const char *regions, int is_file) { (void)readers; (void)regions; (void)is_file; UNSUPPORTED("bcf_sr_set_regions"); return -1; }
int bcf_sr_set_targets(bcf_srs_t *readers, const char *targets, int is_file, int alleles) { (void)readers; (void)targets; (void)is_file; (void)alleles; UNSUPPORTED("bcf_sr_set_targets"); return -1; }
int bcf_sr_add_reader(bcf_srs_t *readers, const char *fname) { (void)readers; (void)fname; UNSUPPORTED("bcf_sr_add_reader"); return -1; }
int bcf_sr_next_line(bcf_srs_t *readers) { (void)readers; UNSUPPORTED("bcf_sr_next_line"); return 0; }
bcf1_t *bcf_sr_get_line(bcf_srs_t *readers, int i) { (void)readers; (void)i; UNSUPPORTED("bcf_sr_get_line"); return NULL; }

} /* extern "C" */
