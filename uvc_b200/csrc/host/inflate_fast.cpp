// inflate_fast.cpp - raw DEFLATE (RFC 1951) decoder for whole BGZF members.
//
// A BGZF member is an independent deflate stream of at most 64 KiB of output whose compressed bytes are all in memory, so the decoder needs no
// streaming state: a 64-bit bit buffer refilled with one unaligned load, table-driven Huffman decoding (11-bit primary table for
// literal/length codes, 8-bit for distances, second-level tables for longer codes), word-wise match copies. Stock zlib inflates the level-1
// streams that aligners and our generator write at ~170 MB/s per core, which made BAM decode 3/4 inflate; this decoder
// is checked against zlib byte for byte in tests/test_inflate.py, and uvc_inflate_member falls back to zlib whenever it reports an error.
#include "inflate_fast.h"

#include <string.h>
#include <zlib.h>

namespace {

const int LT_BITS = 11, DT_BITS = 8;
const int LT_SIZE = (1 << LT_BITS) + 1024, DT_SIZE = (1 << DT_BITS) + 512;    // primary + room for the second-level tables

// table entry: bits 0-7 code length to consume (second level: the bits beyond the primary index), bits 8-12 extra bits (or second-level index
// bits), bits 16-31 value (literal, base length, base distance, or start of the second-level table); flags in bits 13-15
const uint32_t F_LITERAL = 1u << 13, F_EOB = 1u << 14, F_SUB = 1u << 15;
// literal entries of the primary literal/length table may carry TWO literals (bit 12 set; value = first | second << 8, length = both codes): see pair_literals
const uint32_t F_LIT2 = 1u << 12;

const uint16_t LEN_BASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t LEN_XBITS[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t DIST_BASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t DIST_XBITS[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

inline uint32_t reverse_bits(uint32_t code, int len) {
    uint32_t r = 0;
    for (int i = 0; i < len; i++) { r = (r << 1) | ((code >> i) & 1u); }
    return r;
}

// value part of the entry of symbol `sym`; kind 0 = literal/length alphabet, 1 = distance alphabet. Returns false for symbols that must not occur.
inline bool symbol_entry(int kind, int sym, uint32_t & e) {
    if (0 == kind) {
        if (sym < 256) { e = F_LITERAL | ((uint32_t)sym << 16); return true; }
        if (256 == sym) { e = F_EOB; return true; }
        if (sym > 285) { return false; }
        e = ((uint32_t)LEN_BASE[sym - 257] << 16) | ((uint32_t)LEN_XBITS[sym - 257] << 8);
        return true;
    }
    if (sym > 29) { return false; }
    e = ((uint32_t)DIST_BASE[sym] << 16) | ((uint32_t)DIST_XBITS[sym] << 8);
    return true;
}

// Builds the decode table of a canonical Huffman code. lens[i] = code length of symbol i (0 = unused). Returns false if the code is over-subscribed,
// incomplete (except the single-code case the format allows), uses a reserved symbol, or does not fit the table.
bool build_table(uint32_t *tab, int tab_size, int root_bits, const uint8_t *lens, int n_sym, int kind) {
    int count[16];
    memset(count, 0, sizeof(count));
    for (int i = 0; i < n_sym; i++) { count[lens[i]]++; }
    const int n_used = n_sym - count[0];
    for (int i = 0; i < (1 << root_bits); i++) { tab[i] = 0; }      // 0 = invalid (code length 0)
    if (0 == n_used) { return true; }                               // an empty code: decoding a symbol with it is an error (entry 0)
    int left = 1;
    for (int l = 1; l <= 15; l++) { left = (left << 1) - count[l]; if (left < 0) { return false; } }
    if (left > 0 && !(1 == n_used && 1 == count[1])) { return false; }   // incomplete: only "one symbol of length 1" is allowed
    uint32_t next_code[16];
    uint32_t code = 0;
    count[0] = 0;
    for (int l = 1; l <= 15; l++) { code = (code + (uint32_t)count[l - 1]) << 1; next_code[l] = code; }
    // second-level tables: for every primary index, the longest code that shares it
    uint8_t sub_bits[1 << LT_BITS];
    memset(sub_bits, 0, (size_t)1 << root_bits);
    uint32_t codes[288];
    for (int i = 0; i < n_sym; i++) {
        const int l = lens[i];
        if (0 == l) { continue; }
        codes[i] = reverse_bits(next_code[l]++, l);
        if (l > root_bits) { const uint32_t idx = codes[i] & ((1u << root_bits) - 1); if (l - root_bits > sub_bits[idx]) { sub_bits[idx] = (uint8_t)(l - root_bits); } }
    }
    int next_free = 1 << root_bits;
    for (int idx = 0; idx < (1 << root_bits); idx++) {
        if (0 == sub_bits[idx]) { continue; }
        const int n = 1 << sub_bits[idx];
        if (next_free + n > tab_size) { return false; }
        tab[idx] = F_SUB | ((uint32_t)next_free << 16) | ((uint32_t)sub_bits[idx] << 8) | (uint32_t)root_bits;
        for (int k = 0; k < n; k++) { tab[next_free + k] = 0; }
        next_free += n;
    }
    for (int i = 0; i < n_sym; i++) {
        const int l = lens[i];
        if (0 == l) { continue; }
        uint32_t e;
        if (!symbol_entry(kind, i, e)) { return false; }
        if (l <= root_bits) {
            e |= (uint32_t)l;
            for (uint32_t k = codes[i]; k < (1u << root_bits); k += (1u << l)) { tab[k] = e; }
        } else {
            const uint32_t idx = codes[i] & ((1u << root_bits) - 1);
            const uint32_t start = tab[idx] >> 16;
            const int sb = sub_bits[idx], rest = l - root_bits;
            e |= (uint32_t)rest;
            for (uint32_t k = codes[i] >> root_bits; k < (1u << sb); k += (1u << rest)) { tab[start + k] = e; }
        }
    }
    return true;
}

// Quality strings and 4-bit packed bases are literal-heavy: most of a BAM's deflate symbols are literals with short codes. Wherever the bits of
// a primary index that follow a literal's code decide a second literal completely, the entry is replaced by one that yields both.
void pair_literals(uint32_t *tab) {
    uint32_t one[1 << LT_BITS];
    memcpy(one, tab, sizeof(one));
    for (uint32_t i = 0; i < (1u << LT_BITS); i++) {
        const uint32_t e = one[i];
        if (!(e & F_LITERAL) || (e & F_SUB)) { continue; }
        const int l1 = (int)(e & 0xff), rem = LT_BITS - l1;
        if (rem < 1) { continue; }
        const uint32_t e2 = one[i >> l1];               // (the missing high bits read as zeros: fine for a code of at most `rem` bits)
        if (!(e2 & F_LITERAL) || (e2 & F_SUB) || (int)(e2 & 0xff) > rem) { continue; }
        tab[i] = F_LITERAL | F_LIT2 | (((e >> 16) & 0xffu) << 16) | (((e2 >> 16) & 0xffu) << 24) | (uint32_t)(l1 + (int)(e2 & 0xff));
    }
}

struct FixedTables {      // the fixed code of block type 1 (RFC 1951, 3.2.6): both codes contain reserved symbols, so they are filled in directly
    uint32_t lt[LT_SIZE], dt[DT_SIZE];
    FixedTables() {
        uint8_t lens[288];
        for (int i = 0; i < 144; i++) { lens[i] = 8; }
        for (int i = 144; i < 256; i++) { lens[i] = 9; }
        for (int i = 256; i < 280; i++) { lens[i] = 7; }
        for (int i = 280; i < 288; i++) { lens[i] = 8; }
        int count[16];
        memset(count, 0, sizeof(count));
        for (int i = 0; i < 288; i++) { count[lens[i]]++; }
        uint32_t code = 0, next_code[16];
        count[0] = 0;
        for (int l = 1; l <= 15; l++) { code = (code + (uint32_t)count[l - 1]) << 1; next_code[l] = code; }
        for (int i = 0; i < (1 << LT_BITS); i++) { lt[i] = 0; }
        for (int i = 0; i < 288; i++) {
            const int l = lens[i];
            const uint32_t c = reverse_bits(next_code[l]++, l);
            uint32_t e;
            if (!symbol_entry(0, i, e)) { continue; }      // 286 and 287 take part in the code but must not occur: their entries stay invalid
            e |= (uint32_t)l;
            for (uint32_t k = c; k < (1u << LT_BITS); k += (1u << l)) { lt[k] = e; }
        }
        for (int i = 0; i < (1 << DT_BITS); i++) { dt[i] = 0; }
        for (int i = 0; i < 30; i++) {                      // 32 codes of 5 bits, 30 of them used
            uint32_t e;
            symbol_entry(1, i, e);
            e |= 5u;
            for (uint32_t k = reverse_bits((uint32_t)i, 5); k < (1u << DT_BITS); k += 32) { dt[k] = e; }
        }
        pair_literals(lt);
    }
};
const FixedTables & fixed_tables() { static const FixedTables t; return t; }

inline uint64_t load64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }     // (little-endian hosts: x86-64, aarch64)

} // namespace

// The caller guarantees 16 readable bytes after in[in_len - 1] (their values do not matter).
int64_t uvc_inflate_raw(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_cap) {
    const uint8_t *ip = in, *const in_end = in + in_len;
    uint8_t *op = out, *const out_end = out + out_cap;
    uint64_t bb = 0;
    int bc = 0;
    uint32_t lt_dyn[LT_SIZE], dt_dyn[DT_SIZE];
    #define REFILL() { bb |= load64(ip) << bc; ip += (63 - bc) >> 3; bc |= 56; }
    #define DROP(n) { bb >>= (n); bc -= (n); }
    for (;;) {
        if (ip > in_end + 8) { return -1; }
        REFILL();
        const int last = (int)(bb & 1), type = (int)((bb >> 1) & 3);
        DROP(3);
        const uint32_t *lt, *dt;
        if (0 == type) {
            // stored: back to the byte boundary (the bit buffer holds whole bytes that were read ahead)
            DROP(bc & 7);
            const uint8_t *p = ip - (bc >> 3);
            if (p + 4 > in_end) { return -1; }
            const uint32_t len = p[0] | ((uint32_t)p[1] << 8), nlen = p[2] | ((uint32_t)p[3] << 8);
            if ((len ^ 0xffffu) != nlen) { return -1; }
            p += 4;
            if (len > (size_t)(in_end - p) || len > (size_t)(out_end - op)) { return -1; }
            memcpy(op, p, len);
            op += len; ip = p + len; bb = 0; bc = 0;
            if (last) { break; }
            continue;
        } else if (1 == type) {
            lt = fixed_tables().lt; dt = fixed_tables().dt;
        } else if (2 == type) {
            const int hlit = (int)(bb & 31) + 257, hdist = (int)((bb >> 5) & 31) + 1, hclen = (int)((bb >> 10) & 15) + 4;
            DROP(14);
            if (hlit > 286 || hdist > 30) { return -1; }
            static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
            uint8_t cl[19];
            memset(cl, 0, sizeof(cl));
            for (int i = 0; i < hclen; i++) {
                if (bc < 3) { REFILL(); }
                cl[order[i]] = (uint8_t)(bb & 7); DROP(3);
            }
            // the code-length code: at most 7 bits, one flat table
            uint32_t ct[128];
            {
                int count[8]; memset(count, 0, sizeof(count));
                for (int i = 0; i < 19; i++) { count[cl[i]]++; }
                int left = 1, n_used = 19 - count[0];
                for (int l = 1; l <= 7; l++) { left = (left << 1) - count[l]; if (left < 0) { return -1; } }
                if (0 == n_used || (left > 0 && !(1 == n_used && 1 == count[1]))) { return -1; }
                uint32_t code = 0, next_code[8];
                count[0] = 0;
                for (int l = 1; l <= 7; l++) { code = (code + (uint32_t)count[l - 1]) << 1; next_code[l] = code; }
                for (int i = 0; i < 128; i++) { ct[i] = 0; }
                for (int i = 0; i < 19; i++) {
                    const int l = cl[i];
                    if (0 == l) { continue; }
                    for (uint32_t k = reverse_bits(next_code[l]++, l); k < 128; k += (1u << l)) { ct[k] = ((uint32_t)i << 8) | (uint32_t)l; }
                }
            }
            uint8_t lens[286 + 30 + 140];
            int n = 0;
            while (n < hlit + hdist) {
                if (ip > in_end + 8) { return -1; }
                REFILL();
                const uint32_t e = ct[bb & 127];
                const int l = (int)(e & 0xff), sym = (int)(e >> 8);
                if (0 == l) { return -1; }
                DROP(l);
                if (sym < 16) { lens[n++] = (uint8_t)sym; continue; }
                int rep; uint8_t val = 0;
                if (16 == sym) { if (0 == n) { return -1; } val = lens[n - 1]; rep = 3 + (int)(bb & 3); DROP(2); }
                else if (17 == sym) { rep = 3 + (int)(bb & 7); DROP(3); }
                else { rep = 11 + (int)(bb & 127); DROP(7); }
                if (n + rep > hlit + hdist) { return -1; }
                memset(lens + n, val, (size_t)rep);
                n += rep;
            }
            if (0 == lens[256]) { return -1; }      // no end-of-block code
            if (!build_table(lt_dyn, LT_SIZE, LT_BITS, lens, hlit, 0)) { return -1; }
            pair_literals(lt_dyn);
            if (!build_table(dt_dyn, DT_SIZE, DT_BITS, lens + hlit, hdist, 1)) { return -1; }
            lt = lt_dyn; dt = dt_dyn;
        } else { return -1; }
        // ---- the symbols of the block
        bool eob = false;
        while (!eob) {
            // far from both ends: a symbol reads at most 8 bytes beyond ip and writes at most 258 + 8 bytes beyond op, so nothing is checked but the
            // match distance
            while (ip + 16 <= in_end && (size_t)(out_end - op) >= 274) {
                REFILL();
                uint32_t e = lt[bb & ((1u << LT_BITS) - 1)];
                if (e & F_LITERAL) {
                    // up to three table entries (one or two literals each) per refill: at least 56 - 3 * 15 bits remain for the next lookup
                    op[0] = (uint8_t)(e >> 16); op[1] = (uint8_t)(e >> 24); op += 1 + ((e >> 12) & 1u); DROP((int)(e & 0xff));
                    e = lt[bb & ((1u << LT_BITS) - 1)];
                    if (e & F_LITERAL) {
                        op[0] = (uint8_t)(e >> 16); op[1] = (uint8_t)(e >> 24); op += 1 + ((e >> 12) & 1u); DROP((int)(e & 0xff));
                        e = lt[bb & ((1u << LT_BITS) - 1)];
                        if (e & F_LITERAL) { op[0] = (uint8_t)(e >> 16); op[1] = (uint8_t)(e >> 24); op += 1 + ((e >> 12) & 1u); DROP((int)(e & 0xff)); }
                    }
                    continue;
                }
                if (e & F_SUB) { e = lt[(e >> 16) + ((bb >> LT_BITS) & ((1u << ((e >> 8) & 31)) - 1))]; DROP(LT_BITS); }
                int l = (int)(e & 0xff);
                if (0 == l) { return -1; }
                DROP(l);
                if (e & F_LITERAL) { *op++ = (uint8_t)(e >> 16); continue; }      // (a literal with a long code, from a second-level table)
                if (e & F_EOB) { eob = true; break; }
                const int xb = (int)((e >> 8) & 31);
                const uint32_t len = (e >> 16) + (uint32_t)(bb & ((1u << xb) - 1));
                DROP(xb);
                uint32_t d = dt[bb & ((1u << DT_BITS) - 1)];
                if (d & F_SUB) { d = dt[(d >> 16) + ((bb >> DT_BITS) & ((1u << ((d >> 8) & 31)) - 1))]; DROP(DT_BITS); }
                l = (int)(d & 0xff);
                if (0 == l) { return -1; }
                DROP(l);
                const int dxb = (int)((d >> 8) & 31);
                const uint32_t dist = (d >> 16) + (uint32_t)(bb & ((1u << dxb) - 1));
                DROP(dxb);
                if (dist > (size_t)(op - out)) { return -1; }
                const uint8_t *src = op - dist;
                uint8_t *dst = op;
                const uint8_t *const stop = op + len;
                if (dist >= 8) { do { memcpy(dst, src, 8); dst += 8; src += 8; } while (dst < stop); }
                else if (1 == dist) { memset(op, *src, len); }
                else { do { *dst++ = *src++; } while (dst < stop); }
                op += len;
            }
            if (eob) { break; }
            // near an end: one symbol with every check
            if (ip > in_end + 8) { return -1; }
            REFILL();
            uint32_t e = lt[bb & ((1u << LT_BITS) - 1)];
            if (e & F_SUB) { e = lt[(e >> 16) + ((bb >> LT_BITS) & ((1u << ((e >> 8) & 31)) - 1))]; DROP(LT_BITS); }
            int l = (int)(e & 0xff);
            if (0 == l) { return -1; }
            DROP(l);
            if (e & F_LITERAL) {
                const size_t n_lit = 1 + ((e >> 12) & 1u);
                if ((size_t)(out_end - op) < n_lit) { return -1; }
                op[0] = (uint8_t)(e >> 16);
                if (2 == n_lit) { op[1] = (uint8_t)(e >> 24); }
                op += n_lit;
                continue;
            }
            if (e & F_EOB) { break; }
            const int xb = (int)((e >> 8) & 31);
            const uint32_t len = (e >> 16) + (uint32_t)(bb & ((1u << xb) - 1));
            DROP(xb);
            // (at most 15 + 5 bits used since the refill: 36 are left, a distance needs at most 15 + 13)
            uint32_t d = dt[bb & ((1u << DT_BITS) - 1)];
            if (d & F_SUB) { d = dt[(d >> 16) + ((bb >> DT_BITS) & ((1u << ((d >> 8) & 31)) - 1))]; DROP(DT_BITS); }
            l = (int)(d & 0xff);
            if (0 == l) { return -1; }
            DROP(l);
            const int dxb = (int)((d >> 8) & 31);
            const uint32_t dist = (d >> 16) + (uint32_t)(bb & ((1u << dxb) - 1));
            DROP(dxb);
            if (dist > (size_t)(op - out) || len > (size_t)(out_end - op)) { return -1; }
            const uint8_t *src = op - dist;
            for (uint32_t k = 0; k < len; k++) { op[k] = src[k]; }
            op += len;
        }
        if (last) { break; }
    }
    #undef REFILL
    #undef DROP
    // bytes actually consumed: what was loaded minus the whole bytes still in the bit buffer
    if ((ip - (bc >> 3)) > in_end) { return -1; }
    return (int64_t)(op - out);
}

int64_t uvc_inflate_member(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_cap) {
    const int64_t n = uvc_inflate_raw(in, in_len, out, out_cap);
    if (n >= 0) { return n; }
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (inflateInit2(&zs, -15) != Z_OK) { return -1; }
    zs.next_in = (Bytef*)in; zs.avail_in = (uInt)in_len; zs.next_out = out; zs.avail_out = (uInt)out_cap;
    const int rc = inflate(&zs, Z_FINISH);
    const int64_t total = (int64_t)zs.total_out;
    inflateEnd(&zs);
    return (rc == Z_STREAM_END ? total : -1);
}
