// synthgen.cpp - fast seeded generator of the benchmark inputs: reference FASTA (+.fai), coordinate-sorted
// BAM (+.bai) of simulated 2x150 paired-end reads with spiked SNVs/indels at known allele fractions, optional duplex UMIs in the read
// name (QNAME#AGTA+TGGT), target BED and truth table.
//
// It implements the same statistical model as uvc_b200/synth.py (which stays the generator of the small parity-test inputs and of the
// committed golden fixtures), multi-threaded and streaming, so that the BASELINE.json configurations can be produced at their full size
// on the benchmark box within seconds (tens of millions of reads per minute instead of 0.1 M reads/s). Everything is a pure function of
// the command line (seed included): the random stream of a fragment depends only on (seed, region, index), never on the thread count.
//
// BENCH/TEST INFRASTRUCTURE: not part of the product path.
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {

const int READ_LEN = 150;
const int BGZF_MAX = 0xff00;

struct Rng {      // xoshiro256++ seeded through splitmix64
    uint64_t s[4];
    static uint64_t splitmix(uint64_t & x) { uint64_t z = (x += 0x9E3779B97F4A7C15ULL); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; return z ^ (z >> 31); }
    Rng(uint64_t a, uint64_t b = 0, uint64_t c = 0) {
        uint64_t x = a * 0xD1342543DE82EF95ULL + b * 0x2545F4914F6CDD1DULL + c * 0x9E3779B97F4A7C15ULL + 0x1234567ULL;
        for (int i = 0; i < 4; i++) { s[i] = splitmix(x); }
    }
    static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    inline uint64_t next() {
        const uint64_t r = rotl(s[0] + s[3], 23) + s[0];
        const uint64_t t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
        return r;
    }
    inline double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    inline int64_t below(int64_t n) { return (n <= 1 ? 0 : (int64_t)(((unsigned __int128)next() * (unsigned __int128)(uint64_t)n) >> 64)); }
    inline int64_t range(int64_t lo, int64_t hi) { return lo + below(hi - lo); }   // [lo, hi)
    double normal() { double u1 = uniform(), u2 = uniform(); if (u1 < 1e-300) { u1 = 1e-300; } return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2); }
    int64_t geometric(double p) { double u = uniform(); if (u < 1e-300) { u = 1e-300; } const int64_t k = (int64_t)ceil(log(u) / log1p(-p)); return (k < 1 ? 1 : k); }
};

struct Contig { std::string name; int64_t len; std::string seq; };
struct Region { int32_t ci; int64_t beg, end; bool is_target; };
struct Variant { int32_t ci; int64_t pos; int kind; std::string ref, alt; double vaf; };   // kind 0 snv, 1 ins, 2 del

struct Config {
    std::string name = "c1", outdir = ".";
    uint64_t seed = 1001;
    std::vector<Contig> contigs;
    double depth = 100;
    int n_snv = 200, n_indel = 60, max_indel_len = 30;
    std::vector<double> vafs = {0.05, 0.10, 0.25, 0.50};
    // targets: n_targets regular targets (first, step, len) on contig 0; 0 = whole contigs
    int64_t n_targets = 0, target_first = 1000, target_step = 1000, target_len = 200;
    double amplicon_frac = 0;
    bool umi = false; int umi_len = 4; double family_mean = 8, duplex_frac = 0.6, swapped_umi_frac = 0.5, pcr_err_rate = 1e-3;
    double sub_err = 5e-4, indel_err = 1e-5, clip_frac = 0.01, lowmapq_frac = 0.02;
    int64_t str_every = 2000;
    double insert_mean = 350, insert_sd = 50; int64_t insert_min = 160, insert_max = 1000;
    int threads = 0, level = 1;
};

const char BASES[4] = {'A', 'C', 'G', 'T'};
inline int base_code(char c) { return (c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : 3); }
const uint8_t NT16[4] = {1, 2, 4, 8};

void make_reference(Config & cfg) {
    for (size_t ci = 0; ci < cfg.contigs.size(); ci++) {
        Contig & c = cfg.contigs[ci];
        Rng rng(cfg.seed, 100 + ci);
        c.seq.resize((size_t)c.len);
        for (int64_t i = 0; i < c.len; i += 32) {
            uint64_t r = rng.next();
            for (int64_t j = i; j < i + 32 && j < c.len; j++, r >>= 2) { c.seq[(size_t)j] = BASES[r & 3]; }
        }
        int64_t pos = rng.range(200, cfg.str_every);
        while (pos + 80 < c.len) {
            const int kind = (int)rng.below(3);
            if (kind == 0) {
                const int n = (int)rng.range(6, 21);
                const char b = BASES[rng.below(4)];
                for (int k = 0; k < n; k++) { c.seq[(size_t)(pos + k)] = b; }
            } else {
                const int ulen = kind + 1;
                char unit[3];
                for (int k = 0; k < ulen; k++) { unit[k] = BASES[rng.below(4)]; }
                bool same = true;
                for (int k = 1; k < ulen; k++) { if (unit[k] != unit[0]) { same = false; } }
                if (same) { unit[ulen - 1] = BASES[(base_code(unit[0]) + 1) % 4]; }
                const int n = (int)rng.range(4, 13);
                for (int k = 0; k < n * ulen; k++) { c.seq[(size_t)(pos + k)] = unit[k % ulen]; }
            }
            pos += rng.range(cfg.str_every / 2, cfg.str_every * 3 / 2);
        }
    }
}

void write_fasta(const Config & cfg, const std::string & path) {
    FILE *fa = fopen(path.c_str(), "wb"), *fai = fopen((path + ".fai").c_str(), "w");
    if (!fa || !fai) { perror("fasta"); exit(2); }
    int64_t off = 0;
    const int width = 60;
    std::string buf;
    for (const Contig & c : cfg.contigs) {
        off += fprintf(fa, ">%s\n", c.name.c_str());
        fprintf(fai, "%s\t%lld\t%lld\t%d\t%d\n", c.name.c_str(), (long long)c.len, (long long)off, width, width + 1);
        buf.clear();
        buf.reserve((size_t)(c.len + c.len / width + 2));
        for (int64_t i = 0; i < c.len; i += width) {
            const int64_t n = std::min<int64_t>(width, c.len - i);
            buf.append(c.seq, (size_t)i, (size_t)n);
            buf.push_back('\n');
        }
        fwrite(buf.data(), 1, buf.size(), fa);
        off += (int64_t)buf.size();
    }
    fclose(fa); fclose(fai);
}

// regions = units of generation: the targets, or blocks of the contigs sized for ~100 k reads
std::vector<Region> make_regions(const Config & cfg) {
    std::vector<Region> out;
    if (cfg.n_targets > 0) {
        for (int64_t i = 0; i < cfg.n_targets; i++) { out.push_back(Region{0, cfg.target_first + cfg.target_step * i, cfg.target_first + cfg.target_step * i + cfg.target_len, true}); }
        return out;
    }
    int64_t blk = (int64_t)(100000.0 * READ_LEN / cfg.depth);
    blk = std::max<int64_t>(2000, std::min<int64_t>(100000, blk));
    for (size_t ci = 0; ci < cfg.contigs.size(); ci++) {
        for (int64_t b = 0; b < cfg.contigs[ci].len; b += blk) { out.push_back(Region{(int32_t)ci, b, std::min(cfg.contigs[ci].len, b + blk), false}); }
    }
    return out;
}

std::vector<Variant> make_variants(const Config & cfg, const std::vector<Region> & regions) {
    Rng rng(cfg.seed, 200);
    std::vector<Variant> out;
    std::vector<std::vector<int64_t>> taken(cfg.contigs.size());
    const int n_total = cfg.n_snv + cfg.n_indel;
    std::vector<double> cum(regions.size());
    double tot = 0;
    for (size_t i = 0; i < regions.size(); i++) { tot += (double)std::max<int64_t>(1, regions[i].end - regions[i].beg); cum[i] = tot; }
    for (int64_t tries = 0; (int)out.size() < n_total && tries < (int64_t)n_total * 50; tries++) {
        const bool snv = ((int)out.size() < cfg.n_snv);
        const size_t ri = (size_t)(std::lower_bound(cum.begin(), cum.end(), rng.uniform() * tot) - cum.begin());
        const Region & R = regions[std::min(ri, regions.size() - 1)];
        const Contig & C = cfg.contigs[(size_t)R.ci];
        const int64_t lo = std::max<int64_t>(R.beg + 5, 300), hi = std::min<int64_t>(R.end - 5, C.len - 300);
        if (hi <= lo) { continue; }
        const int64_t pos = rng.range(lo, hi);
        bool clash = false;
        for (int64_t p : taken[(size_t)R.ci]) { if (std::llabs(p - pos) < 120) { clash = true; break; } }
        if (clash) { continue; }
        Variant v;
        v.ci = R.ci; v.pos = pos; v.vaf = cfg.vafs[(size_t)rng.below((int64_t)cfg.vafs.size())];
        const char r = C.seq[(size_t)pos];
        if (snv) {
            v.kind = 0; v.ref = std::string(1, r);
            v.alt = std::string(1, BASES[(base_code(r) + 1 + rng.below(3)) % 4]);
        } else {
            const int ilen = (int)std::min<int64_t>(cfg.max_indel_len, 1 + rng.geometric(0.25));
            if (rng.uniform() < 0.5) {
                v.kind = 1; v.ref = std::string(1, r); v.alt = v.ref;
                for (int k = 0; k < ilen; k++) { v.alt.push_back(BASES[rng.below(4)]); }
            } else {
                v.kind = 2; v.ref = C.seq.substr((size_t)pos, (size_t)(1 + ilen)); v.alt = std::string(1, r);
            }
        }
        taken[(size_t)R.ci].push_back(pos);
        out.push_back(v);
    }
    std::sort(out.begin(), out.end(), [](const Variant & a, const Variant & b) { return a.ci != b.ci ? a.ci < b.ci : a.pos < b.pos; });
    return out;
}

inline uint32_t reg2bin(int64_t beg, int64_t end) {
    --end;
    if (beg >> 14 == end >> 14) { return (uint32_t)(((1 << 15) - 1) / 7 + (beg >> 14)); }
    if (beg >> 17 == end >> 17) { return (uint32_t)(((1 << 12) - 1) / 7 + (beg >> 17)); }
    if (beg >> 20 == end >> 20) { return (uint32_t)(((1 << 9) - 1) / 7 + (beg >> 20)); }
    if (beg >> 23 == end >> 23) { return (uint32_t)(((1 << 6) - 1) / 7 + (beg >> 23)); }
    if (beg >> 26 == end >> 26) { return (uint32_t)(((1 << 3) - 1) / 7 + (beg >> 26)); }
    return 0;
}

// One haplotype segment of a fragment: M = `len` bases at refpos (from the reference, or one substituted base), I = inserted bases, D = deleted reference bases
struct Seg { char op; int64_t refpos; int32_t len; const char *bases; char sub; };

struct ReadOut {
    int64_t pos = 0, rend = 0;
    std::vector<uint32_t> cigar;
    std::string seq, qual;
    int nm = 0, mapq = 60;
};

struct RecRef { int32_t ci; int32_t pos; int32_t rend; uint32_t off, len; uint64_t order; };   // a finished BAM record inside a worker's arena

struct Arena {
    std::vector<uint8_t> bytes;
    std::vector<RecRef> recs;
};

inline void put32(std::vector<uint8_t> & v, uint32_t x) { const size_t n = v.size(); v.resize(n + 4); memcpy(&v[n], &x, 4); }
inline void put16(std::vector<uint8_t> & v, uint16_t x) { const size_t n = v.size(); v.resize(n + 2); memcpy(&v[n], &x, 2); }

struct Generator {
    const Config & cfg;
    const std::vector<Region> & regions;
    const std::vector<Variant> & variants;
    std::vector<std::pair<size_t, size_t>> var_range;   // per contig: [first, last) in variants

    Generator(const Config & c, const std::vector<Region> & r, const std::vector<Variant> & v) : cfg(c), regions(r), variants(v) {
        var_range.assign(cfg.contigs.size(), std::make_pair((size_t)0, (size_t)0));
        for (size_t ci = 0; ci < cfg.contigs.size(); ci++) {
            size_t a = 0;
            while (a < v.size() && v[a].ci < (int32_t)ci) { a++; }
            size_t b = a;
            while (b < v.size() && v[b].ci == (int32_t)ci) { b++; }
            var_range[ci] = std::make_pair(a, b);
        }
    }

    void sample_quals(Rng & rng, std::string & q, size_t n) const {
        q.resize(n);
        for (size_t i = 0; i < n; i++) {
            const uint64_t r = rng.next();
            const uint32_t u = (uint32_t)(r & 0xffff);
            if (u < 52429) { q[i] = 37; }                                       // 80 %
            else if (u < 62259) { q[i] = (char)(30 + ((r >> 16) & 0xffff) * 7 / 65536); }    // 15 %: 30..36
            else { q[i] = (char)(2 + ((r >> 16) & 0xffff) * 24 / 65536); }      // 5 %: 2..25
        }
    }

    // takes the first (forward) or last (reverse) READ_LEN query bases of the haplotype
    bool read_from_segs(const std::vector<Seg> & segs, bool forward, const std::string & ref, ReadOut & out) const {
        struct Piece { char op; int64_t refpos; int32_t len; const char *bases; char sub; };
        std::vector<Piece> taken;
        int need = READ_LEN;
        const int n = (int)segs.size();
        for (int k = 0; k < n && need > 0; k++) {
            const Seg & s = segs[(size_t)(forward ? k : n - 1 - k)];
            if (s.op == 'D') { if (!taken.empty()) { taken.push_back(Piece{'D', s.refpos, s.len, NULL, 0}); } continue; }
            if (s.len <= need) { taken.push_back(Piece{s.op, s.refpos, s.len, s.bases, s.sub}); need -= s.len; }
            else {
                if (forward) { taken.push_back(Piece{s.op, s.refpos, need, s.bases, s.sub}); }
                else { taken.push_back(Piece{s.op, (s.op == 'M' ? s.refpos + (s.len - need) : s.refpos), need, (s.bases ? s.bases + (s.len - need) : NULL), s.sub}); }
                need = 0;
            }
        }
        while (!taken.empty() && taken.back().op != 'M') { taken.pop_back(); }
        if (!forward) { std::reverse(taken.begin(), taken.end()); }
        size_t a = 0;
        while (a < taken.size() && taken[a].op != 'M') { a++; }
        if (a >= taken.size()) { return false; }
        out.pos = taken[a].refpos;
        out.cigar.clear(); out.seq.clear();
        int64_t rlen = 0;
        for (size_t k = a; k < taken.size(); k++) {
            const Piece & p = taken[k];
            const uint32_t code = (p.op == 'M' ? 0u : p.op == 'I' ? 1u : 2u);
            if (!out.cigar.empty() && (out.cigar.back() & 0xf) == code) { out.cigar.back() += ((uint32_t)p.len << 4); }
            else { out.cigar.push_back(((uint32_t)p.len << 4) | code); }
            if (p.op == 'M') {
                if (p.sub) { out.seq.push_back(p.sub); } else { out.seq.append(ref, (size_t)p.refpos, (size_t)p.len); }
                rlen += p.len;
            } else if (p.op == 'I') { out.seq.append(p.bases, (size_t)p.len); }
            else { rlen += p.len; }
        }
        out.rend = out.pos + rlen;
        return true;
    }

    static int compute_nm(const std::string & ref, const ReadOut & r) {
        int nm = 0; int64_t q = 0, p = r.pos;
        for (uint32_t c : r.cigar) {
            const int op = (int)(c & 0xf), l = (int)(c >> 4);
            if (op == 0) { for (int k = 0; k < l; k++) { nm += (r.seq[(size_t)(q + k)] != ref[(size_t)(p + k)]); } q += l; p += l; }
            else if (op == 1) { nm += l; q += l; }
            else if (op == 2) { nm += l; p += l; }
            else if (op == 4) { q += l; }
        }
        return nm;
    }

    void emit(Arena & A, int32_t ci, const ReadOut & r, const std::string & qname, uint16_t flag, int64_t mpos, int64_t isize, uint64_t order) const {
        const size_t start = A.bytes.size();
        const uint32_t l_qname = (uint32_t)qname.size() + 1, l_seq = (uint32_t)r.seq.size(), n_cig = (uint32_t)r.cigar.size();
        const uint32_t block = 32 + l_qname + 4 * n_cig + (l_seq + 1) / 2 + l_seq + 4;
        std::vector<uint8_t> & b = A.bytes;
        if (b.capacity() < start + block + 4) { b.reserve(std::max<size_t>((size_t)1 << 20, b.capacity() * 2)); }
        put32(b, block); put32(b, (uint32_t)ci); put32(b, (uint32_t)r.pos);
        b.push_back((uint8_t)l_qname); b.push_back((uint8_t)r.mapq); put16(b, (uint16_t)reg2bin(r.pos, r.rend));
        put16(b, (uint16_t)n_cig); put16(b, flag); put32(b, l_seq); put32(b, (uint32_t)ci); put32(b, (uint32_t)mpos); put32(b, (uint32_t)(int32_t)isize);
        b.insert(b.end(), qname.begin(), qname.end()); b.push_back(0);
        for (uint32_t c : r.cigar) { put32(b, c); }
        for (uint32_t i = 0; i < l_seq; i += 2) {
            const uint8_t hi = NT16[base_code(r.seq[i])], lo = (i + 1 < l_seq ? NT16[base_code(r.seq[i + 1])] : 0);
            b.push_back((uint8_t)((hi << 4) | lo));
        }
        b.insert(b.end(), r.qual.begin(), r.qual.end());
        b.push_back('N'); b.push_back('M'); b.push_back('C'); b.push_back((uint8_t)std::min(255, r.nm));
        A.recs.push_back(RecRef{ci, (int32_t)r.pos, (int32_t)r.rend, (uint32_t)start, (uint32_t)(b.size() - start), order});
    }

    // all fragments (and their two reads each) of one region
    void generate_region(size_t ri, Arena & A) const {
        const Region & R = regions[ri];
        const Contig & C = cfg.contigs[(size_t)R.ci];
        const std::string & ref = C.seq;
        Rng rng(cfg.seed, 300, ri);
        const double n_reads = cfg.depth * (double)(R.end - R.beg) / READ_LEN;
        int64_t n_mol = (int64_t)llround(n_reads / 2.0 / (cfg.umi ? cfg.family_mean * (1.0 + cfg.duplex_frac) : 1.0));
        if (n_mol < 1) { n_mol = 1; }
        const bool amplicon = (R.is_target && rng.uniform() < cfg.amplicon_frac);
        const size_t v0 = var_range[(size_t)R.ci].first, v1 = var_range[(size_t)R.ci].second;
        std::vector<Seg> segs;
        std::vector<const Variant*> carried;
        ReadOut rd[2];
        std::string qname;
        uint64_t frag_local = 0;
        for (int64_t m = 0; m < n_mol; m++) {
            int64_t ins = (int64_t)llround(cfg.insert_mean + cfg.insert_sd * rng.normal());
            ins = std::max(cfg.insert_min, std::min(cfg.insert_max, ins));
            int64_t st, en;
            if (amplicon) {
                st = std::max<int64_t>(0, R.beg - 20); en = std::min<int64_t>(C.len, R.end + 20); en = std::max(en, st + cfg.insert_min);
            } else if (!R.is_target) {
                st = rng.range(R.beg, R.end);
                if (st + ins > C.len) { st = std::max<int64_t>(0, C.len - ins); }
                en = st + ins;
            } else {
                st = rng.range(R.beg - ins + 30, R.end - 30);
                st = std::max<int64_t>(0, std::min(st, C.len - ins));
                en = st + ins;
            }
            en = std::min(en, C.len);
            const int top = (int)(rng.next() & 1);
            int64_t umi_a = 0, umi_b = 0; int64_t fam_top = 1, fam_bot = 0; bool swapped_m = false;
            bool pcr_err = false; int64_t pcr_pos = 0; char pcr_alt = 'A';
            if (cfg.umi) {
                umi_a = rng.below((int64_t)1 << (2 * cfg.umi_len)); umi_b = rng.below((int64_t)1 << (2 * cfg.umi_len));
                const double p = 1.0 / cfg.family_mean;
                fam_top = rng.geometric(p);
                fam_bot = (rng.uniform() < cfg.duplex_frac ? rng.geometric(p) : 0);
                swapped_m = (rng.uniform() < cfg.swapped_umi_frac);
                pcr_err = (rng.uniform() < cfg.pcr_err_rate * (double)(en - st));
                pcr_pos = st + rng.below(std::max<int64_t>(1, en - st));
                pcr_alt = BASES[rng.below(4)];
            }
            // variants this molecule carries
            carried.clear();
            bool has_indel = false;
            {
                size_t lo = v0;
                {   // first variant with pos >= st - 40 (binary search)
                    size_t a = v0, b = v1;
                    while (a < b) { const size_t mid = (a + b) / 2; if (variants[mid].pos < st - 40) { a = mid + 1; } else { b = mid; } }
                    lo = a;
                }
                for (size_t vi = lo; vi < v1 && variants[vi].pos < en; vi++) {
                    const Variant & v = variants[vi];
                    const int64_t span = (v.kind == 2 ? (int64_t)v.ref.size() : 1);
                    if (!(st <= v.pos + span && en > v.pos - 1)) { continue; }
                    Rng vr(cfg.seed, 1000 + vi, ri * 0x100000ULL + (uint64_t)m);
                    if (vr.uniform() < v.vaf) { carried.push_back(&v); if (v.kind != 0) { has_indel = true; } }
                }
            }
            const int64_t n_copies = fam_top + fam_bot;
            for (int64_t cp = 0; cp < n_copies; cp++, frag_local++) {
                const bool is_bottom = (cp >= fam_top);
                const int ftop = (is_bottom ? 1 - top : top);
                const bool swapped = (swapped_m && is_bottom);
                const bool with_pcr = (pcr_err && (rng.next() & 1));
                // haplotype segments of [st, en)
                segs.clear();
                {
                    int64_t p = st;
                    auto match_to = [&](int64_t upto) { if (upto > p) { segs.push_back(Seg{'M', p, (int32_t)(upto - p), NULL, 0}); p = upto; } };
                    bool pcr_done = !with_pcr;
                    auto maybe_pcr = [&](int64_t upto) {   // the PCR substitution, if it lies in the stretch [p, upto)
                        if (!pcr_done && pcr_pos >= p && pcr_pos < upto) { match_to(pcr_pos); segs.push_back(Seg{'M', pcr_pos, 1, NULL, pcr_alt}); p = pcr_pos + 1; pcr_done = true; }
                    };
                    for (const Variant *v : carried) {
                        if (v->kind == 0) {
                            if (v->pos >= p && v->pos < en) { maybe_pcr(v->pos); match_to(v->pos); segs.push_back(Seg{'M', v->pos, 1, NULL, v->alt[0]}); p = v->pos + 1; }
                        } else if (v->kind == 1) {
                            if (st <= v->pos && v->pos + 1 < en && v->pos + 1 >= p) { maybe_pcr(v->pos + 1); match_to(v->pos + 1); segs.push_back(Seg{'I', v->pos + 1, (int32_t)v->alt.size() - 1, v->alt.data() + 1, 0}); }
                        } else {
                            const int64_t dlen = (int64_t)v->ref.size() - 1;
                            if (st <= v->pos && v->pos + 1 + dlen < en && v->pos + 1 >= p) { maybe_pcr(v->pos + 1); match_to(v->pos + 1); segs.push_back(Seg{'D', v->pos + 1, (int32_t)dlen, NULL, 0}); p = v->pos + 1 + dlen; }
                        }
                    }
                    maybe_pcr(en);
                    match_to(en);
                }
                bool ok = true;
                for (int r2 = 0; r2 < 2; r2++) {
                    const bool forward = ((ftop ^ r2) != 0);
                    ReadOut & o = rd[r2];
                    if (!read_from_segs(segs, forward, ref, o)) { ok = false; break; }
                    sample_quals(rng, o.qual, o.seq.size());
                    const uint64_t dice = rng.next();
                    const double u_indel = (double)(dice & 0xffffff) / 16777216.0, u_clip = (double)((dice >> 24) & 0xffffff) / 16777216.0;
                    const double u_mapq = (double)((dice >> 48) & 0xffff) / 65536.0;
                    // sequencing-error indel in the middle of a gap-free read
                    if (u_indel < cfg.indel_err * READ_LEN && o.cigar.size() == 1 && (o.cigar[0] >> 4) > 60 && !has_indel) {
                        const int L = (int)(o.cigar[0] >> 4);
                        const int at = (int)rng.range(25, L - 25);
                        if (rng.next() & 1) {
                            o.seq.insert(o.seq.begin() + at, BASES[rng.below(4)]);
                            o.seq.resize((size_t)L);
                            o.cigar = {((uint32_t)at << 4) | 0u, (1u << 4) | 1u, ((uint32_t)(L - at - 1) << 4) | 0u};
                            o.rend = o.pos + L - 1;
                        } else if (o.pos + L + 1 <= C.len) {
                            o.seq = ref.substr((size_t)o.pos, (size_t)at) + ref.substr((size_t)(o.pos + at + 1), (size_t)(L - at));
                            o.cigar = {((uint32_t)at << 4) | 0u, (1u << 4) | 2u, ((uint32_t)(L - at) << 4) | 0u};
                            o.rend = o.pos + L + 1;
                        }
                    }
                    // soft clip at the 3' end of the read
                    if (u_clip < cfg.clip_frac * 0.6 && (o.cigar.front() & 0xf) == 0 && (o.cigar.back() & 0xf) == 0 && (o.cigar.front() >> 4) > 45 && (o.cigar.back() >> 4) > 45) {
                        const int k = (int)rng.range(5, 31);
                        if (forward) {
                            for (int j = 0; j < k; j++) { o.seq[o.seq.size() - 1 - (size_t)j] = BASES[rng.below(4)]; }
                            o.cigar.back() -= ((uint32_t)k << 4);
                            o.cigar.push_back(((uint32_t)k << 4) | 4u);
                            o.rend -= k;
                        } else {
                            for (int j = 0; j < k; j++) { o.seq[(size_t)j] = BASES[rng.below(4)]; }
                            o.cigar.front() -= ((uint32_t)k << 4);
                            o.cigar.insert(o.cigar.begin(), ((uint32_t)k << 4) | 4u);
                            o.pos += k;
                        }
                    }
                    // substitution errors
                    static const double PERR[26] = {1.0, 0.794328, 0.630957, 0.501187, 0.398107, 0.316228, 0.251189, 0.199526, 0.158489, 0.125893, 0.1, 0.0794328, 0.0630957,
                                                    0.0501187, 0.0398107, 0.0316228, 0.0251189, 0.0199526, 0.0158489, 0.0125893, 0.01, 0.00794328, 0.00630957, 0.00501187, 0.00398107, 0.00316228};
                    const uint32_t thr_hi = (uint32_t)(cfg.sub_err * 4294967296.0);
                    for (size_t i = 0; i < o.seq.size(); i++) {
                        const uint64_t r = rng.next();
                        const int q = o.qual[i];
                        const uint32_t thr = (q <= 25 ? (uint32_t)std::min(4294967295.0, (PERR[q] + cfg.sub_err) * 4294967296.0) : thr_hi);
                        if ((uint32_t)r < thr) { o.seq[i] = BASES[(base_code(o.seq[i]) + 1 + (int)((r >> 32) % 3)) % 4]; }
                    }
                    o.nm = compute_nm(ref, o);
                    o.mapq = (u_mapq < cfg.lowmapq_frac ? (int)rng.below(31) : 60);
                }
                if (!ok) { continue; }
                // mate fields
                char nb[64];
                int nlen = snprintf(nb, sizeof(nb), "r%06llu_%07llu", (unsigned long long)ri, (unsigned long long)frag_local);
                qname.assign(nb, (size_t)nlen);
                if (cfg.umi) {
                    char ua[40], ub[40];
                    for (int k = cfg.umi_len - 1, a = (int)umi_a, b = (int)umi_b; k >= 0; k--, a >>= 2, b >>= 2) { ua[k] = BASES[a & 3]; ub[k] = BASES[b & 3]; }
                    qname.push_back('#');
                    qname.append(swapped ? ub : ua, (size_t)cfg.umi_len);
                    qname.push_back('+');
                    qname.append(swapped ? ua : ub, (size_t)cfg.umi_len);
                }
                const int64_t left = std::min(rd[0].pos, rd[1].pos), right = std::max(rd[0].rend, rd[1].rend), tl = right - left;
                for (int r2 = 0; r2 < 2; r2++) {
                    const ReadOut & me = rd[r2], & mate = rd[r2 ^ 1];
                    const bool fwd = ((ftop ^ r2) != 0), mfwd = !fwd;
                    const int64_t isize = ((me.pos < mate.pos || (me.pos == mate.pos && r2 == 0)) ? tl : -tl);
                    const uint16_t flag = (uint16_t)(0x1 | 0x2 | (r2 ? 0x80 : 0x40) | (fwd ? 0 : 0x10) | (mfwd ? 0 : 0x20));
                    emit(A, R.ci, me, qname, flag, mate.pos, isize, ((uint64_t)ri << 32) | (frag_local * 2 + (uint64_t)r2));
                }
            }
        }
    }
};

struct Block { std::vector<uint8_t> cdata; uint32_t ulen; };

void deflate_block(const uint8_t *src, uint32_t n, int level, Block & out) {
    out.ulen = n;
    out.cdata.resize(18 + compressBound(n) + 8);
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);
    zs.next_in = (Bytef*)src; zs.avail_in = n;
    zs.next_out = out.cdata.data() + 18; zs.avail_out = (uInt)(out.cdata.size() - 18 - 8);
    deflate(&zs, Z_FINISH);
    const uint32_t clen = (uint32_t)zs.total_out;
    deflateEnd(&zs);
    const uint32_t bsize = clen + 25;
    static const uint8_t hdr[16] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 66, 67, 2, 0};
    memcpy(out.cdata.data(), hdr, 16);
    out.cdata[16] = (uint8_t)(bsize & 0xff); out.cdata[17] = (uint8_t)(bsize >> 8);
    const uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), src, n);
    memcpy(out.cdata.data() + 18 + clen, &crc, 4);
    memcpy(out.cdata.data() + 18 + clen + 4, &n, 4);
    out.cdata.resize(18 + clen + 8);
}

struct BaiBuilder {
    struct Chunk { uint64_t beg, end; };
    struct Ref { std::vector<std::pair<uint32_t, Chunk>> runs; std::vector<uint64_t> lin; };
    std::vector<Ref> refs;
    int32_t last_ci = -1; uint32_t last_bin = 0xffffffffu;
    explicit BaiBuilder(size_t n) : refs(n) {}
    void add(int32_t ci, int64_t pos, int64_t rend, uint64_t vbeg, uint64_t vend) {
        Ref & R = refs[(size_t)ci];
        const uint32_t bin = reg2bin(pos, rend);
        if (ci != last_ci || bin != last_bin) { R.runs.push_back(std::make_pair(bin, Chunk{vbeg, vend})); last_ci = ci; last_bin = bin; }
        else { R.runs.back().second.end = vend; }
        const size_t w0 = (size_t)(pos >> 14), w1 = (size_t)((rend - 1) >> 14);
        if (R.lin.size() <= w1) { R.lin.resize(w1 + 1, UINT64_MAX); }
        for (size_t w = w0; w <= w1; w++) { if (R.lin[w] == UINT64_MAX) { R.lin[w] = vbeg; } }
    }
    void write(const std::string & path) {
        FILE *f = fopen(path.c_str(), "wb");
        if (!f) { perror("bai"); exit(2); }
        fwrite("BAI\1", 1, 4, f);
        const int32_t n_ref = (int32_t)refs.size();
        fwrite(&n_ref, 4, 1, f);
        for (Ref & R : refs) {
            std::stable_sort(R.runs.begin(), R.runs.end(), [](const std::pair<uint32_t, Chunk> & a, const std::pair<uint32_t, Chunk> & b) { return a.first < b.first; });
            int32_t n_bin = 0;
            for (size_t i = 0; i < R.runs.size(); i++) { if (i == 0 || R.runs[i].first != R.runs[i - 1].first) { n_bin++; } }
            fwrite(&n_bin, 4, 1, f);
            for (size_t i = 0; i < R.runs.size();) {
                size_t j = i;
                while (j < R.runs.size() && R.runs[j].first == R.runs[i].first) { j++; }
                const uint32_t bin = R.runs[i].first; const int32_t n_chunk = (int32_t)(j - i);
                fwrite(&bin, 4, 1, f); fwrite(&n_chunk, 4, 1, f);
                for (size_t k = i; k < j; k++) { fwrite(&R.runs[k].second.beg, 8, 1, f); fwrite(&R.runs[k].second.end, 8, 1, f); }
                i = j;
            }
            for (size_t w = R.lin.size(); w-- > 0;) { if (R.lin[w] == UINT64_MAX) { R.lin[w] = (w + 1 < R.lin.size() ? R.lin[w + 1] : 0); } }
            const int32_t n_intv = (int32_t)R.lin.size();
            fwrite(&n_intv, 4, 1, f);
            if (n_intv) { fwrite(R.lin.data(), 8, (size_t)n_intv, f); }
        }
        fclose(f);
    }
};

template <class F> void parallel_for(size_t n, int threads, F body) {
    if (threads <= 1 || n <= 1) { for (size_t i = 0; i < n; i++) { body(i, 0); } return; }
    std::atomic<size_t> next(0);
    std::vector<std::thread> pool;
    const int nt = (int)std::min<size_t>((size_t)threads, n);
    for (int t = 0; t < nt; t++) { pool.emplace_back([&, t]() { for (;;) { const size_t i = next.fetch_add(1); if (i >= n) { break; } body(i, t); } }); }
    for (auto & th : pool) { th.join(); }
}

std::vector<double> parse_doubles(const char *s) { std::vector<double> v; for (const char *p = s; *p;) { char *e; v.push_back(strtod(p, &e)); p = (*e == ',' ? e + 1 : e); if (e == p && *e) { break; } } return v; }

} // namespace

int main(int argc, char **argv) {
    Config cfg;
    std::vector<std::pair<std::string, int64_t>> contig_spec;
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        auto val = [&]() -> const char* { if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", a.c_str()); exit(2); } return argv[++i]; };
        if (a == "--name") { cfg.name = val(); } else if (a == "--out") { cfg.outdir = val(); } else if (a == "--seed") { cfg.seed = strtoull(val(), NULL, 10); }
        else if (a == "--contig") { const std::string s = val(); const size_t c = s.find(':'); contig_spec.push_back(std::make_pair(s.substr(0, c), atoll(s.c_str() + c + 1))); }
        else if (a == "--depth") { cfg.depth = atof(val()); } else if (a == "--n-snv") { cfg.n_snv = atoi(val()); } else if (a == "--n-indel") { cfg.n_indel = atoi(val()); }
        else if (a == "--vafs") { cfg.vafs = parse_doubles(val()); } else if (a == "--max-indel-len") { cfg.max_indel_len = atoi(val()); }
        else if (a == "--targets") { const std::vector<double> t = parse_doubles(val()); if (t.size() != 4) { fprintf(stderr, "--targets n,first,step,len\n"); return 2; } cfg.n_targets = (int64_t)t[0]; cfg.target_first = (int64_t)t[1]; cfg.target_step = (int64_t)t[2]; cfg.target_len = (int64_t)t[3]; }
        else if (a == "--amplicon-frac") { cfg.amplicon_frac = atof(val()); } else if (a == "--umi") { cfg.umi = atoi(val()) != 0; } else if (a == "--umi-len") { cfg.umi_len = atoi(val()); }
        else if (a == "--family-mean") { cfg.family_mean = atof(val()); } else if (a == "--duplex-frac") { cfg.duplex_frac = atof(val()); } else if (a == "--swapped-umi-frac") { cfg.swapped_umi_frac = atof(val()); }
        else if (a == "--pcr-err-rate") { cfg.pcr_err_rate = atof(val()); } else if (a == "--sub-err") { cfg.sub_err = atof(val()); } else if (a == "--indel-err") { cfg.indel_err = atof(val()); }
        else if (a == "--clip-frac") { cfg.clip_frac = atof(val()); } else if (a == "--lowmapq-frac") { cfg.lowmapq_frac = atof(val()); } else if (a == "--str-every") { cfg.str_every = atoll(val()); }
        else if (a == "--threads") { cfg.threads = atoi(val()); } else if (a == "--level") { cfg.level = atoi(val()); }
        else { fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
    }
    if (contig_spec.empty()) { fprintf(stderr, "usage: uvc_synthgen --name N --out DIR --seed S --contig NAME:LEN [--contig ...] --depth D [--targets n,first,step,len] [--umi 1] ...\n"); return 2; }
    if (cfg.threads <= 0) { cfg.threads = (int)std::max(1u, std::thread::hardware_concurrency()); }
    for (auto & cs : contig_spec) { Contig c; c.name = cs.first; c.len = cs.second; cfg.contigs.push_back(c); }
    const std::string base = cfg.outdir + "/" + cfg.name;

    make_reference(cfg);
    write_fasta(cfg, base + ".fa");
    const std::vector<Region> regions = make_regions(cfg);
    const std::vector<Variant> variants = make_variants(cfg, regions);
    {
        FILE *f = fopen((base + ".truth.tsv").c_str(), "w");
        static const char *KIND[3] = {"snv", "ins", "del"};
        for (const Variant & v : variants) { fprintf(f, "%s\t%lld\t%s\t%s\t%s\t%g\n", cfg.contigs[(size_t)v.ci].name.c_str(), (long long)v.pos + 1, KIND[v.kind], v.ref.c_str(), v.alt.c_str(), v.vaf); }
        fclose(f);
    }
    int64_t target_bases = 0;
    if (cfg.n_targets > 0) {
        FILE *f = fopen((base + ".bed").c_str(), "w");
        for (const Region & R : regions) { fprintf(f, "%s\t%lld\t%lld\n", cfg.contigs[(size_t)R.ci].name.c_str(), (long long)R.beg, (long long)R.end); target_bases += R.end - R.beg; }
        fclose(f);
    } else { for (const Contig & c : cfg.contigs) { target_bases += c.len; } }

    FILE *bam = fopen((base + ".bam").c_str(), "wb");
    if (!bam) { perror("bam"); return 2; }
    uint64_t coff = 0;
    {   // header
        std::string text = "@HD\tVN:1.6\tSO:coordinate\n";
        for (const Contig & c : cfg.contigs) { text += "@SQ\tSN:" + c.name + "\tLN:" + std::to_string(c.len) + "\n"; }
        std::vector<uint8_t> h;
        h.insert(h.end(), {'B', 'A', 'M', 1});
        put32(h, (uint32_t)text.size()); h.insert(h.end(), text.begin(), text.end());
        put32(h, (uint32_t)cfg.contigs.size());
        for (const Contig & c : cfg.contigs) { put32(h, (uint32_t)c.name.size() + 1); h.insert(h.end(), c.name.begin(), c.name.end()); h.push_back(0); put32(h, (uint32_t)c.len); }
        for (size_t o = 0; o < h.size(); o += BGZF_MAX) {
            Block b; deflate_block(h.data() + o, (uint32_t)std::min<size_t>(BGZF_MAX, h.size() - o), cfg.level, b);
            fwrite(b.cdata.data(), 1, b.cdata.size(), bam); coff += b.cdata.size();
        }
    }

    Generator gen(cfg, regions, variants);
    BaiBuilder bai(cfg.contigs.size());
    int64_t n_reads = 0;
    // streaming: groups of consecutive regions are generated in parallel; records below the next group's lowest possible start are final
    const double reads_per_region = std::max(1.0, cfg.depth * (double)(regions[0].end - regions[0].beg) / READ_LEN * (cfg.umi ? 1.0 : 1.0));
    const size_t group = (size_t)std::max<double>(cfg.threads, std::min<double>(4096, 1.5e6 / reads_per_region));
    std::vector<Arena> pending;              // arenas that still hold unwritten records
    std::vector<std::vector<RecRef>> carry;  // their unwritten records
    for (size_t g0 = 0; g0 < regions.size(); g0 += group) {
        const size_t g1 = std::min(regions.size(), g0 + group);
        std::vector<Arena> arenas((size_t)cfg.threads);
        parallel_for(g1 - g0, cfg.threads, [&](size_t k, int t) { gen.generate_region(g0 + k, arenas[(size_t)t]); });
        // cutoff: nothing generated later can start below it
        int32_t cut_ci = INT32_MAX; int64_t cut_pos = 0;
        if (g1 < regions.size()) { cut_ci = regions[g1].ci; cut_pos = regions[g1].beg - cfg.insert_max - 8; }
        // all candidate records: carried-over ones and the new ones; an index entry names (arena, record)
        struct Ent { int32_t ci, pos; uint64_t order; uint32_t arena, idx; };
        std::vector<Arena*> all;
        for (Arena & a : pending) { all.push_back(&a); }
        std::vector<std::vector<RecRef>> lists;
        for (auto & c : carry) { lists.push_back(std::move(c)); }
        for (Arena & a : arenas) { all.push_back(&a); lists.push_back(std::move(a.recs)); }
        std::vector<Ent> fin;
        std::vector<std::vector<RecRef>> keep(all.size());
        for (size_t ai = 0; ai < all.size(); ai++) {
            for (size_t k = 0; k < lists[ai].size(); k++) {
                const RecRef & r = lists[ai][k];
                if (r.ci < cut_ci || (r.ci == cut_ci && r.pos < cut_pos)) { fin.push_back(Ent{r.ci, r.pos, r.order, (uint32_t)ai, (uint32_t)k}); }
                else { keep[ai].push_back(r); }
            }
        }
        // sort by (contig, pos, order): split into position slices, each sorted by one thread
        const size_t n_slices = (size_t)cfg.threads * 4;
        std::vector<std::vector<Ent>> slices(n_slices);
        if (!fin.empty()) {
            int64_t lo = INT64_MAX, hi = INT64_MIN;
            auto key = [&](const Ent & e) { return (int64_t)e.ci * ((int64_t)1 << 32) + e.pos; };
            for (const Ent & e : fin) { lo = std::min(lo, key(e)); hi = std::max(hi, key(e)); }
            const double width = (double)(hi - lo + 1) / (double)n_slices;
            for (const Ent & e : fin) { size_t k = (size_t)((double)(key(e) - lo) / width); if (k >= n_slices) { k = n_slices - 1; } slices[k].push_back(e); }
            parallel_for(n_slices, cfg.threads, [&](size_t s, int) {
                std::vector<Ent> & out = slices[s];
                std::sort(out.begin(), out.end(), [](const Ent & a, const Ent & b) { return a.ci != b.ci ? a.ci < b.ci : (a.pos != b.pos ? a.pos < b.pos : a.order < b.order); });
            });
        }
        // compress every slice into BGZF blocks (records are not split across blocks)
        struct SliceOut { std::vector<Block> blocks; std::vector<std::pair<uint32_t, uint32_t>> where; };   // per record: (block, offset)
        std::vector<SliceOut> outs(n_slices);
        parallel_for(n_slices, cfg.threads, [&](size_t s, int) {
            SliceOut & o = outs[s];
            std::vector<uint8_t> buf;
            buf.reserve(BGZF_MAX);
            auto flush = [&]() { if (buf.empty()) { return; } o.blocks.emplace_back(); deflate_block(buf.data(), (uint32_t)buf.size(), cfg.level, o.blocks.back()); buf.clear(); };
            for (const Ent & e : slices[s]) {
                const RecRef & r = lists[e.arena][e.idx];
                if (buf.size() + r.len > (size_t)BGZF_MAX) { flush(); }
                o.where.push_back(std::make_pair((uint32_t)o.blocks.size(), (uint32_t)buf.size()));
                buf.insert(buf.end(), all[e.arena]->bytes.begin() + r.off, all[e.arena]->bytes.begin() + r.off + r.len);
            }
            flush();
        });
        // write in order, index
        for (size_t s = 0; s < n_slices; s++) {
            SliceOut & o = outs[s];
            std::vector<uint64_t> bco(o.blocks.size() + 1);
            uint64_t c = coff;
            for (size_t b = 0; b < o.blocks.size(); b++) { bco[b] = c; c += o.blocks[b].cdata.size(); }
            bco[o.blocks.size()] = c;
            for (size_t k = 0; k < slices[s].size(); k++) {
                const Ent & e = slices[s][k];
                const RecRef & r = lists[e.arena][e.idx];
                const uint64_t vbeg = (bco[o.where[k].first] << 16) | o.where[k].second;
                uint64_t vend;
                if (k + 1 < slices[s].size()) { vend = (bco[o.where[k + 1].first] << 16) | o.where[k + 1].second; }
                else { vend = (c << 16); }
                bai.add(r.ci, r.pos, r.rend, vbeg, vend);
            }
            for (Block & b : o.blocks) { fwrite(b.cdata.data(), 1, b.cdata.size(), bam); }
            coff = c;
            n_reads += (int64_t)slices[s].size();
        }
        // carry over what is not final yet
        std::vector<Arena> npend; std::vector<std::vector<RecRef>> ncarry;
        for (size_t ai = 0; ai < all.size(); ai++) {
            if (keep[ai].empty()) { continue; }
            npend.emplace_back(); npend.back().bytes.swap(all[ai]->bytes);
            ncarry.push_back(std::move(keep[ai]));
        }
        pending.swap(npend); carry.swap(ncarry);
    }
    { Block b; deflate_block((const uint8_t*)"", 0, cfg.level, b); fwrite(b.cdata.data(), 1, b.cdata.size(), bam); }
    fclose(bam);
    bai.write(base + ".bam.bai");
    {
        FILE *f = fopen((base + ".meta.json").c_str(), "w");
        fprintf(f, "{\"bam\": \"%s.bam\", \"fasta\": \"%s.fa\", \"bed\": %s, \"n_reads\": %lld, \"n_positions\": %lld, \"n_variants\": %zu, \"contigs\": [",
                base.c_str(), base.c_str(), (cfg.n_targets > 0 ? ("\"" + base + ".bed\"").c_str() : "null"), (long long)n_reads, (long long)target_bases, variants.size());
        for (size_t i = 0; i < cfg.contigs.size(); i++) { fprintf(f, "%s[\"%s\", %lld]", i ? ", " : "", cfg.contigs[i].name.c_str(), (long long)cfg.contigs[i].len); }
        fprintf(f, "]}\n");
        fclose(f);
    }
    fprintf(stderr, "uvc_synthgen: %lld reads, %zu variants, %zu regions -> %s.bam\n", (long long)n_reads, variants.size(), regions.size(), base.c_str());
    return 0;
}
