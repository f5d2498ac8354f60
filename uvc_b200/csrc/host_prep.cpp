// host_prep.cpp - host-side staging of a batch of tiles.
//
// Stage P0 (read filter, dedup centres, family key, grouping) restates reference grouping.cpp:347-442, 608-997
// and MolecularID.hpp:20-69 with sort-based containers; the base-quality fix-ups restate grouping.cpp:459-543;
// stage P1 (repeat context, BAQ offsets) restates main.hpp:699-721, 794-874 and main.cpp:400-429.
// Reference quirks that change results are kept on purpose and marked QUIRK.
#include "host_prep.h"
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include <algorithm>
#include <mutex>
#include <atomic>
#include <string>
#include <thread>
#include <tuple>
#include <unordered_map>
#include <unordered_set>

#include <float.h>
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

namespace {

const int ARRPOS_MARGIN = UVC_MAX_INSERT_SIZE;   // grouping.cpp:22
const int ARRPOS_OUTER_RANGE = 10;               // grouping.cpp:23
const int ARRPOS_INNER_RANGE = 3;                // grouping.cpp:24

inline int64_t nnminus(int64_t a, int64_t b) { return (a > b ? a - b : 0); }

inline int cigar_op(uint32_t c) { return (int)(c & 0xf); }
inline int cigar_len(uint32_t c) { return (int)(c >> 4); }
inline bool op_consumes_ref(int op) { return op == UVC_CMATCH || op == UVC_CDEL || op == UVC_CREF_SKIP || op == UVC_CEQUAL || op == UVC_CDIFF; }
inline bool op_is_match(int op) { return op == UVC_CMATCH || op == UVC_CEQUAL || op == UVC_CDIFF; }

struct Raw {
    int32_t pos, rend, mpos, isize, mtid, l_qseq, n_cigar, nm;
    uint16_t flag; uint8_t mapq;
    const uint8_t *seq, *qual; const uint32_t *cigar; const char *qname;
};

inline Raw get_raw(const uvcgpu_reads_soa & rs, int64_t i) {
    Raw r;
    r.pos = rs.pos[i]; r.mpos = rs.mpos[i]; r.isize = rs.isize[i]; r.mtid = rs.mtid[i];
    r.l_qseq = rs.l_qseq[i]; r.n_cigar = rs.n_cigar[i]; r.nm = rs.nm[i]; r.flag = rs.flag[i]; r.mapq = rs.mapq[i];
    r.seq = rs.seq + rs.seq_off[i]; r.qual = rs.qual + rs.qual_off[i]; r.cigar = rs.cigar + rs.cigar_off[i]; r.qname = rs.qname + rs.qname_off[i];
    // bam_endpos: reference length of the alignment, 1 if it is zero or the read is unmapped
    int64_t rlen = 0;
    if (!(r.flag & 0x4)) { for (int k = 0; k < r.n_cigar; k++) { if (op_consumes_ref(cigar_op(r.cigar[k]))) { rlen += cigar_len(r.cigar[k]); } } }
    if (0 == rlen) { rlen = 1; }
    r.rend = (int32_t)(r.pos + rlen);
    // NORM_INSERT_SIZE (common.hpp:75)
    if (abs(r.isize) >= UVC_MAX_INSERT_SIZE) { r.isize = 0; }
    return r;
}

inline int read_strand(uint16_t flag) { return (((flag & 0x81) == 0x81) ? (!!(flag & 0x20)) : (!!(flag & 0x10))); }

enum Filt { KEEP, NOT_MAPPED, NOT_PRIMARY, LOW_MAPQ, LOW_ALN_LEN, LOW_ISIZE, HIGH_ISIZE, ZERO_ISIZE, OUT_OF_RANGE, NOT_END_TO_END };

// grouping.cpp:347-415
Filt classify(bool & isrc, bool & isr2, int32_t & tBeg, int32_t & tEnd, const Raw & r, int32_t fetch_tbeg, int32_t fetch_tend,
        const uvcgpu_params & par, bool end2end, bool pem) {
    if (r.flag & 0x4) { return NOT_MAPPED; }
    if (r.flag & 0x900) { return NOT_PRIMARY; }
    // QUIRK: the reference's call sites pass (min_aln_len, min_mapqual) in swapped order (grouping.cpp:676-677 vs :351-352)
    const int32_t min_mapqual = par.kept_aln_min_aln_len;
    const int32_t min_aln_len = par.kept_aln_min_mapqual;
    if ((int32_t)r.mapq < min_mapqual) { return LOW_MAPQ; }
    if ((r.rend - r.pos) < min_aln_len) { return LOW_ALN_LEN; }
    if (0 == r.isize) {
        if (par.kept_aln_is_zero_isize_discarded) { return ZERO_ISIZE; }
    } else {
        if (abs(r.isize) < par.kept_aln_min_isize) { return LOW_ISIZE; }
        if (abs(r.isize) > par.kept_aln_max_isize) { return HIGH_ISIZE; }
    }
    isrc = ((r.flag & 0x10) == 0x10);
    isr2 = ((r.flag & 0x80) == 0x80 && (r.flag & 0x1) == 0x1);
    if (!pem) { isr2 = false; }
    const int32_t begpos = r.pos, endpos = r.rend - 1;
    if ((!pem) || ((r.flag & 0x1) == 0) || (r.flag & 0x8) || (0 == r.isize) || (abs(r.isize) >= ARRPOS_MARGIN)) {
        tBeg = (isrc ? endpos : begpos);
        tEnd = (isrc ? begpos : endpos);
    } else {
        const int32_t b1 = std::min(begpos, r.mpos);
        const int32_t e1 = b1 + abs(r.isize) - 1;
        const bool strand = read_strand(r.flag);
        tBeg = (strand ? e1 : b1);
        tEnd = (strand ? b1 : e1);
    }
    const int32_t ob = std::min(tBeg, tEnd), oe = std::max(tBeg, tEnd);
    if (ob + (ARRPOS_MARGIN - ARRPOS_OUTER_RANGE) <= fetch_tbeg || fetch_tend - 1 + (ARRPOS_MARGIN - ARRPOS_OUTER_RANGE) <= oe) { return OUT_OF_RANGE; }
    if (end2end && !(ob <= fetch_tbeg && oe >= fetch_tend)) { return NOT_END_TO_END; }
    return KEEP;
}

// poscounter_to_pos2pcenter (grouping.cpp:422-442) evaluated on demand for one position: the local maximum within +-3 that the end
// position snaps to. The reference fills the whole array, which starts as all zeros (its inicount copy), hence 0 outside the loop range.
int32_t center_at(const std::vector<int32_t> & cnt, int32_t lo, const double *center_pow) {
    const int32_t n = (int32_t)cnt.size();
    if (lo < ARRPOS_INNER_RANGE || lo >= n - ARRPOS_INNER_RANGE) { return 0; }
    const int32_t lo_cnt = cnt[lo];
    int32_t center = lo;
    int32_t max_cnt = lo_cnt;
    for (int32_t hi = lo - ARRPOS_INNER_RANGE; hi < lo + ARRPOS_INNER_RANGE + 1; hi++) {
        const int32_t hi_cnt = cnt[hi];
        const int d = abs(lo - hi);
        if ((hi_cnt > max_cnt) && ((hi_cnt + 1) > (lo_cnt + 1) * center_pow[d])) {
            center = hi;
            max_cnt = hi_cnt;
        }
    }
    return center;
}

// FNV-1a over a NUL-terminated name (only used to bucket names in the per-tile name set; equality is always checked on the strings)
inline uint64_t name_hash(const char *s, size_t n) {
    // 8 bytes per step over the known length (never reads past the name), then the tail byte by byte
    uint64_t h = 1469598103934665603ULL;
    size_t i = 0;
    for (; i + 8 <= n; i += 8) {
        uint64_t w;
        memcpy(&w, s + i, 8);
        h = (h ^ w) * 0x9E3779B97F4A7C15ULL;
        h ^= h >> 29;
    }
    for (; i < n; i++) { h = (h ^ (uint64_t)(uint8_t)s[i]) * 1099511628211ULL; }
    return h;
}

inline uint64_t str_hash(const char *s, uint64_t base) {
    uint64_t h = 0;
    for (size_t i = 0; s[i]; i++) { h = h * base + (uint64_t)(int64_t)s[i]; }
    return h;
}

typedef std::pair<int32_t, int32_t> tidpos_t;

// a string that lives in the caller's qname buffer; ordered like std::string (bytewise, then by length)
struct StrView {
    const char *p; size_t n;
    StrView() : p(""), n(0) {}
    StrView(const char *p_, size_t n_) : p(p_), n(n_) {}
    int cmp(const StrView & o) const { const int r = memcmp(p, o.p, n < o.n ? n : o.n); return (r != 0 ? r : (n < o.n ? -1 : (n > o.n ? 1 : 0))); }
    bool operator==(const StrView & o) const { return n == o.n && 0 == memcmp(p, o.p, n); }
    bool operator!=(const StrView & o) const { return !(*this == o); }
    bool operator<(const StrView & o) const { return cmp(o) < 0; }
};

struct FamKey {
    tidpos_t beg, end;
    StrView qname, umi;
    uint32_t duplexflag, dedup_idflag;
    bool operator<(const FamKey & o) const { // MolecularID.hpp:52-68 (the hash tie-break can never decide: equal fields give equal hashes)
        if (beg != o.beg) { return beg < o.beg; }
        if (end != o.end) { return end < o.end; }
        if (qname != o.qname) { return qname < o.qname; }
        if (umi != o.umi) { return umi < o.umi; }
        if (duplexflag != o.duplexflag) { return duplexflag < o.duplexflag; }
        return dedup_idflag < o.dedup_idflag;
    }
    bool operator==(const FamKey & o) const {
        return beg == o.beg && end == o.end && qname == o.qname && umi == o.umi && duplexflag == o.duplexflag && dedup_idflag == o.dedup_idflag;
    }
};

struct Kept {
    int64_t raw;          // index into the caller's SoA
    Raw r;
    FamKey key;
    StrView umi_full;
    tidpos_t begpair, endpair;   // the read's own (non-key) MolecularBarcode ends
    int strand;
    uint64_t qhash2;
    int32_t fam_local, frag_local;
    int32_t simple, m_qoff, n_ev;   // CIGAR shape (see ReadRec), computed once in stage A
};

// main.hpp:699-721. QUIRK: rank2 is computed with rulen1 when rc2 <= 1.
bool more_str(int32_t rulen1, int32_t rc1, int32_t rulen2, int32_t rc2, int32_t repeatsize_max) {
    if (rulen2 * rc2 == 0) { return true; }
    if (rulen1 > repeatsize_max || rulen2 > repeatsize_max) { return (rulen1 < rulen2 || (rulen1 == rulen2 && rc1 > rc2)); }
    int rank1 = (rc1 <= 1 ? (-rc1 * rulen1) : ((rc1 - 1) * rulen1));
    int rank2 = (rc2 <= 1 ? (-rc2 * rulen1) : ((rc2 - 1) * rulen2));
    if (0 == rc1 || 0 == rulen1) { rank1 = -100; }
    if (0 == rc2 || 0 == rulen2) { rank2 = -100; }
    return rank1 > rank2;
}

// main.hpp:794-801 with prob2phred (main_conversion.hpp:890-893)
int32_t slip_phred(double ampfact, int32_t unit, int32_t nunits) {
    const int32_t region = unit * nunits;
    const double num_slips = (region > 64 ? (double)(region - 8) : log1p(exp((double)region - (double)8))) * ampfact / ((double)(unit * unit));
    return (int32_t)floor(-10 * log((1.0 - DBL_EPSILON) / (num_slips + 1.0)) / log(10));
}

// main.hpp:803-874: best short-tandem-repeat (unit <= str_max) and any-tandem-repeat (unit <= vntr_max) track per reference base
void repeat_context(uvcgpu_rtr *out, const char *ref, int32_t n, const uvcgpu_params & par) {   // out: n + 1 records
    for (int32_t i = 0; i <= n; i++) { uvcgpu_rtr & t = out[i]; t.begpos = 0; t.tracklen = 0; t.unitlen = 0; t.indelphred = par.indel_BQ_max; t.anyTR_begpos = 0; t.anyTR_tracklen = 0; t.anyTR_unitlen = 0; }
    const int32_t str_max = par.indel_str_repeatsize_max, vntr_max = par.indel_vntr_repeatsize_max;
    // cand[p] bit u (2 <= u <= umax): the period-u match run that starts at p is at least u long, i.e. unit u repeats (num >= 2). Units that do
    // not repeat can never win (see below), so the walk only examines the set bits. One backward sweep keeps the 48 run lengths (saturating
    // bytes) in three SSE registers: about 25 instructions per reference base instead of a 35-iteration scalar loop.
    const int32_t umax = std::min(vntr_max, 49);
    std::vector<uint64_t> cand((size_t)n + 1, 0);
#if defined(__SSE2__)
    {
        std::string padded(ref, (size_t)n);
        padded.append(64, '\0');                 // never equal to a reference character: comparisons past the end fail like `q + unit < n`
        __m128i run[3], uvec[3];
        for (int k = 0; k < 3; k++) {
            run[k] = _mm_setzero_si128();
            alignas(16) uint8_t uu[16];
            for (int j = 0; j < 16; j++) { uu[j] = (uint8_t)(2 + 16 * k + j); }
            uvec[k] = _mm_load_si128((const __m128i*)uu);
        }
        const __m128i one = _mm_set1_epi8(1);
        const uint64_t keep = ((umax >= 63) ? ~(uint64_t)0 : (((uint64_t)1 << (umax + 1)) - 1)) & ~(uint64_t)3;
        for (int32_t p = n - 1; p >= 0; p--) {
            const __m128i c = _mm_set1_epi8(padded[(size_t)p]);
            uint64_t m = 0;
            for (int k = 0; k < 3; k++) {
                const __m128i nxt = _mm_loadu_si128((const __m128i*)(padded.data() + p + 2 + 16 * k));
                const __m128i eq = _mm_cmpeq_epi8(nxt, c);
                run[k] = _mm_and_si128(_mm_adds_epu8(run[k], one), eq);
                const __m128i ge = _mm_cmpeq_epi8(_mm_max_epu8(run[k], uvec[k]), run[k]);   // run >= u
                m |= ((uint64_t)(uint32_t)_mm_movemask_epi8(ge)) << (2 + 16 * k);
            }
            cand[(size_t)p] = m & keep;
        }
    }
#else
    for (int32_t p = 0; p < n; p++) { cand[(size_t)p] = ((((uint64_t)1 << (umax + 1)) - 1) & ~(uint64_t)3); }    // no filter: every unit is examined
#endif
    for (int32_t refpos = 0; refpos < n;) {
        int32_t best_unit = 0, best_num = 0, best_end = refpos;
        int32_t any_unit = 0, any_num = 0, any_end = refpos;
        // units in increasing order: 1, then the repeating units among 2..umax (set bits of cand), then every unit above umax.
        // A unit > 1 that does not repeat (num = 1, rank -unit) never beats what unit 1 already set (rank >= -1, or QUIRK -unit).
        auto examine = [&](int32_t unit) {
            int32_t q = refpos;
            while (q + unit < n && ref[q] == ref[q + unit]) { q++; }
            if (unit > 1 && q - refpos < unit) { return; }      // num = 1 again
            const int32_t num = (q - refpos) / unit + 1;
            if (unit <= str_max && more_str(unit, num, best_unit, best_num, str_max)) { best_unit = unit; best_num = num; best_end = q + unit; }
            if (more_str(unit, num, any_unit, any_num, vntr_max)) { any_unit = unit; any_num = num; any_end = q + unit; }
        };
        if (vntr_max >= 1) { examine(1); }
        for (uint64_t todo = cand[(size_t)refpos]; todo; todo &= todo - 1) { examine(__builtin_ctzll(todo)); }
        for (int32_t unit = std::max(2, umax + 1); unit <= vntr_max; unit++) { examine(unit); }
        {
            const int32_t stop = std::min(best_end, n);
            const int32_t tl = stop - refpos;
            // slip_phred costs four libm calls; nearly every position asks for (unit 1, 1 repeat): memoised per thread for small arguments
            const double ampfact = par.indel_polymerase_slip_rate * par.indel_del_to_ins_err_ratio;
            static thread_local double memo_amp = -1;
            static thread_local int32_t memo[40][64];
            if (memo_amp != ampfact) { memo_amp = ampfact; for (auto & row : memo) { for (auto & x : row) { x = INT32_MIN; } } }
            const int32_t nun = tl / best_unit;
            int32_t dec;
            if (best_unit < 40 && nun < 64) {
                if (INT32_MIN == memo[best_unit][nun]) { memo[best_unit][nun] = slip_phred(ampfact, best_unit, nun); }
                dec = memo[best_unit][nun];
            } else {
                dec = slip_phred(ampfact, best_unit, nun);
            }
            for (int32_t i = refpos; i != stop; i++) {
                if (tl > out[i].tracklen) {
                    out[i].begpos = refpos; out[i].tracklen = tl; out[i].unitlen = best_unit;
                    out[i].indelphred = par.indel_BQ_max - std::min(par.indel_BQ_max - 1, dec);
                }
            }
        }
        {
            const int32_t stop = std::min(any_end, n);
            const int32_t tl = stop - refpos;
            for (int32_t i = refpos; i != stop; i++) {
                if (tl > out[i].anyTR_tracklen) { out[i].anyTR_begpos = refpos; out[i].anyTR_tracklen = tl; out[i].anyTR_unitlen = any_unit; }
            }
        }
        const int32_t nb = str_max + best_unit;
        refpos += std::max(best_unit * best_num, nb + 1) - nb;
    }
    if (n > 0) { out[n] = out[n - 1]; }
}

// main.cpp:400-429. QUIRK: the any-tandem-repeat variant still divides by the STR unit length.
void baq_prefix(int32_t *dst, const uvcgpu_rtr *rtr, size_t n, bool any_tr, const uvcgpu_params & par) {
    int64_t sum = 0;
    const int32_t polsize = (int32_t)round(par.indel_polymerase_size);
    for (size_t i = 0; i < n; i++) {
        const int32_t tl = (any_tr ? rtr[i].anyTR_tracklen : rtr[i].tracklen);
        const int32_t ul = rtr[i].unitlen;
        if (tl / ul >= 3 || (tl / ul >= 2 && tl >= polsize)) {
            sum += (par.indel_str_phred_per_region * 10) / tl + 1;
        } else {
            sum += par.indel_nonSTR_phred_per_base * 10;
        }
        dst[i] = (int32_t)sum;
    }
    for (size_t i = 0; i < n; i++) { dst[i] = (int32_t)((int64_t)dst[i] / 10); }
}

inline uint8_t char_to_symbol(char c) { // CHAR_TO_SYMBOL (main_conversion.hpp:473-486)
    switch (c) {
        case 'A': case 'a': return UVC_BASE_A;
        case 'C': case 'c': return UVC_BASE_C;
        case 'G': case 'g': return UVC_BASE_G;
        case 'T': case 't': return UVC_BASE_T;
        case 'I': case 'i': return UVC_LINK_M;
        case '-': case '_': return UVC_LINK_D1;
        default: return UVC_BASE_N;
    }
}

} // namespace

#include <chrono>
static std::atomic<int64_t> g_prof[8];
static inline int64_t prof_now() { return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
#define PROF(i) { const int64_t n_ = prof_now(); g_prof[i] += n_ - prof_t; prof_t = n_; }

void uvc_fill_view_constants(BatchView & v, const uvcgpu_params & par) {
    v.par = par;
    for (int d = 0; d < 4; d++) { v.center_pow[d] = pow(par.dedup_center_mult, (double)d); }
    v.indelphred_half = ((int32_t)round((10.0 / log(10.0)) * log(par.indel_del_to_ins_err_ratio))) / 2;
}

// Stages ONE tile into a private HostBatch whose offsets are all tile-local (the tile keeps its global index ti in the
// records). Tiles are independent (the reference runs them on different threads, main.cpp:1479), so the batch builder
// below stages them on all host cores and then concatenates.
// Stage A (everything that decides sizes: filter, family grouping, fragment/family records) keeps its kept reads in TileWork; after the batch
// offsets are known, stage B packs the reads and the per-position reference context of the tile straight into the batch arrays (one copy of
// the sequence/quality bytes instead of two).
// The kept-read vectors (a few hundred kilobytes per tile) are recycled through a process-wide cache: a fresh allocation of that size is
// mmap'ed, page-faulted in and unmapped again for every tile, which costs more than filling it.
struct KeptCache {
    std::mutex mu;
    std::vector<std::vector<Kept>> free_list;
    void take(std::vector<Kept> & v) { std::lock_guard<std::mutex> lk(mu); if (!free_list.empty()) { v.swap(free_list.back()); free_list.pop_back(); } v.clear(); }
    void give(std::vector<Kept> & v) { v.clear(); std::lock_guard<std::mutex> lk(mu); if (free_list.size() < 4096 && v.capacity() > 0) { free_list.emplace_back(); free_list.back().swap(v); } }
};
static KeptCache & kept_cache() { static KeptCache *c = new KeptCache(); return *c; }

// The same for the tile-private staging batches of stage A (fragment / family records of one tile): recycled with their capacity.
struct PartCache {
    std::mutex mu;
    std::vector<HostBatch> free_list;
    static void reset(HostBatch & b) {
        b.tiles.clear(); b.pos_tile.clear(); b.refsym.clear(); b.rtr.clear(); b.baq.clear(); b.baq2.clear(); b.reads.clear(); b.read_raw_index.clear();
        b.seq.clear(); b.qual.clear(); b.cigar.clear(); b.frags.clear(); b.frag_reads.clear(); b.fams.clear(); b.rfam.clear(); b.fam_umi.clear();
        b.fchunk_frag.clear(); b.mchunk_fs.clear();
        b.n_fcol = b.n_mcol = 0; b.n_pos = b.n_cx = b.n_ev = 0; b.n_reads_in = 0;
    }
    void take(HostBatch & b) { std::lock_guard<std::mutex> lk(mu); if (!free_list.empty()) { b = std::move(free_list.back()); free_list.pop_back(); } reset(b); }
    void give(HostBatch & b) { reset(b); std::lock_guard<std::mutex> lk(mu); if (free_list.size() < 4096) { free_list.emplace_back(std::move(b)); } b = HostBatch(); }
};
static PartCache & part_cache() { static PartCache *c = new PartCache(); return *c; }

struct TileWork {
    std::vector<Kept> kept;
    size_t n_seq = 0, n_qual = 0, n_cig = 0;
    int32_t ext_end_ref = 0;
    const HostContig *contig = NULL;
};
static int build_tile_a(HostBatch & hb, TileWork & tw, const uvcgpu_params & par, const std::map<int32_t, HostContig> & contigs,
        int32_t ti, const uvcgpu_tile & ut, const uvcgpu_reads_soa & rs, const double *center_pow, bool pem, std::string & msg) {
    hb.tiles.resize(1);
    {
        TileInfo & T = hb.tiles[0];
        memset(&T, 0, sizeof(T));
        T.tid = ut.tid; T.beg_pos = ut.beg_pos; T.end_pos = ut.end_pos; T.region_flag = ut.region_flag;
        T.prev_tid = ut.prev_tid; T.prev_beg_pos = ut.prev_beg_pos; T.prev_end_pos = ut.prev_end_pos;
        T.pos_off = 0; T.read_off = 0; T.frag_off = 0; T.fam_off = 0;
        if (ut.read_begin < 0 || ut.read_end > rs.n_reads || ut.read_begin > ut.read_end || ut.beg_pos >= ut.end_pos) { msg = "invalid tile"; return UVCGPU_EINVAL; }
        auto cit = contigs.find(ut.tid);
        hb.n_reads_in += ut.read_end - ut.read_begin;
        const int32_t fetch_tbeg = ut.beg_pos, fetch_tend = ut.end_pos;
        const bool end2end = (ut.region_flag & 0x1);
        const int32_t fetch_size = fetch_tend - fetch_tbeg + (ARRPOS_MARGIN + ARRPOS_OUTER_RANGE) * 2;
        std::vector<int32_t> beg_cnt[4], end_cnt[4];
        std::vector<int64_t> border_psum[4];
        for (int c = 0; c < 4; c++) { beg_cnt[c].assign(fetch_size, 0); end_cnt[c].assign(fetch_size, 0); border_psum[c].assign((size_t)fetch_size + 1, 0); }

        int64_t prof_t = prof_now();
        // pass 1 (grouping.cpp:666-695): end histograms and the set of fragment names that touch the tile
        // set of qnames (grouping.cpp:648 visited_qnames): open addressing over read indices, names compared as strings
        const int64_t n_in = ut.read_end - ut.read_begin;
        size_t vcap = 16;
        while ((int64_t)vcap < 2 * n_in) { vcap *= 2; }
        std::vector<int32_t> vslot(vcap, -1);                 // tile-local read index of the first read that carries the name
        std::vector<uint64_t> vhash((size_t)n_in);            // name hash of every read of the tile
        auto visited_find = [&](const char *name, uint64_t h) -> bool {
            for (size_t k = (size_t)h & (vcap - 1);; k = (k + 1) & (vcap - 1)) {
                const int32_t j = vslot[k];
                if (j < 0) { return false; }
                if (vhash[(size_t)j] == h && 0 == strcmp(rs.qname + rs.qname_off[ut.read_begin + j], name)) { return true; }
            }
        };
        auto visited_insert = [&](const char *name, uint64_t h, int32_t j_new) {
            for (size_t k = (size_t)h & (vcap - 1);; k = (k + 1) & (vcap - 1)) {
                const int32_t j = vslot[k];
                if (j < 0) { vslot[k] = j_new; return; }
                if (vhash[(size_t)j] == h && 0 == strcmp(rs.qname + rs.qname_off[ut.read_begin + j], name)) { return; }
            }
        };
        // what pass 1 learns about a read is kept for pass 2 (the classification, the template ends, whether the read itself put its name in the set)
        struct Pass1 { int32_t rend, tBeg, tEnd; int8_t c; uint8_t keep, touches; };
        std::vector<Pass1> p1((size_t)n_in);
        for (int64_t i = ut.read_begin; i < ut.read_end; i++) {
            const Raw r = get_raw(rs, i);
            Pass1 & P = p1[(size_t)(i - ut.read_begin)];
            P.rend = r.rend; P.keep = 0; P.touches = 0; P.c = 0; P.tBeg = P.tEnd = 0;
            vhash[(size_t)(i - ut.read_begin)] = name_hash(r.qname, strnlen(r.qname, (size_t)(rs.qname_off[i + 1] - rs.qname_off[i])));
            bool isrc = false, isr2 = false; int32_t tBeg = 0, tEnd = 0;
            if (KEEP != classify(isrc, isr2, tBeg, tEnd, r, fetch_tbeg, fetch_tend, par, end2end, pem)) { continue; }
            const int c = isrc * 2 + isr2;
            P.keep = 1; P.c = (int8_t)c; P.tBeg = tBeg; P.tEnd = tEnd;
            const int32_t bi = tBeg + ARRPOS_MARGIN - fetch_tbeg, ei = tEnd + ARRPOS_MARGIN - fetch_tbeg;
            if (bi >= 0 && bi < fetch_size) { beg_cnt[c][bi] += 1; }
            if (ei >= 0 && ei < fetch_size) { end_cnt[c][ei] += 1; }
            const int32_t lo = std::min(tBeg, tEnd), hi = std::max(tBeg, tEnd) + 2;
            if (!((hi <= fetch_tbeg) || (fetch_tend <= lo))) { P.touches = 1; visited_insert(r.qname, vhash[(size_t)(i - ut.read_begin)], (int32_t)(i - ut.read_begin)); }
        }
        PROF(0)
        for (int c = 0; c < 4; c++) {
            int64_t bs = 0, es = 0;
            for (int32_t i = 0; i < fetch_size; i++) { bs += beg_cnt[c][i]; es += end_cnt[c][i]; border_psum[c][i + 1] = bs + es; }
        }

        PROF(1)
        // pass 2 (grouping.cpp:731-977): family key of every kept read
        std::vector<Kept> & kept = tw.kept;
        kept_cache().take(kept);
        kept.reserve((size_t)n_in);
        int32_t bam_beg = INT32_MAX, bam_end = 0;
        int64_t pcrpassed = 0;
        for (int64_t i = ut.read_begin; i < ut.read_end; i++) {
            const Pass1 & P = p1[(size_t)(i - ut.read_begin)];
            // (the order of the reference's tests - position window, name set, classification - does not matter: all must hold)
            if (!P.keep) { continue; }
            if (rs.pos[i] < nnminus(fetch_tbeg, UVC_MAX_INSERT_SIZE + 1) || P.rend > (fetch_tend + UVC_MAX_INSERT_SIZE + 1)) { continue; }
            if (!P.touches && !visited_find(rs.qname + rs.qname_off[i], vhash[(size_t)(i - ut.read_begin)])) { continue; }
            Raw r = get_raw(rs, i);
            const bool isrc = (P.c >> 1) & 1, isr2 = P.c & 1; const int32_t tBeg = P.tBeg, tEnd = P.tEnd;
            bam_beg = std::min(bam_beg, r.pos);
            bam_end = std::max(bam_end, r.rend);
            const char *qname = r.qname;
            // one pass over the name: its length, the first two '#' and the order-defining hash (str_hash(qname, 17), MolecularID / grouping.cpp:932)
            size_t qlen = 0;
            const char *hash1 = NULL, *hash2 = NULL;
            uint64_t qh2 = 0;
            for (const char *c = qname; *c; c++, qlen++) {
                qh2 = qh2 * 17 + (uint64_t)(int64_t)*c;
                if ('#' == *c) { if (!hash1) { hash1 = c; } else if (!hash2) { hash2 = c; } }
            }
            const char *umi_beg = (hash1 ? hash1 + 1 : qname + qlen);
            const char *umi_end = (hash2 ? hash2 : qname + qlen);
            const bool umi_found = ((umi_beg + 1 < umi_end) && (1 /* MOLECULE_TAG_NONE */ != par.molecule_tag));
            bool duplex_found = false;
            const size_t umi_len = umi_end - umi_beg;
            if (umi_found) {
                const size_t half = (umi_len - 1) / 2;
                if ((umi_len % 2 == 1) && ('+' == umi_beg[half]) && (!par.disable_duplex)) { duplex_found = true; }
            }
            const int c = isrc * 2 + isr2;
            const int32_t beg1 = tBeg + ARRPOS_MARGIN - fetch_tbeg, end1 = tEnd + ARRPOS_MARGIN - fetch_tbeg;
            const int32_t beg2 = center_at(beg_cnt[c], beg1, center_pow), end2 = center_at(end_cnt[c], end1, center_pow);
            const int64_t beg2count = beg_cnt[c][beg2], end2count = end_cnt[c][end2];
            const int32_t insL = std::min(beg2 + 6, end2);
            const int32_t insR = std::max((int64_t)beg2, nnminus(end2, 6));
            const int64_t tot = border_psum[c][insR] - border_psum[c][insL];
            const double begratio = (double)(beg2count * (insR - insL) + 1) / (double)(tot + (insR - insL) + 1);
            const double endratio = (double)(end2count * (insR - insL) + 1) / (double)(tot + (insR - insL) + 1);
            const bool beg_amp = (begratio > par.dedup_amplicon_border_to_insert_cov_weak_avgDP_ratio
                    && (beg2count >= par.dedup_amplicon_border_weak_minDP) && (beg2count >= tot * par.dedup_amplicon_border_to_insert_cov_weak_totDP_ratio));
            const bool end_amp = (endratio > par.dedup_amplicon_border_to_insert_cov_weak_avgDP_ratio
                    && (end2count >= par.dedup_amplicon_border_weak_minDP) && (end2count >= tot * par.dedup_amplicon_border_to_insert_cov_weak_totDP_ratio));
            const bool beg_strong = (begratio > par.dedup_amplicon_border_to_insert_cov_strong_avgDP_ratio
                    && (beg2count >= par.dedup_amplicon_border_strong_minDP) && (beg2count >= tot * par.dedup_amplicon_border_to_insert_cov_strong_totDP_ratio));
            const bool end_strong = (endratio > par.dedup_amplicon_border_to_insert_cov_strong_avgDP_ratio
                    && (end2count >= par.dedup_amplicon_border_strong_minDP) && (end2count >= tot * par.dedup_amplicon_border_to_insert_cov_strong_totDP_ratio));
            const bool assay_amplicon = (beg_strong || end_strong || (beg_amp && end_amp));
            pcrpassed += assay_amplicon;
            uint32_t idflag = 0;
            if (par.dedup_flag != 0) {
                idflag = par.dedup_flag;
            } else if (umi_found) {
                if (beg_strong && end_amp && beg2count > end2count * par.dedup_amplicon_end2end_ratio) { idflag = 0x9; }
                else if (end_strong && beg_amp && end2count > beg2count * par.dedup_amplicon_end2end_ratio) { idflag = 0xA; }
                else { idflag = 0xB; }
            } else if (assay_amplicon) {
                idflag = 0x7;
            } else {
                idflag = 0x3;
            }
            const bool preserved = ((r.flag & 0x1) && (!(r.flag & 0x4)) && (!(r.flag & 0x8)) && (abs(r.isize) >= (UVC_MAX_INSERT_SIZE * 3 / 4) || r.isize == 0));
            const int32_t begtid = ut.tid;
            const int32_t endtid = (((r.flag & 0x1) && !(r.flag & 0x8)) ? r.mtid : (INT32_MAX - 1));
            const tidpos_t begpair(begtid, preserved ? r.pos : (beg2 - ARRPOS_MARGIN + fetch_tbeg));
            const tidpos_t endpair(endtid, preserved ? r.mpos : (end2 - ARRPOS_MARGIN + fetch_tbeg));
            Kept k;
            k.raw = i; k.r = r; k.strand = read_strand(r.flag); k.qhash2 = qh2;
            k.umi_full = (umi_found ? StrView(umi_beg, umi_len) : StrView());
            k.begpair = begpair; k.endpair = endpair;
            // MolecularBarcode::createKey (MolecularID.hpp:20-51)
            k.key.beg = tidpos_t(-1, -1); k.key.end = tidpos_t(-1, -1);
            if (0x3 == (0x3 & idflag)) { k.key.beg = std::min(begpair, endpair); k.key.end = std::max(begpair, endpair); }
            else if (0x1 & idflag) { k.key.beg = begpair; }
            else if (0x2 & idflag) { k.key.end = endpair; }
            if (0x4 & idflag) { k.key.qname = StrView(qname, qlen); }
            if (0x8 & idflag) { k.key.umi = k.umi_full; }
            k.key.duplexflag = (umi_found ? 0x1 : 0) + (duplex_found ? 0x2 : 0) + (assay_amplicon ? 0x4 : 0) + (preserved ? 0x8 : 0);
            k.key.dedup_idflag = idflag;
            k.fam_local = k.frag_local = -1;
            kept.push_back(k);
        }
        PROF(2)
        T.num_passed = (int64_t)kept.size();
        T.num_pcrpassed = pcrpassed;
        T.bam_inclu_beg = bam_beg; T.bam_exclu_end = bam_end;
        T.is_amplicon_inferred = !((pcrpassed) * 2 <= (int64_t)kept.size());
        if (kept.empty()) { T.skipped = 1; T.ext_beg = T.ext_end = 0; return 0; }
        if (cit == contigs.end()) { msg = "contig of a tile was not set with uvcgpu_set_contig"; return UVCGPU_EINVAL; }
        const HostContig & contig = cit->second;
        tw.contig = &contig;
        T.rpos_inclu_beg = std::max(ut.beg_pos, bam_beg);
        T.rpos_exclu_end = std::min(ut.end_pos, bam_end);
        T.ext_beg = (int32_t)std::max((int64_t)0, nnminus(std::min(ut.beg_pos, bam_beg), UVC_MAX_STR_N_BASES));
        const int32_t ext_end_ref = (int32_t)std::min((int64_t)ut.contig_len, (int64_t)std::max(ut.end_pos, bam_end) + UVC_MAX_STR_N_BASES);
        T.ext_end = ext_end_ref + 1;
        tw.ext_end_ref = ext_end_ref;
        if (contig.available && (int64_t)ext_end_ref > contig.len) { msg = "tile extends beyond the contig that was set"; return UVCGPU_EINVAL; }

        // families in MolecularBarcode order; fragments by qname hash inside (family, strand); reads in file order inside a fragment
        std::vector<int32_t> order(kept.size());
        for (size_t i = 0; i < kept.size(); i++) { order[i] = (int32_t)i; }
        std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) {
            if (!(kept[a].key == kept[b].key)) { return kept[a].key < kept[b].key; }
            if (kept[a].strand != kept[b].strand) { return kept[a].strand < kept[b].strand; }
            return kept[a].qhash2 < kept[b].qhash2;
        });
        PROF(3)
        const int64_t read_base = (int64_t)hb.reads.size();
        std::vector<int32_t> l2r_end, r2l_end;      // (reused by every family: no allocation per family)
        hb.frag_reads.reserve(kept.size());
        for (size_t oi = 0; oi < order.size();) {
            size_t oj = oi;
            while (oj < order.size() && kept[order[oj]].key == kept[order[oi]].key) { oj++; }
            FamRec F;
            memset(&F, 0, sizeof(F));
            F.tile = ti;
            const FamKey & key = kept[order[oi]].key;
            F.duplexflag = key.duplexflag; F.dedup_idflag = key.dedup_idflag;
            const int32_t fam_index = (int32_t)hb.fams.size();
            // the reference keeps the MolecularBarcode of the first inserted read (file order) as the family's non-key data
            int32_t first_in_file = order[oi];
            for (size_t o = oi; o < oj; o++) { first_in_file = std::min(first_in_file, order[o]); }
            hb.fam_umi.push_back(std::string(kept[first_in_file].umi_full.p, kept[first_in_file].umi_full.n));
            F.beg_tid = kept[first_in_file].begpair.first; F.beg_pos = kept[first_in_file].begpair.second;
            F.end_tid = kept[first_in_file].endpair.first; F.end_pos = kept[first_in_file].endpair.second;
            int32_t both_beg = INT32_MAX, both_end = 0;
            size_t o = oi;
            for (int strand = 0; strand < 2; strand++) {
                F.frag_off[strand] = (int32_t)hb.frags.size();
                int32_t s_beg = INT32_MAX, s_end = 0, s_hi = 0;
                l2r_end.clear(); r2l_end.clear();
                int64_t qseqlen_sum = 0, n_qseqs = 0;
                while (o < oj && kept[order[o]].strand == strand) {
                    size_t p = o;
                    while (p < oj && kept[order[p]].strand == strand && kept[order[p]].qhash2 == kept[order[o]].qhash2) { p++; }
                    FragRec G;
                    memset(&G, 0, sizeof(G));
                    G.tile = ti; G.fam = fam_index; G.strand = strand;
                    G.read_off = (int32_t)hb.frag_reads.size(); G.n_reads = (int32_t)(p - o);
                    int32_t f_beg = INT32_MAX, f_end = 0, f_hi = 0;
                    for (size_t q = o; q < p; q++) { // stable sort kept file order inside the fragment
                        Kept & k = kept[order[q]];
                        k.fam_local = fam_index; k.frag_local = (int32_t)hb.frags.size();
                        hb.frag_reads.push_back((int32_t)(read_base + order[q]));
                        // fillTidBegEndFromAlns1 (main.hpp:659-673). QUIRK: the exclusive end grows by one per alignment visited.
                        f_beg = std::min(f_beg, k.r.pos); f_end = std::max(f_end, k.r.rend) + 1; f_hi = std::max(f_hi, k.r.rend);
                        s_beg = std::min(s_beg, k.r.pos); s_end = std::max(s_end, k.r.rend) + 1; s_hi = std::max(s_hi, k.r.rend);
                        both_beg = std::min(both_beg, k.r.pos); both_end = std::max(both_end, k.r.rend) + 1;
                        G.normMQ = std::max(G.normMQ, (int32_t)k.r.mapq);
                        if (k.r.flag & 0x10) { r2l_end.push_back(k.r.pos); } else { l2r_end.push_back(k.r.rend); }
                        qseqlen_sum += k.r.l_qseq; n_qseqs += 1;
                    }
                    G.beg = f_beg; G.end = f_end;
                    G.lo = f_beg; G.hi = f_hi;
                    G.col_off = hb.n_fcol;
                    hb.n_fcol += ((int64_t)(G.hi - G.lo) + UVC_COL_CHUNK - 1) / UVC_COL_CHUNK * UVC_COL_CHUNK;
                    hb.frags.push_back(G);
                    o = p;
                }
                F.n_frags[strand] = (int32_t)hb.frags.size() - F.frag_off[strand];
                if (F.n_frags[strand] > 65535) { msg = "a molecule family has more than 65535 fragments on one strand (FamCol counts are 16-bit)"; return UVCGPU_EUNSUPPORTED; }
                F.beg2[strand] = s_beg; F.end2[strand] = s_end;
                F.lo[strand] = (F.n_frags[strand] > 0 ? s_beg : 0); F.hi[strand] = (F.n_frags[strand] > 0 ? s_hi : 0);
                F.col_off[strand] = hb.n_mcol;
                F.direct_frag[strand] = -1;
                if (1 == F.n_frags[strand] && !(par.microadjust_padded_deletion_flag & 0x1)) { F.direct_frag[strand] = F.frag_off[strand]; }
                else { hb.n_mcol += ((int64_t)(F.hi[strand] - F.lo[strand]) + UVC_COL_CHUNK - 1) / UVC_COL_CHUNK * UVC_COL_CHUNK; }
                // MEDIAN of the unsorted vectors (main_conversion.hpp:24-28, main.hpp:2939-2940)
                F.l2r_end_median[strand] = (l2r_end.size() ? (l2r_end[(l2r_end.size() - 1) / 2] + l2r_end[l2r_end.size() / 2]) / 2 : s_end);
                F.r2l_end_median[strand] = (r2l_end.size() ? (r2l_end[(r2l_end.size() - 1) / 2] + r2l_end[r2l_end.size() / 2]) / 2 : s_beg);
                F.qlen_ok[strand] = ((F.n_frags[strand] >= par.fam_thres_dup1add) && (qseqlen_sum >= n_qseqs * par.fam_thres_qseqlen));
                F.nsb_min[strand] = s_end; F.nsb_max[strand] = s_beg;
            }
            F.beg_both = both_beg; F.end_both = both_end;
            hb.fams.push_back(F);
            oi = oj;
        }

        PROF(4)
        // sizes of what stage B will write
        int32_t max_span = 0;
        for (Kept & k : kept) {
            tw.n_seq += (size_t)(k.r.l_qseq + 1) / 2; tw.n_qual += (size_t)k.r.l_qseq; tw.n_cig += (size_t)k.r.n_cigar;
            // simple = [S|H|P]* (M|=|X) [S|H|P]*
            int lead = 0, a = 0, b = k.r.n_cigar;
            while (a < b && (cigar_op(k.r.cigar[a]) == UVC_CSOFT_CLIP || cigar_op(k.r.cigar[a]) == UVC_CHARD_CLIP || cigar_op(k.r.cigar[a]) == UVC_CPAD)) {
                if (cigar_op(k.r.cigar[a]) == UVC_CSOFT_CLIP) { lead += cigar_len(k.r.cigar[a]); }
                a++;
            }
            while (b > a && (cigar_op(k.r.cigar[b - 1]) == UVC_CSOFT_CLIP || cigar_op(k.r.cigar[b - 1]) == UVC_CHARD_CLIP || cigar_op(k.r.cigar[b - 1]) == UVC_CPAD)) { b--; }
            k.simple = ((b - a == 1) && op_is_match(cigar_op(k.r.cigar[a])));
            k.m_qoff = lead;
            k.n_ev = 0;
            if (!k.simple) {
                hb.n_cx += (k.r.rend - k.r.pos);
                for (int c = 0; c < k.r.n_cigar; c++) { if (cigar_op(k.r.cigar[c]) == UVC_CINS || cigar_op(k.r.cigar[c]) == UVC_CDEL) { k.n_ev++; } }
                hb.n_ev += k.n_ev;
            }
            max_span = std::max(max_span, k.r.rend - k.r.pos);
        }
        T.n_reads = (int32_t)kept.size();
        T.n_frags = (int32_t)((int64_t)hb.frags.size() - T.frag_off);
        T.n_fams = (int32_t)((int64_t)hb.fams.size() - T.fam_off);
        T.max_read_span = max_span;
        hb.n_pos = T.ext_end - T.ext_beg;
        PROF(5)
    }
    return 0;
}

struct BatchOff { int64_t pos, read, frag, fam, fragread, seq, qual, cigar, cx, ev, fcol, mcol; };

// Stage B: the tile's reads (file order) and reference context written at their final places in the batch arrays.
static void build_tile_b(HostBatch & hb, const TileWork & tw, const TileInfo & T, const BatchOff & o, const uvcgpu_params & par, int32_t ti,
        size_t n_frags_tile, size_t n_fams_tile) {
    int64_t prof_t = prof_now();
    const std::vector<Kept> & kept = tw.kept;
    if (kept.empty()) { return; }
    {
        std::vector<int32_t> frag_maxrend(n_frags_tile, INT32_MIN), fam_maxrend(2 * n_fams_tile, INT32_MIN), famboth_maxrend(n_fams_tile, INT32_MIN);
        uint64_t seq_at = (uint64_t)o.seq, qual_at = (uint64_t)o.qual, cig_at = (uint64_t)o.cigar;
        int64_t cx_at = o.cx, ev_at = o.ev;
        for (size_t i = 0; i < kept.size(); i++) {
            const Kept & k = kept[i];
            ReadRec & R = hb.reads[(size_t)o.read + i];      // built in place
            memset(&R, 0, sizeof(R));
            R.pos = k.r.pos; R.rend = k.r.rend; R.mpos = k.r.mpos; R.isize = k.r.isize;
            R.l_qseq = k.r.l_qseq; R.n_cigar = k.r.n_cigar; R.nm = k.r.nm;
            R.flag = k.r.flag; R.mapq = k.r.mapq; R.strand = (uint8_t)k.strand;
            R.dflag = k.key.duplexflag; R.tile = ti; R.frag = k.frag_local + (int32_t)o.frag; R.fam = k.fam_local + (int32_t)o.fam;
            R.seq_off = seq_at; R.qual_off = qual_at; R.cigar_off = cig_at;
            memcpy(hb.seq.data() + seq_at, k.r.seq, (size_t)(k.r.l_qseq + 1) / 2); seq_at += (uint64_t)(k.r.l_qseq + 1) / 2;
            memcpy(hb.qual.data() + qual_at, k.r.qual, (size_t)k.r.l_qseq); qual_at += (uint64_t)k.r.l_qseq;
            memcpy(hb.cigar.data() + cig_at, k.r.cigar, (size_t)k.r.n_cigar * sizeof(uint32_t)); cig_at += (uint64_t)k.r.n_cigar;
            R.simple = k.simple; R.m_qoff = k.m_qoff;
            R.cx_off = -1; R.ev_off = (int32_t)ev_at; R.n_ev = k.n_ev;
            if (!R.simple) { R.cx_off = (int32_t)cx_at; cx_at += (R.rend - R.pos); ev_at += R.n_ev; }
            R.fragprev_maxrend = frag_maxrend[(size_t)k.frag_local];
            frag_maxrend[(size_t)k.frag_local] = std::max(frag_maxrend[(size_t)k.frag_local], R.rend);
            const size_t fkey = (size_t)k.fam_local * 2 + R.strand;
            R.famprev_maxrend = fam_maxrend[fkey];
            fam_maxrend[fkey] = std::max(fam_maxrend[fkey], R.rend);
            R.fambothprev_maxrend = famboth_maxrend[(size_t)k.fam_local];
            famboth_maxrend[(size_t)k.fam_local] = std::max(famboth_maxrend[(size_t)k.fam_local], R.rend);
            hb.read_raw_index[(size_t)o.read + i] = k.raw;
        }
    }
    PROF(5)
    {
        // stage P1: reference symbols, repeat context, BAQ prefix sums over [ext_beg, ext_end)
        const int32_t npos = T.ext_end - T.ext_beg;
        const int32_t nref = npos - 1;
        std::string refstring;
        if (tw.contig->available) { refstring.assign(tw.contig->bases.data() + T.ext_beg, (size_t)nref); }
        else { refstring.assign((size_t)nref, 'n'); }
        const size_t poff = (size_t)o.pos;
        std::fill(hb.pos_tile.begin() + poff, hb.pos_tile.begin() + poff + npos, ti);
        for (int32_t i = 0; i < nref; i++) { hb.refsym[poff + i] = char_to_symbol(refstring[i]); }
        hb.refsym[poff + nref] = UVC_BASE_N;
        PROF(6)
        repeat_context(hb.rtr.data() + poff, refstring.data(), nref, par);
        PROF(7)
        baq_prefix(hb.baq.data() + poff, hb.rtr.data() + poff, (size_t)npos, false, par);
        baq_prefix(hb.baq2.data() + poff, hb.rtr.data() + poff, (size_t)npos, true, par);
        PROF(6)
    }
}

int uvc_host_threads(int32_t n_tiles, int requested) {
    int n = (requested > 0 ? requested : (int)std::thread::hardware_concurrency());
    const char *e = getenv("UVC_HOST_THREADS");
    if (e && atoi(e) > 0) { n = atoi(e); }
    if (n < 1) { n = 1; }
    if (n > 64) { n = 64; }
    return std::min<int>(n, n_tiles);
}


int uvc_build_host_batch(HostBatch & hb, const uvcgpu_params & par, const std::map<int32_t, HostContig> & contigs,
        int32_t n_tiles, const uvcgpu_tile *tiles, const uvcgpu_reads_soa *sources, const int32_t *tile_source, int n_threads_req, std::string & msg) {
    if (par.inferred_sequencing_platform != 1) { msg = "only the Illumina/BGI platform path is implemented"; return UVCGPU_EUNSUPPORTED; }
    const bool pem = (0 == par.pair_end_merge);
    double center_pow[4];
    for (int d = 0; d < 4; d++) { center_pow[d] = pow(par.dedup_center_mult, (double)d); }
    hb = HostBatch();
    const int n_threads = n_threads_req;
    const int64_t wall0 = prof_now();
    // 1. every tile staged privately, on all host cores
    std::vector<HostBatch> part((size_t)n_tiles);
    std::vector<TileWork> work((size_t)n_tiles);
    std::vector<int> rcs((size_t)n_tiles, 0);
    std::vector<std::string> msgs((size_t)n_tiles);
    uvc_parallel_for(n_tiles, n_threads, [&](int32_t ti) {
        uvc_stage_thread_pinning(false);
        part_cache().take(part[ti]);
        rcs[ti] = build_tile_a(part[ti], work[ti], par, contigs, ti, tiles[ti], sources[tile_source ? tile_source[ti] : 0], center_pow, pem, msgs[ti]);
        uvc_stage_thread_pinning(true);
    });
    for (int32_t ti = 0; ti < n_tiles; ti++) { if (rcs[ti] != 0) { msg = msgs[ti]; return rcs[ti]; } }
    const int64_t wall1 = prof_now();
    // 2. offsets of every tile in the concatenated arrays
    typedef BatchOff Off;
    std::vector<Off> off((size_t)n_tiles + 1);
    memset(&off[0], 0, sizeof(Off));
    for (int32_t ti = 0; ti < n_tiles; ti++) {
        const HostBatch & b = part[ti];
        Off o = off[ti];
        o.pos += b.n_pos; o.read += (int64_t)work[ti].kept.size(); o.frag += (int64_t)b.frags.size(); o.fam += (int64_t)b.fams.size();
        o.fragread += (int64_t)b.frag_reads.size(); o.seq += (int64_t)work[ti].n_seq; o.qual += (int64_t)work[ti].n_qual; o.cigar += (int64_t)work[ti].n_cig;
        o.cx += b.n_cx; o.ev += b.n_ev; o.fcol += b.n_fcol; o.mcol += b.n_mcol;
        off[ti + 1] = o;
        hb.n_reads_in += b.n_reads_in;
    }
    const Off & tot = off[n_tiles];
    if (tot.read > INT32_MAX || tot.frag > INT32_MAX || tot.cx > INT32_MAX || tot.ev > INT32_MAX || tot.fragread > INT32_MAX || tot.fam > INT32_MAX / 2) { msg = "batch too large: submit fewer tiles"; return UVCGPU_EINVAL; }
    hb.tiles.resize((size_t)n_tiles);
    hb.pos_tile.resize((size_t)tot.pos); hb.refsym.resize((size_t)tot.pos); hb.rtr.resize((size_t)tot.pos); hb.baq.resize((size_t)tot.pos); hb.baq2.resize((size_t)tot.pos);
    hb.reads.resize((size_t)tot.read); hb.read_raw_index.resize((size_t)tot.read); hb.rfam.resize((size_t)tot.read);
    hb.seq.resize((size_t)tot.seq); hb.qual.resize((size_t)tot.qual); hb.cigar.resize((size_t)tot.cigar);
    hb.frags.resize((size_t)tot.frag); hb.frag_reads.resize((size_t)tot.fragread); hb.fams.resize((size_t)tot.fam); hb.fam_umi.resize((size_t)tot.fam);
    hb.n_pos = tot.pos; hb.n_cx = tot.cx; hb.n_ev = tot.ev; hb.n_fcol = tot.fcol; hb.n_mcol = tot.mcol;
    hb.fchunk_frag.resize((size_t)(tot.fcol / UVC_COL_CHUNK)); hb.mchunk_fs.resize((size_t)(tot.mcol / UVC_COL_CHUNK));
    const int64_t wall2 = prof_now();
    // 3. stage B of every tile and its fragment / family records with the tile-local indices rebased, again on all cores (disjoint destination ranges)
    uvc_parallel_for(n_tiles, n_threads, [&](int32_t ti) {
        HostBatch & b = part[ti];
        const Off & o = off[ti];
        TileInfo T = b.tiles[0];
        T.pos_off = o.pos; T.read_off = o.read; T.frag_off = o.frag; T.fam_off = o.fam;
        hb.tiles[ti] = T;
        uvc_stage_thread_pinning(false);
        build_tile_b(hb, work[ti], T, o, par, ti, b.frags.size(), b.fams.size());
        uvc_stage_thread_pinning(true);
        for (size_t i = 0; i < b.frags.size(); i++) {
            FragRec G = b.frags[i];
            G.fam += (int32_t)o.fam; G.read_off += (int32_t)o.fragread; G.col_off += o.fcol;
            hb.frags[(size_t)o.frag + i] = G;
            const int64_t c1 = G.col_off / UVC_COL_CHUNK + ((int64_t)(G.hi - G.lo) + UVC_COL_CHUNK - 1) / UVC_COL_CHUNK;
            for (int64_t c = G.col_off / UVC_COL_CHUNK; c < c1; c++) { hb.fchunk_frag[(size_t)c] = (int32_t)(o.frag + (int64_t)i); }
        }
        for (size_t i = 0; i < b.frag_reads.size(); i++) { hb.frag_reads[(size_t)o.fragread + i] = b.frag_reads[i] + (int32_t)o.read; }
        for (size_t i = 0; i < b.fams.size(); i++) {
            FamRec F = b.fams[i];
            F.frag_off[0] += (int32_t)o.frag; F.frag_off[1] += (int32_t)o.frag;
            F.col_off[0] += o.mcol; F.col_off[1] += o.mcol;
            for (int strand = 0; strand < 2; strand++) { if (F.direct_frag[strand] >= 0) { F.direct_frag[strand] += (int32_t)o.frag; } }
            hb.fams[(size_t)o.fam + i] = F;
            for (int strand = 0; strand < 2; strand++) {
                if (F.direct_frag[strand] >= 0) { continue; }
                const int64_t c1 = F.col_off[strand] / UVC_COL_CHUNK + ((int64_t)(F.hi[strand] - F.lo[strand]) + UVC_COL_CHUNK - 1) / UVC_COL_CHUNK;
                for (int64_t c = F.col_off[strand] / UVC_COL_CHUNK; c < c1; c++) { hb.mchunk_fs[(size_t)c] = (int32_t)(2 * (o.fam + (int64_t)i) + strand); }
            }
            hb.fam_umi[(size_t)o.fam + i].swap(b.fam_umi[i]);
        }
        for (size_t i = 0; i < work[ti].kept.size(); i++) {
            const ReadRec & R = hb.reads[(size_t)o.read + i];
            const FamRec & F = hb.fams[(size_t)R.fam];
            ReadFam q;
            q.rend = R.rend; q.famprev_maxrend = R.famprev_maxrend; q.fambothprev_maxrend = R.fambothprev_maxrend; q.fam = R.fam; q.pad = 0;
            q.flags = (R.strand ? UVC_RF_STRAND : 0u) | ((F.duplexflag & 0x2) ? UVC_RF_DUPLEX_UMI : 0u) | ((F.n_frags[0] > 0 && F.n_frags[1] > 0) ? UVC_RF_BOTH_STRANDS : 0u);
            if (F.direct_frag[R.strand] >= 0) {
                const FragRec & G = hb.frags[(size_t)F.direct_frag[R.strand]];
                q.flags |= UVC_RF_DIRECT;
                q.col_base = G.col_off - G.lo;
            } else {
                q.col_base = F.col_off[R.strand] - F.lo[R.strand];
            }
            hb.rfam[(size_t)o.read + i] = q;
        }
        part_cache().give(b);   // hand the private copy back (its vectors keep their capacity for the next tile)
        kept_cache().give(work[ti].kept);
        work[ti] = TileWork();
    });
    if (getenv("UVC_PREP_PROFILE")) { fprintf(stderr, "prep wall ms: tiles %.1f resize %.1f concat %.1f\n", (wall1 - wall0) / 1e6, (wall2 - wall1) / 1e6, (prof_now() - wall2) / 1e6); }
    if (getenv("UVC_PREP_PROFILE")) { fprintf(stderr, "prep ms: pass1 %.1f centers %.1f pass2 %.1f sort %.1f families %.1f pack %.1f P1 %.1f (of which repeat context %.1f)\n", g_prof[0] / 1e6, g_prof[1] / 1e6, g_prof[2] / 1e6, g_prof[3] / 1e6, g_prof[4] / 1e6, g_prof[5] / 1e6, (g_prof[6] + g_prof[7]) / 1e6, g_prof[7] / 1e6); for (auto & x : g_prof) { x = 0; } }
    return 0;
}

// Families of a tile in the order of the reference's std::map<MolecularBarcode, ...> (MolecularID.hpp:52-68). The device groups by exact key
// equality and orders families by their first read in the file; this test hook restores the map order from the key of each family.
std::string uvc_families_text(const HostBatch & hb, int32_t tile_index) {
    std::string out;
    const TileInfo & T = hb.tiles[tile_index];
    struct Key { tidpos_t beg, end; std::string qname, umi; uint32_t dflag, idflag; int64_t fam; std::string umi_full; };
    std::vector<Key> keys;
    for (int64_t fi = T.fam_off; fi < T.fam_off + T.n_fams; fi++) {
        const FamRec & F = hb.fams[fi];
        // the family's first read in file order carries its non-key data (and the strings of the key)
        int32_t first = INT32_MAX;
        for (int strand = 0; strand < 2; strand++) {
            for (int32_t g = F.frag_off[strand]; g < F.frag_off[strand] + F.n_frags[strand]; g++) {
                const FragRec & G = hb.frags[g];
                for (int32_t q = G.read_off; q < G.read_off + G.n_reads; q++) { first = std::min(first, hb.frag_reads[q]); }
            }
        }
        const char *qname = hb.raw_qname(hb.reads[first].raw);
        const char *h1 = strchr(qname, '#');
        const char *h2 = (h1 ? strchr(h1 + 1, '#') : NULL);
        const size_t qlen = strlen(qname);
        const char *umi_beg = (h1 ? h1 + 1 : qname + qlen), *umi_end = (h2 ? h2 : qname + qlen);
        Key k;
        k.fam = fi; k.dflag = F.duplexflag; k.idflag = F.dedup_idflag;
        k.umi_full = ((F.duplexflag & 0x1) ? std::string(umi_beg, umi_end) : std::string());
        const tidpos_t begpair(F.beg_tid, F.beg_pos), endpair(F.end_tid, F.end_pos);
        k.beg = tidpos_t(-1, -1); k.end = tidpos_t(-1, -1);
        if (0x3 == (0x3 & k.idflag)) { k.beg = std::min(begpair, endpair); k.end = std::max(begpair, endpair); }
        else if (0x1 & k.idflag) { k.beg = begpair; }
        else if (0x2 & k.idflag) { k.end = endpair; }
        if (0x4 & k.idflag) { k.qname = qname; }
        if (0x8 & k.idflag) { k.umi = k.umi_full; }
        keys.push_back(k);
    }
    std::sort(keys.begin(), keys.end(), [](const Key & a, const Key & b) {
        if (a.beg != b.beg) { return a.beg < b.beg; }
        if (a.end != b.end) { return a.end < b.end; }
        if (a.qname != b.qname) { return a.qname < b.qname; }
        if (a.umi != b.umi) { return a.umi < b.umi; }
        if (a.dflag != b.dflag) { return a.dflag < b.dflag; }
        return a.idflag < b.idflag;
    });
    for (const Key & k : keys) {
        const FamRec & F = hb.fams[k.fam];
        out += "F\t" + std::to_string(F.beg_tid) + "\t" + std::to_string(F.beg_pos) + "\t" + std::to_string(F.end_tid) + "\t" + std::to_string(F.end_pos)
            + "\t" + std::to_string(F.duplexflag) + "\t" + std::to_string(F.dedup_idflag) + "\t" + k.umi_full
            + "\t" + std::to_string(F.n_frags[0]) + "\t" + std::to_string(F.n_frags[1]) + "\n";
        for (int strand = 0; strand < 2; strand++) {
            for (int32_t g = F.frag_off[strand]; g < F.frag_off[strand] + F.n_frags[strand]; g++) {
                const FragRec & G = hb.frags[g];
                out += "f\t" + std::to_string(strand);
                for (int32_t q = G.read_off; q < G.read_off + G.n_reads; q++) {
                    const ReadRec & R = hb.reads[hb.frag_reads[q]];
                    out += std::string("\t") + hb.raw_qname(R.raw) + "/" + std::to_string(R.flag) + "/" + std::to_string(R.pos);
                }
                out += "\n";
            }
        }
    }
    return out;
}
