// prep_core.cuh - per-work-item bodies of the device staging kernels: stage P0 (read filter, fragment-end histograms, dedup centres,
// family key, family/fragment segmentation) and stage P1 (reference repeat context, BAQ prefix sums).
//
// P0 restates reference grouping.cpp:347-442, 608-997 and MolecularID.hpp:20-69: the reference groups reads with
// std::map<MolecularBarcode, ...>; here the kept reads of a batch are labelled by an exact-match hash table on the family key and then
// SORTED (stable LSD radix sort, engine.cpp) by (family label, strand, qname hash), which yields the same segmentation: families, their two
// strands, the fragments of a strand in qname-hash order, the reads of a fragment in file order. P1 restates main.hpp:699-721, 794-874 and
// main.cpp:400-429. Reference quirks that change results are kept on purpose and marked QUIRK.
//
// Every body is a functor over a work-item index, compiled for the device by nvcc and - for the CPU-only unit tests of the host logic
// (tests/ only, never shipped as a fallback) - as ordinary C++ run in serial loops.
#ifndef UVC_PREP_CORE_CUH_INCLUDED
#define UVC_PREP_CORE_CUH_INCLUDED

#include "batch.h"
#include "kernels_core.cuh"

namespace uvc {

#define UVC_ARRPOS_MARGIN UVC_MAX_INSERT_SIZE   // grouping.cpp:22
#define UVC_ARRPOS_OUTER_RANGE 10               // grouping.cpp:23
#define UVC_ARRPOS_INNER_RANGE 3                // grouping.cpp:24

UVC_HD void atomic_min32(int32_t *p, int32_t v) {
#if defined(__CUDA_ARCH__)
    atomicMin(p, v);
#else
    if (v < *p) { *p = v; }
#endif
}
UVC_HD void atomic_max32(int32_t *p, int32_t v) {
#if defined(__CUDA_ARCH__)
    atomicMax(p, v);
#else
    if (v > *p) { *p = v; }
#endif
}
UVC_HD void atomic_min64(int64_t *p, int64_t v) {
#if defined(__CUDA_ARCH__)
    atomicMin((long long*)p, (long long)v);
#else
    if (v < *p) { *p = v; }
#endif
}
UVC_HD void atomic_max_u64(uint64_t *p, uint64_t v) {
#if defined(__CUDA_ARCH__)
    atomicMax((unsigned long long*)p, (unsigned long long)v);
#else
    if (v > *p) { *p = v; }
#endif
}
UVC_HD int32_t atomic_cas32(int32_t *p, int32_t expected, int32_t desired) {
#if defined(__CUDA_ARCH__)
    return atomicCAS(p, expected, desired);
#else
    const int32_t old = *p; if (old == expected) { *p = desired; } return old;
#endif
}

// What stage P0a derives from one raw BAM record (uvcgpu_reads_soa entry), independent of any tile.
struct RawDer {
    int32_t rend;               // bam_endpos
    int32_t isize;              // NORM_INSERT_SIZE (common.hpp:75)
    int32_t simple, m_qoff, n_ev;   // CIGAR shape (see ReadRec)
    int32_t qlen;               // strlen(qname)
    int32_t umi_off, umi_len;   // UMI substring of the name (grouping.cpp:768-779); umi_len = 0: none
    int32_t umi_found, duplex_found;
    uint64_t qhash2;            // strhash(qname, 17): fragment identity (grouping.cpp:766, Hash.hpp:9-15)
    uint64_t namehash, umihash; // 64-bit hashes that bucket names / UMIs; equality is always confirmed on the bytes
};

// What the first pass learns about a (tile, record) pair (grouping.cpp:666-695), kept for the second pass
struct PairInfo { int32_t tBeg, tEnd; int32_t flags; };   // flags: bit0 keep, bit1-2 class c = isrc * 2 + isr2, bit3 touches the tile
#define UVC_PAIR_KEEP 1
#define UVC_PAIR_TOUCH 8

struct PrepTile {               // host-built per-tile constants of the staging kernels
    int64_t pair_off;           // first (tile, record) pair of the tile; the tile's records are raw[raw_begin, raw_begin + n_in)
    int64_t raw_begin; int32_t n_in;
    int32_t fetch_size;         // tile length + (ARRPOS_MARGIN + ARRPOS_OUTER_RANGE) * 2 (grouping.cpp:646)
    int64_t hist_off;           // into hist: 8 arrays of fetch_size counters (beg_cnt[4], end_cnt[4])
    int64_t psum_off;           // into psum: 4 arrays of fetch_size + 1 prefix sums
    int64_t set_off; int32_t set_mask;   // the tile's name set: open addressing over set_mask + 1 slots
    const char *contig;         // upper-cased bases of the tile's contig on the compute side; NULL: reference not available (all 'n', main.cpp:57-59)
};

struct KeptRec {                // one read that passed the filter of its tile (file order inside a tile)
    int32_t raw, tile;
    int32_t beg_tid, beg_pos, end_tid, end_pos;       // the read's own MolecularBarcode ends (grouping.cpp:927-930)
    int32_t kbeg_tid, kbeg_pos, kend_tid, kend_pos;   // key ends after MolecularBarcode::createKey (MolecularID.hpp:20-51)
    uint32_t dflag, idflag;
    int32_t strand;
    int32_t slot, label;        // hash-table slot of its family; label = smallest kept index of the family (its first read in file order)
    int32_t frag, fam;          // batch-global fragment / family index
};

struct PrepTotals {             // downloaded once per batch
    int64_t n_kept, n_frags, n_fams, n_fs, n_cx, n_ev, n_fcol, n_mcol, n_pos, n_qual;
    int64_t n_list[2];          // lengths of the two position lists (kernels_core.cuh: tile_need_range)
    int32_t err;                // 1: a family strand has more than 65535 fragments
    int32_t pad;
};

struct PrepView {
    uvcgpu_params par;
    double center_pow[4];
    int32_t pem;
    int32_t n_tiles;
    int64_t n_raw, n_pairs;
    // raw records (device copies of the caller's SoA slices)
    const int32_t *pos, *mpos, *isize, *mtid, *l_qseq, *n_cigar, *nm;
    const uint16_t *flag; const uint8_t *mapq;
    const uint64_t *seq_off, *qual_off, *cigar_off, *qname_off;
    const uint8_t *seq; uint8_t *qual; const uint32_t *cigar; const char *qname;
    RawDer *rd;
    const uvcgpu_tile *utiles;
    const PrepTile *pt;
    PairInfo *pair;
    int32_t *hist; const int64_t *psum; int32_t *nameset;   // psum: exclusive prefix sum over the concatenated (tile, class) arrays of begin + end counts
    int32_t *keepflag; const int64_t *keepscan;       // [n_pairs + 1]
    int32_t *tile_bam_beg, *tile_bam_end, *tile_pcr, *tile_span;
    TileInfo *tiles;
    int64_t *tile_len;          // [n_tiles] extended length of every tile
    const int64_t *tile_npos;   // [n_tiles + 1] exclusive scan of tile_len
    int64_t *list_off[2];       // [n_tiles + 1] per position list: first entry of each tile
    int32_t *list_tile[2];      // [n_list / 32] per position list: tile of each chunk
    // kept reads
    int64_t n_kept;
    KeptRec *kept;
    int32_t *famtab; int32_t famtab_mask;
    uint64_t *key_a, *key_b; int32_t *val_a, *val_b;  // sort buffers
    const int32_t *ord2;        // kept reads ordered by (label, strand, qhash2, file order)
    const int32_t *ord1;        // kept reads ordered by (label, strand, file order)
    int32_t *flag_fam, *flag_fs, *flag_frag;          // boundary flags over ord2
    const int64_t *scan_fam, *scan_fs, *scan_frag;    // exclusive scans of the flags (n_kept + 1)
    int32_t *frag_start, *fs_start, *fam_start;       // first ord2 index of every fragment / (family, strand) / family
    int32_t *cx_len, *ev_len, *ql_len; const int64_t *scan_cx, *scan_ev, *scan_ql;
    int64_t *fcol_len, *mcol_len; const int64_t *scan_fcol, *scan_mcol;
    PrepTotals *totals;
    // outputs (the arrays of BatchView)
    ReadRec *reads; FragRec *frags; FamRec *fams; int32_t *frag_reads; ReadFam *rfam;
    int32_t *fchunk_frag, *mchunk_fs;
    // stage P1
    int64_t n_pos;
    int32_t *pos_tile; uint8_t *refsym; uvcgpu_rtr *rtr; int32_t *baq, *baq2;
    uint64_t *p1_cand;          // per position: best STR and any-TR unit and match-run length as seen from that position, packed
    int32_t *p1_jump;           // per position: how far the reference's skip-walk advances from here
    uint8_t *p1_visited;
    uint64_t *p1_str_key, *p1_any_key;
    const int32_t *slip_tab;
};

UVC_HD uint64_t mix64(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }
UVC_HD uint64_t bytes_hash(const char *s, int32_t n) {
    uint64_t h = 1469598103934665603ULL;
    for (int32_t i = 0; i < n; i++) { h = (h ^ (uint64_t)(uint8_t)s[i]) * 1099511628211ULL; }
    return mix64(h ^ (uint64_t)n);
}
UVC_HD bool bytes_equal(const char *a, const char *b, int32_t n) { for (int32_t i = 0; i < n; i++) { if (a[i] != b[i]) { return false; } } return true; }

// tile of a (tile, record) pair: largest t with pt[t].pair_off <= pi
UVC_HD int32_t pair_tile(const PrepView & q, int64_t pi) {
    int32_t a = 0, b = q.n_tiles;
    while (b - a > 1) { const int32_t m = (a + b) >> 1; if (q.pt[m].pair_off <= pi) { a = m; } else { b = m; } }
    return a;
}

// ------------------------------------------------------------------------------------------------ P0z: offsets of directly uploaded sources
// When the caller's SoA buffers are page-locked, the record arrays of every source go to the device straight from them (no staging copy on the
// host); the four offset arrays then still count from the start of their own source's byte arrays. One thread per entry adds its source's
// displacement in the concatenated arrays. tab: [n_src + 1] first raw index of each source, then [n_src][4] displacements (seq, qual, cigar, name).
struct P0zRebase {
    uint64_t *so, *qo, *co, *no; const int64_t *tab; int32_t n_src; int64_t n_raw;
    UVC_HD void operator()(int64_t r) const {
        int32_t s = 0;
        if (r >= n_raw) { s = n_src - 1; while (s > 0 && tab[s] == tab[s + 1]) { s--; } }          // the end entry belongs to the last source that has records
        else { int32_t a = 0, b = n_src; while (b - a > 1) { const int32_t m = (a + b) >> 1; if (tab[m] <= r) { a = m; } else { b = m; } } s = a; }
        const int64_t *d = tab + (n_src + 1) + 4 * (int64_t)s;
        so[r] += (uint64_t)d[0]; qo[r] += (uint64_t)d[1]; co[r] += (uint64_t)d[2]; no[r] += (uint64_t)d[3];
    }
};

// ------------------------------------------------------------------------------------------------ P0a: one thread per raw record
struct P0aRaw {
    PrepView q;
    UVC_HD void operator()(int64_t i) const {
        RawDer d;
        const uint32_t *cigar = q.cigar + q.cigar_off[i];
        const int32_t n_cigar = q.n_cigar[i];
        const uint16_t flag = q.flag[i];
        // bam_endpos: reference length of the alignment, 1 if it is zero or the read is unmapped
        int64_t rlen = 0;
        int32_t n_ev = 0;
        if (!(flag & 0x4)) {
            for (int32_t k = 0; k < n_cigar; k++) {
                const int op = cig_op(cigar[k]);
                if (is_match_op(op) || op == UVC_CDEL || op == UVC_CREF_SKIP) { rlen += cig_len(cigar[k]); }
            }
        }
        for (int32_t k = 0; k < n_cigar; k++) { const int op = cig_op(cigar[k]); if (op == UVC_CINS || op == UVC_CDEL) { n_ev++; } }
        if (0 == rlen) { rlen = 1; }
        d.rend = (int32_t)(q.pos[i] + rlen);
        d.isize = (iabs(q.isize[i]) >= UVC_MAX_INSERT_SIZE ? 0 : q.isize[i]);
        {   // simple = [S|H|P]* (M|=|X) [S|H|P]*
            int32_t lead = 0, a = 0, b = n_cigar;
            while (a < b && (cig_op(cigar[a]) == UVC_CSOFT_CLIP || cig_op(cigar[a]) == UVC_CHARD_CLIP || cig_op(cigar[a]) == UVC_CPAD)) {
                if (cig_op(cigar[a]) == UVC_CSOFT_CLIP) { lead += cig_len(cigar[a]); }
                a++;
            }
            while (b > a && (cig_op(cigar[b - 1]) == UVC_CSOFT_CLIP || cig_op(cigar[b - 1]) == UVC_CHARD_CLIP || cig_op(cigar[b - 1]) == UVC_CPAD)) { b--; }
            d.simple = (((b - a == 1) && is_match_op(cig_op(cigar[a]))) ? 1 : 0);
            d.m_qoff = lead;
            d.n_ev = (d.simple ? 0 : n_ev);
        }
        // one pass over the name: its length, the first two '#' and the order-defining hash strhash(qname, 17)
        const char *qname = q.qname + q.qname_off[i];
        int32_t qlen = 0, h1 = -1, h2 = -1;
        uint64_t qh2 = 0;
        for (; qname[qlen]; qlen++) {
            const char c = qname[qlen];
            qh2 = qh2 * 17 + (uint64_t)(int64_t)c;
            if ('#' == c) { if (h1 < 0) { h1 = qlen; } else if (h2 < 0) { h2 = qlen; } }
        }
        d.qlen = qlen; d.qhash2 = qh2;
        d.namehash = bytes_hash(qname, qlen);
        const int32_t umi_beg = (h1 >= 0 ? h1 + 1 : qlen), umi_end = (h2 >= 0 ? h2 : qlen);
        d.umi_found = (((umi_beg + 1 < umi_end) && (1 /* MOLECULE_TAG_NONE */ != q.par.molecule_tag)) ? 1 : 0);
        d.duplex_found = 0;
        d.umi_off = umi_beg; d.umi_len = (d.umi_found ? umi_end - umi_beg : 0);
        if (d.umi_found) {
            const int32_t ulen = umi_end - umi_beg, half = (ulen - 1) / 2;
            if ((ulen % 2 == 1) && ('+' == qname[umi_beg + half]) && (!q.par.disable_duplex)) { d.duplex_found = 1; }
        }
        d.umihash = (d.umi_found ? bytes_hash(qname + umi_beg, d.umi_len) : 0);
        q.rd[i] = d;
    }
};

// grouping.cpp:347-415 (fill_isrc_isr2_beg_end_with_aln); returns true if the read is kept (NOT_FILTERED)
UVC_HD bool classify_read(const PrepView & q, int64_t i, const RawDer & d, int32_t fetch_tbeg, int32_t fetch_tend, bool end2end,
        bool & isrc, bool & isr2, int32_t & tBeg, int32_t & tEnd) {
    const uvcgpu_params & par = q.par;
    const uint16_t flag = q.flag[i];
    if (flag & 0x4) { return false; }
    if (flag & 0x900) { return false; }
    // QUIRK: the reference's call sites pass (min_aln_len, min_mapqual) in swapped order (grouping.cpp:676-677 vs :351-352)
    const int32_t min_mapqual = par.kept_aln_min_aln_len;
    const int32_t min_aln_len = par.kept_aln_min_mapqual;
    const int32_t pos = q.pos[i];
    if ((int32_t)q.mapq[i] < min_mapqual) { return false; }
    if ((d.rend - pos) < min_aln_len) { return false; }
    if (0 == d.isize) {
        if (par.kept_aln_is_zero_isize_discarded) { return false; }
    } else {
        if (iabs(d.isize) < par.kept_aln_min_isize) { return false; }
        if (iabs(d.isize) > par.kept_aln_max_isize) { return false; }
    }
    isrc = ((flag & 0x10) == 0x10);
    isr2 = ((flag & 0x80) == 0x80 && (flag & 0x1) == 0x1);
    if (!q.pem) { isr2 = false; }
    const int32_t begpos = pos, endpos = d.rend - 1;
    if ((!q.pem) || ((flag & 0x1) == 0) || (flag & 0x8) || (0 == d.isize) || (iabs(d.isize) >= UVC_ARRPOS_MARGIN)) {
        tBeg = (isrc ? endpos : begpos);
        tEnd = (isrc ? begpos : endpos);
    } else {
        const int32_t b1 = tmin(begpos, q.mpos[i]);
        const int32_t e1 = b1 + iabs(d.isize) - 1;
        const bool strand = read_strand(flag);
        tBeg = (strand ? e1 : b1);
        tEnd = (strand ? b1 : e1);
    }
    const int32_t ob = tmin(tBeg, tEnd), oe = tmax(tBeg, tEnd);
    if (ob + (UVC_ARRPOS_MARGIN - UVC_ARRPOS_OUTER_RANGE) <= fetch_tbeg || fetch_tend - 1 + (UVC_ARRPOS_MARGIN - UVC_ARRPOS_OUTER_RANGE) <= oe) { return false; }
    if (end2end && !(ob <= fetch_tbeg && oe >= fetch_tend)) { return false; }
    return true;
}

// the tile's name set (grouping.cpp:648 visited_qnames): open addressing over tile-local record indices, names compared as strings
UVC_HD void nameset_insert(const PrepView & q, const PrepTile & P, int32_t j_new) {
    const int64_t i_new = P.raw_begin + j_new;
    const RawDer & dn = q.rd[i_new];
    const char *name = q.qname + q.qname_off[i_new];
    int32_t *tab = q.nameset + P.set_off;
    for (uint32_t k = (uint32_t)dn.namehash & (uint32_t)P.set_mask;; k = (k + 1) & (uint32_t)P.set_mask) {
        const int32_t j = atomic_cas32(&tab[k], -1, j_new);
        if (j < 0) { return; }
        const RawDer & dj = q.rd[P.raw_begin + j];
        if (dj.namehash == dn.namehash && dj.qlen == dn.qlen && bytes_equal(q.qname + q.qname_off[P.raw_begin + j], name, dn.qlen)) { return; }
    }
}
UVC_HD bool nameset_find(const PrepView & q, const PrepTile & P, int64_t i) {
    const RawDer & dn = q.rd[i];
    const char *name = q.qname + q.qname_off[i];
    const int32_t *tab = q.nameset + P.set_off;
    for (uint32_t k = (uint32_t)dn.namehash & (uint32_t)P.set_mask;; k = (k + 1) & (uint32_t)P.set_mask) {
        const int32_t j = tab[k];
        if (j < 0) { return false; }
        const RawDer & dj = q.rd[P.raw_begin + j];
        if (dj.namehash == dn.namehash && dj.qlen == dn.qlen && bytes_equal(q.qname + q.qname_off[P.raw_begin + j], name, dn.qlen)) { return true; }
    }
}

// ------------------------------------------------------------------------------------------------ P0b: one thread per (tile, record) pair
// pass 1 (grouping.cpp:666-695): classification, fragment-end histograms per class, names of the fragments that touch the tile
struct P0bPair {
    PrepView q;
    UVC_HD void operator()(int64_t pi) const {
        const int32_t t = pair_tile(q, pi);
        const PrepTile & P = q.pt[t];
        const uvcgpu_tile & ut = q.utiles[t];
        const int32_t j = (int32_t)(pi - P.pair_off);
        const int64_t i = P.raw_begin + j;
        const RawDer & d = q.rd[i];
        PairInfo I; I.tBeg = 0; I.tEnd = 0; I.flags = 0;
        bool isrc = false, isr2 = false; int32_t tBeg = 0, tEnd = 0;
        if (classify_read(q, i, d, ut.beg_pos, ut.end_pos, (ut.region_flag & 0x1) != 0, isrc, isr2, tBeg, tEnd)) {
            const int c = (isrc ? 2 : 0) + (isr2 ? 1 : 0);
            I.tBeg = tBeg; I.tEnd = tEnd; I.flags = UVC_PAIR_KEEP | (c << 1);
            const int32_t fs = P.fetch_size;
            const int32_t bi = tBeg + UVC_ARRPOS_MARGIN - ut.beg_pos, ei = tEnd + UVC_ARRPOS_MARGIN - ut.beg_pos;
            int32_t *h = q.hist + P.hist_off;
            if (bi >= 0 && bi < fs) { atomic_add(&h[(int64_t)c * fs + bi], 1); }
            if (ei >= 0 && ei < fs) { atomic_add(&h[(int64_t)(4 + c) * fs + ei], 1); }
            const int32_t lo = tmin(tBeg, tEnd), hi = tmax(tBeg, tEnd) + 2;
            if (!((hi <= ut.beg_pos) || (ut.end_pos <= lo))) { I.flags |= UVC_PAIR_TOUCH; nameset_insert(q, P, j); }
        }
        q.pair[pi] = I;
    }
};

// ------------------------------------------------------------------------------------------------ P0c: one thread per histogram slot
// border_psum (grouping.cpp:697-708): prefix sums of begin + end counts. The sums of all (tile, class) arrays are taken by ONE prefix sum over
// their concatenation (only differences inside one array are ever used): this kernel writes its input, begin + end count per slot.
struct P0cComb {
    PrepView q; int32_t *comb;
    UVC_HD void operator()(int64_t k) const {
        // slot k of the concatenation: tile t (by psum_off), class c, index i
        int32_t a = 0, b = q.n_tiles;
        while (b - a > 1) { const int32_t m = (a + b) >> 1; if (q.pt[m].psum_off <= k) { a = m; } else { b = m; } }
        const PrepTile & P = q.pt[a];
        const int32_t fs = P.fetch_size;
        const int64_t o = k - P.psum_off;
        const int32_t c = (int32_t)(o / (fs + 1)), i = (int32_t)(o % (fs + 1));
        const int32_t *h = q.hist + P.hist_off;
        comb[k] = (i < fs ? h[(int64_t)c * fs + i] + h[(int64_t)(4 + c) * fs + i] : 0);
    }
};

// ------------------------------------------------------------------------------------------------ P0d: one thread per pair -> keep flag
struct P0dKeep {
    PrepView q;
    UVC_HD void operator()(int64_t pi) const {
        const int32_t t = pair_tile(q, pi);
        const PrepTile & P = q.pt[t];
        const uvcgpu_tile & ut = q.utiles[t];
        const int64_t i = P.raw_begin + (pi - P.pair_off);
        const PairInfo I = q.pair[pi];
        int32_t keep = 0;
        if (I.flags & UVC_PAIR_KEEP) {
            // (the order of the reference's tests - position window, name set, classification - does not matter: all must hold)
            const bool in_window = !(q.pos[i] < nnminus(ut.beg_pos, UVC_MAX_INSERT_SIZE + 1) || q.rd[i].rend > (ut.end_pos + UVC_MAX_INSERT_SIZE + 1));
            if (in_window && ((I.flags & UVC_PAIR_TOUCH) || nameset_find(q, P, i))) { keep = 1; }
        }
        q.keepflag[pi] = keep;
    }
};

// poscounter_to_pos2pcenter (grouping.cpp:422-442) evaluated on demand for one position: the local maximum within +-3 that the end
// position snaps to. The reference fills the whole array, which starts as all zeros (its inicount copy), hence 0 outside the loop range.
UVC_HD int32_t center_at(const int32_t *cnt, int32_t n, int32_t lo, const double *center_pow) {
    if (lo < UVC_ARRPOS_INNER_RANGE || lo >= n - UVC_ARRPOS_INNER_RANGE) { return 0; }
    const int32_t lo_cnt = cnt[lo];
    int32_t center = lo;
    int32_t max_cnt = lo_cnt;
    for (int32_t hi = lo - UVC_ARRPOS_INNER_RANGE; hi < lo + UVC_ARRPOS_INNER_RANGE + 1; hi++) {
        const int32_t hi_cnt = cnt[hi];
        const int d = iabs(lo - hi);
        if ((hi_cnt > max_cnt) && ((double)(hi_cnt + 1) > (double)(lo_cnt + 1) * center_pow[d])) {
            center = hi;
            max_cnt = hi_cnt;
        }
    }
    return center;
}

UVC_HD bool tidpos_less(int32_t at, int32_t ap, int32_t bt, int32_t bp) { return (at < bt) || (at == bt && ap < bp); }

// ------------------------------------------------------------------------------------------------ P0e: one thread per pair -> kept read
// pass 2 (grouping.cpp:731-977): dedup centres, amplicon inference, dedup_idflag, MolecularBarcode key of every kept read
struct P0eKept {
    PrepView q;
    UVC_HD void operator()(int64_t pi) const {
        if (!q.keepflag[pi]) { return; }
        const int64_t k = q.keepscan[pi];
        const int32_t t = pair_tile(q, pi);
        const PrepTile & P = q.pt[t];
        const uvcgpu_tile & ut = q.utiles[t];
        const int64_t i = P.raw_begin + (pi - P.pair_off);
        const PairInfo I = q.pair[pi];
        const RawDer & d = q.rd[i];
        const uvcgpu_params & par = q.par;
        const int c = (I.flags >> 1) & 3;
        const int32_t fs = P.fetch_size;
        const int32_t fetch_tbeg = ut.beg_pos;
        const int32_t *beg_cnt = q.hist + P.hist_off + (int64_t)c * fs, *end_cnt = q.hist + P.hist_off + (int64_t)(4 + c) * fs;
        const int64_t *psum = q.psum + P.psum_off + (int64_t)c * (fs + 1);
        const int32_t pos = q.pos[i], mpos = q.mpos[i];
        const uint16_t flag = q.flag[i];
#if defined(__CUDA_ARCH__)
        {   // thousands of reads of a tile update the same three words: one reduction per group of lanes with the same tile instead of one per read
            const uint32_t peers = __match_any_sync(__activemask(), t);
            const int32_t mn = __reduce_min_sync(peers, pos), mx = __reduce_max_sync(peers, d.rend), sp = __reduce_max_sync(peers, d.rend - pos);
            if ((threadIdx.x & 31) == (uint32_t)(__ffs((int)peers) - 1)) { atomic_min32(&q.tile_bam_beg[t], mn); atomic_max32(&q.tile_bam_end[t], mx); atomic_max32(&q.tile_span[t], sp); }
        }
#else
        atomic_min32(&q.tile_bam_beg[t], pos);
        atomic_max32(&q.tile_bam_end[t], d.rend);
        atomic_max32(&q.tile_span[t], d.rend - pos);
#endif
        const int32_t beg1 = I.tBeg + UVC_ARRPOS_MARGIN - fetch_tbeg, end1 = I.tEnd + UVC_ARRPOS_MARGIN - fetch_tbeg;
        const int32_t beg2 = center_at(beg_cnt, fs, beg1, q.center_pow), end2 = center_at(end_cnt, fs, end1, q.center_pow);
        const int64_t beg2count = beg_cnt[beg2], end2count = end_cnt[end2];
        const int32_t insL = tmin(beg2 + 6, end2);
        const int32_t insR = (int32_t)tmax((int64_t)beg2, (int64_t)(end2 > 6 ? end2 - 6 : 0));
        const int64_t tot = psum[insR] - psum[insL];
        const double begratio = (double)(beg2count * (insR - insL) + 1) / (double)(tot + (insR - insL) + 1);
        const double endratio = (double)(end2count * (insR - insL) + 1) / (double)(tot + (insR - insL) + 1);
        const bool beg_amp = (begratio > par.dedup_amplicon_border_to_insert_cov_weak_avgDP_ratio
                && ((double)beg2count >= par.dedup_amplicon_border_weak_minDP) && ((double)beg2count >= (double)tot * par.dedup_amplicon_border_to_insert_cov_weak_totDP_ratio));
        const bool end_amp = (endratio > par.dedup_amplicon_border_to_insert_cov_weak_avgDP_ratio
                && ((double)end2count >= par.dedup_amplicon_border_weak_minDP) && ((double)end2count >= (double)tot * par.dedup_amplicon_border_to_insert_cov_weak_totDP_ratio));
        const bool beg_strong = (begratio > par.dedup_amplicon_border_to_insert_cov_strong_avgDP_ratio
                && ((double)beg2count >= par.dedup_amplicon_border_strong_minDP) && ((double)beg2count >= (double)tot * par.dedup_amplicon_border_to_insert_cov_strong_totDP_ratio));
        const bool end_strong = (endratio > par.dedup_amplicon_border_to_insert_cov_strong_avgDP_ratio
                && ((double)end2count >= par.dedup_amplicon_border_strong_minDP) && ((double)end2count >= (double)tot * par.dedup_amplicon_border_to_insert_cov_strong_totDP_ratio));
        const bool assay_amplicon = (beg_strong || end_strong || (beg_amp && end_amp));
        if (assay_amplicon) { atomic_add(&q.tile_pcr[t], 1); }
        uint32_t idflag = 0;
        if (par.dedup_flag != 0) {
            idflag = par.dedup_flag;
        } else if (d.umi_found) {
            if (beg_strong && end_amp && (double)beg2count > (double)end2count * par.dedup_amplicon_end2end_ratio) { idflag = 0x9; }
            else if (end_strong && beg_amp && (double)end2count > (double)beg2count * par.dedup_amplicon_end2end_ratio) { idflag = 0xA; }
            else { idflag = 0xB; }
        } else if (assay_amplicon) {
            idflag = 0x7;
        } else {
            idflag = 0x3;
        }
        const bool preserved = ((flag & 0x1) && (!(flag & 0x4)) && (!(flag & 0x8)) && (iabs(d.isize) >= (UVC_MAX_INSERT_SIZE * 3 / 4) || d.isize == 0));
        KeptRec K;
        K.raw = (int32_t)i; K.tile = t;
        K.beg_tid = ut.tid;
        K.end_tid = (((flag & 0x1) && !(flag & 0x8)) ? q.mtid[i] : (INT32_MAX - 1));
        K.beg_pos = (preserved ? pos : (beg2 - UVC_ARRPOS_MARGIN + fetch_tbeg));
        K.end_pos = (preserved ? mpos : (end2 - UVC_ARRPOS_MARGIN + fetch_tbeg));
        // MolecularBarcode::createKey (MolecularID.hpp:20-51)
        K.kbeg_tid = -1; K.kbeg_pos = -1; K.kend_tid = -1; K.kend_pos = -1;
        if (0x3 == (0x3 & idflag)) {
            const bool b_first = !tidpos_less(K.end_tid, K.end_pos, K.beg_tid, K.beg_pos);   // min / max of the two (tid, pos) pairs
            K.kbeg_tid = (b_first ? K.beg_tid : K.end_tid); K.kbeg_pos = (b_first ? K.beg_pos : K.end_pos);
            K.kend_tid = (b_first ? K.end_tid : K.beg_tid); K.kend_pos = (b_first ? K.end_pos : K.beg_pos);
        } else if (0x1 & idflag) { K.kbeg_tid = K.beg_tid; K.kbeg_pos = K.beg_pos; }
        else if (0x2 & idflag) { K.kend_tid = K.end_tid; K.kend_pos = K.end_pos; }
        K.dflag = (d.umi_found ? 0x1u : 0u) + (d.duplex_found ? 0x2u : 0u) + (assay_amplicon ? 0x4u : 0u) + (preserved ? 0x8u : 0u);
        K.idflag = idflag;
        K.strand = read_strand(flag);
        K.slot = -1; K.label = -1; K.frag = -1; K.fam = -1;
        q.kept[k] = K;
    }
};

// ------------------------------------------------------------------------------------------------ P0t: one thread per tile
struct P0tTile {
    PrepView q;
    UVC_HD void operator()(int64_t t) const {
        const uvcgpu_tile & ut = q.utiles[t];
        const PrepTile & P = q.pt[t];
        TileInfo T;
        T.tid = ut.tid; T.beg_pos = ut.beg_pos; T.end_pos = ut.end_pos; T.region_flag = ut.region_flag;
        T.prev_tid = ut.prev_tid; T.prev_beg_pos = ut.prev_beg_pos; T.prev_end_pos = ut.prev_end_pos;
        T.ext_beg = 0; T.ext_end = 0; T.rpos_inclu_beg = 0; T.rpos_exclu_end = 0;
        T.pos_off = 0; T.frag_off = 0; T.n_frags = 0; T.fam_off = 0; T.n_fams = 0;
        const int64_t k0 = q.keepscan[P.pair_off], k1 = q.keepscan[P.pair_off + P.n_in];
        T.read_off = k0; T.n_reads = (int32_t)(k1 - k0);
        T.num_passed = k1 - k0; T.num_pcrpassed = q.tile_pcr[t];
        T.bam_inclu_beg = q.tile_bam_beg[t]; T.bam_exclu_end = q.tile_bam_end[t];
        T.is_amplicon_inferred = !((T.num_pcrpassed) * 2 <= T.num_passed);
        T.max_read_span = q.tile_span[t];
        T.skipped = (k1 == k0 ? 1 : 0);
        if (T.skipped) { T.bam_inclu_beg = INT32_MAX; T.bam_exclu_end = 0; }
        if (!T.skipped) {
            T.rpos_inclu_beg = tmax(ut.beg_pos, T.bam_inclu_beg);
            T.rpos_exclu_end = tmin(ut.end_pos, T.bam_exclu_end);
            T.ext_beg = nnminus(tmin(ut.beg_pos, T.bam_inclu_beg), UVC_MAX_STR_N_BASES);
            const int64_t e = (int64_t)tmax(ut.end_pos, T.bam_exclu_end) + UVC_MAX_STR_N_BASES;
            T.ext_end = (int32_t)(e < (int64_t)ut.contig_len ? e : (int64_t)ut.contig_len) + 1;
        }
        q.tiles[t] = T;
        q.tile_len[t] = T.ext_end - T.ext_beg;
    }
};
struct P0tTileOff {     // after the scan of tile_len; fragment / family ranges of a tile are found by minimum and count
    PrepView q;
    UVC_HD void operator()(int64_t t) const {
        q.tiles[t].pos_off = q.tile_npos[t]; q.tiles[t].frag_off = INT64_MAX; q.tiles[t].fam_off = INT64_MAX;
        if (t == q.n_tiles - 1) { q.totals->n_pos = q.tile_npos[q.n_tiles]; }
    }
};
// The two position lists (kernels_core.cuh: tile_need_range, list_position): lengths of the tiles' runs, each padded to a multiple of 32
// (their exclusive scans are the tiles' offsets in the lists)
struct P0tNeedLen {
    PrepView q; int64_t *len0, *len1;
    UVC_HD void operator()(int64_t t) const {
        int32_t b, e;
        tile_need_range(q.par, q.tiles[t], 0, b, e);
        len0[t] = ((int64_t)(e - b) + 31) / 32 * 32;
        tile_need_range(q.par, q.tiles[t], 1, b, e);
        len1[t] = ((int64_t)(e - b) + 31) / 32 * 32;
    }
};
struct P0tNeedTot {
    PrepView q;
    UVC_HD void operator()(int64_t k) const { q.totals->n_list[k] = q.list_off[k][q.n_tiles]; }
};
struct P0tNeedChunks {  // one thread per (tile, kind): the owner of each chunk of 32 list entries
    PrepView q;
    UVC_HD void operator()(int64_t i) const {
        const int kind = (int)(i & 1);
        const int32_t t = (int32_t)(i >> 1);
        int32_t b, e;
        tile_need_range(q.par, q.tiles[t], kind, b, e);
        const int64_t c0 = q.list_off[kind][t] / 32, c1 = c0 + ((int64_t)(e - b) + 31) / 32;
        for (int64_t c = c0; c < c1; c++) { q.list_tile[kind][c] = t; }
    }
};
struct P0tTileFix {     // tiles without fragments
    PrepView q;
    UVC_HD void operator()(int64_t t) const { TileInfo & T = q.tiles[t]; if (0 == T.n_frags) { T.frag_off = 0; } if (0 == T.n_fams) { T.fam_off = 0; } }
};

// ------------------------------------------------------------------------------------------------ P0f/P0g: family labels by exact-match hashing
UVC_HD bool family_key_equal(const PrepView & q, const KeptRec & a, const KeptRec & b) {
    if (a.tile != b.tile || a.kbeg_tid != b.kbeg_tid || a.kbeg_pos != b.kbeg_pos || a.kend_tid != b.kend_tid || a.kend_pos != b.kend_pos
            || a.dflag != b.dflag || a.idflag != b.idflag) { return false; }
    const RawDer & da = q.rd[a.raw], & db = q.rd[b.raw];
    if (0x4 & a.idflag) {     // key.qname
        if (da.namehash != db.namehash || da.qlen != db.qlen || !bytes_equal(q.qname + q.qname_off[a.raw], q.qname + q.qname_off[b.raw], da.qlen)) { return false; }
    }
    if (0x8 & a.idflag) {     // key.umistring (verbatim: A+B and B+A are different families, grouping.cpp:931)
        if (da.umihash != db.umihash || da.umi_len != db.umi_len
                || !bytes_equal(q.qname + q.qname_off[a.raw] + da.umi_off, q.qname + q.qname_off[b.raw] + db.umi_off, da.umi_len)) { return false; }
    }
    return true;
}
struct P0fInsert {
    PrepView q;
    UVC_HD void operator()(int64_t k) const {
        KeptRec & K = q.kept[k];
        const RawDer & d = q.rd[K.raw];
        uint64_t h = mix64((uint64_t)(uint32_t)K.tile * 0x9E3779B97F4A7C15ULL + (uint64_t)(uint32_t)K.kbeg_pos);
        h = mix64(h ^ ((uint64_t)(uint32_t)K.kend_pos << 32 | (uint32_t)K.kbeg_tid) ^ ((uint64_t)K.kend_tid << 17));
        h = mix64(h ^ ((uint64_t)K.dflag << 8) ^ K.idflag ^ ((0x4 & K.idflag) ? d.namehash : 0) ^ ((0x8 & K.idflag) ? (d.umihash * 0x100000001B3ULL) : 0));
        for (uint32_t s = (uint32_t)h & (uint32_t)q.famtab_mask;; s = (s + 1) & (uint32_t)q.famtab_mask) {
            const int32_t old = atomic_cas32(&q.famtab[s], -1, (int32_t)k);
            if (old < 0) { K.slot = (int32_t)s; return; }
            if (family_key_equal(q, q.kept[old], K)) { atomic_min32(&q.famtab[s], (int32_t)k); K.slot = (int32_t)s; return; }
        }
    }
};
struct P0gLabel {       // label = first read of the family in file order; sort keys of both orders
    PrepView q;
    UVC_HD void operator()(int64_t k) const {
        KeptRec & K = q.kept[k];
        K.label = q.famtab[K.slot];
        q.key_a[k] = q.rd[K.raw].qhash2;
        q.val_a[k] = (int32_t)k;
    }
};
struct P0gKey2 {        // second-level key of a (partially sorted) order: (label, strand)
    PrepView q; const int32_t *order; uint64_t *keys;
    UVC_HD void operator()(int64_t i) const { const KeptRec & K = q.kept[order[i]]; keys[i] = ((uint64_t)(uint32_t)K.label << 1) | (uint64_t)(K.strand ? 1 : 0); }
};
struct P0gIota { int32_t *vals; UVC_HD void operator()(int64_t i) const { vals[i] = (int32_t)i; } };

// ------------------------------------------------------------------------------------------------ P0h: boundaries over the sorted order
struct P0hBounds {
    PrepView q;
    UVC_HD void operator()(int64_t i) const {
        const KeptRec & K = q.kept[q.ord2[i]];
        int32_t nf = 1, ns = 1, ng = 1;
        if (i > 0) {
            const KeptRec & L = q.kept[q.ord2[i - 1]];
            nf = (L.label != K.label ? 1 : 0);
            ns = ((nf || L.strand != K.strand) ? 1 : 0);
            ng = ((ns || q.rd[L.raw].qhash2 != q.rd[K.raw].qhash2) ? 1 : 0);
        }
        q.flag_fam[i] = nf; q.flag_fs[i] = ns; q.flag_frag[i] = ng;
    }
};
struct P0hIds {         // ids from the scans; start index of every segment
    PrepView q;
    UVC_HD void operator()(int64_t i) const {
        KeptRec & K = q.kept[q.ord2[i]];
        const int64_t fam = q.scan_fam[i] + q.flag_fam[i] - 1, frag = q.scan_frag[i] + q.flag_frag[i] - 1, fs = q.scan_fs[i] + q.flag_fs[i] - 1;
        K.fam = (int32_t)fam; K.frag = (int32_t)frag;
        q.frag_reads[i] = q.ord2[i];
        if (q.flag_fam[i]) { q.fam_start[fam] = (int32_t)i; }
        if (q.flag_fs[i]) { q.fs_start[fs] = (int32_t)i; }
        if (q.flag_frag[i]) { q.frag_start[frag] = (int32_t)i; }
        if (i == q.n_kept - 1) {
            q.totals->n_fams = fam + 1; q.totals->n_frags = frag + 1; q.totals->n_fs = fs + 1; q.totals->n_kept = q.n_kept;
            q.fam_start[fam + 1] = (int32_t)q.n_kept; q.fs_start[fs + 1] = (int32_t)q.n_kept; q.frag_start[frag + 1] = (int32_t)q.n_kept;
        }
    }
};

// ------------------------------------------------------------------------------------------------ P0i: one thread per kept read -> ReadRec
struct P0iRead {
    PrepView q;
    UVC_HD void operator()(int64_t k) const {
        const KeptRec & K = q.kept[k];
        const int64_t i = K.raw;
        const RawDer & d = q.rd[i];
        ReadRec R;
        R.pos = q.pos[i]; R.rend = d.rend; R.mpos = q.mpos[i]; R.isize = d.isize;
        R.l_qseq = q.l_qseq[i]; R.n_cigar = q.n_cigar[i]; R.nm = q.nm[i];
        R.flag = q.flag[i]; R.mapq = q.mapq[i]; R.strand = (uint8_t)K.strand;
        R.dflag = K.dflag; R.tile = K.tile; R.frag = K.frag; R.fam = K.fam;
        R.seq_off = q.seq_off[i]; R.cigar_off = q.cigar_off[i];
        R.qual_off = (uint64_t)q.scan_ql[k];      // the read's own copy of its base qualities (a record shared by two tiles is corrected once per tile, like the reference's bam_dup1 copies)
        R.simple = d.simple; R.m_qoff = d.m_qoff;
        R.cx_off = (d.simple ? -1 : (int32_t)q.scan_cx[k]);
        R.ev_off = (int32_t)q.scan_ev[k]; R.n_ev = d.n_ev;
        R.fragprev_maxrend = INT32_MIN; R.famprev_maxrend = INT32_MIN; R.fambothprev_maxrend = INT32_MIN;
        R.raw = (int32_t)i;
        q.reads[k] = R;
        if (k == q.n_kept - 1) { q.totals->n_cx = q.scan_cx[q.n_kept]; q.totals->n_ev = q.scan_ev[q.n_kept]; q.totals->n_qual = q.scan_ql[q.n_kept]; }
    }
};
struct P0iLens {        // inputs of the cx / ev offset scans
    PrepView q;
    UVC_HD void operator()(int64_t k) const {
        const RawDer & d = q.rd[q.kept[k].raw];
        q.cx_len[k] = (d.simple ? 0 : d.rend - q.pos[q.kept[k].raw]);
        q.ev_len[k] = d.n_ev;
        q.ql_len[k] = q.l_qseq[q.kept[k].raw];
    }
};

// ------------------------------------------------------------------------------------------------ P0j: one thread per fragment -> FragRec
struct P0jFrag {
    PrepView q;
    UVC_HD void operator()(int64_t f) const {
        if (f >= q.totals->n_frags) { return; }
        const int32_t i0 = q.frag_start[f], i1 = q.frag_start[f + 1];
        const KeptRec & K0 = q.kept[q.ord2[i0]];
        FragRec G;
        G.tile = K0.tile; G.fam = K0.fam; G.strand = K0.strand;
        G.read_off = i0; G.n_reads = i1 - i0;
        G.n_cov = 0; G.n_near_mut = 0; G.normMQ = 0; G.col_off = 0;
        int32_t f_beg = INT32_MAX, f_end = 0, f_hi = 0, prevmax = INT32_MIN;
        for (int32_t i = i0; i < i1; i++) {     // file order inside the fragment (stable sort)
            const int32_t k = q.ord2[i];
            ReadRec & R = q.reads[k];
            // fillTidBegEndFromAlns1 (main.hpp:659-673). QUIRK: the exclusive end grows by one per alignment visited.
            f_beg = tmin(f_beg, R.pos); f_end = tmax(f_end, R.rend) + 1; f_hi = tmax(f_hi, R.rend);
            G.normMQ = tmax(G.normMQ, (int32_t)R.mapq);
            R.fragprev_maxrend = prevmax;
            prevmax = tmax(prevmax, R.rend);
        }
        G.beg = f_beg; G.end = f_end; G.lo = f_beg; G.hi = f_hi;
        q.frags[f] = G;
        q.fcol_len[f] = ((int64_t)(G.hi - G.lo) + UVC_COL_CHUNK - 1) / UVC_COL_CHUNK * UVC_COL_CHUNK;
#if defined(__CUDA_ARCH__)
        {
            const uint32_t peers = __match_any_sync(__activemask(), G.tile);
            const int32_t first = __reduce_min_sync(peers, (int32_t)f);
            if ((threadIdx.x & 31) == (uint32_t)(__ffs((int)peers) - 1)) { atomic_add(&q.tiles[G.tile].n_frags, __popc(peers)); atomic_min64(&q.tiles[G.tile].frag_off, (int64_t)first); }
        }
#else
        atomic_add(&q.tiles[G.tile].n_frags, 1);
        atomic_min64(&q.tiles[G.tile].frag_off, f);
#endif
    }
};

// ------------------------------------------------------------------------------------------------ P0k: one thread per family -> FamRec
struct P0kFam {
    PrepView q;
    UVC_HD void operator()(int64_t fi) const {
        if (fi >= q.totals->n_fams) { return; }
        const uvcgpu_params & par = q.par;
        const int32_t i0 = q.fam_start[fi], i1 = q.fam_start[fi + 1];
        const KeptRec & K0 = q.kept[q.ord2[i0]];
        const KeptRec & Kfirst = q.kept[K0.label];      // the reference keeps the MolecularBarcode of the first inserted read (file order) as the family's non-key data
        FamRec F;
        F.tile = K0.tile; F.duplexflag = K0.dflag; F.dedup_idflag = K0.idflag;
        F.beg_tid = Kfirst.beg_tid; F.beg_pos = Kfirst.beg_pos; F.end_tid = Kfirst.end_tid; F.end_pos = Kfirst.end_pos;
        int32_t both_beg = INT32_MAX, both_end = 0;
        int32_t i = i0;
        for (int strand = 0; strand < 2; strand++) {
            const int32_t s0 = i;
            while (i < i1 && q.kept[q.ord2[i]].strand == strand) { i++; }
            const int32_t s1 = i;
            F.frag_off[strand] = (s1 > s0 ? q.kept[q.ord2[s0]].frag : (strand == 0 ? q.kept[q.ord2[i0]].frag : (s0 < i1 ? q.kept[q.ord2[s0]].frag : q.kept[q.ord2[i1 - 1]].frag + 1)));
            int32_t s_beg = INT32_MAX, s_end = 0, s_hi = 0, n_l2r = 0, n_r2l = 0;
            int64_t qseqlen_sum = 0, n_qseqs = 0;
            for (int32_t j = s0; j < s1; j++) {
                const ReadRec & R = q.reads[q.ord2[j]];
                s_beg = tmin(s_beg, R.pos); s_end = tmax(s_end, R.rend) + 1; s_hi = tmax(s_hi, R.rend);
                both_beg = tmin(both_beg, R.pos); both_end = tmax(both_end, R.rend) + 1;
                if (R.flag & 0x10) { n_r2l++; } else { n_l2r++; }
                qseqlen_sum += R.l_qseq; n_qseqs += 1;
            }
            F.n_frags[strand] = (s1 > s0 ? q.kept[q.ord2[s1 - 1]].frag - q.kept[q.ord2[s0]].frag + 1 : 0);
            if (F.n_frags[strand] > 65535) { q.totals->err = 1; }
            F.beg2[strand] = s_beg; F.end2[strand] = s_end;
            F.lo[strand] = (F.n_frags[strand] > 0 ? s_beg : 0); F.hi[strand] = (F.n_frags[strand] > 0 ? s_hi : 0);
            F.col_off[strand] = 0;
            F.direct_frag[strand] = -1;
            int64_t mlen = 0;
            if (1 == F.n_frags[strand] && !(par.microadjust_padded_deletion_flag & 0x1)) { F.direct_frag[strand] = F.frag_off[strand]; }
            else { mlen = ((int64_t)(F.hi[strand] - F.lo[strand]) + UVC_COL_CHUNK - 1) / UVC_COL_CHUNK * UVC_COL_CHUNK; }
            q.mcol_len[2 * fi + strand] = mlen;
            // MEDIAN of the unsorted vectors (main_conversion.hpp:24-28, main.hpp:2939-2940): elements (n - 1) / 2 and n / 2 in visiting order
            int32_t l2r_a = 0, l2r_b = 0, r2l_a = 0, r2l_b = 0, cl = 0, cr = 0;
            for (int32_t j = s0; j < s1; j++) {
                const ReadRec & R = q.reads[q.ord2[j]];
                if (R.flag & 0x10) { if (cr == (n_r2l - 1) / 2) { r2l_a = R.pos; } if (cr == n_r2l / 2) { r2l_b = R.pos; } cr++; }
                else { if (cl == (n_l2r - 1) / 2) { l2r_a = R.rend; } if (cl == n_l2r / 2) { l2r_b = R.rend; } cl++; }
            }
            F.l2r_end_median[strand] = (n_l2r ? (l2r_a + l2r_b) / 2 : s_end);
            F.r2l_end_median[strand] = (n_r2l ? (r2l_a + r2l_b) / 2 : s_beg);
            F.qlen_ok[strand] = (((F.n_frags[strand] >= par.fam_thres_dup1add) && (qseqlen_sum >= n_qseqs * par.fam_thres_qseqlen)) ? 1 : 0);
            F.nsb_min[strand] = s_end; F.nsb_max[strand] = s_beg;
        }
        F.beg_both = both_beg; F.end_both = both_end;
        q.fams[fi] = F;
#if defined(__CUDA_ARCH__)
        {
            const uint32_t peers = __match_any_sync(__activemask(), F.tile);
            const int32_t first = __reduce_min_sync(peers, (int32_t)fi);
            if ((threadIdx.x & 31) == (uint32_t)(__ffs((int)peers) - 1)) { atomic_add(&q.tiles[F.tile].n_fams, __popc(peers)); atomic_min64(&q.tiles[F.tile].fam_off, (int64_t)first); }
        }
#else
        atomic_add(&q.tiles[F.tile].n_fams, 1);
        atomic_min64(&q.tiles[F.tile].fam_off, fi);
#endif
    }
};

// famprev_maxrend / fambothprev_maxrend: prefix maxima in FILE order over the reads of a (family, strand) / of a family. ord1 holds the kept
// reads ordered by (label, strand, file order): one thread per family start walks its two strand runs and merges them by file index.
struct P0kPrevMax {
    PrepView q;
    UVC_HD void operator()(int64_t i) const {
        const KeptRec & K = q.kept[q.ord1[i]];
        if (i > 0 && q.kept[q.ord1[i - 1]].label == K.label) { return; }
        int64_t a0 = i, a1 = i;
        while (a1 < q.n_kept && q.kept[q.ord1[a1]].label == K.label && q.kept[q.ord1[a1]].strand == 0) { a1++; }
        int64_t b0 = a1, b1 = a1;
        while (b1 < q.n_kept && q.kept[q.ord1[b1]].label == K.label) { b1++; }
        int32_t m = INT32_MIN;
        for (int64_t j = a0; j < a1; j++) { ReadRec & R = q.reads[q.ord1[j]]; R.famprev_maxrend = m; m = tmax(m, R.rend); }
        m = INT32_MIN;
        for (int64_t j = b0; j < b1; j++) { ReadRec & R = q.reads[q.ord1[j]]; R.famprev_maxrend = m; m = tmax(m, R.rend); }
        m = INT32_MIN;
        int64_t a = a0, b = b0;
        while (a < a1 || b < b1) {
            const bool take_a = (b >= b1 || (a < a1 && q.ord1[a] < q.ord1[b]));
            ReadRec & R = q.reads[q.ord1[take_a ? a : b]];
            R.fambothprev_maxrend = m; m = tmax(m, R.rend);
            if (take_a) { a++; } else { b++; }
        }
    }
};

// column offsets (after the scans of fcol_len / mcol_len), chunk owners and the compact per-read family records
struct P0lFragCol {
    PrepView q;
    UVC_HD void operator()(int64_t f) const {     // (launched for the upper bound n_kept)
        if (f >= q.totals->n_frags) { return; }
        FragRec & G = q.frags[f];
        G.col_off = q.scan_fcol[f];
        if (f == q.totals->n_frags - 1) { q.totals->n_fcol = q.scan_fcol[f + 1]; }
    }
};
struct P0lFamCol {
    PrepView q;
    UVC_HD void operator()(int64_t fs) const {    // (launched for the upper bound 2 * n_kept)
        if (fs >= 2 * q.totals->n_fams) { return; }
        FamRec & F = q.fams[fs >> 1];
        F.col_off[fs & 1] = q.scan_mcol[fs];
        if (fs == 2 * q.totals->n_fams - 1) { q.totals->n_mcol = q.scan_mcol[fs + 1]; }
    }
};
struct P0mFragChunks {
    PrepView q;
    UVC_HD void operator()(int64_t f) const {
        const FragRec & G = q.frags[f];
        const int64_t c0 = G.col_off / UVC_COL_CHUNK, c1 = c0 + ((int64_t)(G.hi - G.lo) + UVC_COL_CHUNK - 1) / UVC_COL_CHUNK;
        for (int64_t c = c0; c < c1; c++) { q.fchunk_frag[c] = (int32_t)f; }
    }
};
struct P0mFamChunks {
    PrepView q;
    UVC_HD void operator()(int64_t fs) const {
        const FamRec & F = q.fams[fs >> 1];
        const int strand = (int)(fs & 1);
        if (F.direct_frag[strand] >= 0) { return; }
        const int64_t c0 = F.col_off[strand] / UVC_COL_CHUNK, c1 = c0 + ((int64_t)(F.hi[strand] - F.lo[strand]) + UVC_COL_CHUNK - 1) / UVC_COL_CHUNK;
        for (int64_t c = c0; c < c1; c++) { q.mchunk_fs[c] = (int32_t)fs; }
    }
};
struct P0nReadFam {
    PrepView q;
    UVC_HD void operator()(int64_t k) const {
        const ReadRec & R = q.reads[k];
        const FamRec & F = q.fams[R.fam];
        ReadFam o;
        o.rend = R.rend; o.famprev_maxrend = R.famprev_maxrend; o.fambothprev_maxrend = R.fambothprev_maxrend; o.fam = R.fam; o.pad = 0;
        o.flags = (R.strand ? UVC_RF_STRAND : 0u) | ((F.duplexflag & 0x2) ? UVC_RF_DUPLEX_UMI : 0u) | ((F.n_frags[0] > 0 && F.n_frags[1] > 0) ? UVC_RF_BOTH_STRANDS : 0u);
        if (F.direct_frag[R.strand] >= 0) {
            const FragRec & G = q.frags[F.direct_frag[R.strand]];
            o.flags |= UVC_RF_DIRECT;
            o.col_base = G.col_off - G.lo;
        } else {
            o.col_base = F.col_off[R.strand] - F.lo[R.strand];
        }
        q.rfam[k] = o;
    }
};

// ================================================================================================ stage P1
UVC_HD uint8_t char_to_symbol(char c) { // CHAR_TO_SYMBOL (main_conversion.hpp:473-486)
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
// reference base at offset o of tile T's reference string (length n_ref = ext_end - ext_beg - 1); 'n' if the contig is not available
UVC_HD char ref_char(const PrepView & q, const PrepTile & P, const TileInfo & T, int32_t o) {
    return (P.contig ? P.contig[T.ext_beg + o] : 'n');
}
// main.hpp:699-721. QUIRK: rank2 is computed with rulen1 when rc2 <= 1.
UVC_HD bool more_str(int32_t rulen1, int32_t rc1, int32_t rulen2, int32_t rc2, int32_t repeatsize_max) {
    if (rulen2 * rc2 == 0) { return true; }
    if (rulen1 > repeatsize_max || rulen2 > repeatsize_max) { return (rulen1 < rulen2 || (rulen1 == rulen2 && rc1 > rc2)); }
    int rank1 = (rc1 <= 1 ? (-rc1 * rulen1) : ((rc1 - 1) * rulen1));
    int rank2 = (rc2 <= 1 ? (-rc2 * rulen1) : ((rc2 - 1) * rulen2));
    if (0 == rc1 || 0 == rulen1) { rank1 = -100; }
    if (0 == rc2 || 0 == rulen2) { rank2 = -100; }
    return rank1 > rank2;
}

// P1a: one thread per extended position: tile id, reference symbol, and what refstring2repeatvec (main.hpp:803-874) finds when its walk stops
// here: the best short-tandem-repeat unit (<= indel_str_repeatsize_max) and the best any-tandem-repeat unit (<= indel_vntr_repeatsize_max)
// with the lengths of their match runs, and how far the walk advances from here.
// p1_cand = str unit (8 bits) | any unit (8 bits) << 8 | str run (24 bits) << 16 | any run (24 bits) << 40
struct P1aCand {
    PrepView q;
    UVC_HD void operator()(int64_t gp) const {
        // tile of the position: the last tile whose first position is <= gp (empty tiles share the offset of their successor)
        int32_t a = 0, b = q.n_tiles;
        while (b - a > 1) { const int32_t m = (a + b) >> 1; if (q.tile_npos[m] <= gp) { a = m; } else { b = m; } }
        const int32_t t = a;
        const TileInfo & T = q.tiles[t];
        const PrepTile & P = q.pt[t];
        q.pos_tile[gp] = t;
        const int32_t n = (T.ext_end - T.ext_beg) - 1;        // length of the reference string
        const int32_t refpos = (int32_t)(gp - T.pos_off);
        q.p1_visited[gp] = 0; q.p1_str_key[gp] = 0; q.p1_any_key[gp] = 0;
        if (refpos >= n) { q.refsym[gp] = UVC_BASE_N; q.p1_cand[gp] = 0; q.p1_jump[gp] = 1; return; }
        q.refsym[gp] = char_to_symbol(ref_char(q, P, T, refpos));
        const int32_t str_max = q.par.indel_str_repeatsize_max, vntr_max = q.par.indel_vntr_repeatsize_max;
        int32_t best_unit = 0, best_num = 0, best_run = 0, any_unit = 0, any_num = 0, any_run = 0;
        for (int32_t unit = 1; unit <= vntr_max; unit++) {
            int32_t qq = refpos;
            while (qq + unit < n && ref_char(q, P, T, qq) == ref_char(q, P, T, qq + unit)) { qq++; }
            const int32_t num = (qq - refpos) / unit + 1;
            if (unit <= str_max && more_str(unit, num, best_unit, best_num, str_max)) { best_unit = unit; best_num = num; best_run = qq - refpos; }
            if (more_str(unit, num, any_unit, any_num, vntr_max)) { any_unit = unit; any_num = num; any_run = qq - refpos; }
        }
        q.p1_cand[gp] = (uint64_t)(uint32_t)best_unit | ((uint64_t)(uint32_t)any_unit << 8) | ((uint64_t)(uint32_t)tmin(best_run, 0xffffff) << 16)
                | ((uint64_t)(uint32_t)tmin(any_run, 0xffffff) << 40);
        const int32_t nb = str_max + best_unit;          // the walk's step from here (main.hpp:868-869)
        q.p1_jump[gp] = tmax(best_unit * best_num, nb + 1) - nb;
    }
};
// P1b: one thread per tile follows the reference's skip-walk (refpos += step) and marks the visited positions
struct P1bWalk {
    PrepView q;
    UVC_HD void operator()(int64_t t) const {
        const TileInfo & T = q.tiles[t];
        const int32_t n = (T.ext_end - T.ext_beg) - 1;
        for (int32_t refpos = 0; refpos < n;) { q.p1_visited[T.pos_off + refpos] = 1; refpos += q.p1_jump[T.pos_off + refpos]; }
    }
};
// P1c: every visited position offers its two tracks to the positions they cover: the longest track wins, the earliest start on ties
// (main.hpp:846-866 assigns in walk order with a strict comparison). key = tracklen << 32 | ~start.
struct P1cOffer {
    PrepView q;
    UVC_HD void operator()(int64_t gp) const {
        if (!q.p1_visited[gp]) { return; }
        const TileInfo & T = q.tiles[q.pos_tile[gp]];
        const int32_t n = (T.ext_end - T.ext_beg) - 1;
        const int32_t refpos = (int32_t)(gp - T.pos_off);
        const uint64_t cand = q.p1_cand[gp];
        const int32_t best_unit = (int32_t)(cand & 0xff), any_unit = (int32_t)((cand >> 8) & 0xff);
        const int32_t best_run = (int32_t)((cand >> 16) & 0xffffff), any_run = (int32_t)((cand >> 40) & 0xffffff);
        {
            const int32_t stop = tmin(refpos + best_run + best_unit, n);
            const uint64_t key = ((uint64_t)(uint32_t)(stop - refpos) << 32) | (uint64_t)(0xffffffffu - (uint32_t)refpos);
            for (int32_t i = refpos; i < stop; i++) { atomic_max_u64(&q.p1_str_key[T.pos_off + i], key); }
        }
        {
            const int32_t stop = tmin(refpos + any_run + any_unit, n);
            const uint64_t key = ((uint64_t)(uint32_t)(stop - refpos) << 32) | (uint64_t)(0xffffffffu - (uint32_t)refpos);
            for (int32_t i = refpos; i < stop; i++) { atomic_max_u64(&q.p1_any_key[T.pos_off + i], key); }
        }
    }
};
// P1d: one thread per position: the winning tracks -> RegionalTandemRepeat; increments of the two BAQ prefix sums (main.cpp:400-429).
// QUIRK: the any-tandem-repeat variant still divides by the STR unit length.
struct P1dDecode {
    PrepView q; int32_t *inc1, *inc2;
    UVC_HD void operator()(int64_t gp) const {
        const TileInfo & T = q.tiles[q.pos_tile[gp]];
        const int32_t n = (T.ext_end - T.ext_beg) - 1;
        const int32_t refpos = (int32_t)(gp - T.pos_off);
        const int64_t src = (refpos >= n ? (n > 0 ? gp - 1 : -1) : gp);     // the extra last record repeats its predecessor (main.hpp:872)
        uvcgpu_rtr r;
        r.begpos = 0; r.tracklen = 0; r.unitlen = 0; r.indelphred = q.par.indel_BQ_max; r.anyTR_begpos = 0; r.anyTR_tracklen = 0; r.anyTR_unitlen = 0;
        if (src >= 0) {
            const uint64_t ks = q.p1_str_key[src], ka = q.p1_any_key[src];
            if (ks) {
                const int32_t start = (int32_t)(0xffffffffu - (uint32_t)(ks & 0xffffffffu)), tl = (int32_t)(ks >> 32);
                const int32_t unit = (int32_t)(q.p1_cand[T.pos_off + start] & 0xff);
                r.begpos = start; r.tracklen = tl; r.unitlen = unit;
                // indel_phred (main.hpp:794-801) with ampfact = slip rate x del-to-ins ratio: variant 1 of the host-evaluated table
                const int32_t nun = tl / unit;
                int32_t dec;
                if (unit >= 1 && unit <= UVC_SLIP_MAXUNIT && nun >= 0 && nun < UVC_SLIP_NMAX) { dec = q.slip_tab[(UVC_SLIP_MAXUNIT + (unit - 1)) * UVC_SLIP_NMAX + nun]; }
                else {
                    const double ampfact = q.par.indel_polymerase_slip_rate * q.par.indel_del_to_ins_err_ratio;
                    const int32_t region = unit * nun;
                    const double num_slips = (region > 64 ? (double)(region - 8) : log1p(exp((double)region - 8.0))) * ampfact / ((double)(unit * unit));
                    dec = (int32_t)floor(-10 * log((1.0 - 2.220446049250313e-16) / (num_slips + 1.0)) / log(10.0));
                }
                r.indelphred = q.par.indel_BQ_max - tmin(q.par.indel_BQ_max - 1, dec);
            }
            if (ka) {
                const int32_t start = (int32_t)(0xffffffffu - (uint32_t)(ka & 0xffffffffu)), tl = (int32_t)(ka >> 32);
                r.anyTR_begpos = start; r.anyTR_tracklen = tl; r.anyTR_unitlen = (int32_t)((q.p1_cand[T.pos_off + start] >> 8) & 0xff);
            }
        }
        q.rtr[gp] = r;
        const int32_t polsize = (int32_t)round(q.par.indel_polymerase_size);
        const int32_t ul = r.unitlen;
        for (int v = 0; v < 2; v++) {
            const int32_t tl = (v ? r.anyTR_tracklen : r.tracklen);
            int32_t inc;
            if (ul > 0 && (tl / ul >= 3 || (tl / ul >= 2 && tl >= polsize))) { inc = (q.par.indel_str_phred_per_region * 10) / tl + 1; }
            else { inc = q.par.indel_nonSTR_phred_per_base * 10; }
            (v ? inc2 : inc1)[gp] = inc;
        }
    }
};
struct P1fKeepPhred { const uvcgpu_rtr *rtr; int32_t *out; UVC_HD void operator()(int64_t gp) const { out[gp] = rtr[gp].indelphred; } };
struct P1eBaq {         // after the scans of the increments: prefix sums restart at every tile
    PrepView q; const int64_t *scan1, *scan2;
    UVC_HD void operator()(int64_t gp) const {
        const TileInfo & T = q.tiles[q.pos_tile[gp]];
        q.baq[gp] = (int32_t)((int64_t)(int32_t)(scan1[gp + 1] - scan1[T.pos_off]) / 10);
        q.baq2[gp] = (int32_t)((int64_t)(int32_t)(scan2[gp + 1] - scan2[T.pos_off]) / 10);
    }
};

} // namespace uvc
#endif
