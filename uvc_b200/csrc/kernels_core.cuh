// kernels_core.cuh - per-work-item bodies of the pileup kernels.
//
// Execution model ("pileup columns"): the reference walks every read and scatters ~130 read-modify-writes per aligned
// base into a 6 KB-per-position state (main.hpp:925-1204, 1762-2296, 1363-1595). Here every reference position is
// owned by one thread that GATHERS the reads covering it (reads are sorted by start position, so they form a window
// found by binary search) and keeps the hot counters (the reference-matching base symbol and LINK_M) in registers;
// the position's records are then written once, as contiguous bursts, in the reference's own struct layout.
// No atomics are needed on the dense path because each position's counters have exactly one writer; only the rare
// per-CIGAR-operation events (indels, clips, mismatch runs) use global atomics, from the per-read kernels.
//
// Every body is a plain function of (view, index), compiled for the device by nvcc and - for the CPU-only unit tests
// of the host logic (tests/ only, never shipped as a fallback) - as ordinary C++.
#ifndef UVC_KERNELS_CORE_CUH_INCLUDED
#define UVC_KERNELS_CORE_CUH_INCLUDED

#include "batch.h"

#include <limits.h>
#include <math.h>

#if defined(__CUDACC__)
#define UVC_HD __host__ __device__ __forceinline__
#else
#define UVC_HD inline
#endif

namespace uvc {

template <class T> UVC_HD T tmin(T a, T b) { return a < b ? a : b; }
template <class T> UVC_HD T tmax(T a, T b) { return a > b ? a : b; }
UVC_HD int32_t nnminus(int32_t a, int32_t b) { return (a > b ? a - b : 0); }
UVC_HD int32_t __popc_u32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}
UVC_HD int32_t __ctz_u32(uint32_t x) {   // x != 0
#if defined(__CUDA_ARCH__)
    return __ffs((int)x) - 1;
#else
    return __builtin_ctz(x);
#endif
}
UVC_HD int32_t iabs(int32_t a) { return a < 0 ? -a : a; }
UVC_HD int32_t between(int32_t v, int32_t lo, int32_t hi) { return tmin(tmax(lo, v), hi); }

UVC_HD void atomic_add(int32_t *p, int32_t v) {
#if defined(__CUDA_ARCH__)
    atomicAdd(p, v);
#else
    *p += v;
#endif
}
UVC_HD void atomic_add(int64_t *p, int64_t v) {
#if defined(__CUDA_ARCH__)
    atomicAdd((unsigned long long*)p, (unsigned long long)v);
#else
    *p += v;
#endif
}

UVC_HD int cig_op(uint32_t c) { return (int)(c & 0xf); }
UVC_HD int32_t cig_len(uint32_t c) { return (int32_t)(c >> 4); }
UVC_HD bool is_match_op(int op) { return op == UVC_CMATCH || op == UVC_CEQUAL || op == UVC_CDIFF; }
UVC_HD bool is_ins_symbol(int s) { return s == UVC_LINK_I1 || s == UVC_LINK_I2 || s == UVC_LINK_I3P; }
UVC_HD bool is_del_symbol(int s) { return s == UVC_LINK_D1 || s == UVC_LINK_D2 || s == UVC_LINK_D3P; }

// seq_nt16_int of the 4-bit base at query index i: A,C,G,T -> 0..3, anything else -> 4 (BASE_N) (main.hpp:1829-1833)
// nibble k of the constant = symbol of the 4-bit code k (1, 2, 4, 8 -> A, C, G, T; everything else -> N): a shift instead of a branch chain
UVC_HD int nt16_to_symbol(uint32_t b4) { return (int)((0x4444444344424104ULL >> ((b4 & 0xfu) * 4)) & 0xfu); }
UVC_HD int base3(const uint8_t *seq, int32_t i) {
    return nt16_to_symbol((uint32_t)(seq[i >> 1] >> ((~i & 1) << 2)));
}
UVC_HD int base4(const uint8_t *seq, int32_t i) { return (seq[i >> 1] >> ((~i & 1) << 2)) & 0xf; }

UVC_HD const uvcgpu_params & par_of(const BatchView & v) { return v.par; }
UVC_HD int read_strand(uint16_t flag) { return (((flag & 0x81) == 0x81) ? (!!(flag & 0x20)) : (!!(flag & 0x10))); }

// Range [lo, hi) of the tile's reads whose start lies in (p - max_span, p]: the only reads that can cover p.
UVC_HD void read_window(const BatchView & v, const TileInfo & T, int32_t p, int64_t & lo, int64_t & hi) {
    const int64_t r0 = T.read_off, r1 = T.read_off + T.n_reads;
    const int32_t minpos = p - T.max_read_span;  // reads with pos <= minpos cannot reach p
    int64_t a = r0, b = r1;
    while (a < b) { const int64_t m = (a + b) >> 1; if (v.reads[m].pos <= minpos) { a = m + 1; } else { b = m; } }
    lo = a;
    b = r1;
    while (a < b) { const int64_t m = (a + b) >> 1; if (v.reads[m].pos <= p) { a = m + 1; } else { b = m; } }
    hi = a;
}

// Read window of a position thread. [lo, hi) is the position's own window; [ulo, uhi) is the union over the 32 positions of its warp: every
// lane walks the union in lock-step (skipping reads outside its own window), so that at each step all lanes look at the SAME read - its
// record is one broadcast load and the per-position data of that read (column entries, base qualities) are consecutive addresses.
struct Win { int64_t lo, hi, ulo, uhi; };

UVC_HD void position_window(const BatchView & v, int64_t gp, Win & w) {
    const TileInfo & T = v.tiles[v.pos_tile[gp]];
    read_window(v, T, (int32_t)(gp - T.pos_off) + T.ext_beg, w.lo, w.hi);
    w.ulo = w.lo; w.uhi = w.hi;
}

// What read R shows at reference position p (pos <= p < rend).
struct Locus {
    int32_t qpos;       // query index, -1 if not an aligned base
    bool is_m;          // aligned (M/=/X) base
    bool not_first;     // not the first base of its M run (the junction before it carries LINK_M)
    bool is_del;        // deleted base
    int32_t prev_rpos, next_rpos; // neighbouring low-quality indels (main.hpp:1902-1903); for deleted bases: the two precomputed distances
};

UVC_HD Locus locate(const BatchView & v, const ReadRec & R, int32_t p) {
    Locus L;
    if (R.simple) {
        const int32_t o = p - R.pos;
        L.qpos = R.m_qoff + o; L.is_m = true; L.not_first = (o > 0); L.is_del = false; L.prev_rpos = 0; L.next_rpos = INT32_MAX;
    } else {
        const CxEntry e = v.cx[R.cx_off + (p - R.pos)];
        L.qpos = e.qpos; L.is_m = (e.flags & 1); L.not_first = (e.flags & 2); L.is_del = (e.flags & 4); L.prev_rpos = e.prev_rpos; L.next_rpos = e.next_rpos;
    }
    return L;
}

UVC_HD bool primer_masked(const BatchView & v, const ReadRec & R, const ReadDerived & D, int32_t rpos) {
    // main.hpp:1895: a base counts if the assay is not an amplicon (or the normal is used to filter primers) or it lies inside the insert minus primers
    const bool is_assay_amplicon = ((R.dflag & 0x4) || ((v.par.primerlen > 0) && !(0x2 & v.par.primer_flag)));
    const bool normal_filters_primers = (v.par.tn_is_paired && (0x1 & v.par.primer_flag));
    return !((normal_filters_primers || !is_assay_amplicon) || (D.ibeg <= rpos && rpos < D.iend));
}

// apply_bq_err_correction3 (grouping.cpp:459-543) on the read's own copy of its base qualities: the platform increment and cap, the 3'-tail
// penalty (long end clip and/or a low-complexity tail up to the 2nd distinct base of quality >= 20), and poly-G (>= 4) minus one.
UVC_HD void fix_base_qualities(uint8_t *q, const uint8_t *seq, const uint32_t *cigar, const ReadRec & R, const uvcgpu_params & par) {
    const int32_t l = R.l_qseq;
    if ((0 == l) || (R.flag & 0x4)) { return; }
    for (int32_t i = 0; i < l; i++) { q[i] = (uint8_t)tmin((int32_t)q[i] + par.assay_sequencing_BQ_inc, par.assay_sequencing_BQ_max); }
    const int isrc = ((R.flag & 0x10) ? 1 : 0);
    int32_t inclu_beg[2] = {0, l - 1};
    int32_t exclu_end[2] = {l, -1};
    int32_t end_clip_len = 0;
    if (R.n_cigar > 0) {
        uint32_t c = cigar[0];
        if (cig_op(c) == UVC_CSOFT_CLIP) {
            if (0 == isrc) { inclu_beg[0] += cig_len(c); } else { exclu_end[1] += cig_len(c); end_clip_len = cig_len(c); }
        }
        c = cigar[R.n_cigar - 1];
        if (cig_op(c) == UVC_CSOFT_CLIP) {
            if (1 == isrc) { inclu_beg[1] -= cig_len(c); } else { exclu_end[0] -= cig_len(c); end_clip_len = cig_len(c); }
        }
    }
    const int32_t inc = (isrc ? -1 : 1);
    {
        int prev_b = 0;
        int distinct = 0;
        const int32_t start = exclu_end[isrc] - inc;
        int32_t termpos = start;
        for (; termpos != inclu_beg[isrc] - inc; termpos -= inc) {
            const int b = base4(seq, termpos);
            if (b != prev_b && q[termpos] >= 20) {
                prev_b = b;
                distinct++;
                if (2 == distinct) { break; }
            }
        }
        const int32_t tracklen = iabs(termpos - start);
        const int32_t tail_penal = (end_clip_len >= 20 ? 1 : 0) + (tracklen >= 15 ? 2 : (tracklen >= 10 ? 1 : 0));
        if (tail_penal > 0) {
            for (int32_t p = start; p != (inclu_beg[isrc] - inc) && p != termpos; p -= inc) {
                q[p] = (uint8_t)(tmax((int32_t)q[p], tail_penal + 1) - tail_penal);
            }
        }
    }
    {
        int32_t homopol = 0;
        int prev_b = 0;
        for (int32_t p = inclu_beg[isrc]; p != exclu_end[isrc]; p += inc) {
            const int b = base4(seq, p);
            if (b == prev_b) {
                homopol++;
                if (homopol >= 4 && b == 4 /* nt16 code of G */) { q[p] = (uint8_t)(tmax((int32_t)q[p], 2) - 1); }
            } else {
                prev_b = b;
                homopol = 1;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ K0: one thread per read
// Read-level constants (main.hpp:937-998, 1789-1885), the per-reference-base expansion and the low-quality-indel
// neighbours of complex reads (main.hpp:1817-1859, 1897-1916, 2219-2252), the list of indel events, and the rare
// per-operation contributions to the prep sets (insertions, deletions, clips, mismatch runs; main.hpp:1025-1046, 1069-1199).
// seq_ro / qual_rw: where the read's packed bases and its base qualities live for the duration of the call (NULL: in the batch arrays; the CUDA
// kernel passes shared-memory copies that it loads and stores with coalesced accesses). The qualities are corrected in place.
UVC_HD void k0_read(const BatchView & v, int64_t ri, const uint8_t *seq_ro = NULL, uint8_t *qual_rw = NULL) {
    const ReadRec & R = v.reads[ri];
    ReadDerived D;
    const TileInfo & T = v.tiles[R.tile];
    const uint32_t *cigar = v.cigar + R.cigar_off;
    const uint8_t *seq = (seq_ro ? seq_ro : v.seq + R.seq_off);
    uint8_t *qual_w = (qual_rw ? qual_rw : v.qual + R.qual_off);
    if (!qual_rw) { const uint8_t *raw = v.qual_raw + v.raw_qual_off[R.raw]; for (int32_t i = 0; i < R.l_qseq; i++) { qual_w[i] = raw[i]; } }
    fix_base_qualities(qual_w, seq, cigar, R, par_of(v));
    const uint8_t *qual = qual_w;
    const int64_t po = T.pos_off - T.ext_beg;        // concatenated index of reference position x is po + x
    const int32_t npos = T.ext_end - T.ext_beg;
    const int32_t *baq = v.baq + po;
    const uint8_t *refsym = v.refsym + po;
    const uvcgpu_params & par = v.par;
    const int32_t pos = R.pos, rend = R.rend;

    int32_t nge = 0, ngo = 0, clip_cnt = 0, insbaq = 0, delbaq = 0, inslen = 0, dellen = 0;
    {
        int32_t rpos = pos;
        for (int32_t i = 0; i < R.n_cigar; i++) {
            const int op = cig_op(cigar[i]); const int32_t len = cig_len(cigar[i]);
            if (op == UVC_CINS || op == UVC_CDEL) {
                nge += len; ngo++;
                const int32_t d = baq[tmin(rpos + len, T.ext_end - 1)] - baq[rpos];
                if (op == UVC_CINS) { insbaq += d; inslen += len; } else { delbaq += d; dellen += len; rpos += len; }
            } else if (is_match_op(op) || op == UVC_CREF_SKIP) {
                rpos += len;
            }
            if (op == UVC_CSOFT_CLIP || op == UVC_CHARD_CLIP) { clip_cnt++; }
        }
    }
    const int32_t nm_cnt = (R.nm >= 0 ? R.nm : nge);
    const int32_t xm_cnt = nm_cnt - nge;
    const int32_t span = rend - pos;
    D.xm1500 = xm_cnt * 1500 / span;
    D.go1500 = ngo * 1500 / span;
    D.avg_gaplen = nge / tmax(1, ngo);
    D.nge_cnt = nge; D.ngo_cnt = ngo; D.clip_cnt = clip_cnt;
    D.inslen_sum = inslen; D.dellen_sum = dellen; D.insbaq_sum = insbaq; D.delbaq_sum = delbaq;
    D.lclip = ((R.n_cigar > 0 && cig_op(cigar[0]) == UVC_CSOFT_CLIP) ? cig_len(cigar[0]) : 0);
    D.rclip = ((R.n_cigar > 0 && cig_op(cigar[R.n_cigar - 1]) == UVC_CSOFT_CLIP) ? cig_len(cigar[R.n_cigar - 1]) : 0);
    const int32_t penal_clip = tmax(D.lclip, D.rclip) / 6;
    const int32_t penal_nm = (D.xm1500 + D.go1500) / 30;
    D.micro_indel_penal = tmin(1, penal_nm + penal_clip);
    D.micro_nogap_penal = tmin(4, penal_nm + penal_clip) + 1;
    const bool isrc = ((R.flag & 0x10) == 0x10);
    D.ibeg = ((R.isize != 0) ? (tmin(pos, R.mpos) + par.primerlen)
            : ((isrc && (0x0 == (0x1 & R.flag))) ? 0 : (pos + par.primerlen)));
    D.iend = ((R.isize != 0) ? nnminus(tmin(pos, R.mpos) + iabs(R.isize), par.primerlen)
            : ((isrc && (0x0 == (0x1 & R.flag))) ? nnminus(rend, par.primerlen) : INT32_MAX));

    const int32_t pcr_inc = ((R.dflag & 0x4) ? 1 : 0);
    const int32_t umi_inc = ((R.dflag & 0x1) ? 1 : 0);
    const int32_t frag_pos_L = tmin(pos, R.mpos);
    const int32_t frag_pos_R = frag_pos_L + iabs(R.isize);
    const int32_t adj = par.indel_adj_tracklen_dist;
    const uvcgpu_rtr *rtr = v.rtr + T.pos_off;          // indexed by p - ext_beg
    uvcgpu_prep_set *prep = v.prep + po;

    int32_t bm_cnt[5] = {0, 0, 0, 0, 0};
    int32_t n_ev = 0;
    // low-quality indels of this read in reference order: lowq[0] = 0, ..., INT32_MAX (main.hpp:1817-1858)
    // kept implicitly: we only need, while replaying the walk, the current "previous" and "next" entries.
    {
        int32_t qpos = 0, rpos = pos;
        // first pass: mismatch counts per base type, indel event records, prep-set events
        for (int32_t i = 0; i < R.n_cigar; i++) {
            const int op = cig_op(cigar[i]); const int32_t len = cig_len(cigar[i]);
            if (is_match_op(op)) {
                for (int32_t j = 0; j < len; j++) {
                    const int b = base3(seq, qpos);
                    if ((int)refsym[rpos] != b) {
                        bm_cnt[b] += 1;
                        // mismatch-run detection exactly as written (main.hpp:1025-1046): the look-ahead runs over raw query and
                        // reference indices regardless of the CIGAR, bounded by l_qseq and rend
                        int32_t nq = qpos + 1, nr = rpos + 1;
                        bool mism = true;
                        while (mism && nq < R.l_qseq && nr < rend) {
                            mism = ((int)refsym[nr] != base3(seq, nq));
                            nq++; nr++;
                        }
                        if (nr == rpos + 2) {
                            for (int32_t r = tmax(pos, rpos - 1); r < tmin(nr, rend); r++) { atomic_add(&prep[r].a_snv_dp, 1); }
                        }
                        if (nr > rpos + 2) {
                            for (int32_t r = tmax(pos, rpos - 1); r < tmin(nr, rend); r++) { atomic_add(&prep[r].a_dnv_dp, 1); }
                        }
                    }
                    qpos++; rpos++;
                }
            } else if (op == UVC_CINS || op == UVC_CDEL) {
                const int32_t ridx = rpos - T.ext_beg;
                const uvcgpu_rtr rtr1 = rtr[tmax(adj, ridx) - adj];
                const uvcgpu_rtr rtr2 = rtr[tmin(ridx + adj, npos - 1)];
                const int32_t unitlen2 = tmax(1, (rtr1.tracklen > rtr2.tracklen) ? rtr1.unitlen : rtr2.unitlen);
                const int32_t inv100 = 100 / ((0 == len % unitlen2) ? (len / unitlen2) : 4);
                const int32_t rtr_lo = tmax((T.ext_beg + rtr1.begpos) - adj, pos);
                const int32_t rtr_hi = tmin((T.ext_beg + rtr2.begpos + rtr2.tracklen) + adj, rend);
                if (op == UVC_CINS) {
                    const int32_t nbases = (int32_t)(((uint32_t)len * (uint32_t)par.indel_adj_indellen_perc) / 100u);
                    for (int32_t r2 = tmax(rpos - nbases, pos); r2 < tmin(rpos + nbases, rend); r2++) {
                        atomic_add(&prep[r2].a_near_ins_dp, 1);
                        atomic_add(&prep[r2].a_near_ins_pow2len, (int64_t)((uint32_t)len * (uint32_t)len));
                        atomic_add(&prep[r2].a_near_ins_l_pow2len, (int64_t)((r2 + 1 - (rpos - nbases)) * (r2 + 1 - (rpos - nbases))));
                        atomic_add(&prep[r2].a_near_ins_r_pow2len, (int64_t)(((rpos + nbases) - r2) * ((rpos + nbases) - r2)));
                        atomic_add(&prep[r2].a_near_ins_inv100len, inv100);
                    }
                    for (int32_t r2 = rtr_lo; r2 < rtr_hi; r2++) { atomic_add(&prep[r2].a_near_RTR_ins_dp, 1); }
                    atomic_add(&prep[rpos].a_at_ins_dp, 1);
                } else {
                    // every deleted base gets the dense per-base counters (main.hpp:1127-1161)
                    const int32_t ldist = rpos - pos + 1, rdist = rend - rpos;
                    const int32_t lbaq = baq[rpos] - baq[pos] + 1, rbaq = baq[rend - 1] - baq[rpos] + 1;
                    for (int32_t r2 = rpos; r2 < rpos + len; r2++) {
                        uvcgpu_prep_set & q = prep[r2];
                        atomic_add(&q.a_pcr_dp, pcr_inc); atomic_add(&q.a_umi_dp, umi_inc); atomic_add(&q.a_dp, 1);
                        atomic_add(&q.a_qlen, span); atomic_add(&q.a_highBQ_dp, 1);
                        atomic_add(&q.a_XM1500, D.xm1500); atomic_add(&q.a_GO1500, D.go1500); atomic_add(&q.a_GAPLEN, D.avg_gaplen);
                        if (R.isize != 0) {
                            // QUIRK: distances are taken at the deletion start rpos, not at r2 (main.hpp:1139-1142)
                            if (isrc) { atomic_add(&q.a_LI, (int64_t)tmin(rpos - frag_pos_L + 1, UVC_MAX_INSERT_SIZE)); atomic_add(&q.a_LIDP, 1); }
                            else { atomic_add(&q.a_RI, (int64_t)tmin(frag_pos_R - rpos, UVC_MAX_INSERT_SIZE)); atomic_add(&q.a_RIDP, 1); }
                        }
                        atomic_add(&q.a_l_dist_sum, ldist); atomic_add(&q.a_r_dist_sum, rdist);
                        atomic_add(&q.a_inslen_sum, inslen); atomic_add(&q.a_dellen_sum, dellen);
                        // QUIRK: the BAQ sums go to the deletion start once per deleted base (main.hpp:1156-1157)
                        atomic_add(&prep[rpos].a_l_BAQ_sum, (int64_t)lbaq); atomic_add(&prep[rpos].a_r_BAQ_sum, (int64_t)rbaq);
                        atomic_add(&q.a_insBAQ_sum, (int64_t)insbaq); atomic_add(&q.a_delBAQ_sum, (int64_t)delbaq);
                    }
                    const int32_t nb_l = (int32_t)(((uint32_t)len * (uint32_t)(par.indel_adj_indellen_perc - 100)) / 100u);
                    const int32_t nb_r = (int32_t)(((uint32_t)len * (uint32_t)par.indel_adj_indellen_perc) / 100u);
                    const int32_t lp = tmax(rpos - nb_l, pos);
                    const int32_t rp = tmin(rpos + nb_r, rend) - 1;
                    for (int32_t r2 = lp; r2 <= rp; r2++) {
                        atomic_add(&prep[r2].a_near_del_dp, 1);
                        atomic_add(&prep[r2].a_near_del_pow2len, (int64_t)((uint32_t)len * (uint32_t)len));
                        atomic_add(&prep[r2].a_near_del_l_pow2len, (int64_t)((r2 - lp + 1) * (r2 - lp + 1)));
                        atomic_add(&prep[r2].a_near_del_r_pow2len, (int64_t)((rp - r2 + 1) * (rp - r2 + 1)));
                        atomic_add(&prep[r2].a_near_del_inv100len, inv100);
                    }
                    for (int32_t r2 = rtr_lo; r2 < rtr_hi; r2++) { atomic_add(&prep[r2].a_near_RTR_del_dp, 1); }
                    atomic_add(&prep[rpos].a_at_del_dp, 1);
                }
                if (!R.simple) {
                    IndelEvent & E = v.ev[R.ev_off + n_ev];
                    E.read = (int32_t)ri; E.rpos = rpos; E.oplen = len; E.qpos = qpos; E.is_del = (op == UVC_CDEL); E.cigar_idx = i; E.tile = R.tile; E.raw = R.raw;
                    E.symbol = -1; E.incvalue = 0; E.incvalue2 = 0; E.counted = 0;
                    n_ev++;
                }
                if (op == UVC_CINS) { qpos += len; } else { rpos += len; }
            } else {
                const int32_t rdelta = ((0 == i) ? 0 : -1);
                if ((op == UVC_CSOFT_CLIP || op == UVC_CHARD_CLIP) && pcr_inc) {
                    for (int32_t r2 = rpos + rdelta - par.microadjust_near_clip_dist; r2 <= rpos + rdelta + par.microadjust_near_clip_dist; r2++) {
                        if (T.ext_beg <= r2 && r2 < T.ext_end) { atomic_add(&prep[r2].a_near_pcr_clip_dp, pcr_inc); }
                    }
                }
                if ((op == UVC_CSOFT_CLIP || op == UVC_CHARD_CLIP) && (0 == pcr_inc) && (len >= par.microadjust_alignment_clip_min_len)) {
                    atomic_add(&prep[rpos + rdelta].a_near_long_clip_dp, 1);
                }
                if (op == UVC_CREF_SKIP) { rpos += len; } else if (op == UVC_CSOFT_CLIP) { qpos += len; }
            }
        }
    }
    for (int b = 0; b < 5; b++) { D.bm1500[b] = bm_cnt[b] * 1500 / span; D.bm_term[b] = (D.bm1500[b] > 20 ? (100 * 400 / (D.bm1500[b] * D.bm1500[b])) : 100); }
    D.xm_term = (D.xm1500 > 20 ? (100 * 400 / (D.xm1500 * D.xm1500)) : 100);
    D.baq_pos = baq[pos]; D.baq_rend1 = baq[rend - 1]; D.baq2_rend1 = (v.baq2 + po)[rend - 1];
    v.rd[ri] = D;
    {
        PileRec P;
        P.pos = pos; P.rend = rend; P.frag_l = frag_pos_L; P.frag_r = frag_pos_R;
        P.baq_pos = D.baq_pos; P.baq_rend1 = D.baq_rend1; P.baq2_rend1 = D.baq2_rend1;
        const bool is_assay_amplicon = ((R.dflag & 0x4) || ((par.primerlen > 0) && !(0x2 & par.primer_flag)));
        const bool normal_filters_primers = (par.tn_is_paired && (0x1 & par.primer_flag));
        P.bits = (isrc ? UVC_PR_ISRC : 0u) | ((R.flag & 0x1) ? UVC_PR_PAIRED : 0u) | ((R.flag & 0x8) ? UVC_PR_MATE_UNMAPPED : 0u) | (R.strand ? UVC_PR_STRAND : 0u)
            | ((R.isize != 0) ? UVC_PR_HAS_ISIZE : 0u) | (is_assay_amplicon ? UVC_PR_AMPLICON : 0u) | ((R.dflag & 0x1) ? UVC_PR_UMI : 0u)
            | ((0 == D.clip_cnt) ? UVC_PR_NOCLIP : 0u) | (R.simple ? UVC_PR_SIMPLE : 0u) | ((D.nge_cnt > 0) ? UVC_PR_HAS_GAPS : 0u)
            | ((!(normal_filters_primers || !is_assay_amplicon)) ? UVC_PR_MASK_ON : 0u)
            | (((R.isize != 0) || (0 == (R.flag & 0x1))) ? UVC_PR_IS_NORMAL : 0u) | (((0 == (R.flag & 0x8)) || (0 == (R.flag & 0x1))) ? UVC_PR_MATE_OK : 0u)
            | ((uint32_t)R.mapq << 16) | ((uint32_t)D.micro_nogap_penal << 24);
        P.seq_off = (uint32_t)R.seq_off; P.qual_off = (uint32_t)R.qual_off; P.cx_off = (R.simple ? R.m_qoff : R.cx_off);
        P.terms_lo = (uint32_t)D.xm_term | ((uint32_t)D.bm_term[0] << 7) | ((uint32_t)D.bm_term[1] << 14) | ((uint32_t)D.bm_term[2] << 21);
        P.terms_hi = (uint32_t)D.bm_term[3] | ((uint32_t)D.bm_term[4] << 7);
        P.ibeg = D.ibeg; P.iend = D.iend;
        P.l_qseq = R.l_qseq;
        v.prec[ri] = P;
        PrepRec Q;
        Q.xm1500 = D.xm1500; Q.go1500 = D.go1500; Q.avg_gaplen = D.avg_gaplen; Q.inslen_sum = D.inslen_sum; Q.dellen_sum = D.dellen_sum;
        Q.insbaq_sum = D.insbaq_sum; Q.delbaq_sum = D.delbaq_sum; Q.flags = ((R.dflag & 0x4) ? 1u : 0u);
        v.qrec[ri] = Q;
    }

    if (!R.simple) {
        // Replay of the bias walk's bookkeeping (main.hpp:1817-1859, 1886-2257): for every reference base of the read, what it
        // aligns to and which low-quality indels surround it according to the reference's one-step-per-base cursor.
        CxEntry *cx = v.cx + R.cx_off;
        for (int32_t o = 0; o < span; o++) { cx[o].qpos = -1; cx[o].flags = 0; cx[o].pad = 0; cx[o].prev_rpos = 0; cx[o].next_rpos = INT32_MAX; }
        // the list of low-quality indel positions, materialised in the events' scratch (symbol field) as flags
        int32_t n_low = 0;
        {
            int32_t qpos = 0, rpos = pos, e = 0;
            for (int32_t i = 0; i < R.n_cigar; i++) {
                const int op = cig_op(cigar[i]); const int32_t len = cig_len(cigar[i]);
                if (is_match_op(op)) { qpos += len; rpos += len; }
                else if (op == UVC_CINS) {
                    bool low = false;
                    // QUIRK: the upper bound mixes a query index with the reference coordinate rend (main.hpp:1841)
                    for (int32_t q2 = qpos - tmin(qpos, 1); q2 < tmin(qpos + len + 1, rend); q2++) {
                        if (q2 < R.l_qseq && (int32_t)qual[q2] < par.bias_thres_interfering_indel_BQ) { low = true; }
                    }
                    v.ev[R.ev_off + e].counted = (low ? 1 : 0); if (low) { n_low++; }
                    e++; qpos += len;
                } else if (op == UVC_CDEL) {
                    const int32_t qa = tmax(1, qpos) - 1;
                    const int32_t qb = tmin(qpos, R.l_qseq - 1);
                    const bool low = (tmin((int32_t)qual[qa], (int32_t)qual[qb]) <= par.bias_thres_interfering_indel_BQ);
                    v.ev[R.ev_off + e].counted = (low ? 1 : 0); if (low) { n_low++; }
                    e++; rpos += len;
                } else if (op == UVC_CREF_SKIP) { rpos += len; } else if (op == UVC_CSOFT_CLIP) { qpos += len; }
            }
        }
        // cursor replay: list = {0, low-quality indel rposs..., INT32_MAX}; idx starts at 0
        int32_t idx = 0;           // index into the list
        int32_t ev_cursor = 0;     // event index of list[idx] - 1 bookkeeping is done through the helper below
        (void)ev_cursor;
        // helper: value of list[k]
        #define UVC_LOWQ_LIST(k, out) { \
            if ((k) <= 0) { (out) = 0; } else if ((k) > n_low) { (out) = INT32_MAX; } else { \
                int32_t seen_ = 0; (out) = INT32_MAX; \
                for (int32_t e_ = 0; e_ < R.n_ev; e_++) { if (v.ev[R.ev_off + e_].counted) { seen_++; if (seen_ == (k)) { (out) = v.ev[R.ev_off + e_].rpos; break; } } } } }
        int32_t qpos = 0, rpos = pos;
        for (int32_t i = 0; i < R.n_cigar; i++) {
            const int op = cig_op(cigar[i]); const int32_t len = cig_len(cigar[i]);
            if (is_match_op(op)) {
                for (int32_t j = 0; j < len; j++) {
                    CxEntry & c = cx[rpos - pos];
                    c.qpos = (int16_t)qpos; c.flags = (uint8_t)(1 | (j > 0 ? 2 : 0));
                    if (!primer_masked(v, R, D, rpos) && nge > 0) {
                        int32_t cur; UVC_LOWQ_LIST(idx, cur);
                        if (cur <= rpos) { idx++; }
                        int32_t a, b; UVC_LOWQ_LIST(idx - 1, a); UVC_LOWQ_LIST(idx, b);
                        c.prev_rpos = a; c.next_rpos = b;
                    }
                    qpos++; rpos++;
                }
            } else if (op == UVC_CINS) {
                qpos += len;
            } else if (op == UVC_CDEL) {
                const int32_t nbases2end = tmin(qpos, R.l_qseq - qpos);
                const bool counted = (!primer_masked(v, R, D, rpos)) && (nbases2end >= par.indel_filter_edge_dist);
                for (int32_t r2 = rpos; r2 < rpos + len; r2++) {
                    CxEntry & c = cx[r2 - pos];
                    c.flags = 4;
                    if (counted && r2 < tmin(rpos + len, rend)) {
                        // BASE_NN at r2, then LINK_NN at r2 + 1 (main.hpp:2219-2252); each visit advances the cursor at most once
                        for (int s = 0; s < 2; s++) {
                            const int32_t pp = (s == 0 ? r2 : r2 + 1);
                            if (pp >= rend) { continue; }
                            int32_t cur; UVC_LOWQ_LIST(idx, cur);
                            if (cur <= rpos) { idx++; }
                            int32_t a, b; UVC_LOWQ_LIST(idx - 1, a); UVC_LOWQ_LIST(idx, b);
                            const uint32_t d1 = (uint32_t)rpos - (uint32_t)a, d2 = (uint32_t)b - (uint32_t)rpos;
                            const int32_t dist = (int32_t)(d1 < d2 ? d1 : d2);
                            if (s == 0) { c.prev_rpos = dist; } else { c.next_rpos = dist; }
                        }
                    }
                }
                rpos += len;
            } else if (op == UVC_CREF_SKIP) { rpos += len; } else if (op == UVC_CSOFT_CLIP) { qpos += len; }
        }
        #undef UVC_LOWQ_LIST
        for (int32_t e = 0; e < R.n_ev; e++) { v.ev[R.ev_off + e].counted = 0; }
    }
}

#define UVC_K2_NOBASE 0xffffu
// index of the read's aligned base at p in its query sequence, -1 if there is none (or !mine)
UVC_HD int32_t base_index(const BatchView & v, const ReadRec & R, int32_t p, bool mine) {
    const int32_t o = p - R.pos;
    if (!(mine && o >= 0 && p < R.rend && R.l_qseq > 0)) { return -1; }
    if (R.simple) { return tmin(R.m_qoff + o, R.l_qseq - 1); }
    const CxEntry e = v.cx[(int64_t)R.cx_off + o];
    return ((e.flags & 1) ? tmax(0, tmin((int32_t)e.qpos, R.l_qseq - 1)) : -1);
}
UVC_HD int32_t base_index(const BatchView & v, const PileRec & P, int32_t p, bool mine) {
    const int32_t o = p - P.pos;
    if (!(mine && o >= 0 && p < P.rend && P.l_qseq > 0)) { return -1; }
    if (P.bits & UVC_PR_SIMPLE) { return tmin(P.cx_off + o, P.l_qseq - 1); }
    const CxEntry e = v.cx[(int64_t)P.cx_off + o];
    return ((e.flags & 1) ? tmax(0, tmin((int32_t)e.qpos, P.l_qseq - 1)) : -1);
}
UVC_HD uint32_t pack_base(uint32_t seq_byte, uint32_t qual_byte, int32_t qpos) {
    const uint32_t sym = (uint32_t)nt16_to_symbol(seq_byte >> ((~qpos & 1) << 2));
    return (qpos >= 0 ? ((sym << 8) | qual_byte) : UVC_K2_NOBASE);
}
// ------------------------------------------------------------------------------------------------ K1: one thread per position
// Dense part of walk #1 (main.hpp:1006-1068) gathered per position, then the threshold pass (main.hpp:1206-1299).
struct K1State {
    int64_t gp;
    int32_t p, baq_p;
    uvcgpu_prep_set a;
};
UVC_HD void k1_begin(K1State & s, const BatchView & v, int64_t gp) {
    const TileInfo & T = v.tiles[v.pos_tile[gp]];
    s.gp = gp;
    s.p = (int32_t)(gp - T.pos_off) + T.ext_beg;
    s.a = v.prep[gp];     // starts from the rare-event contributions of K0
    s.baq_p = v.baq[gp];
}
// one read of the window, from its compact records; `packed` = (symbol << 8) | quality of its aligned base at p, UVC_K2_NOBASE if it shows none
UVC_HD void k1_read(K1State & s, const BatchView & v, const PileRec & P, const PrepRec & Q, uint32_t packed) {
    if (packed == UVC_K2_NOBASE) { return; }
    uvcgpu_prep_set & a = s.a;
    const int32_t p = s.p;
    const int32_t span = P.rend - P.pos;
    a.a_pcr_dp += (int32_t)(Q.flags & 1u);
    a.a_umi_dp += ((P.bits & UVC_PR_UMI) ? 1 : 0);
    a.a_dp += 1;
    a.a_qlen += span;
    a.a_XM1500 += Q.xm1500; a.a_GO1500 += Q.go1500; a.a_GAPLEN += Q.avg_gaplen;
    // (conditions as 0/1 factors instead of branches, as in segbias)
    {
        const int32_t has = ((P.bits & UVC_PR_HAS_ISIZE) ? 1 : 0), rc = ((P.bits & UVC_PR_ISRC) ? 1 : 0);
        const int32_t li = has * rc, ri = has * (1 - rc);
        a.a_LI += (int64_t)(li * tmin(p - P.frag_l + 1, UVC_MAX_INSERT_SIZE)); a.a_LIDP += li;
        a.a_RI += (int64_t)(ri * tmin(P.frag_r - p, UVC_MAX_INSERT_SIZE)); a.a_RIDP += ri;
    }
    {
        const int32_t hq = (((int32_t)(packed & 0xffu) >= v.par.bias_thres_highBQ) ? 1 : 0);
        a.a_l_dist_sum += hq * (p - P.pos + 1);
        a.a_r_dist_sum += hq * (P.rend - p);
        a.a_inslen_sum += hq * Q.inslen_sum; a.a_dellen_sum += hq * Q.dellen_sum;
        a.a_l_BAQ_sum += (int64_t)(hq * (s.baq_p - P.baq_pos + 1));
        a.a_r_BAQ_sum += (int64_t)(hq * (P.baq_rend1 - s.baq_p + 1));
        a.a_insBAQ_sum += (int64_t)(hq * Q.insbaq_sum); a.a_delBAQ_sum += (int64_t)(hq * Q.delbaq_sum);
        a.a_highBQ_dp += hq;
    }
}
UVC_HD void k1_end(K1State & s, const BatchView & v);
UVC_HD void k1_position(const BatchView & v, int64_t gp, const Win & w) {
    K1State s;
    k1_begin(s, v, gp);
    for (int64_t ri = w.ulo; ri < w.uhi; ri++) {
        if (ri < w.lo || ri >= w.hi) { continue; }
        const PileRec & P = v.prec[ri];
        const int32_t qpos = base_index(v, P, s.p, true);
        k1_read(s, v, P, v.qrec[ri], pack_base(0, v.qual[(uint64_t)P.qual_off + (uint32_t)tmax(qpos, 0)], qpos));
    }
    k1_end(s, v);
}
// the threshold pass (main.hpp:1206-1299) and the store of the position's prep set
UVC_HD void k1_end(K1State & s, const BatchView & v) {
    const uvcgpu_params & par = v.par;
    const int64_t gp = s.gp;
    const uvcgpu_prep_set & a = s.a;
    v.prep[gp] = a;

    uvcgpu_thres_set t;
    const int32_t segLIDP = tmax(a.a_LIDP, 1), segRIDP = tmax(a.a_RIDP, 1);
    const double ins_l = ceil(sqrt((double)(a.a_near_ins_l_pow2len / tmax(a.a_near_ins_dp, 1))));
    const double del_l = ceil(sqrt((double)(a.a_near_del_l_pow2len / tmax(a.a_near_del_dp, 1))));
    const double ins_r = ceil(sqrt((double)(a.a_near_ins_r_pow2len / tmax(a.a_near_ins_dp, 1))));
    const double del_r = ceil(sqrt((double)(a.a_near_del_r_pow2len / tmax(a.a_near_del_dp, 1))));
    const double dnv_border = 0; // the 10-base DNV border applies to IonTorrent only (main.hpp:1234-1235)
    t.aLPxT = (int32_t)(tmax(ins_l, tmax(del_l, dnv_border)) + par.bias_thres_aLPxT_add);
    t.aRPxT = (int32_t)(tmax(ins_r, tmax(del_r, dnv_border)) + par.bias_thres_aLPxT_add);
    int32_t indelphred = v.rtr[gp].indelphred;
    if (a.a_near_ins_dp * par.indel_del_to_ins_err_ratio < a.a_near_del_dp) { indelphred += v.indelphred_half; }
    if (a.a_near_del_dp * par.indel_del_to_ins_err_ratio < a.a_near_ins_dp) { indelphred -= v.indelphred_half; }
    const int32_t pc_inc1 = (int32_t)(3 * 100 * tmax(1, a.a_near_ins_dp + a.a_near_del_dp) / (tmax(1, a.a_near_ins_inv100len + a.a_near_del_inv100len))) - 3;
    indelphred += between(pc_inc1, 0, 6);
    indelphred = tmax(indelphred, 0);
    v.rtr[gp].indelphred = indelphred;
    const bool is_normal = par.is_tumor_vcf_provided;
    const int64_t LI1T = (is_normal ? par.bias_thres_aLRI1NT_perc : par.bias_thres_aLRI1T_perc);
    const int64_t LI1t = (is_normal ? par.bias_thres_aLRI1Nt_perc : par.bias_thres_aLRI1t_perc);
    t.aLI1T = (int32_t)(a.a_LI * LI1T / (segLIDP * 100) + par.bias_thres_aLRI1T_add);
    t.aLI2T = (int32_t)(a.a_LI * (int64_t)par.bias_thres_aLRI2T_perc / (segLIDP * 100) + par.bias_thres_aLRI2T_add);
    t.aLI1t = (int32_t)(a.a_LI * LI1t / (segLIDP * 100));
    t.aLI2t = (int32_t)(a.a_LI * (int64_t)par.bias_thres_aLRI2t_perc / (segLIDP * 100));
    t.aRI1T = (int32_t)(a.a_RI * LI1T / (segRIDP * 100) + par.bias_thres_aLRI1T_add);
    t.aRI2T = (int32_t)(a.a_RI * (int64_t)par.bias_thres_aLRI2T_perc / (segRIDP * 100) + par.bias_thres_aLRI2T_add);
    t.aRI1t = (int32_t)(a.a_RI * LI1t / (segRIDP * 100));
    t.aRI2t = (int32_t)(a.a_RI * (int64_t)par.bias_thres_aLRI2t_perc / (segRIDP * 100));
    const int64_t P1 = (is_normal ? par.bias_thres_aLRP1Nt_avgmul_perc : par.bias_thres_aLRP1t_avgmul_perc);
    const int64_t P2 = par.bias_thres_aLRP2t_avgmul_perc;
    const int64_t B1 = (is_normal ? par.bias_thres_aLRB1Nt_avgmul_perc : par.bias_thres_aLRB1t_avgmul_perc);
    const int64_t B2 = par.bias_thres_aLRB2t_avgmul_perc;
    const int64_t hb100 = tmax(1, a.a_highBQ_dp * 100);
    #define UVC_NNM64(x, y) ((x) > (y) ? ((x) - (y)) : 0)
    t.aLP1t = (int32_t)UVC_NNM64((int64_t)a.a_l_dist_sum * P1 / hb100, (int64_t)par.bias_thres_aLRP1t_minus);
    t.aLP2t = (int32_t)UVC_NNM64((int64_t)a.a_l_dist_sum * P2 / hb100, (int64_t)par.bias_thres_aLRP2t_minus);
    t.aRP1t = (int32_t)UVC_NNM64((int64_t)a.a_r_dist_sum * P1 / hb100, (int64_t)par.bias_thres_aLRP1t_minus);
    t.aRP2t = (int32_t)UVC_NNM64((int64_t)a.a_r_dist_sum * P2 / hb100, (int64_t)par.bias_thres_aLRP2t_minus);
    const int64_t pdel = a.a_delBAQ_sum / tmax(1, a.a_highBQ_dp);
    t.aLB1t = (int32_t)UVC_NNM64(a.a_l_BAQ_sum * B1 / hb100, par.bias_thres_aLRB1t_minus + pdel);
    t.aLB2t = (int32_t)UVC_NNM64(a.a_l_BAQ_sum * B2 / hb100, (int64_t)par.bias_thres_aLRB2t_minus);
    t.aRB1t = (int32_t)UVC_NNM64(a.a_r_BAQ_sum * B1 / hb100, par.bias_thres_aLRB1t_minus + pdel);
    t.aRB2t = (int32_t)UVC_NNM64(a.a_r_BAQ_sum * B2 / hb100, (int64_t)par.bias_thres_aLRB2t_minus);
    #undef UVC_NNM64
    v.thres[gp] = t;
}

// ------------------------------------------------------------------------------------------------ segment bias (dealwith_segbias)
struct SegAcc {
    uvcgpu_seginfo_set s;
    int32_t a1BQf, a1BQr, a2BQf, a2BQr;
    int32_t bqsum;
};

UVC_HD void segacc_zero(SegAcc & a) {
    a.s.a2XM2 = a.s.a2BM2 = a.s.aPF1 = a.s.aPF2 = a.s.aBQ2 = a.s.aMQs = a.s.aP1 = a.s.aP2 = a.s.aP3 = a.s.aNC = 0;
    a.s.aDPff = a.s.aDPfr = a.s.aDPrf = a.s.aDPrr = 0;
    a.s.aLP1 = a.s.aLP2 = a.s.aLPL = a.s.aRP1 = a.s.aRP2 = a.s.aRPL = 0;
    a.s.aLB1 = a.s.aLB2 = 0; a.s.aLBL = 0; a.s.aRB1 = a.s.aRB2 = 0; a.s.aRBL = 0;
    a.s.aLI1 = a.s.aLI2 = a.s.aRI1 = a.s.aRI2 = a.s.aRIf = a.s.aLIr = 0; a.s.aLIT = 0; a.s.aRIT = 0;
    a.a1BQf = a.a1BQr = a.a2BQf = a.a2BQr = 0; a.bqsum = 0;
}

// adds an accumulator to the global records of (position gp, symbol sym); kAtomic for the per-read event kernels
template <bool kAtomic>
UVC_HD void segacc_flush(const BatchView & v, int64_t gp, int sym, const SegAcc & a) {
    uvcgpu_seginfo_set & g = v.seginfo[gp * UVC_NSYM + sym];
    int32_t *vq = v.vq + (gp * UVC_NSYM + sym) * UVCGPU_NUM_VQ_TAGS;
    int32_t *bq = v.bqsum + gp * UVC_NSYM + sym;
    #define UVC_ADD(dst, val) { if ((val) != 0) { if (kAtomic) { atomic_add(&(dst), (val)); } else { (dst) += (val); } } }
    UVC_ADD(g.a2XM2, a.s.a2XM2) UVC_ADD(g.a2BM2, a.s.a2BM2) UVC_ADD(g.aPF1, a.s.aPF1) UVC_ADD(g.aPF2, a.s.aPF2) UVC_ADD(g.aBQ2, a.s.aBQ2)
    UVC_ADD(g.aMQs, a.s.aMQs) UVC_ADD(g.aP1, a.s.aP1) UVC_ADD(g.aP2, a.s.aP2) UVC_ADD(g.aP3, a.s.aP3) UVC_ADD(g.aNC, a.s.aNC)
    UVC_ADD(g.aDPff, a.s.aDPff) UVC_ADD(g.aDPfr, a.s.aDPfr) UVC_ADD(g.aDPrf, a.s.aDPrf) UVC_ADD(g.aDPrr, a.s.aDPrr)
    UVC_ADD(g.aLP1, a.s.aLP1) UVC_ADD(g.aLP2, a.s.aLP2) UVC_ADD(g.aLPL, a.s.aLPL) UVC_ADD(g.aRP1, a.s.aRP1) UVC_ADD(g.aRP2, a.s.aRP2) UVC_ADD(g.aRPL, a.s.aRPL)
    UVC_ADD(g.aLB1, a.s.aLB1) UVC_ADD(g.aLB2, a.s.aLB2) UVC_ADD(g.aLBL, a.s.aLBL) UVC_ADD(g.aRB1, a.s.aRB1) UVC_ADD(g.aRB2, a.s.aRB2) UVC_ADD(g.aRBL, a.s.aRBL)
    UVC_ADD(g.aLI1, a.s.aLI1) UVC_ADD(g.aLI2, a.s.aLI2) UVC_ADD(g.aRI1, a.s.aRI1) UVC_ADD(g.aRI2, a.s.aRI2) UVC_ADD(g.aRIf, a.s.aRIf) UVC_ADD(g.aLIr, a.s.aLIr)
    UVC_ADD(g.aLIT, a.s.aLIT) UVC_ADD(g.aRIT, a.s.aRIT)
    UVC_ADD(vq[0], a.a1BQf) UVC_ADD(vq[1], a.a1BQr) UVC_ADD(vq[2], a.a2BQf) UVC_ADD(vq[3], a.a2BQr)
    UVC_ADD(*bq, a.bqsum)
    #undef UVC_ADD
}

UVC_HD void bidir_bias(int32_t & lp1, int32_t & lp2, int32_t & rp1, int32_t & rp2, int64_t & lpl, int64_t & rpl,
        int32_t L1, int32_t L2, int32_t R1, int32_t R2, int32_t nl, int32_t nr, bool tier2, int32_t n_indel) {
    // update_bidirectional_bias (main.hpp:1318-1358)
    if (nl + n_indel >= L1) { lp1 += 1; }
    if ((nl + n_indel >= L2) && tier2) { lp2 += 1; }
    if (nr >= R1) { rp1 += 1; }
    if ((nr >= R2) && tier2) { rp2 += 1; }
    lpl += nl; rpl += nr;
}

// The same update on records in global memory that only this thread writes, as fire-and-forget reductions (no read-modify-write round trips)
UVC_HD void bidir_bias_red(int32_t *lp1, int32_t *lp2, int32_t *rp1, int32_t *rp2, int32_t L1, int32_t L2, int32_t R1, int32_t R2,
        int32_t nl, int32_t nr, bool tier2, int32_t n_indel) {
    if (nl + n_indel >= L1) { atomic_add(lp1, 1); }
    if ((nl + n_indel >= L2) && tier2) { atomic_add(lp2, 1); }
    if (nr >= R1) { atomic_add(rp1, 1); }
    if ((nr >= R2) && tier2) { atomic_add(rp2, 1); }
}

// One call of dealwith_segbias<isGap> (main.hpp:1360-1595) for read R at position rpos with quality bq, into accumulator a.
template <bool isGap>
UVC_HD void segbias(SegAcc & a, const BatchView & v, const ReadRec & R, const ReadDerived & D, const uvcgpu_thres_set & th,
        int32_t baq_rpos, int32_t baq2_rpos, int32_t bq, int32_t rpos, int32_t bm_term, bool is_ins_op, int32_t indel_len, int32_t dist_indel) {
    const uvcgpu_params & par = v.par;
    const bool is_assay_amplicon = ((R.dflag & 0x4) || ((par.primerlen > 0) && !(0x2 & par.primer_flag)));
    const bool normal_filters_primers = (par.tn_is_paired && (0x1 & par.primer_flag));
    const bool is_assay_UMI = (R.dflag & 0x1);
    const int32_t pos = R.pos, rend = R.rend;
    const int32_t seg_l_baq1 = baq_rpos - D.baq_pos + 1;
    const int32_t seg_r_baq0 = D.baq_rend1 - baq_rpos + 1;
    const int32_t seg_r_baq1 = (isGap ? tmin(seg_r_baq0, D.baq2_rend1 - baq2_rpos + 7) : seg_r_baq0);
    const int32_t seg_l_nbases = rpos - pos + 1;
    const int32_t seg_r_nbases = rend - rpos;
    const bool is_high_readlen = (par.central_readlen >= par.microadjust_median_readlen_thres);
    const int32_t seg_l_baq = (is_high_readlen ? seg_l_baq1 : tmax(seg_l_baq1, seg_l_nbases * par.microadjust_BAQ_per_base_x1024 / 1024));
    const int32_t seg_r_baq = (is_high_readlen ? seg_r_baq1 : tmax(seg_r_baq1, seg_r_nbases * par.microadjust_BAQ_per_base_x1024 / 1024));
    const int32_t frag_pos_L = tmin(pos, R.mpos);
    const int32_t frag_pos_R = frag_pos_L + iabs(R.isize);
    const int32_t frag_l_nb = ((R.isize != 0) ? tmin(rpos - frag_pos_L + 1, UVC_MAX_INSERT_SIZE) : UVC_MAX_INSERT_SIZE);
    const int32_t frag_r_nb = ((R.isize != 0) ? tmin(frag_pos_R - rpos, UVC_MAX_INSERT_SIZE) : UVC_MAX_INSERT_SIZE);
    const bool is_normal = ((R.isize != 0) || (0 == (R.flag & 0x1)));
    const bool isrc = ((R.flag & 0x10) == 0x10);
    const bool strand = R.strand;

    // From here on the updates are written without control flow (conditions become 0/1 addends, non-short-circuit & and |): the conditions on the
    // read (strand, orientation, pairing) are the same for all lanes, but a taken-or-not branch per counter costs more than the add it guards
    // when only a few warps are resident.
    const int32_t rc = (isrc ? 1 : 0), fw = 1 - rc;
    const int32_t has_isize = ((R.isize != 0) ? 1 : 0);
    const int32_t sq = bq * bq / UVC_SQR_QUAL_DIV;
    a.a1BQr += rc * bq; a.a2BQr += rc * sq; a.a1BQf += fw * bq; a.a2BQf += fw * sq;
    a.s.aMQs += R.mapq;
    const int32_t st1 = (strand ? 1 : 0), st0 = 1 - st1;
    a.s.aDPrr += st1 * rc; a.s.aDPrf += st1 * fw; a.s.aDPfr += st0 * rc; a.s.aDPff += st0 * fw;
    a.s.aP3 += ((tmin(dist_indel, tmin(seg_l_nbases, seg_r_nbases)) >= par.bias_thres_interfering_indel) ? 1 : 0);
    a.s.aNC += ((0 == D.clip_cnt) ? 1 : 0);
    a.s.aLIT += (int64_t)(rc * has_isize * frag_l_nb); a.s.aRIT += (int64_t)(fw * has_isize * frag_r_nb);

    const int32_t LPxT0 = th.aLPxT, RPxT = th.aRPxT;
    const int32_t LPxT = (isGap ? LPxT0 : tmin(LPxT0, RPxT));
    const bool far_from_edge = (seg_l_nbases + (is_ins_op ? nnminus(indel_len, par.microadjust_nobias_pos_indel_maxlen) : 0) >= LPxT) & (seg_r_nbases >= RPxT);
    const int32_t highBAQ = par.bias_thres_highBAQ + (isGap ? 0 : 3);
    const bool unaffected_by_edge = (seg_l_baq >= highBAQ) & (seg_r_baq >= highBAQ);
    const int32_t min_dist2iend = ((R.flag & 0x1) ? tmin(frag_l_nb, frag_r_nb) : (isrc ? seg_r_nbases : seg_l_nbases));
    a.s.aP1 += ((far_from_edge & unaffected_by_edge & ((min_dist2iend > par.primerlen2) | !is_assay_amplicon)) ? 1 : 0);
    a.s.aP2 += ((is_assay_UMI | !is_assay_amplicon) ? 1 : 0);

    int32_t f1, f2;
    if ((uint32_t)bq < 128u) { f1 = v.pf_tab[bq]; f2 = v.pf_tab[128 + bq]; }
    else {
        f1 = ((bq < par.bias_thres_PFBQ1) ? (100 * (bq * bq) / (par.bias_thres_PFBQ1 * par.bias_thres_PFBQ1)) : 100);
        f2 = ((bq < par.bias_thres_PFBQ2) ? (100 * (bq * bq) / (par.bias_thres_PFBQ2 * par.bias_thres_PFBQ2)) : 100);
    }
    if (isGap) {
        a.s.aPF1 += tmin(100, f1);
        a.s.aPF2 += tmin(100, f2);
    } else {
        a.s.aPF1 += (100 * f1 / 100);
        a.s.aPF2 += (100 * f2 / 100);
        a.s.a2XM2 += D.xm_term;
        a.s.a2BM2 += bm_term;
    }
    {
        // update_bidirectional_bias (main.hpp:1318-1358) twice: position bias (if far from the edges) and BAQ bias (if unaffected by the edges)
        const bool counted = (isGap ? (dist_indel >= par.bias_thres_interfering_indel) : (bq >= par.bias_thres_highBQ));
        const bool tier2 = (isGap | (bq >= par.bias_thres_highBQ));
        const int32_t gp_ = ((counted & far_from_edge) ? 1 : 0), gb_ = ((counted & unaffected_by_edge) ? 1 : 0), t2 = (tier2 ? 1 : 0);
        const int32_t nl = seg_l_nbases, nr = seg_r_nbases;
        a.s.aLP1 += gp_ * ((nl + indel_len >= th.aLP1t) ? 1 : 0);
        a.s.aLP2 += gp_ * t2 * ((nl + indel_len >= th.aLP2t) ? 1 : 0);
        a.s.aRP1 += gp_ * ((nr >= th.aRP1t) ? 1 : 0);
        a.s.aRP2 += gp_ * t2 * ((nr >= th.aRP2t) ? 1 : 0);
        a.s.aLPL += gp_ * nl; a.s.aRPL += gp_ * nr;
        a.s.aLB1 += gb_ * ((seg_l_baq >= par.bias_thres_BAQ1) ? 1 : 0);
        a.s.aLB2 += gb_ * t2 * ((seg_l_baq >= par.bias_thres_BAQ2) ? 1 : 0);
        a.s.aRB1 += gb_ * ((seg_r_baq >= par.bias_thres_BAQ1) ? 1 : 0);
        a.s.aRB2 += gb_ * t2 * ((seg_r_baq >= par.bias_thres_BAQ2) ? 1 : 0);
        a.s.aLBL += (int64_t)(gb_ * seg_l_baq); a.s.aRBL += (int64_t)(gb_ * seg_r_baq);
        a.s.aBQ2 += (counted ? 1 : 0);
    }
    {
        // insert-size bias: the reverse-complemented read looks at its left fragment end, the forward read at its right one
        const bool mate_ok = ((0 == (R.flag & 0x8)) | (0 == (R.flag & 0x1)));
        const bool nonbiased = (mate_ok & (isrc ? (seg_l_nbases > seg_r_nbases) : (seg_l_nbases < seg_r_nbases)));
        const bool pos_good = ((!is_assay_amplicon) | (!normal_filters_primers) | (far_from_edge & unaffected_by_edge));
        const int32_t d = (isrc ? frag_l_nb : frag_r_nb);
        const int32_t t1 = (isrc ? th.aLI1t : th.aRI1t), T1 = (isrc ? th.aLI1T : th.aRI1T);
        const int32_t t2 = (isrc ? th.aLI2t : th.aRI2t), T2 = (isrc ? th.aLI2T : th.aRI2T);
        const bool ok = (is_normal | (isGap & nonbiased));
        const int32_t c1 = (((d >= t1) & ((d <= T1) | isGap) & ok) ? 1 : 0);
        const int32_t c2 = (((d >= t2) & ((d <= T2) | isGap) & ok & pos_good) ? 1 : 0);
        const int32_t c3 = (pos_good ? 1 : 0);
        a.s.aLI1 += rc * c1; a.s.aLI2 += rc * c2; a.s.aLIr += rc * c3;
        a.s.aRI1 += fw * c1; a.s.aRI2 += fw * c2; a.s.aRIf += fw * c3;
    }
}

// distance of an aligned base to the nearest low-quality indel of its own read (main.hpp:1897-1916)
UVC_HD int32_t dist_to_interfering_indel(const BatchView & v, const TileInfo & T, const ReadDerived & D, const Locus & L, const uvcgpu_thres_set & th, int32_t rpos) {
    if (D.nge_cnt <= 0) { return 10000; }
    const int32_t adj = v.par.indel_adj_tracklen_dist;
    const int32_t npos = T.ext_end - T.ext_beg;
    const uvcgpu_rtr *rtr = v.rtr + T.pos_off;
    const int32_t ridx = rpos - T.ext_beg;
    const uvcgpu_rtr rtr1 = rtr[tmax(ridx, adj) - adj];
    const uvcgpu_rtr rtr2 = rtr[tmin(ridx + adj, npos - 1)];
    const int32_t prevlen = nnminus(rpos - L.prev_rpos, tmax(rpos - (T.ext_beg + rtr1.begpos), th.aLP1t));
    const int32_t nextlen = nnminus(L.next_rpos - rpos, tmax((T.ext_beg + rtr2.begpos + rtr2.tracklen) - rpos, th.aRP1t));
    return tmin(prevlen, nextlen);
}

// K1b, one thread per position (after K1 has adjusted every indelphred): the junction quality both neighbours agree on
UVC_HD void k1b_noindel(const BatchView & v, int64_t gp) {
    v.noindel[gp] = tmin(v.rtr[gp > 0 ? gp - 1 : 0].indelphred, v.rtr[gp].indelphred);
}
// quality weight of "no indel" at the junction before an aligned base (main.hpp:1918-1924)
UVC_HD int32_t nogap_weight(const BatchView & v, int64_t gp, const ReadDerived & D) {
    const int32_t noindel = v.noindel[gp];
    return nnminus(tmin(80, noindel), D.micro_nogap_penal) + 1;
}

// ------------------------------------------------------------------------------------------------ positions whose counters can reach the output
// The reference fills its per-position arrays over the whole extended range of a tile's reads (main.cpp:529-530, 569), because its loop nest
// is read-major; it then reads them at the tile's own positions only: the per-position loop (main.cpp:608-1172) visits zerobased_pos in
// [rpos_inclu_beg, rpos_exclu_end] and looks at refpos = zerobased_pos - 1 (base symbols) and zerobased_pos (link symbols); an MGVCF block line
// started inside that range reads the fragment / family depths of up to 1000 positions ahead (main.cpp:666-667). Everything else that crosses
// positions (fragment and family columns, haplotype strings, indel events, the tandem-repeat context, the adjusted indel qualities of K1)
// comes from read-, fragment- or family-major kernels or from K1, which run in full.
// kind 0: seginfo / bqsum / vq of the bias pileup (K2); kind 1: fragment and family depths (K3b, K4).
UVC_HD void tile_need_range(const uvcgpu_params & par, const TileInfo & T, int kind, int32_t & b, int32_t & e) {
    if (T.skipped) { b = e = T.ext_beg; return; }
    b = tmax(T.ext_beg, T.rpos_inclu_beg - 1);
    e = T.rpos_exclu_end + 1;
    if (1 == kind && (par.outvar_flag & 0x8)) { e = T.rpos_exclu_end + 1002; }
    e = tmin(T.ext_end, e);
    if (e < b) { e = b; }
}
// concatenated position index of entry i of list `kind`, or -1 (padding / beyond the list)
UVC_HD int64_t list_position(const BatchView & v, int kind, int64_t i) {
    if (NULL == v.list_tile[kind]) { return (i < v.n_pos ? i : -1); }
    if (i >= v.n_list[kind]) { return -1; }
    const int32_t t = v.list_tile[kind][i >> 5];
    const TileInfo & T = v.tiles[t];
    int32_t b, e;
    tile_need_range(v.par, T, kind, b, e);
    const int64_t p = (int64_t)b + (i - v.list_off[kind][t]);
    return (p < (int64_t)e ? T.pos_off + (p - T.ext_beg) : -1);
}

// ------------------------------------------------------------------------------------------------ K2: bias pileup of the dense symbols
// Walk #2 (updateByAln<SUM, bias> over aligned bases, main.hpp:1890-2008) gathered per position.
// dealwith_segbias<isGap> (main.hpp:1360-1595) for an aligned base (isGap = false) or the gap-free junction before it (isGap = true), from the
// compact record of the read (batch.h: PileRec): same arithmetic as segbias() with is_ins_op = false and indel_len = 0, with the per-read
// conditions taken from P.bits.
template <bool isGap>
UVC_HD void segbias_dense(SegAcc & a, const BatchView & v, const PileRec & P, const uvcgpu_thres_set & th,
        int32_t baq_rpos, int32_t baq2_rpos, int32_t bq, int32_t rpos, int32_t bm_term, int32_t xm_term, int32_t dist_indel, bool normal_filters_primers) {
    const uvcgpu_params & par = v.par;
    const uint32_t bits = P.bits;
    const bool is_assay_amplicon = (bits & UVC_PR_AMPLICON);
    const int32_t seg_l_baq1 = baq_rpos - P.baq_pos + 1;
    const int32_t seg_r_baq0 = P.baq_rend1 - baq_rpos + 1;
    const int32_t seg_r_baq1 = (isGap ? tmin(seg_r_baq0, P.baq2_rend1 - baq2_rpos + 7) : seg_r_baq0);
    const int32_t seg_l_nbases = rpos - P.pos + 1;
    const int32_t seg_r_nbases = P.rend - rpos;
    const bool is_high_readlen = (par.central_readlen >= par.microadjust_median_readlen_thres);
    const int32_t seg_l_baq = (is_high_readlen ? seg_l_baq1 : tmax(seg_l_baq1, seg_l_nbases * par.microadjust_BAQ_per_base_x1024 / 1024));
    const int32_t seg_r_baq = (is_high_readlen ? seg_r_baq1 : tmax(seg_r_baq1, seg_r_nbases * par.microadjust_BAQ_per_base_x1024 / 1024));
    const int32_t has_isize = ((bits & UVC_PR_HAS_ISIZE) ? 1 : 0);
    const int32_t frag_l_nb = (has_isize ? tmin(rpos - P.frag_l + 1, UVC_MAX_INSERT_SIZE) : UVC_MAX_INSERT_SIZE);
    const int32_t frag_r_nb = (has_isize ? tmin(P.frag_r - rpos, UVC_MAX_INSERT_SIZE) : UVC_MAX_INSERT_SIZE);
    const bool isrc = (bits & UVC_PR_ISRC);

    const int32_t rc = (isrc ? 1 : 0), fw = 1 - rc;
    const int32_t sq = bq * bq / UVC_SQR_QUAL_DIV;
    a.a1BQr += rc * bq; a.a2BQr += rc * sq; a.a1BQf += fw * bq; a.a2BQf += fw * sq;
    a.s.aMQs += (int32_t)((bits >> 16) & 0xffu);
    const int32_t st1 = ((bits & UVC_PR_STRAND) ? 1 : 0), st0 = 1 - st1;
    a.s.aDPrr += st1 * rc; a.s.aDPrf += st1 * fw; a.s.aDPfr += st0 * rc; a.s.aDPff += st0 * fw;
    a.s.aP3 += ((tmin(dist_indel, tmin(seg_l_nbases, seg_r_nbases)) >= par.bias_thres_interfering_indel) ? 1 : 0);
    a.s.aNC += ((bits & UVC_PR_NOCLIP) ? 1 : 0);
    a.s.aLIT += (int64_t)(rc * has_isize * frag_l_nb); a.s.aRIT += (int64_t)(fw * has_isize * frag_r_nb);

    const int32_t LPxT0 = th.aLPxT, RPxT = th.aRPxT;
    const int32_t LPxT = (isGap ? LPxT0 : tmin(LPxT0, RPxT));
    const bool far_from_edge = (seg_l_nbases >= LPxT) & (seg_r_nbases >= RPxT);
    const int32_t highBAQ = par.bias_thres_highBAQ + (isGap ? 0 : 3);
    const bool unaffected_by_edge = (seg_l_baq >= highBAQ) & (seg_r_baq >= highBAQ);
    const int32_t min_dist2iend = ((bits & UVC_PR_PAIRED) ? tmin(frag_l_nb, frag_r_nb) : (isrc ? seg_r_nbases : seg_l_nbases));
    a.s.aP1 += ((far_from_edge & unaffected_by_edge & ((min_dist2iend > par.primerlen2) | !is_assay_amplicon)) ? 1 : 0);
    a.s.aP2 += (((bits & UVC_PR_UMI) != 0) | !is_assay_amplicon) ? 1 : 0;

    int32_t f1, f2;
    if ((uint32_t)bq < 128u) { f1 = v.pf_tab[bq]; f2 = v.pf_tab[128 + bq]; }
    else {
        f1 = ((bq < par.bias_thres_PFBQ1) ? (100 * (bq * bq) / (par.bias_thres_PFBQ1 * par.bias_thres_PFBQ1)) : 100);
        f2 = ((bq < par.bias_thres_PFBQ2) ? (100 * (bq * bq) / (par.bias_thres_PFBQ2 * par.bias_thres_PFBQ2)) : 100);
    }
    if (isGap) {
        a.s.aPF1 += tmin(100, f1);
        a.s.aPF2 += tmin(100, f2);
    } else {
        a.s.aPF1 += (100 * f1 / 100);
        a.s.aPF2 += (100 * f2 / 100);
        a.s.a2XM2 += xm_term;
        a.s.a2BM2 += bm_term;
    }
    {
        const bool counted = (isGap ? (dist_indel >= par.bias_thres_interfering_indel) : (bq >= par.bias_thres_highBQ));
        const bool tier2 = (isGap | (bq >= par.bias_thres_highBQ));
        const int32_t gp_ = ((counted & far_from_edge) ? 1 : 0), gb_ = ((counted & unaffected_by_edge) ? 1 : 0), t2 = (tier2 ? 1 : 0);
        const int32_t nl = seg_l_nbases, nr = seg_r_nbases;
        a.s.aLP1 += gp_ * ((nl >= th.aLP1t) ? 1 : 0);
        a.s.aLP2 += gp_ * t2 * ((nl >= th.aLP2t) ? 1 : 0);
        a.s.aRP1 += gp_ * ((nr >= th.aRP1t) ? 1 : 0);
        a.s.aRP2 += gp_ * t2 * ((nr >= th.aRP2t) ? 1 : 0);
        a.s.aLPL += gp_ * nl; a.s.aRPL += gp_ * nr;
        a.s.aLB1 += gb_ * ((seg_l_baq >= par.bias_thres_BAQ1) ? 1 : 0);
        a.s.aLB2 += gb_ * t2 * ((seg_l_baq >= par.bias_thres_BAQ2) ? 1 : 0);
        a.s.aRB1 += gb_ * ((seg_r_baq >= par.bias_thres_BAQ1) ? 1 : 0);
        a.s.aRB2 += gb_ * t2 * ((seg_r_baq >= par.bias_thres_BAQ2) ? 1 : 0);
        a.s.aLBL += (int64_t)(gb_ * seg_l_baq); a.s.aRBL += (int64_t)(gb_ * seg_r_baq);
        a.s.aBQ2 += (counted ? 1 : 0);
    }
    {
        const bool mate_ok = (bits & UVC_PR_MATE_OK);
        const bool nonbiased = (mate_ok & (isrc ? (seg_l_nbases > seg_r_nbases) : (seg_l_nbases < seg_r_nbases)));
        const bool pos_good = ((!is_assay_amplicon) | (!normal_filters_primers) | (far_from_edge & unaffected_by_edge));
        const int32_t d = (isrc ? frag_l_nb : frag_r_nb);
        const int32_t t1 = (isrc ? th.aLI1t : th.aRI1t), T1 = (isrc ? th.aLI1T : th.aRI1T);
        const int32_t t2 = (isrc ? th.aLI2t : th.aRI2t), T2 = (isrc ? th.aLI2T : th.aRI2T);
        const bool ok = (((bits & UVC_PR_IS_NORMAL) != 0) | (isGap & nonbiased));
        const int32_t c1 = (((d >= t1) & ((d <= T1) | isGap) & ok) ? 1 : 0);
        const int32_t c2 = (((d >= t2) & ((d <= T2) | isGap) & ok & pos_good) ? 1 : 0);
        const int32_t c3 = (pos_good ? 1 : 0);
        a.s.aLI1 += rc * c1; a.s.aLI2 += rc * c2; a.s.aLIr += rc * c3;
        a.s.aRI1 += fw * c1; a.s.aRI2 += fw * c2; a.s.aRIf += fw * c3;
    }
}

// ------------------------------------------------------------------------------------------------ K2, merged roles: one thread per position
// Both symbols that nearly every read votes for at a position - the reference base and LINK_M - are handled by ONE thread: what
// dealwith_segbias derives from (read, position) alone (distances to the read and fragment ends, BAQ spans, threshold comparisons, strand and
// orientation classes) is computed once and feeds both updates. The 23 pure event counters of the two symbols share registers: the base
// symbol counts in the low and LINK_M in the high 16 bits of one word, so one predicated add updates both; the remaining sums are 32-bit per
// symbol. Every 60 000 visited reads the accumulators are flushed (no half-word can overflow before that).
#define UVC_K2M_FLUSH_EVERY 60000
struct K2Merged {
    const TileInfo *T;
    int64_t gp;
    int32_t p, major;
    int32_t baq_p, baq2_p, noindel;
    int32_t nvis;
    uvcgpu_thres_set th;
    // packed counters: base symbol | LINK_M << 16
    uint32_t cP1, cP2, cP3, cNC, cDPff, cDPfr, cDPrf, cDPrr, cLP1, cLP2, cRP1, cRP2, cLB1, cLB2, cRB1, cRB2, cLI1, cLI2, cRI1, cRI2, cRIf, cLIr, cBQ2;
    // sums, [0] = base symbol, [1] = LINK_M (separate scalars: indexing would force the struct into local memory)
    int32_t MQs0, MQs1, LPL0, LPL1, RPL0, RPL1, PF10, PF11, PF20, PF21, XM2, BM2, LBL0, LBL1, RBL0, RBL1, LIT0, LIT1, RIT0, RIT1;
    int32_t q1f0, q1f1, q1r0, q1r1, q2f0, q2f1, q2r0, q2r1, bqs0, bqs1;
};
UVC_HD void k2m_zero(K2Merged & s) {
    s.nvis = 0;
    s.cP1 = s.cP2 = s.cP3 = s.cNC = s.cDPff = s.cDPfr = s.cDPrf = s.cDPrr = s.cLP1 = s.cLP2 = s.cRP1 = s.cRP2 = 0;
    s.cLB1 = s.cLB2 = s.cRB1 = s.cRB2 = s.cLI1 = s.cLI2 = s.cRI1 = s.cRI2 = s.cRIf = s.cLIr = s.cBQ2 = 0;
    s.MQs0 = s.MQs1 = s.LPL0 = s.LPL1 = s.RPL0 = s.RPL1 = s.PF10 = s.PF11 = s.PF20 = s.PF21 = s.XM2 = s.BM2 = 0;
    s.LBL0 = s.LBL1 = s.RBL0 = s.RBL1 = s.LIT0 = s.LIT1 = s.RIT0 = s.RIT1 = 0;
    s.q1f0 = s.q1f1 = s.q1r0 = s.q1r1 = s.q2f0 = s.q2f1 = s.q2r0 = s.q2r1 = s.bqs0 = s.bqs1 = 0;
}
UVC_HD void k2m_begin(K2Merged & s, const BatchView & v, int64_t gp) {
    const TileInfo & T = v.tiles[v.pos_tile[gp]];
    s.T = &T; s.gp = gp;
    s.p = (int32_t)(gp - T.pos_off) + T.ext_beg;
    s.baq_p = v.baq[gp]; s.baq2_p = v.baq2[gp];
    s.noindel = v.noindel[gp];
    s.th = v.thres[gp];
    s.major = (int)v.refsym[gp];
    k2m_zero(s);
}
// adds the accumulators of both symbols to the position's records (atomic adds: rare events of other kernels' making and this thread's own
// earlier reductions may be in flight to the same words) and clears them
UVC_HD void k2m_flush(K2Merged & s, const BatchView & v) {
    for (int r = 0; r < 2; r++) {
        const int sym = (r ? UVC_LINK_M : s.major);
        if (sym >= UVC_NSYM) { continue; }
        const int sh = (r ? 16 : 0);
        SegAcc a;
        segacc_zero(a);
        #define UVC_H(x) ((int32_t)(((x) >> sh) & 0xffffu))
        a.s.aP1 = UVC_H(s.cP1); a.s.aP2 = UVC_H(s.cP2); a.s.aP3 = UVC_H(s.cP3); a.s.aNC = UVC_H(s.cNC);
        a.s.aDPff = UVC_H(s.cDPff); a.s.aDPfr = UVC_H(s.cDPfr); a.s.aDPrf = UVC_H(s.cDPrf); a.s.aDPrr = UVC_H(s.cDPrr);
        a.s.aLP1 = UVC_H(s.cLP1); a.s.aLP2 = UVC_H(s.cLP2); a.s.aRP1 = UVC_H(s.cRP1); a.s.aRP2 = UVC_H(s.cRP2);
        a.s.aLB1 = UVC_H(s.cLB1); a.s.aLB2 = UVC_H(s.cLB2); a.s.aRB1 = UVC_H(s.cRB1); a.s.aRB2 = UVC_H(s.cRB2);
        a.s.aLI1 = UVC_H(s.cLI1); a.s.aLI2 = UVC_H(s.cLI2); a.s.aRI1 = UVC_H(s.cRI1); a.s.aRI2 = UVC_H(s.cRI2);
        a.s.aRIf = UVC_H(s.cRIf); a.s.aLIr = UVC_H(s.cLIr); a.s.aBQ2 = UVC_H(s.cBQ2);
        #undef UVC_H
        a.s.aMQs = (r ? s.MQs1 : s.MQs0); a.s.aLPL = (r ? s.LPL1 : s.LPL0); a.s.aRPL = (r ? s.RPL1 : s.RPL0);
        a.s.aPF1 = (r ? s.PF11 : s.PF10); a.s.aPF2 = (r ? s.PF21 : s.PF20);
        a.s.a2XM2 = (r ? 0 : s.XM2); a.s.a2BM2 = (r ? 0 : s.BM2);
        a.s.aLBL = (r ? s.LBL1 : s.LBL0); a.s.aRBL = (r ? s.RBL1 : s.RBL0); a.s.aLIT = (r ? s.LIT1 : s.LIT0); a.s.aRIT = (r ? s.RIT1 : s.RIT0);
        a.a1BQf = (r ? s.q1f1 : s.q1f0); a.a1BQr = (r ? s.q1r1 : s.q1r0); a.a2BQf = (r ? s.q2f1 : s.q2f0); a.a2BQr = (r ? s.q2r1 : s.q2r0);
        a.bqsum = (r ? s.bqs1 : s.bqs0);
        segacc_flush<true>(v, s.gp, sym, a);
    }
    k2m_zero(s);
}
// one read of the position's window: `packed` = (symbol << 8) | quality of its aligned base at p, or UVC_K2_NOBASE
UVC_HD void k2m_read(K2Merged & s, const BatchView & v, const PileRec & P, uint32_t packed) {
    const uvcgpu_params & par = v.par;
    const int32_t p = s.p;
    if (P.rend <= p) { return; }
    bool not_first;
    int32_t prev_rpos = 0, next_rpos = INT32_MAX;
    if (P.bits & UVC_PR_SIMPLE) { not_first = (p > P.pos); }
    else {
        const CxEntry e = v.cx[(int64_t)P.cx_off + (p - P.pos)];
        if (!(e.flags & 1)) { return; }
        not_first = (e.flags & 2); prev_rpos = e.prev_rpos; next_rpos = e.next_rpos;
    }
    if ((P.bits & UVC_PR_MASK_ON) && !(P.ibeg <= p && p < P.iend)) { return; }     // primer_masked
    int32_t dist = 10000;
    if (P.bits & UVC_PR_HAS_GAPS) {       // dist_to_interfering_indel
        const TileInfo & T = *s.T;
        const int32_t adj = par.indel_adj_tracklen_dist;
        const int32_t npos = T.ext_end - T.ext_beg;
        const uvcgpu_rtr *rtr = v.rtr + T.pos_off;
        const int32_t ridx = p - T.ext_beg;
        const uvcgpu_rtr rtr1 = rtr[tmax(ridx, adj) - adj];
        const uvcgpu_rtr rtr2 = rtr[tmin(ridx + adj, npos - 1)];
        const int32_t prevlen = nnminus(p - prev_rpos, tmax(p - (T.ext_beg + rtr1.begpos), s.th.aLP1t));
        const int32_t nextlen = nnminus(next_rpos - p, tmax((T.ext_beg + rtr2.begpos + rtr2.tracklen) - p, s.th.aRP1t));
        dist = tmin(prevlen, nextlen);
    }
    const bool nfp = (par.tn_is_paired && (0x1 & par.primer_flag));
    const bool has_base = (packed != UVC_K2_NOBASE);
    const int sym = (int)(packed >> 8);
    const int32_t bqB = (int32_t)(packed & 0xffu) + par.bq_phred_added_misma;
    if (has_base && sym != s.major) {
        // a base that differs from the reference (rare): its own symbol's records are updated with fire-and-forget atomics
        const int32_t xm_term = (int32_t)(P.terms_lo & 127u);
        const int32_t bm_term = (int32_t)((sym < 3 ? (P.terms_lo >> (7 * (sym + 1))) : (P.terms_hi >> (7 * (sym - 3)))) & 127u);
        SegAcc one;
        segacc_zero(one);
        one.bqsum = bqB;
        segbias_dense<false>(one, v, P, s.th, s.baq_p, s.baq2_p, bqB, p, bm_term, xm_term, dist, nfp);
        segacc_flush<true>(v, s.gp, sym, one);
    }
    // activity of the two symbols at this read: aB = the base equals the position's major symbol, aL = the junction before the base is inside the match run
    const uint32_t aB = ((has_base && sym == s.major) ? 1u : 0u), aL = (not_first ? 1u : 0u);
    const uint32_t A = aB | (aL << 16);
    if (0 == A) { return; }
    if (++s.nvis >= UVC_K2M_FLUSH_EVERY) { k2m_flush(s, v); s.nvis = 1; }
    const uint32_t bits = P.bits;
    const uvcgpu_thres_set & th = s.th;
    const int32_t bqL = nnminus(tmin(80, s.noindel), (int32_t)((bits >> 24) & 0xfu)) + 1;      // nogap_weight
    // ---- what depends on (read, position) only
    const bool is_assay_amplicon = (bits & UVC_PR_AMPLICON);
    const int32_t nl = p - P.pos + 1, nr = P.rend - p;
    const int32_t seg_l_baq1 = s.baq_p - P.baq_pos + 1;
    const int32_t seg_r_baq0 = P.baq_rend1 - s.baq_p + 1;
    const int32_t seg_r_baqG = tmin(seg_r_baq0, P.baq2_rend1 - s.baq2_p + 7);
    const bool is_high_readlen = (par.central_readlen >= par.microadjust_median_readlen_thres);
    const int32_t seg_l_baq = (is_high_readlen ? seg_l_baq1 : tmax(seg_l_baq1, nl * par.microadjust_BAQ_per_base_x1024 / 1024));
    const int32_t seg_r_baqB = (is_high_readlen ? seg_r_baq0 : tmax(seg_r_baq0, nr * par.microadjust_BAQ_per_base_x1024 / 1024));
    const int32_t seg_r_baqL = (is_high_readlen ? seg_r_baqG : tmax(seg_r_baqG, nr * par.microadjust_BAQ_per_base_x1024 / 1024));
    const bool has_isize = (bits & UVC_PR_HAS_ISIZE);
    const int32_t frag_l_nb = (has_isize ? tmin(p - P.frag_l + 1, UVC_MAX_INSERT_SIZE) : UVC_MAX_INSERT_SIZE);
    const int32_t frag_r_nb = (has_isize ? tmin(P.frag_r - p, UVC_MAX_INSERT_SIZE) : UVC_MAX_INSERT_SIZE);
    const bool isrc = (bits & UVC_PR_ISRC);
    const bool st1 = (bits & UVC_PR_STRAND);
    const int32_t mapq = (int32_t)((bits >> 16) & 0xffu);
    const int32_t iB = (int32_t)aB, iL = (int32_t)aL;
    // ---- per symbol: counted / tier 2 / far from the edges / unaffected by the edges
    const bool tier2B = (bqB >= par.bias_thres_highBQ), countedB = tier2B;
    const bool countedL = (dist >= par.bias_thres_interfering_indel);
    const int32_t LPxTL = th.aLPxT, RPxT = th.aRPxT, LPxTB = tmin(LPxTL, RPxT);
    const bool farB = (nl >= LPxTB) & (nr >= RPxT), farL = (nl >= LPxTL) & (nr >= RPxT);
    const bool unaffB = (seg_l_baq >= par.bias_thres_highBAQ + 3) & (seg_r_baqB >= par.bias_thres_highBAQ + 3);
    const bool unaffL = (seg_l_baq >= par.bias_thres_highBAQ) & (seg_r_baqL >= par.bias_thres_highBAQ);
    const uint32_t gpB = ((countedB & farB) ? aB : 0u), gpL = ((countedL & farL) ? aL : 0u);
    const uint32_t gbB = ((countedB & unaffB) ? aB : 0u), gbL = ((countedL & unaffL) ? aL : 0u);
    const uint32_t GP = gpB | (gpL << 16), GPT2 = (tier2B ? gpB : 0u) | (gpL << 16);       // (the gap symbol is always tier 2)
    const uint32_t GB = gbB | (gbL << 16), GBT2 = (tier2B ? gbB : 0u) | (gbL << 16);
    // ---- sums of the qualities
    {
        const int32_t sqB = bqB * bqB / UVC_SQR_QUAL_DIV, sqL = bqL * bqL / UVC_SQR_QUAL_DIV;
        if (isrc) { s.q1r0 += iB * bqB; s.q2r0 += iB * sqB; s.q1r1 += iL * bqL; s.q2r1 += iL * sqL; }
        else { s.q1f0 += iB * bqB; s.q2f0 += iB * sqB; s.q1f1 += iL * bqL; s.q2f1 += iL * sqL; }
        s.bqs0 += iB * bqB; s.bqs1 += iL * bqL;
        s.MQs0 += iB * mapq; s.MQs1 += iL * mapq;
    }
    // ---- strand x orientation depth, distance to interfering indels, clips, insert ends
    if (st1) { if (isrc) { s.cDPrr += A; } else { s.cDPrf += A; } } else { if (isrc) { s.cDPfr += A; } else { s.cDPff += A; } }
    if (tmin(dist, tmin(nl, nr)) >= par.bias_thres_interfering_indel) { s.cP3 += A; }
    if (bits & UVC_PR_NOCLIP) { s.cNC += A; }
    if (has_isize) { if (isrc) { s.LIT0 += iB * frag_l_nb; s.LIT1 += iL * frag_l_nb; } else { s.RIT0 += iB * frag_r_nb; s.RIT1 += iL * frag_r_nb; } }
    {
        const int32_t min_dist2iend = ((bits & UVC_PR_PAIRED) ? tmin(frag_l_nb, frag_r_nb) : (isrc ? nr : nl));
        const bool x = ((min_dist2iend > par.primerlen2) | !is_assay_amplicon);
        s.cP1 += ((farB & unaffB & x) ? aB : 0u) | ((farL & unaffL & x) ? (aL << 16) : 0u);
        if (((bits & UVC_PR_UMI) != 0) | !is_assay_amplicon) { s.cP2 += A; }
    }
    // ---- passing-filter weights of the qualities
    {
        int32_t f1B, f2B, f1L, f2L;
        if ((uint32_t)bqB < 128u) { f1B = v.pf_tab[bqB]; f2B = v.pf_tab[128 + bqB]; }
        else {
            f1B = ((bqB < par.bias_thres_PFBQ1) ? (100 * (bqB * bqB) / (par.bias_thres_PFBQ1 * par.bias_thres_PFBQ1)) : 100);
            f2B = ((bqB < par.bias_thres_PFBQ2) ? (100 * (bqB * bqB) / (par.bias_thres_PFBQ2 * par.bias_thres_PFBQ2)) : 100);
        }
        if ((uint32_t)bqL < 128u) { f1L = v.pf_tab[bqL]; f2L = v.pf_tab[128 + bqL]; }
        else {
            f1L = ((bqL < par.bias_thres_PFBQ1) ? (100 * (bqL * bqL) / (par.bias_thres_PFBQ1 * par.bias_thres_PFBQ1)) : 100);
            f2L = ((bqL < par.bias_thres_PFBQ2) ? (100 * (bqL * bqL) / (par.bias_thres_PFBQ2 * par.bias_thres_PFBQ2)) : 100);
        }
        s.PF10 += iB * (100 * f1B / 100); s.PF20 += iB * (100 * f2B / 100);
        s.PF11 += iL * tmin(100, f1L); s.PF21 += iL * tmin(100, f2L);
        s.XM2 += iB * (int32_t)(P.terms_lo & 127u);
        // (sym == major here whenever iB = 1)
        s.BM2 += iB * (int32_t)((s.major < 3 ? (P.terms_lo >> (7 * (s.major + 1))) : (P.terms_hi >> (7 * ((s.major - 3) & 1)))) & 127u);
    }
    // ---- update_bidirectional_bias (main.hpp:1318-1358) twice: position bias and BAQ bias
    if (nl >= th.aLP1t) { s.cLP1 += GP; }
    if (nl >= th.aLP2t) { s.cLP2 += GPT2; }
    if (nr >= th.aRP1t) { s.cRP1 += GP; }
    if (nr >= th.aRP2t) { s.cRP2 += GPT2; }
    s.LPL0 += (int32_t)gpB * nl; s.LPL1 += (int32_t)gpL * nl; s.RPL0 += (int32_t)gpB * nr; s.RPL1 += (int32_t)gpL * nr;
    if (seg_l_baq >= par.bias_thres_BAQ1) { s.cLB1 += GB; }
    if (seg_l_baq >= par.bias_thres_BAQ2) { s.cLB2 += GBT2; }
    s.cRB1 += ((seg_r_baqB >= par.bias_thres_BAQ1) ? gbB : 0u) | ((seg_r_baqL >= par.bias_thres_BAQ1) ? (gbL << 16) : 0u);
    s.cRB2 += (((seg_r_baqB >= par.bias_thres_BAQ2) & tier2B) ? gbB : 0u) | ((seg_r_baqL >= par.bias_thres_BAQ2) ? (gbL << 16) : 0u);
    s.LBL0 += (int32_t)gbB * seg_l_baq; s.LBL1 += (int32_t)gbL * seg_l_baq; s.RBL0 += (int32_t)gbB * seg_r_baqB; s.RBL1 += (int32_t)gbL * seg_r_baqL;
    s.cBQ2 += (countedB ? aB : 0u) | (countedL ? (aL << 16) : 0u);
    // ---- insert-size bias: the reverse-complemented read looks at its left fragment end, the forward read at its right one
    {
        const bool mate_ok = (bits & UVC_PR_MATE_OK);
        const bool nonbiased = (mate_ok & (isrc ? (nl > nr) : (nl < nr)));
        const bool base_good = ((!is_assay_amplicon) | (!nfp));
        const bool goodB = (base_good | (farB & unaffB)), goodL = (base_good | (farL & unaffL));
        const int32_t d = (isrc ? frag_l_nb : frag_r_nb);
        const int32_t t1 = (isrc ? th.aLI1t : th.aRI1t), T1 = (isrc ? th.aLI1T : th.aRI1T);
        const int32_t t2 = (isrc ? th.aLI2t : th.aRI2t), T2 = (isrc ? th.aLI2T : th.aRI2T);
        const bool normal = ((bits & UVC_PR_IS_NORMAL) != 0);
        const bool okL = (normal | nonbiased);
        const uint32_t c1 = (((d >= t1) & (d <= T1) & normal) ? aB : 0u) | (((d >= t1) & okL) ? (aL << 16) : 0u);
        const uint32_t c2 = (((d >= t2) & (d <= T2) & normal & goodB) ? aB : 0u) | (((d >= t2) & okL & goodL) ? (aL << 16) : 0u);
        const uint32_t c3 = (goodB ? aB : 0u) | (goodL ? (aL << 16) : 0u);
        if (isrc) { s.cLI1 += c1; s.cLI2 += c2; s.cLIr += c3; } else { s.cRI1 += c1; s.cRI2 += c2; s.cRIf += c3; }
    }
}
UVC_HD void k2m_position(const BatchView & v, int64_t gp, const Win & w) {
    K2Merged s;
    k2m_begin(s, v, gp);
    for (int64_t ri = w.ulo; ri < w.uhi; ri++) {
        if (ri < w.lo || ri >= w.hi) { continue; }
        const PileRec & P = v.prec[ri];
        const int32_t qpos = base_index(v, P, s.p, true);
        const int32_t qc = tmax(qpos, 0);
        k2m_read(s, v, P, pack_base(v.seq[(uint64_t)P.seq_off + (uint32_t)(qc >> 1)], v.qual[(uint64_t)P.qual_off + (uint32_t)qc], qpos));
    }
    k2m_flush(s, v);
}

// ------------------------------------------------------------------------------------------------ K2e: one thread per indel event
// indel_len_rusize_phred (main.hpp:757-790): round(10*log10(i)) for i = 0..18
UVC_HD int32_t units_phred(int32_t indel_len, int32_t unit) {
    const int32_t tab[19] = {0, 0, 3, 5, 6, 7, 8, 8, 9, 10, 10, 10, 11, 11, 11, 12, 12, 12, 13};
    if (0 == (indel_len % unit)) { return tab[tmin(indel_len / unit, 18)]; }
    return tab[tmin(indel_len, 18)];
}

UVC_HD int32_t slip_phred_lookup(const BatchView & v, int variant, int32_t unit, int32_t nunits) {
    // indel_phred (main.hpp:794-801) from the host-evaluated table; beyond the table the closed form is evaluated here
    if (unit >= 1 && unit <= UVC_SLIP_MAXUNIT && nunits >= 0 && nunits < UVC_SLIP_NMAX) {
        return v.slip_tab[(variant * UVC_SLIP_MAXUNIT + (unit - 1)) * UVC_SLIP_NMAX + nunits];
    }
    const double ampfact = (variant ? v.par.indel_polymerase_slip_rate * v.par.indel_del_to_ins_err_ratio : v.par.indel_polymerase_slip_rate);
    const int32_t region = unit * nunits;
    const double num_slips = (region > 64 ? (double)(region - 8) : log1p(exp((double)region - 8.0))) * ampfact / ((double)(unit * unit));
    return (int32_t)floor(-10 * log((1.0 - 2.220446049250313e-16) / (num_slips + 1.0)) / log(10.0));
}

// ref_to_phredvalue (main.hpp:876-922): repeat context right of the junction decides the prior quality of an indel there
UVC_HD int32_t indel_prior_phred(int32_t & n_units, int32_t & max_num, int32_t & best_unit, const BatchView & v, const TileInfo & T,
        int32_t rpos, int32_t oplen, bool is_del) {
    const uint8_t *ref = v.refsym + T.pos_off;
    const int32_t n = (T.ext_end - T.ext_beg) - 1;
    const int32_t refpos = rpos - T.ext_beg;
    const int32_t str_max = v.par.indel_str_repeatsize_max;
    max_num = 0; best_unit = 0;
    for (int32_t unit = 1; unit <= str_max; unit++) {
        int32_t q = refpos;
        while (q + unit < n && ref[q] == ref[q + unit]) { q++; }
        const int32_t num = (q - refpos) / unit + 1;
        // is_indel_context_more_STR (main.hpp:699-721) with its rank2 quirk
        bool better;
        if (best_unit * max_num == 0) { better = true; }
        else if (unit > str_max || best_unit > str_max) { better = (unit < best_unit || (unit == best_unit && num > max_num)); }
        else {
            int rank1 = (num <= 1 ? (-num * unit) : ((num - 1) * unit));
            int rank2 = (max_num <= 1 ? (-max_num * unit) : ((max_num - 1) * best_unit));
            if (0 == num || 0 == unit) { rank1 = -100; }
            if (0 == max_num || 0 == best_unit) { rank2 = -100; }
            better = (rank1 > rank2);
        }
        if (better) { max_num = num; best_unit = unit; }
    }
    const int variant = ((oplen == best_unit && is_del) ? 1 : 0);
    const int32_t decphred = slip_phred_lookup(v, variant, best_unit, max_num);
    if (best_unit * (max_num - 1) >= 6 - 1) {
        n_units = ((0 == oplen % best_unit) ? (oplen / best_unit) : ((1 == oplen) ? 1 : 0));
    } else {
        n_units = 1 + (oplen / 6);
    }
    const int32_t max_phred = v.par.indel_BQ_max;
    return max_phred - tmin(max_phred, decphred) + units_phred(oplen, best_unit);
}

// Evaluates one insertion/deletion of one read (main.hpp:2009-2257) once: its symbol and quality weights are stored in the event for the
// fragment/family kernels, and the bias walk's contributions (including the BASE_NN/LINK_NN paddings of a deletion) are added here.
UVC_HD void k2e_event(const BatchView & v, int64_t ei) {
    IndelEvent & E = v.ev[ei];
    const ReadRec & R = v.reads[E.read];
    const ReadDerived & D = v.rd[E.read];
    const TileInfo & T = v.tiles[R.tile];
    const uvcgpu_params & par = v.par;
    const int64_t po = T.pos_off - T.ext_beg;
    const int32_t *baq = v.baq + po;
    const int32_t *baq2 = v.baq2 + po;
    const uint8_t *qual = v.qual + R.qual_off;
    const int32_t rpos = E.rpos, qpos = E.qpos, oplen = E.oplen, rend = R.rend, pos = R.pos;
    const bool isrc = ((R.flag & 0x10) == 0x10);
    E.counted = 0; E.symbol = -1; E.incvalue = 0; E.incvalue2 = 0;
    if (primer_masked(v, R, D, rpos)) { return; }
    const uvcgpu_prep_set & pp = v.prep[po + rpos];
    const int32_t added = par.bq_phred_added_indel;
    const int32_t ratiothres = (par.is_tumor_vcf_provided ? 4 : 2);
    int32_t incvalue = 1;
    int32_t n_units = oplen;
    int32_t nbases2end;
    if (!E.is_del) {
        nbases2end = tmin(qpos, R.l_qseq - (qpos + oplen));
        if (nbases2end <= 0) {
            incvalue = (0 != qpos ? (int32_t)qual[qpos - 1] : ((qpos + oplen < R.l_qseq) ? (int32_t)qual[qpos + oplen] : 1)) + added;
        } else {
            int32_t max_num, unit;
            int32_t phredvalue = indel_prior_phred(n_units, max_num, unit, v, T, rpos, oplen, false);
            const int32_t phredinc = (int32_t)round(2 * (v.ten_over_ln10 * log((double)pp.a_dp / (double)(1.0 + nnminus(pp.a_dp, pp.a_at_ins_dp + pp.a_at_del_dp)))));
            const bool multiallelic = (pp.a_near_ins_pow2len * ratiothres > (int64_t)tmax(1, pp.a_near_ins_dp) * (int64_t)((uint32_t)oplen * 3u));
            if (1 == n_units && !multiallelic) { phredvalue += between(phredinc - 3, 0, 4); }
            const int32_t thisdp = pp.a_at_ins_dp;
            const int32_t neardp = tmax(pp.a_near_ins_dp, pp.a_near_RTR_ins_dp);
            int32_t ins_min = 80;
            for (int32_t q2 = qpos; q2 < qpos + oplen; q2++) { ins_min = tmin(ins_min, (int32_t)qual[q2]); }
            int32_t anc_min = 80;
            if (qpos > 0) { anc_min = tmin(anc_min, (int32_t)qual[qpos - 1]); }
            if (qpos + oplen + 1 < R.l_qseq) { anc_min = tmin(anc_min, (int32_t)qual[qpos + oplen + 1]); } // QUIRK: + 1 (main.hpp:2055-2056)
            const int32_t q1 = tmin(anc_min, ins_min);
            const bool lenient = (thisdp * ratiothres <= neardp || (1 == oplen && (D.xm1500 >= par.microadjust_xm
                    || ((D.lclip + par.microadjust_cliplen >= rpos - pos) && isrc)
                    || ((D.rclip + par.microadjust_cliplen >= rend - pos) && !isrc))));
            const int32_t q2v = (lenient ? q1 : 80);
            incvalue = nnminus(tmin(q2v, phredvalue + added), D.micro_indel_penal) + 1;
        }
        if (nbases2end >= par.indel_filter_edge_dist) {
            E.symbol = (1 == n_units ? UVC_LINK_I1 : ((2 == n_units) ? UVC_LINK_I2 : UVC_LINK_I3P));
            E.incvalue = tmax(1, incvalue);
            int32_t inc2 = incvalue;
            for (int32_t q2 = qpos; q2 < qpos + oplen; q2++) { inc2 = tmin(inc2, (int32_t)qual[q2] + added); }
            E.incvalue2 = tmax(1, inc2);
            E.counted = 1;
            SegAcc a; segacc_zero(a);
            a.bqsum = E.incvalue;
            segbias<true>(a, v, R, D, v.thres[po + rpos], baq[rpos], baq2[rpos], E.incvalue, rpos, 100, true, oplen, 10000);
            segacc_flush<true>(v, po + rpos, E.symbol, a);
        }
    } else {
        nbases2end = tmin(qpos, R.l_qseq - qpos);
        if (nbases2end <= 0) {
            incvalue = (0 != qpos ? (int32_t)qual[qpos - 1] : ((qpos < R.l_qseq) ? (int32_t)qual[qpos] : 1)) + added;
        } else {
            int32_t max_num, unit;
            int32_t phredvalue = indel_prior_phred(n_units, max_num, unit, v, T, rpos, oplen, true);
            const int32_t phredinc = (int32_t)round(2 * (v.ten_over_ln10 * log((double)pp.a_dp / (double)(1.0 + nnminus(pp.a_dp, pp.a_at_ins_dp + pp.a_at_del_dp)))));
            if (1 == n_units) { phredvalue += between(phredinc - 3, 0, 4); }
            const int32_t thisdp = pp.a_at_del_dp;
            const int32_t neardp = tmax(pp.a_near_del_dp, pp.a_near_RTR_del_dp);
            const int32_t q1 = tmin(tmin((int32_t)qual[qpos], (int32_t)qual[qpos - 1]), 80);
            const int32_t q2v = ((thisdp * ratiothres <= neardp) ? nnminus(q1, 1) : 80);
            const double delFA = ((double)(thisdp + 0.5) / (double)(pp.a_dp + 1));
            const int32_t delFAQ = tmax(0, par.microadjust_delFAQmax + (int32_t)round(par.powlaw_exponent * (v.ten_over_ln10 * log(delFA))));
            const uint32_t *cigar = v.cigar + R.cigar_off;
            // flanking BAQ up to an insertion of the same length, or else to the read ends (main.hpp:2167-2186)
            int32_t pc = E.cigar_idx, prev_rpos = rpos;
            while ((0 != pc) && (UVC_CINS != cig_op(cigar[pc]) || oplen != cig_len(cigar[pc]))) {
                pc--;
                const int op2 = cig_op(cigar[pc]);
                if (is_match_op(op2) || op2 == UVC_CDEL || op2 == UVC_CREF_SKIP) { prev_rpos -= cig_len(cigar[pc]); }
            }
            int32_t nc = E.cigar_idx, next_rpos = rpos + oplen;
            while ((R.n_cigar - 1 != nc) && (UVC_CINS != cig_op(cigar[nc]) || oplen != cig_len(cigar[nc]))) {
                nc++;
                const int op2 = cig_op(cigar[nc]);
                if (is_match_op(op2) || op2 == UVC_CDEL || op2 == UVC_CREF_SKIP) { next_rpos += cig_len(cigar[nc]); }
            }
            const int32_t baq_l = baq[rpos] - baq[prev_rpos];
            const int32_t baq_r = baq[next_rpos] - baq[rpos + oplen];
            const int32_t qbaq = tmax(delFAQ, tmax(q1, tmin(baq_l, baq_r)));
            incvalue = nnminus(tmin(q2v, tmin(qbaq, phredvalue + added)), D.micro_indel_penal) + 1;
        }
        if (nbases2end >= par.indel_filter_edge_dist) {
            E.symbol = (1 == n_units ? UVC_LINK_D1 : ((2 == n_units) ? UVC_LINK_D2 : UVC_LINK_D3P));
            E.incvalue = tmax(1, incvalue);
            E.incvalue2 = E.incvalue;
            E.counted = 1;
            {
                SegAcc a; segacc_zero(a);
                a.bqsum = E.incvalue;
                segbias<true>(a, v, R, D, v.thres[po + rpos], baq[rpos], baq2[rpos], E.incvalue, rpos, 100, false, oplen, 10000);
                segacc_flush<true>(v, po + rpos, E.symbol, a);
            }
            // padded-deletion symbols: BASE_NN on every deleted base, LINK_NN on the junction after it (main.hpp:2219-2253)
            for (int32_t r2 = rpos; r2 < tmin(rpos + oplen, rend); r2++) {
                const CxEntry c = v.cx[R.cx_off + (r2 - pos)];
                for (int s = 0; s < 2; s++) {
                    const int32_t p2 = (s == 0 ? r2 : r2 + 1);
                    if (p2 >= rend) { continue; }
                    SegAcc a; segacc_zero(a);
                    a.bqsum = E.incvalue;
                    segbias<true>(a, v, R, D, v.thres[po + p2], baq[p2], baq2[p2], E.incvalue, p2, 100, false, oplen, (s == 0 ? c.prev_rpos : c.next_rpos));
                    segacc_flush<true>(v, po + p2, (s == 0 ? UVC_BASE_NN : UVC_LINK_NN), a);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ fragment-level helpers
UVC_HD void votes_zero(int32_t c[UVC_NSYM]) { for (int s = 0; s < UVC_NSYM; s++) { c[s] = 0; } }

// What one read asserts at position p, max-merged into c: updateByAln<BASE_QUALITY_MAX, false, *> (main.hpp:1886-2257, walks #3-#5)
UVC_HD void read_votes(const BatchView & v, int64_t ri, int32_t p, int64_t gp, int32_t c[UVC_NSYM]) {
    const ReadRec & R = v.reads[ri];
    if (p < R.pos || p >= R.rend) { return; }
    const ReadDerived & D = v.rd[ri];
    const Locus L = locate(v, R, p);
    if (L.is_m && !primer_masked(v, R, D, p)) {
        if (L.not_first) { c[UVC_LINK_M] = tmax(c[UVC_LINK_M], nogap_weight(v, gp, D)); }
        const int sym = base3(v.seq + R.seq_off, L.qpos);
        c[sym] = tmax(c[sym], (int32_t)v.qual[R.qual_off + L.qpos] + v.par.bq_phred_added_misma);
    }
    if (!R.simple) {
        for (int32_t e = 0; e < R.n_ev; e++) {
            const IndelEvent & E = v.ev[R.ev_off + e];
            if (!E.counted) { continue; }
            if (E.rpos == p) { c[E.symbol] = tmax(c[E.symbol], E.incvalue); }
            if (E.is_del) {
                const int32_t del_end = tmin(E.rpos + E.oplen, R.rend);
                if (p >= E.rpos && p < del_end) { c[UVC_BASE_NN] = tmax(c[UVC_BASE_NN], E.incvalue); }
                if (p - 1 >= E.rpos && p - 1 < del_end) { c[UVC_LINK_NN] = tmax(c[UVC_LINK_NN], E.incvalue); }
            }
        }
    }
}

UVC_HD bool frag_covers(const BatchView & v, const FragRec & G, int32_t p) {
    for (int32_t k = 0; k < G.n_reads; k++) { const ReadRec & R = v.reads[v.frag_reads[G.read_off + k]]; if (R.pos <= p && p < R.rend) { return true; } }
    return false;
}

UVC_HD void frag_votes(const BatchView & v, const FragRec & G, int32_t p, int64_t gp, int32_t c[UVC_NSYM]) {
    votes_zero(c);
    for (int32_t k = 0; k < G.n_reads; k++) { read_votes(v, v.frag_reads[G.read_off + k], p, gp, c); }
}

// _fillConsensusCounts (main.hpp:374-402). ref_once = TIsRefCountedOnlyOnce (link symbols only).
UVC_HD void consensus(const int32_t c[UVC_NSYM], int lo, int hi, bool ref_once, int & argmax, int32_t & cmax, int32_t & csum) {
    argmax = hi; cmax = 0; csum = 0;
    for (int s = lo; s <= hi; s++) {
        if (ref_once) {
            if (cmax < c[s] || (UVC_LINK_M == argmax && (0 < c[s]))) { argmax = s; cmax = c[s]; csum = cmax; }
        } else {
            if (cmax < c[s]) { argmax = s; cmax = c[s]; }
            csum += c[s];
        }
    }
}
UVC_HD void link_consensus(const int32_t c[UVC_NSYM], bool ref_once, int & argmax, int32_t & cmax, int32_t & csum) { consensus(c, UVC_LINK_M, UVC_LINK_NN, ref_once, argmax, cmax, csum); }
UVC_HD void base_consensus(const int32_t c[UVC_NSYM], bool ignore_padded_del, int & argmax, int32_t & cmax, int32_t & csum) {
    consensus(c, UVC_BASE_A, (ignore_padded_del ? UVC_BASE_T : UVC_BASE_NN), false, argmax, cmax, csum);
}

UVC_HD bool symbols_mutated(int ref, int alt) { // areSymbolsMutated (main_conversion.hpp:364-371)
    if (alt <= UVC_BASE_NN) { return ref != alt && ref < UVC_BASE_N && alt < UVC_BASE_N; }
    return alt != UVC_LINK_M && alt != UVC_LINK_NN;
}

UVC_HD int32_t avg_bq(const BatchView & v, int64_t gp, int s) { // get_avgBQ (main_conversion.hpp:791-796)
    const uvcgpu_seginfo_set & g = v.seginfo[gp * UVC_NSYM + s];
    return v.bqsum[gp * UVC_NSYM + s] / tmax(1, g.aDPff + g.aDPfr + g.aDPrf + g.aDPrr);
}

// PhredMutationTable::toPhredErrRate (main.hpp:238-261)
UVC_HD int32_t sscs_phred(const uvcgpu_params & par, int con, int alt) {
    int32_t r;
    if (is_ins_symbol(con) || is_del_symbol(con)) { r = par.fam_phred_sscs_indel_open; }
    else if (con == UVC_LINK_M) {
        if (alt == UVC_LINK_D1 || alt == UVC_LINK_I1) { r = par.fam_phred_sscs_indel_open; }
        else if (alt == UVC_LINK_D2 || alt == UVC_LINK_I2) { r = par.fam_phred_sscs_indel_open + par.fam_phred_sscs_indel_ext; }
        else { r = par.fam_phred_sscs_indel_open + par.fam_phred_sscs_indel_ext * 2; }
    }
    else if ((con == UVC_BASE_C && alt == UVC_BASE_T) || (con == UVC_BASE_G && alt == UVC_BASE_A)) { r = par.fam_phred_sscs_transition_CG_TA; }
    else if ((con == UVC_BASE_A && alt == UVC_BASE_G) || (con == UVC_BASE_T && alt == UVC_BASE_C)) { r = par.fam_phred_sscs_transition_AT_GC; }
    else if ((con == UVC_BASE_C && alt == UVC_BASE_A) || (con == UVC_BASE_G && alt == UVC_BASE_T)) { r = par.fam_phred_sscs_transversion_CG_AT; }
    else { r = par.fam_phred_sscs_transversion_other; }
    return r + (par.tumor_vcf_fname_nonempty ? 3 : 0); // all_mutation_inc (main.hpp:236)
}

// appends n words to the record stream; returns the write offset or -1 if the stream is full (the cursor keeps growing so the host sees the overflow)
UVC_HD int32_t rec_alloc(const BatchView & v, int32_t n) {
#if defined(__CUDA_ARCH__)
    const int32_t off = atomicAdd(v.rec_cursor, n);
#else
    const int32_t off = *v.rec_cursor; *v.rec_cursor += n;
#endif
    return (off + n <= v.rec_cap ? off : -1);
}
UVC_HD void rec_put6(const BatchView & v, int32_t kind, int32_t strand, int32_t symbol, int32_t pos, int32_t evidx, int32_t cnt) {
    const int32_t off = rec_alloc(v, 6);
    if (off < 0) { return; }
    int32_t *w = v.rec_buf + off;
    w[0] = kind; w[1] = strand; w[2] = symbol; w[3] = pos; w[4] = evidx; w[5] = cnt;
}

// ASCII letter of a 4-bit base code ("=ACMGRSVTWYHKDBN"), used to order inserted sequences like std::string does
UVC_HD int nt16_ascii(int b4) { const char *t = "=ACMGRSVTWYHKDBN"; return (int)t[b4 & 0xf]; }

// -1 / 0 / +1 comparison of the identities of two indel events of the same symbol class: deletion length, or inserted sequence as a string
UVC_HD int indel_cmp(const BatchView & v, const IndelEvent & A, const IndelEvent & B) {
    if (A.is_del) { return (A.oplen < B.oplen ? -1 : (A.oplen > B.oplen ? 1 : 0)); }
    const uint8_t *sa = v.seq + v.reads[A.read].seq_off, *sb = v.seq + v.reads[B.read].seq_off;
    const int32_t n = tmin(A.oplen, B.oplen);
    for (int32_t i = 0; i < n; i++) {
        const int ca = nt16_ascii(base4(sa, A.qpos + i)), cb = nt16_ascii(base4(sb, B.qpos + i));
        if (ca != cb) { return (ca < cb ? -1 : 1); }
    }
    return (A.oplen < B.oplen ? -1 : (A.oplen > B.oplen ? 1 : 0));
}

// posToIndelToCount_updateByConsensus + indelToData_getMajority (main.hpp:50-63, 83-95) over the indel events of class `symbol` that the
// reads [reads, reads + n) show at position p: the identity with the largest summed weight wins, ties go to the larger key.
// Returns the global index of a representative event, or -1.
UVC_HD int32_t indel_majority_of_reads(const BatchView & v, const int32_t *reads, int32_t n, int32_t p, int symbol) {
    int32_t best = -1; int64_t best_w = 0;
    for (int32_t k = 0; k < n; k++) {
        const ReadRec & R = v.reads[reads[k]];
        if (R.simple || p < R.pos || p >= R.rend) { continue; }
        for (int32_t e = 0; e < R.n_ev; e++) {
            const IndelEvent & E = v.ev[R.ev_off + e];
            if (!E.counted || E.rpos != p || E.symbol != symbol) { continue; }
            // summed weight of this identity over all reads
            int64_t w = 0;
            bool is_first = true;
            for (int32_t k2 = 0; k2 < n; k2++) {
                const ReadRec & R2 = v.reads[reads[k2]];
                if (R2.simple || p < R2.pos || p >= R2.rend) { continue; }
                for (int32_t e2 = 0; e2 < R2.n_ev; e2++) {
                    const IndelEvent & E2 = v.ev[R2.ev_off + e2];
                    if (!E2.counted || E2.rpos != p || E2.symbol != symbol) { continue; }
                    if (0 == indel_cmp(v, E, E2)) {
                        w += (E2.is_del ? E2.incvalue : E2.incvalue2);
                        if (k2 < k || (k2 == k && e2 < e)) { is_first = false; }
                    }
                }
            }
            if (!is_first) { continue; }
            if (best < 0 || w > best_w || (w == best_w && indel_cmp(v, E, v.ev[best]) > 0)) { best = R.ev_off + e; best_w = w; }
        }
    }
    return best;
}

// ------------------------------------------------------------------------------------------------ KF: one thread per fragment-column entry
// The reference rebuilds the per-fragment symbol array three times (walks #3, #4, #5). Here it is evaluated once per (fragment, position):
// a warp owns 32 consecutive positions of ONE fragment, so the fragment's read records are warp-uniform and its base qualities are read
// as consecutive bytes; the entry keeps only what fillConsensusCounts (main.hpp:374-417) yields for the two symbol types.
// Returns the entry's UVC_FM_* bits; the caller folds the 32 results of a chunk into v.fmask (one ballot per mask in the CUDA kernel).
// What KF needs from one read of a plain fragment (a read whose CIGAR is [S|H] M [S|H]), loaded once per fragment
struct KfRead {
    int32_t pos, rend, m_qoff, ibeg, iend, nogap_penal;
    bool mask_on;                  // primers of this read are masked outside [ibeg, iend) (primer_masked)
    const uint8_t *seq, *qual;
};
UVC_HD void kf_load_read(KfRead & k, const BatchView & v, int64_t ri) {
    const ReadRec & R = v.reads[ri];
    const ReadDerived & D = v.rd[ri];
    k.pos = R.pos; k.rend = R.rend; k.m_qoff = R.m_qoff; k.ibeg = D.ibeg; k.iend = D.iend; k.nogap_penal = D.micro_nogap_penal;
    const bool is_assay_amplicon = ((R.dflag & 0x4) || ((v.par.primerlen > 0) && !(0x2 & v.par.primer_flag)));
    const bool normal_filters_primers = (v.par.tn_is_paired && (0x1 & v.par.primer_flag));
    k.mask_on = !(normal_filters_primers || !is_assay_amplicon);
    k.seq = v.seq + R.seq_off; k.qual = v.qual + R.qual_off;
}
// true if the fragment is plain: one or two reads, none with indels or skips
UVC_HD bool kf_fragment_is_plain(const BatchView & v, const FragRec & G) {
    bool plain = (G.n_reads <= 2);
    for (int32_t k = 0; plain && k < G.n_reads; k++) { plain = (0 != v.reads[v.frag_reads[G.read_off + k]].simple); }
    return plain;
}
// stores the entry of a covered position from the two consensus triples and returns its UVC_FM_* bits
UVC_HD uint32_t kf_store_entry(const BatchView & v, int64_t i, int64_t gp, bool covered, int la, int32_t lcc, int ba, int32_t bcc, int32_t btc, bool n_votes) {
    FragCol e;
    e.link_cc = 0; e.base_cc = 0; e.base_tc = 0; e.link_sym = UVC_LINK_NN; e.base_sym = UVC_BASE_NN;
    uint32_t bits = 0;
    if (covered) {
        e.link_sym = (uint8_t)(la | (n_votes ? 0x80 : 0)); e.link_cc = (uint16_t)lcc;
        const int ref = v.refsym[gp];
        if (e.link_cc > 0 && symbols_mutated(ref, la)) { bits |= (1u << UVC_FM_MUT_LINK); }
        e.base_sym = (uint8_t)ba; e.base_cc = (uint16_t)bcc; e.base_tc = (uint16_t)btc;
        // (the stored counts are 16-bit: the tests below use them as every later kernel reads them)
        const int32_t con_qual = (int32_t)e.base_cc * 2 - (int32_t)e.base_tc;
        if (e.base_tc > 0 && symbols_mutated(ref, ba)) {
            if (con_qual >= v.par.bias_thres_highBQ) { bits |= (1u << UVC_FM_MUT_BASE_HQ); }
            if (con_qual > 0) { bits |= (1u << UVC_FM_MUT_BASE_ANY); }
        }
        if (e.link_cc | e.base_tc) { bits |= (1u << UVC_FM_COV); }
    }
    v.fcol[i] = e;
    return bits;
}
// Entry i (position p, concatenated index gp) of a plain fragment with reads rd[0 .. n). Such a fragment votes for LINK_M and for at most two base
// symbols: the consensus of both symbol types is written out directly (same results as the general path, which keeps a 14-entry vote array
// in local memory).
UVC_HD uint32_t kf_plain_entry(const BatchView & v, int64_t i, int32_t p, int64_t gp, const KfRead *rd, int32_t n) {
    bool covered = false;
    int32_t linkw = 0, q0 = 0, q1 = 0;
    int s0 = -1, s1 = -1;
    #pragma unroll
    for (int32_t k = 0; k < 2; k++) {      // (n <= 2; constant indices keep rd in registers)
        if (k >= n) { continue; }
        const KfRead & R = rd[k];
        if (p < R.pos || p >= R.rend) { continue; }
        covered = true;
        if (R.mask_on && !(R.ibeg <= p && p < R.iend)) { continue; }      // primer_masked
        if (p > R.pos) {                                                    // nogap_weight
            const int32_t noindel = v.noindel[gp];
            linkw = tmax(linkw, nnminus(tmin(80, noindel), R.nogap_penal) + 1);
        }
        const int32_t qpos = R.m_qoff + (p - R.pos);
        const int sym = base3(R.seq, qpos);
        const int32_t q = tmax(0, (int32_t)R.qual[qpos] + v.par.bq_phred_added_misma);   // votes are max-merged into zeros
        if (s0 < 0 || s0 == sym) { s0 = sym; q0 = tmax(q0, q); } else { s1 = sym; q1 = tmax(q1, q); }
    }
    int la = UVC_LINK_NN, ba = UVC_BASE_NN;
    int32_t lcc = 0, bcc = 0;
    if (linkw > 0) { la = UVC_LINK_M; lcc = linkw; }
    if (s1 >= 0 && s1 < s0) { const int ts = s0; s0 = s1; s1 = ts; const int32_t tq = q0; q0 = q1; q1 = tq; }   // symbol order decides ties
    if (s0 >= 0 && q0 > 0) { ba = s0; bcc = q0; }
    if (s1 >= 0 && q1 > bcc) { ba = s1; bcc = q1; }
    const bool n_votes = ((s0 == UVC_BASE_N && q0 > 0) || (s1 == UVC_BASE_N && q1 > 0));      // votes for BASE_N exist (bit 7 of link_sym)
    return kf_store_entry(v, i, gp, covered, la, lcc, ba, bcc, q0 + q1, n_votes);
}
// Returns the entry's UVC_FM_* bits; the caller folds the 32 results of a chunk into v.fmask (one ballot per mask in the CUDA kernel).
UVC_HD uint32_t kf_fragment_column(const BatchView & v, int64_t i) {
    const FragRec & G = v.frags[v.fchunk_frag[i / UVC_COL_CHUNK]];
    const int32_t o = (int32_t)(i - G.col_off);
    if (o >= G.hi - G.lo) { return 0; }
    const int32_t p = G.lo + o;
    const TileInfo & T = v.tiles[G.tile];
    const int64_t gp = T.pos_off + (p - T.ext_beg);
    if (kf_fragment_is_plain(v, G)) {
        KfRead rd[2];
        for (int32_t k = 0; k < G.n_reads; k++) { kf_load_read(rd[k], v, v.frag_reads[G.read_off + k]); }
        return kf_plain_entry(v, i, p, gp, rd, G.n_reads);
    }
    bool covered = false;
    int la = UVC_LINK_NN, ba = UVC_BASE_NN;
    int32_t lcc = 0, bcc = 0, btc = 0;
    bool n_votes = false;                 // votes for BASE_N / BASE_NN exist (bit 7 of link_sym)
    if (frag_covers(v, G, p)) {
        covered = true;
        int32_t c[UVC_NSYM];
        frag_votes(v, G, p, gp, c);
        int32_t tc;
        link_consensus(c, true, la, lcc, tc);
        n_votes = (0 != (c[UVC_BASE_N] | c[UVC_BASE_NN]));
        base_consensus(c, false, ba, bcc, btc);
    }
    return kf_store_entry(v, i, gp, covered, la, lcc, ba, bcc, btc, n_votes);
}
// folds one entry's bits into the chunk masks (emulation / non-ballot path)
UVC_HD void kf_fold_bits(const BatchView & v, int64_t i, uint32_t bits) {
    uint32_t *m = v.fmask + (i / UVC_COL_CHUNK) * 4;
    for (int k = 0; k < 4; k++) { if ((bits >> k) & 1u) { m[k] |= (1u << (uint32_t)(i % UVC_COL_CHUNK)); } }
}

// the fragment's entry at p (zero entry outside its covered extent)
UVC_HD FragCol frag_entry(const BatchView & v, const FragRec & G, int32_t p) {
    if (p < G.lo || p >= G.hi) { FragCol z; z.link_cc = 0; z.base_cc = 0; z.base_tc = 0; z.link_sym = UVC_LINK_NN; z.base_sym = UVC_BASE_NN; return z; }
    return v.fcol[G.col_off + (p - G.lo)];
}

// ------------------------------------------------------------------------------------------------ K3a: one thread per fragment
// Whole-fragment statistics of updateByAlns3UsingBQ (main.hpp:2650-2756): number of covered positions, number of covered positions within
// +-syserr_mut_region_n_bases of a high-quality mutation, and the fragment's string of mutations (haplotype evidence).
UVC_HD void k3a_fragment(const BatchView & v, int64_t fi) {
    FragRec & G = v.frags[fi];
    const uvcgpu_params & par = v.par;
    const int32_t lo = G.lo, hi = G.hi;
    const int32_t nb = par.syserr_mut_region_n_bases;
    const int64_t c0 = G.col_off / UVC_COL_CHUNK;
    const int32_t n_chunks = (hi - lo + UVC_COL_CHUNK - 1) / UVC_COL_CHUNK;
    const uint32_t *fm = v.fmask + c0 * 4;
    int32_t n_cov = 0, n_near = 0, n_mut_entries = 0;
    if (nb >= 0 && nb <= 32) {
        // n_near = covered positions within nb of a mutated position (each counted once, main.hpp:2747-2756): the covered bits of the mutation
        // mask dilated by nb to both sides, evaluated chunk by chunk on a 96-bit window (previous | current | next chunk)
        uint32_t prev = 0, cur = (n_chunks > 0 ? (fm[UVC_FM_MUT_LINK] | fm[UVC_FM_MUT_BASE_HQ]) : 0u);
        for (int32_t c = 0; c < n_chunks; c++) {
            const uint32_t *m = fm + c * 4;
            const uint32_t nxt = (c + 1 < n_chunks ? (m[4 + UVC_FM_MUT_LINK] | m[4 + UVC_FM_MUT_BASE_HQ]) : 0u);
            n_cov += __popc_u32(m[UVC_FM_COV]);
            n_mut_entries += __popc_u32(m[UVC_FM_MUT_LINK]) + __popc_u32(m[UVC_FM_MUT_BASE_HQ]);
            if (prev | cur | nxt) {
                const uint64_t lowmid = (uint64_t)prev | ((uint64_t)cur << 32);      // bit 32 + k = position k of the current chunk
                const uint64_t midhigh = (uint64_t)cur | ((uint64_t)nxt << 32);      // bit k = position k of the current chunk
                uint64_t right = 0, left = 0;                                        // mutations reaching up (p .. p + nb) / down (p - nb .. p - 1)
                for (int32_t k = 0; k <= nb; k++) { right |= (lowmid << k); }
                for (int32_t k = 1; k <= nb; k++) { left |= (midhigh >> k); }
                const uint32_t near = (uint32_t)(right >> 32) | (uint32_t)left;
                n_near += __popc_u32(near & m[UVC_FM_COV]);
            }
            prev = cur; cur = nxt;
        }
    } else {
        // (a neighbourhood wider than one chunk: distance to the previous and to the next mutated position, bit by bit)
        const int32_t n = hi - lo;
        const int32_t far = 0x3fffffff;
        int32_t prev = -far, nxt = -1;
        for (int32_t i = 0; i < n; i++) {
            const uint32_t *m = fm + (i / UVC_COL_CHUNK) * 4;
            const uint32_t bit = 1u << (uint32_t)(i % UVC_COL_CHUNK);
            n_mut_entries += ((m[UVC_FM_MUT_LINK] & bit) ? 1 : 0) + ((m[UVC_FM_MUT_BASE_HQ] & bit) ? 1 : 0);
            if ((m[UVC_FM_MUT_LINK] | m[UVC_FM_MUT_BASE_HQ]) & bit) { prev = i; }
            if (nxt < i) {      // next mutated position at or after i
                nxt = far;
                for (int32_t j = i; j < n; j++) {
                    const uint32_t *mj = fm + (j / UVC_COL_CHUNK) * 4;
                    if ((mj[UVC_FM_MUT_LINK] | mj[UVC_FM_MUT_BASE_HQ]) & (1u << (uint32_t)(j % UVC_COL_CHUNK))) { nxt = j; break; }
                }
            }
            if (m[UVC_FM_COV] & bit) {
                n_cov++;
                if (nb >= 0 && (i - prev <= nb || nxt - i <= nb)) { n_near++; }
            }
        }
    }
    // the fragment's haplotype string: its mutated high-quality consensus symbols in position order, link before base (main.hpp:2766-2800)
    if (n_mut_entries > 1) {
        int32_t out = rec_alloc(v, 4 + 2 * n_mut_entries);
        if (out >= 0) {
            v.rec_buf[out] = UVC_REC_HAP_BQ; v.rec_buf[out + 1] = G.strand; v.rec_buf[out + 2] = n_mut_entries; v.rec_buf[out + 3] = G.tile;
            out += 4;
            for (int32_t c = 0; c < n_chunks; c++) {
                const uint32_t *m = fm + c * 4;
                for (uint32_t todo = (m[UVC_FM_MUT_LINK] | m[UVC_FM_MUT_BASE_HQ]); todo; todo &= todo - 1) {
                    const int32_t o = c * UVC_COL_CHUNK + __ctz_u32(todo);
                    const uint32_t bit = todo & (0u - todo);
                    const FragCol e = v.fcol[G.col_off + o];
                    if (m[UVC_FM_MUT_LINK] & bit) { v.rec_buf[out] = lo + o; v.rec_buf[out + 1] = (e.link_sym & 0xf); out += 2; }
                    if (m[UVC_FM_MUT_BASE_HQ] & bit) { v.rec_buf[out] = lo + o; v.rec_buf[out + 1] = e.base_sym; out += 2; }
                }
            }
        }
    }
    G.n_cov = n_cov; G.n_near_mut = n_near;
    // the compact per-read records that K3b stages
    for (int32_t k = 0; k < G.n_reads; k++) {
        const int32_t ri = v.frag_reads[G.read_off + k];
        const ReadRec & R = v.reads[ri];
        ReadFrag q;
        q.rend = R.rend; q.fragprev_maxrend = R.fragprev_maxrend;
        q.col_base = G.col_off - lo;
        q.n_cov = n_cov; q.n_near_mut = n_near;
        q.mq_term = (G.normMQ * G.normMQ) / UVC_SQR_QUAL_DIV;
        q.frag_strand = (int32_t)(fi * 2 + (G.strand ? 1 : 0));
        v.rfrag[ri] = q;
    }
}

// infer_max_qual_assuming_independence (main_conversion.hpp:943-974)
UVC_HD void infer_max_qual(int32_t & maxvqual, int32_t & argmaxAD, int32_t & argmaxBQ, const BatchView & v, int32_t max_qual, int32_t dec_qual, const int32_t *distr, int32_t totDP) {
    int32_t currAD = 0;
    maxvqual = 0; argmaxAD = 0; argmaxBQ = 0;
    const int32_t n = tmin(UVC_NUM_BUCKETS, max_qual / dec_qual);
    for (int32_t idx = 0; idx < n; idx++) {
        const int32_t q = distr[idx];
        if (0 == q) { continue; }
        currAD += q;
        const int32_t currBQ = max_qual - (dec_qual * idx);
        const double expBQ = v.ten_over_ln10 * log(((double)totDP / (double)currAD) + 2.220446049250313e-16);
        const int32_t currvqual = (int32_t)(currAD * (currBQ - expBQ));
        if (currvqual > maxvqual) { argmaxAD = currAD; argmaxBQ = currBQ; maxvqual = currvqual; }
    }
}

// ------------------------------------------------------------------------------------------------ K3b: one thread per position
// Fragment-level consensus gathered per position (main.hpp:2620-2733, 2757-2828): bDP, bTA, bTB, bMQ, the quality-bucket histogram and its
// reduction to bIAQb/bIADb/bIDQb, and the indel identities of fragments whose link consensus is an insertion or deletion.
struct K3bHot { int32_t n, cov, near; };       // bDP, bTA, bTB of one (strand, hot symbol)
struct K3bState {
    int64_t gp;
    int32_t p;
    int ref;
    // the two symbols nearly every fragment votes for - the reference base and LINK_M: depths and MQ sums in registers, their quality
    // histograms in a caller-provided array (shared memory in the CUDA kernel): hb[(type * UVC_NUM_BUCKETS + bucket) * hstride]
    int hot[2];
    int32_t hmaxq[2], hmq[2];
    K3bHot h0[2], h1[2];
    int32_t *hb; int hstride;
    uint32_t hb_saddr;     // CUDA: shared-memory address of hb (histogram updates are shared-memory reductions that nothing waits for)
    // thread-private accumulators of every other symbol (this thread is the only writer of the position's records): counted here, stored once
    // at the end. Only a few of the 14 symbols ever occur at one position, so the private arrays are zeroed lazily, per symbol, on first touch
    // (bit s of `touched`), and only touched symbols are reduced and stored: the local-memory traffic follows the data, not the array size.
    // (the arrays live in the caller's frame and are reached through pointers: a state struct made of scalars only stays in registers)
    uint32_t touched;
    int32_t *bucket;       // [UVC_NSYM * UVC_NUM_BUCKETS]
    int32_t *acc;          // [2 * UVC_NSYM * UVCGPU_NUM_FRAG_DEPTHS]
    int32_t *mq, *maxq;    // [UVC_NSYM] each
};
struct K3bArrays {
    int32_t bucket[UVC_NSYM * UVC_NUM_BUCKETS];
    int32_t acc[2 * UVC_NSYM * UVCGPU_NUM_FRAG_DEPTHS];
    int32_t mq[UVC_NSYM], maxq[UVC_NSYM];
};
UVC_HD void k3b_hb_set(K3bState & s, int32_t *hot_buckets, int hstride) {
    s.hb = hot_buckets; s.hstride = hstride; s.hb_saddr = 0;
#if defined(__CUDA_ARCH__)
    s.hb_saddr = (uint32_t)__cvta_generic_to_shared(hot_buckets);
    for (int k = 0; k < 2 * UVC_NUM_BUCKETS; k++) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(s.hb_saddr + (uint32_t)(k * hstride * 4)), "r"(0) : "memory"); }
#else
    for (int k = 0; k < 2 * UVC_NUM_BUCKETS; k++) { hot_buckets[k * hstride] = 0; }
#endif
}
UVC_HD void k3b_hb_inc(K3bState & s, int k) {
#if defined(__CUDA_ARCH__)
    asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(s.hb_saddr + (uint32_t)(k * s.hstride * 4)), "r"(1) : "memory");
#else
    s.hb[k * s.hstride] += 1;
#endif
}
UVC_HD int32_t k3b_hb_get(const K3bState & s, int k) {
#if defined(__CUDA_ARCH__)
    int32_t x;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(x) : "r"(s.hb_saddr + (uint32_t)(k * s.hstride * 4)) : "memory");
    return x;
#else
    return s.hb[k * s.hstride];
#endif
}
UVC_HD void k3b_touch(K3bState & s, const BatchView & v, int con) {
    if (!((s.touched >> con) & 1u)) {
        s.touched |= (1u << con);
        for (int k = 0; k < UVC_NUM_BUCKETS; k++) { s.bucket[con * UVC_NUM_BUCKETS + k] = 0; }
        for (int k = 0; k < UVCGPU_NUM_FRAG_DEPTHS; k++) { s.acc[con * UVCGPU_NUM_FRAG_DEPTHS + k] = 0; s.acc[(UVC_NSYM + con) * UVCGPU_NUM_FRAG_DEPTHS + k] = 0; }
        s.mq[con] = 0; s.maxq[con] = 8 + avg_bq(v, s.gp, con);
    }
}
UVC_HD void k3b_begin(K3bState & s, K3bArrays & arr, const BatchView & v, int64_t gp, int32_t *hot_buckets, int hstride) {
    const TileInfo & T = v.tiles[v.pos_tile[gp]];
    s.bucket = arr.bucket; s.acc = arr.acc; s.mq = arr.mq; s.maxq = arr.maxq;
    s.gp = gp;
    s.p = (int32_t)(gp - T.pos_off) + T.ext_beg;
    s.ref = v.refsym[gp];
    s.touched = 0;
    s.hot[0] = s.ref; s.hot[1] = UVC_LINK_M;
    k3b_hb_set(s, hot_buckets, hstride);
    for (int t = 0; t < 2; t++) {
        s.hmaxq[t] = (s.hot[t] < UVC_NSYM ? 8 + avg_bq(v, gp, s.hot[t]) : 0);
        s.hmq[t] = 0;
        s.h0[t].n = s.h0[t].cov = s.h0[t].near = 0; s.h1[t].n = s.h1[t].cov = s.h1[t].near = 0;
    }
}
// the first read q of its fragment that covers p; e is the fragment's column entry at p
UVC_HD void k3b_read(K3bState & s, const BatchView & v, const ReadFrag & q, const FragCol & e) {
    const uvcgpu_params & par = v.par;
    const int strand = (q.frag_strand & 1);
    #pragma unroll
    for (int type = 1; type >= 0; type--) {
        const int con = (type == 1 ? (e.link_sym & 0xf) : e.base_sym);
        const int32_t cc = (type == 1 ? e.link_cc : e.base_cc), tc = (type == 1 ? e.link_cc : e.base_tc);
        if (0 == tc) { continue; }
        if (con == s.hot[type]) {
            const int32_t max_qual = s.hmaxq[type];
            int32_t phredlike = tmin(cc * 2 - tc, max_qual);
            if (0x1 & par.fam_flag) { phredlike = tmin(phredlike, sscs_phred(par, s.ref, con)); }
            const int32_t pb = tmax(0, max_qual - phredlike);
            if (pb < UVC_NUM_BUCKETS) { k3b_hb_inc(s, type * UVC_NUM_BUCKETS + pb); }
            // the read (hence the strand) is the same for all lanes of the warp
            const int32_t s1 = strand, s0 = 1 - s1;      // (no branch on the strand)
            s.h1[type].n += s1; s.h1[type].cov += s1 * q.n_cov; s.h1[type].near += s1 * q.n_near_mut;
            s.h0[type].n += s0; s.h0[type].cov += s0 * q.n_cov; s.h0[type].near += s0 * q.n_near_mut;
            s.hmq[type] += q.mq_term;
            continue;
        }
        k3b_touch(s, v, con);
        const int32_t max_qual = s.maxq[con];
        int32_t phredlike = tmin(cc * 2 - tc, max_qual);
        if (0x1 & par.fam_flag) { phredlike = tmin(phredlike, sscs_phred(par, s.ref, con)); }
        const int32_t pb = tmax(0, max_qual - phredlike);
        if (pb < UVC_NUM_BUCKETS) { s.bucket[con * UVC_NUM_BUCKETS + pb] += 1; }
        int32_t *fd = s.acc + (strand ? UVC_NSYM * UVCGPU_NUM_FRAG_DEPTHS : 0);
        fd[con * UVCGPU_NUM_FRAG_DEPTHS + 0] += 1;
        fd[con * UVCGPU_NUM_FRAG_DEPTHS + 1] += q.n_cov;
        fd[con * UVCGPU_NUM_FRAG_DEPTHS + 2] += q.n_near_mut;
        s.mq[con] += q.mq_term;
        if (is_ins_symbol(con) || is_del_symbol(con)) {
            const FragRec & G = v.frags[q.frag_strand >> 1];
            const int32_t ev = indel_majority_of_reads(v, v.frag_reads + G.read_off, G.n_reads, s.p, con);
            if (ev >= 0) { rec_put6(v, UVC_REC_FRAG_INDEL, strand, con, s.p, ev, 1); }
        }
    }
}
UVC_HD void k3b_end(K3bState & s, const BatchView & v) {
    const int64_t gp = s.gp;
    // fold the hot symbols into the private arrays
    for (int type = 0; type < 2; type++) {
        if (0 == (s.h0[type].n | s.h1[type].n)) { continue; }
        const int con = s.hot[type];
        k3b_touch(s, v, con);
        for (int k = 0; k < UVC_NUM_BUCKETS; k++) { s.bucket[con * UVC_NUM_BUCKETS + k] += k3b_hb_get(s, type * UVC_NUM_BUCKETS + k); }
        int32_t *f0 = s.acc + con * UVCGPU_NUM_FRAG_DEPTHS, *f1 = s.acc + (UVC_NSYM + con) * UVCGPU_NUM_FRAG_DEPTHS;
        f0[0] += s.h0[type].n; f0[1] += s.h0[type].cov; f0[2] += s.h0[type].near;
        f1[0] += s.h1[type].n; f1[1] += s.h1[type].cov; f1[2] += s.h1[type].near;
        s.mq[con] += s.hmq[type];
    }
    const uint32_t touched = s.touched;
    const int32_t *acc = s.acc;
    int32_t *vq = v.vq + gp * UVC_NSYM * UVCGPU_NUM_VQ_TAGS;
    for (int type = 0; type < 2; type++) {
        const int s0 = (type == 0 ? UVC_BASE_A : UVC_LINK_M), s1 = (type == 0 ? UVC_BASE_NN : UVC_LINK_NN);
        int32_t totDP = 0;
        for (int y = s0; y <= s1; y++) { if ((touched >> y) & 1u) { totDP += acc[y * UVCGPU_NUM_FRAG_DEPTHS] + acc[(UVC_NSYM + y) * UVCGPU_NUM_FRAG_DEPTHS]; } }
        for (int y = s0; y <= s1; y++) {
            if (!((touched >> y) & 1u)) { continue; }
            for (int strand = 0; strand < 2; strand++) {
                int32_t *g = v.fragdepth + ((strand * v.n_pos + gp) * UVC_NSYM + y) * UVCGPU_NUM_FRAG_DEPTHS;
                for (int k = 0; k < UVCGPU_NUM_FRAG_DEPTHS; k++) { g[k] = acc[(strand * UVC_NSYM + y) * UVCGPU_NUM_FRAG_DEPTHS + k]; }
            }
            vq[y * UVCGPU_NUM_VQ_TAGS + 4] = s.mq[y];
            int32_t q, ad, bq;
            infer_max_qual(q, ad, bq, v, s.maxq[y], 1, s.bucket + y * UVC_NUM_BUCKETS, totDP);
            vq[y * UVCGPU_NUM_VQ_TAGS + 5] = q; vq[y * UVCGPU_NUM_VQ_TAGS + 6] = ad; vq[y * UVCGPU_NUM_VQ_TAGS + 7] = bq;
        }
    }
}
UVC_HD void k3b_position(const BatchView & v, int64_t gp, const Win & w) {
    K3bState s;
    K3bArrays arr;
    int32_t hot_buckets[2 * UVC_NUM_BUCKETS];
    k3b_begin(s, arr, v, gp, hot_buckets, 1);
    for (int64_t ri = w.ulo; ri < w.uhi; ri++) {
        if (ri < w.lo || ri >= w.hi) { continue; }
        const ReadFrag q = v.rfrag[ri];
        if (q.rend <= s.p || q.fragprev_maxrend > s.p) { continue; }   // not covering, or an earlier read of the same fragment already handled p
        k3b_read(s, v, q, v.fcol[q.col_base + s.p]);
    }
    k3b_end(s, v);
}

// ------------------------------------------------------------------------------------------------ family-level helpers
// Per (family, strand, position): number of fragments voting for each symbol after the base-quality filter (read_family_con_ampl,
// GenericSymbol2Count::updateByFiltering, main.hpp:466-495) and, optionally, the major-minus-minor quality sums (read_family_mmm_ampl,
// updateByMajorMinusMinor, main.hpp:497-520).
UVC_HD void fam_counts(const BatchView & v, const FamRec & F, int strand, int32_t p, int64_t gp, int32_t con[UVC_NSYM], int32_t *mmm) {
    votes_zero(con);
    if (mmm) { votes_zero(mmm); }
    const bool ignore_padded_del = (v.par.microadjust_padded_deletion_flag & 0x1); // Illumina/BGI branch of main.hpp:2908
    for (int32_t g = F.frag_off[strand]; g < F.frag_off[strand] + F.n_frags[strand]; g++) {
        const FragRec & G = v.frags[g];
        if (p < G.lo || p >= G.hi) { continue; }
        const FragCol e = v.fcol[G.col_off + (p - G.lo)];
        if (0 == (e.link_cc | e.base_tc)) { continue; }
        // link: the reference is counted once, so count_sum == count_max and the adjusted quality is the count itself
        if (e.link_cc > 0) { con[e.link_sym & 0xf] += 1; if (mmm) { mmm[e.link_sym & 0xf] += e.link_cc; } }
        int a = e.base_sym; int32_t cc = e.base_cc, tc = e.base_tc;
        if (ignore_padded_del && (e.link_sym & 0x80)) {   // consensus over A..T differs only when N / padded-deletion votes exist: recompute
            int32_t c[UVC_NSYM];
            frag_votes(v, G, p, gp, c);
            base_consensus(c, true, a, cc, tc);
        }
        int32_t adj = tmax(cc * 2, tc) - tc;
        if (adj >= v.par.fam_thres_highBQ_snv && adj > 0) { con[a] += 1; }
        if (mmm) {
            adj = tmax((int32_t)e.base_cc * 2, (int32_t)e.base_tc) - (int32_t)e.base_tc;
            if (adj > 0) { mmm[e.base_sym] += adj; }
        }
    }
}

// majority indel event of fragment G at p if its link consensus is `symbol`, else -1
UVC_HD int32_t frag_link_indel_event(const BatchView & v, const FragRec & G, int32_t p, int64_t gp, int symbol) {
    (void)gp;
    const FragCol e = frag_entry(v, G, p);
    if ((e.link_sym & 0xf) != symbol || 0 == e.link_cc) { return -1; }
    return indel_majority_of_reads(v, v.frag_reads + G.read_off, G.n_reads, p, symbol);
}

// Majority identity of the family's indel map for `symbol` at p (each fragment whose link consensus is `symbol` adds its own majority identity
// once, main.hpp:1679-1685); *count receives the number of fragments behind the winner.
UVC_HD int32_t fam_indel_majority(const BatchView & v, const FamRec & F, int strand, int32_t p, int64_t gp, int symbol, int32_t *count) {
    int32_t best = -1, best_n = 0;
    const int32_t g0 = F.frag_off[strand], g1 = F.frag_off[strand] + F.n_frags[strand];
    for (int32_t g = g0; g < g1; g++) {
        const int32_t e = frag_link_indel_event(v, v.frags[g], p, gp, symbol);
        if (e < 0) { continue; }
        int32_t n = 0; bool first = true;
        for (int32_t g2 = g0; g2 < g1; g2++) {
            const int32_t e2 = (g2 == g ? e : frag_link_indel_event(v, v.frags[g2], p, gp, symbol));
            if (e2 < 0) { continue; }
            if (0 == indel_cmp(v, v.ev[e], v.ev[e2])) { n++; if (g2 < g) { first = false; } }
        }
        if (!first) { continue; }
        if (best < 0 || n > best_n || (n == best_n && indel_cmp(v, v.ev[e], v.ev[best]) > 0)) { best = e; best_n = n; }
    }
    if (count) { *count = best_n; }
    return best;
}

UVC_HD void plain_consensus(const int32_t c[UVC_NSYM], int type, int & a, int32_t & cc, int32_t & tc) {
    if (type == 1) { link_consensus(c, false, a, cc, tc); } else { base_consensus(c, false, a, cc, tc); }
}

UVC_HD bool fam_is_good(const uvcgpu_params & par, const FamRec & F, int32_t cc, int32_t tc) {
    return (par.fam_thres_dup1add <= tc) && (cc * 100 >= tc * par.fam_thres_dup1perc) && ((F.duplexflag & 0x1) || (par.fam_flag & 0x2));
}

// ------------------------------------------------------------------------------------------------ KM: one thread per family-column entry
// Per (family, strand, position): the fragment votes of the family (read_family_con_ampl / read_family_mmm_ampl of main.hpp:2999-3040, 3392-3420)
// reduced to the consensus triples that family loop 1, loop 2, the duplex step, the family end scan and the haplotype strings consume.
// A warp owns 32 consecutive positions of one (family, strand): the fragment entries it reads are consecutive in memory.
UVC_HD void km_family_column(const BatchView & v, int64_t i) {
    const int32_t fs = v.mchunk_fs[i / UVC_COL_CHUNK];
    const FamRec & F = v.fams[fs >> 1];
    const int strand = (fs & 1);
    const int32_t o = (int32_t)(i - F.col_off[strand]);
    if (o >= F.hi[strand] - F.lo[strand]) { return; }
    const int32_t p = F.lo[strand] + o;
    const TileInfo & T = v.tiles[F.tile];
    int32_t con[UVC_NSYM], mmm[UVC_NSYM];
    fam_counts(v, F, strand, p, T.pos_off + (p - T.ext_beg), con, mmm);
    FamCol m;
    for (int type = 0; type < 2; type++) {
        int a; int32_t cc, tc;
        plain_consensus(con, type, a, cc, tc);
        m.a1[type] = (uint8_t)a; m.cc1[type] = (uint16_t)tmin(cc, 65535); m.tc1[type] = (uint16_t)tmin(tc, 65535);
        plain_consensus(mmm, type, a, cc, tc);
        m.a2[type] = (uint8_t)a; m.mmm_cc[type] = (uint32_t)cc; m.mmm_tot[type] = (uint32_t)tc; m.con_a2[type] = (uint16_t)tmin(con[a], 65535);
    }
    v.mcol[i] = m;
}

// The family entry of a strand that holds a single fragment, derived from that fragment's entry: at most one vote per symbol type
// (read_family_con_ampl / read_family_mmm_ampl with one fragment, main.hpp:466-520), so nothing needs to be materialised.
UVC_HD FamCol famcol_from_frag(const FragCol & e, const uvcgpu_params & par) {
    FamCol m;
    m.a1[0] = m.a2[0] = (uint8_t)UVC_BASE_NN; m.a1[1] = m.a2[1] = (uint8_t)UVC_LINK_NN;   // consensus over all-zero counts returns the last symbol of the type
    m.cc1[0] = m.tc1[0] = m.con_a2[0] = 0; m.cc1[1] = m.tc1[1] = m.con_a2[1] = 0;
    m.mmm_cc[0] = m.mmm_tot[0] = 0; m.mmm_cc[1] = m.mmm_tot[1] = 0;
    if (e.link_cc > 0) {
        m.a1[1] = m.a2[1] = (uint8_t)(e.link_sym & 0xf); m.cc1[1] = m.tc1[1] = m.con_a2[1] = 1; m.mmm_cc[1] = m.mmm_tot[1] = e.link_cc;
    }
    const int32_t adj = tmax((int32_t)e.base_cc * 2, (int32_t)e.base_tc) - (int32_t)e.base_tc;
    if (adj > 0) {
        m.a2[0] = e.base_sym; m.mmm_cc[0] = m.mmm_tot[0] = (uint32_t)adj;
        if (adj >= par.fam_thres_highBQ_snv) { m.a1[0] = e.base_sym; m.cc1[0] = m.tc1[0] = m.con_a2[0] = 1; }
    }
    return m;
}

// the entry of (family, strand) at p; all-zero counts outside the strand's covered extent
UVC_HD FamCol fam_entry(const BatchView & v, const FamRec & F, int strand, int32_t p) {
    if (p < F.lo[strand] || p >= F.hi[strand]) {
        FamCol z;
        for (int t = 0; t < 2; t++) { z.a1[t] = z.a2[t] = (uint8_t)(t ? UVC_LINK_NN : UVC_BASE_NN); z.cc1[t] = z.tc1[t] = z.con_a2[t] = 0; z.mmm_cc[t] = z.mmm_tot[t] = 0; }
        return z;
    }
    if (F.direct_frag[strand] >= 0) { return famcol_from_frag(frag_entry(v, v.frags[F.direct_frag[strand]], p), v.par); }
    return v.mcol[F.col_off[strand] + (p - F.lo[strand])];
}

// the entry a read's (family, strand) has at p, through the read's compact record (p is covered by the read, hence inside the extent)
UVC_HD FamCol fam_entry_of_read(const BatchView & v, const ReadFam & q, int32_t p) {
    if (q.flags & UVC_RF_DIRECT) { return famcol_from_frag(v.fcol[q.col_base + p], v.par); }
    return v.mcol[q.col_base + p];
}

// ------------------------------------------------------------------------------------------------ K4a: one thread per (family, strand)
// no_strict_bias_pos_min/max (main.hpp:2959-2998): the outermost positions, from either end, at which the family forms a tier-2 base consensus.
UVC_HD void k4a_family_strand(const BatchView & v, int64_t i) {
    FamRec & F = v.fams[i >> 1];
    const int strand = (int)(i & 1);
    if (0 == F.n_frags[strand]) { return; }
    F.nsb_min[strand] = F.end2[strand]; F.nsb_max[strand] = F.beg2[strand];
    if (!F.qlen_ok[strand] || !((F.duplexflag & 0x1) || (v.par.fam_flag & 0x2))) { return; }
    const int32_t lo = F.lo[strand], hi = F.hi[strand];   // covered extent (positions outside have no votes)
    for (int dir = 0; dir < 2; dir++) {
        for (int32_t p = (dir ? hi - 1 : lo); (dir ? p >= lo : p < hi); p += (dir ? -1 : 1)) {
            const FamCol m = fam_entry(v, F, strand, p);
            const int a = m.a1[0]; const int32_t cc = m.cc1[0], tc = m.tc1[0];
            if (0 == tc) { continue; }
            if (fam_is_good(v.par, F, cc, tc) && (UVC_BASE_N != a) && (UVC_BASE_NN != a)) {
                if (dir) { F.nsb_max[strand] = p; } else { F.nsb_min[strand] = p; }
                break;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ K4: one thread per position
// updateByAlns3UsingFQ gathered per position: family loop 1 (main.hpp:2999-3355), then - because its only cross-family dependence, cDPM/cDPm,
// is per position - family loop 2 (main.hpp:3392-3551) and the per-strand reduction of the quality buckets (main.hpp:3552-3591).
//
// Loop 2 needs the completed cDPM/cDPm of loop 1 only for families that can enter a quality bucket (tot_nfrags >= fam_thres_dup1add) and its
// single-strand / duplex consensus counters only for duplex-UMI families. For every other (family, strand) - all of them on non-UMI data - the
// whole of loop 2 is "cDP1 += 1", which loop 1 does on the spot from the same column entry; the second pass over the window then only runs for
// positions that saw a family it still has work for (need2).
#define UVC_K4_LIST 24
struct K4Hot { int32_t dp1, dp12, dp2, dp3, dpM, dpm, dp21; };   // cDP1, cDP12, cDP2, cDP3, cDPM, cDPm, cDP21 (cDPD: duplex-UMI families only, never hot)
UVC_HD void k4hot_zero(K4Hot & h) { h.dp1 = h.dp12 = h.dp2 = h.dp3 = h.dpM = h.dpm = h.dp21 = 0; }
struct K4State {
    const TileInfo *T;
    int64_t gp;
    int32_t p;
    int ref;
    int32_t baq_last, tn_add;
    const int32_t *baq, *baq2;
    uvcgpu_thres_set th;
    uvcgpu_faminfo_set *finfo;
    uint32_t touched;      // bit strand * 14 + symbol: facc / bucket rows that are live (rows are zeroed lazily on first touch, see K3b)
    // reads (offsets from the start of the window) that loop 1 leaves work for; more than UVC_K4_LIST of them: the whole window is walked again
    int32_t n_need2;
    uint16_t *need2;       // [UVC_K4_LIST]
    bool lone_is_plain;    // a strand with one fragment (cc = tc = 1) can only count as cDP12 / cDP21 / cDP1 under the current thresholds
    // family depth counters of the two symbols nearly every family votes for - the reference base and LINK_M - per strand: registers
    int hot[2];            // [type]
    K4Hot h0[2], h1[2];    // strand 0 / strand 1, [type]
    // thread-private family depth counters of every other symbol (this thread is the position's only writer): stored once at the end.
    // (the arrays live in the caller's frame and are reached through pointers: a state struct made of scalars only stays in registers)
    int32_t *facc;         // [2 * UVC_NSYM * UVCGPU_NUM_FAM_DEPTHS]
    int32_t *bucket;       // [2 * UVC_NSYM * UVC_NUM_BUCKETS]
};
struct K4Arrays {
    int32_t facc[2 * UVC_NSYM * UVCGPU_NUM_FAM_DEPTHS];
    int32_t bucket[2 * UVC_NSYM * UVC_NUM_BUCKETS];
    uint16_t need2[UVC_K4_LIST];
};
enum { UVC_cDP1 = 0, UVC_cDP12 = 1, UVC_cDP2 = 2, UVC_cDP3 = 3, UVC_cDPM = 4, UVC_cDPm = 5, UVC_cDP21 = 6, UVC_cDPD = 7 };

UVC_HD void k4_touch(K4State & s, int strand, int sym) {
    const int ix = strand * UVC_NSYM + sym;
    if (!((s.touched >> ix) & 1u)) {
        s.touched |= (1u << ix);
        for (int k = 0; k < UVCGPU_NUM_FAM_DEPTHS; k++) { s.facc[ix * UVCGPU_NUM_FAM_DEPTHS + k] = 0; }
        for (int k = 0; k < UVC_NUM_BUCKETS; k++) { s.bucket[ix * UVC_NUM_BUCKETS + k] = 0; }
    }
}

UVC_HD void k4_begin(K4State & s, K4Arrays & arr, const BatchView & v, int64_t gp) {
    const TileInfo & T = v.tiles[v.pos_tile[gp]];
    s.facc = arr.facc; s.bucket = arr.bucket; s.need2 = arr.need2;
    s.T = &T; s.gp = gp;
    s.p = (int32_t)(gp - T.pos_off) + T.ext_beg;
    const int64_t po = T.pos_off - T.ext_beg;
    s.baq = v.baq + po; s.baq2 = v.baq2 + po;
    s.ref = v.refsym[gp];
    s.baq_last = T.ext_end - 1;
    s.tn_add = (v.par.is_tumor_vcf_provided ? 4 : 0);
    s.th = v.thres[gp];
    s.finfo = v.faminfo + gp * UVC_NSYM;
    s.touched = 0;
    s.n_need2 = 0;
    s.hot[0] = s.ref; s.hot[1] = UVC_LINK_M;
    for (int t = 0; t < 2; t++) { k4hot_zero(s.h0[t]); k4hot_zero(s.h1[t]); }
    const uvcgpu_params & par = v.par;
    s.lone_is_plain = (par.fam_thres_dup1add > 1)
            && !(par.fam_thres_dup2add <= 1 && 100 >= par.fam_thres_dup2perc)
            && !(1 >= par.fam_thres_emperr_all_flat_snv && 100 >= par.fam_thres_emperr_con_perc_snv)
            && !(1 >= par.fam_thres_emperr_all_flat_indel && 100 >= par.fam_thres_emperr_con_perc_indel);
}

UVC_HD void k4hot_add(K4Hot & h, const K4Hot & d) {
    h.dp1 += d.dp1; h.dp12 += d.dp12; h.dp2 += d.dp2; h.dp3 += d.dp3; h.dpM += d.dpM; h.dpm += d.dpm; h.dp21 += d.dp21;
}
// adds the increments d to the family depth counters of (strand, symbol a of type `type`)
UVC_HD void k4_add(K4State & s, int strand, int type, int a, const K4Hot & d) {
    if (a == s.hot[type]) {
        // (no branch on the strand: both register sets take the increments, one of them multiplied by zero)
        const int32_t s1 = (strand ? 1 : 0), s0 = 1 - s1;
        K4Hot & h0 = s.h0[type]; K4Hot & h1 = s.h1[type];
        h0.dp1 += s0 * d.dp1; h0.dp12 += s0 * d.dp12; h0.dp2 += s0 * d.dp2; h0.dp3 += s0 * d.dp3; h0.dpM += s0 * d.dpM; h0.dpm += s0 * d.dpm; h0.dp21 += s0 * d.dp21;
        h1.dp1 += s1 * d.dp1; h1.dp12 += s1 * d.dp12; h1.dp2 += s1 * d.dp2; h1.dp3 += s1 * d.dp3; h1.dpM += s1 * d.dpM; h1.dpm += s1 * d.dpm; h1.dp21 += s1 * d.dp21;
    } else {
        k4_touch(s, strand, a);
        int32_t *fd = s.facc + (strand * UVC_NSYM + a) * UVCGPU_NUM_FAM_DEPTHS;
        fd[UVC_cDP1] += d.dp1; fd[UVC_cDP12] += d.dp12; fd[UVC_cDP2] += d.dp2; fd[UVC_cDP3] += d.dp3; fd[UVC_cDPM] += d.dpM; fd[UVC_cDPm] += d.dpm; fd[UVC_cDP21] += d.dp21;
    }
}
// current value of one family depth counter of (strand, symbol a)
UVC_HD void k4_major_minor(const K4State & s, int strand, int type, int a, int32_t & major, int32_t & minor) {
    if (a == s.hot[type]) {
        const K4Hot & h = (strand ? s.h1[type] : s.h0[type]);
        major = h.dpM; minor = h.dpm;
    } else {
        const int32_t *fd = s.facc + (strand * UVC_NSYM + a) * UVCGPU_NUM_FAM_DEPTHS;   // touched by loop 1 or zero
        const bool live = ((s.touched >> (strand * UVC_NSYM + a)) & 1u);
        major = (live ? fd[UVC_cDPM] : 0); minor = (live ? fd[UVC_cDPm] : 0);
    }
}

// does loop 1 (with the loop-2 share described above) cover everything this (family, strand) entry asks of loop 2 for symbol type `type`?
UVC_HD bool k4_loop2_done_in_loop1(const uvcgpu_params & par, const ReadFam & q, const FamCol & m, int type) {
    return (0 == (q.flags & UVC_RF_DUPLEX_UMI)) && ((int32_t)m.tc1[type] < par.fam_thres_dup1add);
}

// Loop 1 (and its share of loop 2) for a single-fragment strand of a family without a duplex UMI, straight from the fragment's column entry:
// every vote has cc = tc = 1, so under the usual thresholds (lone_is_plain) it is "cDP12, cDP21 and cDP1 of the symbol" and nothing else - the
// same increments k4_loop1_read derives through famcol_from_frag. Returns false (nothing done) when the general path has to run: indel
// symbols (their identities are recorded), duplex-UMI families, unusual thresholds.
UVC_HD bool k4_loop1_lone_fragment(K4State & s, const BatchView & v, const ReadFam & q, const FragCol & e) {
    if (!s.lone_is_plain || (q.flags & UVC_RF_DUPLEX_UMI)) { return false; }
    const int la = (e.link_sym & 0xf);
    if (e.link_cc > 0 && (is_ins_symbol(la) || is_del_symbol(la))) { return false; }
    const int strand = (int)(q.flags & UVC_RF_STRAND);
    K4Hot d;
    if (e.link_cc > 0) {
        k4hot_zero(d);
        d.dp12 = 1; d.dp21 = 1; d.dp1 = 1;
        k4_add(s, strand, 1, la, d);
    }
    const int32_t adj = tmax((int32_t)e.base_cc * 2, (int32_t)e.base_tc) - (int32_t)e.base_tc;
    if (adj > 0) {
        k4hot_zero(d);
        d.dp1 = 1;
        if (adj >= v.par.fam_thres_highBQ_snv) { d.dp12 = 1; d.dp21 = 1; }
        k4_add(s, strand, 0, e.base_sym, d);
    }
    return true;
}

// loop 1 for the first read q of its (family, strand) that covers p; m is that (family, strand)'s column entry at p; woff = offset of the read
// from the start of the position's window
UVC_HD void k4_loop1_read(K4State & s, const BatchView & v, const ReadFam & q, const FamCol & m, int64_t woff) {
    const uvcgpu_params & par = v.par;
    const int32_t p = s.p;
    const int64_t gp = s.gp;
    const int32_t *baq = s.baq, *baq2 = s.baq2;
    const uvcgpu_thres_set & th = s.th;
    const FamRec & F = v.fams[q.fam];      // only dereferenced on the tier-2 (UMI family) path
    const int strand = (int)(q.flags & UVC_RF_STRAND);
    #pragma unroll
    for (int type = 1; type >= 0; type--) {
        const int a = m.a1[type]; const int32_t cc = m.cc1[type], tc = m.tc1[type];
        if (0 == tc) { continue; }
        // the read (hence strand) is the same for all lanes of the warp: only `a == hot` can diverge
        K4Hot d;
        k4hot_zero(d);
        d.dp12 = 1;
        if (1 == tc) { d.dp21 = 1; }
        const bool is_indel = (is_ins_symbol(a) || is_del_symbol(a));
        int32_t fam_ev = -2;   // family-majority indel event, computed lazily
        if (fam_is_good(par, F, cc, tc)) {
            d.dp2 = 1;
            if (is_indel) {
                fam_ev = fam_indel_majority(v, F, strand, p, gp, a, NULL);
                if (fam_ev >= 0) { rec_put6(v, UVC_REC_CDP2_INDEL, strand, a, p, fam_ev, 1); }
            }
            // family-level position / BAQ bias (main.hpp:3208-3318)
            int32_t rbeg = tmin(F.nsb_min[strand], p);
            int32_t rend = tmax(F.nsb_max[strand], p);
            const int32_t l2r = F.l2r_end_median[strand], r2l = F.r2l_end_median[strand];
            const bool nonconf_middle = (l2r <= (r2l + par.indel_adj_tracklen_dist));
            if (nonconf_middle && p < r2l) { rend = tmax(tmin(l2r, tmin(r2l, rend)), p); }
            if (nonconf_middle && l2r < p) { rbeg = tmin(tmax(l2r, tmax(r2l, rbeg)), p); }
            uvcgpu_faminfo_set & fi = s.finfo[a];
            const bool isGap = (type == 1);
            const int32_t l_nb = nnminus(p + 1, rbeg);
            const int32_t r_nb = nnminus(rend, p);
            const int32_t LPxT = (isGap ? th.aLPxT : tmin(th.aLPxT, th.aRPxT));
            int32_t indel_len = 0;
            if (is_ins_symbol(a)) {
                // QUIRK: the reference takes the majority COUNT of the family's inserted sequences as "indel_len" (main.hpp:3238-3244)
                for (int sym = UVC_LINK_I3P; sym <= UVC_LINK_I1; sym++) { int32_t n = 0; fam_indel_majority(v, F, strand, p, gp, sym, &n); indel_len = tmax(indel_len, n); }
            } else if (is_del_symbol(a)) {
                for (int sym = UVC_LINK_D3P; sym <= UVC_LINK_D1; sym++) { int32_t n = 0; fam_indel_majority(v, F, strand, p, gp, sym, &n); indel_len = tmax(indel_len, n); }
            }
            const bool far_from_edge = (l_nb + (is_ins_symbol(a) ? nnminus(indel_len, par.microadjust_nobias_pos_indel_maxlen) : 0) >= LPxT) && (r_nb >= th.aRPxT);
            // (the family records of a position have this thread as their only writer; reductions instead of ~15 dependent read-modify-writes)
            if (far_from_edge) {
                bidir_bias_red(&fi.c2LP1, &fi.c2LP2, &fi.c2RP1, &fi.c2RP2, th.aLP1t, th.aLP2t, th.aRP1t, th.aRP2t, l_nb, r_nb, true, 0);
                atomic_add(&fi.c2LPL, l_nb); atomic_add(&fi.c2RPL, r_nb);
            }
            if (nnminus(p + 1, F.nsb_min[strand]) >= par.bias_thres_strict_c2LRP0) { atomic_add(&fi.c2LP0, 1); }
            if (nnminus(F.nsb_max[strand], p) >= par.bias_thres_strict_c2LRP0) { atomic_add(&fi.c2RP0, 1); }
            const int32_t seg_l_baq = baq[p] - baq[tmax(rbeg, nnminus(p, UVC_MAX_STR_N_BASES))] + 1;
            const int32_t ridx = tmin(rend - 1, tmin(p + UVC_MAX_STR_N_BASES, s.baq_last));
            const int32_t seg_r_baq0 = baq[ridx] - baq[p] + 1;
            const int32_t seg_r_baq = (isGap ? tmin(seg_r_baq0, baq2[ridx] - baq2[p] + 7) : seg_r_baq0);
            const int32_t highBAQ = par.bias_thres_highBAQ + (isGap ? 0 : 3);
            if (seg_l_baq >= highBAQ && seg_r_baq >= highBAQ) {
                bidir_bias_red(&fi.c2LB1, &fi.c2LB2, &fi.c2RB1, &fi.c2RB2, par.bias_thres_BAQ1, par.bias_thres_BAQ2, par.bias_thres_BAQ1, par.bias_thres_BAQ2,
                        seg_l_baq, seg_r_baq, true, 0);
                atomic_add(&fi.c2LBL, (int64_t)seg_l_baq); atomic_add(&fi.c2RBL, (int64_t)seg_r_baq);
            }
            atomic_add(&fi.c2BQ2, 1);
        }
        if (par.fam_thres_dup2add <= tc && (cc * 100 >= tc * par.fam_thres_dup2perc)) { d.dp3 = 1; }
        if (is_indel) {
            if (fam_ev == -2) { fam_ev = fam_indel_majority(v, F, strand, p, gp, a, NULL); }
            if (fam_ev >= 0) { rec_put6(v, UVC_REC_FAM_INDEL, strand, a, p, fam_ev, 1); }
        }
        const bool is_subst = (a <= UVC_BASE_NN);
        const int32_t flat = (is_subst ? par.fam_thres_emperr_all_flat_snv : par.fam_thres_emperr_all_flat_indel);
        const int32_t perc = (is_subst ? par.fam_thres_emperr_con_perc_snv : par.fam_thres_emperr_con_perc_indel);
        if (tc >= flat && cc * 100 >= tc * perc) {
            // every other symbol of the type adds its own count to cDPm and, QUIRK, the whole total to cDPM (main.hpp:3343-3352)
            const int32_t n_other = (type == 0 ? (UVC_BASE_NN - UVC_BASE_A) : (UVC_LINK_NN - UVC_LINK_M));
            d.dpm = tc - cc;
            d.dpM = tc * n_other;
        }
        k4_add(s, strand, type, a, d);
    }
    // the share of loop 2 that needs nothing from the other families (see the header comment)
    bool need2 = (0 != (q.flags & UVC_RF_DUPLEX_UMI));
    #pragma unroll
    for (int type = 1; type >= 0; type--) {
        if (0 == m.mmm_tot[type]) { continue; }
        if (k4_loop2_done_in_loop1(par, q, m, type)) {
            K4Hot d;
            k4hot_zero(d);
            d.dp1 = 1;
            k4_add(s, strand, type, m.a2[type], d);
        } else {
            need2 = true;
        }
    }
    if (need2) {
        if (s.n_need2 < UVC_K4_LIST && woff < 65536) { s.need2[s.n_need2] = (uint16_t)woff; s.n_need2 += 1; } else { s.n_need2 = UVC_K4_LIST + 1; }
    }
}

// loop 2 for a read q that covers p (q.rend > p), for what loop 1 left over. Only reads that loop 1 visited can have work here: the duplex
// step wants the first read of the whole family at p, which is also the first read of its own strand.
UVC_HD void k4_loop2_read(K4State & s, const BatchView & v, const ReadFam & q, const FamCol *pre = NULL) {   // pre: the read's (family, strand) entry at p, if the caller holds it
    const uvcgpu_params & par = v.par;
    const int32_t p = s.p;
    const int64_t gp = s.gp;
    const FamRec & F = v.fams[q.fam];      // only dereferenced for indel identities and duplex families
    const bool is_duplex_umi = (0 != (q.flags & UVC_RF_DUPLEX_UMI));
    const bool will_inc_dscs = (is_duplex_umi && (q.flags & UVC_RF_BOTH_STRANDS));
    const bool will_inc_sscs = (is_duplex_umi && !will_inc_dscs);
    if (q.famprev_maxrend <= p) {   // first read of its (family, strand) that covers p
        const int strand = (int)(q.flags & UVC_RF_STRAND);
        int32_t *fd = s.facc + strand * (UVC_NSYM * UVCGPU_NUM_FAM_DEPTHS);
        const FamCol m = (pre ? *pre : fam_entry_of_read(v, q, p));
        #pragma unroll
        for (int type = 1; type >= 0; type--) {
            const int a = m.a2[type]; const int32_t con_sumBQs = (int32_t)m.mmm_cc[type], tot_sumBQs = (int32_t)m.mmm_tot[type];
            if (0 == tot_sumBQs) { continue; }
            if (k4_loop2_done_in_loop1(par, q, m, type)) { continue; }
            const int32_t con_nfrags = m.con_a2[type];
            const int32_t tot_nfrags = m.tc1[type];
            {
                K4Hot d;
                k4hot_zero(d);
                d.dp1 = 1;
                k4_add(s, strand, type, a, d);
            }
            if (will_inc_sscs && (tot_nfrags >= par.fam_thres_dup1add) && (con_nfrags * 100 >= tot_nfrags * par.fam_thres_dup1perc)) {
                k4_touch(s, strand, a);
                fd[a * UVCGPU_NUM_FAM_DEPTHS + UVC_cDPD] += 1;
                if (is_ins_symbol(a) || is_del_symbol(a)) {
                    const int32_t e = fam_indel_majority(v, F, strand, p, gp, a, NULL);
                    if (e >= 0) { rec_put6(v, UVC_REC_C2D_INDEL, strand, a, p, e, 1); }
                }
            }
            if (tot_nfrags >= par.fam_thres_dup1add) {     // the family's consensus quality only matters if it enters a bucket (main.hpp:3445-3455)
                const int32_t avgBQ = ((0 == tot_nfrags) ? 1 : (con_sumBQs / tot_nfrags));
                int32_t majorcount, minorcount;
                k4_major_minor(s, strand, type, a, majorcount, minorcount);
                const double prior_weight = 1.0 / (minorcount + 1.0);
                const double p2p = v.phred2prob_tab[tmin(tmax(avgBQ, 0), 127)];
                const double realphred = -10 * log((minorcount + prior_weight) / (majorcount + minorcount + prior_weight / p2p)) / v.ln10;
                const int32_t indep_frag_phred = (int32_t)round(((con_nfrags * 2) - tot_nfrags) * realphred);
                int32_t confam_qual;
                if (type == 1) {
                    confam_qual = tmax(1, tmin(indep_frag_phred, par.fam_phred_indel_inc_before_barcode_labeling + (int32_t)round(realphred)));
                } else {
                    confam_qual = tmax(1, tmin(indep_frag_phred, (con_sumBQs * 2) - tot_sumBQs));
                }
                const int32_t max_qual = sscs_phred(par, s.ref, a) + s.tn_add;
                const int32_t confam_qual2 = tmin(confam_qual, max_qual);
                const int32_t pb = (max_qual - confam_qual2 + 2) / 4;
                if (pb >= 0 && pb < UVC_NUM_BUCKETS) { k4_touch(s, strand, a); s.bucket[(strand * UVC_NSYM + a) * UVC_NUM_BUCKETS + pb] += 1; }
            }
        }
    }
    if (will_inc_dscs && q.fambothprev_maxrend <= p) {   // first read of the duplex family that covers p
        int32_t *dup = v.duplex + gp * UVC_NSYM * UVCGPU_NUM_DUPLEX_DEPTHS;
        int32_t dcount[UVC_NSYM];
        votes_zero(dcount);
        int link_con[2] = {-1, -1};
        for (int strand = 0; strand < 2; strand++) {
            const FamCol m = fam_entry(v, F, strand, p);
            for (int type = 1; type >= 0; type--) {
                const int a = m.a1[type]; const int32_t cc = m.cc1[type], tc = m.tc1[type];
                const int32_t adj = tmax(cc * 2, tc) - tc;
                if (adj >= 1 && adj > 0) { dcount[a] += 1; }
                if (type == 1 && cc > 0) { link_con[strand] = a; }
            }
        }
        for (int type = 0; type < 2; type++) {
            int a; int32_t cc, tc;
            plain_consensus(dcount, type, a, cc, tc);
            if (0 < tc) { dup[a * UVCGPU_NUM_DUPLEX_DEPTHS + 0] += 1; }
            if (1 < tc) {
                dup[a * UVCGPU_NUM_DUPLEX_DEPTHS + 1] += 1;
                if (is_ins_symbol(a) || is_del_symbol(a)) {
                    // majority identity of the duplex map: one entry per strand whose family link consensus is this symbol
                    int32_t e0 = (link_con[0] == a ? fam_indel_majority(v, F, 0, p, gp, a, NULL) : -1);
                    int32_t e1 = (link_con[1] == a ? fam_indel_majority(v, F, 1, p, gp, a, NULL) : -1);
                    int32_t e = (e0 < 0 ? e1 : (e1 < 0 ? e0 : ((indel_cmp(v, v.ev[e0], v.ev[e1]) >= 0) ? e0 : e1)));
                    if (e >= 0) { rec_put6(v, UVC_REC_C2D_INDEL, 0, a, p, e, 1); rec_put6(v, UVC_REC_C2D_INDEL, 1, a, p, e, 1); }
                }
            }
        }
    }
}

// loop 2 over the short list of reads that loop 1 recorded (the common case on non-UMI data: a few multi-fragment families per position)
UVC_HD void k4_loop2_listed(K4State & s, const BatchView & v, const Win & w) {
    if (s.n_need2 > UVC_K4_LIST) { return; }
    for (int32_t j = 0; j < s.n_need2; j++) { k4_loop2_read(s, v, v.rfam[w.lo + s.need2[j]]); }
}

// per-strand reduction of the family quality buckets (main.hpp:3552-3591) and the store of the position's family depth records
UVC_HD void k4_end(K4State & s, const BatchView & v) {
    const uvcgpu_params & par = v.par;
    const int64_t gp = s.gp;
    int32_t *vq = v.vq + gp * UVC_NSYM * UVCGPU_NUM_VQ_TAGS;
    for (int strand = 0; strand < 2; strand++) {
        for (int type = 0; type < 2; type++) {
            const K4Hot & h = (strand ? s.h1[type] : s.h0[type]);
            if (0 == (h.dp1 | h.dp12 | h.dp2 | h.dp3 | h.dpM | h.dpm | h.dp21)) { continue; }
            k4_touch(s, strand, s.hot[type]);
            int32_t *fd = s.facc + (strand * UVC_NSYM + s.hot[type]) * UVCGPU_NUM_FAM_DEPTHS;
            fd[UVC_cDP1] += h.dp1; fd[UVC_cDP12] += h.dp12; fd[UVC_cDP2] += h.dp2; fd[UVC_cDP3] += h.dp3; fd[UVC_cDPM] += h.dpM; fd[UVC_cDPm] += h.dpm; fd[UVC_cDP21] += h.dp21;
        }
    }
    for (int strand = 0; strand < 2; strand++) {
        const int32_t *fd = s.facc + strand * (UVC_NSYM * UVCGPU_NUM_FAM_DEPTHS);
        for (int type = 0; type < 2; type++) {
            const int s0 = (type == 0 ? UVC_BASE_A : UVC_LINK_M), s1 = (type == 0 ? UVC_BASE_NN : UVC_LINK_NN);
            int32_t totDP = 0;
            for (int y = s0; y <= s1; y++) { if ((s.touched >> (strand * UVC_NSYM + y)) & 1u) { totDP += fd[y * UVCGPU_NUM_FAM_DEPTHS + UVC_cDP1]; } }
            for (int y = s0; y <= s1; y++) {
                if (!((s.touched >> (strand * UVC_NSYM + y)) & 1u)) { continue; }
                int32_t q, ad, bq;
                infer_max_qual(q, ad, bq, v, sscs_phred(par, s.ref, y) + s.tn_add, 4, s.bucket + (strand * UVC_NSYM + y) * UVC_NUM_BUCKETS, totDP);
                vq[y * UVCGPU_NUM_VQ_TAGS + 8 + 3 * strand] = q; vq[y * UVCGPU_NUM_VQ_TAGS + 9 + 3 * strand] = ad; vq[y * UVCGPU_NUM_VQ_TAGS + 10 + 3 * strand] = bq;
                int32_t *g = v.famdepth + ((strand * v.n_pos + gp) * UVC_NSYM + y) * UVCGPU_NUM_FAM_DEPTHS;
                for (int k = 0; k < UVCGPU_NUM_FAM_DEPTHS; k++) { g[k] = fd[y * UVCGPU_NUM_FAM_DEPTHS + k]; }
            }
        }
    }
}

UVC_HD void k4_position(const BatchView & v, int64_t gp, const Win & w) {
    K4State s;
    K4Arrays arr;
    k4_begin(s, arr, v, gp);
    const int32_t p = s.p;
    for (int64_t ri = w.ulo; ri < w.uhi; ri++) {
        if (ri < w.lo || ri >= w.hi) { continue; }
        const ReadFam q = v.rfam[ri];
        if (q.rend <= p || q.famprev_maxrend > p) { continue; }
        if ((q.flags & UVC_RF_DIRECT) && k4_loop1_lone_fragment(s, v, q, v.fcol[q.col_base + p])) { continue; }
        k4_loop1_read(s, v, q, fam_entry_of_read(v, q, p), ri - w.lo);
    }
    k4_loop2_listed(s, v, w);
    if (s.n_need2 > UVC_K4_LIST) {
        for (int64_t ri = w.ulo; ri < w.uhi; ri++) {
            if (ri < w.lo || ri >= w.hi) { continue; }
            const ReadFam q = v.rfam[ri];
            if (q.rend <= p) { continue; }
            k4_loop2_read(s, v, q);
        }
    }
    k4_end(s, v);
}

// ------------------------------------------------------------------------------------------------ K4c: one thread per (family, strand)
// Haplotype evidence of families (main.hpp:3434-3521): the string of high-quality mutated consensus symbols of a single-strand family, and
// the sub-string that is also a tier-2 family consensus. Needs the complete cDPM/cDPm counters, hence runs after K4.
UVC_HD void k4c_family_strand(const BatchView & v, int64_t i) {
    const FamRec & F = v.fams[i >> 1];
    const int strand = (int)(i & 1);
    if (0 == F.n_frags[strand]) { return; }
    const TileInfo & T = v.tiles[F.tile];
    const uvcgpu_params & par = v.par;
    const int64_t po = T.pos_off - T.ext_beg;
    enum { cDPM = 4, cDPm = 5 };
    const int32_t lo = F.lo[strand], hi = F.hi[strand];
    int32_t n_fq = 0, n_f2q = 0;
    for (int pass = 0; pass < 2; pass++) {
        int32_t out_fq = -1, out_f2q = -1;
        if (pass == 1) {
            if (n_fq <= 1 && n_f2q <= 1) { break; }
            if (n_fq > 1) {
                out_fq = rec_alloc(v, 4 + 2 * n_fq);
                if (out_fq >= 0) { int32_t *w = v.rec_buf + out_fq; w[0] = UVC_REC_HAP_FQ; w[1] = strand; w[2] = n_fq; w[3] = F.tile; out_fq += 4; }
            }
            if (n_f2q > 1) {
                out_f2q = rec_alloc(v, 4 + 2 * n_f2q);
                if (out_f2q >= 0) { int32_t *w = v.rec_buf + out_f2q; w[0] = UVC_REC_HAP_F2Q; w[1] = strand; w[2] = n_f2q; w[3] = F.tile; out_f2q += 4; }
            }
        }
        // A strand with a single fragment (every strand of non-UMI data) can only contribute where that fragment's column has a mutated
        // consensus: the scan visits the set bits of the fragment's chunk masks instead of every position.
        const bool by_mask = (F.direct_frag[strand] >= 0);
        const FragRec *Gd = (by_mask ? &v.frags[F.direct_frag[strand]] : NULL);
        const uint32_t *fm = (by_mask ? v.fmask + (Gd->col_off / UVC_COL_CHUNK) * 4 : NULL);
        const int32_t n_chunks = (by_mask ? (Gd->hi - Gd->lo + UVC_COL_CHUNK - 1) / UVC_COL_CHUNK : 0);
        int32_t chunk = 0;
        uint32_t todo = ((by_mask && n_chunks > 0) ? (fm[UVC_FM_MUT_LINK] | fm[UVC_FM_MUT_BASE_ANY]) : 0u);
        for (int32_t p = lo; p < hi; p++) {
            if (by_mask) {
                while (0 == todo && chunk + 1 < n_chunks) { chunk++; todo = (fm[chunk * 4 + UVC_FM_MUT_LINK] | fm[chunk * 4 + UVC_FM_MUT_BASE_ANY]); }
                if (0 == todo) { break; }
                p = Gd->lo + chunk * UVC_COL_CHUNK + __ctz_u32(todo);
                todo &= todo - 1;
                if (p < lo || p >= hi) { continue; }
            }
            const int64_t gp = po + p;
            const int ref = v.refsym[gp];
            const FamCol m = fam_entry(v, F, strand, p);
            const int32_t *fd = v.famdepth + ((strand * v.n_pos + gp) * UVC_NSYM) * UVCGPU_NUM_FAM_DEPTHS;
            for (int type = 1; type >= 0; type--) {
                const int a = m.a2[type]; const int32_t con_sumBQs = (int32_t)m.mmm_cc[type], tot_sumBQs = (int32_t)m.mmm_tot[type];
                if (0 == tot_sumBQs || !symbols_mutated(ref, a)) { continue; }
                bool high = (type == 1);
                if (!high) {
                    const int32_t con_nfrags = m.con_a2[type];
                    const int32_t tot_nfrags = m.tc1[type];
                    const int32_t avgBQ = ((0 == tot_nfrags) ? 1 : (con_sumBQs / tot_nfrags));
                    const int32_t majorcount = fd[a * UVCGPU_NUM_FAM_DEPTHS + cDPM], minorcount = fd[a * UVCGPU_NUM_FAM_DEPTHS + cDPm];
                    const double prior_weight = 1.0 / (minorcount + 1.0);
                    const double p2p = v.phred2prob_tab[tmin(tmax(avgBQ, 0), 127)];
                    const double realphred = -10 * log((minorcount + prior_weight) / (majorcount + minorcount + prior_weight / p2p)) / v.ln10;
                    const int32_t indep_frag_phred = (int32_t)round(((con_nfrags * 2) - tot_nfrags) * realphred);
                    const int32_t confam_qual = tmax(1, tmin(indep_frag_phred, (con_sumBQs * 2) - tot_sumBQs));
                    high = (confam_qual >= par.bias_thres_highBQ);
                }
                if (!high) { continue; }
                const int a1 = m.a1[type]; const int32_t cc1 = m.cc1[type], tc1 = m.tc1[type];
                const bool confam = (a == a1 && par.fam_thres_dup1add <= tc1 && (cc1 * 100 >= tc1 * par.fam_thres_dup1perc));
                if (pass == 0) { n_fq++; if (confam) { n_f2q++; } }
                else {
                    if (out_fq >= 0) { v.rec_buf[out_fq] = p; v.rec_buf[out_fq + 1] = a; out_fq += 2; }
                    if (confam && out_f2q >= 0) { v.rec_buf[out_f2q] = p; v.rec_buf[out_f2q + 1] = a; out_f2q += 2; }
                }
            }
        }
    }
}

} // namespace uvc

#endif
