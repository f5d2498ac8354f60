// batch.h - data model of one submitted batch of tiles.
//
// HostBatch owns the host-side staging vectors; BatchView is the flat pointer view that the kernels read
// (device pointers in the CUDA build, host pointers in the test-only emulation build).
//
// HBM layout: every per-position array is the concatenation of the tiles' extended ranges
// [ext_beg, ext_end) (ext_end = the reference's extended_exclu_end_pos + 1, main.cpp:529-530, 569), so one
// launch covers all tiles of the batch; per-position records are array-of-structs in the reference's own
// struct layout (include/uvcgpu.h), which makes a dump comparable with the oracle by memcmp and makes the
// write-out of a position a single contiguous burst.
#ifndef UVC_BATCH_H_INCLUDED
#define UVC_BATCH_H_INCLUDED

#include "../../include/uvcgpu.h"

#include <stdint.h>

#define UVC_NSYM 14
#define UVC_MAX_INSERT_SIZE 2000   // common.hpp:64
#define UVC_MAX_STR_N_BASES 100    // common.hpp:63
#define UVC_SQR_QUAL_DIV 32        // main_conversion.hpp:20
#define UVC_NUM_BUCKETS 16         // main_conversion.hpp:920
#define UVC_SLIP_MAXUNIT 8
#define UVC_SLIP_NMAX 1024

// BAM CIGAR operators
#define UVC_CMATCH 0
#define UVC_CINS 1
#define UVC_CDEL 2
#define UVC_CREF_SKIP 3
#define UVC_CSOFT_CLIP 4
#define UVC_CHARD_CLIP 5
#define UVC_CPAD 6
#define UVC_CEQUAL 7
#define UVC_CDIFF 8

struct TileInfo {
    int32_t tid, beg_pos, end_pos;
    uint32_t region_flag;
    int32_t prev_tid, prev_beg_pos, prev_end_pos;
    int32_t ext_beg, ext_end;        // per-position arrays cover [ext_beg, ext_end); refstring covers [ext_beg, ext_end - 1)
    int32_t rpos_inclu_beg, rpos_exclu_end;
    int32_t bam_inclu_beg, bam_exclu_end;
    int64_t pos_off;                 // offset of ext_beg in the concatenated per-position arrays
    int64_t read_off; int32_t n_reads;
    int64_t frag_off; int32_t n_frags;
    int64_t fam_off; int32_t n_fams;
    int64_t num_passed, num_pcrpassed;
    int32_t is_amplicon_inferred;    // !is_by_capture (main.cpp:510-511)
    int32_t max_read_span;           // max(rend - pos) over the tile's kept reads
    int32_t skipped;                 // process_batch returned -1 (main.cpp:520-523)
};

// Per kept read (BAM order inside a tile, i.e. sorted by pos).
struct ReadRec {
    int32_t pos, rend, mpos, isize;
    int32_t l_qseq, n_cigar, nm;
    uint16_t flag; uint8_t mapq; uint8_t strand;       // strand = bam_get_strand (common.hpp:90)
    uint32_t dflag;                                    // duplexflag of its family (grouping.cpp:932)
    int32_t tile;
    int32_t frag, fam;                                 // batch-global fragment / family index
    uint64_t seq_off, qual_off, cigar_off;             // into the packed blobs (bytes, bytes, words)
    // read is "simple" if its CIGAR is [S|H] M [S|H] with a single M/=/X run
    int32_t simple, m_qoff, cx_off;                    // cx_off: offset of its per-reference-base expansion (complex reads), -1 otherwise
    int32_t ev_off, n_ev;                              // indel events of this read
    int32_t fragprev_maxrend, famprev_maxrend;         // max rend over earlier reads of the same fragment / (family,strand); INT32_MIN if none
    int32_t fambothprev_maxrend;                       // same over both strands of the family
    int32_t raw;                                       // index of the record in the batch's raw record arrays (the caller's SoA slices)
};

// Derived per-read constants (kernel K0).
struct ReadDerived {
    int32_t xm1500, go1500, avg_gaplen, nge_cnt, ngo_cnt, clip_cnt, lclip, rclip;
    int32_t inslen_sum, dellen_sum, insbaq_sum, delbaq_sum;
    int32_t bm1500[5];
    int32_t micro_indel_penal, micro_nogap_penal;
    int32_t ibeg, iend;                                // amplicon primer window (main.hpp:1872-1875)
    // per-read terms of dealwith_segbias (main.hpp:1360-1595) that do not depend on the position: hoisted out of the per-base work
    int32_t baq_pos, baq_rend1, baq2_rend1;            // baq[pos], baq[rend - 1], baq2[rend - 1]
    int32_t xm_term, bm_term[5];                       // (x > 20 ? 100 * 400 / (x * x) : 100) for x = xm1500, bm1500[b]
};

// What the bias pileup (kernel K2) needs from a read, 64 bytes, written by K0: the per-read terms of dealwith_segbias (main.hpp:1360-1595) that do
// not depend on the position are evaluated once per read instead of once per (read, position), and a warp stages a chunk of these records
// with one bulk asynchronous copy. bits: UVC_PR_* | mapq << 16 | micro_nogap_penal << 24.
struct alignas(16) PileRec {
    int32_t pos, rend, frag_l, frag_r;             // frag_l = min(pos, mpos), frag_r = frag_l + |isize|
    int32_t baq_pos, baq_rend1, baq2_rend1; uint32_t bits;
    uint32_t seq_off, qual_off; int32_t cx_off; uint32_t terms_lo;    // cx_off: simple reads keep m_qoff here; terms_lo = xm_term | bm_term[0] << 7 | bm_term[1] << 14 | bm_term[2] << 21
    uint32_t terms_hi; int32_t ibeg, iend; int32_t l_qseq;           // terms_hi = bm_term[3] | bm_term[4] << 7
};
// What the prep-set walk (kernel K1) needs from a read beyond PileRec, 32 bytes, written by K0
struct alignas(16) PrepRec { int32_t xm1500, go1500, avg_gaplen, inslen_sum, dellen_sum, insbaq_sum, delbaq_sum; uint32_t flags; };   // flags bit 0: duplexflag & 0x4
#define UVC_PR_ISRC 0x1u            // flag & 0x10
#define UVC_PR_PAIRED 0x2u          // flag & 0x1
#define UVC_PR_MATE_UNMAPPED 0x4u   // flag & 0x8
#define UVC_PR_STRAND 0x8u          // bam_get_strand
#define UVC_PR_HAS_ISIZE 0x10u      // isize != 0
#define UVC_PR_AMPLICON 0x20u       // is_assay_amplicon (main.hpp:1385-1387)
#define UVC_PR_UMI 0x40u            // duplexflag & 0x1
#define UVC_PR_NOCLIP 0x80u         // no soft/hard clip
#define UVC_PR_SIMPLE 0x100u        // CIGAR is [S|H] M [S|H]
#define UVC_PR_HAS_GAPS 0x200u      // at least one inserted / deleted base
#define UVC_PR_MASK_ON 0x400u       // primers of this read are masked outside [ibeg, iend)
#define UVC_PR_IS_NORMAL 0x800u     // isize != 0 or not paired (main.hpp:1531)
#define UVC_PR_MATE_OK 0x1000u      // mate mapped or not paired

// Per-reference-base expansion entry of a complex read: what the read shows at reference offset o = p - pos.
struct CxEntry {
    int32_t prev_rpos, next_rpos; // aligned bases: neighbouring entries of the low-quality-indel list (main.hpp:1902-1903);
                                  // deleted bases: the distance passed for the BASE_NN / LINK_NN padding updates (main.hpp:2245)
    int16_t qpos;      // query index of the aligned base, -1 if the reference base is deleted/skipped
    uint8_t flags;     // bit0: M base; bit1: M base that is not the first of its run (i2 > 0); bit2: deleted base (D op)
    uint8_t pad;
};

// One I or D cigar operation of a read, evaluated once (kernel K2e) and reused by all later walks.
struct IndelEvent {
    int32_t read, rpos, oplen, qpos;
    int32_t is_del;
    int32_t symbol;        // LINK_I1/I2/I3P/D1/D2/D3P
    int32_t incvalue;      // value passed to inc()/dealwith_segbias (already MAX(1, incvalue))
    int32_t incvalue2;     // count added to the inserted-sequence map (MAX(1, incvalue2))
    int32_t counted;       // nbases2end >= indel_filter_edge_dist and not primer-masked
    int32_t cigar_idx;
    int32_t tile, raw;     // tile of the read; index of the read in the batch's raw record arrays (the host reads inserted bases from the caller's SoA)
};

struct FragRec {
    int32_t tile, fam, strand;
    int32_t read_off, n_reads;     // into frag_reads
    int32_t beg, end;              // fillTidBegEndFromAlns1 (main.hpp:659-673)
    int32_t n_cov, n_near_mut;     // bTA / bTB numerators (main.hpp:2747-2756), kernel K3a
    int32_t normMQ;
    int32_t lo, hi;                // covered extent: min pos, max rend over the fragment's reads
    int64_t col_off;               // first entry of the fragment's column in fcol (multiple of 32)
};

// What one fragment (R1+R2 max-merged) asserts at one reference position: the result of the reference's per-fragment walks #3, #4 and #5
// (updateByAln<BASE_QUALITY_MAX>, main.hpp:2629, 2887, 3403; they rebuild the same array three times) reduced by fillConsensusCounts
// (main.hpp:374-417) for both symbol types. Computed once per (fragment, position) by kernel KF and read by every fragment/family kernel.
struct FragCol {
    uint16_t link_cc;              // link consensus with the reference counted once: its count (0 = the fragment says nothing here)
    uint16_t base_cc, base_tc;     // base consensus over A..NN: largest count and sum of counts
    uint8_t link_sym;              // low 4 bits: symbol; bit 7: the fragment votes for BASE_N or BASE_NN here (consensus over A..T may differ)
    uint8_t base_sym;
};

// Per (family, strand, position): the fragment votes of the family after the base-quality filter (updateByFiltering, main.hpp:466-495) and the
// major-minus-minor quality sums (updateByMajorMinusMinor, main.hpp:497-520), reduced to what the family loops consume. Index [0] = base, [1] = link.
struct FamCol {
    uint8_t a1[2], a2[2];          // consensus symbol over the fragment counts / over the quality sums
    uint16_t cc1[2], tc1[2];       // count of a1 and total count
    uint16_t con_a2[2];            // fragment count of a2
    uint32_t mmm_cc[2], mmm_tot[2];
};

// per-chunk bit masks of the fragment columns: what the whole-fragment and whole-family scans (K3a, K4c) look for, so that they read 16 bytes per
// 32 positions instead of 32 entries and 32 reference symbols
#define UVC_FM_COV 0               // the fragment says something here (link_cc | base_tc)
#define UVC_FM_MUT_LINK 1          // link consensus present and mutated w.r.t. the reference
#define UVC_FM_MUT_BASE_HQ 2       // base consensus mutated and 2 * cc - tc >= bias_thres_highBQ
#define UVC_FM_MUT_BASE_ANY 3      // base consensus mutated and 2 * cc - tc > 0
#define UVC_COL_CHUNK 32           // column entries are laid out in chunks of one warp; a chunk belongs to one fragment / (family, strand)

struct FamRec {
    int32_t tile;
    uint32_t duplexflag, dedup_idflag;
    int32_t frag_off[2], n_frags[2];   // fragments of each strand, contiguous in frag order
    int32_t beg2[2], end2[2];          // fillTidBegEndFromAlns2 per strand (main.hpp:675-685)
    int32_t beg_both, end_both;        // over both strands (main.hpp:3378-3381)
    int32_t l2r_end_median[2], r2l_end_median[2];
    int32_t qlen_ok[2];                // alns2.size() >= fam_thres_dup1add && qseqlen_sum >= n_qseqs * fam_thres_qseqlen
    int32_t nsb_min[2], nsb_max[2];    // no_strict_bias_pos_min/max (main.hpp:2959-2998), kernel K4a
    int32_t beg_tid, beg_pos, end_tid, end_pos;
    int32_t lo[2], hi[2];              // covered extent of each strand: min pos, max rend over its reads (lo >= hi: no reads)
    int64_t col_off[2];                // first entry of each strand's column in mcol (multiple of 32); unused when direct_frag >= 0
    int32_t direct_frag[2];            // a strand with a single fragment has no column of its own: its entries are derived on the fly
                                       // from that fragment's column (8 B instead of 32 B per position); -1 = materialised in mcol
};

// What the position kernels of the family stage need from a read, 32 B (instead of chasing ReadRec -> FamRec): built on the host.
struct ReadFam {
    int32_t rend, famprev_maxrend, fambothprev_maxrend, fam;
    int64_t col_base;                  // index of the (family, strand) entry at position p is col_base + p (mcol, or fcol when UVC_RF_DIRECT)
    uint32_t flags;                    // UVC_RF_*
    int32_t pad;
};
// What the position kernel of the fragment stage (K3b) needs from a read, 32 B (instead of chasing ReadRec -> FragRec): written by K3a, which
// computes the whole-fragment statistics.
struct ReadFrag {
    int32_t rend, fragprev_maxrend;
    int64_t col_base;                  // the fragment's entry at position p is fcol[col_base + p]
    int32_t n_cov, n_near_mut;         // bTA / bTB numerators of the fragment
    int32_t mq_term;                   // normMQ^2 / UVC_SQR_QUAL_DIV
    int32_t frag_strand;               // fragment index * 2 + strand
};
#define UVC_RF_STRAND 1u
#define UVC_RF_DIRECT 2u               // col_base points into fcol (single-fragment family-strand)
#define UVC_RF_DUPLEX_UMI 4u           // family has a duplex UMI (duplexflag & 0x2)
#define UVC_RF_BOTH_STRANDS 8u         // family has fragments on both strands

struct BatchView {
    uvcgpu_params par;
    double center_pow[4];          // pow(dedup_center_mult, d) computed with the host libm
    int32_t indelphred_half;       // (int)round(numstates2phred(indel_del_to_ins_err_ratio)) / 2 (main.hpp:1244)
    double ten_over_ln10;          // 10.0 / log(10.0) as the host libm evaluates it
    double ln10;                   // log(10)
    const double *phred2prob_tab;  // [128] phred2prob(q) = pow(10, -((float)q) / 10) (main_conversion.hpp:885-888) evaluated with the host libm
    const int32_t *slip_tab;       // [2][UVC_SLIP_MAXUNIT][UVC_SLIP_NMAX] indel_phred (main.hpp:794-801) evaluated with the host libm
    const int32_t *pf_tab;         // [2][128] passing-filter weight of a base quality: bq < PFBQk ? 100 * bq^2 / PFBQk^2 : 100 (main.hpp:1452-1455)
    int32_t n_tiles;
    int64_t n_pos, n_reads, n_frags, n_fams, n_cx, n_ev;
    const TileInfo *tiles;
    const int32_t *pos_tile;       // tile of each concatenated position
    // Positions the output-only position kernels run on (par.all_positions = 0): list k (0: bias pileup K2, 1: fragment / family consensus K3b, K4)
    // is the concatenation of one run of consecutive positions per tile, each padded to a multiple of 32 (a warp never straddles two tiles).
    // NULL list_tile = every position.
    int64_t n_list[2];
    const int32_t *list_tile[2];   // [n_list / 32] tile of each chunk of 32 list entries
    const int64_t *list_off[2];    // [n_tiles] first list entry of each tile
    // reference context
    const uint8_t *refsym;         // AlignmentSymbol of the reference base (CHAR_TO_SYMBOL, main_conversion.hpp:473-486)
    uvcgpu_rtr *rtr;
    const int32_t *baq, *baq2;     // BAQ prefix sums (main.cpp:400-429); values fit 32 bits
    int32_t *noindel;              // [n_pos] min(indelphred of the position before, indelphred of the position) after K1's adjustment: the
                                   // quality of "no indel at the junction before this base" (main.hpp:1918-1924), one coalesced word instead of two 28-byte records
    // reads
    const ReadRec *reads;
    ReadDerived *rd;
    PileRec *prec;                 // [n_reads], kernel K0
    PrepRec *qrec;                 // [n_reads], kernel K0
    const uint8_t *seq;
    uint8_t *qual;                 // per kept read: its base qualities after the reference's quality fix-ups (grouping.cpp:459-543), written by K0
    const uint8_t *qual_raw;       // base qualities as uploaded (a record that two tiles keep is shared here)
    const uint64_t *raw_qual_off;  // offset of raw record i in qual_raw
    const uint32_t *cigar;
    CxEntry *cx;
    IndelEvent *ev;
    const FragRec *frags_in; FragRec *frags;
    const int32_t *frag_reads;
    FamRec *fams;
    const ReadFam *rfam;           // [n_reads]
    ReadFrag *rfrag;               // [n_reads], kernel K3a
    // per-fragment and per-(family, strand) columns
    int64_t n_fcol, n_mcol;        // padded entry counts (multiples of UVC_COL_CHUNK)
    FragCol *fcol;
    FamCol *mcol;
    const int32_t *fchunk_frag;    // [n_fcol / 32] fragment that owns each chunk of fcol
    uint32_t *fmask;               // [n_fcol / 32][4] per chunk of fcol, one bit per entry (kernel KF): UVC_FM_*
    const int32_t *mchunk_fs;      // [n_mcol / 32] 2 * family + strand that owns each chunk of mcol
    // per-position state
    uvcgpu_prep_set *prep;
    uvcgpu_thres_set *thres;
    uvcgpu_seginfo_set *seginfo;   // [n_pos][14]
    int32_t *bqsum;                // [n_pos][14]  bg_seg_bqsum_conslogo (main.hpp:2566)
    int32_t *vq;                   // [n_pos][14][27]
    int32_t *fragdepth;            // [2][n_pos][14][3]
    int32_t *famdepth;             // [2][n_pos][14][8]
    uvcgpu_faminfo_set *faminfo;   // [n_pos][14]
    int32_t *duplex;               // [n_pos][14][2]
    // sparse outputs (indel identities, haplotype strings): a stream of int32 records appended with one atomic cursor
    int32_t *rec_buf; int32_t *rec_cursor; int32_t rec_cap;
};

// record stream entry kinds; fixed records are 6 words {kind, strand, symbol, pos, event index, count}
#define UVC_REC_FRAG_INDEL 1     // fragment-level indel identity   -> symbol_to_frag_format_depth_sets[strand] maps (main.hpp:2710-2717)
#define UVC_REC_FAM_INDEL 2      // family-level                    -> symbol_to_fam_format_depth_sets_2strand[strand] maps (main.hpp:3327-3336)
#define UVC_REC_CDP2_INDEL 3     // tier-2 consensus families       -> pos2iseq2data_cDP2 / pos2dlen2data_cDP2 (main.hpp:3197-3206)
#define UVC_REC_C2D_INDEL 4      // single-strand / duplex consensus -> pos2iseq2data_c2dDP / pos2dlen2data_c2dDP (main.hpp:3460-3469, 3536-3545)
// haplotype strings are variable-length: {kind, strand, n, tile} followed by n pairs {pos, symbol}
#define UVC_REC_HAP_BQ 10
#define UVC_REC_HAP_FQ 11
#define UVC_REC_HAP_F2Q 12

#endif
