/* uvcgpu.h - C ABI of the B200 pileup-and-score library (libuvcgpu.so).
 *
 * This is the drop-in boundary for the reference's in-process seam
 *     template<class T> int process_batch(std::string& uncompressed_vcf_string, ..., BatchArg& arg, const T& tkis)
 * (reference main.cpp:458-464, called per tier-3 region at main.cpp:1508-1521). The reference has no
 * plugin/FFI interface of its own (one statically linked executable, Makefile:33-38), so every entry
 * point below names the reference code it replaces. Plain pointers and sizes only; no exceptions cross
 * the boundary; every function returns 0 on success or a negative UVCGPU_E* code, with the message
 * available from uvcgpu_last_error(). All buffers are caller-owned. The library fails loudly
 * (UVCGPU_ENODEVICE) when no CUDA device is present: there is no CPU fallback.
 */
#ifndef UVCGPU_H_INCLUDED
#define UVCGPU_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UVCGPU_ABI_VERSION 4

enum uvcgpu_error {
    UVCGPU_OK = 0,
    UVCGPU_EINVAL = -1,      /* bad argument */
    UVCGPU_ENODEVICE = -2,   /* no usable CUDA device (never falls back to the CPU) */
    UVCGPU_ECUDA = -3,       /* a CUDA call failed */
    UVCGPU_ENOMEM = -4,
    UVCGPU_EUNSUPPORTED = -5 /* a reference code path this build does not cover (e.g. IonTorrent) */
};

/* AlignmentSymbol (reference main_conversion.hpp:316-334); VTI in the VCF carries these integers. */
enum uvcgpu_symbol {
    UVC_BASE_A = 0, UVC_BASE_C = 1, UVC_BASE_G = 2, UVC_BASE_T = 3, UVC_BASE_N = 4, UVC_BASE_NN = 5,
    UVC_LINK_M = 6, UVC_LINK_D3P = 7, UVC_LINK_D2 = 8, UVC_LINK_D1 = 9,
    UVC_LINK_I3P = 10, UVC_LINK_I2 = 11, UVC_LINK_I1 = 12, UVC_LINK_NN = 13,
    UVC_NUM_SYMBOLS = 14
};

/* Flat POD mirror of the reference's CommandLineArgs AFTER its data-driven inference
 * (CmdLineArgs.hpp:20-438; inference CmdLineArgs.cpp:36-136, 1003-1035). Field names follow the
 * reference's option names. Filled by uvcgpu_params_default() and then overridden by the caller. */
typedef struct uvcgpu_params {
    int32_t abi_version;
    /* inferred */
    int32_t inferred_sequencing_platform; /* 1 = Illumina/BGI, 2 = IonTorrent (unsupported) */
    int32_t central_readlen;
    int32_t inferred_maxMQ;
    int32_t is_tumor_vcf_provided;        /* IS_PROVIDED(vcf_tumor_fname): this run is the normal of a T/N pair */
    int32_t assay_type;                   /* 0 auto, 1 capture, 2 amplicon */
    int32_t molecule_tag;
    int32_t pair_end_merge;               /* 0 = yes */
    int32_t disable_duplex;
    int32_t should_output_all;
    uint32_t outvar_flag;
    /* read filter, grouping (grouping.cpp:347-415, 608-997) */
    int32_t kept_aln_min_aln_len, kept_aln_min_mapqual, kept_aln_min_isize, kept_aln_max_isize, kept_aln_is_zero_isize_discarded;
    int32_t min_altdp_thres;
    uint32_t dedup_flag;
    double dedup_center_mult;
    double dedup_amplicon_end2end_ratio;
    double dedup_amplicon_border_to_insert_cov_weak_avgDP_ratio, dedup_amplicon_border_to_insert_cov_strong_avgDP_ratio;
    double dedup_amplicon_border_to_insert_cov_weak_totDP_ratio, dedup_amplicon_border_to_insert_cov_strong_totDP_ratio;
    double dedup_amplicon_border_weak_minDP, dedup_amplicon_border_strong_minDP;
    int32_t assay_sequencing_BQ_max, assay_sequencing_BQ_inc;
    /* primers */
    int32_t primerlen, primerlen2;
    uint32_t primer_flag;
    int32_t tn_is_paired;
    int32_t bq_phred_added_misma, bq_phred_added_indel;
    /* bias thresholds (CmdLineArgs.hpp:137-198) */
    int32_t bias_thres_highBQ, bias_thres_highBAQ, bias_thres_aLPxT_add, bias_thres_aLPxT_perc;
    int32_t bias_thres_aLRP1t_minus, bias_thres_aLRP2t_minus, bias_thres_aLRB1t_minus, bias_thres_aLRB2t_minus;
    int32_t bias_thres_aLRP1t_avgmul_perc, bias_thres_aLRP2t_avgmul_perc, bias_thres_aLRB1t_avgmul_perc, bias_thres_aLRB2t_avgmul_perc;
    int32_t bias_thres_aLRP1Nt_avgmul_perc, bias_thres_aLRB1Nt_avgmul_perc;
    int32_t bias_thres_aLRI1T_perc, bias_thres_aLRI2T_perc, bias_thres_aLRI1t_perc, bias_thres_aLRI2t_perc;
    int32_t bias_thres_aLRI1NT_perc, bias_thres_aLRI1Nt_perc, bias_thres_aLRI1T_add, bias_thres_aLRI2T_add;
    int32_t bias_thres_PFBQ1, bias_thres_PFBQ2;
    int32_t bias_thres_interfering_indel, bias_thres_interfering_indel_BQ, bias_thres_BAQ1, bias_thres_BAQ2;
    int32_t bias_thres_strict_c2LRP0;
    /* families (CmdLineArgs.hpp:43-51, 235-273) */
    int32_t fam_thres_highBQ_snv, fam_thres_highBQ_indel, fam_thres_dup1add, fam_thres_dup1perc, fam_thres_dup2add, fam_thres_dup2perc, fam_thres_qseqlen;
    int32_t fam_thres_emperr_all_flat_snv, fam_thres_emperr_con_perc_snv, fam_thres_emperr_all_flat_indel, fam_thres_emperr_con_perc_indel;
    int32_t fam_phred_indel_inc_before_barcode_labeling;
    int32_t fam_phred_sscs_transition_CG_TA, fam_phred_sscs_transition_AT_GC, fam_phred_sscs_transversion_CG_AT, fam_phred_sscs_transversion_other;
    int32_t fam_phred_sscs_indel_open, fam_phred_sscs_indel_ext;
    uint32_t fam_flag;
    int32_t syserr_mut_region_n_bases;
    /* indels (CmdLineArgs.hpp:330-354) */
    int32_t indel_BQ_max, indel_str_repeatsize_max, indel_vntr_repeatsize_max;
    double indel_polymerase_size, indel_polymerase_slip_rate, indel_del_to_ins_err_ratio;
    int32_t indel_adj_tracklen_dist, indel_adj_indellen_perc, indel_nonSTR_phred_per_base, indel_str_phred_per_region, indel_filter_edge_dist;
    double powlaw_exponent;
    /* micro-adjustments (CmdLineArgs.hpp:364-408) */
    int32_t microadjust_xm, microadjust_cliplen, microadjust_delFAQmax, microadjust_nobias_pos_indel_maxlen;
    int32_t microadjust_near_clip_dist, microadjust_alignment_clip_min_len, microadjust_padded_deletion_flag;
    int32_t microadjust_median_readlen_thres, microadjust_BAQ_per_base_x1024;
    /* phasing */
    int32_t phasing_haplotype_max_count, phasing_haplotype_min_ad, phasing_haplotype_max_detail_cnt;
    int32_t tumor_vcf_fname_nonempty;     /* vcf_tumor_fname.size() > 0: true even for the default "." (QUIRK, main.hpp:2564, 2858) */
    int32_t reserved[15];
    /* ---- scoring (BcfFormat_symbol_calc_DPv / calc_qual / output_germline / append_vcf_record, main.hpp:4274-6272) ---- */
    double vqual, vfa1, vfa2;
    int32_t vdp1, vad1, vdp2, vad2, min_r_ad, min_a_ad;
    int32_t syserr_minABQ_pcr_snv, syserr_minABQ_pcr_indel, syserr_minABQ_cap_snv, syserr_minABQ_cap_indel;   /* after the platform inference */
    int32_t syserr_BQ_prior, syserr_BQ_sbratio_q_add, syserr_BQ_sbratio_q_max, syserr_BQ_xmratio_q_add, syserr_BQ_xmratio_q_max;
    int32_t syserr_BQ_bmratio_q_add, syserr_BQ_bmratio_q_max, syserr_BQ_strand_favor_mul, syserr_MQ_min, syserr_MQ_max;
    double syserr_MQ_NMR_expfrac, syserr_MQ_NMR_altfrac_coef, syserr_MQ_NMR_nonaltfrac_coef, syserr_MQ_NMR_pl_exponent, syserr_MQ_nonref_base;
    double powlaw_anyvar_base, powlaw_amplicon_allele_fraction_coef;
    int32_t penal4lowdep;
    uint32_t nobias_flag;
    double nobias_pos_indel_lenfrac_thres;
    int32_t nobias_pos_indel_str_track_len;
    int32_t bias_prior_DPadd_perc;
    double bias_priorfreq_pos, bias_priorfreq_indel_in_read_div, bias_priorfreq_indel_in_var_div2, bias_priorfreq_indel_in_str_div2, bias_priorfreq_var_in_str_div2;
    double bias_prior_var_DP_mul;
    int32_t bias_priorfreq_ipos_snv, bias_priorfreq_ipos_indel, bias_priorfreq_strand_snv_base, bias_priorfreq_strand_indel;
    double bias_FA_pseudocount_indel_in_read, bias_priorfreq_orientation_snv_base, bias_priorfreq_orientation_indel_base;
    int32_t bias_FA_powerlaw_noUMI_phred_inc_snv, bias_FA_powerlaw_noUMI_phred_inc_indel, bias_FA_powerlaw_withUMI_phred_inc_snv, bias_FA_powerlaw_withUMI_phred_inc_indel;
    int32_t bias_reduction_by_high_sequencingDP_min_n_totDepth, bias_reduction_by_high_sequencingDP_min_n_altDepth;
    double bias_thres_FTS_FA, bias_orientation_min_effective_allelefrac;
    int32_t bias_is_orientation_artifact_mixed_with_sequencing_error;
    int32_t fam_min_n_copies, fam_min_n_copies_DPxAD, fam_min_overseq_perc, fam_bias_overseq_perc, fam_tier3DP_bias_overseq_perc, fam_indel_nonUMI_phred_dec_per_fold_overseq;
    int32_t fam_phred_dscs_all, fam_phred_dscs_max, fam_phred_dscs_inc_max, fam_phred_pow_sscs_transversion_AT_TA_origin;
    double fam_phred_pow_sscs_snv_origin, fam_phred_pow_sscs_indel_origin, fam_phred_pow_dscs_all_origin;
    double germ_hetero_FA;
    int32_t germ_phred_hetero_snp, germ_phred_hetero_indel, germ_phred_homalt_snp, germ_phred_homalt_indel, germ_phred_het3al_snp, germ_phred_het3al_indel;
    int32_t tn_q_inc_max, tn_q_inc_max_sscs_CG_AT, tn_q_inc_max_sscs_other;
    double tn_syserr_norm_devqual;
    double indel_multiallele_samepos_penal, indel_multiallele_diffpos_penal, indel_multiallele_soma_penal_thres;
    double indel_tetraallele_germline_penal_value, indel_tetraallele_germline_penal_thres;
    int32_t indel_ins_penal_pseudocount;
    double contam_any_mul_frac, contam_t2n_mul_frac;
    double microadjust_bias_pos_indel_fold, microadjust_bias_pos_indel_misma_to_indel_ratio, microadjust_nobias_pos_indel_misma_to_indel_ratio;
    int32_t microadjust_nobias_pos_indel_bMQ, microadjust_nobias_pos_indel_perc;
    double microadjust_nobias_strand_all_fold, microadjust_refbias_indel_max, microadjust_counterbias_pos_odds_ratio, microadjust_counterbias_pos_fold_ratio;
    int32_t microadjust_fam_binom_qual_halving_thres, microadjust_ref_MQ_dec_max;
    int32_t microadjust_syserr_MQ_NMR_tn_syserr_no_penal_qual_min, microadjust_syserr_MQ_NMR_tn_syserr_no_penal_qual_max;
    int32_t microadjust_longfrag_sidelength_min, microadjust_longfrag_sidelength_max;
    double microadjust_longfrag_sidelength_zeroMQpenalty;
    int32_t microadjust_alignment_clip_min_count, microadjust_alignment_tracklen_min;
    double microadjust_alignment_clip_min_frac;
    int32_t microadjust_germline_mix_with_del_snv_penalty, microadjust_strand_orientation_absence_DP_fold, microadjust_orientation_absence_snv_penalty;
    int32_t microadjust_strand_absence_snv_penalty, microadjust_dedup_absence_indel_penalty;
    int32_t lib_wgs_min_avg_fraglen, lib_nonwgs_clip_penal_min_indelsize, lib_nonwgs_normal_max_rescued_MQ, lib_wgs_normal_max_rescued_MQ;
    double lib_nonwgs_ad_pseudocount, lib_nonwgs_normal_full_self_rescue_fa, lib_nonwgs_normal_min_self_rescue_fa_ratio, lib_nonwgs_normal_add_mul_ad;
    int32_t should_output_all_germline;
    /* Not a reference parameter. 0 (default): the position kernels that only feed the output (bias pileup, fragment and family consensus) run
     * on the positions whose counters can reach it - the tile's own positions and one before them, plus the span the MGVCF block lines read
     * ahead (main.cpp:666-667) - instead of on the whole extended range of the tile's reads; the VCF is the same. 1: every array holds what the
     * reference's does over the whole extended range (uvcgpu_dump_counters then agrees with the reference everywhere; used by the parity tests). */
    int32_t all_positions;
    int32_t reserved2[6];
} uvcgpu_params;

/* One tier-3 region (the reference's BedLine, iohts.hpp:14-35) plus the previous one, as process_batch receives
 * them through BatchArg (main.cpp:143-144). Reads of the region are records [read_begin, read_end) of the batch. */
typedef struct uvcgpu_tile {
    int32_t tid;
    int32_t beg_pos, end_pos;       /* BedLine.beg_pos / end_pos */
    uint32_t region_flag;           /* BedLine.region_flag (0x1 = END_TO_END) */
    int32_t prev_tid, prev_beg_pos, prev_end_pos;
    int32_t contig_len;             /* target_len of tid */
    int64_t read_begin, read_end;   /* slice of uvcgpu_reads_soa */
} uvcgpu_tile;

/* BAM records of all tiles of a batch, structure-of-arrays. For every tile the caller supplies what the reference
 * fetches with sam_itr_queryi(tid, beg-2000, end+2000) (grouping.cpp:664, 730), in file order, already decoded (the slices of
 * different tiles may overlap, and a slice may be a file-order superset of the window: records that end before it are dropped by the
 * read filter like the reference's OUT_OF_RANGE test, grouping.cpp:408-409):
 * the fields of bam1_core_t, the NM aux tag (or -1), 4-bit packed bases, base qualities, CIGAR words, NUL-terminated qname.
 * Offsets are in bytes (seq, qual, qname) or 32-bit words (cigar). */
typedef struct uvcgpu_reads_soa {
    int64_t n_reads;
    const int32_t *pos, *mpos, *isize, *mtid, *l_qseq, *n_cigar, *nm;
    const uint16_t *flag;
    const uint8_t *mapq;
    const uint64_t *seq_off, *qual_off, *cigar_off, *qname_off; /* n_reads + 1 entries each */
    const uint8_t *seq;      /* (l_qseq+1)/2 bytes per read */
    const uint8_t *qual;     /* l_qseq bytes per read */
    const uint32_t *cigar;   /* n_cigar words per read */
    const char *qname;       /* l_qname bytes per read incl. NUL */
} uvcgpu_reads_soa;

/* Per-position counter records, byte-compatible with the reference's structs so that a dump can be compared
 * with the oracle by memcmp: SegFormatPrepSet (main_conversion.hpp:541-605, 208 B), SegFormatThresSet (:614-643, 72 B),
 * SegFormatInfoSet (:645-691, 152 B), FamFormatInfoSet (:701-720, 72 B), RegionalTandemRepeat (common.hpp:150-160, 28 B). */
typedef struct uvcgpu_prep_set {
    int32_t a_dp, a_near_ins_dp, a_near_del_dp, a_near_RTR_ins_dp, a_near_RTR_del_dp;
    int32_t a_pcr_dp, a_umi_dp, a_snv_dp, a_dnv_dp, a_highBQ_dp;
    int32_t a_near_pcr_clip_dp, a_near_long_clip_dp, a_at_ins_dp, a_at_del_dp;
    int32_t a_XM1500, a_GO1500, a_GAPLEN, a_qlen;
    int64_t a_near_ins_pow2len, a_near_del_pow2len;
    int32_t a_near_ins_inv100len, a_near_del_inv100len;
    int64_t a_near_ins_l_pow2len, a_near_ins_r_pow2len, a_near_del_l_pow2len, a_near_del_r_pow2len;
    int64_t a_LI; int32_t a_LIDP;
    int64_t a_RI; int32_t a_RIDP;
    int32_t a_l_dist_sum, a_r_dist_sum, a_inslen_sum, a_dellen_sum;
    int64_t a_l_BAQ_sum, a_r_BAQ_sum, a_insBAQ_sum, a_delBAQ_sum;
} uvcgpu_prep_set;

typedef struct uvcgpu_thres_set {
    int32_t aLPxT, aRPxT;
    int32_t aLI1T, aLI2T, aRI1T, aRI2T, aLI1t, aLI2t, aRI1t, aRI2t;
    int32_t aLP1t, aLP2t, aRP1t, aRP2t;
    int32_t aLB1t, aLB2t, aRB1t, aRB2t;
} uvcgpu_thres_set;

typedef struct uvcgpu_seginfo_set {
    int32_t a2XM2, a2BM2, aPF1, aPF2, aBQ2, aMQs, aP1, aP2, aP3, aNC;
    int32_t aDPff, aDPfr, aDPrf, aDPrr;
    int32_t aLP1, aLP2, aLPL, aRP1, aRP2, aRPL;
    int32_t aLB1, aLB2; int64_t aLBL;
    int32_t aRB1, aRB2; int64_t aRBL;
    int32_t aLI1, aLI2, aRI1, aRI2, aRIf, aLIr;
    int64_t aLIT, aRIT;
} uvcgpu_seginfo_set;

typedef struct uvcgpu_faminfo_set {
    int32_t c2LP1, c2LP2, c2LPL, c2RP1, c2RP2, c2RPL, c2LP0, c2RP0;
    int32_t c2LB1, c2LB2; int64_t c2LBL;
    int32_t c2RB1, c2RB2; int64_t c2RBL;
    int32_t c2BQ2;
} uvcgpu_faminfo_set;

typedef struct uvcgpu_rtr {
    int32_t begpos, tracklen, unitlen, indelphred, anyTR_begpos, anyTR_tracklen, anyTR_unitlen;
} uvcgpu_rtr;

#define UVCGPU_NUM_FRAG_DEPTHS 3   /* FRAG_bDP, bTA, bTB (main_conversion.hpp:693-698) */
#define UVCGPU_NUM_FAM_DEPTHS 8    /* FAM_cDP1, cDP12, cDP2, cDP3, cDPM, cDPm, cDP21, cDPD (:722-733) */
#define UVCGPU_NUM_DUPLEX_DEPTHS 2 /* DUPLEX_dDP1, dDP2 (:736-740) */
#define UVCGPU_NUM_VQ_TAGS 27      /* VQFormatTagSet (:743-782) */

/* Sections of the per-position state that uvcgpu_dump_counters can return (element = one reference struct/array per
 * position of the tile's extended range [ext_beg, ext_end)). */
enum uvcgpu_section {
    UVCGPU_SEC_META = 0,       /* int64[16]: see uvcgpu_tile_meta below */
    UVCGPU_SEC_RTR = 1,        /* uvcgpu_rtr per position (after the threshold pass adjusted indelphred) */
    UVCGPU_SEC_BAQ = 2,        /* int64 per position */
    UVCGPU_SEC_BAQ2 = 3,
    UVCGPU_SEC_PREP = 4,       /* uvcgpu_prep_set */
    UVCGPU_SEC_THRES = 5,      /* uvcgpu_thres_set */
    UVCGPU_SEC_SEGINFO = 6,    /* uvcgpu_seginfo_set[14] */
    UVCGPU_SEC_FAMINFO = 7,    /* uvcgpu_faminfo_set[14] */
    UVCGPU_SEC_FRAGDEPTH0 = 8, /* int32[14][3], strand 0 */
    UVCGPU_SEC_FRAGDEPTH1 = 9,
    UVCGPU_SEC_FAMDEPTH0 = 10, /* int32[14][8], strand 0 */
    UVCGPU_SEC_FAMDEPTH1 = 11,
    UVCGPU_SEC_DUPLEX = 12,    /* int32[14][2] */
    UVCGPU_SEC_VQ = 13,        /* int32[14][27] */
    UVCGPU_SEC_FAMILIES = 14,  /* text: family grouping, same format as the oracle harness */
    UVCGPU_SEC_RTR_INITIAL = 15,
    UVCGPU_SEC_INDELMAPS = 16, /* text: indel identity maps, same format as the oracle harness */
    UVCGPU_SEC_HAPLINKS = 17,  /* text: haplotype links (updateHapMap), same format as the oracle harness */
    UVCGPU_SEC_VCF = 18,       /* text: the tile's VCF body lines (same as uvcgpu_tile_vcf) */
    UVCGPU_NUM_SECTIONS
};

/* Per-tile scalars computed on the way (indices into the META section). */
enum uvcgpu_tile_meta {
    UVCGPU_META_NUM_PASSED = 0, UVCGPU_META_NUM_PCRPASSED = 1, UVCGPU_META_BAM_BEG = 2, UVCGPU_META_BAM_END = 3,
    UVCGPU_META_RPOS_BEG = 4, UVCGPU_META_RPOS_END = 5, UVCGPU_META_EXT_BEG = 6, UVCGPU_META_EXT_END = 7,
    UVCGPU_META_NUM_FAMILIES = 8
};

typedef struct uvcgpu_ctx uvcgpu_ctx;
typedef int64_t uvcgpu_ticket;

/* Throughput/timing record of one submitted batch (device times from CUDA events on the context's stream). */
typedef struct uvcgpu_batch_stats {
    int64_t n_tiles, n_reads_in, n_reads_kept, n_positions, n_ext_positions, n_families, n_fragments;
    int64_t h2d_bytes, d2h_bytes, gpu_launches;
    double host_prep_ms, h2d_ms, kernel_ms, d2h_ms;
    double kernel_ms_by_stage[16];   /* 0-10: K0, K1, K2, K2e, KF, K3a, K3b, KM, K4a, K4, K4c; 11: K6 block-line inputs; 12: K5 candidate scoring;
                                        13: staging kernels (P0 read filter + family segmentation, P1 reference context) */
    int64_t n_vcf_records;           /* candidate records the scoring stage kept */
    double host_score_ms;            /* host part of the scoring stage (indel allele table, record ordering) */
    double reserved[6];              /* [0]: ms the submit call waited for the sizes of the batch (staging kernels), [1]: ms of the whole staging part of submit,
                                        [2]: ms of the sparse-record downloads, [3]: ms of the host-side sparse maps, [4]: ms of the indel allele table,
                                        [5]: ms of the scoring kernels and the downloads of their results */
    int64_t n_positions_pileup;      /* positions the bias pileup (K2) ran on: n_ext_positions with all_positions, else the needed ones */
    int64_t n_positions_consensus;   /* positions the fragment / family consensus kernels (K3b, K4) ran on */
} uvcgpu_batch_stats;

/* CommandLineArgs defaults (CmdLineArgs.hpp) with the Illumina inference applied (CmdLineArgs.cpp:127-134). */
void uvcgpu_params_default(uvcgpu_params *p);

/* Replaces the per-thread handle set-up of main() (main.cpp:1297-1319). device = CUDA ordinal. */
int uvcgpu_create(uvcgpu_ctx **out, int device, const uvcgpu_params *params);
void uvcgpu_destroy(uvcgpu_ctx *ctx);
const char *uvcgpu_last_error(const uvcgpu_ctx *ctx);

/* Replaces load_refstring/faidx_fetch_seq (main.cpp:54-70, 553): the contig's bases are uploaded once and stay in HBM.
 * bases = ASCII, any case; NULL means "reference not available" (all 'n', main.cpp:57-59). */
int uvcgpu_set_contig(uvcgpu_ctx *ctx, int32_t tid, const char *bases, int64_t len);

/* Drops the bases of a contig that no later tile needs (the uvc1 host walks the genome in order). */
int uvcgpu_unset_contig(uvcgpu_ctx *ctx, int32_t tid);

/* Number of host threads the context may use for staging (stage P0/P1 of uvcgpu_submit) and for VCF text (uvcgpu_tile_vcf); 0 = all cores.
 * Replaces the reference's -t for these two stages (main.cpp:1479: one OpenMP thread per tier-2 region). */
int uvcgpu_set_host_threads(uvcgpu_ctx *ctx, int32_t n_threads);

/* Name of the contig as it is printed in the CHROM column (bam_hdr->target_name[tid], main.cpp:1462). */
int uvcgpu_set_contig_name(uvcgpu_ctx *ctx, int32_t tid, const char *name);

/* Replaces the body of process_batch up to scoring for a batch of tiles (main.cpp:481-591:
 * grouping.cpp:608-997 read filter + family grouping, :459-567 BQ fix-ups, main.hpp:803-874 repeat context,
 * main.cpp:400-429 BAQ offsets, main.hpp:3665-3742 updateByRegion3Aln). Asynchronous: the call registers the batch and returns; a worker
 * thread of the context stages it (upload, staging kernels, pileup launches) in submission order while the caller goes on - typically to
 * finish the previous batch. Arguments are checked and `tiles` is copied before the call returns; the SoA buffers are borrowed until
 * uvcgpu_release. Whatever goes wrong during staging (an invalid tile, a contig that was not set, out of memory) is reported by
 * uvcgpu_collect of the ticket, with uvcgpu_last_error set. */
int uvcgpu_submit(uvcgpu_ctx *ctx, int32_t n_tiles, const uvcgpu_tile *tiles, const uvcgpu_reads_soa *reads, uvcgpu_ticket *ticket);

/* Same as uvcgpu_submit, but the records of the tiles come from several SoA buffers (the uvc1 host decodes the tiles of a batch on several
 * threads, each into its own buffer): tile k's [read_begin, read_end) indexes sources[tile_source[k]]. */
int uvcgpu_submit_multi(uvcgpu_ctx *ctx, int32_t n_tiles, const uvcgpu_tile *tiles, int32_t n_sources, const uvcgpu_reads_soa *sources,
                        const int32_t *tile_source, uvcgpu_ticket *ticket);

/* Waits for the batch; fills stats. */
int uvcgpu_collect(uvcgpu_ctx *ctx, uvcgpu_ticket ticket, uvcgpu_batch_stats *stats);

/* Replaces the scoring half of process_batch for a collected batch (main.cpp:593-1172: BcfFormat_symboltype_init, BcfFormat_symbol_init,
 * BcfFormat_symbol_calc_DPv, BcfFormat_symbol_sum_DPv, BcfFormat_symbol_calc_qual, output_germline, append_vcf_record keep decision):
 * the indel identity streams are aggregated on the host, candidates are scored on the device. Synchronous. stats may be NULL. */
int uvcgpu_score(uvcgpu_ctx *ctx, uvcgpu_ticket ticket, uvcgpu_batch_stats *stats);

/* The uncompressed VCF body of one tile, exactly the bytes process_batch appends to its output string (main.cpp:1184): MGVCF block lines,
 * additional-indel-candidate lines and variant records in the reference's order. Runs uvcgpu_score first if it has not run.
 * Returns the number of bytes in *needed; copies min(cap, needed) bytes into dst (dst may be NULL). */
int uvcgpu_tile_vcf(uvcgpu_ctx *ctx, uvcgpu_ticket ticket, int32_t tile_index, char *dst, size_t cap, size_t *needed);

/* The VCF bodies of ALL tiles of the batch, concatenated in tile order (what a caller appends to its output for the batch): one call instead of
 * one per tile. Same conventions as uvcgpu_tile_vcf. */
int uvcgpu_batch_vcf(uvcgpu_ctx *ctx, uvcgpu_ticket ticket, char *dst, size_t cap, size_t *needed);

/* Test hook: raw per-position arrays of one tile of a collected batch for bit-exact parity against the oracle.
 * Returns the number of bytes the section needs in *needed; copies min(cap, needed) bytes into dst (dst may be NULL). */
int uvcgpu_dump_counters(uvcgpu_ctx *ctx, uvcgpu_ticket ticket, int32_t tile_index, int32_t section, void *dst, size_t cap, size_t *needed);

/* Test hook: evaluates scoring functions of the DEVICE code at caller-given points, so that the reference's compile-time known-answer checks
 * (static_asserts of main_conversion.hpp:205-209, 251-254) can be run against the CUDA implementation itself. in holds n triples (x, a, b):
 * which = 0: calc_binom_10log10_likeratio<false,false>(x, a, b) (main_conversion.hpp:222-237); which = 1: prob2odds(odds2prob(x));
 * which = 2: odds2prob(prob2odds(x)); which = 3: logit2(a, b). out receives n doubles. */
int uvcgpu_selftest_math(uvcgpu_ctx *ctx, int32_t which, const double *in, int32_t n, double *out);

/* Frees the device/host state of a collected batch. */
int uvcgpu_release(uvcgpu_ctx *ctx, uvcgpu_ticket ticket);

int uvcgpu_device_count(void);

/* Creates the CUDA context of a device ahead of the first uvcgpu_create (driver and context initialisation take a few hundred milliseconds: a
 * host program calls this from a helper thread while it parses its inputs; the reference has no counterpart - its per-thread handles are opened
 * in main.cpp:1297-1319). Returns UVCGPU_OK or UVCGPU_ENODEVICE / UVCGPU_ECUDA. */
int uvcgpu_device_warmup(int device);

/* Number of host staging blocks that the library's background thread still has to page-lock (process-wide). Staging never waits for
 * page-locking: a batch that finds no page-locked block of the size it needs stages through pageable memory (slower copies) and the block is
 * provisioned for the next one. A benchmark polls this after its warm-up to know that the steady state has been reached. */
int uvcgpu_staging_backlog(void);
/* Bytes of host staging memory the library has page-locked so far (process-wide); constant once the cache has seen the caller's steady state. */
int64_t uvcgpu_staging_pinned_bytes(void);

/* Page-locks every array of a caller's SoA buffer (cudaHostRegister), so that uvcgpu_submit uploads the records straight from it instead of
 * copying them into the library's own page-locked staging first (which costs a core about 0.1 s per gigabyte and the same again in host memory
 * traffic). Worth it for buffers that are submitted more than once or stay alive for a while (page-locking costs about as much as one copy);
 * a buffer that is not registered works the same, through the staging copy. The arrays must not be reallocated while registered (a
 * uvchost_readbuf may not grow); unregister before freeing them. No reference counterpart (htslib hands out one bam1_t at a time).
 * Returns UVCGPU_OK, or UVCGPU_ECUDA if an array could not be page-locked (the buffer then simply takes the staging path). */
int uvcgpu_host_register_reads(const uvcgpu_reads_soa *reads);
int uvcgpu_host_unregister_reads(const uvcgpu_reads_soa *reads);

/* Where the host's time inside the library's driver calls went so far (process-wide, all contexts and threads): for each kind of call, in this
 * order - device allocation, free, memset, download enqueue, kernel launches, waits for events, upload enqueue - three numbers: total milliseconds, number
 * of calls, longest single call in milliseconds. Writes at most `cap` doubles to `out` (may be NULL) and returns how many there are (21).
 * Diagnostics only (no reference counterpart): several contexts per GPU meet at the driver's locks, and this shows it. */
int uvcgpu_host_call_stats(double *out, int32_t cap);

/* sizeof(uvcgpu_params) as the library was compiled, so that foreign-language bindings can verify their mirror of the struct. */
size_t uvcgpu_sizeof_params(void);

#ifdef __cplusplus
}
#endif
#endif
