// score_core.cuh - per-candidate scoring (SURVEY.md rows a-13 ... a-18) as per-work-item bodies.
//
// One thread handles one "zero-based position" of a tile exactly like one iteration of the reference's second hot loop
// (main.cpp:608-1172): the base-symbol group at refpos = zb - 1 and the link-symbol group at refpos = zb. For each group it gathers the
// candidate alleles' counters from the per-position arrays (BcfFormat_symboltype_init / BcfFormat_symbol_init, main.hpp:3889-4251),
// computes the bias-reduced allele fractions and filter flags (BcfFormat_symbol_calc_DPv, :4274-4844), the cross-candidate sums
// (BcfFormat_symbol_sum_DPv, :4888-4906), the variant qualities (BcfFormat_symbol_calc_qual, :4908-5343), the germline likelihoods
// (output_germline, :5483-5616) and the record-level qualities / keep decision (append_vcf_record, :6027-6263). Candidates that the
// reference would print are appended to an output array; the host turns them into VCF text.
//
// All floating point is double (float for QUAL, as in the reference). Covered configuration: tumor-only runs on the Illumina/BGI platform.
#ifndef UVC_SCORE_CORE_CUH_INCLUDED
#define UVC_SCORE_CORE_CUH_INCLUDED

#include "kernels_core.cuh"

#include <float.h>

#define UVC_MAX_GROUP_CANDS 12
#define UVC_NUM_FTS 19

// Upper-case FORMAT tags: one value set per (refpos, symbol type). [0] = sum over the symbols of the type, [1] = the padded-deletion symbol.
struct GroupFmt {
    int32_t APDP[12]; int64_t APXM[8]; int64_t APLRID[4]; int64_t APLRI[4]; int32_t APLRP[4];
    int32_t ALRPxT[2], ALRIT[4], ALRIt[4], ALRPt[4], ALRBt[4];
    int32_t AMQs[2], A1BQf[2], A1BQr[2], ADPff[2], ADPfr[2], ADPrf[2], ADPrr[2];
    int32_t ALP1[2], ALP2[2]; int64_t ALPL[2]; int32_t ARP1[2], ARP2[2]; int64_t ARPL[2];
    int32_t ALB2[2]; int64_t ALBL[2]; int32_t ARB2[2]; int64_t ARBL[2];
    int32_t ALI2[2], ALIr[2], ARI2[2], ARIf[2], ABQ2[2], APF2[2], AP1[2], AP2[2];
    int32_t BDPb[2], BTAb[2], BTBb[2], CDP1b[2], CDP1d[2], CDP12b[2], CDP2b[2], CDP3b[2], CDP21b[2], CDPMb[2], CDPmb[2], CDPDb[2];
    int32_t C2BQ2[2], C2LP0[2], C2RP0[2], C2LP2[2], C2RP2[2]; int64_t C2LPL[2], C2RPL[2];
    int32_t C2LB2[2], C2RB2[2]; int64_t C2LBL[2], C2RBL[2];
    int32_t DDP1[2], DDP2[2];
    int32_t CDP1v[2], CDP1w[2], CDP1x[2], CDP2v[2], CDP2w[2], CDP2x[2];
};

// Lower-case FORMAT tags of one candidate allele (the "R" tags are printed as [reference allele's value, this value]).
struct CandFmt {
    int32_t symbol, ev, gap_len;            // ev: representative indel event (identity) or -1
    int32_t aMQs, a1BQf, a1BQr, aP1, aP2, aDPff, aDPfr, aDPrf, aDPrr;
    int32_t aLP1, aLP2; int64_t aLPL; int32_t aRP1, aRP2; int64_t aRPL;
    int32_t aLB1, aLB2; int64_t aLBL; int32_t aRB1, aRB2; int64_t aRBL;
    int32_t a2XM2, a2BM2, aBQ2, aPF1, aPF2, aLI1, aLI2, aLIr, aRI1, aRI2, aRIf;
    int64_t aLIT, aRIT; int32_t aP3, aNC;
    int32_t bDPf, bTAf, bTBf, bDPr, bTAr, bTBr;
    int32_t cDP1f, cDP12f, cDP2f, cDP3f, cDP21f, cDPMf, cDPmf, cDPDf, cDP1r, cDP12r, cDP2r, cDP3r, cDP21r, cDPMr, cDPmr, cDPDr;
    int32_t c2LP1, c2LP2; int64_t c2LPL; int32_t c2RP1, c2RP2; int64_t c2RPL;
    int32_t c2LB1, c2LB2; int64_t c2LBL; int32_t c2RB1, c2RB2; int64_t c2RBL;
    int32_t c2BQ2, c2LP0, c2RP0, dDP1, dDP2;
    int32_t AD, bAD, c2AD;
    int32_t aBQ, a2BQf, a2BQr, aBQQ, bMQ, bIAQb, bIADb, bIDQb, cIAQf, cIADf, cIDQf, cIAQr, cIADr, cIDQr;
    int32_t bDPa, cDP0a;
    // BcfFormat_symbol_calc_DPv
    int32_t enable_tier2, nPF[2], nNFA[6], nAFA[9], nBCFA[10];
    uint32_t fts_mask; int32_t fts_pct[UVC_NUM_FTS];
    int32_t bNMa, bNMb, bNMQ, cDP1v, cDP1w, cDP1x, cDP2v, cDP2w, cDP2x;
    // BcfFormat_symbol_calc_qual
    int32_t cMmQ, aAaMQ, bMQQ, bIAQ, cIAQ, cPCQ1, cPLQ1, cPCQ2, cPLQ2, bTINQ, cTINQ, gVQ1, cVQ1, cVQ2, dVQinc, CONTQ;
};

// One record the reference would print (append_vcf_record): group values, the reference allele's and this allele's values, record-level numbers.
struct VarRec {
    int64_t gp;             // concatenated position index of refpos
    int32_t tile, refpos, symboltype, refsymbol, cand_index, pad0;
    GroupFmt g;
    CandFmt ref, alt;
    int32_t DP, bDP, c2DP;  // fmt.DP, fmt.bDP, fmt.c2DP
    int32_t cVQ1M[2], cVQ2M[2], cVQAM[2], cVQSM_ev[2];
    int32_t vHGQ, vAC[2], vNLODQ[2];
    int32_t tlodq, nlodq, somaticq, TNBQF[4], TNCQF[4];
    int32_t tbDP, tDP, tAD[2], t2DP, t2AD[2];
    float vcfqual, lowestVAQ;
    int32_t repeatnum, repeatunit_len, rtr_info[6];
};

// Per-position inputs of the MGVCF block lines (main.cpp:655-757): for the two symbol types in VCF order (link, base).
struct GvcfPos { int32_t bdepth[2], cdepth[2], cdep12[2], refQ[2]; };
// Per-position numbers of the ADDITIONAL_INDEL_CANDIDATE lines and the INFO RU/RC fields (main.cpp:608-614, 759-790).
struct GvcfExtra { int32_t tracklen, repeatnum, unitlen, a_dp, a_clip; };

// Candidate indel alleles of a (position, symbol), produced on the host from the indel identity maps exactly as indel_get_majority does.
struct IndelAllele { int64_t key; int32_t bAD, cAD, ev, len; };   // key = gp * 16 + symbol; several alleles per key in output order

// One (refpos, symbol type) group of a candidate position and the slots of its candidate alleles (kernels K5b ... K5f)
struct GroupRec {
    int64_t gp;                 // concatenated position index of refpos
    int32_t tile, refpos, type, refsymbol;
    int32_t first, n;           // candidates [first, first + n) in the reference's order
    int32_t partner;            // the other group of the same zero-based position, or -1
    int32_t ins_cdepth, del_cdepth, ins1_cdepth, del1_cdepth, repeatunit_len, repeatnum;
    GroupFmt g;
};
// what BcfFormat_symbol_init needs to know about a candidate besides its position
struct CandDesc { int32_t group, symbol, bDPa, cDP0a, ev, gap_len, minABQ, pad; };

// QUIRK: the reference runs BcfFormat_symbol_init + BcfFormat_symbol_calc_DPv for every allele of an indel symbol on ONE format object
// (main.cpp:850, 905-946), and calc_DPv appends to nNFA / nAFA / nBCFA and to the last FTS string without clearing them (main.hpp:4261-4268,
// 4737-4769): the record of the second allele of a symbol prints the first allele's values in front of its own. A kept record therefore comes
// with one of these per earlier allele of its symbol (rare: a site with two different insertions or deletions of the same length class).
struct PrevAllele { int32_t rec_slot, order; int32_t nNFA[6], nAFA[9], nBCFA[10]; uint32_t fts_mask; int32_t fts_pct[UVC_NUM_FTS]; };

struct ScoreView {
    const IndelAllele *alleles; int64_t n_alleles;
    VarRec *out; int32_t *out_cursor; int32_t out_cap;      // out_cursor[0..4]: records, candidate positions, groups, candidates, earlier-allele records
    PrevAllele *prev; int32_t prev_cap;
    GroupRec *groups; CandDesc *desc; CandFmt *cands; int32_t group_cap, cand_cap;
    GvcfPos *gvcf;          // [n_pos]
    GvcfExtra *gextra;      // [n_pos]
    int32_t *cand_list;     // [n_pos] zero-based positions that have at least one candidate allele (compacted by K5a; order is irrelevant)
    int32_t *cand_cursor;   // out_cursor[1]
};

namespace uvc {

UVC_HD double dmin(double a, double b) { return a < b ? a : b; }
UVC_HD double dmax(double a, double b) { return a > b ? a : b; }
// double -> integer conversion with x86-64 semantics (cvttsd2si): NaN and out-of-range values give the most negative integer
// ("integer indefinite"). The reference runs on x86-64 and several of its intermediate allele fractions are NaN or infinite for
// zero-depth alleles (0/0); its results therefore depend on this behaviour, while CUDA's cvt.rzi saturates (NaN -> 0, +inf -> INT_MAX).
UVC_HD int32_t d2i(double x) {
#if defined(__CUDA_ARCH__)
    return ((x > -2147483649.0 && x < 2147483648.0) ? (int32_t)x : INT32_MIN);
#else
    return (int32_t)x;
#endif
}
UVC_HD int64_t d2l(double x) {
#if defined(__CUDA_ARCH__)
    return ((x >= -9223372036854775808.0 && x < 9223372036854775808.0) ? (int64_t)x : INT64_MIN);
#else
    return (int64_t)x;
#endif
}
UVC_HD bool is_subst(int s) { return s >= UVC_BASE_A && s <= UVC_BASE_NN; }

// SYMBOL_TYPE_TO_SYMBOLS order (main_conversion.hpp:397-400)
UVC_HD int type_symbol(int type, int k) {
    if (type == 0) { return k; }
    const int link_order[8] = {UVC_LINK_M, UVC_LINK_I1, UVC_LINK_I2, UVC_LINK_I3P, UVC_LINK_D1, UVC_LINK_D2, UVC_LINK_D3P, UVC_LINK_NN};
    return link_order[k];
}
UVC_HD int type_nsym(int type) { return type == 0 ? 6 : 8; }

struct PosPtrs {
    const uvcgpu_prep_set *prep; const uvcgpu_thres_set *thres; const uvcgpu_seginfo_set *seg; const uvcgpu_faminfo_set *fam;
    const int32_t *vq, *fd0, *fd1, *fm0, *fm1, *dup;
};
UVC_HD PosPtrs pos_ptrs(const BatchView & v, int64_t gp) {
    PosPtrs P;
    P.prep = v.prep + gp; P.thres = v.thres + gp; P.seg = v.seginfo + gp * UVC_NSYM; P.fam = v.faminfo + gp * UVC_NSYM;
    P.vq = v.vq + gp * UVC_NSYM * UVCGPU_NUM_VQ_TAGS;
    P.fd0 = v.fragdepth + ((0 * v.n_pos + gp) * UVC_NSYM) * UVCGPU_NUM_FRAG_DEPTHS;
    P.fd1 = v.fragdepth + ((1 * v.n_pos + gp) * UVC_NSYM) * UVCGPU_NUM_FRAG_DEPTHS;
    P.fm0 = v.famdepth + ((0 * v.n_pos + gp) * UVC_NSYM) * UVCGPU_NUM_FAM_DEPTHS;
    P.fm1 = v.famdepth + ((1 * v.n_pos + gp) * UVC_NSYM) * UVCGPU_NUM_FAM_DEPTHS;
    P.dup = v.duplex + gp * UVC_NSYM * UVCGPU_NUM_DUPLEX_DEPTHS;
    return P;
}

// BcfFormat_symboltype_init (main.hpp:3889-4081)
UVC_HD void group_init(GroupFmt & g, const PosPtrs & P, int type) {
    const uvcgpu_prep_set & p = *P.prep;
    const uvcgpu_thres_set & t = *P.thres;
    const int s0 = (type == 0 ? UVC_BASE_A : UVC_LINK_M), s1 = (type == 0 ? UVC_BASE_NN : UVC_LINK_NN), nn = s1;
    g.APDP[0] = p.a_dp; g.APDP[1] = p.a_near_ins_dp; g.APDP[2] = p.a_near_del_dp; g.APDP[3] = p.a_near_RTR_ins_dp; g.APDP[4] = p.a_near_RTR_del_dp;
    g.APDP[5] = p.a_pcr_dp; g.APDP[6] = p.a_snv_dp; g.APDP[7] = p.a_dnv_dp; g.APDP[8] = p.a_highBQ_dp; g.APDP[9] = p.a_near_pcr_clip_dp;
    g.APDP[10] = p.a_near_long_clip_dp; g.APDP[11] = p.a_umi_dp;
    g.APXM[0] = p.a_XM1500; g.APXM[1] = p.a_GO1500; g.APXM[2] = p.a_qlen; g.APXM[3] = p.a_GAPLEN; g.APXM[4] = p.a_near_ins_pow2len;
    g.APXM[5] = p.a_near_del_pow2len; g.APXM[6] = p.a_near_ins_inv100len; g.APXM[7] = p.a_near_del_inv100len;
    g.APLRID[0] = p.a_near_ins_l_pow2len; g.APLRID[1] = p.a_near_ins_r_pow2len; g.APLRID[2] = p.a_near_del_l_pow2len; g.APLRID[3] = p.a_near_del_r_pow2len;
    g.APLRI[0] = p.a_LI; g.APLRI[1] = p.a_LIDP; g.APLRI[2] = p.a_RI; g.APLRI[3] = p.a_RIDP;
    g.APLRP[0] = p.a_l_dist_sum; g.APLRP[1] = p.a_r_dist_sum; g.APLRP[2] = p.a_inslen_sum; g.APLRP[3] = p.a_dellen_sum;
    g.ALRPxT[0] = t.aLPxT; g.ALRPxT[1] = t.aRPxT;
    g.ALRIT[0] = t.aLI1T; g.ALRIT[1] = t.aLI2T; g.ALRIT[2] = t.aRI1T; g.ALRIT[3] = t.aRI2T;
    g.ALRIt[0] = t.aLI1t; g.ALRIt[1] = t.aLI2t; g.ALRIt[2] = t.aRI1t; g.ALRIt[3] = t.aRI2t;
    g.ALRPt[0] = t.aLP1t; g.ALRPt[1] = t.aLP2t; g.ALRPt[2] = t.aRP1t; g.ALRPt[3] = t.aRP2t;
    g.ALRBt[0] = t.aLB1t; g.ALRBt[1] = t.aLB2t; g.ALRBt[2] = t.aRB1t; g.ALRBt[3] = t.aRB2t;
    #define UVC_SUMSEG(dst, field) { int64_t r_ = 0; for (int s = s0; s <= s1; s++) { r_ += (int64_t)P.seg[s].field; } dst[0] = r_; dst[1] = P.seg[nn].field; }
    #define UVC_SUMFAM(dst, field) { int64_t r_ = 0; for (int s = s0; s <= s1; s++) { r_ += (int64_t)P.fam[s].field; } dst[0] = r_; dst[1] = P.fam[nn].field; }
    { int32_t r = 0; for (int s = s0; s <= s1; s++) { r += P.vq[s * UVCGPU_NUM_VQ_TAGS + 0]; } g.A1BQf[0] = r; g.A1BQf[1] = P.vq[nn * UVCGPU_NUM_VQ_TAGS + 0]; }
    { int32_t r = 0; for (int s = s0; s <= s1; s++) { r += P.vq[s * UVCGPU_NUM_VQ_TAGS + 1]; } g.A1BQr[0] = r; g.A1BQr[1] = P.vq[nn * UVCGPU_NUM_VQ_TAGS + 1]; }
    UVC_SUMSEG(g.AMQs, aMQs) UVC_SUMSEG(g.AP1, aP1) UVC_SUMSEG(g.AP2, aP2)
    UVC_SUMSEG(g.ADPff, aDPff) UVC_SUMSEG(g.ADPfr, aDPfr) UVC_SUMSEG(g.ADPrf, aDPrf) UVC_SUMSEG(g.ADPrr, aDPrr)
    UVC_SUMSEG(g.ALP1, aLP1) UVC_SUMSEG(g.ALP2, aLP2) UVC_SUMSEG(g.ALPL, aLPL) UVC_SUMSEG(g.ARP1, aRP1) UVC_SUMSEG(g.ARP2, aRP2) UVC_SUMSEG(g.ARPL, aRPL)
    UVC_SUMSEG(g.ALB2, aLB2) UVC_SUMSEG(g.ALBL, aLBL) UVC_SUMSEG(g.ARB2, aRB2) UVC_SUMSEG(g.ARBL, aRBL)
    UVC_SUMSEG(g.ABQ2, aBQ2) UVC_SUMSEG(g.APF2, aPF2) UVC_SUMSEG(g.ALI2, aLI2) UVC_SUMSEG(g.ARIf, aRIf) UVC_SUMSEG(g.ARI2, aRI2) UVC_SUMSEG(g.ALIr, aLIr)
    #define UVC_FR(dst, arr0, arr1, stride, idx) { int32_t a_ = 0, b_ = 0; for (int s = s0; s <= s1; s++) { a_ += arr0[s * stride + idx]; b_ += arr1[s * stride + idx]; } dst[0] = a_; dst[1] = b_; }
    UVC_FR(g.BDPb, P.fd0, P.fd1, UVCGPU_NUM_FRAG_DEPTHS, 0) UVC_FR(g.BTAb, P.fd0, P.fd1, UVCGPU_NUM_FRAG_DEPTHS, 1) UVC_FR(g.BTBb, P.fd0, P.fd1, UVCGPU_NUM_FRAG_DEPTHS, 2)
    UVC_FR(g.CDP1b, P.fm0, P.fm1, UVCGPU_NUM_FAM_DEPTHS, 0) UVC_FR(g.CDP12b, P.fm0, P.fm1, UVCGPU_NUM_FAM_DEPTHS, 1) UVC_FR(g.CDP2b, P.fm0, P.fm1, UVCGPU_NUM_FAM_DEPTHS, 2)
    UVC_FR(g.CDP3b, P.fm0, P.fm1, UVCGPU_NUM_FAM_DEPTHS, 3) UVC_FR(g.CDPMb, P.fm0, P.fm1, UVCGPU_NUM_FAM_DEPTHS, 4) UVC_FR(g.CDPmb, P.fm0, P.fm1, UVCGPU_NUM_FAM_DEPTHS, 5)
    UVC_FR(g.CDP21b, P.fm0, P.fm1, UVCGPU_NUM_FAM_DEPTHS, 6) UVC_FR(g.CDPDb, P.fm0, P.fm1, UVCGPU_NUM_FAM_DEPTHS, 7)
    // QUIRK: fill_symboltype_nn_fmt reads strand 0 twice (main.hpp:3783-3784)
    g.CDP1d[0] = P.fm0[nn * UVCGPU_NUM_FAM_DEPTHS + 0]; g.CDP1d[1] = P.fm0[nn * UVCGPU_NUM_FAM_DEPTHS + 0];
    UVC_SUMFAM(g.C2LP2, c2LP2) UVC_SUMFAM(g.C2LPL, c2LPL) UVC_SUMFAM(g.C2RP2, c2RP2) UVC_SUMFAM(g.C2RPL, c2RPL)
    UVC_SUMFAM(g.C2LB2, c2LB2) UVC_SUMFAM(g.C2LBL, c2LBL) UVC_SUMFAM(g.C2RB2, c2RB2) UVC_SUMFAM(g.C2RBL, c2RBL)
    UVC_SUMFAM(g.C2BQ2, c2BQ2) UVC_SUMFAM(g.C2LP0, c2LP0) UVC_SUMFAM(g.C2RP0, c2RP0)
    { int32_t r = 0; for (int s = s0; s <= s1; s++) { r += P.dup[s * 2 + 0]; } g.DDP1[0] = r; g.DDP1[1] = P.dup[nn * 2 + 0]; }
    { int32_t r = 0; for (int s = s0; s <= s1; s++) { r += P.dup[s * 2 + 1]; } g.DDP2[0] = r; g.DDP2[1] = P.dup[nn * 2 + 1]; }
    #undef UVC_SUMSEG
    #undef UVC_SUMFAM
    #undef UVC_FR
    for (int k = 0; k < 2; k++) { g.CDP1v[k] = g.CDP1w[k] = g.CDP1x[k] = g.CDP2v[k] = g.CDP2w[k] = g.CDP2x[k] = 0; }
}

// BcfFormat_symbol_init + fill_symbol_VQ_fmts (main.hpp:4095-4251, 3820-3887)
UVC_HD void cand_init(CandFmt & c, const GroupFmt & g, const BatchView & v, const PosPtrs & P, int symbol, int32_t bDPa, int32_t cDP0a, int32_t ev, int32_t gap_len, int32_t minABQ) {
    const uvcgpu_params & par = v.par;
    const uvcgpu_seginfo_set & s = P.seg[symbol];
    const uvcgpu_faminfo_set & f = P.fam[symbol];
    const int32_t *vq = P.vq + symbol * UVCGPU_NUM_VQ_TAGS;
    c.symbol = symbol; c.ev = ev; c.gap_len = gap_len;
    c.a1BQf = vq[0]; c.a1BQr = vq[1]; c.aMQs = s.aMQs; c.aP1 = s.aP1; c.aP2 = s.aP2;
    c.aDPff = s.aDPff; c.aDPfr = s.aDPfr; c.aDPrf = s.aDPrf; c.aDPrr = s.aDPrr;
    c.aLP1 = s.aLP1; c.aLP2 = s.aLP2; c.aLPL = s.aLPL; c.aRP1 = s.aRP1; c.aRP2 = s.aRP2; c.aRPL = s.aRPL;
    c.aLB1 = s.aLB1; c.aLB2 = s.aLB2; c.aLBL = s.aLBL; c.aRB1 = s.aRB1; c.aRB2 = s.aRB2; c.aRBL = s.aRBL;
    c.a2XM2 = s.a2XM2; c.a2BM2 = s.a2BM2; c.aBQ2 = s.aBQ2; c.aPF1 = s.aPF1; c.aPF2 = s.aPF2;
    c.aLI1 = s.aLI1; c.aLI2 = s.aLI2; c.aLIr = s.aLIr; c.aRI1 = s.aRI1; c.aRI2 = s.aRI2; c.aRIf = s.aRIf;
    c.aLIT = s.aLIT; c.aRIT = s.aRIT; c.aP3 = s.aP3; c.aNC = s.aNC;
    c.bDPf = P.fd0[symbol * 3 + 0]; c.bTAf = P.fd0[symbol * 3 + 1]; c.bTBf = P.fd0[symbol * 3 + 2];
    c.bDPr = P.fd1[symbol * 3 + 0]; c.bTAr = P.fd1[symbol * 3 + 1]; c.bTBr = P.fd1[symbol * 3 + 2];
    const int32_t *m0 = P.fm0 + symbol * UVCGPU_NUM_FAM_DEPTHS, *m1 = P.fm1 + symbol * UVCGPU_NUM_FAM_DEPTHS;
    c.cDP1f = m0[0]; c.cDP12f = m0[1]; c.cDP2f = m0[2]; c.cDP3f = m0[3]; c.cDPMf = m0[4]; c.cDPmf = m0[5]; c.cDP21f = m0[6]; c.cDPDf = m0[7];
    c.cDP1r = m1[0]; c.cDP12r = m1[1]; c.cDP2r = m1[2]; c.cDP3r = m1[3]; c.cDPMr = m1[4]; c.cDPmr = m1[5]; c.cDP21r = m1[6]; c.cDPDr = m1[7];
    c.c2LP1 = f.c2LP1; c.c2LP2 = f.c2LP2; c.c2LPL = f.c2LPL; c.c2RP1 = f.c2RP1; c.c2RP2 = f.c2RP2; c.c2RPL = f.c2RPL;
    c.c2LB1 = f.c2LB1; c.c2LB2 = f.c2LB2; c.c2LBL = f.c2LBL; c.c2RB1 = f.c2RB1; c.c2RB2 = f.c2RB2; c.c2RBL = f.c2RBL;
    c.c2BQ2 = f.c2BQ2; c.c2LP0 = f.c2LP0; c.c2RP0 = f.c2RP0;
    c.dDP1 = P.dup[symbol * 2 + 0]; c.dDP2 = P.dup[symbol * 2 + 1];
    c.AD = c.cDP1f + c.cDP1r; c.bAD = c.bDPf + c.bDPr; c.c2AD = c.cDP2f + c.cDP2r;
    // fill_symbol_VQ_fmts
    const int32_t a2BQf = vq[2], a2BQr = vq[3];
    const int32_t aDPf = c.aDPff + c.aDPrf, aDPr = c.aDPfr + c.aDPrr;
    const int32_t ADP = g.ADPff[0] + g.ADPrf[0] + g.ADPfr[0] + g.ADPrr[0];
    const int32_t rssDPfBQ = (int32_t)(aDPf * sqrt((double)(((int64_t)a2BQf * UVC_SQR_QUAL_DIV) / tmax(1, aDPf))));
    const int32_t rssDPrBQ = (int32_t)(aDPr * sqrt((double)(((int64_t)a2BQr * UVC_SQR_QUAL_DIV) / tmax(1, aDPr))));
    const int32_t rssDPbBQ = (int32_t)((aDPf + aDPr) * sqrt((double)((a2BQf + a2BQr) * UVC_SQR_QUAL_DIV / tmax(1, aDPf + aDPr))));
    const double excess = dmax(0.0, ((aDPf + aDPr + 0.5) * 2.0 / (ADP + 1.0) - 1.0));
    int32_t minABQa = minABQ - d2i((5 * 10.0 * (excess * excess)));
    const double sbratio = (double)(tmax(aDPf, aDPr) * 10 + 10) / (double)(tmin(aDPf, aDPr) * 10 + 10);
    minABQa += between(d2i((sbratio * sbratio)) - par.syserr_BQ_sbratio_q_add, 0, par.syserr_BQ_sbratio_q_max);
    const int32_t xmratio = (par.syserr_BQ_xmratio_q_max * 10 * (aDPf + aDPr) / tmax(1, c.a2XM2));
    const int32_t bmratio = (par.syserr_BQ_bmratio_q_max * 10 * (aDPf + aDPr) / tmax(1, c.a2BM2));
    minABQa += between(xmratio - par.syserr_BQ_xmratio_q_add, 0, par.syserr_BQ_xmratio_q_max) + between(bmratio - par.syserr_BQ_bmratio_q_add, 0, par.syserr_BQ_bmratio_q_max);
    const int32_t m = par.syserr_BQ_strand_favor_mul;
    const int32_t q_fw = (rssDPfBQ * m - minABQa * aDPf * m / 10 + rssDPrBQ - minABQa * aDPr / 10) / m;
    const int32_t q_rv = (rssDPrBQ * m - minABQa * aDPr * m / 10 + rssDPfBQ - minABQa * aDPf / 10) / m;
    const int32_t q_2d = (rssDPbBQ) - minABQa * (aDPf + aDPr) / 10;
    const int32_t a_rmsBQ = (rssDPbBQ) / tmax(1, aDPf + aDPr);
    c.bMQ = (int32_t)round(sqrt((double)(((int64_t)vq[4] * UVC_SQR_QUAL_DIV) / tmax(c.bDPf + c.bDPr, 1))) + (double)(1.0 - FLT_EPSILON));
    c.aBQQ = tmax(a_rmsBQ, par.syserr_BQ_prior + tmax(q_2d, tmax(q_fw, q_rv)));
    c.a2BQf = rssDPfBQ; c.a2BQr = rssDPrBQ; c.aBQ = a_rmsBQ;
    c.bIAQb = vq[5]; c.bIADb = vq[6]; c.bIDQb = vq[7]; c.cIAQf = vq[8]; c.cIADf = vq[9]; c.cIDQf = vq[10]; c.cIAQr = vq[11]; c.cIADr = vq[12]; c.cIDQr = vq[13];
    c.bDPa = bDPa; c.cDP0a = cDP0a;
    c.enable_tier2 = 0; c.fts_mask = 0;
    for (int k = 0; k < UVC_NUM_FTS; k++) { c.fts_pct[k] = 0; }
}

// dp4_to_pcFA (main_conversion.hpp:798-849); returns {bias-reduced FA, no-bias FA}
template <bool TBidirectional, bool TIsOverseqFracDisabled>
UVC_HD void dp4_to_pcFA(double out[2], double overseq_frac, double aADpass, double aADfail, double aDPpass, double aDPfail, double pl_exponent, double n_nats,
        double aADavgKeyVal, double aDPavgKeyVal, double priorAD, double priorDP) {
    if (!TIsOverseqFracDisabled) { aDPfail *= overseq_frac; aDPpass *= overseq_frac; aADfail *= overseq_frac; aADpass *= overseq_frac; }
    aDPfail += priorDP; aDPpass += priorDP; aADfail += priorAD; aADpass += priorAD;
    const double nobiasFA = (aADfail + aADpass) / (aDPfail + aDPpass);
    if ((aADpass / aDPpass) >= (aADfail / aDPfail)) {
        if (TBidirectional) {
            double t = aDPfail; aDPfail = aDPpass; aDPpass = t;
            t = aADfail; aADfail = aADpass; aADpass = t;
        } else {
            out[0] = (aADpass / aDPpass); out[1] = nobiasFA; return;
        }
    }
    const double aBDfail = aDPfail * 2 - aADfail * 1;
    const double aBDpass = aDPpass * 2 - aADpass * 1;
    double aADpassfrac = aADpass / (aADpass + aADfail);
    double aBDpassfrac = aBDpass / (aBDpass + aBDfail);
    if ((!TBidirectional) && (aADavgKeyVal >= 0) && (aDPavgKeyVal >= 0)) {
        aADpassfrac = aADavgKeyVal / (aADavgKeyVal + aDPavgKeyVal * 0.9);
        aBDpassfrac = 1.0 - aADpassfrac;
    }
    double infogain = aADfail * log((1.0 - aADpassfrac) / (1.0 - aBDpassfrac));
    if (TBidirectional) { infogain += aADpass * log(aADpassfrac / aBDpassfrac); }
    if (infogain <= n_nats) { out[0] = aADfail / aDPfail; out[1] = nobiasFA; }
    else { out[0] = dmax(aADpass / aDPpass, (aADfail / aDPfail) * exp((n_nats - infogain) / pl_exponent)); out[1] = nobiasFA; }
}

UVC_HD double prob2odds(double p) { return p / (1.0 - p); }
UVC_HD double logit2(double a, double b) { return log(prob2odds((a + DBL_EPSILON) / (a + b + 2.0 * DBL_EPSILON))); }
UVC_HD double phred2nat(const BatchView & v, double x) { return (v.ln10 / 10.0) * x; }
UVC_HD double numstates2phred(const BatchView & v, double x) { return v.ten_over_ln10 * log(x); }
UVC_HD int32_t numstates2deciphred(const BatchView & v, double x) { return d2i(round((100.0 / v.ln10) * log(x))); }

// calc_binom_10log10_likeratio<false,false> (main_conversion.hpp:222-237)
UVC_HD double binom_10log10_likeratio(const BatchView & v, double prob, double a, double b) {
    prob = (prob + DBL_EPSILON) / (1.0 + (2.0 * DBL_EPSILON));
    a += DBL_EPSILON; b += DBL_EPSILON;
    const double A = (prob) * (a + b), B = (1.0 - prob) * (a + b);
    if (a > A) { return 10.0 / v.ln10 * (a * log(a / A) + b * log(b / B)); }
    return 0.0;
}

UVC_HD double odds2prob(double odds) { return odds / (odds + 1.0); }   // main_conversion.hpp:199-203
// test hook (uvcgpu_selftest_math): one scoring function at one point
UVC_HD double selftest_math(const BatchView & v, int32_t which, double x, double a, double b) {
    switch (which) {
        case 0: return binom_10log10_likeratio(v, x, a, b);
        case 1: return prob2odds(odds2prob(x));
        case 2: return odds2prob(prob2odds(x));
        default: return logit2(a, b);
    }
}

UVC_HD bool implies_short_frag(const GroupFmt & g, int32_t wgs_min_avg_fragsize) { // does_fmt_imply_short_frag (main.hpp:170-174)
    return (g.APLRI[0] + g.APLRI[2]) < (g.APLRI[1] + g.APLRI[3]) * (int64_t)wgs_min_avg_fragsize;
}

UVC_HD double norm_fa_refbias(double FA, double refbias) { return (FA + FA * refbias) / (FA + (1.0 - FA) / (1.0 + refbias) + FA * refbias); }

// fmt_bias_push (main.hpp:4258-4272): the deciphred value, and the FTS entry when the bias-reduced FA is below thres x the plain FA
UVC_HD int32_t bias_push(CandFmt & c, const BatchView & v, int fts_index, double refFA, double biasFA) {
    if (biasFA < refFA * v.par.bias_thres_FTS_FA) {
        c.fts_mask |= (1u << fts_index);
        c.fts_pct[fts_index] = d2i(round(100.0 * biasFA / refFA));
    }
    return -numstates2deciphred(v, biasFA);
}

// FTS entries in the order BcfFormat_symbol_calc_DPv pushes them (main.hpp:4745-4769)
enum { FTS_aStrand, FTS_aBQXM, FTS_aInsertSize, FTS_aAlignL, FTS_aAlignR, FTS_aPositionL, FTS_aPositionR, FTS_abPositionL, FTS_abPositionR,
       FTS_bcDup, FTS_cbDup, FTS_c0Orientation, FTS_c2Orientation, FTS_c2PositionL, FTS_c2PositionR, FTS_c2AlignL, FTS_c2AlignR, FTS_c2StrictPosL, FTS_c2StrictPosR };

// BcfFormat_symbol_calc_DPv (main.hpp:4274-4844), tumor-only branch (tpfa = -1)
UVC_HD void calc_DPv(CandFmt & c, const GroupFmt & g, const BatchView & v, const uvcgpu_prep_set & prep, const uvcgpu_rtr & rtr1, const uvcgpu_rtr & rtr2, int refsymbol) {
    const uvcgpu_params & par = v.par;
    const double unbias_ratio = 1.0, unbias_qualadd = 0;
    const int32_t allbias_allprior = 0;
    const bool is_strong_amplicon = ((prep.a_pcr_dp) * 100 > prep.a_dp * 50);
    const bool is_weak_amplicon = ((prep.a_pcr_dp) * 100 > prep.a_dp * 30);
    const double pfa = 0.5, c2altpc = 0.025;
    const int symbol = c.symbol;
    const bool subst = is_subst(symbol);
    const bool isins = is_ins_symbol(symbol), isdel = is_del_symbol(symbol);
    const int32_t ADP1 = (g.ADPff[0] + g.ADPfr[0] + g.ADPrf[0] + g.ADPrr[0]);
    const int32_t aDP1 = (c.aDPff + c.aDPfr + c.aDPrf + c.aDPrr);
    const int32_t aDP = aDP1;
    const int32_t ADP = tmax(ADP1, prep.a_near_pcr_clip_dp);
    const int32_t cDP1 = (c.cDP1f + c.cDP1r);
    const int32_t CDP1 = g.CDP1b[0] + g.CDP1b[1];
    const double cFA2 = (c.cDP2f + c.cDP2r + c2altpc) / (g.CDP2b[0] + g.CDP2b[1] + 1.0);
    const double cFA3 = (c.cDP3f + c.cDP3r + c2altpc) / (g.CDP3b[0] + g.CDP3b[1] + 1.0);

    double counterbias_P_FA = 1e-9, counterbias_BQ_FA = 1e-9, dir_bias_div = 1.0;
    const bool is_nmore_amplicon = is_strong_amplicon;
    if ((is_nmore_amplicon && (0x2 == (0x2 & par.nobias_flag))) || ((!is_nmore_amplicon) && (0x1 == (0x1 & par.nobias_flag)))) {
        const double using_bias_oddsA = prob2odds((aDP - c.aP1 + 0.5) / (ADP - g.AP1[0] + 1.0));
        const double using_nobias_oddsA = prob2odds((c.aP1 + 0.5) / (g.AP1[0] + 1.0));
        const bool is_pos_counterbias = (
                   (using_bias_oddsA * par.microadjust_counterbias_pos_odds_ratio < using_nobias_oddsA * (unbias_ratio - DBL_EPSILON))
                && (c.aP1 * (unbias_ratio - DBL_EPSILON) > aDP - c.aP1)
                && ((ADP - g.AP1[0]) * par.microadjust_counterbias_pos_fold_ratio * (unbias_ratio - DBL_EPSILON) > g.AP1[0])
                && ((0 == par.primerlen && 0 != par.primerlen2) || !subst));
        if (is_pos_counterbias) {
            counterbias_P_FA = dmax(counterbias_P_FA, (c.aP1 + 0.5) / (tmax(g.AP1[0], prep.a_near_pcr_clip_dp) + 1.0));
        } else {
            counterbias_P_FA = dmax(counterbias_P_FA, 2e-9);
        }
        if (subst) {
            const bool f_good = ((g.ADPfr[0] + g.ADPrr[0]) + 150 <= (g.ADPff[0] + g.ADPrf[0]) * 5 * unbias_ratio);
            const bool r_good = ((g.ADPff[0] + g.ADPrf[0]) + 150 <= (g.ADPfr[0] + g.ADPrr[0]) * 5 * unbias_ratio);
            const int32_t avg_f_aBQ = (c.a1BQf / tmax(1, c.aDPff + c.aDPrf));
            const int32_t avg_r_aBQ = (c.a1BQr / tmax(1, c.aDPfr + c.aDPrr));
            const int32_t avg_f_ABQ = (g.A1BQf[0] / tmax(1, g.ADPff[0] + g.ADPrf[0]));
            const int32_t avg_r_ABQ = (g.A1BQr[0] / tmax(1, g.ADPfr[0] + g.ADPrr[0]));
            if ((c.a1BQf >= c.a1BQr) && (f_good && r_good) && (avg_f_aBQ + unbias_qualadd >= avg_r_ABQ + 14) && (avg_r_ABQ <= 14 + unbias_qualadd)) {
                counterbias_BQ_FA = dmax(counterbias_BQ_FA, (c.aDPff + c.aDPrf + 0.5) / (g.ADPff[0] + g.ADPrf[0] + 1.0));
            }
            if ((c.a1BQr >= c.a1BQf) && (f_good && r_good) && (avg_r_aBQ + unbias_qualadd >= avg_f_ABQ + 14) && (avg_f_ABQ <= 14 + unbias_qualadd)) {
                counterbias_BQ_FA = dmax(counterbias_BQ_FA, (c.aDPfr + c.aDPrr + 0.5) / (g.ADPfr[0] + g.ADPrr[0] + 1.0));
            }
        } else {
            dir_bias_div = (1.0 + (uint64_t)c.gap_len / (uint64_t)par.indel_str_repeatsize_max); // size_t / int: integer division first
        }
    }
    const int32_t aDPgap = nnminus(tmax(g.APDP[1], g.APDP[2]), c.aP3);
    const double aDPFAgap = ((rtr1.tracklen + rtr2.tracklen < par.indel_str_repeatsize_max) ? 1.0 : ((c.aP3 + pfa) / (aDPgap + 1.0)));
    const double aDPFA1 = ((aDP + pfa) / (ADP + 1.0));
    const double labelFA = (c.aP2 + 1.5 + c.aP2) / (g.AP2[0] + 2.0 + c.aP2);
    const double aDPFA = dmin((subst ? dmin(aDPFA1, dmax(aDPFA1 / 3, aDPFAgap)) : (aDPFA1)), labelFA * (ADP + 1.0) / (g.AP2[0] + 0.5) * unbias_ratio);
    const int32_t aDPplus = (subst ? 0 : ((aDP + 1) * par.bias_prior_DPadd_perc / 100));
    const double dp_coef = ((symbol == UVC_LINK_M)
            ? dmax(par.contam_any_mul_frac, 1.0 - tmax(rtr1.tracklen, rtr2.tracklen) / (tmax((int64_t)1, tmax(g.ALPL[0], g.ARPL[0])) / dmax(1.0 / 150.0, (double)g.ABQ2[0]))) : 1.0);
    double aPpriorfreq0 = par.bias_priorfreq_pos;
    double aBpriorfreq0 = aPpriorfreq0;
    const bool is_in_indel_read = ((g.APXM[1]) / 15.0 * par.microadjust_bias_pos_indel_fold * (par.bias_prior_var_DP_mul) > (aDP + aDPplus) * dp_coef);
    const bool is_in_indel_len = (tmax(g.APDP[1], g.APDP[2]) * (par.bias_prior_var_DP_mul) > (aDP + aDPplus) * dp_coef);
    const bool is_in_indel_rtr = (tmax(g.APDP[3], g.APDP[4]) * (par.bias_prior_var_DP_mul) > (aDP + aDPplus) * dp_coef);
    const bool is_in_rtr = (tmax(rtr1.tracklen, rtr2.tracklen) > round(par.indel_polymerase_size));
    if (is_in_indel_read || ((isins || isdel) && (g.APXM[0] > g.APXM[1] * par.microadjust_bias_pos_indel_misma_to_indel_ratio))) {
        aPpriorfreq0 -= par.bias_priorfreq_indel_in_read_div;
        aBpriorfreq0 -= par.bias_priorfreq_indel_in_read_div;
    }
    if (UVC_LINK_M != symbol && UVC_LINK_NN != symbol) {
        double maxpf = 0;
        if (is_in_indel_len) { maxpf = dmax(maxpf, par.bias_priorfreq_indel_in_var_div2); }
        if (is_in_indel_rtr) { maxpf = dmax(maxpf, par.bias_priorfreq_indel_in_str_div2); }
        if (is_in_rtr) { maxpf = dmax(maxpf, par.bias_priorfreq_var_in_str_div2); }
        aBpriorfreq0 -= maxpf; aPpriorfreq0 -= maxpf;
    }
    const double aPpriorfreq = aPpriorfreq0 + allbias_allprior;
    const double aBpriorfreq = aBpriorfreq0 + allbias_allprior;
    c.nPF[0] = d2i(round(aPpriorfreq)); c.nPF[1] = d2i(round(aBpriorfreq));
    const double aIpriorfreq = (subst ? par.bias_priorfreq_ipos_snv : par.bias_priorfreq_ipos_indel) + allbias_allprior;
    const double aSBpriorfreq = (subst ? (tmin(nnminus(c.aBQ, 0), c.bMQ) + par.bias_priorfreq_strand_snv_base) : (par.bias_priorfreq_strand_indel)) + allbias_allprior;
    const double dedup_A2C1_frac = dmin(1.0, (double)tmax(CDP1, par.bias_reduction_by_high_sequencingDP_min_n_totDepth) / (double)tmax(ADP1, 1));
    const double dedup_a2c1_frac = dmin(1.0, (double)tmax(cDP1, par.bias_reduction_by_high_sequencingDP_min_n_altDepth) / (double)tmax(aDP1, 1));
    const double dedup_frag_frac = dmax(dedup_A2C1_frac, dedup_a2c1_frac);
    const double pc_read = ((is_in_indel_read) ? par.bias_FA_pseudocount_indel_in_read : 0.5);
    double x2[2];
    dp4_to_pcFA<false, false>(x2, dedup_frag_frac, c.aLP1, aDP, g.ALP2[0] + c.aLP1 - c.aLP2, ADP, par.powlaw_exponent, phred2nat(v, aPpriorfreq),
            tmax((int64_t)1, c.aLPL) / (double)tmax(1, c.aBQ2), tmax((int64_t)1, g.ALPL[0]) / (double)tmax(1, g.ABQ2[0]), pc_read, 1.0);
    double aLPFA = x2[0];
    dp4_to_pcFA<false, false>(x2, dedup_frag_frac, c.aRP1, aDP, g.ARP2[0] + c.aRP1 - c.aRP2, ADP, par.powlaw_exponent, phred2nat(v, aPpriorfreq),
            tmax((int64_t)1, c.aRPL) / (double)tmax(1, c.aBQ2), tmax((int64_t)1, g.ARPL[0]) / (double)tmax(1, g.ABQ2[0]), pc_read, 1.0);
    double aRPFA = x2[0];
    dp4_to_pcFA<false, false>(x2, dedup_frag_frac, c.aLB1, aDP, g.ALB2[0] + c.aLB1 - c.aLB2, ADP, par.powlaw_exponent, phred2nat(v, aBpriorfreq),
            tmax((int64_t)1, c.aLBL) / (double)tmax(1, c.aBQ2), tmax((int64_t)1, g.ALBL[0]) / (double)tmax(1, g.ABQ2[0]), pc_read, 1.0);
    double aLBFA = x2[0];
    dp4_to_pcFA<false, false>(x2, dedup_frag_frac, c.aRB1, aDP, g.ARB2[0] + c.aRB1 - c.aRB2, ADP, par.powlaw_exponent, phred2nat(v, aBpriorfreq),
            tmax((int64_t)1, c.aRBL) / (double)tmax(1, c.aBQ2), tmax((int64_t)1, g.ARBL[0]) / (double)tmax(1, g.ABQ2[0]), pc_read, 1.0);
    double aRBFA = x2[0];
    const bool is_tmore_amplicon = is_weak_amplicon;

    const int32_t normCDP1 = g.CDP12b[0] + g.CDP12b[1] + 1;
    const int32_t normBDP = (g.BDPb[0] + g.BDPb[1] + 1);
    const int32_t c2DP = (c.cDP2f + c.cDP2r);
    c.enable_tier2 = ((c2DP >= 2) && (normBDP * par.fam_bias_overseq_perc >= normCDP1 * 100) && (prep.a_umi_dp * 100 > prep.a_dp * 50));
    // QUIRK: the divisor uses the first element of the lower-case vector, which at this point still is this candidate's own value (main.hpp:4477-4478)
    const double cFA2L = (c.enable_tier2 ? ((((int64_t)c.c2LP0 * (int64_t)c.c2LP0) * 2 / tmax(1, tmin(c2DP, c.c2LP0 * 4)) + c2altpc) / (g.C2LP0[0] + 1.0)) : 1.0);
    const double cFA2R = (c.enable_tier2 ? ((((int64_t)c.c2RP0 * (int64_t)c.c2RP0) * 2 / tmax(1, tmin(c2DP, c.c2RP0 * 4)) + c2altpc) / (g.C2RP0[0] + 1.0)) : 1.0);
    double c2LPFA = 1.0, c2RPFA = 1.0, c2LBFA = 1.0, c2RBFA = 1.0;
    if (c.enable_tier2) {
        const int32_t C2DP = g.CDP2b[0] + g.CDP2b[1];
        const double c2Ppriorfreq = dmax(0, aPpriorfreq - 0), c2Bpriorfreq = dmax(0, aBpriorfreq - 0);
        dp4_to_pcFA<false, true>(x2, -1, c.c2LP1, c2DP, g.C2LP2[0] + c.c2LP1 - c.c2LP2, C2DP, par.powlaw_exponent, phred2nat(v, c2Ppriorfreq),
                tmax((int64_t)1, c.c2LPL) / (double)tmax(1, c.c2BQ2), tmax((int64_t)1, g.C2LPL[0]) / (double)tmax(1, g.C2BQ2[0]), c2altpc, 1.0);
        c2LPFA = x2[0];
        dp4_to_pcFA<false, true>(x2, -1, c.c2RP1, c2DP, g.C2RP2[0] + c.c2RP1 - c.c2RP2, C2DP, par.powlaw_exponent, phred2nat(v, c2Ppriorfreq),
                tmax((int64_t)1, c.c2RPL) / (double)tmax(1, c.c2BQ2), tmax((int64_t)1, g.C2RPL[0]) / (double)tmax(1, g.C2BQ2[0]), c2altpc, 1.0);
        c2RPFA = x2[0];
        dp4_to_pcFA<false, true>(x2, -1, c.c2LB1, c2DP, g.C2LB2[0] + c.c2LB1 - c.c2LB2, C2DP, par.powlaw_exponent, phred2nat(v, c2Bpriorfreq),
                tmax((int64_t)1, c.c2LBL) / (double)tmax(1, c.c2BQ2), tmax((int64_t)1, g.C2LBL[0]) / (double)tmax(1, g.C2BQ2[0]), c2altpc, 1.0);
        c2LBFA = x2[0];
        dp4_to_pcFA<false, true>(x2, -1, c.c2RB1, c2DP, g.C2RB2[0] + c.c2RB1 - c.c2RB2, C2DP, par.powlaw_exponent, phred2nat(v, c2Bpriorfreq),
                tmax((int64_t)1, c.c2RBL) / (double)tmax(1, c.c2BQ2), tmax((int64_t)1, g.C2RBL[0]) / (double)tmax(1, g.C2BQ2[0]), c2altpc, 1.0);
        c2RBFA = x2[0];
    }
    double aLIFAx2[2], aRIFAx2[2];
    {
        const double ALpd = (g.ALI2[0] + 0.5) / (g.ADPfr[0] + g.ADPrr[0] - g.ALI2[0] + 0.5);
        const double aLpd = (c.aLI1 + ALpd / (1.0 + ALpd)) / (c.aDPfr + c.aDPrr - c.aLI1 + 1.0 / (1.0 + ALpd));
        dp4_to_pcFA<false, false>(aLIFAx2, dedup_frag_frac, (c.aLI1), (c.aDPfr + c.aDPrr), (g.ALI2[0] + c.aLI1 - c.aLI2), (g.ADPfr[0] + g.ADPrr[0]),
                par.powlaw_exponent, phred2nat(v, aIpriorfreq), aLpd, ALpd, 0.25, 0.5);
    }
    double aLIFA = aLIFAx2[0] * ((is_tmore_amplicon) ? (dir_bias_div) : dmax(dir_bias_div, aDPFA / aLIFAx2[1]));
    {
        const double ARpd = (g.ARI2[0] + 0.5) / (g.ADPff[0] + g.ADPrf[0] - g.ARI2[0] + 0.5);
        const double aRpd = (c.aRI1 + ARpd / (1.0 + ARpd)) / (c.aDPff + c.aDPrf - c.aRI1 + 1.0 / (1.0 + ARpd));
        dp4_to_pcFA<false, false>(aRIFAx2, dedup_frag_frac, (c.aRI1), (c.aDPff + c.aDPrf), (g.ARI2[0] + c.aRI1 - c.aRI2), (g.ADPff[0] + g.ADPrf[0]),
                par.powlaw_exponent, phred2nat(v, aIpriorfreq), aRpd, ARpd, 0.25, 0.5);
    }
    double aRIFA = aRIFAx2[0] * ((is_tmore_amplicon) ? (dir_bias_div) : dmax(dir_bias_div, aDPFA / aRIFAx2[1]));
    const double aSIFA = dmax((c.aLI1 + 0.5) / (g.ALI2[0] + c.aLI1 - c.aLI2 + 1.0), (c.aRI1 + 0.5) / (g.ARI2[0] + c.aRI1 - c.aRI2 + 1.0));
    if (isins || isdel) {
        const double coef = tmax(1, c.bDPa) / (double)tmax(1, c.bDPf + c.bDPr);
        const bool in_major_reg = ((tmax(g.APDP[1], g.APDP[3]) + tmax(g.APDP[2], g.APDP[4])) * 0.5 * (1.0 + (double)FLT_EPSILON) < aDP * coef);
        if ((tmin(c.gap_len, par.microadjust_nobias_pos_indel_maxlen) * aDPFA * coef >= par.nobias_pos_indel_lenfrac_thres)
                || (tmax(rtr1.tracklen, rtr2.tracklen) >= par.nobias_pos_indel_str_track_len && in_major_reg
                    && !(g.APXM[0] > g.APXM[1] * par.microadjust_nobias_pos_indel_misma_to_indel_ratio))) {
            aLPFA += 2.0; aRPFA += 2.0; aLBFA += 2.0; aRBFA += 2.0;
            if (c.enable_tier2) { c2LPFA += 2.0; c2RPFA += 2.0; c2LBFA += 2.0; c2RBFA += 2.0; }
        }
        if (c.bMQ >= par.microadjust_nobias_pos_indel_bMQ && c.a2XM2 * 100 >= aDP * 100 * par.microadjust_nobias_pos_indel_perc) { aLIFA += 2.0; aRIFA += 2.0; }
    } else if (UVC_LINK_M == symbol || UVC_LINK_NN == symbol) {
        const double pc = par.bias_FA_pseudocount_indel_in_read;
        aLBFA = dmin(aLBFA, (pc + c.aLB1) / (double)(pc * 2 + ADP));
        aRBFA = dmin(aRBFA, (pc + c.aRB1) / (double)(pc * 2 + ADP));
    } else if (refsymbol == symbol) {
        aLIFA = aRIFA = dmax(aLIFA, aRIFA);
    }
    const int64_t avg_sqr_indel_len = tmax(g.APXM[4] / tmax(1, g.APDP[1]), g.APXM[5] / tmax(1, g.APDP[2]));
    if ((!subst) && ((int64_t)(par.microadjust_nobias_pos_indel_maxlen * par.microadjust_nobias_pos_indel_maxlen) < avg_sqr_indel_len)
            && (UVC_LINK_M == symbol || UVC_LINK_NN == symbol || ((int64_t)((uint64_t)(c.gap_len * 2) * (uint64_t)(c.gap_len * 2)) < avg_sqr_indel_len))) {
        const double pc = par.bias_FA_pseudocount_indel_in_read;
        const double aLPFA_minA = (pc + c.aLP1) / (double)(pc * 2 + g.ALP1[0]);
        const double aRPFA_minA = (pc + c.aRP1) / (double)(pc * 2 + g.ALP1[0]);   // QUIRK: ALP1 for the right side too (main.hpp:4576)
        aLPFA = dmin(aLPFA, aLPFA_minA); aRPFA = dmin(aRPFA, aRPFA_minA);
        if (c.enable_tier2) { c2LPFA = dmin(c2LPFA, aLPFA_minA); c2RPFA = dmin(c2RPFA, aRPFA_minA); }
    }
    const double aPFFA = (c.aPF1 + pfa * 100.0) / (g.APF2[0] + (c.aPF1 - c.aPF2) + 100.0);
    double aSSFAx2[2], cROFA1x2[2], cROFA2x2[2];
    dp4_to_pcFA<true, false>(aSSFAx2, dedup_frag_frac, c.aRIf, c.aLIr, g.ARIf[0], g.ALIr[0], par.powlaw_exponent, phred2nat(v, aSBpriorfreq), -1, -1, 0.5, 1.0);
    const double ori_base = (subst ? par.bias_priorfreq_orientation_snv_base : par.bias_priorfreq_orientation_indel_base) + allbias_allprior;
    const double eff = dmax(aDPFA, par.bias_orientation_min_effective_allelefrac);
    const double ori_all = log(eff * eff) + phred2nat(v, ori_base);
    dp4_to_pcFA<true, false>(cROFA1x2, dedup_frag_frac, c.cDP1f, c.cDP1r, g.CDP1b[0], g.CDP1b[1], par.powlaw_exponent, ori_all, -1, -1, 0.5, 1.0);
    dp4_to_pcFA<true, true>(cROFA2x2, -1, c.cDP2f, c.cDP2r, g.CDP2b[0], g.CDP2b[1], par.powlaw_exponent, ori_all, -1, -1, c2altpc, 1.0);
    double aSSFA = aSSFAx2[0] * dir_bias_div;
    double cROFA1 = cROFA1x2[0] * dir_bias_div;
    double cROFA2 = cROFA2x2[0] * dir_bias_div;
    if (isins || isdel) { c.bAD = tmin(c.bAD, c.bDPa); c.AD = tmin(c.AD, c.cDP0a); }
    const double bFA = (c.bDPa + pfa) / (g.BDPb[0] + g.BDPb[1] + 1.0);
    const double cFA0 = (c.cDP0a + pfa * (implies_short_frag(g, par.lib_wgs_min_avg_fraglen) ? par.lib_nonwgs_ad_pseudocount : 1.0)) / (g.CDP1b[0] + g.CDP1b[1] + 1.0);
    const bool strand_r_weak = ((g.ADPfr[0] + g.ADPrr[0]) * par.microadjust_nobias_strand_all_fold < (g.ADPff[0] + g.ADPrf[0]) * unbias_ratio);
    const bool strand_f_weak = ((g.ADPff[0] + g.ADPrf[0]) * par.microadjust_nobias_strand_all_fold < (g.ADPfr[0] + g.ADPrr[0]) * unbias_ratio);
    if (strand_r_weak) { aLIFA += 4.0; aSSFA += 4.0; }
    if (strand_f_weak) { aRIFA += 4.0; aSSFA += 4.0; }
    const double aLPFA2 = dmax(aDPFA * 0.01, aLPFA), aRPFA2 = dmax(aDPFA * 0.01, aRPFA), aLBFA2 = dmax(aDPFA * 0.01, aLBFA), aRBFA2 = dmax(aDPFA * 0.01, aRBFA);
    const double c2LPFA2 = dmax(cFA2 * 0.01, c2LPFA), c2RPFA2 = dmax(cFA2 * 0.01, c2RPFA), c2LBFA2 = dmax(cFA2 * 0.01, c2LBFA), c2RBFA2 = dmax(cFA2 * 0.01, c2RBFA);
    const double aLIFA2 = dmax(aDPFA * 0.01, aLIFA), aRIFA2 = dmax(aDPFA * 0.01, aRIFA), aSSFA2 = dmax(aDPFA * 0.05, aSSFA);
    cROFA1 = dmax(aDPFA * 1e-4, cROFA1); cROFA2 = dmax(aDPFA * 1e-4, cROFA2);
    // systematic error from the density of nearby mutations
    const double fBTA = (double)(g.BTAb[0] + g.BTAb[1] + 200), fBTB = (double)(g.BTBb[0] + g.BTBb[1] + 6);
    const double fbTA = (double)(c.bTAf + c.bTAr + 100), fbTB = (double)(c.bTBf + c.bTBr + 3);
    const double frag_sidelen_frac = 1.0 - tmin(
            (int64_t)tmin(tmax((int64_t)0, c.aLIT / tmax(1, c.aDPfr + c.aDPrr) - par.microadjust_longfrag_sidelength_min), (int64_t)par.microadjust_longfrag_sidelength_max),
            (int64_t)tmin(tmax((int64_t)0, c.aRIT / tmax(1, c.aDPff + c.aDPrf) - par.microadjust_longfrag_sidelength_min), (int64_t)par.microadjust_longfrag_sidelength_max))
            / par.microadjust_longfrag_sidelength_zeroMQpenalty;
    const double alt_frac0 = fbTB / fbTA;
    const double alt_frac = (is_nmore_amplicon ? (dmax(0, alt_frac0 - 0.2) * 1.25) : alt_frac0);
    const double nonalt_frac = (fBTB + par.contam_any_mul_frac * fbTB - fbTB) / (fBTA + par.contam_any_mul_frac * fbTA - fbTA);
    const double frac_mut = dmax(par.syserr_MQ_NMR_expfrac, par.syserr_MQ_NMR_altfrac_coef * alt_frac * frag_sidelen_frac - par.syserr_MQ_NMR_nonaltfrac_coef * nonalt_frac);
    c.bNMQ = d2i(round(numstates2phred(v, pow(frac_mut / par.syserr_MQ_NMR_expfrac, (par.syserr_MQ_NMR_pl_exponent))) * (frac_mut)));
    c.bNMa = d2i(round(100 * alt_frac)); c.bNMb = d2i(round(100 * nonalt_frac));
    const bool tmore_with_primerlen = (is_tmore_amplicon || ((par.primerlen > 0) && !(0x4 & par.primer_flag)));
    const double bFAa = bFA;
    double t1only = dmin(cROFA1, dmin(aLPFA2, dmin(aRPFA2, dmin(aLBFA2, dmin(aRBFA2, cFA0)))));
    t1only = dmin(t1only, aDPFA * dmin(dmax(1.0 + aDPFA - alt_frac, 0.1), 1.0));
    t1only = dmin(t1only, aPFFA * aSSFA2 / dmax(aSSFA2, aSSFAx2[1]));
    const double t1plus = dmin(aSSFA2, dmin(aLIFA2, dmin(aRIFA2, dmin(dmax(aDPFA * 0.01, aSIFA), bFAa))));
    const double cFA2a = (tmore_with_primerlen ? (cFA2 * (par.powlaw_amplicon_allele_fraction_coef)) : cFA2);
    const double cFA3a = ((normBDP * 100 > normCDP1 * ((par.fam_tier3DP_bias_overseq_perc - 100) / 1 + 100)) ? cFA3 : 1.0);
    const double c23FA = cFA2a;
    const double t2only = dmin(cROFA2, dmin(c2LPFA2, dmin(c2RPFA2, dmin(c2LBFA2, dmin(c2RBFA2, dmin(cFA2a, dmin(cFA3a, dmin(cFA2L, cFA2R))))))));
    c.nNFA[0] = -numstates2deciphred(v, counterbias_P_FA); c.nNFA[1] = -numstates2deciphred(v, counterbias_BQ_FA);
    c.nNFA[2] = -numstates2deciphred(v, aDPFA); c.nNFA[3] = -numstates2deciphred(v, bFA); c.nNFA[4] = -numstates2deciphred(v, cFA0); c.nNFA[5] = -numstates2deciphred(v, cFA2);
    c.nAFA[0] = bias_push(c, v, FTS_aStrand, aDPFA, aSSFA2); c.nAFA[1] = bias_push(c, v, FTS_aBQXM, aDPFA, aPFFA); c.nAFA[2] = bias_push(c, v, FTS_aInsertSize, aDPFA, aSIFA);
    c.nAFA[3] = bias_push(c, v, FTS_aAlignL, aDPFA, aLBFA2); c.nAFA[4] = bias_push(c, v, FTS_aAlignR, aDPFA, aRBFA2);
    c.nAFA[5] = bias_push(c, v, FTS_aPositionL, aDPFA, aLPFA2); c.nAFA[6] = bias_push(c, v, FTS_aPositionR, aDPFA, aRPFA2);
    c.nAFA[7] = bias_push(c, v, FTS_abPositionL, aDPFA, aLIFA2); c.nAFA[8] = bias_push(c, v, FTS_abPositionR, aDPFA, aRIFA2);
    c.nBCFA[0] = bias_push(c, v, FTS_bcDup, bFA, cFA0); c.nBCFA[1] = bias_push(c, v, FTS_cbDup, cFA0, bFA);
    c.nBCFA[2] = bias_push(c, v, FTS_c0Orientation, cFA0, cROFA1); c.nBCFA[3] = bias_push(c, v, FTS_c2Orientation, cFA2, cROFA2);
    c.nBCFA[4] = bias_push(c, v, FTS_c2PositionL, cFA2, c2LPFA2); c.nBCFA[5] = bias_push(c, v, FTS_c2PositionR, cFA2, c2RPFA2);
    c.nBCFA[6] = bias_push(c, v, FTS_c2AlignL, cFA2, c2LBFA2); c.nBCFA[7] = bias_push(c, v, FTS_c2AlignR, cFA2, c2RBFA2);
    c.nBCFA[8] = bias_push(c, v, FTS_c2StrictPosL, cFA2, cFA2L); c.nBCFA[9] = bias_push(c, v, FTS_c2StrictPosR, cFA2, cFA2R);
    const double aNCFA = ((implies_short_frag(g, par.lib_wgs_min_avg_fraglen) && (isins || isdel) && c.gap_len >= par.lib_nonwgs_clip_penal_min_indelsize)
            ? dmax((c.aNC + 0.5) / (ADP + 1.0), dmin(dmax((c.cDP1f + c.cDP1r) / 300.0, 1.0 / 3.0), 2.0 / 3.0) * aDPFA) : 2.0);
    const double counterbias_FA = dmax(counterbias_P_FA, dmax(counterbias_BQ_FA, 1e-9));
    const double dedup_FA = dmin(bFA, cFA0);
    const double frac_umi2seg = dmin(1.0, dmin(c23FA / aDPFA, aDPFA / c23FA));
    const double refbias = 0;
    const int32_t CDP1sum = g.CDP1b[0] + g.CDP1b[1], CDP2sum = g.CDP2b[0] + g.CDP2b[1];
    const double min_abcFA_v = dmax(dmin(dmin(t1plus, t1only), aNCFA), counterbias_FA);
    c.cDP1v = d2i((norm_fa_refbias(min_abcFA_v, refbias) * CDP1sum * 100));
    const double min_abcFA_w = dmax(dmin(aLPFA2, dmin(aRPFA2, dmin(aLBFA2, dmin(aRBFA2, dmin(bFA, aNCFA))))), counterbias_FA);
    c.cDP1w = d2i((norm_fa_refbias(min_abcFA_w, refbias) * CDP1sum * 100));
    const double min_abcFA_x = dmin(aPFFA, dedup_FA);
    c.cDP1x = 1 + d2i((min_abcFA_x * CDP1sum * 100));
    const double cube = cFA2 * cFA2 * cFA2;
    const double c2XBFA2 = dmin(dmax(3.0 * c2LBFA2 * c2RBFA2 * aSSFA2 / cube, dmin(c2LBFA2, c2RBFA2) / 8.0), dmin(c2LBFA2, c2RBFA2));
    const double c2XPFA2 = dmin(dmax(3.0 * c2LPFA2 * c2RPFA2 * aSSFA2 / cube, dmin(c2LPFA2, c2RPFA2) / 8.0), dmin(c2LPFA2, c2RPFA2));
    const double c2XXFA2 = dmin(c2XBFA2, c2XPFA2);
    const double min_c23FA_v = dmax(dmin(dmin(t1plus, dmin(t2only, c2XXFA2)), aNCFA), counterbias_FA * frac_umi2seg);
    c.cDP2v = d2i((norm_fa_refbias(min_c23FA_v, refbias) * CDP2sum * 100));
    const double min_c23FA_w = dmax(dmin(c2LPFA2, dmin(c2RPFA2, dmin(c2XXFA2, dmin(c2LBFA2, dmin(c2RBFA2, dmin(cFA2, aNCFA)))))), counterbias_FA * frac_umi2seg);
    c.cDP2w = d2i((norm_fa_refbias(min_c23FA_w, refbias) * CDP2sum * 100));
    const double min_c23FA_x = dmin(aPFFA, c23FA);
    c.cDP2x = 1 + d2i((min_c23FA_x * CDP2sum * 100));
}

// indelpos_to_context (main.hpp:733-755): best short tandem repeat starting at reference index refidx of the tile's reference string
UVC_HD void repeat_at(int32_t & unit_len, int32_t & repeatnum, const BatchView & v, const TileInfo & T, int32_t refidx) {
    const uint8_t *ref = v.refsym + T.pos_off;
    const int32_t n = (T.ext_end - T.ext_beg) - 1;
    repeatnum = 0; unit_len = 0;
    if (refidx >= n) { return; }
    const int32_t str_max = v.par.indel_str_repeatsize_max;
    for (int32_t unit = 1; unit <= str_max; unit++) {
        int32_t q = refidx;
        while ((q + unit < n) && ref[q] == ref[q + unit]) { q++; }
        const int32_t num = (q - refidx) / unit + 1;
        bool better;
        if (unit_len * repeatnum == 0) { better = true; }
        else {
            int rank1 = (num <= 1 ? (-num * unit) : ((num - 1) * unit));
            int rank2 = (repeatnum <= 1 ? (-repeatnum * unit) : ((repeatnum - 1) * unit_len));
            if (0 == num) { rank1 = -100; }
            if (0 == repeatnum || 0 == unit_len) { rank2 = -100; }
            better = (rank1 > rank2);
        }
        if (better) { repeatnum = num; unit_len = unit; }
    }
}

// BcfFormat_symbol_calc_qual (main.hpp:4908-5343), tumor-only branch (tpfa = -1, is_rescued = false)
UVC_HD void calc_qual(CandFmt & c, const GroupFmt & g, const BatchView & v, const uvcgpu_prep_set & prep, const uvcgpu_rtr & rtr1, const uvcgpu_rtr & rtr2,
        int refsymbol, int32_t ins_cdepth, int32_t del_cdepth, int32_t ins1_cdepth, int32_t del1_cdepth, int32_t repeatunit_len, int32_t repeatnum) {
    const uvcgpu_params & par = v.par;
    const int symbol = c.symbol;
    const bool subst = is_subst(symbol), isins = is_ins_symbol(symbol), isdel = is_del_symbol(symbol);
    const int32_t CDP2sum = g.CDP2b[0] + g.CDP2b[1], CDP1sum = g.CDP1b[0] + g.CDP1b[1];
    const double cFA2 = (c.cDP2f + c.cDP2r + 0.5) / (CDP2sum + 1.0);
    const int32_t powlaw_sscs_phrederr = sscs_phred(par, refsymbol, symbol) + 0;
    const double umi_cFA = (((double)(c.cDP2v) + 0.5) / ((double)(CDP2sum * 100 + 1.0)));
    const double umi_cFA_w = (((double)(c.cDP2w) + 0.5) / ((double)(CDP2sum * 100 + 1.0)));
    const int32_t powlaw_sscs_inc1 = (int32_t)(powlaw_sscs_phrederr - (subst
            ? (((UVC_BASE_A == refsymbol && UVC_BASE_T == symbol) || (UVC_BASE_T == refsymbol && UVC_BASE_A == symbol))
                ? (double)par.fam_phred_pow_sscs_transversion_AT_TA_origin : par.fam_phred_pow_sscs_snv_origin)
            : par.fam_phred_pow_sscs_indel_origin));
    int32_t powlaw_sscs_inc4tn = (subst
            ? (int32_t)(tmax(tmax(par.fam_phred_sscs_transition_CG_TA, par.fam_phred_sscs_transition_AT_GC), tmax(par.fam_phred_sscs_transversion_CG_AT, par.fam_phred_sscs_transversion_other))
                - (par.fam_phred_pow_sscs_snv_origin))
            : powlaw_sscs_inc1);
    const bool is_oxidation = ((UVC_BASE_C == refsymbol && UVC_BASE_A == symbol) || (UVC_BASE_G == refsymbol && UVC_BASE_T == symbol));
    powlaw_sscs_inc4tn += (is_oxidation ? par.tn_q_inc_max_sscs_CG_AT : par.tn_q_inc_max_sscs_other);
    const double t2n_contam_frac = 0 * par.contam_t2n_mul_frac;
    const double contamfrac = par.contam_any_mul_frac + (1.0 - par.contam_any_mul_frac) * t2n_contam_frac;
    const int32_t aDP = (c.aDPff + c.aDPfr + c.aDPrf + c.aDPrr);
    const int32_t ADP = (g.ADPff[0] + g.ADPrf[0] + g.ADPfr[0] + g.ADPrr[0]);
    const int32_t cDP0 = (c.cDP1f + c.cDP1r), CDP0 = CDP1sum;
    const int32_t cDP2 = (c.cDP2f + c.cDP2r), CDP2 = CDP2sum;
    const int32_t aavgMQ = (int32_t)(c.aMQs / tmax(1, aDP));
    const int32_t diffAaMQs = (int32_t)((g.AMQs[0] - c.aMQs) / tmax(1, ADP - aDP)) - aavgMQ;
    const int32_t tn_q_inc_max = par.tn_q_inc_max;
    const int32_t noUMI_bias_inc = tmin(par.bias_FA_powerlaw_noUMI_phred_inc_snv, aDP / 2);
    const double pl_noUMI_phred_inc = par.powlaw_anyvar_base + (subst ? noUMI_bias_inc : par.bias_FA_powerlaw_noUMI_phred_inc_indel);
    const int32_t withUMI_bias_inc = tmin(par.bias_FA_powerlaw_withUMI_phred_inc_snv - par.bias_FA_powerlaw_noUMI_phred_inc_snv, cDP2 / 2) + noUMI_bias_inc;
    const double pl_withUMI_phred_inc = par.powlaw_anyvar_base + (subst ? withUMI_bias_inc : par.bias_FA_powerlaw_withUMI_phred_inc_indel);
    const double prior_weight = 1.0 / (c.cDPmf + c.cDPmr + 1.0);
    const int32_t fam_thres_highBQ = (subst ? par.fam_thres_highBQ_snv : par.fam_thres_highBQ_indel);
    const int32_t cMmQ = d2i(round(numstates2phred(v, (c.cDPMf + c.cDPmf + c.cDPMr + c.cDPmr + pow(10, fam_thres_highBQ / 10.0) * prior_weight) / (c.cDPmf + c.cDPmr + prior_weight))));
    const int32_t nbases_x100_1 = c.bIADb * 100 + 1;
    const int32_t nbases_x100_2 = tmin(nbases_x100_1, c.cDP1v + 1);
    const int64_t perbase_q_x10_1 = 10 * c.bIAQb / tmax(1, c.bIADb);
    const int64_t perbase_q_x10_2 = perbase_q_x10_1 + d2l(round(10 * numstates2phred(v, (double)nbases_x100_2 / (double)nbases_x100_1)));
    int64_t duped_frag_binom_qual = ((isins || isdel) ? perbase_q_x10_1 : perbase_q_x10_2) * nbases_x100_2 / (10 * 100);
    const int64_t contam_frag_withmin_qual = d2l(round(binom_10log10_likeratio(v, t2n_contam_frac, cDP0, CDP0 - cDP0))) + 9 - 3;
    const int32_t inc_snp = tmax(0, 2 * par.germ_phred_hetero_snp - par.germ_phred_het3al_snp - 0);
    const int32_t inc_indel = tmax(0, 2 * par.germ_phred_hetero_indel - par.germ_phred_het3al_indel - 0);
    int32_t phred_het3al_chance_inc = (subst ? inc_snp : inc_indel);
    if (isins || isdel) { phred_het3al_chance_inc = nnminus(inc_indel + 1, c.gap_len); }
    const int32_t contam_syserr_phred_bypassed = phred_het3al_chance_inc;
    const int32_t normcDP1 = (c.cDP12f + c.cDP12r + 1);
    const int32_t normCDP1 = g.CDP12b[0] + g.CDP12b[1] + 1;
    const int32_t normBDP = g.BDPb[0] + g.BDPb[1] + 1;
    const int64_t sscs_dec1a = (((par.fam_min_n_copies / 1 <= normCDP1) || (par.fam_min_n_copies_DPxAD / 1 <= (int64_t)normCDP1 * (int64_t)normcDP1)) ? 0 : (powlaw_sscs_inc1 + 3));
    const int64_t sscs_dec1b = (((int64_t)((par.fam_min_overseq_perc - 100) / 1 + 100) * (int64_t)normCDP1 <= (int64_t)100 * (int64_t)normBDP) ? 0 : (powlaw_sscs_inc1 + 3));
    const int64_t sscs_dec1 = tmax(sscs_dec1a, sscs_dec1b);
    const int64_t sscs_dec2 = nnminus(fam_thres_highBQ, cMmQ);
    const int64_t cIADnormcnt = (int64_t)(c.cIADf + c.cIADr) * 100 + 1;
    const int64_t cIADmincnt = tmin(cIADnormcnt, (int64_t)(c.cDP2v + 1));
    const int64_t sscs_binom_qual_fw = c.cIAQf + ((int64_t)c.cIAQr * (int64_t)tmin(par.fam_phred_dscs_all - c.cIDQf, c.cIDQr)) / tmax(c.cIDQr, 1);
    const int64_t sscs_binom_qual_rv = c.cIAQr + ((int64_t)c.cIAQf * (int64_t)tmin(par.fam_phred_dscs_all - c.cIDQr, c.cIDQf)) / tmax(c.cIDQf, 1);
    const int64_t contam_sscs_withmin_qual = d2l(round(binom_10log10_likeratio(v, t2n_contam_frac, cDP2, CDP2 - cDP2))) + 9 - 3;
    const int64_t sscs_max = tmax(sscs_binom_qual_fw, sscs_binom_qual_rv);
    // non_neg_minus(int64, double) evaluated in double, then int64mul(double -> int64, cIADmincnt)
    const double sub_d = numstates2phred(v, cIADnormcnt / (double)cIADmincnt) * cIADnormcnt / 100.0;
    const double nnm_d = (((double)sscs_max > sub_d) ? ((double)sscs_max - sub_d) : 0);
    int64_t sscs_binom_qual = (d2l(nnm_d) * cIADmincnt) / (cIADnormcnt);
    if (sscs_max > par.microadjust_fam_binom_qual_halving_thres && subst) {
        sscs_binom_qual = tmin(sscs_binom_qual, (int64_t)(par.microadjust_fam_binom_qual_halving_thres + (sscs_max - par.microadjust_fam_binom_qual_halving_thres) / 2));
    }
    sscs_binom_qual -= sscs_dec1 + sscs_dec2;
    const double min_bcFA_v = (((double)(c.cDP1v) + 0.5) / (double)(CDP1sum * 100 + 1.0));
    int32_t dedup_frag_powlaw_qual_v = d2i(round(par.powlaw_exponent * numstates2phred(v, min_bcFA_v) + (pl_noUMI_phred_inc)));
    const double min_bcFA_w = (((double)(c.cDP1w) + 0.5) / (double)(CDP1sum * 100 + 1.0));
    int32_t dedup_frag_powlaw_qual_w = d2i(round(par.powlaw_exponent * numstates2phred(v, min_bcFA_w) + (pl_noUMI_phred_inc) + tn_q_inc_max));
    const int32_t ds_vq_inc_powlaw = d2i(round(10 / v.ln10 * dmin(log((c.cDP12f + 0.5) / (g.CDP12b[0] + 1.0)), log((c.cDP12r + 0.5) / (g.CDP12b[1] + 1.0))))) + (powlaw_sscs_phrederr);
    const int32_t ds_vq_inc_binom = 3 * tmin(c.cDP2f, c.cDP2r);
    const int64_t m5 = tmin(tmin(sscs_binom_qual_fw, sscs_binom_qual_rv), tmin((int64_t)ds_vq_inc_powlaw, tmin((int64_t)ds_vq_inc_binom, (int64_t)3)));
    const int32_t powlaw_sscs_inc2 = (int32_t)(tmax((int64_t)0, m5) * ((cFA2 > 0.002) ? 1 : 0));
    const int32_t sscs_dec3 = ((cFA2 >= 0.003) ? 0 : 5);
    const int32_t sscs_base_2 = (int32_t)(pl_withUMI_phred_inc + powlaw_sscs_inc1 + powlaw_sscs_inc2 - sscs_dec1 - sscs_dec2 - sscs_dec3);
    const int32_t sscs_base_2tn = (int32_t)(pl_withUMI_phred_inc + powlaw_sscs_inc4tn + powlaw_sscs_inc2 - sscs_dec1 - sscs_dec2 - sscs_dec3);
    int32_t sscs_powlaw_qual_v = d2i(round((par.powlaw_exponent * numstates2phred(v, umi_cFA) + sscs_base_2)));
    int32_t sscs_powlaw_qual_w = d2i(round((par.powlaw_exponent * numstates2phred(v, umi_cFA_w) + sscs_base_2tn)));
    const double dFA = (double)(c.dDP2 + 0.5) / (double)(g.DDP1[0] + 1.0);
    const double dSNR = (double)(c.dDP2 + 0.5) / (double)(c.dDP1 + 1.0);
    const double dnormFA = dFA * pow(dSNR, 1.0 / par.powlaw_exponent);
    const int64_t fam_phred_dscs_estimated = d2l(round((par.fam_phred_dscs_max + powlaw_sscs_phrederr) / 2.0));
    const int64_t dFA_vq_binom = (fam_phred_dscs_estimated - d2l(round(numstates2phred(v, 1.0 / (dnormFA))))) * (int64_t)c.dDP2 * (int64_t)cIADmincnt / (int64_t)cIADnormcnt;
    const int32_t dFA_vq_powlaw = d2i((par.powlaw_anyvar_base + (fam_phred_dscs_estimated - par.fam_phred_pow_dscs_all_origin)
            + d2i(round(numstates2phred(v, (dnormFA) * dmin(1.0, (double)((c.cDP1v) + 0.5) / (double)(CDP1sum * 100 + 1.0)))))));
    c.cMmQ = cMmQ;
    const double eps = (double)FLT_EPSILON;
    const int32_t indel_penal_base = 0;     // IonTorrent only
    int32_t indel_penal4multialleles = 0, indel_penal4multialleles_g = 0, indel_penal4multialleles_soma = 0, indel_UMI_penal = 0;
    if (c.gap_len > 0 && c.cDP0a > 0) {
        const double indel_pq = (double)tmin(slip_phred_lookup(v, 0, repeatunit_len, repeatnum), 24) + 2 - (double)10;
        const int32_t eff_tracklen1 = (repeatunit_len * tmax(1, repeatnum) - repeatunit_len);
        const int32_t eff_tracklen2 = (tmax(rtr1.tracklen - rtr1.unitlen, rtr2.tracklen - rtr2.unitlen) / 3);
        const double indel_ic = numstates2phred(v, (double)tmax((int64_t)c.gap_len + (isins ? 1 : 0), (int64_t)1) / (double)(tmax(eff_tracklen1, eff_tracklen2) + 1))
                + (isins ? (numstates2phred(v, par.indel_del_to_ins_err_ratio) * tmin(200, c.cDP0a) / 200) : 0);
        double indelcdepth = (isins ? ins_cdepth : del_cdepth);
        int32_t indelcdepth_i = (isins ? ins_cdepth : del_cdepth);
        if (UVC_LINK_D1 == symbol) { indelcdepth_i += ins1_cdepth; }
        if (UVC_LINK_I1 == symbol) { indelcdepth_i += d2i((del1_cdepth / par.indel_del_to_ins_err_ratio)); } // auto is int: += double truncates
        indelcdepth = indelcdepth_i;
        const int32_t nearInDelDP = (isins ? g.APDP[1] : g.APDP[2]);
        const int32_t penal1 = d2i(round(par.indel_multiallele_samepos_penal / log(2.0) * log((double)(indelcdepth + eps) / (double)(c.cDP0a + eps))));
        const int32_t penal2 = d2i(round(par.indel_multiallele_diffpos_penal / log(2.0) * log((double)(nearInDelDP + eps) / (double)(tmax(aDP, nearInDelDP) + eps))));
        indel_penal4multialleles_g = d2i((d2i(round(par.indel_tetraallele_germline_penal_value / log(2.0) * log((double)(ins_cdepth + del_cdepth + eps) / (double)(c.cDP0a + eps))))
                - par.indel_tetraallele_germline_penal_thres));
        if (isins) {
            indel_penal4multialleles = (penal1 * par.indel_ins_penal_pseudocount / (int32_t)(par.indel_ins_penal_pseudocount + c.gap_len));
            indel_penal4multialleles_soma = indel_penal4multialleles;
        } else {
            indel_penal4multialleles = tmax(penal1, penal2);
            indel_penal4multialleles_soma = penal1;
        }
        dedup_frag_powlaw_qual_v += d2i(round(indel_ic));
        dedup_frag_powlaw_qual_w += d2i(round(indel_ic));
        duped_frag_binom_qual += d2l(round(indel_pq));
        const uint32_t gl = (uint32_t)tmax(c.gap_len, 1);
        const double sscs_indel_ic = numstates2phred(v, (double)(gl * gl) / (double)(tmax(eff_tracklen1, eff_tracklen2) + 1));
        const int32_t sscs_ins_vs_del_inc = d2i(round(par.powlaw_exponent * numstates2phred(v, par.indel_del_to_ins_err_ratio)));
        const double rhs = sscs_indel_ic * (isins ? 0 : tmax(eff_tracklen1, eff_tracklen2)) / round(par.indel_polymerase_size);
        const int32_t extra_reward = d2i(((((double)sscs_ins_vs_del_inc > rhs) ? ((double)sscs_ins_vs_del_inc - rhs) : 0) - sscs_ins_vs_del_inc / 2));
        sscs_powlaw_qual_v += d2i((round(sscs_indel_ic) + extra_reward));
        sscs_powlaw_qual_w += d2i((round(sscs_indel_ic) + extra_reward));
        sscs_binom_qual += d2l((round(indel_pq) + extra_reward));
        if (c.enable_tier2) {
            const double lhs = (g.BDPb[0] + g.BDPb[1] + 1.0) / (double)(CDP1sum + 1.0) * par.fam_indel_nonUMI_phred_dec_per_fold_overseq;
            const double rr = (par.fam_thres_emperr_all_flat_indel + 1) * par.fam_indel_nonUMI_phred_dec_per_fold_overseq;
            indel_UMI_penal = d2i(((lhs > rr) ? (lhs - rr) : 0));
        }
    }
    c.aAaMQ = diffAaMQs;
    const int32_t readlenMQcap = (int32_t)((g.APXM[2]) / tmax(1, g.APDP[0]) - 17);
    const int32_t diffMQ = (tmax(0, diffAaMQs));
    const bool is_aln_extra_accurate = (par.inferred_maxMQ > 60);
    const int32_t sysMQVQadd = ((symbol == refsymbol) ? 0 : (tmin(par.germ_phred_homalt_snp, ADP * 3)));
    const int32_t sysMQVQadd_somatic = ((symbol != refsymbol) ? 0 : (tmin(par.germ_phred_homalt_snp, ADP * 3)));
    const bool is_MQ_unadjusted = (is_aln_extra_accurate || (!subst) || (aDP > ADP * 3 / 4));
    const int32_t sysMQVQminus = (is_MQ_unadjusted ? 0 : (nnminus((60 - 30), aavgMQ) * 2 / 5))
            + ((is_MQ_unadjusted || (refsymbol != symbol)) ? 0 : nnminus(tmin(15, diffMQ), aavgMQ));
    int32_t diffMQ2 = diffMQ;
    if (c.bMQ < 20) {
        const double aDPxf = (c.aDPff + c.aDPrf + 0.5), aDPxr = (c.aDPfr + c.aDPrr + 0.5);
        const double ADPxf = (g.ADPff[0] + g.ADPrf[0] + 1.0), ADPxr = (g.ADPfr[0] + g.ADPrr[0] + 1.0);
        if ((aDPxr / ADPxr) * 2 < (aDPxf / ADPxf) || (aDPxf / ADPxf) * 2 < (aDPxr / ADPxr)
                || (c.aLI1 + 0.5) / (g.ALI2[0] + 1.0) * (2 * (1.0 + DBL_EPSILON)) < (aDPxr) / (ADPxr)
                || (c.aRI1 + 0.5) / (g.ARI2[0] + 1.0) * (2 * (1.0 + DBL_EPSILON)) < (aDPxf) / (ADPxf)) {
            diffMQ2 = tmax(diffMQ2, 20 - tmin(c.bMQ, 20));
        }
    }
    const int32_t sysMQ_base = d2i(((c.bMQ * (par.syserr_MQ_max - par.syserr_MQ_nonref_base) / par.syserr_MQ_max + par.syserr_MQ_nonref_base))) - (int32_t)(diffMQ2) - (int32_t)(c.bNMQ);
    const int32_t sysMQ = (((refsymbol == symbol) && (ADP > aDP * 2)) ? c.bMQ : (sysMQ_base - d2i((numstates2phred(v, (ADP + 1.0) / (aDP + 0.5))))));
    const bool is_nonWGS = implies_short_frag(g, par.lib_wgs_min_avg_fraglen);
    const int32_t normal_rescued_MQ = tmin(nnminus(readlenMQcap, 60), (is_nonWGS ? par.lib_nonwgs_normal_max_rescued_MQ : par.lib_wgs_normal_max_rescued_MQ));
    int32_t sysMQVQ1 = tmin((tmax(sysMQ, par.syserr_MQ_min) + sysMQVQadd), readlenMQcap);
    const int32_t sysBQVQ = (subst ? c.aBQQ : (200));
    const bool is_weak_amplicon = ((prep.a_pcr_dp * 100) > g.APDP[0] * 30);
    const bool is_tmore_amplicon = is_weak_amplicon;
    if (is_tmore_amplicon && (isins || isdel) && (sysMQVQ1 > 70) && (g.APXM[1] / tmax(g.APDP[0], 1) > 20)) {
        sysMQVQ1 = (int32_t)(70 + ((sysMQVQ1 - 70) * 5 / (g.APXM[1] / tmax(g.APDP[0], 1) - 15)));
    }
    int32_t indel_penal_base_add = 0;
    {
        const int32_t delAPDP = tmax(g.APDP[2], g.APDP[4]);
        if ((g.APDP[0] < 3 * delAPDP) && (g.APDP[0] < 3 * prep.a_snv_dp) && (aDP * 3 < delAPDP) && (aDP * 3 < prep.a_snv_dp) && subst && (rtr2.tracklen >= 8 * rtr2.unitlen)) {
            indel_penal_base_add = par.microadjust_germline_mix_with_del_snv_penalty;
        }
        if (is_tmore_amplicon && isdel) {
            if (aDP * 4 < g.APDP[2]) { indel_penal_base_add = tmax(indel_penal_base_add, 5); }
            else if (c.cDP0a * 3 < 2 * (del_cdepth)) { indel_penal_base_add = tmax(indel_penal_base_add, 2); }
        }
    }
    const int32_t sysMQVQ = tmax(0, sysMQVQ1);
    const int32_t indel_penal_base2 = indel_penal_base + indel_penal_base_add;
    const int32_t ADPfx = g.ADPff[0] + g.ADPfr[0], ADPrx = g.ADPrf[0] + g.ADPrr[0], ADPxf = g.ADPff[0] + g.ADPrf[0], ADPxr = g.ADPfr[0] + g.ADPrr[0];
    const bool frx_imba = (tmax(ADPfx, ADPrx) > par.microadjust_strand_orientation_absence_DP_fold * (tmin(ADPfx, ADPrx) + 1));
    const bool xfr_imba = (tmax(ADPxf, ADPxr) > par.microadjust_strand_orientation_absence_DP_fold * (tmin(ADPxf, ADPxr) + 1));
    const int32_t powlaw_v_minus = (subst ? ((frx_imba ? par.microadjust_orientation_absence_snv_penalty : 0) + (xfr_imba ? par.microadjust_strand_absence_snv_penalty : 0))
            : (is_tmore_amplicon ? par.microadjust_dedup_absence_indel_penalty : 0));
    const int32_t tn_syserr_q = sysMQVQ + par.tn_q_inc_max + normal_rescued_MQ;
    c.bMQQ = sysMQVQ;
    c.bIAQ = (int32_t)(duped_frag_binom_qual - indel_penal_base2);
    c.cIAQ = (int32_t)(sscs_binom_qual - indel_penal_base);
    c.cPCQ1 = tmin(dedup_frag_powlaw_qual_w - indel_penal_base2, tn_syserr_q);
    c.cPLQ1 = dedup_frag_powlaw_qual_v - indel_penal_base2 - powlaw_v_minus;
    c.cPCQ2 = tmin(sscs_powlaw_qual_w - indel_penal_base, tn_syserr_q);
    c.cPLQ2 = sscs_powlaw_qual_v - indel_penal_base;
    c.bTINQ = (int32_t)(contam_frag_withmin_qual + contam_syserr_phred_bypassed);
    c.cTINQ = (int32_t)(contam_sscs_withmin_qual + contam_syserr_phred_bypassed);
    const int32_t aDPpc = ((refsymbol == symbol) ? 1 : 0);
    const int64_t dd = (int64_t)tmax(1, aDP + aDPpc);
    const int32_t penal4BQerr = (subst ? (5 + (int32_t)(((int64_t)par.penal4lowdep) / (dd * dd))) : 0);
    const int32_t indel_q_inc = (((!isins) && (!isdel)) ? 0 : units_phred(c.gap_len, repeatnum));
    c.gVQ1 = tmax(0, indel_q_inc + tmin(tmin(sysBQVQ, nnminus(sysMQVQ, sysMQVQminus)), tmin(c.bIAQ - penal4BQerr, c.cPLQ1))
            - 2 * tmax(0, tmax(d2i((indel_penal4multialleles - par.indel_multiallele_soma_penal_thres)), indel_penal4multialleles_g)));
    const int32_t sysVQsomatic_minus = (15 - tmin(tmin(ADP * 15 / 100, aDP), 15));
    const int32_t sysVQsomatic = nnminus(tmin(sysBQVQ, sysMQVQ + sysMQVQadd_somatic), sysVQsomatic_minus);
    const int32_t bcVQ1 = tmin(tmin(sysVQsomatic, c.bIAQ - penal4BQerr), c.cPLQ1) - indel_penal4multialleles_soma;
    c.cVQ1 = tmax(0, tmin(bcVQ1, c.bTINQ) - indel_UMI_penal);
    int32_t mincVQ2 = 0;
    if (isins || isdel) {
        const int32_t floor_v = d2i((dmin(par.germ_phred_homalt_indel + numstates2phred(v, umi_cFA), (double)(c.cDP2v * 3 / 100)) + ((isins ? 1 : 0) - 1) * 3));
        mincVQ2 = tmax(mincVQ2, floor_v);
    }
    const int64_t dVQinc = tmin(tmin(dFA_vq_binom, (int64_t)dFA_vq_powlaw) - tmax(0, tmin(c.cIAQ, c.cPLQ2)), (int64_t)par.fam_phred_dscs_inc_max);
    c.dVQinc = (int32_t)dVQinc;
    const int32_t cVQ2 = (int32_t)(tmin((int64_t)sysVQsomatic, tmin(c.cIAQ + tmax((int64_t)0, dVQinc), c.cPLQ2 + tmax((int64_t)0, dVQinc))) - indel_penal4multialleles);
    c.cVQ2 = tmax(mincVQ2, tmin(cVQ2, c.cTINQ));
    // CONTQ uses this candidate's cDP1v and the group's CDP1v[0]
    const double binom_contam = binom_10log10_likeratio(v, contamfrac, c.cDP1v, g.CDP1v[0]);
    const double power_contam = round(10.0 / v.ln10 * par.powlaw_exponent * dmax(logit2((c.cDP1v + 1) / (double)(g.CDP1v[0] + 1), contamfrac), 0.0));
    c.CONTQ = d2i(dmin(binom_contam, power_contam));
}

// hetLODQ (main.hpp:5457-5462)
UVC_HD int32_t het_lodq(const BatchView & v, double a1, double a2, double expfrac, double pl_exponent) {
    const int32_t binomLODQ = d2i(binom_10log10_likeratio(v, expfrac, a1, a2));
    const int32_t powerLODQ = d2i(round(10.0 / v.ln10 * pl_exponent * dmax(logit2((a1 + 0.5) * 0.5 / expfrac, (a2 + 0.5) * 0.5 / (1.0 - expfrac)), 0.0)));
    return tmin(binomLODQ, powerLODQ);
}

// The part of output_germline that every record needs (main.hpp:5483-5616): normal-LOD of the site from the four genotype likelihoods.
// cands: this group's candidates; n: their number. Padding entries behave like the reference's modified init_fmt (VTI = END, gVQ1 = 0, cDP1v = 50).
UVC_HD int32_t germline_nlodq(const BatchView & v, const CandFmt *cands, int n, int type, int refsymbol) {
    const uvcgpu_params & par = v.par;
    // symbol_format_vec: candidates except BASE_NN, padded to at least 5; stable descending sort by gVQ1 (insertion sort on <= 16 elements)
    int order[UVC_MAX_GROUP_CANDS + 5]; int m = 0;
    for (int i = 0; i < n; i++) { if (cands[i].symbol != UVC_BASE_NN) { order[m++] = i; } }
    while (m <= 4) { order[m++] = -1; }
    #define UVC_G(i) ((i) < 0 ? 0 : cands[i].gVQ1)
    #define UVC_SYM(i) ((i) < 0 ? UVC_NSYM : cands[i].symbol)
    #define UVC_V(i) ((i) < 0 ? 50 : cands[i].cDP1v)
    // std::sort(rbegin, rend, less) on <= 16 elements is an insertion sort over the reversed range: equal keys keep their relative order in the
    // reversed range, i.e. among equal gVQ1 the LATER element (in forward order) comes first after sorting the reversed view ... which in forward
    // order means: descending by gVQ1, ties in ORIGINAL order reversed twice = original order preserved.
    for (int i = 1; i < m; i++) {
        const int x = order[i]; int j = i - 1;
        while (j >= 0 && UVC_G(order[j]) < UVC_G(x)) { order[j + 1] = order[j]; j--; }
        order[j + 1] = x;
    }
    int sel[4] = {-2, -2, -2, -2};
    int allele_idx = 1; int32_t ref_alodq = INT32_MIN;
    for (int k = 0; k < m; k++) {
        const int sy = UVC_SYM(order[k]);
        const bool isref = (refsymbol == sy || UVC_BASE_NN == sy || UVC_LINK_NN == sy);
        if (isref && UVC_G(order[k]) > ref_alodq) { sel[0] = order[k]; ref_alodq = UVC_G(order[k]); }
        if ((!isref) && allele_idx <= 3) { sel[allele_idx] = order[k]; allele_idx++; }
    }
    int32_t a0 = UVC_G(sel[0]), a1 = UVC_G(sel[1]), a2 = UVC_G(sel[2]), a3 = UVC_G(sel[3]);
    const bool isSubst = is_subst(refsymbol);
    const int symbolNN = UVC_BASE_NN;   // (isSubst || !is_rescued) ? BASE_NN : LINK_NN
    double ad0 = UVC_V(sel[0]) / 100.0, ad1 = UVC_V(sel[1]) / 100.0, ad2 = UVC_V(sel[2]) / 100.0;
    if (symbolNN == UVC_SYM(sel[1])) { ad0 += ad1; ad1 = 0; }
    if (symbolNN == UVC_SYM(sel[2])) { ad0 += ad2; ad2 = 0; }
    const int32_t a0a1 = het_lodq(v, ad0, ad1, 1.0 - par.germ_hetero_FA, par.powlaw_exponent);
    const int32_t a1a0 = het_lodq(v, ad1, ad0, par.germ_hetero_FA, par.powlaw_exponent);
    const int32_t a1a2 = het_lodq(v, ad1, ad2, 0.5, par.powlaw_exponent);
    const int32_t a2a1 = het_lodq(v, ad2, ad1, 0.5, par.powlaw_exponent);
    const int32_t phred_homref = 0;
    const int32_t phred_hetero = (isSubst ? par.germ_phred_hetero_snp : par.germ_phred_hetero_indel);
    const int32_t phred_homalt = (isSubst ? par.germ_phred_homalt_snp : par.germ_phred_homalt_indel);
    const int32_t phred_tri_al = (isSubst ? par.germ_phred_het3al_snp : par.germ_phred_het3al_indel);
    a0 = tmin(a0, (sel[0] < 0 ? 0 : cands[sel[0]].CONTQ));
    const int32_t a2penal = tmax(a2 - (phred_tri_al - phred_hetero), 0);
    const int32_t a3penal = tmax(a3 - phred_hetero, 0);
    const int32_t a01hetp = tmax(tmax(a0a1, a1a0) - (0 - 0), 0);
    const int32_t a12hetp = tmax(tmax(a1a2, a2a1) - (3 - 0), 0);
    const int32_t a03trip = tmax(a0, a3);
    int32_t tri_al_penal = 0;
    const int symb1 = UVC_SYM(sel[1]), symb2 = UVC_SYM(sel[2]);
    if (is_ins_symbol(symb1) && is_ins_symbol(symb2)) {
        tri_al_penal += 3;
        if (symb1 == symb2) { tri_al_penal += 3; if (UVC_LINK_I3P == symb1) { tri_al_penal += 3; } }
    }
    {
        const int32_t nunits[UVC_NSYM + 1] = {0, 0, 0, 0, 0, 0, 0, -3, -2, -1, 3, 2, 1, 0, 0};   // SYMBOL_TO_INDEL_N_UNITS (main.hpp:271-279)
        const int32_t n1 = nunits[symb1], n2 = nunits[symb2];
        if (n1 != 0 && n2 != 0) { tri_al_penal -= between(iabs(n1 - n2) * 3 - 5, 0, 9); }
    }
    // Genotype log-likelihoods GL4raw (main.hpp:5601-5607). The three alternative ones are kept NEGATED (n_k = -GL_k) and reduced with a minimum:
    // ptxas 12.9 for sm_100a drops the sign of one operand when it fuses max(-x, -y, -z) into VIMNMX3 (found with the parity tests).
    const int32_t gl0 = (-phred_homref - a1 - a2penal - a3penal);
    const int32_t n1 = (phred_hetero + tmax(a01hetp, a2) + tmax(tmin(a01hetp, a2) - phred_hetero, 0) + a3penal);
    const int32_t n2 = (phred_homalt + tmax(a0, a2) + tmax(tmin(a0, a2) - phred_hetero, 0) + a3penal);
    const int32_t n3 = (phred_tri_al + tmax(a12hetp, a03trip) + tmax(tmin(a12hetp, a03trip) - phred_hetero, 0) + tmax(tmin(a12hetp, tmin(a0, a3)) - phred_hetero, 0) + tri_al_penal);
    #undef UVC_G
    #undef UVC_SYM
    #undef UVC_V
    (void)type;
    return gl0 + tmin(n1, tmin(n2, n3));
}

// calc_binom_powlaw_syserr_normv_quals (main.hpp:5982-6010)
UVC_HD void tn_quals(int32_t out[4], const BatchView & v, double tAD, double tDP, int32_t tVQ, int32_t tnVQcap, double nAD, double nDP, int32_t nVQ,
        double penal_dimret_coef, int32_t prior_phred, int32_t tn_dec_by_xm, double pl_exponent) {
    const int32_t binom = d2i(binom_10log10_likeratio(v, (tDP - tAD) / (tDP), nDP - nAD, nAD));
    const double nADplus = nAD * dmin(dmax(nDP / tDP - 1.0, 0), 1);
    const double bjpfrac = ((tAD + 0.5) / (tDP + 1.0)) / ((nAD + 0.5 + nADplus) / (nDP + 1.0 + nADplus));
    const int32_t powlaw = d2i(round(pl_exponent * numstates2phred(v, bjpfrac)));
    const int32_t tnVQinc = tmax(-prior_phred, tmax((-d2i(nAD)) * 3, tmin(binom - prior_phred, powlaw - prior_phred)));
    const double lg = log(dmax(bjpfrac, 1.001)) / log(2.0);
    int32_t tnVQdec = tmax(0, nVQ - tmax(0, tmin(binom - prior_phred, d2i(((lg * lg) * penal_dimret_coef)))));
    tnVQdec = tmax(tnVQdec, tmin(nVQ + 9, tn_dec_by_xm));
    const int32_t tnVQ = tmin(tnVQcap, tVQ + tnVQinc) - tnVQdec;
    out[0] = binom; out[1] = powlaw; out[2] = tnVQdec; out[3] = tnVQ;
}

UVC_HD const IndelAllele *find_alleles(const ScoreView & sv, int64_t gp, int symbol, int32_t & n) {
    const int64_t key = gp * 16 + symbol;
    int64_t a = 0, b = sv.n_alleles;
    while (a < b) { const int64_t m = (a + b) >> 1; if (sv.alleles[m].key < key) { a = m + 1; } else { b = m; } }
    int64_t e = a;
    while (e < sv.n_alleles && sv.alleles[e].key == key) { e++; }
    n = (int32_t)(e - a);
    return sv.alleles + a;
}

// Lexicographic comparison of the indel strings of two alleles of the same symbol (the last key of the top-2 ordering, main.cpp:1000).
UVC_HD int allele_string_cmp(const BatchView & v, const CandFmt & a, const CandFmt & b) {
    if (a.ev < 0 || b.ev < 0) { return 0; }
    if (is_del_symbol(a.symbol)) { return (a.gap_len > b.gap_len) - (a.gap_len < b.gap_len); }   // same start: the shorter is a prefix of the longer
    const IndelEvent & ea = v.ev[a.ev]; const IndelEvent & eb = v.ev[b.ev];
    const uint8_t *sa = v.seq + v.reads[ea.read].seq_off, *sb = v.seq + v.reads[eb.read].seq_off;
    const int32_t n = tmin(ea.oplen, eb.oplen);
    for (int32_t i = 0; i < n; i++) {
        const int32_t qa = ea.qpos + i, qb = eb.qpos + i;
        const int ca = (sa[qa >> 1] >> ((~qa & 1) << 2)) & 0xf, cb = (sb[qb >> 1] >> ((~qb & 1) << 2)) & 0xf;
        if (ca != cb) {
            const char nt16[17] = "=ACMGRSVTWYHKDBN";
            return (nt16[ca] > nt16[cb]) - (nt16[ca] < nt16[cb]);
        }
    }
    return (ea.oplen > eb.oplen) - (ea.oplen < eb.oplen);
}

// ------------------------------------------------------------------------------------------------ K5a: one thread per zero-based position
// The candidate test of the reference's per-position loop (main.cpp:832-837): a symbol is a candidate if it is an alternative allele with at
// least min_altdp_thres fragments, or the reference allele next to at least that many non-reference fragments. Most positions have none, so
// the positions that do are compacted into a list and the heavy scoring kernel (K5) runs on full warps of candidate positions only.
UVC_HD bool symbol_is_candidate(const uvcgpu_params & par, int refsymbol, int symbol, int32_t bdepth, int32_t BDP, int32_t ref_bdepth) {
    return !((((refsymbol != symbol) && (bdepth < par.min_altdp_thres)) || ((refsymbol == symbol) && (BDP - ref_bdepth < par.min_altdp_thres))) && (!par.should_output_all));
}

UVC_HD void k5a_flag_position(const BatchView & v, const ScoreView & sv, int64_t gp_zb) {
    const TileInfo & T = v.tiles[v.pos_tile[gp_zb]];
    if (T.skipped) { return; }
    const int32_t zb = (int32_t)(gp_zb - T.pos_off) + T.ext_beg;
    if (zb < T.rpos_inclu_beg || zb > T.rpos_exclu_end) { return; }
    const uvcgpu_params & par = v.par;
    const int32_t nref = (T.ext_end - T.ext_beg) - 1;
    const uint8_t *refsyms = v.refsym + T.pos_off;
    const int refsym_base = ((nref == (zb - 1 - T.ext_beg)) || (-1 == (zb - 1 - T.ext_beg))) ? UVC_BASE_NN : (int)refsyms[zb - 1 - T.ext_beg];
    bool any = false;
    for (int type = 0; type < 2 && !any; type++) {
        if (zb == T.rpos_inclu_beg && type == 0) { continue; }
        const int32_t refpos = (type == 0 ? zb - 1 : zb);
        const int64_t gp = T.pos_off + (refpos - T.ext_beg);
        const int refsymbol = (type == 0 ? refsym_base : UVC_LINK_M);
        const int32_t *fd0 = v.fragdepth + ((0 * v.n_pos + gp) * UVC_NSYM) * UVCGPU_NUM_FRAG_DEPTHS;
        const int32_t *fd1 = v.fragdepth + ((1 * v.n_pos + gp) * UVC_NSYM) * UVCGPU_NUM_FRAG_DEPTHS;
        int32_t BDP = 0;
        for (int k = 0; k < type_nsym(type); k++) { const int sy = type_symbol(type, k); BDP += fd0[sy * 3] + fd1[sy * 3]; }
        const int32_t ref_bdepth = fd0[refsymbol * 3] + fd1[refsymbol * 3];
        for (int k = 0; k < type_nsym(type); k++) {
            const int sy = type_symbol(type, k);
            if (symbol_is_candidate(par, refsymbol, sy, fd0[sy * 3] + fd1[sy * 3], BDP, ref_bdepth)) { any = true; break; }
        }
    }
    if (!any) { return; }
#if defined(__CUDA_ARCH__)
    const int32_t slot = atomicAdd(sv.cand_cursor, 1);
#else
    const int32_t slot = *sv.cand_cursor; *sv.cand_cursor += 1;
#endif
    sv.cand_list[slot] = (int32_t)gp_zb;
}

// ------------------------------------------------------------------------------------------------ K5: the reference's per-position loop as a pipeline
// One iteration of the reference's per-position loop (main.cpp:608-1172) without the text, cut where its data dependences are:
//   K5b  thread / candidate position : enumerates the candidate alleles of the two symbol types in the reference's order and reserves their slots
//   K5g  thread / group              : BcfFormat_symboltype_init (the upper-case tags of one (refpos, symbol type))
//   K5c  thread / candidate          : BcfFormat_symbol_init + BcfFormat_symbol_calc_DPv; BcfFormat_symbol_sum_DPv as integer atomic adds
//   K5e  thread / candidate          : BcfFormat_symbol_calc_qual (needs the sums of all candidates of the group)
//   K5f  thread / group              : output_germline's site likelihood, the top-2 alleles, append_vcf_record's qualities and keep decision
// Candidates and groups live in global arrays between the kernels: a thread holds one candidate (no per-thread candidate tables), and every
// candidate of the batch is a thread of its own instead of a loop iteration of its position's thread.
UVC_HD int32_t atomic_add_i32(int32_t *p, int32_t x) {
#if defined(__CUDA_ARCH__)
    return atomicAdd(p, x);
#else
    const int32_t old = *p; *p += x; return old;
#endif
}

struct K5Site {
    const TileInfo *T; int32_t ti, zb, nref, refsym_base, prev_base1, prev_base2, next_base1, next_base2, minABQ_snv, minABQ_indel;
};
UVC_HD bool k5_site(K5Site & S, const BatchView & v, int64_t gp_zb) {
    S.ti = v.pos_tile[gp_zb];
    const TileInfo & T = v.tiles[S.ti];
    S.T = &T;
    if (T.skipped) { return false; }
    S.zb = (int32_t)(gp_zb - T.pos_off) + T.ext_beg;
    if (S.zb < T.rpos_inclu_beg || S.zb > T.rpos_exclu_end) { return false; }
    const uvcgpu_params & par = v.par;
    S.nref = (T.ext_end - T.ext_beg) - 1;      // refstring.size()
    S.minABQ_snv = (T.is_amplicon_inferred ? par.syserr_minABQ_pcr_snv : par.syserr_minABQ_cap_snv);
    S.minABQ_indel = (T.is_amplicon_inferred ? par.syserr_minABQ_pcr_indel : par.syserr_minABQ_cap_indel);
    const uint8_t *refsyms = v.refsym + T.pos_off;
    const int32_t refidx = S.zb - T.ext_beg;
    S.refsym_base = ((S.nref == (refidx - 1)) || (-1 == (refidx - 1))) ? UVC_BASE_NN : (int)refsyms[refidx - 1];
    S.prev_base1 = ((refidx >= 2) ? (int)refsyms[refidx - 2] : UVC_BASE_NN);
    S.prev_base2 = ((refidx >= 3) ? (int)refsyms[refidx - 3] : UVC_BASE_NN);
    S.next_base1 = ((refidx < S.nref) ? (int)refsyms[refidx] : UVC_BASE_NN);
    S.next_base2 = ((refidx + 1 < S.nref) ? (int)refsyms[refidx + 1] : UVC_BASE_NN);
    return true;
}

// K5b. Pass 0 counts the candidates of the two groups, pass 1 writes their descriptors into the reserved slots (same enumeration twice: the
// order of the candidates inside a group is the reference's, main.cpp:832-905).
UVC_HD void k5b_list_position(const BatchView & v, const ScoreView & sv, int64_t gp_zb) {
    K5Site S;
    if (!k5_site(S, v, gp_zb)) { return; }
    const TileInfo & T = *S.T;
    const uvcgpu_params & par = v.par;
    int32_t ins_cdepth = 0, del_cdepth = 0, ins1_cdepth = 0, del1_cdepth = 0;
    int32_t ncand[2] = {0, 0}, first[2] = {0, 0}, gidx[2] = {-1, -1};
    for (int pass = 0; pass < 2; pass++) {
        if (1 == pass) {
            const int32_t ng = (ncand[0] > 0 ? 1 : 0) + (ncand[1] > 0 ? 1 : 0);
            if (0 == ng) { return; }
            const int32_t g0 = atomic_add_i32(sv.out_cursor + 2, ng);
            const int32_t c0 = atomic_add_i32(sv.out_cursor + 3, ncand[0] + ncand[1]);
            if (g0 + ng > sv.group_cap || c0 + ncand[0] + ncand[1] > sv.cand_cap) { return; }     // (the host sees the totals and runs again with room for all)
            int32_t gi = g0;
            if (ncand[0] > 0) { gidx[0] = gi++; }
            if (ncand[1] > 0) { gidx[1] = gi++; }
            first[0] = c0; first[1] = c0 + ncand[0];
            int32_t repeatunit_len = 0, repeatnum = 0;
            repeat_at(repeatunit_len, repeatnum, v, T, S.zb - T.ext_beg);
            for (int type = 0; type < 2; type++) {
                if (gidx[type] < 0) { continue; }
                GroupRec & G = sv.groups[gidx[type]];
                const int32_t refpos = (type == 0 ? S.zb - 1 : S.zb);
                G.gp = T.pos_off + (refpos - T.ext_beg);
                G.tile = S.ti; G.refpos = refpos; G.type = type; G.refsymbol = (type == 0 ? S.refsym_base : UVC_LINK_M);
                G.first = first[type]; G.n = ncand[type]; G.partner = gidx[1 - type];
                G.ins_cdepth = ins_cdepth; G.del_cdepth = del_cdepth; G.ins1_cdepth = ins1_cdepth; G.del1_cdepth = del1_cdepth;
                G.repeatunit_len = repeatunit_len; G.repeatnum = repeatnum;
            }
            ncand[0] = ncand[1] = 0;
        }
        for (int type = 0; type < 2; type++) {
            if (S.zb == T.rpos_inclu_beg && type == 0) { continue; }
            const int32_t refpos = (type == 0 ? S.zb - 1 : S.zb);
            const int64_t gp = T.pos_off + (refpos - T.ext_beg);
            const int refsymbol = (type == 0 ? S.refsym_base : UVC_LINK_M);
            const PosPtrs P = pos_ptrs(v, gp);
            int32_t BDP = 0;
            for (int k = 0; k < type_nsym(type); k++) { const int sy = type_symbol(type, k); BDP += P.fd0[sy * 3] + P.fd1[sy * 3]; }
            const int32_t ref_bdepth = P.fd0[refsymbol * 3] + P.fd1[refsymbol * 3];
            for (int k = 0; k < type_nsym(type); k++) {
                const int symbol = type_symbol(type, k);
                const int32_t bdepth = P.fd0[symbol * 3] + P.fd1[symbol * 3];
                const int32_t cdepth = tmax(P.fm0[symbol * UVCGPU_NUM_FAM_DEPTHS + 0], P.fm0[symbol * UVCGPU_NUM_FAM_DEPTHS + 1])
                                     + tmax(P.fm1[symbol * UVCGPU_NUM_FAM_DEPTHS + 0], P.fm1[symbol * UVCGPU_NUM_FAM_DEPTHS + 1]);
                if (0 == pass) {
                    if (is_ins_symbol(symbol)) { ins_cdepth += cdepth; if (UVC_LINK_I1 == symbol) { ins1_cdepth += cdepth; } }
                    else if (is_del_symbol(symbol)) { del_cdepth += cdepth; if (UVC_LINK_D1 == symbol) { del1_cdepth += cdepth; } }
                }
                if (!symbol_is_candidate(par, refsymbol, symbol, bdepth, BDP, ref_bdepth)) { continue; }
                const bool is_homopol_1bp = (S.prev_base1 == refsymbol && S.next_base1 == refsymbol);
                const bool is_homopol_2bp = (S.prev_base2 == refsymbol && S.next_base2 == refsymbol);
                const int32_t minABQ = (is_subst(symbol) ? nnminus(S.minABQ_snv, (is_homopol_1bp ? (is_homopol_2bp ? 20 : 10) : 0)) : S.minABQ_indel);
                #define UVC_K5B_ADD(bDPa_, cDP0a_, ev_, len_) { \
                    if (ncand[type] < UVC_MAX_GROUP_CANDS) { \
                        if (1 == pass) { CandDesc & d = sv.desc[first[type] + ncand[type]]; d.group = gidx[type]; d.symbol = symbol; d.bDPa = (bDPa_); d.cDP0a = (cDP0a_); d.ev = (ev_); d.gap_len = (len_); d.minABQ = minABQ; } \
                        ncand[type]++; } }
                if (is_ins_symbol(symbol) || is_del_symbol(symbol)) {
                    int32_t na = 0;
                    const IndelAllele *al = find_alleles(sv, gp, symbol, na);
                    for (int32_t a = 0; a < na; a++) { UVC_K5B_ADD(al[a].bAD, al[a].cAD, al[a].ev, al[a].len) }
                    if (0 == na) {
                        // no read carries this indel symbol here (only reachable with all-out): one placeholder allele whose string is the symbol's
                        // description, e.g. "<LI1>" (indel_get_majority, main.hpp:5412-5418)
                        const int32_t desc_len = ((UVC_LINK_D3P == symbol || UVC_LINK_I3P == symbol) ? 6 : 5);
                        UVC_K5B_ADD(0, 0, -1, desc_len)
                    }
                } else { UVC_K5B_ADD(bdepth, cdepth, -1, 0) }
                #undef UVC_K5B_ADD
            }
        }
    }
}

// K5g: BcfFormat_symboltype_init of one group
UVC_HD void k5g_group(const BatchView & v, const ScoreView & sv, int64_t gi) {
    GroupRec & G = sv.groups[gi];
    group_init(G.g, pos_ptrs(v, G.gp), G.type);
}

// K5c: one candidate through BcfFormat_symbol_init and BcfFormat_symbol_calc_DPv; its six depth values are added to the group's sums
// (BcfFormat_symbol_sum_DPv, main.hpp:4888-4906: integer adds, any order)
UVC_HD void k5c_candidate(const BatchView & v, const ScoreView & sv, int64_t ci) {
    const CandDesc d = sv.desc[ci];
    GroupRec & G = sv.groups[d.group];
    const TileInfo & T = v.tiles[G.tile];
    const uvcgpu_rtr *rtr = v.rtr + T.pos_off;
    const int32_t nrtr = T.ext_end - T.ext_beg;
    const uvcgpu_rtr & rtr1 = rtr[tmax(G.refpos - T.ext_beg, 3) - 3];
    const uvcgpu_rtr & rtr2 = rtr[tmin(G.refpos - T.ext_beg + 3, nrtr - 1)];
    const PosPtrs P = pos_ptrs(v, G.gp);
    CandFmt c;
    cand_init(c, G.g, v, P, d.symbol, d.bDPa, d.cDP0a, d.ev, d.gap_len, d.minABQ);
    calc_DPv(c, G.g, v, *P.prep, rtr1, rtr2, G.refsymbol);
    c.cMmQ = c.aAaMQ = c.bMQQ = c.bIAQ = c.cIAQ = c.cPCQ1 = c.cPLQ1 = c.cPCQ2 = c.cPLQ2 = c.bTINQ = c.cTINQ = c.gVQ1 = c.cVQ1 = c.cVQ2 = c.dVQinc = c.CONTQ = 0;
    sv.cands[ci] = c;
    atomic_add_i32(&G.g.CDP1v[0], c.cDP1v); atomic_add_i32(&G.g.CDP1w[0], c.cDP1w); atomic_add_i32(&G.g.CDP1x[0], c.cDP1x);
    atomic_add_i32(&G.g.CDP2v[0], c.cDP2v); atomic_add_i32(&G.g.CDP2w[0], c.cDP2w); atomic_add_i32(&G.g.CDP2x[0], c.cDP2x);
    if (UVC_BASE_NN == c.symbol || UVC_LINK_NN == c.symbol) {
        G.g.CDP1v[1] = c.cDP1v; G.g.CDP1w[1] = c.cDP1w; G.g.CDP1x[1] = c.cDP1x; G.g.CDP2v[1] = c.cDP2v; G.g.CDP2w[1] = c.cDP2w; G.g.CDP2x[1] = c.cDP2x;
    }
}

// K5e: BcfFormat_symbol_calc_qual of one candidate
UVC_HD void k5e_candidate(const BatchView & v, const ScoreView & sv, int64_t ci) {
    const CandDesc d = sv.desc[ci];
    const GroupRec & G = sv.groups[d.group];
    const TileInfo & T = v.tiles[G.tile];
    const uvcgpu_rtr *rtr = v.rtr + T.pos_off;
    const int32_t nrtr = T.ext_end - T.ext_beg;
    const uvcgpu_rtr & rtr1 = rtr[tmax(G.refpos - T.ext_beg, 3) - 3];
    const uvcgpu_rtr & rtr2 = rtr[tmin(G.refpos - T.ext_beg + 3, nrtr - 1)];
    CandFmt c = sv.cands[ci];
    calc_qual(c, G.g, v, v.prep[G.gp], rtr1, rtr2, G.refsymbol, G.ins_cdepth, G.del_cdepth, G.ins1_cdepth, G.del1_cdepth, G.repeatunit_len, G.repeatnum);
    CandFmt & o = sv.cands[ci];
    o.cMmQ = c.cMmQ; o.aAaMQ = c.aAaMQ; o.bMQQ = c.bMQQ; o.bIAQ = c.bIAQ; o.cIAQ = c.cIAQ; o.cPCQ1 = c.cPCQ1; o.cPLQ1 = c.cPLQ1; o.cPCQ2 = c.cPCQ2; o.cPLQ2 = c.cPLQ2;
    o.bTINQ = c.bTINQ; o.cTINQ = c.cTINQ; o.gVQ1 = c.gVQ1; o.cVQ1 = c.cVQ1; o.cVQ2 = c.cVQ2; o.dVQinc = c.dVQinc; o.CONTQ = c.CONTQ;
}

// number of non-reference alleles of a group that reach the tri-allelic threshold (curr_vAC, main.cpp:975-980)
UVC_HD int32_t k5_vac(const BatchView & v, const ScoreView & sv, int32_t gi) {
    if (gi < 0) { return 0; }
    const GroupRec & G = sv.groups[gi];
    const int32_t het3al = (G.type == 0 ? v.par.germ_phred_het3al_snp : v.par.germ_phred_het3al_indel);
    int32_t n = 0;
    for (int i = 0; i < G.n; i++) { const CandFmt & c = sv.cands[G.first + i]; if (G.refsymbol != c.symbol && tmax(c.cVQ1, c.cVQ2) >= het3al) { n++; } }
    return n;
}

// K5f: the records of one group (third loop, main.cpp:1073-1171 + append_vcf_record)
UVC_HD void k5f_group(const BatchView & v, const ScoreView & sv, int64_t gi) {
    const GroupRec & G = sv.groups[gi];
    const uvcgpu_params & par = v.par;
    const TileInfo & T = v.tiles[G.tile];
    const int type = G.type, refsymbol = G.refsymbol;
    const CandFmt *C = sv.cands + G.first;
    const int ncand = G.n;
    const GroupFmt & g = G.g;
    int32_t curr_vAC[2];
    curr_vAC[type] = k5_vac(v, sv, (int32_t)gi); curr_vAC[1 - type] = k5_vac(v, sv, G.partner);
    int refi = -1;
    for (int i = 0; i < ncand; i++) { if (C[i].symbol == refsymbol) { refi = i; } }
    if (refi < 0) { return; }   // the reference aborts here ("has no REF allele")
    const int32_t nlodq_site = germline_nlodq(v, C, ncand, type, refsymbol);
    // top-2 non-reference alleles by (max(cVQ1, cVQ2), cVQ1, cVQ2, symbol, indel string) descending (main.cpp:1000):
    int top[2] = {-1, -1};
    for (int r = 0; r < 2; r++) {
        for (int i = 0; i < ncand; i++) {
            const CandFmt & c = C[i];
            if (c.symbol == refsymbol || i == top[0]) { continue; }
            if (top[r] < 0) { top[r] = i; continue; }
            const CandFmt & b = C[top[r]];
            const int32_t mc = tmax(c.cVQ1, c.cVQ2), mb = tmax(b.cVQ1, b.cVQ2);
            if (mc > mb || (mc == mb && (c.cVQ1 > b.cVQ1 || (c.cVQ1 == b.cVQ1 && (c.cVQ2 > b.cVQ2 || (c.cVQ2 == b.cVQ2 && (c.symbol > b.symbol || (c.symbol == b.symbol && allele_string_cmp(v, c, b) > 0)))))))) { top[r] = i; }
        }
    }
    const CandFmt & R = C[refi];
    const uvcgpu_rtr *rtr = v.rtr + T.pos_off;
    const int32_t nrtr = T.ext_end - T.ext_beg;
    const int32_t refpos = G.refpos;
    for (int i = 0; i < ncand; i++) {
        const CandFmt & c = C[i];
        const int symbol = c.symbol;
        if (!(par.outvar_flag & 0x4)) { continue; }
        if (((UVC_BASE_NN == symbol) && !(0x20 & par.outvar_flag)) || ((UVC_LINK_NN == symbol) && !(0x40 & par.outvar_flag))) { continue; }
        const int32_t germ_phred = (is_subst(symbol) ? par.germ_phred_hetero_snp : par.germ_phred_hetero_indel);
        const int32_t nlodq1 = nlodq_site - 3 + germ_phred;
        // fill_tki (a = 1: this allele) and fill_conditional_tki<true>
        const int32_t tki_BDP = g.BDPb[0] + g.BDPb[1], tki_bDP = c.bDPf + c.bDPr;
        const int32_t tki_CDP1x = g.CDP1x[0], tki_cDP1x = c.cDP1x, tki_CDP2x = g.CDP2x[0], tki_cDP2x = c.cDP2x;
        const int32_t inc_snp = tmax(0, 2 * par.germ_phred_hetero_snp - par.germ_phred_het3al_snp - 0);
        const int32_t inc_indel = tmax(0, 2 * par.germ_phred_hetero_indel - par.germ_phred_het3al_indel - 0);
        int32_t het3al_inc = (is_subst(symbol) ? inc_snp : inc_indel);
        if (is_ins_symbol(symbol) || is_del_symbol(symbol)) { het3al_inc = nnminus(inc_indel + 1, c.gap_len); }
        const int32_t tn_dec_by_xm = between(tmin(c.bNMQ, c.bNMQ), par.microadjust_syserr_MQ_NMR_tn_syserr_no_penal_qual_min, par.microadjust_syserr_MQ_NMR_tn_syserr_no_penal_qual_max)
                - par.microadjust_syserr_MQ_NMR_tn_syserr_no_penal_qual_min;
        const int32_t prior_phred = 3;
        int32_t bq4[4], cq4[4];
        tn_quals(bq4, v, (tki_cDP1x + 0.5) / 100.0 + 0.0, (tki_CDP1x + 1.0) / 100.0 + 0.0, c.cVQ1, c.cPCQ1, (0 + 0.5) / 100.0 + 0.0, (0 + 1.0) / 100.0 + 0.0,
                nnminus(0, het3al_inc), par.tn_syserr_norm_devqual, prior_phred, tn_dec_by_xm, par.powlaw_exponent);
        // FORMAT_UNCOV: converted_nfm_cVQ2 = 0 - 3 * (0 + 1) / (0 + 1) = -3
        tn_quals(cq4, v, (tki_cDP2x + 0.5) / 100.0 + 0.0, (tki_CDP2x + 1.0) / 100.0 + 0.0, c.cVQ2, c.cPCQ2, (0 + 0.5) / 100.0 + 0.0, (0 + 1.0) / 100.0 + 0.0,
                nnminus(0, tmax(het3al_inc, 3) - 3), par.tn_syserr_norm_devqual, prior_phred, tmax(tn_dec_by_xm, tmin(tmax(0, -3), 8 + 4)), par.powlaw_exponent);
        const int32_t tlodq1 = tmax(bq4[3], cq4[3]);
        const bool is_CT = ((UVC_BASE_C == refsymbol && UVC_BASE_T == symbol) || (UVC_BASE_G == refsymbol && UVC_BASE_A == symbol));
        const double b_min_tlodq = 2 + 3 - (-10 * log((tki_bDP + 1e-3) / (tki_BDP + 1)) / v.ln10) / 10.0;
        const double c2v_min_tlodq = 2 + 5 - (-10 * log((tki_cDP2x * 0.01 + 1e-5) / (tki_CDP2x * 0.01 + 1) / (is_CT ? 5 : 1)) / v.ln10) / 10.0;
        const float lowestVAQ = (float)dmax(b_min_tlodq, c2v_min_tlodq);
        const int32_t tlodq = ((tlodq1 >= 10) ? tlodq1 : (tlodq1 * 3 - 20));
        const int32_t nlodq = nlodq1;
        const int32_t somaticq = tmin(tlodq, nlodq);
        float vq = ((float)tlodq > lowestVAQ ? (float)tlodq : lowestVAQ);
        if (vq < 10.0f) { const float base = (float)pow(10.0, 0.1); vq = log1pf(powf(base, vq)) / logf(base); }   // calc_non_negative<float>
        const int32_t vad1curr = c.aBQ2, vdp1curr = g.ABQ2[0], vad2curr = tki_bDP, vdp2curr = tki_BDP;
        const bool keep_var = (((vq >= par.vqual)
                || ((vad1curr >= par.vad1 && vdp1curr >= par.vdp1 && (vdp1curr * par.vfa1) <= vad1curr)
                 || (vad2curr >= par.vad2 && vdp2curr >= par.vdp2 && (vdp2curr * par.vfa2) <= vad2curr)))
                && (symbol != refsymbol || (par.should_output_all)));
        const int32_t min_ad = ((symbol == refsymbol) ? par.min_r_ad : par.min_a_ad);
        if (!(keep_var && tki_bDP >= min_ad)) { continue; }
        const int32_t slot = atomic_add_i32(sv.out_cursor, 1);
        if (slot >= sv.out_cap) { continue; }
        VarRec & o = sv.out[slot];
        o.gp = G.gp; o.tile = G.tile; o.refpos = refpos; o.symboltype = type; o.refsymbol = refsymbol; o.cand_index = i; o.pad0 = slot;
        for (int j = i - 1; j >= 0 && C[j].symbol == symbol; j--) {     // earlier alleles of the same symbol (see PrevAllele); order = distance back
            const int32_t ps = atomic_add_i32(sv.out_cursor + 4, 1);
            if (ps >= sv.prev_cap) { continue; }
            PrevAllele & q = sv.prev[ps];
            q.rec_slot = slot; q.order = i - j;
            for (int k = 0; k < 6; k++) { q.nNFA[k] = C[j].nNFA[k]; }
            for (int k = 0; k < 9; k++) { q.nAFA[k] = C[j].nAFA[k]; }
            for (int k = 0; k < 10; k++) { q.nBCFA[k] = C[j].nBCFA[k]; }
            q.fts_mask = C[j].fts_mask;
            for (int k = 0; k < UVC_NUM_FTS; k++) { q.fts_pct[k] = C[j].fts_pct[k]; }
        }
        o.g = g; o.ref = R; o.alt = c;
        o.DP = g.CDP1b[0] + g.CDP1b[1]; o.bDP = g.BDPb[0] + g.BDPb[1]; o.c2DP = g.CDP2b[0] + g.CDP2b[1];
        for (int r = 0; r < 2; r++) {
            o.cVQ1M[r] = (top[r] >= 0 ? C[top[r]].cVQ1 : (r == 0 ? -999 : 0));
            o.cVQ2M[r] = (top[r] >= 0 ? C[top[r]].cVQ2 : (r == 0 ? -999 : 0));
            o.cVQAM[r] = (top[r] >= 0 ? C[top[r]].symbol : (r == 0 ? UVC_NSYM : -1));
            o.cVQSM_ev[r] = (top[r] >= 0 ? C[top[r]].ev : -1);
        }
        o.vHGQ = nlodq1; o.vAC[0] = curr_vAC[0]; o.vAC[1] = curr_vAC[1];
        o.vNLODQ[0] = (type == 0 ? nlodq_site : 0); o.vNLODQ[1] = (type == 1 ? nlodq_site : 0);
        o.tlodq = tlodq; o.nlodq = nlodq; o.somaticq = somaticq;
        for (int r = 0; r < 4; r++) { o.TNBQF[r] = bq4[r]; o.TNCQF[r] = cq4[r]; }
        o.tbDP = tki_BDP; o.tDP = o.DP; o.tAD[0] = R.AD; o.tAD[1] = c.AD;
        o.t2DP = (g.CDPDb[0] + g.CDPDb[1]) + (g.DDP2[0] + g.DDP2[1]);
        o.t2AD[0] = R.cDPDf + R.cDPDr + R.dDP2;
        o.t2AD[1] = c.cDPDf + c.cDPDr + c.dDP2;   // indel alleles: replaced on the host by the allele's own gc2dAD sum
        o.vcfqual = vq; o.lowestVAQ = lowestVAQ;
        o.repeatnum = G.repeatnum; o.repeatunit_len = G.repeatunit_len;
        const int32_t adj = par.indel_adj_tracklen_dist;
        const uvcgpu_rtr & q1 = rtr[tmax(refpos - T.ext_beg, adj) - adj];
        const uvcgpu_rtr & q2 = rtr[tmin(refpos - T.ext_beg + adj, nrtr - adj)];   // QUIRK: size - dist, not size - 1 (main.hpp:6101)
        o.rtr_info[0] = ((0 == q1.tracklen) ? 0 : (T.ext_beg + q1.begpos)); o.rtr_info[1] = q1.tracklen; o.rtr_info[2] = q1.unitlen;
        o.rtr_info[3] = ((0 == q2.tracklen) ? 0 : (T.ext_beg + q2.begpos)); o.rtr_info[4] = q2.tracklen; o.rtr_info[5] = q2.unitlen;
    }
}

// ------------------------------------------------------------------------------------------------ K6: one thread per position
// Inputs of the MGVCF block lines (main.cpp:667-721) and of the additional-indel-candidate lines, for every position of every tile.
UVC_HD void k6_gvcf_position(const BatchView & v, const ScoreView & sv, int64_t gp) {
    const int32_t ti = v.pos_tile[gp];
    const TileInfo & T = v.tiles[ti];
    GvcfPos & o = sv.gvcf[gp];
    GvcfExtra & x = sv.gextra[gp];
    if (T.skipped) { return; }
    const uvcgpu_params & par = v.par;
    const int32_t nref = (T.ext_end - T.ext_beg) - 1;
    const int32_t off = (int32_t)(gp - T.pos_off);
    const PosPtrs P = pos_ptrs(v, gp);
    const int base_m = ((off < nref) ? (int)v.refsym[gp] : UVC_BASE_N);
    for (int k = 0; k < 2; k++) {
        const int type = (k == 0 ? 1 : 0);
        const int s0 = (type == 0 ? UVC_BASE_A : UVC_LINK_M), s1 = (type == 0 ? UVC_BASE_NN : UVC_LINK_NN);
        const int refsymbol = (type == 0 ? base_m : UVC_LINK_M);
        int32_t b = 0, c = 0, c12 = 0;
        for (int s = s0; s <= s1; s++) {
            b += P.fd0[s * UVCGPU_NUM_FRAG_DEPTHS + 0] + P.fd1[s * UVCGPU_NUM_FRAG_DEPTHS + 0];
            c += P.fm0[s * UVCGPU_NUM_FAM_DEPTHS + 0] + P.fm1[s * UVCGPU_NUM_FAM_DEPTHS + 0];
            c12 += P.fm0[s * UVCGPU_NUM_FAM_DEPTHS + 1] + P.fm1[s * UVCGPU_NUM_FAM_DEPTHS + 1];
        }
        const int32_t ref_cdepth = P.fm0[refsymbol * UVCGPU_NUM_FAM_DEPTHS + 1] + P.fm1[refsymbol * UVCGPU_NUM_FAM_DEPTHS + 1];
        const int32_t nonref_cdepth = c12 - ref_cdepth;
        const double ref_like_binom = -binom_10log10_likeratio(v, par.contam_any_mul_frac, nonref_cdepth + 0.5, c + 1.0);
        const double ref_like_powlaw = -dmax(0, par.powlaw_exponent * (10 / v.ln10) * logit2((nonref_cdepth + 0.5) / (c + 1.0), par.contam_any_mul_frac));
        const double nonref_like_binom = -binom_10log10_likeratio(v, par.germ_hetero_FA, ref_cdepth + 0.5, c + 1.0);
        const double nonref_like_powlaw = -dmax(0, par.powlaw_exponent * (10 / v.ln10) * logit2((ref_cdepth + 0.5) / (c + 1.0), par.germ_hetero_FA));
        o.bdepth[k] = b; o.cdepth[k] = c; o.cdep12[k] = c12;
        o.refQ[k] = par.germ_phred_hetero_snp + d2i(round(dmax(ref_like_binom, ref_like_powlaw) - d2i(round(dmax(nonref_like_binom, nonref_like_powlaw)))));
    }
    int32_t unit = 0, num = 0;
    repeat_at(unit, num, v, T, off);
    x.unitlen = unit; x.repeatnum = num; x.tracklen = unit * num;
    x.a_dp = P.prep->a_dp; x.a_clip = P.prep->a_near_long_clip_dp;
}

} // namespace uvc

#endif
