/* Instrumented oracle harness (TEST INFRASTRUCTURE, never linked into the product).
 *
 * Textually includes the reference's main.cpp (found through -I$(REF); nothing is copied into this
 * repository) with its main() renamed, then drives the reference's OWN functions for one tier-3
 * region the way process_batch does up to the end of the pileup (reference main.cpp:481-591):
 *   bamfname_to_strand_to_familyuid_to_reads -> fill_strand_umi_readset_with_strand_to_umi_to_reads
 *   -> load_refstring -> refstring2repeatvec -> region_repeatvec_to_baq_offsetarr x2
 *   -> Symbol2CountCoverageSet::updateByRegion3Aln
 * and serialises every per-position counter array of Symbol2CountCoverageSet (members main.hpp:2368-2383)
 * in the reference's own struct layout, plus the family grouping, the indel maps and the haplotype links.
 * The CUDA path's uvcgpu_dump_counters output is compared with this file byte for byte.
 *
 * usage: uvc_ref_dump <bam> <fasta> <tid> <beg> <end> <region_flag> <out.bin> <prev_tid> <prev_beg> <prev_end> [uvc1 options...]
 */
#define main uvc_reference_main
#include "main.cpp"
#undef main

#include <cstdio>
#include <cstring>

namespace {

struct SectionWriter {
    FILE *f;
    explicit SectionWriter(const char *path) : f(fopen(path, "wb")) {
        if (NULL == f) { fprintf(stderr, "cannot open %s for writing\n", path); exit(2); }
        fwrite("UVCDUMP1", 1, 8, f);
    }
    void put(const char *name, const void *data, uint64_t elem_size, uint64_t count) {
        char nm[32];
        memset(nm, 0, sizeof(nm));
        strncpy(nm, name, 31);
        fwrite(nm, 1, 32, f);
        fwrite(&elem_size, 8, 1, f);
        fwrite(&count, 8, 1, f);
        if (elem_size * count > 0) { fwrite(data, elem_size, count, f); }
    }
    void put_text(const char *name, const std::string & s) { put(name, s.data(), 1, s.size()); }
    ~SectionWriter() { fclose(f); }
};

template <class TCov>
void dump_cov(SectionWriter & w, const char *name, const TCov & cov, uvc1_refgpos_t beg, uvc1_refgpos_t end) {
    typedef typename std::remove_cv<typename std::remove_reference<decltype(cov.getByPos(beg))>::type>::type elem_t;
    std::vector<elem_t> buf;
    buf.reserve(end - beg);
    for (uvc1_refgpos_t p = beg; p < end; p++) { buf.push_back(cov.getByPos(p)); }
    w.put(name, buf.data(), sizeof(elem_t), buf.size());
}

template <class TMap>
void indelmap_to_text(std::string & out, const char *label, int strand, int idx, const TMap & m) {
    for (const auto & pos_kv : m) {
        for (const auto & indel_cnt : pos_kv.second) {
            std::ostringstream oss;
            oss << label << "\t" << strand << "\t" << idx << "\t" << pos_kv.first << "\t" << indel_cnt.first << "\t" << indel_cnt.second << "\n";
            out += oss.str();
        }
    }
}

void haplinks_to_text(std::string & out, const char *label, const std::vector<HapLink> & v) {
    for (const auto & h : v) {
        std::ostringstream oss;
        oss << label << "\t" << h.fr_cnts[0] << "\t" << h.fr_cnts[1] << "\t" << h.other_hap_cnts[0] << "\t" << h.other_hap_cnts[1] << "\t";
        for (const auto & ps : h.pos_symb_string) { oss << ps.first << ":" << (int)ps.second << ","; }
        oss << "\n";
        out += oss.str();
    }
}

} // namespace

int main(int argc, char **argv) {
    if (argc < 11) {
        fprintf(stderr, "usage: %s <bam> <fasta> <tid> <beg> <end> <region_flag> <out.bin> <prev_tid> <prev_beg> <prev_end> [uvc1 options...]\n", argv[0]);
        return 1;
    }
    const char *bamfn = argv[1];
    const char *fafn = argv[2];
    const uvc1_refgpos_t tid = atoi(argv[3]);
    const uvc1_refgpos_t beg = atoi(argv[4]);
    const uvc1_refgpos_t end = atoi(argv[5]);
    const uvc1_flag_t region_flag = (uvc1_flag_t)atoi(argv[6]);
    const char *outfn = argv[7];
    const BedLine prev_bedline(atoi(argv[8]), atoi(argv[9]), atoi(argv[10]), 0, 0);
    const BedLine bedline(tid, beg, end, region_flag, 0);

    /* the reference's own option parser + data-driven inference (CmdLineArgs.cpp:174, 1003-1035) */
    std::vector<std::string> args = { "uvc1", bamfn, "-f", fafn, "-o", "/dev/null" };
    for (int i = 11; i < argc; i++) { args.push_back(argv[i]); }
    std::vector<char*> cargs;
    for (auto & a : args) { cargs.push_back(&a[0]); }
    CommandLineArgs paramset;
    int parsing_result_flag = -1;
    int parsing_result_ret = paramset.initFromArgCV(parsing_result_flag, (int)cargs.size(), cargs.data());
    if (parsing_result_ret || parsing_result_flag) { fprintf(stderr, "option parsing failed\n"); return 3; }
    const char *UMI_STRUCT = getenv("ONE_STEP_UMI_STRUCT");
    const std::string UMI_STRUCT_STRING = ((UMI_STRUCT != NULL && strlen(UMI_STRUCT) > 0) ? std::string(UMI_STRUCT) : std::string(""));

    std::vector<std::tuple<std::string, uvc1_refgpos_t>> tid_to_tname_tseqlen_tuple_vec;
    samfname_to_tid_to_tname_tseq_tup_vec(tid_to_tname_tseqlen_tuple_vec, paramset.bam_input_fname);
    samFile *samfile = sam_open(bamfn, "r");
    hts_idx_t *hts_idx = sam_index_load(samfile, bamfn);
    faidx_t *ref_faidx = fai_load(fafn);
    if (NULL == samfile || NULL == hts_idx || NULL == ref_faidx) { fprintf(stderr, "failed to open inputs\n"); return 4; }
    const auto tname_tseqlen_tuple = tid_to_tname_tseqlen_tuple_vec.at(tid);

    std::map<MolecularBarcode, std::pair<std::array<std::map<uvc1_hash_t, std::vector<bam1_t *>>, 2>, MolecularBarcode>> umi_to_strand_to_reads;
    uvc1_refgpos_t bam_inclu_beg_pos, bam_exclu_end_pos;
    std::vector<std::pair<std::array<std::vector<std::vector<bam1_t *>>, 2>, MolecularBarcode>> umi_strand_readset;
    const bool end2end = (bedline.region_flag & BED_END_TO_END_BIT);
    std::array<uvc1_readnum_big_t, 3> passed_pcrpassed_umipassed = bamfname_to_strand_to_familyuid_to_reads(
            umi_to_strand_to_reads, bam_inclu_beg_pos, bam_exclu_end_pos,
            tid, beg, end, end2end, 0, 1, UMI_STRUCT_STRING, samfile, hts_idx, 0, paramset, 0);
    const auto num_passed_reads = passed_pcrpassed_umipassed[0];
    const auto num_pcrpassed_reads = passed_pcrpassed_umipassed[1];
    fill_strand_umi_readset_with_strand_to_umi_to_reads(umi_strand_readset, umi_to_strand_to_reads, paramset, 0);

    SectionWriter w(outfn);
    int64_t meta[16];
    memset(meta, 0, sizeof(meta));
    meta[0] = num_passed_reads;
    meta[1] = num_pcrpassed_reads;
    meta[2] = bam_inclu_beg_pos;
    meta[3] = bam_exclu_end_pos;
    meta[10] = paramset.central_readlen;
    meta[11] = paramset.inferred_maxMQ;
    meta[12] = (int64_t)paramset.inferred_sequencing_platform;
    if ((0 == num_passed_reads) || (-1 == num_passed_reads)) {
        w.put("meta", meta, 8, 16);
        return 0;
    }

    const uvc1_refgpos_t rpos_inclu_beg = MAX(beg, bam_inclu_beg_pos);
    const uvc1_refgpos_t rpos_exclu_end = MIN(end, bam_exclu_end_pos);
    const uvc1_refgpos_t extended_inclu_beg_pos = MAX(0, non_neg_minus(MIN(beg, bam_inclu_beg_pos), MAX_STR_N_BASES));
    const uvc1_refgpos_t extended_exclu_end_pos = MIN(std::get<1>(tname_tseqlen_tuple), MAX(end, bam_exclu_end_pos) + MAX_STR_N_BASES);
    meta[4] = rpos_inclu_beg;
    meta[5] = rpos_exclu_end;
    meta[6] = extended_inclu_beg_pos;
    meta[7] = extended_exclu_end_pos;
    meta[8] = (int64_t)umi_strand_readset.size();
    w.put("meta", meta, 8, 16);

    /* family grouping as text: key fields, then per strand the fragments (qname_hash2 order) with their reads */
    {
        std::string fams;
        for (const auto & fam : umi_strand_readset) {
            const MolecularBarcode & mb = fam.second;
            std::ostringstream oss;
            oss << "F\t" << mb.beg_tidpos_pair.first << "\t" << mb.beg_tidpos_pair.second << "\t" << mb.end_tidpos_pair.first << "\t" << mb.end_tidpos_pair.second
                << "\t" << mb.duplexflag << "\t" << mb.dedup_idflag << "\t" << mb.umistring << "\t" << fam.first[0].size() << "\t" << fam.first[1].size() << "\n";
            for (int strand = 0; strand < 2; strand++) {
                for (const auto & alns1 : fam.first[strand]) {
                    oss << "f\t" << strand;
                    for (const bam1_t *aln : alns1) { oss << "\t" << bam_get_qname(aln) << "/" << aln->core.flag << "/" << aln->core.pos; }
                    oss << "\n";
                }
            }
            fams += oss.str();
        }
        w.put_text("families", fams);
    }

    const std::string refstring = load_refstring(ref_faidx, tid, extended_inclu_beg_pos, extended_exclu_end_pos);
    std::vector<RegionalTandemRepeat> region_repeatvec = refstring2repeatvec(
            refstring, paramset.indel_str_repeatsize_max, paramset.indel_vntr_repeatsize_max, paramset.indel_BQ_max,
            paramset.indel_polymerase_slip_rate, paramset.indel_del_to_ins_err_ratio, 0);
    w.put("rtr_initial", region_repeatvec.data(), sizeof(RegionalTandemRepeat), region_repeatvec.size());
    const auto & baq_offsetarr = region_repeatvec_to_baq_offsetarr(region_repeatvec, tid, extended_inclu_beg_pos, extended_exclu_end_pos + 1, paramset);
    const auto & baq_offsetarr2 = region_repeatvec_to_baq_offsetarr<true>(region_repeatvec, tid, extended_inclu_beg_pos, extended_exclu_end_pos + 1, paramset);

    Symbol2CountCoverageSet cset(tid, extended_inclu_beg_pos, extended_exclu_end_pos + 1);
    std::vector<HapLink> hap_bq, hap_fq, hap_f2q;
    std::array<std::string, NUM_FQLIKE_CON_OUT_FILES> fqdata3;
    cset.updateByRegion3Aln(fqdata3, hap_bq, hap_fq, hap_f2q, umi_strand_readset, refstring, region_repeatvec,
            baq_offsetarr, baq_offsetarr2, prev_bedline, bedline, paramset, 0);

    const uvc1_refgpos_t b = extended_inclu_beg_pos, e = extended_exclu_end_pos + 1;
    w.put_text("refstring", refstring);
    w.put("rtr_final", region_repeatvec.data(), sizeof(RegionalTandemRepeat), region_repeatvec.size());
    dump_cov(w, "baq", baq_offsetarr, b, e);
    dump_cov(w, "baq2", baq_offsetarr2, b, e);
    dump_cov(w, "prep", cset.seg_format_prep_sets, b, e);
    dump_cov(w, "thres", cset.seg_format_thres_sets, b, e);
    dump_cov(w, "seginfo", cset.symbol_to_seg_format_info_sets, b, e);
    dump_cov(w, "faminfo", cset.symbol_to_fam_format_info_sets, b, e);
    dump_cov(w, "fragdepth0", cset.symbol_to_frag_format_depth_sets[0], b, e);
    dump_cov(w, "fragdepth1", cset.symbol_to_frag_format_depth_sets[1], b, e);
    dump_cov(w, "famdepth0", cset.symbol_to_fam_format_depth_sets_2strand[0], b, e);
    dump_cov(w, "famdepth1", cset.symbol_to_fam_format_depth_sets_2strand[1], b, e);
    dump_cov(w, "duplex", cset.symbol_to_duplex_format_depth_sets, b, e);
    dump_cov(w, "vq", cset.symbol_to_VQ_format_tag_sets, b, e);

    {
        std::string t;
        for (int strand = 0; strand < 2; strand++) {
            for (const AlignmentSymbol s : INS_SYMBOLS) {
                indelmap_to_text(t, "frag_ins", strand, (int)s, cset.symbol_to_frag_format_depth_sets[strand].getPosToIseqToData(s));
                indelmap_to_text(t, "fam_ins", strand, (int)s, cset.symbol_to_fam_format_depth_sets_2strand[strand].getPosToIseqToData(s));
                indelmap_to_text(t, "cDP2_ins", strand, (int)s, cset.pos2iseq2data_cDP2[strand][insSymbolToInsIdx(s)]);
                indelmap_to_text(t, "c2dDP_ins", strand, (int)s, cset.pos2iseq2data_c2dDP[strand][insSymbolToInsIdx(s)]);
            }
            for (const AlignmentSymbol s : DEL_SYMBOLS) {
                indelmap_to_text(t, "frag_del", strand, (int)s, cset.symbol_to_frag_format_depth_sets[strand].getPosToDlenToData(s));
                indelmap_to_text(t, "fam_del", strand, (int)s, cset.symbol_to_fam_format_depth_sets_2strand[strand].getPosToDlenToData(s));
                indelmap_to_text(t, "cDP2_del", strand, (int)s, cset.pos2dlen2data_cDP2[strand][delSymbolToDelIdx(s)]);
                indelmap_to_text(t, "c2dDP_del", strand, (int)s, cset.pos2dlen2data_c2dDP[strand][delSymbolToDelIdx(s)]);
            }
        }
        w.put_text("indelmaps", t);
    }
    {
        std::string t;
        haplinks_to_text(t, "bq", hap_bq);
        haplinks_to_text(t, "fq", hap_fq);
        haplinks_to_text(t, "f2q", hap_f2q);
        w.put_text("haplinks", t);
    }
    umi_strand_readset_uvc_destroy(umi_strand_readset);
    fai_destroy(ref_faidx);
    hts_idx_destroy(hts_idx);
    sam_close(samfile);
    return 0;
}
