// vcf_header.cpp - the VCF 4.2 header of the uvc1 host. Same line order and the same IDs / Number / Type as the reference's
// generate_vcf_header (main.hpp:5777-5877) so that any parser of the reference's output reads ours; Description texts are this project's own
// wording (the reference's prose is not reproduced), and fileDate / variantCallerVersion / variantCallerCommand differ by nature.
#include "vcf_header.h"

#include "vcf_header_table.inc"

#include <time.h>

namespace {

std::string info(const char *id, const char *number, const char *type, const char *desc) {
    return std::string("##INFO=<ID=") + id + ",Number=" + number + ",Type=" + type + ",Description=\"" + desc + "\">\n";
}

} // namespace

std::string uvc_vcf_header(int argc, const char *const *argv, const std::vector<std::pair<std::string, int64_t>> & contigs, const std::string & fasta_ref_fname,
        const std::string & sample_name, const uvcgpu_params & par, const std::string & version) {
    time_t rawtime;
    time(&rawtime);
    char timestring[80];
    strftime(timestring, 80, "%F %T", localtime(&rawtime));
    std::string h;
    h += "##fileformat=VCFv4.2\n";
    h += std::string("##fileDate=") + timestring + "\n";
    h += "##reference=" + fasta_ref_fname + "\n";
    for (const auto & c : contigs) { h += "##contig=<ID=" + c.first + ",length=" + std::to_string(c.second) + ">\n"; }
    h += "##ALT=<ID=NON_REF,Description=\"Any possible alternative allele at this location (one-based inclusive POS). This record is similar to, but not, a GVCF block.\">\n";
    for (const char *id : UVC_FILTER_IDS) {
        const bool is_q = (id[0] == 'Q' && id[1] >= '1' && id[1] <= '6');
        h += std::string("##FILTER=<ID=") + id + ",Description=\"" + (is_q ? "Variant quality below the number in the ID" : "UVC filter (for FORMAT/FTS when it names a bias): same meaning as in UVC 0.15.1")
           + ".\">\n";
    }
    h += info("ANY_VAR", "0", "Flag", "Variant of any origin (germline polymorphism and/or somatic mutation)");
    h += info("GERMLINE", "0", "Flag", "Germline variant");
    h += info("SOMATIC", "0", "Flag", "Somatic variant");
    h += info("MGVCF_BLOCK", "0", "Flag", "Multi-sample-GVCF-like block of 1000 consecutive positions, detailed in FORMAT/POS_VT_BDP_CDP_HomRefQ");
    h += info("ADDITIONAL_INDEL_CANDIDATE", "0", "Flag", "Position next to an abnormally high number of clipped sequences or followed by a long STR track");
    h += info("SomaticQ", "A", "Float", "Phred-scaled odds that the variant is not somatic");
    h += info("TLODQ", "A", "Float", "Tumor log-of-data-likelihood quality: Phred-scaled odds that the variant is an artifact");
    h += info("NLODQ", "A", "Float", "Normal log-of-data-likelihood quality: Phred-scaled odds that the variant is of germline origin");
    h += info("NLODV", "A", "String", "The variant symbol that minimizes NLODQ");
    h += info("TNBQF", "4", "Float", "Binomial reward, power-law reward, systematic-error penalty and normal-adjusted tumor quality from deduplicated fragments");
    h += info("TNCQF", "4", "Float", "Binomial reward, power-law reward, systematic-error penalty and normal-adjusted tumor quality from consensus families");
    h += info("tbDP", "1", "Integer", "Tumor total non-deduplicated depth");
    h += info("tDP", "1", "Integer", "Tumor total deduplicated depth");
    h += info("tAD", "R", "Integer", "Tumor deduplicated depth of each allele");
    h += info("t2DP", "1", "Integer", "Tumor total UMI-family depth for duplex-rescued SSCS");
    h += info("t2AD", "R", "Integer", "Tumor UMI-family depth of each allele for duplex-rescued SSCS");
    h += info("nDP", "1", "Integer", "Normal total deduplicated depth");
    h += info("nAD", "R", "Integer", "Normal deduplicated depth of each allele");
    h += info("n2AD", "R", "Integer", "Normal UMI-family depth of each allele");
    h += info("RU", "1", "String", "The shortest repeating unit in the reference");
    h += info("RC", "1", "Integer", "The number of non-interrupted RUs in the reference");
    h += info("R3X2", "6", "Integer", "Repeat start, track length and unit size at the two positions before and after this position");
    for (const auto & t : UVC_FORMAT_TAGS) {
        const bool sub = (t.id[0] == '_');
        h += std::string("##FORMAT=<ID=") + t.id + ",Number=" + t.number + ",Type=" + t.type + ",Description=\""
           + (sub ? "Sub-header that separates groups of FORMAT tags" : "Same definition as FORMAT/" + std::string(t.id) + " of UVC 0.15.1") + ".\">\n";
    }
    h += "##FORMAT=<ID=GL4,Number=4,Type=Integer,Description=\"The four genotype likelihoods for 0/0, 0/1, 1/1, and 1/2\">\n";
    h += "##FORMAT=<ID=GST,Number=.,Type=Integer,Description=\"The genotype statistics\">\n";
    h += "##FORMAT=<ID=CDP1,Number=2,Type=Integer,Description=\"(CDP1f + CDP1r) for all alleles by sum and for the padded deletion allele\">\n";
    h += "##FORMAT=<ID=cDP1,Number=2,Type=Integer,Description=\"(cDP1f + cDP1r)\">\n";
    h += "##FORMAT=<ID=POS_VT_BDP_CDP_HomRefQ,Number=.,Type=Integer,Description=\"Regions of one MGVCF line as ((<pos>,<postype>,<.>,<dup>,<dedup>,<dedupBQ>,<homrefQ>,<.>)+<endpos>): "
         "<pos> separates adjacent regions, <postype> is 1 for the SNV and 2 for the InDel sub-position, <dup>/<dedup>/<dedupBQ> are minimum fragment depths without deduplication, "
         "with deduplication, and with deduplication and the base-quality threshold, <homrefQ> is the minimum homozygous-reference likelihood with the SNV prior "
         "(add " + std::to_string(par.germ_phred_hetero_indel - par.germ_phred_hetero_snp) + " for InDels), <endpos> ends the last region.\">\n";
    h += "##FORMAT=<ID=clipDP,Number=2,Type=Integer,Description=\"Total segment depth and segment depth with adjacent long clips (for the <ADDITIONAL_INDEL_CANDIDATE> symbolic allele)\">\n";
    h += "##phasing=partial\n";
    h += "##variantCallerVersion=" + version + "\n";
    h += "##variantCallerCommand=";
    for (int i = 0; i < argc; i++) { h += std::string(argv[i]) + "  "; }
    h += "\n";
    h += "##variantCallerInferredParameters=(inferred_sequencing_platform=" + std::string(par.inferred_sequencing_platform == 1 ? "Illumina/BGI" : "IonTorrent/LifeTechnologies/ThermoFisher")
       + ",central_readlen=" + std::to_string(par.central_readlen) + ")\n";
    h += "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + sample_name + "\n";
    return h;
}
