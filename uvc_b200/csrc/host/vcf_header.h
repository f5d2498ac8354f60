// vcf_header.h - VCF header text of the uvc1 host (replaces generate_vcf_header, reference main.hpp:5777-5877).
#ifndef UVC_VCF_HEADER_H_INCLUDED
#define UVC_VCF_HEADER_H_INCLUDED

#include "../../../include/uvcgpu.h"

#include <string>
#include <utility>
#include <vector>

std::string uvc_vcf_header(int argc, const char *const *argv, const std::vector<std::pair<std::string, int64_t>> & contigs, const std::string & fasta_ref_fname,
        const std::string & sample_name, const uvcgpu_params & par, const std::string & version);

#endif
