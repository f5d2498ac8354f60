// vcf_emit.h - host side of the scoring stage: candidate indel alleles for the device (indel_get_majority), and the VCF text of a tile
// from the device's records (append_vcf_record + streamAppendBcfFormat, MGVCF block lines, ADDITIONAL_INDEL_CANDIDATE lines).
#ifndef UVC_VCF_EMIT_H_INCLUDED
#define UVC_VCF_EMIT_H_INCLUDED

#include "host_prep.h"
#include "score_core.cuh"
#include "sparse_out.h"

#include <map>
#include <string>
#include <utility>
#include <vector>

// gap* FORMAT tags of one (position, indel symbol) and its candidate alleles (fill_by_indel_info + indel_get_majority, instcode.hpp, main.hpp:5405-5455)
struct IndelSite {
    std::vector<int32_t> gapNf, gapNr;
    std::vector<std::string> gapSeq;
    std::vector<int32_t> gapbAD1, gapcAD1, gc2AD, gc2dAD;
    struct Allele { int32_t bAD, cAD, ev; std::string seq; };
    std::vector<Allele> alleles;
};

typedef std::map<std::pair<int32_t, int32_t>, IndelSite> TileIndelSites;   // (refpos, symbol) -> site

// Builds, for every tile, the indel sites and the flat device table (sorted by key = gp * 16 + symbol).
void uvc_build_indel_sites(std::vector<TileIndelSites> & sites, std::vector<IndelAllele> & table, const HostBatch & hb,
        const std::vector<TileSparse> & sparse, const std::map<int32_t, HostContig> & contigs, const StageVec<IndelEvent> & ev);

// Shared state of a tile's text (reference string, haplotype indices, records by position), and the text of a range of its positions: a deep
// tile is formatted by several threads (uvcgpu_tile_vcf / uvcgpu_batch_vcf), each on its own range; the concatenation is the tile's text.
struct TileTextPlan;
TileTextPlan *uvc_tile_text_plan_new(const HostBatch & hb, int32_t tile_index, const HostContig & contig, const std::vector<const VarRec*> & recs, const TileSparse & sparse);
void uvc_tile_text_plan_free(TileTextPlan *plan);
std::string uvc_tile_vcf_text_range(const TileTextPlan & plan, const HostBatch & hb, int32_t tile_index, const uvcgpu_params & par, const std::string & tname,
        const TileIndelSites & sites, const TileSparse & sparse, const StageVec<IndelEvent> & ev, const GvcfPos *gvcf, const GvcfExtra *gextra,
        int32_t zb_begin, int32_t zb_end, const std::vector<PrevAllele> *prev_alleles = NULL);

// The uncompressed VCF fragment of one tile (what process_batch appends to uncompressed_vcf_string, main.cpp:1184).
std::string uvc_tile_vcf_text(const HostBatch & hb, int32_t tile_index, const uvcgpu_params & par, const std::string & tname, const HostContig & contig,
        const std::vector<const VarRec*> & recs_of_tile, const TileIndelSites & sites, const TileSparse & sparse, const StageVec<IndelEvent> & ev,
        const GvcfPos *gvcf, const GvcfExtra *gextra);

#endif
