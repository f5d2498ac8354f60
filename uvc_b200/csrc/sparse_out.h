// sparse_out.h - host-side assembly of the sparse outputs of a batch from the device record stream:
// the indel-identity maps (the reference's map<pos, map<inserted sequence | deletion length, count>> members of
// CoveredRegion and Symbol2CountCoverageSet, main.hpp:529-530, 2380-2383) and the haplotype links (updateHapMap, main.hpp:3596-3663).
#ifndef UVC_SPARSE_OUT_H_INCLUDED
#define UVC_SPARSE_OUT_H_INCLUDED

#include "batch.h"
#include "host_prep.h"

#include <array>
#include <map>
#include <string>
#include <utility>
#include <vector>

struct IndelKey {
    int32_t kind, strand, symbol, pos;   // kind = UVC_REC_*_INDEL
    bool operator<(const IndelKey & o) const {
        if (kind != o.kind) { return kind < o.kind; }
        if (strand != o.strand) { return strand < o.strand; }
        if (symbol != o.symbol) { return symbol < o.symbol; }
        return pos < o.pos;
    }
};

struct HapLinkOut {
    std::vector<std::pair<int32_t, int32_t>> pos_symb;   // (position, symbol)
    std::array<int32_t, 2> fr_cnts;
    std::array<int32_t, 2> other_hap_cnts;
};

struct IdCount { int32_t count = 0; int32_t ev = -1; };   // count of an indel identity and one representative device event

struct TileSparse {
    // insertions keyed by inserted sequence, deletions keyed by length (kept separately like the reference)
    std::map<IndelKey, std::map<std::string, IdCount>> ins;
    std::map<IndelKey, std::map<int32_t, IdCount>> del;
    std::vector<HapLinkOut> hap_bq, hap_fq, hap_f2q;
};

// rec: the downloaded record stream (n_words valid words); ev: the downloaded indel events.
void uvc_build_sparse(std::vector<TileSparse> & out, const HostBatch & hb, const uvcgpu_params & par,
        const int32_t *rec, int64_t n_words, const IndelEvent *ev);

std::string uvc_indelmaps_text(const TileSparse & ts);
std::string uvc_haplinks_text(const TileSparse & ts);

#endif
