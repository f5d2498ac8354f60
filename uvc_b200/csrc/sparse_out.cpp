// sparse_out.cpp - see sparse_out.h
#include "sparse_out.h"

#include <algorithm>
#include <tuple>

namespace {

typedef std::vector<std::pair<int32_t, int32_t>> mutform_t;

// updateHapMap (main.hpp:3596-3663)
std::vector<HapLinkOut> hap_links(const std::map<mutform_t, std::array<int32_t, 2>> & m, const TileInfo & T, const uvcgpu_params & par) {
    std::vector<HapLinkOut> ret;
    std::vector<std::tuple<int32_t, mutform_t, std::array<int32_t, 2>>> v;
    for (const auto & kv : m) { v.push_back(std::make_tuple(kv.second[0] + kv.second[1], kv.first, kv.second)); }
    std::sort(v.rbegin(), v.rend());
    const size_t num_dst = std::min((size_t)par.phasing_haplotype_max_detail_cnt, v.size());
    std::vector<int32_t> inc_fw(num_dst, 0), inc_rv(num_dst, 0);
    for (size_t i = 0; i < num_dst; i++) {
        const mutform_t & dst = std::get<1>(v[i]);
        for (size_t j = i + 1; j < v.size(); j++) {
            const mutform_t & src = std::get<1>(v[j]);
            bool skipped = false;
            for (const auto & allele : dst) { if (std::find(src.begin(), src.end(), allele) == src.end()) { skipped = true; break; } }
            if (!skipped) { inc_fw[i] += std::get<2>(v[j])[0]; inc_rv[i] += std::get<2>(v[j])[1]; }
        }
    }
    std::vector<int32_t> used((size_t)std::max(1, T.ext_end - T.ext_beg), 0);
    for (size_t i = 0; i < v.size(); i++) {
        const mutform_t & mf = std::get<1>(v[i]);
        const auto & counts = std::get<2>(v[i]);
        if ((counts[0] + counts[1]) < (par.phasing_haplotype_min_ad + (int32_t)mf.size())) { continue; }
        int32_t haplo_totDP = 0;
        for (const auto & sm : mf) { used[sm.first - T.ext_beg] += 1; haplo_totDP += used[sm.first - T.ext_beg]; }
        if (haplo_totDP > (int64_t)par.phasing_haplotype_max_count * (int64_t)mf.size()) { continue; }
        HapLinkOut h;
        h.pos_symb = mf;
        h.fr_cnts = counts;
        if (i >= num_dst) { h.other_hap_cnts = {{-1, -1}}; } else { h.other_hap_cnts = {{inc_fw[i], inc_rv[i]}}; }
        ret.push_back(h);
    }
    return ret;
}

} // namespace

void uvc_build_sparse(std::vector<TileSparse> & out, const HostBatch & hb, const uvcgpu_params & par,
        const int32_t *rec, int64_t n_words, const IndelEvent *ev) {
    out.assign(hb.tiles.size(), TileSparse());
    std::vector<std::map<mutform_t, std::array<int32_t, 2>>> hbq(hb.tiles.size()), hfq(hb.tiles.size()), hf2q(hb.tiles.size());
    static const char *nt16 = "=ACMGRSVTWYHKDBN";
    int64_t o = 0;
    while (o < n_words) {
        const int32_t kind = rec[o];
        if (kind >= UVC_REC_FRAG_INDEL && kind <= UVC_REC_C2D_INDEL) {
            if (o + 6 > n_words) { break; }
            const IndelEvent & E = ev[rec[o + 4]];
            IndelKey k; k.kind = kind; k.strand = rec[o + 1]; k.symbol = rec[o + 2]; k.pos = rec[o + 3];
            TileSparse & ts = out[E.tile];
            if (E.is_del) {
                IdCount & ic = ts.del[k][E.oplen]; ic.count += rec[o + 5]; if (ic.ev < 0) { ic.ev = rec[o + 4]; }
            } else {
                std::string seq;
                const uint8_t *s = hb.raw_seq(E.raw);
                for (int32_t i = 0; i < E.oplen; i++) { const int32_t q = E.qpos + i; seq.push_back(nt16[(s[q >> 1] >> ((~q & 1) << 2)) & 0xf]); }
                IdCount & ic = ts.ins[k][seq]; ic.count += rec[o + 5]; if (ic.ev < 0) { ic.ev = rec[o + 4]; }
            }
            o += 6;
        } else if (kind >= UVC_REC_HAP_BQ && kind <= UVC_REC_HAP_F2Q) {
            if (o + 4 > n_words) { break; }
            const int32_t strand = rec[o + 1], n = rec[o + 2], tile = rec[o + 3];
            if (o + 4 + 2 * (int64_t)n > n_words) { break; }
            mutform_t mf;
            for (int32_t i = 0; i < n; i++) { mf.push_back(std::make_pair(rec[o + 4 + 2 * i], rec[o + 5 + 2 * i])); }
            auto & m = (kind == UVC_REC_HAP_BQ ? hbq[tile] : (kind == UVC_REC_HAP_FQ ? hfq[tile] : hf2q[tile]));
            auto it = m.insert(std::make_pair(mf, std::array<int32_t, 2>{{0, 0}})).first;
            it->second[strand] += 1;
            o += 4 + 2 * (int64_t)n;
        } else {
            break; // zero padding / unknown: end of stream
        }
    }
    for (size_t t = 0; t < hb.tiles.size(); t++) {
        out[t].hap_bq = hap_links(hbq[t], hb.tiles[t], par);
        out[t].hap_fq = hap_links(hfq[t], hb.tiles[t], par);
        out[t].hap_f2q = hap_links(hf2q[t], hb.tiles[t], par);
    }
}

std::string uvc_indelmaps_text(const TileSparse & ts) {
    static const char *labels_ins[5] = {"", "frag_ins", "fam_ins", "cDP2_ins", "c2dDP_ins"};
    static const char *labels_del[5] = {"", "frag_del", "fam_del", "cDP2_del", "c2dDP_del"};
    std::string out;
    for (const auto & kv : ts.ins) {
        for (const auto & sc : kv.second) {
            out += std::string(labels_ins[kv.first.kind]) + "\t" + std::to_string(kv.first.strand) + "\t" + std::to_string(kv.first.symbol) + "\t"
                + std::to_string(kv.first.pos) + "\t" + sc.first + "\t" + std::to_string(sc.second.count) + "\n";
        }
    }
    for (const auto & kv : ts.del) {
        for (const auto & sc : kv.second) {
            out += std::string(labels_del[kv.first.kind]) + "\t" + std::to_string(kv.first.strand) + "\t" + std::to_string(kv.first.symbol) + "\t"
                + std::to_string(kv.first.pos) + "\t" + std::to_string(sc.first) + "\t" + std::to_string(sc.second.count) + "\n";
        }
    }
    return out;
}

std::string uvc_haplinks_text(const TileSparse & ts) {
    std::string out;
    const std::vector<HapLinkOut> *lists[3] = {&ts.hap_bq, &ts.hap_fq, &ts.hap_f2q};
    const char *labels[3] = {"bq", "fq", "f2q"};
    for (int k = 0; k < 3; k++) {
        for (const auto & h : *lists[k]) {
            out += std::string(labels[k]) + "\t" + std::to_string(h.fr_cnts[0]) + "\t" + std::to_string(h.fr_cnts[1]) + "\t"
                + std::to_string(h.other_hap_cnts[0]) + "\t" + std::to_string(h.other_hap_cnts[1]) + "\t";
            for (const auto & ps : h.pos_symb) { out += std::to_string(ps.first) + ":" + std::to_string(ps.second) + ","; }
            out += "\n";
        }
    }
    return out;
}
