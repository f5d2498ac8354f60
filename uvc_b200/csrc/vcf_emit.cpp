// vcf_emit.cpp - see vcf_emit.h. Text layout follows the reference's generated serializer (bcf_formats_generator1.cpp:644-681:
// tags in FORMAT_VEC order, Number=R tags as "ref,alt", separator pseudo-tags printed as their own name, SSCS-only tags skipped
// unless enable_tier2_consensus_format_tags) and append_vcf_record (main.hpp:6027-6272).
#include "vcf_emit.h"
#include <type_traits>

#include <algorithm>
#include <string.h>
#include <set>
#include <tuple>

#include <limits.h>
#include <math.h>
#include <stdio.h>

namespace {

const char *SYMBOL_DESC[17] = {"A", "C", "G", "T", "N", "*", "<LR>", "<LD3P>", "<LD2>", "<LD1>", "<LI3P>", "<LI2>", "<LI1>", "*", "<NONE>", "<NON_REF>", "<ADDITIONAL_INDEL_CANDIDATE>"};
const char *FTS_NAMES[UVC_NUM_FTS] = {"aStrand", "aBQXM", "aInsertSize", "aAlignL", "aAlignR", "aPositionL", "aPositionR", "abPositionL", "abPositionR",
    "bcDup", "cbDup", "c0Orientation", "c2Orientation", "c2PositionL", "c2PositionR", "c2AlignL", "c2AlignR", "c2StrictPosL", "c2StrictPosR"};

inline bool sym_is_ins(int s) { return s == UVC_LINK_I1 || s == UVC_LINK_I2 || s == UVC_LINK_I3P; }
inline bool sym_is_del(int s) { return s == UVC_LINK_D1 || s == UVC_LINK_D2 || s == UVC_LINK_D3P; }
inline bool sym_is_subst(int s) { return s >= UVC_BASE_A && s <= UVC_BASE_NN; }

inline int char_to_symbol(char c) {
    switch (c) {
        case 'A': case 'a': return UVC_BASE_A; case 'C': case 'c': return UVC_BASE_C; case 'G': case 'g': return UVC_BASE_G; case 'T': case 't': return UVC_BASE_T;
        case 'I': case 'i': return UVC_LINK_M; case '-': case '_': return UVC_LINK_D1; default: return UVC_BASE_N;
    }
}

std::string event_string(const HostBatch & hb, const StageVec<IndelEvent> & ev, int32_t e, const std::string & refstring, int32_t ext_beg) {
    static const char *nt16 = "=ACMGRSVTWYHKDBN";
    if (e < 0) { return ""; }
    const IndelEvent & E = ev[e];
    if (E.is_del) { return refstring.substr(E.rpos - ext_beg, E.oplen); }
    const uint8_t *s = hb.raw_seq(E.raw);
    std::string out;
    for (int32_t i = 0; i < E.oplen; i++) { const int32_t q = E.qpos + i; out.push_back(nt16[(s[q >> 1] >> ((~q & 1) << 2)) & 0xf]); }
    return out;
}

// decimal text of a number appended in place: integers by hand (the records are ~600 integers each; std::to_string allocates a temporary
// string per number), everything else as std::to_string prints it
template <class T> inline typename std::enable_if<std::is_integral<T>::value, void>::type append_num(std::string & s, T v) {
    char buf[24];
    char *e = buf + sizeof(buf), *p = e;
    typedef typename std::make_unsigned<T>::type U;
    U u = (U)v;
    const bool neg = (std::is_signed<T>::value && v < 0);
    if (neg) { u = (U)0 - u; }
    do { *--p = (char)('0' + (int)(u % 10)); u /= 10; } while (u);
    if (neg) { *--p = '-'; }
    s.append(p, (size_t)(e - p));
}
template <class T> inline typename std::enable_if<!std::is_integral<T>::value, void>::type append_num(std::string & s, T v) { s += std::to_string(v); }

// decimal text of an integer written at p; returns the end
template <class T> inline char *put_num(char *p, T v) {
    char buf[24];
    char *e = buf + sizeof(buf), *q = e;
    typedef typename std::make_unsigned<T>::type U;
    U u = (U)v;
    const bool neg = (std::is_signed<T>::value && v < 0);
    if (neg) { u = (U)0 - u; }
    do { *--q = (char)('0' + (int)(u % 10)); u /= 10; } while (u);
    if (neg) { *--q = '-'; }
    const size_t n = (size_t)(e - q);
    memcpy(p, q, n);
    return p + n;
}

// The FORMAT values of a record: ':' between fields, ',' inside a field. A field is assembled in a local buffer and appended once (a record has
// ~260 fields with ~700 numbers: one append per number and separator was a third of the formatting time).
struct Out {
    std::string & s;
    bool first = true;
    explicit Out(std::string & str) : s(str) {}
    void sepc() { if (!first) { s += ":"; } first = false; }
    char *begin(char *buf) { char *p = buf; if (!first) { *p++ = ':'; } first = false; return p; }
    void tag(const char *name) { sepc(); s += name; }
    template <class T> void one(T v) { char buf[32]; char *p = put_num(begin(buf), v); s.append(buf, (size_t)(p - buf)); }
    template <class T> void pair(T a, T b) { char buf[64]; char *p = put_num(begin(buf), a); *p++ = ','; p = put_num(p, b); s.append(buf, (size_t)(p - buf)); }
    template <class T> void arr(const T *v, int n) {
        char buf[32 * 24];
        char *p = begin(buf);
        for (int i = 0; i < n; i++) {
            if (i) { *p++ = ','; }
            p = put_num(p, v[i]);
            if (p > buf + sizeof(buf) - 32) { s.append(buf, (size_t)(p - buf)); p = buf; }
        }
        s.append(buf, (size_t)(p - buf));
    }
    void str(const std::string & v) { sepc(); if (v.empty()) { s += "."; } s += v; }
    void ints_or_dot(const std::vector<int32_t> & v) { sepc(); if (v.empty()) { s += '.'; } for (size_t i = 0; i < v.size(); i++) { if (i) { s += ','; } append_num(s, v[i]); } }
    void strs_or_dot(const std::vector<std::string> & v) { sepc(); if (v.empty()) { s += "."; } for (size_t i = 0; i < v.size(); i++) { if (i) { s += ","; } s += v[i]; } }
};

// mutform2count4map_to_phase (main.hpp:5380-5404) restricted to the links that contain (refpos, symbol) and have at least two supporting fragments.
// A deep tile has thousands of links and a record per position: the links that contain a (position, symbol) are indexed once per tile, in link
// order, so that a record only visits its own links.
typedef std::map<std::pair<int32_t, int32_t>, std::vector<int32_t>> PhaseIndex;
PhaseIndex phase_index(const std::vector<HapLinkOut> & links) {
    PhaseIndex idx;
    for (size_t i = 0; i < links.size(); i++) {
        const HapLinkOut & h = links[i];
        if (h.fr_cnts[0] + h.fr_cnts[1] < 2) { continue; }
        for (const auto & ps : h.pos_symb) {
            std::vector<int32_t> & v = idx[ps];
            if (v.empty() || v.back() != (int32_t)i) { v.push_back((int32_t)i); }
        }
    }
    return idx;
}
void append_phase_string(std::string & out, const std::vector<HapLinkOut> & links, const PhaseIndex & idx, int32_t refpos, int32_t symbol) {
    auto it = idx.find(std::make_pair(refpos, symbol));
    if (it == idx.end()) { return; }
    for (int32_t li : it->second) {
        const HapLinkOut & h = links[(size_t)li];
        out += '(';
        for (const auto & ps : h.pos_symb) {
            out += '('; append_num(out, ps.first + (sym_is_subst(ps.second) ? 1 : 0)); out += '&'; out += SYMBOL_DESC[ps.second]; out += ')';
        }
        out += '&'; append_num(out, h.fr_cnts[0]); out += '&'; append_num(out, h.fr_cnts[1]);
        if (-1 < h.other_hap_cnts[0]) { out += "&&"; append_num(out, h.other_hap_cnts[0] + h.fr_cnts[0]); out += '&'; append_num(out, h.other_hap_cnts[1] + h.fr_cnts[1]); }
        out += ')';
    }
}

// One allele's pass over the FTS string (fmt_bias_push, main.hpp:4258-4272, and the PASS default :4771-4773): the failed filters are appended
// to the string the format object already holds (QUIRK: the alleles of an indel symbol share the object, see PrevAllele), "PASS" if it stays empty
void fts_append(std::string & out, uint32_t fts_mask, const int32_t *fts_pct) {
    for (int k = 0; k < UVC_NUM_FTS; k++) {
        if (fts_mask & (1u << k)) {
            if (!out.empty()) { out += "|"; }
            out += std::string(FTS_NAMES[k]) + "-" + std::to_string(fts_pct[k]);
        }
    }
    if (out.empty()) { out = "PASS"; }
}
std::string fts_string(const CandFmt & c) {
    std::string out;
    fts_append(out, c.fts_mask, c.fts_pct);
    return out;
}

} // namespace

void uvc_build_indel_sites(std::vector<TileIndelSites> & sites, std::vector<IndelAllele> & table, const HostBatch & hb,
        const std::vector<TileSparse> & sparse, const std::map<int32_t, HostContig> & contigs, const StageVec<IndelEvent> & ev) {
    sites.assign(hb.tiles.size(), TileIndelSites());
    table.clear();
    for (size_t ti = 0; ti < hb.tiles.size(); ti++) {
        const TileInfo & T = hb.tiles[ti];
        if (T.skipped) { continue; }
        const TileSparse & ts = sparse[ti];
        const HostContig & contig = contigs.at(T.tid);
        const int32_t nref = (T.ext_end - T.ext_beg) - 1;
        const std::string refstring = (contig.available ? contig.bases.substr(T.ext_beg, nref) : std::string((size_t)nref, 'n'));
        // every (pos, symbol) with fragment-level indel evidence on some strand
        std::set<std::pair<int32_t, int32_t>> keys;
        for (const auto & kv : ts.ins) { if (kv.first.kind == UVC_REC_FRAG_INDEL) { keys.insert(std::make_pair(kv.first.pos, kv.first.symbol)); } }
        for (const auto & kv : ts.del) { if (kv.first.kind == UVC_REC_FRAG_INDEL) { keys.insert(std::make_pair(kv.first.pos, kv.first.symbol)); } }
        for (const auto & ps : keys) {
            const int32_t pos = ps.first, symbol = ps.second;
            const bool isins = sym_is_ins(symbol);
            IndelSite site;
            std::map<std::string, int32_t> seq2ev;
            for (int strand = 0; strand < 2; strand++) {
                // fill_by_indel_info2_x (instcode.hpp) for this strand if the fragment map has the position
                typedef std::tuple<int32_t, int32_t, int32_t, int32_t, std::string> tup_t;   // (fq, bq, c2DP, c2dDP, indel string)
                std::vector<tup_t> tuples;
                auto lookup_ins = [&](int kind, const std::string & id) -> int32_t {
                    IndelKey k; k.kind = kind; k.strand = strand; k.symbol = symbol; k.pos = pos;
                    auto it = ts.ins.find(k); if (it == ts.ins.end()) { return 0; }
                    auto jt = it->second.find(id); return (jt == it->second.end() ? 0 : jt->second.count);
                };
                auto lookup_del = [&](int kind, int32_t id) -> int32_t {
                    IndelKey k; k.kind = kind; k.strand = strand; k.symbol = symbol; k.pos = pos;
                    auto it = ts.del.find(k); if (it == ts.del.end()) { return 0; }
                    auto jt = it->second.find(id); return (jt == it->second.end() ? 0 : jt->second.count);
                };
                IndelKey k; k.kind = UVC_REC_FRAG_INDEL; k.strand = strand; k.symbol = symbol; k.pos = pos;
                bool present = false;
                if (isins) {
                    auto it = ts.ins.find(k);
                    if (it != ts.ins.end()) {
                        present = true;
                        for (const auto & idc : it->second) {
                            if (idc.first.empty()) { continue; }
                            tuples.push_back(std::make_tuple(lookup_ins(UVC_REC_FAM_INDEL, idc.first), idc.second.count, lookup_ins(UVC_REC_CDP2_INDEL, idc.first),
                                    lookup_ins(UVC_REC_C2D_INDEL, idc.first), idc.first));
                            if (!seq2ev.count(idc.first)) { seq2ev[idc.first] = idc.second.ev; }
                        }
                    }
                } else {
                    auto it = ts.del.find(k);
                    if (it != ts.del.end()) {
                        present = true;
                        for (const auto & idc : it->second) {
                            const std::string dseq = refstring.substr(pos - T.ext_beg, idc.first);
                            if (dseq.empty()) { continue; }
                            tuples.push_back(std::make_tuple(lookup_del(UVC_REC_FAM_INDEL, idc.first), idc.second.count, lookup_del(UVC_REC_CDP2_INDEL, idc.first),
                                    lookup_del(UVC_REC_C2D_INDEL, idc.first), dseq));
                            if (!seq2ev.count(dseq)) { seq2ev[dseq] = idc.second.ev; }
                        }
                    }
                }
                if (!present) { continue; }
                std::sort(tuples.rbegin(), tuples.rend());
                (strand == 0 ? site.gapNf : site.gapNr).push_back((int32_t)tuples.size());
                for (const auto & t : tuples) {
                    site.gapSeq.push_back(std::get<4>(t)); site.gapbAD1.push_back(std::get<1>(t)); site.gapcAD1.push_back(std::get<0>(t));
                    site.gc2AD.push_back(std::get<2>(t)); site.gc2dAD.push_back(std::get<3>(t));
                }
            }
            // indel_get_majority (main.hpp:5405-5455)
            int32_t nsum = 0;
            for (auto n : site.gapNf) { nsum += n; }
            for (auto n : site.gapNr) { nsum += n; }
            if (0 == nsum) {
                IndelSite::Allele a; a.bAD = 0; a.cAD = 0; a.ev = -1; a.seq = SYMBOL_DESC[symbol];
                site.alleles.push_back(a);
            } else {
                std::map<std::string, std::pair<int32_t, int32_t>> indelmap;
                for (size_t i = 0; i < site.gapSeq.size(); i++) {
                    auto & e = indelmap[site.gapSeq[i]];
                    e.first += site.gapbAD1[i]; e.second += site.gapcAD1[i];
                }
                int32_t max_bAD1 = 0;
                for (const auto & kv : indelmap) { max_bAD1 = std::max(max_bAD1, kv.second.first); }
                std::vector<IndelSite::Allele> als;
                for (const auto & kv : indelmap) {
                    if (kv.second.first >= (max_bAD1 + 3) / 4) {
                        IndelSite::Allele a; a.bAD = kv.second.first; a.cAD = kv.second.second; a.seq = kv.first; a.ev = seq2ev[kv.first];
                        als.push_back(a);
                    }
                }
                // descending by bAD^2 * length; equal keys keep map order (insertion sort on the reversed range for <= 16 elements)
                std::stable_sort(als.begin(), als.end(), [](const IndelSite::Allele & x, const IndelSite::Allele & y) {
                    return ((int64_t)x.bAD * x.bAD * (int64_t)x.seq.size()) > ((int64_t)y.bAD * y.bAD * (int64_t)y.seq.size());
                });
                site.alleles = als;
            }
            const int64_t gp = T.pos_off + (pos - T.ext_beg);
            for (const auto & a : site.alleles) {
                IndelAllele d; d.key = gp * 16 + symbol; d.bAD = a.bAD; d.cAD = a.cAD; d.ev = a.ev; d.len = (int32_t)a.seq.size();
                table.push_back(d);
            }
            sites[ti][ps] = site;
        }
    }
    std::stable_sort(table.begin(), table.end(), [](const IndelAllele & a, const IndelAllele & b) { return a.key < b.key; });
    (void)ev;
}

// What the text of every position of a tile needs: built once per tile, then the tile's positions are formatted in independent ranges
struct TileTextPlan {
    int32_t nref = 0;
    std::string refstring;
    PhaseIndex pidx_bq, pidx_fq, pidx_f2q;
    std::map<std::pair<int32_t, int32_t>, std::vector<const VarRec*>> by_zb;   // records grouped by (zero-based position, symbol type)
    size_t n_recs = 0;
};
TileTextPlan *uvc_tile_text_plan_new(const HostBatch & hb, int32_t tile_index, const HostContig & contig, const std::vector<const VarRec*> & recs, const TileSparse & sparse) {
    TileTextPlan *plan = new TileTextPlan();
    const TileInfo & T = hb.tiles[tile_index];
    if (T.skipped) { return plan; }
    plan->nref = (T.ext_end - T.ext_beg) - 1;
    plan->refstring = (contig.available ? contig.bases.substr(T.ext_beg, plan->nref) : std::string((size_t)plan->nref, 'n'));
    plan->pidx_bq = phase_index(sparse.hap_bq); plan->pidx_fq = phase_index(sparse.hap_fq); plan->pidx_f2q = phase_index(sparse.hap_f2q);
    for (const VarRec *r : recs) { plan->by_zb[std::make_pair(r->symboltype == 0 ? r->refpos + 1 : r->refpos, r->symboltype)].push_back(r); }
    for (auto & kv : plan->by_zb) {
        // the device appends with an atomic cursor: restore the reference's candidate order
        std::stable_sort(kv.second.begin(), kv.second.end(), [](const VarRec *a, const VarRec *b) { return a->cand_index < b->cand_index; });
    }
    plan->n_recs = recs.size();
    return plan;
}
void uvc_tile_text_plan_free(TileTextPlan *plan) { delete plan; }

std::string uvc_tile_vcf_text(const HostBatch & hb, int32_t tile_index, const uvcgpu_params & par, const std::string & tname, const HostContig & contig,
        const std::vector<const VarRec*> & recs, const TileIndelSites & sites, const TileSparse & sparse, const StageVec<IndelEvent> & ev,
        const GvcfPos *gvcf, const GvcfExtra *gextra) {
    const TileInfo & T = hb.tiles[tile_index];
    if (T.skipped) { return std::string(); }
    TileTextPlan *plan = uvc_tile_text_plan_new(hb, tile_index, contig, recs, sparse);
    std::string out = uvc_tile_vcf_text_range(*plan, hb, tile_index, par, tname, sites, sparse, ev, gvcf, gextra, T.rpos_inclu_beg, T.rpos_exclu_end + 1);
    uvc_tile_text_plan_free(plan);
    return out;
}

// the text of the zero-based positions [zb_begin, zb_end) of a tile (a sub-range of [rpos_inclu_beg, rpos_exclu_end]): ranges are independent
std::string uvc_tile_vcf_text_range(const TileTextPlan & plan, const HostBatch & hb, int32_t tile_index, const uvcgpu_params & par, const std::string & tname,
        const TileIndelSites & sites, const TileSparse & sparse, const StageVec<IndelEvent> & ev, const GvcfPos *gvcf, const GvcfExtra *gextra,
        int32_t zb_begin, int32_t zb_end, const std::vector<PrevAllele> *prev_alleles) {
    std::string out;
    const TileInfo & T = hb.tiles[tile_index];
    if (T.skipped || zb_end <= zb_begin) { return out; }
    const int32_t nref = plan.nref;
    const std::string & refstring = plan.refstring;
    const PhaseIndex & pidx_bq = plan.pidx_bq, & pidx_fq = plan.pidx_fq, & pidx_f2q = plan.pidx_f2q;
    const auto & by_zb = plan.by_zb;
    {
        size_t n_here = 0;
        for (auto it = by_zb.lower_bound(std::make_pair(zb_begin, 0)); it != by_zb.end() && it->first.first < zb_end; ++it) { n_here += it->second.size(); }
        out.reserve(n_here * 3200 + (size_t)(zb_end - zb_begin) * 24 + 4096);
    }
    // (the tracklen of the position before the range: what the loop carries from one position to the next)
    int32_t prev_tracklen = (zb_begin > T.rpos_inclu_beg ? gextra[T.pos_off + (zb_begin - 1 - T.ext_beg)].tracklen : 0);
    for (int32_t zb = zb_begin; zb < zb_end; zb++) {
        const int64_t gp_zb = T.pos_off + (zb - T.ext_beg);
        const GvcfExtra & X = gextra[gp_zb];
        const int32_t curr_tracklen = X.tracklen;
        const std::string repeatunit = ((zb - T.ext_beg) < nref ? refstring.substr(zb - T.ext_beg, X.unitlen) : std::string());
        if (zb != T.rpos_inclu_beg) {
            const int32_t refpos = zb - 1;
            const int64_t gp = gp_zb - 1;
            if ((par.outvar_flag & 0x8) && (((refpos % 1000) == 0) || (refpos == T.beg_pos))) {
                // MGVCF block line (main.cpp:655-757)
                const int32_t init_refQ = (INT_MAX / 2 + 1);
                int32_t prev_b = 0, prev_c = 0, prev_c12 = 0, prev_refQ = init_refQ;
                std::string body;
                const int32_t rp2end = std::min(refpos + 1000 + 1, T.ext_end);
                for (int32_t rp2 = refpos; rp2 < rp2end; rp2++) {
                    const GvcfPos & g = gvcf[T.pos_off + (rp2 - T.ext_beg)];
                    for (int k = 0; k < 2; k++) {   // k = 0: link, 1: base (SYMBOL_TYPES_IN_VCF_ORDER)
                        const int stype = (k == 0 ? 1 : 0);
                        const int32_t b = g.bdepth[k], c = g.cdepth[k], c12 = g.cdep12[k], q = g.refQ[k];
                        auto diff = [](int32_t cur, int32_t prev) { const int32_t lo = std::min(cur, prev), hi = std::max(cur, prev); if (lo * 130 >= hi * 100) { return false; } if (lo + 3 >= hi) { return false; } return true; };
                        if ((init_refQ == prev_refQ) || (abs(q - prev_refQ) > 10) || diff(b, prev_b) || diff(c, prev_c) || diff(c12, prev_c12)) {
                            append_num(body, rp2 + ((0 == stype) ? 1 : 0)); body += ','; append_num(body, 1 + stype); body += ",.,"; append_num(body, b); body += ',';
                            append_num(body, c); body += ','; append_num(body, c12); body += ','; append_num(body, q); body += ",.,";
                            prev_b = b; prev_c = c; prev_c12 = c12; prev_refQ = q;
                        }
                    }
                }
                if (!body.empty()) { body.pop_back(); }
                const std::string vcfREF = refstring.substr(refpos - T.ext_beg, 1);
                out += tname + "\t" + std::to_string(refpos + 1) + "\t.\t" + vcfREF + "\t<NON_REF>\t.\t.\tMGVCF_BLOCK\tGT:VTI:POS_VT_BDP_CDP_HomRefQ\t.:"
                     + std::to_string(char_to_symbol(vcfREF[0])) + ",15:" + body + "," + std::to_string(rp2end) + "\n";
            }
            const GvcfExtra & XR = gextra[gp];
            const int32_t aCDP = XR.a_clip, ADP = XR.a_dp;
            const bool in_long_track = (curr_tracklen > std::max(par.microadjust_alignment_tracklen_min - 1, prev_tracklen));
            const bool in_clip_region = ((aCDP >= par.microadjust_alignment_clip_min_count) && (aCDP >= ADP * (par.microadjust_alignment_clip_min_frac - DBL_EPSILON)));
            if ((0x10 & par.outvar_flag) && (in_long_track || in_clip_region) && (ADP >= 2 * par.microadjust_alignment_clip_min_count)) {
                const std::string vcfREF = refstring.substr(refpos - T.ext_beg, 1);
                out += tname + "\t" + std::to_string(refpos + 1) + "\t.\t" + vcfREF + "\t<ADDITIONAL_INDEL_CANDIDATE>\t.\t.\tADDITIONAL_INDEL_CANDIDATE;RU=" + repeatunit + ";RC="
                     + std::to_string(X.repeatnum) + "\tGT:VTI:clipDP\t.:" + std::to_string(char_to_symbol(vcfREF[0])) + ",16:" + std::to_string(ADP) + "," + std::to_string(aCDP) + "\n";
            }
        }
        for (int type = 0; type < 2; type++) {
            auto it = by_zb.find(std::make_pair(zb, type));
            if (it == by_zb.end()) { continue; }
            for (const VarRec *rp : it->second) {
                const VarRec & r = *rp;
                const CandFmt & A = r.alt, & R = r.ref;
                const GroupFmt & g = r.g;
                const int symbol = A.symbol, refpos = r.refpos;
                const int32_t regionpos = refpos - T.ext_beg;
                const std::string indelstring = ((sym_is_ins(symbol) || sym_is_del(symbol)) ? (A.ev >= 0 ? event_string(hb, ev, A.ev, refstring, T.ext_beg) : std::string(SYMBOL_DESC[symbol])) : std::string());
                int32_t vcfpos; std::string vcfref, vcfalt;
                if (indelstring.size() > 0) {
                    vcfpos = refpos;
                    vcfref = (regionpos > 0 ? refstring.substr(regionpos - 1, 1) : "n");
                    vcfalt = vcfref;
                    if ('<' == indelstring[0]) { vcfalt = indelstring; } else if (sym_is_ins(symbol)) { vcfalt += indelstring; } else { vcfref += indelstring; }
                } else {
                    if (sym_is_subst(symbol)) { vcfpos = refpos + 1; vcfref = refstring.substr(regionpos, 1); }
                    else { vcfpos = refpos; vcfref = (regionpos > 0 ? refstring.substr(regionpos - 1, 1) : "n"); }
                    vcfalt = SYMBOL_DESC[symbol];
                }
                // QUAL (main.hpp:6206): calc_non_negative<float> is re-evaluated here with the host libm from the device's integers, so that the
                // printed value does not depend on the last bit of the device's powf/log1pf (the keep decision on the device only compares with vqual)
                float vq = ((float)r.tlodq > r.lowestVAQ ? (float)r.tlodq : r.lowestVAQ);
                if (vq < 10.0f) { const float base = (float)pow(10.0, 0.1); vq = log1pf(powf(base, vq)) / logf(base); }
                const char *filter = (vq < 10 ? "Q10" : (vq < 20 ? "Q20" : (vq < 30 ? "Q30" : (vq < 40 ? "Q40" : (vq < 50 ? "Q50" : (vq < 60 ? "Q60" : "PASS"))))));
                // t2AD of an indel allele: sum of gc2dAD over the entries with this sequence (indelstring_gapSeq_gapAD_to_AD, main.hpp:5930-5939)
                int32_t t2AD1 = r.t2AD[1];
                const IndelSite *site = NULL;
                if (sym_is_ins(symbol) || sym_is_del(symbol)) {
                    auto st = sites.find(std::make_pair(refpos, symbol));
                    if (st != sites.end()) { site = &st->second; }
                    t2AD1 = 0;
                    if (site) { for (size_t i = 0; i < site->gapSeq.size(); i++) { if (site->gapSeq[i] == indelstring) { t2AD1 += site->gc2dAD[i]; } } }
                }
                out += tname; out += '\t'; append_num(out, vcfpos); out += "\t.\t"; out += vcfref; out += '\t'; out += vcfalt; out += '\t'; out += std::to_string(vq); out += '\t'; out += filter; out += '\t';
                out += "ANY_VAR;SomaticQ="; append_num(out, r.somaticq); out += ";TLODQ="; append_num(out, r.tlodq); out += ";NLODQ="; append_num(out, r.nlodq); out += ";NLODV=<NONE>";
                out += ";TNBQF="; append_num(out, r.TNBQF[0]); out += ','; append_num(out, r.TNBQF[1]); out += ','; append_num(out, r.TNBQF[2]); out += ','; append_num(out, r.TNBQF[3]);
                out += ";TNCQF="; append_num(out, r.TNCQF[0]); out += ','; append_num(out, r.TNCQF[1]); out += ','; append_num(out, r.TNCQF[2]); out += ','; append_num(out, r.TNCQF[3]);
                out += ";tbDP="; append_num(out, r.tbDP); out += ";tDP="; append_num(out, r.tDP); out += ";tAD="; append_num(out, r.tAD[0]); out += ','; append_num(out, r.tAD[1]);
                out += ";t2DP="; append_num(out, r.t2DP); out += ";t2AD="; append_num(out, r.t2AD[0]); out += ','; append_num(out, t2AD1);
                out += ";RU="; out += repeatunit; out += ";RC="; append_num(out, r.repeatnum);
                out += ";R3X2="; append_num(out, r.rtr_info[0]); out += ','; append_num(out, r.rtr_info[1]); out += ','; append_num(out, r.rtr_info[2]); out += ','; append_num(out, r.rtr_info[3]); out += ',';
                append_num(out, r.rtr_info[4]); out += ','; append_num(out, r.rtr_info[5]);
                const bool sscs = (A.enable_tier2 != 0);
                out += "\t";
                {   // FORMAT key string (bcf_formats_generator1.cpp:599-622: with or without the tier-2 consensus tags)
                    static const std::string f_head = "GT:GQ:HQ:FT:FTS:_A_:DP:AD:bDP:bAD:c2DP:c2AD:_Aa:APDP:APXM:_Ab:APLRID:APLRI:APLRP:_Ac:ALRPxT:ALRIT:ALRIt:ALRPt:ALRBt:_AQ:aMQs:AMQs:a1BQf:A1BQf:a1BQr:A1BQr:"
                        "_ADPf:aDPff:ADPff:aDPfr:ADPfr:_ADPr:aDPrf:ADPrf:aDPrr:ADPrr:_ALP:aLP1:ALP1:aLP2:ALP2:aLPL:ALPL:_ARP:aRP1:ARP1:aRP2:ARP2:aRPL:ARPL:_ALB:aLB1:aLB2:ALB2:aLBL:ALBL:"
                        "_ARB:aRB1:aRB2:ARB2:aRBL:ARBL:_ALI:aLI1:aLI2:ALI2:aLIr:ALIr:_ARI:aRI1:aRI2:ARI2:aRIf:ARIf:_AX:aBQ2:ABQ2:aPF2:APF2:aP1:AP1:aP2:AP2:_Ax:aPF1:aLIT:aRIT:aP3:aNC:"
                        "_BDP:bDPf:bDPr:BDPb:BDPd:bTAf:bTAr:BTAb:bTBf:bTBr:BTBb:_CDP1:cDP1f:cDP1r:CDP1b:CDP1d:cDP12f:cDP12r:CDP12b:_CDP2:cDP2f:cDP2r:CDP2b:CDP2d:";
                    static const std::string f_sscs = "c2BQ2:C2BQ2:c2LP0:C2LP0:c2RP0:C2RP0:_C2XP:c2LP1:c2LP2:c2RP1:c2RP2:c2LPL:c2RPL:_C2XB:c2LB1:c2LB2:c2RB1:c2RB2:c2LBL:c2RBL:_CDPx:cDP3f:cDP3r:CDP3b:cDP21f:cDP21r:CDP21b:"
                             "_cDPMm:cDPMf:cDPMr:CDPMb:cDPmf:cDPmr:CDPmb:";
                    static const std::string f_tail = "CDPDb:cDPDf:cDPDr:_DDP:DDP1:dDP1:DDP2:dDP2:_ea:aBQ:a2BQf:a2BQr:a2XM2:a2BM2:aBQQ:_eb:bMQ:aAaMQ:bNMQ:bNMa:bNMb:bMQQ:_eB:bIAQb:bIADb:bIDQb:_eC:cIAQf:cIADf:cIDQf:cIAQr:cIADr:cIDQr:"
                         "_eE:bIAQ:cIAQ:bTINQ:cTINQ:_eQ1:cPCQ1:cPLQ1:cVQ1:gVQ1:_eQ2:cPCQ2:cPLQ2:cVQ2:cMmQ:dVQinc:_CDP1vx:cDP1v:CDP1v:cDP1w:CDP1w:cDP1x:CDP1x:_CDP2vx:cDP2v:CDP2v:cDP2w:CDP2w:cDP2x:CDP2x:"
                         "_f1:CONTQ:nPF:nNFA:nAFA:nBCFA:_g1:VTI:VTD:cVQ1M:cVQ2M:cVQAM:cVQSM:_g2:gapNf:gapNr:gapSeq:gapbAD1:gapcAD1:gc2AD:gc2dAD:_g3:bDPa:cDP0a:gapSa:_h1:bHap:cHap:c2Hap:_i1:vHGQ:vAC:vNLODQ:note";
                    static const std::string f_with = f_head + f_sscs + f_tail + "\t", f_without = f_head + f_tail + "\t";
                    out += (sscs ? f_with : f_without);
                }
                Out o(out);
                #define RR(field) o.pair(R.field, A.field)
                // earlier alleles of the same indel symbol (rare; see PrevAllele in score_core.cuh): their values come first
                const PrevAllele *pa0 = NULL, *pa1 = NULL;
                if (prev_alleles && !prev_alleles->empty()) {
                    auto lo = std::lower_bound(prev_alleles->begin(), prev_alleles->end(), r.pad0, [](const PrevAllele & a, int32_t slot) { return a.rec_slot < slot; });
                    auto hi = lo;
                    while (hi != prev_alleles->end() && hi->rec_slot == r.pad0) { ++hi; }
                    if (hi != lo) { pa0 = &*lo; pa1 = pa0 + (hi - lo); }
                }
                std::string fts;
                for (const PrevAllele *pa = pa0; pa != pa1; pa++) { fts_append(fts, pa->fts_mask, pa->fts_pct); }
                fts_append(fts, A.fts_mask, A.fts_pct);
                o.str("./1"); o.one(0); o.pair(0, 0); o.str(""); o.str(fts); o.tag("_A_");
                o.one(r.DP); RR(AD); o.one(r.bDP); RR(bAD); o.one(r.c2DP); RR(c2AD); o.tag("_Aa");
                o.arr(g.APDP, 12); o.arr(g.APXM, 8); o.tag("_Ab"); o.arr(g.APLRID, 4); o.arr(g.APLRI, 4); o.arr(g.APLRP, 4); o.tag("_Ac");
                o.arr(g.ALRPxT, 2); o.arr(g.ALRIT, 4); o.arr(g.ALRIt, 4); o.arr(g.ALRPt, 4); o.arr(g.ALRBt, 4); o.tag("_AQ");
                RR(aMQs); o.one(g.AMQs[0]); RR(a1BQf); o.one(g.A1BQf[0]); RR(a1BQr); o.one(g.A1BQr[0]); o.tag("_ADPf");
                RR(aDPff); o.arr(g.ADPff, 2); RR(aDPfr); o.arr(g.ADPfr, 2); o.tag("_ADPr"); RR(aDPrf); o.arr(g.ADPrf, 2); RR(aDPrr); o.arr(g.ADPrr, 2); o.tag("_ALP");
                RR(aLP1); o.one(g.ALP1[0]); RR(aLP2); o.one(g.ALP2[0]); RR(aLPL); o.one(g.ALPL[0]); o.tag("_ARP");
                RR(aRP1); o.one(g.ARP1[0]); RR(aRP2); o.one(g.ARP2[0]); RR(aRPL); o.one(g.ARPL[0]); o.tag("_ALB");
                RR(aLB1); RR(aLB2); o.one(g.ALB2[0]); RR(aLBL); o.one(g.ALBL[0]); o.tag("_ARB");
                RR(aRB1); RR(aRB2); o.one(g.ARB2[0]); RR(aRBL); o.one(g.ARBL[0]); o.tag("_ALI");
                RR(aLI1); RR(aLI2); o.one(g.ALI2[0]); RR(aLIr); o.one(g.ALIr[0]); o.tag("_ARI");
                RR(aRI1); RR(aRI2); o.one(g.ARI2[0]); RR(aRIf); o.one(g.ARIf[0]); o.tag("_AX");
                RR(aBQ2); o.one(g.ABQ2[0]); RR(aPF2); o.one(g.APF2[0]); RR(aP1); o.one(g.AP1[0]); RR(aP2); o.one(g.AP2[0]); o.tag("_Ax");
                RR(aPF1); RR(aLIT); RR(aRIT); RR(aP3); RR(aNC); o.tag("_BDP");
                RR(bDPf); RR(bDPr); o.arr(g.BDPb, 2); o.pair(0, 0); RR(bTAf); RR(bTAr); o.arr(g.BTAb, 2); RR(bTBf); RR(bTBr); o.arr(g.BTBb, 2); o.tag("_CDP1");
                RR(cDP1f); RR(cDP1r); o.arr(g.CDP1b, 2); o.arr(g.CDP1d, 2); RR(cDP12f); RR(cDP12r); o.arr(g.CDP12b, 2); o.tag("_CDP2");
                RR(cDP2f); RR(cDP2r); o.arr(g.CDP2b, 2); o.pair(0, 0);
                if (sscs) {
                    RR(c2BQ2); o.one(g.C2BQ2[0]); RR(c2LP0); o.one(g.C2LP0[0]); RR(c2RP0); o.one(g.C2RP0[0]); o.tag("_C2XP");
                    RR(c2LP1); RR(c2LP2); RR(c2RP1); RR(c2RP2); RR(c2LPL); RR(c2RPL); o.tag("_C2XB");
                    RR(c2LB1); RR(c2LB2); RR(c2RB1); RR(c2RB2); RR(c2LBL); RR(c2RBL); o.tag("_CDPx");
                    RR(cDP3f); RR(cDP3r); o.arr(g.CDP3b, 2); RR(cDP21f); RR(cDP21r); o.arr(g.CDP21b, 2); o.tag("_cDPMm");
                    RR(cDPMf); RR(cDPMr); o.arr(g.CDPMb, 2); RR(cDPmf); RR(cDPmr); o.arr(g.CDPmb, 2);
                }
                o.arr(g.CDPDb, 2); RR(cDPDf); RR(cDPDr); o.tag("_DDP");
                o.arr(g.DDP1, 2); RR(dDP1); o.arr(g.DDP2, 2); RR(dDP2); o.tag("_ea");
                RR(aBQ); RR(a2BQf); RR(a2BQr); RR(a2XM2); RR(a2BM2); RR(aBQQ); o.tag("_eb");
                RR(bMQ); RR(aAaMQ); RR(bNMQ); RR(bNMa); RR(bNMb); RR(bMQQ); o.tag("_eB");
                RR(bIAQb); RR(bIADb); RR(bIDQb); o.tag("_eC");
                RR(cIAQf); RR(cIADf); RR(cIDQf); RR(cIAQr); RR(cIADr); RR(cIDQr); o.tag("_eE");
                RR(bIAQ); RR(cIAQ); RR(bTINQ); RR(cTINQ); o.tag("_eQ1");
                RR(cPCQ1); RR(cPLQ1); RR(cVQ1); RR(gVQ1); o.tag("_eQ2");
                RR(cPCQ2); RR(cPLQ2); RR(cVQ2); RR(cMmQ); RR(dVQinc); o.tag("_CDP1vx");
                RR(cDP1v); o.arr(g.CDP1v, 2); RR(cDP1w); o.one(g.CDP1w[0]); RR(cDP1x); o.one(g.CDP1x[0]); o.tag("_CDP2vx");
                RR(cDP2v); o.arr(g.CDP2v, 2); RR(cDP2w); o.one(g.CDP2w[0]); RR(cDP2x); o.one(g.CDP2x[0]); o.tag("_f1");
                RR(CONTQ); o.arr(A.nPF, 2);
                if (pa0 == pa1) { o.arr(A.nNFA, 6); o.arr(A.nAFA, 9); o.arr(A.nBCFA, 10); }
                else {
                    o.sepc(); for (const PrevAllele *pa = pa0; pa != pa1; pa++) { for (int k = 0; k < 6; k++) { append_num(out, pa->nNFA[k]); out += ','; } }
                    for (int k = 0; k < 6; k++) { if (k) { out += ','; } append_num(out, A.nNFA[k]); }
                    o.sepc(); for (const PrevAllele *pa = pa0; pa != pa1; pa++) { for (int k = 0; k < 9; k++) { append_num(out, pa->nAFA[k]); out += ','; } }
                    for (int k = 0; k < 9; k++) { if (k) { out += ','; } append_num(out, A.nAFA[k]); }
                    o.sepc(); for (const PrevAllele *pa = pa0; pa != pa1; pa++) { for (int k = 0; k < 10; k++) { append_num(out, pa->nBCFA[k]); out += ','; } }
                    for (int k = 0; k < 10; k++) { if (k) { out += ','; } append_num(out, A.nBCFA[k]); }
                }
                o.tag("_g1");
                o.pair(R.symbol, A.symbol);
                o.sepc(); out += std::string(SYMBOL_DESC[R.symbol]) + "," + SYMBOL_DESC[A.symbol];
                o.arr(r.cVQ1M, 2); o.arr(r.cVQ2M, 2);
                o.sepc(); out += std::string(r.cVQAM[0] >= 0 ? SYMBOL_DESC[r.cVQAM[0]] : "") + "," + (r.cVQAM[1] >= 0 ? SYMBOL_DESC[r.cVQAM[1]] : "");
                auto top_string = [&](int k) -> std::string {
                    const int sy = r.cVQAM[k];
                    if (sy < 0 || !(sym_is_ins(sy) || sym_is_del(sy))) { return std::string(); }
                    return (r.cVQSM_ev[k] >= 0 ? event_string(hb, ev, r.cVQSM_ev[k], refstring, T.ext_beg) : std::string(SYMBOL_DESC[sy]));
                };
                o.sepc(); out += top_string(0) + "," + top_string(1);
                o.tag("_g2");
                static const std::vector<int32_t> no_ints; static const std::vector<std::string> no_strs;
                o.ints_or_dot(site ? site->gapNf : no_ints); o.ints_or_dot(site ? site->gapNr : no_ints); o.strs_or_dot(site ? site->gapSeq : no_strs);
                o.ints_or_dot(site ? site->gapbAD1 : no_ints); o.ints_or_dot(site ? site->gapcAD1 : no_ints); o.ints_or_dot(site ? site->gc2AD : no_ints); o.ints_or_dot(site ? site->gc2dAD : no_ints);
                o.tag("_g3");
                RR(bDPa); RR(cDP0a);
                o.sepc(); out += "," + indelstring;
                o.tag("_h1");
                {
                    const std::vector<HapLinkOut> *hl[3] = {&sparse.hap_bq, &sparse.hap_fq, &sparse.hap_f2q};
                    const PhaseIndex *hi[3] = {&pidx_bq, &pidx_fq, &pidx_f2q};
                    for (int k = 0; k < 3; k++) { o.sepc(); const size_t before = out.size(); append_phase_string(out, *hl[k], *hi[k], refpos, symbol); if (out.size() == before) { out += '.'; } }
                }
                o.tag("_i1");
                o.one(r.vHGQ); o.arr(r.vAC, 2); o.arr(r.vNLODQ, 2); o.str("");
                #undef RR
                out += "\n";
            }
        }
        prev_tracklen = curr_tracklen;
    }
    return out;
}
