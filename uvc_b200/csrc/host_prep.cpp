// host_prep.cpp - what is left of the batch staging on the host: constants of the kernel view, the host-thread policy and the text form of
// the family grouping (test hook). Stage P0 (read filter, dedup centres, family key, grouping: reference grouping.cpp:347-442, 608-997,
// MolecularID.hpp:20-69) and stage P1 (repeat context, BAQ offsets: main.hpp:699-721, 794-874, main.cpp:400-429) run on the device
// (prep_core.cuh, prep_device.inc).
#include "host_prep.h"

#include <algorithm>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

typedef std::pair<int32_t, int32_t> tidpos_t;

void uvc_fill_view_constants(BatchView & v, const uvcgpu_params & par) {
    v.par = par;
    for (int d = 0; d < 4; d++) { v.center_pow[d] = pow(par.dedup_center_mult, (double)d); }
    v.indelphred_half = ((int32_t)round((10.0 / log(10.0)) * log(par.indel_del_to_ins_err_ratio))) / 2;
}


int uvc_host_threads(int32_t n_tiles, int requested) {
    int n = (requested > 0 ? requested : (int)std::thread::hardware_concurrency());
    const char *e = getenv("UVC_HOST_THREADS");
    if (e && atoi(e) > 0) { n = atoi(e); }
    if (n < 1) { n = 1; }
    if (n > 64) { n = 64; }
    return std::min<int>(n, n_tiles);
}



// Families of a tile in the order of the reference's std::map<MolecularBarcode, ...> (MolecularID.hpp:52-68). The device groups by exact key
// equality and orders families by their first read in the file; this test hook restores the map order from the key of each family.
std::string uvc_families_text(const HostBatch & hb, int32_t tile_index) {
    std::string out;
    const TileInfo & T = hb.tiles[tile_index];
    struct Key { tidpos_t beg, end; std::string qname, umi; uint32_t dflag, idflag; int64_t fam; std::string umi_full; };
    std::vector<Key> keys;
    for (int64_t fi = T.fam_off; fi < T.fam_off + T.n_fams; fi++) {
        const FamRec & F = hb.fams[fi];
        // the family's first read in file order carries its non-key data (and the strings of the key)
        int32_t first = INT32_MAX;
        for (int strand = 0; strand < 2; strand++) {
            for (int32_t g = F.frag_off[strand]; g < F.frag_off[strand] + F.n_frags[strand]; g++) {
                const FragRec & G = hb.frags[g];
                for (int32_t q = G.read_off; q < G.read_off + G.n_reads; q++) { first = std::min(first, hb.frag_reads[q]); }
            }
        }
        const char *qname = hb.raw_qname(hb.reads[first].raw);
        const char *h1 = strchr(qname, '#');
        const char *h2 = (h1 ? strchr(h1 + 1, '#') : NULL);
        const size_t qlen = strlen(qname);
        const char *umi_beg = (h1 ? h1 + 1 : qname + qlen), *umi_end = (h2 ? h2 : qname + qlen);
        Key k;
        k.fam = fi; k.dflag = F.duplexflag; k.idflag = F.dedup_idflag;
        k.umi_full = ((F.duplexflag & 0x1) ? std::string(umi_beg, umi_end) : std::string());
        const tidpos_t begpair(F.beg_tid, F.beg_pos), endpair(F.end_tid, F.end_pos);
        k.beg = tidpos_t(-1, -1); k.end = tidpos_t(-1, -1);
        if (0x3 == (0x3 & k.idflag)) { k.beg = std::min(begpair, endpair); k.end = std::max(begpair, endpair); }
        else if (0x1 & k.idflag) { k.beg = begpair; }
        else if (0x2 & k.idflag) { k.end = endpair; }
        if (0x4 & k.idflag) { k.qname = qname; }
        if (0x8 & k.idflag) { k.umi = k.umi_full; }
        keys.push_back(k);
    }
    std::sort(keys.begin(), keys.end(), [](const Key & a, const Key & b) {
        if (a.beg != b.beg) { return a.beg < b.beg; }
        if (a.end != b.end) { return a.end < b.end; }
        if (a.qname != b.qname) { return a.qname < b.qname; }
        if (a.umi != b.umi) { return a.umi < b.umi; }
        if (a.dflag != b.dflag) { return a.dflag < b.dflag; }
        return a.idflag < b.idflag;
    });
    for (const Key & k : keys) {
        const FamRec & F = hb.fams[k.fam];
        out += "F\t" + std::to_string(F.beg_tid) + "\t" + std::to_string(F.beg_pos) + "\t" + std::to_string(F.end_tid) + "\t" + std::to_string(F.end_pos)
            + "\t" + std::to_string(F.duplexflag) + "\t" + std::to_string(F.dedup_idflag) + "\t" + k.umi_full
            + "\t" + std::to_string(F.n_frags[0]) + "\t" + std::to_string(F.n_frags[1]) + "\n";
        for (int strand = 0; strand < 2; strand++) {
            for (int32_t g = F.frag_off[strand]; g < F.frag_off[strand] + F.n_frags[strand]; g++) {
                const FragRec & G = hb.frags[g];
                out += "f\t" + std::to_string(strand);
                for (int32_t q = G.read_off; q < G.read_off + G.n_reads; q++) {
                    const ReadRec & R = hb.reads[hb.frag_reads[q]];
                    out += std::string("\t") + hb.raw_qname(R.raw) + "/" + std::to_string(R.flag) + "/" + std::to_string(R.pos);
                }
                out += "\n";
            }
        }
    }
    return out;
}
