// tiler.cpp - see tiler.h. Integer types follow the reference's (int64 totals, size_t products) because the memory model mixes signed
// and unsigned 64-bit arithmetic (grouping.cpp:28-67).
#include "tiler.h"

#include <algorithm>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace {

const size_t NUM_BYTES_PER_REF_POS = 1024 * 8;   // grouping.cpp:9
const size_t NUM_BYTES_PER_READ = 512;           // grouping.cpp:10
const int64_t MAX_STR_N_BASES = 100;             // common.hpp:63
const size_t NUM_WORKING_UNITS_PER_THREAD = 8;   // common.hpp:45

inline size_t square_big(int64_t x) { return ((size_t)x) * ((size_t)x); }

// grouping.cpp:28-46
bool over_mem_lim(int64_t n_reads, int64_t n_reads_sq, int64_t n_rposs, int64_t n_rposs_sq, size_t nthreads, size_t mem_per_thread, bool is_fastq_gen) {
    const size_t a = (size_t)(n_reads_sq / (n_reads > 1 ? n_reads : 1)) * nthreads;
    const size_t reads_bytes = (size_t)((int64_t)(a < (size_t)n_reads ? a : (size_t)n_reads) * (int64_t)NUM_BYTES_PER_READ);
    const size_t b = (size_t)(n_rposs_sq / (n_rposs > 1 ? n_rposs : 1)) * nthreads;
    const size_t rposs_bytes = (size_t)((int64_t)((b < (size_t)n_rposs ? b : (size_t)n_rposs) + (2 * (size_t)MAX_STR_N_BASES * nthreads)) * (int64_t)NUM_BYTES_PER_REF_POS);
    const size_t vcf_bytes = (size_t)(n_rposs * (int64_t)1024);
    const size_t fqs_bytes = (is_fastq_gen ? (size_t)((n_reads * (int64_t)NUM_BYTES_PER_READ) / 4) : 0);
    return (reads_bytes + rposs_bytes + vcf_bytes + fqs_bytes) > ((1024UL * 1024UL) * mem_per_thread * nthreads);
}

// grouping.cpp:48-67
bool sub_over_mem_lim(int64_t region_n_reads, int64_t region_n_rposs, size_t mem_per_thread, size_t curr_beg, size_t block_running_end) {
    const size_t reads_bytes = (size_t)(region_n_reads * (int64_t)NUM_BYTES_PER_READ);
    const size_t rposs_bytes = (size_t)(region_n_rposs * (int64_t)(NUM_BYTES_PER_REF_POS + 1024));
    const size_t memfree = ((1024UL * 1024UL) / NUM_WORKING_UNITS_PER_THREAD) * mem_per_thread;
    const size_t gap = (block_running_end > curr_beg ? block_running_end - curr_beg : 0);
    const size_t overlap_bytes = memfree * (gap < 150 ? gap : 150) / 150;   // more overlap -> more memory allowed
    return (reads_bytes + rposs_bytes) > memfree + overlap_bytes;
}

} // namespace

struct uvchost_tiler {
    uvchost_bam *bam = NULL;
    std::string err;
    size_t nthreads = 1, mem_per_thread = 1536;
    int64_t bed_dp = -1;
    bool is_fastq_gen = false;
    std::vector<uvchost_bedline> given;     // -R / --targets intervals
    std::vector<int64_t> given_counts;      // reads overlapping each interval, when counted by one sweep over the BAM
    size_t given_idx = 0;
    int32_t last_tid = -1, last_beg = -1, last_end = -1;
    uvchost_core rec;                       // the reference's single alnrecord: keeps its content at end of file
    bool started = false;
    std::vector<uvchost_bedline> out;
    uvchost_tile_cb cb = NULL;
    void *cb_user = NULL;
    int32_t scan_threads = 1;
    void emit(const uvchost_bedline & l) { out.push_back(l); if (cb) { cb(&l, cb_user); } }
};

// Number of records overlapping each given interval, for all intervals in ONE sequential pass over the BAM (read-ahead inflation on
// scan_threads threads). Gives exactly the per-interval counts of SamIter::iternext (grouping.cpp:178-195), which queries the index once
// per interval and re-reads, for dense panels, the same 16 kbp index window for every target inside it.
static void count_by_sweep(uvchost_tiler *t) {
    struct Iv { int32_t beg, end; size_t idx; };
    std::map<int32_t, std::vector<Iv>> by_tid;
    for (size_t i = 0; i < t->given.size(); i++) { Iv v; v.beg = t->given[i].beg_pos; v.end = t->given[i].end_pos; v.idx = i; by_tid[t->given[i].tid].push_back(v); }
    std::map<int32_t, std::vector<int32_t>> prefmax;
    for (auto & kv : by_tid) {
        std::sort(kv.second.begin(), kv.second.end(), [](const Iv & a, const Iv & b) { return a.beg < b.beg; });
        std::vector<int32_t> & pm = prefmax[kv.first];
        int32_t m = INT_MIN;
        for (const auto & v : kv.second) { m = (v.end > m ? v.end : m); pm.push_back(m); }
    }
    t->given_counts.assign(t->given.size(), 0);
    if (uvchost_bam_rewind_parallel(t->bam, t->scan_threads > 1 ? t->scan_threads : 2) != 0) { return; }
    uvchost_core c;
    int32_t cur_tid = -2; const std::vector<Iv> *ivs = NULL; const std::vector<int32_t> *pm = NULL;
    while (uvchost_bam_next_core(t->bam, &c) > 0) {
        if (c.tid != cur_tid) {
            cur_tid = c.tid;
            auto it = by_tid.find(c.tid);
            ivs = (it == by_tid.end() ? NULL : &it->second);
            pm = (it == by_tid.end() ? NULL : &prefmax[c.tid]);
        }
        if (NULL == ivs) { continue; }
        // intervals with beg < endpos, walked from the last one down while any earlier interval can still end after pos
        size_t hi = std::lower_bound(ivs->begin(), ivs->end(), c.endpos, [](const Iv & a, int32_t e) { return a.beg < e; }) - ivs->begin();
        while (hi > 0 && (*pm)[hi - 1] > c.pos) {
            if ((*ivs)[hi - 1].end > c.pos) { t->given_counts[(*ivs)[hi - 1].idx] += 1; }
            hi--;
        }
    }
}

extern "C" {

const char *uvchost_tiler_error(const uvchost_tiler *t) { return t->err.c_str(); }
void uvchost_tiler_set_callback(uvchost_tiler *t, uvchost_tile_cb cb, void *user) { t->cb = cb; t->cb_user = user; }
void uvchost_tiler_set_scan_threads(uvchost_tiler *t, int32_t scan_threads) { t->scan_threads = scan_threads; }

void uvchost_tiler_close(uvchost_tiler *t) {
    if (NULL == t) { return; }
    if (t->bam) { uvchost_bam_close(t->bam); }
    delete t;
}

uvchost_tiler *uvchost_tiler_open(const char *bam_path, const char *bed_fname, const char *targets, int32_t nthreads, int64_t mem_per_thread_mb,
        int64_t bed_in_avg_sequencing_DP, int32_t is_fastq_gen) {
    uvchost_tiler *t = new uvchost_tiler();
    t->bam = uvchost_bam_open(bam_path);
    t->nthreads = (size_t)(nthreads > 0 ? nthreads : 1);
    t->mem_per_thread = (size_t)mem_per_thread_mb;
    t->bed_dp = bed_in_avg_sequencing_DP;
    t->is_fastq_gen = (0 != is_fastq_gen);
    memset(&t->rec, 0, sizeof(t->rec));
    t->rec.endpos = 1;
    if (NULL == t->bam) { t->err = std::string("failed to open ") + bam_path; return t; }
    std::map<std::string, int32_t> name2tid;
    for (int32_t i = 0; i < uvchost_bam_n_targets(t->bam); i++) { name2tid[uvchost_bam_target_name(t->bam, i)] = i; }
    const bool has_targets = (targets && targets[0] && strcmp(targets, ".") != 0);
    const bool has_bed = (bed_fname && bed_fname[0] && strcmp(bed_fname, ".") != 0);
    if (has_targets) {   // target_region_to_contigs (grouping.cpp:69-107): chr:beg-end[,chr:beg-end...] or chr:pos
        std::istringstream ss(targets);
        std::string region;
        while (getline(ss, region, ',')) {
            std::vector<char> name(region.size() + 1);
            unsigned long b1 = 0, e1 = 0;
            int n = sscanf(region.c_str(), "%[^:]:%lu-%lu", name.data(), &b1, &e1);
            if (n < 3) { n = sscanf(region.c_str(), "%[^:]:%lu", name.data(), &b1); e1 = b1 + 1; }
            if (n < 2) { t->err = "the region " + region + " is neither TEMPLATE:START-END nor TEMPLATE:POS"; return t; }
            auto it = name2tid.find(name.data());
            if (it == name2tid.end()) { t->err = "the template name of " + region + " is not in the BAM header"; return t; }
            uvchost_bedline l; l.tid = it->second; l.beg_pos = (int32_t)b1; l.end_pos = (int32_t)e1; l.region_flag = 0;
            l.n_reads = ((-1 == t->bed_dp) ? 0 : (t->bed_dp * (l.end_pos - l.beg_pos) + 1));
            t->given.push_back(l);
        }
    } else if (has_bed) { // bed_fname_to_contigs (grouping.cpp:109-155)
        std::ifstream f(bed_fname);
        if (!f.good()) { t->err = std::string("failed to open ") + bed_fname; return t; }
        while (f.good()) {
            std::string line;
            getline(f, line);
            if (line.empty() || line[0] == '#') { continue; }
            std::istringstream ls(line);
            std::string name, token;
            int32_t b = 0, e = 0;
            ls >> name; ls >> b; ls >> e;
            if (!(b < e)) { t->err = "BED interval does not end after its start: " + line; return t; }
            auto it = name2tid.find(name);
            if (it == name2tid.end()) { t->err = "the template name " + name + " of the BED file is not in the BAM header"; return t; }
            uvchost_bedline l; l.tid = it->second; l.beg_pos = b; l.end_pos = e; l.region_flag = 0;
            l.n_reads = (int32_t)((-1 == t->bed_dp) ? 0 : (t->bed_dp * (e - b) + 1));
            while (ls.good()) {
                ls >> token;
                if (token == "BedLineFlag") { ls >> l.region_flag; }
                else if (token == "NumberOfReadsInThisInterval") { int32_t n = 0; ls >> n; l.n_reads = n; }
            }
            t->given.push_back(l);
        }
    }
    return t;
}

int64_t uvchost_tiler_emit_given(uvchost_tiler *t) {
    if (NULL == t || NULL == t->bam || t->given.empty() || NULL == t->cb) { return 0; }
    for (const uvchost_bedline & l : t->given) {
        uvchost_bedline lc = l;
        if (lc.n_reads <= 0) { lc.n_reads = uvchost_bam_estimate_reads(t->bam, l.tid, l.beg_pos, l.end_pos); }
        t->cb(&lc, t->cb_user);
    }
    t->cb = NULL;      // the iterations that follow (if the caller wants them) only report
    return (int64_t)t->given.size();
}

int64_t uvchost_tiler_next(uvchost_tiler *t, const uvchost_bedline **lines, int64_t *n_lines) {
    t->out.clear();
    *lines = NULL; *n_lines = 0;
    if (NULL == t->bam) { return -1; }
    int64_t total_n_reads = 0, total_n_rposs = 0, total_n_reads_sq = 0, total_n_rposs_sq = 0;
    if (!t->given.empty()) {
        // grouping.cpp:170-213
        if (-1 == t->bed_dp && t->given_counts.empty() && t->given.size() >= 32) { count_by_sweep(t); }
        for (; t->given_idx < t->given.size(); t->given_idx++) {
            uvchost_bedline l = t->given[t->given_idx];
            int64_t region_n_reads = t->bed_dp * (int64_t)(l.end_pos - l.beg_pos);
            if (-1 == t->bed_dp) { region_n_reads = (t->given_counts.empty() ? uvchost_bam_count(t->bam, l.tid, l.beg_pos, l.end_pos) : t->given_counts[t->given_idx]); }
            // the tile list keeps the line as given (n_reads = 0 like the reference's BedLine); the callback gets the counted reads, which size the GPU batches
            t->out.push_back(l);
            if (t->cb) { uvchost_bedline lc = l; lc.n_reads = region_n_reads; t->cb(&lc, t->cb_user); }
            const int64_t region_n_rposs = l.end_pos - l.beg_pos;
            total_n_reads += region_n_reads; total_n_rposs += region_n_rposs;
            total_n_reads_sq += (int64_t)square_big(region_n_reads); total_n_rposs_sq += (int64_t)square_big(region_n_rposs);
            if (over_mem_lim(total_n_reads, total_n_reads_sq, total_n_rposs, total_n_rposs_sq, t->nthreads, t->mem_per_thread, t->is_fastq_gen)) {
                t->given_idx++;
                break;
            }
        }
    } else {
        // grouping.cpp:214-310
        if (!t->started) { if (uvchost_bam_rewind_parallel(t->bam, t->scan_threads) != 0) { t->err = "seek failed"; return -1; } t->started = true; }
        int32_t block_tid = t->last_tid, block_beg = t->last_beg, block_running_end = t->last_end;
        int64_t region_n_reads = 0, region_n_rposs = 0, region_n_rposs_add = 0;
        int ret = -1;
        bool returned_early = false;
        do {
            const int r = uvchost_bam_next_core(t->bam, &t->rec);
            ret = (r > 0 ? 0 : (0 == r ? -1 : -2));
            if (ret < -1) { t->err = "error while reading the BAM file"; break; }
            if (t->rec.flag & 0x4) { continue; }   // QUIRK: also at end of file, so a trailing unmapped record suppresses the final cut
            const int32_t curr_tid = t->rec.tid, curr_beg = t->rec.pos, curr_end = t->rec.endpos;
            const bool sub_over = sub_over_mem_lim(region_n_reads, region_n_rposs + region_n_rposs_add, t->mem_per_thread, (size_t)(int64_t)curr_beg, (size_t)(int64_t)block_running_end);
            const bool tid_changed = (curr_tid != block_tid);
            const bool far_jumped = ((curr_tid == block_tid) && ((int64_t)block_running_end + (MAX_STR_N_BASES * 2) < curr_beg));
            const uint32_t region_flag = (tid_changed ? 16u : 0u) + (far_jumped ? 8u : 0u) + (sub_over ? 4u : 0u) + ((-1 == ret) ? 2u : 0u);
            if (region_flag) {
                const bool is_1st_read = (-1 == block_tid);
                const int64_t cap = (is_1st_read ? (int64_t)INT_MAX : (int64_t)(int32_t)uvchost_bam_target_len(t->bam, block_tid));
                const int64_t block_norm_end = ((int64_t)block_running_end < cap ? (int64_t)block_running_end : cap);
                const bool zero_sized = ((int64_t)block_beg >= block_norm_end);
                if ((!is_1st_read) && (!zero_sized)) {
                    uvchost_bedline l; l.tid = block_tid; l.beg_pos = block_beg; l.end_pos = (int32_t)block_norm_end; l.region_flag = region_flag; l.n_reads = region_n_reads;
                    t->emit(l);
                    const int64_t region_s_rposs = region_n_rposs + region_n_rposs_add;
                    total_n_reads += region_n_reads; total_n_rposs += region_s_rposs;
                    total_n_reads_sq += (int64_t)square_big(region_n_reads); total_n_rposs_sq += (int64_t)square_big(region_s_rposs);
                    region_n_rposs = 0; region_n_rposs_add = 0; region_n_reads = 0;
                }
                block_tid = curr_tid;
                const int32_t new_block_beg = (block_beg > curr_beg ? block_beg : curr_beg);   // skip over non-covered bases
                block_beg = (tid_changed ? curr_beg : (int32_t)((int64_t)new_block_beg > block_norm_end ? (int64_t)new_block_beg : block_norm_end));
                if (over_mem_lim(total_n_reads, total_n_reads_sq, total_n_rposs, total_n_rposs_sq, t->nthreads, t->mem_per_thread, t->is_fastq_gen)) {
                    t->last_tid = block_tid; t->last_beg = block_beg;
                    t->last_end = (int32_t)((int64_t)block_beg > block_norm_end ? (int64_t)block_beg : block_norm_end);
                    returned_early = true;   // QUIRK: the current read is neither counted nor extends the running end
                    break;
                }
            }
            if (tid_changed) {
                block_beg = curr_beg;
                block_running_end = curr_end;
                region_n_rposs_add += region_n_rposs;
            } else {
                block_running_end = (block_running_end > curr_end ? block_running_end : curr_end);
            }
            region_n_reads++;
            region_n_rposs = (int64_t)block_running_end - (int64_t)block_beg;
        } while (ret >= 0);
        (void)returned_early;
    }
    *lines = t->out.data(); *n_lines = (int64_t)t->out.size();
    return total_n_reads;
}

} // extern "C"
