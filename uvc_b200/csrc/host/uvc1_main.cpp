// uvc1_main.cpp - the uvc1 command line on B200: same positional argument, options, per-region batching and block-gzipped VCF output as
// the reference's main() (main.cpp:1219-1602), with process_batch (main.cpp:458-1193) replaced by the C ABI of include/uvcgpu.h.
//
//   tiler thread      : tier-3 tile list, bit-identical to SamIter::iternext for the same -t / --mem-per-thread (tiler.h)
//   lanes (>= 1 / GPU): each lane owns a uvcgpu context (one CUDA stream) and BAM/FASTA handles; it takes the next batch of consecutive tiles,
//                       decodes the tiles' fetch windows on its decode threads, submits, collects, scores, formats and block-compresses.
//                       Several lanes per GPU overlap host staging / text formatting of one batch with the kernels of another.
//   writer            : concatenates the compressed batches in tile order, so the output is identical for any number of GPUs or lanes
//                       (main.cpp:1541-1551 semantics).
// Genomic regions are independent, so GPUs never exchange data: no collective is used.
#include "../../../include/uvcgpu.h"
#include "bam_reader.h"
#include "bgzf_writer.h"
#include "tiler.h"
#include "vcf_header.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <fstream>
#include <map>
#include <mutex>
#include <set>
#include <string>
#include <thread>
#include <vector>

#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "params_table.inc"

#define UVC_B200_VERSION "uvc-b200 0.1 (drop-in for the pileup-and-score path of UVC 0.15.1)"

namespace {

struct Options {
    std::string bam, fasta = "NA", out = "-", sample = "-", bed, targets, bed_out;
    int threads = 8;               // reference default (CmdLineArgs.hpp:33)
    int64_t mem_per_thread = 1536; // CmdLineArgs.hpp:34
    int64_t bed_in_avg_sequencing_DP = -1;
    int gpus = 0;                  // 0 = all visible
    int lanes_per_gpu = 0;         // 0 = automatic
    // a batch is full when it has enough positions to fill the GPU (every position is a thread of the position kernels) or when its reads reach
    // the memory bound (column caches and records grow with the reads): deep small panels need many reads per batch to get enough positions
    int64_t batch_positions = 512 * 1000, batch_reads = 8 * 1000 * 1000;
    int compress_level = 5;
    bool stats = false;
};

struct Batch {
    int64_t seq = -1;
    std::vector<uvchost_bedline> tiles, prevs;
};

template <class T> class Channel {
public:
    explicit Channel(size_t cap) : cap_(cap) {}
    void push(T && v) {
        std::unique_lock<std::mutex> lk(m_);
        not_full_.wait(lk, [&]() { return q_.size() < cap_ || aborted_; });
        q_.push_back(std::move(v));
        not_empty_.notify_one();
    }
    bool pop(T & v) {
        std::unique_lock<std::mutex> lk(m_);
        not_empty_.wait(lk, [&]() { return !q_.empty() || closed_ || aborted_; });
        if (aborted_ || q_.empty()) { return false; }
        v = std::move(q_.front());
        q_.pop_front();
        not_full_.notify_one();
        return true;
    }
    void close() { std::lock_guard<std::mutex> lk(m_); closed_ = true; not_empty_.notify_all(); }
    void abort() { std::lock_guard<std::mutex> lk(m_); aborted_ = true; not_empty_.notify_all(); not_full_.notify_all(); }
private:
    std::mutex m_;
    std::condition_variable not_full_, not_empty_;
    std::deque<T> q_;
    size_t cap_;
    bool closed_ = false, aborted_ = false;
};

struct Shared {
    Options opt;
    uvcgpu_params par;
    std::vector<std::pair<std::string, int64_t>> contigs;
    Channel<Batch> batches{8};
    std::mutex out_mutex;
    std::condition_variable out_cv;
    std::map<int64_t, std::string> done;     // seq -> bytes to write
    std::atomic<int> failed{0};
    std::string fail_msg;
    std::mutex fail_mutex;
    // totals
    std::atomic<int64_t> n_reads_kept{0}, n_positions{0}, n_records{0}, n_batches{0}, n_launches{0};
    std::atomic<int64_t> us_fetch{0}, us_prep{0}, us_gpu_wait{0}, us_score{0}, us_text{0}, us_compress{0};
    std::atomic<int64_t> t_tiler_open{0}, t_tiler_done{0}, t_ctx_ready{0}, t_first_batch{0};   // --stats timeline (microsecond clock values)
    std::mutex kernel_ms_mutex;
    double kernel_ms = 0;
    void fail(const std::string & msg) {
        std::lock_guard<std::mutex> lk(fail_mutex);
        if (!failed.exchange(1)) { fail_msg = msg; }
        batches.abort();
        { std::lock_guard<std::mutex> lk2(out_mutex); }     // the writer is either before its predicate check (it will see `failed`) or already waiting (it gets the notify)
        out_cv.notify_all();
    }
};

int64_t now_us() { return std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

void usage(const char *prog) {
    fprintf(stderr,
        "%s\nusage: %s <inputBAM> -f <ref.fa|NA> -o <out.vcf.gz|-> [-t threads] [-s sample] [-R regions.bed] [--targets chr:beg-end,...]\n"
        "        [-q vqual] [-A] [--outvar-flag N] [--mem-per-thread MB] [--bed-out-fname f] [--gpus N] [--lanes-per-gpu N]\n"
        "        [--gpu-batch-positions N] [--gpu-batch-reads N] [--<parameter-name> value ...]\n"
        "  <inputBAM> must be coordinate-sorted with <inputBAM>.bai next to it; <ref.fa> needs <ref.fa>.fai.\n"
        "  Every numeric field of uvcgpu_params (include/uvcgpu.h) is an option named like the reference's, e.g. --fam-thres-dup1add 2.\n",
        UVC_B200_VERSION, prog);
}

bool file_exists(const std::string & f) { std::ifstream s(f.c_str()); return (bool)s; }

// returns 0, or an exit code
int parse_args(Options & o, uvcgpu_params & par, int argc, char **argv) {
    bool have_bam = false;
    // uvcgpu_params_default() holds these four with the Illumina inference already applied; on the command line the user value comes first and
    // the platform offsets are added to it afterwards (CmdLineArgs.hpp:288-291 default 0; CmdLineArgs.cpp:127-134 adds 200 / 100)
    par.syserr_minABQ_pcr_snv = 0; par.syserr_minABQ_pcr_indel = 0; par.syserr_minABQ_cap_snv = 0; par.syserr_minABQ_cap_indel = 0;
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        auto need = [&](const char *name) -> const char* { if (i + 1 >= argc) { fprintf(stderr, "option %s needs a value\n", name); exit(109); } return argv[++i]; };
        if (a == "-h" || a == "--help") { usage(argv[0]); exit(0); }
        else if (a == "-v" || a == "--version") { printf("%s\n", UVC_B200_VERSION); exit(0); }
        else if (a == "-f" || a == "--fasta") { o.fasta = need("-f"); }
        else if (a == "-o" || a == "--output") { o.out = need("-o"); }
        else if (a == "-R" || a == "--regions-file" || a == "--bed-in-fname") { o.bed = need("-R"); }
        else if (a == "--targets") { o.targets = need("--targets"); }
        else if (a == "-s" || a == "--sample") { o.sample = need("-s"); }
        else if (a == "-t" || a == "--threads") { o.threads = atoi(need("-t")); }
        else if (a == "--mem-per-thread") { o.mem_per_thread = atoll(need(a.c_str())); }
        else if (a == "--bed-out-fname") { o.bed_out = need(a.c_str()); }
        else if (a == "--bed-in-avg-sequencing-DP") { o.bed_in_avg_sequencing_DP = atoll(need(a.c_str())); }
        else if (a == "-A" || a == "--all-out") { par.should_output_all = 1; }
        else if (a == "--all-germline-out") { par.should_output_all_germline = 1; }
        else if (a == "-q") { par.vqual = atof(need("-q")); }
        else if (a == "--gpus") { o.gpus = atoi(need(a.c_str())); }
        else if (a == "--lanes-per-gpu") { o.lanes_per_gpu = atoi(need(a.c_str())); }
        else if (a == "--gpu-batch-positions") { o.batch_positions = atoll(need(a.c_str())); }
        else if (a == "--gpu-batch-reads") { o.batch_reads = atoll(need(a.c_str())); }
        else if (a == "--compress-level") { o.compress_level = atoi(need(a.c_str())); }
        else if (a == "--stats") { o.stats = true; }
        else if (a == "--tumor-vcf") { fprintf(stderr, "--tumor-vcf (the normal pass of a tumor-normal pair) is not implemented in this build\n"); return 105; }
        else if (a == "--fam-consensus-out-fastq") { fprintf(stderr, "--fam-consensus-out-fastq is not implemented in this build\n"); return 105; }
        else if (a.size() > 2 && a[0] == '-' && a[1] == '-') {
            bool found = false;
            for (const auto & e : UVC_PARAM_OPTS) {
                if (a == e.opt) {
                    const char *v = need(e.opt);
                    char *base = (char*)&par + e.off;
                    if (e.type == 'i') { *(int32_t*)base = (int32_t)strtol(v, NULL, 0); } else if (e.type == 'u') { *(uint32_t*)base = (uint32_t)strtoul(v, NULL, 0); } else { *(double*)base = atof(v); }
                    found = true;
                    break;
                }
            }
            if (!found) { fprintf(stderr, "unknown option %s\n", a.c_str()); return 109; }
        }
        else if (a.size() > 1 && a[0] == '-') { fprintf(stderr, "unknown option %s\n", a.c_str()); return 109; }
        else if (!have_bam) { o.bam = a; have_bam = true; }
        else { fprintf(stderr, "unexpected argument %s\n", a.c_str()); return 109; }
    }
    if (!have_bam) { usage(argv[0]); return 106; }
    if (par.outvar_flag & 0x1u) { fprintf(stderr, "--outvar-flag bit 0x1 (GERMLINE text records, main.hpp:5736-5773) is not implemented in this build\n"); return 105; }
    return 0;
}

// The reference's tier-2 split of one tier-1 iteration (main.cpp:1372-1399): only used to print the same --bed-out-fname columns.
std::string bed_out_text(const std::vector<uvchost_bedline> & lines, const std::vector<std::pair<std::string, int64_t>> & contigs, int nthreads, int64_t tier1_idx) {
    int64_t nreads = 0, npositions = 0;
    for (const auto & l : lines) { nreads += l.n_reads; npositions += l.end_pos - l.beg_pos; }
    std::vector<std::pair<size_t, size_t>> pairs;
    int32_t last_tid = (lines.empty() ? -1 : lines[0].tid);
    int64_t cr = 0; int32_t cp = 0; size_t cur = 0;
    for (size_t j = 0; j < lines.size(); j++) {
        cr += lines[j].n_reads; cp += lines[j].end_pos - lines[j].beg_pos;
        if ((j > 0) && ((last_tid != lines[j].tid) || ((size_t)cr * (size_t)nthreads * 8 > (size_t)nreads) || ((size_t)(int64_t)cp * (size_t)nthreads * 8 > (size_t)npositions))) {
            pairs.push_back(std::make_pair(cur, j));
            cur = j; last_tid = lines[j].tid; cr = 0; cp = 0;
        }
    }
    pairs.push_back(std::make_pair(cur, lines.size()));
    std::string s;
    for (size_t t2 = 0; t2 < pairs.size(); t2++) {
        for (size_t t3 = pairs[t2].first; t3 < pairs[t2].second; t3++) {
            const auto & l = lines[t3];
            s += contigs[(size_t)l.tid].first + "\t" + std::to_string(l.beg_pos) + "\t" + std::to_string(l.end_pos) + "\tBedLineFlag\t" + std::to_string(l.region_flag)
               + "\tNumberOfReadsInThisInterval\t" + std::to_string(l.n_reads) + "\tNumberOfRefBasesInThisInterval\t" + std::to_string(l.end_pos - l.beg_pos)
               + "\tTier1regionIndex\t" + std::to_string(tier1_idx) + "\tTier2regionIndex\t" + std::to_string(t2) + "\tTier3regionIndex\t" + std::to_string(t3) + "\n";
        }
    }
    return s;
}

// Tiles are handed to the lanes the moment the tiler cuts them (the scan of the rest of the BAM continues meanwhile): consecutive tiles are
// packed into batches bounded by positions and reads; a batch never spans two contigs.
struct BatchPacker {
    Shared *sh;
    Batch cur;
    uvchost_bedline prev;
    int64_t seq = 0, bp = 0, br = 0;
    explicit BatchPacker(Shared *s) : sh(s) { prev.tid = -1; prev.beg_pos = 0; prev.end_pos = 0; prev.region_flag = 0; prev.n_reads = 0; }
    void flush() { if (!cur.tiles.empty()) { cur.seq = seq++; sh->batches.push(std::move(cur)); cur = Batch(); bp = 0; br = 0; } }
    void add(const uvchost_bedline & l) {
        const int64_t lp = (int64_t)(l.end_pos - l.beg_pos) + 1000, lr = l.n_reads;     // (reads reach about a fragment length beyond the tile on either side)
        if (!cur.tiles.empty() && (bp + lp > sh->opt.batch_positions || br + lr > sh->opt.batch_reads || cur.tiles.back().tid != l.tid)) { flush(); }
        cur.tiles.push_back(l); cur.prevs.push_back(prev);
        prev = l; bp += lp; br += lr;
    }
    static void on_tile(const uvchost_bedline *l, void *user) { ((BatchPacker*)user)->add(*l); }
};

void tiler_thread(Shared *sh, int scan_threads) {
    const Options & o = sh->opt;
    uvchost_tiler *t = uvchost_tiler_open(o.bam.c_str(), o.bed.c_str(), o.targets.c_str(), o.threads, o.mem_per_thread, o.bed_in_avg_sequencing_DP, 0);
    if (uvchost_tiler_error(t)[0]) { sh->fail(std::string("tiler: ") + uvchost_tiler_error(t)); uvchost_tiler_close(t); sh->batches.close(); return; }
    sh->t_tiler_open = now_us();
    std::ofstream bed_out;
    if (!o.bed_out.empty() && o.bed_out != ".") { bed_out.open(o.bed_out.c_str(), std::ios::out); }
    BatchPacker packer(sh);
    uvchost_tiler_set_callback(t, BatchPacker::on_tile, &packer);
    uvchost_tiler_set_scan_threads(t, scan_threads);
    int64_t iter = 0;
    // -R / --targets: the tiles are known up front; the lanes get them all at once. The reference's tier-1 iterations (a pass over the whole BAM
    // that counts the reads of every interval) only decide how --bed-out-fname groups the tiles: they run afterwards, and only if asked for.
    const bool eager = (uvchost_tiler_emit_given(t) > 0);
    if (eager) { packer.flush(); sh->batches.close(); sh->t_tiler_done = now_us(); }
    for (; !eager || bed_out.is_open();) {
        const uvchost_bedline *lines = NULL; int64_t n = 0;
        const int64_t nreads = uvchost_tiler_next(t, &lines, &n);
        if (nreads < 0) { sh->fail(std::string("tiler: ") + uvchost_tiler_error(t)); break; }
        if (!(nreads > 0 || n > 0)) { break; }     // main.cpp:1339
        if (!eager) { packer.flush(); }
        if (bed_out.is_open()) { bed_out << bed_out_text(std::vector<uvchost_bedline>(lines, lines + n), sh->contigs, o.threads, iter); }
        iter++;
        if (sh->failed.load()) { break; }
    }
    uvchost_tiler_close(t);
    if (!eager) { sh->batches.close(); sh->t_tiler_done = now_us(); }
}

struct Lane {
    Shared *sh = NULL;
    int device = 0, n_threads = 1, index = 0;
};

void lane_thread(Lane lane) {
    Shared *sh = lane.sh;
    const Options & o = sh->opt;
    uvcgpu_ctx *ctx = NULL;
    int rc = uvcgpu_create(&ctx, lane.device, &sh->par);
    if (rc != 0) { sh->fail("uvcgpu_create failed on device " + std::to_string(lane.device) + " with code " + std::to_string(rc) + (rc == UVCGPU_ENODEVICE ? " (no CUDA device; there is no CPU fallback)" : "")); return; }
    uvcgpu_set_host_threads(ctx, lane.n_threads);
    { int64_t z = 0; sh->t_ctx_ready.compare_exchange_strong(z, now_us()); }
    const int n_dec = std::max(1, lane.n_threads);
    std::vector<uvchost_bam*> bams((size_t)n_dec, NULL);
    // Two sets of read buffers: uvcgpu_submit is asynchronous, so the lane decodes batch k+1 into one set while the GPU works on batch k, whose
    // records stay in the other set until it is released.
    std::vector<uvchost_readbuf*> rbs2[2];
    for (int k = 0; k < n_dec; k++) { bams[k] = uvchost_bam_open(o.bam.c_str()); if (NULL == bams[k]) { sh->fail("failed to open " + o.bam); } }
    for (int h = 0; h < 2; h++) { for (int k = 0; k < n_dec; k++) { rbs2[h].push_back(uvchost_readbuf_new()); } }
    uvchost_fasta *fa = NULL;
    if (!o.fasta.empty()) { fa = uvchost_fasta_open(o.fasta.c_str()); if (NULL == fa) { sh->fail("failed to open " + o.fasta + " (or its .fai)"); } }
    std::set<int32_t> loaded;
    struct InFlight {
        bool active = false;
        uvcgpu_ticket ticket = 0;
        int64_t seq = 0;
        int32_t n_tiles = 0;
        std::set<int32_t> contigs;
        std::vector<uvcgpu_tile> tiles; std::vector<int32_t> tile_source; std::vector<uvcgpu_reads_soa> sources;
        int64_t us_fetch = 0, us_prep = 0;
    };
    InFlight fl[2];
    // collect, score, text, release, compress and publish a submitted batch
    auto finish = [&](InFlight & f) -> bool {
        if (!f.active) { return true; }
        f.active = false;
        uvcgpu_batch_stats st;
        const int64_t t2 = now_us();
        int rc2 = uvcgpu_collect(ctx, f.ticket, &st);
        const int64_t t3 = now_us();
        if (0 == rc2) { rc2 = uvcgpu_score(ctx, f.ticket, &st); }
        const int64_t t4 = now_us();
        std::string text;
        if (0 == rc2) {      // the bodies of the batch's tiles in tile order (formatted and copied out on the lane's host threads)
            size_t total = 0;
            rc2 = uvcgpu_batch_vcf(ctx, f.ticket, NULL, 0, &total);
            if (0 == rc2 && total > 0) { text.resize(total); rc2 = uvcgpu_batch_vcf(ctx, f.ticket, &text[0], total, &total); }
        }
        const int64_t t5 = now_us();
        if (rc2 != 0) { sh->fail(std::string("batch ") + std::to_string(f.seq) + " failed (" + std::to_string(rc2) + "): " + uvcgpu_last_error(ctx)); return false; }
        uvcgpu_release(ctx, f.ticket);
        std::string outbytes;
        if (o.out == "-") { outbytes.swap(text); }
        else if (uvchost_bgzf_compress(outbytes, text.data(), text.size(), o.compress_level, lane.n_threads) != 0) { sh->fail("BGZF compression failed"); return false; }
        const int64_t t6 = now_us();
        sh->n_reads_kept += st.n_reads_kept; sh->n_positions += st.n_positions; sh->n_records += st.n_vcf_records; sh->n_batches += 1; sh->n_launches += st.gpu_launches;
        sh->us_fetch += f.us_fetch; sh->us_prep += f.us_prep; sh->us_gpu_wait += t3 - t2; sh->us_score += t4 - t3; sh->us_text += t5 - t4; sh->us_compress += t6 - t5;
        { std::lock_guard<std::mutex> lk(sh->kernel_ms_mutex); sh->kernel_ms += st.kernel_ms; }
        {
            std::lock_guard<std::mutex> lk(sh->out_mutex);
            sh->done[f.seq].swap(outbytes);
            { int64_t z = 0; sh->t_first_batch.compare_exchange_strong(z, now_us()); }
        }
        sh->out_cv.notify_all();
        return true;
    };
    Batch b;
    int cur = 0;
    while (!sh->failed.load() && sh->batches.pop(b)) {
        InFlight & f = fl[cur];
        InFlight & prev = fl[cur ^ 1];
        std::vector<uvchost_readbuf*> & rbs = rbs2[cur];
        const int32_t n_tiles = (int32_t)b.tiles.size();
        // reference bases of the contigs this batch touches (load_refstring, main.cpp:54-70); the batch in flight keeps its own
        std::set<int32_t> needed;
        for (const auto & l : b.tiles) { needed.insert(l.tid); }
        for (auto it = loaded.begin(); it != loaded.end();) {
            if (!needed.count(*it) && !(prev.active && prev.contigs.count(*it))) { uvcgpu_unset_contig(ctx, *it); it = loaded.erase(it); } else { ++it; }
        }
        for (int32_t tid : needed) {
            if (loaded.count(tid)) { continue; }
            int64_t len = 0;
            char *bases = (fa ? uvchost_fasta_fetch_contig(fa, sh->contigs[(size_t)tid].first.c_str(), &len) : NULL);
            if (fa && NULL == bases) { sh->fail("contig " + sh->contigs[(size_t)tid].first + " is not in " + o.fasta); break; }
            uvcgpu_set_contig(ctx, tid, bases, bases ? len : sh->contigs[(size_t)tid].second);
            uvcgpu_set_contig_name(ctx, tid, sh->contigs[(size_t)tid].first.c_str());
            free(bases);
            loaded.insert(tid);
        }
        if (sh->failed.load()) { break; }
        // decode: the tiles are cut into n_dec runs of consecutive tiles, one decode thread and one SoA buffer per run
        const int64_t t0 = now_us();
        f.tiles.assign((size_t)n_tiles, uvcgpu_tile());
        f.tile_source.assign((size_t)n_tiles, 0);
        std::vector<uvcgpu_tile> & tiles = f.tiles;
        std::vector<int32_t> & tile_source = f.tile_source;
        const int n_src = std::min(n_dec, (int)n_tiles);
        std::vector<std::thread> pool;
        std::atomic<int> dec_failed(0);
        for (int s = 0; s < n_src; s++) {
            pool.emplace_back([&, s]() {
                uvchost_readbuf_clear(rbs[s]);
                const int32_t k0 = (int32_t)((int64_t)n_tiles * s / n_src), k1 = (int32_t)((int64_t)n_tiles * (s + 1) / n_src);
                // sam_itr_queryi(tid, beg - MAX_INSERT_SIZE, end + MAX_INSERT_SIZE) of every tile (grouping.cpp:664, 730); a batch holds one contig
                std::vector<int64_t> begs, ends, rb0((size_t)(k1 - k0)), rb1((size_t)(k1 - k0));
                for (int32_t k = k0; k < k1; k++) { begs.push_back(std::max(0, b.tiles[k].beg_pos - 2000)); ends.push_back((int64_t)b.tiles[k].end_pos + 2000); }
                if (k1 > k0) {
                    // every record decoded once, neighbouring tiles share their halos (overlapping slices); unsorted BED lines fall back to one copy per tile
                    int64_t got = uvchost_bam_fetch_span(bams[s], b.tiles[k0].tid, k1 - k0, begs.data(), ends.data(), rbs[s], rb0.data(), rb1.data());
                    if (-2 == got) { got = uvchost_bam_fetch_tiles(bams[s], b.tiles[k0].tid, k1 - k0, begs.data(), ends.data(), rbs[s], rb0.data(), rb1.data()); }
                    if (got < 0) { dec_failed.store(1); }
                }
                for (int32_t k = k0; k < k1; k++) {
                    const uvchost_bedline & l = b.tiles[k];
                    uvcgpu_tile & T = tiles[k];
                    T.tid = l.tid; T.beg_pos = l.beg_pos; T.end_pos = l.end_pos; T.region_flag = l.region_flag;
                    T.prev_tid = b.prevs[k].tid; T.prev_beg_pos = b.prevs[k].beg_pos; T.prev_end_pos = b.prevs[k].end_pos;
                    T.contig_len = (int32_t)sh->contigs[(size_t)l.tid].second;
                    T.read_begin = rb0[(size_t)(k - k0)]; T.read_end = rb1[(size_t)(k - k0)];
                    tile_source[k] = s;
                }
            });
        }
        for (auto & th : pool) { th.join(); }
        if (dec_failed.load()) { sh->fail("error while reading " + o.bam); break; }
        f.sources.assign((size_t)n_src, uvcgpu_reads_soa());
        for (int s = 0; s < n_src; s++) { uvchost_readbuf_view(rbs[s], &f.sources[s]); }
        const int64_t t1 = now_us();
        rc = uvcgpu_submit_multi(ctx, n_tiles, tiles.data(), n_src, f.sources.data(), tile_source.data(), &f.ticket);
        const int64_t t2 = now_us();
        if (rc != 0) { sh->fail(std::string("batch ") + std::to_string(b.seq) + " could not be submitted (" + std::to_string(rc) + "): " + uvcgpu_last_error(ctx)); break; }
        f.active = true; f.seq = b.seq; f.n_tiles = n_tiles; f.contigs = needed; f.us_fetch = t1 - t0; f.us_prep = t2 - t1;
        // the batch before this one had the GPU while this one was decoded: finish it now
        if (!finish(prev)) { break; }
        cur ^= 1;
    }
    for (int h = 0; h < 2; h++) { if (!sh->failed.load()) { finish(fl[cur ^ 1 ^ h]); } }
    for (int h = 0; h < 2; h++) { if (fl[h].active) { uvcgpu_release(ctx, fl[h].ticket); fl[h].active = false; } }     // (after a failure)
    for (int h = 0; h < 2; h++) { for (uvchost_readbuf *r : rbs2[h]) { uvchost_readbuf_free(r); } }
    for (int k = 0; k < n_dec; k++) { if (bams[k]) { uvchost_bam_close(bams[k]); } }
    if (fa) { uvchost_fasta_close(fa); }
    uvcgpu_destroy(ctx);
}

} // namespace

int main(int argc, char **argv) {
    const clock_t c_start = clock();
    const int64_t t_start = now_us();
    Shared sh;
    uvcgpu_params_default(&sh.par);
    sh.par.central_readlen = 0;   // 0 = estimate from the data (CmdLineArgs.hpp:100)
    Options & o = sh.opt;
    int rc = parse_args(o, sh.par, argc, argv);
    if (rc != 0) { return rc; }
    // CmdLineArgs.cpp:1015-1022
    if (!file_exists(o.bam)) { fprintf(stderr, "The file %s of type (BAM) does not exist. \n", o.bam.c_str()); return -4; }
    if (!file_exists(o.bam + ".bai")) { fprintf(stderr, "The file %s.bai of type (BAM index) does not exist. \n", o.bam.c_str()); return -4; }
    if (o.fasta != "NA") {
        if (!file_exists(o.fasta)) { fprintf(stderr, "The file %s of type (FASTA) does not exist. \n", o.fasta.c_str()); return -4; }
        if (!file_exists(o.fasta + ".fai")) { fprintf(stderr, "The file %s.fai of type (FASTA index) does not exist. \n", o.fasta.c_str()); return -4; }
    } else { o.fasta = ""; }
    if (o.threads < 1) { o.threads = 1; }

    // the CUDA driver and the device contexts come up on a helper thread while the inputs are inspected and the tiler starts its scan
    int n_dev = 0;
    std::thread gpu_warmup([&]() {
        n_dev = uvcgpu_device_count();
        const int n = (o.gpus > 0 ? std::min(o.gpus, n_dev) : n_dev);
        for (int d = 0; d < n; d++) { uvcgpu_device_warmup(d); }
    });
    struct Joiner { std::thread & t; ~Joiner() { if (t.joinable()) { t.join(); } } } warmup_joiner{gpu_warmup};

    // data-driven inference (CommandLineArgs::selfUpdateByPlatform, CmdLineArgs.cpp:36-136)
    {
        uvchost_bam *b = uvchost_bam_open(o.bam.c_str());
        if (NULL == b) { fprintf(stderr, "Failed to load BAM file %s\n", o.bam.c_str()); return -3; }
        // the reference stops when the index cannot be loaded (main.cpp:1307-1311); a missing, truncated or foreign .bai must not end in an empty VCF
        if (uvchost_bam_index_status(b) != 1) { fprintf(stderr, "Failed to load BAM index %s.bai (%s)\n", o.bam.c_str(), uvchost_bam_index_status(b) == 0 ? "not found" : "invalid or truncated"); uvchost_bam_close(b); return -5; }
        for (int32_t i = 0; i < uvchost_bam_n_targets(b); i++) { sh.contigs.push_back(std::make_pair(std::string(uvchost_bam_target_name(b, i)), uvchost_bam_target_len(b, i))); }
        uvchost_infer_stats is;
        if (uvchost_bam_infer(b, 5000, &is) != 0) { fprintf(stderr, "Failed to read %s\n", o.bam.c_str()); return -3; }
        uvchost_bam_close(b);
        sh.par.inferred_maxMQ = std::max(0, is.max_mapq);
        if (0 == sh.par.central_readlen) { sh.par.central_readlen = is.median_qlen; }
        const bool is_pe = (is.count_pe > 0);
        const bool q2x = (2 * (uint64_t)(is.q30_n_fail_bases - is.q20_n_fail_bases) < (uint64_t)is.q30_n_pass_bases);
        const bool q4x = (4 * (uint64_t)(is.q30_n_fail_bases - is.q20_n_fail_bases) < (uint64_t)is.q30_n_pass_bases);
        const bool fixqlen = ((int64_t)is.median_qlen * 100 > (int64_t)is.max_qlen * 95);
        const bool illumina = (is_pe || q4x || (q2x && fixqlen));
        fprintf(stderr, "Inferred_sequencing_platform=%s\n IsPairedEnd=%d\n", illumina ? "Illumina/BGI" : "IonTorrent/LifeTechnologies/ThermoFisher", (int)is_pe);
        if (!illumina) { fprintf(stderr, "The IonTorrent-specific code path (TIsProton) is not implemented in this build.\n"); return 105; }
        sh.par.inferred_sequencing_platform = 1;
        // SYSERR_MINABQ_SNV_ILLUMINA / SYSERR_MINABQ_INDEL_ILLUMINA are added to whatever the command line set (CmdLineArgs.cpp:127-134)
        sh.par.syserr_minABQ_pcr_snv += 200; sh.par.syserr_minABQ_pcr_indel += 100; sh.par.syserr_minABQ_cap_snv += 200; sh.par.syserr_minABQ_cap_indel += 100;
    }

    const int64_t t_setup = now_us();
    std::thread tiler(tiler_thread, &sh, std::max(1, std::min(4, std::min(std::max(o.threads, 1), std::max(1, (int)std::thread::hardware_concurrency())) / 2)));
    struct TilerGuard { Shared & sh; std::thread & t; ~TilerGuard() { if (t.joinable()) { sh.failed.store(true); sh.batches.abort(); t.join(); } } } tiler_guard{sh, tiler};   // early returns below
    gpu_warmup.join();
#ifdef UVC_EMU_HOST
    const int n_gpus = 1;
    (void)n_dev;
#else
    if (n_dev <= 0) { fprintf(stderr, "no CUDA device found: uvc1 (B200 build) has no CPU fallback\n"); return 107; }
    const int n_gpus = (o.gpus > 0 ? std::min(o.gpus, n_dev) : n_dev);
#endif
    const int hw = std::max(1, (int)std::thread::hardware_concurrency());
    const int cpu_budget = std::min(std::max(o.threads, 1), hw);
    int lanes_per_gpu = (o.lanes_per_gpu > 0 ? o.lanes_per_gpu : std::max(1, std::min(3, cpu_budget / (2 * n_gpus))));
    const int n_lanes = lanes_per_gpu * n_gpus;
    const int threads_per_lane = std::max(1, (cpu_budget + n_lanes - 1) / n_lanes);

    FILE *fout = NULL;
    const bool to_stdout = (o.out == "-");
    if (!to_stdout && !o.out.empty()) {
        fout = fopen(o.out.c_str(), "wb");
        if (NULL == fout) { fprintf(stderr, "Unable to open the bgzip file %s\n", o.out.c_str()); return -8; }
    }
    {
        const std::string header = uvc_vcf_header(argc, argv, sh.contigs, o.fasta, o.sample, sh.par, UVC_B200_VERSION);
        if (to_stdout) { fwrite(header.data(), 1, header.size(), stdout); }
        else if (fout) { std::string z; uvchost_bgzf_compress(z, header.data(), header.size(), o.compress_level, 1); fwrite(z.data(), 1, z.size(), fout); }
    }

    std::vector<std::thread> lanes;
    for (int k = 0; k < n_lanes; k++) {
        Lane l; l.sh = &sh; l.device = k % n_gpus; l.n_threads = threads_per_lane; l.index = k;
        lanes.emplace_back(lane_thread, l);
    }
    // writer: batches in tile order
    std::atomic<int> lanes_done(0);
    std::thread joiner([&]() { for (auto & th : lanes) { th.join(); } { std::lock_guard<std::mutex> lk(sh.out_mutex); lanes_done.store(1); } sh.out_cv.notify_all(); });
    int64_t next_seq = 0, bytes_written = 0;
    for (;;) {
        std::string chunk;
        {
            std::unique_lock<std::mutex> lk(sh.out_mutex);
            sh.out_cv.wait(lk, [&]() { return sh.done.count(next_seq) || lanes_done.load() || sh.failed.load(); });
            auto it = sh.done.find(next_seq);
            if (it == sh.done.end()) { break; }
            chunk.swap(it->second);
            sh.done.erase(it);
        }
        if (!chunk.empty()) {
            if (to_stdout) { fwrite(chunk.data(), 1, chunk.size(), stdout); } else if (fout) { fwrite(chunk.data(), 1, chunk.size(), fout); }
            bytes_written += (int64_t)chunk.size();
        }
        next_seq++;
    }
    tiler.join();
    joiner.join();
    if (sh.failed.load()) {
        fprintf(stderr, "uvc1: %s\n", sh.fail_msg.c_str());
        if (fout) { fclose(fout); }
        return 110;
    }
    if (fout) { std::string z; uvchost_bgzf_eof(z); fwrite(z.data(), 1, z.size(), fout); fclose(fout); }
    if (to_stdout) { fflush(stdout); }
    const double wall = (now_us() - t_start) / 1e6;
    if (o.stats) {
        fprintf(stderr, "uvc1-b200 stats: gpus=%d lanes=%d threads_per_lane=%d batches=%lld reads_kept=%lld positions=%lld records=%lld launches=%lld kernel_ms=%.1f\n"
                        "uvc1-b200 stage seconds summed over lanes: decode=%.2f stage+h2d=%.2f gpu_wait=%.2f score=%.2f text=%.2f compress=%.2f\n"
                        "uvc1-b200 timeline seconds since start: inference+header %.2f, first context ready %.2f, tiler open %.2f, first batch out %.2f, tiler done %.2f, all written %.2f\n"
                        "uvc1-b200 throughput: %.0f reads/s %.0f positions/s\n",
                n_gpus, n_lanes, threads_per_lane, (long long)sh.n_batches.load(), (long long)sh.n_reads_kept.load(), (long long)sh.n_positions.load(), (long long)sh.n_records.load(),
                (long long)sh.n_launches.load(), sh.kernel_ms,
                sh.us_fetch / 1e6, sh.us_prep / 1e6, sh.us_gpu_wait / 1e6, sh.us_score / 1e6, sh.us_text / 1e6, sh.us_compress / 1e6,
                (t_setup - t_start) / 1e6, (sh.t_ctx_ready.load() - t_start) / 1e6, (sh.t_tiler_open.load() - t_start) / 1e6, (sh.t_first_batch.load() - t_start) / 1e6,
                (sh.t_tiler_done.load() - t_start) / 1e6, wall,
                sh.n_reads_kept.load() / wall, sh.n_positions.load() / wall);
    }
    fprintf(stderr, "CPU time used: %.2f seconds\nWall clock time passed: %.2f seconds\n", (double)(clock() - c_start) / CLOCKS_PER_SEC, wall);
    return 0;
}
