// host_prep.h - host-side pieces of the batch staging: page-locked staging memory, the host mirror of a batch (tiles; the grouping only when a
// test hook asks for it) and the host-thread policy. Stages P0 and P1 themselves run on the device (prep_core.cuh, prep_device.inc).
#ifndef UVC_HOST_PREP_H_INCLUDED
#define UVC_HOST_PREP_H_INCLUDED

#include "batch.h"

#include <map>
#include <new>
#include <utility>
#include <string>
#include <vector>

// Staging memory. Large blocks come from a process-wide cache of page-locked host memory (CUDA build; plain malloc in the emulation build), so
// that the uploads and downloads of a batch are true asynchronous DMA transfers and no page-locking cost is paid per batch; small blocks are malloc'ed.
void *uvc_stage_alloc(size_t bytes);
void uvc_stage_free(void *p, size_t bytes);
void uvc_stage_thread_pinning(bool enabled);   // per thread: the tile-private scratch of the staging workers is not worth page-locking
template <class T> struct UvcStageAlloc {
    typedef T value_type;
    UvcStageAlloc() {}
    template <class U> UvcStageAlloc(const UvcStageAlloc<U> &) {}
    T *allocate(size_t n) { return (T*)uvc_stage_alloc(n * sizeof(T)); }
    void deallocate(T *p, size_t n) { uvc_stage_free((void*)p, n * sizeof(T)); }
    // resize(n) default-initialises (no zero fill): staging arrays are written in full before they are uploaded, and a single-threaded memset of
    // hundreds of megabytes per batch is not free
    template <class U> void construct(U *p) { ::new ((void*)p) U; }
    template <class U, class A0, class... Args> void construct(U *p, A0 && a0, Args &&... args) { ::new ((void*)p) U(std::forward<A0>(a0), std::forward<Args>(args)...); }
    template <class U> bool operator==(const UvcStageAlloc<U> &) const { return true; }
    template <class U> bool operator!=(const UvcStageAlloc<U> &) const { return false; }
};
template <class T> using StageVec = std::vector<T, UvcStageAlloc<T>>;

struct HostContig {
    std::string bases;   // upper-cased; empty if not available
    int64_t len = 0;
    bool available = false;
};

struct HostBatch {
    StageVec<TileInfo> tiles;
    StageVec<int32_t> pos_tile;
    StageVec<uint8_t> refsym;
    StageVec<uvcgpu_rtr> rtr;          // as computed from the reference string (before the threshold pass adjusts indelphred)
    StageVec<int32_t> baq, baq2;
    StageVec<ReadRec> reads;
    std::vector<int64_t> read_raw_index;  // index into the caller's uvcgpu_reads_soa
    StageVec<uint8_t> seq, qual;
    StageVec<uint32_t> cigar;
    StageVec<FragRec> frags;
    StageVec<int32_t> frag_reads;
    StageVec<FamRec> fams;
    StageVec<ReadFam> rfam;               // per read, see batch.h
    std::vector<std::string> fam_umi;     // umistring of each family (for the grouping dump)
    StageVec<int32_t> fchunk_frag, mchunk_fs;   // owner of every 32-entry chunk of the fragment / family-strand columns
    int64_t n_fcol = 0, n_mcol = 0;       // padded column entries
    int64_t n_pos = 0, n_cx = 0, n_ev = 0;
    int64_t n_reads_in = 0;
    // where the batch's raw record arrays (device staging) came from: source s contributed its records [raw_host_begin[s], ...) as raw indices
    // [raw_dev_begin[s], raw_dev_begin[s + 1]); the host reads inserted bases and names from the caller's SoA (borrowed until release)
    std::vector<uvcgpu_reads_soa> raw_sources;
    std::vector<int64_t> raw_host_begin, raw_dev_begin;
    int64_t raw_to_host(int64_t raw, size_t & s) const {
        s = 0;
        while (s + 1 < raw_sources.size() && raw_dev_begin[s + 1] <= raw) { s++; }
        return raw_host_begin[s] + (raw - raw_dev_begin[s]);
    }
    const uint8_t *raw_seq(int64_t raw) const { size_t s; const int64_t i = raw_to_host(raw, s); return raw_sources[s].seq + raw_sources[s].seq_off[i]; }
    const char *raw_qname(int64_t raw) const { size_t s; const int64_t i = raw_to_host(raw, s); return raw_sources[s].qname + raw_sources[s].qname_off[i]; }
};

// Text form of the family grouping of one tile (same format as oracle/harness_dump.cpp writes).
std::string uvc_families_text(const HostBatch & hb, int32_t tile_index);

void uvc_fill_view_constants(BatchView & v, const uvcgpu_params & par);

// Number of host threads to use for n independent items (requested <= 0: all cores; the UVC_HOST_THREADS environment variable overrides).
int uvc_host_threads(int32_t n_items, int requested);

// Dynamic parallel loop over [0, n) on n_threads host threads (requested semantics as above).
template <class F> void uvc_parallel_for(int32_t n, int requested_threads, F body);


#include <atomic>
#include <thread>

template <class F> void uvc_parallel_for(int32_t n, int requested_threads, F body) {
    const int n_threads = uvc_host_threads(n, requested_threads);
    if (n_threads <= 1) { for (int32_t i = 0; i < n; i++) { body(i); } return; }
    std::atomic<int32_t> next(0);
    std::vector<std::thread> pool;
    for (int t = 0; t < n_threads; t++) {
        pool.emplace_back([&]() { for (;;) { const int32_t i = next.fetch_add(1); if (i >= n) { break; } body(i); } });
    }
    for (auto & th : pool) { th.join(); }
}

#endif
