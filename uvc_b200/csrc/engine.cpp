// engine.cpp - the C ABI of include/uvcgpu.h.
//
// Compiled as CUDA (nvcc -x cu) into libuvcgpu.so: the product. The same file compiled by g++ with -DUVC_EMU gives
// tests/emu/libuvcgpu_emu.so, which runs the identical per-item kernel bodies in serial loops so that the host logic and
// the kernel logic can be unit-tested in a container without a GPU. The emulation library is test infrastructure: nothing in
// the product (bench.py, __graft_entry__.py, the uvc1 host) loads it, and libuvcgpu.so has no CPU path - uvcgpu_create fails
// with UVCGPU_ENODEVICE when no CUDA device is present.
#include "../../include/uvcgpu.h"
#include "batch.h"
#include "host_prep.h"
#include "kernels_core.cuh"
#include "prep_core.cuh"
#include "score_core.cuh"
#include "sparse_out.h"
#include "vcf_emit.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#if defined(__CUDACC__) && !defined(UVC_EMU)
#define UVC_CUDA 1
#include <cuda_runtime.h>
#else
#define UVC_CUDA 0
#endif

#define UVC_CURSOR_SLOTS 64
#define UVC_CURSOR_WORDS 16     // per slot: 4 words of the sparse stream's cursor, 8 of the scoring pipeline's counters
#define UVC_N_PILEUP_STAGES 12   // K0, K1, K2, K2e, KF, K3a, K3b, KM, K4a, K4, K4c, K6 (kernel_ms_by_stage[0..11]); [12] = K5

namespace {

double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct BatchState {
    HostBatch hb;
    BatchView view;                       // pointers valid on the compute side (device or, in the emulation, host)
    std::vector<void*> allocs;            // compute-side allocations
    std::vector<void*> temp_allocs;       // scratch of the device staging (stages P0/P1), freed as soon as the staging kernels are enqueued
    // CUDA build: device memory of the batch is carved from slabs of the process-wide slab cache (see DevArena below)
    struct Slab { char *base; size_t size; };
    struct Arena { std::vector<Slab> slabs; size_t used = 0, requested = 0; };
    Arena keep_arena, temp_arena;
    bool temps_done = false;              // the scratch is no longer needed once the work enqueued so far has run: released at collect
    StageVec<uint8_t> raw_stage;          // page-locked copy of the caller's records that the batch needs (source of the one host-to-device copy)
    uvc::PrepView prep_view;              // arrays of the staging kernels that outlive them (reads, fragments, families)
    int32_t *indelphred0 = nullptr;       // compute side: indelphred of every position before the threshold pass adjusts it (dump hook)
    bool groups_on_host = false;          // hb.reads / frags / fams / frag_reads downloaded (dump hook only)
    std::vector<std::pair<void*, size_t>> alloc_sizes;
    std::vector<uvcgpu_reads_soa> sources; // caller's SoA buffers (borrowed until release)
    std::vector<int32_t> tile_source;
    std::vector<std::vector<std::string>> vcf_text;   // per tile: its VCF body as the strings of its position ranges, in order (formatted once on all host threads)
    bool vcf_built = false;
    uvcgpu_batch_stats stats;
    bool submitted = true;                 // false while the context's worker thread still stages the batch
    int submit_rc = 0;
    std::string submit_err;
    bool collected = false;
    bool sparse_built = false;
    std::vector<TileSparse> sparse;
    StageVec<IndelEvent> ev_host;
    bool scored = false;
    std::vector<TileIndelSites> sites;
    StageVec<VarRec> recs;                               // candidate records as downloaded
    std::vector<std::vector<const VarRec*>> recs_by_tile; // the records of every tile (pointers into recs: a record is ~2 KB)
    StageVec<GvcfPos> gvcf;
    StageVec<GvcfExtra> gextra;
    // K6's outputs and the cursor of the sparse record stream travel to the host at the end of the batch's own kernels (one wait: collect)
    GvcfPos *d_gvcf = nullptr; GvcfExtra *d_gextra = nullptr;
    int32_t *rec_cursor_host = nullptr, *score_cursor_host = nullptr;   // 4 + 4 words of the context's page-locked cursor slab
    StageVec<IndelAllele> allele_table;
    std::vector<PrevAllele> prev_alleles;                 // sorted by (record slot, allele order): see score_core.cuh
#if UVC_CUDA
    cudaEvent_t ev[UVC_N_PILEUP_STAGES + 1];
    cudaEvent_t ev_done;                  // after the downloads that ride behind the batch's kernels
    bool have_events = false;
    cudaEvent_t ev_prep[2];               // around the staging kernels (stages P0/P1)
    bool have_prep_events = false;
#endif
};

} // namespace

struct SubmitJob { BatchState *bs; std::vector<uvcgpu_tile> tiles; };

struct uvcgpu_ctx {
    int device = 0;
    uvcgpu_params par;
    std::string err;
    std::map<int32_t, HostContig> contigs;
    std::map<int32_t, char*> d_contigs;   // device copies of the contigs' bases (stage P1 reads the reference there); absent = not available
    std::map<int32_t, std::string> contig_names;
    std::map<uvcgpu_ticket, std::unique_ptr<BatchState>> batches;
    uvcgpu_ticket next_ticket = 1;
    std::vector<int32_t> slip_tab;
    // compute-side copies of the per-context constant tables (uploaded once: a per-batch upload from pageable memory would make every submit
    // wait for the stream to drain)
    double *c_phred2prob = nullptr; int32_t *c_pf_tab = nullptr; int32_t *c_slip_tab = nullptr;
    int host_threads = 0;
    // page-locked words the small downloads of a batch land in (cursors of the sparse record stream and of the scoring pipeline): a download
    // into pageable memory would block the enqueuing thread until the stream gets there. UVC_CURSOR_SLOTS batches may be alive per context.
    int32_t *cursor_slab = nullptr;
    std::atomic<size_t> keep_hint{0}, temp_hint{0}, score_hint{0};  // device bytes the context's batches needed so far: arrays, staging scratch, scoring scratch (size of the first slab a batch asks for)
    double k5_groups_per_pos = 0.25, k5_cands_per_pos = 0.75;   // capacities of the scoring pipeline per position, grown to what the batches need
#if UVC_CUDA
    cudaStream_t stream = nullptr;        // submit: staging copies and the pileup kernels of every batch, in submission order
    cudaStream_t prep_stream = nullptr;   // staging kernels (stages P0/P1) of a batch and their two short synchronisations: they depend on nothing that the
                                          // pileup kernels of the previous batch compute, so they do not queue behind them
    cudaStream_t post_stream = nullptr;   // everything after collect (scoring kernels, downloads) of a batch, so that it does not queue behind the next batch
    cudaEvent_t wait_ev[3] = {nullptr, nullptr, nullptr};   // blocking-sync events the host waits on (one per stream)
    // Submits are asynchronous: the caller's thread copies nothing and waits for nothing, the context's worker thread stages the batch (record
    // copy, upload, staging kernels and their two short waits, allocations, pileup launches) in submission order while the caller finishes
    // the previous batch. Failures of the worker are reported by uvcgpu_collect.
    std::thread worker;
    std::mutex wmu;
    std::condition_variable wcv, wdone;
    std::deque<SubmitJob> wq;
    bool wstop = false;
#endif
};
#if UVC_CUDA
// the stream the backend helpers use on this thread: `stream` (the default at every entry point), `prep_stream` / `stream` on the worker
// thread, `post_stream` inside PostScope
static thread_local cudaStream_t t_active = nullptr;
#endif
// where error messages of this thread go: the context's string, or the batch's own one on the worker thread
static thread_local std::string *t_err = nullptr;
#define UVC_ERR(ctx) (*(t_err ? t_err : &(ctx)->err))

// ------------------------------------------------------------------------------------------------ staging memory (see host_prep.h)
// Page-locking memory is slow (~1 GB/s), so it is never done on the critical path: a request that finds no cached page-locked block of its
// size class is served from pageable memory at once, and a background thread page-locks a block of that class for the NEXT batch. Short
// runs therefore start immediately, long runs converge to fully page-locked staging (asynchronous DMA both ways).
#include <type_traits>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <unordered_map>
namespace {
const size_t kStageSmall = (size_t)1 << 20;
// upper bound of page-locked staging memory: UVCGPU_MAX_PINNED_GB (default 24), never more than a quarter of the physical memory
static size_t stage_max_pinned() {
    static const size_t v = []() {
        size_t gb = 24;
        const char *e = getenv("UVCGPU_MAX_PINNED_GB");
        if (e && atoi(e) >= 0) { gb = (size_t)atoi(e); }
        size_t cap = gb << 30;
        const long pages = sysconf(_SC_PHYS_PAGES), psz = sysconf(_SC_PAGE_SIZE);
        if (pages > 0 && psz > 0) { cap = std::min(cap, (size_t)pages * (size_t)psz / 4); }
        return cap;
    }();
    return v;
}
struct StageState {
    std::mutex mu;
    std::condition_variable cv;
    std::multimap<size_t, void*> free_blocks;         // size class -> cached page-locked blocks
    std::unordered_map<void*, size_t> pinned;         // every page-locked block (handed out or cached) -> its size class
    std::deque<size_t> wanted;                        // size classes the provisioning thread should page-lock
    size_t total_pinned = 0;
    int in_progress = 0;                              // blocks being page-locked right now
    bool thread_started = false, stop = false;
    std::thread worker;
};
StageState & stage_state() { static StageState *s = new StageState(); return *s; }   // leaked on purpose: used by a detached thread
size_t stage_class(size_t bytes) {                    // next of {1, 1.25, 1.5, 1.75} x 2^k
    size_t p2 = kStageSmall;
    while (p2 * 2 <= bytes) { p2 *= 2; }
    for (int q = 4; q <= 8; q++) { const size_t c = p2 / 4 * q; if (c >= bytes) { return c; } }
    return p2 * 2;
}
#if UVC_CUDA
void stage_provision_loop() {
    StageState & st = stage_state();
    for (;;) {
        size_t cls = 0;
        {
            std::unique_lock<std::mutex> lk(st.mu);
            st.cv.wait(lk, [&]() { return st.stop || !st.wanted.empty(); });
            if (st.stop) { return; }
            cls = st.wanted.front(); st.wanted.pop_front();
            if (st.total_pinned + cls > stage_max_pinned()) { continue; }
            st.total_pinned += cls;
            st.in_progress = 1;
        }
        void *p = NULL;
        if (cudaHostAlloc(&p, cls, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); std::lock_guard<std::mutex> lk(st.mu); st.total_pinned -= cls; st.in_progress = 0; continue; }
        std::lock_guard<std::mutex> lk(st.mu);
        st.in_progress = 0;
        st.pinned[p] = cls;
        st.free_blocks.insert(std::make_pair(cls, p));
    }
}
#endif
}
static thread_local bool g_stage_pinning = true;
void uvc_stage_thread_pinning(bool enabled) { g_stage_pinning = enabled; }
void *uvc_stage_alloc(size_t bytes) {
    if (0 == bytes) { bytes = 1; }
    // (small blocks are page-locked too, in the smallest class: they are the targets of asynchronous downloads - sparse records, indel events,
    //  candidate records - and a download into pageable memory blocks the enqueuing thread inside the driver)
    if (!g_stage_pinning) { return (getenv("UVC_DEBUG_FILL") ? memset(malloc(bytes), atoi(getenv("UVC_DEBUG_FILL")), bytes) : malloc(bytes)); }
#if UVC_CUDA
    StageState & st = stage_state();
    const size_t cls = stage_class(bytes);
    void *larger = NULL;
    {
        std::lock_guard<std::mutex> lk(st.mu);
        auto it = st.free_blocks.find(cls);
        if (it != st.free_blocks.end()) { void *p = it->second; st.free_blocks.erase(it); return p; }
        // no block of this class: a cached block of a somewhat larger class (up to 4x) serves too - page-locking a new one stalls every
        // driver call of the process while it lasts, and pageable staging makes the copies synchronous
        it = st.free_blocks.lower_bound(cls);
        if (it != st.free_blocks.end() && it->first <= 4 * cls) { larger = it->second; st.free_blocks.erase(it); return larger; }
        // one block for this request and a spare (the number of batches alive at once varies with the caller's pipelining), unless the same
        // class is already queued twice: concurrent misses of one class must not queue a pile of blocks that nobody will use
        int queued = 0;
        for (size_t c : st.wanted) { if (c == cls) { queued++; } }
        for (; queued < 2; queued++) { st.wanted.push_back(cls); }
        if (!st.thread_started && !st.stop) {
            st.thread_started = true;
            st.worker = std::thread(stage_provision_loop);
            // joined at exit, before the CUDA runtime (initialised earlier, hence torn down later) goes away
            atexit([]() { StageState & s2 = stage_state(); { std::lock_guard<std::mutex> lk2(s2.mu); s2.stop = true; } s2.cv.notify_all(); if (s2.worker.joinable()) { s2.worker.join(); } });
        }
    }
    st.cv.notify_one();
#endif
    return (getenv("UVC_DEBUG_FILL") ? memset(malloc(bytes), atoi(getenv("UVC_DEBUG_FILL")), bytes) : malloc(bytes));   // debug aid: poison fresh staging memory
}
void uvc_stage_free(void *p, size_t bytes) {
    if (NULL == p) { return; }
    (void)bytes;
#if UVC_CUDA
    {
        StageState & st = stage_state();
        std::lock_guard<std::mutex> lk(st.mu);
        auto it = st.pinned.find(p);
        if (it != st.pinned.end()) { st.free_blocks.insert(std::make_pair(it->second, p)); return; }
    }
#endif
    free(p);
}

// ------------------------------------------------------------------------------------------------ where the host's time goes
// Process-wide totals of the time host threads spend inside the driver, by kind of call (uvcgpu_host_call_stats): with several contexts per GPU
// the calls of all threads meet at the driver's locks, and an end-to-end run is as fast as these stay short.
enum { UVC_T_MALLOC, UVC_T_FREE, UVC_T_MEMSET, UVC_T_COPY, UVC_T_LAUNCH, UVC_T_WAIT, UVC_T_HOSTCOPY, UVC_T_N };
static std::atomic<int64_t> g_call_ns[UVC_T_N], g_call_n[UVC_T_N], g_call_max_ns[UVC_T_N];
struct CallTimer {
    int kind; std::chrono::steady_clock::time_point t0;
    explicit CallTimer(int k) : kind(k), t0(std::chrono::steady_clock::now()) {}
    ~CallTimer() {
        const int64_t ns = std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
        g_call_ns[kind].fetch_add(ns, std::memory_order_relaxed); g_call_n[kind].fetch_add(1, std::memory_order_relaxed);
        int64_t m = g_call_max_ns[kind].load(std::memory_order_relaxed);
        while (ns > m && !g_call_max_ns[kind].compare_exchange_weak(m, ns, std::memory_order_relaxed)) {}
    }
};

// ------------------------------------------------------------------------------------------------ compute backend
#if UVC_CUDA

#define UVC_CUDA_CHECK(ctx, call) { cudaError_t e_ = (call); if (e_ != cudaSuccess) { UVC_ERR(ctx) = std::string(#call) + ": " + cudaGetErrorString(e_); return UVCGPU_ECUDA; } }

// One named __global__ per stage (so that profiles list them by name); every thread handles one work item.
#define UVC_DEFINE_KERNEL(name, call) \
    __global__ void __launch_bounds__(128) name(const BatchView v, int64_t n) { \
        const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; \
        if (i < n) { call; } \
    }
// Position kernels: every thread owns one reference position; the 32 lanes of a warp walk the UNION of their read windows in lock-step
// (kernels_core.cuh, struct Win), which turns the per-read loads into warp-wide broadcasts and the per-position loads into consecutive addresses.
__device__ __forceinline__ void uvc_warp_window(const BatchView & v, int64_t gp, bool active, uvc::Win & w) {
    w.lo = 0; w.hi = 0;
    if (active) { uvc::position_window(v, gp, w); }
    const int32_t lo32 = (active && w.hi > w.lo ? (int32_t)w.lo : INT32_MAX), hi32 = (active && w.hi > w.lo ? (int32_t)w.hi : 0);
    w.ulo = (int64_t)__reduce_min_sync(0xffffffffu, lo32);
    w.uhi = (int64_t)__reduce_max_sync(0xffffffffu, hi32);
}
#define UVC_DEFINE_POS_KERNEL(name, call) \
    __global__ void __launch_bounds__(128) name(const BatchView v, int64_t n) { \
        const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; \
        uvc::Win w; \
        uvc_warp_window(v, i, i < n, w); \
        if (i < n) { call; } \
    }
// K0: one thread per read, a warp = 32 consecutive reads. The per-base loops of a read (quality fix-ups, mismatch runs) touch its bases and
// qualities byte by byte; a warp therefore first copies the bytes of its 32 reads into shared memory with coalesced loads (lane = byte), every
// lane then works on its own row, and the corrected qualities go back with coalesced stores. Rows are padded to an odd number of words, so
// that the 32 lanes reading byte i of their own rows hit different banks. Reads longer than the row work in place.
#define UVC_K0_MAXQ 160
#define UVC_K0_QSTRIDE 164      // 41 words
#define UVC_K0_SSTRIDE 84       // 21 words (80 bytes of packed bases)
__global__ void __launch_bounds__(128) uvc_k0_read_consts(const BatchView v, int64_t n) {
    __shared__ __align__(16) uint8_t s_qual[4][32 * UVC_K0_QSTRIDE];
    __shared__ __align__(16) uint8_t s_seq[4][32 * UVC_K0_SSTRIDE];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t r0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) - lane;     // first read of the warp
    if (r0 >= n) { return; }
    const int nr = (int)(n - r0 < 32 ? n - r0 : 32);
    uint8_t *sq = s_qual[warp], *ss = s_seq[warp];
    // every lane fetches the length and the byte addresses of its own read once; the copy loop takes them from there by shuffle, so that its
    // iterations carry no dependent loads and the byte loads of several reads are in flight together
    int32_t my_l = 0;
    const uint8_t *my_gq = v.qual_raw, *my_gs = v.seq;
    uint8_t *my_out = v.qual;
    if (lane < nr) {
        const ReadRec & R = v.reads[r0 + lane];
        my_l = R.l_qseq; my_gq = v.qual_raw + v.raw_qual_off[R.raw]; my_gs = v.seq + R.seq_off; my_out = v.qual + R.qual_off;
    }
    #pragma unroll 4
    for (int k = 0; k < nr; k++) {
        const int32_t l = __shfl_sync(0xffffffffu, my_l, k);
        const uint8_t *gq = (const uint8_t*)__shfl_sync(0xffffffffu, (unsigned long long)my_gq, k), *gs = (const uint8_t*)__shfl_sync(0xffffffffu, (unsigned long long)my_gs, k);
        if (l > UVC_K0_MAXQ) { continue; }
        for (int32_t i = lane; i < l; i += 32) { sq[k * UVC_K0_QSTRIDE + i] = gq[i]; }
        for (int32_t i = lane; i < (l + 1) / 2; i += 32) { ss[k * UVC_K0_SSTRIDE + i] = gs[i]; }
    }
    __syncwarp();
    const int64_t ri = r0 + lane;
    if (lane < nr) {
        const bool staged = (v.reads[ri].l_qseq <= UVC_K0_MAXQ);
        uvc::k0_read(v, ri, staged ? ss + lane * UVC_K0_SSTRIDE : NULL, staged ? sq + lane * UVC_K0_QSTRIDE : NULL);
    }
    __syncwarp();
    for (int k = 0; k < nr; k++) {
        const int32_t l = __shfl_sync(0xffffffffu, my_l, k);
        uint8_t *gq = (uint8_t*)__shfl_sync(0xffffffffu, (unsigned long long)my_out, k);
        if (l > UVC_K0_MAXQ) { continue; }
        for (int32_t i = lane; i < l; i += 32) { gq[i] = sq[k * UVC_K0_QSTRIDE + i]; }
    }
}
// ---- warp-private staging of per-read records in shared memory, double-buffered with cp.async (LDGSTS)
// A warp walks its union window in chunks of UVC_STAGE_READS reads. Chunk bases are multiples of 4 reads, so that every record array slice
// starts on a 16-byte boundary (record sizes are multiples of 4 bytes) and is copied with 16-byte asynchronous copies; the copy of chunk
// i + 1 is in flight while chunk i is processed.
#ifndef UVC_STAGE_READS
#define UVC_STAGE_READS 16    // a multiple of 8 (gather groups) and of 4 (16-byte slices)
#endif
__device__ __forceinline__ void uvc_cp_async16(void *smem_dst, const void *gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void uvc_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void uvc_cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }
// copies records [first, first + n) of src to dst (n <= UVC_STAGE_READS; first is a multiple of 4; the slice is rounded up to 16 bytes: the
// arrays are allocated with slack)
template <class T> __device__ __forceinline__ void uvc_warp_stage_async(T *dst, const T *src, int64_t first, int n, int lane) {
    static_assert(sizeof(T) % 4 == 0, "records are word multiples");
    const int n16 = (n * (int)sizeof(T) + 15) / 16;
    const char *g = (const char*)(src + first);
    char *d = (char*)dst;
    for (int k = lane; k < n16; k += 32) { uvc_cp_async16(d + 16 * k, g + 16 * k); }
}

// ---- bulk asynchronous copies (TMA unit, cp.async.bulk) completed on an mbarrier: one elected lane moves a whole chunk of records
__device__ __forceinline__ void uvc_mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void uvc_mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void uvc_mbar_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
// bytes is a multiple of 16; source and destination are 16-byte aligned
__device__ __forceinline__ void uvc_bulk_g2s(void *smem_dst, const void *gmem_src, unsigned bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
            :: "r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void uvc_mbar_wait(uint64_t *bar, unsigned parity) {
    asm volatile("{\n"
                 ".reg .pred P1;\n"
                 "LAB_WAIT:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
                 "@P1 bra DONE;\n"
                 "bra LAB_WAIT;\n"
                 "DONE:\n"
                 "}" :: "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}

// gather of the (base, quality) byte pairs of a staged chunk of compact records, in groups of 8 reads: all 16 byte loads of a group are issued
// before the first one is consumed (memory-level parallelism instead of one dependent load pair per read)
__device__ __forceinline__ void uvc_gather_bases_p(const BatchView & v, const PileRec *sP, uint16_t (*bq)[32], int nc, int64_t cb, const uvc::Win & w, int32_t p, bool active, int lane) {
    for (int k0 = 0; k0 < nc; k0 += 8) {
        int32_t qp[8];
        uint32_t sb[8], qb[8];
        #pragma unroll
        for (int j = 0; j < 8; j++) {
            const int k = (k0 + j < nc ? k0 + j : nc - 1);
            const int64_t ri = cb + k;
            const PileRec & P = sP[k];
            qp[j] = uvc::base_index(v, P, p, active && ri >= w.lo && ri < w.hi);
            const int32_t qc = (qp[j] > 0 ? qp[j] : 0);
            sb[j] = v.seq[(uint64_t)P.seq_off + (uint32_t)(qc >> 1)];
            qb[j] = v.qual[(uint64_t)P.qual_off + (uint32_t)qc];
        }
        #pragma unroll
        for (int j = 0; j < 8; j++) { bq[k0 + j][lane] = (uint16_t)uvc::pack_base(sb[j], qb[j], qp[j]); }
    }
}

// K2: one thread per position (both dense symbols, see kernels_core.cuh: K2Merged).
// Each warp stages the compact records (PileRec, 64 bytes, written by K0) of UVC_STAGE_READS reads of its union window in shared memory: ONE
// elected lane issues one bulk asynchronous copy per chunk (cp.async.bulk, completion on the warp's mbarrier), one chunk ahead; the warp
// then gathers the (base, quality) byte pairs of the whole chunk with independent loads (memory-level parallelism instead of one dependent
// load pair per read), and the per-read work runs entirely from shared memory.
struct __align__(16) K2StageP {
    PileRec P[2][UVC_STAGE_READS];
    uint16_t bq[UVC_STAGE_READS][32];
    uint64_t bar[2];
};
#ifndef UVC_K2_MINBLOCKS
#define UVC_K2_MINBLOCKS 4
#endif
__global__ void __launch_bounds__(128, UVC_K2_MINBLOCKS) uvc_k2_bias_pileup(const BatchView v, int64_t n) {
    extern __shared__ __align__(16) unsigned char uvc_smem[];
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t gp_ = (idx < n ? uvc::list_position(v, 0, idx) : -1);
    const bool active = (gp_ >= 0);
    const int64_t gp = (active ? gp_ : 0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    K2StageP & S = ((K2StageP*)uvc_smem)[warp];
    if (0 == lane) { uvc_mbar_init(&S.bar[0], 1); uvc_mbar_init(&S.bar[1], 1); uvc_mbar_fence_init(); }
    __syncwarp();
    uvc::Win w;
    uvc_warp_window(v, gp, active, w);
    uvc::K2Merged st;
    st.p = 0;
    if (active) { uvc::k2m_begin(st, v, gp); }
    const int64_t c0 = w.ulo;
    auto issue = [&](int64_t cb, int buf) {
        if (cb < w.uhi && 0 == lane) {
            const unsigned bytes = (unsigned)((w.uhi - cb < UVC_STAGE_READS ? w.uhi - cb : UVC_STAGE_READS) * sizeof(PileRec));
            uvc_mbar_expect_tx(&S.bar[buf], bytes);
            uvc_bulk_g2s(S.P[buf], v.prec + cb, bytes, &S.bar[buf]);
        }
    };
    issue(c0, 0);
    unsigned phase0 = 0, phase1 = 0;
    int buf = 0;
    // the lane's own window relative to the chunk base, as 32-bit numbers
    for (int64_t cb = c0; cb < w.uhi; cb += UVC_STAGE_READS, buf ^= 1) {
        const int nc = (int)(w.uhi - cb < UVC_STAGE_READS ? w.uhi - cb : UVC_STAGE_READS);
        issue(cb + UVC_STAGE_READS, buf ^ 1);          // (every lane finished with that buffer before the __syncwarp that ended the previous iteration)
        if (buf) { uvc_mbar_wait(&S.bar[1], phase1); phase1 ^= 1u; } else { uvc_mbar_wait(&S.bar[0], phase0); phase0 ^= 1u; }
        const PileRec *sP = S.P[buf];
        uvc_gather_bases_p(v, sP, S.bq, nc, cb, w, st.p, active, lane);
        if (active) {
            const int k_lo = (int)(w.lo - cb > 0 ? (w.lo - cb < nc ? w.lo - cb : nc) : 0), k_hi = (int)(w.hi - cb < nc ? (w.hi - cb > 0 ? w.hi - cb : 0) : nc);
            for (int k = k_lo; k < k_hi; k++) { uvc::k2m_read(st, v, sP[k], (uint32_t)S.bq[k][lane]); }
        }
        __syncwarp();
    }
    if (active) { uvc::k2m_flush(st, v); }
}
// K1: one thread per position; the compact records of UVC_STAGE_READS reads per warp (PileRec + PrepRec, two bulk asynchronous copies on one
// mbarrier, one chunk ahead) and the same grouped byte gather as K2
struct __align__(16) K1StageP {
    PileRec P[2][UVC_STAGE_READS];
    PrepRec Q[2][UVC_STAGE_READS];
    uint16_t bq[UVC_STAGE_READS][32];
    uint64_t bar[2];
};
#ifndef UVC_K1_MINBLOCKS
#define UVC_K1_MINBLOCKS 4
#endif
__global__ void __launch_bounds__(128, UVC_K1_MINBLOCKS) uvc_k1_prep_thres(const BatchView v, int64_t n) {
    extern __shared__ __align__(16) unsigned char uvc_smem[];
    const int64_t gp = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    K1StageP & S = ((K1StageP*)uvc_smem)[warp];
    const bool active = (gp < n);
    if (0 == lane) { uvc_mbar_init(&S.bar[0], 1); uvc_mbar_init(&S.bar[1], 1); uvc_mbar_fence_init(); }
    __syncwarp();
    uvc::Win w;
    uvc_warp_window(v, gp, active, w);
    uvc::K1State st;
    st.p = 0;
    if (active) { uvc::k1_begin(st, v, gp); }
    const int64_t c0 = w.ulo;
    auto issue = [&](int64_t cb, int buf) {
        if (cb < w.uhi && 0 == lane) {
            const unsigned cnt = (unsigned)(w.uhi - cb < UVC_STAGE_READS ? w.uhi - cb : UVC_STAGE_READS);
            uvc_mbar_expect_tx(&S.bar[buf], cnt * (unsigned)(sizeof(PileRec) + sizeof(PrepRec)));
            uvc_bulk_g2s(S.P[buf], v.prec + cb, cnt * (unsigned)sizeof(PileRec), &S.bar[buf]);
            uvc_bulk_g2s(S.Q[buf], v.qrec + cb, cnt * (unsigned)sizeof(PrepRec), &S.bar[buf]);
        }
    };
    issue(c0, 0);
    unsigned phase0 = 0, phase1 = 0;
    int buf = 0;
    for (int64_t cb = c0; cb < w.uhi; cb += UVC_STAGE_READS, buf ^= 1) {
        const int nc = (int)(w.uhi - cb < UVC_STAGE_READS ? w.uhi - cb : UVC_STAGE_READS);
        issue(cb + UVC_STAGE_READS, buf ^ 1);
        if (buf) { uvc_mbar_wait(&S.bar[1], phase1); phase1 ^= 1u; } else { uvc_mbar_wait(&S.bar[0], phase0); phase0 ^= 1u; }
        const PileRec *sP = S.P[buf];
        const PrepRec *sQ = S.Q[buf];
        uvc_gather_bases_p(v, sP, S.bq, nc, cb, w, st.p, active, lane);
        if (active) {
            for (int k = 0; k < nc; k++) { uvc::k1_read(st, v, sP[k], sQ[k], (uint32_t)S.bq[k][lane]); }   // lanes outside their own window hold NOBASE
        }
        __syncwarp();
    }
    if (active) { uvc::k1_end(st, v); }
}
UVC_DEFINE_KERNEL(uvc_k1b_noindel, uvc::k1b_noindel(v, i))
UVC_DEFINE_KERNEL(uvc_k2e_indel_events, uvc::k2e_event(v, i))
// KF: one warp per fragment walks the fragment's column chunk by chunk (lane = position within the chunk): the records of the fragment's
// reads are loaded once per fragment instead of once per entry; a chunk's four bit masks are four ballots.
#ifndef UVC_KF_MINBLOCKS
#define UVC_KF_MINBLOCKS 12    // latency-bound on the loads of base qualities: measured per sub-batch 3.10 ms at 93 registers, 2.41 at 64 (8 blocks per SM), 2.01 at 42 (12 blocks)
#endif
__global__ void __launch_bounds__(128, UVC_KF_MINBLOCKS) uvc_kf_fragment_columns(const BatchView v, int64_t n_threads) {
    const int64_t fi = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (fi >= (n_threads >> 5)) { return; }
    const FragRec & G = v.frags[fi];
    const int32_t lo = G.lo, n = G.hi - G.lo;
    const int64_t col_off = G.col_off;
    const TileInfo & T = v.tiles[G.tile];
    const int64_t gp0 = T.pos_off + (lo - T.ext_beg);
    const bool plain = uvc::kf_fragment_is_plain(v, G);
    const int32_t n_reads = G.n_reads;
    uvc::KfRead rd[2];
    if (plain) {
        if (n_reads > 0) { uvc::kf_load_read(rd[0], v, v.frag_reads[G.read_off]); }
        if (n_reads > 1) { uvc::kf_load_read(rd[1], v, v.frag_reads[G.read_off + 1]); }
    }
    for (int32_t c = 0; c * UVC_COL_CHUNK < n; c++) {
        const int32_t o = c * UVC_COL_CHUNK + lane;
        const int64_t i = col_off + o;
        uint32_t bits = 0;
        if (o < n) { bits = (plain ? uvc::kf_plain_entry(v, i, lo + o, gp0 + o, rd, n_reads) : uvc::kf_fragment_column(v, i)); }
        const uint32_t m0 = __ballot_sync(0xffffffffu, bits & 1u), m1 = __ballot_sync(0xffffffffu, bits & 2u);
        const uint32_t m2 = __ballot_sync(0xffffffffu, bits & 4u), m3 = __ballot_sync(0xffffffffu, bits & 8u);
        if (0 == lane) { *(uint4*)(v.fmask + (i / UVC_COL_CHUNK) * 4) = make_uint4(m0, m1, m2, m3); }
    }
}
UVC_DEFINE_KERNEL(uvc_k3a_fragment_stats, uvc::k3a_fragment(v, i))
UVC_DEFINE_KERNEL(uvc_km_family_columns, uvc::km_family_column(v, i))
UVC_DEFINE_KERNEL(uvc_k4a_family_ends, uvc::k4a_family_strand(v, i))
// ---- column-entry pipeline of K3b and K4
// Both kernels walk the warp's union window in chunks of UVC_COL_READS reads. Per chunk a warp needs (1) the compact 32-byte records of the
// reads and (2) for every lane the 8-byte column entry of each read at the lane's position. Both are fetched with cp.async: records two chunks
// ahead (ring of 3), entries one chunk ahead (ring of 2, issued as soon as the chunk's records have landed), so the entry loads of chunk
// i + 1 are in flight during all of the work on chunk i. Commit order: R0 R1 E0 | R2 E1 | R3 E2 | ...; waiting for "all but the most recent
// group" after committing R(i+2) guarantees R(i+1) and E(i).
#ifndef UVC_COL_READS
#define UVC_COL_READS 8     // measured (tools/gpu_variants.sh): 8-read chunks leave room for five K3b blocks per SM (3.70 vs 4.05 ms per sub-batch)
#endif
__device__ __forceinline__ void uvc_cp_async8(void *smem_dst, const void *gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(d), "l"(gmem_src) : "memory");
}
template <class Q> struct __align__(16) ColStage {
    Q q[3][UVC_COL_READS];
    FragCol e[2][UVC_COL_READS][32];
};
__device__ __forceinline__ int uvc_chunk_len(int64_t cb, int64_t uhi) { return (int)(uhi - cb < UVC_COL_READS ? uhi - cb : UVC_COL_READS); }

// K3b: records = ReadFrag (written by K3a). The quality histograms of the two hot symbols live in shared memory ([bucket][thread]: conflict-free,
// updated with reductions), their depth counters in registers.
#ifndef UVC_K3B_MINBLOCKS
#define UVC_K3B_MINBLOCKS 6    // 80 registers; six blocks per SM with 8-read chunks (2.88 vs 3.05 ms per sub-batch at five)
#endif
__global__ void __launch_bounds__(128, UVC_K3B_MINBLOCKS) uvc_k3b_fragment_consensus(const BatchView v, int64_t n) {
    extern __shared__ __align__(16) unsigned char uvc_smem[];
    typedef ColStage<ReadFrag> Stage;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t gp_ = (idx < n ? uvc::list_position(v, 1, idx) : -1);
    const bool active = (gp_ >= 0);
    const int64_t gp = (active ? gp_ : 0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    Stage & S = ((Stage*)uvc_smem)[warp];
    int32_t *hot_buckets = (int32_t*)(uvc_smem + (blockDim.x >> 5) * sizeof(Stage)) + threadIdx.x;
    uvc::Win w;
    uvc_warp_window(v, gp, active, w);
    uvc::K3bState st;
    uvc::K3bArrays arr;
    st.p = 0;
    if (active) { uvc::k3b_begin(st, arr, v, gp, hot_buckets, (int)blockDim.x); }
    const int32_t p = st.p;
    const int64_t c0 = w.ulo & ~(int64_t)3;
    auto issue_records = [&](int64_t cb, int slot) { if (cb < w.uhi) { uvc_warp_stage_async(S.q[slot], v.rfrag, cb, uvc_chunk_len(cb, w.uhi), lane); } uvc_cp_async_commit(); };
    auto issue_entries = [&](int64_t cb, int rslot, int eslot) {
        if (cb < w.uhi && active) {
            const int nc = uvc_chunk_len(cb, w.uhi);
            for (int k = 0; k < nc; k++) {
                const int64_t ri = cb + k;
                const ReadFrag & q = S.q[rslot][k];
                if (ri >= w.lo && ri < w.hi && q.rend > p && q.fragprev_maxrend <= p) { uvc_cp_async8(&S.e[eslot][k][lane], v.fcol + (q.col_base + p)); }
            }
        }
        uvc_cp_async_commit();
    };
    issue_records(c0, 0);
    issue_records(c0 + UVC_COL_READS, 1);
    uvc_cp_async_wait<1>();
    __syncwarp();
    issue_entries(c0, 0, 0);
    int i = 0;
    for (int64_t cb = c0; cb < w.uhi; cb += UVC_COL_READS, i++) {
        const int rs = i % 3, es = i & 1;
        issue_records(cb + 2 * UVC_COL_READS, (i + 2) % 3);
        uvc_cp_async_wait<1>();
        __syncwarp();
        issue_entries(cb + UVC_COL_READS, (i + 1) % 3, es ^ 1);
        if (active) {
            const int nc = uvc_chunk_len(cb, w.uhi);
            for (int k = 0; k < nc; k++) {
                const int64_t ri = cb + k;
                if (ri < w.lo || ri >= w.hi) { continue; }
                const ReadFrag q = S.q[rs][k];
                if (q.rend <= p || q.fragprev_maxrend > p) { continue; }
                uvc::k3b_read(st, v, q, S.e[es][k][lane]);
            }
        }
        __syncwarp();
    }
    uvc_cp_async_wait<0>();
    if (active) { uvc::k3b_end(st, v); }
}
// K4: records = ReadFam (built on the host). Two shapes of the column-entry pipeline:
//   narrow (non-UMI data: nearly every (family, strand) is a single fragment): 16 reads per chunk, 8-byte fragment entries through the
//          pipeline, the few 32-byte entries of multi-fragment families read in place;
//   wide   (UMI data: most strands hold several fragments): 8 reads per chunk, 32-byte slots that receive either entry kind, so that the
//          family entries of both walks of the window are in flight one chunk ahead as well.
#ifndef UVC_K4_MINBLOCKS
#define UVC_K4_MINBLOCKS 4     // 128 registers: four blocks per SM
#endif
template <bool kWide> struct __align__(16) K4StageT {
    static const int kReads = (kWide ? 8 : UVC_COL_READS);
    ReadFam q[3][kReads];
    typename std::conditional<kWide, FamCol, FragCol>::type e[2][kReads][32];
};
__device__ __forceinline__ void uvc_cp_async16x2(void *smem_dst, const void *gmem_src) { uvc_cp_async16(smem_dst, gmem_src); uvc_cp_async16((char*)smem_dst + 16, (const char*)gmem_src + 16); }

template <bool kWide> __device__ __forceinline__ void uvc_k4_body(const BatchView & v, int64_t n) {
    extern __shared__ __align__(16) unsigned char uvc_smem[];
    typedef K4StageT<kWide> Stage;
    const int kReads = Stage::kReads;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t gp_ = (idx < n ? uvc::list_position(v, 1, idx) : -1);
    const bool active = (gp_ >= 0);
    const int64_t gp = (active ? gp_ : 0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    Stage & S = ((Stage*)uvc_smem)[warp];
    uvc::Win w;
    uvc_warp_window(v, gp, active, w);
    uvc::K4State st;
    uvc::K4Arrays arr;
    st.p = 0; st.n_need2 = 0;
    if (active) { uvc::k4_begin(st, arr, v, gp); }
    const int32_t p = st.p;
    const int64_t c0 = w.ulo & ~(int64_t)3;
    auto chunk_len = [&](int64_t cb) { return (int)(w.uhi - cb < kReads ? w.uhi - cb : kReads); };
    auto issue_records = [&](int64_t cb, int slot) { if (cb < w.uhi) { uvc_warp_stage_async(S.q[slot], v.rfam, cb, chunk_len(cb), lane); } uvc_cp_async_commit(); };
    // entries of the reads of this lane's own window that are the first read of their (family, strand) at p
    auto issue_entries = [&](int64_t cb, int rslot, int eslot, bool wanted) {
        if (cb < w.uhi && wanted) {
            const int nc = chunk_len(cb);
            for (int k = 0; k < nc; k++) {
                const int64_t ri = cb + k;
                const ReadFam & q = S.q[rslot][k];
                if (!(ri >= w.lo && ri < w.hi && q.rend > p && q.famprev_maxrend <= p)) { continue; }
                if (q.flags & UVC_RF_DIRECT) { uvc_cp_async8(&S.e[eslot][k][lane], v.fcol + (q.col_base + p)); }
                else if (kWide) { uvc_cp_async16x2(&S.e[eslot][k][lane], v.mcol + (q.col_base + p)); }
            }
        }
        uvc_cp_async_commit();
    };
    // the (family, strand) entry of staged read q: from the slot, or (narrow shape, multi-fragment strand) from global memory
    auto entry_of = [&](const ReadFam & q, int eslot, int k) -> FamCol {
        if (q.flags & UVC_RF_DIRECT) { return uvc::famcol_from_frag(*(const FragCol*)&S.e[eslot][k][lane], v.par); }
        if (kWide) { return *(const FamCol*)&S.e[eslot][k][lane]; }
        return v.mcol[q.col_base + p];
    };
    // one walk of the window through the pipeline; pass 0 = loop 1 (with the share of loop 2 that needs nothing from other families),
    // pass 1 = loop 2 for the positions whose need-list overflowed (UMI data)
    auto walk = [&](const int pass) {
        const bool mine = (active && (pass == 0 || st.n_need2 > UVC_K4_LIST));
        issue_records(c0, 0);
        issue_records(c0 + kReads, 1);
        uvc_cp_async_wait<1>();
        __syncwarp();
        issue_entries(c0, 0, 0, mine);
        int i = 0;
        for (int64_t cb = c0; cb < w.uhi; cb += kReads, i++) {
            const int rs = i % 3, es = i & 1;
            issue_records(cb + 2 * kReads, (i + 2) % 3);
            uvc_cp_async_wait<1>();
            __syncwarp();
            issue_entries(cb + kReads, (i + 1) % 3, es ^ 1, mine);
            if (mine) {
                const int nc = chunk_len(cb);
                for (int k = 0; k < nc; k++) {
                    const int64_t ri = cb + k;
                    if (ri < w.lo || ri >= w.hi) { continue; }
                    const ReadFam q = S.q[rs][k];
                    if (q.rend <= p) { continue; }
                    if (pass == 0) {
                        if (q.famprev_maxrend > p) { continue; }
                        if ((q.flags & UVC_RF_DIRECT) && uvc::k4_loop1_lone_fragment(st, v, q, *(const FragCol*)&S.e[es][k][lane])) { continue; }
                        uvc::k4_loop1_read(st, v, q, entry_of(q, es, k), ri - w.lo);
                    } else if (q.famprev_maxrend <= p) {
                        const FamCol m = entry_of(q, es, k);
                        uvc::k4_loop2_read(st, v, q, &m);
                    } else {
                        uvc::k4_loop2_read(st, v, q, NULL);     // (not the first read of its strand here: nothing to fetch, nothing to do)
                    }
                }
            }
            __syncwarp();
        }
        uvc_cp_async_wait<0>();
        __syncwarp();
    };
    walk(0);
    // loop 2 for what is left: the short list first; positions whose list overflowed walk the window again
    if (active) { uvc::k4_loop2_listed(st, v, w); }
    if (__any_sync(0xffffffffu, active && st.n_need2 > UVC_K4_LIST)) { walk(1); }
    if (active) { uvc::k4_end(st, v); }
}
__global__ void __launch_bounds__(128, UVC_K4_MINBLOCKS) uvc_k4_family_consensus(const BatchView v, int64_t n) { uvc_k4_body<false>(v, n); }
__global__ void __launch_bounds__(128, 3) uvc_k4_family_consensus_umi(const BatchView v, int64_t n) { uvc_k4_body<true>(v, n); }
UVC_DEFINE_KERNEL(uvc_k4c_family_haplotypes, uvc::k4c_family_strand(v, i))

// scoring stage: K6 one thread per extended position, K5 one thread per zero-based position (heavy local state: 64 threads per block)
__global__ void __launch_bounds__(128) uvc_k6_gvcf_inputs(const BatchView v, const ScoreView sv, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { uvc::k6_gvcf_position(v, sv, i); }
}
__global__ void __launch_bounds__(128) uvc_k5a_flag_candidates(const BatchView v, const ScoreView sv, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { uvc::k5a_flag_position(v, sv, i); }
}
// The scoring pipeline (score_core.cuh): grids are sized for the capacities, threads beyond the counts exit at once. A kernel whose input
// overflowed its capacity does nothing: the host sees the counts and runs the pipeline again with room for everything.
__device__ __forceinline__ bool uvc_k5_fits(const ScoreView & sv) { return sv.out_cursor[2] <= sv.group_cap && sv.out_cursor[3] <= sv.cand_cap; }
__global__ void __launch_bounds__(128) uvc_k5b_list_candidates(const BatchView v, const ScoreView sv, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && i < (int64_t)*sv.cand_cursor) { uvc::k5b_list_position(v, sv, (int64_t)sv.cand_list[i]); }
}
__global__ void __launch_bounds__(128) uvc_k5g_group_init(const BatchView v, const ScoreView sv, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && i < (int64_t)sv.out_cursor[2] && uvc_k5_fits(sv)) { uvc::k5g_group(v, sv, i); }
}
#ifndef UVC_K5C_MINBLOCKS
#define UVC_K5C_MINBLOCKS 2    // BcfFormat_symbol_calc_DPv is ~60 double-precision allele fractions per candidate
#endif
__global__ void __launch_bounds__(128, UVC_K5C_MINBLOCKS) uvc_k5c_candidate_depths(const BatchView v, const ScoreView sv, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && i < (int64_t)sv.out_cursor[3] && uvc_k5_fits(sv)) { uvc::k5c_candidate(v, sv, i); }
}
__global__ void __launch_bounds__(128) uvc_k5e_candidate_quals(const BatchView v, const ScoreView sv, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && i < (int64_t)sv.out_cursor[3] && uvc_k5_fits(sv)) { uvc::k5e_candidate(v, sv, i); }
}
__global__ void __launch_bounds__(128) uvc_k5f_group_records(const BatchView v, const ScoreView sv, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && i < (int64_t)sv.out_cursor[2] && uvc_k5_fits(sv)) { uvc::k5f_group(v, sv, i); }
}

__global__ void uvc_selftest_math_kernel(const BatchView v, int32_t which, const double *in, int32_t n, double *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { out[i] = uvc::selftest_math(v, which, in[3 * i], in[3 * i + 1], in[3 * i + 2]); }
}

typedef void (*uvc_kernel_t)(const BatchView, int64_t);
static void launch(uvc_kernel_t k, cudaStream_t s, const BatchView & v, int64_t n, int64_t & launches) {
    if (n <= 0) { return; }
    const int threads = 128;
    k<<<(unsigned)((n + threads - 1) / threads), threads, 0, s>>>(v, n);
    launches++;
}

// Waits of the host: the event is polled with sleeps in between whose length grows with the time already waited (20 - 250 us). Measured
// alternatives: cudaEventSynchronize on a default event spins on a core per waiting thread (contexts x ranks threads wait at the same time and
// starve the record copies and the VCF text: 2-GPU e2e 39.9 M reads/s instead of 68.7 M); cudaEventBlockingSync sleeps until the driver's
// interrupt arrives, which took up to hundreds of milliseconds per wait on some boxes (1-GPU e2e 7.4 - 48.6 M reads/s from run to run).
static cudaError_t uvc_event_wait(cudaEvent_t e) {
    CallTimer ct(UVC_T_WAIT);
    const auto t0 = std::chrono::steady_clock::now();
    for (;;) {
        const cudaError_t q = cudaEventQuery(e);
        if (q != cudaErrorNotReady) { return q; }
        const int64_t us = std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count();
        if (us < 30) { std::this_thread::yield(); continue; }
        std::this_thread::sleep_for(std::chrono::microseconds(us / 8 < 20 ? 20 : (us / 8 > 250 ? 250 : us / 8)));
    }
}
static int backend_wait_stream(uvcgpu_ctx *ctx, cudaStream_t s) {
    cudaEvent_t & e = (s == ctx->stream ? ctx->wait_ev[0] : (s == ctx->prep_stream ? ctx->wait_ev[1] : ctx->wait_ev[2]));
    if (NULL == e) { UVC_CUDA_CHECK(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); }
    UVC_CUDA_CHECK(ctx, cudaEventRecord(e, s));
    UVC_CUDA_CHECK(ctx, uvc_event_wait(e));
    return 0;
}
// Device memory. The stream-ordered pool (cudaMallocAsync) was measured at milliseconds per call with several contexts allocating and freeing
// on two dozen streams (up to seconds for single calls: ~10 s of host-thread time per 26.7 M-read step), so batches carve their arrays from
// big slabs instead. A slab cache per device keeps every slab the process ever allocated; a batch takes one slab sized like the context's
// previous batches (more if it turns out larger), bump-allocates from it, and hands the slabs back when it is released - by then everything
// that used them has completed (collect waited for the pileup kernels, scoring synchronises before it returns), so the next user needs no
// stream ordering. In the steady state no driver call allocates or frees memory.
struct SlabCache {
    std::mutex mu;
    std::multimap<size_t, std::pair<char*, uint64_t>> free_slabs;   // size -> (slab, tick at which it came back)
    uint64_t tick = 0;                                              // counts acquisitions
    int live_contexts = 0;
};
static SlabCache & slab_cache(int device) { static SlabCache *c = new SlabCache[64]; return c[device & 63]; }
static void slab_purge(SlabCache & sc) {       // (caller holds the lock)
    for (auto & kv : sc.free_slabs) { cudaFree(kv.second.first); }
    sc.free_slabs.clear();
}
static int slab_acquire(uvcgpu_ctx *ctx, size_t min_bytes, BatchState::Slab & out) {
    CallTimer ct(UVC_T_MALLOC);
    SlabCache & sc = slab_cache(ctx->device);
    // size classes of 1/8 octave (and at least 64 MB): batches of slightly different sizes share slabs
    size_t unit = (size_t)64 << 20;
    while (unit * 16 <= min_bytes) { unit *= 2; }
    const size_t want = (min_bytes + unit - 1) / unit * unit;
    std::lock_guard<std::mutex> lk(sc.mu);
    sc.tick++;
    auto it = sc.free_slabs.lower_bound(want);
    if (it != sc.free_slabs.end() && it->first <= 2 * want + ((size_t)256 << 20)) { out.base = it->second.first; out.size = it->first; sc.free_slabs.erase(it); return 0; }
    // a miss (warm-up, or the batches grew): slabs nobody has asked for in a while (the sizes of the first batches) go back to the driver first
    for (auto jt = sc.free_slabs.begin(); jt != sc.free_slabs.end(); ) {
        if (sc.tick - jt->second.second > 48) { cudaFree(jt->second.first); jt = sc.free_slabs.erase(jt); } else { ++jt; }
    }
    void *p = NULL;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) { cudaGetLastError(); slab_purge(sc); e = cudaMalloc(&p, want); }       // make room: give the cached slabs back and try once more
    if (e != cudaSuccess) { cudaGetLastError(); UVC_ERR(ctx) = std::string("out of device memory: ") + cudaGetErrorString(e); return UVCGPU_ENOMEM; }
    out.base = (char*)p; out.size = want;
    return 0;
}
static void arena_release(uvcgpu_ctx *ctx, BatchState::Arena & a) {
    if (a.slabs.empty()) { return; }
    SlabCache & sc = slab_cache(ctx->device);
    std::lock_guard<std::mutex> lk(sc.mu);
    for (auto & sl : a.slabs) { sc.free_slabs.insert(std::make_pair(sl.size, std::make_pair(sl.base, sc.tick))); }
    a.slabs.clear(); a.used = 0; a.requested = 0;
}
static int arena_alloc(uvcgpu_ctx *ctx, BatchState::Arena & a, size_t hint, void **out, size_t bytes) {
    bytes = (bytes + 64 + 255) / 256 * 256;     // slack: staged record slices are rounded up to 16 bytes, and clamped byte loads may touch offset 0 of an empty blob
    a.requested += bytes;
    if (a.slabs.empty() || a.used + bytes > a.slabs.back().size) {
        BatchState::Slab sl;
        const size_t next = (a.slabs.empty() ? std::max(bytes, hint) : std::max(bytes, std::max(hint / 4, (size_t)256 << 20)));
        const int rc = slab_acquire(ctx, next, sl);
        if (rc != 0) { return rc; }
        a.slabs.push_back(sl); a.used = 0;
    }
    *out = a.slabs.back().base + a.used;
    a.used += bytes;
    return 0;
}
static int backend_alloc(uvcgpu_ctx *ctx, BatchState & bs, void **out, size_t bytes, bool zero) {
    const int rc = arena_alloc(ctx, bs.keep_arena, ctx->keep_hint.load(), out, bytes);
    if (rc != 0) { return rc; }
    if (zero) { CallTimer ct(UVC_T_MEMSET); UVC_CUDA_CHECK(ctx, cudaMemsetAsync(*out, 0, bytes + 64, t_active)); }
    return 0;
}
static int backend_upload(uvcgpu_ctx *ctx, BatchState & bs, void *dst, const void *src, size_t bytes) {
    if (bytes) { CallTimer ct(UVC_T_HOSTCOPY); UVC_CUDA_CHECK(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, t_active)); bs.stats.h2d_bytes += (int64_t)bytes; }
    return 0;
}
// downloads are enqueued (page-locked destinations) and waited for together: every wait of the host costs the turn-around of the stream behind
// the pileup blocks of other batches that occupy the SMs
static int backend_download_async(uvcgpu_ctx *ctx, void *dst, const void *src, size_t bytes) {
    if (bytes) { CallTimer ct(UVC_T_COPY); UVC_CUDA_CHECK(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, t_active)); }
    return 0;
}
static int backend_sync(uvcgpu_ctx *ctx) { return backend_wait_stream(ctx, t_active); }
// A synchronous download of a few bytes up to a few megabytes (sizes of a batch, its tiles, ...) lands in a page-locked bounce buffer of the
// calling thread: a copy into pageable memory makes the driver spin inside the call until the stream gets there (measured: ~0.5 s of a core per
// 26.7 M-read step, with the driver's lock held against every other thread).
static void *uvc_bounce(size_t bytes) {
    static thread_local void *p = nullptr;
    static thread_local size_t cap = 0;
    if (bytes > cap) {
        if (p) { cudaFreeHost(p); p = nullptr; cap = 0; }
        const size_t want = std::max(bytes, (size_t)256 << 10);
        if (cudaHostAlloc(&p, want, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); p = nullptr; return nullptr; }
        cap = want;
    }
    return p;
}
static int backend_download(uvcgpu_ctx *ctx, void *dst, const void *src, size_t bytes) {
    if (0 == bytes) { return 0; }
    void *b = (bytes <= ((size_t)4 << 20) ? uvc_bounce(bytes) : nullptr);
    int rc_ = backend_download_async(ctx, (b ? b : dst), src, bytes);
    if (0 == rc_) { rc_ = backend_sync(ctx); }
    if (rc_ != 0) { return rc_; }
    if (b) { memcpy(dst, b, bytes); }
    return 0;
}
static int backend_zero(uvcgpu_ctx *ctx, void *dst, size_t bytes) { CallTimer ct(UVC_T_MEMSET); UVC_CUDA_CHECK(ctx, cudaMemsetAsync(dst, 0, bytes, t_active)); return 0; }
// the scratch is not needed any more once the work enqueued so far has run (staging: released at collect; scoring: `now`, after its synchronisation)
static void backend_free_temps(uvcgpu_ctx *ctx, BatchState & bs, bool now = false) {
    std::atomic<size_t> & hint = (bs.collected ? ctx->score_hint : ctx->temp_hint);
    const size_t need = bs.temp_arena.requested + bs.temp_arena.requested / 8;
    if (bs.temp_arena.requested > hint.load()) { hint.store(need); }
    if (now) { arena_release(ctx, bs.temp_arena); bs.temps_done = false; } else { bs.temps_done = true; }
}
// A batch is released after everything of it has completed: its slabs go straight back to the cache.
static void backend_free(uvcgpu_ctx *ctx, BatchState & bs) {
    CallTimer ct(UVC_T_FREE);
    const size_t need = bs.keep_arena.requested + bs.keep_arena.requested / 8;
    if (bs.keep_arena.requested > ctx->keep_hint.load()) { ctx->keep_hint.store(need); }
    arena_release(ctx, bs.temp_arena); arena_release(ctx, bs.keep_arena);
}
// scratch of the staging and scoring kernels; fill >= 0: every byte is set to it
static int backend_alloc_temp(uvcgpu_ctx *ctx, BatchState & bs, void **out, size_t bytes, int fill = -1) {
    const int rc = arena_alloc(ctx, bs.temp_arena, (bs.collected ? ctx->score_hint.load() : ctx->temp_hint.load()), out, bytes);
    if (rc != 0) { return rc; }
    if (fill >= 0) { CallTimer ct(UVC_T_MEMSET); UVC_CUDA_CHECK(ctx, cudaMemsetAsync(*out, fill, bytes + 64, t_active)); }
    return 0;
}

#include "prep_device.inc"

static int backend_run(uvcgpu_ctx *ctx, BatchState & bs) {
    CallTimer ct_run(UVC_T_LAUNCH);      // (the pileup launches, their events and the three small downloads)
    const BatchView & v = bs.view;
    int64_t launches = 0;
    for (int i = 0; i < UVC_N_PILEUP_STAGES + 1; i++) { UVC_CUDA_CHECK(ctx, cudaEventCreate(&bs.ev[i])); }
    bs.have_events = true;
    int e = 0;
    // Threads per block of the position kernels (every warp is self-contained: its own staging slot, no block-wide synchronisation). A batch
    // with few positions and deep windows (a small panel at very high depth) is cut into one-warp blocks so that every SM gets some.
    // (four 128-thread blocks are resident per SM: below 148 x 4 x 128 positions smaller blocks spread the warps over the SMs more evenly)
    auto pos_block = [](int64_t n) {
        int b = (n >= (int64_t)148 * 128 * 4 ? 128 : (n >= (int64_t)148 * 64 * 2 ? 64 : 32));
        const char *f = getenv("UVC_POS_BLOCK");      // tests force every shape
        if (f && (atoi(f) == 32 || atoi(f) == 64 || atoi(f) == 128)) { b = atoi(f); }
        return b;
    };
    // K1 runs on every position; K2 and K3b / K4 on their position lists (kernels_core.cuh: tile_need_range) unless par.all_positions
    const int64_t n_k2 = (v.list_tile[0] ? v.n_list[0] : v.n_pos), n_k34 = (v.list_tile[1] ? v.n_list[1] : v.n_pos);
    int pb = pos_block(v.n_pos);
    UVC_CUDA_CHECK(ctx, cudaEventRecord(bs.ev[e++], ctx->stream));
    #define UVC_STAGE(kernel, n) { launch(kernel, ctx->stream, v, (n), launches); UVC_CUDA_CHECK(ctx, cudaEventRecord(bs.ev[e++], ctx->stream)); }
    UVC_STAGE(uvc_k0_read_consts, v.n_reads)
    if (v.n_pos > 0) {
        static_assert(sizeof(K1StageP) % 16 == 0 && sizeof(PrepRec) == 32, "per-warp staging slots keep 16-byte alignment");
        const size_t smem = 4 * sizeof(K1StageP);
        UVC_CUDA_CHECK(ctx, cudaFuncSetAttribute(uvc_k1_prep_thres, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        uvc_k1_prep_thres<<<(unsigned)((v.n_pos + pb - 1) / pb), pb, smem * pb / 128, ctx->stream>>>(v, v.n_pos);
        launches++;
        launch(uvc_k1b_noindel, ctx->stream, v, v.n_pos, launches);
    }
    UVC_CUDA_CHECK(ctx, cudaEventRecord(bs.ev[e++], ctx->stream));
    pb = pos_block(n_k2);
    if (n_k2 > 0) {
        static_assert(sizeof(K2StageP) % 16 == 0 && sizeof(PileRec) == 64, "per-warp staging slots keep 16-byte alignment");
        const size_t smem = 4 * sizeof(K2StageP);
        UVC_CUDA_CHECK(ctx, cudaFuncSetAttribute(uvc_k2_bias_pileup, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        uvc_k2_bias_pileup<<<(unsigned)((n_k2 + pb - 1) / pb), pb, smem * pb / 128, ctx->stream>>>(v, n_k2);
        launches++;
    }
    UVC_CUDA_CHECK(ctx, cudaEventRecord(bs.ev[e++], ctx->stream));
    UVC_STAGE(uvc_k2e_indel_events, v.n_ev)
    UVC_STAGE(uvc_kf_fragment_columns, v.n_frags * 32)
    UVC_STAGE(uvc_k3a_fragment_stats, v.n_frags)
    pb = pos_block(n_k34);
    if (n_k34 > 0) {
        static_assert(sizeof(ColStage<ReadFrag>) % 16 == 0, "per-warp staging slots keep 16-byte alignment");
        const size_t smem = (pb / 32) * sizeof(ColStage<ReadFrag>) + 2 * UVC_NUM_BUCKETS * pb * sizeof(int32_t);
        const size_t smem_max = 4 * sizeof(ColStage<ReadFrag>) + 2 * UVC_NUM_BUCKETS * 128 * sizeof(int32_t);   // (the attribute is shared by all contexts: always the largest shape)
        UVC_CUDA_CHECK(ctx, cudaFuncSetAttribute(uvc_k3b_fragment_consensus, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
        uvc_k3b_fragment_consensus<<<(unsigned)((n_k34 + pb - 1) / pb), pb, smem, ctx->stream>>>(v, n_k34);
        launches++;
    }
    UVC_CUDA_CHECK(ctx, cudaEventRecord(bs.ev[e++], ctx->stream));
    UVC_STAGE(uvc_km_family_columns, v.n_mcol)
    UVC_STAGE(uvc_k4a_family_ends, 2 * v.n_fams)
    if (n_k34 > 0) {
        static_assert(sizeof(K4StageT<false>) % 16 == 0 && sizeof(K4StageT<true>) % 16 == 0, "per-warp staging slots keep 16-byte alignment");
        // the wide shape when a good part of the column entries belongs to multi-fragment (UMI) families
        const bool wide = (v.n_mcol * 16 >= v.n_fcol);
        const size_t smem = 4 * (wide ? sizeof(K4StageT<true>) : sizeof(K4StageT<false>));
        UVC_CUDA_CHECK(ctx, cudaFuncSetAttribute(uvc_k4_family_consensus, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(4 * sizeof(K4StageT<false>))));
        UVC_CUDA_CHECK(ctx, cudaFuncSetAttribute(uvc_k4_family_consensus_umi, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(4 * sizeof(K4StageT<true>))));
        if (wide) { uvc_k4_family_consensus_umi<<<(unsigned)((n_k34 + pb - 1) / pb), pb, smem * pb / 128, ctx->stream>>>(v, n_k34); }
        else { uvc_k4_family_consensus<<<(unsigned)((n_k34 + pb - 1) / pb), pb, smem * pb / 128, ctx->stream>>>(v, n_k34); }
        launches++;
    }
    UVC_CUDA_CHECK(ctx, cudaEventRecord(bs.ev[e++], ctx->stream));
    UVC_STAGE(uvc_k4c_family_haplotypes, 2 * v.n_fams)
    #undef UVC_STAGE
    {
        // K6 needs nothing from the host: it and the downloads of its outputs and of the sparse stream's cursor ride behind the batch's kernels
        ScoreView sv6;
        memset(&sv6, 0, sizeof(sv6));
        sv6.gvcf = bs.d_gvcf; sv6.gextra = bs.d_gextra;
        if (v.n_pos > 0) { uvc_k6_gvcf_inputs<<<(unsigned)((v.n_pos + 127) / 128), 128, 0, ctx->stream>>>(v, sv6, v.n_pos); launches++; }
        UVC_CUDA_CHECK(ctx, cudaEventRecord(bs.ev[e++], ctx->stream));
        UVC_CUDA_CHECK(ctx, cudaMemcpyAsync(bs.gvcf.data(), bs.d_gvcf, bs.gvcf.size() * sizeof(GvcfPos), cudaMemcpyDeviceToHost, ctx->stream));
        UVC_CUDA_CHECK(ctx, cudaMemcpyAsync(bs.gextra.data(), bs.d_gextra, bs.gextra.size() * sizeof(GvcfExtra), cudaMemcpyDeviceToHost, ctx->stream));
        UVC_CUDA_CHECK(ctx, cudaMemcpyAsync(bs.rec_cursor_host, v.rec_cursor, 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        UVC_CUDA_CHECK(ctx, cudaEventCreateWithFlags(&bs.ev_done, cudaEventDisableTiming));
        UVC_CUDA_CHECK(ctx, cudaEventRecord(bs.ev_done, ctx->stream));      // (the batch is complete when its downloads are)
        bs.stats.d2h_bytes += (int64_t)(bs.gvcf.size() * sizeof(GvcfPos) + bs.gextra.size() * sizeof(GvcfExtra) + 16);
    }
    UVC_CUDA_CHECK(ctx, cudaGetLastError());
    bs.stats.gpu_launches += launches;
    return 0;
}

static int backend_wait(uvcgpu_ctx *ctx, BatchState & bs) {
    // only this batch's last kernel is waited for: a later batch may already be queued on the stream
    if (bs.have_events) { UVC_CUDA_CHECK(ctx, uvc_event_wait(bs.ev_done)); cudaEventDestroy(bs.ev_done); }
    else { const int rc_ = backend_wait_stream(ctx, ctx->stream); if (rc_ != 0) { return rc_; } }
    if (bs.have_events) {
        float ms = 0;
        double total = 0;
        for (int i = 0; i < UVC_N_PILEUP_STAGES; i++) {
            UVC_CUDA_CHECK(ctx, cudaEventElapsedTime(&ms, bs.ev[i], bs.ev[i + 1]));
            bs.stats.kernel_ms_by_stage[i] = ms;
            total += ms;
        }
        bs.stats.kernel_ms = total;
        for (int i = 0; i < UVC_N_PILEUP_STAGES + 1; i++) { cudaEventDestroy(bs.ev[i]); }
        bs.have_events = false;
    }
    if (bs.have_prep_events) {
        // the staging kernels (P0/P1): includes the two short host round trips inside them (sizes of the batch)
        float ms = 0;
        UVC_CUDA_CHECK(ctx, cudaEventElapsedTime(&ms, bs.ev_prep[0], bs.ev_prep[1]));
        bs.stats.kernel_ms_by_stage[13] = ms;
        bs.stats.kernel_ms += ms;
        cudaEventDestroy(bs.ev_prep[0]); cudaEventDestroy(bs.ev_prep[1]);
        bs.have_prep_events = false;
    }
    return 0;
}

static int backend_score(uvcgpu_ctx *ctx, BatchState & bs, const ScoreView & sv) {
    const BatchView & v = bs.view;
    cudaEvent_t e[2];
    for (int i = 0; i < 2; i++) { UVC_CUDA_CHECK(ctx, cudaEventCreate(&e[i])); }
    UVC_CUDA_CHECK(ctx, cudaEventRecord(e[0], t_active));
    if (v.n_pos > 0) { uvc_k5a_flag_candidates<<<(unsigned)((v.n_pos + 127) / 128), 128, 0, t_active>>>(v, sv, v.n_pos); }
    if (v.n_pos > 0) {
        uvc_k5b_list_candidates<<<(unsigned)((v.n_pos + 127) / 128), 128, 0, t_active>>>(v, sv, v.n_pos);
        uvc_k5g_group_init<<<(unsigned)((sv.group_cap + 127) / 128), 128, 0, t_active>>>(v, sv, sv.group_cap);
        uvc_k5c_candidate_depths<<<(unsigned)((sv.cand_cap + 127) / 128), 128, 0, t_active>>>(v, sv, sv.cand_cap);
        uvc_k5e_candidate_quals<<<(unsigned)((sv.cand_cap + 127) / 128), 128, 0, t_active>>>(v, sv, sv.cand_cap);
        uvc_k5f_group_records<<<(unsigned)((sv.group_cap + 127) / 128), 128, 0, t_active>>>(v, sv, sv.group_cap);
    }
    UVC_CUDA_CHECK(ctx, cudaEventRecord(e[1], t_active));
    UVC_CUDA_CHECK(ctx, cudaGetLastError());
    // the results the host always needs ride behind the kernels: the cursors and the first records
    { int rc_ = backend_download_async(ctx, bs.score_cursor_host, sv.out_cursor, 8 * sizeof(int32_t));
      if (0 == rc_) { rc_ = backend_download_async(ctx, bs.recs.data(), sv.out, bs.recs.size() * sizeof(VarRec)); }
      if (0 == rc_) { rc_ = backend_sync(ctx); }
      if (rc_ != 0) { return rc_; } }
    float ms = 0;
    UVC_CUDA_CHECK(ctx, cudaEventElapsedTime(&ms, e[0], e[1])); bs.stats.kernel_ms_by_stage[12] = ms; bs.stats.kernel_ms += ms;
    for (int i = 0; i < 2; i++) { cudaEventDestroy(e[i]); }
    if (v.n_pos > 0) { bs.stats.gpu_launches += 6; }
    return 0;
}

#else // ------------------------------------------------------------------------------------------- emulation (tests only)

static int backend_alloc(uvcgpu_ctx *, BatchState & bs, void **out, size_t bytes, bool) {
    bytes += 64;     // same slack as the CUDA build
    *out = calloc(1, bytes);
    if (NULL == *out) { return UVCGPU_ENOMEM; }
    bs.allocs.push_back(*out);
    return 0;
}
static int backend_upload(uvcgpu_ctx *, BatchState & bs, void *dst, const void *src, size_t bytes) {
    if (bytes) { memcpy(dst, src, bytes); bs.stats.h2d_bytes += (int64_t)bytes; }
    return 0;
}
static int backend_download(uvcgpu_ctx *, void *dst, const void *src, size_t bytes) { if (bytes) { memcpy(dst, src, bytes); } return 0; }
static int backend_download_async(uvcgpu_ctx *ctx, void *dst, const void *src, size_t bytes) { return backend_download(ctx, dst, src, bytes); }
static int backend_sync(uvcgpu_ctx *) { return 0; }
static int backend_zero(uvcgpu_ctx *, void *dst, size_t bytes) { memset(dst, 0, bytes); return 0; }
static void backend_free_temps(uvcgpu_ctx *, BatchState & bs, bool = false) { for (void *p : bs.temp_allocs) { free(p); } bs.temp_allocs.clear(); }
static void backend_free(uvcgpu_ctx *ctx, BatchState & bs) { backend_free_temps(ctx, bs); for (void *p : bs.allocs) { free(p); } bs.allocs.clear(); }
static int backend_alloc_temp(uvcgpu_ctx *, BatchState & bs, void **out, size_t bytes, int fill = -1) {
    bytes += 64;
    *out = malloc(bytes);
    if (NULL == *out) { return UVCGPU_ENOMEM; }
    memset(*out, (fill >= 0 ? fill : 0xA5), bytes);      // (unfilled scratch is poisoned: every element must be written before it is read)
    bs.temp_allocs.push_back(*out);
    return 0;
}

#include "prep_device.inc"

static int backend_run(uvcgpu_ctx *, BatchState & bs) {
    const BatchView & v = bs.view;
    if (getenv("UVC_EMU_PREP_ONLY")) { return 0; }   // host-staging profiling runs (tools/prep_profile.py)
    for (int64_t i = 0; i < v.n_reads; i++) { uvc::k0_read(v, i); }
    uvc::Win w;
    for (int64_t i = 0; i < v.n_pos; i++) { uvc::position_window(v, i, w); uvc::k1_position(v, i, w); }
    for (int64_t i = 0; i < v.n_pos; i++) { uvc::k1b_noindel(v, i); }
    const int64_t n_k2 = (v.list_tile[0] ? v.n_list[0] : v.n_pos), n_k34 = (v.list_tile[1] ? v.n_list[1] : v.n_pos);
    for (int64_t j = 0; j < n_k2; j++) { const int64_t i = uvc::list_position(v, 0, j); if (i < 0) { continue; } uvc::position_window(v, i, w); uvc::k2m_position(v, i, w); }
    for (int64_t i = 0; i < v.n_ev; i++) { uvc::k2e_event(v, i); }
    for (int64_t i = 0; i < v.n_fcol; i++) { uvc::kf_fold_bits(v, i, uvc::kf_fragment_column(v, i)); }
    for (int64_t i = 0; i < v.n_frags; i++) { uvc::k3a_fragment(v, i); }
    for (int64_t j = 0; j < n_k34; j++) { const int64_t i = uvc::list_position(v, 1, j); if (i < 0) { continue; } uvc::position_window(v, i, w); uvc::k3b_position(v, i, w); }
    for (int64_t i = 0; i < v.n_mcol; i++) { uvc::km_family_column(v, i); }
    for (int64_t i = 0; i < 2 * v.n_fams; i++) { uvc::k4a_family_strand(v, i); }
    for (int64_t j = 0; j < n_k34; j++) { const int64_t i = uvc::list_position(v, 1, j); if (i < 0) { continue; } uvc::position_window(v, i, w); uvc::k4_position(v, i, w); }
    for (int64_t i = 0; i < 2 * v.n_fams; i++) { uvc::k4c_family_strand(v, i); }
    ScoreView sv6;
    memset(&sv6, 0, sizeof(sv6));
    sv6.gvcf = bs.d_gvcf; sv6.gextra = bs.d_gextra;
    for (int64_t i = 0; i < v.n_pos; i++) { uvc::k6_gvcf_position(v, sv6, i); }
    memcpy(bs.gvcf.data(), bs.d_gvcf, bs.gvcf.size() * sizeof(GvcfPos)); memcpy(bs.gextra.data(), bs.d_gextra, bs.gextra.size() * sizeof(GvcfExtra));
    memcpy(bs.rec_cursor_host, v.rec_cursor, 4 * sizeof(int32_t));
    return 0;
}
static int backend_wait(uvcgpu_ctx *, BatchState &) { return 0; }
static int backend_score(uvcgpu_ctx *, BatchState & bs, const ScoreView & sv) {
    const BatchView & v = bs.view;
    for (int64_t i = 0; i < v.n_pos; i++) { uvc::k5a_flag_position(v, sv, i); }
    for (int64_t i = 0; i < (int64_t)*sv.cand_cursor; i++) { uvc::k5b_list_position(v, sv, (int64_t)sv.cand_list[i]); }
    if (sv.out_cursor[2] <= sv.group_cap && sv.out_cursor[3] <= sv.cand_cap) {
        for (int64_t i = 0; i < (int64_t)sv.out_cursor[2]; i++) { uvc::k5g_group(v, sv, i); }
        for (int64_t i = 0; i < (int64_t)sv.out_cursor[3]; i++) { uvc::k5c_candidate(v, sv, i); }
        for (int64_t i = 0; i < (int64_t)sv.out_cursor[3]; i++) { uvc::k5e_candidate(v, sv, i); }
        for (int64_t i = 0; i < (int64_t)sv.out_cursor[2]; i++) { uvc::k5f_group(v, sv, i); }
    }
    memcpy(bs.score_cursor_host, sv.out_cursor, 8 * sizeof(int32_t));
    memcpy(bs.recs.data(), sv.out, bs.recs.size() * sizeof(VarRec));
    return 0;
}

#endif

// Entry points may be called from any host thread: the context's device is made current first (CUDA's current device is per thread).
static inline void enter_ctx(const uvcgpu_ctx *ctx) {
#if UVC_CUDA
    if (ctx) { cudaSetDevice(ctx->device); t_active = ctx->stream; }
#else
    (void)ctx;
#endif
}

// ------------------------------------------------------------------------------------------------ C ABI
extern "C" {

static int wait_submitted(uvcgpu_ctx *ctx, BatchState & bs);
static void wait_worker_idle(uvcgpu_ctx *ctx);
#if UVC_CUDA
static void submit_worker(uvcgpu_ctx *ctx);
#endif


void uvcgpu_params_default(uvcgpu_params *p) {
    memset(p, 0, sizeof(*p));
    p->abi_version = UVCGPU_ABI_VERSION;
    p->inferred_sequencing_platform = 1;
    p->central_readlen = 150;
    p->inferred_maxMQ = 60;
    p->outvar_flag = 62;
    p->kept_aln_max_isize = INT32_MAX;
    p->min_altdp_thres = 2;
    p->dedup_center_mult = 5;
    p->dedup_amplicon_end2end_ratio = 1.5;
    p->dedup_amplicon_border_to_insert_cov_weak_avgDP_ratio = 5;
    p->dedup_amplicon_border_to_insert_cov_strong_avgDP_ratio = 20;
    p->dedup_amplicon_border_to_insert_cov_weak_totDP_ratio = 0.05;
    p->dedup_amplicon_border_to_insert_cov_strong_totDP_ratio = 0.20;
    p->dedup_amplicon_border_weak_minDP = 100;
    p->dedup_amplicon_border_strong_minDP = 400;
    p->assay_sequencing_BQ_max = 37;
    p->primerlen2 = 23;
    p->bias_thres_highBQ = 20; p->bias_thres_highBAQ = 20; p->bias_thres_aLPxT_add = 5; p->bias_thres_aLPxT_perc = 160;
    p->bias_thres_aLRP1t_minus = 10; p->bias_thres_aLRP2t_minus = 5; p->bias_thres_aLRB1t_minus = 50; p->bias_thres_aLRB2t_minus = 25;
    p->bias_thres_aLRP1t_avgmul_perc = 100; p->bias_thres_aLRP2t_avgmul_perc = 100; p->bias_thres_aLRB1t_avgmul_perc = 100; p->bias_thres_aLRB2t_avgmul_perc = 100;
    p->bias_thres_aLRP1Nt_avgmul_perc = 80; p->bias_thres_aLRB1Nt_avgmul_perc = 80;
    p->bias_thres_aLRI1T_perc = 200; p->bias_thres_aLRI2T_perc = 150; p->bias_thres_aLRI1t_perc = 50; p->bias_thres_aLRI2t_perc = 67;
    p->bias_thres_aLRI1NT_perc = 250; p->bias_thres_aLRI1Nt_perc = 40; p->bias_thres_aLRI1T_add = 180; p->bias_thres_aLRI2T_add = 150;
    p->bias_thres_PFBQ1 = 25; p->bias_thres_PFBQ2 = 30;
    p->bias_thres_interfering_indel = 5; p->bias_thres_interfering_indel_BQ = 21; p->bias_thres_BAQ1 = 23; p->bias_thres_BAQ2 = 33;
    p->bias_thres_strict_c2LRP0 = 5;
    p->fam_thres_highBQ_snv = 25; p->fam_thres_highBQ_indel = 13; p->fam_thres_dup1add = 2; p->fam_thres_dup1perc = 80;
    p->fam_thres_dup2add = 3; p->fam_thres_dup2perc = 70; p->fam_thres_qseqlen = 75;
    p->fam_thres_emperr_all_flat_snv = 4; p->fam_thres_emperr_con_perc_snv = 67; p->fam_thres_emperr_all_flat_indel = 4; p->fam_thres_emperr_con_perc_indel = 67;
    p->fam_phred_indel_inc_before_barcode_labeling = 14;
    p->fam_phred_sscs_transition_CG_TA = 40; p->fam_phred_sscs_transition_AT_GC = 44; p->fam_phred_sscs_transversion_CG_AT = 48; p->fam_phred_sscs_transversion_other = 48;
    p->fam_phred_sscs_indel_open = 58; p->fam_phred_sscs_indel_ext = 0;
    p->syserr_mut_region_n_bases = 11;
    p->indel_BQ_max = 42; p->indel_str_repeatsize_max = 6; p->indel_vntr_repeatsize_max = 35;
    p->indel_polymerase_size = 8.0; p->indel_polymerase_slip_rate = 8.0; p->indel_del_to_ins_err_ratio = 5.0;
    p->indel_adj_tracklen_dist = 6; p->indel_adj_indellen_perc = 160; p->indel_nonSTR_phred_per_base = 5; p->indel_str_phred_per_region = 10; p->indel_filter_edge_dist = 5;
    p->powlaw_exponent = 3.0;
    p->microadjust_xm = 7; p->microadjust_cliplen = 5; p->microadjust_delFAQmax = 49; p->microadjust_nobias_pos_indel_maxlen = 16;
    p->microadjust_near_clip_dist = 2; p->microadjust_alignment_clip_min_len = 12; p->microadjust_padded_deletion_flag = 0x2;
    p->microadjust_median_readlen_thres = 125; p->microadjust_BAQ_per_base_x1024 = 1024;
    p->tumor_vcf_fname_nonempty = 1;
    // scoring defaults (CmdLineArgs.hpp:39, 72-82, 106-108, 110, 196-232, 242-247, 263-272, 277-303, 309-326, 341-347, 361-417) with the
    // Illumina inference of CmdLineArgs.cpp:127-134 applied to syserr_minABQ_*
    p->vqual = 15; p->vfa1 = 0.002; p->vfa2 = 0.0002; p->vdp1 = 1000; p->vad1 = 4; p->vdp2 = 10000; p->vad2 = 8; p->min_r_ad = 0; p->min_a_ad = 0;
    p->syserr_minABQ_pcr_snv = 200; p->syserr_minABQ_pcr_indel = 100; p->syserr_minABQ_cap_snv = 200; p->syserr_minABQ_cap_indel = 100;
    p->syserr_BQ_prior = 30; p->syserr_BQ_sbratio_q_add = 5; p->syserr_BQ_sbratio_q_max = 40; p->syserr_BQ_xmratio_q_add = 5; p->syserr_BQ_xmratio_q_max = 40;
    p->syserr_BQ_bmratio_q_add = 5; p->syserr_BQ_bmratio_q_max = 40; p->syserr_BQ_strand_favor_mul = 3; p->syserr_MQ_min = 0; p->syserr_MQ_max = 60;
    p->syserr_MQ_NMR_expfrac = 0.03; p->syserr_MQ_NMR_altfrac_coef = 2.0; p->syserr_MQ_NMR_nonaltfrac_coef = 2.0; p->syserr_MQ_NMR_pl_exponent = 3.0; p->syserr_MQ_nonref_base = 40;
    p->powlaw_anyvar_base = (double)(60 + 25 + 5); p->powlaw_amplicon_allele_fraction_coef = (5.0 / 8.0);
    p->penal4lowdep = 37; p->nobias_flag = 0x2; p->nobias_pos_indel_lenfrac_thres = 2.0; p->nobias_pos_indel_str_track_len = 16;
    p->bias_prior_DPadd_perc = 50;
    p->bias_priorfreq_pos = 40; p->bias_priorfreq_indel_in_read_div = 20; p->bias_priorfreq_indel_in_var_div2 = 15; p->bias_priorfreq_indel_in_str_div2 = 10; p->bias_priorfreq_var_in_str_div2 = 5;
    p->bias_prior_var_DP_mul = 1.25 + (double)FLT_EPSILON;
    p->bias_priorfreq_ipos_snv = 45; p->bias_priorfreq_ipos_indel = 45; p->bias_priorfreq_strand_snv_base = 10; p->bias_priorfreq_strand_indel = 45;
    p->bias_FA_pseudocount_indel_in_read = 0.5 / 10.0; p->bias_priorfreq_orientation_snv_base = 45; p->bias_priorfreq_orientation_indel_base = 45;
    p->bias_FA_powerlaw_noUMI_phred_inc_snv = 5; p->bias_FA_powerlaw_noUMI_phred_inc_indel = 7; p->bias_FA_powerlaw_withUMI_phred_inc_snv = 8; p->bias_FA_powerlaw_withUMI_phred_inc_indel = 7;
    p->bias_reduction_by_high_sequencingDP_min_n_totDepth = 800; p->bias_reduction_by_high_sequencingDP_min_n_altDepth = 3;
    p->bias_thres_FTS_FA = 0.6; p->bias_orientation_min_effective_allelefrac = 0.004; p->bias_is_orientation_artifact_mixed_with_sequencing_error = 0;
    p->fam_min_n_copies = 800; p->fam_min_n_copies_DPxAD = 20 * 1000; p->fam_min_overseq_perc = 200; p->fam_bias_overseq_perc = 150; p->fam_tier3DP_bias_overseq_perc = 350;
    p->fam_indel_nonUMI_phred_dec_per_fold_overseq = 9;
    p->fam_phred_dscs_all = 58; p->fam_phred_dscs_max = 68; p->fam_phred_dscs_inc_max = (68 - 48); p->fam_phred_pow_sscs_transversion_AT_TA_origin = 44 - (41 - 6) + 4;
    p->fam_phred_pow_sscs_snv_origin = 44 - (41 - 6); p->fam_phred_pow_sscs_indel_origin = 58 - 9 * 3; p->fam_phred_pow_dscs_all_origin = 0;
    p->germ_hetero_FA = 0.47;
    p->germ_phred_hetero_snp = 31; p->germ_phred_hetero_indel = 40; p->germ_phred_homalt_snp = 33; p->germ_phred_homalt_indel = 42; p->germ_phred_het3al_snp = 59; p->germ_phred_het3al_indel = 49;
    p->tn_q_inc_max = 9; p->tn_q_inc_max_sscs_CG_AT = 0; p->tn_q_inc_max_sscs_other = 5; p->tn_syserr_norm_devqual = 15.0;
    p->indel_multiallele_samepos_penal = 11.0; p->indel_multiallele_diffpos_penal = 8.0; p->indel_multiallele_soma_penal_thres = 11.0;
    p->indel_tetraallele_germline_penal_value = 8.0 * 2; p->indel_tetraallele_germline_penal_thres = 22.0; p->indel_ins_penal_pseudocount = 16;
    p->contam_any_mul_frac = 0.02; p->contam_t2n_mul_frac = 0.05;
    p->microadjust_bias_pos_indel_fold = 2; p->microadjust_bias_pos_indel_misma_to_indel_ratio = 4 * (1.0 - DBL_EPSILON); p->microadjust_nobias_pos_indel_misma_to_indel_ratio = 4 * (1.0 - DBL_EPSILON);
    p->microadjust_nobias_pos_indel_bMQ = 50; p->microadjust_nobias_pos_indel_perc = 50;
    p->microadjust_nobias_strand_all_fold = 5; p->microadjust_refbias_indel_max = 2.0; p->microadjust_counterbias_pos_odds_ratio = 3.5; p->microadjust_counterbias_pos_fold_ratio = 5.0;
    p->microadjust_fam_binom_qual_halving_thres = 70; p->microadjust_ref_MQ_dec_max = 15;
    p->microadjust_syserr_MQ_NMR_tn_syserr_no_penal_qual_min = 30; p->microadjust_syserr_MQ_NMR_tn_syserr_no_penal_qual_max = 30 + 12;
    p->microadjust_longfrag_sidelength_min = 300; p->microadjust_longfrag_sidelength_max = 600; p->microadjust_longfrag_sidelength_zeroMQpenalty = 300;
    p->microadjust_alignment_clip_min_count = 2; p->microadjust_alignment_tracklen_min = 25; p->microadjust_alignment_clip_min_frac = 0.05;
    p->microadjust_germline_mix_with_del_snv_penalty = 9; p->microadjust_strand_orientation_absence_DP_fold = 5; p->microadjust_orientation_absence_snv_penalty = 4;
    p->microadjust_strand_absence_snv_penalty = 4; p->microadjust_dedup_absence_indel_penalty = 1;
    p->lib_wgs_min_avg_fraglen = 300; p->lib_nonwgs_clip_penal_min_indelsize = 8; p->lib_nonwgs_normal_max_rescued_MQ = 30; p->lib_wgs_normal_max_rescued_MQ = 0;
    p->lib_nonwgs_ad_pseudocount = 0.1; p->lib_nonwgs_normal_full_self_rescue_fa = 0.1; p->lib_nonwgs_normal_min_self_rescue_fa_ratio = 0.2; p->lib_nonwgs_normal_add_mul_ad = 1.0;
    p->phasing_haplotype_max_count = 8; p->phasing_haplotype_min_ad = 1; p->phasing_haplotype_max_detail_cnt = 3;
}

size_t uvcgpu_sizeof_params(void) { return sizeof(uvcgpu_params); }

int uvcgpu_device_count(void) {
#if UVC_CUDA
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { return 0; }
    return n;
#else
    return 0;
#endif
}

int uvcgpu_staging_backlog(void) {
#if UVC_CUDA
    StageState & st = stage_state();
    std::lock_guard<std::mutex> lk(st.mu);
    return (int)st.wanted.size() + st.in_progress;
#else
    return 0;
#endif
}

int64_t uvcgpu_staging_pinned_bytes(void) {
#if UVC_CUDA
    StageState & st = stage_state();
    std::lock_guard<std::mutex> lk(st.mu);
    return (int64_t)st.total_pinned;
#else
    return 0;
#endif
}

// page-locks (or releases) every array of a caller's SoA buffer
static int host_register_reads(const uvcgpu_reads_soa *r, bool on) {
    if (NULL == r) { return UVCGPU_EINVAL; }
#if UVC_CUDA
    const int64_t n = r->n_reads;
    if (n <= 0) { return UVCGPU_OK; }
    const void *ptrs[17] = {r->pos, r->mpos, r->isize, r->mtid, r->l_qseq, r->n_cigar, r->nm, r->flag, r->mapq, r->seq_off, r->qual_off, r->cigar_off, r->qname_off,
                            r->seq, r->qual, r->cigar, r->qname};
    const size_t bytes[17] = {(size_t)n * 4, (size_t)n * 4, (size_t)n * 4, (size_t)n * 4, (size_t)n * 4, (size_t)n * 4, (size_t)n * 4, (size_t)n * 2, (size_t)n,
                              (size_t)(n + 1) * 8, (size_t)(n + 1) * 8, (size_t)(n + 1) * 8, (size_t)(n + 1) * 8,
                              (size_t)r->seq_off[n], (size_t)r->qual_off[n], (size_t)r->cigar_off[n] * 4, (size_t)r->qname_off[n]};
    int rc = UVCGPU_OK;
    for (int k = 0; k < 17; k++) {
        if (NULL == ptrs[k] || 0 == bytes[k]) { continue; }
        const cudaError_t e = (on ? cudaHostRegister((void*)ptrs[k], bytes[k], cudaHostRegisterPortable) : cudaHostUnregister((void*)ptrs[k]));
        if (e != cudaSuccess) { cudaGetLastError(); if (e != cudaErrorHostMemoryAlreadyRegistered && e != cudaErrorHostMemoryNotRegistered) { rc = UVCGPU_ECUDA; } }
    }
    return rc;
#else
    (void)on;
    return UVCGPU_OK;
#endif
}
int uvcgpu_host_register_reads(const uvcgpu_reads_soa *reads) { return host_register_reads(reads, true); }
int uvcgpu_host_unregister_reads(const uvcgpu_reads_soa *reads) { return host_register_reads(reads, false); }

int uvcgpu_host_call_stats(double *out, int32_t cap) {
    // per kind (device allocation, free, memset, download enqueue, kernel launches, waits for events, upload enqueue): total ms, calls, longest call in ms
    const int n = 3 * UVC_T_N;
    if (out) {
        for (int k = 0; k < UVC_T_N && 3 * k + 2 < cap; k++) {
            out[3 * k] = (double)g_call_ns[k].load() * 1e-6; out[3 * k + 1] = (double)g_call_n[k].load(); out[3 * k + 2] = (double)g_call_max_ns[k].load() * 1e-6;
        }
    }
    return n;
}

int uvcgpu_device_warmup(int device) {
#if UVC_CUDA
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) { return UVCGPU_ENODEVICE; }
    if (cudaSetDevice(device) != cudaSuccess || cudaFree(0) != cudaSuccess) { return UVCGPU_ECUDA; }
    return UVCGPU_OK;
#else
    (void)device;
    return UVCGPU_OK;
#endif
}

int uvcgpu_create(uvcgpu_ctx **out, int device, const uvcgpu_params *params) {
    if (NULL == out || NULL == params || params->abi_version != UVCGPU_ABI_VERSION) { return UVCGPU_EINVAL; }
    *out = NULL;
#if UVC_CUDA
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) { return UVCGPU_ENODEVICE; }
    if (cudaSetDevice(device) != cudaSuccess) { return UVCGPU_ENODEVICE; }
#endif
    uvcgpu_ctx *ctx = new uvcgpu_ctx();
    ctx->device = device;
    ctx->par = *params;
    if (params->indel_str_repeatsize_max > UVC_SLIP_MAXUNIT) { delete ctx; return UVCGPU_EUNSUPPORTED; }
    if (params->outvar_flag & 0x1u) { delete ctx; return UVCGPU_EUNSUPPORTED; }     // GERMLINE text records (main.hpp:5736-5773) are not emitted by this build
    // indel_phred (main.hpp:794-801) tabulated with the host libm so that device and reference agree on every floor()
    ctx->slip_tab.assign((size_t)2 * UVC_SLIP_MAXUNIT * UVC_SLIP_NMAX, 0);
    for (int variant = 0; variant < 2; variant++) {
        const double ampfact = (variant ? params->indel_polymerase_slip_rate * params->indel_del_to_ins_err_ratio : params->indel_polymerase_slip_rate);
        for (int unit = 1; unit <= UVC_SLIP_MAXUNIT; unit++) {
            for (int nu = 0; nu < UVC_SLIP_NMAX; nu++) {
                const int region = unit * nu;
                const double num_slips = (region > 64 ? (double)(region - 8) : log1p(exp((double)region - (double)8))) * ampfact / ((double)(unit * unit));
                ctx->slip_tab[((size_t)variant * UVC_SLIP_MAXUNIT + (unit - 1)) * UVC_SLIP_NMAX + nu] = (int32_t)floor(-10 * log((1.0 - DBL_EPSILON) / (num_slips + 1.0)) / log(10));
            }
        }
    }
    std::vector<double> p2p(128);
    for (int q = 0; q < 128; q++) { p2p[q] = pow(10, -((float)q) / 10); }   // phred2prob (main_conversion.hpp:885-888): float exponent, double pow
    std::vector<int32_t> pf(256);
    for (int k = 0; k < 2; k++) {
        const int32_t t = (k ? params->bias_thres_PFBQ2 : params->bias_thres_PFBQ1);
        for (int bq = 0; bq < 128; bq++) { pf[k * 128 + bq] = ((bq < t) ? (100 * (bq * bq) / (t * t)) : 100); }
    }
#if UVC_CUDA
    {
        bool ok = (cudaMalloc((void**)&ctx->c_phred2prob, p2p.size() * sizeof(double)) == cudaSuccess) && (cudaMalloc((void**)&ctx->c_pf_tab, pf.size() * sizeof(int32_t)) == cudaSuccess)
            && (cudaMalloc((void**)&ctx->c_slip_tab, ctx->slip_tab.size() * sizeof(int32_t)) == cudaSuccess);
        ok = ok && (cudaMemcpy(ctx->c_phred2prob, p2p.data(), p2p.size() * sizeof(double), cudaMemcpyHostToDevice) == cudaSuccess)
            && (cudaMemcpy(ctx->c_pf_tab, pf.data(), pf.size() * sizeof(int32_t), cudaMemcpyHostToDevice) == cudaSuccess)
            && (cudaMemcpy(ctx->c_slip_tab, ctx->slip_tab.data(), ctx->slip_tab.size() * sizeof(int32_t), cudaMemcpyHostToDevice) == cudaSuccess);
        if (!ok) { cudaGetLastError(); delete ctx; return UVCGPU_ECUDA; }
    }
#else
    ctx->c_phred2prob = (double*)malloc(p2p.size() * sizeof(double)); memcpy(ctx->c_phred2prob, p2p.data(), p2p.size() * sizeof(double));
    ctx->c_pf_tab = (int32_t*)malloc(pf.size() * sizeof(int32_t)); memcpy(ctx->c_pf_tab, pf.data(), pf.size() * sizeof(int32_t));
    ctx->c_slip_tab = (int32_t*)malloc(ctx->slip_tab.size() * sizeof(int32_t)); memcpy(ctx->c_slip_tab, ctx->slip_tab.data(), ctx->slip_tab.size() * sizeof(int32_t));
#endif
#if UVC_CUDA
    // The host waits for the short staging and scoring kernels (sizes of the batch, kept records), never for the long pileup kernels: the
    // streams of the short kernels get the higher priority, so that their blocks are scheduled ahead of the pending blocks of pileup kernels
    // of other contexts on the same GPU instead of behind them.
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, prio_lo) != cudaSuccess) { delete ctx; return UVCGPU_ECUDA; }
    if (cudaStreamCreateWithPriority(&ctx->post_stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess) { cudaStreamDestroy(ctx->stream); delete ctx; return UVCGPU_ECUDA; }
    if (cudaStreamCreateWithPriority(&ctx->prep_stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess) { cudaStreamDestroy(ctx->stream); cudaStreamDestroy(ctx->post_stream); delete ctx; return UVCGPU_ECUDA; }
    t_active = ctx->stream;
    {   // keep freed device blocks in the stream-ordered pool instead of returning them to the driver after every batch
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) { uint64_t keep = UINT64_MAX; cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep); }
    }
#endif
#if UVC_CUDA
    { SlabCache & sc = slab_cache(device); std::lock_guard<std::mutex> lk(sc.mu); sc.live_contexts++; }
    if (cudaHostAlloc((void**)&ctx->cursor_slab, (size_t)UVC_CURSOR_SLOTS * UVC_CURSOR_WORDS * sizeof(int32_t), cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); delete ctx; return UVCGPU_ECUDA; }
    ctx->worker = std::thread(submit_worker, ctx);
#else
    ctx->cursor_slab = (int32_t*)calloc((size_t)UVC_CURSOR_SLOTS * UVC_CURSOR_WORDS, sizeof(int32_t));
#endif
    *out = ctx;
    return UVCGPU_OK;
}

void uvcgpu_destroy(uvcgpu_ctx *ctx) {
    if (NULL == ctx) { return; }
    enter_ctx(ctx);
    wait_worker_idle(ctx);
#if UVC_CUDA
    { std::lock_guard<std::mutex> lk(ctx->wmu); ctx->wstop = true; }
    ctx->wcv.notify_all();
    if (ctx->worker.joinable()) { ctx->worker.join(); }
#endif
#if UVC_CUDA
    cudaStreamSynchronize(ctx->prep_stream); cudaStreamSynchronize(ctx->stream); cudaStreamSynchronize(ctx->post_stream);   // (slabs go back to the cache only after their last use)
#endif
    for (auto & kv : ctx->batches) { backend_free(ctx, *kv.second); }
#if UVC_CUDA
    {
        SlabCache & sc = slab_cache(ctx->device);
        std::lock_guard<std::mutex> lk(sc.mu);
        if (--sc.live_contexts <= 0) { sc.live_contexts = 0; slab_purge(sc); }     // the last context of the device gives the memory back
    }
    for (auto & kv : ctx->d_contigs) { cudaFree(kv.second); }
    cudaFree(ctx->c_phred2prob); cudaFree(ctx->c_pf_tab); cudaFree(ctx->c_slip_tab);
    if (ctx->cursor_slab) { cudaFreeHost(ctx->cursor_slab); }
    for (auto & e : ctx->wait_ev) { if (e) { cudaEventDestroy(e); } }
    if (ctx->post_stream) { cudaStreamDestroy(ctx->post_stream); }
    if (ctx->prep_stream) { cudaStreamDestroy(ctx->prep_stream); }
    if (ctx->stream) { cudaStreamDestroy(ctx->stream); }
#else
    free(ctx->cursor_slab); free(ctx->c_phred2prob); free(ctx->c_pf_tab); free(ctx->c_slip_tab);
#endif
    delete ctx;
}

const char *uvcgpu_last_error(const uvcgpu_ctx *ctx) { return (ctx ? ctx->err.c_str() : "null context"); }

int uvcgpu_set_contig(uvcgpu_ctx *ctx, int32_t tid, const char *bases, int64_t len) {
    if (NULL == ctx || tid < 0 || len < 0) { return UVCGPU_EINVAL; }
    enter_ctx(ctx);
    wait_worker_idle(ctx);
    HostContig & c = ctx->contigs[tid];
    c.len = len;
    c.available = (NULL != bases);
    c.bases.clear();
    if (bases) {
        c.bases.assign(bases, (size_t)len);
        for (auto & ch : c.bases) { ch = (char)toupper(ch); } // load_refstring (main.cpp:65-67)
    }
#if UVC_CUDA
    {   // the bases stay in HBM until the contig is unset: stage P1 reads the reference there
        auto it = ctx->d_contigs.find(tid);
        if (it != ctx->d_contigs.end()) { cudaStreamSynchronize(ctx->stream); cudaFree(it->second); ctx->d_contigs.erase(it); }
        if (bases) {
            char *d = NULL;
            UVC_CUDA_CHECK(ctx, cudaMalloc((void**)&d, (size_t)len + 64));
            cudaError_t e = cudaMemcpy(d, c.bases.data(), (size_t)len, cudaMemcpyHostToDevice);
            if (e != cudaSuccess) { cudaFree(d); UVC_ERR(ctx) = std::string("cudaMemcpy of the contig: ") + cudaGetErrorString(e); return UVCGPU_ECUDA; }
            ctx->d_contigs[tid] = d;
        }
    }
#endif
    return UVCGPU_OK;
}

int uvcgpu_unset_contig(uvcgpu_ctx *ctx, int32_t tid) {
    if (NULL == ctx) { return UVCGPU_EINVAL; }
    enter_ctx(ctx);
    wait_worker_idle(ctx);
    ctx->contigs.erase(tid);
#if UVC_CUDA
    auto it = ctx->d_contigs.find(tid);
    if (it != ctx->d_contigs.end()) { cudaStreamSynchronize(ctx->stream); cudaFree(it->second); ctx->d_contigs.erase(it); }
#endif
    return UVCGPU_OK;
}

int uvcgpu_set_host_threads(uvcgpu_ctx *ctx, int32_t n_threads) {
    if (NULL == ctx || n_threads < 0) { return UVCGPU_EINVAL; }
    ctx->host_threads = n_threads;
    return UVCGPU_OK;
}

int uvcgpu_set_contig_name(uvcgpu_ctx *ctx, int32_t tid, const char *name) {
    if (NULL == ctx || tid < 0 || NULL == name) { return UVCGPU_EINVAL; }
    ctx->contig_names[tid] = name;
    return UVCGPU_OK;
}

// a failed submit: whatever was enqueued for the batch is drained before its memory goes back to the pool
static void backend_abort(uvcgpu_ctx *ctx, BatchState & bs) {
#if UVC_CUDA
    cudaStreamSynchronize(ctx->prep_stream); cudaStreamSynchronize(ctx->stream);
#endif
    backend_free(ctx, bs);
}
#define UVC_TRY(expr) { int rc_ = (expr); if (rc_ != 0) { backend_abort(ctx, *bs); return rc_; } }

int uvcgpu_submit(uvcgpu_ctx *ctx, int32_t n_tiles, const uvcgpu_tile *tiles, const uvcgpu_reads_soa *reads, uvcgpu_ticket *ticket) {
    return uvcgpu_submit_multi(ctx, n_tiles, tiles, 1, reads, NULL, ticket);
}

// the staging of a batch (on the context's worker thread in the CUDA build)
static int submit_body(uvcgpu_ctx *ctx, BatchState *bs, int32_t n_tiles, const uvcgpu_tile *tiles) {
    const double t0 = now_ms();
    HostBatch & hb = bs->hb;
    BatchView & v = bs->view;
    memset(&v, 0, sizeof(v));
    uvc_fill_view_constants(v, ctx->par);
    v.ten_over_ln10 = 10.0 / log(10.0);
    v.ln10 = log(10);

#define UVC_UP(field, type, vec) { void *d_ = NULL; UVC_TRY(backend_alloc(ctx, *bs, &d_, (vec).size() * sizeof(type), false)); \
        UVC_TRY(backend_upload(ctx, *bs, d_, (vec).data(), (vec).size() * sizeof(type))); v.field = (type*)d_; }
#define UVC_ZERO(field, type, count) { void *d_ = NULL; UVC_TRY(backend_alloc(ctx, *bs, &d_, (size_t)(count) * sizeof(type), true)); v.field = (type*)d_; }
    v.phred2prob_tab = ctx->c_phred2prob; v.pf_tab = ctx->c_pf_tab; v.slip_tab = ctx->c_slip_tab;
    // stages P0 and P1 on the device: read filter, family segmentation, reference context (prep_device.inc); fills the view's input arrays
#if UVC_CUDA
    t_active = ctx->prep_stream;
    {
        const int rc_prep = prep_on_device(ctx, *bs, n_tiles, tiles);
        t_active = ctx->stream;
        if (rc_prep != 0) { backend_abort(ctx, *bs); return rc_prep; }
    }
    UVC_CUDA_CHECK(ctx, cudaStreamWaitEvent(ctx->stream, bs->ev_prep[1], 0));     // the pileup kernels start when the staging kernels are done
#else
    UVC_TRY(prep_on_device(ctx, *bs, n_tiles, tiles));
#endif
    const double t1 = now_ms();
    v.n_fcol = hb.n_fcol; v.n_mcol = hb.n_mcol;
    { void *d_ = NULL; UVC_TRY(backend_alloc(ctx, *bs, &d_, (size_t)v.n_fcol * sizeof(FragCol), false)); v.fcol = (FragCol*)d_; }
    UVC_ZERO(fmask, uint32_t, (v.n_fcol / UVC_COL_CHUNK) * 4)
    { void *d_ = NULL; UVC_TRY(backend_alloc(ctx, *bs, &d_, (size_t)v.n_mcol * sizeof(FamCol), false)); v.mcol = (FamCol*)d_; }
    UVC_ZERO(rd, ReadDerived, v.n_reads)
    { void *d_ = NULL; UVC_TRY(backend_alloc(ctx, *bs, &d_, (size_t)v.n_reads * sizeof(PileRec), false)); v.prec = (PileRec*)d_; }
    { void *d_ = NULL; UVC_TRY(backend_alloc(ctx, *bs, &d_, (size_t)v.n_reads * sizeof(PrepRec), false)); v.qrec = (PrepRec*)d_; }
    UVC_ZERO(rfrag, ReadFrag, v.n_reads)
    UVC_ZERO(cx, CxEntry, v.n_cx)
    UVC_ZERO(ev, IndelEvent, v.n_ev)
    UVC_ZERO(prep, uvcgpu_prep_set, v.n_pos)
    UVC_ZERO(thres, uvcgpu_thres_set, v.n_pos)
    UVC_ZERO(seginfo, uvcgpu_seginfo_set, v.n_pos * UVC_NSYM)
    UVC_ZERO(bqsum, int32_t, v.n_pos * UVC_NSYM)
    UVC_ZERO(vq, int32_t, v.n_pos * UVC_NSYM * UVCGPU_NUM_VQ_TAGS)
    UVC_ZERO(fragdepth, int32_t, 2 * v.n_pos * UVC_NSYM * UVCGPU_NUM_FRAG_DEPTHS)
    UVC_ZERO(famdepth, int32_t, 2 * v.n_pos * UVC_NSYM * UVCGPU_NUM_FAM_DEPTHS)
    UVC_ZERO(faminfo, uvcgpu_faminfo_set, v.n_pos * UVC_NSYM)
    UVC_ZERO(duplex, int32_t, v.n_pos * UVC_NSYM * UVCGPU_NUM_DUPLEX_DEPTHS)
    v.frags_in = v.frags;
    v.rec_cap = (int32_t)std::min<int64_t>((int64_t)1 << 30, 16 * v.n_reads + (1 << 20));
    UVC_ZERO(rec_buf, int32_t, v.rec_cap)
    UVC_ZERO(rec_cursor, int32_t, 4)
    { void *d_ = NULL; UVC_TRY(backend_alloc(ctx, *bs, &d_, (size_t)v.n_pos * sizeof(int32_t), false)); v.noindel = (int32_t*)d_; }
    { void *d_ = NULL; UVC_TRY(backend_alloc(ctx, *bs, &d_, (size_t)v.n_pos * sizeof(GvcfPos), true)); bs->d_gvcf = (GvcfPos*)d_; }
    { void *d_ = NULL; UVC_TRY(backend_alloc(ctx, *bs, &d_, (size_t)v.n_pos * sizeof(GvcfExtra), true)); bs->d_gextra = (GvcfExtra*)d_; }
    bs->gvcf.resize((size_t)v.n_pos); bs->gextra.resize((size_t)v.n_pos);
    const double t2 = now_ms();
    UVC_TRY(backend_run(ctx, *bs));
    uvcgpu_batch_stats & st = bs->stats;
    st.n_tiles = n_tiles; st.n_reads_in = hb.n_reads_in; st.n_reads_kept = v.n_reads; st.n_ext_positions = v.n_pos;
    st.n_families = v.n_fams; st.n_fragments = v.n_frags;
    st.n_positions_pileup = (v.list_tile[0] ? v.n_list[0] : v.n_pos); st.n_positions_consensus = (v.list_tile[1] ? v.n_list[1] : v.n_pos);
    for (const auto & T : hb.tiles) { st.n_positions += T.end_pos - T.beg_pos; }
    st.h2d_ms = (t1 - t0) - st.host_prep_ms;    // the rest of the staging call: uploads, staging kernels and their two synchronisations
    (void)t2;
    return UVCGPU_OK;
}

#if UVC_CUDA
static void submit_worker(uvcgpu_ctx *ctx) {
    cudaSetDevice(ctx->device);
    for (;;) {
        SubmitJob job;
        {
            std::unique_lock<std::mutex> lk(ctx->wmu);
            ctx->wcv.wait(lk, [&]() { return ctx->wstop || !ctx->wq.empty(); });
            if (ctx->wq.empty()) { return; }
            job = std::move(ctx->wq.front());
            ctx->wq.pop_front();
        }
        t_active = ctx->stream;
        t_err = &job.bs->submit_err;
        const int rc = submit_body(ctx, job.bs, (int32_t)job.tiles.size(), job.tiles.data());
        t_err = nullptr;
        {
            std::lock_guard<std::mutex> lk(ctx->wmu);
            job.bs->submit_rc = rc;
            job.bs->submitted = true;
        }
        ctx->wdone.notify_all();
    }
}
#endif
// waits until the worker has staged the batch; returns the result of its staging
static int wait_submitted(uvcgpu_ctx *ctx, BatchState & bs) {
#if UVC_CUDA
    std::unique_lock<std::mutex> lk(ctx->wmu);
    ctx->wdone.wait(lk, [&]() { return bs.submitted; });
#endif
    if (bs.submit_rc != 0 && !bs.submit_err.empty()) { ctx->err = bs.submit_err; }
    return bs.submit_rc;
}
// (contigs and parameters must not change under a batch that is being staged)
static void wait_worker_idle(uvcgpu_ctx *ctx) {
    for (auto & kv : ctx->batches) { wait_submitted(ctx, *kv.second); }
}

int uvcgpu_submit_multi(uvcgpu_ctx *ctx, int32_t n_tiles, const uvcgpu_tile *tiles, int32_t n_sources, const uvcgpu_reads_soa *sources,
        const int32_t *tile_source, uvcgpu_ticket *ticket) {
    enter_ctx(ctx);
    if (NULL == ctx || NULL == tiles || NULL == sources || NULL == ticket || n_tiles <= 0 || n_sources <= 0) { return UVCGPU_EINVAL; }
    std::unique_ptr<BatchState> bs(new BatchState());
    memset(&bs->stats, 0, sizeof(bs->stats));
    bs->sources.assign(sources, sources + n_sources);
    bs->tile_source.assign((size_t)n_tiles, 0);
    if (tile_source) {
        for (int32_t k = 0; k < n_tiles; k++) {
            if (tile_source[k] < 0 || tile_source[k] >= n_sources) { UVC_ERR(ctx) = "tile_source out of range"; return UVCGPU_EINVAL; }
            bs->tile_source[(size_t)k] = tile_source[k];
        }
    }
    BatchState *raw = bs.get();
    if (ctx->batches.size() >= (size_t)UVC_CURSOR_SLOTS) { UVC_ERR(ctx) = "too many batches alive in one context"; return UVCGPU_EINVAL; }
    *ticket = ctx->next_ticket++;
    {   // a free slot of the cursor slab
        int slot = (int)(*ticket % UVC_CURSOR_SLOTS);
        for (;;) {
            bool used = false;
            for (auto & kv : ctx->batches) { if (kv.second->rec_cursor_host == ctx->cursor_slab + UVC_CURSOR_WORDS * slot) { used = true; break; } }
            if (!used) { break; }
            slot = (slot + 1) % UVC_CURSOR_SLOTS;
        }
        raw->rec_cursor_host = ctx->cursor_slab + UVC_CURSOR_WORDS * slot; raw->score_cursor_host = raw->rec_cursor_host + 4;
        memset(raw->rec_cursor_host, 0, UVC_CURSOR_WORDS * sizeof(int32_t));
    }
    ctx->batches[*ticket] = std::move(bs);
#if UVC_CUDA
    raw->submitted = false;
    {
        std::lock_guard<std::mutex> lk(ctx->wmu);
        SubmitJob job;
        job.bs = raw; job.tiles.assign(tiles, tiles + n_tiles);
        ctx->wq.push_back(std::move(job));
    }
    ctx->wcv.notify_one();
    return UVCGPU_OK;
#else
    raw->submit_rc = submit_body(ctx, raw, n_tiles, tiles);
    if (raw->submit_rc != 0) { const int rc = raw->submit_rc; ctx->batches.erase(*ticket); return rc; }
    return UVCGPU_OK;
#endif
}

int uvcgpu_collect(uvcgpu_ctx *ctx, uvcgpu_ticket ticket, uvcgpu_batch_stats *stats) {
    enter_ctx(ctx);
    if (NULL == ctx) { return UVCGPU_EINVAL; }
    auto it = ctx->batches.find(ticket);
    if (it == ctx->batches.end()) { UVC_ERR(ctx) = "unknown ticket"; return UVCGPU_EINVAL; }
    BatchState & bs = *it->second;
    if (!bs.collected) {
        int rc = wait_submitted(ctx, bs);
        if (rc != 0) { return rc; }
        rc = backend_wait(ctx, bs);
        if (rc != 0) { return rc; }
#if UVC_CUDA
        if (bs.temps_done) { arena_release(ctx, bs.temp_arena); bs.temps_done = false; }     // the staging kernels ran before the pileup kernels
#endif
        bs.collected = true;
    }
    if (stats) { *stats = bs.stats; }
    return UVCGPU_OK;
}

// Work on a collected batch (downloads, scoring kernels and their allocations) goes to the context's second stream: the batch's own kernels
// are complete (collect waited for its last event), and a later batch may already be queued on the submit stream.
struct PostScope {
#if UVC_CUDA
    uvcgpu_ctx *c;
    explicit PostScope(uvcgpu_ctx *ctx) : c(ctx) { t_active = c->post_stream; }
    ~PostScope() { t_active = c->stream; }
#else
    explicit PostScope(uvcgpu_ctx *) {}
#endif
};

static int ensure_sparse(uvcgpu_ctx *ctx, BatchState & bs) {
    if (bs.sparse_built) { return 0; }
    PostScope post_scope(ctx);
    const BatchView & v = bs.view;
    const int32_t *cursor = bs.rec_cursor_host;      // (arrived with the batch)
    int rc = 0;
    if (cursor[0] > v.rec_cap) { UVC_ERR(ctx) = "sparse record stream overflow: submit a smaller batch"; return UVCGPU_ENOMEM; }
    const double t_sp0 = now_ms();
    StageVec<int32_t> rec((size_t)cursor[0]);
    bs.ev_host.resize((size_t)v.n_ev);
    rc = backend_download_async(ctx, rec.data(), v.rec_buf, rec.size() * sizeof(int32_t));
    if (0 == rc) { rc = backend_download_async(ctx, bs.ev_host.data(), v.ev, bs.ev_host.size() * sizeof(IndelEvent)); }
    if (0 == rc) { rc = backend_sync(ctx); }
    if (rc != 0) { return rc; }
    bs.stats.d2h_bytes += (int64_t)(rec.size() * sizeof(int32_t) + bs.ev_host.size() * sizeof(IndelEvent));
    const double t_sp1 = now_ms();
    uvc_build_sparse(bs.sparse, bs.hb, ctx->par, rec.data(), (int64_t)rec.size(), bs.ev_host.data());
    bs.stats.reserved[2] += t_sp1 - t_sp0;          // download of the sparse record stream and the indel events
    bs.stats.reserved[3] += now_ms() - t_sp1;       // host: indel identity maps, haplotype links
    bs.sparse_built = true;
    return 0;
}

static int ensure_scored(uvcgpu_ctx *ctx, BatchState & bs) {
    if (bs.scored) { return 0; }
    int rc = ensure_sparse(ctx, bs);
    if (rc != 0) { return rc; }
    PostScope post_scope(ctx);
    const double t0 = now_ms();
    BatchView & v = bs.view;
    std::vector<IndelAllele> table_;
    uvc_build_indel_sites(bs.sites, table_, bs.hb, bs.sparse, ctx->contigs, bs.ev_host);
    StageVec<IndelAllele> & table = bs.allele_table;      // (page-locked and alive until release: its upload is asynchronous)
    table.assign(table_.begin(), table_.end());
    const double t1 = now_ms();
    ScoreView sv;
    memset(&sv, 0, sizeof(sv));
    void *d = NULL;
    if ((rc = backend_alloc(ctx, bs, &d, table.size() * sizeof(IndelAllele), false)) != 0) { return rc; }
    if ((rc = backend_upload(ctx, bs, d, table.data(), table.size() * sizeof(IndelAllele))) != 0) { return rc; }
    sv.alleles = (const IndelAllele*)d; sv.n_alleles = (int64_t)table.size();
    sv.gvcf = bs.d_gvcf; sv.gextra = bs.d_gextra;
    if ((rc = backend_alloc(ctx, bs, &d, 32, true)) != 0) { return rc; }
    sv.out_cursor = (int32_t*)d;
    sv.cand_cursor = sv.out_cursor + 1;
    if (v.n_pos > INT32_MAX) { UVC_ERR(ctx) = "batch too large: submit fewer tiles"; return UVCGPU_EINVAL; }
    if ((rc = backend_alloc(ctx, bs, &d, (size_t)v.n_pos * sizeof(int32_t), false)) != 0) { return rc; }
    sv.cand_list = (int32_t*)d;
    int64_t cap = v.n_pos / 16 + 4096, prev_cap = v.n_pos / 256 + 1024;
    // groups and candidates per position: sized from what the context's earlier batches needed (a batch that needs more runs twice)
    int64_t group_cap = (int64_t)(ctx->k5_groups_per_pos * (double)v.n_pos) + 1024, cand_cap = (int64_t)(ctx->k5_cands_per_pos * (double)v.n_pos) + 4096;
    StageVec<VarRec> & recs = bs.recs;
    const size_t n_first = 256;       // records downloaded together with their count (most batches have fewer: one wait instead of two)
    for (int attempt = 0; attempt < 3; attempt++) {
        if ((rc = backend_alloc_temp(ctx, bs, &d, (size_t)cap * sizeof(VarRec))) != 0) { return rc; }
        sv.out = (VarRec*)d; sv.out_cap = (int32_t)cap;
        if ((rc = backend_alloc_temp(ctx, bs, &d, (size_t)group_cap * sizeof(GroupRec))) != 0) { return rc; }
        sv.groups = (GroupRec*)d; sv.group_cap = (int32_t)group_cap;
        if ((rc = backend_alloc_temp(ctx, bs, &d, (size_t)cand_cap * sizeof(CandDesc))) != 0) { return rc; }
        sv.desc = (CandDesc*)d;
        if ((rc = backend_alloc_temp(ctx, bs, &d, (size_t)cand_cap * sizeof(CandFmt))) != 0) { return rc; }
        sv.cands = (CandFmt*)d; sv.cand_cap = (int32_t)cand_cap;
        if ((rc = backend_alloc_temp(ctx, bs, &d, (size_t)prev_cap * sizeof(PrevAllele))) != 0) { return rc; }
        sv.prev = (PrevAllele*)d; sv.prev_cap = (int32_t)prev_cap;
        if ((rc = backend_zero(ctx, sv.out_cursor, 32)) != 0) { return rc; }
        recs.resize(std::min<size_t>(n_first, (size_t)cap));
        if ((rc = backend_score(ctx, bs, sv)) != 0) { return rc; }
        const int32_t n = bs.score_cursor_host[0], n_groups = bs.score_cursor_host[2], n_cands = bs.score_cursor_host[3], n_prev = bs.score_cursor_host[4];
        ctx->k5_groups_per_pos = std::max(ctx->k5_groups_per_pos, std::min(2.0, 1.25 * (double)n_groups / (double)std::max<int64_t>(1, v.n_pos)));
        ctx->k5_cands_per_pos = std::max(ctx->k5_cands_per_pos, std::min(2.0 * UVC_MAX_GROUP_CANDS, 1.25 * (double)n_cands / (double)std::max<int64_t>(1, v.n_pos)));
        if (n <= cap && n_groups <= group_cap && n_cands <= cand_cap && n_prev <= prev_cap) {
            const size_t have = recs.size();
            recs.resize((size_t)n);
            if ((size_t)n > have && (rc = backend_download(ctx, recs.data() + have, sv.out + have, ((size_t)n - have) * sizeof(VarRec))) != 0) { return rc; }
            bs.prev_alleles.resize((size_t)n_prev);
            if (n_prev > 0 && (rc = backend_download(ctx, bs.prev_alleles.data(), sv.prev, (size_t)n_prev * sizeof(PrevAllele))) != 0) { return rc; }
            std::sort(bs.prev_alleles.begin(), bs.prev_alleles.end(), [](const PrevAllele & a, const PrevAllele & b) { return a.rec_slot != b.rec_slot ? a.rec_slot < b.rec_slot : a.order > b.order; });
            break;
        }
        if (attempt == 2) { UVC_ERR(ctx) = "candidate record buffer overflow"; return UVCGPU_ENOMEM; }
        // the kernels counted everything they wanted to write: run again with room for all of it
        cap = std::max<int64_t>(cap, n); group_cap = std::max<int64_t>(group_cap, n_groups); cand_cap = std::max<int64_t>(cand_cap, n_cands); prev_cap = std::max<int64_t>(prev_cap, n_prev);
        backend_free_temps(ctx, bs, true);
    }
    backend_free_temps(ctx, bs, true);      // (the scoring kernels and the downloads of their results have completed)
    bs.stats.d2h_bytes += (int64_t)(recs.size() * sizeof(VarRec));
    const double t2 = now_ms();
    bs.recs_by_tile.assign(bs.hb.tiles.size(), std::vector<const VarRec*>());
    for (const auto & r : recs) { bs.recs_by_tile[(size_t)r.tile].push_back(&r); }
    bs.stats.n_vcf_records = (int64_t)recs.size();
    bs.stats.host_score_ms = (t1 - t0) + (now_ms() - t2);
    bs.stats.reserved[4] += t1 - t0;                // host: indel allele table
    bs.stats.reserved[5] += t2 - t1;                // scoring kernels and the downloads of their results
    bs.scored = true;
    return 0;
}

int uvcgpu_score(uvcgpu_ctx *ctx, uvcgpu_ticket ticket, uvcgpu_batch_stats *stats) {
    enter_ctx(ctx);
    if (NULL == ctx) { return UVCGPU_EINVAL; }
    auto it = ctx->batches.find(ticket);
    if (it == ctx->batches.end()) { UVC_ERR(ctx) = "unknown ticket"; return UVCGPU_EINVAL; }
    BatchState & bs = *it->second;
    if (!bs.collected) { UVC_ERR(ctx) = "batch not collected yet"; return UVCGPU_EINVAL; }
    int rc = ensure_scored(ctx, bs);
    if (rc != 0) { return rc; }
    if (stats) { *stats = bs.stats; }
    return UVCGPU_OK;
}

// VCF body of every tile of the batch, formatted once on the context's host threads (tiles are independent)
static int ensure_vcf_text(uvcgpu_ctx *ctx, BatchState & bs) {
    if (bs.vcf_built) { return 0; }
    int rc = ensure_scored(ctx, bs);
    if (rc != 0) { return rc; }
    const int32_t n_tiles = (int32_t)bs.hb.tiles.size();
    bs.vcf_text.assign((size_t)n_tiles, std::vector<std::string>());
    for (int32_t ti = 0; ti < n_tiles; ti++) {
        const TileInfo & T = bs.hb.tiles[ti];
        if (!T.skipped && ctx->contigs.find(T.tid) == ctx->contigs.end()) { UVC_ERR(ctx) = "contig of a tile was unset before its VCF text was requested"; return UVCGPU_EINVAL; }
    }
    // Tiles are independent, and so are the positions of a tile: a tile is cut into ranges of positions with about the same number of records
    // (a deep tile of a small panel has a record per position and symbol: its text alone is hundreds of megabytes), every range is formatted
    // by one thread, and the ranges' strings are concatenated in order.
    std::vector<TileTextPlan*> plans((size_t)n_tiles, NULL);
    uvc_parallel_for(n_tiles, ctx->host_threads, [&](int32_t ti) {
        if (!bs.hb.tiles[ti].skipped) { plans[ti] = uvc_tile_text_plan_new(bs.hb, ti, ctx->contigs.at(bs.hb.tiles[ti].tid), bs.recs_by_tile[ti], bs.sparse[ti]); }
    });
    struct Range { int32_t ti, zb0, zb1; };
    std::vector<Range> ranges;
    size_t kRecsPerRange = 512;
    { const char *f = getenv("UVC_TEXT_RECS_PER_RANGE"); if (f && atoi(f) > 0) { kRecsPerRange = (size_t)atoi(f); } }      // tests force many small ranges
    for (int32_t ti = 0; ti < n_tiles; ti++) {
        const TileInfo & T = bs.hb.tiles[ti];
        if (T.skipped) { continue; }
        // cut after the position at which the running record count passes a multiple of kRecsPerRange (records are listed by refpos: close enough to zb)
        std::vector<int32_t> zbs;
        for (const VarRec *r : bs.recs_by_tile[ti]) { zbs.push_back(r->symboltype == 0 ? r->refpos + 1 : r->refpos); }
        std::sort(zbs.begin(), zbs.end());
        int32_t zb0 = T.rpos_inclu_beg;
        for (size_t k = kRecsPerRange; k < zbs.size(); k += kRecsPerRange) {
            const int32_t cut = zbs[k] + 1;
            if (cut > zb0 && cut <= T.rpos_exclu_end) { ranges.push_back(Range{ti, zb0, cut}); zb0 = cut; }
        }
        ranges.push_back(Range{ti, zb0, T.rpos_exclu_end + 1});
    }
    std::vector<std::string> parts(ranges.size());
    uvc_parallel_for((int32_t)ranges.size(), ctx->host_threads, [&](int32_t k) {
        const Range & R = ranges[(size_t)k];
        const TileInfo & T = bs.hb.tiles[R.ti];
        auto nm = ctx->contig_names.find(T.tid);
        const std::string tname = (nm == ctx->contig_names.end() ? std::to_string(T.tid) : nm->second);
        parts[(size_t)k] = uvc_tile_vcf_text_range(*plans[R.ti], bs.hb, R.ti, ctx->par, tname, bs.sites[R.ti], bs.sparse[R.ti], bs.ev_host, bs.gvcf.data(), bs.gextra.data(), R.zb0, R.zb1, &bs.prev_alleles);
    });
    for (size_t k = 0; k < ranges.size(); k++) { bs.vcf_text[(size_t)ranges[k].ti].emplace_back(std::move(parts[k])); }     // (no concatenation: the callers copy the parts out)
    for (TileTextPlan *pl : plans) { if (pl) { uvc_tile_text_plan_free(pl); } }
    bs.vcf_built = true;
    return 0;
}

static int tile_vcf_text(uvcgpu_ctx *ctx, BatchState & bs, int32_t tile_index, std::string & out) {
    int rc = ensure_vcf_text(ctx, bs);
    if (rc != 0) { return rc; }
    out.clear();
    size_t total = 0;
    for (const auto & t : bs.vcf_text[tile_index]) { total += t.size(); }
    out.reserve(total);
    for (const auto & t : bs.vcf_text[tile_index]) { out += t; }
    return 0;
}

int uvcgpu_tile_vcf(uvcgpu_ctx *ctx, uvcgpu_ticket ticket, int32_t tile_index, char *dst, size_t cap, size_t *needed) {
    enter_ctx(ctx);
    if (NULL == ctx || NULL == needed) { return UVCGPU_EINVAL; }
    auto it = ctx->batches.find(ticket);
    if (it == ctx->batches.end()) { UVC_ERR(ctx) = "unknown ticket"; return UVCGPU_EINVAL; }
    BatchState & bs = *it->second;
    if (!bs.collected) { UVC_ERR(ctx) = "batch not collected yet"; return UVCGPU_EINVAL; }
    if (tile_index < 0 || tile_index >= (int32_t)bs.hb.tiles.size()) { return UVCGPU_EINVAL; }
    int rc = ensure_vcf_text(ctx, bs);
    if (rc != 0) { return rc; }
    size_t total = 0;
    for (const auto & t : bs.vcf_text[tile_index]) { total += t.size(); }
    *needed = total;
    if (dst && cap) {
        size_t at = 0;
        for (const auto & t : bs.vcf_text[tile_index]) {
            if (at >= cap) { break; }
            const size_t n = (t.size() < cap - at ? t.size() : cap - at);
            memcpy(dst + at, t.data(), n);
            at += n;
        }
    }
    return UVCGPU_OK;
}

int uvcgpu_batch_vcf(uvcgpu_ctx *ctx, uvcgpu_ticket ticket, char *dst, size_t cap, size_t *needed) {
    enter_ctx(ctx);
    if (NULL == ctx || NULL == needed) { return UVCGPU_EINVAL; }
    auto it = ctx->batches.find(ticket);
    if (it == ctx->batches.end()) { UVC_ERR(ctx) = "unknown ticket"; return UVCGPU_EINVAL; }
    BatchState & bs = *it->second;
    if (!bs.collected) { UVC_ERR(ctx) = "batch not collected yet"; return UVCGPU_EINVAL; }
    int rc = ensure_vcf_text(ctx, bs);
    if (rc != 0) { return rc; }
    size_t total = 0;
    std::vector<std::pair<const std::string*, size_t>> pieces;     // (part, its offset in the batch's text)
    for (const auto & tile : bs.vcf_text) { for (const auto & t : tile) { pieces.push_back(std::make_pair(&t, total)); total += t.size(); } }
    *needed = total;
    if (dst && cap) {
        uvc_parallel_for((int32_t)pieces.size(), (total >= ((size_t)8 << 20) ? ctx->host_threads : 1), [&](int32_t k) {
            const std::string & t = *pieces[(size_t)k].first;
            const size_t at = pieces[(size_t)k].second;
            if (at >= cap) { return; }
            memcpy(dst + at, t.data(), (t.size() < cap - at ? t.size() : cap - at));
        });
    }
    return UVCGPU_OK;
}

int uvcgpu_selftest_math(uvcgpu_ctx *ctx, int32_t which, const double *in, int32_t n, double *out) {
    enter_ctx(ctx);
    if (NULL == ctx || NULL == in || NULL == out || n <= 0 || which < 0 || which > 3) { return UVCGPU_EINVAL; }
    BatchView v;
    memset(&v, 0, sizeof(v));
    uvc_fill_view_constants(v, ctx->par);
    v.ten_over_ln10 = 10.0 / log(10.0);
    v.ln10 = log(10);
#if UVC_CUDA
    double *d_in = NULL, *d_out = NULL;
    UVC_CUDA_CHECK(ctx, cudaMalloc((void**)&d_in, (size_t)n * 3 * sizeof(double)));
    if (cudaMalloc((void**)&d_out, (size_t)n * sizeof(double)) != cudaSuccess) { cudaFree(d_in); UVC_ERR(ctx) = "cudaMalloc"; return UVCGPU_ECUDA; }
    cudaMemcpy(d_in, in, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice);
    uvc_selftest_math_kernel<<<(unsigned)((n + 63) / 64), 64, 0, ctx->stream>>>(v, which, d_in, n, d_out);
    const cudaError_t e = cudaMemcpy(out, d_out, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(d_in); cudaFree(d_out);
    if (e != cudaSuccess) { UVC_ERR(ctx) = std::string("uvcgpu_selftest_math: ") + cudaGetErrorString(e); return UVCGPU_ECUDA; }
#else
    for (int32_t i = 0; i < n; i++) { out[i] = uvc::selftest_math(v, which, in[3 * i], in[3 * i + 1], in[3 * i + 2]); }
#endif
    return UVCGPU_OK;
}

int uvcgpu_release(uvcgpu_ctx *ctx, uvcgpu_ticket ticket) {
    enter_ctx(ctx);
    if (NULL == ctx) { return UVCGPU_EINVAL; }
    auto it = ctx->batches.find(ticket);
    if (it == ctx->batches.end()) { return UVCGPU_EINVAL; }
    if (!it->second->collected && 0 == wait_submitted(ctx, *it->second)) { backend_wait(ctx, *it->second); }   // (a collected batch has nothing left on the submit stream)
#if UVC_CUDA
    backend_wait_stream(ctx, ctx->post_stream);                        // idle unless a scoring call failed half-way
#endif
    backend_free(ctx, *it->second);
    ctx->batches.erase(it);
    return UVCGPU_OK;
}

int uvcgpu_dump_counters(uvcgpu_ctx *ctx, uvcgpu_ticket ticket, int32_t tile_index, int32_t section, void *dst, size_t cap, size_t *needed) {
    enter_ctx(ctx);
    if (NULL == ctx || NULL == needed) { return UVCGPU_EINVAL; }
    auto it = ctx->batches.find(ticket);
    if (it == ctx->batches.end()) { UVC_ERR(ctx) = "unknown ticket"; return UVCGPU_EINVAL; }
    BatchState & bs = *it->second;
    if (!bs.collected) { UVC_ERR(ctx) = "batch not collected yet"; return UVCGPU_EINVAL; }
    if (tile_index < 0 || tile_index >= (int32_t)bs.hb.tiles.size()) { return UVCGPU_EINVAL; }
    PostScope post_scope(ctx);
    const TileInfo & T = bs.hb.tiles[tile_index];
    const BatchView & v = bs.view;
    const size_t npos = (size_t)(T.ext_end - T.ext_beg);
    const size_t off = (size_t)T.pos_off;
    std::vector<uint8_t> tmp;
    const void *src = NULL;      // compute-side source
    size_t bytes = 0;
    bool host_side = false;
    switch (section) {
        case UVCGPU_SEC_META: {
            int64_t meta[16];
            memset(meta, 0, sizeof(meta));
            meta[UVCGPU_META_NUM_PASSED] = T.num_passed; meta[UVCGPU_META_NUM_PCRPASSED] = T.num_pcrpassed;
            meta[UVCGPU_META_BAM_BEG] = T.bam_inclu_beg; meta[UVCGPU_META_BAM_END] = T.bam_exclu_end;
            meta[UVCGPU_META_RPOS_BEG] = T.rpos_inclu_beg; meta[UVCGPU_META_RPOS_END] = T.rpos_exclu_end;
            meta[UVCGPU_META_EXT_BEG] = T.ext_beg; meta[UVCGPU_META_EXT_END] = T.ext_end - 1;
            meta[UVCGPU_META_NUM_FAMILIES] = T.n_fams;
            tmp.assign((uint8_t*)meta, (uint8_t*)meta + sizeof(meta));
            host_side = true; break;
        }
        case UVCGPU_SEC_FAMILIES: {
            if (!bs.groups_on_host) {     // the segmentation lives on the device: fetched for this test hook only
                bs.hb.reads.resize((size_t)v.n_reads); bs.hb.frags.resize((size_t)v.n_frags); bs.hb.fams.resize((size_t)v.n_fams); bs.hb.frag_reads.resize((size_t)v.n_reads);
                int rc = backend_download(ctx, bs.hb.reads.data(), v.reads, bs.hb.reads.size() * sizeof(ReadRec));
                if (0 == rc) { rc = backend_download(ctx, bs.hb.frags.data(), v.frags, bs.hb.frags.size() * sizeof(FragRec)); }
                if (0 == rc) { rc = backend_download(ctx, bs.hb.fams.data(), v.fams, bs.hb.fams.size() * sizeof(FamRec)); }
                if (0 == rc) { rc = backend_download(ctx, bs.hb.frag_reads.data(), v.frag_reads, bs.hb.frag_reads.size() * sizeof(int32_t)); }
                if (rc != 0) { return rc; }
                bs.groups_on_host = true;
            }
            const std::string s = uvc_families_text(bs.hb, tile_index);
            tmp.assign(s.begin(), s.end());
            host_side = true; break;
        }
        case UVCGPU_SEC_INDELMAPS: case UVCGPU_SEC_HAPLINKS: {
            int rc = ensure_sparse(ctx, bs);
            if (rc != 0) { return rc; }
            const std::string s = (section == UVCGPU_SEC_INDELMAPS ? uvc_indelmaps_text(bs.sparse[tile_index]) : uvc_haplinks_text(bs.sparse[tile_index]));
            tmp.assign(s.begin(), s.end());
            host_side = true; break;
        }
        case UVCGPU_SEC_VCF: {
            std::string s;
            int rc = tile_vcf_text(ctx, bs, tile_index, s);
            if (rc != 0) { return rc; }
            tmp.assign(s.begin(), s.end());
            host_side = true; break;
        }
        case UVCGPU_SEC_RTR_INITIAL: {
            std::vector<uvcgpu_rtr> r(npos);
            std::vector<int32_t> ip(npos);
            int rc = (T.skipped ? 0 : backend_download(ctx, r.data(), v.rtr + off, npos * sizeof(uvcgpu_rtr)));
            if (0 == rc && !T.skipped) { rc = backend_download(ctx, ip.data(), bs.indelphred0 + off, npos * sizeof(int32_t)); }
            if (rc != 0) { return rc; }
            for (size_t i = 0; i < npos; i++) { r[i].indelphred = ip[i]; }
            tmp.assign((const uint8_t*)r.data(), (const uint8_t*)(r.data() + npos));
            host_side = true; break;
        }
        case UVCGPU_SEC_BAQ: case UVCGPU_SEC_BAQ2: {
            std::vector<int32_t> b(npos);
            int rc = (T.skipped ? 0 : backend_download(ctx, b.data(), (section == UVCGPU_SEC_BAQ ? v.baq : v.baq2) + off, npos * sizeof(int32_t)));
            if (rc != 0) { return rc; }
            std::vector<int64_t> w(npos);
            for (size_t i = 0; i < npos; i++) { w[i] = b[i]; }
            tmp.assign((uint8_t*)w.data(), (uint8_t*)(w.data() + npos));
            host_side = true; break;
        }
        case UVCGPU_SEC_RTR: src = v.rtr + off; bytes = npos * sizeof(uvcgpu_rtr); break;
        case UVCGPU_SEC_PREP: src = v.prep + off; bytes = npos * sizeof(uvcgpu_prep_set); break;
        case UVCGPU_SEC_THRES: src = v.thres + off; bytes = npos * sizeof(uvcgpu_thres_set); break;
        case UVCGPU_SEC_SEGINFO: src = v.seginfo + off * UVC_NSYM; bytes = npos * UVC_NSYM * sizeof(uvcgpu_seginfo_set); break;
        case UVCGPU_SEC_FAMINFO: src = v.faminfo + off * UVC_NSYM; bytes = npos * UVC_NSYM * sizeof(uvcgpu_faminfo_set); break;
        case UVCGPU_SEC_FRAGDEPTH0: case UVCGPU_SEC_FRAGDEPTH1: {
            const size_t strand = (section == UVCGPU_SEC_FRAGDEPTH1);
            src = v.fragdepth + (strand * (size_t)v.n_pos + off) * UVC_NSYM * UVCGPU_NUM_FRAG_DEPTHS; bytes = npos * UVC_NSYM * UVCGPU_NUM_FRAG_DEPTHS * 4; break;
        }
        case UVCGPU_SEC_FAMDEPTH0: case UVCGPU_SEC_FAMDEPTH1: {
            const size_t strand = (section == UVCGPU_SEC_FAMDEPTH1);
            src = v.famdepth + (strand * (size_t)v.n_pos + off) * UVC_NSYM * UVCGPU_NUM_FAM_DEPTHS; bytes = npos * UVC_NSYM * UVCGPU_NUM_FAM_DEPTHS * 4; break;
        }
        case UVCGPU_SEC_DUPLEX: src = v.duplex + off * UVC_NSYM * UVCGPU_NUM_DUPLEX_DEPTHS; bytes = npos * UVC_NSYM * UVCGPU_NUM_DUPLEX_DEPTHS * 4; break;
        case UVCGPU_SEC_VQ: src = v.vq + off * UVC_NSYM * UVCGPU_NUM_VQ_TAGS; bytes = npos * UVC_NSYM * UVCGPU_NUM_VQ_TAGS * 4; break;
        default: UVC_ERR(ctx) = "unknown section"; return UVCGPU_EINVAL;
    }
    if (T.skipped && !host_side) { bytes = 0; }
    if (host_side) {
        *needed = tmp.size();
        if (dst && cap) { memcpy(dst, tmp.data(), tmp.size() < cap ? tmp.size() : cap); }
        return UVCGPU_OK;
    }
    *needed = bytes;
    if (dst && cap && bytes) {
        int rc = backend_download(ctx, dst, src, bytes < cap ? bytes : cap);
        if (rc != 0) { return rc; }
    }
    return UVCGPU_OK;
}

} // extern "C"
