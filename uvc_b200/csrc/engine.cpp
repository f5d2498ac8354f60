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
#include "sparse_out.h"

#include <algorithm>
#include <chrono>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#if defined(__CUDACC__) && !defined(UVC_EMU)
#define UVC_CUDA 1
#include <cuda_runtime.h>
#else
#define UVC_CUDA 0
#endif

namespace {

double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct BatchState {
    HostBatch hb;
    BatchView view;                       // pointers valid on the compute side (device or, in the emulation, host)
    std::vector<void*> allocs;            // compute-side allocations
    std::vector<std::pair<void*, size_t>> alloc_sizes;
    uvcgpu_reads_soa reads;               // caller's SoA (borrowed until release)
    uvcgpu_batch_stats stats;
    bool collected = false;
    bool sparse_built = false;
    std::vector<TileSparse> sparse;
    std::vector<IndelEvent> ev_host;
#if UVC_CUDA
    cudaEvent_t ev[12];
    bool have_events = false;
#endif
};

} // namespace

struct uvcgpu_ctx {
    int device = 0;
    uvcgpu_params par;
    std::string err;
    std::map<int32_t, HostContig> contigs;
    std::map<uvcgpu_ticket, std::unique_ptr<BatchState>> batches;
    uvcgpu_ticket next_ticket = 1;
    std::vector<int32_t> slip_tab;
#if UVC_CUDA
    cudaStream_t stream = nullptr;
#endif
};

// ------------------------------------------------------------------------------------------------ compute backend
#if UVC_CUDA

#define UVC_CUDA_CHECK(ctx, call) { cudaError_t e_ = (call); if (e_ != cudaSuccess) { (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e_); return UVCGPU_ECUDA; } }

// One named __global__ per stage (so that profiles list them by name); every thread handles one work item.
#define UVC_DEFINE_KERNEL(name, call) \
    __global__ void __launch_bounds__(128) name(const BatchView v, int64_t n) { \
        const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; \
        if (i < n) { call; } \
    }
UVC_DEFINE_KERNEL(uvc_k0_read_consts, uvc::k0_read(v, i))
UVC_DEFINE_KERNEL(uvc_k1_prep_thres, uvc::k1_position(v, i))
// both roles of a position sit in different warps of the same block: block = 64 positions x 2 roles
UVC_DEFINE_KERNEL(uvc_k2_bias_pileup, { const int64_t gp = (i / 128) * 64 + (i % 64); if (gp < v.n_pos) { uvc::k2_position(v, gp, (int)((i % 128) / 64)); } })
UVC_DEFINE_KERNEL(uvc_k2e_indel_events, uvc::k2e_event(v, i))
UVC_DEFINE_KERNEL(uvc_k3a_fragment_stats, uvc::k3a_fragment(v, i))
UVC_DEFINE_KERNEL(uvc_k3b_fragment_consensus, uvc::k3b_position(v, i))
UVC_DEFINE_KERNEL(uvc_k4a_family_ends, uvc::k4a_family_strand(v, i))
UVC_DEFINE_KERNEL(uvc_k4_family_consensus, uvc::k4_position(v, i))
UVC_DEFINE_KERNEL(uvc_k4c_family_haplotypes, uvc::k4c_family_strand(v, i))

typedef void (*uvc_kernel_t)(const BatchView, int64_t);
static void launch(uvc_kernel_t k, cudaStream_t s, const BatchView & v, int64_t n, int64_t & launches) {
    if (n <= 0) { return; }
    const int threads = 128;
    k<<<(unsigned)((n + threads - 1) / threads), threads, 0, s>>>(v, n);
    launches++;
}

static int backend_alloc(uvcgpu_ctx *ctx, BatchState & bs, void **out, size_t bytes, bool zero) {
    if (0 == bytes) { bytes = 16; }
    UVC_CUDA_CHECK(ctx, cudaMalloc(out, bytes));
    bs.allocs.push_back(*out);
    if (zero) { UVC_CUDA_CHECK(ctx, cudaMemsetAsync(*out, 0, bytes, ctx->stream)); }
    return 0;
}
static int backend_upload(uvcgpu_ctx *ctx, BatchState & bs, void *dst, const void *src, size_t bytes) {
    if (bytes) { UVC_CUDA_CHECK(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream)); bs.stats.h2d_bytes += (int64_t)bytes; }
    return 0;
}
static int backend_download(uvcgpu_ctx *ctx, void *dst, const void *src, size_t bytes) {
    if (bytes) { UVC_CUDA_CHECK(ctx, cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost)); }
    return 0;
}
static void backend_free(BatchState & bs) { for (void *p : bs.allocs) { cudaFree(p); } bs.allocs.clear(); }

static int backend_run(uvcgpu_ctx *ctx, BatchState & bs) {
    const BatchView & v = bs.view;
    int64_t launches = 0;
    for (int i = 0; i < 12; i++) { UVC_CUDA_CHECK(ctx, cudaEventCreate(&bs.ev[i])); }
    bs.have_events = true;
    UVC_CUDA_CHECK(ctx, cudaEventRecord(bs.ev[0], ctx->stream));
    launch(uvc_k0_read_consts, ctx->stream, v, v.n_reads, launches);
    UVC_CUDA_CHECK(ctx, cudaEventRecord(bs.ev[1], ctx->stream));
    launch(uvc_k1_prep_thres, ctx->stream, v, v.n_pos, launches);
    UVC_CUDA_CHECK(ctx, cudaEventRecord(bs.ev[2], ctx->stream));
    launch(uvc_k2_bias_pileup, ctx->stream, v, ((v.n_pos + 63) / 64) * 128, launches);
    UVC_CUDA_CHECK(ctx, cudaEventRecord(bs.ev[3], ctx->stream));
    launch(uvc_k2e_indel_events, ctx->stream, v, v.n_ev, launches);
    UVC_CUDA_CHECK(ctx, cudaEventRecord(bs.ev[4], ctx->stream));
    launch(uvc_k3a_fragment_stats, ctx->stream, v, v.n_frags, launches);
    UVC_CUDA_CHECK(ctx, cudaEventRecord(bs.ev[5], ctx->stream));
    launch(uvc_k3b_fragment_consensus, ctx->stream, v, v.n_pos, launches);
    UVC_CUDA_CHECK(ctx, cudaEventRecord(bs.ev[6], ctx->stream));
    launch(uvc_k4a_family_ends, ctx->stream, v, 2 * v.n_fams, launches);
    UVC_CUDA_CHECK(ctx, cudaEventRecord(bs.ev[7], ctx->stream));
    launch(uvc_k4_family_consensus, ctx->stream, v, v.n_pos, launches);
    UVC_CUDA_CHECK(ctx, cudaEventRecord(bs.ev[8], ctx->stream));
    launch(uvc_k4c_family_haplotypes, ctx->stream, v, 2 * v.n_fams, launches);
    UVC_CUDA_CHECK(ctx, cudaEventRecord(bs.ev[9], ctx->stream));
    UVC_CUDA_CHECK(ctx, cudaGetLastError());
    bs.stats.gpu_launches = launches;
    return 0;
}

static int backend_wait(uvcgpu_ctx *ctx, BatchState & bs) {
    UVC_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    if (bs.have_events) {
        float ms = 0;
        double total = 0;
        for (int i = 0; i < 9; i++) {
            UVC_CUDA_CHECK(ctx, cudaEventElapsedTime(&ms, bs.ev[i], bs.ev[i + 1]));
            bs.stats.kernel_ms_by_stage[i] = ms;
            total += ms;
        }
        bs.stats.kernel_ms = total;
        for (int i = 0; i < 12; i++) { cudaEventDestroy(bs.ev[i]); }
        bs.have_events = false;
    }
    return 0;
}

#else // ------------------------------------------------------------------------------------------- emulation (tests only)

static int backend_alloc(uvcgpu_ctx *, BatchState & bs, void **out, size_t bytes, bool) {
    if (0 == bytes) { bytes = 16; }
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
static void backend_free(BatchState & bs) { for (void *p : bs.allocs) { free(p); } bs.allocs.clear(); }
static int backend_run(uvcgpu_ctx *, BatchState & bs) {
    const BatchView & v = bs.view;
    for (int64_t i = 0; i < v.n_reads; i++) { uvc::k0_read(v, i); }
    for (int64_t i = 0; i < v.n_pos; i++) { uvc::k1_position(v, i); }
    for (int64_t i = 0; i < v.n_pos; i++) { uvc::k2_position(v, i, 0); uvc::k2_position(v, i, 1); }
    for (int64_t i = 0; i < v.n_ev; i++) { uvc::k2e_event(v, i); }
    for (int64_t i = 0; i < v.n_frags; i++) { uvc::k3a_fragment(v, i); }
    for (int64_t i = 0; i < v.n_pos; i++) { uvc::k3b_position(v, i); }
    for (int64_t i = 0; i < 2 * v.n_fams; i++) { uvc::k4a_family_strand(v, i); }
    for (int64_t i = 0; i < v.n_pos; i++) { uvc::k4_position(v, i); }
    for (int64_t i = 0; i < 2 * v.n_fams; i++) { uvc::k4c_family_strand(v, i); }
    bs.stats.gpu_launches = 0;
    return 0;
}
static int backend_wait(uvcgpu_ctx *, BatchState &) { return 0; }

#endif

// ------------------------------------------------------------------------------------------------ C ABI
extern "C" {

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
#if UVC_CUDA
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return UVCGPU_ECUDA; }
#endif
    *out = ctx;
    return UVCGPU_OK;
}

void uvcgpu_destroy(uvcgpu_ctx *ctx) {
    if (NULL == ctx) { return; }
    for (auto & kv : ctx->batches) { backend_free(*kv.second); }
#if UVC_CUDA
    if (ctx->stream) { cudaStreamDestroy(ctx->stream); }
#endif
    delete ctx;
}

const char *uvcgpu_last_error(const uvcgpu_ctx *ctx) { return (ctx ? ctx->err.c_str() : "null context"); }

int uvcgpu_set_contig(uvcgpu_ctx *ctx, int32_t tid, const char *bases, int64_t len) {
    if (NULL == ctx || tid < 0 || len < 0) { return UVCGPU_EINVAL; }
    HostContig & c = ctx->contigs[tid];
    c.len = len;
    c.available = (NULL != bases);
    c.bases.clear();
    if (bases) {
        c.bases.assign(bases, (size_t)len);
        for (auto & ch : c.bases) { ch = (char)toupper(ch); } // load_refstring (main.cpp:65-67)
    }
    return UVCGPU_OK;
}

#define UVC_TRY(expr) { int rc_ = (expr); if (rc_ != 0) { backend_free(*bs); return rc_; } }

int uvcgpu_submit(uvcgpu_ctx *ctx, int32_t n_tiles, const uvcgpu_tile *tiles, const uvcgpu_reads_soa *reads, uvcgpu_ticket *ticket) {
    if (NULL == ctx || NULL == tiles || NULL == reads || NULL == ticket || n_tiles <= 0) { return UVCGPU_EINVAL; }
    std::unique_ptr<BatchState> bs(new BatchState());
    memset(&bs->stats, 0, sizeof(bs->stats));
    bs->reads = *reads;
    const double t0 = now_ms();
    std::string msg;
    int rc = uvc_build_host_batch(bs->hb, ctx->par, ctx->contigs, n_tiles, tiles, *reads, msg);
    if (rc != 0) { ctx->err = msg; return rc; }
    const double t1 = now_ms();
    HostBatch & hb = bs->hb;
    BatchView & v = bs->view;
    memset(&v, 0, sizeof(v));
    uvc_fill_view_constants(v, ctx->par);
    v.ten_over_ln10 = 10.0 / log(10.0);
    v.ln10 = log(10);
    v.n_tiles = n_tiles;
    v.n_pos = hb.n_pos; v.n_reads = (int64_t)hb.reads.size(); v.n_frags = (int64_t)hb.frags.size(); v.n_fams = (int64_t)hb.fams.size();
    v.n_cx = hb.n_cx; v.n_ev = hb.n_ev;

#define UVC_UP(field, type, vec) { void *d_ = NULL; UVC_TRY(backend_alloc(ctx, *bs, &d_, (vec).size() * sizeof(type), false)); \
        UVC_TRY(backend_upload(ctx, *bs, d_, (vec).data(), (vec).size() * sizeof(type))); v.field = (type*)d_; }
#define UVC_ZERO(field, type, count) { void *d_ = NULL; UVC_TRY(backend_alloc(ctx, *bs, &d_, (size_t)(count) * sizeof(type), true)); v.field = (type*)d_; }
    {
        std::vector<double> tab(128);
        for (int q = 0; q < 128; q++) { tab[q] = pow(10, -((float)q) / 10); }   // phred2prob (main_conversion.hpp:885-888): float exponent, double pow
        UVC_UP(phred2prob_tab, double, tab)
    }
    UVC_UP(tiles, TileInfo, hb.tiles)
    UVC_UP(pos_tile, int32_t, hb.pos_tile)
    UVC_UP(refsym, uint8_t, hb.refsym)
    UVC_UP(rtr, uvcgpu_rtr, hb.rtr)
    UVC_UP(baq, int32_t, hb.baq)
    UVC_UP(baq2, int32_t, hb.baq2)
    UVC_UP(reads, ReadRec, hb.reads)
    UVC_UP(seq, uint8_t, hb.seq)
    UVC_UP(qual, uint8_t, hb.qual)
    UVC_UP(cigar, uint32_t, hb.cigar)
    UVC_UP(frags, FragRec, hb.frags)
    UVC_UP(frag_reads, int32_t, hb.frag_reads)
    UVC_UP(fams, FamRec, hb.fams)
    UVC_UP(slip_tab, int32_t, ctx->slip_tab)
    UVC_ZERO(rd, ReadDerived, v.n_reads)
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
    const double t2 = now_ms();
    UVC_TRY(backend_run(ctx, *bs));
    uvcgpu_batch_stats & st = bs->stats;
    st.n_tiles = n_tiles; st.n_reads_in = hb.n_reads_in; st.n_reads_kept = v.n_reads; st.n_ext_positions = v.n_pos;
    st.n_families = v.n_fams; st.n_fragments = v.n_frags;
    for (const auto & T : hb.tiles) { st.n_positions += T.end_pos - T.beg_pos; }
    st.host_prep_ms = t1 - t0;
    st.h2d_ms = t2 - t1;
    *ticket = ctx->next_ticket++;
    ctx->batches[*ticket] = std::move(bs);
    return UVCGPU_OK;
}

int uvcgpu_collect(uvcgpu_ctx *ctx, uvcgpu_ticket ticket, uvcgpu_batch_stats *stats) {
    if (NULL == ctx) { return UVCGPU_EINVAL; }
    auto it = ctx->batches.find(ticket);
    if (it == ctx->batches.end()) { ctx->err = "unknown ticket"; return UVCGPU_EINVAL; }
    BatchState & bs = *it->second;
    if (!bs.collected) {
        int rc = backend_wait(ctx, bs);
        if (rc != 0) { return rc; }
        bs.collected = true;
    }
    if (stats) { *stats = bs.stats; }
    return UVCGPU_OK;
}

static int ensure_sparse(uvcgpu_ctx *ctx, BatchState & bs) {
    if (bs.sparse_built) { return 0; }
    const BatchView & v = bs.view;
    int32_t cursor[4] = {0, 0, 0, 0};
    int rc = backend_download(ctx, cursor, v.rec_cursor, sizeof(cursor));
    if (rc != 0) { return rc; }
    if (cursor[0] > v.rec_cap) { ctx->err = "sparse record stream overflow: submit a smaller batch"; return UVCGPU_ENOMEM; }
    std::vector<int32_t> rec((size_t)cursor[0]);
    rc = backend_download(ctx, rec.data(), v.rec_buf, rec.size() * sizeof(int32_t));
    if (rc != 0) { return rc; }
    bs.ev_host.resize((size_t)v.n_ev);
    rc = backend_download(ctx, bs.ev_host.data(), v.ev, bs.ev_host.size() * sizeof(IndelEvent));
    if (rc != 0) { return rc; }
    bs.stats.d2h_bytes += (int64_t)(rec.size() * sizeof(int32_t) + bs.ev_host.size() * sizeof(IndelEvent));
    uvc_build_sparse(bs.sparse, bs.hb, ctx->par, rec.data(), (int64_t)rec.size(), bs.ev_host.data());
    bs.sparse_built = true;
    return 0;
}

int uvcgpu_release(uvcgpu_ctx *ctx, uvcgpu_ticket ticket) {
    if (NULL == ctx) { return UVCGPU_EINVAL; }
    auto it = ctx->batches.find(ticket);
    if (it == ctx->batches.end()) { return UVCGPU_EINVAL; }
    backend_wait(ctx, *it->second);
    backend_free(*it->second);
    ctx->batches.erase(it);
    return UVCGPU_OK;
}

int uvcgpu_dump_counters(uvcgpu_ctx *ctx, uvcgpu_ticket ticket, int32_t tile_index, int32_t section, void *dst, size_t cap, size_t *needed) {
    if (NULL == ctx || NULL == needed) { return UVCGPU_EINVAL; }
    auto it = ctx->batches.find(ticket);
    if (it == ctx->batches.end()) { ctx->err = "unknown ticket"; return UVCGPU_EINVAL; }
    BatchState & bs = *it->second;
    if (!bs.collected) { ctx->err = "batch not collected yet"; return UVCGPU_EINVAL; }
    if (tile_index < 0 || tile_index >= (int32_t)bs.hb.tiles.size()) { return UVCGPU_EINVAL; }
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
            const std::string s = uvc_families_text(bs.hb, tile_index, bs.reads);
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
        case UVCGPU_SEC_RTR_INITIAL: {
            tmp.assign((const uint8_t*)(bs.hb.rtr.data() + off), (const uint8_t*)(bs.hb.rtr.data() + off + npos));
            host_side = true; break;
        }
        case UVCGPU_SEC_BAQ: case UVCGPU_SEC_BAQ2: {
            const std::vector<int32_t> & b = (section == UVCGPU_SEC_BAQ ? bs.hb.baq : bs.hb.baq2);
            std::vector<int64_t> w(npos);
            for (size_t i = 0; i < npos; i++) { w[i] = b[off + i]; }
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
        default: ctx->err = "unknown section"; return UVCGPU_EINVAL;
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
