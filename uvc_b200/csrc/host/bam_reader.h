// bam_reader.h - host-side input substrate of the B200 uvc1 host: BGZF/BAM/BAI and FASTA/FAI readers that decode
// straight into the structure-of-arrays record layout of include/uvcgpu.h (uvcgpu_reads_soa).
// Replaces the reference's use of htslib (sam_open/sam_hdr_read/sam_index_load/sam_itr_queryi/sam_itr_next,
// grouping.cpp:179-193, 664-666, 730-731; fai_load/faidx_fetch_seq, main.cpp:54-70).
#ifndef UVC_BAM_READER_H_INCLUDED
#define UVC_BAM_READER_H_INCLUDED

#include "../../../include/uvcgpu.h"

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct uvchost_bam uvchost_bam;
typedef struct uvchost_readbuf uvchost_readbuf;
typedef struct uvchost_fasta uvchost_fasta;

uvchost_bam *uvchost_bam_open(const char *path);          /* also loads <path>.bai */
/* 1: <path>.bai loaded and consistent with the BAM header; 0: no index file; -1: an index file exists but is invalid (bad magic, truncated, or a
 * different number of reference sequences). Region queries fail without a valid index (the reference exits with "Failed to load BAM index",
 * main.cpp:1307-1311). */
int uvchost_bam_index_status(const uvchost_bam *b);
void uvchost_bam_close(uvchost_bam *b);
int32_t uvchost_bam_n_targets(const uvchost_bam *b);
const char *uvchost_bam_target_name(const uvchost_bam *b, int32_t tid);
int64_t uvchost_bam_target_len(const uvchost_bam *b, int32_t tid);

uvchost_readbuf *uvchost_readbuf_new(void);
void uvchost_readbuf_free(uvchost_readbuf *rb);
void uvchost_readbuf_clear(uvchost_readbuf *rb);
int64_t uvchost_readbuf_size(const uvchost_readbuf *rb);
/* Fills out with pointers into rb (valid until the next append/clear/free). */
void uvchost_readbuf_view(const uvchost_readbuf *rb, uvcgpu_reads_soa *out);

/* Appends, in file order, every record with this tid, pos < end and bam_endpos > beg (the records htslib's
 * sam_itr_queryi(idx, tid, beg, end) iterates). Returns the number appended or a negative error. */
int64_t uvchost_bam_fetch(uvchost_bam *b, int32_t tid, int64_t beg, int64_t end, uvchost_readbuf *rb);

/* The fetch windows [begs[k], ends[k]) of n tiles of one contig, given in ascending order of begs: fills, for every tile, the records that
 * uvchost_bam_fetch would append for it, in the same order, and reports their slice [read_begin[k], read_end[k]) of rb. When the windows are
 * dense (small panel targets, tile halos) the span is inflated and parsed ONCE and the records are copied to every tile that they overlap,
 * instead of once per tile from the start of its 16 kbp index window. Returns the number of records appended or a negative error. */
int64_t uvchost_bam_fetch_tiles(uvchost_bam *b, int32_t tid, int32_t n, const int64_t *begs, const int64_t *ends, uvchost_readbuf *rb,
                                int64_t *read_begin, int64_t *read_end);

/* Like uvchost_bam_fetch_tiles for windows given in ascending order of BOTH begs and ends, but every record is appended ONCE: tile k's slice
 * [read_begin[k], read_end[k]) runs from the first record that ends after begs[k] to the last record that starts before ends[k], so the slices
 * of neighbouring tiles overlap (tile halos are not duplicated: five times fewer bytes on a dense panel). A slice is a file-order superset of
 * what sam_itr_queryi yields for the window: the few extra records (pos < beg, end <= beg) start at least 2000 bases before the tile and are
 * dropped by the read filter exactly like the reference's OUT_OF_RANGE test (grouping.cpp:408-409), so uvcgpu_submit gives identical results.
 * Returns the number of records appended, -1 on a read error, -2 if the windows are not ascending. */
int64_t uvchost_bam_fetch_span(uvchost_bam *b, int32_t tid, int32_t n, const int64_t *begs, const int64_t *ends, uvchost_readbuf *rb,
                               int64_t *read_begin, int64_t *read_end);

/* Sequential scan of core fields (for the region tiler): calls cb(tid, pos, endpos, flag, isize, user) for every record in file order
 * until cb returns non-zero or the file ends. */
typedef int (*uvchost_scan_cb)(int32_t tid, int32_t pos, int32_t endpos, uint16_t flag, int32_t isize, int32_t l_qseq, uint8_t mapq, void *user);
int uvchost_bam_scan(uvchost_bam *b, uvchost_scan_cb cb, void *user);

/* Resumable sequential reader of core fields (used by the region tiler, tiler.h): rewind to the first record, then pull records one by one.
 * next returns 1 and fills *c, 0 at end of file (*c untouched, like htslib's sam_read1), negative on error. */
typedef struct uvchost_core { int32_t tid, pos, endpos, isize, l_qseq; uint16_t flag; uint8_t mapq; } uvchost_core;
int uvchost_bam_rewind(uvchost_bam *b);
/* Same, with the following blocks inflated ahead on n_threads threads (whole-file scans); any later seek returns to on-demand inflation. */
int uvchost_bam_rewind_parallel(uvchost_bam *b, int n_threads);
int uvchost_bam_next_core(uvchost_bam *b, uvchost_core *c);
/* Positions the sequential reader at the first record that can overlap [beg, end) of tid (sam_itr_queryi); records are then pulled with
 * uvchost_bam_next_core and filtered by the caller. Returns 0, or 1 if the index has no data for the region. */
int uvchost_bam_seek_region(uvchost_bam *b, int32_t tid, int64_t beg);
/* Rough number of records that overlap [beg, end), from the linear index alone (compressed bytes per reference position of the surrounding
 * 16 kbp windows): no record is read. Used to size GPU batches before (or without) counting; errs on the high side. */
int64_t uvchost_bam_estimate_reads(const uvchost_bam *b, int32_t tid, int64_t beg, int64_t end);
/* Number of records of tid that overlap [beg, end) (the count SamIter::iternext takes per BED line, grouping.cpp:178-195). */
int64_t uvchost_bam_count(uvchost_bam *b, int32_t tid, int64_t beg, int64_t end);

/* Statistics of the first max_records records, as CommandLineArgs::selfUpdateByPlatform gathers them (CmdLineArgs.cpp:36-108). */
typedef struct uvchost_infer_stats {
    int64_t count_pe, count_se, q20_n_fail_bases, q30_n_fail_bases, q30_n_pass_bases;
    int32_t max_mapq, median_qlen, max_qlen;   /* median/max over {150} + the records' l_qseq */
} uvchost_infer_stats;
int uvchost_bam_infer(uvchost_bam *b, int64_t max_records, uvchost_infer_stats *out);

uvchost_fasta *uvchost_fasta_open(const char *path);      /* needs <path>.fai */
void uvchost_fasta_close(uvchost_fasta *f);
/* Returns a malloc'ed buffer with the whole sequence of the named contig (caller frees), length in *len; NULL if absent. */
char *uvchost_fasta_fetch_contig(uvchost_fasta *f, const char *name, int64_t *len);

#ifdef __cplusplus
}
#endif
#endif
