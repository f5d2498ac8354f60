// tiler.h - tier-3 region list of the uvc1 host.
//
// Restates the reference's SamIter (grouping.hpp:26-117, grouping.cpp:28-67 memory model, :69-155 region parsing, :157-314 iternext):
// one sequential pass over the BAM cuts the genome into tiles ("BedLines", iohts.hpp:14-35) on contig change (flag 16), on a gap of more
// than 200 uncovered bases (8), when the sub-memory model overflows (4) and at end of file (2); with -R / --targets the tiles are the
// given intervals. Every call of next() returns the tiles of one tier-1 iteration exactly as the reference's iternext() does for the same
// -t / --mem-per-thread, including its read-order-dependent state (the read that triggers a cut at a tier-1 boundary is dropped from the
// running end, grouping.cpp:296-300), so that the tile list - and therefore every tile-dependent counter - is bit-identical.
#ifndef UVC_TILER_H_INCLUDED
#define UVC_TILER_H_INCLUDED

#include "bam_reader.h"

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct uvchost_bedline {
    int32_t tid, beg_pos, end_pos;
    uint32_t region_flag;
    int64_t n_reads;
} uvchost_bedline;

typedef struct uvchost_tiler uvchost_tiler;

/* bed_fname / targets may be NULL or "" (not provided). nthreads and mem_per_thread_mb are the reference's -t and --mem-per-thread:
 * they only enter the memory model that decides where tier-1 iterations (and, through the sub-model, tiles) are cut.
 * bed_in_avg_sequencing_DP = -1 counts the reads of each BED line with the index (the reference's default). */
uvchost_tiler *uvchost_tiler_open(const char *bam_path, const char *bed_fname, const char *targets, int32_t nthreads, int64_t mem_per_thread_mb,
        int64_t bed_in_avg_sequencing_DP, int32_t is_fastq_gen);
void uvchost_tiler_close(uvchost_tiler *t);
const char *uvchost_tiler_error(const uvchost_tiler *t);   /* non-empty after a failed open/next */

/* Optional: cb(line, user) is called for every tile the moment it is cut, before uvchost_tiler_next returns the whole iteration, so that the
 * caller can start working on the first tiles while the scan continues. scan_threads > 1 inflates the BAM ahead on that many threads. */
typedef void (*uvchost_tile_cb)(const uvchost_bedline *line, void *user);
void uvchost_tiler_set_callback(uvchost_tiler *t, uvchost_tile_cb cb, void *user);
void uvchost_tiler_set_scan_threads(uvchost_tiler *t, int32_t scan_threads);

/* Given intervals (-R / --targets): the tiles are the intervals themselves, known before any record is read. Hands ALL of them to the callback
 * at once, with read counts estimated from the index, and disables the callback; the tier-1 iterations (uvchost_tiler_next), which need the
 * exact per-interval read counts of a whole-file pass and only decide how --bed-out-fname groups the tiles, can then run while the GPU works -
 * or not at all. Returns the number of tiles handed out (0: no given intervals). */
int64_t uvchost_tiler_emit_given(uvchost_tiler *t);

/* One tier-1 iteration. Returns the iteration's total number of reads (the reference's iternext return value) or a negative error;
 * *lines / *n_lines point into the tiler and stay valid until the next call. */
int64_t uvchost_tiler_next(uvchost_tiler *t, const uvchost_bedline **lines, int64_t *n_lines);

#ifdef __cplusplus
}
#endif
#endif
