/* htslib-compat shim (TEST INFRASTRUCTURE, not product code).
 * Minimal re-declaration of the htslib-1.11 API surface that the unmodified
 * reference (genetronhealth/uvc 0.15.1) needs, written from the public htslib
 * interface/BAM-SAM spec so that the reference can be built here without the
 * un-vendored htslib dependency (reference Makefile:16-17). Only behaviour the
 * reference's hot path exercises is implemented (see htslib_compat.cpp). */
#ifndef HTSLIB_COMPAT_HTS_H
#define HTSLIB_COMPAT_HTS_H
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef int64_t hts_pos_t;
#define HTS_POS_MAX ((((int64_t)INT32_MAX)<<32)|INT32_MAX)

typedef struct kstring_t { size_t l, m; char *s; } kstring_t;

struct BGZF;
typedef struct htsFile {
    struct BGZF *bgzf;
    char *fn;
    int is_write;
} htsFile;

typedef struct hts_idx_t hts_idx_t;

typedef struct hts_itr_t {
    int tid;
    hts_pos_t beg, end;
    int64_t start_voffset; /* -1: nothing to read */
    int positioned;
    int finished;
} hts_itr_t;

htsFile *hts_open(const char *fn, const char *mode);
int hts_close(htsFile *fp);
void hts_idx_destroy(hts_idx_t *idx);
void hts_itr_destroy(hts_itr_t *iter);

extern const unsigned char seq_nt16_table[256];
extern const char seq_nt16_str[];
extern const int seq_nt16_int[];

#ifdef __cplusplus
}
#endif
#endif
