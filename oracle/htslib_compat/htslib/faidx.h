/* htslib-compat shim (test infrastructure): faidx subset (uncompressed FASTA + .fai). */
#ifndef HTSLIB_COMPAT_FAIDX_H
#define HTSLIB_COMPAT_FAIDX_H
#include "hts.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct faidx_t faidx_t;
faidx_t *fai_load(const char *fn);
void fai_destroy(faidx_t *fai);
const char *faidx_iseq(const faidx_t *fai, int i);
char *faidx_fetch_seq(const faidx_t *fai, const char *c_name, int p_beg_i, int p_end_i, int *len);
#ifdef __cplusplus
}
#endif
#endif
