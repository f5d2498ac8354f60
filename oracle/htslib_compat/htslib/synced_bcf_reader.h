/* htslib-compat shim (test infrastructure): synced BCF reader declarations (stubs). */
#ifndef HTSLIB_COMPAT_SYNCED_BCF_READER_H
#define HTSLIB_COMPAT_SYNCED_BCF_READER_H
#include "hts.h"
#include "vcf.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef enum bcf_sr_opt_t { BCF_SR_REQUIRE_IDX, BCF_SR_PAIR_LOGIC } bcf_sr_opt_t;
typedef struct bcf_srs_t { int nreaders; void *priv; } bcf_srs_t;
bcf_srs_t *bcf_sr_init(void);
void bcf_sr_destroy(bcf_srs_t *readers);
int bcf_sr_set_opt(bcf_srs_t *readers, bcf_sr_opt_t opt, ...);
int bcf_sr_set_regions(bcf_srs_t *readers, const char *regions, int is_file);
int bcf_sr_set_targets(bcf_srs_t *readers, const char *targets, int is_file, int alleles);
int bcf_sr_add_reader(bcf_srs_t *readers, const char *fname);
int bcf_sr_next_line(bcf_srs_t *readers);
bcf1_t *bcf_sr_get_line(bcf_srs_t *readers, int i);
#ifdef __cplusplus
}
#endif
#endif
