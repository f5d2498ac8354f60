/* htslib-compat shim (test infrastructure): VCF/BCF declarations. Only the
 * --tumor-vcf path of the reference uses these; the shim implements them as
 * loud failures (see htslib_compat.cpp), which is enough for tumor-only runs. */
#ifndef HTSLIB_COMPAT_VCF_H
#define HTSLIB_COMPAT_VCF_H
#include "hts.h"
#ifdef __cplusplus
extern "C" {
#endif
#define BCF_DT_ID     0
#define BCF_DT_CTG    1
#define BCF_DT_SAMPLE 2
#define BCF_HT_INT  1
#define BCF_HT_REAL 2
#define BCF_HT_STR  3
#define BCF_UN_ALL 15

typedef struct bcf_hdr_t {
    int32_t n[3];
    void *id[3];
    void *dict[3];
    char **samples;
} bcf_hdr_t;

typedef struct bcf_dec_t {
    int m_fmt, m_info, m_id, m_als, m_allele, m_flt;
    int n_flt;
    int *flt;
    char *id, *als;
    char **allele;
} bcf_dec_t;

typedef struct bcf1_t {
    hts_pos_t pos;
    hts_pos_t rlen;
    int32_t rid;
    float qual;
    uint32_t n_info:16, n_allele:16;
    uint32_t n_fmt:8, n_sample:24;
    kstring_t shared, indiv;
    bcf_dec_t d;
    int max_unpack;
    int unpacked;
    int unpack_size[3];
    int errcode;
} bcf1_t;

#define bcf_hdr_nsamples(hdr) (hdr)->n[BCF_DT_SAMPLE]
#define bcf_close(fp) hts_close(fp)

bcf_hdr_t *bcf_hdr_read(htsFile *fp);
void bcf_hdr_destroy(bcf_hdr_t *h);
int bcf_unpack(bcf1_t *b, int which);
bcf1_t *bcf_dup(bcf1_t *src);
void bcf_destroy(bcf1_t *v);
int bcf_get_format_values(const bcf_hdr_t *hdr, bcf1_t *line, const char *tag, void **dst, int *ndst, int type);
#define bcf_get_format_int32(hdr,line,tag,dst,ndst) bcf_get_format_values(hdr,line,tag,(void**)(dst),ndst,BCF_HT_INT)
#define bcf_get_format_float(hdr,line,tag,dst,ndst) bcf_get_format_values(hdr,line,tag,(void**)(dst),ndst,BCF_HT_REAL)
#define bcf_get_format_char(hdr,line,tag,dst,ndst)  bcf_get_format_values(hdr,line,tag,(void**)(dst),ndst,BCF_HT_STR)
int vcf_format(const bcf_hdr_t *h, const bcf1_t *v, kstring_t *s);
#ifdef __cplusplus
}
#endif
#endif
