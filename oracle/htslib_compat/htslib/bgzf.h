/* htslib-compat shim (test infrastructure): BGZF subset. */
#ifndef HTSLIB_COMPAT_BGZF_H
#define HTSLIB_COMPAT_BGZF_H
#include <stdint.h>
#include <stdio.h>
#include <sys/types.h>
#ifdef __cplusplus
extern "C" {
#endif
#define BGZF_BLOCK_SIZE     0xff00
#define BGZF_MAX_BLOCK_SIZE 0x10000

typedef struct BGZF {
    FILE *fp;
    int is_write;
    int compress_level;
    /* reader state */
    uint8_t *ublock;        /* uncompressed data of the current block */
    int ublock_len;         /* its length */
    int ublock_off;         /* read cursor inside it */
    int64_t block_address;  /* compressed file offset of the current block */
    int block_clen;         /* compressed length of the current block */
    /* writer state */
    uint8_t *wbuf;
    int wbuf_len;
} BGZF;

BGZF *bgzf_open(const char *path, const char *mode);
int bgzf_close(BGZF *fp);
ssize_t bgzf_write(BGZF *fp, const void *data, size_t length);
ssize_t bgzf_raw_write(BGZF *fp, const void *data, size_t length);
int bgzf_flush(BGZF *fp);
int bgzf_compress(void *dst, size_t *dlen, const void *src, size_t slen, int level);
ssize_t bgzf_read(BGZF *fp, void *data, size_t length);
int64_t bgzf_tell_compat(BGZF *fp);
int bgzf_seek_compat(BGZF *fp, int64_t voffset);
#ifdef __cplusplus
}
#endif
#endif
