// shim_selfcheck.cpp - TEST INFRASTRUCTURE. Prints what the oracle's htslib shim (oracle/htslib_compat) decodes from a BAM, through the
// very calls the reference makes (sam_open, sam_hdr_read, sam_read1 / sam_index_load + sam_itr_queryi + sam_itr_next, bam_endpos,
// bam_aux_get + bam_aux2i, faidx), so that a test can compare it with an independent decoder: parity at the htslib boundary is otherwise
// unpinned.
// usage: shim_selfcheck <bam> all | <bam> region <tid> <beg> <end> | <fasta> fetch <name> <beg> <end>
#include "htslib/faidx.h"
#include "htslib/sam.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static void print_record(const bam1_t *b) {
    const uint8_t *nm = bam_aux_get(b, "NM");
    printf("%d\t%lld\t%lld\t%d\t%d\t%lld\t%lld\t%d\t%s\t%lld\t", (int)b->core.tid, (long long)b->core.pos, (long long)bam_endpos(b), (int)b->core.flag, (int)b->core.qual,
           (long long)b->core.mpos, (long long)b->core.isize, (int)b->core.mtid, bam_get_qname(b), (long long)(nm ? bam_aux2i(nm) : -1));
    const uint32_t *cigar = bam_get_cigar(b);
    for (uint32_t k = 0; k < b->core.n_cigar; k++) { printf("%u%c", bam_cigar_oplen(cigar[k]), "MIDNSHP=XB"[bam_cigar_op(cigar[k])]); }
    printf("\t");
    for (int32_t i = 0; i < b->core.l_qseq; i++) { putchar(seq_nt16_str[bam_seqi(bam_get_seq(b), i)]); }
    printf("\t");
    for (int32_t i = 0; i < b->core.l_qseq; i++) { putchar(33 + bam_get_qual(b)[i]); }
    printf("\n");
}

int main(int argc, char **argv) {
    if (argc >= 6 && 0 == strcmp(argv[2], "fetch")) {
        faidx_t *fai = fai_load(argv[1]);
        if (NULL == fai) { return 3; }
        int len = 0;
        char *s = faidx_fetch_seq(fai, argv[3], atoi(argv[4]), atoi(argv[5]), &len);
        printf("%d\t%s\n", len, s ? s : "");
        free(s);
        fai_destroy(fai);
        return 0;
    }
    if (argc < 3) { return 2; }
    samFile *fp = sam_open(argv[1], "r");
    if (NULL == fp) { return 3; }
    sam_hdr_t *h = sam_hdr_read(fp);
    if (NULL == h) { return 4; }
    for (int i = 0; i < h->n_targets; i++) { printf("@\t%s\t%u\n", h->target_name[i], (unsigned)h->target_len[i]); }
    bam1_t *b = bam_init1();
    if (0 == strcmp(argv[2], "all")) {
        while (sam_read1(fp, h, b) >= 0) { print_record(b); }
    } else if (argc >= 6) {
        hts_idx_t *idx = sam_index_load(fp, argv[1]);
        if (NULL == idx) { return 5; }
        hts_itr_t *it = sam_itr_queryi(idx, atoi(argv[3]), atoll(argv[4]), atoll(argv[5]));
        while (it && sam_itr_next(fp, it, b) >= 0) { print_record(b); }
        if (it) { hts_itr_destroy(it); }
        hts_idx_destroy(idx);
    }
    bam_destroy1(b);
    bam_hdr_destroy(h);
    sam_close(fp);
    return 0;
}
