/*
 * oracle/hl.h -- TEST INFRASTRUCTURE ONLY (CPU oracle).
 *
 * Minimal record / header / FASTA / BED substrate for the CPU oracle.  The
 * oracle restates, in plain scalar C, the algorithm of the reference hot path
 * (samtools 1.23.1 mpileup/depth/coverage + the htslib 1.23.1 pileup engine,
 * BAQ and errmod).  htslib is NOT present under /root/reference, so the
 * engine pieces are restated from the published algorithm and pinned by the
 * reference's golden outputs (test/mpileup/expected/ etc, see tests/).
 *
 * Nothing in the product (samtools_b200/) may include, link or execute
 * anything under oracle/.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py use it, as the checker.
 */
#ifndef ORACLE_HL_H
#define ORACLE_HL_H

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef int64_t hpos_t;
#define HPOS_MAX ((((int64_t)INT32_MAX)<<32)|UINT32_MAX)

/* SAM flag bits (SAM spec section 1.4) */
#define F_PAIRED 1
#define F_PROPER 2
#define F_UNMAP 4
#define F_MUNMAP 8
#define F_REVERSE 16
#define F_MREVERSE 32
#define F_READ1 64
#define F_READ2 128
#define F_SECONDARY 256
#define F_QCFAIL 512
#define F_DUP 1024
#define F_SUPP 2048

/* CIGAR ops, BAM encoding: MIDNSHP=XB */
enum { C_M = 0, C_I, C_D, C_N, C_S, C_H, C_P, C_EQ, C_X, C_B };
#define cop(c) ((c) & 0xf)
#define cln(c) ((c) >> 4)

/* one alignment record (the role of htslib's bam1_t) */
typedef struct rec_t {
    hpos_t pos;          /* 0-based */
    int32_t tid;
    uint16_t flag;
    uint8_t mapq;
    uint32_t n_cigar;
    int32_t l_qseq;
    int32_t mtid;
    hpos_t mpos, isize;
    char *qname;
    uint32_t *cigar;
    uint8_t *seq;        /* 4-bit packed, high nibble first */
    uint8_t *qual;       /* l_qseq bytes, 0xff.. when absent */
    uint8_t *aux;        /* BAM-encoded aux block */
    int l_aux;
} rec_t;

#define seqi(s, i) ((s)[(i) >> 1] >> ((~(i) & 1) << 2) & 0xf)

void rec_init(rec_t *r);
void rec_free(rec_t *r);
void rec_copy(rec_t *dst, const rec_t *src);
hpos_t rec_rlen(const rec_t *r);    /* reference length of CIGAR (bam_cigar2rlen) */
hpos_t rec_qlen(const rec_t *r);    /* query length of CIGAR (bam_cigar2qlen) */
hpos_t rec_endpos(const rec_t *r);  /* bam_endpos */
const uint8_t *rec_aux_get(const rec_t *r, const char tag[2]); /* points at type byte */
int rec_aux_del(rec_t *r, const uint8_t *s); /* s = pointer at type byte */
void rec_aux_append(rec_t *r, const char tag[2], char type, int len, const uint8_t *data);

typedef struct {
    int n_ref;
    char **name;
    hpos_t *len;
    char *text;          /* full header text (for @RG parsing) */
} hdr_t;
int hdr_name2tid(const hdr_t *h, const char *name);
void hdr_free(hdr_t *h);

typedef struct reader_t reader_t;
/* fai_for_headerless: optional .fai path giving contigs for headerless SAM */
reader_t *reader_open(const char *fn, const char *fai_for_headerless);
hdr_t *reader_hdr(reader_t *rd);
/* region restriction (the role of sam_itr_querys/sam_itr_next): linear scan,
 * result-identical to an index lookup */
int reader_set_region(reader_t *rd, const char *reg, int *tid, hpos_t *beg, hpos_t *end);
int reader_next(reader_t *rd, rec_t *r);  /* 0 ok, -1 EOF, < -1 error */
void reader_close(reader_t *rd);

/* FASTA (the role of fai_load + faidx_fetch_seq64 of a whole contig) */
typedef struct { int n; char **name; char **seq; hpos_t *len; } fasta_t;
fasta_t *fasta_load(const char *fn);
int fasta_find(const fasta_t *fa, const char *name);
void fasta_free(fasta_t *fa);

/* BED / position list (bedidx.c semantics) */
typedef struct bed_t bed_t;
bed_t *bed_load(const char *fn);
int bed_hit(const bed_t *b, const char *chr, hpos_t beg, hpos_t end);
void bed_free(bed_t *b);

int parse_flag(const char *s);  /* bam_str2flag */
int parse_region(const hdr_t *h, const char *reg, int *tid, hpos_t *beg, hpos_t *end);

extern const unsigned char nt16_table[256];  /* IUPAC char -> 4-bit */
extern const char nt16_str[];                /* "=ACMGRSVTWYHKDBN" */
extern const int nt16_int[];                 /* 4-bit -> 0..3, 4 for ambiguous */

/* growable output string */
typedef struct { size_t l, m; char *s; } str_t;
void s_putc(str_t *s, int c);
void s_puts(str_t *s, const char *p);
void s_putn(str_t *s, const char *p, size_t n);
void s_putll(str_t *s, long long v);

#endif
