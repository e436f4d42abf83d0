/*
 * oracle/plp.h -- TEST INFRASTRUCTURE ONLY (CPU oracle).
 * Scalar restatement of htslib 1.23.1's pileup iterators (sam.c: bam_plp_*,
 * bam_mplp_*, resolve_cigar2, overlap_push/tweak_overlap_quality,
 * bam_plp_insertion_mod), BAQ (realn.c sam_prob_realn/sam_cap_mapq,
 * probaln.c probaln_glocal) and errmod (errmod.c).  See SURVEY.md section 8a rows
 * a2-a9, a17 and Appendix A.  htslib's source is absent from /root/reference;
 * behaviour is pinned by the reference's golden outputs (tests/).
 */
#ifndef ORACLE_PLP_H
#define ORACLE_PLP_H
#include "hl.h"

/* one (read, column) entry: the role of bam_pileup1_t */
typedef struct {
    rec_t *b;
    int32_t qpos;
    int indel;
    unsigned is_del:1, is_head:1, is_tail:1, is_refskip:1;
    int cigar_ind;
} pile1_t;

typedef int (*plp_pull_f)(void *data, rec_t *b);  /* >=0 ok, -1 EOF, < -1 error */

typedef struct plp_t plp_t;
plp_t *plp_init(plp_pull_f func, void *data);
void plp_destroy(plp_t *it);
void plp_set_maxcnt(plp_t *it, int maxcnt);
void plp_init_overlaps(plp_t *it);
int plp_push(plp_t *it, const rec_t *b);     /* b == NULL marks EOF */
const pile1_t *plp_next(plp_t *it, int *tid, hpos_t *pos, int *n_plp);
const pile1_t *plp_auto(plp_t *it, int *tid, hpos_t *pos, int *n_plp);

typedef struct mplp_t mplp_t;
mplp_t *mplp_init(int n, plp_pull_f func, void **data);
void mplp_destroy(mplp_t *it);
void mplp_set_maxcnt(mplp_t *it, int maxcnt);
void mplp_init_overlaps(mplp_t *it);
int mplp_auto(mplp_t *it, int *tid, hpos_t *pos, int *n_plp, const pile1_t **plp);

/* insertion text after a column (bam_plp_insertion_mod without base mods) */
int plp_insertion(const pile1_t *p, str_t *ins, int *del_len);

/* BAQ / mapq cap (realn.c) */
int baq_realn(rec_t *b, const char *ref, hpos_t ref_len, int flag);
int cap_mapq(const rec_t *b, const char *ref, hpos_t ref_len, int thres);

/* errmod (errmod.c) + glfgen (bam2bcf.c:65-123) */
typedef struct errmod_t errmod_t;
errmod_t *errmod_new(double depcorr);
void errmod_free(errmod_t *em);
int errmod_calc(const errmod_t *em, int n, int m, uint16_t *bases, float *q);
int glfgen(int n, const pile1_t *pl, int ref_base4, int min_baseQ, int capQ,
           const errmod_t *em, float qsum[4], float p[25]);
void hts_srand48_(long seed);
double hts_drand48_(void);

#endif
