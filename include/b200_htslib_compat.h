/*
 * b200_htslib_compat.h -- tier T1: htslib's pileup iterator API, served by the CUDA engine.
 *
 * Re-declares, with htslib 1.23's public layouts (htslib/sam.h; restated in
 * SURVEY.md section 8b because htslib is absent from the reference tree), the
 * types and entry points the reference's pileup callers use:
 *
 *   bam_plp_init / bam_plp_push / bam_plp64_next / bam_plp64_auto / bam_plp_auto /
 *   bam_plp_set_maxcnt / bam_plp_reset / bam_plp_destroy        bam_plbuf.c:40-66, cut_target.c:223-224, phase.c:699-718
 *   bam_mplp_init / bam_mplp_init_overlaps / bam_mplp_set_maxcnt /
 *   bam_mplp_auto / bam_mplp64_auto / bam_mplp_destroy            bam_plcmd.c:581-607,922; coverage.c:572-589,698; bedcov.c:303-316
 *   bam_plp_insertion                                             bam_plcmd.c:119 (via _mod), bam_tview.c:223,255
 *   bam_plp_constructor / bam_plp_destructor / bam_mplp_constructor / bam_mplp_destructor      bam_plcmd.c:582-585, bedcov.c:313-314
 *   bam_plp_next / bam_mplp_reset                                 phase.c:718, bam_plcmd.c (region restarts)
 *
 * Behaviour is htslib's: columns with n_plp > 0 in
 * (tid,pos) order, reads inside a column in push order, bam_pileup1_t fields
 * as resolve_cigar2 sets them, the max-depth rule, and -- when overlaps are
 * enabled -- the mate-overlap quality tweak applied to the iterator's own
 * copies of the reads.  Internally the iterator pulls every read of one
 * reference sequence through the callback, stages it as one batch
 * (b200_stage), obtains all (read, column) entries from the device
 * (b200_pileup_entries) and then hands the columns out one by one; `plp[i].b`
 * points at the iterator's copy of the read, valid until the next call, as in
 * htslib.  Constructor/destructor hooks: the constructor runs at bam_plp_push time for every mapped read (htslib
 * skips reads its max-depth rule drops; here that rule is evaluated later, on the device, so such reads get a
 * constructor AND a destructor call); destructors run when the reference sequence a read belongs to has been handed
 * out, and for everything still buffered at bam_plp_reset / bam_plp_destroy.
 *
 * Link with -lb200pileup.  No CPU fallback: bam_plp_init() returns NULL when
 * no CUDA device is available.
 */
#ifndef B200_HTSLIB_COMPAT_H
#define B200_HTSLIB_COMPAT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int64_t hts_pos_t;

typedef struct bam1_core_t {
    hts_pos_t pos;
    int32_t tid;
    uint16_t bin;
    uint8_t qual;
    uint8_t l_extranul;
    uint16_t flag;
    uint16_t l_qname;
    uint32_t n_cigar;
    int32_t l_qseq;
    int32_t mtid;
    hts_pos_t mpos;
    hts_pos_t isize;
} bam1_core_t;

typedef struct bam1_t {
    bam1_core_t core;
    uint64_t id;
    uint8_t *data;      /* qname (NUL padded to l_qname) | cigar u32[n_cigar] | seq 4-bit | qual | aux */
    int l_data;
    uint32_t m_data;
    uint32_t mempolicy:2, :30;
} bam1_t;

#define bam_get_qname(b) ((char*)(b)->data)
#define bam_get_cigar(b) ((uint32_t*)((b)->data + (b)->core.l_qname))
#define bam_get_seq(b)   ((b)->data + ((b)->core.n_cigar<<2) + (b)->core.l_qname)
#define bam_get_qual(b)  ((b)->data + ((b)->core.n_cigar<<2) + (b)->core.l_qname + (((b)->core.l_qseq + 1)>>1))
#define bam_seqi(s, i) ((s)[(i)>>1] >> ((~(i)&1)<<2) & 0xf)
#define bam_is_rev(b) (((b)->core.flag&16) != 0)

typedef union { void *p; int64_t i; double f; } bam_pileup_cd;

typedef struct bam_pileup1_t {
    bam1_t *b;
    int32_t qpos;
    int indel, level;
    uint32_t is_del:1, is_head:1, is_tail:1, is_refskip:1, :1, aux:27;
    bam_pileup_cd cd;
    int cigar_ind;
} bam_pileup1_t;

typedef int (*bam_plp_auto_f)(void *data, bam1_t *b);   /* >=0 ok, -1 EOF, < -1 error */

typedef struct b200_plp *bam_plp_t;
typedef struct b200_mplp *bam_mplp_t;

bam_plp_t bam_plp_init(bam_plp_auto_f func, void *data);
void bam_plp_destroy(bam_plp_t iter);
void bam_plp_reset(bam_plp_t iter);
int bam_plp_push(bam_plp_t iter, const bam1_t *b);   /* b == NULL marks end of input */
const bam_pileup1_t *bam_plp64_next(bam_plp_t iter, int *_tid, hts_pos_t *_pos, int *_n_plp);
const bam_pileup1_t *bam_plp64_auto(bam_plp_t iter, int *_tid, hts_pos_t *_pos, int *_n_plp);
const bam_pileup1_t *bam_plp_auto(bam_plp_t iter, int *_tid, int *_pos, int *_n_plp);
const bam_pileup1_t *bam_plp_next(bam_plp_t iter, int *_tid, int *_pos, int *_n_plp);
void bam_plp_set_maxcnt(bam_plp_t iter, int maxcnt);
/* per-read client data (bam_pileup1_t.cd): func(data, b, cd) with the iterator's callback data */
typedef int (*bam_plp_cd_f)(void *data, const bam1_t *b, bam_pileup_cd *cd);
void bam_plp_constructor(bam_plp_t iter, bam_plp_cd_f func);
void bam_plp_destructor(bam_plp_t iter, bam_plp_cd_f func);

bam_mplp_t bam_mplp_init(int n, bam_plp_auto_f func, void **data);
int bam_mplp_init_overlaps(bam_mplp_t iter);
void bam_mplp_destroy(bam_mplp_t iter);
void bam_mplp_set_maxcnt(bam_mplp_t iter, int maxcnt);
void bam_mplp_reset(bam_mplp_t iter);
void bam_mplp_constructor(bam_mplp_t iter, bam_plp_cd_f func);
void bam_mplp_destructor(bam_mplp_t iter, bam_plp_cd_f func);
int bam_mplp_auto(bam_mplp_t iter, int *_tid, int *_pos, int *n_plp, const bam_pileup1_t **plp);
int bam_mplp64_auto(bam_mplp_t iter, int *_tid, hts_pos_t *_pos, int *n_plp, const bam_pileup1_t **plp);

/* insertion sequence following a column with p->indel > 0 (bam_plp_insertion).  `ins` must hold at least
 * the returned length + 1 bytes (ins_cap); del_len receives the length of a deletion that follows. */
int b200_plp_insertion(const bam_pileup1_t *p, char *ins, int ins_cap, int *del_len);
/* htslib's own signature (kstring.h layout): ins->s is grown with realloc; returns the insertion length, -1 on failure */
typedef struct kstring_t { size_t l, m; char *s; } kstring_t;
int bam_plp_insertion(const bam_pileup1_t *p, kstring_t *ins, int *del_len);

/* ---- per-read and per-column entry points of the same hot path (samtools_b200/csrc/host/hts_read_ops.cpp) ----------
 * htslib realn.c:  sam_prob_realn (BAQ; call sites bam_plcmd.c:451, bam_md.c:475), sam_cap_mapq (bam_plcmd.c:453, bam_md.c:481)
 * htslib errmod.h: errmod_init / errmod_cal / errmod_destroy (bam2bcf.c:46,121; phase.c:754; cut_target.c:84)
 * bam2bcf.h:51-53: bcf_call_init / bcf_call_glfgen / bcf_call_destroy (tv_pl_func, bam_tview.c:193-205)
 * htslib sam.h:    bam_plp_insertion_mod (pileup_seq, bam_plcmd.c:119; m == NULL, i.e. without -M markup)
 * Same signatures, return codes and in-place effects as upstream; each call runs one small batch on the device. */
int sam_prob_realn(bam1_t *b, const char *ref, hts_pos_t ref_len, int flag);   /* flag: 1 apply, 2 extend, 4 redo */
int sam_cap_mapq(bam1_t *b, const char *ref, hts_pos_t ref_len, int thres);
typedef struct errmod_t errmod_t;
errmod_t *errmod_init(double depcorr);
void errmod_destroy(errmod_t *em);
int errmod_cal(const errmod_t *em, int n, int m, uint16_t *bases, float *q);
typedef struct __bcf_callaux_t {          /* bam2bcf.h:33-39 */
    int capQ, min_baseQ;
    int max_bases;
    uint16_t *bases;
    errmod_t *e;
} bcf_callaux_t;
typedef struct {                          /* bam2bcf.h:42-45 */
    float qsum[4];
    float p[25];
} bcf_callret1_t;
bcf_callaux_t *bcf_call_init(double theta, int min_baseQ);
void bcf_call_destroy(bcf_callaux_t *bca);
int bcf_call_glfgen(int _n, const bam_pileup1_t *pl, int ref_base, bcf_callaux_t *bca, bcf_callret1_t *r);
typedef struct hts_base_mod_state hts_base_mod_state;
int bam_plp_insertion_mod(const bam_pileup1_t *p, hts_base_mod_state *m, kstring_t *ins, int *del_len);

bam1_t *bam_init1(void);
void bam_destroy1(bam1_t *b);
bam1_t *bam_copy1(bam1_t *dst, const bam1_t *src);

#ifdef __cplusplus
}
#endif
#endif
