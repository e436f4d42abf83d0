/*
 * oracle/plp.c -- TEST INFRASTRUCTURE ONLY (CPU oracle).
 *
 * Sequential pileup iterator: restates htslib 1.23.1 sam.c bam_plp_push /
 * bam_plp64_next / bam_plp64_auto / bam_mplp64_auto, the CIGAR cursor
 * resolve_cigar2, the mate-overlap machinery (overlap_push,
 * tweak_overlap_quality, overlap_remove) and bam_plp_insertion_mod.
 * Call sites in the reference: bam_plcmd.c:581-607, coverage.c:572-589,
 * bam_plbuf.c:47-66.  Semantics: SURVEY.md Appendix A1-A5.
 */
#include "plp.h"
#include <assert.h>

typedef struct node_t {
    rec_t b;
    hpos_t beg, end;          /* [beg,end): raw CIGAR reference span */
    int k; hpos_t x; int y;   /* CIGAR cursor: op index, ref start, query start */
    struct node_t *next;
} node_t;

typedef struct ovl_ent { char *key; node_t *val; struct ovl_ent *next; } ovl_ent;
#define OVL_NB 4096

struct plp_t {
    node_t *head, *tail;
    int live;                 /* linked nodes; mempool count = live + 1 (sentinel) */
    int tid, max_tid; hpos_t pos, max_pos;
    int is_eof, error, maxcnt;
    pile1_t *plp; int max_plp;
    plp_pull_f func; void *data;
    rec_t b;
    ovl_ent **ovl;            /* NULL unless overlap detection enabled */
};

static uint32_t x31(const char *s)
{
    uint32_t h = (uint32_t)*s;
    if (h) for (++s; *s; ++s) h = (h << 5) - h + (uint32_t)*s;
    return h;
}
static uint32_t wang(uint32_t key)
{
    key += ~(key << 15); key ^= (key >> 10); key += (key << 3);
    key ^= (key >> 6); key += ~(key << 11); key ^= (key >> 16);
    return key;
}

static node_t *node_new(void) { node_t *n = calloc(1, sizeof(*n)); rec_init(&n->b); return n; }
static void node_free(node_t *n) { rec_free(&n->b); free(n); }

plp_t *plp_init(plp_pull_f func, void *data)
{
    plp_t *it = calloc(1, sizeof(*it));
    it->head = it->tail = node_new();
    it->max_tid = -1; it->max_pos = -1;
    it->maxcnt = 8000;
    it->func = func; it->data = data;
    rec_init(&it->b);
    return it;
}
void plp_destroy(plp_t *it)
{
    node_t *p = it->head;
    while (p) { node_t *q = p->next; node_free(p); p = q; }
    if (it->ovl) {
        int i;
        for (i = 0; i < OVL_NB; i++) { ovl_ent *e = it->ovl[i]; while (e) { ovl_ent *q = e->next; free(e->key); free(e); e = q; } }
        free(it->ovl);
    }
    rec_free(&it->b);
    free(it->plp); free(it);
}
void plp_set_maxcnt(plp_t *it, int m) { it->maxcnt = m; }
void plp_init_overlaps(plp_t *it) { if (!it->ovl) it->ovl = calloc(OVL_NB, sizeof(ovl_ent *)); }

static ovl_ent **ovl_find(plp_t *it, const char *key)
{
    ovl_ent **pp = &it->ovl[x31(key) % OVL_NB];
    while (*pp && strcmp((*pp)->key, key)) pp = &(*pp)->next;
    return pp;
}
static void ovl_remove(plp_t *it, const rec_t *b)
{
    if (!it->ovl) return;
    ovl_ent **pp = ovl_find(it, b->qname);
    if (*pp) { ovl_ent *e = *pp; *pp = e->next; free(e->key); free(e); }
}

/* --- walking two CIGARs in lock-step on reference coordinates (A5) --- */
typedef struct { const uint32_t *c, *cmax; hpos_t icig, iseq, iref; } cwalk_t;

/* position the walker on reference offset *iref (relative to read start) */
static int cw_set(cwalk_t *w)
{
    hpos_t pos = w->iref;
    if (pos < 0) return -1;
    w->icig = 0; w->iseq = 0; w->iref = 0;
    while (w->c < w->cmax) {
        int op = cop(*w->c); int n = (int)cln(*w->c);
        if (op == C_S) { w->c++; w->iseq += n; w->icig = 0; continue; }
        if (op == C_H || op == C_P) { w->c++; w->icig = 0; continue; }
        if (op == C_M || op == C_EQ || op == C_X) {
            pos -= n;
            if (pos < 0) { w->icig = n + pos; w->iseq += w->icig; w->iref += w->icig; return C_M; }
            w->c++; w->iseq += n; w->icig = 0; w->iref += n;
            continue;
        }
        if (op == C_I) { w->c++; w->iseq += n; w->icig = 0; continue; }
        if (op == C_D || op == C_N) {
            pos -= n; if (pos < 0) pos = 0;
            w->c++; w->iref += n;
            continue;
        }
        return -2;
    }
    w->iseq = -1;
    return -1;
}
static int cw_next(cwalk_t *w)
{
    while (w->c < w->cmax) {
        int op = cop(*w->c); int n = (int)cln(*w->c);
        if (op == C_M || op == C_EQ || op == C_X) {
            if (w->icig >= n - 1) { w->icig = -1; w->c++; continue; }
            w->iseq++; w->icig++; w->iref++;
            return C_M;
        }
        if (op == C_D || op == C_N) { w->c++; w->iref += n; w->icig = -1; continue; }
        if (op == C_I) { w->c++; w->iseq += n; w->icig = -1; continue; }
        if (op == C_S) { w->c++; w->iseq += n; w->icig = -1; continue; }
        if (op == C_H || op == C_P) { w->c++; w->icig = -1; continue; }
        return -2;
    }
    w->iseq = -1; w->iref = -1;
    return -1;
}

/* a = mate seen first, b = mate arriving now; rewrites qualities of both */
static int tweak_overlap(rec_t *a, rec_t *b)
{
    cwalk_t A = { a->cigar, a->cigar + a->n_cigar, 0, 0, 0 };
    cwalk_t B = { b->cigar, b->cigar + b->n_cigar, 0, 0, 0 };
    hpos_t iref = b->pos;
    A.iref = iref - a->pos; B.iref = iref - b->pos;
    int ra = cw_set(&A);
    if (ra < 0) return ra < -1 ? -1 : 0;
    int rb = cw_set(&B);
    if (rb < 0) return rb < -1 ? -1 : 0;

    /* which mate keeps the evidence on ties / agreement: hash of the name */
    int amul, bmul;
    if (wang(x31(a->qname)) & 1) { amul = 1; bmul = 0; } else { amul = 0; bmul = 1; }

    int err = 0;
    for (;;) {
        while (ra >= 0 && A.iref >= 0 && A.iref < iref - a->pos) ra = cw_next(&A);
        if (ra < 0) { err = ra < -1 ? -1 : 0; break; }
        while (rb >= 0 && B.iref >= 0 && B.iref < iref - b->pos) rb = cw_next(&B);
        if (rb < 0) { err = rb < -1 ? -1 : 0; break; }
        if (iref < A.iref + a->pos) iref = A.iref + a->pos;
        if (iref < B.iref + b->pos) iref = B.iref + b->pos;
        iref++;

        if (A.iref + a->pos != B.iref + b->pos) {
            /* one mate sits after a deletion: catch the other one up, treating
             * the bases opposite the deletion like mismatches */
            if (A.iref + a->pos < B.iref + b->pos && B.c > b->cigar && cop(B.c[-1]) == C_D) {
                do {
                    a->qual[A.iseq] = amul ? (uint8_t)(a->qual[A.iseq] * 0.8) : 0;
                    ra = cw_next(&A);
                    if (ra < 0) return -(ra < -1);
                } while (A.iref + a->pos < B.iref + b->pos);
            } else if (A.c > a->cigar && cop(A.c[-1]) == C_D) {
                do {
                    b->qual[B.iseq] = bmul ? (uint8_t)(b->qual[B.iseq] * 0.8) : 0;
                    rb = cw_next(&B);
                    if (rb < 0) return -(rb < -1);
                } while (B.iref + b->pos < A.iref + a->pos);
            } else continue; /* e.g. ref-skip: unsupported */
        }
        if (A.iseq > a->l_qseq || B.iseq > b->l_qseq) return -1;

        if (seqi(a->seq, A.iseq) == seqi(b->seq, B.iseq)) {
            int q = a->qual[A.iseq] + b->qual[B.iseq];
            if (q > 200) q = 200;
            a->qual[A.iseq] = (uint8_t)(amul * q);
            b->qual[B.iseq] = (uint8_t)(bmul * q);
        } else if (a->qual[A.iseq] > b->qual[B.iseq]) {
            a->qual[A.iseq] = (uint8_t)(0.8 * a->qual[A.iseq]);
            b->qual[B.iseq] = 0;
        } else if (a->qual[A.iseq] < b->qual[B.iseq]) {
            b->qual[B.iseq] = (uint8_t)(0.8 * b->qual[B.iseq]);
            a->qual[A.iseq] = 0;
        } else {
            a->qual[A.iseq] = (uint8_t)(amul * 0.8 * a->qual[A.iseq]);
            b->qual[B.iseq] = (uint8_t)(bmul * 0.8 * b->qual[B.iseq]);
        }
    }
    return err;
}

static int ovl_push(plp_t *it, node_t *nd)
{
    if (!it->ovl) return 0;
    rec_t *b = &nd->b;
    if ((b->flag & F_MUNMAP) || !(b->flag & F_PROPER)) return 0;
    if ((b->mtid >= 0 && b->tid != b->mtid) ||
        (llabs(b->isize) >= 2 * (long long)b->l_qseq && b->mpos >= nd->end)) return 0;
    ovl_ent **pp = ovl_find(it, b->qname);
    if (!*pp) {
        if (b->mpos >= b->pos || ((b->flag & F_PAIRED) && b->mpos == -1)) {
            ovl_ent *e = calloc(1, sizeof(*e));
            e->key = strdup(b->qname); e->val = nd;
            *pp = e;
        }
        return 0;
    }
    ovl_ent *e = *pp;
    int err = tweak_overlap(&e->val->b, b);
    *pp = e->next; free(e->key); free(e);
    return err;
}

int plp_push(plp_t *it, const rec_t *b)
{
    if (it->error) return -1;
    if (!b) { it->is_eof = 1; return 0; }
    if (b->tid < 0 || (b->flag & F_UNMAP)) { ovl_remove(it, b); return 0; }
    if (it->tid == b->tid && it->pos == b->pos && it->live + 1 > it->maxcnt) { ovl_remove(it, b); return 0; }
    node_t *t = it->tail;
    rec_copy(&t->b, b);
    t->beg = b->pos;
    t->end = b->pos + rec_rlen(b);
    t->k = -1; t->x = 0; t->y = 0;
    if (b->tid < it->max_tid || (b->tid == it->max_tid && t->beg < it->max_pos)) {
        fprintf(stderr, "[oracle] The input is not sorted\n");
        it->error = 1;
        return -1;
    }
    it->max_tid = b->tid; it->max_pos = t->beg;
    if (t->end > it->pos || t->b.tid > it->tid) {
        if (ovl_push(it, t) < 0) { it->error = 1; return -1; }
        t->next = node_new();
        it->tail = t->next;
        it->live++;
    }
    return 0;
}

/* CIGAR cursor for column pos (resolve_cigar2; Appendix A2) */
static void resolve(pile1_t *p, hpos_t pos, node_t *s)
{
    rec_t *b = p->b;
    const uint32_t *cg = b->cigar;
    int n = (int)b->n_cigar, k;
    if (s->k == -1) {
        s->x = b->pos; s->y = 0;
        for (k = 0; k < n; k++) {
            int op = cop(cg[k]);
            if (op == C_M || op == C_D || op == C_N || op == C_EQ || op == C_X) break;
            if (op == C_I || op == C_S) s->y += (int)cln(cg[k]);
        }
        assert(k < n);
        s->k = k;
    } else {
        int l = (int)cln(cg[s->k]);
        if (pos - s->x >= l) {
            int op = cop(cg[s->k]);
            if (op == C_M || op == C_EQ || op == C_X) s->y += l;
            s->x += l;
            for (k = s->k + 1; k < n; k++) {
                op = cop(cg[k]);
                if (op == C_M || op == C_D || op == C_N || op == C_EQ || op == C_X) break;
                if (op == C_I || op == C_S) s->y += (int)cln(cg[k]);
            }
            assert(k < n);
            s->k = k;
        }
    }
    int op = cop(cg[s->k]), l = (int)cln(cg[s->k]);
    p->is_del = p->is_refskip = 0; p->indel = 0;
    if (s->x + l - 1 == pos && s->k + 1 < n) {
        int op2 = cop(cg[s->k + 1]), l2 = (int)cln(cg[s->k + 1]);
        if (op2 == C_D && op != C_D) {
            p->indel = -l2;
            for (k = s->k + 2; k < n; k++) { if (cop(cg[k]) == C_D) p->indel -= (int)cln(cg[k]); else break; }
        } else if (op2 == C_I) {
            p->indel = l2;
            for (k = s->k + 2; k < n; k++) {
                int o = cop(cg[k]);
                if (o == C_I) p->indel += (int)cln(cg[k]);
                else if (o != C_P) break;
            }
        } else if (op2 == C_P && s->k + 2 < n) {
            int l3 = 0;
            for (k = s->k + 2; k < n; k++) {
                int o = cop(cg[k]);
                if (o == C_I) l3 += (int)cln(cg[k]);
                else if (o == C_D || o == C_M || o == C_N || o == C_EQ || o == C_X) break;
            }
            if (l3 > 0) p->indel = l3;
        }
    }
    if (op == C_M || op == C_EQ || op == C_X) p->qpos = s->y + (int)(pos - s->x);
    else if (op == C_D || op == C_N) { p->is_del = 1; p->qpos = s->y; p->is_refskip = (op == C_N); }
    p->is_head = (pos == b->pos);
    p->is_tail = (pos == s->end - 1);
    p->cigar_ind = s->k;
}

const pile1_t *plp_next(plp_t *it, int *_tid, hpos_t *_pos, int *_n)
{
    if (it->error) { *_n = -1; return NULL; }
    *_n = 0;
    if (it->is_eof && it->head == it->tail) return NULL;
    while (it->is_eof || it->max_tid > it->tid || (it->max_tid == it->tid && it->max_pos > it->pos)) {
        int n = 0;
        node_t **pp = &it->head;
        while (*pp != it->tail) {
            node_t *p = *pp;
            if (p->b.tid < it->tid || (p->b.tid == it->tid && p->end <= it->pos)) {
                ovl_remove(it, &p->b);
                *pp = p->next; node_free(p); it->live--;
            } else {
                if (p->b.tid == it->tid && p->beg <= it->pos) {
                    if (n == it->max_plp) {
                        it->max_plp = it->max_plp ? it->max_plp << 1 : 256;
                        it->plp = realloc(it->plp, sizeof(pile1_t) * (size_t)it->max_plp);
                    }
                    it->plp[n].b = &p->b;
                    resolve(&it->plp[n], it->pos, p);
                    n++;
                }
                pp = &p->next;
            }
        }
        *_n = n; *_tid = it->tid; *_pos = it->pos;
        if (it->head != it->tail) {
            if (it->tid > it->head->b.tid) { it->error = 1; *_n = -1; return NULL; }
            if (it->tid < it->head->b.tid) { it->tid = it->head->b.tid; it->pos = it->head->beg; }
            else if (it->pos < it->head->beg) it->pos = it->head->beg;
            else ++it->pos;
        } else ++it->pos;
        if (n) return it->plp;
        if (it->is_eof && it->head == it->tail) break;
    }
    return NULL;
}

const pile1_t *plp_auto(plp_t *it, int *_tid, hpos_t *_pos, int *_n)
{
    const pile1_t *plp;
    if (!it->func || it->error) { *_n = -1; return NULL; }
    if ((plp = plp_next(it, _tid, _pos, _n)) != NULL) return plp;
    *_n = 0;
    if (it->is_eof) return NULL;
    int ret;
    while ((ret = it->func(it->data, &it->b)) >= 0) {
        if (plp_push(it, &it->b) < 0) { *_n = -1; return NULL; }
        if ((plp = plp_next(it, _tid, _pos, _n)) != NULL) return plp;
    }
    if (ret < -1) { it->error = ret; *_n = -1; return NULL; }
    if (plp_push(it, NULL) < 0) { *_n = -1; return NULL; }
    if ((plp = plp_next(it, _tid, _pos, _n)) != NULL) return plp;
    return NULL;
}

/* ---- multi-file merge (bam_mplp64_auto; Appendix A4) ---- */
struct mplp_t {
    int n;
    uint32_t min_tid, *tid;
    uint64_t min_pos, *pos;
    plp_t **it;
    int *n_plp;
    const pile1_t **plp;
};
mplp_t *mplp_init(int n, plp_pull_f func, void **data)
{
    mplp_t *m = calloc(1, sizeof(*m));
    int i;
    m->n = n;
    m->tid = calloc((size_t)n, sizeof(uint32_t)); m->pos = calloc((size_t)n, sizeof(uint64_t));
    m->it = calloc((size_t)n, sizeof(plp_t *)); m->n_plp = calloc((size_t)n, sizeof(int));
    m->plp = calloc((size_t)n, sizeof(pile1_t *));
    m->min_pos = (uint64_t)-1; m->min_tid = (uint32_t)-1;
    for (i = 0; i < n; i++) { m->it[i] = plp_init(func, data[i]); m->pos[i] = m->min_pos; m->tid[i] = m->min_tid; }
    return m;
}
void mplp_destroy(mplp_t *m)
{
    int i;
    for (i = 0; i < m->n; i++) plp_destroy(m->it[i]);
    free(m->tid); free(m->pos); free(m->it); free(m->n_plp); free(m->plp); free(m);
}
void mplp_set_maxcnt(mplp_t *m, int c) { int i; for (i = 0; i < m->n; i++) plp_set_maxcnt(m->it[i], c); }
void mplp_init_overlaps(mplp_t *m) { int i; for (i = 0; i < m->n; i++) plp_init_overlaps(m->it[i]); }
int mplp_auto(mplp_t *m, int *_tid, hpos_t *_pos, int *n_plp, const pile1_t **plp)
{
    int i, ret = 0;
    uint64_t new_pos = (uint64_t)-1; uint32_t new_tid = (uint32_t)-1;
    for (i = 0; i < m->n; i++) {
        if (m->pos[i] == m->min_pos && m->tid[i] == m->min_tid) {
            int tid; hpos_t pos;
            m->plp[i] = plp_auto(m->it[i], &tid, &pos, &m->n_plp[i]);
            if (m->it[i]->error) return -1;
            if (m->plp[i]) { m->tid[i] = (uint32_t)tid; m->pos[i] = (uint64_t)pos; }
            else { m->tid[i] = 0; m->pos[i] = 0; }
        }
        if (m->plp[i]) {
            if (m->tid[i] < new_tid) { new_tid = m->tid[i]; new_pos = m->pos[i]; }
            else if (m->tid[i] == new_tid && m->pos[i] < new_pos) new_pos = m->pos[i];
        }
    }
    m->min_pos = new_pos; m->min_tid = new_tid;
    if (new_pos == (uint64_t)-1) return 0;
    *_tid = (int)new_tid; *_pos = (hpos_t)new_pos;
    for (i = 0; i < m->n; i++) {
        if (m->pos[i] == m->min_pos && m->tid[i] == m->min_tid) { n_plp[i] = m->n_plp[i]; plp[i] = m->plp[i]; ret++; }
        else { n_plp[i] = 0; plp[i] = NULL; }
    }
    return ret;
}

/* insertion string following column p (bam_plp_insertion_mod, no base mods; A3) */
int plp_insertion(const pile1_t *p, str_t *ins, int *del_len)
{
    ins->l = 0; if (ins->s) ins->s[0] = 0;
    if (p->indel <= 0) return 0;
    if (del_len) *del_len = 0;
    const rec_t *b = p->b;
    int k, j = 1, nb = 0;
    for (k = p->cigar_ind + 1; k < (int)b->n_cigar; k++) {
        int op = cop(b->cigar[k]), l = (int)cln(b->cigar[k]), i;
        if (op == C_P) { for (i = 0; i < l; i++) s_putc(ins, '*'); nb += l; }
        else if (op == C_I) {
            for (i = 0; i < l; i++, j++) {
                int q = p->qpos + j - (int)p->is_del;
                s_putc(ins, q < b->l_qseq ? nt16_str[seqi(b->seq, q)] : 'N');
            }
            nb += l;
        } else { if (op == C_D && del_len) *del_len = l; break; }
    }
    return nb;
}
