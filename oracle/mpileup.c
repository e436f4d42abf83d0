/*
 * oracle/mpileup.c -- TEST INFRASTRUCTURE ONLY (CPU oracle).
 *
 * Restates the reference's mpileup driver, bam_plcmd.c: read filter callback
 * mplp_func (:400-461), column loop (:607-868), per-read column text
 * pileup_seq (:54-169), zero-depth rows print_empty_pileup (:372-398), -a/-aa
 * gap filling (:610-660, :880-910) and option parsing (:1075-1272).
 * Base modifications (-M) and CRAM are not restated (out of scope).
 * `gl` mode prints per-column genotype likelihoods through glfgen/errmod
 * (bam2bcf.c:65-123 front end; numbers unpinned by reference tests).
 */
#include "plp.h"
#include <ctype.h>
#include <getopt.h>
#include <limits.h>
#include <errno.h>

#define FL_NO_ORPHAN 1
#define FL_REALN 2
#define FL_REDO_BAQ 4
#define FL_ILLUMINA13 8
#define FL_IGNORE_RG 16
#define FL_OVERLAPS 32
/* optional output columns, in the reference's print order (bam_plcmd.c:182-199) */
enum { X_MAPQ_CHAR, X_QPOS, X_QNAME, X_FLAG, X_RNAME, X_POS, X_MAPQ, X_RNEXT, X_PNEXT, X_RLEN, X_QPOS5, X_N };

typedef struct {
    int min_mq, flag, min_baseQ, capQ_thres, max_depth, all, rev_del;
    int rflag_require, rflag_filter;
    char *reg; const char *fa_fn, *out_fn;
    fasta_t *fa; bed_t *bed;
    char **rg_excl; int n_rg_excl;
    int xcol[X_N];
    char **tags; int n_tags;
    char sep, empty; int no_ins, no_del, no_ends;
    int gl;              /* `gl` mode */
} conf_t;

typedef struct {
    reader_t *rd; hdr_t *h; const conf_t *conf;
} aux_t;

static int get_ref(const conf_t *c, const hdr_t *h, int tid, const char **ref, hpos_t *len)
{
    *ref = NULL; *len = 0;
    if (!c->fa || tid < 0 || tid >= h->n_ref) return 0;
    int i = fasta_find(c->fa, h->name[tid]);
    if (i < 0) return 0;
    *ref = c->fa->seq[i]; *len = c->fa->len[i];
    return 1;
}

/* bam_plcmd.c:400-461 */
static int pull(void *data, rec_t *b)
{
    aux_t *ma = data;
    const conf_t *c = ma->conf;
    int ret, skip = 0;
    do {
        const char *ref; hpos_t ref_len; int has_ref;
        ret = reader_next(ma->rd, b);
        if (ret < 0) break;
        if (b->tid < 0 || (b->flag & F_UNMAP)) { skip = 1; continue; }
        if (c->rflag_require && !(c->rflag_require & b->flag)) { skip = 1; continue; }
        if (c->rflag_filter && (c->rflag_filter & b->flag)) { skip = 1; continue; }
        if (c->bed && c->all == 0) {
            skip = !bed_hit(c->bed, ma->h->name[b->tid], b->pos, rec_endpos(b));
            if (skip) continue;
        }
        if (c->n_rg_excl) {
            const uint8_t *rg = rec_aux_get(b, "RG");
            skip = 0;
            if (rg) { int i; for (i = 0; i < c->n_rg_excl; i++) if (!strcmp(c->rg_excl[i], (const char *)rg + 1)) skip = 1; }
            if (skip) continue;
        }
        if (c->flag & FL_ILLUMINA13) {
            int i;
            for (i = 0; i < b->l_qseq; ++i) b->qual[i] = b->qual[i] > 31 ? b->qual[i] - 31 : 0;
        }
        if (c->fa) {
            has_ref = get_ref(c, ma->h, b->tid, &ref, &ref_len);
            if (has_ref && ref_len <= b->pos) {
                fprintf(stderr, "[mplp_func] Skipping because %lld is outside of %lld [ref:%d]\n",
                        (long long)b->pos, (long long)ref_len, b->tid);
                skip = 1; continue;
            }
        } else has_ref = 0;
        skip = 0;
        if (has_ref && (c->flag & FL_REALN)) baq_realn(b, ref, ref_len, (c->flag & FL_REDO_BAQ) ? 7 : 3);
        if (has_ref && c->capQ_thres > 10) {
            int q = cap_mapq(b, ref, ref_len, c->capQ_thres);
            if (q < 0) skip = 1;
            else if (b->mapq > q) b->mapq = (uint8_t)q;
        }
        if (b->mapq < c->min_mq) skip = 1;
        else if ((c->flag & FL_NO_ORPHAN) && (b->flag & F_PAIRED) && !(b->flag & F_PROPER)) skip = 1;
    } while (skip);
    return ret;
}

/* bam_plcmd.c:54-169 (without base modifications) */
static void pileup_seq(str_t *o, const pile1_t *p, hpos_t pos, hpos_t ref_len, const char *ref,
                       str_t *ins, const conf_t *c)
{
    int j, rev = (p->b->flag & F_REVERSE) != 0;
    if (!c->no_ends && p->is_head) { s_putc(o, '^'); s_putc(o, p->b->mapq > 93 ? 126 : p->b->mapq + 33); }
    if (!p->is_del) {
        static const char lc[] = ",acmgrsvtwyhkdbn", uc[] = ".ACMGRSVTWYHKDBN";
        int ch = p->qpos < p->b->l_qseq ? seqi(p->b->seq, p->qpos) : 15;
        if (ref) {
            int rb = pos < ref_len ? nt16_table[(uint8_t)ref[pos]] : 15;
            if (ch == rb) ch = 0;
        }
        s_putc(o, rev ? lc[ch] : uc[ch]);
    } else s_putc(o, p->is_refskip ? (rev ? '<' : '>') : ((rev && c->rev_del) ? '#' : '*'));
    int del_len = -p->indel;
    if (p->indel > 0) {
        int len = plp_insertion(p, ins, &del_len);
        if (c->no_ins < 2) { s_putc(o, '+'); s_putll(o, len); }
        if (!c->no_ins) {
            for (j = 0; j < (int)ins->l; j++) {
                int ch = (unsigned char)ins->s[j];
                if (rev) s_putc(o, ch != '*' ? tolower(ch) : (c->rev_del ? '#' : '*'));
                else s_putc(o, toupper(ch));
            }
        }
    }
    if (del_len > 0) {
        if (c->no_del < 2) s_putll(o, -del_len);
        if (!c->no_del)
            for (j = 1; j <= del_len; ++j) {
                int ch = (ref && (int)pos + j < ref_len) ? ref[pos + j] : 'N';
                s_putc(o, rev ? tolower(ch) : toupper(ch));
            }
    }
    if (!c->no_ends && p->is_tail) s_putc(o, '$');
}

static int n_xcols(const conf_t *c) { int i, n = 0; for (i = 0; i < X_N; i++) n += c->xcol[i] != 0; return n + c->n_tags; }

/* bam_plcmd.c:372-398 */
static void empty_row(str_t *o, const conf_t *c, const char *name, hpos_t pos, int nfn, const char *ref, hpos_t ref_len)
{
    int i, j, nx = n_xcols(c);
    s_puts(o, name); s_putc(o, '\t'); s_putll(o, pos + 1); s_putc(o, '\t');
    s_putc(o, (ref && pos < ref_len) ? ref[pos] : 'N');
    for (i = 0; i < nfn; i++) { s_putn(o, "\t0\t*\t*", 6); for (j = 0; j < nx; j++) s_putn(o, "\t*", 2); }
    s_putc(o, '\n');
}

static void flush(str_t *o, FILE *fp) { if (o->l) fwrite(o->s, 1, o->l, fp); o->l = 0; }

static void aux_value(str_t *o, const uint8_t *t, const conf_t *c)
{
    int ty = *t; const uint8_t *v = t + 1;
    if (ty == 'Z' || ty == 'H') s_puts(o, (const char *)v);
    else if (ty == 'c') s_putll(o, *(const int8_t *)v);
    else if (ty == 'C') s_putll(o, *v);
    else if (ty == 's') { int16_t x; memcpy(&x, v, 2); s_putll(o, x); }
    else if (ty == 'S') { uint16_t x; memcpy(&x, v, 2); s_putll(o, x); }
    else if (ty == 'i') { int32_t x; memcpy(&x, v, 4); s_putll(o, x); }
    else if (ty == 'I') { uint32_t x; memcpy(&x, v, 4); s_putll(o, x); }
    else if (ty == 'f') { float x; char b[64]; memcpy(&x, v, 4); snprintf(b, sizeof b, "%g", x); s_puts(o, b); }
    else if (ty == 'd') { double x; char b[64]; memcpy(&x, v, 8); snprintf(b, sizeof b, "%g", x); s_puts(o, b); }
    else if (ty == 'A') s_putc(o, *v);
    else s_putc(o, '*');
    (void)c;
}

static int mpileup(conf_t *conf, int nfn, char **fn)
{
    aux_t **data = calloc((size_t)nfn, sizeof(aux_t *));
    const pile1_t **plp = calloc((size_t)nfn, sizeof(pile1_t *));
    int *n_plp = calloc((size_t)nfn, sizeof(int));
    int i, tid, tid0 = 0, max_depth, ret;
    hpos_t pos, beg0 = 0, end0 = HPOS_MAX, ref_len = 0;
    const char *ref = NULL;
    hdr_t *h = NULL;
    str_t buf = {0, 0, NULL}, ks_seq = {0, 0, NULL}, ks_qual = {0, 0, NULL}, ks_ins = {0, 0, NULL};
    errmod_t *em = conf->gl ? errmod_new(1. - 0.83) : NULL;

    if (nfn == 0) { fprintf(stderr, "[mpileup] no input file/data given\n"); return 1; }
    /* sample counting for the stderr line (sample.c:79-121): distinct SM values, or file names */
    int n_smpl = 0; char **smpl = NULL;
    for (i = 0; i < nfn; i++) {
        char *fai = NULL;
        if (conf->fa_fn) { fai = malloc(strlen(conf->fa_fn) + 5); sprintf(fai, "%s.fai", conf->fa_fn); }
        data[i] = calloc(1, sizeof(aux_t));
        data[i]->rd = reader_open(fn[i], fai);
        free(fai);
        if (!data[i]->rd) { fprintf(stderr, "[mpileup] failed to open %s: %s\n", fn[i], strerror(errno)); return 1; }
        data[i]->conf = conf;
        hdr_t *ht = reader_hdr(data[i]->rd);
        {   /* bam_smpl_add */
            const char *p = (conf->flag & FL_IGNORE_RG) ? NULL : ht->text, *q;
            int n = 0, j;
            while (p && (q = strstr(p, "@RG")) != NULL) {
                const char *id, *sm;
                p = q + 3;
                id = strstr(p, "\tID:"); sm = strstr(p, "\tSM:");
                if (!(id && sm)) break;
                sm += 4;
                size_t L = strcspn(sm, "\t\n");
                char *v = strndup(sm, L);
                for (j = 0; j < n_smpl; j++) if (!strcmp(smpl[j], v)) break;
                if (j == n_smpl) { smpl = realloc(smpl, sizeof(char *) * (size_t)(n_smpl + 1)); smpl[n_smpl++] = v; } else free(v);
                p = (id + 4 > sm) ? id + 4 : sm;
                n++;
            }
            if (n == 0) {
                for (j = 0; j < n_smpl; j++) if (!strcmp(smpl[j], fn[i])) break;
                if (j == n_smpl) { smpl = realloc(smpl, sizeof(char *) * (size_t)(n_smpl + 1)); smpl[n_smpl++] = strdup(fn[i]); }
            }
        }
        if (conf->reg) {
            int t; hpos_t b, e;
            if (reader_set_region(data[i]->rd, conf->reg, &t, &b, &e) < 0) {
                fprintf(stderr, "[E::mpileup] fail to parse region '%s' with %s\n", conf->reg, fn[i]);
                return 1;
            }
            if (i == 0) { beg0 = b; end0 = e; tid0 = t; }
        }
        if (i == 0) h = ht;
        data[i]->h = h;
    }
    fprintf(stderr, "[mpileup] %d samples in %d input files\n", n_smpl, nfn);
    FILE *fp = conf->out_fn ? fopen(conf->out_fn, "w") : stdout;
    if (!fp) { fprintf(stderr, "[mpileup] failed to write to %s\n", conf->out_fn); return 1; }

    mplp_t *iter = mplp_init(nfn, pull, (void **)data);
    if (conf->flag & FL_OVERLAPS) mplp_init_overlaps(iter);
    if (!conf->max_depth) {
        max_depth = INT_MAX;
        fprintf(stderr, "[mpileup] Max depth set to maximum value (%d)\n", INT_MAX);
    } else {
        max_depth = conf->max_depth;
        if (max_depth * nfn > 1 << 20) fprintf(stderr, "[mpileup] Combined max depth is above 1M. Potential memory hog!\n");
    }
    mplp_set_maxcnt(iter, max_depth);
    int last_tid = -1, got_ref = 0, one_seq = 0;
    hpos_t last_pos = -1;

    while ((ret = mplp_auto(iter, &tid, &pos, n_plp, plp)) > 0) {
        one_seq = 1;
        if (conf->reg && (pos < beg0 || pos >= end0)) continue;
        if (conf->all) {
            while (tid > last_tid) {
                if (last_tid >= 0 && !conf->reg) {
                    while (++last_pos < h->len[last_tid]) {
                        if (conf->bed && bed_hit(conf->bed, h->name[last_tid], last_pos, last_pos + 1) == 0) continue;
                        empty_row(&buf, conf, h->name[last_tid], last_pos, nfn, ref, ref_len);
                        flush(&buf, fp);
                    }
                }
                last_tid++; got_ref = 0; last_pos = -1;
                if (conf->all < 2) break;
                if (tid > last_tid) got_ref = get_ref(conf, h, last_tid, &ref, &ref_len);
            }
        }
        if (!got_ref || last_tid != tid) { got_ref = get_ref(conf, h, tid, &ref, &ref_len); last_tid = tid; }
        if (conf->all) {
            while (++last_pos < pos) {
                if (conf->reg && last_pos < beg0) continue;
                if (conf->bed && bed_hit(conf->bed, h->name[tid], last_pos, last_pos + 1) == 0) continue;
                empty_row(&buf, conf, h->name[tid], last_pos, nfn, ref, ref_len);
                flush(&buf, fp);
            }
            last_pos = pos;
        }
        if (conf->bed && tid >= 0 && !bed_hit(conf->bed, h->name[tid], pos, pos + 1)) continue;

        s_puts(&buf, h->name[tid]); s_putc(&buf, '\t'); s_putll(&buf, pos + 1); s_putc(&buf, '\t');
        s_putc(&buf, (ref && pos < ref_len) ? ref[pos] : 'N');
        if (conf->gl) {
            /* GL mode: all files pooled per file, SNP likelihoods only */
            int rb4 = (ref && pos < ref_len) ? nt16_table[(uint8_t)ref[pos]] : 15;
            for (i = 0; i < nfn; ++i) {
                float qsum[4], p25[25]; int j;
                int n = glfgen(n_plp[i], plp[i], rb4, conf->min_baseQ, 60, em, qsum, p25);
                char t[64];
                s_putc(&buf, '\t'); s_putll(&buf, n < 0 ? 0 : n);
                for (j = 0; j < 4; j++) { snprintf(t, sizeof t, "\t%.9g", qsum[j]); s_puts(&buf, t); }
                for (j = 0; j < 25; j++) { snprintf(t, sizeof t, "\t%.9g", p25[j]); s_puts(&buf, t); }
            }
            s_putc(&buf, '\n');
            flush(&buf, fp);
            continue;
        }
        for (i = 0; i < nfn; ++i) {
            int j, cnt, x, t;
            ks_seq.l = ks_qual.l = 0;
            for (j = cnt = 0; j < n_plp[i]; ++j) {
                const pile1_t *p = plp[i] + j;
                int c = p->qpos < p->b->l_qseq ? p->b->qual[p->qpos] : 0;
                if (c >= conf->min_baseQ) {
                    pileup_seq(&ks_seq, p, pos, ref_len, ref, &ks_ins, conf);
                    s_putc(&ks_qual, c + 33 < 126 ? c + 33 : 126);
                    cnt++;
                }
            }
            s_putc(&buf, '\t'); s_putll(&buf, cnt); s_putc(&buf, '\t');
            if (n_plp[i] == 0) {
                s_putn(&buf, "*\t*", 3);
                for (x = 0; x < n_xcols(conf); x++) s_putn(&buf, "\t*", 2);
                continue;
            }
            if (ks_seq.l) s_putn(&buf, ks_seq.s, ks_seq.l); else s_putc(&buf, '*');
            s_putc(&buf, '\t');
            if (ks_qual.l) s_putn(&buf, ks_qual.s, ks_qual.l); else s_putc(&buf, '*');
            for (x = 0; x < X_N; x++) {
                if (!conf->xcol[x]) continue;
                int n = 0;
                s_putc(&buf, '\t');
                for (j = 0; j < n_plp[i]; ++j) {
                    const pile1_t *p = plp[i] + j;
                    int c = p->qpos < p->b->l_qseq ? p->b->qual[p->qpos] : 0;
                    if (c < conf->min_baseQ) continue;
                    if (n > 0 && x != X_MAPQ_CHAR) s_putc(&buf, ',');
                    n++;
                    switch (x) {
                    case X_MAPQ_CHAR: c = p->b->mapq + 33; if (c > 126) c = 126; s_putc(&buf, c); break;
                    case X_QPOS: s_putll(&buf, p->qpos + 1); break;
                    case X_QPOS5: s_putll(&buf, (p->b->flag & F_REVERSE) ? p->b->l_qseq - p->qpos + (int)p->is_del : p->qpos + 1); break;
                    case X_QNAME: s_puts(&buf, p->b->qname); break;
                    case X_FLAG: s_putll(&buf, p->b->flag); break;
                    case X_RNAME: if (p->b->tid >= 0) s_puts(&buf, h->name[p->b->tid]); else s_putc(&buf, '*'); break;
                    case X_POS: s_putll(&buf, p->b->pos + 1); break;
                    case X_MAPQ: s_putll(&buf, p->b->mapq); break;
                    case X_RNEXT: if (p->b->mtid >= 0) s_puts(&buf, h->name[p->b->mtid]); else s_putc(&buf, '*'); break;
                    case X_PNEXT: s_putll(&buf, p->b->mpos + 1); break;
                    case X_RLEN: s_putll(&buf, p->b->l_qseq); break;
                    }
                }
                if (!n) s_putc(&buf, '*');
            }
            for (t = 0; t < conf->n_tags; t++) {
                int n = 0;
                s_putc(&buf, '\t');
                for (j = 0; j < n_plp[i]; ++j) {
                    const pile1_t *p = plp[i] + j;
                    int c = p->qpos < p->b->l_qseq ? p->b->qual[p->qpos] : 0;
                    if (c < conf->min_baseQ) continue;
                    if (n > 0) s_putc(&buf, conf->sep);
                    n++;
                    const uint8_t *tg = rec_aux_get(p->b, conf->tags[t]);
                    if (!tg) { s_putc(&buf, conf->empty); continue; }
                    aux_value(&buf, tg, conf);
                }
                if (!n) s_putc(&buf, '*');
            }
        }
        s_putc(&buf, '\n');
        flush(&buf, fp);
    }
    if (ret < 0) { fprintf(stderr, "samtools mpileup: error reading from input file\n"); return 1; }

    if (conf->all && !conf->gl) {
        if (last_tid < 0 && conf->reg && conf->all > 1) {
            last_tid = tid0; last_pos = beg0 - 1;
            get_ref(conf, h, tid0, &ref, &ref_len);
        } else if (last_tid < 0 && !one_seq && conf->all > 1) last_tid = 0;
        while (last_tid >= 0 && last_tid < h->n_ref) {
            get_ref(conf, h, last_tid, &ref, &ref_len);
            while (++last_pos < h->len[last_tid]) {
                if (last_pos >= end0) break;
                if (conf->bed && bed_hit(conf->bed, h->name[last_tid], last_pos, last_pos + 1) == 0) continue;
                empty_row(&buf, conf, h->name[last_tid], last_pos, nfn, ref, ref_len);
                flush(&buf, fp);
            }
            last_tid++; last_pos = -1;
            if (conf->all < 2 || conf->reg) break;
        }
    }
    if (conf->out_fn) fclose(fp); else fflush(fp);
    mplp_destroy(iter);
    for (i = 0; i < nfn; i++) { reader_close(data[i]->rd); free(data[i]); }
    free(data); free(plp); free(n_plp); free(buf.s); free(ks_seq.s); free(ks_qual.s); free(ks_ins.s);
    for (i = 0; i < n_smpl; i++) free(smpl[i]);
    free(smpl);
    errmod_free(em);
    return 0;
}

int read_file_list(const char *fn, int *n, char ***files)
{
    FILE *f = fopen(fn, "r");
    char buf[1024];
    *n = 0; *files = NULL;
    if (!f) { fprintf(stderr, "%s: %s\n", fn, strerror(errno)); return 1; }
    while (fgets(buf, sizeof buf, f)) {
        size_t l = strlen(buf);
        while (l > 0 && isspace((unsigned char)buf[l - 1])) l--;
        if (!l) continue;
        buf[l] = 0;
        if (strncmp(buf, "file://", 7) == 0) memmove(buf, buf + 7, l - 6);
        *files = realloc(*files, sizeof(char *) * (size_t)(*n + 1));
        (*files)[(*n)++] = strdup(buf);
    }
    fclose(f);
    if (!*n) { fprintf(stderr, "No files read from %s\n", fn); return 1; }
    return 0;
}

int main_mpileup(int argc, char **argv, int gl)
{
    conf_t c;
    const char *file_list = NULL;
    int ch, use_orphan = 0, nfiles = 0, ret;
    char **fn = NULL;
    memset(&c, 0, sizeof c);
    c.min_baseQ = 13; c.max_depth = 8000;
    c.flag = FL_NO_ORPHAN | FL_REALN | FL_OVERLAPS;
    c.rflag_filter = F_UNMAP | F_SECONDARY | F_QCFAIL | F_DUP;
    c.sep = ','; c.empty = '*'; c.gl = gl;
    static const struct option lo[] = {
        {"rf", 1, 0, 1}, {"ff", 1, 0, 2}, {"incl-flags", 1, 0, 1}, {"excl-flags", 1, 0, 2},
        {"output", 1, 0, 3}, {"output-QNAME", 0, 0, 5}, {"output-qname", 0, 0, 5},
        {"illumina1.3+", 0, 0, '6'}, {"count-orphans", 0, 0, 'A'}, {"bam-list", 1, 0, 'b'},
        {"no-BAQ", 0, 0, 'B'}, {"no-baq", 0, 0, 'B'}, {"adjust-MQ", 1, 0, 'C'}, {"adjust-mq", 1, 0, 'C'},
        {"max-depth", 1, 0, 'd'}, {"redo-BAQ", 0, 0, 'E'}, {"redo-baq", 0, 0, 'E'}, {"fasta-ref", 1, 0, 'f'},
        {"reference", 1, 0, 'f'},
        {"exclude-RG", 1, 0, 'G'}, {"exclude-rg", 1, 0, 'G'}, {"positions", 1, 0, 'l'}, {"region", 1, 0, 'r'},
        {"ignore-RG", 0, 0, 'R'}, {"ignore-rg", 0, 0, 'R'}, {"min-MQ", 1, 0, 'q'}, {"min-mq", 1, 0, 'q'},
        {"min-BQ", 1, 0, 'Q'}, {"min-bq", 1, 0, 'Q'}, {"ignore-overlaps-removal", 0, 0, 'x'},
        {"disable-overlap-removal", 0, 0, 'x'}, {"output-BP", 0, 0, 'O'}, {"output-bp", 0, 0, 'O'},
        {"output-BP-5", 0, 0, 14}, {"output-bp-5", 0, 0, 14}, {"output-MQ", 0, 0, 's'}, {"output-mq", 0, 0, 's'},
        {"reverse-del", 0, 0, 6}, {"output-extra", 1, 0, 7}, {"output-sep", 1, 0, 8}, {"output-empty", 1, 0, 9},
        {"no-output-ins", 0, 0, 10}, {"no-output-del", 0, 0, 12}, {"no-output-ends", 0, 0, 13},
        {0, 0, 0, 0}
    };
    optind = 1;
    while ((ch = getopt_long(argc, argv, "Af:r:l:q:Q:RC:Bd:b:o:EG:6OsxXa", lo, NULL)) >= 0) {
        switch (ch) {
        case 'x': c.flag &= ~FL_OVERLAPS; break;
        case 1: c.rflag_require = parse_flag(optarg); if (c.rflag_require < 0) { fprintf(stderr, "Could not parse --rf %s\n", optarg); return 1; } break;
        case 2: c.rflag_filter = parse_flag(optarg); if (c.rflag_filter < 0) { fprintf(stderr, "Could not parse --ff %s\n", optarg); return 1; } break;
        case 3: case 'o': c.out_fn = optarg; break;
        case 5: c.xcol[X_QNAME] = 1; break;
        case 6: c.rev_del = 1; break;
        case 7: {
            static const struct { const char *n; int x; } cols[] = {
                {"QNAME", X_QNAME}, {"FLAG", X_FLAG}, {"RNAME", X_RNAME}, {"POS", X_POS}, {"MAPQ", X_MAPQ},
                {"RNEXT", X_RNEXT}, {"PNEXT", X_PNEXT}, {"RLEN", X_RLEN} };
            char *save, *tag = strtok_r(optarg, ",", &save);
            for (; tag; tag = strtok_r(NULL, ",", &save)) {
                size_t i;
                for (i = 0; i < sizeof cols / sizeof *cols; i++) if (!strcmp(cols[i].n, tag)) { c.xcol[cols[i].x] = 1; break; }
                if (i < sizeof cols / sizeof *cols) continue;
                if (strlen(tag) != 2) fprintf(stderr, "[build_auxlist] tag '%s' has more than two characters or not supported\n", tag);
                else { c.tags = realloc(c.tags, sizeof(char *) * (size_t)(c.n_tags + 1)); c.tags[c.n_tags++] = tag; }
            }
            break;
        }
        case 8: c.sep = optarg[0]; break;
        case 9: c.empty = optarg[0]; break;
        case 10: c.no_ins++; break;
        case 12: c.no_del++; break;
        case 13: c.no_ends = 1; break;
        case 14: c.xcol[X_QPOS5] = 1; break;
        case 'f': c.fa = fasta_load(optarg); if (!c.fa) return 1; c.fa_fn = optarg; break;
        case 'd': c.max_depth = atoi(optarg); break;
        case 'r': c.reg = strdup(optarg); break;
        case 'l': c.bed = bed_load(optarg); if (!c.bed) { fprintf(stderr, "samtools mpileup: Could not read file \"%s\"\n", optarg); return 1; } break;
        case 'B': c.flag &= ~FL_REALN; break;
        case 'X': break;
        case 'E': c.flag |= FL_REDO_BAQ; break;
        case '6': c.flag |= FL_ILLUMINA13; break;
        case 'R': c.flag |= FL_IGNORE_RG; break;
        case 's': c.xcol[X_MAPQ_CHAR] = 1; break;
        case 'O': c.xcol[X_QPOS] = 1; break;
        case 'C': c.capQ_thres = atoi(optarg); break;
        case 'q': c.min_mq = atoi(optarg); break;
        case 'Q': c.min_baseQ = atoi(optarg); break;
        case 'b': file_list = optarg; break;
        case 'A': use_orphan = 1; break;
        case 'G': {
            FILE *f = fopen(optarg, "r"); char b[1024];
            if (f) { while (fscanf(f, "%1023s", b) > 0) { c.rg_excl = realloc(c.rg_excl, sizeof(char *) * (size_t)(c.n_rg_excl + 1)); c.rg_excl[c.n_rg_excl++] = strdup(b); } fclose(f); }
            break;
        }
        case 'a': c.all++; break;
        default: fprintf(stderr, "Usage: mpileup [options] in1.bam [in2.bam [...]]\n"); return 1;
        }
    }
    if (!(c.flag & FL_REALN) && (c.flag & FL_REDO_BAQ)) { fprintf(stderr, "Error: The -B option cannot be combined with -E\n"); return 1; }
    if (use_orphan) c.flag &= ~FL_NO_ORPHAN;
    if (argc == 1) { fprintf(stderr, "Usage: mpileup [options] in1.bam [in2.bam [...]]\n"); return 1; }
    if (file_list) {
        if (read_file_list(file_list, &nfiles, &fn)) return 1;
        ret = mpileup(&c, nfiles, fn);
    } else ret = mpileup(&c, argc - optind, argv + optind);
    return ret;
}

/* `pileup-dump [-o] [-d maxcnt] files...`: every pile1_t field per column, for checking the
 * htslib-compatible iterator tier of the product (tests/compat/plp_dump.cpp) */
static int dump_pull(void *data, rec_t *b) { return reader_next((reader_t *)data, b); }
int main_pileup_dump(int argc, char **argv)
{
    int overlaps = 0, maxcnt = 8000, a = 1, i, j;
    for (; a < argc && argv[a][0] == '-'; ++a) {
        if (!strcmp(argv[a], "-o")) overlaps = 1;
        else if (!strcmp(argv[a], "-d") && a + 1 < argc) maxcnt = atoi(argv[++a]);
    }
    int n = argc - a;
    if (n < 1) return 1;
    void **data = calloc((size_t)n, sizeof(void *));
    for (i = 0; i < n; i++) { data[i] = reader_open(argv[a + i], NULL); if (!data[i]) return 1; }
    mplp_t *it = mplp_init(n, dump_pull, data);
    if (overlaps) mplp_init_overlaps(it);
    mplp_set_maxcnt(it, maxcnt);
    int *n_plp = calloc((size_t)n, sizeof(int)); const pile1_t **plp = calloc((size_t)n, sizeof(pile1_t *));
    int tid, ret; hpos_t pos; str_t ins = {0, 0, NULL};
    while ((ret = mplp_auto(it, &tid, &pos, n_plp, plp)) > 0) {
        printf("%d\t%lld", tid, (long long)pos);
        for (i = 0; i < n; i++) {
            printf("\t%d:", n_plp[i]);
            for (j = 0; j < n_plp[i]; j++) {
                const pile1_t *p = plp[i] + j;
                int q = p->qpos < p->b->l_qseq ? p->b->qual[p->qpos] : -1, dl = 0;
                int il = plp_insertion(p, &ins, &dl);
                printf(" %s/%d/%d/%d%d%d%d/%d/%d/%s/%d", p->b->qname, p->qpos, p->indel, p->is_del, p->is_head, p->is_tail, p->is_refskip,
                       p->cigar_ind, q, il > 0 ? ins.s : "-", dl);
            }
        }
        putchar('\n');
    }
    return ret < 0 ? 1 : 0;
}
