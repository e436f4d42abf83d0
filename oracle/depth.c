/*
 * oracle/depth.c -- TEST INFRASTRUCTURE ONLY (CPU oracle).
 *
 * Restates the reference's `samtools depth`, bam2depth.c: per-read CIGAR
 * accumulation add_depth (:209-477), k-way file merge and read filters
 * fastdepth_core (:486-699), zero_region (:88-118), qlen_used (:124-159) and
 * option parsing main_depth (:732-1006).  The reference's power-of-two ring
 * buffer is replaced by a sliding window array with the same visible
 * behaviour (counts, end_pos guard, flush order).
 * Pinned by test/mpileup/depth.reg (expected/d*.out) and test/large_pos.
 */
#include "hl.h"
#include <getopt.h>
#include <limits.h>
#include <errno.h>

int read_file_list(const char *fn, int *n, char ***files);

typedef struct {
    int header, flag, incl_flag, require_flag, min_qual, min_mqual, min_len, skip_del, all_pos, remove_overlaps;
    FILE *out; char *reg; bed_t *bed;
} dopt_t;

typedef struct {
    int nfiles;
    int **hist;          /* hist[file][i - win0] */
    hpos_t win0; size_t wsize;
    hpos_t *end_pos;
    hpos_t last_output;
    int last_ref;
    const char *ref;
    hpos_t beg, end; int tid;
    str_t ks; size_t pre;  /* ks holds "name\t", pre = its length */
} dhist_t;

static void zero_region(dopt_t *o, dhist_t *dh, const char *name, hpos_t start, hpos_t end)
{
    str_t s = {0, 0, NULL};
    hpos_t i; int n;
    s_puts(&s, name); s_putc(&s, '\t');
    size_t cur = s.l;
    if (dh->beg >= 0 && start < dh->beg) start = dh->beg;
    if (dh->end >= 0 && end > dh->end) end = dh->end;
    for (i = start; i < end; i++) {
        if (o->bed && bed_hit(o->bed, name, i, i + 1) == 0) continue;
        s.l = cur;
        s_putll(&s, i + 1);
        for (n = 0; n < dh->nfiles; n++) { s_putc(&s, '\t'); s_putc(&s, '0'); }
        s_putc(&s, '\n');
        fputs(s.s, o->out);
    }
    free(s.s);
}

static hpos_t qlen_used(const rec_t *b)
{
    int n = (int)b->n_cigar, kl, kr, k;
    hpos_t l;
    if (b->l_qseq) {
        l = b->l_qseq;
        for (kl = 0; kl < n; kl++) { if (cop(b->cigar[kl]) == C_S) l -= cln(b->cigar[kl]); else break; }
        for (kr = n - 1; kr > kl; kr--) { if (cop(b->cigar[kr]) == C_S) l -= cln(b->cigar[kr]); else break; }
    } else {
        for (k = 0, l = 0; k < n; k++) {
            int op = cop(b->cigar[k]);
            if (op == C_M || op == C_I || op == C_EQ || op == C_X) l += cln(b->cigar[k]);
        }
    }
    return l;
}

/* drop window entries before `from`, make sure [from, upto) is addressable */
static void win_fit(dhist_t *dh, hpos_t from, hpos_t upto)
{
    int n;
    if (from > dh->win0) {
        size_t shift = (size_t)(from - dh->win0);
        for (n = 0; n < dh->nfiles; n++) {
            if (shift < dh->wsize) {
                memmove(dh->hist[n], dh->hist[n] + shift, (dh->wsize - shift) * sizeof(int));
                memset(dh->hist[n] + (dh->wsize - shift), 0, shift * sizeof(int));
            } else memset(dh->hist[n], 0, dh->wsize * sizeof(int));
        }
        dh->win0 = from;
    }
    if ((size_t)(upto - dh->win0) > dh->wsize) {
        size_t ns = dh->wsize ? dh->wsize : 2048;
        while (ns < (size_t)(upto - dh->win0)) ns <<= 1;
        for (n = 0; n < dh->nfiles; n++) {
            dh->hist[n] = realloc(dh->hist[n], ns * sizeof(int));
            memset(dh->hist[n] + dh->wsize, 0, (ns - dh->wsize) * sizeof(int));
        }
        dh->wsize = ns;
    }
}
#define H(f, i) dh->hist[f][(i) - dh->win0]

/* print rows [last_output, lim) while any file still has data; returns first unprinted */
static hpos_t flush_rows(dopt_t *o, dhist_t *dh, hpos_t lim, int bounded)
{
    hpos_t i; int n, nf = dh->nfiles;
    for (i = dh->last_output; nf && (!bounded || i < lim); i++) {
        nf = 0;
        for (n = 0; n < dh->nfiles; n++) if (i < dh->end_pos[n]) nf++;
        if (!nf) break;
        if (o->bed && bed_hit(o->bed, dh->ref, i, i + 1) == 0) continue;
        dh->ks.l = dh->pre;
        s_putll(&dh->ks, i + 1);
        for (n = 0; n < dh->nfiles; n++) {
            s_putc(&dh->ks, '\t');
            s_putll(&dh->ks, i < dh->end_pos[n] ? H(n, i) : 0);
        }
        s_putc(&dh->ks, '\n');
        fputs(dh->ks.s, o->out);
    }
    return i;
}

/* bam2depth.c:209-477 */
static int add_depth(dopt_t *o, dhist_t *dh, hdr_t *h, rec_t *b, hpos_t overlap_clip, int file)
{
    hpos_t i; int n;
    if (!b || b->tid != dh->last_ref) {
        if (dh->last_ref >= 0) {
            i = flush_rows(o, dh, 0, 0);
            if (o->all_pos) zero_region(o, dh, h->name[dh->last_ref], i, h->len[dh->last_ref]);
        }
        if (o->all_pos > 1 && !o->reg) {
            int lr = dh->last_ref < 0 ? 0 : dh->last_ref + 1, rr = b ? b->tid : h->n_ref, r;
            for (r = lr; r < rr; r++) zero_region(o, dh, h->name[r], 0, h->len[r]);
        }
        if (!b) {
            if (o->all_pos && o->reg && dh->last_ref < 0) {
                hpos_t e = dh->end < h->len[dh->tid] ? dh->end : h->len[dh->tid];
                zero_region(o, dh, h->name[dh->tid], dh->beg, e);
            }
            return 0;
        }
        for (n = 0; n < dh->nfiles; n++) dh->end_pos[n] = 0;
        dh->last_output = dh->beg >= 0 ? (b->pos > dh->beg ? b->pos : dh->beg) : b->pos;
        dh->last_ref = b->tid;
        dh->ref = h->name[b->tid];
        dh->ks.l = 0; s_puts(&dh->ks, dh->ref); s_putc(&dh->ks, '\t'); dh->pre = dh->ks.l;
        for (n = 0; n < dh->nfiles; n++) if (dh->wsize) memset(dh->hist[n], 0, dh->wsize * sizeof(int));
        dh->win0 = b->pos;
        if (o->all_pos) zero_region(o, dh, dh->ref, 0, b->pos);
    } else if (dh->last_output < b->pos) {
        i = flush_rows(o, dh, b->pos, 1);
        if (o->all_pos && i < b->pos) zero_region(o, dh, dh->ref, i, b->pos);
        dh->last_output = b->pos;
    }

    hpos_t end_pos = rec_endpos(b);
    if (b->tid < dh->last_ref || (dh->last_ref == b->tid && end_pos < dh->last_output)) {
        fprintf(stderr, "samtools depth: Data is not position sorted\n");
        return -1;
    }
    /* window covers [min(last_output, pos), end_pos]; both bounds only move
     * right within a reference, so older entries can be dropped */
    {
        hpos_t lo = dh->last_output < b->pos ? dh->last_output : b->pos;
        win_fit(dh, lo, end_pos + 1);
    }
    /* clear any stale counts at never-seen coordinates of this file */
    {
        hpos_t e = dh->end_pos[file] > b->pos ? dh->end_pos[file] : b->pos;
        for (i = e; i < end_pos; i++) H(file, i) = 0;
    }

    const uint32_t *cig = b->cigar;
    int ncig = (int)b->n_cigar, j, k, spos = 0;
    const uint8_t *qual = b->qual;
    int min_qual = o->min_qual;
    i = b->pos;
    for (j = 0; j < ncig; j++) {
        int op = cop(cig[j]), oplen = (int)cln(cig[j]);
        switch (op) {
        case C_D: case C_N:
            if (op != C_D || o->skip_del) { i += oplen; }
            else {
                k = 0;
                if (overlap_clip) {
                    if (i + oplen <= overlap_clip) { i += oplen; break; }
                    else if (i < overlap_clip) { k = (int)(overlap_clip - i); i = overlap_clip; }
                }
                if (spos < b->l_qseq) for (; k < oplen; k++, i++) H(file, i) += qual[spos] >= min_qual;
                else for (; k < oplen; k++, i++) H(file, i)++;
            }
            break;
        case C_M: case C_EQ: case C_X:
            if (overlap_clip) {
                if (i + oplen <= overlap_clip) { i += oplen; spos += oplen; break; }
                else if (i < overlap_clip) { oplen -= (int)(overlap_clip - i); spos += (int)(overlap_clip - i); i = overlap_clip; }
            }
            if (!min_qual) for (k = 0; k < oplen; k++) H(file, i + k)++;
            else for (k = 0; k < oplen; k++) H(file, i + k) += qual[spos + k] >= min_qual;
            spos += oplen; i += oplen;
            break;
        case C_I: case C_S: spos += oplen; break;
        case C_P: case C_H: break;
        default:
            fprintf(stderr, "samtools depth: Unsupported cigar op '%d'\n", op);
            return -1;
        }
    }
    if (dh->end >= 0 && end_pos > dh->end) end_pos = dh->end;
    if (dh->end_pos[file] < end_pos) dh->end_pos[file] = end_pos;
    return 0;
}

typedef struct oent { char *key; hpos_t val; struct oent *next; } oent;
#define ONB 4096
static uint32_t shash(const char *s) { uint32_t h = 5381; while (*s) h = h * 33 + (uint8_t)*s++; return h; }

static int pass_filters(const dopt_t *o, const rec_t *b)
{
    if (b->tid < 0) return 0;
    if (b->flag & o->flag) return 0;
    if (o->incl_flag && (b->flag & o->incl_flag) == 0) return 0;
    if ((b->flag & o->require_flag) != o->require_flag) return 0;
    if (b->mapq < o->min_mqual) return 0;
    if (o->min_len && qlen_used(b) < o->min_len) return 0;
    return 1;
}

/* bam2depth.c:486-699 */
static int fastdepth_core(dopt_t *o, int nfiles, char **fn, reader_t **rd, hdr_t **h, int have_reg, int rtid, hpos_t rbeg, hpos_t rend)
{
    int i, ret = 0, to_go = nfiles;
    rec_t *b = calloc((size_t)nfiles, sizeof(rec_t));
    int *finished = calloc((size_t)nfiles, sizeof(int));
    oent ***ovl = NULL;
    dhist_t dh; memset(&dh, 0, sizeof dh);
    for (i = 0; i < nfiles; i++) rec_init(&b[i]);
    if (o->remove_overlaps) { ovl = calloc((size_t)nfiles, sizeof(oent **)); for (i = 0; i < nfiles; i++) ovl[i] = calloc(ONB, sizeof(oent *)); }
    dh.nfiles = nfiles;
    dh.hist = calloc((size_t)nfiles, sizeof(int *));
    dh.end_pos = calloc((size_t)nfiles, sizeof(hpos_t));
    dh.last_ref = -99;
    dh.last_output = have_reg ? rbeg : 0;
    dh.beg = dh.end = -1; dh.tid = 0;
    if (have_reg) { dh.tid = rtid; dh.beg = rbeg; dh.end = rend; }
    if (o->header) {
        fprintf(o->out, "#CHROM\tPOS");
        for (i = 0; i < nfiles; i++) fprintf(o->out, "\t%s", fn[i]);
        fputc('\n', o->out);
    }
    for (i = 0; i < nfiles; i++) {
        for (;;) {
            ret = reader_next(rd[i], &b[i]);
            if (ret < -1) goto err;
            if (ret == -1) { to_go--; finished[i] = 1; break; }
            if (pass_filters(o, &b[i])) break;
        }
    }
    while (to_go) {
        int best_tid = INT_MAX, best_file = 0; hpos_t best_pos = HPOS_MAX;
        for (i = 0; i < nfiles; i++) {
            if (finished[i]) continue;
            if (best_tid > b[i].tid) { best_tid = b[i].tid; best_pos = b[i].pos; best_file = i; }
            else if (best_tid == b[i].tid && best_pos > b[i].pos) { best_pos = b[i].pos; best_file = i; }
        }
        i = best_file;
        hpos_t clip = 0;
        if (ovl && (b[i].flag & F_PAIRED) && !(b[i].flag & F_MUNMAP)) {
            oent **pp = &ovl[i][shash(b[i].qname) % ONB];
            while (*pp && strcmp((*pp)->key, b[i].qname)) pp = &(*pp)->next;
            if (!*pp) {
                hpos_t endpos = rec_endpos(&b[i]);
                if (b[i].mpos == -1 || (b[i].tid == b[i].mtid && b[i].mpos <= endpos)) {
                    oent *e = calloc(1, sizeof(*e)); e->key = strdup(b[i].qname); e->val = endpos; *pp = e;
                }
            } else { oent *e = *pp; clip = e->val; *pp = e->next; free(e->key); free(e); }
        }
        if (add_depth(o, &dh, h[i], &b[i], clip, i) < 0) { ret = -1; goto err; }
        for (; !finished[i];) {
            ret = reader_next(rd[i], &b[i]);
            if (ret < -1) { ret = -1; goto err; }
            if (ret == -1) { to_go--; finished[i] = 1; break; }
            if (pass_filters(o, &b[i])) break;
        }
    }
    ret = add_depth(o, &dh, h[0], NULL, 0, 0);
err:
    for (i = 0; i < nfiles; i++) { rec_free(&b[i]); free(dh.hist[i]); }
    free(b); free(finished); free(dh.hist); free(dh.end_pos); free(dh.ks.s);
    return ret < 0 ? -1 : 0;
}

int main_depth(int argc, char **argv)
{
    dopt_t o; memset(&o, 0, sizeof o);
    o.flag = F_UNMAP | F_SECONDARY | F_DUP | F_QCFAIL; o.skip_del = 1; o.out = stdout;
    static const struct option lo[] = {
        {"min-MQ", 1, 0, 'Q'}, {"min-mq", 1, 0, 'Q'}, {"min-BQ", 1, 0, 'q'}, {"min-bq", 1, 0, 'q'},
        {"excl-flags", 1, 0, 'G'}, {"incl-flags", 1, 0, 1}, {"require-flags", 1, 0, 2}, {"threads", 1, 0, '@'}, {0, 0, 0, 0}
    };
    int c, tmp, nfiles, i; char *file_list = NULL, **fn = NULL;
    optind = 1;
    while ((c = getopt_long(argc, argv, "@:q:Q:JHd:m:l:g:G:o:ar:Xf:b:s", lo, NULL)) >= 0) {
        switch (c) {
        case 'a': o.all_pos++; break;
        case 'b': o.bed = bed_load(optarg); if (!o.bed) { fprintf(stderr, "samtools depth: Could not read file \"%s\"\n", optarg); return 1; } break;
        case 'f': file_list = optarg; break;
        case 'd': case 'm': case '@': case 'X': break;
        case 'g': tmp = parse_flag(optarg); if (tmp < 0) { fprintf(stderr, "samtools depth: Unknown flag '%s'\n", optarg); return 1; } o.flag &= ~tmp; break;
        case 'G': tmp = parse_flag(optarg); if (tmp < 0) { fprintf(stderr, "samtools depth: Unknown flag '%s'\n", optarg); return 1; } o.flag |= tmp; break;
        case 1: tmp = parse_flag(optarg); if (tmp < 0) return 1; o.incl_flag |= tmp; break;
        case 2: tmp = parse_flag(optarg); if (tmp < 0) return 1; o.require_flag |= tmp; break;
        case 'l': o.min_len = atoi(optarg); break;
        case 'H': o.header = 1; break;
        case 'q': o.min_qual = atoi(optarg); break;
        case 'Q': o.min_mqual = atoi(optarg); break;
        case 'J': o.skip_del = 0; break;
        case 'o': if (o.out == stdout) { o.out = fopen(optarg, "w"); if (!o.out) return 1; } break;
        case 'r': o.reg = optarg; break;
        case 's': o.remove_overlaps = 1; break;
        default: fprintf(stderr, "Usage: depth [options] in.bam [in.bam ...]\n"); return 1;
        }
    }
    if (file_list) { if (read_file_list(file_list, &nfiles, &fn)) return 1; }
    else { nfiles = argc - optind; fn = argv + optind; }
    if (nfiles < 1) { fprintf(stderr, "Usage: depth [options] in.bam [in.bam ...]\n"); return 1; }
    reader_t **rd = calloc((size_t)nfiles, sizeof(reader_t *));
    hdr_t **h = calloc((size_t)nfiles, sizeof(hdr_t *));
    int rtid = 0; hpos_t rbeg = 0, rend = 0;
    for (i = 0; i < nfiles; i++) {
        rd[i] = reader_open(fn[i], NULL);
        if (!rd[i]) { fprintf(stderr, "samtools depth: Cannot open input file \"%s\"\n", fn[i]); return 1; }
        h[i] = reader_hdr(rd[i]);
        if (o.reg) {
            int t; hpos_t b, e;
            if (reader_set_region(rd[i], o.reg, &t, &b, &e) < 0) { fprintf(stderr, "samtools depth: cannot parse region \"%s\"\n", o.reg); return 1; }
            if (i == 0) { rtid = t; rbeg = b; rend = e; }
        }
    }
    int ret = fastdepth_core(&o, nfiles, fn, rd, h, o.reg != NULL, rtid, rbeg, rend) ? 1 : 0;
    for (i = 0; i < nfiles; i++) reader_close(rd[i]);
    if (o.out != stdout) fclose(o.out); else fflush(stdout);
    return ret;
}
