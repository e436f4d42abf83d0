/*
 * oracle/coverage.c -- TEST INFRASTRUCTURE ONLY (CPU oracle).
 *
 * Restates the reference's `samtools coverage` tabular mode, coverage.c: read
 * callback read_bam (:178-198), per-column reducers (:589-661), row printing
 * print_tabular_line (:200-221) and option parsing (:306-461).  The ASCII /
 * UTF-8 histogram views (-m/-A/-D/-w) are terminal UI and not restated.
 * Pinned by test/coverage/{1..5}.expected on test/dat/sample.sam.
 */
#include "plp.h"
#include <getopt.h>
#include <limits.h>
#include <stdbool.h>

int read_file_list(const char *fn, int *n, char ***files);

typedef struct {
    unsigned long long n_covered_bases, summed_coverage, summed_baseQ, summed_mapQ, quality_bases;
    unsigned int n_reads, n_selected_reads;
    bool covered;
    hpos_t beg, end;
} cstat_t;

typedef struct { reader_t *rd; hdr_t *h; int min_mapQ, min_len, fail_flags, required_flags; cstat_t *stats; } caux_t;

static int read_bam(void *data, rec_t *b)
{
    caux_t *a = data;
    int nref = a->h->n_ref, ret;
    for (;;) {
        if ((ret = reader_next(a->rd, b)) < 0) break;
        if (b->tid >= 0 && b->tid < nref) a->stats[b->tid].n_reads++;
        if (a->fail_flags && (b->flag & a->fail_flags)) continue;
        if (a->required_flags && !(b->flag & a->required_flags)) continue;
        if (b->mapq < a->min_mapQ) continue;
        if (a->min_len && rec_qlen(b) < a->min_len) continue;
        if (b->tid >= 0 && b->tid < nref) { a->stats[b->tid].n_selected_reads++; a->stats[b->tid].summed_mapQ += b->mapq; }
        break;
    }
    return ret;
}

static void print_row(FILE *out, const hdr_t *h, const cstat_t *s, int tid, bool *header)
{
    if (*header) { fputs("#rname\tstartpos\tendpos\tnumreads\tcovbases\tcoverage\tmeandepth\tmeanbaseq\tmeanmapq\n", out); *header = false; }
    fputs(h->name[tid], out);
    double region_len = (double)s[tid].end - s[tid].beg;
    fprintf(out, "\t%lld\t%lld\t%u\t%llu\t%g\t%g\t%.3g\t%.3g\n",
            (long long)s[tid].beg + 1, (long long)s[tid].end, s[tid].n_selected_reads, s[tid].n_covered_bases,
            100.0 * s[tid].n_covered_bases / region_len,
            s[tid].summed_coverage / region_len,
            s[tid].quality_bases > 0 ? s[tid].summed_baseQ / (double)s[tid].quality_bases : 0,
            s[tid].n_selected_reads > 0 ? s[tid].summed_mapQ / (double)s[tid].n_selected_reads : 0);
}

int main_coverage(int argc, char **argv)
{
    int max_depth = 1000000, min_baseQ = 0, min_mapQ = 0, min_len = 0, mindepth = 1;
    int fail_flags = F_UNMAP | F_SECONDARY | F_QCFAIL | F_DUP, required_flags = 0;
    char *opt_reg = NULL, *file_list = NULL, *out_fn = NULL, **fn = NULL;
    bool print_header = true;
    int c, i, j, n;
    static const struct option lo[] = {
        {"rf", 1, 0, 1}, {"ff", 1, 0, 2}, {"incl-flags", 1, 0, 1}, {"excl-flags", 1, 0, 2},
        {"bam-list", 1, 0, 'b'}, {"min-read-len", 1, 0, 'l'}, {"min-MQ", 1, 0, 'q'}, {"min-mq", 1, 0, 'q'},
        {"min-BQ", 1, 0, 'Q'}, {"min-bq", 1, 0, 'Q'}, {"output", 1, 0, 'o'}, {"no-header", 0, 0, 'H'},
        {"region", 1, 0, 'r'}, {"depth", 1, 0, 'd'}, {"min-depth", 1, 0, 3}, {0, 0, 0, 0}
    };
    optind = 1;
    while ((c = getopt_long(argc, argv, "o:l:q:Q:Hr:b:d:", lo, NULL)) != -1) {
        switch (c) {
        case 1: if ((required_flags = parse_flag(optarg)) < 0) { fprintf(stderr, "Could not parse --rf %s\n", optarg); return 1; } break;
        case 2: if ((fail_flags = parse_flag(optarg)) < 0) { fprintf(stderr, "Could not parse --ff %s\n", optarg); return 1; } break;
        case 3: if ((i = atoi(optarg)) > 0) mindepth = i; break;
        case 'o': out_fn = optarg; break;
        case 'l': min_len = atoi(optarg); break;
        case 'q': min_mapQ = atoi(optarg); break;
        case 'Q': min_baseQ = atoi(optarg); break;
        case 'd': max_depth = atoi(optarg); break;
        case 'r': opt_reg = optarg; break;
        case 'b': file_list = optarg; break;
        case 'H': print_header = false; break;
        default: fprintf(stderr, "Usage: coverage [options] in1.bam [in2.bam [...]]\n"); return 1;
        }
    }
    if (file_list) { if (read_file_list(file_list, &n, &fn)) return 1; }
    else { n = argc - optind; fn = argv + optind; }
    if (n < 1) { fprintf(stderr, "Usage: coverage [options] in1.bam [in2.bam [...]]\n"); return 1; }
    FILE *out = (out_fn && strcmp(out_fn, "-")) ? fopen(out_fn, "w") : stdout;
    if (!out) return 1;
    caux_t **data = calloc((size_t)n, sizeof(caux_t *));
    int rtid = -1; hpos_t rbeg = 0, rend = 0;
    for (i = 0; i < n; i++) {
        data[i] = calloc(1, sizeof(caux_t));
        data[i]->rd = reader_open(fn[i], NULL);
        if (!data[i]->rd) { fprintf(stderr, "samtools coverage: Could not open \"%s\"\n", fn[i]); return 1; }
        data[i]->h = reader_hdr(data[i]->rd);
        data[i]->min_mapQ = min_mapQ; data[i]->min_len = min_len;
        data[i]->fail_flags = fail_flags; data[i]->required_flags = required_flags;
        if (opt_reg) {
            int t; hpos_t b, e;
            if (reader_set_region(data[i]->rd, opt_reg, &t, &b, &e) < 0) { fprintf(stderr, "samtools coverage: Failed to parse region \"%s\"\n", opt_reg); return 1; }
            if (i == 0) { rtid = t; rbeg = b; rend = e; }
        }
    }
    hdr_t *h = data[0]->h;
    int n_targets = h->n_ref;
    cstat_t *stats = calloc((size_t)n_targets + 1, sizeof(cstat_t));
    if (opt_reg) {
        cstat_t *s = stats + rtid;
        s->beg = rbeg; s->end = rend;
        if (s->end == HPOS_MAX) s->end = h->len[rtid];
    }
    for (i = 0; i < n; i++) data[i]->stats = stats;
    mplp_t *mplp = mplp_init(n, read_bam, (void **)data);
    if (max_depth > 0) mplp_set_maxcnt(mplp, max_depth);
    else if (!max_depth) mplp_set_maxcnt(mplp, INT_MAX);
    int *n_plp = calloc((size_t)n, sizeof(int));
    const pile1_t **plp = calloc((size_t)n, sizeof(pile1_t *));
    int ret, tid = -1, old_tid = -1, warn = 0; hpos_t pos;
    while ((ret = mplp_auto(mplp, &tid, &pos, n_plp, plp)) > 0) {
        if (tid != old_tid) {
            if (old_tid >= 0) print_row(out, h, stats, old_tid, &print_header);
            stats[tid].covered = true;
            if (!opt_reg) stats[tid].end = h->len[tid];
            old_tid = tid;
        }
        if (pos < stats[tid].beg || pos >= stats[tid].end) continue;
        if (tid >= n_targets) continue;
        bool count_base = false;
        unsigned long long sbq = 0, qb = 0, depth = 0;
        for (i = 0; i < n; ++i) {
            int d = n_plp[i];
            for (j = 0; j < n_plp[i]; ++j) {
                const pile1_t *p = plp[i] + j;
                if (p->is_del || p->is_refskip) --d;
                else if (p->qpos < p->b->l_qseq) {
                    if (p->b->qual[p->qpos] < min_baseQ) --d;
                    else { sbq += p->b->qual[p->qpos]; ++qb; }
                } else warn = 1;
            }
            if (d > 0) { count_base = true; depth += (unsigned long long)d; }
        }
        if (count_base && depth >= (unsigned long long)mindepth) {
            stats[tid].summed_coverage += depth;
            stats[tid].summed_baseQ += sbq;
            stats[tid].quality_bases += qb;
            stats[tid].n_covered_bases++;
        }
    }
    if (ret < 0) return 1;
    if (tid == -1 && opt_reg && *opt_reg != '*') tid = rtid;
    if (tid < n_targets && tid >= 0) print_row(out, h, stats, tid, &print_header);
    if (!opt_reg)
        for (i = 0; i < n_targets; ++i)
            if (!stats[i].covered) { stats[i].end = h->len[i]; print_row(out, h, stats, i, &print_header); }
    if (warn) fprintf(stderr, "samtools coverage: Warning:  Missing quality values in alignments.  Mean base quality calculated only on available values.\n");
    mplp_destroy(mplp);
    for (i = 0; i < n; i++) { reader_close(data[i]->rd); free(data[i]); }
    free(data); free(stats); free(n_plp); free(plp);
    if (out != stdout) fclose(out); else fflush(out);
    return 0;
}
