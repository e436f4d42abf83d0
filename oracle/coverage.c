/*
 * oracle/coverage.c -- TEST INFRASTRUCTURE ONLY (CPU oracle).
 *
 * Restates the reference's `samtools coverage` tabular mode, coverage.c: read
 * callback read_bam (:178-198), per-column reducers (:589-661), row printing
 * print_tabular_line (:200-221), option parsing (:306-461) and the histogram views
 * (-m/-A/-D/-w: per-bin breadth / depth counters :609-660, print_hist :223-304).
 * Tabular mode is pinned by test/coverage/{1..5}.expected on test/dat/sample.sam; the
 * histogram views have no golden file in the reference (parity unpinned); one
 * behaviour is DEFINED here where the reference's is undefined: an all-zero histogram
 * divides 0 by 0 and converts NaN to int -- the x86-64 build ends up with the full
 * block in every cell, which is what is printed here.
 */
#include "plp.h"
#include <getopt.h>
#include <limits.h>
#include <stdbool.h>
#include <math.h>
#include <sys/ioctl.h>

int read_file_list(const char *fn, int *n, char ***files);

typedef struct {
    unsigned long long n_covered_bases, summed_coverage, summed_baseQ, summed_mapQ, quality_bases;
    unsigned int n_reads, n_selected_reads;
    bool covered;
    hpos_t beg, end;
    int64_t bin_width;
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

/* ---- histogram view (coverage.c:160-176 center_text / readable_bps, :223-304 print_hist) */
static const char *const kBlocks8[8] = {"\xE2\x96\x81", "\xE2\x96\x82", "\xE2\x96\x83", "\xE2\x96\x84", "\xE2\x96\x85", "\xE2\x96\x86", "\xE2\x96\x87", "\xE2\x96\x88"};
static const char *const kBlocks2[2] = {".", ":"};

static char *fmt_bp(double bp, char *buf)            /* 1234567 -> "1.23M": as many decimals as thousands were divided out */
{
    static const char *unit[] = {"", "K", "M", "G", "T"};
    int u = 0;
    while (bp >= 1000 && u < 4) { bp /= 1000; ++u; }
    sprintf(buf, "%.*f%s", u, bp, unit[u]);
    return buf;
}

static char *centred(const char *text, char *buf, int width)
{
    int len = (int)strlen(text), pad = (width - len) / 2, odd = (width - len) % 2;
    if (pad >= 1) sprintf(buf, " %*s%*s", len + pad, text, pad - 1 + odd, " ");
    else sprintf(buf, "%s", text);
    return buf;
}

static void print_histogram(FILE *out, const hdr_t *h, const cstat_t *st, int tid, const uint32_t *hist, int n_bins, bool utf, bool plot_depth)
{
    const cstat_t *s = st + tid;
    const int rows = 10, steps = utf ? 8 : 2;
    const char *const *glyph = utf ? kBlocks8 : kBlocks2;
    const char *bar = utf ? "\xE2\x94\x82" : "|";
    double region_len = (double)(s->end - s->beg), top = 0.0;
    double *val = calloc((size_t)(n_bins > 0 ? n_bins : 1), sizeof(double));
    char b1[64], b2[64];
    int i, col;
    for (i = 0; i < n_bins; ++i) {
        val[i] = (uint32_t)((plot_depth ? 1u : 100u) * hist[i]) / (double)s->bin_width;     /* the product is formed in 32 bits */
        if (val[i] > top) top = val[i];
    }
    fprintf(out, "%s (%sbp)\n", h->name[tid], fmt_bp((double)h->len[tid], b1));
    const double step = top / rows;
    for (i = rows - 1; i >= 0; --i) {
        const double floor_ = step * i;
        if (plot_depth) fprintf(out, ">%8.1f ", i * step); else fprintf(out, ">%7.2f%% ", floor_);
        fputs(bar, out);
        for (col = 0; col < n_bins; ++col) {
            int g;
            if (step == 0.0) g = steps - 1;                           /* see the header: 0/0 in the reference */
            else {
                g = (int)round(steps * (val[col] - floor_) / step) - 1;
                if (g >= steps) g = steps - 1;
            }
            if (g < 0) fputc(' ', out); else fputs(glyph[g], out);
        }
        fputs(bar, out); fputc(' ', out);
        switch (i) {
        case 9: fprintf(out, "Number of reads: %u", s->n_selected_reads); break;
        case 8: if (s->n_reads - s->n_selected_reads > 0) fprintf(out, "    (%i filtered)", s->n_reads - s->n_selected_reads); break;
        case 7: fprintf(out, "Covered bases:   %sbp", fmt_bp((double)s->n_covered_bases, b1)); break;
        case 6: fprintf(out, "Percent covered: %.4g%%", 100.0 * s->n_covered_bases / region_len); break;
        case 5: fprintf(out, "Mean coverage:   %.3gx", s->summed_coverage / region_len); break;
        case 4: fprintf(out, "Mean baseQ:      %.3g", s->quality_bases > 0 ? s->summed_baseQ / (double)s->quality_bases : 0); break;
        case 3: fprintf(out, "Mean mapQ:       %.3g", s->summed_mapQ / (double)s->n_selected_reads); break;
        case 1: fprintf(out, "Histo bin width: %sbp", fmt_bp((double)s->bin_width, b1)); break;
        case 0: if (plot_depth) fprintf(out, "Histo max cov:   %.5g", top); else fprintf(out, "Histo max bin:   %.5g%%", top); break;
        default: break;
        }
        fputc('\n', out);
    }
    fprintf(out, "     %s", centred(fmt_bp((double)(s->beg + 1), b1), b2, 10));
    for (col = 10; col < 10 * (n_bins / 10); col += 10) fprintf(out, "%s", centred(fmt_bp((double)(s->beg + s->bin_width * col), b1), b2, 10));
    fprintf(out, "%*s%s", n_bins % 10, " ", centred(fmt_bp((double)s->end, b1), b2, 10));
    fputc('\n', out);
    free(val);
}

int main_coverage(int argc, char **argv)
{
    int max_depth = 1000000, min_baseQ = 0, min_mapQ = 0, min_len = 0, mindepth = 1;
    int fail_flags = F_UNMAP | F_SECONDARY | F_QCFAIL | F_DUP, required_flags = 0;
    char *opt_reg = NULL, *file_list = NULL, *out_fn = NULL, **fn = NULL;
    bool print_header = true, want_hist = false, want_table = true, utf = true, plot_depth = false, full_width = true;
    int n_bins_opt = 50;
    int c, i, j, n;
    static const struct option lo[] = {
        {"rf", 1, 0, 1}, {"ff", 1, 0, 2}, {"incl-flags", 1, 0, 1}, {"excl-flags", 1, 0, 2},
        {"bam-list", 1, 0, 'b'}, {"min-read-len", 1, 0, 'l'}, {"min-MQ", 1, 0, 'q'}, {"min-mq", 1, 0, 'q'},
        {"min-BQ", 1, 0, 'Q'}, {"min-bq", 1, 0, 'Q'}, {"output", 1, 0, 'o'}, {"no-header", 0, 0, 'H'},
        {"region", 1, 0, 'r'}, {"depth", 1, 0, 'd'}, {"min-depth", 1, 0, 3},
        {"histogram", 0, 0, 'm'}, {"ascii", 0, 0, 'A'}, {"plot-depth", 0, 0, 'D'}, {"n-bins", 1, 0, 'w'}, {0, 0, 0, 0}
    };
    optind = 1;
    while ((c = getopt_long(argc, argv, "Ao:l:q:Q:Hw:r:b:md:D", lo, NULL)) != -1) {
        switch (c) {
        case 1: if ((required_flags = parse_flag(optarg)) < 0) { fprintf(stderr, "Could not parse --rf %s\n", optarg); return 1; } break;
        case 2: if ((fail_flags = parse_flag(optarg)) < 0) { fprintf(stderr, "Could not parse --ff %s\n", optarg); return 1; } break;
        case 3: if ((i = atoi(optarg)) > 0) mindepth = i; break;
        case 'o': out_fn = optarg; full_width = false; break;
        case 'w': n_bins_opt = atoi(optarg); full_width = false; want_hist = true; want_table = false; break;
        case 'm': want_hist = true; want_table = false; break;
        case 'A': utf = false; want_hist = true; want_table = false; break;
        case 'D': want_hist = true; want_table = false; plot_depth = true; break;
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
    if (n_bins_opt <= 0 || full_width) {      /* terminal width - 40, at least 40 (coverage.c:437-461) */
        const char *ec = getenv("COLUMNS");
        int columns = 0;
        if (ec) columns = atoi(ec);
        else { struct winsize ws; if (ioctl(2, TIOCGWINSZ, &ws) == 0) columns = ws.ws_col; }
        n_bins_opt = columns > 60 ? columns - 40 : 40;
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
    int64_t n_bins = n_bins_opt, cur_bin = 0;
    uint32_t *hist = calloc((size_t)n_bins_opt + 1, sizeof(uint32_t));
    if (opt_reg) {
        cstat_t *s = stats + rtid;
        s->beg = rbeg; s->end = rend;
        if (s->end == HPOS_MAX) s->end = h->len[rtid];
        if (n_bins_opt > s->end - s->beg) n_bins = s->end - s->beg;
        s->bin_width = (s->end - s->beg) / (n_bins > 0 ? n_bins : 1);
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
            if (old_tid >= 0) {
                if (want_hist) { print_histogram(out, h, stats, old_tid, hist, (int)n_bins, utf, plot_depth); fputc('\n', out); memset(hist, 0, (size_t)n_bins * sizeof(uint32_t)); }
                else if (want_table) print_row(out, h, stats, old_tid, &print_header);
            }
            stats[tid].covered = true;
            if (!opt_reg) stats[tid].end = h->len[tid];
            if (want_hist) {
                n_bins = n_bins_opt > stats[tid].end - stats[tid].beg ? stats[tid].end - stats[tid].beg : n_bins_opt;
                stats[tid].bin_width = (stats[tid].end - stats[tid].beg) / n_bins;
            }
            old_tid = tid;
        }
        if (pos < stats[tid].beg || pos >= stats[tid].end) continue;
        if (tid >= n_targets) continue;
        if (want_hist) cur_bin = (pos - stats[tid].beg) / stats[tid].bin_width;
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
            if (cur_bin < n_bins && plot_depth) hist[cur_bin] += (uint32_t)d;
        }
        if (count_base && depth >= (unsigned long long)mindepth) {
            stats[tid].summed_coverage += depth;
            stats[tid].summed_baseQ += sbq;
            stats[tid].quality_bases += qb;
            stats[tid].n_covered_bases++;
            if (want_hist && cur_bin < n_bins && !plot_depth) ++hist[cur_bin];
        }
    }
    if (ret < 0) return 1;
    if (tid == -1 && opt_reg && *opt_reg != '*') tid = rtid;
    if (tid < n_targets && tid >= 0) {
        if (want_hist) print_histogram(out, h, stats, tid, hist, (int)n_bins, utf, plot_depth);
        else if (want_table) print_row(out, h, stats, tid, &print_header);
    }
    if (!opt_reg && want_table)
        for (i = 0; i < n_targets; ++i)
            if (!stats[i].covered) { stats[i].end = h->len[i]; print_row(out, h, stats, i, &print_header); }
    if (warn) fprintf(stderr, "samtools coverage: Warning:  Missing quality values in alignments.  Mean base quality calculated only on available values.\n");
    mplp_destroy(mplp);
    for (i = 0; i < n; i++) { reader_close(data[i]->rd); free(data[i]); }
    free(data); free(stats); free(n_plp); free(plp);
    if (out != stdout) fclose(out); else fflush(out);
    return 0;
}
