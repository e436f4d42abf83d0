/*
 * oracle/bedcov.c -- TEST INFRASTRUCTURE ONLY (CPU oracle).
 *
 * Restates the reference's `samtools bedcov`, bedcov.c: read callback read_bam (:54-70), header line output_header (:84-124),
 * per BED line a fresh multi-file pileup over [beg,end) with the column reducers of :303-331 (sum of per-column depth,
 * optionally without deletions / reference skips (-j), bases at or above a depth threshold (-d), reads pushed (-c)).
 * The reference queries a BAI index per line (sam_itr_queryi); here every line re-opens the file and scans it, which is
 * result-identical.  Pinned by the four .expected files of test/bedcov (test/test.pl:3817-3866).
 */
#include "plp.h"
#include <ctype.h>
#include <getopt.h>
#include <limits.h>
#include <zlib.h>

typedef struct { reader_t *rd; int min_mapQ; uint32_t flags; long long rcnt; } baux_t;

static int read_bam(void *data, rec_t *b)
{
    baux_t *a = data;
    int ret;
    for (;;) {
        if ((ret = reader_next(a->rd, b)) < 0) break;
        if (b->flag & a->flags) continue;
        if ((int)b->mapq < a->min_mapQ) continue;
        break;
    }
    /* bedcov counts reads through the iterator's constructor hook, which bam_plp_push runs for every mapped read it buffers */
    if (ret >= 0 && b->tid >= 0 && !(b->flag & F_UNMAP)) a->rcnt++;
    return ret;
}

static void output_header(FILE *fp, const char *hdr, int fields, int n, char **fn, int depth, int rcount)
{
    static const char *bedcols[] = { "chrom", "chromStart", "chromEnd", "name", "score", "strand", "thickStart", "thickEnd", "itemRgb",
                                     "blockCount", "blockSizes", "blockStarts" };
    int i;
    if (hdr) fprintf(fp, "%s", hdr);
    else for (i = 0; i < fields; ++i) fprintf(fp, "%s%s", i ? "\t" : "#", i < 12 ? bedcols[i] : ".");
    for (i = 0; i < n; ++i) fprintf(fp, "\t%s_cov", fn[i]);
    if (depth >= 0) for (i = 0; i < n; ++i) fprintf(fp, "\t%s_depth", fn[i]);
    if (rcount) for (i = 0; i < n; ++i) fprintf(fp, "\t%s_count", fn[i]);
    fprintf(fp, "\n");
}

int main_bedcov(int argc, char **argv)
{
    int c, i, j, n, min_mapQ = 0, skip_DN = 0, do_rcount = 0, min_depth = -1, max_depth = INT_MAX, print_header = 0, hdr = 0, status = 0, tflags;
    uint32_t flags = F_UNMAP | F_SECONDARY | F_QCFAIL | F_DUP;
    static const struct option lo[] = { {"min-MQ", 1, 0, 'Q'}, {"min-mq", 1, 0, 'Q'}, {"max-depth", 1, 0, 1000}, {0, 0, 0, 0} };
    optind = 1;
    while ((c = getopt_long(argc, argv, "Q:Xg:G:jd:Hc", lo, NULL)) >= 0) {
        switch (c) {
        case 'Q': min_mapQ = atoi(optarg); break;
        case 'X': break;
        case 'c': do_rcount = 1; break;
        case 'H': print_header = 1; break;
        case 'g': tflags = parse_flag(optarg); if (tflags < 0 || tflags > 4095) { fprintf(stderr, "[bedcov] Flag value \"%s\" is not supported\n", optarg); return 1; } flags &= ~(uint32_t)tflags; break;
        case 'G': tflags = parse_flag(optarg); if (tflags < 0 || tflags > 4095) { fprintf(stderr, "[bedcov] Flag value \"%s\" is not supported\n", optarg); return 1; } flags |= (uint32_t)tflags; break;
        case 'j': skip_DN = 1; break;
        case 'd': min_depth = atoi(optarg); break;
        case 1000: max_depth = atoi(optarg); break;
        default: fprintf(stderr, "Usage: samtools bedcov [options] <in.bed> <in1.bam> [...]\n"); return 1;
        }
    }
    if (optind + 2 > argc) { fprintf(stderr, "Usage: samtools bedcov [options] <in.bed> <in1.bam> [...]\n"); return 1; }
    n = argc - optind - 1;
    char **fn = argv + optind + 1;
    if (!print_header) hdr = 1;
    gzFile fp = gzopen(argv[optind], "rb");
    if (!fp) { fprintf(stderr, "[bedcov] can't open BED file '%s'\n", argv[optind]); return 2; }
    reader_t *r0 = reader_open(fn[0], NULL);
    if (!r0) { fprintf(stderr, "ERROR: fail to open index BAM file '%s'\n", fn[0]); return 2; }
    hdr_t *h0 = reader_hdr(r0);
    long long *cnt = calloc((size_t)n, sizeof *cnt), *pcov = calloc((size_t)n, sizeof *pcov);
    int *n_plp = calloc((size_t)n, sizeof(int));
    const pile1_t **plp = calloc((size_t)n, sizeof *plp);
    static char line[1 << 16];
    while (gzgets(fp, line, sizeof line)) {
        size_t l = strlen(line);
        while (l && (line[l - 1] == '\n' || line[l - 1] == '\r')) line[--l] = 0;
        if (l == 0) continue;
        if (line[0] == '#') {
            if (!hdr && !strncmp(line, "#chrom", 6)) { output_header(stdout, line, -1, n, fn, min_depth, do_rcount); hdr = 1; }
            continue;
        }
        if (!strncmp(line, "track ", 6) || !strncmp(line, "browser ", 8)) continue;
        if (!hdr) {
            int fields = 0; const char *t = line;
            while (*t) if (*t++ == '\t') fields++;
            output_header(stdout, NULL, fields + 1, n, fn, min_depth, do_rcount);
            hdr = 1;
        }
        char *p = line, *q = line; int tid; long long beg = 0, end = 0;
        while (*p && !isspace((unsigned char)*p)) ++p;
        if (*p == 0) goto bed_error;
        { char sv = *p; *p = 0; tid = hdr_name2tid(h0, q); *p = sv; }
        if (tid < 0) goto bed_error;
        if (sscanf(p + 1, "%lld %lld", &beg, &end) < 2 || end < beg) goto bed_error;
        {
            baux_t *aux = calloc((size_t)n, sizeof *aux); void **data = calloc((size_t)n, sizeof *data);
            char reg[1200];
            snprintf(reg, sizeof reg, "%s:%lld-%lld", h0->name[tid], beg + 1, end);
            for (i = 0; i < n; ++i) {
                int t2; hpos_t b2, e2;
                aux[i].rd = reader_open(fn[i], NULL); aux[i].min_mapQ = min_mapQ; aux[i].flags = flags; aux[i].rcnt = 0;
                if (!aux[i].rd) { fprintf(stderr, "ERROR: fail to open index BAM file '%s'\n", fn[i]); return 2; }
                if (end > beg) reader_set_region(aux[i].rd, reg, &t2, &b2, &e2);
                data[i] = &aux[i];
            }
            mplp_t *mp = mplp_init(n, read_bam, data);
            mplp_set_maxcnt(mp, min_depth > max_depth ? min_depth : max_depth);
            memset(cnt, 0, sizeof *cnt * (size_t)n); memset(pcov, 0, sizeof *pcov * (size_t)n);
            int ret, t3; hpos_t pos;
            while (end > beg && (ret = mplp_auto(mp, &t3, &pos, n_plp, plp)) > 0)
                if (t3 == tid && pos >= beg && pos < end)
                    for (i = 0; i < n; ++i) {
                        int m = 0;
                        if (skip_DN || min_depth >= 0) for (j = 0; j < n_plp[i]; ++j) if (plp[i][j].is_del || plp[i][j].is_refskip) ++m;
                        int pd = n_plp[i] - m;
                        cnt[i] += pd;
                        if (min_depth >= 0 && pd >= min_depth) pcov[i]++;
                    }
            fputs(line, stdout);
            for (i = 0; i < n; ++i) printf("\t%lld", cnt[i]);
            if (min_depth >= 0) for (i = 0; i < n; ++i) printf("\t%lld", pcov[i]);
            if (do_rcount) for (i = 0; i < n; ++i) printf("\t%lld", aux[i].rcnt);
            putchar('\n');
            mplp_destroy(mp);
            for (i = 0; i < n; ++i) reader_close(aux[i].rd);
            free(aux); free(data);
        }
        continue;
bed_error:
        fprintf(stderr, "Errors in BED line '%s'\n", line);
        status = 2;
    }
    gzclose(fp); reader_close(r0);
    free(cnt); free(pcov); free(n_plp); free(plp);
    return status;
}
