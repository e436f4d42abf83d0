/*
 * oracle/errmod.c -- TEST INFRASTRUCTURE ONLY (CPU oracle).
 *
 * Restates htslib 1.23.1 errmod.c (errmod_init / errmod_cal), the sampling
 * helpers it uses (ksort.h ks_shuffle over hts_drand48, ascending sort) and
 * the in-tree front end bam2bcf.c:65-123 (bcf_call_glfgen).  Semantics:
 * SURVEY.md section 8a rows a16-a17, Appendix A7.
 *
 * PARITY UNPINNED: no reference test holds errmod_cal / glfgen numbers (only
 * the tview consensus letters, test/large_pos/tview.expected.out:3).
 */
#include "plp.h"
#include <math.h>

#define ETA 0.03

struct errmod_t { double depcorr; double *fk, *beta, *lhet; };

errmod_t *errmod_new(double depcorr)
{
    errmod_t *em = calloc(1, sizeof(*em));
    int k, n, q;
    double *lC;
    em->depcorr = depcorr;
    em->fk = calloc(256, sizeof(double));
    em->fk[0] = 1.0;
    for (n = 1; n < 256; ++n) em->fk[n] = pow(1. - depcorr, n) * (1.0 - ETA) + ETA;
    em->beta = calloc(256 * 256 * 64, sizeof(double));
    lC = calloc(256 * 256, sizeof(double));
    for (n = 1; n != 256; ++n) {
        double lgn = lgamma(n + 1);
        for (k = 1; k <= n; ++k) lC[n << 8 | k] = lgn - lgamma(k + 1) - lgamma(n - k + 1);
    }
    for (q = 1; q != 64; ++q) {
        double e = pow(10.0, -q / 10.0);
        double le = log(e);
        double le1 = log(1.0 - e);
        for (n = 1; n <= 255; ++n) {
            /* binomial tail ratio in LOG space, long double accumulators (htslib errmod.c cal_coef): the plain-space
             * running sums underflow to 0/0 = NaN for the deep k of a high-quality column (n = 167, q = 40, k >= 100) */
            double *beta = em->beta + (q << 16 | n << 8);
            long double sum, sum1;
            sum1 = lC[n << 8 | n] + n * le;
            beta[n] = HUGE_VAL;
            for (k = n - 1; k >= 0; --k, sum1 = sum) {
                sum = sum1 + log1pl(expl(lC[n << 8 | k] + k * le + (n - k) * le1 - sum1));
                beta[k] = -10. / M_LN10 * (sum1 - sum);
            }
        }
    }
    em->lhet = calloc(256 * 256, sizeof(double));
    for (n = 0; n < 256; ++n)
        for (k = 0; k < 256; ++k) em->lhet[n << 8 | k] = lC[n << 8 | k] - M_LN2 * n;
    free(lC);
    return em;
}
void errmod_free(errmod_t *em) { if (em) { free(em->fk); free(em->beta); free(em->lhet); free(em); } }

/* 48-bit LCG of drand48 (hts_os.c hts_drand48), default seed as srand48(0) */
static uint64_t g_rs = 0x330EULL;
void hts_srand48_(long seed) { g_rs = (((uint64_t)seed & 0xffffffffULL) << 16) | 0x330EULL; }
double hts_drand48_(void)
{
    g_rs = (g_rs * 0x5DEECE66DULL + 0xBULL) & 0xffffffffffffULL;
    return (double)g_rs / 281474976710656.0;
}
static int cmp_u16(const void *a, const void *b) { return (int)*(const uint16_t *)a - (int)*(const uint16_t *)b; }

int errmod_calc(const errmod_t *em, int n, int m, uint16_t *bases, float *q)
{
    double fsum[16], bsum[16];
    uint32_t c[16];
    int i, j, k, w[32];
    memset(q, 0, (size_t)(m * m) * sizeof(float));
    if (n == 0) return 0;
    if (n > 255) { /* ks_shuffle then keep the first 255 */
        for (i = n; i > 1; --i) {
            uint16_t tmp;
            j = (int)(hts_drand48_() * i);
            tmp = bases[j]; bases[j] = bases[i - 1]; bases[i - 1] = tmp;
        }
        n = 255;
    }
    qsort(bases, (size_t)n, 2, cmp_u16);
    memset(w, 0, sizeof w); memset(fsum, 0, sizeof fsum); memset(bsum, 0, sizeof bsum); memset(c, 0, sizeof c);
    for (j = n - 1; j >= 0; --j) {
        uint16_t b = bases[j];
        int qual = b >> 5 < 4 ? 4 : b >> 5;
        if (qual > 63) qual = 63;
        int basestrand = b & 0x1f, base = b & 0xf;
        fsum[base] += em->fk[w[basestrand]];
        bsum[base] += em->fk[w[basestrand]] * em->beta[qual << 16 | n << 8 | c[base]];
        ++c[base]; ++w[basestrand];
    }
    for (j = 0; j < m; ++j) {
        float tmp1, tmp3; int tmp2;
        for (k = 0, tmp1 = tmp3 = 0.0, tmp2 = 0; k < m; ++k) {
            if (k == j) continue;
            tmp1 += bsum[k]; tmp2 += c[k]; tmp3 += fsum[k];
        }
        if (tmp2) q[j * m + j] = tmp1;
        for (k = j + 1; k < m; ++k) {
            int cjk = c[j] + c[k];
            for (i = 0, tmp2 = 0, tmp1 = tmp3 = 0.0; i < m; ++i) {
                if (i == j || i == k) continue;
                tmp1 += bsum[i]; tmp2 += c[i]; tmp3 += fsum[i];
            }
            if (tmp2) q[j * m + k] = q[k * m + j] = -4.343 * em->lhet[cjk << 8 | c[k]] + tmp1;
            else q[j * m + k] = q[k * m + j] = -4.343 * em->lhet[cjk << 8 | c[k]];
        }
        for (k = 0; k < m; ++k) if (q[j * m + k] < 0.0) q[j * m + k] = 0.0;
    }
    return 0;
}

/* bcf_call_glfgen for SNP columns (bam2bcf.c:65-123; ref_base4 >= 0) */
int glfgen(int _n, const pile1_t *pl, int ref_base4, int min_baseQ, int capQ,
           const errmod_t *em, float qsum[4], float p25[25])
{
    int i, n;
    memset(qsum, 0, sizeof(float) * 4);
    memset(p25, 0, sizeof(float) * 25);
    if (_n <= 0) return -1;
    uint16_t *bases = malloc(2 * (size_t)_n);
    for (i = n = 0; i < _n; ++i) {
        const pile1_t *p = pl + i;
        int q, b, mapQ;
        if (p->is_del || p->is_refskip || (p->b->flag & F_UNMAP)) continue;
        mapQ = p->b->mapq < 255 ? p->b->mapq : 20;
        q = p->qpos < p->b->l_qseq ? (int)p->b->qual[p->qpos] : 0;
        if (q < min_baseQ) continue;
        if (q > 99) q = 99;
        mapQ = mapQ < capQ ? mapQ : capQ;
        if (q > mapQ) q = mapQ;
        if (q > 63) q = 63;
        if (q < 4) q = 4;
        if (p->qpos < p->b->l_qseq) {
            b = seqi(p->b->seq, p->qpos);
            b = nt16_int[b ? b : ref_base4];
        } else b = 4;
        bases[n++] = (uint16_t)(q << 5 | ((p->b->flag & F_REVERSE) ? 1 : 0) << 4 | b);
        if (b < 4) qsum[b] += q;
    }
    errmod_calc(em, n, 5, bases, p25);
    free(bases);
    return n;
}
