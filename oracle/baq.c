/*
 * oracle/baq.c -- TEST INFRASTRUCTURE ONLY (CPU oracle).
 *
 * Restates htslib 1.23.1 realn.c (sam_prob_realn, sam_cap_mapq) and
 * probaln.c (probaln_glocal).  Reference call sites: bam_plcmd.c:451 (flag 3
 * = APPLY|EXTEND, 7 adds REDO) and bam_plcmd.c:453.  Semantics: SURVEY.md
 * section 8a rows a2-a4 and Appendix A6.  Arithmetic is IEEE double evaluated in the
 * written association order; build WITHOUT FMA contraction (-ffp-contract=off)
 * to match the reference's plain x86-64 -O2 build (Makefile:29).
 * Pinned by test/mpileup/expected/{16,19,21,23,33,34}.out (tests/).
 * sam_cap_mapq (-C) is unpinned by any reference test.
 */
#include "plp.h"
#include <math.h>

#define BAQ_APPLY 1
#define BAQ_EXTEND 2
#define BAQ_REDO 4

#define EI .25
#define EM .33333333333

typedef struct { double d, e; int bw; } hmm_par_t;

static double g_q2p[256];

/* storage offset of band cell (i,k): 3 states per cell, first cell at 3 */
#define SET_U(u, b, i, k) { int x_ = (i) - (b); x_ = x_ > 0 ? x_ : 0; (u) = ((k) - x_ + 1) * 3; }

/* banded glocal pair-HMM; fills state[] (MAP ref offset<<2|state) and q[] */
static int glocal(const uint8_t *ref, int l_ref, const uint8_t *query, int l_query,
                  const uint8_t *iqual, const hmm_par_t *c, int *state, uint8_t *q)
{
    double *fm, *bm, *s, m[9], sI, sM, bI, bM;
    float *qual;
    int bw, bw2, i, k, Pr;
    size_t stride;

    if (l_ref <= 0 || l_query <= 0) return 0;
    bw = l_ref > l_query ? l_ref : l_query;
    if (bw > c->bw) bw = c->bw;
    if (bw < abs(l_ref - l_query)) bw = abs(l_ref - l_query);
    bw2 = bw * 2 + 1;
    stride = (size_t)bw2 * 3 + 6;
    fm = calloc((size_t)(l_query + 1) * stride, sizeof(double));
    bm = calloc((size_t)(l_query + 1) * stride, sizeof(double));
    s = calloc((size_t)l_query + 2, sizeof(double));
    qual = calloc((size_t)l_query, sizeof(float));
    if (g_q2p[0] == 0)
        for (i = 0; i < 256; ++i) g_q2p[i] = pow(10, -i / 10.);
    for (i = 0; i < l_query; ++i) qual[i] = (float)g_q2p[iqual ? iqual[i] : 30];
#define F(i) (fm + (size_t)(i) * stride)
#define B(i) (bm + (size_t)(i) * stride)
    /* transitions */
    sM = sI = 1. / (2 * l_query + 2);
    m[0*3+0] = (1 - c->d - c->d) * (1 - sM); m[0*3+1] = m[0*3+2] = c->d * (1 - sM);
    m[1*3+0] = (1 - c->e) * (1 - sI); m[1*3+1] = c->e * (1 - sI); m[1*3+2] = 0.;
    m[2*3+0] = 1 - c->e; m[2*3+1] = 0.; m[2*3+2] = c->e;
    bM = (1 - c->d) / l_ref; bI = c->d / l_ref;
    /*** forward ***/
    SET_U(k, bw, 0, 0);
    F(0)[k] = s[0] = 1.;
    { /* row 1 */
        double *fi = F(1), sum;
        int beg = 1, end = l_ref < bw + 1 ? l_ref : bw + 1, _beg, _end;
        for (k = beg, sum = 0.; k <= end; ++k) {
            int u;
            double e = (ref[k - 1] > 3 || query[0] > 3) ? 1. : ref[k - 1] == query[0] ? 1. - qual[0] : qual[0] * EM;
            SET_U(u, bw, 1, k);
            fi[u + 0] = e * bM; fi[u + 1] = EI * bI;
            sum += fi[u] + fi[u + 1];
        }
        s[1] = sum;
        SET_U(_beg, bw, 1, beg); SET_U(_end, bw, 1, end); _end += 2;
        for (k = _beg; k <= _end; ++k) fi[k] /= sum;
    }
    for (i = 2; i <= l_query; ++i) {
        double *fi = F(i), *fi1 = F(i - 1), sum, qli = qual[i - 1];
        int beg = 1, end = l_ref, x, _beg, _end;
        uint8_t qyi = query[i - 1];
        x = i - bw; beg = beg > x ? beg : x;
        x = i + bw; end = end < x ? end : x;
        for (k = beg, sum = 0.; k <= end; ++k) {
            int u, v11, v01, v10;
            double e;
            e = (ref[k - 1] > 3 || qyi > 3) ? 1. : ref[k - 1] == qyi ? 1. - qli : qli * EM;
            SET_U(u, bw, i, k); SET_U(v11, bw, i - 1, k - 1); SET_U(v10, bw, i - 1, k); SET_U(v01, bw, i, k - 1);
            fi[u + 0] = e * (m[0] * fi1[v11 + 0] + m[3] * fi1[v11 + 1] + m[6] * fi1[v11 + 2]);
            fi[u + 1] = EI * (m[1] * fi1[v10 + 0] + m[4] * fi1[v10 + 1]);
            fi[u + 2] = m[2] * fi[v01 + 0] + m[8] * fi[v01 + 2];
            sum += fi[u] + fi[u + 1] + fi[u + 2];
        }
        s[i] = sum;
        SET_U(_beg, bw, i, beg); SET_U(_end, bw, i, end); _end += 2;
        for (k = _beg, sum = 1. / sum; k <= _end; ++k) fi[k] *= sum;
    }
    { /* termination */
        double sum;
        for (k = 1, sum = 0.; k <= l_ref; ++k) {
            int u;
            SET_U(u, bw, l_query, k);
            if (u < 3 || u >= bw2 * 3 + 3) continue;
            sum += F(l_query)[u + 0] * sM + F(l_query)[u + 1] * sI;
        }
        s[l_query + 1] = sum;
    }
    { /* likelihood */
        double p = 1., Pr1 = 0.;
        for (i = 0; i <= l_query + 1; ++i) {
            p *= s[i];
            if (p < 1e-100) Pr1 += -4.343 * log(p), p = 1.;
        }
        Pr1 += -4.343 * log(p * l_ref * l_query);
        Pr = (int)(Pr1 + .499);
    }
    /*** backward ***/
    for (k = 1; k <= l_ref; ++k) {
        int u;
        double *bi = B(l_query);
        SET_U(u, bw, l_query, k);
        if (u < 3 || u >= bw2 * 3 + 3) continue;
        bi[u + 0] = sM / s[l_query] / s[l_query + 1]; bi[u + 1] = sI / s[l_query] / s[l_query + 1];
    }
    for (i = l_query - 1; i >= 1; --i) {
        int beg = 1, end = l_ref, x, _beg, _end;
        double *bi = B(i), *bi1 = B(i + 1), y = (i > 1), qli1 = qual[i];
        uint8_t qyi1 = query[i];
        x = i - bw; beg = beg > x ? beg : x;
        x = i + bw; end = end < x ? end : x;
        for (k = end; k >= beg; --k) {
            int u, v11, v01, v10;
            double e;
            SET_U(u, bw, i, k); SET_U(v11, bw, i + 1, k + 1); SET_U(v10, bw, i + 1, k); SET_U(v01, bw, i, k + 1);
            e = (k >= l_ref ? 0 : (ref[k] > 3 || qyi1 > 3) ? 1. : ref[k] == qyi1 ? 1. - qli1 : qli1 * EM) * bi1[v11];
            bi[u + 0] = e * m[0] + EI * m[1] * bi1[v10 + 1] + m[2] * bi[v01 + 2];
            bi[u + 1] = e * m[3] + EI * m[4] * bi1[v10 + 1];
            bi[u + 2] = (e * m[6] + m[8] * bi[v01 + 2]) * y;
        }
        SET_U(_beg, bw, i, beg); SET_U(_end, bw, i, end); _end += 2;
        for (k = _beg, y = 1. / s[i]; k <= _end; ++k) bi[k] *= y;
    }
    /*** MAP ***/
    for (i = 1; i <= l_query; ++i) {
        double sum = 0., *fi = F(i), *bi = B(i), max = 0.;
        int beg = 1, end = l_ref, x, max_k = -1;
        x = i - bw; beg = beg > x ? beg : x;
        x = i + bw; end = end < x ? end : x;
        for (k = beg; k <= end; ++k) {
            int u;
            double z;
            SET_U(u, bw, i, k);
            z = fi[u + 0] * bi[u + 0]; if (z > max) max = z, max_k = (k - 1) << 2 | 0; sum += z;
            z = fi[u + 1] * bi[u + 1]; if (z > max) max = z, max_k = (k - 1) << 2 | 1; sum += z;
        }
        max /= sum;
        if (state) state[i - 1] = max_k;
        if (q) {
            double v = -4.343 * log(1. - max) + .499;
            /* x86-64 cvttsd2si of inf/NaN yields INT_MIN; keep that behaviour */
            k = (v == v && v < 2147483648.0 && v > -2147483649.0) ? (int)v : (-2147483647 - 1);
            q[i - 1] = (uint8_t)(k > 100 ? 99 : k);
        }
    }
    free(fm); free(bm); free(s); free(qual);
    return Pr;
#undef F
#undef B
}

int baq_realn(rec_t *b, const char *ref, hpos_t ref_len, int flag)
{
    int k, bw, y, yb, ye, apply_baq = flag & BAQ_APPLY, extend_baq = flag & BAQ_EXTEND, redo_baq = flag & BAQ_REDO;
    hpos_t i, x, xb, xe;
    const uint32_t *cigar = b->cigar;
    hmm_par_t conf = { 0.001, 0.1, 10 };
    const uint8_t *bqt, *zqt;
    uint8_t *qual = b->qual;

    if ((b->flag & F_UNMAP) || b->l_qseq == 0 || qual[0] == 0xff) return -1;
    bqt = rec_aux_get(b, "BQ"); if (bqt && *bqt != 'Z') bqt = NULL;
    zqt = rec_aux_get(b, "ZQ"); if (zqt && *zqt != 'Z') zqt = NULL;
    if (bqt && redo_baq) { rec_aux_del(b, bqt); bqt = NULL; zqt = rec_aux_get(b, "ZQ"); if (zqt && *zqt != 'Z') zqt = NULL; }
    if (bqt && zqt) { rec_aux_del(b, zqt); zqt = NULL; bqt = rec_aux_get(b, "BQ"); }
    if (bqt || zqt) {
        if ((apply_baq && zqt) || (!apply_baq && bqt)) return -3;
        if (bqt && apply_baq) {
            const uint8_t *bq = bqt + 1;
            for (i = 0; i < b->l_qseq; ++i)
                qual[i] = qual[i] + 64 < bq[i] ? 0 : (uint8_t)(qual[i] - ((int)bq[i] - 64));
            ((uint8_t *)bqt)[-2] = 'Z'; /* BQ -> ZQ */
        } else if (zqt && !apply_baq) {
            const uint8_t *zq = zqt + 1;
            for (i = 0; i < b->l_qseq; ++i) qual[i] = (uint8_t)(qual[i] + ((int)zq[i] - 64));
            ((uint8_t *)zqt)[-2] = 'B';
        }
        return 0;
    }
    /* aligned span */
    x = b->pos; y = 0; yb = ye = -1; xb = xe = -1;
    for (k = 0; k < (int)b->n_cigar; ++k) {
        int op = cop(cigar[k]), l = (int)cln(cigar[k]);
        if (op == C_M || op == C_EQ || op == C_X) {
            if (yb < 0) yb = y;
            if (xb < 0) xb = x;
            ye = y + l; xe = x + l;
            x += l; y += l;
        } else if (op == C_S || op == C_I) y += l;
        else if (op == C_D) x += l;
        else if (op == C_N) return -1;
    }
    if (xb == -1) return -1;
    bw = 7;
    if (llabs((xe - xb) - (ye - yb)) > bw) bw = (int)llabs((xe - xb) - (ye - yb)) + 3;
    conf.bw = bw;
    xb -= yb + bw / 2; if (xb < 0) xb = 0;
    xe += b->l_qseq - ye + bw / 2;
    if (xe - xb - b->l_qseq > bw) {
        xb += (xe - xb - b->l_qseq - bw) / 2; xe -= (xe - xb - b->l_qseq - bw) / 2;
    }
    {
        int l = b->l_qseq;
        size_t lref = xe > xb ? (size_t)(xe - xb) : 1;
        if (lref < (size_t)l) lref = (size_t)l;
        uint8_t *bq = malloc((size_t)l + 1), *tseq = malloc((size_t)l + lref + 16), *tref = malloc(lref + (size_t)l + 16), *q = malloc((size_t)l + 1);
        int *state = malloc(sizeof(int) * (size_t)l);
        memcpy(bq, qual, (size_t)l);
        for (i = 0; i < l; ++i) tseq[i] = (uint8_t)nt16_int[seqi(b->seq, i)];
        for (i = xb; i < xe; ++i) {
            if (i >= ref_len || ref[i] == '\0') { xe = i; break; }
            tref[i - xb] = (uint8_t)nt16_int[nt16_table[(unsigned char)ref[i]]];
        }
        glocal(tref, (int)(xe - xb), tseq, l, qual, &conf, state, q);
        if (xe - xb <= 0) { /* nothing aligned: glocal returned early, leave untouched */
            free(bq); free(tseq); free(tref); free(q); free(state);
            return 0;
        }
        if (!extend_baq) {
            for (k = 0, x = b->pos, y = 0; k < (int)b->n_cigar; ++k) {
                int op = cop(cigar[k]), len = (int)cln(cigar[k]);
                if (op == C_M || op == C_EQ || op == C_X) {
                    if (len > l - y) len = l - y;
                    for (i = y; i < y + len; ++i) {
                        if ((state[i] & 3) != 0 || state[i] >> 2 != x - xb + (i - y)) bq[i] = 0;
                        else bq[i] = bq[i] < q[i] ? bq[i] : q[i];
                    }
                    x += len; y += len;
                } else if (op == C_S || op == C_I) { if (len > l - y) len = l - y; y += len; }
                else if (op == C_D) x += len;
            }
            for (i = 0; i < l; ++i) bq[i] = (uint8_t)(qual[i] - bq[i] + 64);
        } else {
            uint8_t *left = tseq, *rght = tref;
            for (k = 0, x = b->pos, y = 0; k < (int)b->n_cigar; ++k) {
                int op = cop(cigar[k]), len = (int)cln(cigar[k]);
                if (op == C_M || op == C_EQ || op == C_X) {
                    if (len > l - y) len = l - y;
                    if (len > 0) {
                        for (i = y; i < y + len; ++i)
                            bq[i] = ((state[i] & 3) != 0 || state[i] >> 2 != x - xb + (i - y)) ? 0 : q[i];
                        for (left[y] = bq[y], i = y + 1; i < y + len; ++i)
                            left[i] = bq[i] > left[i - 1] ? bq[i] : left[i - 1];
                        for (rght[y + len - 1] = bq[y + len - 1], i = y + len - 2; i >= y; --i)
                            rght[i] = bq[i] > rght[i + 1] ? bq[i] : rght[i + 1];
                        for (i = y; i < y + len; ++i) bq[i] = left[i] < rght[i] ? left[i] : rght[i];
                    }
                    x += len; y += len;
                } else if (op == C_S || op == C_I) { if (len > l - y) len = l - y; y += len; }
                else if (op == C_D) x += len;
            }
            for (i = 0; i < l; ++i) bq[i] = (uint8_t)(64 + (qual[i] <= bq[i] ? 0 : qual[i] - bq[i]));
        }
        if (apply_baq) {
            for (i = 0; i < l; ++i) qual[i] = (uint8_t)(qual[i] - (bq[i] - 64));
            bq[l] = 0;
            rec_aux_append(b, "ZQ", 'Z', l + 1, bq);
        } else { bq[l] = 0; rec_aux_append(b, "BQ", 'Z', l + 1, bq); }
        free(bq); free(tseq); free(tref); free(q); free(state);
    }
    return 0;
}

/* sam_cap_mapq: formula in doc/samtools-mpileup.1:219-241 (unpinned) */
int cap_mapq(const rec_t *b, const char *ref, hpos_t ref_len, int thres)
{
    const uint8_t *seq = b->seq, *qual = b->qual;
    int i, y, mm, q, len, clip_l, clip_q;
    hpos_t x;
    double t;
    if (thres < 0) thres = 40;
    mm = q = len = clip_l = clip_q = 0;
    for (i = y = 0, x = b->pos; i < (int)b->n_cigar; ++i) {
        int j, l = (int)cln(b->cigar[i]), op = cop(b->cigar[i]);
        if (op == C_M || op == C_EQ || op == C_X) {
            for (j = 0; j < l; ++j) {
                int c1, c2, z = y + j;
                if (x + j >= ref_len || ref[x + j] == '\0') break;
                c1 = seqi(seq, z); c2 = nt16_table[(unsigned char)ref[x + j]];
                if (c2 != 15 && c1 != 15 && qual[z] >= 13) {
                    ++len;
                    if (c1 && c1 != c2 && qual[z] >= 13) { ++mm; q += qual[z] > 33 ? 33 : qual[z]; }
                }
            }
            if (j < l) break;
            x += l; y += l; len += l;
        } else if (op == C_D) {
            for (j = 0; j < l; ++j) if (x + j >= ref_len || ref[x + j] == '\0') break;
            if (j < l) break;
            x += l;
        } else if (op == C_S) {
            for (j = 0; j < l; ++j) clip_q += qual[y + j];
            clip_l += l; y += l;
        } else if (op == C_H) { clip_q += 13 * l; clip_l += l; }
        else if (op == C_I) y += l;
        else if (op == C_N) x += l;
    }
    for (i = 0, t = 1; i < mm; ++i) t *= (double)len / (i + 1);
    t = q - 4.343 * log(t) + clip_q / 5.;
    if (t > thres) return -1;
    if (t < 0) t = 0;
    t = sqrt((thres - t) / thres) * thres;
    return (int)(t + .499);
}
