// read_ops_check.cpp -- TEST CLIENT of the per-read / per-column htslib entry points exported by libb200pileup.so
// (include/b200_htslib_compat.h: sam_prob_realn, sam_cap_mapq, errmod_cal, bcf_call_glfgen, bam_plp_insertion_mod).
// Every result is compared in-process with the CPU oracle's restatement (oracle/baq.c, oracle/errmod.c), which this
// test -- not the product -- links.
//   read_ops_check in.sam ref.fa [max_reads]     exit 0 when everything agrees
#include "b200_htslib_compat.h"
extern "C" {
#include "../../oracle/plp.h"
}
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

static void to_bam1(const rec_t *r, bam1_t *b)
{
    const size_t lq = strlen(r->qname) + 1, l_qname = (lq + 3) & ~(size_t)3;
    const size_t need = l_qname + 4 * (size_t)r->n_cigar + (size_t)(r->l_qseq + 1) / 2 + (size_t)r->l_qseq + (size_t)r->l_aux;
    b->data = (uint8_t *)realloc(b->data, need ? need : 1); b->m_data = (uint32_t)need;
    memset(b->data, 0, l_qname); memcpy(b->data, r->qname, lq);
    uint8_t *p = b->data + l_qname;
    memcpy(p, r->cigar, 4 * (size_t)r->n_cigar); p += 4 * (size_t)r->n_cigar;
    memcpy(p, r->seq, (size_t)(r->l_qseq + 1) / 2); p += (size_t)(r->l_qseq + 1) / 2;
    memcpy(p, r->qual, (size_t)r->l_qseq); p += (size_t)r->l_qseq;
    if (r->l_aux) memcpy(p, r->aux, (size_t)r->l_aux);
    b->l_data = (int)need;
    b->core.pos = r->pos; b->core.tid = r->tid; b->core.qual = r->mapq; b->core.flag = r->flag; b->core.l_qname = (uint16_t)l_qname;
    b->core.l_extranul = (uint8_t)(l_qname - lq); b->core.n_cigar = r->n_cigar; b->core.l_qseq = r->l_qseq;
    b->core.mtid = r->mtid; b->core.mpos = r->mpos; b->core.isize = r->isize; b->core.bin = 0;
}

static long g_bad = 0;
#define FAIL(...) do { if (++g_bad <= 10) { fprintf(stderr, __VA_ARGS__); fputc('\n', stderr); } } while (0)

int main(int argc, char **argv)
{
    if (argc < 3) { fprintf(stderr, "usage: read_ops_check in.sam ref.fa [max_reads]\n"); return 2; }
    const long max_reads = argc > 3 ? atol(argv[3]) : 200;
    reader_t *rd = reader_open(argv[1], NULL);
    fasta_t *fa = fasta_load(argv[2]);
    if (!rd || !fa) { fprintf(stderr, "cannot open inputs\n"); return 2; }
    hdr_t *h = reader_hdr(rd);
    std::vector<rec_t> recs;
    {
        rec_t r; rec_init(&r);
        while ((long)recs.size() < max_reads && reader_next(rd, &r) >= 0) {
            if (r.tid < 0 || fasta_find(fa, h->name[r.tid]) < 0) continue;
            rec_t c; rec_init(&c); rec_copy(&c, &r); recs.push_back(c);
        }
    }
    long n_realn = 0, n_cap = 0, n_em = 0, n_gl = 0;
    // ---- sam_prob_realn (flags 3 = mpileup default, 1 / 0 / 2 = the calmd variants) and sam_cap_mapq
    const int flags[4] = {3, 1, 0, 2};
    for (size_t i = 0; i < recs.size(); ++i) {
        const int fi = fasta_find(fa, h->name[recs[i].tid]);
        const char *ref = fa->seq[fi]; const hpos_t ref_len = fa->len[fi];
        for (int f = 0; f < 4; ++f) {
            if (f > 0 && i % 4 != 0) continue;              // the calmd variants on a quarter of the reads
            rec_t want; rec_init(&want); rec_copy(&want, &recs[i]);
            const int rc_want = baq_realn(&want, ref, ref_len, flags[f]);
            bam1_t *b = bam_init1(); to_bam1(&recs[i], b);
            const int rc_got = sam_prob_realn(b, ref, ref_len, flags[f]);
            ++n_realn;
            if (rc_got != rc_want) FAIL("sam_prob_realn(%s, flag %d): rc %d, oracle %d", recs[i].qname, flags[f], rc_got, rc_want);
            else {
                if (memcmp(bam_get_qual(b), want.qual, (size_t)want.l_qseq) != 0) FAIL("sam_prob_realn(%s, flag %d): qualities differ", recs[i].qname, flags[f]);
                const uint8_t *aux = bam_get_qual(b) + b->core.l_qseq; const int l_aux = (int)(b->data + b->l_data - aux);
                if (l_aux != want.l_aux || (l_aux && memcmp(aux, want.aux, (size_t)l_aux) != 0)) FAIL("sam_prob_realn(%s, flag %d): tags differ (%d vs %d bytes)", recs[i].qname, flags[f], l_aux, want.l_aux);
            }
            bam_destroy1(b); rec_free(&want);
        }
        if (i % 2 == 0) {
            bam1_t *b = bam_init1(); to_bam1(&recs[i], b);
            const int got = sam_cap_mapq(b, ref, ref_len, 50), want = cap_mapq(&recs[i], ref, ref_len, 50);
            ++n_cap;
            if (got != want) FAIL("sam_cap_mapq(%s): %d, oracle %d", recs[i].qname, got, want);
            bam_destroy1(b);
        }
    }
    // ---- errmod_cal: random packed bases; n > 255 exercises ks_shuffle over the drand48 stream (both sides start at
    // srand48(0) and see the same calls in the same order)
    {
        errmod_t *em = errmod_init(1. - 0.83);      // the product's handle (CUDA engine)
        errmod_t *om = errmod_new(1. - 0.83);       // the oracle's tables (same opaque type name, different library)
        if (!em) { fprintf(stderr, "errmod_init failed (no CUDA device?)\n"); return 3; }
        uint64_t s = 12345;
        const int ns[] = {0, 1, 2, 7, 30, 200, 255, 256, 300, 1000, 3000, 31, 256};
        for (int m = 5; m >= 4; --m)
            for (int n : ns) {
                std::vector<uint16_t> a((size_t)n + 1), b;
                for (int i = 0; i < n; ++i) {
                    s = s * 6364136223846793005ULL + 1442695040888963407ULL;
                    const int q = 4 + (int)((s >> 33) % 60), strand = (int)((s >> 20) & 1);
                    int base = (int)((s >> 40) % 100); base = base < 80 ? 0 : base < 90 ? 1 : base < 95 ? 2 : base < 98 ? 3 : 4;
                    if (base >= m) base = m - 1;
                    a[(size_t)i] = (uint16_t)(q << 5 | strand << 4 | base);
                }
                b = a;
                float qa[25], qb[25];
                errmod_cal(em, n, m, a.data(), qa);
                errmod_calc(om, n, m, b.data(), qb);
                ++n_em;
                if (memcmp(qa, qb, sizeof(float) * (size_t)(m * m)) != 0) FAIL("errmod_cal(n=%d, m=%d): likelihoods differ (q[1] %g vs %g)", n, m, qa[1], qb[1]);
                const int ns_ = n > 255 ? 255 : n;
                if (memcmp(a.data(), b.data(), 2 * (size_t)ns_) != 0) FAIL("errmod_cal(n=%d, m=%d): sorted bases differ", n, m);
            }
        errmod_destroy(em);
    }
    // ---- bcf_call_glfgen: columns made of the loaded reads at varying query positions
    {
        bcf_callaux_t *bca = bcf_call_init(0.83, 13);
        errmod_t *om = errmod_new(1. - 0.83);
        if (!bca) { fprintf(stderr, "bcf_call_init failed\n"); return 3; }
        std::vector<bam1_t *> bs;
        for (auto &r : recs) { bam1_t *b = bam_init1(); to_bam1(&r, b); bs.push_back(b); }
        const int depths[] = {1, 3, 17, 64, 255, 256, 400};
        uint64_t s = 99;
        for (int d : depths) {
            if (bs.empty()) break;
            std::vector<bam_pileup1_t> pl((size_t)d); std::vector<pile1_t> po((size_t)d);
            for (int i = 0; i < d; ++i) {
                s = s * 6364136223846793005ULL + 1442695040888963407ULL;
                const size_t ri = (size_t)((s >> 33) % bs.size());
                memset(&pl[(size_t)i], 0, sizeof(bam_pileup1_t)); memset(&po[(size_t)i], 0, sizeof(pile1_t));
                const int lq = bs[ri]->core.l_qseq;
                int qpos = lq > 0 ? (int)((s >> 12) % (uint64_t)(lq + 2)) : 0;     // sometimes past the end
                pl[(size_t)i].b = bs[ri]; pl[(size_t)i].qpos = qpos; pl[(size_t)i].is_del = ((s >> 50) % 10) == 0; pl[(size_t)i].is_refskip = ((s >> 54) % 23) == 0;
                po[(size_t)i].b = &recs[ri]; po[(size_t)i].qpos = qpos; po[(size_t)i].is_del = pl[(size_t)i].is_del; po[(size_t)i].is_refskip = pl[(size_t)i].is_refskip;
            }
            for (int rb4 : {1, 2, 4, 8, 15}) {
                bcf_callret1_t r; float qs[4], p25[25];
                const int got = bcf_call_glfgen(d, pl.data(), rb4, bca, &r);
                const int want = glfgen(d, po.data(), rb4, 13, 60, om, qs, p25);
                ++n_gl;
                if (got != want) FAIL("bcf_call_glfgen(depth %d): n %d, oracle %d", d, got, want);
                else if (memcmp(r.qsum, qs, sizeof qs) != 0 || memcmp(r.p, p25, sizeof p25) != 0) FAIL("bcf_call_glfgen(depth %d, ref %d): qsum / p differ (p[1] %g vs %g)", d, rb4, r.p[1], p25[1]);
            }
        }
        for (bam1_t *b : bs) bam_destroy1(b);
        bcf_call_destroy(bca);
    }
    printf("sam_prob_realn %ld calls, sam_cap_mapq %ld, errmod_cal %ld, bcf_call_glfgen %ld: %ld mismatches\n", n_realn, n_cap, n_em, n_gl, g_bad);
    return g_bad ? 1 : 0;
}
