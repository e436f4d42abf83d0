// hts_read_ops.cpp -- htslib's per-read and per-column entry points of the hot path, served by the CUDA engine
// (tier T1, include/b200_htslib_compat.h).  These are the symbols an UNMODIFIED caller links against:
//   sam_prob_realn, sam_cap_mapq            mplp_func bam_plcmd.c:451,453; calmd bam_md.c:475,481
//   errmod_init / errmod_cal / errmod_destroy   bam2bcf.c:46,53,121; phase.c:754; cut_target.c:84
//   bcf_call_init / bcf_call_glfgen / bcf_call_destroy   bam2bcf.h:51-53; tv_pl_func bam_tview.c:197
//   bam_plp_insertion_mod                   pileup_seq bam_plcmd.c:119
// One small batch (a read, a column) per call: a correctness surface with htslib's signatures and return codes, not
// the fast path -- the batch tier (include/b200_pileup.h) is.  All arithmetic of the path runs on the device: BAQ in
// k_baq_*, the mapq cap in k_cap_mapq, glfgen / errmod_cal in k_glfgen_one / k_errmod_one.  Host side: packing, the
// integer BQ:Z/ZQ:Z tag path of sam_prob_realn (pure byte bookkeeping, realn.c) and tag edits on bam1_t.
// No CPU fallback: every call fails (error return) when no CUDA device is available.
#include "../../../include/b200_htslib_compat.h"
#include "../../../include/b200_pileup.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {

b200_engine_t *g_eng = nullptr;
b200_engine_t *engine()
{
    if (!g_eng) {
        const char *d = getenv("B200_DEVICE");
        if (b200_engine_create(d ? atoi(d) : 0, &g_eng) != 0) g_eng = nullptr;
    }
    return g_eng;
}

// ---- aux block of a bam1_t (SAMv1 4.2.4) ------------------------------------
uint8_t *aux_begin(bam1_t *b) { return bam_get_qual(b) + b->core.l_qseq; }
uint8_t *aux_end(bam1_t *b) { return b->data + b->l_data; }
// length of the value that starts at the type byte s (type byte included), or -1 when malformed
long aux_value_len(const uint8_t *s, const uint8_t *end)
{
    if (s >= end) return -1;
    switch (*s) {
    case 'A': case 'c': case 'C': return 2;
    case 's': case 'S': return 3;
    case 'i': case 'I': case 'f': return 5;
    case 'd': return 9;
    case 'Z': case 'H': { const uint8_t *p = s + 1; while (p < end && *p) ++p; return p < end ? (long)(p - s) + 1 : -1; }
    case 'B': {
        if (s + 6 > end) return -1;
        uint32_t n; memcpy(&n, s + 2, 4);
        int w;
        switch (s[1]) { case 'c': case 'C': w = 1; break; case 's': case 'S': w = 2; break; case 'i': case 'I': case 'f': w = 4; break; default: return -1; }
        return 6 + (long)n * w;
    }
    default: return -1;
    }
}
// -> type byte of tag, or nullptr
uint8_t *aux_get(bam1_t *b, const char tag[2])
{
    uint8_t *s = aux_begin(b), *end = aux_end(b);
    while (s + 3 <= end) {
        const long l = aux_value_len(s + 2, end);
        if (l < 0) return nullptr;
        if (s[0] == (uint8_t)tag[0] && s[1] == (uint8_t)tag[1]) return s + 2;
        s += 2 + l;
    }
    return nullptr;
}
void aux_del(bam1_t *b, uint8_t *type_byte)
{
    uint8_t *end = aux_end(b);
    const long l = aux_value_len(type_byte, end);
    if (l < 0) return;
    uint8_t *from = type_byte + l, *to = type_byte - 2;
    memmove(to, from, (size_t)(end - from));
    b->l_data -= (int)(from - to);
}
int aux_append_z(bam1_t *b, const char tag[2], const uint8_t *str, int len_with_nul)
{
    const size_t need = (size_t)b->l_data + 3 + (size_t)len_with_nul;
    if (b->m_data < need) {
        uint8_t *d = (uint8_t *)realloc(b->data, need + 32);
        if (!d) return -1;
        b->data = d; b->m_data = (uint32_t)(need + 32);
    }
    uint8_t *p = b->data + b->l_data;
    p[0] = (uint8_t)tag[0]; p[1] = (uint8_t)tag[1]; p[2] = 'Z';
    memcpy(p + 3, str, (size_t)len_with_nul);
    b->l_data = (int)need;
    return 0;
}

// one read as a batch; the reference travels as a window around the read (the engine treats bases outside the
// window as 'N', so the margin covers everything sam_prob_realn / sam_cap_mapq can look at)
struct OneRead {
    int64_t pos, mpos, isize, file_start[2], prev; uint16_t flag; uint8_t mapq, rbits; int32_t l_qseq, mtid; uint32_t n_cigar; uint64_t cigar_off, qual_off;
    std::vector<uint32_t> cigar; std::vector<uint8_t> seq4, qual;
    b200_batch_t bt;
    void pack(const bam1_t *b, const char *ref, hts_pos_t ref_len)
    {
        const bam1_core_t &c = b->core;
        pos = c.pos; mpos = c.mpos; isize = c.isize; flag = c.flag; mapq = c.qual; l_qseq = c.l_qseq; mtid = c.mtid; n_cigar = c.n_cigar;
        cigar_off = 0; qual_off = 0; prev = -1; rbits = 0; file_start[0] = 0; file_start[1] = 1;
        const uint32_t *cg = bam_get_cigar(b);
        cigar.assign(cg, cg + c.n_cigar); cigar.push_back(0);
        const uint8_t *q = bam_get_qual(b), *s = bam_get_seq(b);
        qual.assign(q, q + c.l_qseq); qual.resize((size_t)c.l_qseq + 16, 0);
        seq4.assign(s, s + (c.l_qseq + 1) / 2); seq4.resize((size_t)(c.l_qseq + 1) / 2 + 16, 0);
        int64_t rl = 0;
        for (uint32_t k = 0; k < c.n_cigar; ++k) { const int op = cg[k] & 0xf; if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rl += cg[k] >> 4; }
        memset(&bt, 0, sizeof bt);
        bt.n_files = 1; bt.n_reads = 1; bt.file_start = file_start;
        bt.pos = &pos; bt.flag = &flag; bt.mapq = &mapq; bt.l_qseq = &l_qseq; bt.n_cigar = &n_cigar; bt.cigar_off = &cigar_off; bt.qual_off = &qual_off;
        bt.mtid = &mtid; bt.mpos = &mpos; bt.isize = &isize; bt.prev_same_name = nullptr; bt.rbits = nullptr;
        bt.cigar = cigar.data(); bt.n_cigar_total = c.n_cigar; bt.seq4 = seq4.data(); bt.qual = qual.data(); bt.qual_bytes = (uint64_t)c.l_qseq;
        bt.tid = c.tid; bt.tid_len = ref_len; bt.tid_name = "";
        const int64_t margin = 2 * ((int64_t)c.l_qseq + rl) + 64;
        int64_t rb = c.pos - margin, re = c.pos + rl + margin;
        if (rb < 0) rb = 0;
        if (re > ref_len) re = ref_len;
        if (ref && re > rb) { bt.ref = ref + rb; bt.ref_beg = rb; bt.ref_n = re - rb; bt.ref_len = ref_len; }
    }
};

b200_stage_conf_t read_conf(const OneRead &r, int baq)
{
    b200_stage_conf_t sc; memset(&sc, 0, sizeof sc);
    sc.mode = B200_MODE_MPILEUP; sc.baq = baq;
    sc.beg = r.pos > 0 ? r.pos : 0; sc.end = ((int64_t)INT32_MAX << 32) | UINT32_MAX;
    return sc;
}

}  // namespace

extern "C" {

// htslib realn.c sam_prob_realn: flag bit 0 = apply, bit 1 = extend, bit 2 = redo.  Return codes as upstream:
// 0 done, -1 nothing to do (unmapped / no sequence / no qualities / ref-skip / no aligned base), -3 inconsistent tags,
// -4 failure (upstream: allocation; here also: no CUDA device).
int sam_prob_realn(bam1_t *b, const char *ref, hts_pos_t ref_len, int flag)
{
    const int apply = flag & 1, extend = flag & 2, redo = flag & 4;
    uint8_t *qual = bam_get_qual(b);
    const int l = b->core.l_qseq;
    if ((b->core.flag & 4) || l == 0 || qual[0] == 0xff) return -1;
    uint8_t *bq = aux_get(b, "BQ"); if (bq && *bq != 'Z') bq = nullptr;
    uint8_t *zq = aux_get(b, "ZQ"); if (zq && *zq != 'Z') zq = nullptr;
    if (bq && redo) { aux_del(b, bq); bq = nullptr; zq = aux_get(b, "ZQ"); if (zq && *zq != 'Z') zq = nullptr; }
    if (bq && zq) { aux_del(b, zq); zq = nullptr; bq = aux_get(b, "BQ"); }
    qual = bam_get_qual(b);
    if (bq || zq) {   // stored tag: integer bookkeeping only
        if ((apply && zq) || (!apply && bq)) return -3;
        if (bq && apply) {
            const uint8_t *t = bq + 1;
            for (int i = 0; i < l; ++i) qual[i] = qual[i] + 64 < t[i] ? 0 : (uint8_t)(qual[i] - ((int)t[i] - 64));
            bq[-2] = 'Z';
        } else if (zq && !apply) {
            const uint8_t *t = zq + 1;
            for (int i = 0; i < l; ++i) qual[i] = (uint8_t)(qual[i] + ((int)t[i] - 64));
            zq[-2] = 'B';
        }
        return 0;
    }
    {   // reads the HMM never sees (realn.c prologue): a reference skip, or no aligned base
        const uint32_t *cg = bam_get_cigar(b);
        bool any_m = false;
        for (uint32_t k = 0; k < b->core.n_cigar; ++k) { const int op = cg[k] & 0xf; if (op == 3) return -1; if (op == 0 || op == 7 || op == 8) any_m = true; }
        if (!any_m) return -1;
    }
    b200_engine_t *e = engine();
    if (!e || !ref) return -4;
    OneRead r; r.pack(b, ref, ref_len);
    const b200_stage_conf_t sc = read_conf(r, extend ? 1 : 3);      // 3: without the extension of realn.c's BAQ_EXTEND
    b200_stage_stats_t st;
    if (b200_stage(e, &r.bt, &sc, &st) != 0) { fprintf(stderr, "[b200 sam_prob_realn] %s\n", b200_last_error(e)); return -4; }
    std::vector<uint8_t> nq((size_t)l + 16);
    if (b200_fetch_qual(e, nq.data(), (size_t)l) != 0) return -4;
    // tag = 64 + (quality taken away); APPLY rewrites the qualities and stores the tag as ZQ, else BQ
    std::vector<uint8_t> tag((size_t)l + 1);
    for (int i = 0; i < l; ++i) tag[(size_t)i] = (uint8_t)(64 + (qual[i] - nq[(size_t)i]));
    tag[(size_t)l] = 0;
    if (apply) memcpy(qual, nq.data(), (size_t)l);
    if (aux_append_z(b, apply ? "ZQ" : "BQ", tag.data(), l + 1) != 0) return -4;
    return 0;
}

// htslib realn.c sam_cap_mapq: the capped mapping quality, -1 when the read should be dropped
int sam_cap_mapq(bam1_t *b, const char *ref, hts_pos_t ref_len, int thres)
{
    b200_engine_t *e = engine();
    if (!e || !ref) return -1;
    OneRead r; r.pack(b, ref, ref_len);
    const b200_stage_conf_t sc = read_conf(r, 0);
    b200_stage_stats_t st;
    if (b200_stage(e, &r.bt, &sc, &st) != 0) { fprintf(stderr, "[b200 sam_cap_mapq] %s\n", b200_last_error(e)); return -1; }
    int32_t q = -1;
    if (b200_cap_mapq(e, thres, &q, 1) != 0) { fprintf(stderr, "[b200 sam_cap_mapq] %s\n", b200_last_error(e)); return -1; }
    return q;
}

// ---- errmod (htslib errmod.h) -------------------------------------------------
struct errmod_t { double depcorr; };
errmod_t *errmod_init(double depcorr)
{
    if (!engine()) return nullptr;
    errmod_t *em = (errmod_t *)calloc(1, sizeof(errmod_t));
    if (em) em->depcorr = depcorr;
    return em;
}
void errmod_destroy(errmod_t *em) { free(em); }
int errmod_cal(const errmod_t *em, int n, int m, uint16_t *bases, float *q)
{
    if (m > 16) return -1;                       // "m > m" check of upstream is vacuous; 4-bit allele codes bound m
    for (int i = 0; i < m * m; ++i) q[i] = 0.f;
    if (n == 0) return 0;
    b200_engine_t *e = engine();
    if (!e || !em) return -1;
    if (b200_errmod_cal(e, em->depcorr, n, m, bases, q) != 0) { fprintf(stderr, "[b200 errmod_cal] %s\n", b200_last_error(e)); return -1; }
    return 0;
}

// ---- bam2bcf.h ------------------------------------------------------------------
bcf_callaux_t *bcf_call_init(double theta, int min_baseQ)
{
    if (theta <= 0.) theta = 0.83;               // CALL_DEFTHETA
    bcf_callaux_t *bca = (bcf_callaux_t *)calloc(1, sizeof(bcf_callaux_t));
    if (!bca) return nullptr;
    bca->capQ = 60; bca->min_baseQ = min_baseQ;
    bca->e = errmod_init(1. - theta);
    if (!bca->e) { free(bca); return nullptr; }
    return bca;
}
void bcf_call_destroy(bcf_callaux_t *bca)
{
    if (!bca) return;
    errmod_destroy(bca->e);
    free(bca->bases); free(bca);
}
int bcf_call_glfgen(int _n, const bam_pileup1_t *pl, int ref_base, bcf_callaux_t *bca, bcf_callret1_t *r)
{
    memset(r->qsum, 0, sizeof(float) * 4);
    memset(r->p, 0, sizeof(float) * 25);
    if (_n <= 0) return -1;
    if (ref_base < 0) return -1;                 // indel columns (p->aux packing) are not part of the samtools callers' use
    b200_engine_t *e = engine();
    if (!e) return -1;
    std::vector<uint8_t> buf((size_t)_n * 4);
    uint8_t *q = buf.data(), *mq = q + _n, *b4 = mq + _n, *fl = b4 + _n;
    for (int i = 0; i < _n; ++i) {
        const bam_pileup1_t *p = pl + i;
        const bam1_t *b = p->b;
        const bool in = p->qpos < b->core.l_qseq;
        q[i] = in ? bam_get_qual(b)[p->qpos] : 0;
        mq[i] = b->core.qual;
        b4[i] = in ? (uint8_t)bam_seqi(bam_get_seq(b), p->qpos) : 0xff;
        fl[i] = (uint8_t)(((p->is_del || p->is_refskip || (b->core.flag & 4)) ? 1 : 0) | (bam_is_rev(b) ? 2 : 0));
    }
    const int n = b200_glfgen(e, bca->e->depcorr, _n, q, mq, b4, fl, ref_base, bca->min_baseQ, bca->capQ, r->qsum, r->p);
    if (n < 0) fprintf(stderr, "[b200 bcf_call_glfgen] %s\n", b200_last_error(e));
    return n;
}

// htslib sam.c bam_plp_insertion_mod without base-modification markup (m == NULL is how pileup_seq calls it unless -M)
int bam_plp_insertion_mod(const bam_pileup1_t *p, hts_base_mod_state *m, kstring_t *ins, int *del_len)
{
    (void)m;
    return bam_plp_insertion(p, ins, del_len);
}

}  // extern "C"
