// baq_reg.h -- the banded glocal pair-HMM of BAQ with the band held in registers (band = 7: the shape
// sam_prob_realn gives every read whose aligned reference and query spans differ by <= 7).
//
// Replaces htslib probaln_glocal (probaln.c) + the sam_prob_realn epilogue (realn.c) as called from
// bam_plcmd.c:451; semantics SURVEY.md 8a rows a2/a3, Appendix A6.  One thread owns one read.
//
// The reference stores a (l_query+1) x (2bw+1) x 3 matrix per direction.  Here a row of the band lives in
// registers in DIAGONAL coordinates: cell j = 0..14 of row i is reference offset k = i - 7 + j, so that
//   forward   M(i,j) <- row i-1 cell j      I(i,j) <- row i-1 cell j+1     D(i,j) <- row i cell j-1
//   backward  b(i,j) <- row i+1 cell j (M), row i+1 cell j-1 (I), row i cell j+1 (D)
// and a row is updated IN PLACE (ascending j forward, descending j backward) with compile-time register
// indices.  Cells outside 1 <= k <= l_ref are zero, which is what the reference's calloc'ed guard slots hold.
// Only the scaled forward M and I states (what the MAP step multiplies) leave the SM: 30 doubles + the row's
// scaling factor per row, written once and read back once -- the D state, the backward matrix and the
// unscaled values never touch memory.
//
// Bit-exactness: every product / sum is evaluated in the association order of the C source, in IEEE double,
// with no FMA contraction (the translation unit is built with -fmad=false; the host test build with
// -ffp-contract=off); the sequential-in-k D recurrences and row sums stay sequential.  pow()/log() never run
// on the device: host-libm tables (see baq.cuh).
//
// The arithmetic is __host__ __device__ so that the test harness tests/emul/baq_host.cpp can single-step it on the
// CPU against the CPU restatement of sam_prob_realn; memory traffic goes through a policy object (Mem).
#pragma once
#include <stdint.h>
#include <math.h>
#include <string.h>
#include "plp_core.h"

namespace baqr {

constexpr int BW = 7, NB = 2 * BW + 1;     // band half-width, cells per row
#define BAQR_EI .25
#define BAQR_EM .33333333333

struct Par { double m0, m1, m2, m3, m4, m6, m8, EIm1, EIm4, bM, bI, sM; };

PLP_HD Par make_par(int l_query, int l_ref)
{
    const double cd = 0.001, ce = 0.1;
    Par p;
    p.sM = 1. / (2 * l_query + 2);
    const double sI = p.sM;
    p.m0 = (1 - cd - cd) * (1 - p.sM); p.m1 = p.m2 = cd * (1 - p.sM);
    p.m3 = (1 - ce) * (1 - sI); p.m4 = ce * (1 - sI);
    p.m6 = 1 - ce; p.m8 = ce;
    p.bM = (1 - cd) / l_ref; p.bI = cd / l_ref;
    p.EIm1 = BAQR_EI * p.m1; p.EIm4 = BAQR_EI * p.m4;
    return p;
}

// emission of cell j: xw = (window of reference codes, one nibble per cell) XOR (query code in every nibble)
//   nibble 0 -> equal codes (1 - eps); bit 2 set -> the reference base is ambiguous (1.0); else eps/3.
// A query base that is ambiguous passes em_match = em_mis = 1.0.
PLP_HD double emis_sel(uint64_t xw, int j, double em_match, double em_mis)
{
    const uint32_t t = (uint32_t)(xw >> (4 * j)) & 0xfu;
    return t == 0 ? em_match : ((t & 4u) ? 1.0 : em_mis);
}

PLP_HD bool cell_valid(int i, int j, int l_ref) { const int k = i - BW + j; return k >= 1 && k <= l_ref; }

// ---- per-row inputs, eight rows per load.  A row consumes one quality byte, one base nibble and one reference code; each is
// a dependent, poorly cached access when fetched byte by byte (a warp's 32 reads touch 32 different sectors).  The streams
// below hold the eight bytes that cover eight consecutive rows in one 64-bit register, with the next eight already in flight,
// so a thread issues one load per stream per eight rows and has eight rows of FP64 work to hide it behind.
// ld8: unaligned 8-byte little-endian load (device: two aligned 8-byte loads and a funnel shift -- up to 15 bytes past p are
// touched, the staged arrays carry that slack; host harness: memcpy from padded buffers).
PLP_HD uint64_t ld8(const uint8_t *p)
{
#if defined(__CUDA_ARCH__)
    const unsigned long long a = (unsigned long long)p;
    const uint64_t *w = reinterpret_cast<const uint64_t *>(a & ~7ull);
    const uint32_t sh = (uint32_t)(a & 7ull) * 8u;
    const uint64_t lo = w[0], hi = w[1];
    return sh ? (lo >> sh) | (hi << (64u - sh)) : lo;
#else
    uint64_t v; memcpy(&v, p, 8); return v;
#endif
}
PLP_HD uint32_t byte_of(uint64_t w, int k) { return (uint32_t)(w >> (8 * k)) & 0xffu; }
// query base qi of a read whose first base is nibble qoff of seq4, from the 8-byte word loaded at byte (qoff + (qi & ~7)) >> 1
PLP_HD int nib_at(uint64_t w, uint32_t qoff, int base, int qi)     // word loaded at byte (qoff + base) >> 1, base <= qi < base + 8
{
    const uint32_t g = qoff + (uint32_t)qi, g0 = qoff + (uint32_t)base;
    return (int)((w >> (8u * ((g >> 1) - (g0 >> 1)) + ((~g & 1u) << 2))) & 0xfu);
}
PLP_HD int nib_of(uint64_t w, uint32_t qoff, int qi) { return nib_at(w, qoff, qi & ~7, qi); }

// forward row i >= 2 from row i-1 (scaled) held in M/I/D; returns the row sum s[i] (cells left UNSCALED)
template <bool EDGE>
PLP_HD double fwd_row(double (&M)[NB], double (&I)[NB], double (&D)[NB], const Par &p, uint64_t xw, double em_match, double em_mis, int i, int l_ref)
{
    double sum = 0., pM = 0., pD = 0.;
#pragma unroll
    for (int j = 0; j < NB; ++j) {
        const double e = emis_sel(xw, j, em_match, em_mis);
        double Mn = e * (p.m0 * M[j] + p.m3 * I[j] + p.m6 * D[j]);
        double In = 0.;
        if (j + 1 < NB) In = BAQR_EI * (p.m1 * M[j + 1] + p.m4 * I[j + 1]);
        double Dn = p.m2 * pM + p.m8 * pD;
        if (EDGE) { if (!cell_valid(i, j, l_ref)) { Mn = 0.; In = 0.; Dn = 0.; } }
        sum += Mn + In + Dn;
        M[j] = Mn; I[j] = In; D[j] = Dn;
        pM = Mn; pD = Dn;
    }
    return sum;
}

// backward row i (1 <= i < l_query) from row i+1 (scaled) held in M/I, unscaled result left in M/I
template <bool EDGE>
PLP_HD void bwd_row(double (&M)[NB], double (&I)[NB], const Par &p, uint64_t xw, double em_match, double em_mis, int i, int l_ref)
{
    double nD = 0.;
#pragma unroll
    for (int j = NB - 1; j >= 0; --j) {
        double em = emis_sel(xw, j, em_match, em_mis);
        if (EDGE) { if (i - BW + j >= l_ref) em = 0.; }            // "k >= l_ref ? 0"
        const double e = em * M[j];
        const double b10 = j > 0 ? I[j - 1] : 0.;
        double B0 = e * p.m0 + p.EIm1 * b10 + p.m2 * nD;
        double B1 = e * p.m3 + p.EIm4 * b10;
        double B2 = e * p.m6 + p.m8 * nD;
        if (i == 1) B2 = B2 * 0.;                                     // "* y", y = (i > 1): x * 1. is x
        if (EDGE) { if (!cell_valid(i, j, l_ref)) { B0 = 0.; B1 = 0.; B2 = 0.; } }
        M[j] = B0; I[j] = B1; nD = B2;
    }
}

// (int)(-4.343 * log(1 - max) + .499) capped like the reference, from the break points of that step function
PLP_HD int phred_of(double xx, const double *qthr)
{
    if (!(xx > 0.)) return 0;            // log(0) / NaN: x86 cvttsd2si gives INT_MIN, stored as uint8 0
    int lo_ = 0, hi_ = 101;              // count thresholds T[1..101] with xx <= T[j]
    while (lo_ < hi_) { const int mid = (lo_ + hi_ + 1) >> 1; if (xx <= qthr[mid]) lo_ = mid; else hi_ = mid - 1; }
    return lo_ > 100 ? 99 : lo_;
}

// One read.  Mem provides
//   int  ref_code(int p)                         code 0..4 of window position p (0-based, p < l_ref)
//   uint64_t ref8(int p)                         codes of window positions p..p+7 (p >= 0, multiple of 8), one per byte; 4 beyond the sequence
//   void put_row(int i, M, I, double inv)        scaled forward M/I states of row i and 1/s[i]
//   void fence()                                 rows written so far are visible to the fetches that follow
//   void fetch(int i)                            start bringing row i back (asynchronous on the device)
//   void wait(int pending)                       all but the `pending` most recent fetches have landed
//   void get(int i, int j, double &fM, double &fI);   double inv(int i)
//   void put_word(int j, int32_t w);  int32_t get_word(int j)        per-base scratch
// q2pf[q] = (double)(float)pow(10, -q/10.)   qthr = break points (see baq.cuh)
template <class Mem>
PLP_HD void baq_read(Mem &mem, uint8_t *qual, const uint8_t *seq4, uint32_t qoff, int lq, int l_ref, int64_t pos, int64_t xb,
                     const uint32_t *cg, int n_cigar, const double *q2pf, const double *qthr, bool extend = true)
{
    const Par p = make_par(lq, l_ref);
    double M[NB], I[NB], D[NB];
    const uint64_t kRep = 0x0111111111111111ull, kMask = 0x0fffffffffffffffull;
    // window of reference codes: nibble j of `win` at row i is the code of window position i + j - 8
    uint64_t win = 0;
#pragma unroll
    for (int j = BW; j < NB; ++j) { const int pp = j - BW; win |= (uint64_t)(pp < l_ref ? mem.ref_code(pp) : 4) << (4 * j); }
    double s_lq;
    {   // ---- forward, row 1
        const int qc = plp::nt16_int_of(plp::base4(seq4, qoff, 0));
        const double ql = q2pf[qual[0]];
        const double em_match = qc > 3 ? 1. : 1. - ql, em_mis = qc > 3 ? 1. : ql * BAQR_EM;
        const uint64_t xw = win ^ (kRep * (uint64_t)(qc & 7));
        double sum = 0.;
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            double a = 0., b = 0.;
            if (j >= BW && cell_valid(1, j, l_ref)) { a = emis_sel(xw, j, em_match, em_mis) * p.bM; b = BAQR_EI * p.bI; sum += a + b; }
            M[j] = a; I[j] = b; D[j] = 0.;
        }
#pragma unroll
        for (int j = BW; j < NB; ++j) if (cell_valid(1, j, l_ref)) { M[j] /= sum; I[j] /= sum; }
        mem.put_row(1, M, I, 1. / sum);
        s_lq = sum;
    }
    // the inputs of a row (query base i-1, window position i + 6 entering at cell 14) come from the 8-row streams (see ld8)
    const uint8_t *sq = seq4 + (qoff >> 1);                 // byte of base 0 (its high or low nibble: qoff & 1)
    const uint32_t qpar = qoff & 1u;
    uint64_t qw = ld8(qual), sw = ld8(sq), rw = 0;
    uint64_t qn = lq > 8 ? ld8(qual + 8) : 0, sn = lq > 8 ? ld8(sq + ((qpar + 8u) >> 1)) : 0, rn = mem.ref8(BW + 1);   // row 2 opens block 8 of the reference stream
    for (int i = 2; i <= lq; ++i) {
        const int qi = i - 1, rp = i + BW - 1;
        if ((qi & 7) == 0) { qw = qn; sw = sn; if (qi + 8 < lq) { qn = ld8(qual + qi + 8); sn = ld8(sq + ((qpar + (uint32_t)qi + 8u) >> 1)); } }
        if ((rp & 7) == 0) { rw = rn; rn = mem.ref8(rp + 8); }
        const uint32_t cq = byte_of(qw, qi & 7);
        const int cr = rp < l_ref ? (int)byte_of(rw, rp & 7) : 4;
        win = (win >> 4) | ((uint64_t)cr << (4 * (NB - 1)));
        const int qc = plp::nt16_int_of(nib_of(sw, qpar, qi));
        const double ql = q2pf[cq];
        const double em_match = qc > 3 ? 1. : 1. - ql, em_mis = qc > 3 ? 1. : ql * BAQR_EM;
        const uint64_t xw = win ^ (kRep * (uint64_t)(qc & 7));
        double sum, inv;
        if (i > BW && i + BW <= l_ref) {
            sum = fwd_row<false>(M, I, D, p, xw, em_match, em_mis, i, l_ref);
            inv = 1. / sum;
#pragma unroll
            for (int j = 0; j < NB; ++j) { M[j] *= inv; I[j] *= inv; D[j] *= inv; }
        } else {
            sum = fwd_row<true>(M, I, D, p, xw, em_match, em_mis, i, l_ref);
            inv = 1. / sum;
#pragma unroll
            for (int j = 0; j < NB; ++j) if (cell_valid(i, j, l_ref)) { M[j] *= inv; I[j] *= inv; D[j] *= inv; }
        }
        mem.put_row(i, M, I, inv);
        s_lq = sum;
    }
    mem.fence();
    mem.fetch(lq);
    double s_last = 0.;      // ---- termination
#pragma unroll
    for (int j = 0; j < NB; ++j) if (cell_valid(lq, j, l_ref)) s_last += M[j] * p.sM + I[j] * p.sM;
    {   // ---- backward, row l_query
        const double bv = p.sM / s_lq / s_last;
#pragma unroll
        for (int j = 0; j < NB; ++j) { const double v = cell_valid(lq, j, l_ref) ? bv : 0.; M[j] = v; I[j] = v; }
    }
    // `win` holds positions lq + j - 8: exactly what backward row lq-1 compares against (position k = i - 7 + j)
    // backward row i uses query base i and, from row lq-2 down, window position i - 7 entering at cell 0: the same 8-row
    // streams, walked downwards (block of index x = x & ~7; the block below is in flight while a block is consumed)
    const int r0 = lq - 2 - BW;                              // first position to enter (row lq-2)
    {
        const int q0 = (lq - 1) & ~7, rb = r0 >= 0 ? (r0 & ~7) : 0;
        qw = ld8(qual + q0); sw = ld8(sq + ((qpar + (uint32_t)q0) >> 1));
        if (q0 >= 8) { qn = ld8(qual + q0 - 8); sn = ld8(sq + ((qpar + (uint32_t)q0 - 8u) >> 1)); }
        rw = mem.ref8(rb); if (rb >= 8) rn = mem.ref8(rb - 8);
    }
    for (int i = lq; i >= 1; --i) {
        if (i > 1) mem.fetch(i - 1);
        if (i < lq) {
            const int qi = i, rp = i - BW;
            if ((qi & 7) == 7 && qi != lq - 1) { qw = qn; sw = sn; if (qi >= 15) { qn = ld8(qual + qi - 15); sn = ld8(sq + ((qpar + (uint32_t)qi - 15u) >> 1)); } }
            if (rp < r0 && rp >= 0 && (rp & 7) == 7) { rw = rn; if (rp >= 15) rn = mem.ref8(rp - 15); }
            const uint32_t cq = byte_of(qw, qi & 7);
            const int cr = (rp >= 0 && rp < l_ref) ? (int)byte_of(rw, rp & 7) : 4;
            if (i < lq - 1) win = ((win << 4) & kMask) | (uint64_t)cr;
            const int qc = plp::nt16_int_of(nib_of(sw, qpar, qi));
            const double ql = q2pf[cq];
            const double em_match = qc > 3 ? 1. : 1. - ql, em_mis = qc > 3 ? 1. : ql * BAQR_EM;
            const uint64_t xw = win ^ (kRep * (uint64_t)(qc & 7));
            if (i > BW && i + BW < l_ref) {
                bwd_row<false>(M, I, p, xw, em_match, em_mis, i, l_ref);
                mem.wait(1);                                         // row i (fetched one iteration ago) has landed
                const double ys = mem.inv(i);
#pragma unroll
                for (int j = 0; j < NB; ++j) { M[j] *= ys; I[j] *= ys; }
            } else {
                bwd_row<true>(M, I, p, xw, em_match, em_mis, i, l_ref);
                mem.wait(i > 1 ? 1 : 0);
                const double ys = mem.inv(i);
#pragma unroll
                for (int j = 0; j < NB; ++j) if (cell_valid(i, j, l_ref)) { M[j] *= ys; I[j] *= ys; }
            }
        } else {
            // row lq: nothing to compute (the streams are primed above)
            mem.wait(i > 1 ? 1 : 0);
        }
        // ---- MAP of row i (cells outside the band are zero on both sides: they add 0 and never exceed the maximum)
        double sum = 0., mx = 0.; int best = -1;
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            double fM, fI; mem.get(i, j, fM, fI);
            const double z0 = fM * M[j], z1 = fI * I[j];
            const bool g0 = z0 > mx; mx = g0 ? z0 : mx; best = g0 ? 2 * j : best; sum += z0;
            const bool g1 = z1 > mx; mx = g1 ? z1 : mx; best = g1 ? 2 * j + 1 : best; sum += z1;
        }
        mx /= sum;
        const int max_k = best < 0 ? -1 : (((i - BW - 1 + (best >> 1)) << 2) | (best & 1));
        const int kq = phred_of(1. - mx, qthr);
        mem.put_word(i - 1, (int32_t)((uint32_t)max_k << 8 | (uint32_t)kq));
    }
    // ---- sam_prob_realn epilogue (APPLY): per match run, zero the bases whose MAP state is not the aligned match,
    // with EXTEND replace each by the smaller of the running maxima from both ends of the run, cap the quality
    int64_t x = pos; int y = 0;
    for (int kk = 0; kk < n_cigar; ++kk) {
        const int op = cg[kk] & 0xf; int l = (int)(cg[kk] >> 4);
        if (plp::is_mop(op)) {
            if (l > lq - y) l = lq - y;
            if (l > 0) {
                // eight bases per step: the scratch words (and, on the way back, the quality bytes) of a step are independent loads,
                // issued together -- a base-by-base loop exposes one full memory round trip per base
                int left = 0;
                for (int j0 = y; j0 < y + l; j0 += 8) {
                    const int n = y + l - j0 < 8 ? y + l - j0 : 8;
                    int32_t w[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) w[u] = u < n ? mem.get_word(j0 + u) : 0;
#pragma unroll
                    for (int u = 0; u < 8; ++u) if (u < n) {
                        const int j = j0 + u;
                        const int st = w[u] >> 8, kq = w[u] & 0xff;
                        const int t = ((st & 3) != 0 || (int64_t)(st >> 2) != x - xb + (j - y)) ? 0 : kq;
                        left = (extend && left > t) ? left : t;
                        mem.put_word(j, t | left << 8);
                    }
                }
                int rght = 0;
                for (int j1 = y + l - 1; j1 >= y; j1 -= 8) {
                    const int n = j1 - y + 1 < 8 ? j1 - y + 1 : 8;
                    int32_t w[8]; int qv[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) { w[u] = u < n ? mem.get_word(j1 - u) : 0; qv[u] = u < n ? qual[j1 - u] : 0; }
#pragma unroll
                    for (int u = 0; u < 8; ++u) if (u < n) {
                        const int t = w[u] & 0xff, lf = w[u] >> 8;
                        rght = (extend && rght > t) ? rght : t;
                        const int bq = lf < rght ? lf : rght;
                        qual[j1 - u] = (uint8_t)(qv[u] - (qv[u] <= bq ? 0 : qv[u] - bq));
                    }
                }
            }
            x += l; y += l;
        } else if (op == plp::OP_S || op == plp::OP_I) { if (l > lq - y) l = lq - y; y += l; }
        else if (op == plp::OP_D) x += l;
    }
}

// host: tables with the box's own libm (what the reference binary would use here)
//   q2p[q]  = pow(10, -q/10.)            (the kernels round it to float per base like the reference's `float qual[]`)
//   qthr[j] = the largest x in (0,1] with (int)(-4.343*log(x)+.499) >= j, j = 1..101 (qthr[0] = 2)
inline void host_tables(double *q2p, double *qthr)
{
    for (int i = 0; i < 256; ++i) q2p[i] = pow(10, -i / 10.);
    qthr[0] = 2.0;
    for (int j = 1; j <= 101; ++j) {
        double x = exp(-((double)j - .499) / 4.343);
        auto val = [](double v) { return (int)(-4.343 * log(v) + .499); };
        while (x > 0 && val(x) < j) x = nextafter(x, 0.0);
        while (true) { double nx = nextafter(x, 2.0); if (nx <= 1.0 && val(nx) >= j) x = nx; else break; }
        qthr[j] = x;
    }
}

}  // namespace baqr
