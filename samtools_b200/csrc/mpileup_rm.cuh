// mpileup_rm.cuh -- read-major mpileup kernels (the fast path of b200_mpileup_text).
//
// Same output as the column-major formatter in plp_core.h (mp_line_size /
// mp_line_write, kept as the selectable reference variant), different
// traversal.  A CTA owns RM_COLS adjacent columns, each of its warps owns
// RM_SUB of them.  A warp streams the reads that can cover its columns in
// file order (far-reaching list, then the contiguous slice) and, for every
// read, runs its lanes ALONG THE READ: lane j handles base j, j+32, ...  so
// quality / base loads are coalesced 32-byte / 16-byte requests instead of one
// byte per lane from 30 different reads.  Because exactly one warp touches a
// column and it visits the reads in file order, per-column counters and write
// cursors are plain shared-memory words (no atomics), and within-column order
// -- which is part of bit-exactness -- is preserved by construction.
//
//   k_mp_rm_size   counts per (column,file): n_plp, bases passing -Q, bytes of
//                  the sequence column, digits of the optional columns;
//                  writes MpFileSz[file][col], the line length and tile totals
//   (scan of tile totals -> byte offset of every tile)
//   k_mp_rm_write  lays the lines out in shared memory with the destination's
//                  16-byte phase, streams the reads again filling sequence and
//                  quality strings through per-column cursors, and leaves
//                  through one cp.async.bulk (TMA) shared->global store.
// Reads that are not of the simple [S]<n>M[S] shape (indels, ref-skips, pads:
// a few % of a WGS run) take the generic per-column functions of plp_core.h.
#pragma once

constexpr int RM_COLS = 256;     // columns per CTA
constexpr int RM_WARPS = 4;
constexpr int RM_SUB = RM_COLS / RM_WARPS;   // columns per warp (two 32-column range groups)
static_assert(RM_SUB == 64, "a warp owns two 32-column range groups");

struct RmList { int32_t n_ovf, lo, n; const int32_t *ovf; };

// reads that can cover the warp's 64 columns: far-reaching list of the first
// group, then the slice from the first group's lo to the second group's hi
__device__ __forceinline__ RmList rm_list(const View &v, int f, int g0)
{
    const int64_t k0 = (int64_t)f * v.n_tiles + g0;
    const int g1 = g0 + 1 < v.n_tiles ? g0 + 1 : g0;
    const int64_t k1 = (int64_t)f * v.n_tiles + g1;
    RmList r;
    const int32_t o0 = v.ovf_off[k0];
    r.n_ovf = v.ovf_off[k0 + 1] - o0;
    r.ovf = v.ovf_idx + o0;
    r.lo = v.tile_lo[k0];
    int32_t hi = v.tile_hi[k1];
    if (hi < r.lo) hi = r.lo;
    r.n = r.n_ovf + (hi - r.lo);
    return r;
}

struct RmSizeSm {   // per-column counters of the current file (shared memory)
    uint32_t nplp[RM_COLS], cnt[RM_COLS], seq[RM_COLS], bp[RM_COLS], bp5[RM_COLS];
};

__global__ void __launch_bounds__(RM_WARPS * 32) k_mp_rm_size(View v, MpConf cf, uint32_t *len_out, MpFileSz *fsz /* [n_files][ncols] */,
                                                              uint32_t *tile_total)
{
    __shared__ RmSizeSm sm;
    __shared__ uint32_t s_len[RM_COLS];
    __shared__ uint32_t s_any[RM_COLS];
    __shared__ uint32_t s_ws[RM_WARPS];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int32_t c0 = (int32_t)blockIdx.x * RM_COLS;          // first column of the CTA
    const int32_t s0 = c0 + w * RM_SUB, s1 = s0 + RM_SUB;      // the warp's columns [s0,s1)
    const uint32_t ends = cf.no_ends ? 0u : 1u;
    for (int j = lane; j < RM_SUB; j += 32) { s_len[w * RM_SUB + j] = 0; s_any[w * RM_SUB + j] = 0; }
    for (int f = 0; f < v.n_files; ++f) {
        for (int j = lane; j < RM_SUB; j += 32) {
            const int c = w * RM_SUB + j;
            sm.nplp[c] = 0; sm.cnt[c] = 0; sm.seq[c] = 0; sm.bp[c] = 0; sm.bp5[c] = 0;
        }
        __syncwarp();
        if (s0 < v.ncols) {
            const RmList rl = rm_list(v, f, s0 >> 5);
            for (int32_t t = 0; t < rl.n; ++t) {
                const int32_t i = t < rl.n_ovf ? rl.ovf[t] : rl.lo + (t - rl.n_ovf);
                ReadDesc d = load_hot(v.desc + i);
                const int32_t a = d.rpos > s0 ? d.rpos : s0, b = d.rend < s1 ? d.rend : s1;   // covered columns of this warp
                if (a >= b) continue;
                if (d.fl & RD_SIMPLE) {
                    const uint32_t qbase = d.qoff + (uint32_t)d.qstart - (uint32_t)d.rpos;   // query index = column + (qstart - rpos)
                    int32_t l5 = 0;
                    if (cf.out_qpos5) { load_cold(d, v.desc + i); l5 = d.l_qseq; }
                    for (int32_t c = a + lane; c < b; c += 32) {
                        const int cl = c - c0;
                        const int q = (int)v.qual[qbase + (uint32_t)c];
                        sm.nplp[cl] += 1;
                        if (q >= cf.min_baseQ) {
                            sm.cnt[cl] += 1;
                            sm.seq[cl] += 1u + (ends & (uint32_t)(c == d.rpos)) * 2u + (ends & (uint32_t)(c == d.rend - 1));
                            if (cf.out_qpos | cf.out_qpos5) {
                                const int32_t qpos = (int32_t)d.qstart + (c - d.rpos);
                                if (cf.out_qpos) sm.bp[cl] += (uint32_t)ndigits32((uint32_t)(qpos + 1)) + 1;
                                if (cf.out_qpos5) { const int32_t q5 = (d.fl & RD_REV) ? l5 - qpos : qpos + 1; sm.bp5[cl] += (uint32_t)(q5 < 0 ? 1 + ndigits32((uint32_t)(-q5)) : ndigits32((uint32_t)q5)) + 1; }
                            }
                        }
                    }
                } else {
                    load_cold(d, v.desc + i);
                    const uint32_t *cg = v.cigar + d.cig_off;
                    for (int32_t c = a + lane; c < b; c += 32) {
                        const int cl = c - c0;
                        Ent e; resolve(v, d, c, e);
                        sm.nplp[cl] += 1;
                        if (ent_qual(v, d, e) >= cf.min_baseQ) {
                            sm.cnt[cl] += 1;
                            sm.seq[cl] += (uint32_t)mp_entry_size(cf, d, cg, e);
                            if (cf.out_qpos) sm.bp[cl] += (uint32_t)ndigits((uint64_t)(e.qpos + 1)) + 1;
                            if (cf.out_qpos5) { const int32_t q5 = qpos5_of(d, e); sm.bp5[cl] += (uint32_t)(q5 < 0 ? 1 + ndigits((uint64_t)(-(int64_t)q5)) : ndigits((uint64_t)q5)) + 1; }
                        }
                    }
                }
                __syncwarp();
            }
        }
        __syncwarp();
        for (int j = lane; j < RM_SUB; j += 32) {
            const int cl = w * RM_SUB + j; const int32_t c = c0 + cl;
            if (c < v.ncols) {
                MpFileSz s; s.nplp = (int32_t)sm.nplp[cl]; s.cnt = (int32_t)sm.cnt[cl]; s.seq_len = sm.seq[cl]; s.bp_len = sm.bp[cl]; s.bp5_len = sm.bp5[cl];
                fsz[(int64_t)f * v.ncols + c] = s;
                s_len[cl] += mp_file_section_len(cf, s);
                s_any[cl] |= (uint32_t)(s.nplp > 0);
            }
        }
        __syncwarp();
    }
    uint32_t tot = 0;
    for (int j = lane; j < RM_SUB; j += 32) {
        const int cl = w * RM_SUB + j; const int32_t c = c0 + cl;
        if (c < v.ncols) {
            uint32_t len = 0;
            if ((s_any[cl] || (cf.all && c < v.ncols_all)) && bed_pass(v, c)) len = mp_head_len(v, c) + s_len[cl] + 1;
            len_out[c] = len;
            tot += len;
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    if (lane == 0) s_ws[w] = tot;
    __syncthreads();
    // totals per 128 columns (the granularity of the tile-offset scan, shared with the column-major write kernel)
    if (threadIdx.x < 2) tile_total[blockIdx.x * 2 + threadIdx.x] = s_ws[2 * threadIdx.x] + s_ws[2 * threadIdx.x + 1];
}

struct RmWriteSm {   // per-column layout of the current file section (offsets into the tile text)
    uint32_t line[RM_COLS];     // running end of what has been laid out for the line
    uint32_t seq[RM_COLS], qual[RM_COLS], mq[RM_COLS], bp[RM_COLS], bp5[RM_COLS];   // next byte to write in each string
    uint32_t n[RM_COLS];        // entries written so far (comma logic of the position lists)
    uint32_t emit[RM_COLS];
};

__global__ void __launch_bounds__(RM_WARPS * 32) k_mp_rm_write(View v, MpConf cf, const uint32_t *len_in, const MpFileSz *fsz,
                                                               const uint64_t *tile_base, char *out, uint32_t smem_cap, int use_tma)
{
    extern __shared__ __align__(16) char s_text[];
    __shared__ RmWriteSm sm;
    __shared__ uint32_t s_ws[RM_WARPS];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int32_t c0 = (int32_t)blockIdx.x * RM_COLS;
    const int32_t s0 = c0 + w * RM_SUB, s1 = s0 + RM_SUB;
    // ---- line offsets inside the tile (each lane owns columns s0+lane and s0+32+lane)
    uint32_t l0 = 0, l1 = 0;
    { const int32_t ca = s0 + lane, cb = s0 + 32 + lane; if (ca < v.ncols) l0 = len_in[ca]; if (cb < v.ncols) l1 = len_in[cb]; }
    uint32_t x0 = l0, x1 = l1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x0, o); if (lane >= o) x0 += y; }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x1, o); if (lane >= o) x1 += y; }
    const uint32_t sum0 = __shfl_sync(0xffffffffu, x0, 31), sum1 = __shfl_sync(0xffffffffu, x1, 31);
    if (lane == 0) s_ws[w] = sum0 + sum1;
    __syncthreads();
    uint32_t wbase = 0, total = 0;
    for (int k = 0; k < RM_WARPS; ++k) { if (k < w) wbase += s_ws[k]; total += s_ws[k]; }
    if (total == 0) return;
    const uint64_t base = tile_base[blockIdx.x * 2];
    const uint32_t phase = (uint32_t)(base & 15);
    const bool in_smem = total + phase <= smem_cap;
    char *text = in_smem ? s_text + phase : out + base;     // very deep tiles format straight into HBM
    {
        const uint32_t off0 = wbase + x0 - l0, off1 = wbase + sum0 + x1 - l1;
        const int cla = w * RM_SUB + lane, clb = cla + 32;
        sm.emit[cla] = l0 != 0; sm.emit[clb] = l1 != 0;
        if (l0) sm.line[cla] = off0 + (uint32_t)(mp_head_write(v, s0 + lane, text + off0) - (text + off0));
        if (l1) sm.line[clb] = off1 + (uint32_t)(mp_head_write(v, s0 + 32 + lane, text + off1) - (text + off1));
    }
    __syncwarp();
    const int nopt = mp_n_opt_cols(cf);
    for (int f = 0; f < v.n_files; ++f) {
        // ---- lay out this file's section of every line: "\tcnt\t" seq "\t" qual ["\t" opt]*
        for (int j = lane; j < RM_SUB; j += 32) {
            const int cl = w * RM_SUB + j; const int32_t c = c0 + cl;
            if (c >= v.ncols || !sm.emit[cl]) continue;
            const MpFileSz s = fsz[(int64_t)f * v.ncols + c];
            char *p = text + sm.line[cl];
            *p++ = '\t'; p += put_u64(p, (uint64_t)s.cnt); *p++ = '\t';
            if (s.nplp == 0) {
                *p++ = '*'; *p++ = '\t'; *p++ = '*';
                for (int k = 0; k < nopt; ++k) { *p++ = '\t'; *p++ = '*'; }
            } else {
                char *ps = p, *pq = ps + (s.seq_len ? s.seq_len : 1) + 1, *pm = pq + (s.cnt ? s.cnt : 1), *pb = pm, *pb5;
                if (cf.out_mapq) pb = pm + 1 + (s.cnt ? s.cnt : 1);
                pb5 = pb;
                if (cf.out_qpos) pb5 = pb + 1 + (s.cnt ? s.bp_len - 1 : 1);
                char *pend = pb5;
                if (cf.out_qpos5) pend = pb5 + 1 + (s.cnt ? s.bp5_len - 1 : 1);
                pq[-1] = '\t';
                if (!s.cnt) {
                    *ps = '*'; *pq = '*';
                    if (cf.out_mapq) { pm[0] = '\t'; pm[1] = '*'; }
                    if (cf.out_qpos) { pb[0] = '\t'; pb[1] = '*'; }
                    if (cf.out_qpos5) { pb5[0] = '\t'; pb5[1] = '*'; }
                } else {
                    if (cf.out_mapq) *pm++ = '\t';
                    if (cf.out_qpos) *pb++ = '\t';
                    if (cf.out_qpos5) *pb5++ = '\t';
                }
                sm.seq[cl] = (uint32_t)(ps - text); sm.qual[cl] = (uint32_t)(pq - text); sm.mq[cl] = (uint32_t)(pm - text);
                sm.bp[cl] = (uint32_t)(pb - text); sm.bp5[cl] = (uint32_t)(pb5 - text); sm.n[cl] = 0;
                p = pend;
                for (int k = 0; k < cf.n_star_cols; ++k) { *p++ = '\t'; *p++ = '*'; }
            }
            sm.line[cl] = (uint32_t)(p - text);
        }
        __syncwarp();
        // ---- stream the reads of this file over the warp's columns, in file order
        if (s0 < v.ncols) {
            const RmList rl = rm_list(v, f, s0 >> 5);
            for (int32_t t = 0; t < rl.n; ++t) {
                const int32_t i = t < rl.n_ovf ? rl.ovf[t] : rl.lo + (t - rl.n_ovf);
                ReadDesc d = load_hot(v.desc + i);
                const int32_t a = d.rpos > s0 ? d.rpos : s0, b = d.rend < s1 ? d.rend : s1;
                if (a >= b) continue;
                const bool rev = d.fl & RD_REV;
                if (d.fl & RD_SIMPLE) {
                    const uint32_t qbase = d.qoff + (uint32_t)d.qstart - (uint32_t)d.rpos;
                    int32_t l5 = 0;
                    if (cf.out_qpos5) { load_cold(d, v.desc + i); l5 = d.l_qseq; }
                    const char mqc = (char)(d.mapq > 93 ? 126 : d.mapq + 33);
                    for (int32_t c = a + lane; c < b; c += 32) {
                        const int cl = c - c0;
                        const uint32_t qi = qbase + (uint32_t)c;
                        const int q = (int)v.qual[qi];
                        int ch = (v.seq4[qi >> 1] >> ((~qi & 1) << 2)) & 0xf;
                        if (q < cf.min_baseQ || !sm.emit[cl]) continue;
                        if (v.ref) {
                            int rb = 15;
                            if ((int64_t)c < v.ref_len_rel) { const int64_t ri = (int64_t)c - v.ref_off; if (ri >= 0 && ri < v.ref_n) rb = nt16_of((unsigned char)v.ref[ri]); }
                            if (ch == rb) ch = 0;
                        }
                        char *ps = text + sm.seq[cl];
                        if (!cf.no_ends && c == d.rpos) { *ps++ = '^'; *ps++ = mqc; }
                        *ps++ = base_char(ch, rev);
                        if (!cf.no_ends && c == d.rend - 1) *ps++ = '$';
                        sm.seq[cl] = (uint32_t)(ps - text);
                        text[sm.qual[cl]++] = (char)(q + 33 < 126 ? q + 33 : 126);
                        if (cf.out_mapq | cf.out_qpos | cf.out_qpos5) {
                            const uint32_t n = sm.n[cl]++;
                            const int32_t qpos = (int32_t)d.qstart + (c - d.rpos);
                            if (cf.out_mapq) { const int m = d.mapq + 33; text[sm.mq[cl]++] = (char)(m > 126 ? 126 : m); }
                            if (cf.out_qpos) { char *pb = text + sm.bp[cl]; if (n) *pb++ = ','; pb += put_i32(pb, qpos + 1); sm.bp[cl] = (uint32_t)(pb - text); }
                            if (cf.out_qpos5) { char *pb = text + sm.bp5[cl]; if (n) *pb++ = ','; pb += put_i32(pb, rev ? l5 - qpos : qpos + 1); sm.bp5[cl] = (uint32_t)(pb - text); }
                        }
                    }
                } else {
                    load_cold(d, v.desc + i);
                    const uint32_t *cg = v.cigar + d.cig_off;
                    for (int32_t c = a + lane; c < b; c += 32) {
                        const int cl = c - c0;
                        if (!sm.emit[cl]) continue;
                        Ent e; resolve(v, d, c, e);
                        const int q = ent_qual(v, d, e);
                        if (q < cf.min_baseQ) continue;
                        char *ps = text + sm.seq[cl];
                        ps += mp_entry_write(v, cf, d, cg, e, c, ps);
                        sm.seq[cl] = (uint32_t)(ps - text);
                        text[sm.qual[cl]++] = (char)(q + 33 < 126 ? q + 33 : 126);
                        if (cf.out_mapq | cf.out_qpos | cf.out_qpos5) {
                            const uint32_t n = sm.n[cl]++;
                            if (cf.out_mapq) { const int m = d.mapq + 33; text[sm.mq[cl]++] = (char)(m > 126 ? 126 : m); }
                            if (cf.out_qpos) { char *pb = text + sm.bp[cl]; if (n) *pb++ = ','; pb += put_i32(pb, e.qpos + 1); sm.bp[cl] = (uint32_t)(pb - text); }
                            if (cf.out_qpos5) { char *pb = text + sm.bp5[cl]; if (n) *pb++ = ','; pb += put_i32(pb, qpos5_of(d, e)); sm.bp5[cl] = (uint32_t)(pb - text); }
                        }
                    }
                }
                __syncwarp();
            }
        }
        __syncwarp();
    }
    for (int j = lane; j < RM_SUB; j += 32) { const int cl = w * RM_SUB + j; if (c0 + cl < v.ncols && sm.emit[cl]) text[sm.line[cl]] = '\n'; }
    if (!in_smem) return;
    __syncthreads();
    char *g = out + base; const char *sb = s_text + phase;
    const uint32_t head = min(total, (16u - phase) & 15u);
    const uint32_t body = (total - head) & ~15u;
    const uint32_t tail = total - head - body;
    if (threadIdx.x < head) g[threadIdx.x] = sb[threadIdx.x];
    if (threadIdx.x < tail) g[head + body + threadIdx.x] = sb[head + body + threadIdx.x];
    if (body) {
        if (use_tma) { if (threadIdx.x == 0) bulk_store_s2g(g + head, sb + head, body); }
        else {
            const uint4 *src = reinterpret_cast<const uint4 *>(sb + head);
            uint4 *dst = reinterpret_cast<uint4 *>(g + head);
            for (uint32_t k = threadIdx.x; k < body / 16; k += RM_WARPS * 32) dst[k] = src[k];
        }
    }
}
