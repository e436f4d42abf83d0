// mpileup_w4.cuh -- write pass with FOUR ADJACENT COLUMNS PER THREAD (single-file, no optional
// columns: the standard `mpileup` line).  Same bytes as mp_line_write (plp_core.h); the point
// is the instruction budget.  With one column per thread every (read, column) entry pays for
// a descriptor load, a range test and two byte loads of its own (~100 warp-instructions per
// warp iteration, one entry per lane).  A thread that owns columns 4t..4t+3 shares the
// descriptor, the range test and ONE 32-bit quality word + ONE 32-bit base word among four
// entries, and a warp's lanes read consecutive words (fully coalesced 128 B requests).
//
// CTA = W4_THREADS threads = W4_COLS columns = two 128-column tiles of the offset scan.
// Lines are formatted into shared memory laid out with the destination's 16-byte phase and
// leave through one cp.async.bulk (TMA) shared->global store, as in the other write kernels.
#pragma once

constexpr int W4_THREADS = 64;
constexpr int W4_COLS = W4_THREADS * 4;   // 256

__device__ __forceinline__ uint32_t ld_u32_unaligned(const uint8_t *base, uint32_t off)
{
    const uint32_t *p = reinterpret_cast<const uint32_t *>(base + (off & ~3u));
    return __funnelshift_r(__ldg(p), __ldg(p + 1), (off & 3u) * 8u);
}

__global__ void __launch_bounds__(W4_THREADS) k_mpileup_write4(View v, MpConf cf, const uint32_t *len_in, const MpFileSz *fsz,
                                                               const uint64_t *tile_base, char *out, uint32_t smem_cap, int use_tma)
{
    extern __shared__ __align__(16) char s_text[];
    __shared__ uint32_t s_ws[W4_THREADS / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int32_t cb = (int32_t)blockIdx.x * W4_COLS + (int32_t)threadIdx.x * 4;   // this thread's first column
    uint32_t len[4] = {0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < 4; ++k) if (cb + k < v.ncols) len[k] = len_in[cb + k];
    const uint32_t mine = len[0] + len[1] + len[2] + len[3];
    uint32_t x = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) s_ws[w] = x;
    __syncthreads();
    uint32_t wbase = 0, total = 0;
#pragma unroll
    for (int k = 0; k < W4_THREADS / 32; ++k) { if (k < w) wbase += s_ws[k]; total += s_ws[k]; }
    if (total == 0) return;
    const uint64_t base = tile_base[blockIdx.x * 2];
    const uint32_t phase = (uint32_t)(base & 15);
    const bool in_smem = total + phase <= smem_cap;
    char *text = in_smem ? s_text + phase : out + base;
    uint32_t off = wbase + x - mine;
    // ---- lay out the four lines: header, "\tcnt\t", placeholders, separators, newline
    uint32_t ps[4], pq[4];     // next byte of the sequence / quality string (offsets into text); 0xffffffff: nothing to fill
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        ps[k] = pq[k] = 0xffffffffu;
        if (!len[k]) continue;
        const int32_t c = cb + k;
        const MpFileSz s = fsz[c];
        char *p0 = text + off, *p = mp_head_write(v, c, p0);
        *p++ = '\t'; p += put_u64(p, (uint64_t)s.cnt); *p++ = '\t';
        if (s.nplp == 0) { *p++ = '*'; *p++ = '\t'; *p++ = '*'; }
        else if (s.cnt == 0) { *p++ = '*'; *p++ = '\t'; *p++ = '*'; }
        else {
            ps[k] = (uint32_t)(p - text);
            p += s.seq_len; *p++ = '\t';
            pq[k] = (uint32_t)(p - text);
            p += s.cnt;
        }
        for (int j = 0; j < cf.n_star_cols; ++j) { *p++ = '\t'; *p++ = '*'; }
        *p = '\n';
        off += len[k];
    }
    // ---- reads that can cover this warp's 128 columns, in file order
    const int32_t wc0 = (int32_t)blockIdx.x * W4_COLS + w * 128;
    if (wc0 < v.ncols) {
        const int g0 = wc0 >> 5;
        const int g3 = g0 + 3 < v.n_tiles ? g0 + 3 : v.n_tiles - 1;
        const int32_t o0 = v.ovf_off[g0], n_ovf = v.ovf_off[g0 + 1] - o0;
        const int32_t lo = v.tile_lo[g0];
        int32_t hi = v.tile_hi[g3]; if (hi < lo) hi = lo;
        const bool ends = !cf.no_ends;
        const int minq = cf.min_baseQ;
        int rb[4] = {-1, -1, -1, -1};
        if (v.ref) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                rb[k] = 15;
                const int64_t c = cb + k;
                if (c < v.ref_len_rel) { const int64_t ri = c - v.ref_off; if (ri >= 0 && ri < v.ref_n) rb[k] = nt16_of((unsigned char)v.ref[ri]); }
            }
        }
        auto body = [&](ReadDesc d, int32_t i) {
            const int32_t a = d.rpos > cb ? d.rpos : cb, b = d.rend < cb + 4 ? d.rend : cb + 4;
            if (a >= b) return;
            if (d.fl & RD_SIMPLE) {
                const uint32_t qi0 = d.qoff + (uint32_t)d.qstart + (uint32_t)(cb - d.rpos);   // query index of column cb (may wrap for columns left of the read)
                const bool rev = d.fl & RD_REV;
                uint32_t qw, sw;
                if (a == cb && b == cb + 4) {                  // all four columns inside the read: one word of qualities, one of bases
                    qw = ld_u32_unaligned(v.qual, qi0);
                    sw = ld_u32_unaligned(v.seq4, qi0 >> 1);
                } else {
                    qw = 0; sw = 0;
                    for (int32_t c = a; c < b; ++c) {
                        const int k = c - cb; const uint32_t qi = qi0 + (uint32_t)k;
                        qw |= (uint32_t)v.qual[qi] << (8 * k);
                        // place the nibble where the word path would have it
                        const uint32_t nb = (v.seq4[qi >> 1] >> ((~qi & 1) << 2)) & 0xf;
                        const uint32_t pos = (uint32_t)k + (qi0 & 1);
                        sw |= nb << (8 * (pos >> 1) + ((~pos & 1) << 2));
                    }
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int32_t c = cb + k;
                    if (c < a || c >= b || ps[k] == 0xffffffffu) continue;
                    const int q = (int)((qw >> (8 * k)) & 0xff);
                    if (q < minq) continue;
                    const uint32_t pos = (uint32_t)k + (qi0 & 1);
                    int ch = (int)((sw >> (8 * (pos >> 1) + ((~pos & 1) << 2))) & 0xf);
                    if (ch == rb[k]) ch = 0;
                    char *p = text + ps[k];
                    if (ends && c == d.rpos) { *p++ = '^'; *p++ = (char)(d.mapq > 93 ? 126 : d.mapq + 33); }
                    *p++ = base_char(ch, rev);
                    if (ends && c == d.rend - 1) *p++ = '$';
                    ps[k] = (uint32_t)(p - text);
                    text[pq[k]++] = (char)(q + 33 < 126 ? q + 33 : 126);
                }
            } else {
                load_cold(d, v.desc + i);
                const uint32_t *cg = v.cigar + d.cig_off;
                for (int32_t c = a; c < b; ++c) {
                    const int k = c - cb;
                    // (k is not a compile-time constant here: go through local copies)
                    uint32_t psk = k == 0 ? ps[0] : k == 1 ? ps[1] : k == 2 ? ps[2] : ps[3];
                    uint32_t pqk = k == 0 ? pq[0] : k == 1 ? pq[1] : k == 2 ? pq[2] : pq[3];
                    if (psk == 0xffffffffu) continue;
                    Ent e; resolve(v, d, c, e);
                    const int q = ent_qual(v, d, e);
                    if (q < minq) continue;
                    psk += (uint32_t)mp_entry_write(v, cf, d, cg, e, c, text + psk);
                    text[pqk++] = (char)(q + 33 < 126 ? q + 33 : 126);
                    if (k == 0) { ps[0] = psk; pq[0] = pqk; } else if (k == 1) { ps[1] = psk; pq[1] = pqk; }
                    else if (k == 2) { ps[2] = psk; pq[2] = pqk; } else { ps[3] = psk; pq[3] = pqk; }
                }
            }
        };
        for (int32_t t = 0; t < n_ovf; ++t) { const int32_t i = v.ovf_idx[o0 + t]; body(load_hot(v.desc + i), i); }
        if (lo < hi) {
            ReadDesc dn = load_hot(v.desc + lo);
            for (int32_t i = lo; i < hi; ++i) {
                const ReadDesc d = dn;
                if (i + 1 < hi) dn = load_hot(v.desc + i + 1);
                body(d, i);
            }
        }
    }
    if (!in_smem) return;
    __syncthreads();
    char *g = out + base; const char *sb = s_text + phase;
    const uint32_t head = min(total, (16u - phase) & 15u);
    const uint32_t body_b = (total - head) & ~15u;
    const uint32_t tail = total - head - body_b;
    if (threadIdx.x < head) g[threadIdx.x] = sb[threadIdx.x];
    if (threadIdx.x < tail) g[head + body_b + threadIdx.x] = sb[head + body_b + threadIdx.x];
    if (body_b) {
        if (use_tma) { if (threadIdx.x == 0) bulk_store_s2g(g + head, sb + head, body_b); }
        else {
            const uint4 *src = reinterpret_cast<const uint4 *>(sb + head);
            uint4 *dst = reinterpret_cast<uint4 *>(g + head);
            for (uint32_t k = threadIdx.x; k < body_b / 16; k += W4_THREADS) dst[k] = src[k];
        }
    }
}
