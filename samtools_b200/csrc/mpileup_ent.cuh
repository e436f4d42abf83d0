// mpileup_ent.cuh -- default single-file mpileup text path: entry strings + gather.
//
//   k_mp_entries  READ-MAJOR, one warp per read, lanes along the read.  Simple reads ([S]<n>M[S]): each lane takes
//                 eight consecutive query bases -- one aligned 8-byte quality load, one aligned 4-byte base load --
//                 turns them into eight 16-bit entries (plp_core.h "entry strings": sequence character, quality
//                 character, "^"/"$" flags; 0 = fails -Q) and stores them with one 16-byte store at the index of the
//                 quality bytes.  Other reads go column by column through the generic cursor into a slice of a second
//                 array.  The same pass feeds the line-length sums of mpileup_ss.cuh (coverage difference array,
//                 failing bases and extra bytes per column): it IS the size pass.
//   k_mp_gather   COLUMN-MAJOR, one thread per reference position: walks the reads of its 32-column slice in file
//                 order and appends the non-empty entries to its line (2 bytes per entry, no decoding, no CIGAR walk);
//                 the tile leaves through the TMA bulk store of text_write_tile.
// Replaces the per-(read, column) formatting loop of the round-1 write kernel (~100 instructions per pair).
#pragma once

__device__ __forceinline__ uint32_t ref_nt16_at(const View &v, const uint8_t *refc, int32_t c)
{
    if ((int64_t)c < v.ref_len_rel) { const int64_t ri = (int64_t)c - v.ref_off; if (ri >= 0 && ri < v.ref_n) return (uint32_t)refc[ri] & 0xfu; }
    return 15u;
}

// eight consecutive bases of a simple read that all lie inside the read and the window: no range tests, the "^" / "$"
// flags are patched in afterwards by the one lane that holds the read's first / last base
template <bool HAS_REF>
__device__ __forceinline__ uint32_t ent_group8(const View &v, const uint8_t *refc, uint2 qq, uint32_t s4, int32_t c_of_g, uint32_t rev, int minq,
                                               const uint8_t *tab, uint32_t (&ent)[8])
{
    uint32_t failmask = 0;
    const uint8_t *t = tab + rev * 16u;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint32_t q = ((k < 4 ? qq.x : qq.y) >> (8 * (k & 3))) & 0xffu;
        uint32_t code = (s4 >> (8 * (k >> 1) + ((k & 1) ? 0 : 4))) & 0xfu;
        if (HAS_REF) { if (code == ref_nt16_at(v, refc, c_of_g + k)) code = 0; }
        uint32_t x = (uint32_t)t[code] | umin32(q + 33u, 126u) << 8;
        if ((int)q < minq) { x = 0; failmask |= 1u << k; }
        ent[k] = x;
    }
    return failmask;
}

template <bool HAS_REF>
__global__ void __launch_bounds__(256) k_mp_entries(View v, MpConf cf, int64_t n_reads, const uint8_t *refc /* per staged reference byte: nt16 code | 0..4 code << 4 */,
                                                    int32_t *diff, uint32_t *fail, uint32_t *extra, uint16_t *E, uint16_t *E2,
                                                    unsigned long long *e2_cursor, ReadDesc *desc_rw)
{
    __shared__ uint8_t s_tab[32];
    if (threadIdx.x < 32) s_tab[threadIdx.x] = (uint8_t)".ACMGRSVTWYHKDBN,acmgrsvtwyhkdbn"[threadIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const bool ends = !cf.no_ends;
    const int minq = cf.min_baseQ;
    for (int64_t i = warp; i < n_reads; i += n_warps) {
        ReadDesc d = load_hot(v.desc + i);
        if (d.rend <= d.rpos) continue;                         // filtered read
        const int32_t a = d.rpos > 0 ? d.rpos : 0, b = d.rend < v.ncols ? d.rend : v.ncols;   // columns inside the window
        if (a >= b) continue;
        if (lane == 0) { atomicAdd(&diff[a], 1); atomicAdd(&diff[b], -1); }
        const uint32_t rev = (d.fl & RD_REV) ? 1u : 0u;
        if (d.fl & RD_SIMPLE) {
            const uint32_t q0 = d.qoff + (uint32_t)d.qstart;                 // query index of column rpos
            const uint32_t lo = q0 + (uint32_t)(a - d.rpos), hi = q0 + (uint32_t)(b - d.rpos);
            const uint32_t qtail = q0 + (uint32_t)(d.rend - d.rpos) - 1u;
            for (uint32_t g = (lo & ~7u) + 8u * (uint32_t)lane; g < hi; g += 256u) {
                const uint2 qq = __ldg(reinterpret_cast<const uint2 *>(v.qual + g));
                const uint32_t s4 = __ldg(reinterpret_cast<const uint32_t *>(v.seq4 + (g >> 1)));
                const int32_t c_of_g = d.rpos + (int32_t)(g - q0);          // column of query index g
                uint32_t ent[8];
                // all eight bases of the group are formatted (bytes beyond the read's ends belong to its neighbours in the
                // arrays and are harmless to read); the ones inside [lo,hi) are kept
                uint32_t failmask = ent_group8<HAS_REF>(v, refc, qq, s4, c_of_g, rev, minq, s_tab, ent);
                const uint32_t kb = lo > g ? lo - g : 0u, ke = hi - g < 8u ? hi - g : 8u;
                const uint32_t vmask = ((1u << ke) - 1u) & ~((1u << kb) - 1u);
                failmask &= vmask;
                if (ends) {   // "^"+mapq at the read's first base, "$" at its last: at most one lane each
                    const uint32_t kh = q0 - g, kt = qtail - g;
                    if (kh < 8u && ((vmask & ~failmask) >> kh) & 1u) {
#pragma unroll
                        for (int k = 0; k < 8; ++k) if ((uint32_t)k == kh) ent[k] |= 0x80u;
                        atomicAdd(&extra[c_of_g + (int32_t)kh], 2u);
                    }
                    if (kt < 8u && ((vmask & ~failmask) >> kt) & 1u) {
#pragma unroll
                        for (int k = 0; k < 8; ++k) if ((uint32_t)k == kt) ent[k] |= 0x8000u;
                        atomicAdd(&extra[c_of_g + (int32_t)kt], 1u);
                    }
                }
                while (failmask) { const int k = __ffs(failmask) - 1; failmask &= failmask - 1u; atomicAdd(&fail[c_of_g + k], 1u); }
                if (vmask == 0xffu) {
                    uint4 w;
                    w.x = ent[0] | ent[1] << 16; w.y = ent[2] | ent[3] << 16; w.z = ent[4] | ent[5] << 16; w.w = ent[6] | ent[7] << 16;
                    *reinterpret_cast<uint4 *>(E + g) = w;
                } else {
#pragma unroll
                    for (int k = 0; k < 8; ++k) if ((vmask >> k) & 1u) E[g + (uint32_t)k] = (uint16_t)ent[k];
                }
            }
        } else {
            load_cold(d, v.desc + i);
            unsigned long long eo = 0;
            if (lane == 0) { eo = atomicAdd(e2_cursor, (unsigned long long)(uint32_t)(d.rend - d.rpos)); desc_rw[i].pad_ = (uint32_t)eo; }
            eo = __shfl_sync(0xffffffffu, eo, 0);
            for (int32_t c = a + lane; c < b; c += 32) {
                const uint32_t rb = HAS_REF ? ref_nt16_at(v, refc, c) : 0x10u;
                uint32_t xb;
                const uint32_t x = ent_generic(v, cf, d, c, rb, s_tab, xb);
                E2[eo + (uint32_t)(c - d.rpos)] = (uint16_t)x;
                if (!x) atomicAdd(&fail[c], 1u);
                else if (xb) atomicAdd(&extra[c], xb);
            }
        }
    }
}

struct MpEntFmt {
    View v; MpConf cf; const uint16_t *E, *E2;
    typedef MpFileSz State;
    __device__ __forceinline__ void write(int32_t c, const State &s, char *p) const { mp_line_write_ent(v, cf, c, s, p, E, E2); }
};

// The gather of one line by its thread, with the warp's help: the 32 columns of a warp share one read slice, so the
// warp stages the slice's descriptors (first 16 bytes: rpos, rend, qoff, qstart|mapq|flags) in shared memory, 32 at a time
// with one coalesced load, and every thread then walks them with broadcast shared-memory loads -- no per-thread descriptor
// loads from global memory, and the entry loads of EIGHT reads are in flight before the first append (the entry of a read
// that is not over the column, or not of the simple shape, is entry 0 of the array, discarded).
// Same bytes as mp_line_write_ent (plp_core.h), which stays the reference implementation (emulation harness, deep tiles).
template <bool OUT_MAPQ>
__device__ __forceinline__ void gather_line_warp(const View &v, const MpConf &cf, int32_t c, bool active, const MpFileSz &s, char *p,
                                                 const uint16_t *E, const uint16_t *E2, uint4 *s_desc)
{
    const int lane = threadIdx.x & 31;
    EntCur cur; cur.ps = nullptr; cur.pq = nullptr; cur.pm = nullptr;
    if (active) cur = ent_layout(v, cf, c, s, p);
    const bool on = cur.ps != nullptr;
    uint32_t so = on ? (uint32_t)(cur.ps - p) : 0u, qo = on ? (uint32_t)(cur.pq - p) : 0u, mo = on ? (uint32_t)(cur.pm - p) : 0u;
    const ReadRange rr = read_range(v, 0, c >> 5);            // the same for the 32 lanes
    const uint32_t kSimple = (uint32_t)RD_SIMPLE << 24;
    // append entry e of read i.  The common entry (no flag, not special) is two byte stores and two cursor bumps, with no branch
    // on e == 0: an empty entry stores a byte at the cursor without advancing it, the next real entry overwrites it -- and the two
    // bytes just behind the sequence / quality strings (a tab each) are written AFTER the gather (see the end of this function)
    auto emit = [&](uint32_t pk, int32_t i, uint32_t e) {
        if (e & 0x8080u) {                                   // "^" / "$" flags or the special marker: rare
            const uint32_t mapq = (pk >> 16) & 0xffu;
            if (e == ENT_SPECIAL) {
                int q;
                so += (uint32_t)ent_special(v, cf, i, c, p + so, q);
                p[qo++] = (char)(q + 33 < 126 ? q + 33 : 126);
            } else {
                if (e & 0x80u) { p[so++] = '^'; p[so++] = (char)(mapq > 93u ? 126u : mapq + 33u); }
                p[so++] = (char)(e & 0x7fu);
                if (e & 0x8000u) p[so++] = '$';
                p[qo++] = (char)((e >> 8) & 0x7fu);
            }
            if (OUT_MAPQ) p[mo++] = (char)umin32(mapq + 33u, 126u);
        } else {
            const uint32_t adv = e != 0u;
            p[so] = (char)e; so += adv;
            p[qo] = (char)(e >> 8); qo += adv;
            if (OUT_MAPQ) { p[mo] = (char)umin32(((pk >> 16) & 0xffu) + 33u, 126u); mo += adv; }
        }
    };
    auto one = [&](const uint4 &d, int32_t i) {      // general route for one read
        const uint32_t rel = (uint32_t)(c - (int32_t)d.x);
        if (rel >= (uint32_t)((int32_t)d.y - (int32_t)d.x)) return;
        const uint32_t e = (d.w & kSimple) ? E[d.z + (d.w & 0xffffu) + rel] : E2[v.desc[i].pad_ + rel];
        emit(d.w, i, e);
    };
    if (on) for (int32_t t = 0; t < rr.n_ovf; ++t) { const int32_t i = rr.ovf[t]; one(__ldg(reinterpret_cast<const uint4 *>(v.desc + i)), i); }
    const int32_t hi = rr.lo + (rr.n - rr.n_ovf);
    for (int32_t base = rr.lo; base < hi; base += 32) {
        const int32_t cnt = hi - base < 32 ? hi - base : 32;
        __syncwarp();
        if (lane < cnt) s_desc[lane] = __ldg(reinterpret_cast<const uint4 *>(v.desc + base + lane));
        __syncwarp();
        if (!on) continue;
        int32_t r = 0;
        for (; r + 8 <= cnt; r += 8) {
            uint32_t pk[8], rel[8], e[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint4 d = s_desc[r + k];
                pk[k] = d.w;
                rel[k] = (uint32_t)(c - (int32_t)d.x);
                const bool in = rel[k] < (uint32_t)((int32_t)d.y - (int32_t)d.x);
                const bool fast = in && (d.w & kSimple);
                e[k] = E[fast ? d.z + (d.w & 0xffffu) + rel[k] : 0u];
                if (!fast) e[k] = in ? ENT_SPECIAL + 1u : 0u;       // in, not simple: marker resolved below through the second array
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (e[k] == ENT_SPECIAL + 1u) e[k] = E2[v.desc[base + r + k].pad_ + rel[k]];
                emit(pk[k], base + r + k, e[k]);
            }
        }
        for (; r < cnt; ++r) one(s_desc[r], base + r);
    }
    if (on) {   // the separators an empty trailing entry may have scribbled on
        cur.pq[-1] = '\t';                                            // behind the sequence string
        if (OUT_MAPQ) { cur.pm[-1] = '\t'; p[mo] = (cf.n_star_cols ? '\t' : '\n'); }
        else p[qo] = (cf.n_star_cols ? '\t' : '\n');                   // behind the quality string: next column's tab or the newline
    }
}

template <int MIN_CTAS, bool OUT_MAPQ>
__global__ void __launch_bounds__(TILE, MIN_CTAS) k_mp_gather(MpEntFmt fmt, const uint32_t *len_in, const MpFileSz *st_in, const uint64_t *tile_base,
                                                              char *out, uint32_t smem_cap, int use_tma)
{
    extern __shared__ __align__(16) char s_text[];
    __shared__ uint32_t s_ws[TILE / 32];
    __shared__ uint4 s_desc[TILE / 32][32];
    const int32_t ncols = fmt.v.ncols;
    const int32_t c = (int32_t)blockIdx.x * TILE + (int32_t)threadIdx.x;
    MpFileSz stt;
    uint32_t len = 0;
    if (c < ncols) { len = len_in[c]; if (len) stt = st_in[c]; }
    uint32_t total;
    const uint32_t off = block_excl_scan<TILE>(len, s_ws, total);
    if (total == 0) return;
    const uint64_t base = tile_base[blockIdx.x];
    const uint32_t phase = (uint32_t)(base & 15);
    if (total + phase <= smem_cap) {
        char *sb = s_text + phase;
        gather_line_warp<OUT_MAPQ>(fmt.v, fmt.cf, c < ncols ? c : ncols - 1, len != 0, stt, sb + off, fmt.E, fmt.E2, s_desc[threadIdx.x >> 5]);
        __syncthreads();
        // ragged head (to the next 16 B boundary of the destination), aligned body, ragged tail
        char *g = out + base;
        const uint32_t head = min(total, (16u - phase) & 15u);
        const uint32_t body = (total - head) & ~15u;
        const uint32_t tail = total - head - body;
        if (threadIdx.x < head) g[threadIdx.x] = sb[threadIdx.x];
        if (threadIdx.x < tail) g[head + body + threadIdx.x] = sb[head + body + threadIdx.x];
        if (body) {
            if (use_tma) {
                if (threadIdx.x == 0) bulk_store_s2g(g + head, sb + head, body);
            } else {
                const uint4 *src = reinterpret_cast<const uint4 *>(sb + head);
                uint4 *dst = reinterpret_cast<uint4 *>(g + head);
                for (uint32_t i = threadIdx.x; i < body / 16; i += TILE) dst[i] = src[i];
            }
        }
    } else if (len) {
        fmt.write(c, stt, out + base + off);   // very deep tile: format straight into HBM
    }
}
