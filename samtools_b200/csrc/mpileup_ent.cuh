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

__global__ void __launch_bounds__(256) k_mp_entries(View v, MpConf cf, int64_t n_reads, const uint8_t *refc /* per staged reference byte: nt16 code | 0..4 code << 4; null without a FASTA */,
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
                uint32_t ent[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const uint32_t qi = g + (uint32_t)k;
                    const uint32_t q = ((k < 4 ? qq.x : qq.y) >> (8 * (k & 3))) & 0xffu;
                    const uint32_t code = (s4 >> (8 * (k >> 1) + ((k & 1) ? 0 : 4))) & 0xfu;
                    const bool in = qi >= lo && qi < hi;
                    const int32_t c = d.rpos + (int32_t)(qi - q0);
                    uint32_t rb = 0x10u;
                    if (refc && in) rb = ref_nt16_at(v, refc, c);
                    uint32_t fl = 0;
                    if (ends) fl = (qi == q0 ? 0x80u : 0u) | (qi == qtail ? 0x8000u : 0u);
                    const uint32_t x = ent_plain(q, code, rb, rev, minq, fl, s_tab);
                    ent[k] = x;
                    if (in) {
                        if (!x) atomicAdd(&fail[c], 1u);
                        else if (fl) atomicAdd(&extra[c], ((fl & 0x80u) ? 2u : 0u) + ((fl & 0x8000u) ? 1u : 0u));
                    }
                }
                if (g >= lo && g + 8u <= hi) {
                    uint4 w;
                    w.x = ent[0] | ent[1] << 16; w.y = ent[2] | ent[3] << 16; w.z = ent[4] | ent[5] << 16; w.w = ent[6] | ent[7] << 16;
                    *reinterpret_cast<uint4 *>(E + g) = w;
                } else {
#pragma unroll
                    for (int k = 0; k < 8; ++k) { const uint32_t qi = g + (uint32_t)k; if (qi >= lo && qi < hi) E[qi] = (uint16_t)ent[k]; }
                }
            }
        } else {
            load_cold(d, v.desc + i);
            unsigned long long eo = 0;
            if (lane == 0) { eo = atomicAdd(e2_cursor, (unsigned long long)(uint32_t)(d.rend - d.rpos)); desc_rw[i].pad_ = (uint32_t)eo; }
            eo = __shfl_sync(0xffffffffu, eo, 0);
            for (int32_t c = a + lane; c < b; c += 32) {
                const uint32_t rb = refc ? ref_nt16_at(v, refc, c) : 0x10u;
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
__global__ void __launch_bounds__(TILE) k_mp_gather(MpEntFmt fmt, const uint32_t *len, const MpFileSz *st, const uint64_t *tile_base,
                                                    char *out, uint32_t smem_cap, int use_tma)
{
    text_write_tile(fmt, fmt.v.ncols, len, st, tile_base, out, smem_cap, use_tma);
}
