// mpileup_sr.cuh -- staged-reads write pass for the standard single-file mpileup line.
//
// k_mpileup_write (thread per column, every thread walking the reads in global memory) spends
// ~115 instructions per (read, 32-column group): descriptor unpacking, 64-bit address
// arithmetic, dependent byte loads and the software pipeline that hides them.  Here the CTA
// stages what its 128 columns need in shared memory first (plp_core.h, "building blocks of the
// staged-reads write pass"):
//   A  warp per read, 4 reads per warp and 16-read chunk: the descriptors are fetched by 4 lanes
//      one chunk ahead; the part of the read's quality string and base nibbles that lies over
//      the tile is copied with asynchronous, coalesced word copies (cp.async, <= 33 + 17 words)
//      into the buffer that is NOT being consumed, and one 32-bit slot word describes the read
//      relative to the tile.  No per-base work, no register staging, one barrier per chunk.
//   B  thread = column walks the slots: slot word (broadcast LDS), coverage test, one byte LDS
//      for the quality, one for the base; appends to its line, which is staged in shared memory
//      exactly as in k_mpileup_write and leaves through the same cp.async.bulk (TMA) store.
// Reads that are not [S]<n>M[S] are not staged; their columns call the generic functions (out of
// line, so the common loop stays small).  Within-column order is the slot order = file order.
// Used when the line has no position columns (-O) and one input file; anything else takes
// k_mpileup_write.  `test_c2_size_properties` and the golden tests run both.
#pragma once

constexpr int SR_SLOTS = 16;          // reads staged per chunk (two chunks in flight: one being consumed, one being copied)
constexpr int SR_WSLOTS = SR_SLOTS / (TILE / 32);   // slots staged by one warp per chunk

// byte access through 32-bit shared-window addresses (keeps the append loop free of generic-address arithmetic)
__device__ __forceinline__ void sts8(uint32_t addr, uint32_t val)
{
    asm volatile("st.shared.u8 [%0], %1;" :: "r"(addr), "r"(val) : "memory");
}
__device__ __forceinline__ uint32_t lds8(uint32_t addr)
{
    uint32_t r;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(r) : "r"(addr));
    return r;
}
// asynchronous 4-byte global->shared copy (LDGSTS): no register staging, the warp does not wait
__device__ __forceinline__ void cp_async4(uint32_t saddr, const void *g)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.wait_all;" ::: "memory");
}

__device__ __noinline__ int sr_slow_dev(const View &v, const MpConf &cf, int32_t i, int32_t c, char *ps, int *q)
{
    int qq;
    const int n = sr_slow_entry(v, cf, i, c, ps, qq);
    *q = qq;
    return n;
}
__device__ __noinline__ void sr_deep_line(const View &v, const MpConf &cf, int32_t c, const MpFileSz &s, char *p)
{
    mp_line_write(v, cf, c >> 5, c, s, p);
}

struct SrSm {
    uint32_t q[2][SR_SLOTS][SR_QROW];   // staged quality bytes
    uint32_t s[2][SR_SLOTS][SR_SROW];   // staged base nibbles
    // per slot: x = slot word (sr_meta), y = byte offset (from q) of tile column 0's quality, z = nibble offset (from s) of
    // tile column 0's base -- both may point before the row, only covered columns are read --, w = read index (generic path)
    int4 meta[2][SR_SLOTS];
    uint8_t mq[2][SR_SLOTS];            // mapq character
    uint8_t tab[32];                    // ".ACMGRSVTWYHKDBN,acmgrsvtwyhkdbn"
};

__global__ void __launch_bounds__(TILE) k_mp_sr_write(const __grid_constant__ View v, const __grid_constant__ MpConf cf, const uint32_t *len_in,
                                                      const MpFileSz *st_in, const uint64_t *tile_base, char *out, uint32_t smem_cap, int use_tma)
{
    static_assert(TILE == SR_COLS, "one thread per tile column");
    extern __shared__ __align__(16) char s_text[];
    __shared__ __align__(16) SrSm sm;
    __shared__ uint32_t s_ws[TILE / 32];
    const int tid = (int)threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int32_t c0 = (int32_t)blockIdx.x * TILE, c = c0 + tid;
    const ReadRange rr = sr_range(v, c0 >> 5);
    const int nch = (rr.n + SR_SLOTS - 1) / SR_SLOTS;
    // descriptor of "my" slot of chunk j (lanes 0..SR_WSLOTS-1 of every warp), zero past the end of the list
    auto fetch = [&](int j, int32_t &idx) -> uint4 {
        const int32_t t = j * SR_SLOTS + SR_WSLOTS * w + lane;
        idx = 0;
        if (lane < SR_WSLOTS && t < rr.n) { idx = range_at(rr, t); return __ldg(reinterpret_cast<const uint4 *>(v.desc + idx)); }
        return make_uint4(0, 0, 0, 0);          // rpos == rend: covers nothing
    };
    int32_t i_next, i_cur;
    uint4 raw_cur = fetch(0, i_cur);
    uint4 raw_next = fetch(1, i_next);

    MpFileSz stt;
    uint32_t len = 0;
    if (c < v.ncols) { len = len_in[c]; if (len) stt = st_in[c]; }
    uint32_t total;
    const uint32_t off = block_excl_scan<TILE>(len, s_ws, total);
    if (total == 0) return;
    const uint64_t base = tile_base[blockIdx.x];
    const uint32_t phase = (uint32_t)(base & 15);
    if (total + phase > smem_cap) {            // very deep tile: format straight into HBM, column by column
        if (len) sr_deep_line(v, cf, c, stt, out + base + off);
        return;
    }
    const uint32_t ends = cf.no_ends ? 0u : 1u;
    const int minq = cf.min_baseQ, out_mapq = cf.out_mapq;
    const uint32_t *qual32 = reinterpret_cast<const uint32_t *>(v.qual), *seq32 = reinterpret_cast<const uint32_t *>(v.seq4);
    const uint32_t a_q0 = (uint32_t)__cvta_generic_to_shared(&sm.q[0][0][0]), a_s0 = (uint32_t)__cvta_generic_to_shared(&sm.s[0][0][0]);
    const uint32_t a_tab = (uint32_t)__cvta_generic_to_shared(&sm.tab[0]), a_mq = (uint32_t)__cvta_generic_to_shared(&sm.mq[0][0]);

    // ---- A: warp w stages slots SR_WSLOTS*w.. of a chunk into buffer b: slot words by the lanes that hold the descriptors,
    //         then the part of each read's qualities / bases that lies over the tile, as asynchronous word copies
    auto stage = [&](int b, const uint4 &raw, int32_t idx) {
        uint32_t my_m = 0, my_qi = 0;
        if (lane < SR_WSLOTS) {
            ReadDesc d;
            d.rpos = (int32_t)raw.x; d.rend = (int32_t)raw.y; d.qoff = raw.z;
            d.qstart = (uint16_t)(raw.w & 0xffffu); d.mapq = (uint8_t)((raw.w >> 16) & 0xffu); d.fl = (uint8_t)(raw.w >> 24);
            my_m = sr_meta(d, c0, ends, my_qi);
            const int slot = SR_WSLOTS * w + lane;
            const int32_t rel_a = (int32_t)(my_m & 0xffu), row = b * SR_SLOTS + slot;
            sm.meta[b][slot] = make_int4((int32_t)my_m, row * (SR_QROW * 4) + (int32_t)((my_m >> 16) & 3u) - rel_a,
                                         row * (SR_SROW * 8) + (int32_t)((my_m >> 18) & 7u) - rel_a, idx);
            sm.mq[b][slot] = (uint8_t)(d.mapq > 93 ? 126 : d.mapq + 33);
        }
#pragma unroll
        for (int k = 0; k < SR_WSLOTS; ++k) {
            const uint32_t m = __shfl_sync(0xffffffffu, my_m, k), qi = __shfl_sync(0xffffffffu, my_qi, k);
            if (m & SR_SIMPLE) {                 // warp-uniform
                const uint32_t row = (uint32_t)(b * SR_SLOTS + SR_WSLOTS * w + k);
                const uint32_t nb = (m >> 8) & 0xffu;
                const uint32_t nwq = ((qi & 3u) + nb + 3u) >> 2, nws = ((qi & 7u) + nb + 7u) >> 3;
                const uint32_t *q32 = qual32 + (qi >> 2), *b32 = seq32 + (qi >> 3);
                const uint32_t dq = a_q0 + (row * SR_QROW + (uint32_t)lane) * 4u, ds = a_s0 + (row * SR_SROW + (uint32_t)lane) * 4u;
                if ((uint32_t)lane < nwq) cp_async4(dq, q32 + lane);
                if (nwq > 32u && lane == 0) cp_async4(dq + 128u, q32 + 32);
                if ((uint32_t)lane < nws) cp_async4(ds, b32 + lane);
            }
        }
    };
    if (nch > 0) stage(0, raw_cur, i_cur);

    // ---- my line: everything except the entries (overlaps with the copies of chunk 0)
    char *sb = s_text + phase;
    if (tid < 32) sm.tab[tid] = (uint8_t)".ACMGRSVTWYHKDBN,acmgrsvtwyhkdbn"[tid];
    const uint32_t rb = sr_ref_code(v, c);
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sb);
    uint32_t ps = 0, pq = 0, pm = 0;            // cursors of my line's strings as shared-window addresses (0: nothing to append)
    if (len) {
        const SrCur cur = sr_layout(v, cf, c, stt, sb + off);
        if (cur.ps) { ps = sbase + (uint32_t)(cur.ps - sb); pq = sbase + (uint32_t)(cur.pq - sb); pm = sbase + (uint32_t)(cur.pm - sb); }
    }
    const uint32_t a_q = a_q0 + (uint32_t)tid;

    for (int k = 0; k < nch; ++k) {
        const int b = k & 1;
        cp_async_wait_all();
        __syncthreads();                        // chunk k has landed and is visible; everybody is done with chunk k-1
        if (k + 1 < nch) {                      // copy chunk k+1 while chunk k is consumed
            stage(b ^ 1, raw_next, i_next);
            raw_next = fetch(k + 2, i_next);
        }
        // ---- B: my column, slot by slot
        if (ps) {
            const int ns = rr.n - k * SR_SLOTS < SR_SLOTS ? rr.n - k * SR_SLOTS : SR_SLOTS;
            int4 nx = sm.meta[b][0];
#pragma unroll 4
            for (int s = 0; s < ns; ++s) {
                const int4 mt = nx;
                nx = sm.meta[b][s + 1 < SR_SLOTS ? s + 1 : s];
                const uint32_t m = (uint32_t)mt.x;
                const uint32_t r = (uint32_t)tid - (m & 0xffu);
                if (r >= ((m >> 8) & 0xffu)) continue;
                const uint32_t a_mqs = a_mq + (uint32_t)(b * SR_SLOTS + s);
                if (m & SR_SIMPLE) {
                    // same arithmetic as sr_entry (plp_core.h), on shared-window addresses; both loads issued before the -Q test
                    const uint32_t nib = (uint32_t)(mt.z + tid);            // includes 2 x the row offset (even: parity is the nibble's)
                    const uint32_t q = lds8(a_q + (uint32_t)mt.y);
                    const uint32_t sbyte = lds8(a_s0 + (nib >> 1));
                    if ((int)q < minq) continue;
                    uint32_t code = (sbyte >> ((~nib & 1u) << 2)) & 0xfu;
                    if (code == rb) code = 0;
                    const uint32_t ch = lds8(a_tab + (((m >> 20) & 0x10u) | code));
                    if ((m >> 26) & (uint32_t)(r == 0)) { sts8(ps, '^'); sts8(ps + 1, lds8(a_mqs)); ps += 2; }
                    sts8(ps++, ch);
                    if ((m >> 27) & (uint32_t)(r + 1u == ((m >> 8) & 0xffu))) sts8(ps++, '$');
                    sts8(pq++, min(q + 33u, 126u));
                } else {
                    int q;
                    const int nb = sr_slow_dev(v, cf, mt.w, c, sb + (ps - sbase), &q);
                    if (nb < 0) continue;
                    ps += (uint32_t)nb;
                    sts8(pq++, (uint32_t)(q + 33 < 126 ? q + 33 : 126));
                }
                if (out_mapq) sts8(pm++, lds8(a_mqs));
            }
        }
    }
    __syncthreads();
    char *g = out + base;
    const uint32_t head = min(total, (16u - phase) & 15u);
    const uint32_t body = (total - head) & ~15u;
    const uint32_t tail = total - head - body;
    if ((uint32_t)tid < head) g[tid] = sb[tid];
    if ((uint32_t)tid < tail) g[head + body + tid] = sb[head + body + tid];
    if (body) {
        if (use_tma) { if (tid == 0) bulk_store_s2g(g + head, sb + head, body); }
        else {
            const uint4 *src = reinterpret_cast<const uint4 *>(sb + head);
            uint4 *dst = reinterpret_cast<uint4 *>(g + head);
            for (uint32_t i = (uint32_t)tid; i < body / 16; i += TILE) dst[i] = src[i];
        }
    }
}
