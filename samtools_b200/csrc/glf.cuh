// glf.cuh -- genotype likelihoods per column on the device.
//
// Replaces bcf_call_glfgen (bam2bcf.c:65-123: per-read filter, quality caps,
// packing q<<5|strand<<4|base, qsum) and htslib errmod_cal / errmod_init
// (errmod.c; tables fk, beta, lhet; ks_shuffle over hts_drand48 when a column
// holds more than 255 usable bases; ascending sort consumed from the top).
// Semantics: SURVEY.md section 8a rows a16/a17.  One warp per (column, file).
// No reference test pins these numbers ("parity unpinned"); parity is against
// the CPU oracle's restatement.
#pragma once
#include <math.h>

#define GL_CAP 4096   // usable bases per (column,file) kept in shared memory

__device__ __forceinline__ uint64_t lcg_jump(uint64_t s, uint64_t k)
{
    // s_{n+1} = A*s_n + C (mod 2^48); advance k steps by repeated squaring
    uint64_t A = 0x5DEECE66DULL, C = 0xBULL, accA = 1, accC = 0;
    const uint64_t M = 0xffffffffffffULL;
    while (k) {
        if (k & 1) { accC = (accC * A + C) & M; accA = (accA * A) & M; }
        C = ((A + 1) * C) & M; A = (A * A) & M;
        k >>= 1;
    }
    return (accA * s + accC) & M;
}

// pass 1: usable bases per (column,file) -> random draws ks_shuffle will consume
__device__ __forceinline__ bool gl_code(const View &v, const ReadDesc &d, int32_t c, int min_baseQ, int rb4, uint16_t &code, int &qv, int &bv)
{
    Ent e; resolve(v, d, c, e);
    if (e.is_del || e.is_refskip) return false;
    int q = e.qpos < d.l_qseq ? (int)v.qual[d.qoff + (uint64_t)e.qpos] : 0;
    if (q < min_baseQ) return false;
    int mapQ = d.mapq < 255 ? d.mapq : 20;
    if (q > 99) q = 99;
    if (mapQ > 60) mapQ = 60;
    if (q > mapQ) q = mapQ;
    if (q > 63) q = 63;
    if (q < 4) q = 4;
    int b;
    if (e.qpos < d.l_qseq) { b = base4(v.seq4, d.qoff, e.qpos); b = nt16_int_of(b ? b : rb4); }
    else b = 4;
    code = (uint16_t)(q << 5 | ((d.fl & RD_REV) ? 1 : 0) << 4 | b);
    qv = q; bv = b;
    return true;
}

__global__ void k_gl_count(View v, int min_baseQ, uint32_t *draws, int32_t *nplp_any)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)v.ncols * v.n_files) return;
    const int32_t c = (int32_t)(idx / v.n_files); const int f = (int)(idx % v.n_files);
    const int g = c >> 5;
    const ReadRange rr = read_range(v, f, g);
    int rb4 = 15;
    { const char rc = ref_char(v, c); rb4 = (v.ref && (int64_t)c < v.ref_len_rel) ? nt16_of((unsigned char)rc) : 15; }
    uint32_t n = 0, np = 0;
    for (int32_t t_ = 0; t_ < rr.n; ++t_) {
        const ReadDesc d = v.desc[range_at(rr, t_)];
        if (c < d.rpos || c >= d.rend) continue;
        ++np;
        uint16_t code; int q, b;
        if (gl_code(v, d, c, min_baseQ, rb4, code, q, b)) ++n;
    }
    draws[idx] = n > 255 ? n - 1 : 0;
    if (np) atomicOr(&nplp_any[c], 1);
}

__global__ void __launch_bounds__(128) k_gl(View v, int min_baseQ, const uint64_t *draw_off, uint64_t rng_base_draws,
                                             const double *fk, const double *beta, const double *lhet,
                                             int32_t *out_n, float *out_qp /* 29 floats */, uint32_t *overflow)
{
    __shared__ uint16_t s_b[4][GL_CAP];
    const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
    const int64_t idx = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (idx >= (int64_t)v.ncols * v.n_files) return;
    const int32_t c = (int32_t)(idx / v.n_files); const int f = (int)(idx % v.n_files);
    const int g = c >> 5;
    const ReadRange rr = read_range(v, f, g);
    uint16_t *bs = s_b[wl];
    int rb4;
    { const char rc = ref_char(v, c); rb4 = (v.ref && (int64_t)c < v.ref_len_rel) ? nt16_of((unsigned char)rc) : 15; }
    int n = 0, nplp = 0;
    float qsum[4] = {0, 0, 0, 0};   // exact small integers, order-independent below 2^24
    for (int32_t base = 0; base < rr.n; base += 32) {
        const int32_t t_ = base + lane;
        bool ok = false, cov = false; uint16_t code = 0; int q = 0, b = 4;
        if (t_ < rr.n) {
            const ReadDesc d = v.desc[range_at(rr, t_)];
            if (c >= d.rpos && c < d.rend) { cov = true; ok = gl_code(v, d, c, min_baseQ, rb4, code, q, b); }
        }
        const unsigned mk = __ballot_sync(0xffffffffu, ok);
        nplp += __popc(__ballot_sync(0xffffffffu, cov));
        if (ok) { const int slot = n + __popc(mk & ((1u << lane) - 1)); if (slot < GL_CAP) bs[slot] = code; }
        n += __popc(mk);
#pragma unroll
        for (int bb = 0; bb < 4; ++bb) {
            float x = (ok && b == bb) ? (float)q : 0.f;
            for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            qsum[bb] += x;
        }
    }
    __syncwarp();
    float *o = out_qp + idx * 29;
    if (nplp == 0) { if (lane == 0) out_n[idx] = -1; if (lane < 29) o[lane] = 0.f; return; }
    if (n > GL_CAP) { if (lane == 0) atomicOr(overflow, 1u); n = GL_CAP; }
    const int n_used = n;
    if (n > 255) {   // ks_shuffle with the process-wide drand48 stream, then keep the first 255
        if (lane == 0) {
            uint64_t s = lcg_jump(0x330EULL, rng_base_draws + draw_off[idx]);
            for (int i = n; i > 1; --i) {
                s = (s * 0x5DEECE66DULL + 0xBULL) & 0xffffffffffffULL;
                const int j = (int)(((double)s / 281474976710656.0) * i);
                const uint16_t t = bs[j]; bs[j] = bs[i - 1]; bs[i - 1] = t;
            }
        }
        n = 255;
        __syncwarp();
    }
    // ascending sort of n <= 255 codes: pad to 256 and bitonic-sort in shared memory
    for (int i = n + lane; i < 256; i += 32) bs[i] = 0xffff;
    __syncwarp();
    for (int k = 2; k <= 256; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = lane; t < 256; t += 32) {
                const int p = t ^ j;
                if (p > t) {
                    const uint16_t a = bs[t], b = bs[p];
                    const bool up = (t & k) == 0;
                    if ((a > b) == up) { bs[t] = b; bs[p] = a; }
                }
            }
            __syncwarp();
        }
    if (lane == 0) {
        double fsum[5] = {0, 0, 0, 0, 0}, bsum[5] = {0, 0, 0, 0, 0};
        int cc[5] = {0, 0, 0, 0, 0}, w[32];
        for (int i = 0; i < 32; ++i) w[i] = 0;
        for (int j = n - 1; j >= 0; --j) {
            const uint16_t b = bs[j];
            int qual = (b >> 5) < 4 ? 4 : (b >> 5);
            if (qual > 63) qual = 63;
            const int basestrand = b & 0x1f, base = b & 0xf;
            const double fkw = fk[w[basestrand]];
            fsum[base] += fkw;
            bsum[base] += fkw * beta[qual << 16 | n << 8 | cc[base]];
            ++cc[base]; ++w[basestrand];
        }
        float q[25];
        for (int i = 0; i < 25; ++i) q[i] = 0.f;
        const int m = 5;
        for (int j = 0; j < m; ++j) {
            float tmp1 = 0.f; int tmp2 = 0;
            for (int k = 0; k < m; ++k) { if (k == j) continue; tmp1 = (float)((double)tmp1 + bsum[k]); tmp2 += cc[k]; }
            if (tmp2) q[j * m + j] = tmp1;
            for (int k = j + 1; k < m; ++k) {
                const int cjk = cc[j] + cc[k];
                tmp1 = 0.f; tmp2 = 0;
                for (int i = 0; i < m; ++i) { if (i == j || i == k) continue; tmp1 = (float)((double)tmp1 + bsum[i]); tmp2 += cc[i]; }
                const float val = tmp2 ? (float)(-4.343 * lhet[cjk << 8 | cc[k]] + (double)tmp1) : (float)(-4.343 * lhet[cjk << 8 | cc[k]]);
                q[j * m + k] = q[k * m + j] = val;
            }
            for (int k = 0; k < m; ++k) if (q[j * m + k] < 0.0f) q[j * m + k] = 0.0f;
        }
        out_n[idx] = n_used;
        for (int i = 0; i < 4; ++i) o[i] = qsum[i];
        for (int i = 0; i < 25; ++i) o[4 + i] = q[i];
    }
}

// errmod_init(depcorr = 1 - 0.83) tables, built once with the host libm (long double exp/log as upstream)
static void errmod_host_tables(std::vector<double> &fk, std::vector<double> &beta, std::vector<double> &lhet)
{
    const double depcorr = 1. - 0.83, eta = 0.03;
    fk.assign(256, 0.); beta.assign((size_t)256 * 256 * 64, 0.); lhet.assign(256 * 256, 0.);
    fk[0] = 1.0;
    for (int n = 1; n < 256; ++n) fk[n] = pow(1. - depcorr, n) * (1.0 - eta) + eta;
    std::vector<double> lC(256 * 256, 0.);
    for (int n = 1; n != 256; ++n) {
        const double lgn = lgamma(n + 1);
        for (int k = 1; k <= n; ++k) lC[n << 8 | k] = lgn - lgamma(k + 1) - lgamma(n - k + 1);
    }
    for (int q = 1; q != 64; ++q) {
        const double e = pow(10.0, -q / 10.0), le = log(e), le1 = log(1.0 - e);
        for (int n = 1; n <= 255; ++n) {
            double *b = beta.data() + (q << 16 | n << 8);
            double sum, sum1;
            sum1 = sum = 0.0;
            for (int k = n; k >= 0; --k, sum1 = sum) {
                sum = sum1 + expl(lC[n << 8 | k] + k * le + (n - k) * le1);
                b[k] = -10. / M_LN10 * logl(sum1 / sum);
            }
        }
    }
    for (int n = 0; n < 256; ++n)
        for (int k = 0; k < 256; ++k) lhet[n << 8 | k] = lC[n << 8 | k] - M_LN2 * n;
}

extern "C" int b200_glf(b200_engine_t *e, int32_t min_baseQ, int64_t *n_cols, int64_t *col_pos, int32_t *n_bases,
                        float *qsum, float *p25, size_t cap_cols)
{
    if (!e || !e->staged) { if (e) snprintf(e->err, sizeof e->err, "no staged batch"); return -1; }
    CK(cudaSetDevice(e->device));
    if (!e->d_beta) {
        std::vector<double> fk, beta, lhet;
        errmod_host_tables(fk, beta, lhet);
        CK(cudaMalloc((void **)&e->d_fk, fk.size() * 8)); CK(cudaMalloc((void **)&e->d_beta, beta.size() * 8)); CK(cudaMalloc((void **)&e->d_lhet, lhet.size() * 8));
        CK(cudaMemcpy(e->d_fk, fk.data(), fk.size() * 8, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(e->d_beta, beta.data(), beta.size() * 8, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(e->d_lhet, lhet.data(), lhet.size() * 8, cudaMemcpyHostToDevice));
    }
    View v; fill_view(e, v, nullptr, nullptr, 0, 0, 0);
    *n_cols = 0;
    const int64_t tot = (int64_t)v.ncols * v.n_files;
    if (tot == 0) return 0;
    ENSURE(col_n, (size_t)tot + 1); ENSURE(col_off, (size_t)tot + 2); ENSURE(gl_n, (size_t)tot + 1);
    ENSURE(gl_out, (size_t)tot * 29 + 1); ENSURE(gl_flag, (size_t)v.ncols + 2);
    CK(cudaMemsetAsync(e->gl_flag, 0, ((size_t)v.ncols + 2) * 4, e->stream));
    CK(cudaEventRecord(e->ev0, e->stream));
    k_gl_count<<<nblk(tot, 256), 256, 0, e->stream>>>(v, min_baseQ, e->col_n, (int32_t *)e->gl_flag); e->launches++;
    const int nb = nblk(tot, 256);
    ENSURE(status, (size_t)nb + 1);
    CK(cudaMemsetAsync(e->status, 0, ((size_t)nb + 1) * 8, e->stream));
    CK(cudaMemsetAsync(e->d_misc, 0, 8, e->stream));
    k_scan_u32_to_u64<<<nb, 256, 0, e->stream>>>(e->col_n, e->col_off, (int32_t)tot, e->status, (uint32_t *)e->d_misc); e->launches++;
    uint32_t *ovf = (uint32_t *)(e->gl_flag + v.ncols);
    k_gl<<<nblk(tot * 32, 128), 128, 0, e->stream>>>(v, min_baseQ, e->col_off, e->gl_rng_draws, e->d_fk, e->d_beta, e->d_lhet,
                                                    e->gl_n, e->gl_out, ovf); e->launches++;
    CK(cudaEventRecord(e->ev1, e->stream));
    uint64_t total_draws = 0;
    CK(cudaMemcpyAsync(&total_draws, e->col_off + tot, 8, cudaMemcpyDeviceToHost, e->stream));
    if (!col_pos) {   // device-only: the likelihoods stay in HBM
        uint32_t h_ovf = 0;
        CK(cudaMemcpyAsync(&h_ovf, ovf, 4, cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
        CK(cudaGetLastError());
        float ms0 = 0; cudaEventElapsedTime(&ms0, e->ev0, e->ev1); e->last_kernel_ms = ms0;
        e->gl_rng_draws += total_draws;
        if (h_ovf) { snprintf(e->err, sizeof e->err, "GL: a column holds more than %d usable bases", GL_CAP); return -1; }
        *n_cols = v.ncols;
        return 0;
    }
    std::vector<uint32_t> flag((size_t)v.ncols + 1);
    std::vector<int32_t> hn((size_t)tot);
    std::vector<float> ho((size_t)tot * 29);
    CK(cudaMemcpyAsync(flag.data(), e->gl_flag, ((size_t)v.ncols + 1) * 4, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaMemcpyAsync(hn.data(), e->gl_n, (size_t)tot * 4, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaMemcpyAsync(ho.data(), e->gl_out, (size_t)tot * 29 * 4, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaGetLastError());
    float ms = 0; cudaEventElapsedTime(&ms, e->ev0, e->ev1); e->last_kernel_ms = ms;
    e->gl_rng_draws += total_draws;
    if (flag[(size_t)v.ncols]) { snprintf(e->err, sizeof e->err, "GL: a column holds more than %d usable bases", GL_CAP); return -1; }
    int64_t k = 0;
    for (int32_t c = 0; c < v.ncols; ++c) {
        if (!flag[(size_t)c]) continue;
        if ((size_t)k >= cap_cols) { snprintf(e->err, sizeof e->err, "GL output capacity too small"); return -2; }
        col_pos[k] = e->win_base + c;
        for (int f = 0; f < v.n_files; ++f) {
            const size_t s = (size_t)c * v.n_files + f, d = (size_t)k * v.n_files + f;
            n_bases[d] = hn[s];
            memcpy(qsum + d * 4, ho.data() + s * 29, 16);
            memcpy(p25 + d * 25, ho.data() + s * 29 + 4, 100);
        }
        ++k;
    }
    *n_cols = k;
    return 0;
}
