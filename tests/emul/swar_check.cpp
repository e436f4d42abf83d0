// swar_check.cpp -- TEST HARNESS (not shipped): exhaustive check of the SIMD-within-a-register entry formatter the device entry
// pass uses (plp_core.h::ent_group8_swar) against the per-base definition (ent_plain): every quality byte 0..255 x every
// 4-bit base code x both strands x -Q 0..127 x reference code equal / different / absent, in every one of the eight positions
// of a group.  build: tests/emul/build.sh
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include "../../samtools_b200/csrc/plp_core.h"
using namespace plp;

int main()
{
    const uint8_t *tab = (const uint8_t *)".ACMGRSVTWYHKDBN,acmgrsvtwyhkdbn";
    long n = 0, bad = 0;
    uint32_t rng = 12345;
    auto rnd = [&]() { rng = rng * 1664525u + 1013904223u; return rng >> 8; };
    for (int minq = 0; minq <= 127; ++minq) {
        const uint32_t minq4 = (uint32_t)minq * 0x01010101u;
        for (uint32_t rev = 0; rev < 2; ++rev) {
            const EntTab t = ent_tab(rev);
            for (int q = 0; q < 256; ++q)
                for (int code = 0; code < 16; ++code)
                    for (int refmode = 0; refmode < 3; ++refmode) {          // 0: no reference, 1: reference equals the base, 2: differs
                        const int k = (q + code + minq) & 7;                 // position of the probed base inside the group
                        uint32_t qb[8], cb[8], rb[8];
                        for (int j = 0; j < 8; ++j) { qb[j] = rnd() & 0xff; cb[j] = rnd() & 15; rb[j] = rnd() & 15; }
                        qb[k] = (uint32_t)q; cb[k] = (uint32_t)code; rb[k] = refmode == 1 ? (uint32_t)code : ((uint32_t)code + 1 + (rnd() % 15)) & 15;
                        uint32_t qx = 0, qy = 0, s4 = 0, r8 = 0;
                        for (int j = 0; j < 8; ++j) {
                            if (j < 4) qx |= qb[j] << (8 * j); else qy |= qb[j] << (8 * (j - 4));
                            s4 |= cb[j] << (8 * (j >> 1) + ((j & 1) ? 0 : 4));
                            r8 |= rb[j] << (4 * j);
                        }
                        uint32_t w[4];
                        const uint32_t fm = ent_group8_swar(qx, qy, s4, refmode != 0, r8, t, minq4, w);
                        for (int j = 0; j < 8; ++j) {
                            const uint32_t want = ent_plain(qb[j], cb[j], refmode ? rb[j] : 0x10u, rev, minq, 0, tab);
                            const uint32_t got = (w[j >> 1] >> (16 * (j & 1))) & 0xffffu;
                            ++n;
                            if (got != want || (((fm >> j) & 1u) != (want == 0 ? 1u : 0u))) {
                                if (++bad <= 5) fprintf(stderr, "MISMATCH q=%u code=%u ref=%u rev=%u minq=%d pos=%d: got %04x want %04x fail bit %u\n",
                                                        qb[j], cb[j], refmode ? rb[j] : 16u, rev, minq, j, got, want, (fm >> j) & 1u);
                            }
                        }
                    }
        }
    }
    printf("checked %ld entries, %ld mismatches\n", n, bad);
    return bad ? 1 : 0;
}
