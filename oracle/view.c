/*
 * oracle/view.c -- TEST INFRASTRUCTURE ONLY (CPU oracle).
 * A tiny `view` (SAM text out, -h/-H/-f/-F/-o) so that the reference's
 * regression tables (test/mpileup/mpileup.reg INIT lines and `samtools view
 * ... | samtools mpileup -` pipelines) can be replayed without samtools.
 * "-b"/"-S" are accepted and ignored: the oracle reader sniffs SAM vs BAM.
 */
#include "hl.h"
#include <getopt.h>

static void put_aux(str_t *o, const uint8_t *s, const uint8_t *end)
{
    while (s + 3 <= end) {
        int ty = s[2];
        s_putc(o, '\t'); s_putc(o, s[0]); s_putc(o, s[1]); s_putc(o, ':');
        const uint8_t *v = s + 3;
        char b[64];
        switch (ty) {
        case 'A': s_puts(o, "A:"); s_putc(o, *v); s = v + 1; break;
        case 'c': s_puts(o, "i:"); s_putll(o, *(const int8_t *)v); s = v + 1; break;
        case 'C': s_puts(o, "i:"); s_putll(o, *v); s = v + 1; break;
        case 's': { int16_t x; memcpy(&x, v, 2); s_puts(o, "i:"); s_putll(o, x); s = v + 2; break; }
        case 'S': { uint16_t x; memcpy(&x, v, 2); s_puts(o, "i:"); s_putll(o, x); s = v + 2; break; }
        case 'i': { int32_t x; memcpy(&x, v, 4); s_puts(o, "i:"); s_putll(o, x); s = v + 4; break; }
        case 'I': { uint32_t x; memcpy(&x, v, 4); s_puts(o, "i:"); s_putll(o, x); s = v + 4; break; }
        case 'f': { float x; memcpy(&x, v, 4); snprintf(b, sizeof b, "f:%g", x); s_puts(o, b); s = v + 4; break; }
        case 'Z': case 'H': s_putc(o, ty); s_putc(o, ':'); s_puts(o, (const char *)v); s = v + strlen((const char *)v) + 1; break;
        case 'B': {
            int st = v[0]; uint32_t n, i; memcpy(&n, v + 1, 4);
            int esz = (st == 'c' || st == 'C') ? 1 : (st == 's' || st == 'S') ? 2 : 4;
            s_puts(o, "B:"); s_putc(o, st);
            const uint8_t *e = v + 5;
            for (i = 0; i < n; i++, e += esz) {
                s_putc(o, ',');
                if (st == 'f') { float x; memcpy(&x, e, 4); snprintf(b, sizeof b, "%g", x); s_puts(o, b); }
                else if (st == 'c') s_putll(o, *(const int8_t *)e);
                else if (st == 'C') s_putll(o, *e);
                else if (st == 's') { int16_t x; memcpy(&x, e, 2); s_putll(o, x); }
                else if (st == 'S') { uint16_t x; memcpy(&x, e, 2); s_putll(o, x); }
                else if (st == 'i') { int32_t x; memcpy(&x, e, 4); s_putll(o, x); }
                else { uint32_t x; memcpy(&x, e, 4); s_putll(o, x); }
            }
            s = e;
            break;
        }
        default: return;
        }
    }
}

int main_view(int argc, char **argv)
{
    int c, with_hdr = 0, hdr_only = 0, req = 0, excl = 0;
    const char *out_fn = NULL;
    optind = 1;
    while ((c = getopt(argc, argv, "hHbSCf:F:o:")) >= 0) {
        switch (c) {
        case 'h': with_hdr = 1; break;
        case 'H': hdr_only = 1; break;
        case 'f': req = parse_flag(optarg); break;
        case 'F': excl = parse_flag(optarg); break;
        case 'o': out_fn = optarg; break;
        default: break;
        }
    }
    if (optind >= argc) return 1;
    reader_t *rd = reader_open(argv[optind], NULL);
    if (!rd) return 1;
    hdr_t *h = reader_hdr(rd);
    FILE *fp = out_fn ? fopen(out_fn, "w") : stdout;
    if (with_hdr || hdr_only) {
        if (h->text && *h->text) fputs(h->text, fp);
        else { int i; for (i = 0; i < h->n_ref; i++) fprintf(fp, "@SQ\tSN:%s\tLN:%lld\n", h->name[i], (long long)h->len[i]); }
    }
    rec_t b; rec_init(&b);
    str_t o = {0, 0, NULL};
    while (!hdr_only && reader_next(rd, &b) >= 0) {
        int i;
        if ((b.flag & req) != req || (b.flag & excl)) continue;
        o.l = 0;
        s_puts(&o, b.qname); s_putc(&o, '\t'); s_putll(&o, b.flag); s_putc(&o, '\t');
        s_puts(&o, b.tid >= 0 ? h->name[b.tid] : "*"); s_putc(&o, '\t'); s_putll(&o, b.pos + 1); s_putc(&o, '\t');
        s_putll(&o, b.mapq); s_putc(&o, '\t');
        if (b.n_cigar == 0) s_putc(&o, '*');
        for (i = 0; i < (int)b.n_cigar; i++) { s_putll(&o, cln(b.cigar[i])); s_putc(&o, "MIDNSHP=XB"[cop(b.cigar[i])]); }
        s_putc(&o, '\t');
        if (b.mtid < 0) s_putc(&o, '*'); else if (b.mtid == b.tid) s_putc(&o, '='); else s_puts(&o, h->name[b.mtid]);
        s_putc(&o, '\t'); s_putll(&o, b.mpos + 1); s_putc(&o, '\t'); s_putll(&o, b.isize); s_putc(&o, '\t');
        if (b.l_qseq == 0) s_putc(&o, '*');
        for (i = 0; i < b.l_qseq; i++) s_putc(&o, nt16_str[seqi(b.seq, i)]);
        s_putc(&o, '\t');
        if (b.l_qseq == 0 || b.qual[0] == 0xff) s_putc(&o, '*');
        else for (i = 0; i < b.l_qseq; i++) s_putc(&o, b.qual[i] + 33);
        put_aux(&o, b.aux, b.aux + b.l_aux);
        s_putc(&o, '\n');
        fwrite(o.s, 1, o.l, fp);
    }
    if (out_fn) fclose(fp);
    reader_close(rd);
    return 0;
}
