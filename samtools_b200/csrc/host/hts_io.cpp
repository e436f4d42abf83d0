// hts_io.cpp -- see hts_io.hpp.
#include "hts_io.hpp"
#include <zlib.h>
#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cerrno>
#include <cstdlib>
#include <cstring>
#include <strings.h>
#include <sys/stat.h>

namespace b200 {

static const char kOps[] = "MIDNSHP=XB";

static uint8_t nt16(unsigned char c)
{
    switch (c | 0x20) {
    case 'a': return 1; case 'c': return 2; case 'm': return 3; case 'g': return 4; case 'r': return 5; case 's': return 6;
    case 'v': return 7; case 't': return 8; case 'w': return 9; case 'y': return 10; case 'h': return 11; case 'k': return 12;
    case 'd': return 13; case 'b': return 14; case 'n': return 15;
    default: break;
    }
    if (c == '=') return 0;
    if (c >= '0' && c <= '3') return (uint8_t)(1 << (c - '0'));
    return 15;
}

int64_t Record::rlen() const
{
    int64_t l = 0;
    for (uint32_t c : cigar) { int op = c & 0xf; if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) l += c >> 4; }
    return l;
}
int64_t Record::endpos() const
{
    int64_t rl = 1;
    if (!(flag & F_UNMAP) && !cigar.empty()) { rl = rlen(); if (rl == 0) rl = 1; }
    return pos + rl;
}
static int aux_size(int t)
{
    switch (t) { case 'A': case 'c': case 'C': return 1; case 's': case 'S': return 2; case 'i': case 'I': case 'f': return 4; case 'd': return 8; default: return 0; }
}
static const uint8_t *aux_skip(const uint8_t *s, const uint8_t *end)
{
    int t = *s++;
    if (int sz = aux_size(t)) return s + sz <= end ? s + sz : nullptr;
    if (t == 'Z' || t == 'H') { while (s < end && *s) ++s; return s < end ? s + 1 : nullptr; }
    if (t == 'B') {
        if (s + 5 > end) return nullptr;
        int esz = aux_size(*s); uint32_t n; memcpy(&n, s + 1, 4);
        s += 5 + (size_t)esz * n;
        return (esz && s <= end) ? s : nullptr;
    }
    return nullptr;
}
const uint8_t *Record::aux_get(const char tag[2]) const
{
    const uint8_t *s = aux.data(), *end = s + aux.size();
    while (s && s + 3 <= end) {
        if (s[0] == (uint8_t)tag[0] && s[1] == (uint8_t)tag[1]) return s + 2;
        s = aux_skip(s + 2, end);
    }
    return nullptr;
}

int Header::name2tid(const std::string &n) const
{
    for (size_t i = 0; i < names.size(); ++i) if (names[i] == n) return (int)i;
    return -1;
}

bool parse_region(const Header &h, const std::string &reg, int &tid, int64_t &beg, int64_t &end)
{
    beg = 0; end = POS_MAX;
    int t = h.name2tid(reg);
    if (t >= 0) { tid = t; return true; }
    size_t colon = reg.rfind(':');
    if (colon == std::string::npos) return false;
    t = h.name2tid(reg.substr(0, colon));
    if (t < 0) return false;
    std::string a, b; bool dash = false;
    for (size_t i = colon + 1; i < reg.size(); ++i) {
        char c = reg[i];
        if (c == ',') continue;
        if (c == '-' && !dash) { dash = true; continue; }
        (dash ? b : a).push_back(c);
    }
    long long lb = a.empty() ? 0 : atoll(a.c_str());
    beg = lb > 0 ? lb - 1 : 0;
    end = (dash && !b.empty()) ? atoll(b.c_str()) : POS_MAX;
    if (beg >= end) return false;
    tid = t;
    return true;
}

int parse_flag(const std::string &s)
{
    char *e; long v = strtol(s.c_str(), &e, 0);
    if (e != s.c_str() && *e == 0) return v < 0 ? -1 : (int)v;
    static const struct { const char *n; int f; } names[] = {
        {"PAIRED", 1}, {"PROPER_PAIR", 2}, {"UNMAP", 4}, {"MUNMAP", 8}, {"REVERSE", 16}, {"MREVERSE", 32},
        {"READ1", 64}, {"READ2", 128}, {"SECONDARY", 256}, {"QCFAIL", 512}, {"DUP", 1024}, {"SUPPLEMENTARY", 2048} };
    int flag = 0; size_t p = 0;
    while (p < s.size()) {
        size_t q = s.find(',', p); if (q == std::string::npos) q = s.size();
        std::string tok = s.substr(p, q - p); bool ok = false;
        for (auto &n : names) if (strcasecmp(tok.c_str(), n.n) == 0) { flag |= n.f; ok = true; break; }
        if (!ok) return -1;
        p = q + 1;
    }
    return flag;
}

uint32_t qname_hash_bit(const std::string &q)
{
    const char *s = q.c_str();
    uint32_t h = (uint32_t)*s;
    if (h) for (++s; *s; ++s) h = (h << 5) - h + (uint32_t)*s;
    h += ~(h << 15); h ^= (h >> 10); h += (h << 3); h ^= (h >> 6); h += ~(h << 11); h ^= (h >> 16);
    return h & 1;
}

// ------------------------------------------------------------------ reader
AlnReader::~AlnReader() { if (gz_) gzclose((gzFile)gz_); }

bool AlnReader::getline(std::string &s)
{
    char buf[1 << 16];
    s.clear();
    bool got = false;
    while (gzgets((gzFile)gz_, buf, sizeof buf)) {
        got = true;
        size_t n = strlen(buf);
        if (n && buf[n - 1] == '\n') { s.append(buf, n - 1); if (!s.empty() && s.back() == '\r') s.pop_back(); return true; }
        s.append(buf, n);
    }
    return got;
}

std::unique_ptr<AlnReader> AlnReader::open(const std::string &path, const std::string &fai)
{
    gzFile fp = path == "-" ? gzdopen(0, "rb") : gzopen(path.c_str(), "rb");
    if (!fp) return nullptr;
    gzbuffer(fp, 1 << 18);
    std::unique_ptr<AlnReader> rd(new AlnReader());
    rd->gz_ = fp;
    int c0 = gzgetc(fp);
    if (c0 < 0) return rd;
    gzungetc(c0, fp);
    auto add_ref = [&](const std::string &n, int64_t l) { rd->hdr_.names.push_back(n); rd->hdr_.lens.push_back(l); };
    bool bam = false;
    if (c0 == 'B') {   // "BAM\1" magic, or a SAM record whose name starts with B: look at four bytes
        char magic[4]; int got = gzread(fp, magic, 4);
        if (got == 4 && memcmp(magic, "BAM\1", 4) == 0) bam = true;
        else for (int j = got - 1; j >= 0; --j) gzungetc((unsigned char)magic[j], fp);
    }
    if (bam) {
        int32_t l_text, n_ref;
        if (gzread(fp, &l_text, 4) != 4) return nullptr;
        rd->is_bam_ = true;
        rd->hdr_.text.resize((size_t)l_text);
        if (l_text && gzread(fp, &rd->hdr_.text[0], (unsigned)l_text) != l_text) return nullptr;
        while (!rd->hdr_.text.empty() && rd->hdr_.text.back() == '\0') rd->hdr_.text.pop_back();
        if (gzread(fp, &n_ref, 4) != 4) return nullptr;
        for (int i = 0; i < n_ref; ++i) {
            int32_t ln, lr;
            if (gzread(fp, &ln, 4) != 4) return nullptr;
            std::string nm((size_t)ln, '\0');
            if (gzread(fp, &nm[0], (unsigned)ln) != ln || gzread(fp, &lr, 4) != 4) return nullptr;
            nm.resize(strlen(nm.c_str()));
            add_ref(nm, lr);
        }
        return rd;
    }
    std::string ln;
    while (rd->getline(ln)) {
        if (!ln.empty() && ln[0] == '@') {
            rd->hdr_.text += ln; rd->hdr_.text += '\n';
            if (ln.compare(0, 3, "@SQ") == 0) {
                std::string sn; int64_t len = 0; size_t p = 0;
                while (p < ln.size()) {
                    size_t q = ln.find('\t', p); if (q == std::string::npos) q = ln.size();
                    if (ln.compare(p, 3, "SN:") == 0) sn = ln.substr(p + 3, q - p - 3);
                    else if (ln.compare(p, 3, "LN:") == 0) len = atoll(ln.c_str() + p + 3);
                    p = q + 1;
                }
                if (!sn.empty()) add_ref(sn, len);
            }
        } else { rd->pending_ = ln; rd->have_pending_ = true; break; }
    }
    if (rd->hdr_.names.empty() && !fai.empty()) {
        if (FILE *f = fopen(fai.c_str(), "r")) {
            char buf[4096], nm[1024]; long long l;
            while (fgets(buf, sizeof buf, f)) if (sscanf(buf, "%1023s %lld", nm, &l) == 2) add_ref(nm, l);
            fclose(f);
        }
    }
    return rd;
}

bool AlnReader::set_region(const std::string &reg, int &tid, int64_t &beg, int64_t &end)
{
    if (!parse_region(hdr_, reg, rtid_, rbeg_, rend_)) return false;
    has_reg_ = true;
    tid = rtid_; beg = rbeg_; end = rend_;
    return true;
}

static void aux_put(std::vector<uint8_t> &a, const char *tag, char type, const void *d, size_t n)
{
    a.push_back((uint8_t)tag[0]); a.push_back((uint8_t)tag[1]); a.push_back((uint8_t)type);
    const uint8_t *p = (const uint8_t *)d; a.insert(a.end(), p, p + n);
}
static void aux_put_int(std::vector<uint8_t> &a, const char *tag, long long v)
{
    if (v < 0) {
        if (v >= -128) { int8_t x = (int8_t)v; aux_put(a, tag, 'c', &x, 1); }
        else if (v >= -32768) { int16_t x = (int16_t)v; aux_put(a, tag, 's', &x, 2); }
        else { int32_t x = (int32_t)v; aux_put(a, tag, 'i', &x, 4); }
    } else {
        if (v < 256) { uint8_t x = (uint8_t)v; aux_put(a, tag, 'C', &x, 1); }
        else if (v < 65536) { uint16_t x = (uint16_t)v; aux_put(a, tag, 'S', &x, 2); }
        else { uint32_t x = (uint32_t)v; aux_put(a, tag, 'I', &x, 4); }
    }
}

int AlnReader::parse_sam(char *line, Record &r)
{
    char *f[11]; int nf = 0; char *p = line;
    while (nf < 11) {
        f[nf++] = p;
        char *t = strchr(p, '\t');
        if (!t) { p = nullptr; break; }
        *t = 0; p = t + 1;
    }
    if (nf < 11) return -2;
    r = Record();
    r.qname = f[0];
    r.flag = (uint16_t)strtol(f[1], nullptr, 10);
    r.tid = strcmp(f[2], "*") ? hdr_.name2tid(f[2]) : -1;
    r.pos = atoll(f[3]) - 1;
    r.mapq = (uint8_t)strtol(f[4], nullptr, 10);
    if (strcmp(f[5], "*")) {
        const char *c = f[5];
        while (*c) {
            char *e; unsigned long l = strtoul(c, &e, 10);
            const char *o = *e ? strchr(kOps, *e) : nullptr;
            if (!o || o - kOps > 8) return -2;                            // 'B' (and anything else) is not an alignment op
            r.cigar.push_back((uint32_t)l << 4 | (uint32_t)(o - kOps));
            c = e + 1;
        }
    }
    if (!strcmp(f[6], "=")) r.mtid = r.tid; else r.mtid = strcmp(f[6], "*") ? hdr_.name2tid(f[6]) : -1;
    r.mpos = atoll(f[7]) - 1;
    r.isize = atoll(f[8]);
    if (strcmp(f[9], "*")) {
        size_t l = strlen(f[9]);
        r.l_qseq = (int32_t)l;
        r.seq4.assign((l + 1) / 2, 0);
        for (size_t i = 0; i < l; ++i) r.seq4[i >> 1] |= (uint8_t)(nt16((unsigned char)f[9][i]) << ((~i & 1) << 2));
        r.qual.resize(l);
        if (!strcmp(f[10], "*")) std::fill(r.qual.begin(), r.qual.end(), 0xff);
        else for (size_t i = 0; i < l; ++i) r.qual[i] = (uint8_t)(f[10][i] - 33);
    }
    while (p && *p) {
        char *t = strchr(p, '\t');
        if (t) *t = 0;
        size_t L = strlen(p);
        if (L >= 5 && p[2] == ':' && p[4] == ':') {
            char type = p[3]; const char *v = p + 5;
            if (type == 'A') aux_put(r.aux, p, 'A', v, 1);
            else if (type == 'i') aux_put_int(r.aux, p, atoll(v));
            else if (type == 'f') { float x = strtof(v, nullptr); aux_put(r.aux, p, 'f', &x, 4); }
            else if (type == 'Z' || type == 'H') aux_put(r.aux, p, type, v, strlen(v) + 1);
            else if (type == 'B') {
                char st = v[0]; int esz = aux_size(st); uint32_t n = 0;
                for (const char *q = v + 1; *q; ++q) if (*q == ',') ++n;
                std::vector<uint8_t> buf(5 + (size_t)esz * n);
                buf[0] = (uint8_t)st; memcpy(&buf[1], &n, 4);
                uint8_t *o = buf.data() + 5; const char *q = v + 1;
                while (*q == ',') {
                    char *e; ++q;
                    if (st == 'f') { float x = strtof(q, &e); memcpy(o, &x, 4); }
                    else { long long x = strtoll(q, &e, 10); memcpy(o, &x, (size_t)esz); }
                    o += esz; q = e;
                }
                aux_put(r.aux, p, 'B', buf.data(), buf.size());
            }
        }
        p = t ? t + 1 : nullptr;
    }
    return 0;
}

int AlnReader::read_bam(Record &r)
{
    int32_t bs;
    int n = gzread((gzFile)gz_, &bs, 4);
    if (n == 0) return -1;
    if (n != 4 || bs < 32) return -2;
    std::vector<uint8_t> d((size_t)bs);
    if (gzread((gzFile)gz_, d.data(), (unsigned)bs) != bs) return -2;
    r = Record();
    int32_t i32; uint16_t u16;
    memcpy(&r.tid, &d[0], 4);
    memcpy(&i32, &d[4], 4); r.pos = i32;
    int l_name = d[8]; r.mapq = d[9];
    memcpy(&u16, &d[12], 2); uint32_t ncig = u16;
    memcpy(&r.flag, &d[14], 2);
    memcpy(&r.l_qseq, &d[16], 4);
    memcpy(&r.mtid, &d[20], 4);
    memcpy(&i32, &d[24], 4); r.mpos = i32;
    memcpy(&i32, &d[28], 4); r.isize = i32;
    // the variable part must fit the block (a truncated / corrupt record must not be read past its end)
    if (r.l_qseq < 0 || l_name < 1 ||
        32ull + (uint64_t)l_name + 4ull * ncig + ((uint64_t)r.l_qseq + 1) / 2 + (uint64_t)r.l_qseq > (uint64_t)bs) return -2;
    const uint8_t *p = d.data() + 32;
    r.qname.assign((const char *)p, strnlen((const char *)p, (size_t)l_name)); p += l_name;
    r.cigar.resize(ncig); if (ncig) memcpy(r.cigar.data(), p, 4 * (size_t)ncig); p += 4 * (size_t)ncig;
    r.seq4.assign(p, p + (r.l_qseq + 1) / 2); p += (r.l_qseq + 1) / 2;
    r.qual.assign(p, p + r.l_qseq); p += r.l_qseq;
    r.aux.assign(p, (const uint8_t *)d.data() + bs);
    for (uint32_t c : r.cigar) if ((c & 0xf) > 8) return -2;          // only MIDNSHP=X are defined for alignments
    // long CIGARs (> 65535 ops) live in the CG:B,I tag behind a <l_qseq>S<rlen>N placeholder (SAMv1 4.2.2; htslib bam_tag2cigar)
    if (r.cigar.size() == 2 && (r.cigar[0] & 0xf) == 4 && (int32_t)(r.cigar[0] >> 4) == r.l_qseq && (r.cigar[1] & 0xf) == 3) {
        const uint8_t *cg = r.aux_get("CG");
        if (cg && cg[0] == 'B' && cg[1] == 'I') {
            uint32_t n; memcpy(&n, cg + 2, 4);
            const size_t off = (size_t)(cg - r.aux.data());
            if (n > 0 && off + 6 + 4ull * n <= r.aux.size()) {
                std::vector<uint32_t> real(n);
                memcpy(real.data(), cg + 6, 4ull * n);
                for (uint32_t c : real) if ((c & 0xf) > 8) return -2;
                r.cigar.swap(real);
                r.aux.erase(r.aux.begin() + (long)off - 2, r.aux.begin() + (long)(off + 6 + 4ull * n));   // drop the tag like htslib does
            }
        }
    }
    return 0;
}

int AlnReader::next_raw(Record &r)
{
    if (!gz_) return -1;
    if (is_bam_) return read_bam(r);
    for (;;) {
        if (have_pending_) { line_ = pending_; have_pending_ = false; }
        else if (!getline(line_)) return -1;
        if (line_.empty()) continue;
        std::vector<char> buf(line_.begin(), line_.end()); buf.push_back(0);
        return parse_sam(buf.data(), r);
    }
}

int AlnReader::next(Record &r)
{
    for (;;) {
        int ret = next_raw(r);
        if (ret < 0) return ret;
        if (has_reg_) {
            if (r.tid != rtid_) continue;
            if (!(r.pos < rend_ && r.endpos() > rbeg_)) continue;
        }
        return 0;
    }
}

// ------------------------------------------------------------------ FASTA
std::unique_ptr<Fasta> Fasta::load(const std::string &path)
{
    gzFile fp = gzopen(path.c_str(), "rb");
    if (!fp) return nullptr;
    gzbuffer(fp, 1 << 18);
    std::unique_ptr<Fasta> fa(new Fasta());
    char buf[1 << 16];
    bool in_name = false;
    while (gzgets(fp, buf, sizeof buf)) {
        size_t n = strlen(buf);
        bool eol = n && buf[n - 1] == '\n';
        if (in_name) { in_name = !eol; continue; }   // tail of a very long header line
        if (buf[0] == '>') {
            char *e = buf + 1;
            while (*e && !isspace((unsigned char)*e)) ++e;
            fa->names.emplace_back(buf + 1, e);
            fa->seqs.emplace_back();
            in_name = !eol;
        } else if (!fa->seqs.empty()) {
            std::string &s = fa->seqs.back();
            for (size_t i = 0; i < n; ++i) if (isgraph((unsigned char)buf[i])) s.push_back(buf[i]);
        }
    }
    gzclose(fp);
    return fa;
}
int Fasta::find(const std::string &n) const
{
    for (size_t i = 0; i < names.size(); ++i) if (names[i] == n) return (int)i;
    return -1;
}

// ------------------------------------------------------------------ BED (bedidx.c:102-191, :258-360)
static constexpr int kBedShift = 13;

std::unique_ptr<Bed> Bed::load(const std::string &path)
{
    gzFile fp = gzopen(path.c_str(), "rb");
    if (!fp) return nullptr;
    std::unique_ptr<Bed> bed(new Bed());
    char buf[1 << 16];
    while (gzgets(fp, buf, sizeof buf)) {
        char *ref = buf;
        while (*ref && isspace((unsigned char)*ref)) ++ref;
        if (!*ref || *ref == '#') continue;
        char *re = ref;
        while (*re && !isspace((unsigned char)*re)) ++re;
        unsigned long long b = 0, e = 0; int num = 0;
        if (*re) { *re = 0; num = sscanf(re + 1, "%llu %llu", &b, &e); }
        if (num == 1) e = b--;
        if (num < 1 || e < b) {
            if (!strcmp(ref, "browser") || !strcmp(ref, "track")) continue;
            fprintf(stderr, "[bed_read] Parse error reading \"%s\"\n", path.c_str());
            gzclose(fp);
            return nullptr;
        }
        bed->chr[ref].iv.emplace_back((int64_t)b, (int64_t)e);
    }
    gzclose(fp);
    for (auto &kv : bed->chr) {
        Chr &c = kv.second;
        std::stable_sort(c.iv.begin(), c.iv.end(), [](const std::pair<int64_t, int64_t> &x, const std::pair<int64_t, int64_t> &y) { return x.first < y.first; });
        int64_t last_end = 0;
        for (size_t i = 0; i < c.iv.size(); ++i) {
            int64_t bb = c.iv[i].first >= 0 ? c.iv[i].first >> kBedShift : 0, ee = c.iv[i].second >= 0 ? c.iv[i].second >> kBedShift : 0;
            if (ee < last_end) continue;
            if ((size_t)ee + 1 > c.idx.size()) c.idx.resize((size_t)ee + 1, 0);
            int64_t j;
            for (j = last_end; j < bb; ++j) c.idx[(size_t)j] = i > 0 ? (int)i - 1 : 0;
            for (; j <= ee; ++j) c.idx[(size_t)j] = (int)i;
            last_end = ee + 1;
        }
        c.max_idx = last_end;
    }
    return bed;
}

bool Bed::overlap(const std::string &name, int64_t beg, int64_t end) const
{
    auto it = chr.find(name);
    if (it == chr.end() || it->second.iv.empty()) return false;
    const Chr &c = it->second;
    size_t off = 0;
    if (!c.idx.empty() && c.max_idx > 0 && beg >= 0)
        off = (size_t)((beg >> kBedShift) >= c.max_idx ? c.idx[(size_t)c.max_idx - 1] : c.idx[(size_t)(beg >> kBedShift)]);
    for (size_t i = off; i < c.iv.size(); ++i) {
        if (c.iv[i].first >= end) break;
        if (c.iv[i].second > beg && c.iv[i].first < end) return true;
    }
    return false;
}

void Bed::merged(const std::string &name, std::vector<int64_t> &b, std::vector<int64_t> &e) const
{
    b.clear(); e.clear();
    auto it = chr.find(name);
    if (it == chr.end()) return;
    for (auto &iv : it->second.iv) {   // already sorted by start
        if (iv.second <= iv.first) continue;   // empty interval never contains a position
        if (!b.empty() && iv.first <= e.back()) { if (iv.second > e.back()) e.back() = iv.second; }
        else { b.push_back(iv.first); e.push_back(iv.second); }
    }
}

bool read_file_list(const std::string &path, std::vector<std::string> &out)
{
    FILE *f = fopen(path.c_str(), "r");
    if (!f) { fprintf(stderr, "%s: %s\n", path.c_str(), strerror(errno)); return false; }
    char buf[1024];
    while (fgets(buf, sizeof buf, f)) {
        size_t l = strlen(buf);
        while (l && isspace((unsigned char)buf[l - 1])) --l;
        if (!l) continue;
        buf[l] = 0;
        std::string s = buf;
        struct stat sb;
        bool url = s.compare(0, 7, "file://") == 0;
        if (url) s = s.substr(7);
        if (stat(s.c_str(), &sb) != 0) {
            fprintf(stderr, "The file list \"%s\" appears broken, could not locate: %s\n", path.c_str(), buf);
            fclose(f);
            return false;
        }
        out.push_back(s);
    }
    fclose(f);
    if (out.empty()) { fprintf(stderr, "No files read from %s\n", path.c_str()); return false; }
    return true;
}

}  // namespace b200
