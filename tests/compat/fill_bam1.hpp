// fill_bam1.hpp -- test helper: a pull callback in the style of an htslib caller (fills a bam1_t from this
// repository's SAM/BAM reader).  Shared by plp_dump.cpp and plbuf_dump.cpp.
#pragma once
#include "b200_htslib_compat.h"
#include "../../samtools_b200/csrc/host/hts_io.hpp"
#include <cstdlib>
#include <cstring>
#include <memory>

struct Src { std::unique_ptr<b200::AlnReader> rd; };

static int pull(void *data, bam1_t *b)
{
    Src *s = (Src *)data;
    b200::Record r;
    int ret = s->rd->next(r);
    if (ret < 0) return ret;
    const size_t lq = r.qname.size() + 1, l_qname = (lq + 3) & ~(size_t)3;
    const size_t need = l_qname + 4 * r.cigar.size() + r.seq4.size() + r.qual.size() + r.aux.size();
    if (b->m_data < need) { b->data = (uint8_t *)realloc(b->data, need); b->m_data = (uint32_t)need; }
    memset(b->data, 0, l_qname);
    memcpy(b->data, r.qname.c_str(), lq);
    uint8_t *p = b->data + l_qname;
    memcpy(p, r.cigar.data(), 4 * r.cigar.size()); p += 4 * r.cigar.size();
    memcpy(p, r.seq4.data(), r.seq4.size()); p += r.seq4.size();
    memcpy(p, r.qual.data(), r.qual.size()); p += r.qual.size();
    memcpy(p, r.aux.data(), r.aux.size());
    b->l_data = (int)need;
    b->core.pos = r.pos; b->core.tid = r.tid; b->core.qual = r.mapq; b->core.flag = r.flag; b->core.l_qname = (uint16_t)l_qname;
    b->core.l_extranul = (uint8_t)(l_qname - lq); b->core.n_cigar = (uint32_t)r.cigar.size(); b->core.l_qseq = r.l_qseq;
    b->core.mtid = r.mtid; b->core.mpos = r.mpos; b->core.isize = r.isize;
    return 0;
}

