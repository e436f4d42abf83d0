// packer.hpp -- decoded records -> structure-of-arrays batch for the engine.
//
// The staging image of a bam1_t stream (include/b200_pileup.h, b200_batch_t):
// one array per bam1_core_t field, CIGAR ops / 4-bit bases / qualities packed
// back to back (each read's qualities start on an even offset so its bases sit
// at the same nibble index), plus the per-read host bits and the
// previous-record-with-the-same-QNAME link that replaces htslib's name hash.
#pragma once
#include "hts_io.hpp"
#include "../../../include/b200_pileup.h"
#include <unordered_map>

namespace b200 {

struct PackedBatch {
    std::vector<int64_t> file_start, pos, mpos, isize, prev, depth_clip;
    std::vector<uint16_t> flag;
    std::vector<uint8_t> mapq, rbits, seq4, qual;
    std::vector<int32_t> l_qseq, mtid;
    std::vector<uint32_t> n_cigar, cigar;
    std::vector<uint64_t> cigar_off, qual_off;
    std::string name;

    void clear()
    {
        file_start.clear(); pos.clear(); mpos.clear(); isize.clear(); prev.clear(); flag.clear(); mapq.clear(); rbits.clear();
        seq4.clear(); qual.clear(); l_qseq.clear(); mtid.clear(); n_cigar.clear(); cigar.clear(); cigar_off.clear(); qual_off.clear();
        depth_clip.clear();
    }
    void begin_file() { file_start.push_back((int64_t)pos.size()); names_.clear(); }
    void finish() { file_start.push_back((int64_t)pos.size()); if (seq4.size() * 2 < qual.size() + 2) seq4.resize((qual.size() + 2) / 2, 0); }

    // rb: B200_RB_* bits decided by the host; name hash bit is added here
    void add(const Record &r, uint8_t rb, bool link_names)
    {
        const int64_t i = (int64_t)pos.size();
        pos.push_back(r.pos); mpos.push_back(r.mpos); isize.push_back(r.isize); flag.push_back(r.flag); mapq.push_back(r.mapq);
        l_qseq.push_back(r.l_qseq); mtid.push_back(r.mtid); n_cigar.push_back((uint32_t)r.cigar.size());
        cigar_off.push_back(cigar.size());
        cigar.insert(cigar.end(), r.cigar.begin(), r.cigar.end());
        if (qual.size() & 1) qual.push_back(0);
        const uint64_t qo = qual.size();
        qual_off.push_back(qo);
        qual.insert(qual.end(), r.qual.begin(), r.qual.end());
        seq4.resize((qo + (uint64_t)r.l_qseq + 1) / 2 + 1, 0);
        for (int32_t k = 0; k < r.l_qseq; ++k) {
            const uint8_t b = (r.seq4[(size_t)k >> 1] >> ((~k & 1) << 2)) & 0xf;
            const uint64_t n = qo + (uint64_t)k;
            seq4[n >> 1] |= (uint8_t)(b << ((~n & 1) << 2));
        }
        int64_t pv = -1;
        if (link_names) {
            auto it = names_.find(r.qname);
            if (it != names_.end()) { pv = it->second; it->second = i; }
            else names_.emplace(r.qname, i);
            if (qname_hash_bit(r.qname)) rb |= B200_RB_NAME_ODD;
        }
        prev.push_back(pv);
        rbits.push_back(rb);
    }

    b200_batch_t view(int32_t tid, int64_t tid_len, const std::string &tid_name, const std::string *ref) const
    {
        b200_batch_t b;
        b.n_files = (int32_t)file_start.size() - 1; b.n_reads = (int64_t)pos.size(); b.file_start = file_start.data();
        b.pos = pos.data(); b.flag = flag.data(); b.mapq = mapq.data(); b.l_qseq = l_qseq.data(); b.n_cigar = n_cigar.data();
        b.cigar_off = cigar_off.data(); b.qual_off = qual_off.data(); b.mtid = mtid.data(); b.mpos = mpos.data(); b.isize = isize.data();
        b.prev_same_name = prev.data(); b.rbits = rbits.data();
        b.depth_clip = depth_clip.size() == pos.size() && !pos.empty() ? depth_clip.data() : nullptr;
        b.cigar = cigar.data(); b.n_cigar_total = cigar.size(); b.seq4 = seq4.data(); b.qual = qual.data(); b.qual_bytes = qual.size();
        b.tid = tid; b.tid_len = tid_len; b.tid_name = tid_name.c_str();
        b.ref = ref ? ref->data() : nullptr; b.ref_beg = 0; b.ref_n = ref ? (int64_t)ref->size() : 0; b.ref_len = ref ? (int64_t)ref->size() : 0;
        return b;
    }

private:
    std::unordered_map<std::string, int64_t> names_;
};

}  // namespace b200
