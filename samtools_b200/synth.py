"""Seeded synthetic workloads of SURVEY.md section 8(d), as SoA batches (and SAM text).

One contig of `length` bp at `depth`x with `read_len` bp paired reads:
  reference  uniform ACGT, 0.1 % of positions inside short N runs
  pairs      insert size ~ N(400, 50), proper-pair flags 99/147 (and 83/163),
             so ~15 % of pairs overlap; both mates on this contig
  mapq       60 for 90 % of reads, uniform 0..59 otherwise
  qualities  Illumina-like: Q37 plateau decaying towards Q20 at the 3' end, +-3 noise, floor 2
  bases      reference with substitutions at the per-base error rate (+0.1 % SNPs)
  CIGAR      97 % <L>M, 1.5 % one insertion (<=10), 1.5 % one deletion (<=10),
             2 % soft-clipped end, 0.1 % ref-skip (N)  (fractions configurable)
  flags      0.5 % get DUP / SECONDARY / QCFAIL to exercise the filters
Everything derives from numpy Generator(seed); the same seed gives the same
batch on every machine.  `write_sam()` writes the identical records as SAM text
so the CPU oracle (or a real samtools) reads the same data.
"""
import numpy as np

NT16 = {'A': 1, 'C': 2, 'G': 4, 'T': 8, 'N': 15}
_CODE2CH = np.frombuffer(b'=ACMGRSVTWYHKDBN', dtype=np.uint8)


def make_reference(length, seed=1, n_frac=0.001):
    rng = np.random.default_rng(seed)
    ref = np.frombuffer(b'ACGT', dtype=np.uint8)[rng.integers(0, 4, size=length)]
    n_runs = max(1, int(length * n_frac / 20)) if n_frac > 0 else 0
    for s in rng.integers(0, max(1, length - 20), size=n_runs):
        ref[s:s + int(rng.integers(5, 35))] = ord('N')
    return ref


def _name_odd(pair_id, width=9):
    """__ac_Wang_hash(__ac_X31_hash_string("p%0*d")) & 1, vectorised (uint32 wraparound)."""
    h = np.full(pair_id.shape, ord('p'), dtype=np.uint32)
    div = 10 ** (width - 1)
    with np.errstate(over='ignore'):
        for _ in range(width):
            d = ((pair_id // div) % 10).astype(np.uint32) + np.uint32(ord('0'))
            h = h * np.uint32(31) + d
            div //= 10
        h = h + ~(h << np.uint32(15)); h ^= (h >> np.uint32(10)); h = h + (h << np.uint32(3)); h ^= (h >> np.uint32(6))
        h = h + ~(h << np.uint32(11)); h ^= (h >> np.uint32(16))
    return (h & 1).astype(np.uint8)


def make_batch(length=1_000_000, depth=30, read_len=150, seed=2, ref=None, ref_seed=1, with_ref=True, tid=0,
               tid_name='chr1', frac_ins=0.015, frac_del=0.015, frac_clip=0.02, frac_skip=0.001, frac_flag=0.005,
               paired=True, n_pairs=None, start_lo=0, start_span=None):
    """Returns a dict with the b200_batch_t arrays (+ 'pair_id', 'ref' ...).
    n_pairs / start_lo / start_span override the uniform whole-contig placement (used by make_panel: the
    leftmost mates start inside [start_lo, start_lo + start_span))."""
    rng = np.random.default_rng(seed)
    L = read_len
    if ref is None:
        ref = make_reference(length, ref_seed)
    if n_pairs is None:
        n_pairs = int(round(depth * length / (2.0 * L)))
    isize = np.clip(rng.normal(400, 50, n_pairs).round().astype(np.int64), L, 1000)
    if start_span is None:
        start_span = max(1, length - 3200 - start_lo)                               # reads (even with a 2 kb N skip) stay inside the contig
    start = (start_lo + rng.integers(0, max(1, start_span), size=n_pairs)).astype(np.int64)
    # which mate is first / forward
    fwd_first = rng.random(n_pairs) < 0.5
    n = 2 * n_pairs
    pair_id = np.repeat(np.arange(n_pairs, dtype=np.int64), 2)
    is_second = np.tile(np.array([0, 1], dtype=np.int8), n_pairs)          # 0: leftmost mate, 1: rightmost
    pos = np.where(is_second == 0, np.repeat(start, 2), np.repeat(start + isize - L, 2))
    rev = is_second.astype(bool)                                           # rightmost mate is the reverse one
    read1 = np.where(np.repeat(fwd_first, 2), is_second == 0, is_second == 1)
    flag = np.full(n, 1 | 2, dtype=np.uint16)
    flag |= np.where(rev, 16, 32).astype(np.uint16)
    flag |= np.where(read1, 64, 128).astype(np.uint16)
    if not paired:
        flag = np.where(rev, 16, 0).astype(np.uint16)
    # CIGAR class: 0 M, 1 ins, 2 del, 3 clip-left, 4 clip-right, 5 ref-skip
    u = rng.random(n)
    cls = np.zeros(n, dtype=np.int8)
    edges = np.cumsum([frac_ins, frac_del, frac_clip / 2, frac_clip / 2, frac_skip])
    cls[u < edges[4]] = 5; cls[u < edges[3]] = 4; cls[u < edges[2]] = 3; cls[u < edges[1]] = 2; cls[u < edges[0]] = 1
    k = np.minimum(rng.geometric(0.4, n), 10).astype(np.int64)              # indel / clip length
    k = np.where(cls == 5, rng.integers(200, 2000, n), k)
    k = np.where((cls == 3) | (cls == 4), rng.integers(1, 30, n), k)
    a = rng.integers(10, L - 20, n).astype(np.int64)                       # bases before the event
    # query index -> reference index (or -1 = random base); only the ~6 % non-<L>M reads need fixing up
    j = np.arange(L, dtype=np.int32)[None, :]
    refidx = pos.astype(np.int32)[:, None] + j
    sp = np.nonzero(cls != 0)[0]
    if len(sp):
        A, K, P, CL = a[sp, None].astype(np.int32), k[sp, None].astype(np.int32), pos[sp, None].astype(np.int32), cls[sp, None]
        r = P + j
        r = np.where((CL == 1) & (j >= A) & (j < A + K), -1, r)
        r = np.where((CL == 1) & (j >= A + K), P + j - K, r)
        r = np.where(((CL == 2) | (CL == 5)) & (j >= A), P + j + K, r)
        r = np.where((CL == 3) & (j < K), -1, r)
        r = np.where((CL == 3) & (j >= K), P + j - K, r)
        r = np.where((CL == 4) & (j >= L - K), -1, r)
        refidx[sp] = r
    # qualities
    prof = np.where(np.arange(L) < L * 0.6, 37.0, 37.0 - 17.0 * (np.arange(L) - L * 0.6) / (L * 0.4)).astype(np.float32)
    q = np.clip(np.rint(prof[None, :] + 3.0 * rng.standard_normal((n, L), dtype=np.float32)), 2, 41).astype(np.uint8)
    q[rev] = q[rev, ::-1]                                                  # 3' end is the left end of a reverse read
    # bases
    base = ref[np.clip(refidx, 0, length - 1)]
    perr = (10.0 ** (-np.arange(64, dtype=np.float32) / 10.0) + 0.001).astype(np.float32)[q]
    mut = rng.random((n, L), dtype=np.float32) < perr
    mut |= (refidx < 0) | (refidx >= length)
    nm = int(mut.sum())
    base[mut] = np.frombuffer(b'ACGT', dtype=np.uint8)[rng.integers(0, 4, size=nm, dtype=np.uint8)]
    lut = np.full(256, 15, dtype=np.uint8)
    for ch, v in NT16.items():
        lut[ord(ch)] = v
    code = lut[base]
    # mapq + flags
    mapq = np.where(rng.random(n) < 0.9, 60, rng.integers(0, 60, n)).astype(np.uint8)
    ff = rng.random(n)
    flag = flag | np.where(ff < frac_flag / 3, 1024, 0).astype(np.uint16)
    flag = flag | np.where((ff >= frac_flag / 3) & (ff < 2 * frac_flag / 3), 256, 0).astype(np.uint16)
    flag = flag | np.where((ff >= 2 * frac_flag / 3) & (ff < frac_flag), 512, 0).astype(np.uint16)
    # sort by position (stable: keeps pair order for ties)
    order = np.argsort(pos, kind='stable')
    pos, flag, mapq, cls, k, a, pair_id, code, q = pos[order], flag[order], mapq[order], cls[order], k[order], a[order], pair_id[order], code[order], q[order]
    # CIGAR ops
    n_cigar = np.where(cls == 0, 1, np.where((cls == 3) | (cls == 4), 2, 3)).astype(np.uint32)
    cigar_off = np.concatenate([[0], np.cumsum(n_cigar)[:-1]]).astype(np.uint64)
    cigar = np.zeros(int(n_cigar.sum()), dtype=np.uint32)
    M, I, D, N_, S = 0, 1, 2, 3, 4
    o = cigar_off.astype(np.int64)
    m0 = cls == 0; cigar[o[m0]] = (L << 4) | M
    for c_, op in ((1, I), (2, D), (5, N_)):
        m = cls == c_
        cigar[o[m]] = (a[m].astype(np.uint32) << 4) | M
        cigar[o[m] + 1] = (k[m].astype(np.uint32) << 4) | op
        rest = (L - a[m] - (k[m] if c_ == 1 else 0)).astype(np.uint32)
        cigar[o[m] + 2] = (rest << 4) | M
    m = cls == 3; cigar[o[m]] = (k[m].astype(np.uint32) << 4) | S; cigar[o[m] + 1] = ((L - k[m]).astype(np.uint32) << 4) | M
    m = cls == 4; cigar[o[m]] = ((L - k[m]).astype(np.uint32) << 4) | M; cigar[o[m] + 1] = (k[m].astype(np.uint32) << 4) | S
    rlen = np.where(cls == 1, L - k, np.where((cls == 2) | (cls == 5), L + k, np.where((cls == 3) | (cls == 4), L - k, L))).astype(np.int64)
    # mates
    by_pair = np.argsort(pair_id, kind='stable')
    first, second = by_pair[0::2], by_pair[1::2]
    prev = np.full(n, -1, dtype=np.int64)
    prev[second] = first
    mpos = np.zeros(n, dtype=np.int64); mpos[first] = pos[second]; mpos[second] = pos[first]
    tl = np.zeros(n, dtype=np.int64)
    span = np.maximum(pos[second] + rlen[second], pos[first] + rlen[first]) - pos[first]
    tl[first] = span; tl[second] = -span
    if not paired:
        prev[:] = -1; mpos[:] = -1; tl[:] = 0
    # pack
    lq = L + (L & 1)
    qual = np.zeros((n, lq), dtype=np.uint8); qual[:, :L] = q
    cpad = np.zeros((n, lq), dtype=np.uint8); cpad[:, :L] = code
    seq4 = ((cpad[:, 0::2] << 4) | cpad[:, 1::2]).reshape(-1)
    qual_off = (np.arange(n, dtype=np.uint64) * np.uint64(lq))
    rbits = (_name_odd(pair_id) * 2).astype(np.uint8)
    return dict(file_start=np.array([0, n], dtype=np.int64), pos=pos.astype(np.int64), flag=flag, mapq=mapq,
                l_qseq=np.full(n, L, dtype=np.int32), n_cigar=n_cigar, cigar_off=cigar_off, qual_off=qual_off,
                mtid=np.full(n, tid if paired else -1, dtype=np.int32), mpos=mpos, isize=tl, prev_same_name=prev, rbits=rbits,
                cigar=cigar, seq4=np.ascontiguousarray(seq4), qual=np.ascontiguousarray(qual.reshape(-1)),
                tid=tid, tid_len=length, tid_name=tid_name, ref=ref if with_ref else None, ref_beg=0, ref_len=length,
                pair_id=pair_id, read_len=L, ref_full=ref)


def make_region(length, depth=30, read_len=150, seed=2, chunk=1_000_000, with_ref=False, tid_name='chr1', **kw):
    """A `length`-bp window built from independent `chunk`-sized pieces (bounded generator memory).
    Reads never cross a chunk edge, so the concatenation stays coordinate sorted."""
    parts, off = [], 0
    i = 0
    while off < length:
        ln = min(chunk, length - off)
        parts.append((off, make_batch(length=ln, depth=depth, read_len=read_len, seed=seed * 1000 + i, ref_seed=seed * 1000 + i + 500,
                                      with_ref=True, tid_name=tid_name, **kw)))
        off += ln; i += 1
    out = {}
    nread = np.cumsum([0] + [len(p['pos']) for _, p in parts])
    ncig = np.cumsum([0] + [len(p['cigar']) for _, p in parts])
    nq = np.cumsum([0] + [len(p['qual']) for _, p in parts])
    npair = np.cumsum([0] + [int(p['pair_id'].max()) + 1 if len(p['pair_id']) else 0 for _, p in parts])
    cat = lambda key, add=None: np.concatenate([(p[key] + (add[i] if add is not None else 0)).astype(p[key].dtype) for i, (_, p) in enumerate(parts)])
    offs = [o for o, _ in parts]
    out['pos'] = cat('pos', offs); out['mpos'] = cat('mpos', offs)
    for k_ in ('flag', 'mapq', 'l_qseq', 'n_cigar', 'mtid', 'isize', 'cigar', 'seq4', 'qual'):
        out[k_] = cat(k_)
    out['cigar_off'] = cat('cigar_off', [np.uint64(x) for x in ncig[:-1]]); out['qual_off'] = cat('qual_off', [np.uint64(x) for x in nq[:-1]])
    out['prev_same_name'] = np.concatenate([np.where(p['prev_same_name'] >= 0, p['prev_same_name'] + nread[i], -1) for i, (_, p) in enumerate(parts)])
    out['pair_id'] = cat('pair_id', list(npair[:-1]))
    out['rbits'] = (_name_odd(out['pair_id']) * 2).astype(np.uint8)
    out['file_start'] = np.array([0, nread[-1]], dtype=np.int64)
    ref = np.concatenate([p['ref_full'] for _, p in parts])
    out.update(tid=0, tid_len=length, tid_name=tid_name, ref=ref if with_ref else None, ref_beg=0, ref_len=length, read_len=read_len, ref_full=ref)
    return out


def algorithmic_bytes_in(soa, overlap=True):
    """sum over reads of ceil(l/2) + l + 4*n_cigar + 24 (+24 mate fields) (+ ref bytes touched)."""
    l = soa['l_qseq'].astype(np.int64)
    b = int(((l + 1) // 2 + l + 4 * soa['n_cigar'].astype(np.int64) + 24 + (24 if overlap else 0)).sum())
    return b


def write_fasta(path, name, ref, width=60):
    with open(path, 'w') as f:
        f.write(f'>{name}\n')
        s = ref.tobytes().decode()
        for i in range(0, len(s), width):
            f.write(s[i:i + width] + '\n')


def write_sam(path, soa, max_reads=None):
    """The same records as SAM text (names p%09d, so the name-hash bit matches rbits)."""
    name = soa['tid_name']; L = soa['read_len']
    lq = L + (L & 1)
    n = len(soa['pos']) if max_reads is None else min(max_reads, len(soa['pos']))
    seq4 = soa['seq4'].reshape(-1, lq // 2); qual = soa['qual'].reshape(-1, lq)
    ops = 'MIDNSHP=XB'
    with open(path, 'w') as f:
        f.write(f'@HD\tVN:1.6\tSO:coordinate\n@SQ\tSN:{name}\tLN:{soa["tid_len"]}\n')
        for i in range(n):
            co = int(soa['cigar_off'][i]); nc = int(soa['n_cigar'][i])
            cg = ''.join(f'{int(c) >> 4}{ops[int(c) & 15]}' for c in soa['cigar'][co:co + nc])
            nib = np.empty(lq, dtype=np.uint8); nib[0::2] = seq4[i] >> 4; nib[1::2] = seq4[i] & 15
            seq = _CODE2CH[nib[:L]].tobytes().decode()
            ql = (qual[i, :L] + 33).tobytes().decode()
            paired = bool(soa['flag'][i] & 1)
            f.write(f'p{int(soa["pair_id"][i]):09d}\t{int(soa["flag"][i])}\t{name}\t{int(soa["pos"][i]) + 1}\t{int(soa["mapq"][i])}\t{cg}\t'
                    f'{"=" if paired else "*"}\t{int(soa["mpos"][i]) + 1}\t{int(soa["isize"][i])}\t{seq}\t{ql}\n')


def ref_span(soa):
    """bam_cigar2rlen per read (reference bases consumed: M, D, N, =, X)."""
    cg = soa['cigar']
    op = cg & 0xf
    ln = np.where((op == 0) | (op == 2) | (op == 3) | (op == 7) | (op == 8), cg >> 4, 0).astype(np.int64)
    cs = np.concatenate([[0], np.cumsum(ln)])
    o = soa['cigar_off'].astype(np.int64)
    return cs[o + soa['n_cigar'].astype(np.int64)] - cs[o]


def _gather_ranges(off, cnt):
    """indices off[i] .. off[i]+cnt[i]-1 for every i, concatenated."""
    cnt = cnt.astype(np.int64)
    tot = int(cnt.sum())
    if tot == 0:
        return np.zeros(0, dtype=np.int64)
    starts = np.concatenate([[0], np.cumsum(cnt)[:-1]])
    return np.repeat(off.astype(np.int64) - starts, cnt) + np.arange(tot, dtype=np.int64)


def take_reads(soa, idx):
    """Sub-batch holding reads `idx` (ascending batch indices, or a permutation), payload re-packed; name links that
    leave the selection become -1.  All reads must have the same l_qseq (true of this generator)."""
    idx = np.asarray(idx, dtype=np.int64)
    n0 = len(soa['pos'])
    L = soa['read_len']; lq = L + (L & 1)
    out = dict(soa)
    for k in ('pos', 'flag', 'mapq', 'l_qseq', 'n_cigar', 'mtid', 'mpos', 'isize', 'rbits', 'pair_id'):
        out[k] = np.ascontiguousarray(soa[k][idx])
    inv = np.full(n0 + 1, -1, dtype=np.int64)
    inv[idx] = np.arange(len(idx), dtype=np.int64)
    prev = soa['prev_same_name'][idx]
    out['prev_same_name'] = np.where(prev >= 0, inv[np.where(prev >= 0, prev, n0)], -1).astype(np.int64)
    rows = (soa['qual_off'][idx] // np.uint64(lq)).astype(np.int64)
    out['qual'] = np.ascontiguousarray(soa['qual'].reshape(-1, lq)[rows].reshape(-1))
    out['seq4'] = np.ascontiguousarray(soa['seq4'].reshape(-1, lq // 2)[rows].reshape(-1))
    out['qual_off'] = np.arange(len(idx), dtype=np.uint64) * np.uint64(lq)
    out['cigar'] = np.ascontiguousarray(soa['cigar'][_gather_ranges(soa['cigar_off'][idx], out['n_cigar'])])
    out['cigar_off'] = np.concatenate([[0], np.cumsum(out['n_cigar'].astype(np.int64))[:-1]]).astype(np.uint64)
    out['file_start'] = np.array([0, len(idx)], dtype=np.int64)
    return out


def make_panel(n_targets=500, target_len=2000, depth=200, n_hot=20, hot_depth=2000, hot_span=100, spacing=10_000,
               read_len=150, seed=4, with_ref=True, tid_name='chr1', **kw):
    """BASELINE config 4 shape: `n_targets` targets of `target_len` bp at `depth`x mean, `n_hot` of them
    carrying a hotspot where another `hot_depth`x of pairs start inside `hot_span` bp (columns with thousands of reads,
    i.e. errmod_cal's n > 255 shuffle path).  One contig, targets `spacing` bp apart; coordinate sorted."""
    L = read_len
    length = n_targets * spacing
    rng = np.random.default_rng(seed)
    hot = set(rng.choice(n_targets, size=min(n_hot, n_targets), replace=False).tolist()) if n_hot else set()
    ref = make_reference(length, seed * 1000 + 1)
    parts = []
    npair = 0
    for t in range(n_targets):
        off = t * spacing
        sub = ref[off:off + spacing]
        pieces = [dict(n_pairs=int(round(depth * target_len / (2.0 * L))), start_lo=200, start_span=target_len, seed=seed * 100003 + 2 * t)]
        if t in hot:
            pieces.append(dict(n_pairs=int(round(hot_depth * (hot_span + L) / (2.0 * L))), start_lo=200 + target_len // 2, start_span=hot_span,
                               seed=seed * 100003 + 2 * t + 1))
        for pc in pieces:
            b = make_batch(length=spacing, read_len=L, ref=sub, with_ref=True, tid_name=tid_name, **pc, **kw)
            parts.append((off, b, npair))
            npair += int(b['pair_id'].max()) + 1 if len(b['pair_id']) else 0
    nread = np.cumsum([0] + [len(p['pos']) for _, p, _ in parts])
    ncig = np.cumsum([0] + [len(p['cigar']) for _, p, _ in parts])
    nq = np.cumsum([0] + [len(p['qual']) for _, p, _ in parts])
    out = {}
    out['pos'] = np.concatenate([p['pos'] + o for o, p, _ in parts]); out['mpos'] = np.concatenate([p['mpos'] + o for o, p, _ in parts])
    for k_ in ('flag', 'mapq', 'l_qseq', 'n_cigar', 'mtid', 'isize', 'cigar', 'seq4', 'qual'):
        out[k_] = np.concatenate([p[k_] for _, p, _ in parts])
    out['cigar_off'] = np.concatenate([p['cigar_off'] + np.uint64(ncig[i]) for i, (_, p, _) in enumerate(parts)])
    out['qual_off'] = np.concatenate([p['qual_off'] + np.uint64(nq[i]) for i, (_, p, _) in enumerate(parts)])
    out['prev_same_name'] = np.concatenate([np.where(p['prev_same_name'] >= 0, p['prev_same_name'] + nread[i], -1) for i, (_, p, _) in enumerate(parts)])
    out['pair_id'] = np.concatenate([p['pair_id'] + base for _, p, base in parts])
    out['rbits'] = (_name_odd(out['pair_id']) * 2).astype(np.uint8)
    out['file_start'] = np.array([0, nread[-1]], dtype=np.int64)
    out.update(tid=0, tid_len=length, tid_name=tid_name, ref=ref if with_ref else None, ref_beg=0, ref_len=length, read_len=L, ref_full=ref)
    order = np.argsort(out['pos'], kind='stable')              # hotspot pieces interleave with their target's reads
    res = take_reads(out, order)
    res['hot_targets'] = sorted(hot)
    return res


def _reg2bin(beg, end):
    """SAMv1 5.3 reg2bin, vectorised (end exclusive)."""
    end = end - 1
    b = np.zeros(len(beg), dtype=np.int64)
    done = np.zeros(len(beg), dtype=bool)
    for shift, base in ((14, 4681), (17, 585), (20, 73), (23, 9), (26, 1)):
        m = ~done & ((beg >> shift) == (end >> shift))
        b[m] = base + (beg[m] >> shift)
        done |= m
    return b.astype(np.uint16)


def write_bam(path, soa, level=1, block=0xff00):
    """The same records as a coordinate-sorted BAM (BGZF, SAMv1 4.2): what `samtools mpileup` really reads.
    Names are p%09d like write_sam().  No index is written (the readers here scan)."""
    import struct
    import zlib
    name = soa['tid_name'].encode(); L = soa['read_len']; lq = L + (L & 1)
    n = len(soa['pos'])
    text = f'@HD\tVN:1.6\tSO:coordinate\n@SQ\tSN:{soa["tid_name"]}\tLN:{soa["tid_len"]}\n'.encode()
    hdr = b'BAM\1' + struct.pack('<i', len(text)) + text + struct.pack('<i', 1) + struct.pack('<i', len(name) + 1) + name + b'\0' + \
        struct.pack('<i', int(soa['tid_len']))
    ncg = soa['n_cigar'].astype(np.int64)
    fixed = 32 + 11 + (L + 1) // 2 + L
    size = fixed + 4 * ncg                                   # block_size field excluded
    off = np.concatenate([[0], np.cumsum(size + 4)])
    buf = np.zeros(int(off[-1]), dtype=np.uint8)
    rl = ref_span(soa)
    pos = soa['pos'].astype(np.int64)
    core = np.zeros(n, dtype=np.dtype([('bs', '<i4'), ('tid', '<i4'), ('pos', '<i4'), ('lname', 'u1'), ('mapq', 'u1'), ('bin', '<u2'), ('ncig', '<u2'),
                                        ('flag', '<u2'), ('lseq', '<i4'), ('mtid', '<i4'), ('mpos', '<i4'), ('tlen', '<i4')]))
    core['bs'] = size; core['tid'] = soa.get('tid', 0); core['pos'] = pos; core['lname'] = 11; core['mapq'] = soa['mapq']
    core['bin'] = _reg2bin(pos, pos + np.maximum(rl, 1)); core['ncig'] = ncg; core['flag'] = soa['flag']; core['lseq'] = L
    core['mtid'] = soa['mtid']; core['mpos'] = soa['mpos']; core['tlen'] = soa['isize']
    cb = core.view(np.uint8).reshape(n, 36)
    digits = np.zeros((n, 11), dtype=np.uint8); digits[:, 0] = ord('p')
    pid = soa['pair_id'].astype(np.int64)
    for k in range(9):
        digits[:, 9 - k] = ord('0') + (pid // 10 ** k) % 10
    rows = (soa['qual_off'] // np.uint64(lq)).astype(np.int64)
    seq = soa['seq4'].reshape(-1, lq // 2)[rows][:, :(L + 1) // 2].copy()
    if L & 1:
        seq[:, -1] &= 0xf0
    qual = soa['qual'].reshape(-1, lq)[rows][:, :L]
    o = off[:-1]
    for cols, at in ((cb, 0), (digits, 36)):
        buf[(o[:, None] + at + np.arange(cols.shape[1])[None, :]).reshape(-1)] = cols.reshape(-1)
    cig_bytes = soa['cigar'].astype('<u4').view(np.uint8)
    dst = np.repeat(o + 47, 4 * ncg) + (np.arange(int(4 * ncg.sum())) - np.repeat(np.concatenate([[0], np.cumsum(4 * ncg)[:-1]]), 4 * ncg))
    src = _gather_ranges(4 * soa['cigar_off'].astype(np.int64), 4 * ncg)
    buf[dst] = cig_bytes[src]
    so = o + 47 + 4 * ncg
    buf[(so[:, None] + np.arange(seq.shape[1])[None, :]).reshape(-1)] = seq.reshape(-1)
    qo = so + seq.shape[1]
    buf[(qo[:, None] + np.arange(L)[None, :]).reshape(-1)] = qual.reshape(-1)
    data = hdr + buf.tobytes()
    with open(path, 'wb') as f:
        for i in range(0, len(data), block):
            chunk = data[i:i + block]
            c = zlib.compressobj(level, zlib.DEFLATED, -15)
            comp = c.compress(chunk) + c.flush()
            f.write(b'\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0' + struct.pack('<H', len(comp) + 25) + comp +
                    struct.pack('<II', zlib.crc32(chunk) & 0xffffffff, len(chunk)))
        f.write(bytes.fromhex('1f8b08040000000000ff0600424302001b0003000000000000000000'))
