"""ctypes binding of the C ABI in include/b200_pileup.h (the CUDA engine).

This is the reference-side binding a maintainer would add around the batch
tier.  It mirrors the C structs field by field; numpy
arrays are passed as plain pointers.  There is no fallback of any kind: if the
shared library is missing or no CUDA device is present the call raises.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'lib', 'libb200pileup.so')

MODE_MPILEUP, MODE_DEPTH, MODE_COVERAGE = 0, 1, 2
RB_HOST_SKIP, RB_NAME_ODD, RB_BAQ_DONE = 1, 2, 4
POS_MAX = (0x7fffffff << 32) | 0xffffffff

EXPORTS = ['b200_engine_create', 'b200_engine_destroy', 'b200_last_error', 'b200_version', 'b200_stage',
           'b200_mpileup_text', 'b200_depth_text', 'b200_coverage', 'b200_coverage_hist', 'b200_glf', 'b200_fetch_qual',
           'b200_fetch_mapq_keep', 'b200_pileup_entries', 'b200_last_kernel_ms', 'b200_last_stage_ms', 'b200_set_keep_raw', 'b200_restage', 'b200_last_stage_device_ms',
           'b200_launch_count', 'b200_last_mpileup_parts_ms', 'b200_gl_rng_draws', 'b200_last_baq_ms',
           'b200_errmod_cal', 'b200_glfgen', 'b200_cap_mapq', 'b200_mpileup_text_bound', 'b200_depth_text_bound', 'b200_bedcov']


class Batch(C.Structure):
    _fields_ = [('n_files', C.c_int32), ('n_reads', C.c_int64), ('file_start', C.c_void_p),
                ('pos', C.c_void_p), ('flag', C.c_void_p), ('mapq', C.c_void_p), ('l_qseq', C.c_void_p),
                ('n_cigar', C.c_void_p), ('cigar_off', C.c_void_p), ('qual_off', C.c_void_p), ('mtid', C.c_void_p),
                ('mpos', C.c_void_p), ('isize', C.c_void_p), ('prev_same_name', C.c_void_p), ('rbits', C.c_void_p), ('depth_clip', C.c_void_p),
                ('cigar', C.c_void_p), ('n_cigar_total', C.c_uint64), ('seq4', C.c_void_p), ('qual', C.c_void_p),
                ('qual_bytes', C.c_uint64), ('tid', C.c_int32), ('tid_len', C.c_int64), ('tid_name', C.c_char_p),
                ('ref', C.c_void_p), ('ref_beg', C.c_int64), ('ref_n', C.c_int64), ('ref_len', C.c_int64)]


class StageConf(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ('mode', 'rflag_require', 'rflag_filter', 'min_mq', 'no_orphan', 'illumina13', 'baq',
                                          'capq_thres', 'overlaps', 'max_depth', 'd_flag_excl', 'd_flag_incl', 'd_flag_require',
                                          'd_min_mapq', 'd_min_len', 'd_remove_overlaps', 'c_min_len')] + \
               [('beg', C.c_int64), ('end', C.c_int64)]


class StageStats(C.Structure):
    _fields_ = [('n_kept', C.c_int64), ('n_kept_in_window', C.c_int64), ('out_bound', C.c_uint64), ('n_cols', C.c_int64),
                ('n_reads', C.c_uint64), ('n_selected_reads', C.c_uint64), ('summed_mapq', C.c_uint64)]


class MpileupConf(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ('min_baseQ', 'all', 'rev_del', 'no_ins', 'no_del', 'no_ends', 'out_mapq', 'out_qpos',
                                          'out_qpos5', 'n_star_cols')] + \
               [('bed_beg', C.c_void_p), ('bed_end', C.c_void_p), ('n_bed', C.c_int32), ('bed_active', C.c_int32)] + \
               [('n_x', C.c_int32), ('x_off', C.c_void_p), ('x_dat', C.c_void_p), ('x_bytes', C.c_uint64), ('x_sep', C.c_char * 16)]


class DepthConf(C.Structure):
    _fields_ = [('min_qual', C.c_int32), ('count_del', C.c_int32), ('all', C.c_int32),
                ('bed_beg', C.c_void_p), ('bed_end', C.c_void_p), ('n_bed', C.c_int32), ('bed_active', C.c_int32)]


class CoverageConf(C.Structure):
    _fields_ = [('min_baseQ', C.c_int32), ('min_depth', C.c_int32)]


class CoverageSums(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ('n_covered_bases', 'summed_coverage', 'summed_baseQ', 'quality_bases', 'missing_qual')]


class Pileup1(C.Structure):
    _fields_ = [('read', C.c_int64), ('qpos', C.c_int32), ('indel', C.c_int32), ('cigar_ind', C.c_int32), ('bits', C.c_uint32)]


_lib = None


def load_library():
    """Load libb200pileup.so; raises (never falls back) when it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f'{LIB_PATH} is missing: build it with `python samtools_b200/build.py` '
                               '(there is no CPU fallback for the pileup engine)')
        lib = C.CDLL(LIB_PATH)
        lib.b200_engine_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        lib.b200_engine_destroy.argtypes = [C.c_void_p]
        lib.b200_last_error.argtypes = [C.c_void_p]; lib.b200_last_error.restype = C.c_char_p
        lib.b200_version.restype = C.c_char_p
        lib.b200_stage.argtypes = [C.c_void_p, C.POINTER(Batch), C.POINTER(StageConf), C.POINTER(StageStats)]
        lib.b200_mpileup_text.argtypes = [C.c_void_p, C.POINTER(MpileupConf), C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        lib.b200_depth_text.argtypes = [C.c_void_p, C.POINTER(DepthConf), C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        lib.b200_coverage.argtypes = [C.c_void_p, C.POINTER(CoverageConf), C.POINTER(CoverageSums)]
        lib.b200_glf.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_int64), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        lib.b200_fetch_qual.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        lib.b200_fetch_mapq_keep.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        lib.b200_pileup_entries.argtypes = [C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        lib.b200_last_kernel_ms.argtypes = [C.c_void_p]; lib.b200_last_kernel_ms.restype = C.c_double
        lib.b200_last_stage_ms.argtypes = [C.c_void_p]; lib.b200_last_stage_ms.restype = C.c_double
        lib.b200_last_stage_device_ms.argtypes = [C.c_void_p]; lib.b200_last_stage_device_ms.restype = C.c_double
        lib.b200_set_keep_raw.argtypes = [C.c_void_p, C.c_int]; lib.b200_set_keep_raw.restype = C.c_int
        lib.b200_restage.argtypes = [C.c_void_p, C.c_void_p]; lib.b200_restage.restype = C.c_int
        lib.b200_launch_count.argtypes = [C.c_void_p]; lib.b200_launch_count.restype = C.c_int64
        lib.b200_last_mpileup_parts_ms.argtypes = [C.c_void_p, C.POINTER(C.c_double * 3)]
        lib.b200_last_baq_ms.argtypes = [C.c_void_p]; lib.b200_last_baq_ms.restype = C.c_double
        lib.b200_gl_rng_draws.argtypes = [C.c_void_p]; lib.b200_gl_rng_draws.restype = C.c_uint64
        _lib = lib
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def default_stage_conf(mode=MODE_MPILEUP, **kw):
    """Defaults of the reference CLIs (bam_plcmd.c:1083-1093, bam2depth.c:740-754, coverage.c:311-330)."""
    c = StageConf()
    c.mode = mode
    c.rflag_filter = 4 | 256 | 512 | 1024
    c.no_orphan = 1
    c.overlaps = 1
    c.baq = 1
    c.max_depth = 8000 if mode == MODE_MPILEUP else 1000000
    c.d_flag_excl = 4 | 256 | 1024 | 512
    c.beg, c.end = 0, POS_MAX
    for k, v in kw.items():
        if not hasattr(c, k):
            raise AttributeError(k)
        setattr(c, k, v)
    return c


class Engine:
    """One engine handle = one CUDA device + stream (not thread-safe)."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = C.c_void_p()
        if self.lib.b200_engine_create(device, C.byref(h)) != 0:
            raise RuntimeError('b200_engine_create failed: no usable CUDA device (the engine has no CPU fallback)')
        self.h = h
        self._keep = None

    def close(self):
        if self.h:
            self.lib.b200_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _err(self, what):
        raise RuntimeError(f'{what}: {self.lib.b200_last_error(self.h).decode()}')

    def stage(self, soa, conf):
        """soa: dict of numpy arrays with the b200_batch_t fields (see synth.make_batch)."""
        b = Batch()
        b.n_files = len(soa['file_start']) - 1
        b.n_reads = len(soa['pos'])
        for k in ('file_start', 'pos', 'flag', 'mapq', 'l_qseq', 'n_cigar', 'cigar_off', 'qual_off', 'mtid', 'mpos', 'isize',
                  'prev_same_name', 'rbits', 'depth_clip', 'cigar', 'seq4', 'qual'):
            setattr(b, k, _ptr(soa.get(k)))
        b.n_cigar_total = len(soa['cigar'])
        b.qual_bytes = len(soa['qual'])
        b.tid = int(soa.get('tid', 0)); b.tid_len = int(soa['tid_len'])
        name = soa['tid_name'].encode()
        b.tid_name = name
        ref = soa.get('ref')
        b.ref = _ptr(ref); b.ref_beg = int(soa.get('ref_beg', 0)); b.ref_n = 0 if ref is None else len(ref)
        b.ref_len = 0 if ref is None else int(soa.get('ref_len', len(ref)))
        st = StageStats()
        self._keep = (soa, name)
        if self.lib.b200_stage(self.h, C.byref(b), C.byref(conf), C.byref(st)) != 0:
            self._err('b200_stage')
        return st

    def set_keep_raw(self, on=True):
        """keep pristine qualities / mapq resident so that restage() can repeat the device side of the read stage"""
        if self.lib.b200_set_keep_raw(self.h, 1 if on else 0) != 0:
            self._err('b200_set_keep_raw')

    def restage(self):
        st = StageStats()
        if self.lib.b200_restage(self.h, C.byref(st)) != 0:
            self._err('b200_restage')
        return st

    def _text(self, fn, conf, out=None, fetch=True):
        n = C.c_size_t(0)
        if not fetch:
            if fn(self.h, C.byref(conf), None, 0, C.byref(n)) != 0:
                self._err('column stage')
            return n.value
        if out is None:
            if fn(self.h, C.byref(conf), None, 0, C.byref(n)) != 0:
                self._err('column stage')
            out = np.empty(n.value + 1, dtype=np.uint8)
        if fn(self.h, C.byref(conf), _ptr(out), out.nbytes, C.byref(n)) != 0:
            self._err('column stage')
        return out[:n.value].tobytes()

    def mpileup_text(self, conf=None, out=None, fetch=True, **kw):
        conf = conf or mpileup_conf(**kw)
        return self._text(self.lib.b200_mpileup_text, conf, out, fetch)

    def depth_text(self, conf=None, out=None, fetch=True, **kw):
        if conf is None:
            conf = DepthConf()
            for k, v in kw.items():
                setattr(conf, k, v)
        return self._text(self.lib.b200_depth_text, conf, out, fetch)

    def coverage(self, min_baseQ=0, min_depth=1):
        c = CoverageConf(min_baseQ, min_depth); s = CoverageSums()
        if self.lib.b200_coverage(self.h, C.byref(c), C.byref(s)) != 0:
            self._err('b200_coverage')
        return {k: getattr(s, k) for k, _ in CoverageSums._fields_}

    def glf(self, min_baseQ, cap_cols, n_files=1, fetch=True):
        n = C.c_int64(0)
        if not fetch:     # compute only; the likelihoods stay in HBM
            if self.lib.b200_glf(self.h, min_baseQ, C.byref(n), None, None, None, None, 0) != 0:
                self._err('b200_glf')
            return n.value
        pos = np.zeros(cap_cols, np.int64); nb = np.zeros(cap_cols * n_files, np.int32)
        qs = np.zeros(cap_cols * n_files * 4, np.float32); p25 = np.zeros(cap_cols * n_files * 25, np.float32)
        if self.lib.b200_glf(self.h, min_baseQ, C.byref(n), _ptr(pos), _ptr(nb), _ptr(qs), _ptr(p25), cap_cols) != 0:
            self._err('b200_glf')
        k = n.value
        return pos[:k], nb[:k * n_files].reshape(k, n_files), qs[:k * n_files * 4].reshape(k, n_files, 4), p25[:k * n_files * 25].reshape(k, n_files, 25)

    def fetch_qual(self, nbytes):
        q = np.zeros(nbytes, np.uint8)
        if self.lib.b200_fetch_qual(self.h, _ptr(q), nbytes) != 0:
            self._err('b200_fetch_qual')
        return q

    def pileup_entries(self, file, beg, end, cap):
        ncol = end - beg
        col_n = np.zeros(max(ncol, 1), np.uint32)
        ents = np.zeros(cap, dtype=np.dtype([('read', '<i8'), ('qpos', '<i4'), ('indel', '<i4'), ('cigar_ind', '<i4'), ('bits', '<u4')]))
        n = C.c_size_t(0)
        if self.lib.b200_pileup_entries(self.h, file, beg, end, _ptr(col_n), _ptr(ents), cap, C.byref(n)) != 0:
            self._err('b200_pileup_entries')
        return col_n, ents[:n.value]

    @property
    def last_kernel_ms(self):
        return self.lib.b200_last_kernel_ms(self.h)

    @property
    def last_stage_ms(self):
        return self.lib.b200_last_stage_ms(self.h)

    @property
    def last_stage_device_ms(self):
        return self.lib.b200_last_stage_device_ms(self.h)

    @property
    def last_baq_ms(self):
        return self.lib.b200_last_baq_ms(self.h)

    @property
    def last_mpileup_parts_ms(self):
        a = (C.c_double * 3)()
        self.lib.b200_last_mpileup_parts_ms(self.h, C.byref(a))
        return list(a)

    @property
    def gl_rng_draws(self):
        """hts_drand48 draws consumed so far by errmod_cal's ks_shuffle (columns with more than 255 usable bases)"""
        return self.lib.b200_gl_rng_draws(self.h)

    @property
    def launches(self):
        return self.lib.b200_launch_count(self.h)


def mpileup_conf(**kw):
    c = MpileupConf()
    c.min_baseQ = 13
    for k, v in kw.items():
        if not hasattr(c, k):
            raise AttributeError(k)
        setattr(c, k, v)
    return c
