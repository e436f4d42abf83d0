#!/usr/bin/env python3
"""bench.py -- mpileup reference positions/sec on the BASELINE C2-shape workload.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--region-mb M]
  (N>1: python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...)

A "step" is one pass of the hot path over one staged batch: a 30x / 150 bp
paired-read window of M Mb (default 8 Mb = 1.6 M reads, ~0.41 GB of SoA input
and ~0.66 GB of pileup text, both far larger than the 126 MB L2, so successive
steps cannot be served from cache), `mpileup -a` semantics without a FASTA
(BASELINE.json configs[1] scaled up so one step lasts milliseconds, not
microseconds).  With N GPUs every rank owns its own M-Mb region (weak scaling,
no data-path collective; the per-step collective is the all_gather of the
per-region column summaries rank 0 needs to emit shards in genome order).

value  device-resident: inputs already staged in HBM, K column-stage launches,
       wall time between two device synchronisations (max over ranks).
e2e    through the C ABI with HOST buffers: pinned-host SoA -> b200_stage (H2D +
       read stage) -> b200_mpileup_text -> pinned-host text (D2H), every step.
roofline  algorithmic bytes / CUDA-event time of the k_mpileup launch.
cpu_baseline / --impl reference: the CPU oracle (the reference cannot be built
       here: htslib is absent) on the box's host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

METRIC = 'mpileup reference positions/sec'
UNIT = 'positions/s'


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        return float(json.load(open(p))['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    """SM clock + throttle reasons of one GPU, polled through NVML every few ms while the timed
    regions run (the recipe's nvidia-smi query, without the process start-up latency)."""
    REASONS = (('hw_slowdown', 0x8), ('sw_power_cap', 0x4), ('sw_thermal_slowdown', 0x20), ('hw_thermal_slowdown', 0x40))

    def __init__(self, index, period=0.004):
        self.index, self.period, self.rows, self.stop = index, period, [], False
        self.h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get('CUDA_VISIBLE_DEVICES')
            phys = int(vis.split(',')[index]) if vis and vis.split(',')[index].isdigit() else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.h = None

    def _poll(self):
        nv = self.nv
        while not self.stop:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((float(mhz), int(rs)))
            except Exception:
                pass
            time.sleep(self.period)

    def __enter__(self):
        if self.h is not None:
            self.t = threading.Thread(target=self._poll, daemon=True); self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        if self.h is not None:
            self.t.join(timeout=2)

    def summary(self):
        sm = [r[0] for r in self.rows]
        reasons = sorted({name for _, bits in self.rows for name, mask in self.REASONS if bits & mask})
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': getattr(self, 'max_mhz', None), 'reasons': reasons,
                'samples': len(sm), 'source': 'nvml, polled during the timed regions'}


def oracle_path():
    exe = os.path.join(ROOT, 'oracle', '_build', 'plp_oracle')
    if not os.path.exists(exe):
        subprocess.run(['make', '-s', '-C', os.path.join(ROOT, 'oracle')], check=True)
    return exe


def cpu_sample(td, sample_mb):
    """SAM text of a `sample_mb`-Mb window of the same workload shape (seeded)."""
    from samtools_b200 import synth
    n = int(sample_mb * 1e6)
    soa = synth.make_region(n, seed=99, chunk=500_000)
    sam = os.path.join(td, 'sample.sam')
    synth.write_sam(sam, soa)
    return sam, n


def run_oracle_parallel(exe, sam, procs):
    t0 = time.perf_counter()
    ps = [subprocess.Popen([exe, 'mpileup', '-a', sam], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for _ in range(procs)]
    for p in ps:
        if p.wait() != 0:
            raise RuntimeError('oracle failed')
    return time.perf_counter() - t0


def reference_arm(args):
    """--impl reference: the CPU implementation of the path on all host cores (oracle port)."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    exe = oracle_path()
    cores = os.cpu_count() or 1
    sample_mb = 1.0
    with tempfile.TemporaryDirectory() as td:
        sam, ncols = cpu_sample(td, sample_mb)
        for _ in range(max(args.warmup, 1) if args.warmup else 0):
            run_oracle_parallel(exe, sam, cores)
        t = [run_oracle_parallel(exe, sam, cores) for _ in range(args.steps)]
    total = sum(t)
    val = cores * ncols * args.steps / total
    sample = f'{cores} concurrent single-threaded oracle processes (the reference pileup is single-threaded, bam_plcmd.c:1098), ' \
             f'each `mpileup -a` over a {sample_mb:g} Mb 30x/150bp SAM-text window to /dev/null'
    line = {'metric': METRIC, 'value': val, 'unit': UNIT, 'impl': 'reference', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': 1e3 * total / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8',
            'data': 'synthetic', 'config': workload_config(args),
            'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
            'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}, 'gpu_launches': 0}
    print(json.dumps(line))


def workload_config(args):
    return {'workload': f'BASELINE C2 shape x{args.region_mb:g}: synthetic {args.region_mb:g} Mb region per GPU, 30x depth, 150 bp paired reads, '
                        f'`mpileup -a` (no FASTA -> no BAQ, overlap removal on, -Q13, -d 8000)',
            'region_mb_per_gpu': args.region_mb, 'depth': 30, 'read_len': 150,
            'l2_policy': 'inputs (0.05 GB/Mb) and outputs (0.08 GB/Mb) per step exceed the 126 MB L2; no explicit flush',
            'parallelism': f'region-shard x{args.gpus}'}


def pin(soa):
    """Keep the SoA arrays in pinned host memory so cudaMemcpyAsync runs at PCIe speed."""
    import torch
    out = dict(soa)
    for k, v in soa.items():
        if isinstance(v, np.ndarray) and v.nbytes > 0:
            t = torch.from_numpy(np.ascontiguousarray(v)).pin_memory()
            out[k] = t.numpy()
            out['_pin_' + k] = t
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--region-mb', type=float, default=8.0)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--e2e-handles', type=int, default=3,
                    help='engine handles (one host thread each) used by the end-to-end loop; handles are per-thread objects like the htslib iterators they replace')
    args = ap.parse_args()
    if args.impl == 'reference':
        return reference_arm(args)

    import torch
    import torch.distributed as dist
    from samtools_b200 import engine, synth, shard

    rank = int(os.environ.get('RANK', '0')); world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device (the pileup engine has no CPU fallback)')
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    dev = torch.device('cuda', local)

    ncols = int(args.region_mb * 1e6)
    soa = synth.make_region(ncols, seed=2 + rank)          # this rank's region
    soa['ref'] = None
    soa = pin(soa)
    n_reads = len(soa['pos'])
    h2d = int(sum(soa[k].nbytes for k in ('pos', 'flag', 'mapq', 'l_qseq', 'n_cigar', 'cigar_off', 'qual_off', 'mtid', 'mpos', 'isize',
                                          'prev_same_name', 'rbits', 'cigar', 'seq4', 'qual')))
    bytes_in = synth.algorithmic_bytes_in(soa, overlap=True)

    eng = engine.Engine(local)
    sconf = engine.default_stage_conf(engine.MODE_MPILEUP)
    mconf = engine.mpileup_conf(all=1)
    eng.stage(soa, sconf)
    out_len = eng.mpileup_text(mconf, fetch=False)
    out_host_t = torch.empty(out_len + 4096, dtype=torch.uint8).pin_memory()
    out_host = out_host_t.numpy()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # device-resident measurement: its own handle, with pristine qualities / mapq kept resident so that the read stage
    # (which edits them in place) can be repeated without the PCIe copies
    eng_dev = engine.Engine(local)
    eng_dev.set_keep_raw(True)
    eng_dev.stage(soa, sconf)
    stage_ms = []

    def step_device():
        """the whole hot path on resident inputs: read stage (filters, overlap tweak, descriptors, read slices) + column stage"""
        eng_dev.restage()
        stage_ms.append(eng_dev.last_stage_device_ms)
        n = eng_dev.mpileup_text(mconf, fetch=False)
        if world > 1:   # column summaries of every region, for ordered emission at rank 0
            shard.gather_summaries([n, ncols, n_reads], device=dev)
        return n

    # end-to-end: the caller owns a stream of windows; like htslib handles, an engine handle serves one
    # thread, so a pipeline uses one handle per host thread (H2D of one window overlaps kernels / D2H of another)
    import ctypes as C
    import threading as _th
    n_h = max(1, args.e2e_handles)
    handles = [eng] + [engine.Engine(local) for _ in range(n_h - 1)]
    outs = [out_host] + [torch.empty(out_len + 4096, dtype=torch.uint8).pin_memory().numpy() for _ in range(n_h - 1)]

    def one_window(h, buf):
        h.stage(soa, sconf)
        n = C.c_size_t(0)
        if h.lib.b200_mpileup_text(h.h, C.byref(mconf), buf.ctypes.data_as(C.c_void_p), buf.nbytes, C.byref(n)) != 0:
            raise RuntimeError(h.lib.b200_last_error(h.h).decode())
        return n.value

    def run_e2e(k_steps):
        """k_steps windows through n_h handles; returns bytes of the last window"""
        res = [0] * n_h
        def worker(j):
            for s_ in range(j, k_steps, n_h):
                res[j] = one_window(handles[j], outs[j])
        ths = [_th.Thread(target=worker, args=(j,)) for j in range(n_h)]
        [t.start() for t in ths]; [t.join() for t in ths]
        if world > 1:
            shard.gather_summaries([max(res), ncols, n_reads], device=dev)
        return max(res)

    def step_e2e():
        return run_e2e(n_h)

    for _ in range(args.warmup):
        step_e2e(); step_device()

    # ---- device-resident: K passes of read stage + column stage
    kernel_ms, parts_ms = [], []
    stage_ms.clear()
    with ClockSampler(local) as clk:
        sync_all()
        l0 = eng_dev.launches
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_device()
            kernel_ms.append(eng_dev.last_kernel_ms)
            parts_ms.append(eng_dev.last_mpileup_parts_ms)
        sync_all()
        dt = time.perf_counter() - t0
        launches = eng_dev.launches - l0
        # ---- end to end through the C ABI with host buffers
        sync_all()
        t0 = time.perf_counter()
        run_e2e(args.steps)
        sync_all()
        dt_e2e = time.perf_counter() - t0
    dt = shard.max_over_ranks(dt, dev); dt_e2e = shard.max_over_ranks(dt_e2e, dev)
    clocks = clk.summary()

    if rank == 0:
        peak, peak_src = peaks()
        kms = float(np.mean(kernel_ms))
        size_ms, scan_ms, write_ms = [float(x) for x in np.mean(np.array(parts_ms), axis=0)]
        # dominant kernel = the write pass: it re-reads every staged read (descriptor, qualities, bases) and
        # writes every text byte once -> its algorithmic bytes are bytes_in + bytes_out
        alg = bytes_in + out_len
        achieved = alg / (write_ms * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, 'profiles', 'r01_traffic.json')
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            if abs(tj.get('region_mb', 0) - args.region_mb) < 1e-9:
                traffic = tj.get('k_mpileup_write_dram_bytes')
        value = world * ncols * args.steps / dt
        e2e = world * ncols * args.steps / dt_e2e
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
                'ms_per_step': 1e3 * dt / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8',
                'data': 'synthetic', 'config': workload_config(args), 'clocks': clocks,
                'e2e': {'value': e2e, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': int(out_len), 'ms_per_step': 1e3 * dt_e2e / args.steps,
                        'handles': n_h},
                'gpu_launches': int(launches),
                'roofline': {'bound': 'hbm', 'kernel': 'k_mpileup_write', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                             'traffic': traffic, 'peak_source': peak_src, 'algorithmic_bytes_per_launch': int(alg),
                             'bytes_in': int(bytes_in), 'bytes_out': int(out_len), 'kernel_ms': write_ms,
                             'step_kernels_ms': {'read_stage(k_prep*,k_build_desc,k_overlap,ranges; includes its host syncs)': float(np.mean(stage_ms)),
                                                 'size_pass(k_ss_reads+k_ss_scan+k_ss_cols)': size_ms, 'k_scan_u32_to_u64': scan_ms, 'k_mpileup_write': write_ms,
                                                 'column_stage_total': kms, 'total': kms + float(np.mean(stage_ms))},
                             'step_frac': alg / ((kms + float(np.mean(stage_ms))) * 1e-3) / 1e9 / peak},
                'reads_per_step_per_gpu': n_reads}
        if world == 1 and not args.no_cpu_baseline:
            exe = oracle_path()
            with tempfile.TemporaryDirectory() as td:
                sam, nc = cpu_sample(td, 2.0)
                reps = 4
                t0 = time.perf_counter()
                for _ in range(reps):
                    subprocess.run([exe, 'mpileup', '-a', sam], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=True)
                cpu_dt = time.perf_counter() - t0
            line['cpu_baseline'] = {'value': reps * nc / cpu_dt, 'unit': UNIT, 'cores': 1, 'kind': 'port',
                                    'sample': f'{reps} x CPU oracle `mpileup -a` over a 2 Mb 30x/150bp SAM-text window (same generator), 1 thread'}
        print(json.dumps(line))
    for h in handles:
        h.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
