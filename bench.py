#!/usr/bin/env python
"""Benchmark of the B200-native CLEDB 2-line database search (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--mode iquv|iqud] [--impl reference]

One "step" = one pass of the hot path (cledb_match: classify, bucket, chi^2 search + top-nsearch, finalise) over one
synthetic observation map per GPU.  N=1 workload = BASELINE config 2: 256x256-pixel 2-line Fe XIII IQUV map,
one synthetic PyCELP-shaped database (340 200 entries, int16-encoded in HBM), nsearch 8.  For N>1 every rank
processes its own 256x256 map (weak scaling, database replicated) and the result maps are gathered with NCCL.

Prints ONE JSON line (rank 0).  `value` = pixels/s with inputs resident in HBM; `e2e` = the same through the
host-buffer C-ABI call (pinned host memory, H2D and D2H inside the timed region); `roofline` = the search kernel
against the FP32 FFMA peak measured in the same run; `cpu_baseline` = the oracle port (NumPy restatement of the
reference, fork pool) on a bounded pixel sample on this box's host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(REPO, 'solar-coronal-inversion_b200')
for p in (REPO, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

NSEARCH = 8


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--mode', default='iquv', choices=['iquv', 'iqud'])
    ap.add_argument('--nx', type=int, default=256)
    ap.add_argument('--encoding', type=int, default=1)
    ap.add_argument('--ndb', type=int, default=1, help='number of distinct (height, density) databases the map touches')
    ap.add_argument('--cpu-seconds', type=float, default=20.0)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--dbg', type=int, default=0, help='kernel timing experiment (results invalid)')
    ap.add_argument('--prune', action='store_true', help='(default) exact tile pruning, the library default')
    ap.add_argument('--no-prune', action='store_true', help='brute-force search as the measured path (the kernel the FP32 roofline is quoted on)')
    ap.add_argument('--no-prune-pass', action='store_true', help='skip the extra brute-force pass of the default run')
    ap.add_argument('--ppi', type=int, default=0, help='tuning: pixels per warp item (0 = library default)')
    ap.add_argument('--split', type=int, default=0, help='tuning: chunks per candidate list (0 = library default)')
    return ap.parse_args()


def workload_name(args):
    dbs = 'single synthetic PyCELP-shaped DB' if args.ndb == 1 else '%d synthetic PyCELP-shaped DBs (height x density grid points)' % args.ndb
    return ('2-line Fe XIII %s inversion, %dx%d synthetic map per GPU, %s '
            '(21x180x90 = 340200 entries each), nsearch %d' % (args.mode.upper(), args.nx, args.nx, dbs, NSEARCH))


def make_inputs(args, rank):
    """Returns (list of databases, observation dict).  With --ndb K the map touches K databases: rows of the map
    fall into height bands (as yobs does on a real map) and the density index varies inside a band."""
    from cledb_b200 import synth
    if args.ndb == 1:
        dbs = [synth.make_database(9, 40)]
        enc = None
    else:
        nh = max(1, int(round(args.ndb ** 0.5)))
        nd = (args.ndb + nh - 1) // nh
        grid = [(h, d) for h in range(nh) for d in range(nd)][:args.ndb]
        dbs = [synth.make_database(5 + 3 * h, 30 + 2 * d) for h, d in grid]
        rng = np.random.default_rng(1234 + rank)
        band = (np.arange(args.nx) * nh // args.nx)[:, None]                       # height band of a map row
        dens = rng.integers(0, nd, (args.nx, args.nx))
        enc = np.minimum(band * nd + dens, args.ndb - 1).astype(np.int32)
    obs = synth.make_observation(args.nx, args.nx, dbs, db_enc=enc, seed=7 + rank)
    return dbs, obs


# ---------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (NumPy restatement of cledb_invproc, validated bit-for-bit against the reference)
# ---------------------------------------------------------------------------------------------------------------
def cpu_run(args, db, obs, npixels, workers, seed):
    import oracle.cledb_oracle as orc
    from cledb_b200 import synth, codec
    dec = [codec.decoded_database(d, args.encoding).reshape(d.shape) for d in db]
    hdr = synth.make_dbhdr()
    nx = args.nx
    rng = np.random.default_rng(seed)
    flat = rng.choice(nx * nx, npixels, replace=False)
    pixels = [(int(f // nx), int(f % nx)) for f in flat]
    iss = obs['issuemask'].copy()
    t0 = time.perf_counter()
    orc.invert_map(obs['sobs_totrot'], obs['sobs_dopp'], dec, obs['db_enc'], obs['yobs'], obs['aobs'], obs['dobs'],
                   obs['snr'], iss, hdr, obs['keyvals'], NSEARCH, 1000, 3 if args.mode == 'iqud' else 0,
                   args.mode == 'iqud', workers=workers, pixels=pixels)
    return time.perf_counter() - t0


def cpu_workers():
    n = os.cpu_count() or 1
    return max(1, min(100, n - 2))          # the reference's pool size: min(ncpu, cpu_count()-2), ncpu = 100


def cpu_baseline(args, db, obs):
    workers = cpu_workers()
    n0 = 16 * workers                   # probe large enough that pool start-up does not dominate the per-pixel estimate
    t = cpu_run(args, db, obs, n0, workers, 0)
    n1 = int(max(n0, min(args.nx * args.nx, n0 * args.cpu_seconds / max(t, 1e-3))))
    t1 = cpu_run(args, db, obs, n1, workers, 1)
    return dict(value=n1 / t1, unit='pixels/s', cores=workers, kind='port',
                sample='%d of %d pixels of the same map (seeded subset), oracle/cledb_oracle.invert_map in a fork pool of '
                       '%d workers (reference pool size min(ncpu, cpu_count-2)), %.1f s' % (n1, args.nx * args.nx, workers, t1))


def run_reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    db, obs = make_inputs(args, 0)
    workers = cpu_workers()
    n = 32 * workers
    for w in range(args.warmup):
        cpu_run(args, db, obs, max(workers, n // 4), workers, 100 + w)
    times = [cpu_run(args, db, obs, n, workers, 200 + s) for s in range(args.steps)]
    t = float(np.mean(times))
    val = n / t
    line = dict(impl='reference', metric='2-line pixels/s', value=val, unit='pixels/s', n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=t * 1e3, higher_is_better=True, scaling='weak', vs_baseline=None,
                dtype='f64', data='synthetic',
                config=dict(workload=workload_name(args), sample_pixels_per_step=n,
                            note='reference is pure Python/NumPy and cannot travel to the GPU box; this arm times the '
                                 'oracle port of it (bit-identical outputs, tests/test_oracle_golden.py)'),
                cpu_baseline=dict(value=val, unit='pixels/s', cores=workers, kind='port',
                                  sample='%d seeded pixels per step of the %dx%d map' % (n, args.nx, args.nx)),
                e2e=dict(value=val, unit='pixels/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                evals_per_s=val * 340200)
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '50'], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(',')])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        try:
            sm = [float(r[0]) for r in self.rows if len(r) >= 7]
            if sm:
                out['sm_mhz'] = float(np.median(sm))
                out['sm_max_mhz'] = float(self.rows[0][1])
                out['power_w_max'] = max(float(r[2]) for r in self.rows if len(r) >= 7)
                names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
                out['reasons'] = [n for i, n in enumerate(names) if any(r[3 + i] == 'Active' for r in self.rows if len(r) >= 7)]
                out['samples'] = len(sm)
        except Exception:
            pass
        return out


def main():
    args = parse()
    if args.impl == 'reference':
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    from cledb_b200 import native
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    dev = torch.device('cuda', local)
    mode = native.MODE_IQUD if args.mode == 'iqud' else native.MODE_IQUV

    db, obs = make_inputs(args, rank)
    ctx = native.Context(local)
    pruned = not args.no_prune
    ctx.set_pruning(pruned)
    for i, d in enumerate(db):
        ctx.db_put(i, d, args.encoding)
    props = ctx.device_props()
    if args.ppi or args.split:
        ctx.set_tuning(args.ppi, args.split)
    if args.dbg:
        ctx.set_tuning(-args.dbg, 0)
    npix = args.nx * args.nx
    k = NSEARCH

    sobs_h = torch.from_numpy(obs['sobs_totrot'].reshape(-1, 8)).pin_memory()
    snr_h = torch.from_numpy(obs['snr'].reshape(-1, 8)).pin_memory()
    id_h = torch.from_numpy(np.ascontiguousarray(obs['db_enc'].reshape(-1).astype(np.int32))).pin_memory()
    sobs_d, snr_d, id_d = sobs_h.to(dev), snr_h.to(dev), id_h.to(dev)
    maps_d = torch.empty((2, npix, k), dtype=torch.int32, device=dev)      # idx and chi side by side: one collective gathers both
    idx_d = maps_d[0]
    chi_d = maps_d[1].view(torch.float32)
    kpos_d = torch.empty((npix, k), dtype=torch.int32, device=dev)
    rows_d = torch.empty((npix, k, 8), dtype=torch.float32, device=dev)
    st_d = torch.empty(npix, dtype=torch.uint8, device=dev)
    nc_d = torch.empty(npix, dtype=torch.int64, device=dev)
    gathered = torch.empty((world, 2, npix, k), dtype=torch.int32, device=dev) if world > 1 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)          # > 126 MB L2

    stream = torch.cuda.Stream(device=dev)      # an explicit stream: the library and the CUDA events share it
    torch.cuda.set_stream(stream)

    def step():
        ctx.match_dev(mode, npix, sobs_d.data_ptr(), snr_d.data_ptr(), id_d.data_ptr(), k, idx_d.data_ptr(), chi_d.data_ptr(),
                      None, kpos_d.data_ptr(), rows_d.data_ptr(), st_d.data_ptr(), nc_d.data_ptr(), stream.cuda_stream)
        if world > 1:                       # the only collective of the path: gather the result maps
            dist.all_gather_into_tensor(gathered, maps_d)

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    fp32_peak = ctx.measure_fp32_peak()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    def timed(nsteps):
        ctx.profile_begin()
        evs = []
        for _ in range(nsteps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            step()
            b.record(stream)
            evs.append((a, b))
        torch.cuda.synchronize()
        return [a.elapsed_time(b) for a, b in evs], ctx.profile_collect()

    launches0 = ctx.launch_count()
    step_ms, kern_ms = timed(args.steps)
    if world > 1:
        dist.barrier()
    launches = ctx.launch_count() - launches0
    prune_stats = ctx.pruning_stats() if pruned else None

    ms = float(np.mean(step_ms))
    ncand = int(nc_d.sum().item())
    nsearched = int((st_d == 0).sum().item())
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        c = torch.tensor([ncand, nsearched], device=dev, dtype=torch.int64)
        dist.all_reduce(c)
        ncand_all, nsearched_all = int(c[0].item()), int(c[1].item())
    else:
        ncand_all, nsearched_all = ncand, nsearched

    # ---- end to end through the host-buffer C-ABI call a reference binding makes (cledb_invert: search + solution
    #      building, pinned host memory in, the finished invout / sfound maps in pinned host memory out)
    from cledb_b200 import synth
    hdr = synth.make_dbhdr()
    out_h = dict(invout=torch.empty((npix, k, 11), dtype=torch.float32).pin_memory().numpy(),
                 sfound=torch.empty((npix, k, 8), dtype=torch.float32).pin_memory().numpy(),
                 status=torch.empty(npix, dtype=torch.uint8).pin_memory().numpy())
    sobs_n, snr_n, id_n = sobs_h.numpy(), snr_h.numpy(), id_h.numpy()
    yobs_n = torch.from_numpy(obs['yobs'].reshape(-1).copy()).pin_memory().numpy()
    dobs_n = torch.from_numpy(obs['dobs'].reshape(-1).copy()).pin_memory().numpy()
    aobs_n = obs['aobs'].reshape(-1)
    dopp_n = np.ascontiguousarray(obs['sobs_dopp'].reshape(npix, -1)[:, 0]) if args.mode == 'iqud' else None
    bcalc = 3 if args.mode == 'iqud' else 0

    def e2e_step():
        ctx.invert(mode, sobs_n, snr_n, id_n, k, yobs_n, dobs_n, aobs_n, dopp_n, hdr, 1000, bcalc, out=out_h)

    for _ in range(3):
        e2e_step()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    e2e_s = (time.perf_counter() - t0) / args.steps
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    # ---- a stream of maps through TWO contexts on this GPU (two host threads, each the same blocking cledb_invert call on its
    #      own pinned buffers): the result copy of one map overlaps the search of the next.  Extra information next to `e2e`
    #      (which is the single synchronous call a reference binding makes), N = 1 only.
    e2e_stream = None
    if world == 1 and not args.no_prune_pass:
        ctx2 = native.Context(local)
        ctx2.set_pruning(pruned)
        for i, d in enumerate(db):
            ctx2.db_put(i, d, args.encoding)
        out_h2 = dict(invout=torch.empty((npix, k, 11), dtype=torch.float32).pin_memory().numpy(),
                      sfound=torch.empty((npix, k, 8), dtype=torch.float32).pin_memory().numpy(),
                      status=torch.empty(npix, dtype=torch.uint8).pin_memory().numpy())
        nmaps = max(4, args.steps)

        def worker(c, o, n):
            for _ in range(n):
                c.invert(mode, sobs_n, snr_n, id_n, k, yobs_n, dobs_n, aobs_n, dopp_n, hdr, 1000, bcalc, out=o)

        worker(ctx2, out_h2, 3)
        ths = [threading.Thread(target=worker, args=(ctx, out_h, nmaps)), threading.Thread(target=worker, args=(ctx2, out_h2, nmaps))]
        t0 = time.perf_counter()
        for th in ths:
            th.start()
        for th in ths:
            th.join()
        e2e_stream = (time.perf_counter() - t0) / (2 * nmaps)
        same = bool(np.array_equal(out_h['invout'], out_h2['invout'], equal_nan=True) and np.array_equal(out_h['sfound'], out_h2['sfound'], equal_nan=True))
        ctx2.close()
    # ---- the same maps once more through the brute-force search (pruning off): the kernel that evaluates every
    #      (pixel, entry) pair, quoted against the FP32 peak
    brute = None
    if pruned and not args.no_prune_pass:
        for i in range(len(db)):
            ctx.db_drop(i)
        ctx.set_pruning(False)
        for i, d in enumerate(db):
            ctx.db_put(i, d, args.encoding)
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        b_step, b_kern = timed(min(args.steps, 10))
        brute = (float(np.mean(b_step)), float(np.mean(b_kern)) if len(b_kern) else float('nan'))
    clocks = sampler.stop() if rank == 0 else None
    h2d = npix * (32 + 32 + 4 + 4 + 4 + 4 + 4 + (4 if args.mode == 'iqud' else 0))
    d2h = npix * (k * (44 + 32) + 1)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    kms = float(np.mean(kern_ms)) if len(kern_ms) else float('nan')

    def traffic_of(name):                                 # DRAM bytes of the search kernel per launch, from the committed ncu capture
        if args.nx == 256 and args.ndb == 1 and args.encoding == 1:
            try:
                return json.load(open(os.path.join(REPO, 'profiles', 'k_match_traffic.json'))).get(name)
            except Exception:
                return None
        return None

    flops = 20.0 * ncand                                  # 10 FMA per candidate evaluation, this rank
    theo = props['sm_count'] * 128 * 2 * (clocks['sm_max_mhz'] or props['sm_clock_khz'] / 1e3) * 1e6 / 1e12
    peak = fp32_peak
    enc_name = 'int16' if args.encoding == 1 else 'float32'

    def roof(kernel, kernel_ms, step_mean_ms, traffic):
        achieved = flops / (kernel_ms * 1e-3) / 1e12
        return dict(bound='fp32', kernel=kernel, achieved=achieved, peak=peak, unit='TFLOP/s',
                    frac=achieved / peak if peak else None, traffic=traffic, kernel_ms=kernel_ms,
                    kernel_share_of_step=kernel_ms / step_mean_ms, flop_per_candidate=20, candidates_per_launch=ncand,
                    peak_source='FFMA microkernel measured in this run (cledb_measure_fp32_peak); theoretical %.1f TFLOP/s '
                                '= %d SMs x 128 lanes x 2 x max SM clock' % (theo, props['sm_count']))

    roofline_pruned = None
    if pruned:
        roofline_pruned = roof('k_match_pruned<%s>' % enc_name, kms, float(np.mean(step_ms)), traffic_of(args.mode + '_pruned'))
        pairs = prune_stats['evaluated'] + prune_stats['skipped']
        roofline_pruned['frac_algorithmic'] = roofline_pruned.pop('frac')
        roofline_pruned['frac'] = (20.0 * 256 * prune_stats['evaluated'] / (kms * 1e-3) / 1e12 / peak) if peak else None
        roofline_pruned['bound'] = 'instruction issue (selection and bound code); not FP32-bound'
        roofline_pruned['note'] = ('the kernel of the timed steps (value, e2e): exact tile bounds skip %.1f %% of the (pixel, tile) pairs. '
                                   'frac counts the FLOPs of the pairs it evaluates, frac_algorithmic the FLOPs of every (pixel, candidate) '
                                   'pair as in SURVEY 8d (> 1: work avoided, not a throughput claim)'
                                   % (100.0 * prune_stats['skipped'] / max(1, pairs)))
    if brute is not None:
        # the FP32-bound kernel of the path - the brute-force chi^2 search north_star's roofline target is set on - timed in
        # this run on the same maps with pruning off
        roofline = roof('k_match<%s,fp32>' % enc_name, brute[1], brute[0], traffic_of(args.mode))
        roofline['path'] = 'pruning off (cledb_set_pruning(ctx, 0)): %.3f ms per step, %.4g pixels/s; value / e2e above are measured with ' \
                           'pruning on, see roofline_pruned' % (brute[0], world * npix / (brute[0] * 1e-3))
        roofline['value_pruning_off'] = world * npix / (brute[0] * 1e-3)
    elif pruned:
        roofline = roofline_pruned
    else:
        roofline = roof('k_match<%s,fp32>' % enc_name, kms, float(np.mean(step_ms)), traffic_of(args.mode))
    line = dict(
        metric='2-line pixels/s', value=world * npix / (ms * 1e-3), unit='pixels/s', n_gpus=world, steps=args.steps,
        warmup=max(args.warmup, 3), ms_per_step=ms, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32',
        data='synthetic',
        config=dict(workload=workload_name(args), db_encoding='int16' if args.encoding == 1 else 'float32',
                    parallelism='pixel-sharded x%d, database replicated, NCCL all_gather of idx/chi maps' % world if world > 1 else 'single GPU',
                    search='exact tile pruning (library default)' if pruned else 'brute force (pruning off)',
                    l2='flushed between timed steps (256 MiB memset, outside the per-step CUDA events)',
                    timing='CUDA events per step on the launching stream, mean of K, max over ranks'),
        pruning=prune_stats,
        evals_per_s=world * npix * 340200 / (ms * 1e-3),
        candidate_evals_per_s=ncand_all / (ms * 1e-3),
        pixels_searched=nsearched_all,
        gpu_launches=int(launches),
        e2e=dict(value=world * npix / e2e_s, unit='pixels/s', h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                 api='cledb_invert (C-ABI, pinned host buffers in, finished invout/sfound maps out)', ms_per_step=e2e_s * 1e3,
                 stream_of_maps=None if e2e_stream is None else dict(
                     value=npix / e2e_stream, ms_per_map=e2e_stream * 1e3, identical_outputs=same,
                     how='two contexts on this GPU, one host thread each, the same blocking call on its own pinned buffers: '
                         'copies of one map overlap the search of the next')),
        roofline=roofline,
        clocks=clocks,
        device=dict(sm_count=props['sm_count'], match_ctas_per_sm=props['match_ctas_per_sm'], match_regs=props['match_regs'],
                    match_smem=props['match_smem']))
    if roofline_pruned is not None and roofline is not roofline_pruned:
        line['roofline_pruned'] = roofline_pruned
    if not args.no_cpu_baseline:
        line['cpu_baseline'] = cpu_baseline(args, db, obs)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
