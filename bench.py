#!/usr/bin/env python
"""Benchmark of the B200-native CLEDB 2-line database search (BASELINE.json metric and configs).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--configs 2,3,4,5]

Prints ONE JSON line (rank 0).  Top level = BASELINE config 2, the configuration the metric is quoted on: one 256x256-pixel
2-line Fe XIII IQUV map per GPU against one synthetic PyCELP-shaped database (340 200 entries, int16 in HBM), nsearch 8,
through the library's default search (exact tile pruning).  A "step" is one pass of the hot path over one map (cledb_match:
classify, bucket, chi^2 search + top-nsearch, finalise).  N > 1: every rank searches its own map (weak scaling, database
replicated) and the result maps are all-gathered over NVLink by the library (cledb_allgather_dev: its own peer-memory
stores where CUDA IPC works, else NCCL).

  value      pixels/s, inputs resident in HBM, CUDA events on the launching stream, L2 flushed between steps, max over ranks
  e2e        the same through the host-buffer C-ABI call (cledb_invert: H2D of the inputs, search, solutions, issue flags,
             D2H of the finished invout / sfound maps inside the timed region); e2e.python = CLEDB_PROC.cledb_invproc itself
  roofline   the kernel `value` timed (k_match_pruned) against the FP32 peak measured in the same run
  by_config  one record per BASELINE config, each with its own value / e2e / roofline / cpu_baseline:
               config2_iquv_256_pruned (= top level), config2_iquv_256_bruteforce (the FP32-roofline record: every (pixel,
               candidate) pair evaluated), config3_iqud_256_*, config4_blos_2048 / config4_qurotate_2048* (HBM roofline),
               config5_iquv_1024_multidb (ONE 1024x1024 map over several hundred databases, pixel-sharded over the N ranks
               through cledb_invert_dev + one all-gather: strong scaling; at N = 1 the whole map on one GPU)
  cpu_baseline  the oracle port (NumPy restatement of the reference, fork pool) on a bounded pixel sample, this box's cores (N = 1 only)
"""
import argparse
import ctypes
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
NPOOL = 8          # observation maps the top-level record cycles through
NDB_ROWS = 340200


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--mode', default='iquv', choices=['iquv', 'iqud'], help='mode of the top-level record')
    ap.add_argument('--nx', type=int, default=256)
    ap.add_argument('--encoding', type=int, default=1)
    ap.add_argument('--ndb', type=int, default=1, help='databases the top-level map touches')
    ap.add_argument('--ndb5', type=int, default=247, help='databases of the config-5 map (the reference notebook loads 247 for a 1024^2 map)')
    ap.add_argument('--nx5', type=int, default=1024)
    ap.add_argument('--configs', default='', help='comma list of extra configs to run (default: 2b,3,4,5 at N=1; 2b,5 at N>1)')
    ap.add_argument('--transport', default='auto', choices=['auto', 'peer', 'nccl'],
                    help='multi-GPU gathers: peer = the library\'s own stores into the peers\' windows (default where CUDA IPC works), nccl = ncclAllGather')
    ap.add_argument('--push-blocks', type=int, default=12, help='N > 1: blocks of the gather kernel when it runs beside the next search')
    ap.add_argument('--pieces', type=int, default=0, help='config 5: pieces of a shard in the pipelined gather (0 = library default)')
    ap.add_argument('--no-overlap', action='store_true', help='N > 1: time search + gather back to back on one stream (no pipelining of the gather)')
    ap.add_argument('--cpu-seconds', type=float, default=15.0)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--dbg', type=int, default=0, help='kernel timing experiment (results invalid)')
    ap.add_argument('--no-prune', action='store_true', help='brute-force search as the top-level path')
    ap.add_argument('--ppi', type=int, default=0, help='tuning: pixels per warp item (0 = library default)')
    ap.add_argument('--split', type=int, default=0, help='tuning: chunks per candidate list (0 = library default)')
    return ap.parse_args()


def workload_name(mode, nx, ndb):
    dbs = 'single synthetic PyCELP-shaped DB' if ndb == 1 else '%d synthetic PyCELP-shaped DBs (height x density grid points)' % ndb
    return ('2-line Fe XIII %s inversion, %dx%d synthetic map per GPU, %s (21x180x90 = 340200 entries each), nsearch %d'
            % (mode.upper(), nx, nx, dbs, NSEARCH))


def make_inputs(nx, ndb, rank):
    """(list of databases, observation dict).  With ndb = K the map touches K databases: rows of the map fall into height
    bands (as yobs does on a real map) and the density index varies inside a band."""
    from cledb_b200 import synth
    if ndb == 1:
        dbs, enc = [synth.make_database(9, 40)], None
    else:
        nh = max(1, int(round(ndb ** 0.5)))
        nd = (ndb + nh - 1) // nh
        grid = [(h, d) for h in range(nh) for d in range(nd)][:ndb]
        dbs = [synth.make_database(5 + 3 * h, 30 + 2 * d) for h, d in grid]
        rng = np.random.default_rng(1234 + rank)
        band = (np.arange(nx) * nh // nx)[:, None]
        dens = rng.integers(0, nd, (nx, nx))
        enc = np.minimum(band * nd + dens, ndb - 1).astype(np.int32)
    return dbs, synth.make_observation(nx, nx, dbs, db_enc=enc, seed=7 + rank)


# ---------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (NumPy restatement of cledb_invproc, validated bit-for-bit against the reference)
# ---------------------------------------------------------------------------------------------------------------
def cpu_run(mode, encoding, db, obs, npixels, workers, seed):
    import oracle.cledb_oracle as orc
    from cledb_b200 import synth, codec
    dec = [codec.decoded_database(d, encoding).reshape(d.shape) for d in db]
    hdr = synth.make_dbhdr()
    nx, ny = obs['sobs_totrot'].shape[:2]
    flat = np.random.default_rng(seed).choice(nx * ny, npixels, replace=False)
    pixels = [(int(f // ny), int(f % ny)) for f in flat]
    iss = obs['issuemask'].copy()
    t0 = time.perf_counter()
    orc.invert_map(obs['sobs_totrot'], obs['sobs_dopp'], dec, obs['db_enc'], obs['yobs'], obs['aobs'], obs['dobs'],
                   obs['snr'], iss, hdr, obs['keyvals'], NSEARCH, 1000, 3 if mode == 'iqud' else 0, mode == 'iqud',
                   workers=workers, pixels=pixels)
    return time.perf_counter() - t0


def cpu_workers():
    n = os.cpu_count() or 1
    return max(1, min(100, n - 2))          # the reference's pool size: min(ncpu, cpu_count()-2), ncpu = 100


def cpu_baseline(mode, encoding, db, obs, seconds):
    workers = cpu_workers()
    n0 = 16 * workers                   # probe large enough that pool start-up does not dominate the per-pixel estimate
    t = cpu_run(mode, encoding, db, obs, n0, workers, 0)
    npix = obs['sobs_totrot'].shape[0] * obs['sobs_totrot'].shape[1]
    n1 = int(max(2048, min(npix, n0 * seconds / max(t, 1e-3))))
    n1 = min(n1, npix)
    t1 = cpu_run(mode, encoding, db, obs, n1, workers, 1)
    return dict(value=n1 / t1, unit='pixels/s', cores=workers, kind='port',
                sample='%d of %d pixels of the same map (seeded subset), oracle/cledb_oracle.invert_map in a fork pool of '
                       '%d workers (reference pool size min(ncpu, cpu_count-2)), %.1f s' % (n1, npix, workers, t1))


def run_reference_arm(args):
    if int(os.environ.get('RANK', '0')) != 0:
        return
    db, obs = make_inputs(args.nx, args.ndb, 0)
    workers = cpu_workers()
    n = min(args.nx * args.nx, max(2048, 32 * workers))          # a seeded subset of >= 2048 pixels
    for w in range(min(args.warmup, 1)):
        cpu_run(args.mode, args.encoding, db, obs, max(workers, n // 4), workers, 100 + w)
    steps = max(1, min(args.steps, 12))                          # bounded: each step is ~2-4 s of all host cores
    times = [cpu_run(args.mode, args.encoding, db, obs, n, workers, 200 + s) for s in range(steps)]
    t = float(np.mean(times))
    val = n / t
    line = dict(impl='reference', metric='2-line pixels/s', value=val, unit='pixels/s', n_gpus=args.gpus, steps=steps,
                warmup=min(args.warmup, 1), ms_per_step=t * 1e3, higher_is_better=True, scaling='weak', vs_baseline=None,
                dtype='f64', data='synthetic',
                config=dict(workload=workload_name(args.mode, args.nx, args.ndb), sample_pixels_per_step=n,
                            note='reference is pure Python/NumPy and cannot travel to the GPU box; this arm times the oracle port of '
                                 'it (bit-identical outputs, tests/test_oracle_golden.py) on all host cores; steps capped at 12'),
                cpu_baseline=dict(value=val, unit='pixels/s', cores=workers, kind='port',
                                  sample='%d seeded pixels per step of the %dx%d map' % (n, args.nx, args.nx)),
                e2e=dict(value=val, unit='pixels/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                scanned_evals_per_s=val * NDB_ROWS)
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons while the timed regions run."""
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

    def mark(self):
        return len(self.rows)

    def summary(self, lo=0, hi=None):
        rows = [r for r in self.rows[lo:hi] if len(r) >= 7]
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        try:
            sm = [float(r[0]) for r in rows]
            if sm:
                out['sm_mhz'] = float(np.median(sm))
                out['sm_max_mhz'] = float(rows[0][1])
                out['power_w_max'] = max(float(r[2]) for r in rows)
                names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
                out['reasons'] = [n for i, n in enumerate(names) if any(r[3 + i] == 'Active' for r in rows)]
                out['samples'] = len(sm)
        except Exception:
            pass
        return out

    def stop(self):
        if self.proc:
            self.proc.terminate()


class _DeviceBytes:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = dict(shape=(int(nbytes),), typestr='|u1', data=(int(ptr), False), version=2)


class Bench:
    """Shared state of one rank: device, stream, context, timing helpers."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        from cledb_b200 import native
        self.torch, self.dist, self.native, self.args = torch, dist, native, args
        self.rank = int(os.environ.get('RANK', '0'))
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        self.local = int(os.environ.get('LOCAL_RANK', '0'))
        torch.cuda.set_device(self.local)
        self.dev = torch.device('cuda', self.local)
        if self.world > 1:
            dist.init_process_group('nccl', device_id=self.dev)
        self.ctx = native.Context(self.local)
        ids = [None]
        if self.world > 1:                                   # the library's own communicator (NCCL bound at run time)
            ids = [native.comm_unique_id() if self.rank == 0 else None]
            dist.broadcast_object_list(ids, src=0)
        self.ctx.comm_init(self.rank, self.world, ids[0])    # one rank: no NCCL involved
        if args.transport != 'auto':
            self.ctx.comm_transport(args.transport)
        self.transport = self.ctx.comm_transport()           # 'peer': the library's own stores into the peers' windows; 'nccl'
        self.stream = torch.cuda.Stream(device=self.dev)     # an explicit stream: the library and the CUDA events share it
        torch.cuda.set_stream(self.stream)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)          # > 126 MB L2
        self.props = self.ctx.device_props()
        self.fp32_peak = None
        self.sampler = ClockSampler(self.local)
        if self.rank == 0:
            self.sampler.start()
            time.sleep(0.3)
        try:
            self.hbm_peak = json.load(open(os.path.join(REPO, 'MEASURED_PEAKS.json')))['hbm_gbs']
            self.hbm_src = 'MEASURED_PEAKS.json hbm_gbs (burst copy bandwidth)'
        except Exception:
            self.hbm_peak, self.hbm_src = 6553.9, 'fallback 6553.9 GB/s (B200_PROFILING.md)'

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, nsteps, profile=True):
        """K steps, each between two CUDA events on the launching stream, L2 flushed before each step (outside the events).
        Returns (per-step ms, per-launch ms of the search kernel from the library's own events)."""
        torch = self.torch
        if profile:
            self.ctx.profile_begin()
        evs = []
        for _ in range(nsteps):
            self.flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(self.stream)
            fn()
            b.record(self.stream)
            evs.append((a, b))
        torch.cuda.synchronize()
        return [a.elapsed_time(b) for a, b in evs], (self.ctx.profile_collect() if profile else None)

    def timed_pipelined(self, search, gather, nsteps, side):
        """K maps as a software pipeline: the gather of map i-1 (side stream) runs beside the search of map i (launching stream).
        K + 1 brackets of CUDA events on the launching stream, L2 flushed before each (outside the events); bracket i opens, the
        side stream waits for the opening event and gathers map i-1, the launching stream searches map i and joins the side
        stream before the bracket closes - so every search and every gather lies inside exactly one bracket (the last bracket
        holds the last gather alone) and nothing hides behind the untimed flush.  Returns ([sum of brackets / K], kernel ms)."""
        torch = self.torch
        self.ctx.profile_begin()
        evs = []
        for i in range(nsteps + 1):
            self.flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(self.stream)
            join = None
            if i > 0:
                side.wait_event(a)
                gather((i - 1) % 2, side.cuda_stream)
                join = torch.cuda.Event()
                join.record(side)
            if i < nsteps:
                search(i % 2)
            if join is not None:
                self.stream.wait_event(join)
            b.record(self.stream)
            evs.append((a, b))
        torch.cuda.synchronize()
        return [sum(a.elapsed_time(b) for a, b in evs) / nsteps], self.ctx.profile_collect()

    def max_over_ranks(self, x):
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([float(x)], device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def all_ranks(self, x):
        if self.world == 1:
            return [float(x)]
        t = self.torch.zeros(self.world, device=self.dev, dtype=self.torch.float64)
        t[self.rank] = float(x)
        self.dist.all_reduce(t)
        return [round(float(v), 5) for v in t.tolist()]

    def sum_over_ranks(self, xs):
        if self.world == 1:
            return [int(x) for x in xs]
        t = self.torch.tensor([int(x) for x in xs], device=self.dev, dtype=self.torch.int64)
        self.dist.all_reduce(t)
        return [int(v) for v in t.tolist()]

    def pinned(self, a):
        return self.torch.from_numpy(np.ascontiguousarray(a)).pin_memory()

    def view(self, ptr, nbytes):
        """uint8 tensor over device memory the library owns (the receive window)."""
        return self.torch.as_tensor(_DeviceBytes(ptr, nbytes), device=self.dev)

    def traffic_of(self, name):
        """DRAM bytes per launch of a kernel from the committed ncu --set full capture (profiles/kernel_traffic.json)."""
        try:
            return json.load(open(os.path.join(REPO, 'profiles', 'kernel_traffic.json'))).get(name)
        except Exception:
            return None


# ---------------------------------------------------------------------------------------------------------------
# configs 2 and 3: the 256x256 map against one database, both searches
# ---------------------------------------------------------------------------------------------------------------
def bench_search(B, mode_name, nx, ndb, encoding, steps, warmup, with_gather, want_e2e=True, want_python=True, want_stream=True):
    """Returns {'pruned': record, 'bruteforce': record} for one mode on the nx x nx map."""
    torch, native, ctx = B.torch, B.native, B.ctx
    mode = native.MODE_IQUD if mode_name == 'iqud' else native.MODE_IQUV
    db, obs = make_inputs(nx, ndb, B.rank)
    npix, k = nx * nx, NSEARCH
    sobs_h, snr_h = B.pinned(obs['sobs_totrot'].reshape(-1, 8)), B.pinned(obs['snr'].reshape(-1, 8))
    id_h = B.pinned(obs['db_enc'].reshape(-1).astype(np.int32))
    sobs_d, snr_d, id_d = sobs_h.to(B.dev), snr_h.to(B.dev), id_h.to(B.dev)
    maps_d = torch.empty((2, npix, k), dtype=torch.int32, device=B.dev)      # idx and chi side by side: one collective gathers both
    kpos_d = torch.empty((npix, k), dtype=torch.int32, device=B.dev)
    rows_d = torch.empty((npix, k, 8), dtype=torch.float32, device=B.dev)
    st_d = torch.empty(npix, dtype=torch.uint8, device=B.dev)
    nc_d = torch.empty(npix, dtype=torch.int64, device=B.dev)
    # the gathered maps live in the library's receive window: with the peer transport every rank stores its block straight
    # into every rank's window (k_push_block between the free / done flags), with the NCCL one it is a plain receive buffer
    # Two sets of result maps and two receive regions: in the timed pipeline the gather of one map runs beside the search of the next.
    map_bytes = maps_d.numel() * 4
    gathered = ctx.comm_window(2 * B.world * map_bytes) if with_gather and B.world > 1 else None
    maps2 = [maps_d, torch.empty_like(maps_d)] if gathered is not None else [maps_d]
    side = torch.cuda.Stream(device=B.dev) if gathered is not None else None

    # The top-level record walks a pool of NPOOL observation maps (seeds 7 .. 7+NPOOL-1, the same pool at every N), rank r starting
    # at map r: the pruned search's time depends on the map (its longest work item), so every rank - and the N = 1 run - sees the
    # same mix of maps and the per-GPU work is the same at every N.
    pool = [(sobs_d, snr_d, id_d)]
    if with_gather:
        for m in range(1, NPOOL):
            o = make_inputs(nx, ndb, (B.rank + m) % NPOOL)[1]
            pool.append(tuple(torch.from_numpy(np.ascontiguousarray(a)).to(B.dev) for a in
                              (o['sobs_totrot'].reshape(-1, 8), o['snr'].reshape(-1, 8), o['db_enc'].reshape(-1).astype(np.int32))))
    cursor = [0]

    def search(b=0):
        s_d, n_d, i_d = pool[cursor[0] % len(pool)]
        cursor[0] += 1
        ctx.match_dev(mode, npix, s_d.data_ptr(), n_d.data_ptr(), i_d.data_ptr(), k, maps2[b][0].data_ptr(), maps2[b][1].data_ptr(),
                      None, kpos_d.data_ptr(), rows_d.data_ptr(), st_d.data_ptr(), nc_d.data_ptr(), B.stream.cuda_stream)

    def gather(b, stream):                  # the only collective of the path: gather the result maps (library call)
        ctx.allgather_dev(maps2[b].data_ptr(), gathered + b * B.world * map_bytes, map_bytes, stream)

    def step():                             # search, then the gather on the same stream
        search(0)
        if gathered is not None:
            gather(0, B.stream.cuda_stream)

    from cledb_b200 import synth
    hdr = synth.make_dbhdr()
    out_h = dict(invout=torch.empty((npix, k, 11), dtype=torch.float32).pin_memory().numpy(),
                 sfound=torch.empty((npix, k, 8), dtype=torch.float32).pin_memory().numpy(),
                 status=torch.empty(npix, dtype=torch.uint8).pin_memory().numpy(),
                 issue=torch.empty(npix, dtype=torch.int32).pin_memory().numpy())
    sobs_n, snr_n, id_n = sobs_h.numpy(), snr_h.numpy(), id_h.numpy()
    yobs_n, dobs_n = B.pinned(obs['yobs'].reshape(-1)).numpy(), B.pinned(obs['dobs'].reshape(-1)).numpy()
    aobs_n = obs['aobs'].reshape(-1)
    dopp_n = np.ascontiguousarray(obs['sobs_dopp'].reshape(npix, -1)[:, 0]) if mode_name == 'iqud' else None
    bcalc = 3 if mode_name == 'iqud' else 0

    def e2e_step(c=ctx, o=out_h):
        c.invert(mode, sobs_n, snr_n, id_n, k, yobs_n, dobs_n, aobs_n, dopp_n, hdr, 1000, bcalc, out=o)

    h2d = npix * (32 + 32 + 4 + 4 + 4 + 4 + 4 + (4 if mode_name == 'iqud' else 0))
    d2h = npix * (k * (44 + 32) + 1 + 4)
    enc_name = 'int16' if encoding == 1 else 'float32'
    recs = {}
    for variant in (1, 2, 0):                                # exact tile pruning; seeded brute force (same k-d layout); plain brute force
        pruned = variant == 1
        ctx.set_pruning(variant)
        if variant != 2:
            for i, d in enumerate(db):
                ctx.db_put(i, d, encoding)
        if B.args.ppi or B.args.split:
            ctx.set_tuning(B.args.ppi, B.args.split)
        if B.args.dbg:
            ctx.set_tuning(-B.args.dbg, 0)
        for _ in range(max(warmup, 3)):
            step()
        torch.cuda.synchronize()
        if B.fp32_peak is None:
            B.fp32_peak = ctx.measure_fp32_peak()
        B.barrier()
        mark = B.sampler.mark()
        launches0 = ctx.launch_count()
        nst = steps if pruned else max(3, min(steps, 10 if variant == 2 else 5))
        serial_ms = None
        cursor[0] = 0
        if pruned and gathered is not None and not B.args.no_overlap:
            s_ms, _ = B.timed(step, nst)
            serial_ms = B.max_over_ranks(float(np.mean(s_ms)))
            B.barrier()
            launches0 = ctx.launch_count()
            cursor[0] = 0
            if B.transport == 'peer':
                ctx.comm_push_blocks(B.args.push_blocks)    # the gather runs beside a search: it leaves most SMs to the search
            step_ms, kern_ms = B.timed_pipelined(search, gather, nst, side)
            if B.transport == 'peer':
                ctx.comm_push_blocks(0)
        else:
            step_ms, kern_ms = B.timed(step, nst)
        cursor[0] = 0
        B.barrier()
        launches = ctx.launch_count() - launches0
        stats = ctx.pruning_stats() if pruned else None
        if pruned and len(pool) > 1:                        # (pixel, tile) pairs of the average map of the pool
            acc = {}
            for m in range(len(pool)):
                cursor[0] = m
                search(0)
                torch.cuda.synchronize()
                for kk, v in ctx.pruning_stats().items():
                    acc[kk] = acc.get(kk, 0) + v
            stats = {kk: int(round(v / len(pool))) for kk, v in acc.items()}
            cursor[0] = 0
            search(0)                                       # leave the results of this rank's first map in the buffers
            torch.cuda.synchronize()
        ms = B.max_over_ranks(float(np.mean(step_ms)))
        kms = float(np.mean(kern_ms)) if len(kern_ms) else float('nan')
        ncand, nsearched = int(nc_d.sum().item()), int((st_d == 0).sum().item())
        ncand_all, nsearched_all = B.sum_over_ranks([ncand, nsearched])
        peak = B.fp32_peak
        clocks = B.sampler.summary(mark) if B.rank == 0 else None
        theo = B.props['sm_count'] * 128 * 2 * ((clocks or {}).get('sm_max_mhz') or B.props['sm_clock_khz'] / 1e3) * 1e6 / 1e12
        peak_src = ('FFMA microkernel measured in this run (cledb_measure_fp32_peak); theoretical %.1f TFLOP/s = %d SMs x 128 lanes '
                    'x 2 x max SM clock.  MEASURED_PEAKS.json holds HBM and bf16 tensor peaks, neither bounds this kernel'
                    % (theo, B.props['sm_count']))
        if pruned:
            evaluated = 256.0 * stats['evaluated']                      # entries actually evaluated (pairs x 256-entry tiles)
            achieved = 20.0 * evaluated / (kms * 1e-3) / 1e12
            pairs = stats['evaluated'] + stats['skipped']
            roof = dict(bound='fp32 (instruction-issue limited: selection and bound code)', kernel='k_match_pruned<%s>' % enc_name,
                        achieved=achieved, peak=peak, unit='TFLOP/s', frac=achieved / peak if peak else None,
                        traffic=B.traffic_of('k_match_pruned_' + mode_name) if (nx, ndb, encoding) == (256, 1, 1) else None, kernel_ms=kms, kernel_share_of_step=kms / float(np.mean(step_ms)),
                        flop_per_evaluated_entry=20, evaluated_entries_per_launch=evaluated,
                        frac_if_every_candidate_counted=20.0 * ncand / (kms * 1e-3) / 1e12 / peak if peak else None,
                        note='FLOPs of the entries the kernel really evaluates: exact tile bounds skip %.1f %% of the (pixel, tile) pairs. '
                             'frac_if_every_candidate_counted credits the skipped candidates too (SURVEY 8d algorithmic count; > 1 means '
                             'work avoided, not throughput)' % (100.0 * stats['skipped'] / max(1, pairs)),
                        peak_source=peak_src)
            evals = evaluated * B.world
        else:
            achieved = 20.0 * ncand / (kms * 1e-3) / 1e12
            roof = dict(bound='fp32', kernel='k_match<%s,fp32>%s' % (enc_name, ' seeded, k-d tile layout' if variant == 2 else ' plain layout'),
                        achieved=achieved, peak=peak, unit='TFLOP/s',
                        frac=achieved / peak if peak else None,
                        traffic=B.traffic_of('k_match_%s_%s' % ('seeded' if variant == 2 else 'plain', mode_name)) if (nx, ndb, encoding) == (256, 1, 1) else None,
                        kernel_ms=kms,
                        kernel_share_of_step=kms / float(np.mean(step_ms)), flop_per_candidate=20, candidates_per_launch=ncand,
                        peak_source=peak_src)
            evals = ncand_all
        rec = dict(value=B.world * npix / (ms * 1e-3), unit='pixels/s', ms_per_step=ms, steps=nst,
                   workload=workload_name(mode_name, nx, ndb), search={1: 'exact tile pruning (library default, cledb_set_pruning 1)',
                                                                      2: 'brute force, every (pixel, entry) pair evaluated; k-d tile layout, every pixel started from its seed tile (cledb_set_pruning 2)',
                                                                      0: 'brute force, every (pixel, entry) pair evaluated; plain stride-permuted layout (cledb_set_pruning 0)'}[variant],
                   evals_per_s=evals / (ms * 1e-3), evals_definition='(pixel, database entry) chi^2 evaluations actually executed per second',
                   nominal_scanned_evals_per_s=B.world * npix * float(NDB_ROWS) * max(1, ndb) / max(1, ndb) / (ms * 1e-3),
                   nominal_candidate_evals_per_s=ncand_all / (ms * 1e-3), pixels_searched=nsearched_all, gpu_launches=int(launches),
                   pruning=stats, roofline=roof, clocks=clocks)
        if B.world > 1:
            rec['per_rank'] = dict(step_ms=B.all_ranks(float(np.mean(step_ms))), search_kernel_ms=B.all_ranks(kms),
                                   note='mean step and mean search-kernel time of every rank (value uses the slowest rank)')
        if len(pool) > 1:
            rec['map_pool'] = ('%d observation maps (seeds 7..%d) on the same database, rank r starts at map r and takes the next one each step: '
                               'the same mix of maps on every rank and at every N' % (len(pool), 6 + len(pool)))
        if serial_ms is not None:
            rec['gather_overlap'] = dict(ms_per_step_serial=serial_ms, value_serial=B.world * npix / (serial_ms * 1e-3),
                                         how='ms_per_step: the gather of map i-1 runs on a second stream beside the search of map i (two sets of '
                                             'result maps, two receive regions; every search and every gather lies inside one timed bracket, '
                                             'the last gather in a bracket of its own).  ms_per_step_serial: search and gather back to back on '
                                             'one stream.')
        if want_e2e and variant != 0:
            for _ in range(3):
                e2e_step()
            B.barrier()
            t0 = time.perf_counter()
            for _ in range(nst):
                e2e_step()
            e2e_s = B.max_over_ranks((time.perf_counter() - t0) / nst)
            rec['e2e'] = dict(value=B.world * npix / e2e_s, unit='pixels/s', h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                              ms_per_step=e2e_s * 1e3,
                              api='cledb_invert (C-ABI): pinned host buffers in, search + solutions + issue flags on the device, finished '
                                  'invout / sfound maps out')
        recs[{1: 'pruned', 2: 'bruteforce', 0: 'bruteforce_plain'}[variant]] = rec
        if pruned and want_stream and B.world == 1:
            # a stream of maps through TWO contexts on this GPU (two host threads, the same blocking call): the result copy of
            # one map overlaps the search of the next
            ctx2 = native.Context(B.local)
            ctx2.set_pruning(True)
            for i, d in enumerate(db):
                ctx2.db_put(i, d, encoding)
            out_h2 = {kk: torch.empty(v.shape, dtype=torch.from_numpy(v).dtype).pin_memory().numpy() for kk, v in out_h.items()}
            nmaps = max(4, nst)

            def worker(c, o, n):
                for _ in range(n):
                    e2e_step(c, o)

            worker(ctx2, out_h2, 3)
            ths = [threading.Thread(target=worker, args=(ctx, out_h, nmaps)), threading.Thread(target=worker, args=(ctx2, out_h2, nmaps))]
            t0 = time.perf_counter()
            for th in ths:
                th.start()
            for th in ths:
                th.join()
            es = (time.perf_counter() - t0) / (2 * nmaps)
            same = bool(np.array_equal(out_h['invout'], out_h2['invout'], equal_nan=True) and np.array_equal(out_h['sfound'], out_h2['sfound'], equal_nan=True))
            rec['e2e']['stream_of_maps'] = dict(value=npix / es, ms_per_map=es * 1e3, identical_outputs=same,
                                                how='two contexts on this GPU, one host thread each, the same blocking call')
            ctx2.close()
        if variant != 1:
            for i in range(len(db)):
                ctx.db_drop(i)
    ctx.set_pruning(1)
    ctx.set_tuning(0, 0)
    if want_python:
        # the reference-shaped Python call itself (CLEDB_PROC.cledb_invproc): same map, same encoding, its own context; the
        # database is resident after the first call, results land in the binding's pinned pool
        import CLEDB_PROC.CLEDB_PROC as procinv
        procinv.OPTIONS.update(device=B.local, dbencoding=encoding, precision=0, solution='device', pruning=True)
        pargs = (obs['sobs_totrot'], obs['sobs_dopp'] if mode_name == 'iqud' else 0, db, obs['db_enc'], obs['yobs'], obs['aobs'],
                 obs['dobs'], obs['snr'])
        ptail = (hdr, obs['keyvals'], k, 1000, bcalc, mode_name == 'iqud', 1, False, 0)
        keep = None
        for _ in range(4):                                   # the binding's pinned pool reaches its steady state: two sets of result
            keep = procinv.cledb_invproc(*pargs, obs['issuemask'].copy(), *ptail)      # arrays alive at a time, as in the loop below
        B.barrier()
        nst = max(3, min(steps, 20))
        iss = obs['issuemask'].copy()
        per_call = []
        for _ in range(nst):
            t0 = time.perf_counter()
            inv, sf = procinv.cledb_invproc(*pargs, iss, *ptail)
            per_call.append(time.perf_counter() - t0)
        del keep
        py_s = B.max_over_ranks(float(np.mean(per_call)))
        agree = bool(np.array_equal(inv.reshape(npix, k, 11), recs['pruned']['e2e'] and out_h['invout'], equal_nan=True)) if want_e2e else None
        recs['pruned'].setdefault('e2e', {})['python'] = dict(
            value=B.world * npix / py_s, unit='pixels/s', ms_per_step=py_s * 1e3, same_invout_as_c_abi=agree,
            ms_per_call_median=float(np.median(per_call)) * 1e3, ms_per_call_min=float(np.min(per_call)) * 1e3, ms_per_call_max=float(np.max(per_call)) * 1e3,
            api='CLEDB_PROC.cledb_invproc(...) - the reference-shaped call: pageable NumPy inputs, fresh result arrays from the pinned pool, '
                'issuemask updated in place')
        del inv, sf
        procinv.release_databases()
    recs['_inputs'] = (db, obs)
    return recs


# ---------------------------------------------------------------------------------------------------------------
# config 4: the elementwise neighbours on a 2048 x 2048 map (HBM roofline)
# ---------------------------------------------------------------------------------------------------------------
def bench_elementwise(B, nx, steps, cpu=True):
    """Both kernels are far shorter than the gap between two host-issued launches is stable (23 - 55 us), so a timed step is a
    batch of NB launches between one pair of CUDA events, every launch on its own input / output buffers: NSET buffer sets,
    0.6 - 1.3 GB in total, cycled through - no launch finds its data in the 126 MB L2, no flush needed between them."""
    torch, ctx = B.torch, B.ctx
    from cledb_b200 import synth
    import constants as consts
    import oracle.cledb_oracle as orc
    npix = nx * nx
    out = {}
    NSET, NB = 4, 8

    def timed(fn_of_set):
        for i in range(NSET):
            fn_of_set(i)
        torch.cuda.synchronize()

        def batch():
            for i in range(NB):
                fn_of_set(i % NSET)
        ms, _ = B.timed(batch, steps, profile=False)
        return float(np.mean(ms)) / NB

    def record(kernel, metric, workload, ms, bpp, e2e_s, h2d, d2h, cpu_rec):
        gbs = bpp * npix / (ms * 1e-3) / 1e9
        tr = B.traffic_of(kernel + ('_%d' % nx))
        return dict(metric=metric, value=npix / (ms * 1e-3), unit='pixels/s', ms_per_step=ms, steps=steps * NB, workload=workload,
                    timing='%d launches per CUDA-event pair over %d rotating buffer sets (working set %.0f MB, larger than L2)'
                           % (NB, NSET, NSET * bpp * npix / 1e6),
                    roofline=dict(bound='hbm', kernel=kernel, achieved=gbs, peak=B.hbm_peak, unit='GB/s', frac=gbs / B.hbm_peak, traffic=tr,
                                  bytes_per_pixel=bpp, algorithmic_bytes_per_launch=bpp * npix, peak_source=B.hbm_src),
                    e2e=dict(value=npix / e2e_s, unit='pixels/s', h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h, ms_per_step=e2e_s * 1e3),
                    cpu_baseline=cpu_rec)

    one = synth.make_oneline_map(nx, nx)
    c = consts.Constants('fe-xiii_1074')
    kfac = 1.e9 * (c.planckconst * c.l_speed / c.bohrmagneton / (c.line_ref ** 2))
    d_in = [torch.from_numpy(one.reshape(-1, 4)).to(B.dev) for _ in range(NSET)]
    d_out = [torch.empty_like(d_in[0]) for _ in range(NSET)]
    d_fl = [torch.empty((npix, 2), dtype=torch.uint8, device=B.dev) for _ in range(NSET)]
    ms = timed(lambda i: ctx.blos_dev(npix, d_in[i].data_ptr(), kfac, c.g_eff, c.F_factor, d_out[i].data_ptr(), d_fl[i].data_ptr(),
                                      B.stream.cuda_stream))
    one_p = B.pinned(one.reshape(-1, 4)).numpy()
    ctx.blos(one_p, kfac, c.g_eff, c.F_factor)
    t0 = time.perf_counter()
    for _ in range(3):
        out_h, fl_h = ctx.blos(one_p, kfac, c.g_eff, c.F_factor)
    e2e = (time.perf_counter() - t0) / 3
    assert np.array_equal(out_h, d_out[0].cpu().numpy(), equal_nan=True)
    cpu_rec = None
    if cpu:
        ns = 8192
        t0 = time.perf_counter()
        ref = np.stack([orc.blos_pixel(one.reshape(-1, 4)[i], c)[0] for i in range(ns)])
        cpu_rec = dict(value=ns / (time.perf_counter() - t0), unit='pixels/s', cores=1, kind='port', sample='%d pixels, oracle blos_pixel loop' % ns)
        assert np.allclose(ref[:, :3], out_h[:ns, :3], rtol=2e-6, atol=1e-30, equal_nan=True)          # field strengths
        assert np.allclose(ref[:, 3], out_h[:ns, 3], rtol=0, atol=1e-6, equal_nan=True)                   # azimuth
    out['config4_blos_%d' % nx] = record('k_blos', '1-line B_LOS pixels/s',
                                         'blos_proc_1pix on a %dx%d 1-line map (%.0f MB in+out per launch)' % (nx, nx, 34.0 * npix / 1e6),
                                         ms, 34, e2e, 16 * npix, 18 * npix, cpu_rec)
    del d_in, d_out, d_fl

    two = synth.make_twoline_raw(nx, nx)
    kv = synth.make_keyvals(nx, nx, cdelt=5e-5)
    xs = np.array([kv[8] + i * kv[11] for i in range(nx)], dtype=np.float64)
    ys = np.array([kv[9] + i * kv[12] for i in range(nx)], dtype=np.float64)
    d_xs, d_ys = torch.from_numpy(xs).to(B.dev), torch.from_numpy(ys).to(B.dev)
    d2 = [torch.from_numpy(two).to(B.dev) for _ in range(NSET)]
    o2 = [torch.empty_like(d2[0]) for _ in range(NSET)]
    d_a = [torch.empty((nx, nx), dtype=torch.float32, device=B.dev) for _ in range(NSET)]
    d_y = [torch.empty((nx, nx), dtype=torch.float32, device=B.dev) for _ in range(NSET)]
    two_p = B.pinned(two).numpy()
    for label, hp in (('', None), ('_hpxy', 'grid')):
        d_hp = hp_np = None
        if hp:
            hp_np = np.stack(np.meshgrid(xs.astype(np.float32), ys.astype(np.float32), indexing='ij')).astype(np.float32)
            d_hp = [torch.from_numpy(hp_np).to(B.dev) for _ in range(NSET)]
        ms = timed(lambda i: ctx.qurotate_dev(nx, nx, d_xs.data_ptr(), d_ys.data_ptr(), kv[14], d_hp[i].data_ptr() if hp else None,
                                              d2[i].data_ptr(), o2[i].data_ptr(), d_a[i].data_ptr(), d_y[i].data_ptr(), B.stream.cuda_stream))
        ctx.qurotate(two_p, xs, ys, kv[14], hp_np)
        t0 = time.perf_counter()
        rot_h, a_h, y_h = ctx.qurotate(two_p, xs, ys, kv[14], hp_np)
        e2e = time.perf_counter() - t0
        cpu_rec = None
        if cpu:
            nsx = 48
            sub_kv = list(kv)
            sub_kv[0], sub_kv[1] = nsx, nsx
            t0 = time.perf_counter()
            r_ref, a_ref, y_ref = orc.qurotate_map(two[:nsx, :nsx], sub_kv, hp_np[:, :nsx, :nsx] if hp else np.zeros((2, nsx, nsx)))
            cpu_rec = dict(value=nsx * nsx / (time.perf_counter() - t0), unit='pixels/s', cores=1, kind='port',
                           sample='%dx%d pixels, oracle qurotate_map' % (nsx, nsx))
            assert np.allclose(a_ref, a_h[:nsx, :nsx], atol=1e-6) and np.allclose(y_ref, y_h[:nsx, :nsx], rtol=1e-6)
        bpp = 72 + (8 if hp else 0)
        out['config4_qurotate_%d%s' % (nx, label)] = record(
            'k_qurotate' + label, 'QU rotation + height pixels/s',
            'obs_qurotate + obs_calcheight on a %dx%d 2-line map, %s (%.0f MB per launch)' % (nx, nx, 'Cryo-NIRSP coordinate grid' if hp else 'rectangular grid', bpp * npix / 1e6),
            ms, bpp, e2e, (32 + (8 if hp else 0)) * npix, 40 * npix, cpu_rec)
    return out


# ---------------------------------------------------------------------------------------------------------------
# config 5: ONE 1024 x 1024 map over several hundred databases, pixel-sharded over the ranks (strong scaling)
# ---------------------------------------------------------------------------------------------------------------
def bench_config5(B, nx, ndb, steps):
    torch, native, ctx = B.torch, B.native, B.ctx
    from cledb_b200 import synth
    k, npix = NSEARCH, nx * nx
    nh = max(1, int(round(ndb ** 0.5)))
    nd = (ndb + nh - 1) // nh
    grid = [(h, d) for h in range(nh) for d in range(nd)][:ndb]
    rng = np.random.default_rng(4321)                        # the same map on every rank
    band = (np.arange(nx) * nh // nx)[:, None]
    enc = np.minimum(band * nd + rng.integers(0, nd, (nx, nx)), ndb - 1).astype(np.int32).reshape(-1)
    kv = synth.make_keyvals(nx, nx)
    aobs, yobs = synth.geometry(nx, nx, kv)
    dobs = rng.uniform(7.0, 9.5, npix).astype(np.float32)
    order, bounds = native.partition(enc, np.full(ndb, float(NDB_ROWS)), B.world)
    mine = order[bounds[B.rank]:bounds[B.rank + 1]]
    n = int(mine.size)
    emine = enc[mine]
    used = np.unique(emine)
    t0 = time.perf_counter()
    # this rank generates, uploads and draws its observation from only the databases of its range
    sobs, snr = np.empty((n, 8), np.float32), np.empty((n, 8), np.float32)
    ctx.set_pruning(True)
    lut = np.full(ndb, -1, np.int32)
    for j, e in enumerate(used):
        h, d = grid[int(e)]
        db = synth.make_database((5 + 3 * h) % 99, (30 + 2 * d) % 121)
        sel = np.flatnonzero(emine == e)
        o = synth.make_observation(sel.size, 1, [db], seed=1000 + int(e))
        sobs[sel], snr[sel] = o['sobs_totrot'].reshape(-1, 8), o['snr'].reshape(-1, 8)
        ctx.db_put(j, db, B.args.encoding)
        lut[e] = j
    setup_s = time.perf_counter() - t0
    hdr = synth.make_dbhdr()
    ang = aobs.reshape(-1)[mine]
    with np.errstate(all='ignore'):
        rc, rs = np.cos(-2. * ang), np.sin(-2. * ang)
    ids = lut[emine]
    yl, dl = yobs.reshape(-1)[mine], dobs[mine]
    maxcount = int(np.max(np.diff(bounds)))
    lay = np.zeros(5, np.int64)
    native.load_library().cledb_shard_layout(maxcount, k, lay.ctypes.data_as(ctypes.c_void_p))
    o_inv, o_sf, o_st, o_iss, rank_bytes = (int(v) for v in lay)
    dt = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(B.dev)
    d_sobs, d_snr, d_ids, d_y, d_d, d_rc, d_rs = dt(sobs), dt(snr), dt(ids), dt(yl), dt(dl), dt(rc), dt(rs)
    g = native.DbGrid.from_dbhdr(hdr)
    tabs = [dt(t) for t in g._tables]
    g.sin_bphi, g.cos_bphi, g.sin_btheta, g.cos_btheta = (t.data_ptr() for t in tabs)
    send = torch.empty(rank_bytes, dtype=torch.uint8, device=B.dev)
    sp = send.data_ptr()
    mlay = np.zeros(5, np.int64)
    native.load_library().cledb_map_layout(npix, k, mlay.ctypes.data_as(ctypes.c_void_p))
    m_inv, m_sf, m_st, m_iss, map_bytes = (int(v) for v in mlay)
    win = ctx.comm_window(map_bytes)                        # the full maps of this rank: [invout | sfound | status | issue]
    d_mine = dt(mine.astype(np.int32))
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    state = dict(recv=None, order=None, bounds=None, parts=[])

    def search():
        if n:
            ctx.invert_dev(native.MODE_IQUV, n, d_sobs.data_ptr(), d_snr.data_ptr(), d_ids.data_ptr(), k, d_y.data_ptr(), d_d.data_ptr(),
                           d_rc.data_ptr(), d_rs.data_ptr(), None, 0, g, 1000, 0, sp + o_inv, sp + o_sf, sp + o_st, sp + o_iss,
                           B.stream.cuda_stream)

    def step_peer(record=False):
        """search + solutions of the shard -> ONE kernel stores every record at its pixel in every rank's maps (NVLink)"""
        if record:
            ev[0].record(B.stream)
        search()
        if record:
            ev[1].record(B.stream)
        ctx.gather_maps_dev(npix, n, d_mine.data_ptr(), maxcount, k, sp, win, -1, B.stream.cuda_stream)
        if record:
            ev[2].record(B.stream)
            torch.cuda.synchronize()
            state['parts'].append([ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), 0.0])

    def step_pipe(record=False):
        """the same as a pipeline (cledb_invert_gather_dev): the shard in pieces, the records of a piece cross NVLink while the
        next piece is searched"""
        ctx.invert_gather_dev(native.MODE_IQUV, n, d_sobs.data_ptr(), d_snr.data_ptr(), d_ids.data_ptr(), k, d_y.data_ptr(), d_d.data_ptr(),
                              d_rc.data_ptr(), d_rs.data_ptr(), None, 0, g, 1000, 0, npix, d_mine.data_ptr(), win, -1, B.args.pieces,
                              B.stream.cuda_stream)
        if record:
            torch.cuda.synchronize()
            state['parts'].append([0.0, 0.0, 0.0])

    def step_nccl(record=False):
        """search + solutions of the shard -> ncclAllGather of the packed records -> scatter kernel"""
        if record:
            ev[0].record(B.stream)
        search()
        if record:
            ev[1].record(B.stream)
        if B.world > 1:
            ctx.allgather_dev(sp, state['recv'].data_ptr(), rank_bytes, B.stream.cuda_stream)
        if record:
            ev[2].record(B.stream)
        ctx.unshard_dev(B.world, state['bounds'].data_ptr(), state['order'].data_ptr(), npix, maxcount, k,
                        state['recv'].data_ptr() if B.world > 1 else sp, win + m_inv, win + m_sf, win + m_st, win + m_iss, B.stream.cuda_stream)
        if record:
            ev[3].record(B.stream)
            torch.cuda.synchronize()
            state['parts'].append([ev[i].elapsed_time(ev[i + 1]) for i in range(3)])

    def measure(step, nst):
        state['parts'] = []
        for _ in range(3):
            step()
        B.barrier()
        for _ in range(3):
            B.flush.zero_()
            step(record=True)                               # where the step's time goes (search+build / gather / scatter)
        B.barrier()
        launches0 = ctx.launch_count()
        step_ms, kern_ms = B.timed(step, nst)
        B.barrier()
        part = [B.max_over_ranks(float(v)) for v in np.mean(np.array(state['parts']), axis=0)]
        return B.max_over_ranks(float(np.mean(step_ms))), kern_ms, part, ctx.launch_count() - launches0

    nst = max(3, min(steps, 10))
    peer = B.transport == 'peer'
    nccl_rec = None
    if not peer or B.world > 1:                             # the collective route (also measured beside the peer one when N > 1)
        ctx.comm_transport('nccl')
        state['recv'] = torch.empty(rank_bytes * B.world, dtype=torch.uint8, device=B.dev) if B.world > 1 else None
        state['order'], state['bounds'] = dt(order), dt(bounds)
        ms, kern_ms, part, launches = measure(step_nccl, nst)
        maps_nccl = B.view(win, map_bytes).clone() if peer else None
        nccl_rec = dict(value=npix / (ms * 1e-3), unit='pixels/s', ms_per_step=ms, steps=nst,
                        step_breakdown_ms=dict(search_and_build=part[0], all_gather=part[1], scatter=part[2]),
                        how='ncclAllGather of the packed records (%.0f MB into every GPU) + k_unshard' % (rank_bytes * (B.world - 1) / 1e6))
        if peer:
            state['recv'] = None
    step = step_nccl
    unpiped = None
    if peer:
        ctx.comm_transport('peer')
        ms, kern_ms, part, launches = measure(step_peer, nst)
        step = step_peer
        if nccl_rec is not None:
            nccl_rec['same_maps_as_peer_transport'] = bool(torch.equal(maps_nccl, B.view(win, map_bytes)))
        if B.world > 1:                                     # the pipelined call is the product's path (cledb_invert_sharded uses it)
            unpiped = dict(ms_per_step=ms, value=npix / (ms * 1e-3), how='search + solutions of the whole shard, then one k_push_records')
            maps_one = B.view(win, map_bytes).clone()
            ms, kern_ms, _, launches = measure(step_pipe, nst)
            unpiped['same_maps_as_pipelined'] = bool(torch.equal(maps_one, B.view(win, map_bytes)))
            del maps_one
            step = step_pipe
        if nccl_rec is not None:
            del maps_nccl
    stats = ctx.pruning_stats()
    f_st = B.view(win + m_st, npix)
    nsearched = B.sum_over_ranks([int((f_st == 0).sum().item()) if B.rank == 0 else 0])[0]
    how = ('ONE kernel (k_push_records) stores every finished record (%d B/pixel) straight at its pixel in the maps of every rank over '
           'NVLink / NVSwitch peer memory, between two rounds of flags' if peer else
           'one ncclAllGather of the packed finished records (%d B/pixel) -> scatter into the full maps on every rank') % (k * 76 + 5)
    rec = dict(metric='2-line pixels/s', value=npix / (ms * 1e-3), unit='pixels/s', ms_per_step=ms, steps=nst, scaling='strong',
               workload='ONE %dx%d 2-line IQUV map over %d synthetic PyCELP-shaped databases (height bands x densities; %.1f GB of float32 '
                        'databases, %s in HBM), pixel-sharded over %d GPU(s): cledb_partition -> cledb_invert_dev on the shard -> %s'
                        % (nx, nx, ndb, ndb * NDB_ROWS * 32 / 1e9, 'int16' if B.args.encoding == 1 else 'float32', B.world, how),
               transport=B.transport,
               search='exact tile pruning (library default)', pixels_searched=nsearched, databases_on_this_rank=int(used.size),
               shard_pixels=n, gpu_launches=int(launches),
               step_breakdown_ms=dict(search_and_build=part[0], gather=part[1], scatter=part[2],
                                      note='CUDA events between the calls of one step, max over ranks; every GPU receives %.0f MB of '
                                           'records from its peers' % ((npix - n) * (k * 76 + 5) / 1e6)),
               search_kernel_ms=float(np.sum(kern_ms)) / nst if len(kern_ms) else None, pruning=stats, setup_seconds=setup_s)
    if nccl_rec is not None and peer:
        rec['nccl_transport'] = nccl_rec
    if unpiped is not None:
        rec['unpipelined'] = unpiped
        rec['pipeline'] = ('cledb_invert_gather_dev: the shard goes through in pieces (2 from six ranks on when a piece still holds 64 Ki pixels, else 1); the records '
                           'of a piece cross NVLink on a second stream (k_push_records on 24 blocks of 1024 threads) while the next piece is searched; '
                           'step_breakdown_ms is the unpipelined step')
    # the same step with the brute-force search (k-d layout, seeded: every (pixel, entry) pair evaluated - the FP32-bound kernel)
    ctx.set_pruning(2)
    for _ in range(2):
        step()
    B.barrier()
    nb = max(2, min(steps, 4))
    b_ms, b_kern = B.timed(step, nb)
    B.barrier()
    ctx.set_pruning(1)
    bms = B.max_over_ranks(float(np.mean(b_ms)))
    st_local = send[o_st:o_st + n].cpu().numpy() if n else np.zeros(0, np.uint8)
    info = [ctx.db_info(j) for j in range(used.size)]
    cnt_of = np.array([[i['npos'], i['nneg'], i['nzero']] for i in info], dtype=np.int64).reshape(-1, 3)
    v = sobs[:, 3]
    cls_of = np.where(v > 0, 0, np.where(v < 0, 1, 2))
    ncand = int(cnt_of[ids, cls_of][st_local == 0].sum()) if n else 0
    bk = float(np.sum(b_kern)) / nb if len(b_kern) else float('nan')        # search-kernel time per step (a piece = a launch)
    frac = B.max_over_ranks(20.0 * ncand / (bk * 1e-3) / 1e12 / B.fp32_peak) if B.fp32_peak else None
    rec['bruteforce'] = dict(value=npix / (bms * 1e-3), unit='pixels/s', ms_per_step=bms, steps=nb, search_kernel_ms=B.max_over_ranks(bk),
                             search='brute force, every (pixel, entry) pair evaluated; k-d tile layout, seeded (cledb_set_pruning 2)',
                             roofline=dict(bound='fp32', kernel='k_match<%s,fp32> seeded' % ('int16' if B.args.encoding == 1 else 'float32'),
                                           frac=frac, peak=B.fp32_peak, unit='TFLOP/s', flop_per_candidate=20,
                                           note='max over ranks of 20 FLOP x the rank\'s candidates / its kernel time / measured FP32 peak'))
    # end to end: the host-buffer collective call (every rank its shard in, the full maps out on rank 0)
    out = ctx.result_buffers(npix, k) if B.rank == 0 else None
    for _ in range(2):
        ctx.invert_sharded(native.MODE_IQUV, sobs, snr, ids, k, yl, dl, ang, None, hdr, 1000, 0, npix, order, bounds, root=0, out=out)
    B.barrier()
    ne = max(2, min(steps, 5))
    t0 = time.perf_counter()
    for _ in range(ne):
        ctx.invert_sharded(native.MODE_IQUV, sobs, snr, ids, k, yl, dl, ang, None, hdr, 1000, 0, npix, order, bounds, root=0, out=out)
    e2e_s = B.max_over_ranks((time.perf_counter() - t0) / ne)
    rec['e2e'] = dict(value=npix / e2e_s, unit='pixels/s', ms_per_step=e2e_s * 1e3, h2d_bytes_per_step=int(n) * 84,
                      d2h_bytes_per_step=npix * (k * 76 + 5),
                      api='cledb_invert_sharded (C-ABI, collective): every rank uploads its shard, rank 0 receives the full invout / sfound '
                          'maps in pinned host memory (%.0f MB over one PCIe link)' % (npix * (k * 76 + 5) / 1e6))
    if B.rank == 0:
        f_inv = B.view(win + m_inv, npix * k * 44).cpu().numpy().view(np.float32).reshape(npix, k, 11)
        same = bool(np.array_equal(out['invout'].reshape(npix, k, 11), f_inv, equal_nan=True))
        rec['e2e']['same_maps_as_device_path'] = same
    for j in range(used.size):
        ctx.db_drop(j)
    return rec


def main():
    args = parse()
    if args.impl == 'reference':
        return run_reference_arm(args)
    B = Bench(args)
    world = B.world
    extra = [c for c in args.configs.split(',') if c] or (['2b', '3', '4', '5'] if world == 1 else ['2b', '5'])
    top = bench_search(B, args.mode, args.nx, args.ndb, args.encoding, args.steps, args.warmup, with_gather=True,
                       want_python=True, want_stream=True)
    db, obs = top.pop('_inputs')
    main_rec = top['bruteforce' if args.no_prune else 'pruned']
    by = {}
    tag = 'config%s_%s_%d' % ('2' if args.mode == 'iquv' else '3', args.mode, args.nx)
    by[tag + '_pruned'] = top['pruned']
    if '2b' in extra:
        by[tag + '_bruteforce'] = top['bruteforce']
        by[tag + '_bruteforce_plain'] = top['bruteforce_plain']
    cpu_main = None
    # the CPU baseline is timed at N = 1 only (rank 0 of a larger job would keep the other ranks waiting in the next collective)
    if world > 1:
        args.no_cpu_baseline = True
    if B.rank == 0 and not args.no_cpu_baseline:
        cpu_main = cpu_baseline(args.mode, args.encoding, db, obs, args.cpu_seconds)
        for r in (top['pruned'], top['bruteforce'], top['bruteforce_plain']):
            r['cpu_baseline'] = cpu_main
    del db, obs
    if '3' in extra and args.mode == 'iquv':
        other = bench_search(B, 'iqud', args.nx, args.ndb, args.encoding, max(5, args.steps // 2), args.warmup, with_gather=False,
                             want_python=False, want_stream=False)
        db3, obs3 = other.pop('_inputs')
        by['config3_iqud_%d_pruned' % args.nx] = other['pruned']
        by['config3_iqud_%d_bruteforce' % args.nx] = other['bruteforce']
        by['config3_iqud_%d_bruteforce_plain' % args.nx] = other['bruteforce_plain']
        if B.rank == 0 and not args.no_cpu_baseline:
            c3 = cpu_baseline('iqud', args.encoding, db3, obs3, min(args.cpu_seconds, 10.0))
            other['pruned']['cpu_baseline'] = other['bruteforce']['cpu_baseline'] = other['bruteforce_plain']['cpu_baseline'] = c3
        del db3, obs3
    if '4' in extra:
        by.update(bench_elementwise(B, 2048, max(10, args.steps), cpu=not args.no_cpu_baseline))
    if '5' in extra:
        rec5 = bench_config5(B, args.nx5, args.ndb5, args.steps)
        if cpu_main is not None:
            rec5['cpu_baseline'] = dict(cpu_main, sample=cpu_main['sample'] + ' (per-pixel cost of the CPU path does not depend on how many '
                                                                             'databases the map touches: same figure as config 2)')
        by['config5_iquv_%d_multidb' % args.nx5] = rec5
    B.barrier()
    clocks = None
    if B.rank == 0:
        clocks = B.sampler.summary()
        B.sampler.stop()
    if B.rank != 0:
        if world > 1:
            B.dist.destroy_process_group()
        return
    line = dict(
        metric='2-line pixels/s', value=main_rec['value'], unit='pixels/s', n_gpus=world, steps=main_rec['steps'],
        warmup=max(args.warmup, 3), ms_per_step=main_rec['ms_per_step'], higher_is_better=True, scaling='weak', vs_baseline=None,
        dtype='f32', data='synthetic',
        config=dict(workload=main_rec['workload'], db_encoding='int16' if args.encoding == 1 else 'float32',
                    parallelism=('pixel-sharded x%d (one map per rank), database replicated, one all-gather of the idx/chi maps by the '
                                 'library (cledb_allgather_dev: %s)'
                                 % (world, 'every rank stores its block straight into every rank\'s window over NVLink / NVSwitch peer '
                                           'memory, k_push_block between two rounds of flags' if B.transport == 'peer' else
                                    'ncclAllGather over NVLink')) if world > 1 else 'single GPU',
                    transport=B.transport if world > 1 else None,
                    search=main_rec['search'],
                    l2='flushed between timed steps (256 MiB memset, outside the per-step CUDA events)',
                    timing='CUDA events per step on the launching stream, mean of K, max over ranks' +
                           ('; the gather of map i-1 (second stream) overlaps the search of map i, every gather joined inside a timed '
                            'bracket (gather_overlap.ms_per_step_serial = without the overlap)' if main_rec.get('gather_overlap') else '')),
        evals_per_s=main_rec['evals_per_s'], evals_definition=main_rec['evals_definition'],
        nominal_scanned_evals_per_s=main_rec['nominal_scanned_evals_per_s'],
        nominal_candidate_evals_per_s=main_rec['nominal_candidate_evals_per_s'],
        pixels_searched=main_rec['pixels_searched'], gpu_launches=main_rec['gpu_launches'], pruning=main_rec['pruning'],
        e2e=main_rec.get('e2e'), roofline=main_rec['roofline'], clocks=clocks, gather_overlap=main_rec.get('gather_overlap'),
        device=dict(sm_count=B.props['sm_count'], match_ctas_per_sm=B.props['match_ctas_per_sm'], match_regs=B.props['match_regs'],
                    match_smem=B.props['match_smem']),
        by_config=by)
    if cpu_main is not None:
        line['cpu_baseline'] = cpu_main
    if 'bruteforce' in top and not args.no_prune:
        line['roofline_bruteforce'] = dict(top['bruteforce']['roofline'], value=top['bruteforce']['value'],
                                           ms_per_step=top['bruteforce']['ms_per_step'],
                                           note='the FP32-bound kernel of the path - every (pixel, entry) pair evaluated -, same maps, same run: '
                                                'by_config.%s_bruteforce (k-d layout, seeded) and %s_bruteforce_plain' % (tag, tag))
    print(json.dumps(line))
    if world > 1:
        B.dist.destroy_process_group()


if __name__ == '__main__':
    main()
