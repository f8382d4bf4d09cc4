#!/usr/bin/env python
"""When do the warps of the pruned search run out of work?  (timing experiment, dbg 7: every warp records the globaltimer
when the item queue is empty.)  Usage: python profiles/exp_tail.py [nx]"""
import os, sys, ctypes
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, 'solar-coronal-inversion_b200')]
import numpy as np, torch
from cledb_b200 import native, synth
dev = torch.device('cuda', 0)
ctx = native.Context(0)
lib = native.load_library()
lib.cledb_debug_stats.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
db = synth.make_database(9, 40)
ctx.db_put(0, db, native.ENC_I16)
for nx in (128, 256, 512, 1024):
    obs = synth.make_observation(nx, nx, [db], seed=7)
    npix, k = nx * nx, 8
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    sobs, snr, ids = d(obs['sobs_totrot'].reshape(-1, 8)), d(obs['snr'].reshape(-1, 8)), d(np.zeros(npix, np.int32))
    idx = torch.empty((npix, k), dtype=torch.int32, device=dev); chi = torch.empty((npix, k), dtype=torch.float32, device=dev)
    st = torch.empty(npix, dtype=torch.uint8, device=dev)
    for mode in (0, 1):
        f = lambda: ctx.match_dev(mode, npix, sobs.data_ptr(), snr.data_ptr(), ids.data_ptr(), k, idx.data_ptr(), chi.data_ptr(), None, None, None, st.data_ptr(), None, stream.cuda_stream)
        f(); f(); torch.cuda.synchronize()
        ctx.set_tuning(-8, 0); f(); torch.cuda.synchronize()
        ctx.profile_begin(); f(); f(); f(); k8 = float(np.mean(ctx.profile_collect()))
        ctx.set_tuning(0, 0)
        ks = []
        for qs in ('0', '2', '4', '6', '8'):
            os.environ['CLEDB_B200_QSPLIT'] = qs
            f(); torch.cuda.synchronize()
            ctx.profile_begin(); f(); f(); f(); ks.append(float(np.mean(ctx.profile_collect())) * 1e3)
        os.environ.pop('CLEDB_B200_QSPLIT')
        print('%d^2 %s kernel: items in index order %.1f us; heaviest first, heavy items in pieces from median bin + 0(never)/2/4/6/8: %s us'
              % (nx, 'IQUD' if mode else 'IQUV', k8 * 1e3, ' '.join('%.1f' % v for v in ks)))
        ctx.set_tuning(-7, 0)
        ctx.profile_begin(); f(); kms = float(ctx.profile_collect()[0])
        buf = np.zeros(1 << 17, np.uint64)
        lib.cledb_debug_stats(ctx.h, buf.ctypes.data_as(ctypes.c_void_p), buf.size)
        ctx.set_tuning(0, 0)
        t0 = int(buf[2]); ends = buf[4:4 + 148 * 16].astype(np.int64); ends = ends[ends > 0] - t0
        items = buf[4096:].reshape(-1, 4)
        items = items[items[:, 1] > 0]
        dur = (items[:, 1].astype(np.int64) - items[:, 0].astype(np.int64)) / 1e3
        start = (items[:, 0].astype(np.int64) - t0) / 1e3
        ev, exact, npx = items[:, 2].astype(np.int64), items[:, 3] & 1, (items[:, 3] >> 8).astype(np.int64)
        o = np.argsort(-dur)
        print('   items %d: duration us median %.1f p90 %.1f p99 %.1f max %.1f; evaluated (pixel, tile) pairs per item median %d p99 %d max %d; exact-only items %d'
              % (dur.size, np.median(dur), np.percentile(dur, 90), np.percentile(dur, 99), dur.max(), np.median(ev), np.percentile(ev, 99), ev.max(), int(exact.sum())))
        print('   the 8 longest items: ' + '; '.join('%.0f us from %.0f us, %d pairs, %d px%s' % (dur[i], start[i], ev[i], npx[i], ' EXACT' if exact[i] else '') for i in o[:8]))
        e = np.sort(ends) / 1e3
        print('%d^2 %s: kernel %.1f us; warps: %d; idle-at (us after the first warp started): min %.0f  p10 %.0f  median %.0f  p90 %.0f  max %.0f; mean busy fraction %.2f'
              % (nx, 'IQUD' if mode else 'IQUV', kms * 1e3, e.size, e[0], e[int(.1 * e.size)], e[e.size // 2], e[int(.9 * e.size)], e[-1], e.mean() / e[-1]), flush=True)
ctx.close()
