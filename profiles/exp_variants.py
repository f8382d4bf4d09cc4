#!/usr/bin/env python
"""Kernel tuning experiment: time the search (cledb_match_dev, inputs resident, CUDA events) and the fused host call for
the library named by $CLEDB_B200_LIB.  Usage: CLEDB_B200_LIB=... python profiles/exp_variants.py [tag]"""
import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, 'solar-coronal-inversion_b200')]
import numpy as np, torch
from cledb_b200 import native, synth
tag = sys.argv[1] if len(sys.argv) > 1 else os.path.basename(native.LIB_PATH)
dev = torch.device('cuda', 0)
ctx = native.Context(0)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
db = synth.make_database(9, 40)
ctx.db_put(0, db, native.ENC_I16)
hdr = synth.make_dbhdr()
ref = {}
for nx in (256, 1024):
    obs = synth.make_observation(nx, nx, [db], seed=7)
    npix, k = nx * nx, 8
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    sobs, snr, ids = d(obs['sobs_totrot'].reshape(-1, 8)), d(obs['snr'].reshape(-1, 8)), d(np.zeros(npix, np.int32))
    idx = torch.empty((npix, k), dtype=torch.int32, device=dev); chi = torch.empty((npix, k), dtype=torch.float32, device=dev)
    kpos = torch.empty_like(idx); rows = torch.empty((npix, k, 8), dtype=torch.float32, device=dev)
    st = torch.empty(npix, dtype=torch.uint8, device=dev)
    for mode in (0, 1):
        def step():
            ctx.match_dev(mode, npix, sobs.data_ptr(), snr.data_ptr(), ids.data_ptr(), k, idx.data_ptr(), chi.data_ptr(), None,
                          kpos.data_ptr(), rows.data_ptr(), st.data_ptr(), None, stream.cuda_stream)
        for _ in range(3): step()
        torch.cuda.synchronize()
        ctx.profile_begin(); evs = []
        n = 20 if nx == 256 else 6
        for _ in range(n):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream); step(); b.record(stream); evs.append((a, b))
        torch.cuda.synchronize()
        ms = np.mean([a.elapsed_time(b) for a, b in evs]); kms = np.mean(ctx.profile_collect())
        chk = int(idx.to(torch.int64).sum().item()) ^ int(chi.view(torch.int32).to(torch.int64).sum().item())
        print('%-14s %4d^2 %s  step %.3f ms  kernel %.3f ms  %.1f Mpx/s  checksum %x' % (tag, nx, 'IQUD' if mode else 'IQUV', ms, kms, npix / ms / 1e3, chk & 0xffffffffffff), flush=True)
    if nx == 256:
        out = ctx.result_buffers(npix, k)
        s_h, n_h = obs['sobs_totrot'].reshape(-1, 8), obs['snr'].reshape(-1, 8)
        for mode, bcalc in ((0, 0), (1, 3)):
            dopp = obs['sobs_dopp'].reshape(npix, -1)[:, 0].copy() if mode else None
            f = lambda: ctx.invert(mode, s_h, n_h, np.zeros(npix, np.int32), k, obs['yobs'], obs['dobs'], obs['aobs'], dopp, hdr, 1000, bcalc, out=out)
            for _ in range(3): f()
            t0 = time.perf_counter()
            for _ in range(20): f()
            print('%-14s  256^2 %s  cledb_invert (pageable in, pinned out) %.3f ms' % (tag, 'IQUD' if mode else 'IQUV', (time.perf_counter() - t0) / 20 * 1e3), flush=True)
ctx.close()
