#!/usr/bin/env python
"""BASELINE config 4: 1-line analytic B_LOS and QU-rotation(+height) on a 2048x2048 Cryo-NIRSP-shaped synthetic map.

Times k_blos / k_qurotate through the device-pointer C-ABI (cledb_blos_dev / cledb_qurotate_dev) with CUDA events on
the launching stream, inputs resident in HBM and larger than L2 (128-300 MB per pass), reports GB/s of ALGORITHMIC
bytes (34 B/pixel/line and 72 B/pixel) against the measured HBM copy bandwidth in MEASURED_PEAKS.json,
the end-to-end rate through the host-buffer calls, and the oracle port on a bounded sample on the host.
Prints one JSON line per kernel.  Usage (GPU box): python profiles/bench_elementwise.py [nx] [steps]
"""
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, 'solar-coronal-inversion_b200')]
import numpy as np  # noqa: E402
import torch  # noqa: E402
from cledb_b200 import native, synth  # noqa: E402
import constants as consts  # noqa: E402
import oracle.cledb_oracle as orc  # noqa: E402

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
peak = 6553.9
try:
    peak = json.load(open(os.path.join(REPO, 'MEASURED_PEAKS.json')))['hbm_gbs']
    peak_src = 'MEASURED_PEAKS.json hbm_gbs'
except Exception:
    peak_src = 'fallback 6553.9 GB/s'
dev = torch.device('cuda', 0)
ctx = native.Context(0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
npix = nx * nx


def timed(fn):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    evs = []
    for _ in range(steps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        fn()
        b.record(stream)
        evs.append((a, b))
    torch.cuda.synchronize()
    return float(np.mean([a.elapsed_time(b) for a, b in evs]))


# ---- B_LOS -------------------------------------------------------------------------------------------------------
one = synth.make_oneline_map(nx, nx)
c = consts.Constants('fe-xiii_1074')
kfac = 1.e9 * (c.planckconst * c.l_speed / c.bohrmagneton / (c.line_ref ** 2))
d_in = torch.from_numpy(one.reshape(-1, 4)).to(dev)
d_out = torch.empty_like(d_in)
d_fl = torch.empty((npix, 2), dtype=torch.uint8, device=dev)
ms = timed(lambda: ctx.blos_dev(npix, d_in.data_ptr(), kfac, c.g_eff, c.F_factor, d_out.data_ptr(), d_fl.data_ptr(), stream.cuda_stream))
t0 = time.perf_counter()
for _ in range(3):
    out_h, fl_h = ctx.blos(one.reshape(-1, 4), kfac, c.g_eff, c.F_factor)
e2e = (time.perf_counter() - t0) / 3
assert np.array_equal(out_h, d_out.cpu().numpy(), equal_nan=True)
ns = 4096
t0 = time.perf_counter()
ref = np.stack([orc.blos_pixel(one.reshape(-1, 4)[i], c)[0] for i in range(ns)])
cpu = ns / (time.perf_counter() - t0)
assert np.allclose(ref, out_h[:ns], rtol=2e-6, atol=1e-30, equal_nan=True)
gbs = 34.0 * npix / (ms * 1e-3) / 1e9
print(json.dumps(dict(metric='1-line B_LOS pixels/s', value=npix / (ms * 1e-3), unit='pixels/s', ms_per_step=ms, steps=steps,
                      config=dict(workload='blos_proc_1pix on a %dx%d 1-line map (%.0f MB in+out, larger than L2)' % (nx, nx, 34.0 * npix / 1e6)),
                      roofline=dict(bound='hbm', kernel='k_blos', achieved=gbs, peak=peak, unit='GB/s', frac=gbs / peak, traffic=None,
                                    bytes_per_pixel=34, peak_source=peak_src),
                      e2e=dict(value=npix / e2e, unit='pixels/s', h2d_bytes_per_step=16 * npix, d2h_bytes_per_step=18 * npix),
                      cpu_baseline=dict(value=cpu, unit='pixels/s', cores=1, kind='port', sample='%d pixels, oracle blos_pixel loop' % ns))))

# ---- QU rotation + height ------------------------------------------------------------------------------------------
two = synth.make_twoline_raw(nx, nx)
kv = synth.make_keyvals(nx, nx, cdelt=5e-5)
xs = np.array([kv[8] + i * kv[11] for i in range(nx)], dtype=np.float64)
ys = np.array([kv[9] + i * kv[12] for i in range(nx)], dtype=np.float64)
d_xs, d_ys = torch.from_numpy(xs).to(dev), torch.from_numpy(ys).to(dev)
d2 = torch.from_numpy(two).to(dev)
o2 = torch.empty_like(d2)
d_a = torch.empty((nx, nx), dtype=torch.float32, device=dev)
d_y = torch.empty((nx, nx), dtype=torch.float32, device=dev)
for label, hp in (('rectangular grid', None), ('Cryo-NIRSP coordinate grid', 'grid')):
    d_hp = None
    hp_np = None
    if hp:
        hp_np = np.stack(np.meshgrid(xs.astype(np.float32), ys.astype(np.float32), indexing='ij')).astype(np.float32)
        d_hp = torch.from_numpy(hp_np).to(dev)
    ms = timed(lambda: ctx.qurotate_dev(nx, nx, d_xs.data_ptr(), d_ys.data_ptr(), kv[14], d_hp.data_ptr() if hp else None,
                                        d2.data_ptr(), o2.data_ptr(), d_a.data_ptr(), d_y.data_ptr(), stream.cuda_stream))
    t0 = time.perf_counter()
    rot_h, a_h, y_h = ctx.qurotate(two, xs, ys, kv[14], hp_np)
    e2e = time.perf_counter() - t0
    nsx = 32
    sub_kv = list(kv)
    sub_kv[0], sub_kv[1] = nsx, nsx
    t0 = time.perf_counter()
    r_ref, a_ref, y_ref = orc.qurotate_map(two[:nsx, :nsx], sub_kv, hp_np[:, :nsx, :nsx] if hp else np.zeros((2, nsx, nsx)))
    cpu = nsx * nsx / (time.perf_counter() - t0)
    assert np.allclose(a_ref, a_h[:nsx, :nsx], atol=1e-6) and np.allclose(y_ref, y_h[:nsx, :nsx], rtol=1e-6)
    assert np.allclose(r_ref, rot_h[:nsx, :nsx], rtol=1e-5, atol=1e-16)
    bpp = 72 + (8 if hp else 0)
    gbs = bpp * npix / (ms * 1e-3) / 1e9
    print(json.dumps(dict(metric='QU rotation + height pixels/s', value=npix / (ms * 1e-3), unit='pixels/s', ms_per_step=ms, steps=steps,
                          config=dict(workload='obs_qurotate + obs_calcheight on a %dx%d 2-line map, %s (%.0f MB)' % (nx, nx, label, bpp * npix / 1e6)),
                          roofline=dict(bound='hbm', kernel='k_qurotate', achieved=gbs, peak=peak, unit='GB/s', frac=gbs / peak,
                                        traffic=None, bytes_per_pixel=bpp, peak_source=peak_src),
                          e2e=dict(value=npix / e2e, unit='pixels/s', h2d_bytes_per_step=(32 + (8 if hp else 0)) * npix, d2h_bytes_per_step=40 * npix),
                          cpu_baseline=dict(value=cpu, unit='pixels/s', cores=1, kind='port', sample='%dx%d pixels, oracle qurotate_map' % (nsx, nsx)))))
ctx.close()
