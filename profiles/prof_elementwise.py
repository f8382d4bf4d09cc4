#!/usr/bin/env python
"""Small driver for ncu: BASELINE config 4 (2048x2048 map) through cledb_blos_dev / cledb_qurotate_dev, 3 launches each.
    ncu --set full --clock-control none --import-source on -k regex:k_blos -s 1 -c 1 -o gpurun_out/k_blos python profiles/prof_elementwise.py"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, 'solar-coronal-inversion_b200')]
import numpy as np, torch
from cledb_b200 import native, synth
import constants as consts
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
npix = nx * nx
dev = torch.device('cuda', 0)
ctx = native.Context(0)
s = torch.cuda.Stream(device=dev)
one = synth.make_oneline_map(nx, nx)
c = consts.Constants('fe-xiii_1074')
kfac = 1.e9 * (c.planckconst * c.l_speed / c.bohrmagneton / (c.line_ref ** 2))
d_in = torch.from_numpy(one.reshape(-1, 4)).to(dev); d_out = torch.empty_like(d_in)
d_fl = torch.empty((npix, 2), dtype=torch.uint8, device=dev)
two = synth.make_twoline_raw(nx, nx)
kv = synth.make_keyvals(nx, nx, cdelt=5e-5)
xs = np.array([kv[8] + i * kv[11] for i in range(nx)]); ys = np.array([kv[9] + i * kv[12] for i in range(nx)])
d_xs, d_ys, d2 = torch.from_numpy(xs).to(dev), torch.from_numpy(ys).to(dev), torch.from_numpy(two).to(dev)
o2 = torch.empty_like(d2); d_a = torch.empty((nx, nx), dtype=torch.float32, device=dev); d_y = torch.empty_like(d_a)
torch.cuda.synchronize()
for _ in range(3):
    ctx.blos_dev(npix, d_in.data_ptr(), kfac, c.g_eff, c.F_factor, d_out.data_ptr(), d_fl.data_ptr(), s.cuda_stream)
    ctx.qurotate_dev(nx, nx, d_xs.data_ptr(), d_ys.data_ptr(), kv[14], None, d2.data_ptr(), o2.data_ptr(), d_a.data_ptr(), d_y.data_ptr(), s.cuda_stream)
torch.cuda.synchronize()
print('ok', float(d_out.nan_to_num().sum()), float(o2.nan_to_num().sum()))
