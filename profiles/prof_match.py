#!/usr/bin/env python
"""Small driver for ncu: BASELINE config 2 (256x256 IQUV map, one 340200-entry database, int16) through
cledb_match_dev, 2 warm-up calls + 2 profiled calls.  Usage (on the GPU box):
    python profiles/prof_match.py [iquv|iqud] [nx] [prune|brute|seeded]
    ncu --set full --clock-control none --import-source on -k regex:k_match -s 2 -c 1 -o gpurun_out/prof python profiles/prof_match.py iquv 256 brute
"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, 'solar-coronal-inversion_b200')]
import numpy as np  # noqa: E402
import torch  # noqa: E402
from cledb_b200 import native, synth  # noqa: E402

mode = native.MODE_IQUD if len(sys.argv) > 1 and sys.argv[1] == 'iqud' else native.MODE_IQUV
nx = int(sys.argv[2]) if len(sys.argv) > 2 else 256
layout = sys.argv[3] if len(sys.argv) > 3 else 'prune'      # 'prune' (library default), 'brute' (plain layout) or 'seeded' (k-d layout, brute force)
db = synth.make_database(9, 40)
obs = synth.make_observation(nx, nx, [db], seed=7)
ctx = native.Context(0)
ctx.set_pruning({'prune': 1, 'brute': 0, 'seeded': 2}[layout])
ctx.db_put(0, db, native.ENC_I16)
dev = torch.device('cuda', 0)
npix, k = nx * nx, 8
sobs = torch.from_numpy(obs['sobs_totrot'].reshape(-1, 8)).to(dev)
snr = torch.from_numpy(obs['snr'].reshape(-1, 8)).to(dev)
ids = torch.zeros(npix, dtype=torch.int32, device=dev)
idx = torch.empty((npix, k), dtype=torch.int32, device=dev)
chi = torch.empty((npix, k), dtype=torch.float32, device=dev)
st = torch.empty(npix, dtype=torch.uint8, device=dev)
s = torch.cuda.Stream(device=dev)
torch.cuda.synchronize()
ctx.profile_begin()
for i in range(4):
    ctx.match_dev(mode, npix, sobs.data_ptr(), snr.data_ptr(), ids.data_ptr(), k, idx.data_ptr(), chi.data_ptr(),
                  status_ptr=st.data_ptr(), stream=s.cuda_stream)
torch.cuda.synchronize()
print('kernel ms:', ctx.profile_collect(), 'searched', int((st == 0).sum()), 'idx checksum', int(idx.sum()))
