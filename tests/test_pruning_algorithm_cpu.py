"""CPU walk-through of the pruned search in NumPy with float32-rounded arithmetic: k-d tiles with component
boxes, own-seed-tile bootstrap, tile bounds against the running threshold - and the same top-k (values and first-occurrence
ties) as the brute-force evaluation of every entry.  Algorithm-level check that needs no GPU; the kernels are compared on
the device in tests/test_gpu_parity.py."""
import numpy as np
import pytest

import cledb_b200.synth as synth
from test_bound_property import chain, lower_bound

F32 = np.float32
COMPS = [1, 2, 4, 5, 6]
TILE = 32


def kd_order(vals, ids):
    """Leaves of TILE entries: median split at a multiple of TILE along the component with the widest weighted extent."""
    nt = -(-len(ids) // TILE)
    if nt <= 1:
        return ids
    w = np.array([1, 1, 0.1, 1, 1])
    ext = (vals[ids].max(axis=0) - vals[ids].min(axis=0)) * w
    dim = int(np.argmax(ext))
    mid = (nt // 2) * TILE
    order = ids[np.lexsort((ids, vals[ids, dim]))]
    return np.concatenate([kd_order(vals, order[:mid]), kd_order(vals, order[mid:])])


def topk_brute(v, rows, k):
    order = np.lexsort((rows, v))               # value, then original row: first-occurrence ties
    order = order[~np.isnan(v[order])][:k]
    return v[order], rows[order]


@pytest.mark.parametrize('seed', [0, 1, 2])
def test_pruned_search_equals_brute_force(seed):
    rng = np.random.default_rng(seed)
    db = synth.make_database(3 + seed, 17, ngx=3, nbphi=24, nbtheta=16).reshape(-1, 8)
    db = np.concatenate([db, db[rng.choice(len(db), 40)]])          # duplicated rows: exact ties
    vals = db[:, COMPS].astype(F32)
    k = 8
    evaluated = total = 0
    for sign in (1, -1):
        rows = np.flatnonzero(np.sign(db[:, 3]) == sign)
        order = kd_order(vals, rows)
        tiles = [order[i:i + TILE] for i in range(0, len(order), TILE)]
        lo = np.array([vals[t].min(axis=0) for t in tiles])
        hi = np.array([vals[t].max(axis=0) for t in tiles])
        rep = np.array([vals[t[np.argmin((((vals[t] - 0.5 * (lo[i] + hi[i])) * [1, 1, 0.1, 1, 1]) ** 2).sum(axis=1))]] for i, t in enumerate(tiles)])
        for _ in range(25):
            truth = vals[rng.choice(rows)]
            sig = (np.abs(truth) * rng.uniform(0.002, 0.3) + 1e-6).astype(F32)
            s = (truth + sig * rng.standard_normal(5)).astype(F32)
            a = (np.array([1, 1, 0.01, 1, 1]) / sig.astype(np.float64)).astype(F32)
            nb = (-(a.astype(np.float64)) * s).astype(F32)
            # brute force over the class
            want_v, want_r = topk_brute(chain(vals[rows], a, nb, F32(0)), rows, k)
            # pruned: seed tile = best representative, exact top-k of it, then tiles whose bound does not exceed tau
            seed_tile = int(np.argmin(chain(rep, a, nb, F32(0))))
            cand_v = chain(vals[tiles[seed_tile]], a, nb, F32(0))
            got_v, got_r = topk_brute(cand_v, tiles[seed_tile], k)
            evaluated += 1
            for t in range(len(tiles)):
                total += 1
                if t == seed_tile:
                    continue
                tau = got_v[k - 1] if len(got_v) == k else F32(np.inf)
                if lower_bound(lo[t], hi[t], a, nb, F32(0)) > tau:
                    continue                                             # the whole tile is above the threshold
                evaluated += 1
                v = np.concatenate([got_v, chain(vals[tiles[t]], a, nb, F32(0))])
                r = np.concatenate([got_r, tiles[t]])
                got_v, got_r = topk_brute(v, r, k)
            assert np.array_equal(got_v, want_v) and np.array_equal(got_r, want_r)
    assert evaluated < 0.6 * total                                         # and most tiles were skipped
