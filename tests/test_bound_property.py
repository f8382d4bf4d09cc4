"""CPU statement of the exactness argument behind the pruned search (csrc/match.cu:tile_lower_bound): with every
operation rounded to float32 the way the kernel rounds it, the bound computed from a tile's component box with the SAME
fma chain is <= the chain value of every entry in the box - for coefficients of either sign, boxes that straddle zero,
zeros and huge dynamic ranges.  fma is emulated as round32(round64(a*b + c)): a*b is exact in float64, and a composition of
monotone roundings is monotone, which is all the argument needs."""
import numpy as np
import pytest

F32 = np.float32


def fma32(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(F32)


def chain(d, a, nb, c0):
    """acc = c0; for c: r = fma(d_c, a_c, nb_c); acc = fma(r, r, acc)   (eval_fp32 in csrc/cledb_internal.h)"""
    acc = np.full(d.shape[:-1], c0, F32)
    for c in range(d.shape[-1]):
        r = fma32(d[..., c], np.broadcast_to(a[c], d.shape[:-1]), np.broadcast_to(nb[c], d.shape[:-1]))
        acc = fma32(r, r, acc)
    return acc


def lower_bound(lo, hi, a, nb, c0):
    lb = F32(c0)
    for c in range(lo.shape[0]):
        rl = fma32(lo[c:c + 1], a[c:c + 1], nb[c:c + 1])[0]
        rh = fma32(hi[c:c + 1], a[c:c + 1], nb[c:c + 1])[0]
        mn = max(max(min(rl, rh), -max(rl, rh)), F32(0))
        if np.isnan(rl + rh):
            mn = F32(0)
        lb = fma32(np.array([mn], F32), np.array([mn], F32), np.array([lb], F32))[0]
    return lb


@pytest.mark.parametrize('seed', range(8))
def test_box_bound_never_exceeds_an_entry(seed):
    rng = np.random.default_rng(seed)
    for trial in range(150):
        n = int(rng.integers(1, 257))
        scale = 10.0 ** rng.integers(-6, 3, 5)
        centre = rng.standard_normal(5) * scale * rng.choice([0.0, 1.0, 30.0])
        d = (centre + rng.standard_normal((n, 5)) * scale * 10.0 ** rng.integers(-4, 1)).astype(F32)
        if trial % 7 == 0:
            d[rng.integers(0, n), rng.integers(0, 5)] = 0.0
        lo, hi = d.min(axis=0), d.max(axis=0)
        a = (rng.choice([-1.0, 1.0], 5) * 10.0 ** rng.uniform(-2, 8, 5)).astype(F32)           # w/sigma: either sign, any size
        target = (centre + rng.standard_normal(5) * scale * 10.0 ** rng.integers(-3, 2)).astype(F32)
        nb = (-(a.astype(np.float64)) * target).astype(F32)
        c0 = F32(0.0 if trial % 2 else rng.uniform(0, 1e6))
        with np.errstate(over='ignore', invalid='ignore'):
            v = chain(d, a, nb, c0)
            lb = lower_bound(lo, hi, a, nb, c0)
        finite = np.isfinite(v)
        if finite.any():
            assert lb <= v[finite].min(), (seed, trial, lb, v[finite].min())
        assert not (lb > v[~np.isnan(v)]).any()
