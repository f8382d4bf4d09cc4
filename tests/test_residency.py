"""CPU tests of the database residency cache (cledb_b200.residency) with a stub device context: identity + lifetime keyed
(a freed array whose address NumPy reuses must never be taken for the old database), content re-checked on
every hit with a fingerprint that sees all eight Stokes components, least-recently-used eviction under an HBM budget."""
import gc

import numpy as np

from cledb_b200.residency import Resident, fingerprint
import cledb_b200.synth as synth


class StubCtx:
    def __init__(self):
        self.dbs, self.log = {}, []

    def db_put(self, db_id, arr, enc):
        self.dbs[db_id] = np.array(arr, copy=True)
        self.log.append(('put', db_id))

    def db_drop(self, db_id):
        del self.dbs[db_id]
        self.log.append(('drop', db_id))

    def db_info(self, db_id):
        return dict(hbm_bytes=self.dbs[db_id].nbytes)


def test_freed_array_with_recycled_address_is_a_new_database():
    ctx, res = StubCtx(), Resident()
    for trial in range(20):
        a = synth.make_database(8, 39 + trial, ngx=3, nbphi=12, nbtheta=8)
        ids = res.ensure_all(ctx, {0: a}, 0)
        assert np.array_equal(ctx.dbs[ids[0]].reshape(a.shape), a), 'the device holds THIS array'
        addr = a.__array_interface__['data'][0]
        del a
        gc.collect()
    assert res.uploads == 20
    res.ensure_all(ctx, {}, 0)
    assert len(ctx.dbs) == 0 and len(res.entries) == 0, 'dead arrays are dropped from the device'


def test_hit_reupload_on_change_and_fingerprint_sees_every_component():
    ctx, res = StubCtx(), Resident()
    a = synth.make_database(9, 40)                     # canonical shape: size // 4096 is a multiple of 8
    b = synth.make_database(9, 41)
    ids = res.ensure_all(ctx, {0: a, 1: b}, 0)
    assert res.uploads == 2 and ids[0] != ids[1]
    assert res.ensure_all(ctx, {0: a, 1: b}, 0) == ids and res.uploads == 2       # hits
    assert fingerprint(a) != fingerprint(b), 'two databases of the same shape differ in the sampled fingerprint'
    a[..., 1] *= np.float32(1.5)                       # in-place edit of one component
    assert res.ensure_all(ctx, {0: a}, 0)[0] == ids[0] and res.uploads == 3
    assert np.array_equal(ctx.dbs[ids[0]].reshape(a.shape), a)
    res.verify = 'full'
    a[3, 7, 5, 2] += np.float32(1e-3)                  # a single element: only the full hash sees it
    res.ensure_all(ctx, {0: a}, 0)
    assert res.uploads == 4
    f64 = a.astype(np.float64)                         # not float32: keyed on the caller's object, converted for the upload
    i1 = res.ensure_all(ctx, {0: f64}, 0)
    i2 = res.ensure_all(ctx, {0: f64}, 0)
    assert i1 == i2 and res.uploads == 5
    assert res.ensure_all(ctx, {0: a}, 1)[0] == ids[0] and res.uploads == 6        # other encoding: re-upload


def test_lru_eviction_under_budget():
    ctx = StubCtx()
    dbs = [synth.make_database(3, i, ngx=3, nbphi=12, nbtheta=8) for i in range(6)]
    res = Resident(budget_bytes=dbs[0].nbytes * 3)
    for i in range(6):
        res.ensure_all(ctx, {0: dbs[i]}, 0)
    assert len(ctx.dbs) == 3 and res.resident_bytes() <= res.budget_bytes
    ids = res.ensure_all(ctx, {i: dbs[i] for i in range(6)}, 0)                    # one call needs all six: nothing in use is evicted
    assert len(ctx.dbs) == 6 and len(set(ids.values())) == 6
    res.ensure_all(ctx, {0: dbs[5]}, 0)
    assert len(ctx.dbs) == 3 and any(np.array_equal(v.reshape(dbs[5].shape), dbs[5]) for v in ctx.dbs.values())
    res.clear(ctx)
    assert not ctx.dbs
