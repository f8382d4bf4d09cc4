"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU and exports every symbol the header
declares; the pure-host codec matches NumPy; the Python surface keeps the reference's guards and fails loudly
(no CPU fallback) when asked to compute without a device."""
import os
import re
import subprocess

import numpy as np
import pytest

from cledb_b200 import native, codec

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(REPO, 'include', 'cledb_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(cledb_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    lib = native.load_library()
    declared = header_functions()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), 'libcledb_b200.so does not export %s' % name
    assert sorted(native.SIGNATURES) == declared, 'ctypes prototypes and header drifted apart'
    out = subprocess.run(['nm', '-D', '--defined-only', native.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r' T (cledb_[a-z0-9_]+)', out))
    assert set(declared) <= exported
    assert lib.cledb_version().decode().startswith('cledb_b200')


def test_library_is_built_for_sm100a():
    out = subprocess.run(['cuobjdump', '-lelf', native.LIB_PATH], capture_output=True, text=True).stdout
    assert 'sm_100a' in out, out


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    with pytest.raises(native.CledbError) as e:
        native.Context(0)
    assert 'no CPU fallback' in str(e.value)
    import CLEDB_PROC.CLEDB_PROC as procinv
    import cledb_b200.synth as synth
    db = synth.make_database(8, 39, ngx=3, nbphi=8, nbtheta=6)
    obs = synth.make_observation(2, 2, [db])
    with pytest.raises(native.CledbError):
        procinv.cledb_invproc(obs['sobs_totrot'], 0, [db], obs['db_enc'], obs['yobs'], obs['aobs'], obs['dobs'], obs['snr'],
                              obs['issuemask'], synth.make_dbhdr(3, 8, 6), obs['keyvals'], 8, 1000, 0, False, 4, False, 0)


def test_codec_matches_numpy_half():
    rng = np.random.default_rng(0)
    x = (rng.standard_normal(200000) * 10.0 ** rng.integers(-14, 4, 200000)).astype(np.float32)
    x[:8] = [0.0, -0.0, np.inf, -np.inf, np.nan, 65504.0, 65520.0, 2.0 ** -25]
    for shift in (0, 3, 17, -4):
        c = native.codec_encode(x, shift)
        assert np.array_equal(c, codec.encode(x, shift))
        d = native.codec_decode(c, shift)
        assert np.array_equal(d, codec.decode(c, shift), equal_nan=True)
    for scale in (1e-6, 0.37, 1.0, 5.0, 3e4):
        y = (rng.standard_normal(1000) * scale).astype(np.float32)
        s = native.codec_shift(y)
        assert s == codec.component_shift(y)
        top = np.abs(y).max() * 2.0 ** s
        assert 2.0 ** 14 <= top < 2.0 ** 15
    assert native.codec_shift(np.zeros(5, np.float32)) == 0
    # relative error of a round trip: half-precision significand over 9 decades below the maximum
    y = (rng.uniform(1e-9, 1.0, 100000)).astype(np.float32)
    s = codec.component_shift(y)
    r = codec.decode(codec.encode(y, s), s)
    big = y > y.max() * 1e-8
    assert (np.abs(r - y)[big] / y[big]).max() <= 2.0 ** -11


def test_decoded_database_mirror():
    import cledb_b200.synth as synth
    db = synth.make_database(8, 39, ngx=3, nbphi=8, nbtheta=6).reshape(-1, 8).copy()
    db[5, 3] = np.nan
    dec = codec.decoded_database(db, 1)
    assert np.isnan(dec[5, 3]) and dec[5, 0] == 1 and (dec[5, [1, 2, 4, 5, 6, 7]] == 0).all()
    assert np.array_equal(codec.decoded_database(db, 0)[[0, 1, 2]], db[[0, 1, 2]])
    keep = ~np.isnan(db[:, 3])
    assert np.all(dec[keep, 0] == 1.0)
    assert np.all(np.sign(dec[keep][:, [1, 2, 4, 5, 6]]) == np.sign(db[keep][:, [1, 2, 4, 5, 6]]))      # zero and sign exact


def test_invproc_guards_need_no_gpu():
    """Fatal-input guards of cledb_invproc return idx = -1 everywhere before anything touches the device
    (reference CLEDB_PROC.py:166-178); blos_proc refuses IQUD data (CLEDB_PROC.py:426-428)."""
    import CLEDB_PROC.CLEDB_PROC as procinv
    import constants as consts
    import ctrlparams
    import cledb_b200.synth as synth
    db = synth.make_database(8, 39, ngx=3, nbphi=8, nbtheta=6)
    obs = synth.make_observation(3, 2, [db])
    hdr = synth.make_dbhdr(3, 8, 6)
    args = (obs['sobs_totrot'], obs['sobs_dopp'], [db], obs['db_enc'], obs['yobs'], obs['aobs'], obs['dobs'], obs['snr'],
            obs['issuemask'], hdr, obs['keyvals'])
    inv, sf = procinv.cledb_invproc(*args, 8, 1000, 0, True, 4, False, 0)          # iqud without bcalc == 3
    assert inv.shape == (3, 2, 8, 11) and (inv[..., 0] == -1).all() and not sf.any()
    inv, sf = procinv.cledb_invproc(*args, 8, 1000, 3, False, 4, False, 0)         # bcalc == 3 without iqud
    assert (inv[..., 0] == -1).all()
    bad = (obs['sobs_totrot'][..., :4],) + args[1:]
    inv, sf = procinv.cledb_invproc(*bad, 8, 1000, 0, False, 4, False, 0)          # one-line observation
    assert (inv[..., 0] == -1).all()
    from cledb_b200.native import CledbError
    with pytest.raises(CledbError, match='no CPU fallback'):
        procinv.cledb_invproc(*args, 8, 1000, 0, False, 4, True, 0)                # a real search needs the GPU: no CPU fallback
    params = ctrlparams.ctrlparams()
    params.iqud, params.verbose = True, 0
    assert procinv.blos_proc(obs['sobs_totrot'], obs['snr'], obs['issuemask'], obs['keyvals'], consts, params) == 0
    assert procinv.iss_block(512 + 32) == [0, 0, 0, 0, 0, 32, 0, 0, 0, 512]


def test_ctrlparams_and_constants_surface():
    import constants as consts
    import ctrlparams
    p = ctrlparams.ctrlparams()
    for name, val in (('nsearch', 8), ('maxchisq', 1000), ('bcalc', 0), ('reduced', False), ('iqud', False), ('ncpu', 100),
                      ('verbose', 2), ('integrated', False), ('dblinpolref', 0), ('jitdisable', True), ('gaussfit', 2)):
        assert getattr(p, name) == val
    assert p.lookuptb == './CLEDB_BUILD/chianti_v10.1_pycelp_fe13_h99_d120_ratio.npz'
    c = consts.Constants('fe-xiii_1074')
    assert c.g_eff == 1.5 and c.F_factor == 0.0 and c.line_ref == 1074.6153
    assert consts.Constants('fe-xiii_1079').g_eff == 1.5 and consts.Constants('si-x_1430').F_factor == 0.5
    k = 1.e9 * (c.planckconst * c.l_speed / c.bohrmagneton / (c.line_ref ** 2))
    assert abs(k - 185481.06) < 0.5


def test_out_of_scope_names_pass_through_to_the_reference(tmp_path, monkeypatch):
    """The drop-in modules shadow the reference's; names they do not define resolve to the reference's module by file path,
    names of the hot path never do."""
    import importlib
    ref = tmp_path / 'ref'
    (ref / 'CLEDB_PREPINV').mkdir(parents=True)
    (ref / 'CLEDB_PREPINV' / 'CLEDB_PREPINV.py').write_text(
        "def sobs_preprocess(a, b, c):\n    return ('from the reference', a)\n"
        "def calls_hot_path():\n    return obs_qurotate\n"
        "def obs_qurotate(*a):\n    raise RuntimeError('hot-path name must not be taken from the reference')\n")
    monkeypatch.setenv('CLEDB_REFERENCE_ROOT', str(ref))
    import cledb_b200.passthrough as pt
    pt._loaded.clear()
    prepinv = importlib.import_module('CLEDB_PREPINV.CLEDB_PREPINV')
    assert prepinv.sobs_preprocess(1, 2, 3) == ('from the reference', 1)
    assert prepinv.obs_qurotate.__module__ == 'CLEDB_PREPINV.CLEDB_PREPINV'
    assert prepinv.calls_hot_path() is prepinv.obs_qurotate      # reference code that calls a hot-path function gets the B200 one
    with pytest.raises(AttributeError):
        prepinv.no_such_function
    pt._loaded.clear()
    monkeypatch.setenv('CLEDB_REFERENCE_ROOT', str(tmp_path / 'missing'))
    with pytest.raises(AttributeError, match='CLEDB_REFERENCE_ROOT'):
        prepinv.sobs_preprocess
    pt._loaded.clear()
