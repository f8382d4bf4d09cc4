"""NumPy mirror of the int16 database codec of libcledb_b200.so (cledb_codec_* in include/cledb_b200.h).

    code   = IEEE binary16 bit pattern of  x * 2^shift   (round to nearest even), stored as int16
    decode = float32(binary16) * 2^-shift                 (exact in float32)

One shift per Stokes component and database: the largest finite |x| over the candidate rows lands in
[2^14, 2^15).  Relative step 2^-11 = 4.9e-4 over 9 decades below the component's maximum, absolute step
2^-24 / 2^shift below that; zero is exact, sign is exact.  (The reference's own int16 scheme, sdb_dcompress at
CLEDB_PREPINV.py:1618-1644, is a 15-decade log scale with a 1.05e-3 step that cannot represent 0 or -1 and is
disabled upstream, TODO.md:15.)

``decoded_database`` returns the database exactly as the kernels see it; parity tests hand THAT array to the
oracle, because quantisation (not arithmetic) changes which entries win.
"""
import numpy as np


def component_shift(x):
    x = np.abs(np.asarray(x, dtype=np.float32))
    x = x[np.isfinite(x)]
    mx = float(x.max()) if x.size else 0.0
    if mx == 0.0:
        return 0
    _, e = np.frexp(np.float32(mx))
    return int(max(-100, min(100, 15 - int(e))))


def encode(x, shift):
    return np.ldexp(np.asarray(x, dtype=np.float32), shift).astype(np.float32).astype(np.float16).view(np.int16)


def decode(codes, shift):
    return np.ldexp(np.asarray(codes, dtype=np.int16).view(np.float16).astype(np.float32), -shift).astype(np.float32)


def decoded_database(db, encoding):
    """[N,8] float32 as seen by the search kernels: rows with NaN V1 are never candidates and come back as
    [1,0,0,NaN,0,0,0,0]; with ``encoding == 1`` components 1..7 go through the int16 codec."""
    flat = np.ascontiguousarray(db, dtype=np.float32).reshape(-1, 8)
    out = flat.copy()
    dropped = np.isnan(flat[:, 3])
    if encoding == 1:
        kept = ~dropped
        for c in range(1, 8):
            s = component_shift(flat[kept, c])
            out[:, c] = decode(encode(flat[:, c], s), s)
    out[dropped] = np.array([1, 0, 0, np.nan, 0, 0, 0, 0], dtype=np.float32)
    return out
