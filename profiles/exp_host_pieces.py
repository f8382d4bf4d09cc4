import sys, time
sys.path[:0]=['solar-coronal-inversion_b200']
import numpy as np
from cledb_b200 import native, solution
def t(f, reps=400):
    f(); t0=time.perf_counter()
    for _ in range(reps): f()
    return (time.perf_counter()-t0)/reps*1e3
n=65536
rng=np.random.default_rng(0)
codes=rng.choice([0,0,0,0,0,0,0,1024,2048,512],n).astype(np.int32); codes[::30000]|=1
inv=np.zeros((n,8,11),np.float32); inv[:,:,3]=1
print('issue_codes new      %.4f ms' % t(lambda: native.issue_codes(codes.copy(), inv)))
def old_issue(codes, invout):
    amb = np.flatnonzero(codes & 1)
    if amb.size:
        codes = codes & ~np.int32(1)
        vv = solution.van_vleck_last(invout[amb])
        codes[amb] = (codes[amb] & ~np.int32(512)) | (vv.astype(np.int32) << 9)
    return codes
print('issue_codes old      %.4f ms' % t(lambda: old_issue(codes.copy(), inv)))
print('codes.copy           %.4f ms' % t(lambda: codes.copy()))
iss=np.zeros((256,256,2),np.int32); c2=codes.reshape(256,256) & ~np.int32(1)
print('apply new (int32)    %.4f ms' % t(lambda: solution.apply_issue_codes(iss, c2)))
def old_apply():
    iss[:, :, 0:2] |= c2[:, :, None]
print('apply old (int32)    %.4f ms' % t(old_apply))
issf=np.zeros((256,256,2),np.float64)
print('apply new (float64)  %.4f ms' % t(lambda: solution.apply_issue_codes(issf, c2), 100))
ang=rng.uniform(-3,3,n).astype(np.float32)
print('cos+sin              %.4f ms' % t(lambda: (np.cos(-2.*ang), np.sin(-2.*ang))))
