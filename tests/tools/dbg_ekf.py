import sys, os, ctypes, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import ekf_script, ekf_common as C
from hybvio_b200 import capi
from oracle import ekf_oracle
hv = capi.Context(0)
def dp():
    p = capi.EkfParams(); capi.load().hv_ekf_default_params(ctypes.byref(p)); return p
for trail, frames, nlist in ((6, 8, (8,20)), (20, 8, (8,20,40,84))):
    p = C.params_with(dp, trail)
    a, b = capi.Ekf(hv, p), ekf_oracle.OracleEKF(p)
    sa, sb, ca, cb = [], [], [], []
    ta = ekf_script.run_frames(a, frames=frames, n_list=nlist, snapshots=sa, checks=ca)
    tb = ekf_script.run_frames(b, frames=frames, n_list=nlist, snapshots=sb, checks=cb)
    print('status eq', [c[0] for c in ca] == [c[0] for c in cb])
    for i,((ma,Pa),(mb,Pb)) in enumerate(zip(sa,sb)):
        print(trail, 'frame', i, 'dm %.2e' % np.abs(ma-mb).max(), 'relP %.2e' % ekf_script.rel_err(Pa,Pb), 'maxP %.2e' % np.abs(Pb).max(), 'asym %.1e' % np.abs(Pa-Pa.T).max())
    ma_, mb_ = [], []
    ekf_script.run_misc_ops(a, ma_, ta); ekf_script.run_misc_ops(b, mb_, tb)
    for i,((ma,Pa),(mb,Pb)) in enumerate(zip(ma_,mb_)):
        print(trail, 'misc', i, 'dm %.2e' % np.abs(ma-mb).max(), 'relP %.2e' % ekf_script.rel_err(Pa,Pb), 'maxP %.2e' % np.abs(Pb).max())
