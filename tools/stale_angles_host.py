import sys, time, ctypes, numpy as np
sys.path.insert(0,'/root/repo')
from tests.host import harness
from cvxpnpl_b200 import synth
def run(lib_path, d):
    harness._lib = ctypes.CDLL(lib_path)
    t0=time.time(); r = harness.solve(d); return r, time.time()-t0
base='/root/repo/tests/host/_build/'
for (n_pts,n_lines,B,noise,cop) in ((3,0,300,0.0,False),(0,3,300,0.0,False),(2,1,300,0.0,False),(8,0,300,0.0,True),(0,6,3000,1.0,False),(5,3,1000,1.0,False),(8,4,2000,2.0,False)):
    d = synth.make_batch(B, n_pts, n_lines, noise=noise, seed=11, coplanar=cop)
    a,ta = run(base+'libhost_harness.so', d)
    b,tb = run(base+'libhost_stale.so', d)
    ia, ib = a['iters'], b['iters']
    ok = (a['status']==0)&(b['status']==0)&(a['n_poses']==1)&(b['n_poses']==1)
    ang = synth.rotation_angle(a['R'][ok,0], b['R'][ok,0]) if ok.any() else np.zeros(1)
    print(f"{n_pts}+{n_lines} cop={cop}: exact mean {ia.mean():.1f} med {np.median(ia)} max {ia.max()} capped {(a['status']==1).sum()} | stale mean {ib.mean():.1f} med {np.median(ib)} max {ib.max()} capped {(b['status']==1).sum()} | status differ {(a['status']!=b['status']).sum()} nposes differ {(a['n_poses']!=b['n_poses']).sum()} both ok {ok.sum()} max ang {ang.max():.2e}  ({ta:.1f}s/{tb:.1f}s)", flush=True)
