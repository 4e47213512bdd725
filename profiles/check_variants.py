import os, sys, subprocess, glob
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE=r'''
import sys, numpy as np, torch
sys.path.insert(0, %r)
from learning3d_b200.utils import knn
import oracle
rng=np.random.default_rng(0)
ok=True
for (B,N,k) in [(4,1024,20),(2,2500,24),(3,333,7)]:
    x=rng.random((B,3,N),dtype=np.float32)
    ok &= np.array_equal(knn(torch.from_numpy(x).cuda(),k).cpu().numpy(), oracle.knn_expansion(x,k))
x=np.tile(rng.random((1,3,16),dtype=np.float32),(1,1,64))
ok &= np.array_equal(knn(torch.from_numpy(x).cuda(),20).cpu().numpy(), oracle.knn_expansion(x,20))
print("parity", ok)
''' % ROOT
for lib in sorted(glob.glob(ROOT+'/profiles/variants/lib_*.so')):
    out=subprocess.run([sys.executable,'-c',CODE],env=dict(os.environ,L3D_B200_LIB=lib),capture_output=True,text=True)
    print(os.path.basename(lib), out.stdout.strip() or out.stderr[-400:])
