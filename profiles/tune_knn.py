"""Time the C2 kNN launch for every library variant under profiles/variants/ (one subprocess each)."""
import glob, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys, torch
sys.path.insert(0, %r)
from learning3d_b200 import _C
lib = _C.lib()
B, N, k = 32, 1024, 20
pool = 34
xs = [torch.rand(B, 3, N, device="cuda") for _ in range(pool)]
outs = [torch.empty(B, N, k, dtype=torch.int64, device="cuda") for _ in range(pool)]
s = _C.stream()
def step(i):
    j = i %% pool
    lib.l3d_knn_expansion(_C.ptr(xs[j]), B, N, k, _C.ptr(outs[j]), None, s)
for i in range(3000): step(i)
torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(2000): step(i)
    e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 2000 * 1e3)
print("%%.2f" %% best)
''' % ROOT
for lib in sorted(glob.glob(os.path.join(ROOT, "profiles", "variants", "lib_*.so"))):
    env = dict(os.environ, L3D_B200_LIB=lib)
    out = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True)
    print(os.path.basename(lib), out.stdout.strip() or out.stderr[-300:], "us", flush=True)
