"""l3d_knn_expansion_host at C2 (B=32, N=1024, k=20, pinned host buffers): wall time per call for the kernel-slice /
copy-chunk / widening-thread settings of capi.cu (each setting in a fresh process: the library reads them once).
Usage: python profiles/sweep_knn_host.py >> profiles/r02/knn_host_path_sweep.txt"""
import json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import ctypes, json, os, sys, time
sys.path.insert(0, %r)
import torch
from learning3d_b200 import _C
B, N, k = 32, 1024, 20
x = torch.rand(B, 3, N).pin_memory()
out = torch.empty(B, N, k, dtype=torch.int64).pin_memory()
L = _C.lib()
f = lambda: _C.check(L.l3d_knn_expansion_host(ctypes.c_void_p(x.data_ptr()), B, N, k, ctypes.c_void_p(out.data_ptr())))
for _ in range(20): f()
best = 1e9
for rep in range(5):
    t0 = time.perf_counter()
    for _ in range(200): f()
    best = min(best, (time.perf_counter() - t0) / 200)
print(json.dumps({"slices": os.environ.get("L3D_HOST_SLICES"), "d2h_chunks": os.environ.get("L3D_HOST_D2H_CHUNKS"),
                  "threads": os.environ.get("L3D_HOST_THREADS"), "pool": os.environ.get("L3D_HOST_POOL"), "graph": os.environ.get("L3D_HOST_GRAPH"), "us_per_call": round(best * 1e6, 1)}))
''' % ROOT

if __name__ == "__main__":
    # (kernel slices, copy chunks per slice, widening threads, 1 = persistent worker pool / 0 = OpenMP region per chunk)
    # last column: L3D_HOST_GRAPH (experiment, removed again: one cudaGraphLaunch of the captured sequence per call, 4-byte
    # flag copies instead of events — measured SLOWER, 106.0 vs 90.6 us: the call is bound by the copy-engine / kernel timeline,
    # not by the driver calls; profiles/r02/knn_host_path_sweep.txt)
    configs = [(2, 1, 16, 1, 0), (2, 1, 16, 1, 1), (2, 2, 16, 1, 1), (4, 1, 16, 1, 1), (4, 2, 16, 1, 1), (1, 4, 16, 1, 1),
               (1, 2, 16, 1, 1), (8, 1, 16, 1, 1), (2, 1, 8, 1, 1)]
    for s, c, t, pl, gr in configs:
        env = dict(os.environ, L3D_HOST_SLICES=str(s), L3D_HOST_D2H_CHUNKS=str(c), L3D_HOST_THREADS=str(t),
                   L3D_HOST_POOL=str(pl), L3D_HOST_GRAPH=str(gr))
        r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=300)
        print(r.stdout.strip() or r.stderr[-400:], flush=True)
