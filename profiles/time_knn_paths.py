"""Device time of knn() on xyz clouds per kernel: the warp-per-row-pair kernel of knn.cu (path 1) against the
thread-per-row kernel of knn_tpr.cu (path 2), CUDA-graph replay of 20 launches over a buffer pool larger than L2.
Usage: python profiles/time_knn_paths.py  > profiles/r02/knn_paths_time.txt"""
import json
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from learning3d_b200 import _C


def time_path(B, N, k, path, feat=False, reps=20, rounds=30):
    dev = torch.device("cuda:0")
    per = B * N * (12 + 8 * k + (24 * k if feat else 0))
    pool = max(2, min(64, int(200e6 // per) + 1))
    xs = [torch.rand(B, 3, N, device=dev) for _ in range(pool)]
    idxs = [torch.empty(B, N, k, dtype=torch.int64, device=dev) for _ in range(pool)]
    feats = [torch.empty(B, 6, N, k, device=dev) for _ in range(pool)] if feat else None
    L = _C.lib()
    L.l3d_debug_knn_path(path)
    s = torch.cuda.Stream()
    def launch(i):
        if feat:
            _C.check(L.l3d_knn_graph_feature(_C.ptr(xs[i]), B, N, k, _C.ptr(idxs[i]), _C.ptr(feats[i]), s.cuda_stream))
        else:
            _C.check(L.l3d_knn_expansion(_C.ptr(xs[i]), B, N, k, _C.ptr(idxs[i]), None, s.cuda_stream))
    with torch.cuda.stream(s):
        for i in range(3):
            launch(i % pool)
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for i in range(reps):
                launch(i % pool)
        g.replay(); s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(rounds):
            g.replay()
        e1.record(s)
        s.synchronize()
    L.l3d_debug_knn_path(0)
    return e0.elapsed_time(e1) * 1e3 / (reps * rounds)


if __name__ == "__main__":
    torch.rand(1, device="cuda:0")
    quick = len(sys.argv) > 1 and sys.argv[1] in ("quick", "phases")
    phases = len(sys.argv) > 1 and sys.argv[1] == "phases"
    shapes = [(32, 1024, 20), (4, 1024, 20), (64, 1024, 20), (32, 2048, 20)] if quick else [
        (32, 1024, 20), (64, 1024, 20), (16, 1024, 20), (8, 1024, 20), (4, 1024, 20), (32, 1024, 16),
        (32, 512, 20), (16, 2048, 20), (32, 2048, 20)]
    print(json.dumps({"lib": os.environ.get("L3D_B200_LIB", "libl3d_b200.so")}), flush=True)
    for (B, N, k) in shapes:
        row = {"B": B, "N": N, "k": k}
        for path, name in ((1, "warp_pair_us"), (2, "thread_per_row_us")):
            if phases and path == 1:
                continue
            row[name] = round(time_path(B, N, k, path), 2)
        row["pairs_per_s_tpr"] = B * N * N / (row["thread_per_row_us"] * 1e-6)
        print(json.dumps(row), flush=True)
    for (B, N, k) in ([] if quick else [(32, 1024, 20)]):
        row = {"B": B, "N": N, "k": k, "fused_graph_feature": True}
        for path, name in ((1, "warp_pair_us"), (2, "thread_per_row_us")):
            row[name] = round(time_path(B, N, k, path, feat=True), 2)
        print(json.dumps(row), flush=True)
