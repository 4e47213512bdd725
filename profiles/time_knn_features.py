"""Device time of knn() on C-channel features (l3d_knn_features: |x|^2, Gram tiles + keys on tcgen05, selection) at
PRNet's call sequence shapes (models/prnet.py:78-90: C = 64 / 64 / 128, B=32, N=1024, k=20), per launch and in total.
Usage: python profiles/time_knn_features.py >> profiles/r02/knn_features_time.txt"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from learning3d_b200.utils import knn


def ev_us(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n


if __name__ == "__main__":
    dev = torch.device("cuda:0")
    for (B, C, N, k) in [(32, 64, 1024, 20), (32, 128, 1024, 20), (16, 64, 2048, 20), (32, 64, 512, 16)]:
        x = torch.relu(torch.randn(B, C, N, device=dev))
        us = ev_us(lambda: knn(x, k))
        torch.backends.cuda.matmul.allow_tf32 = False
        def ref():
            inner = -2 * torch.matmul(x.transpose(2, 1), x)
            xx = torch.sum(x ** 2, dim=1, keepdim=True)
            return (-xx - inner - xx.transpose(2, 1)).topk(k=k, dim=-1)[1]
        print(json.dumps({"B": B, "C": C, "N": N, "k": k, "l3d_knn_features_us": round(us, 1),
                          "reference_torch_ops_same_gpu_us": round(ev_us(ref, 10), 1)}), flush=True)
