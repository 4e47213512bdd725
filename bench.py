#!/usr/bin/env python
"""bench.py — headline benchmark of the learning3d_b200 hot path (contract: see DESIGN.md §5).

    python bench.py --gpus N --steps K --warmup W          # our arm (CUDA, one rank per GPU)
    python bench.py --impl reference --gpus N ...          # reference CPU path on the host cores

A "step" is one pass of the fused pairwise-distance + kNN kernel over one batch of BASELINE
config C2 (B=32 clouds of N=1024 points, k=20 — the DGCNN graph).  Weak scaling: every rank owns
its own B=32 batch, no data-path collective.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

B_PER_GPU, N_PTS, K_NN = 32, 1024, 20
METRIC, UNIT = "point_pairs_per_sec", "pairs/s"
WORKLOAD = ("C2 DGCNN graph: fused pairwise-distance + top-k, B=%d clouds/GPU x N=%d pts, k=%d, fp32, "
            "int64 indices" % (B_PER_GPU, N_PTS, K_NN))
# compulsory bytes per query row = 12 B read + 8*k B int64 index write
ALG_BYTES_PER_ROW = 12 + 8 * K_NN
FLOP_PER_PAIR = 8


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the ~10 s oracle timing")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-graph", action="store_true", help="launch every step directly (no CUDA graph)")
    ap.add_argument("--profile", action="store_true",
                    help="for runs under ncu: no clock ramp, no CPU baseline, few secondary iterations")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


def cpu_baseline_port(seconds):
    """The C oracle (OpenMP over rows) on the host cores: a bounded sample of the same workload."""
    import numpy as np
    import oracle
    rng = np.random.default_rng(1234)
    x = rng.random((B_PER_GPU, 3, N_PTS), dtype=np.float32)
    oracle.knn_expansion(x, K_NN, mt=True)          # warm-up
    reps, t0 = 0, time.perf_counter()
    while True:
        oracle.knn_expansion(x, K_NN, mt=True)
        reps += 1
        el = time.perf_counter() - t0
        if el >= seconds:
            break
    return {"value": reps * B_PER_GPU * N_PTS * N_PTS / el, "unit": UNIT,
            "cores": oracle.num_threads(), "kind": "port",
            "sample": "oracle/l3d_oracle.c knn_expansion (OpenMP), B=%d N=%d k=%d batch repeated %d x (%.1f s)"
                      % (B_PER_GPU, N_PTS, K_NN, reps, el)}


def run_reference(args, rank):
    """Reference arm: the reference's own CPU implementation of knn() (torch matmul + topk,
    utils/model_common_utils.py:3-9, restated call-for-call in oracle/ref_torch.py) on all host
    threads.  Each step is one full B=32 batch (~0.1-0.2 s)."""
    if rank != 0:
        return
    import torch
    from oracle import ref_torch
    # "all the host threads it can use": torch's intra-op pool is sized to the thread count that runs
    # this workload fastest among {all logical CPUs, half, torch's default} (oversubscribing SMT
    # siblings slows matmul + topk down), decided by one untimed call each.
    torch.manual_seed(1234)
    x = torch.rand(B_PER_GPU, 3, N_PTS)
    ncpu = os.cpu_count() or 1
    cands = sorted({ncpu, max(1, ncpu // 2), torch.get_num_threads()}, reverse=True)
    best_t, cores = None, ncpu
    for c in cands:
        torch.set_num_threads(c)
        ref_torch.knn(x, K_NN)
        t0 = time.perf_counter()
        ref_torch.knn(x, K_NN)
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best_t, cores = dt, c
    torch.set_num_threads(cores)
    steps = max(1, min(args.steps, 200))
    for _ in range(max(1, min(args.warmup, 3))):
        ref_torch.knn(x, K_NN)
    t0 = time.perf_counter()
    for _ in range(steps):
        ref_torch.knn(x, K_NN)
    el = time.perf_counter() - t0
    value = steps * B_PER_GPU * N_PTS * N_PTS / el
    sample = "torch-CPU knn() restatement, %d steps of one B=%d N=%d k=%d batch, %d threads" % (
        steps, B_PER_GPU, N_PTS, K_NN, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": WORKLOAD, "device": "host CPU, %d threads" % cores},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args, rank, local_rank, world):
    import torch
    from learning3d_b200 import _C

    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    dist_on = world > 1
    if dist_on:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    lib = _C.lib()
    B, N, k = B_PER_GPU, N_PTS, K_NN

    # inputs+outputs cycled over a pool larger than L2 so no step finds its data in cache
    in_bytes, out_bytes = B * 3 * N * 4, B * N * k * 8
    pool = max(4, int(1.5 * 126e6 / (in_bytes + out_bytes)) + 1)
    torch.manual_seed(1234 + rank)
    xs = [torch.rand(B, 3, N, device=dev) for _ in range(pool)]
    outs = [torch.empty(B, N, k, dtype=torch.int64, device=dev) for _ in range(pool)]
    stream = torch.cuda.current_stream()
    sp = _C._P(stream.cuda_stream)
    null = _C._P(None)

    def step(i, sptr=None):
        j = i % pool
        rc = lib.l3d_knn_expansion(_C._P(xs[j].data_ptr()), B, N, k, _C._P(outs[j].data_ptr()), null,
                                   sptr if sptr is not None else sp)
        if rc:
            _C.check(rc, "knn")

    # clock ramp (untimed) + the W warm-up steps
    t0 = time.perf_counter()
    i = 0
    while not args.profile and time.perf_counter() - t0 < 0.3:
        for _ in range(50):
            step(i); i += 1
        torch.cuda.synchronize()
    for w in range(max(args.warmup, 3)):
        step(w)
    torch.cuda.synchronize()

    # The step is a ~30 us launch: capture one pass over the buffer pool (`pool` launches of our
    # kernel, nothing else) in a CUDA graph and replay it, so the timed region measures the kernel and
    # not the per-launch driver gap.  Steps that do not fill a whole replay are launched directly.
    graph, per_replay = None, pool
    if not args.no_graph and not args.profile and args.steps >= pool:
        side = torch.cuda.Stream()
        side.wait_stream(stream)
        with torch.cuda.stream(side):
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                cap = _C._P(torch.cuda.current_stream().cuda_stream)
                for s in range(per_replay):
                    step(s, cap)
        stream.wait_stream(side)
        for _ in range(3):
            graph.replay()
        torch.cuda.synchronize()
    replays = (args.steps // per_replay) if graph is not None else 0
    direct = args.steps - replays * per_replay

    sampler = ClockSampler(local_rank)
    sampler.start()
    if dist_on:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = _C.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(replays):
        graph.replay()
    for s in range(direct):
        step(s)
    e1.record(stream)
    torch.cuda.synchronize()
    # kernels of ours executed in the timed region: graph nodes replayed + direct launches
    launches = replays * per_replay + (_C.launch_count() - l0)
    ms = e0.elapsed_time(e1)
    if dist_on:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
        ms = float(t.item())
    sampler.stop_flag = True
    sampler.join(timeout=1.0)

    # ---- end-to-end through the host-buffer C ABI (pinned host memory, copies inside) -----
    e2e_steps = 3 if args.profile else max(10, min(args.steps, 200))
    hx = torch.rand(B, 3, N).pin_memory()
    hidx = torch.empty(B, N, k, dtype=torch.int64).pin_memory()
    for _ in range(3):
        _C.check(lib.l3d_knn_expansion_host(_C._P(hx.data_ptr()), B, N, k, _C._P(hidx.data_ptr())), "e2e")
    if dist_on:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        rc = lib.l3d_knn_expansion_host(_C._P(hx.data_ptr()), B, N, k, _C._P(hidx.data_ptr()))
        if rc:
            _C.check(rc, "e2e")
    e2e_s = time.perf_counter() - t0
    if dist_on:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())

    extra = {}
    try:
        extra.update(chamfer_bench(torch, dev, dist_on, world, 5 if args.profile else 200))
    except ImportError:
        pass
    if world == 1:
        extra.update(tensor_core_bench(torch, dev, 2 if args.profile else 30))

    if rank == 0:
        pairs_per_step = world * B * N * N
        value = pairs_per_step * args.steps / (ms * 1e-3)
        kernel_s = ms * 1e-3 / args.steps
        peak, peak_src = peaks()
        alg_bytes = B * N * ALG_BYTES_PER_ROW
        achieved = alg_bytes / kernel_s / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "parallelism": "batch-shard dp%d (no data-path collective)" % world,
                       "global_batch": world * B,
                       "launch": ("cuda-graph replay (%d launches per graph) + %d direct" % (per_replay, direct)
                                  if graph is not None else "direct launches"),
                       "l2": "inputs+outputs cycled over a %d-buffer pool = %.0f MB > 126 MB L2"
                             % (pool, pool * (in_bytes + out_bytes) / 1e6)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu_traffic(),
                         "peak_source": peak_src, "kernel": "l3d::knn_kernel<EXPANSION_NEG,KS=1>",
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "fp32_gflops_achieved": B * N * N * FLOP_PER_PAIR / kernel_s / 1e9,
                         "note": "fp32-issue/selection bound once the NxN matrix is not materialised "
                                 "(SURVEY.md §8d): a perfect kernel reaches ~20% of HBM peak at this shape"},
            "e2e": {"value": world * B * N * N * e2e_steps / e2e_s, "unit": UNIT,
                    "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": out_bytes,
                    "ms_per_step": 1e3 * e2e_s / e2e_steps,
                    "path": "l3d_knn_expansion_host (pinned host buffers, H2D + kernel + D2H + sync)"},
            "gpu_launches": launches,
            "clocks": sampler.summary(),
        }
        if extra:
            line["extra"] = extra
        if world == 1 and not args.no_cpu_baseline and not args.profile:
            line["cpu_baseline"] = cpu_baseline_port(args.cpu_seconds)
        print(json.dumps(line), flush=True)
    if dist_on:
        dist.destroy_process_group()


def ncu_traffic():
    """dram bytes per launch of the kNN kernel from the committed ncu capture (profiles/), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "knn_traffic.json")) as f:
            return json.load(f)["dram_bytes_per_launch"]
    except Exception:
        return None


def chamfer_bench(torch, dev, dist_on, world, iters=200):
    """Secondary metric of BASELINE.json: Chamfer fwd+bwd clouds/s (config C1 shape per GPU and a
    B=32 batch), through the public ChamferDistanceLoss API."""
    from learning3d_b200.losses import ChamferDistanceLoss
    crit = ChamferDistanceLoss()
    out = {}
    for B in (4, 32):
        a = torch.rand(B, 1024, 3, device=dev, requires_grad=True)
        b = torch.rand(B, 1024, 3, device=dev, requires_grad=True)
        for _ in range(min(10, iters)):
            a.grad = b.grad = None
            crit(a, b).backward()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            a.grad = b.grad = None
            crit(a, b).backward()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        if dist_on:
            import torch.distributed as dist
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        out["chamfer_fwd_bwd_clouds_per_sec_B%d" % B] = world * B / (ms * 1e-3)
        out["chamfer_fwd_bwd_ms_B%d" % B] = ms
    return out


def tensor_core_bench(torch, dev, iters=30):
    """Tensor-core rows of the path (per GPU, not aggregated): DCP SVD-head front half at C3
    (B=32, d_k=512, N=1024; svd.py:23-28 fused) and the feature-space kNN graph (B=32, C=64, N=1024, k=20)."""
    from learning3d_b200.utils import knn
    from learning3d_b200.utils.svd import soft_correspondence
    out = {}
    es = torch.randn(32, 512, 1024, device=dev)
    et = torch.randn(32, 512, 1024, device=dev)
    tg = torch.rand(32, 3, 1024, device=dev)
    xf = torch.randn(32, 64, 1024, device=dev)
    for name, fn, flop in (("svd_head_front_C3", lambda: soft_correspondence(es, et, tg), 2.0 * 32 * 1024 * 1024 * 512),
                           ("knn_features_C64", lambda: knn(xf, 20), None)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        out[name + "_us"] = ms * 1e3
        if flop:
            out[name + "_fp32_equiv_tflops"] = flop / (ms * 1e-3) / 1e12
            # tensor-pipe view: 3xTF32 issues three TF32 MMAs per fp32-equivalent product; the TF32 peak is
            # taken as half of the measured dense bf16 cuBLAS throughput (MEASURED_PEAKS.json)
            issued = 3.0 * flop / (ms * 1e-3) / 1e12
            out[name + "_issued_tf32_tflops"] = issued
            try:
                with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                    out[name + "_frac_of_measured_tf32_peak"] = issued / (float(json.load(f)["bf16_tflops"]) / 2.0)
            except Exception:
                pass
    return out


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
