"""Compile the reference's OWN sources (where they lie under /root/reference) into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference); outputs are
git-ignored but travel to the GPU box with the snapshot.  No reference source is copied into the
repository: the compiler reads the files in place.

Built here:
  * cd_ref — losses/cuda/chamfer_distance/chamfer_distance.{cpp,cu}: the reference's Chamfer
    extension (CPU `forward/backward` = nnsearch, and its CUDA kernels `forward_cuda/
    backward_cuda` compiled as ordinary CUDA for sm_100).  The CPU entry points are the
    `"kind": "reference"` baseline and pin oracle/l3d_oracle.c's Chamfer restatement.
  * libpn2_ref.so — utils/lib/src/*_gpu.cu (pointnet2 kernels + launchers) behind ref_shims/pn2_shim.cu.
  * libemd_ref.so — losses/cuda/emd_torch/pkg/include/cuda/emd.cuh (EMD kernels + launchers) behind
    ref_shims/emd_shim.cu.
The extensions' OWN build files do not work any more (emd_torch: AT_CHECK / tensor.type() removed from
torch 2.11; utils/lib: THC removed): only their kernels/launchers are compiled, in place.
  * learning3d/ — the reference's Python package (+ three checkpoints) staged so that the GPU box can run
    the reference's OWN models beside the rebound ones (stage_python below).
"""
import os
import sys

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")


def _fresh(dst, srcs):
    """True when dst exists and is newer than every source (skip the multi-minute rebuild)."""
    if not os.path.exists(dst):
        return False
    t = os.path.getmtime(dst)
    return all(os.path.getmtime(f) <= t for f in srcs)


def build_chamfer():
    src0 = os.path.join(REF, "losses", "cuda", "chamfer_distance")
    if _fresh(os.path.join(OUT, "cd_ref.so"), [os.path.join(src0, "chamfer_distance.cpp"),
                                               os.path.join(src0, "chamfer_distance.cu"), __file__]):
        print("up to date: cd_ref.so")
        return
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    os.environ["CC"] = "/usr/bin/gcc"
    os.environ["CXX"] = "/usr/bin/g++"
    from torch.utils.cpp_extension import load
    src = os.path.join(REF, "losses", "cuda", "chamfer_distance")
    bdir = os.path.join(OUT, "cd_ref_build")
    os.makedirs(bdir, exist_ok=True)
    load(name="cd_ref", sources=[os.path.join(src, "chamfer_distance.cpp"),
                                 os.path.join(src, "chamfer_distance.cu")],
         build_directory=bdir, verbose=False)
    so = os.path.join(bdir, "cd_ref.so")
    assert os.path.exists(so), so
    dst = os.path.join(OUT, "cd_ref.so")
    if os.path.exists(dst):
        os.remove(dst)
    os.link(so, dst)
    print("built", dst)


def build_pointnet2():
    """utils/lib/src/{ball_query,group_points,sampling,interpolate}_gpu.cu — the reference's CUDA
    kernels and launchers, compiled in place with nvcc (their own setup.py needs THC, which torch
    2.11 no longer ships; the .cu files only need torch's headers on the include path) together
    with oracle/ref_shims/pn2_shim.cu (extern "C" wrappers, our own code)."""
    import subprocess
    import sysconfig
    from torch.utils.cpp_extension import include_paths
    src = os.path.join(REF, "utils", "lib", "src")
    dst = os.path.join(OUT, "libpn2_ref.so")
    srcs = [os.path.join(src, f + "_gpu.cu") for f in ("ball_query", "group_points", "sampling", "interpolate")]
    srcs.append(os.path.join(HERE, "ref_shims", "pn2_shim.cu"))
    if _fresh(dst, srcs + [__file__]):
        print("up to date: libpn2_ref.so")
        return
    cmd = ["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17",
           "-shared", "-Xcompiler", "-fPIC", "-w", "-I" + src, "-I" + sysconfig.get_paths()["include"]]
    cmd += ["-I" + p for p in include_paths()]
    cmd += ["-o", dst] + srcs
    env = dict(os.environ, CC="/usr/bin/gcc", CXX="/usr/bin/g++")
    subprocess.run(cmd, check=True, env=env)
    print("built", dst)


def build_emd():
    """losses/cuda/emd_torch/pkg/include/cuda/emd.cuh — the reference's EMD kernels + launchers, included in
    place by oracle/ref_shims/emd_shim.cu, which only swaps the dispatch macro that no longer compiles
    (tensor.type()) for a float-only one.  The extension's own host glue (emd.h: AT_CHECK) is not used."""
    import subprocess
    import sysconfig
    import torch
    from torch.utils.cpp_extension import include_paths
    inc = os.path.join(REF, "losses", "cuda", "emd_torch", "pkg", "include")
    shim = os.path.join(HERE, "ref_shims", "emd_shim.cu")
    dst = os.path.join(OUT, "libemd_ref.so")
    if _fresh(dst, [shim, os.path.join(inc, "cuda", "emd.cuh"), __file__]):
        print("up to date: libemd_ref.so")
        return
    tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
    cmd = ["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17",
           "-shared", "-Xcompiler", "-fPIC", "-w", "-I" + inc, "-I" + sysconfig.get_paths()["include"]]
    cmd += ["-I" + p for p in include_paths()]
    cmd += ["-o", dst, shim, "-L" + tlib, "-ltorch", "-ltorch_cpu", "-ltorch_cuda", "-lc10", "-lc10_cuda",
            "-Xlinker", "-rpath=" + tlib]
    subprocess.run(cmd, check=True, env=dict(os.environ, CC="/usr/bin/gcc", CXX="/usr/bin/g++"))
    print("built", dst)


def stage_python():
    """Stage the reference's own Python package + the three checkpoints the C3/C4/C5 parity tests load under
    oracle/_ref/learning3d/ (git-ignored like the .so files above, so it never enters the history, but it
    travels to the GPU box with the snapshot).  The GPU tests import it as `learning3d`, run the UNMODIFIED
    reference models on the B200 and compare them with the same models rebound to libl3d_b200.so
    (tests/test_gpu_reference_models.py).  Nothing under learning3d_b200/ reads it."""
    import shutil
    dst = os.path.join(OUT, "learning3d")
    n = 0
    for sub in ("models", "utils", "losses", "ops", "data_utils"):
        for root, _dirs, files in os.walk(os.path.join(REF, sub)):
            rel = os.path.relpath(root, REF)
            for f in files:
                if not f.endswith(".py"):
                    continue
                os.makedirs(os.path.join(dst, rel), exist_ok=True)
                d = os.path.join(dst, rel, f)
                s = os.path.join(root, f)
                if not os.path.exists(d) or os.path.getmtime(d) < os.path.getmtime(s):
                    shutil.copyfile(s, d)
                    n += 1
    for ck in ("exp_dcp/models/best_model.t7", "exp_flownet/models/model.best.t7", "exp_pcn/models/best_model.t7"):
        s = os.path.join(REF, "pretrained", ck)
        d = os.path.join(dst, "pretrained", ck)
        if os.path.exists(s) and not os.path.exists(d):
            os.makedirs(os.path.dirname(d), exist_ok=True)
            shutil.copyfile(s, d)
            n += 1
    print("staged reference python: %d file(s) updated under %s" % (n, dst))


def main():
    if not os.path.isdir(REF):
        print("no /root/reference here: keeping prebuilt oracle/_ref (if any)")
        return 0
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1:] or ["chamfer", "pointnet2", "emd", "python"]
    if "chamfer" in which:
        build_chamfer()
    if "pointnet2" in which:
        build_pointnet2()
    if "emd" in which:
        build_emd()
    if "python" in which:
        stage_python()
    return 0


if __name__ == "__main__":
    sys.exit(main())
