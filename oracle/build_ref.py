"""Compile the reference's OWN sources (where they lie under /root/reference) into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference); outputs are
git-ignored but travel to the GPU box with the snapshot.  No reference source is copied into the
repository: the compiler reads the files in place.

Built here:
  * cd_ref — losses/cuda/chamfer_distance/chamfer_distance.{cpp,cu}: the reference's Chamfer
    extension (CPU `forward/backward` = nnsearch, and its CUDA kernels `forward_cuda/
    backward_cuda` compiled as ordinary CUDA for sm_100).  The CPU entry points are the
    `"kind": "reference"` baseline and pin oracle/l3d_oracle.c's Chamfer restatement.
Not buildable (recorded in DESIGN.md): losses/cuda/emd_torch (AT_CHECK / tensor.type() removed
from torch 2.11) and utils/lib (THC removed) through their own build files.
"""
import os
import sys

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")


def build_chamfer():
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


def main():
    if not os.path.isdir(REF):
        print("no /root/reference here: keeping prebuilt oracle/_ref (if any)")
        return 0
    os.makedirs(OUT, exist_ok=True)
    build_chamfer()
    return 0


if __name__ == "__main__":
    sys.exit(main())
