"""ORACLE — TEST INFRASTRUCTURE ONLY.

Compiles the reference's OWN pointnet2 CUDA extension, from the sources where
they lie under /root/reference (nothing is copied into this repo), for sm_100
into oracle/_ref/ (git-ignored, travels to the GPU box with the snapshot):

  oracle/_ref/pointnet2_ref_ext.so          nvcc defaults (-fmad=true): the oracle of record
  oracle/_ref/pointnet2_ref_ext_nofmad.so   -fmad=false: separates algorithmic from
                                            FMA-contraction mismatches

The reference's setup.py (modules/third_party/pointnet2/setup.py:17) cannot be
used as written: its arch list starts at compute_37, which CUDA 12.9 dropped.
The sources themselves build unmodified against torch 2.11.

Used as the GPU-side checker in tests (`-m gpu`) and to generate
tests/golden/ext_refcuda_*.npz (tests/golden/gen_ext_refcuda.py).
"""
import glob
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("MSR3D_REFERENCE_ROOT", "/root/reference")
SRC = os.path.join(REF, "modules", "third_party", "pointnet2", "_ext_src")
OUT = os.path.join(HERE, "_ref")


def build_one(name, extra):
    from torch.utils.cpp_extension import load
    bdir = os.path.join(OUT, "_build_" + name)
    os.makedirs(bdir, exist_ok=True)
    sources = sorted(glob.glob(os.path.join(SRC, "src", "*.cpp")) +
                     glob.glob(os.path.join(SRC, "src", "*.cu")))
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    load(name=name, sources=sources, extra_include_paths=[os.path.join(SRC, "include")],
         extra_cflags=["-O2"],
         extra_cuda_cflags=["-O3", "-gencode", "arch=compute_100,code=sm_100"] + extra,
         build_directory=bdir, verbose=False, is_python_module=False)
    shutil.copy(os.path.join(bdir, name + ".so"), os.path.join(OUT, name + ".so"))
    shutil.rmtree(bdir, ignore_errors=True)
    print("built", os.path.join(OUT, name + ".so"))


def main():
    if not os.path.isdir(SRC):
        print(f"reference sources not found at {SRC}; keeping any prebuilt oracle/_ref", file=sys.stderr)
        return 0
    os.makedirs(OUT, exist_ok=True)
    build_one("pointnet2_ref_ext", [])
    build_one("pointnet2_ref_ext_nofmad", ["-fmad=false"])
    return 0


if __name__ == "__main__":
    sys.exit(main())
