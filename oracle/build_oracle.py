"""Build the oracle's native parts (checker only, never shipped in the product).

    python oracle/build_oracle.py

1. gcc: oracle/c/*.c -> oracle/_build/liboracle.so (plain C restatement).
2. When /root/reference exists (build container), compile the reference's OWN
   CPU voxelization sources, unmodified and where they lie, into
   oracle/_ref/ref_voxel_layer*.so (a torch C++ extension without WITH_CUDA):
       mmdet3d/ops/voxel/src/voxelization.cpp
       mmdet3d/ops/voxel/src/voxelization_cpu.cpp
       mmdet3d/ops/voxel/src/scatter_points_cpu.cpp
3. Likewise the reference's vendored spconv extension (all 7 sources of
   mmdet3d/ops/spconv/src with mmdet3d/ops/spconv/include; its CPU and CUDA functors
   live in one module, so nvcc compiles the .cu files for sm_100) into
   oracle/_ref/ref_sparse_conv_ext*.so. Its CPU path pins oracle/spconv_oracle.py.
   oracle/_ref/ is git-ignored but travels to the GPU box with the snapshot.
4. The reference's CUDA extensions for bev_pool (bev_pool.cpp + bev_pool_cuda.cu) and voxelization
   (voxelization*.cpp/.cu, scatter_points*.cpp/.cu, -DWITH_CUDA), unmodified, for sm_100:
   oracle/_ref/ref_bev_pool_ext*.so, ref_voxel_layer_cuda*.so - the kernels BASELINE.md §4 says to beat, timed beside
   ours by tools/microbench_reference_cuda.py.
   No reference source is copied into this repository.
"""
import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "_build")
REF_OUT = os.path.join(HERE, "_ref")
REF_SRC = "/root/reference/mmdet3d/ops/voxel/src"


def build_c():
    os.makedirs(BUILD, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(HERE, "c", "*.c")))
    out = os.path.join(BUILD, "liboracle.so")
    if os.path.exists(out) and all(os.path.getmtime(s) <= os.path.getmtime(out) for s in srcs):
        return out
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-std=c99", "-ffp-contract=off", "-o", out] + srcs + ["-lm"]
    subprocess.run(cmd, check=True)
    return out


def build_ref():
    if not os.path.isdir(REF_SRC):
        return None
    os.makedirs(REF_OUT, exist_ok=True)
    existing = glob.glob(os.path.join(REF_OUT, "ref_voxel_layer*.so"))
    if existing:
        return existing[0]
    from torch.utils.cpp_extension import load
    tmp = os.path.join(REF_OUT, "_jit")
    os.makedirs(tmp, exist_ok=True)
    srcs = [os.path.join(REF_SRC, f) for f in
            ("voxelization.cpp", "voxelization_cpu.cpp", "scatter_points_cpu.cpp")]
    load(name="ref_voxel_layer", sources=srcs, build_directory=tmp, extra_cflags=["-O2", "-w"],
         verbose=False)
    so = glob.glob(os.path.join(tmp, "ref_voxel_layer*.so"))[0]
    dst = os.path.join(REF_OUT, os.path.basename(so))
    shutil.copy(so, dst)
    shutil.rmtree(tmp, ignore_errors=True)
    return dst


def build_ref_spconv():
    src_dir = "/root/reference/mmdet3d/ops/spconv"
    if not os.path.isdir(src_dir):
        return None
    os.makedirs(REF_OUT, exist_ok=True)
    existing = glob.glob(os.path.join(REF_OUT, "ref_sparse_conv_ext*.so"))
    if existing:
        return existing[0]
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    from torch.utils.cpp_extension import load
    tmp = os.path.join(REF_OUT, "_jit_spconv")
    os.makedirs(tmp, exist_ok=True)
    srcs = [os.path.join(src_dir, "src", f) for f in
            ("all.cc", "reordering.cc", "reordering_cuda.cu", "indice.cc", "indice_cuda.cu",
             "maxpool.cc", "maxpool_cuda.cu")]
    load(name="ref_sparse_conv_ext", sources=srcs,
         extra_include_paths=[os.path.join(src_dir, "include")],
         extra_cflags=["-O2", "-w", "-DWITH_CUDA", "-std=c++17"],
         extra_cuda_cflags=["-w", "-DWITH_CUDA", "-std=c++17"], build_directory=tmp, verbose=False)
    so = glob.glob(os.path.join(tmp, "ref_sparse_conv_ext*.so"))[0]
    dst = os.path.join(REF_OUT, os.path.basename(so))
    shutil.copy(so, dst)
    shutil.rmtree(tmp, ignore_errors=True)
    return dst


def _build_ref_cuda(name, srcs, extra=()):
    """One of the reference's CUDA extensions, UNMODIFIED sources compiled where they lie for sm_100 (nvcc cross-compiles
    without a GPU): the "kernels to beat" of BASELINE.md §4, timed on the GPU box by tools/microbench_reference_cuda.py.
    Checker / evidence only - never imported by the product path."""
    if not all(os.path.exists(f) for f in srcs):
        return None
    os.makedirs(REF_OUT, exist_ok=True)
    existing = glob.glob(os.path.join(REF_OUT, name + "*.so"))
    if existing:
        return existing[0]
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    from torch.utils.cpp_extension import load
    tmp = os.path.join(REF_OUT, "_jit_" + name)
    os.makedirs(tmp, exist_ok=True)
    load(name=name, sources=srcs, extra_cflags=["-O2", "-w", "-DWITH_CUDA"] + list(extra),
         extra_cuda_cflags=["-w", "-DWITH_CUDA"] + list(extra), build_directory=tmp, verbose=False)
    so = glob.glob(os.path.join(tmp, name + "*.so"))[0]
    dst = os.path.join(REF_OUT, os.path.basename(so))
    shutil.copy(so, dst)
    shutil.rmtree(tmp, ignore_errors=True)
    return dst


def build_ref_bev_pool_cuda():
    d = "/root/reference/mmdet3d/ops/bev_pool/src"
    return _build_ref_cuda("ref_bev_pool_ext", [os.path.join(d, "bev_pool.cpp"), os.path.join(d, "bev_pool_cuda.cu")])


def build_ref_voxel_cuda():
    return _build_ref_cuda("ref_voxel_layer_cuda", [os.path.join(REF_SRC, f) for f in (
        "voxelization.cpp", "voxelization_cpu.cpp", "scatter_points_cpu.cpp", "voxelization_cuda.cu", "scatter_points_cuda.cu")])


if __name__ == "__main__":
    print(build_c())
    try:
        print(build_ref())
        print(build_ref_spconv())
        print(build_ref_bev_pool_cuda())
        print(build_ref_voxel_cuda())
    except Exception as e:  # the reference build is optional evidence, not a dependency
        print("reference CPU build unavailable: %s" % e)
        sys.exit(0)
