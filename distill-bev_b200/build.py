"""Build recipe for libdistill_bev_b200.so (nvcc, sm_100a only, in-tree).

    python distill-bev_b200/build.py [--force] [--verbose]

Every .cu under csrc/ is compiled to an object with
``-gencode arch=compute_100a,code=sm_100a -lineinfo`` and linked into
``distill-bev_b200/lib/libdistill_bev_b200.so`` (plain C-ABI, cudart linked
statically, no torch dependency). nvcc cross-compiles without a GPU, so this
runs in the CPU-only build container; the .so travels to the GPU box.
"""
import argparse
import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ_DIR = os.path.join(HERE, "build")
LIB_DIR = os.path.join(HERE, "lib")
LIB_NAME = "libdistill_bev_b200.so"
LIB_PATH = os.path.join(LIB_DIR, LIB_NAME)

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; cannot build the sm_100a extension")
    return exe


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(ROOT, "include", "distill_bev_b200.h"))
    return hs


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile_one(src, force, verbose):
    obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
    if not force and not _stale(obj, [src] + _headers() + [os.path.abspath(__file__)]):
        return obj, ""
    cmd = [_nvcc()] + NVCC_FLAGS + ["-I", os.path.join(ROOT, "include"), "-c", src, "-o", obj]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, p.stdout, p.stderr))
    log = p.stdout + p.stderr
    with open(obj[:-2] + ".ptxas.log", "w") as f:
        f.write(log)
    if verbose:
        print(log)
    return obj, log


def build(force=False, verbose=False):
    """Compile and link; returns the path of the shared library."""
    os.makedirs(OBJ_DIR, exist_ok=True)
    os.makedirs(LIB_DIR, exist_ok=True)
    srcs = _sources()
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(lambda s: _compile_one(s, force, verbose), srcs))
    objs = [o for o, _ in results]
    if force or _stale(LIB_PATH, objs):
        tmp_path = LIB_PATH + ".tmp%d" % os.getpid()      # link aside, then rename: the library on disk is always whole
        cmd = [_nvcc(), "-shared", "-o", tmp_path] + objs + [
            "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static",
            "-Xcompiler", "-fPIC",
        ]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (p.stdout, p.stderr))
        os.replace(tmp_path, LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(a.force, a.verbose))
    sys.exit(0)
