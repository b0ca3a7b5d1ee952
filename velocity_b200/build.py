"""Builds velocity_b200/libvelocity_b200.so (the C-ABI library of include/velocity_b200.h) in-tree.

    python -m velocity_b200.build [--force] [--verbose]

nvcc cross-compiles for sm_100a without a GPU.  The .so is git-ignored but travels to the GPU box
with the repo snapshot.
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libvelocity_b200.so")
OBJ_DIR = os.path.join(HERE, "csrc", "_obj")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC,-fvisibility=hidden", "--expt-relaxed-constexpr",
]
LINK_LIBS = []                                  # the shipped library links no vendor BLAS / solver
VENDOR_FLAGS, VENDOR_LIBS = ["-DVEL_WITH_VENDOR_SOLVER"], ["-lcublas", "-lcusolver"]   # --vendor-solver: A/B build


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _deps():
    return _sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(HERE, "..", "include", "velocity_b200.h")]


def build(force=False, verbose=False, vendor=False):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ_DIR, exist_ok=True)
    hdr_mtime = max(os.path.getmtime(p) for p in _deps() if not p.endswith(".cu"))
    objs, rebuilt = [], False
    procs = []
    for src in _sources():
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_mtime):
            cmd = [nvcc] + NVCC_FLAGS + (VENDOR_FLAGS if vendor else []) + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
            rebuilt = True
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed on %s" % src)
    if rebuilt or not os.path.exists(LIB):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-L/usr/local/cuda/lib64"] + LINK_LIBS + (VENDOR_LIBS if vendor else [])
        subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    vendor = "--vendor-solver" in sys.argv
    print(build(force="--force" in sys.argv or vendor, verbose="--verbose" in sys.argv, vendor=vendor))
