"""Build libvgl_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m vectorgraphlibrary_b200.build [--force] [--verbose]

One object per csrc/*.cu (compiled in parallel), linked into vectorgraphlibrary_b200/libvgl_b200.so. The shared
object travels to the GPU box with the repo snapshot (it is git-ignored, not gpurun-ignored).
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libvgl_b200.so")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "--expt-relaxed-constexpr", "--extended-lambda", "-Xcompiler", "-fPIC,-O2,-fopenmp",
    "-ccbin", "/usr/bin/g++", "-I", os.path.join(ROOT, "include"), "-I", CSRC,
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libvgl_b200 cannot be built (there is no CPU fallback)")
    return exe


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    sources = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    headers = sorted(glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(ROOT, "include", "*.h"))
                     + glob.glob(os.path.join(ROOT, "include", "vgl_b200", "*")))
    nvcc = _nvcc()
    jobs = []
    for src in sources:
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        if force or _stale(obj, [src] + headers):
            cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            jobs.append((src, cmd))

    def run(job):
        src, cmd = job
        p = subprocess.run(cmd, capture_output=True, text=True)
        return src, p.returncode, p.stdout + p.stderr

    with cf.ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for src, rc, out in ex.map(run, jobs):
            if verbose or rc != 0:
                sys.stderr.write(f"--- {os.path.basename(src)}\n{out}\n")
            if rc != 0:
                raise RuntimeError(f"nvcc failed on {src}")
    objs = [os.path.join(OBJ, os.path.basename(s)[:-3] + ".o") for s in sources]
    if force or jobs or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-ccbin", "/usr/bin/g++", "-o", LIB] + objs + ["-Xcompiler", "-fopenmp", "-lcudart_static", "-ldl", "-lrt", "-lpthread"]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            sys.stderr.write(p.stdout + p.stderr)
            raise RuntimeError("link of libvgl_b200.so failed")
    return LIB


SHIM_SRC = os.path.join(ROOT, "tests", "shim", "shim_algorithms.cu")
SHIM_BIN = os.path.join(ROOT, "tests", "shim", "shim_algorithms")


def build_shim_test(force: bool = False) -> str:
    """The C++ host side above the C ABI: tests/shim/shim_algorithms.cu instantiates the user-lambda operator API
    (include/vgl_b200/graph_abstractions_b200.cuh) in its own translation unit, exactly as a VGL algorithm source
    would, and links libvgl_b200.so. Built here so the binary travels to the GPU box with the snapshot."""
    headers = sorted(glob.glob(os.path.join(ROOT, "include", "*.h")) + glob.glob(os.path.join(ROOT, "include", "vgl_b200", "*")))
    if force or _stale(SHIM_BIN, [SHIM_SRC, LIB] + headers):
        cmd = [_nvcc(), "-O2", "-std=c++17", "--extended-lambda", "--expt-relaxed-constexpr", "-gencode",
               "arch=compute_100a,code=sm_100a", "-lineinfo", "-ccbin", "/usr/bin/g++", "-I", os.path.join(ROOT, "include"),
               SHIM_SRC, "-L", HERE, "-lvgl_b200", "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN/../../vectorgraphlibrary_b200",
               "-o", SHIM_BIN]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            sys.stderr.write(p.stdout + p.stderr)
            raise RuntimeError("nvcc failed on tests/shim/shim_algorithms.cu")
    return SHIM_BIN


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(a.force, a.verbose))
    print(build_shim_test(a.force))
