"""Build libdistegnn_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m distegnn_b200.build [--force] [--verbose]

The shared library is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
ROOT = os.path.dirname(PKG)
LIB = os.path.join(PKG, "libdistegnn_b200.so")
OBJ = os.path.join(CSRC, "build")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
          "-Xptxas", "-v", "--expt-relaxed-constexpr"]


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_mtime():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(ROOT, "include", "distegnn_b200.h"))
    return max(os.path.getmtime(h) for h in hdrs)


def _compile(src, verbose, objdir=None, defs=None):
    obj = os.path.join(objdir or OBJ, src[:-3] + ".o")
    extra = os.environ.get("DISTEGNN_NVCC_DEFS", "").split() + list(defs or [])   # e.g. "-DT16_CHUNK_UNROLL=4"
    cmd = [NVCC, *ARCH, *CFLAGS, *extra, "-c", os.path.join(CSRC, src), "-o", obj]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{p.stdout}\n{p.stderr}")
    log = p.stderr
    with open(obj + ".ptxas.log", "w") as f:
        f.write(log)
    if verbose:
        print(log)
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    srcs = sources()
    newest = max([os.path.getmtime(os.path.join(CSRC, s)) for s in srcs] + [_deps_mtime(),
                                                                          os.path.getmtime(__file__)])
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= newest:
        return LIB
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose), srcs))
    cmd = [NVCC, *ARCH, "-shared", "-o", LIB, *objs, "-cudart", "shared",
           "-Xlinker", "-rpath,/usr/local/cuda/lib64"]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError(f"link failed:\n{p.stdout}\n{p.stderr}")
    return LIB


def build_variant(tag: str, defs, verbose: bool = False) -> str:
    """A/B build: the same sources with extra -D flags -> distegnn_b200/variants/libdistegnn_b200.<tag>.so
    (select it at run time with DISTEGNN_B200_LIB=<path>; the variants travel to the GPU box like the main .so)."""
    vdir = os.path.join(PKG, "variants")
    objdir = os.path.join(vdir, "build_" + tag)
    os.makedirs(objdir, exist_ok=True)
    srcs = sources()
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose, objdir, defs), srcs))
    lib = os.path.join(vdir, f"libdistegnn_b200.{tag}.so")
    p = subprocess.run([NVCC, *ARCH, "-shared", "-o", lib, *objs, "-cudart", "shared",
                        "-Xlinker", "-rpath,/usr/local/cuda/lib64"], capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError(f"link failed:\n{p.stdout}\n{p.stderr}")
    return lib


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--variant", metavar="TAG", help="build an A/B variant; flags via --defs")
    ap.add_argument("--defs", default="", help="extra nvcc flags for --variant, e.g. --defs=-DDEGNN_SILU_MODE=0")
    a = ap.parse_args()
    if a.variant:
        print(build_variant(a.variant, a.defs.split(), a.verbose))
    else:
        print(build(a.force, a.verbose))
