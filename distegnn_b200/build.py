"""Build the CUDA libraries in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m distegnn_b200.build [--force] [--verbose]

  libdistegnn_b200.so          the product: every entry point of include/distegnn_b200.h (csrc/*.cu)
  libdistegnn_b200_testing.so  cross-check twins of include/distegnn_b200_testing.h (csrc/testing/*.cu); loaded only by
                               tests/twin_backend.py and the A/B scripts, never by the package

The shared libraries are git-ignored but travel to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import os
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
TSRC = os.path.join(CSRC, "testing")
ROOT = os.path.dirname(PKG)
INC = os.path.join(ROOT, "include")
LIB = os.path.join(PKG, "libdistegnn_b200.so")
LIB_TESTING = os.path.join(PKG, "libdistegnn_b200_testing.so")
OBJ = os.path.join(CSRC, "build")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
          "-Xptxas", "-v", "--expt-relaxed-constexpr", "-I", CSRC]
# shared by both libraries (error string, parameter layout, device queries)
COMMON = ["api.cu"]


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def testing_sources():
    return sorted(os.path.join("testing", f) for f in os.listdir(TSRC) if f.endswith(".cu"))


def _deps_mtime():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs += [os.path.join(INC, f) for f in os.listdir(INC) if f.endswith(".h")]
    return max(os.path.getmtime(h) for h in hdrs)


def _compile(src, verbose, objdir=None, defs=None):
    obj = os.path.join(objdir or OBJ, src.replace(os.sep, "_")[:-3] + ".o")
    extra = os.environ.get("DISTEGNN_NVCC_DEFS", "").split() + list(defs or [])   # e.g. "-DT16_CHUNK_UNROLL=4"
    if src.startswith("testing" + os.sep):        # the twins' declarations (default visibility) come from the testing header
        extra += ["-include", os.path.join(INC, "distegnn_b200_testing.h")]
    cmd = [NVCC, *ARCH, *CFLAGS, *extra, "-c", os.path.join(CSRC, src), "-o", obj]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{p.stdout}\n{p.stderr}")
    with open(obj + ".ptxas.log", "w") as f:
        f.write(p.stderr)
    if verbose:
        print(p.stderr)
    return obj


def _link(lib, objs):
    tmp = lib + ".tmp"                      # link aside, then rename: a snapshot of the tree never sees a half-written .so
    p = subprocess.run([NVCC, *ARCH, "-shared", "-o", tmp, *objs, "-cudart", "shared",
                        "-Xlinker", "-rpath,/usr/local/cuda/lib64"], capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError(f"link failed:\n{p.stdout}\n{p.stderr}")
    os.replace(tmp, lib)
    return lib


def _build_into(objdir, lib, lib_testing, verbose, defs=None):
    os.makedirs(objdir, exist_ok=True)
    srcs, tsrcs = sources(), testing_sources()
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs) + len(tsrcs))) as ex:
        objs = dict(zip(srcs + tsrcs, ex.map(lambda s: _compile(s, verbose, objdir, defs), srcs + tsrcs)))
    _link(lib, [objs[s] for s in srcs])
    _link(lib_testing, [objs[s] for s in COMMON] + [objs[s] for s in tsrcs])
    return lib


def build(force: bool = False, verbose: bool = False) -> str:
    srcs = [os.path.join(CSRC, s) for s in sources() + testing_sources()]
    newest = max([os.path.getmtime(s) for s in srcs] + [_deps_mtime(), os.path.getmtime(__file__)])
    if not force and all(os.path.exists(l) and os.path.getmtime(l) >= newest for l in (LIB, LIB_TESTING)):
        return LIB
    return _build_into(OBJ, LIB, LIB_TESTING, verbose)


def build_variant(tag: str, defs, verbose: bool = False) -> str:
    """A/B build: the same sources with extra -D flags -> distegnn_b200/variants/libdistegnn_b200.<tag>.so
    (select it at run time with DISTEGNN_B200_LIB=<path>; the variants travel to the GPU box like the main .so)."""
    vdir = os.path.join(PKG, "variants")
    return _build_into(os.path.join(vdir, "build_" + tag), os.path.join(vdir, f"libdistegnn_b200.{tag}.so"),
                       os.path.join(vdir, f"libdistegnn_b200_testing.{tag}.so"), verbose, defs)


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
