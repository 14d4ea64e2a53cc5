"""In-tree nvcc build of the sm_100a kernels + C ABI -> viddet_b200/libviddet_b200.so.

nvcc cross-compiles without a GPU; the built .so is git-ignored but travels to the GPU box with
the gpurun snapshot.  `python -m viddet_b200.build [--force] [--verbose]`.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_build")
SO = os.path.join(HERE, "libviddet_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest():
    h = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)) + ["../../include/viddet_b200.h"]:
        p = os.path.join(CSRC, f)
        if os.path.isfile(p):
            h.update(f.encode())
            h.update(open(p, "rb").read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build_variant(tag: str, extra_flags) -> str:
    """Tuning aid: the same sources with extra -D flags -> viddet_b200/variants/libviddet_b200_<tag>.so (git-ignored); select it at
    run time with VD_LIB=<path>.  The product library is always the default build()."""
    vdir = os.path.join(HERE, "variants")
    odir = os.path.join(OBJ, "variant_" + tag)
    os.makedirs(vdir, exist_ok=True); os.makedirs(odir, exist_ok=True)
    out = os.path.join(vdir, "libviddet_b200_%s.so" % tag)

    def one(src):
        obj = os.path.join(odir, src[:-3] + ".o")
        r = subprocess.run([NVCC] + FLAGS + list(extra_flags) + ["-c", os.path.join(CSRC, src), "-o", obj], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, r.stdout + r.stderr))
        open(os.path.join(odir, src[:-3] + ".ptxas.log"), "w").write(r.stdout + r.stderr)
        return obj
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(one, sources()))
    r = subprocess.run([NVCC, "-shared", "-o", out] + objs + ["-lcudart"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return out


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    stamp = os.path.join(OBJ, "digest.txt")
    dig = _digest()
    if not force and os.path.exists(SO) and os.path.exists(stamp) and open(stamp).read() == dig:
        return SO
    if not os.path.exists(NVCC):
        if os.path.exists(SO):
            return SO        # GPU box without a toolchain change: use the snapshot's prebuilt library
        raise RuntimeError("nvcc not found at %s and no prebuilt %s" % (NVCC, SO))

    def compile_one(src):
        obj = os.path.join(OBJ, src[:-3] + ".o")
        cmd = [NVCC] + FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = r.stdout + r.stderr
        open(os.path.join(OBJ, src[:-3] + ".ptxas.log"), "w").write(log)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, log))
        if verbose:
            print(log)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, sources()))
    cmd = [NVCC, "-shared", "-o", SO] + objs + ["-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    open(stamp, "w").write(dig)
    return SO


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--variant":
        print(build_variant(sys.argv[2], sys.argv[3:]))
    else:
        print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
