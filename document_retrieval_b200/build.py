"""Builds csrc/*.cu into document_retrieval_b200/libbr_b200.so for sm_100a (in-tree; the .so is
git-ignored but travels to the GPU box).  nvcc cross-compiles without a GPU."""
from __future__ import annotations

import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, os.environ.get("BR_SO_NAME", "libbr_b200.so"))   # BR_SO_NAME + BR_NVCC_EXTRA: build-variant experiments
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"] + os.environ.get("BR_NVCC_EXTRA", "").split()


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def stale() -> bool:
    if not os.path.isfile(SO):
        return True
    t = os.path.getmtime(SO)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(HERE, "..", "include", "br_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return SO
    objs = []
    bdir = "build" if SO.endswith("libbr_b200.so") else "build_" + os.path.basename(SO)[:-3]
    os.makedirs(os.path.join(HERE, bdir), exist_ok=True)
    procs = []
    for src in sources():
        obj = os.path.join(HERE, bdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = [NVCC, *FLAGS, "-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, p in procs:
        out, _ = p.communicate()
        log.append(f"== {os.path.basename(src)}\n{out}")
        if p.returncode != 0:
            sys.stderr.write("\n".join(log))
            raise RuntimeError(f"nvcc failed on {src}")
    cmd = [NVCC, "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a", "-o", SO, *objs]
    subprocess.check_call(cmd)
    with open(os.path.join(HERE, bdir, "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
