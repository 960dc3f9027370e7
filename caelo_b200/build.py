"""Build libcaelo_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with
the repo snapshot to the GPU box).  ``python -m caelo_b200.build`` or ``build()``."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libcaelo_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]
# pose.cu carries float64 contract K1: no FMA contraction (see oracle/caelo_oracle.c)
SOURCES = {
    "api.cu": [],
    "respond_select.cu": [],
    "patches.cu": [],
    "scan.cu": [],
    "extend.cu": [],
    "encoder.cu": [],
    "match.cu": [],
    "pose.cu": ["-fmad=false"],
    "icp.cu": ["-fmad=false"],
    "umma_debug.cu": [],
}


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isfile(c) or c == "nvcc"):
            return c
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.isfile(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [
        os.path.join(os.path.dirname(HERE), "include", "caelo.h"), os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, extra_flags=(), out_path: str = OUT, objdir_name: str = "build") -> str:
    """``extra_flags`` / ``out_path`` / ``objdir_name``: an experimental variant next to the product library (A/B timing on
    the GPU box: CAELO_SO_PATH selects the library _lib.load() opens)."""
    if not force and not extra_flags and not needs_build():
        return OUT
    nvcc = _nvcc()
    objdir = os.path.join(HERE, objdir_name)
    os.makedirs(objdir, exist_ok=True)
    procs = []
    objs = []
    for src, extra in SOURCES.items():
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [nvcc, *ARCH, *COMMON, *extra, *extra_flags, "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        log.append("== %s ==\n%s" % (src, out))
        failed |= p.returncode != 0
    with open(os.path.join(objdir, "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if failed or verbose:
        sys.stderr.write("\n".join(log))
    if failed:
        raise RuntimeError("nvcc failed (see above)")
    subprocess.check_call([nvcc, *ARCH, "-shared", "-o", out_path, *objs, "-lcudart"])
    return out_path


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
