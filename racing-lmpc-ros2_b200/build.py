"""In-tree build of the CUDA C-ABI library (csrc/liblmpc_b200.so) for sm_100a.

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box with the
repo snapshot, so the GPU tests and bench load exactly what was built here.
"""
import os
import shutil
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
OUT = os.path.join(CSRC, "liblmpc_b200.so")
SOURCES = ["lmpc_capi.cu", "lmpc_qp_tu0.cu", "lmpc_qp_tu1.cu", "lmpc_qp_tu2.cu", "lmpc_qp_tu3.cu", "lmpc_qp_tu4.cu"]
# every header of csrc/ is a dependency of every translation unit (globbed: a hand-kept list went stale once)
DEPS = SOURCES + sorted(f for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))) + [os.path.join("..", "..", "include", "lmpc_b200.h")]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def nvcc_path():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def build(force=False, verbose=False):
    """One nvcc process per translation unit (the QP kernel instantiations are spread over four), run in
    parallel, then one link step.  Objects live in build_dbg/ at the repo root (git-ignored, not shipped to the GPU box)."""
    deps = [os.path.join(CSRC, d) for d in DEPS if os.path.exists(os.path.join(CSRC, d))]
    if not force and os.path.exists(OUT):
        t = os.path.getmtime(OUT)
        if all(os.path.getmtime(d) <= t for d in deps):
            return OUT
    nvcc = nvcc_path()
    objdir = os.path.join(_HERE, "..", "build_dbg")   # git- and gpurun-ignored
    os.makedirs(objdir, exist_ok=True)
    hdr_t = max(os.path.getmtime(d) for d in deps if not d.endswith(".cu"))
    procs, log = [], ""
    for s in SOURCES:
        src, obj = os.path.join(CSRC, s), os.path.join(objdir, s[:-3] + ".o")
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(hdr_t, os.path.getmtime(src)):
            continue
        extra = os.environ.get("LMPC_NVCC_FLAGS", "").split()   # experiments only, e.g. -DLMPC_K1_MINBLOCKS=3
        cmd = [nvcc, "-O3", "-std=c++17"] + ARCH + ["-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v"] + extra + ["-c", "-o", obj, src]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for cmd, p in procs:
        out, _ = p.communicate()
        log += " ".join(cmd) + "\n" + out
        failed |= p.returncode != 0
    if not failed:
        cmd = [nvcc, "-shared"] + ARCH + ["-o", OUT] + [os.path.join(objdir, s[:-3] + ".o") for s in SOURCES]
        res = subprocess.run(cmd, capture_output=True, text=True)
        log += " ".join(cmd) + "\n" + res.stdout + res.stderr
        failed = res.returncode != 0
    with open(os.path.join(CSRC, "build.log"), "w") as f:
        f.write(log)
    if failed:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed")
    if verbose:
        print(log)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
