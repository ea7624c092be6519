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
SOURCES = ["lmpc_capi.cu"]
DEPS = ["lmpc_capi.cu", "lmpc_kernels.cuh", "lmpc_qp_core.cuh", "lmpc_ss_core.cuh", "lmpc_model.cuh",
        "lmpc_warp.cuh", "lmpc_host_params.h", os.path.join("..", "..", "include", "lmpc_b200.h")]


def nvcc_path():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def build(force=False, verbose=False):
    if not force and os.path.exists(OUT):
        t = os.path.getmtime(OUT)
        if all(os.path.getmtime(os.path.join(CSRC, d)) <= t for d in DEPS):
            return OUT
    cmd = [nvcc_path(), "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
           "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v", "-o", OUT] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(CSRC, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if res.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed")
    if verbose:
        print(log)
    return OUT


if __name__ == "__main__":
    build(force=True, verbose=True)
