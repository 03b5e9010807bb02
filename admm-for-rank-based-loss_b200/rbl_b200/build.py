"""Build librbl_b200.so (sm_100a only) in-tree with nvcc.  No torch / pybind dependency: the library
exports a plain C ABI (include/rbl_b200.h) and is loaded with ctypes."""
import os
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "librbl_b200.so")
SOURCES = ["api.cu", "pass_kernels.cu", "vec_kernels.cu", "sort_kernels.cu", "pav_kernels.cu", "batch_kernels.cu",
           "gram_kernels.cu", "metrics_kernels.cu", "ingest_kernels.cu"]
HEADERS = ["common.cuh", "prox_core.h", "pav_core.h", "lbfgs_core.h", os.path.join("..", "..", "include", "rbl_b200.h")]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--shared",
              "-Xcompiler", "-fPIC"]


def have_nvcc():
    import shutil

    return shutil.which(os.environ.get("NVCC", "nvcc")) is not None


def stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    if not force and not stale():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + SOURCES
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        sys.stderr.write(res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
