"""Build the CUDA engine in-tree: bnmtf_ard_b200/libbnmtf_b200.so (sm_100a only)."""
import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
SRC = os.path.join(PKG, "csrc", "engine.cu")
DEPS = [SRC] + [os.path.join(PKG, "csrc", f) for f in ("kernels.cuh", "kernels_v2.cuh", "kernels_v3.cuh", "kernels_nmtf.cuh", "kernels_kmeans.cuh", "kernels_summary.cuh", "kernels_batch.cuh", "kernels_long.cuh", "rng.cuh",
                                                        "engine_nmtf.inc", "engine_summary.inc", "engine_batch.inc")] + [os.path.join(ROOT, "include", "bnmtf_b200.h")]
LIB = os.path.join(PKG, "libbnmtf_b200.so")

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared"]


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build %s" % LIB)


def up_to_date():
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(d) <= t for d in DEPS if os.path.exists(d))


def build(force=False, verbose=False):
    """Compile csrc/engine.cu for sm_100a.  Returns the path of the shared library."""
    if not force and up_to_date():
        return LIB
    cmd = [find_nvcc()] + NVCC_FLAGS + ["-o", LIB, SRC, "-ldl"]
    if verbose:
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n%s\n%s" % (res.stdout, res.stderr))
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
