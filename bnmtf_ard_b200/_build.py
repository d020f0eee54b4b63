"""Build the CUDA engine in-tree: bnmtf_ard_b200/libbnmtf_b200.so (sm_100a only)."""
import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
SRC = os.path.join(PKG, "csrc", "engine.cu")


def deps():
    """Every header / include file of the engine, globbed so that a new one cannot be forgotten."""
    import glob
    csrc = os.path.join(PKG, "csrc")
    return sorted(glob.glob(os.path.join(csrc, "*.cu")) + glob.glob(os.path.join(csrc, "*.cuh")) + glob.glob(os.path.join(csrc, "*.inc"))
                  + glob.glob(os.path.join(ROOT, "include", "*.h")))


LIB = os.path.join(PKG, "libbnmtf_b200.so")

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared", "-split-compile", "0"]


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build %s" % LIB)


def up_to_date():
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(d) <= t for d in deps())


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
