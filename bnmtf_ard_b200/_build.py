"""Build the CUDA engine in-tree: bnmtf_ard_b200/libbnmtf_b200.so (sm_100a only)."""
import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)


def deps():
    """Every header / include file of the engine, globbed so that a new one cannot be forgotten."""
    import glob
    csrc = os.path.join(PKG, "csrc")
    return sorted(glob.glob(os.path.join(csrc, "*.cu")) + glob.glob(os.path.join(csrc, "*.cuh")) + glob.glob(os.path.join(csrc, "*.inc"))
                  + glob.glob(os.path.join(ROOT, "include", "*.h")))


LIB = os.path.join(PKG, "libbnmtf_b200.so")

# -split-compile speeds the build up but costs the cluster sweep 18 % (measured on the B200: 10.09 ms against 8.49 ms
# per sweep for the same source), so each translation unit is compiled whole; the units build in parallel instead.
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC"]
UNITS = ["engine.cu", "sweep4.cu", "sweep5.cu"]


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


def build(force=False, verbose=False, defines=(), lib=None):
    """Compile csrc/*.cu for sm_100a (one nvcc process per translation unit) and link the shared library.
    Returns its path."""
    lib = lib or LIB
    if not force and lib == LIB and up_to_date():
        return lib
    from concurrent.futures import ThreadPoolExecutor
    nvcc = find_nvcc()
    objdir = os.path.join(ROOT, "build", "obj" + ("_" + "_".join(defines) if defines else ""))
    os.makedirs(objdir, exist_ok=True)
    dflags = ["-D" + d for d in defines]

    def compile_unit(unit):
        obj = os.path.join(objdir, unit[:-3] + ".o")
        cmd = [nvcc] + NVCC_FLAGS + dflags + ["-c", "-o", obj, os.path.join(PKG, "csrc", unit)]
        if verbose:
            print(" ".join(cmd))
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s\n%s" % (unit, res.stdout, res.stderr))
        return obj

    with ThreadPoolExecutor(len(UNITS)) as ex:
        objs = list(ex.map(compile_unit, UNITS))
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs + ["-ldl"]
    if verbose:
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (res.stdout, res.stderr))
    return lib


if __name__ == "__main__":
    print(build(force=True, verbose=True))
