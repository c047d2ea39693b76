"""Builds libmadb.so (CUDA kernels + C ABI) in-tree for sm_100a with nvcc.

    python mfem-ad_b200/build.py [--force]

Each csrc/*.cu / *.cpp is compiled to build/*.o in parallel, then linked into
mfem-ad_b200/libmadb.so (static cudart; the .so travels to the GPU box).
"""
import concurrent.futures as cf
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libmadb.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-std=c++17", "-O3", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
         "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall",
         "-diag-suppress", "177,550,128"] + os.environ.get("MADB_CFLAGS", "").split()


def _deps_newer(obj):
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    hdrs = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.hpp")) + \
        glob.glob(os.path.join(HERE, "..", "include", "*.h"))
    return any(os.path.getmtime(h) > t for h in hdrs)


def _compile(src, force, verbose):
    obj = os.path.join(OBJ, os.path.basename(src) + ".o")
    if not force and not _deps_newer(obj) and os.path.getmtime(obj) > os.path.getmtime(src):
        return obj, ""
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-x", "cu", "-c", src, "-o", obj]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError("nvcc failed on %s:\n%s\n%s" % (src, p.stdout, p.stderr))
    return obj, p.stderr


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cpp")))
    objs, logs = [], []
    with cf.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        for obj, log in ex.map(lambda s: _compile(s, force, verbose), srcs):
            objs.append(obj)
            logs.append(log)
    if force or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        subprocess.check_call([NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"])
    if verbose:
        with open(os.path.join(OBJ, "ptxas.log"), "w") as f:
            f.write("\n".join(logs))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
