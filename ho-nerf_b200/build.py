"""Build libhonerf_b200.so in-tree with nvcc for sm_100a (no torch headers involved: the library is
a plain C-ABI shared object, see include/honerf_b200.h).

    python ho-nerf_b200/build.py [--force]
"""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libhonerf_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_mtime():
    inc = os.path.join(os.path.dirname(HERE), "include", "honerf_b200.h")
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [inc]
    return max(os.path.getmtime(f) for f in files)


def up_to_date():
    return os.path.isfile(LIB) and os.path.getmtime(LIB) >= _deps_mtime()


def _compile(src):
    out = os.path.join(OBJ, src[:-3] + ".o")
    cmd = [NVCC] + FLAGS + ["-c", os.path.join(CSRC, src), "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    return out


def build(force=False, verbose=True):
    if not force and up_to_date():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    srcs = _sources()
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(_compile, srcs))
    cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    if verbose:
        print("built", LIB, "from", ", ".join(srcs))
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
