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
SELFTEST_LIB = os.path.join(HERE, "libhonerf_b200_selftest.so")   # csrc/selftest/*.cu + runtime.cu: tests only
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _selftest_sources():
    d = os.path.join(CSRC, "selftest")
    return sorted(os.path.join("selftest", f) for f in os.listdir(d) if f.endswith(".cu"))


def _deps_mtime():
    inc = os.path.join(os.path.dirname(HERE), "include")
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if os.path.isfile(os.path.join(CSRC, f))]
    files += [os.path.join(CSRC, f) for f in _selftest_sources()] + [os.path.join(inc, f) for f in os.listdir(inc)]
    return max(os.path.getmtime(f) for f in files)


def up_to_date():
    m = _deps_mtime()
    return all(os.path.isfile(l) and os.path.getmtime(l) >= m for l in (LIB, SELFTEST_LIB))


def _compile(src):
    out = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
    cmd = [NVCC] + FLAGS + ["-c", os.path.join(CSRC, src), "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    return out


def build(force=False, verbose=True):
    if not force and up_to_date():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    srcs, tsrcs = _sources(), _selftest_sources()
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(srcs) + len(tsrcs))) as ex:
        objs = list(ex.map(_compile, srcs + tsrcs))
    prod, test = objs[:len(srcs)], objs[len(srcs):]
    runtime = [o for o in prod if os.path.basename(o) == "runtime.o"]
    for lib, parts in ((LIB, prod), (SELFTEST_LIB, test + runtime)):
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + parts
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    if verbose:
        print("built", LIB, "from", ", ".join(srcs), "and", SELFTEST_LIB, "from", ", ".join(tsrcs))
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
