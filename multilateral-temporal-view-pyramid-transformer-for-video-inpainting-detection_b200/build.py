"""nvcc recipe for libmumpy_b200.so (sm_100a only, -lineinfo so ncu source pages map to the .cu files).

Every csrc/*.cu is compiled to its own object (in parallel, only when it or a header changed) and the objects are linked
into one shared library; there is no relocatable device code, each translation unit is self-contained.
`python build.py [--force] [-v]`; -v adds -Xptxas=-v and prints the per-kernel register / spill report of what was rebuilt.
"""
import concurrent.futures
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
OUT = os.path.join(HERE, "libmumpy_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _headers():
    return glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(HERE, "..", "include", "mumpy_b200.h")]


def _obj(src):
    return os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def needs_build():
    return _stale(OUT, sources() + _headers())


def _compile(src, verbose):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + ARCH + ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-c", "-o", _obj(src)] \
        + os.environ.get("MUMPY_NVCC_FLAGS", "").split() + [src]      # e.g. -DWTC_TIMING
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (os.path.basename(src), r.stdout, r.stderr))
    return src, r.stderr


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    os.makedirs(OBJ, exist_ok=True)
    hdrs = _headers()
    todo = [s for s in sources() if force or _stale(_obj(s), [s] + hdrs)]
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        for src, log in pool.map(lambda s: _compile(s, verbose), todo):
            if verbose:
                print("==== %s\n%s" % (os.path.basename(src), log))
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    subprocess.run([nvcc] + ARCH + ["-shared", "-Xcompiler", "-fPIC", "-o", OUT] + [_obj(s) for s in sources()], check=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
