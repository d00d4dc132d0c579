"""Build libgripnet_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch headers).

    python gripnet_b200/build.py [--force] [--verbose]

(run it by path, or through ``__graft_entry__.build()``: importing the package itself
needs the library to exist already.)  Every ``csrc/*.cu`` is compiled to an object in
parallel (only the stale ones), then linked into one shared library.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libgripnet_b200.so")
HEADER = os.path.join(os.path.dirname(HERE), "include", "gripnet_b200.h")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default",
]


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))] + [HEADER]


def _newer(path, deps):
    if not os.path.isfile(path):
        return True
    t = os.path.getmtime(path)
    return any(os.path.getmtime(d) > t for d in deps)


def is_stale():
    return _newer(LIB, [os.path.join(CSRC, s) for s in sources()] + _headers())


def build(force=False, verbose=False):
    if not force and not is_stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ, exist_ok=True)
    hdrs = _headers()
    flags = NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else [])
    jobs = []
    for s in sources():
        src, obj = os.path.join(CSRC, s), os.path.join(OBJ, s[:-3] + ".o")
        if force or verbose or _newer(obj, [src] + hdrs):
            jobs.append((s, [nvcc] + flags + ["-c", src, "-o", obj]))

    def run(job):
        name, cmd = job
        return name, subprocess.run(cmd, capture_output=True, text=True)

    with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as ex:
        results = list(ex.map(run, jobs))
    failed = False
    for name, res in results:
        if res.returncode != 0:
            sys.stderr.write(f"--- {name}\n{res.stdout}{res.stderr}")
            failed = True
        elif verbose:
            print(f"--- {name}\n{res.stdout}{res.stderr}")
    if failed:
        raise RuntimeError("nvcc failed building libgripnet_b200.so")
    objs = [os.path.join(OBJ, s[:-3] + ".o") for s in sources()]
    res = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a"] + objs + ["-o", LIB],
                         capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("link of libgripnet_b200.so failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
