"""Builds libdogs_b200.so (plain nvcc, sm_100a, no torch headers) in-tree.

Run as `python -m dogs_b200.build`; `__graft_entry__.build()` calls build().  The object files
and the shared library are git-ignored but travel to the GPU box with gpurun.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libdogs_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _newer(target, deps):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def build(verbose=False, force=False):
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "dogs_b200.h"))
    jobs = []
    for src in _sources():
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src[:-3] + ".o")
        if force or not _newer(o, [s] + headers):
            jobs.append((s, o))

    def cc(job):
        s, o = job
        r = subprocess.run([NVCC] + FLAGS + ["-c", s, "-o", o], capture_output=True, text=True)
        return s, r

    failed = False
    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for s, r in ex.map(cc, jobs):
            if verbose or r.returncode != 0:
                sys.stderr.write(f"== {os.path.basename(s)}\n{r.stderr}\n")
            with open(os.path.join(OBJ, os.path.basename(s) + ".ptxas.log"), "w") as f:
                f.write(r.stderr)
            failed |= r.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed building libdogs_b200.so")
    objs = [os.path.join(OBJ, src[:-3] + ".o") for src in _sources()]
    if force or jobs or not _newer(LIB, objs):
        r = subprocess.run([NVCC, "-shared", "-o", LIB] + objs + ["-lcudart"],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
