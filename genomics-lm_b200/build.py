"""Build libcgpt_b200.so (sm_100a only) in-tree with nvcc.  Usage: python genomics-lm_b200/build.py [--force]"""
import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "codonlm_b200")
LIB = os.path.join(OUT_DIR, "libcgpt_b200.so")
SOURCES = ["core.cu", "elementwise.cu", "lmhead.cu", "gemm_sm100.cu", "attn_sm100.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-diag-suppress", "177"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=True):
    hdrs = [os.path.join(CSRC, "common.cuh"), os.path.join(HERE, "..", "include", "cgpt.h")]
    objs, jobs = [], []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(HERE, "build", s.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [src] + hdrs):
            jobs.append([NVCC] + FLAGS + ["-c", src, "-o", obj])
    with cf.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        for cmd, res in zip(jobs, ex.map(lambda c: subprocess.run(c, capture_output=True, text=True), jobs)):
            if verbose:
                print("[build]", os.path.basename(cmd[-3]), "rc", res.returncode)
            if res.returncode != 0:
                sys.stderr.write(res.stdout + res.stderr)
                raise RuntimeError("nvcc failed: " + " ".join(cmd))
    if force or jobs or _stale(LIB, objs):
        cmd = [NVCC, "-shared", "-o", LIB] + objs
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("link failed")
        if verbose:
            print("[build] linked", LIB)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
