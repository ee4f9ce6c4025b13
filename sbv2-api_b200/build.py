"""Builds libsbv2_b200.so in-tree with nvcc for sm_100a (no other architecture, no fallback)."""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")
OUT_DIR = os.path.join(HERE, "sbv2_b200")
OBJ_DIR = os.path.join(HERE, "build")
LIB = os.path.join(OUT_DIR, "libsbv2_b200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function",
         "--expt-relaxed-constexpr", "-I", os.path.join(HERE, "..", "include"), "-I", CSRC, "-I", HOST]


def sources():
    out = []
    for d in (CSRC, HOST):
        if not os.path.isdir(d):
            continue
        for f in sorted(os.listdir(d)):
            if f.endswith((".cu", ".cpp")):
                out.append(os.path.join(d, f))
    return out


def _digest(path: str) -> str:
    h = hashlib.sha1()
    for d in (CSRC, HOST, os.path.join(HERE, "..", "include")):
        if not os.path.isdir(d):
            continue
        for f in sorted(os.listdir(d)):
            if f.endswith((".h", ".cuh", ".hpp")):
                h.update(open(os.path.join(d, f), "rb").read())
    h.update(open(path, "rb").read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(verbose: bool = False, force: bool = False) -> str:
    os.makedirs(OBJ_DIR, exist_ok=True)
    objs, jobs = [], []
    for src in sources():
        obj = os.path.join(OBJ_DIR, os.path.basename(src) + ".o")
        stamp = obj + ".sha1"
        dig = _digest(src)
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
            continue
        extra = ["-x", "cu"] if src.endswith(".cpp") else []
        jobs.append((src, obj, stamp, dig, [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + extra + ["-c", src, "-o", obj]))

    def run(job):
        src, obj, stamp, dig, cmd = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        open(stamp, "w").write(dig)
        return src

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        list(ex.map(run, jobs))
    if jobs or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-lcudart", "-ldl", "-Xlinker", "-rpath,/usr/local/cuda/lib64"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
